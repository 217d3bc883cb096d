// Fused renderer: one persistent kernel per batch; the OSGDecoder runs on the 5th-gen tensor cores.
//
// A CTA owns a *ray group* (G = 192/S rays: 2 rays at S=Sf=96, 4 at S=Sf=48) and walks it through
//   coarse pass (2 MMA tiles of 128 rows: 128 + 64 valid) -> per-ray importance sampling ->
//   fine pass (2 tiles) -> per-ray merge + alpha/transmittance scan -> weighted colour reduction.
// Nothing but the final (rgb, depth, wsum, xyz) leaves the SM:
//   * tri-plane taps: 8 lanes x 16 B = one 128 B texel per tap (same cooperative gather as v1);
//   * features -> bf16 hi/lo split -> canonical K-major (no-swizzle) UMMA tiles in shared memory;
//   * layer 1: D1[128x64] (TMEM) = A1[128x32] * W1^T, three tcgen05.mma passes (hi*hi + hi*lo + lo*hi)
//     per K=16 step = fp32-class accuracy out of bf16 tensor cores (single pass in fast mode);
//   * epilogue 1: tcgen05.ld D1 -> +b1 -> softplus -> hi/lo split -> A2 tile in shared memory;
//   * layer 2: D2[t][128x48] (TMEM) = A2[128x64] * W2^T (33 outputs padded to 48);
//     the logits of all four tiles of the group STAY in TMEM (64 + 4*48 = 256 columns) until the
//     compositing weights are known, so colours are never stored anywhere;
//   * sigma (column 0) is read back right away for the importance pass / weights;
//   * final: tcgen05.ld logits -> sigmoid -> * omega[row] -> warp transpose-reduce -> 32 floats/ray.
// 256 TMEM columns and ~80 KB of shared memory per CTA => 2 CTAs per SM, which is what overlaps
// one CTA's gather (L1/L2 bound) with the other's epilogues (issue bound).
#include "render_device.cuh"

namespace p3d {

namespace {

using namespace dev;

constexpr int kThreads = 256;            // 8 warps: two per TMEM lane quarter
constexpr int kWarps = kThreads / 32;
constexpr int kRows = 192;               // rows (samples) per pass per group
constexpr int kN2 = 48;                  // layer-2 N (33 padded to a multiple of 16)
constexpr int kTmemCols = 256;           // D1: [0,64)   D2[t]: [64+48t, 64+48t+48), t = 0..3

// canonical K-major / no-swizzle operand tiles: 8x8 bf16 core matrices of 128 contiguous bytes,
// 8-row groups 128 B apart (SBO), K-cores LBO apart.
constexpr int kSBO = 128;
constexpr int kLBO_A = 128 * 16;         // 128-row A tiles: 16 row-groups per K-core -> 2048 B
constexpr int kLBO_W1 = 128 * 8;         // W1: 64 rows
constexpr int kLBO_W2 = 128 * 6;         // W2: 48 rows
constexpr int kA1Bytes = 128 * 32 * 2;   // 8 KB
constexpr int kA2Bytes = 128 * 64 * 2;   // 16 KB
constexpr int kW1Bytes = 64 * 32 * 2;    // 4 KB
constexpr int kW2Bytes = 48 * 64 * 2;    // 6 KB

struct __align__(1024) FusedSmem {
    unsigned char a1_hi[kA1Bytes], a1_lo[kA1Bytes];
    unsigned char a2_hi[kA2Bytes], a2_lo[kA2Bytes];      // also reused as per-ray scratch between passes
    unsigned char w1_hi[kW1Bytes], w1_lo[kW1Bytes];
    unsigned char w2_hi[kW2Bytes], w2_lo[kW2Bytes];
    float b1[kHidden], b2[kN2];
    float t_c[kRows], sg_c[kRows], t_f[kRows], sg_f[kRows];   // per-row depth / density of both passes
    float om_c[kRows], om_f[kRows];                           // per-row composite weight omega
    float xz[kRows * 2];                                      // raw x,z of the current pass (crop mask)
    float acc[4][kRgb];                                       // per-ray colour accumulators
    float ray_back[4];
    unsigned long long mbar;
    unsigned int tmem_base;
    unsigned int pad;
};

struct FusedArgs {
    Geom g;
    const void* planes;
    const float *w1, *b1, *w2, *b2;
    const float *ro, *rd, *u_c, *u_f;
    const float *ray_t0, *ray_t1;
    unsigned int* bounds;
    float *out_rgb, *out_depth, *out_wsum, *out_xyz;
    long long R;              // total rays
    int n_groups, G;          // ray groups, rays per group
    int single_pass;          // 1: fast mode (bf16 hi*hi only)
    int sort_pow2;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) { asm volatile("trap;"); }     // ~2 s: never hang the GPU
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned int* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M=128, K=16
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                                   // descriptor version (sm_100)
    return d;                                          // layout_type = 0 (no swizzle), base_offset = 0
}
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
    // c_format F32 (1<<4) | a_format BF16 (1<<7) | b_format BF16 (1<<10) | K-major A,B | N>>3 @17 | M>>4 @24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r) : "r"(taddr) : "memory");
    return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------ math
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// x = hi + lo (+ 2^-17 x): two bf16 that together carry 16 significant bits
__device__ __forceinline__ void split_bf16(float x, unsigned short& hi, unsigned short& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}
// byte offset of element (row, k) inside a canonical K-major tile whose K-cores are `lbo` bytes apart
__device__ __forceinline__ int tile_off(int row, int k, int lbo) { return (row >> 3) * kSBO + (k >> 3) * lbo + (row & 7) * 16 + (k & 7) * 2; }

// ------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kThreads, 2) k_render_fused(const FusedArgs a) {
    extern __shared__ unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Geom& g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = g.S, Sf = g.Sf, G = a.G;

    // ---------------- one-time setup: TMEM, barrier, decoder weights as UMMA B tiles
    if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
    if (tid == 32) { mbar_init(&sm.mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < kHidden * kC; i += kThreads) {            // W1 (64,32): row n = i/32, k = i%32
        const int n = i / kC, k = i % kC;
        unsigned short hi, lo;
        split_bf16(__fmul_rn(a.w1[i], g.w1_gain), hi, lo);
        const int off = tile_off(n, k, kLBO_W1);
        *reinterpret_cast<unsigned short*>(sm.w1_hi + off) = hi;
        *reinterpret_cast<unsigned short*>(sm.w1_lo + off) = lo;
    }
    for (int i = tid; i < kN2 * kHidden; i += kThreads) {           // W2 (33,64) zero-padded to 48 rows
        const int n = i / kHidden, k = i % kHidden;
        unsigned short hi = 0, lo = 0;
        if (n < kOut) split_bf16(__fmul_rn(a.w2[n * kHidden + k], g.w2_gain), hi, lo);
        const int off = tile_off(n, k, kLBO_W2);
        *reinterpret_cast<unsigned short*>(sm.w2_hi + off) = hi;
        *reinterpret_cast<unsigned short*>(sm.w2_lo + off) = lo;
    }
    if (tid < kHidden) sm.b1[tid] = __fmul_rn(a.b1[tid], g.b1_gain);
    if (tid < kN2) sm.b2[tid] = tid < kOut ? __fmul_rn(a.b2[tid], g.b2_gain) : 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
    uint32_t phase = 0;

    const uint32_t idesc1 = umma_idesc(128, kHidden), idesc2 = umma_idesc(128, kN2);
    const uint32_t a1h = smem_u32(sm.a1_hi), a1l = smem_u32(sm.a1_lo), a2h = smem_u32(sm.a2_hi), a2l = smem_u32(sm.a2_lo);
    const uint32_t w1h = smem_u32(sm.w1_hi), w1l = smem_u32(sm.w1_lo), w2h = smem_u32(sm.w2_hi), w2l = smem_u32(sm.w2_lo);

    // per-ray scratch (importance / merge) lives in the A2 region, which is dead between passes
    float* scratch = reinterpret_cast<float*>(sm.a2_hi);            // 32 KB = 8192 floats

    for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
        const long long ray0 = (long long)grp * G;

        // =================================================================== two passes x two tiles
        for (int pass = 0; pass < 2; ++pass) {
            const int per_ray = pass == 0 ? S : Sf;
            float* t_arr = pass == 0 ? sm.t_c : sm.t_f;
            float* sg_arr = pass == 0 ? sm.sg_c : sm.sg_f;
            for (int tile = 0; tile < 2; ++tile) {
                const int rows_valid = tile == 0 ? 128 : kRows - 128;
                // ------------------------------------------------ gather: 16 rows per warp, 4 per round
                {
                    const int sub = lane >> 3, q = lane & 7;
#pragma unroll 1
                    for (int round = 0; round < 4; ++round) {
                        const int trow = warp * 16 + round * 4 + sub;      // row inside the tile
                        if (trow >= rows_valid) continue;                  // warp-uniform (16-row granularity)
                        const int prow = tile * 128 + trow;                // row inside the pass
                        const int rl = prow / per_ray, s = prow - rl * per_ray;
                        const long long ray = ray0 + rl;
                        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                        float px = 0.f, pz = 0.f, tval = 0.f;
                        if (ray < a.R) {
                            if (pass == 0) {
                                const long long gidx = ray * S + s;
                                const float u = a.u_c ? a.u_c[gidx] : philox_uniform(g.seed, (uint64_t)gidx, 0u);
                                float t0 = 0.f, t1 = 0.f;
                                if (g.ray_mode == P3D_RAYS_AUTOBOX) {
                                    t0 = a.ray_t0[ray]; t1 = a.ray_t1[ray];
                                    if (!(t1 > t0) && a.bounds[4]) { t0 = ordered_to_float(a.bounds[2]); t1 = ordered_to_float(a.bounds[3]); }
                                }
                                tval = coarse_depth(g, s, u, t0, t1);
                            } else {
                                tval = t_arr[prow];
                            }
                            const float* o = a.ro + ray * 3;
                            const float* d = a.rd + ray * 3;
                            px = __fadd_rn(o[0], __fmul_rn(tval, d[0]));
                            const float py = __fadd_rn(o[1], __fmul_rn(tval, d[1]));
                            pz = __fadd_rn(o[2], __fmul_rn(tval, d[2]));
                            const int view = (int)(ray / g.M);
                            f = gather_features<BF16>(a.planes, g, view, px, py, pz, q);
                        }
                        unsigned short h0, h1, h2, h3, l0, l1, l2, l3;
                        split_bf16(f.x, h0, l0); split_bf16(f.y, h1, l1); split_bf16(f.z, h2, l2); split_bf16(f.w, h3, l3);
                        const int off = tile_off(trow, 4 * q, kLBO_A);
                        *reinterpret_cast<uint2*>(sm.a1_hi + off) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
                        *reinterpret_cast<uint2*>(sm.a1_lo + off) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
                        if (q == 0) {
                            if (pass == 0) t_arr[prow] = tval;
                            sm.xz[prow * 2] = px; sm.xz[prow * 2 + 1] = pz;
                        }
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                // ------------------------------------------------ layer 1 on the tensor cores
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kC / 16; ++ks) {
                        const uint32_t ao = ks * 2 * kLBO_A, bo = ks * 2 * kLBO_W1;
                        umma_bf16(tmem, umma_desc(a1h + ao, kLBO_A, kSBO), umma_desc(w1h + bo, kLBO_W1, kSBO), idesc1, ks > 0);
                        if (!a.single_pass) {
                            umma_bf16(tmem, umma_desc(a1h + ao, kLBO_A, kSBO), umma_desc(w1l + bo, kLBO_W1, kSBO), idesc1, 1);
                            umma_bf16(tmem, umma_desc(a1l + ao, kLBO_A, kSBO), umma_desc(w1h + bo, kLBO_W1, kSBO), idesc1, 1);
                        }
                    }
                    umma_commit(&sm.mbar);
                }
                mbar_wait(&sm.mbar, phase); phase ^= 1;
                tc_fence_after();
                // ------------------------------------------------ epilogue 1: softplus -> A2 (hi/lo)
                {
                    const int trow = (warp & 3) * 32 + lane;
                    const int chunk = warp >> 2;                           // columns [32*chunk, 32*chunk+32)
                    if ((warp & 3) * 32 < rows_valid) {                    // warp-uniform
                        float v[32];
                        tmem_ld32(tmem + lane_base + 32 * chunk, v);
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int j = c8 * 8 + 2 * e;
                                unsigned short h0, l0, h1, l1;
                                split_bf16(softplus_fast(v[j] + sm.b1[32 * chunk + j]), h0, l0);
                                split_bf16(softplus_fast(v[j + 1] + sm.b1[32 * chunk + j + 1]), h1, l1);
                                ph[e] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                                pl[e] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                            }
                            const int off = tile_off(trow, 32 * chunk + 8 * c8, kLBO_A);
                            *reinterpret_cast<uint4*>(sm.a2_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(sm.a2_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                // ------------------------------------------------ layer 2 -> D2[2*pass + tile]
                const uint32_t d2 = tmem + 64 + kN2 * (2 * pass + tile);
                if (tid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < kHidden / 16; ++ks) {
                        const uint32_t ao = ks * 2 * kLBO_A, bo = ks * 2 * kLBO_W2;
                        umma_bf16(d2, umma_desc(a2h + ao, kLBO_A, kSBO), umma_desc(w2h + bo, kLBO_W2, kSBO), idesc2, ks > 0);
                        if (!a.single_pass) {
                            umma_bf16(d2, umma_desc(a2h + ao, kLBO_A, kSBO), umma_desc(w2l + bo, kLBO_W2, kSBO), idesc2, 1);
                            umma_bf16(d2, umma_desc(a2l + ao, kLBO_A, kSBO), umma_desc(w2h + bo, kLBO_W2, kSBO), idesc2, 1);
                        }
                    }
                    umma_commit(&sm.mbar);
                }
                mbar_wait(&sm.mbar, phase); phase ^= 1;
                tc_fence_after();
                // ------------------------------------------------ sigma = column 0 (+ masks)
                if (warp < 4 && warp * 32 < rows_valid) {
                    const int prow = tile * 128 + warp * 32 + lane;
                    const float sraw = tmem_ld1(d2 + lane_base) + sm.b2[0];
                    sg_arr[prow] = apply_masks(g, sraw, sm.xz[prow * 2], sm.xz[prow * 2 + 1]);
                }
                tc_fence_before();
                __syncthreads();
            }  // tile

            if (pass == 0) {
                // -------------------------------------------- importance sampling, one warp per ray
                if (Sf > 0) {
                    for (int rl = warp; rl < G; rl += kWarps) {
                        const long long ray = ray0 + rl;
                        float* wv = scratch + rl * (2 * S + a.sort_pow2);
                        float* cdf = wv + S;
                        float* fine = cdf + S;
                        if (ray < a.R) {
                            importance_ray(g, sm.t_c + rl * S, sm.sg_c + rl * S, wv, cdf, fine, a.sort_pow2,
                                           a.u_f ? a.u_f + ray * Sf : nullptr, (unsigned long long)(ray * Sf), lane);
                            for (int f = lane; f < Sf; f += 32) sm.t_f[rl * Sf + f] = fine[f];
                        }
                    }
                }
                __syncthreads();
            }
        }  // pass

        // =================================================================== per-ray weights
        if (tid < 4 * kRgb) sm.acc[tid >> 5][tid & 31] = 0.f;
        for (int rl = warp; rl < G; rl += kWarps) {
            const long long ray = ray0 + rl;
            if (ray >= a.R) continue;
            const int L = S + Sf;
            float* t = scratch + rl * 4 * L;
            float* sg = t + L;
            float* w = sg + L;
            int* src = reinterpret_cast<int*>(w + L);
            float wsum, dnum;
            composite_weights(sm.t_c + rl * S, sm.sg_c + rl * S, sm.t_f + rl * Sf, sm.sg_f + rl * Sf, S, Sf, t, sg, w, src, lane, wsum, dnum);
            for (int j = lane; j < L; j += 32) {
                const int s = src[j];
                if (s < S) sm.om_c[rl * S + s] = w[j]; else sm.om_f[rl * Sf + (s - S)] = w[j];
            }
            const float back = g.white_back ? __fsub_rn(1.f, wsum) : 0.f;
            if (lane < 3) {
                const float v = fmaf(a.ro[ray * 3 + lane], wsum, a.rd[ray * 3 + lane] * dnum);
                a.out_xyz[ray * 3 + lane] = __fsub_rn(__fmul_rn(__fadd_rn(v, back), 2.f), 1.f);
            }
            if (lane == 0) {
                sm.ray_back[rl] = back;
                a.out_depth[ray] = __fdiv_rn(dnum, wsum);
                a.out_wsum[ray] = wsum;
                atomicMin(&a.bounds[0], float_to_ordered(t[0]));
                atomicMax(&a.bounds[1], float_to_ordered(t[L - 1]));
            }
        }
        __syncthreads();
        // =================================================================== colours: sum_j omega_j * rgb_j
        tc_fence_after();
        for (int tt = warp >> 2; tt < 4; tt += 2) {                  // warps 0-3: tiles 0,2   warps 4-7: tiles 1,3
            const int pass = tt >> 1, tile = tt & 1;
            if (pass == 1 && Sf == 0) continue;
            const int rows_valid = tile == 0 ? 128 : kRows - 128;
            if ((warp & 3) * 32 >= rows_valid) continue;
            const int per_ray = pass == 0 ? S : Sf;
            const int prow = tile * 128 + (warp & 3) * 32 + lane;
            const int rl = prow / per_ray;
            const bool live = ray0 + rl < a.R;
            const float om = live ? (pass == 0 ? sm.om_c[prow] : sm.om_f[prow]) : 0.f;
            float v[32];
            tmem_ld32(tmem + 64 + kN2 * tt + 1 + lane_base, v);
#pragma unroll
            for (int c = 0; c < kRgb; ++c) {
                float col = sigmoid_fast(v[c] + sm.b2[1 + c]);
                if (!g.force_sigmoid) col = fmaf(col, 1.002f, -0.001f);
                v[c] = om * col;
            }
            // rows of one warp belong to at most two rays (per_ray >= 32): reduce each separately
            const int rl_lo = __shfl_sync(0xffffffffu, rl, 0), rl_hi = __shfl_sync(0xffffffffu, rl, 31);
            for (int target = rl_lo; target <= rl_hi; ++target) {
                float r[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) r[c] = rl == target ? v[c] : 0.f;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        const bool up = (lane & off) != 0;
                        const float send = up ? r[i] : r[i + off];
                        const float keep = up ? r[i + off] : r[i];
                        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                atomicAdd(&sm.acc[target][lane], r[0]);                 // lane == colour channel
            }
        }
        tc_fence_before();
        __syncthreads();
        if (tid < G * kRgb) {
            const int rl = tid >> 5, c = tid & 31;
            const long long ray = ray0 + rl;
            if (ray < a.R) a.out_rgb[ray * kRgb + c] = __fsub_rn(__fmul_rn(__fadd_rn(sm.acc[rl][c], sm.ray_back[rl]), 2.f), 1.f);
        }
        __syncthreads();
    }  // groups

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace

// kernels of render_v1.cu reused around the fused kernel
int launch_bounds_init(unsigned int* bounds, cudaStream_t stream);
int launch_ray_limits(const float* ro, const float* rd, long long R, float h, float* t0, float* t1, unsigned int* bounds, cudaStream_t stream);
int launch_depth_finalize(float* depth, long long R, const unsigned int* bounds, cudaStream_t stream);

bool fused_supported(const Geom& g) {
    return (g.S == 96 || g.S == 48) && (g.Sf == g.S) && ((long long)g.M % (kRows / g.S) == 0);
}

int render_forward_fused(const Geom& g, const p3d_render_params* p, const void* planes, const float* w1, const float* b1,
                         const float* w2, const float* b2, const float* ro, const float* rd, const float* u_c,
                         const float* u_f, const Workspace& ws, float* out_rgb, float* out_depth, float* out_wsum,
                         float* out_xyz, cudaStream_t stream) {
    if (!fused_supported(g)) {
        set_error("fused tcgen05 renderer supports depth_resolution == depth_resolution_importance in {48, 96} (got %d, %d)", g.S, g.Sf);
        return P3D_EUNSUPPORTED;
    }
    const long long R = (long long)g.N * g.M;
    int rc;
    if ((rc = launch_bounds_init(ws.bounds, stream))) return rc;
    if (g.ray_mode == P3D_RAYS_AUTOBOX)
        if ((rc = launch_ray_limits(ro, rd, R, g.half_box, ws.ray_t0, ws.ray_t1, ws.bounds, stream))) return rc;
    FusedArgs a{};
    a.g = g; a.planes = planes; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.ro = ro; a.rd = rd; a.u_c = u_c; a.u_f = u_f;
    a.ray_t0 = ws.ray_t0; a.ray_t1 = ws.ray_t1; a.bounds = ws.bounds;
    a.out_rgb = out_rgb; a.out_depth = out_depth; a.out_wsum = out_wsum; a.out_xyz = out_xyz;
    a.R = R; a.G = kRows / g.S; a.n_groups = (int)((R + a.G - 1) / a.G);
    a.single_pass = p->mlp_mode == P3D_MLP_TC_BF16;
    int p2 = 1; while (p2 < g.Sf) p2 <<= 1;
    a.sort_pow2 = p2;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        P3D_CUDA_TRY(cudaGetDevice(&dev));
        P3D_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = sizeof(FusedSmem) + 1024;
    auto kern = p->planes_bf16 ? k_render_fused<true> : k_render_fused<false>;
    P3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = a.n_groups < 2 * n_sm ? a.n_groups : 2 * n_sm;
    {
        ProfileScope prof(PROF_FUSED, stream);
        kern<<<grid, kThreads, smem, stream>>>(a);
        P3D_LAUNCH_CHECK();
    }
    return launch_depth_finalize(out_depth, R, ws.bounds, stream);
}

}  // namespace p3d
