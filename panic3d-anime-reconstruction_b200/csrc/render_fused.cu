// Fused renderer: one persistent kernel per batch; the OSGDecoder runs on the 5th-gen tensor cores.
//
// A CTA owns a *ray group* (G = 192/S rays: 2 rays at S=Sf=96, 4 at S=Sf=48) and walks it through
//   coarse pass (2 MMA tiles of 128 rows: 128 + 64 valid) -> per-ray importance sampling ->
//   fine pass (2 tiles) -> per-ray merge + alpha/transmittance scan -> weighted colour reduction.
// Nothing but the final (rgb, depth, wsum, xyz) leaves the SM:
//   * taps: each sample's 12 (texel offset, bilinear weight) records are computed ONCE by a lane pair
//     (32-bit offsets inside the view's tri-plane) and parked in shared memory; the gather proper is
//     8 lanes x 16 B = one 128 B texel per tap, 12 independent 128-bit loads in flight per lane;
//   * features -> bf16 hi/lo split (cvt.rn.bf16x2) -> canonical K-major no-swizzle UMMA tiles in smem;
//   * layer 1: D1[128x64] (TMEM) = A1[128x32] * W1'^T, three tcgen05.mma passes per K=16 step
//     (hi*hi + hi*lo + lo*hi) = fp32-class accuracy out of bf16 tensor cores (one pass in fast mode);
//     W1' = W1*gain*log2(e) so the epilogue's softplus is lg2(1 + ex2(.)) - two MUFU ops;
//   * epilogue 1: tcgen05.ld D1 -> +b1' -> softplus2 -> hi/lo split -> A2 tile in shared memory;
//   * layer 2: D2[t][128x48] (TMEM) = A2[128x64] * W2'^T; row 0 of W2' (sigma) carries ln2, rows 1..32
//     are negated so the colour epilogue is rcp(1 + ex2(.)); the logits of all four tiles of the group
//     STAY in TMEM (64 + 4*48 = 256 columns) until the compositing weights are known;
//   * sigma (column 0) is read back right away for the importance pass / weights;
//   * final: tcgen05.ld logits -> sigmoid -> * omega[row] -> warp transpose-reduce -> 32 floats/ray.
// 256 TMEM columns and ~90 KB of shared memory per CTA => 2 CTAs per SM, which is what overlaps
// one CTA's gather (L1/L2 latency) with the other's epilogues (issue bound).
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
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct __align__(1024) FusedSmem {
    unsigned char a1_hi[kA1Bytes], a1_lo[kA1Bytes];
    unsigned char a2_hi[kA2Bytes], a2_lo[kA2Bytes];      // also reused as per-ray scratch between passes
    unsigned char w1_hi[kW1Bytes], w1_lo[kW1Bytes];
    unsigned char w2_hi[kW2Bytes], w2_lo[kW2Bytes];
    int2 taps[128][12];                                       // per tile row: 12 x (element offset, weight bits)
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
    int n_groups;
    int single_pass;          // 1: fast mode (bf16 hi*hi only)
    int srow, scol, splane;   // plane strides in elements (32-bit: host checked)
    float sigma_cull;         // cull/binarize as a threshold on sigma: alpha(sigma) < thr  <=>  sigma < sigma_cull
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int it = 0; it < (1 << 22); ++it) {            // try_wait suspends in hardware; the cap only guards against a hang
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
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

// ------------------------------------------------------------------------------------------ bf16 split
// two floats -> packed bf16x2 (element 0 in the low half)
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
    return d;
}
// (a, b) = hi + lo with hi, lo bf16: 16 significant bits per value
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void split1(float x, unsigned short& hi, unsigned short& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
}
// byte offset of element (row, k) inside a canonical K-major tile whose K-cores are `lbo` bytes apart
__device__ __forceinline__ int tile_off(int row, int k, int lbo) { return (row >> 3) * kSBO + (k >> 3) * lbo + (row & 7) * 16 + (k & 7) * 2; }

// 4 bilinear taps of one plane as (32-bit element offset, weight); invalid taps -> (0, 0).  renderer.py:68-81
__device__ __forceinline__ void plane_taps32(const Geom& g, const FusedArgs& a, int pbase, float ca, float cb, int2* out) {
    const float gx = __fmul_rn(ca, g.coord_scale), gy = __fmul_rn(cb, g.coord_scale);
    const float fx = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)g.W), 1.f), 0.5f);
    const float fy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)g.H), 1.f), 0.5f);
    const bool sane = (fx > -2.f) && (fx < (float)g.W + 1.f) && (fy > -2.f) && (fy < (float)g.H + 1.f);
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float wx1 = fx - x0f, wy1 = fy - y0f;
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = sane ? (int)x0f : -4, y0 = sane ? (int)y0f : -4;
    const bool vx0 = (unsigned)x0 < (unsigned)g.W, vx1 = (unsigned)(x0 + 1) < (unsigned)g.W;
    const bool vy0 = (unsigned)y0 < (unsigned)g.H, vy1 = (unsigned)(y0 + 1) < (unsigned)g.H;
    const int o00 = pbase + y0 * a.srow + x0 * a.scol;
    int4 lo, hi;
    lo.x = (vx0 && vy0) ? o00 : 0;                     lo.y = __float_as_int((vx0 && vy0) ? wx0 * wy0 : 0.f);
    lo.z = (vx1 && vy0) ? o00 + a.scol : 0;            lo.w = __float_as_int((vx1 && vy0) ? wx1 * wy0 : 0.f);
    hi.x = (vx0 && vy1) ? o00 + a.srow : 0;            hi.y = __float_as_int((vx0 && vy1) ? wx0 * wy1 : 0.f);
    hi.z = (vx1 && vy1) ? o00 + a.srow + a.scol : 0;   hi.w = __float_as_int((vx1 && vy1) ? wx1 * wy1 : 0.f);
    reinterpret_cast<int4*>(out)[0] = lo;
    reinterpret_cast<int4*>(out)[1] = hi;
}

// texel quad load: per-lane 64-bit base (view + 4*q channels) + 32-bit element offset.  Written in PTX so the address
// is ONE mad.wide + the load (nvcc otherwise re-derives every address from the kernel argument with 4 extra ops).
template <bool BF16>
__device__ __forceinline__ float4 load_quad32(const void* qbase, int off) {
    float4 r;
    if (BF16) {
        unsigned int lo, hi;
        asm("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %3, 2, %2;\n\tld.global.nc.v2.u32 {%0,%1}, [a];\n\t}" : "=r"(lo), "=r"(hi) : "l"(qbase), "r"(off));
        r = make_float4(__uint_as_float(lo << 16), __uint_as_float(lo & 0xffff0000u), __uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
    } else {
        asm("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %5, 4, %4;\n\tld.global.nc.v4.f32 {%0,%1,%2,%3}, [a];\n\t}"
            : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(qbase), "r"(off));
    }
    return r;
}

// ------------------------------------------------------------------------------------------
template <bool BF16, int S>
__global__ void __launch_bounds__(kThreads, 2) k_render_fused(const FusedArgs a) {
    constexpr int Sf = S;
    constexpr int G = kRows / S;
    constexpr int SORT_P2 = S <= 64 ? 64 : 128;
    extern __shared__ unsigned char smem_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Geom& g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---------------- one-time setup: TMEM, barrier, decoder weights as UMMA B tiles (scalings folded in)
    if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
    if (tid == 32) { mbar_init(&sm.mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < kHidden * kC; i += kThreads) {            // W1' = W1 * gain * log2(e)
        const int n = i / kC, k = i % kC;
        unsigned short hi, lo;
        split1(__fmul_rn(a.w1[i], g.w1_gain) * kLog2e, hi, lo);
        const int off = tile_off(n, k, kLBO_W1);
        *reinterpret_cast<unsigned short*>(sm.w1_hi + off) = hi;
        *reinterpret_cast<unsigned short*>(sm.w1_lo + off) = lo;
    }
    for (int i = tid; i < kN2 * kHidden; i += kThreads) {           // W2' : row 0 * ln2 ; rows 1..32 negated ; rows 33..47 zero
        const int n = i / kHidden, k = i % kHidden;
        unsigned short hi = 0, lo = 0;
        if (n < kOut) {
            const float w = __fmul_rn(a.w2[n * kHidden + k], g.w2_gain);
            split1(n == 0 ? w * kLn2 : -w, hi, lo);
        }
        const int off = tile_off(n, k, kLBO_W2);
        *reinterpret_cast<unsigned short*>(sm.w2_hi + off) = hi;
        *reinterpret_cast<unsigned short*>(sm.w2_lo + off) = lo;
    }
    if (tid < kHidden) sm.b1[tid] = __fmul_rn(a.b1[tid], g.b1_gain) * kLog2e;
    if (tid < kN2) sm.b2[tid] = tid == 0 ? __fmul_rn(a.b2[0], g.b2_gain) : (tid < kOut ? -__fmul_rn(a.b2[tid], g.b2_gain) * kLog2e : 0.f);
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
        const int view = (int)(ray0 / g.M);                          // a group never straddles views (M % G == 0)
        const void* vplanes = BF16 ? (const void*)(reinterpret_cast<const __nv_bfloat16*>(a.planes) + (long long)view * g.stride_view)
                                   : (const void*)(reinterpret_cast<const float*>(a.planes) + (long long)view * g.stride_view);

        // =================================================================== two passes x two tiles
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            float* t_arr = pass == 0 ? sm.t_c : sm.t_f;
            float* sg_arr = pass == 0 ? sm.sg_c : sm.sg_f;
#pragma unroll 1
            for (int tile = 0; tile < 2; ++tile) {
                const int rows_valid = tile == 0 ? 128 : kRows - 128;
                const bool warp_live = warp * 16 < rows_valid;           // 16 rows per warp
                if (warp_live) {
                    // -------------------------------------------- A) taps: one lane pair per row
                    {
                        const int trow = warp * 16 + (lane >> 1), half = lane & 1;
                        const int prow = tile * 128 + trow;
                        const int rl = prow / S, s = prow - rl * S;
                        const long long ray = ray0 + rl;
                        const bool live = ray < a.R;
                        float tval = 0.f;
                        if (pass == 0) {
                            float u = 0.f;
                            if (half == 0 && live) {
                                const long long gidx = ray * S + s;
                                u = a.u_c ? a.u_c[gidx] : philox_uniform(g.seed, (uint64_t)gidx, 0u);
                            }
                            u = __shfl_sync(0xffffffffu, u, lane & ~1);
                            float t0 = 0.f, t1 = 0.f;
                            if (g.ray_mode == P3D_RAYS_AUTOBOX && live) {
                                t0 = a.ray_t0[ray]; t1 = a.ray_t1[ray];
                                if (!(t1 > t0) && a.bounds[4]) { t0 = ordered_to_float(a.bounds[2]); t1 = ordered_to_float(a.bounds[3]); }
                            }
                            tval = coarse_depth(g, s, u, t0, t1);
                        } else {
                            tval = t_arr[prow];
                        }
                        float px = 0.f, py = 0.f, pz = 0.f;
                        if (live) {
                            const float* o = a.ro + ray * 3;
                            const float* d = a.rd + ray * 3;
                            px = __fadd_rn(o[0], __fmul_rn(tval, d[0]));
                            py = __fadd_rn(o[1], __fmul_rn(tval, d[1]));
                            pz = __fadd_rn(o[2], __fmul_rn(tval, d[2]));
                        } else {
                            px = py = pz = 1e30f;                         // all taps invalid -> zero features
                        }
                        int2* tp = sm.taps[trow];
                        if (half == 0) {
                            plane_taps32(g, a, 0, px, py, tp);
                            plane_taps32(g, a, a.splane, px, pz, tp + 4);
                        } else {
                            const bool pm = g.plane_mode == P3D_PLANES_PANIC3D;
                            plane_taps32(g, a, 2 * a.splane, pm ? py : pz, pm ? pz : px, tp + 8);
                            if (pass == 0) t_arr[prow] = tval;
                            sm.xz[prow * 2] = px; sm.xz[prow * 2 + 1] = pz;
                        }
                    }
                    __syncwarp();
                    // -------------------------------------------- B) gather: 4 rows per round, 8 lanes per row
                    {
                        const int sub = lane >> 3, q = lane & 7;
                        const void* qplanes = BF16 ? (const void*)(reinterpret_cast<const __nv_bfloat16*>(vplanes) + 4 * q)
                                                   : (const void*)(reinterpret_cast<const float*>(vplanes) + 4 * q);
#pragma unroll 1
                        for (int round = 0; round < 4; ++round) {
                            const int trow = warp * 16 + round * 4 + sub;
                            const int4* tp = reinterpret_cast<const int4*>(sm.taps[trow]);
                            int4 tk[6];
#pragma unroll
                            for (int k = 0; k < 6; ++k) tk[k] = tp[k];
                            float4 v[12];
#pragma unroll
                            for (int k = 0; k < 6; ++k) {
                                v[2 * k] = load_quad32<BF16>(qplanes, tk[k].x);
                                v[2 * k + 1] = load_quad32<BF16>(qplanes, tk[k].z);
                            }
                            float4 f[3];
#pragma unroll
                            for (int p = 0; p < 3; ++p) {
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    const int4 t2 = tk[2 * p + k];
                                    const float wa = __int_as_float(t2.y), wb = __int_as_float(t2.w);
                                    const float4 va = v[4 * p + 2 * k], vb = v[4 * p + 2 * k + 1];
                                    acc.x = fmaf(va.x, wa, acc.x); acc.y = fmaf(va.y, wa, acc.y); acc.z = fmaf(va.z, wa, acc.z); acc.w = fmaf(va.w, wa, acc.w);
                                    acc.x = fmaf(vb.x, wb, acc.x); acc.y = fmaf(vb.y, wb, acc.y); acc.z = fmaf(vb.z, wb, acc.z); acc.w = fmaf(vb.w, wb, acc.w);
                                }
                                f[p] = acc;
                            }
                            const float third = 1.f / 3.f;               // mean over planes (x 1/3: within 1 ulp of the divide)
                            const float fx = ((f[0].x + f[1].x) + f[2].x) * third, fy = ((f[0].y + f[1].y) + f[2].y) * third;
                            const float fz = ((f[0].z + f[1].z) + f[2].z) * third, fw = ((f[0].w + f[1].w) + f[2].w) * third;
                            uint32_t h01, l01, h23, l23;
                            split2(fx, fy, h01, l01);
                            split2(fz, fw, h23, l23);
                            const int off = tile_off(trow, 4 * q, kLBO_A);
                            *reinterpret_cast<uint2*>(sm.a1_hi + off) = make_uint2(h01, h23);
                            *reinterpret_cast<uint2*>(sm.a1_lo + off) = make_uint2(l01, l23);
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
                // ------------------------------------------------ epilogue 1: softplus2 -> A2 (hi/lo)
                if ((warp & 3) * 32 < rows_valid) {                    // warp-uniform
                    const int trow = (warp & 3) * 32 + lane;
                    const int chunk = warp >> 2;                           // columns [32*chunk, 32*chunk+32)
                    float v[32];
                    tmem_ld32(tmem + lane_base + 32 * chunk, v);
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        uint32_t ph[4], pl[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = c8 * 8 + 2 * e;
                            split2(softplus2(v[j] + sm.b1[32 * chunk + j]), softplus2(v[j + 1] + sm.b1[32 * chunk + j + 1]), ph[e], pl[e]);
                        }
                        const int off = tile_off(trow, 32 * chunk + 8 * c8, kLBO_A);
                        *reinterpret_cast<uint4*>(sm.a2_hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                        *reinterpret_cast<uint4*>(sm.a2_lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
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
                // ------------------------------------------------ sigma = column 0 (+ crop / cull masks as thresholds)
                if (warp < 4 && warp * 32 < rows_valid) {
                    const int prow = tile * 128 + warp * 32 + lane;
                    float sg = tmem_ld1(d2 + lane_base) + sm.b2[0];
                    if (g.crop_on && !((fabsf(sm.xz[prow * 2]) <= g.crop_limit) && (fabsf(sm.xz[prow * 2 + 1]) <= g.crop_limit))) sg = -1e3f;
                    if (g.binarize_on) sg = sg < a.sigma_cull ? -1e3f : 1e3f;
                    else if (g.cull_on && sg < a.sigma_cull) sg = -1e3f;
                    sg_arr[prow] = sg;
                }
                tc_fence_before();
                __syncthreads();
            }  // tile

            if (pass == 0) {
                // -------------------------------------------- importance sampling (renderer.py:328-387), whole CTA:
                //   (1) thread/interval: alpha   (2) warp/ray: transmittance scan -> weights -> pooled pdf -> cdf
                //   (3) thread/fine sample: inverse CDF   (4) thread/fine sample: rank sort -> ascending t_f
                float* i_alpha = scratch;              // [192]
                float* i_fac = scratch + 192;          // [192]
                float* i_w = scratch + 384;            // [192]
                float* i_cdf = scratch + 576;          // [192]
                float* i_tfu = scratch + 768;          // [192] unsorted importance depths
                if (tid < kRows) {
                    const int rl = tid / S, i = tid - rl * S;
                    float alpha = 0.f, fac = 1.f;
                    if (i < S - 1) {
                        const float smid = __fsub_rn(__fmul_rn(__fadd_rn(sm.sg_c[tid], sm.sg_c[tid + 1]), 0.5f), 1.f);
                        alpha = 1.f - ex2_approx(-kLog2e * __fmul_rn(softplus_mufu(smid), sm.t_c[tid + 1] - sm.t_c[tid]));
                        fac = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
                    }
                    i_alpha[tid] = alpha; i_fac[tid] = fac;
                }
                __syncthreads();
                if (warp < G) {
                    const int rl = warp, nb = S - 3;
                    float carry = 1.f;
#pragma unroll
                    for (int c = 0; c < (S + 31) / 32; ++c) {
                        const int i = c * 32 + lane;
                        const float f = i < S - 1 ? i_fac[rl * S + i] : 1.f;
                        const float incl = warp_scan_mul(f, lane);
                        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                        if (lane == 0) excl = 1.f;
                        if (i < S - 1) i_w[rl * S + i] = i_alpha[rl * S + i] * (carry * excl);
                        carry *= __shfl_sync(0xffffffffu, incl, 31);
                    }
                    __syncwarp();
                    float my[(S + 31) / 32];
                    float part = 0.f;
#pragma unroll
                    for (int c = 0; c < (S + 31) / 32; ++c) {
                        const int k = c * 32 + lane;
                        float v = 0.f;
                        if (k < nb) {
                            const float* w = i_w + rl * S + k;          // pooled[i = k+1]
                            v = __fadd_rn(__fadd_rn(__fmul_rn(__fadd_rn(fmaxf(w[0], w[1]), fmaxf(w[1], w[2])), 0.5f), 0.01f), 1e-5f);
                        }
                        my[c] = v; part += v;
                    }
                    const float total = warp_sum(part);
                    float csum = 0.f;
#pragma unroll
                    for (int c = 0; c < (S + 31) / 32; ++c) {
                        const int k = c * 32 + lane;
                        const float incl = warp_scan_add(k < nb ? __fdiv_rn(my[c], total) : 0.f, lane) + csum;
                        if (k < nb) i_cdf[rl * S + k + 1] = incl;
                        csum = __shfl_sync(0xffffffffu, incl, 31);
                    }
                    if (lane == 0) i_cdf[rl * S] = 0.f;
                }
                __syncthreads();
                if (tid < kRows) {
                    const int rl = tid / S, f = tid - rl * S, nb = S - 3;
                    const long long ray = ray0 + rl;
                    float val = INFINITY;
                    if (ray < a.R) {
                        const float u = a.u_f ? a.u_f[ray * Sf + f] : philox_uniform(g.seed, (uint64_t)(ray * Sf + f), 1u);
                        const float* cdf = i_cdf + rl * S;
                        const float* t = sm.t_c + rl * S;
                        int lo = 0, hi = nb + 1;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
                        const int below = max(lo - 1, 0), above = min(lo, nb);
                        const float c0 = cdf[below], c1 = cdf[above];
                        const float b0 = __fmul_rn(0.5f, __fadd_rn(t[below], t[below + 1]));
                        const float b1 = __fmul_rn(0.5f, __fadd_rn(t[above], t[above + 1]));
                        float den = __fsub_rn(c1, c0);
                        if (den < 1e-5f) den = 1.f;
                        val = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(u, c0), den), __fsub_rn(b1, b0)));
                    }
                    i_tfu[tid] = val;
                }
                __syncthreads();
                if (tid < kRows) {                                   // rank sort (stable): position = #smaller + #equal-before
                    const int rl = tid / S, j = tid - rl * S;
                    const float tj = i_tfu[tid];
                    const float4* row = reinterpret_cast<const float4*>(i_tfu + rl * S);
                    int rank = 0;
#pragma unroll 4
                    for (int k4 = 0; k4 < S / 4; ++k4) {
                        const float4 t4 = row[k4];
                        const int k = 4 * k4;
                        rank += (t4.x < tj || (t4.x == tj && k + 0 < j)) + (t4.y < tj || (t4.y == tj && k + 1 < j)) +
                                (t4.z < tj || (t4.z == tj && k + 2 < j)) + (t4.w < tj || (t4.w == tj && k + 3 < j));
                    }
                    sm.t_f[rl * Sf + rank] = tj;
                }
                __syncthreads();
            }
        }  // pass

        // =================================================================== per-ray weights, whole CTA
        // unify_samples (renderer.py:289-301) + weight part of the final MipRayMarcher2 (ray_marcher.py:25-44):
        //   (1) thread/sample: merged rank (coarse before fine on ties)   (2) thread/interval: alpha
        //   (3) warp/ray: transmittance scan -> w_i -> omega_j = (w_{j-1}+w_j)/2, ray outputs
        constexpr int L = S + Sf;
        float* m_t = scratch;                      // [G][L] merged depths
        float* m_sg = scratch + 384;               // [G][L] merged densities
        float* m_alpha = scratch + 768;            // [G][L]
        float* m_fac = scratch + 1152;             // [G][L]
        float* m_om = scratch + 1536;              // [G][L] omega per merged position
        int* m_pos = reinterpret_cast<int*>(scratch + 1920);   // [2][192] merged position of each coarse / fine row
        if (tid < 4 * kRgb) sm.acc[tid >> 5][tid & 31] = 0.f;
        for (int r = tid; r < 2 * kRows; r += kThreads) {
            const bool is_f = r >= kRows;
            const int row = is_f ? r - kRows : r;
            const int rl = row / S, i = row - rl * S;
            const float* tc = sm.t_c + rl * S;
            const float* tf = sm.t_f + rl * Sf;
            const bool rev = tc[0] > tc[S - 1];                     // degenerate 'auto' limits: coarse depths descend
            int pos;
            float tv, sv;
            if (!is_f) {
                const int ci = rev ? S - 1 - i : i;
                tv = tc[ci]; sv = sm.sg_c[rl * S + ci];
                int lo = 0, hi = Sf;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (tf[mid] < tv) lo = mid + 1; else hi = mid; }
                pos = i + lo;
                m_pos[rl * S + ci] = pos;
            } else {
                tv = tf[i]; sv = sm.sg_f[row];
                int lo = 0, hi = S;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (tc[rev ? S - 1 - mid : mid] <= tv) lo = mid + 1; else hi = mid; }
                pos = i + lo;
                m_pos[kRows + row] = pos;
            }
            m_t[rl * L + pos] = tv; m_sg[rl * L + pos] = sv;
        }
        __syncthreads();
        for (int r = tid; r < 2 * kRows; r += kThreads) {
            const int rl = r / L, i = r - rl * L;
            float alpha = 0.f, fac = 1.f;
            if (i < L - 1) {
                const float smid = __fsub_rn(__fmul_rn(__fadd_rn(m_sg[r], m_sg[r + 1]), 0.5f), 1.f);
                alpha = 1.f - ex2_approx(-kLog2e * __fmul_rn(softplus_mufu(smid), m_t[r + 1] - m_t[r]));
                fac = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
            }
            m_alpha[r] = alpha; m_fac[r] = fac;
        }
        __syncthreads();
        if (warp < G && ray0 + warp < a.R) {
            const int rl = warp;
            const long long ray = ray0 + rl;
            float carry = 1.f, acc_w = 0.f, acc_d = 0.f, wprev = 0.f;
#pragma unroll
            for (int c = 0; c < L / 32; ++c) {
                const int i = c * 32 + lane;
                const float incl = warp_scan_mul(m_fac[rl * L + i], lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.f;
                const float wi = m_alpha[rl * L + i] * (carry * excl);      // 0 for i = L-1 (alpha = 0)
                carry *= __shfl_sync(0xffffffffu, incl, 31);
                acc_w += wi;
                if (i < L - 1) acc_d = fmaf(wi, __fmul_rn(__fadd_rn(m_t[rl * L + i], m_t[rl * L + i + 1]), 0.5f), acc_d);
                float wl = __shfl_up_sync(0xffffffffu, wi, 1);               // w_{i-1}
                if (lane == 0) wl = wprev;
                wprev = __shfl_sync(0xffffffffu, wi, 31);
                m_om[rl * L + i] = __fmul_rn(__fadd_rn(wl, wi), 0.5f);
            }
            const float wsum = warp_sum(acc_w), dnum = warp_sum(acc_d);
            const float back = g.white_back ? __fsub_rn(1.f, wsum) : 0.f;
            if (lane < 3) {
                const float v = fmaf(a.ro[ray * 3 + lane], wsum, a.rd[ray * 3 + lane] * dnum);
                a.out_xyz[ray * 3 + lane] = __fsub_rn(__fmul_rn(__fadd_rn(v, back), 2.f), 1.f);
            }
            if (lane == 0) {
                sm.ray_back[rl] = back;
                a.out_depth[ray] = __fdiv_rn(dnum, wsum);
                a.out_wsum[ray] = wsum;
                atomicMin(&a.bounds[0], float_to_ordered(m_t[rl * L]));
                atomicMax(&a.bounds[1], float_to_ordered(m_t[rl * L + L - 1]));
            }
        }
        __syncthreads();
        // =================================================================== colours: sum_j omega_j * rgb_j
        tc_fence_after();
        for (int tt = warp >> 2; tt < 4; tt += 2) {                  // warps 0-3: tiles 0,2   warps 4-7: tiles 1,3
            const int pass = tt >> 1, tile = tt & 1;
            const int rows_valid = tile == 0 ? 128 : kRows - 128;
            if ((warp & 3) * 32 >= rows_valid) continue;
            const int prow = tile * 128 + (warp & 3) * 32 + lane;
            const int rl = prow / S;
            const bool live = ray0 + rl < a.R;
            const float om = live ? m_om[rl * L + m_pos[pass * kRows + prow]] : 0.f;
            const float ca = g.force_sigmoid ? om : 1.002f * om, cb = g.force_sigmoid ? 0.f : -0.001f * om;
            float v[32];
            tmem_ld32(tmem + 64 + kN2 * tt + 1 + lane_base, v);
#pragma unroll
            for (int c = 0; c < kRgb; ++c)                           // omega * (sigmoid(o) [*1.002 - 0.001]),  z = -o*log2(e)
                v[c] = fmaf(rcp_approx(1.f + ex2_approx(v[c] + sm.b2[1 + c])), ca, cb);
            // rows of one warp belong to at most two rays (S >= 32): reduce each separately
            const int rl_lo = __shfl_sync(0xffffffffu, rl, 0), rl_hi = __shfl_sync(0xffffffffu, rl, 31);
            for (int target = rl_lo; target <= rl_hi; ++target) {
                float r[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) r[c] = (S % 32 == 0 || rl == target) ? v[c] : 0.f;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
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
    if (!((g.S == 96 || g.S == 48) && (g.Sf == g.S) && ((long long)g.M % (kRows / g.S) == 0))) return false;
    // 32-bit tap offsets inside one view's tri-plane
    const long long span = 2 * g.stride_plane + (long long)(g.H - 1) * g.stride_row + (long long)(g.W - 1) * g.stride_col + kC;
    return g.stride_plane >= 0 && g.stride_row >= 0 && g.stride_col >= 0 && span < (1ll << 31);
}

int render_forward_fused(const Geom& g, const p3d_render_params* p, const void* planes, const float* w1, const float* b1,
                         const float* w2, const float* b2, const float* ro, const float* rd, const float* u_c,
                         const float* u_f, const Workspace& ws, float* out_rgb, float* out_depth, float* out_wsum,
                         float* out_xyz, cudaStream_t stream) {
    if (!fused_supported(g)) {
        set_error("fused tcgen05 renderer supports depth_resolution == depth_resolution_importance in {48, 96} (got %d, %d) "
                  "with non-negative plane strides below 2^31 elements per view", g.S, g.Sf);
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
    const int G = kRows / g.S;
    a.R = R; a.n_groups = (int)((R + G - 1) / G);
    a.single_pass = p->mlp_mode == P3D_MLP_TC_BF16;
    a.srow = (int)g.stride_row; a.scol = (int)g.stride_col; a.splane = (int)g.stride_plane;
    // alpha(sigma) = 1 - exp(-softplus(sigma - 1)) < thr   <=>   sigma < 1 + log(expm1(-log1p(-thr)))   (monotone)
    if (g.cull_on || g.binarize_on) {
        const double thr = (double)g.cull_thresh;
        a.sigma_cull = thr >= 1.0 ? INFINITY : (thr <= 0.0 ? -INFINITY : (float)(1.0 + log(expm1(-log1p(-thr)))));
    }
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        P3D_CUDA_TRY(cudaGetDevice(&dev));
        P3D_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = sizeof(FusedSmem) + 1024;
    void (*kern)(FusedArgs) = nullptr;
    if (g.S == 96) kern = p->planes_bf16 ? k_render_fused<true, 96> : k_render_fused<false, 96>;
    else kern = p->planes_bf16 ? k_render_fused<true, 48> : k_render_fused<false, 48>;
    P3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = a.n_groups < 2 * n_sm ? a.n_groups : 2 * n_sm;
    {
        ProfileScope prof(PROF_FUSED, stream);
        kern<<<grid, kThreads, smem, stream>>>(a);
        P3D_LAUNCH_CHECK();
    }
    if (p->defer_depth_clamp) return P3D_OK;
    return launch_depth_finalize(out_depth, R, ws.bounds, stream);
}

}  // namespace p3d
