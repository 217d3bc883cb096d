// Gather-role microbenchmark (round 2): how fast can one SM's gather warps pull the 12 x 128 B bilinear taps of the
// tri-plane renderer out of L2, in isolation (no MMA / epilogue roles), for the bench geometry
// (8 views x 128x128 rays x 192 samples, 3 x 512 x 512 x 32 ch planes per view)?
//
//   MODE 0  round-synchronous LDG.128 (the r1 kernel's gather): 12 loads / lane, wait, lerp, store
//   MODE 1  same, invalid (out-of-plane) taps predicated off
//   MODE 2  rotating registers: the loads of round r+1 are issued slot by slot as round r is consumed
//   MODE 3  LDGSTS (cp.async 16 B) into a DEPTH-deep shared ring per warp, lerp from shared
//   MODE 4  cp.async.bulk (128 B per tap, mbarrier complete_tx) into the same ring
// HALF = 1: planes stored as 16-bit (64 B per tap) - measures what halving the L2 traffic buys.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kViews = 8, kRes = 128, kRays = kRes * kRes, kSamples = 192, kP = 512, kC = 32;
constexpr long long kPlaneElems = (long long)kP * kP * kC;

struct View { float o[3]; float r[9]; };
struct Args {
    const void* planes;      // (view, plane, y, x, 32) fp32 or fp16
    View views[kViews];
    float* sink;             // per-CTA dump so the work is not dead
    unsigned long long* cyc; // per-CTA cycles
    int units;               // work units (32 samples each)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 4 taps of one plane: (element offset or -1, weight)
__device__ __forceinline__ void plane_taps(int pbase, float ca, float cb, int2* out) {
    const float cs = 2.f / 0.7f;
    const float fx = ((ca * cs + 1.f) * (float)kP - 1.f) * 0.5f, fy = ((cb * cs + 1.f) * (float)kP - 1.f) * 0.5f;
    const bool sane = (fx > -2.f) && (fx < kP + 1.f) && (fy > -2.f) && (fy < kP + 1.f);
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float wx1 = fx - x0f, wy1 = fy - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = sane ? (int)x0f : -4, y0 = sane ? (int)y0f : -4;
    const bool vx0 = (unsigned)x0 < (unsigned)kP, vx1 = (unsigned)(x0 + 1) < (unsigned)kP;
    const bool vy0 = (unsigned)y0 < (unsigned)kP, vy1 = (unsigned)(y0 + 1) < (unsigned)kP;
    const int o00 = pbase + (y0 * kP + x0) * kC;
    out[0] = make_int2((vx0 && vy0) ? o00 : -1, __float_as_int((vx0 && vy0) ? wx0 * wy0 : 0.f));
    out[1] = make_int2((vx1 && vy0) ? o00 + kC : -1, __float_as_int((vx1 && vy0) ? wx1 * wy0 : 0.f));
    out[2] = make_int2((vx0 && vy1) ? o00 + kP * kC : -1, __float_as_int((vx0 && vy1) ? wx0 * wy1 : 0.f));
    out[3] = make_int2((vx1 && vy1) ? o00 + kP * kC + kC : -1, __float_as_int((vx1 && vy1) ? wx1 * wy1 : 0.f));
}

template <bool HALF, bool PRED>
__device__ __forceinline__ float4 load_quad(const void* qbase, int off) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HALF) {
        unsigned int lo = 0, hi = 0;
        if (PRED) {
            asm volatile("{\n\t.reg .u64 a;\n\t.reg .pred p;\n\tsetp.ge.s32 p, %3, 0;\n\tmad.wide.s32 a, %3, 2, %2;\n\t@p ld.global.nc.v2.u32 {%0,%1}, [a];\n\t}"
                         : "+r"(lo), "+r"(hi) : "l"(qbase), "r"(off));
        } else {
            const int o2 = off < 0 ? 0 : off;
            asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %3, 2, %2;\n\tld.global.nc.v2.u32 {%0,%1}, [a];\n\t}" : "=r"(lo), "=r"(hi) : "l"(qbase), "r"(o2));
        }
        const __half2 a = *reinterpret_cast<__half2*>(&lo), b = *reinterpret_cast<__half2*>(&hi);
        const float2 fa = __half22float2(a), fb = __half22float2(b);
        r = make_float4(fa.x, fa.y, fb.x, fb.y);
    } else {
        if (PRED) {
            asm volatile("{\n\t.reg .u64 a;\n\t.reg .pred p;\n\tsetp.ge.s32 p, %5, 0;\n\tmad.wide.s32 a, %5, 4, %4;\n\t@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [a];\n\t}"
                         : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w) : "l"(qbase), "r"(off));
        } else {
            const int o2 = off < 0 ? 0 : off;
            asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.s32 a, %5, 4, %4;\n\tld.global.nc.v4.f32 {%0,%1,%2,%3}, [a];\n\t}"
                         : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(qbase), "r"(o2));
        }
    }
    return r;
}

__device__ __forceinline__ void fma4(float4& acc, const float4 v, float w) {
    acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
    uint32_t d; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0)); return d;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// finish one row-quad: mean of 3 planes, split, store 8 B hi + 8 B lo into the warp's A1-like sink
__device__ __forceinline__ void finish(const float4 (&f)[3], unsigned char* a1, int row, int qd) {
    const float third = 1.f / 3.f;
    const float fx = ((f[0].x + f[1].x) + f[2].x) * third, fy = ((f[0].y + f[1].y) + f[2].y) * third;
    const float fz = ((f[0].z + f[1].z) + f[2].z) * third, fw = ((f[0].w + f[1].w) + f[2].w) * third;
    uint32_t h01, l01, h23, l23;
    split2(fx, fy, h01, l01); split2(fz, fw, h23, l23);
    // K-major no-swizzle core matrices with the k-chunk stride offset by 64 B (conflict-free for 4 rows x 8 quads)
    const int off = (row >> 3) * 128 + (qd >> 1) * (512 + 64) + (row & 7) * 16 + (qd & 1) * 8;
    *reinterpret_cast<uint2*>(a1 + off) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(a1 + 4096 + off) = make_uint2(l01, l23);
}

template <int MODE, bool HALF, int NW, int DEPTH>
__global__ void __launch_bounds__(NW * 32, 1) k_gather(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kTapBytes = HALF ? 64 : 128;
    constexpr int kRoundBytes = 4 * 12 * kTapBytes;                       // 4 rows x 12 taps
    // layout: per warp: taps[32][12] int2 (3 KB) | a1 sink (8 KB + slack 1 KB) | ring DEPTH x round | mbar
    constexpr int kWarpBytes = 3072 + 9216 + ((MODE >= 3) ? DEPTH * kRoundBytes : 0) + 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbase = smem + (size_t)warp * kWarpBytes;
    int2(*taps)[12] = reinterpret_cast<int2(*)[12]>(wbase);
    unsigned char* a1 = wbase + 3072;
    unsigned char* ring = wbase + 3072 + 9216;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(wbase + kWarpBytes - 64);
    if (MODE == 4 && lane == 0)
        for (int d = 0; d < DEPTH; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + d)) : "memory");
    if (MODE >= 3) for (int i = lane * 16; i < DEPTH * kRoundBytes; i += 512) *reinterpret_cast<uint4*>(ring + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long t_start = clock64();
    const int sub = lane >> 3, qd = lane & 7;
    uint32_t phase_bits = 0;
    for (int unit = blockIdx.x * NW + warp; unit < a.units; unit += gridDim.x * NW) {
        // unit -> (ray, chunk of 32 samples)
        const int chunk = unit % 6, gray = unit / 6;
        const int view = gray / kRays, pix = gray % kRays;
        const View& V = a.views[view];
        // ---- taps of 32 rows: lane = row
        {
            const int px_ = pix % kRes, py_ = pix / kRes;
            const float tanh_ = 0.2679491924f;                              // tan(15 deg)
            const float cx = ((px_ + 0.5f) / kRes * 2.f - 1.f) * tanh_, cy = ((py_ + 0.5f) / kRes * 2.f - 1.f) * tanh_;
            float d0 = V.r[0] * cx + V.r[1] * cy - V.r[2], d1 = V.r[3] * cx + V.r[4] * cy - V.r[5], d2 = V.r[6] * cx + V.r[7] * cy - V.r[8];
            const float inv = rsqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            d0 *= inv; d1 *= inv; d2 *= inv;
            const int s = chunk * 32 + lane;
            const float t = 0.5f + (s + 0.37f) * (1.f / 192.f);
            const float x = V.o[0] + t * d0, y = V.o[1] + t * d1, z = V.o[2] + t * d2;
            plane_taps(0, x, y, taps[lane]);
            plane_taps((int)kPlaneElems, x, z, taps[lane] + 4);
            plane_taps(2 * (int)kPlaneElems, y, z, taps[lane] + 8);
        }
        __syncwarp();
        const void* qplanes = HALF ? (const void*)(reinterpret_cast<const __half*>(a.planes) + (long long)view * 3 * kPlaneElems + 4 * qd)
                                   : (const void*)(reinterpret_cast<const float*>(a.planes) + (long long)view * 3 * kPlaneElems + 4 * qd);
        if (MODE == 0 || MODE == 1) {
#pragma unroll 1
            for (int round = 0; round < 8; ++round) {
                const int row = round * 4 + sub;
                const int4* tp = reinterpret_cast<const int4*>(taps[row]);
                int4 tk[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) tk[k] = tp[k];
                float4 v[12];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    v[2 * k] = load_quad<HALF, MODE == 1>(qplanes, tk[k].x);
                    v[2 * k + 1] = load_quad<HALF, MODE == 1>(qplanes, tk[k].z);
                }
                float4 f[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    f[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        fma4(f[p], v[4 * p + 2 * k], __int_as_float(tk[2 * p + k].y));
                        fma4(f[p], v[4 * p + 2 * k + 1], __int_as_float(tk[2 * p + k].w));
                    }
                }
                finish(f, a1, row, qd);
            }
        } else if (MODE == 2) {
            float4 v[12];
            {
                const int4* tp = reinterpret_cast<const int4*>(taps[sub]);
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int4 t2 = tp[k];
                    v[2 * k] = load_quad<HALF, true>(qplanes, t2.x);
                    v[2 * k + 1] = load_quad<HALF, true>(qplanes, t2.z);
                }
            }
#pragma unroll 1
            for (int round = 0; round < 8; ++round) {
                const int row = round * 4 + sub;
                const int nrow = (round < 7) ? row + 4 : row;
                const int4* tp = reinterpret_cast<const int4*>(taps[row]);
                const int4* tn = reinterpret_cast<const int4*>(taps[nrow]);
                float4 f[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    f[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int4 t2 = tp[2 * p + k];
                        const int4 n2 = tn[2 * p + k];
                        fma4(f[p], v[4 * p + 2 * k], __int_as_float(t2.y));
                        fma4(f[p], v[4 * p + 2 * k + 1], __int_as_float(t2.w));
                        if (round < 7) {
                            v[4 * p + 2 * k] = load_quad<HALF, true>(qplanes, n2.x);
                            v[4 * p + 2 * k + 1] = load_quad<HALF, true>(qplanes, n2.z);
                        }
                    }
                }
                finish(f, a1, row, qd);
            }
        } else {
            // ---- staged: ring of DEPTH rounds; lane (sub, qd) copies / reads its own 16 B (8 B) of each of its row's 12 taps
            auto issue = [&](int round) {
                const int row = round * 4 + sub;
                unsigned char* st = ring + (round % DEPTH) * kRoundBytes;
                if (MODE == 3) {
                    const int4* tp = reinterpret_cast<const int4*>(taps[row]);
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int4 t2 = tp[k];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int off = h ? t2.z : t2.x;
                            const uint32_t dst = smem_u32(st + (sub * 12 + 2 * k + h) * kTapBytes + qd * (kTapBytes / 8));
                            const int o2 = off < 0 ? 0 : off;
                            const int sz = off < 0 ? 0 : (kTapBytes / 8);                 // src-size 0 -> zero fill
                            if (HALF) {
                                const void* src = reinterpret_cast<const __half*>(qplanes) + o2;
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
                            } else {
                                const void* src = reinterpret_cast<const float*>(qplanes) + o2;
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
                            }
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                } else {
                    // bulk: 48 (row, tap) copies per round spread over the 32 lanes; invalid taps are skipped (weight 0,
                    // stale data finite because the ring starts zeroed and only ever holds plane values)
                    uint32_t bytes = 0;
                    const int2* flat = &taps[round * 4][0];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int idx = lane + 32 * j;
                        if (idx < 48 && flat[idx].x >= 0) bytes += kTapBytes;
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                    const uint32_t bar = smem_u32(mbar + (round % DEPTH));
                    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int idx = lane + 32 * j;
                        if (idx < 48) {
                            const int off = flat[idx].x;
                            if (off >= 0) {
                                const void* src = HALF ? (const void*)(reinterpret_cast<const __half*>(a.planes) + (long long)view * 3 * kPlaneElems + off)
                                                       : (const void*)(reinterpret_cast<const float*>(a.planes) + (long long)view * 3 * kPlaneElems + off);
                                asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                             ::"r"(smem_u32(st + idx * kTapBytes)), "l"(src), "r"(kTapBytes), "r"(bar) : "memory");
                            }
                        }
                    }
                }
            };
#pragma unroll 1
            for (int r = 0; r < DEPTH - 1; ++r) issue(r);
#pragma unroll 1
            for (int round = 0; round < 8; ++round) {
                if (round + DEPTH - 1 < 8) issue(round + DEPTH - 1);
                else if (MODE == 3) asm volatile("cp.async.commit_group;" ::: "memory");
                if (MODE == 3) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
                } else {
                    const int slot = round % DEPTH;
                    const uint32_t bar = smem_u32(mbar + slot), par = (phase_bits >> slot) & 1u;
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
                    phase_bits ^= 1u << slot;
                }
                const int row = round * 4 + sub;
                const unsigned char* st = ring + (round % DEPTH) * kRoundBytes + sub * 12 * kTapBytes + qd * (kTapBytes / 8);
                const int4* tp = reinterpret_cast<const int4*>(taps[row]);
                float4 f[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    f[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int2 t1 = reinterpret_cast<const int2*>(tp)[4 * p + k];
                        float4 vv;
                        if (HALF) {
                            const uint2 raw = *reinterpret_cast<const uint2*>(st + (4 * p + k) * kTapBytes);
                            const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), fb = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                            vv = make_float4(fa.x, fa.y, fb.x, fb.y);
                        } else {
                            vv = *reinterpret_cast<const float4*>(st + (4 * p + k) * kTapBytes);
                        }
                        fma4(f[p], vv, __int_as_float(t1.y));
                    }
                }
                finish(f, a1, row, qd);
                __syncwarp();
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) a.cyc[blockIdx.x] = (unsigned long long)(clock64() - t_start);
    // keep the sink alive
    if (a.sink) a.sink[blockIdx.x * blockDim.x + threadIdx.x] = reinterpret_cast<float*>(smem + (size_t)warp * kWarpBytes + 3072)[lane];
}

template <int MODE, bool HALF, int NW, int DEPTH>
static void run(const char* name, Args a, int n_sm) {
    constexpr int kTapBytes = HALF ? 64 : 128;
    constexpr int kWarpBytes = 3072 + 9216 + ((MODE >= 3) ? DEPTH * 4 * 12 * kTapBytes : 0) + 64;
    const size_t smem = (size_t)NW * kWarpBytes;
    if (smem > 227 * 1024) { printf("%-34s skipped (smem %zu)\n", name, smem); return; }
    auto kern = k_gather<MODE, HALF, NW, DEPTH>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<n_sm, NW * 32, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    unsigned long long* h = (unsigned long long*)malloc(n_sm * 8);
    CK(cudaMemcpy(h, a.cyc, n_sm * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < n_sm; ++i) if ((double)h[i] > mx) mx = (double)h[i];
    free(h);
    const double samples = (double)a.units * 32.0;
    printf("%-34s NW=%2d D=%d smem=%6zu  %7.3f ms  %6.1f cyc/sample/SM  %7.1f GB/s tap bytes\n", name, NW, DEPTH, smem, best,
           mx / (samples / n_sm), samples * 12 * kTapBytes / (best * 1e-3) / 1e9);
}

int main() {
    int n_sm = 0; CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    Args a{};
    const long long total = (long long)kViews * 3 * kPlaneElems;
    void *p32, *p16;
    CK(cudaMalloc(&p32, total * 4)); CK(cudaMalloc(&p16, total * 2));
    CK(cudaMemset(p32, 0, total * 4)); CK(cudaMemset(p16, 0, total * 2));
    CK(cudaMalloc(&a.sink, 148 * 1024 * 4)); CK(cudaMalloc(&a.cyc, 148 * 8));
    for (int v = 0; v < kViews; ++v) {
        const float az = (float)(v * 30 - 180) * 3.14159265f / 180.f;
        const float c = cosf(az), s = sinf(az);
        View V{{s, 0.f, c}, {c, 0.f, s, 0.f, 1.f, 0.f, -s, 0.f, c}};     // camera on a circle of radius 1 looking at the origin
        a.views[v] = V;
    }
    a.units = kViews * kRays * 6;
    printf("gather microbenchmark: %d SMs, %d units x 32 samples = %.1f M samples, fp32 tap bytes %.1f GB\n", n_sm, a.units,
           a.units * 32.0 / 1e6, a.units * 32.0 * 1536 / 1e9);
    a.planes = p32;
    run<0, false, 12, 1>("fp32 round-sync LDG (r1)", a, n_sm);
    run<1, false, 12, 1>("fp32 round-sync LDG pred", a, n_sm);
    run<1, false, 16, 1>("fp32 round-sync LDG pred", a, n_sm);
    run<1, false, 20, 1>("fp32 round-sync LDG pred", a, n_sm);
    run<2, false, 8, 1>("fp32 rotating LDG pred", a, n_sm);
    run<2, false, 12, 1>("fp32 rotating LDG pred", a, n_sm);
    run<2, false, 16, 1>("fp32 rotating LDG pred", a, n_sm);
    run<2, false, 20, 1>("fp32 rotating LDG pred", a, n_sm);
    run<3, false, 8, 2>("fp32 LDGSTS ring", a, n_sm);
    run<3, false, 8, 3>("fp32 LDGSTS ring", a, n_sm);
    run<3, false, 12, 2>("fp32 LDGSTS ring", a, n_sm);
    run<3, false, 6, 4>("fp32 LDGSTS ring", a, n_sm);
    run<4, false, 8, 2>("fp32 bulk ring", a, n_sm);
    run<4, false, 8, 3>("fp32 bulk ring", a, n_sm);
    run<4, false, 12, 2>("fp32 bulk ring", a, n_sm);
    run<4, false, 6, 4>("fp32 bulk ring", a, n_sm);
    run<4, false, 4, 6>("fp32 bulk ring", a, n_sm);
    a.planes = p16;
    run<1, true, 12, 1>("fp16 round-sync LDG pred", a, n_sm);
    run<2, true, 12, 1>("fp16 rotating LDG pred", a, n_sm);
    run<2, true, 16, 1>("fp16 rotating LDG pred", a, n_sm);
    run<3, true, 12, 3>("fp16 LDGSTS ring", a, n_sm);
    run<4, true, 12, 3>("fp16 bulk ring", a, n_sm);
    run<4, true, 8, 6>("fp16 bulk ring", a, n_sm);
    return 0;
}
