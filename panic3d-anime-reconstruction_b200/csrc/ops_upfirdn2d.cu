// upfirdn2d for sm_100a: zero-insert upsample -> pad/crop -> 2-D FIR -> decimate, x gain.
//
//   y[n,c,oy,ox] = gain * sum_{ky,kx} P[oy*downy + ky, ox*downx + kx] * w[ky][kx]
//   P[py,px]     = x[(py-pady0)/upy, (px-padx0)/upx] where both divisions are exact and in range, else 0
//   w[ky][kx]    = flip ? f[ky][kx] : f[fH-1-ky][fW-1-kx]          (reference semantics: ops/upfirdn2d.py:169-213,
//                                                                    kernels ops/upfirdn2d.cu:33-204)
// The op is HBM-bound (read x once, write y once, a handful of FMAs per output), so the B200 kernel is built
// around memory behaviour rather than per-(up,down,filter) template specialisations (the reference has ~100):
//   * tiled path (input rows contiguous along W): a CTA produces a 64x16 output tile of one (n,c) image; the
//     input footprint is staged ONCE in shared memory with coalesced row reads, zero-filled outside the image,
//     so the polyphase inner loop is pure LDS+FFMA with no bounds checks; only the taps whose phase hits a
//     real sample are visited (up>1 never multiplies the inserted zeros);
//   * strided path (channels-last or any other layout): same maths straight from global memory through L1,
//     with threads mapped channel-fastest so channels-last stays coalesced;
//   * fp32 / fp16 / bf16 / fp64 storage, fp32 (fp64) accumulation, 64-bit safe indexing.
#include "p3d_common.cuh"
#include "../../include/p3d_ops.h"

namespace p3d {
namespace {

template <typename T> struct Px;
template <> struct Px<float> { using acc = float; static __device__ float ld(float v) { return v; } static __device__ float st(float v) { return v; } };
template <> struct Px<double> { using acc = double; static __device__ double ld(double v) { return v; } static __device__ double st(double v) { return v; } };
template <> struct Px<__half> { using acc = float; static __device__ float ld(__half v) { return __half2float(v); } static __device__ __half st(float v) { return __float2half_rn(v); } };
template <> struct Px<__nv_bfloat16> { using acc = float; static __device__ float ld(__nv_bfloat16 v) { return __bfloat162float(v); } static __device__ __nv_bfloat16 st(float v) { return __float2bfloat16_rn(v); } };

struct UpfirdnParams {
    const void* x;
    const float* f;
    void* y;
    int N, C, inH, inW, outH, outW, fH, fW;
    long long xs[4], ys[4];          // element strides {n, c, h, w}
    long long fsh, fsw;
    int upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
    int tilesX, tilesY, inTileW, inTileH;
};

constexpr int kTileW = 64, kTileH = 16, kThreads = 256, kMaxTaps = 1024;

__device__ __forceinline__ int floor_div(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ int pos_mod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// ------------------------------------------------------------------------------------------ tiled (W-contiguous input)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_upfirdn2d_tiled(const UpfirdnParams p) {
    using S = typename Px<T>::acc;
    extern __shared__ unsigned char smem_raw[];
    float* s_w = reinterpret_cast<float*>(smem_raw);                         // fH*fW taps (gain, flip folded in)
    S* s_x = reinterpret_cast<S*>(smem_raw + ((p.fH * p.fW * 4 + 15) & ~15));   // inTileH x inTileW
    const int tid = threadIdx.x;
    for (int i = tid; i < p.fH * p.fW; i += kThreads) {
        const int ky = i / p.fW, kx = i - ky * p.fW;
        const int sy = p.flip ? ky : p.fH - 1 - ky, sx = p.flip ? kx : p.fW - 1 - kx;
        s_w[i] = p.f[sy * p.fsh + sx * p.fsw] * p.gain;
    }
    const int tile = blockIdx.x;
    const int tyi = tile / p.tilesX, txi = tile - tyi * p.tilesX;
    const int ox0 = txi * kTileW, oy0 = tyi * kTileH;
    const int ix0 = floor_div(ox0 * p.downx - p.padx0, p.upx);               // first input column/row that can contribute
    const int iy0 = floor_div(oy0 * p.downy - p.pady0, p.upy);
    const T* xp = reinterpret_cast<const T*>(p.x);
    T* yp = reinterpret_cast<T*>(p.y);
    for (long long img = blockIdx.y; img < (long long)p.N * p.C; img += gridDim.y) {
        const int n = (int)(img / p.C), c = (int)(img - (long long)n * p.C);
        const T* xi = xp + n * p.xs[0] + c * p.xs[1];
        __syncthreads();                                                     // previous image's reads are done
        for (int i = tid; i < p.inTileH * p.inTileW; i += kThreads) {
            const int r = i / p.inTileW, q = i - r * p.inTileW;
            const int iy = iy0 + r, ix = ix0 + q;
            S v = (S)0;
            if (iy >= 0 && iy < p.inH && ix >= 0 && ix < p.inW) v = Px<T>::ld(xi[iy * p.xs[2] + ix * p.xs[3]]);
            s_x[i] = v;
        }
        __syncthreads();
        const int tx = tid & (kTileW - 1);
        const int ox = ox0 + tx;
        if (ox < p.outW) {
            const int ux0 = ox * p.downx - p.padx0;                          // up-sampled coordinate of tap kx = 0
            const int kx0 = pos_mod(-ux0, p.upx);
            const int sx0 = (ux0 + kx0) / p.upx - ix0;                       // exact division
#pragma unroll 1
            for (int ty = tid / kTileW; ty < kTileH; ty += kThreads / kTileW) {
                const int oy = oy0 + ty;
                if (oy >= p.outH) break;
                const int uy0 = oy * p.downy - p.pady0;
                const int ky0 = pos_mod(-uy0, p.upy);
                const int sy0 = (uy0 + ky0) / p.upy - iy0;
                S acc = (S)0;
                for (int ky = ky0, sy = sy0; ky < p.fH; ky += p.upy, ++sy) {
                    const S* row = s_x + sy * p.inTileW + sx0;
                    const float* wr = s_w + ky * p.fW;
                    for (int kx = kx0, j = 0; kx < p.fW; kx += p.upx, ++j) acc += row[j] * (S)wr[kx];
                }
                yp[n * p.ys[0] + c * p.ys[1] + oy * p.ys[2] + ox * p.ys[3]] = Px<T>::st(acc);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ fast path (4x4 FIR)
// The shapes panic3d's generator/discriminator hit: a 4x4 [1,3,3,1]^2 filter at (up,down) = (1,1) blur
// after a transposed conv, (2,1) skip-image upsample, (1,2) discriminator downsample; W-contiguous in and out.
// Persistent CTAs walk output tiles (128 x 32|16) with a two-stage cp.async pipeline:
//   * the input footprint of tile i+1 streams into shared memory as ALIGNED 16-byte vectors while tile i is being
//     computed.  Rows of odd-width images (513 x fp16 = 1026 B) are not 16-byte aligned, so each shared row is
//     shifted by the row's own misalignment a_r = (address / sizeof(T)) mod VEC: the aligned global vector k of a row
//     lands on the aligned shared vector k, and element q of the footprint sits at column a_r + q.  Vectors that
//     straddle the image border (<= 2 per row) are assembled element-wise; rows above/below the image are zeros;
//   * a thread owns 2 adjacent output columns x RPT consecutive rows.  Every input sample of its window is read from
//     shared memory ONCE (conflict-free: a warp reads consecutive elements of one row) and scattered into the
//     accumulators of the outputs it feeds - tap indices ky = sr*UP - r*DOWN, kx = sc*UP - e*DOWN are compile-time,
//     so the polyphase structure (up=2 visits 2x2 of the 4x4 taps per output) costs nothing and the 16 taps live
//     in registers;
//   * fp32 accumulation for every storage type, one rounding at the store (as the reference's scalar_t=float path).
template <typename T, int UP, int DOWN, int F>
struct FastCfg {
    static constexpr int VEC = 16 / (int)sizeof(T);
    static constexpr int TW = 128;
    static constexpr int RPT = (DOWN == 1) ? 8 : 4;
    static constexpr int TH = 4 * RPT;
    static constexpr int IN_W = ((TW - 1) * DOWN + F - 1) / UP + 1;
    static constexpr int IN_H = ((TH - 1) * DOWN + F - 1) / UP + 1;
    static constexpr int NIN = (DOWN + F - 1) / UP + 1;
    static constexpr int NINROWS = ((RPT - 1) * DOWN + F - 1) / UP + 1;
    static constexpr int PITCH = (IN_W + VEC - 1 + VEC - 1) / VEC * VEC;
    static constexpr int NV = PITCH / VEC;
    static constexpr size_t kBufBytes = (size_t)IN_H * PITCH * sizeof(T);
    static constexpr size_t kSmem = 2 * kBufBytes + 2 * IN_H * sizeof(int);
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int UP, int DOWN, int F>
__global__ void __launch_bounds__(kThreads) k_upfirdn2d_fast(const UpfirdnParams p, long long total_tiles, int pair_store) {
    using Cfg = FastCfg<T, UP, DOWN, F>;
    using V = uint4;
    constexpr int VEC = Cfg::VEC, PITCH = Cfg::PITCH, NV = Cfg::NV, IN_H = Cfg::IN_H, RPT = Cfg::RPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_buf[2] = {reinterpret_cast<T*>(smem_raw), reinterpret_cast<T*>(smem_raw + Cfg::kBufBytes)};
    int* s_a = reinterpret_cast<int*>(smem_raw + 2 * Cfg::kBufBytes);           // [2][IN_H] row shifts
    const int tid = threadIdx.x;
    const T* xp = reinterpret_cast<const T*>(p.x);
    T* yp = reinterpret_cast<T*>(p.y);

    float w[F * F];
#pragma unroll
    for (int i = 0; i < F * F; ++i) {
        const int ky = i / F, kx = i % F;
        const int sy = p.flip ? ky : F - 1 - ky, sx = p.flip ? kx : F - 1 - kx;
        w[i] = p.f[sy * p.fsh + sx * p.fsw] * p.gain;
    }
    const long long tiles_per_img = (long long)p.tilesX * p.tilesY;

    auto issue = [&](long long t, int b) {
        const long long img = t / tiles_per_img;
        const int rem = (int)(t - img * tiles_per_img);
        const int tyi = rem / p.tilesX, txi = rem - tyi * p.tilesX;
        const int n = (int)(img / p.C), c = (int)(img - (long long)n * p.C);
        const int ix0 = (txi * Cfg::TW * DOWN - p.padx0) / UP, iy0 = (tyi * Cfg::TH * DOWN - p.pady0) / UP;   // exact
        const T* xi = xp + n * p.xs[0] + c * p.xs[1];
        T* sb = s_buf[b];
        for (int i = tid; i < IN_H * NV; i += kThreads) {
            const int r = i / NV, k = i - r * NV;
            const int iy = iy0 + r;
            V* dst = reinterpret_cast<V*>(sb + r * PITCH + k * VEC);
            if (iy < 0 || iy >= p.inH) {
                *dst = make_uint4(0u, 0u, 0u, 0u);
                if (k == 0) s_a[b * IN_H + r] = 0;
                continue;
            }
            const T* row0 = xi + (long long)iy * p.xs[2] + ix0;              // footprint element q = 0 (may be outside the row)
            const int a = (int)((reinterpret_cast<uintptr_t>(row0) / sizeof(T)) & (VEC - 1));
            if (k == 0) s_a[b * IN_H + r] = a;
            const int ix = ix0 + k * VEC - a;                                // image column of the vector's first element
            const T* src = row0 + (k * VEC - a);                             // 16-byte aligned
            if (ix >= 0 && ix + VEC <= p.inW) {
                cp_async16(dst, src);
            } else {
                alignas(16) T tmp[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) tmp[j] = (ix + j >= 0 && ix + j < p.inW) ? src[j] : Px<T>::st(0.f);
                *dst = *reinterpret_cast<const V*>(tmp);
            }
        }
    };

    long long t = blockIdx.x;
    if (t >= total_tiles) return;
    issue(t, 0);
    cp_async_commit();
    const int tx = tid & 63, tr = tid >> 6;
    const int x0 = 2 * tx, y0 = tr * RPT;
    const int bx = (x0 * DOWN) / UP, by = (y0 * DOWN) / UP;
    int b = 0;
    for (; t < total_tiles; t += gridDim.x, b ^= 1) {
        const long long tn = t + gridDim.x;
        if (tn < total_tiles) issue(tn, b ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        float acc[RPT][2];
#pragma unroll
        for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = 0.f;
        const T* sb = s_buf[b];
#pragma unroll
        for (int sr = 0; sr < Cfg::NINROWS; ++sr) {
            const T* rowp = sb + (by + sr) * PITCH + s_a[b * IN_H + by + sr] + bx;
            float v[Cfg::NIN];
#pragma unroll
            for (int sc = 0; sc < Cfg::NIN; ++sc) v[sc] = Px<T>::ld(rowp[sc]);
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int ky = sr * UP - r * DOWN;
                if (ky < 0 || ky >= F) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
#pragma unroll
                    for (int sc = 0; sc < Cfg::NIN; ++sc) {
                        const int kx = sc * UP - e * DOWN;
                        if (kx < 0 || kx >= F) continue;
                        acc[r][e] = fmaf(v[sc], w[ky * F + kx], acc[r][e]);
                    }
                }
            }
        }
        {
            const long long img = t / tiles_per_img;
            const int rem = (int)(t - img * tiles_per_img);
            const int tyi = rem / p.tilesX, txi = rem - tyi * p.tilesX;
            const int n = (int)(img / p.C), c = (int)(img - (long long)n * p.C);
            const int ox = txi * Cfg::TW + x0, oy0 = tyi * Cfg::TH + y0;
            T* yi = yp + n * p.ys[0] + c * p.ys[1] + ox;
            if (ox < p.outW) {
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const int oy = oy0 + r;
                    if (oy >= p.outH) break;
                    T* dst = yi + (long long)oy * p.ys[2];
                    if (ox + 1 < p.outW) {
                        if (pair_store) {
                            struct alignas(2 * sizeof(T)) Pair { T a, b; };
                            *reinterpret_cast<Pair*>(dst) = Pair{Px<T>::st(acc[r][0]), Px<T>::st(acc[r][1])};
                        } else {
                            dst[0] = Px<T>::st(acc[r][0]);
                            dst[1] = Px<T>::st(acc[r][1]);
                        }
                    } else {
                        dst[0] = Px<T>::st(acc[r][0]);
                    }
                }
            }
        }
        __syncthreads();                                                     // buffer b is free for the prefetch after next
    }
    cp_async_wait<0>();
}

template <typename T, int UP, int DOWN, int F>
int launch_fast(UpfirdnParams p, int n_sm, cudaStream_t stream) {
    using Cfg = FastCfg<T, UP, DOWN, F>;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        P3D_CUDA_TRY(cudaFuncSetAttribute(k_upfirdn2d_fast<T, UP, DOWN, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
        P3D_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_upfirdn2d_fast<T, UP, DOWN, F>, kThreads, Cfg::kSmem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    p.tilesX = (p.outW + Cfg::TW - 1) / Cfg::TW;
    p.tilesY = (p.outH + Cfg::TH - 1) / Cfg::TH;
    const long long total = (long long)p.N * p.C * p.tilesX * p.tilesY;
    const long long cap = (long long)n_sm * ctas_per_sm;
    const int grid = (int)(total < cap ? total : cap);
    const uintptr_t pa = 2 * sizeof(T) - 1;
    const int pair_store = ((reinterpret_cast<uintptr_t>(p.y) & pa) == 0 && p.ys[0] % 2 == 0 && p.ys[1] % 2 == 0 && p.ys[2] % 2 == 0) ? 1 : 0;
    k_upfirdn2d_fast<T, UP, DOWN, F><<<grid, kThreads, Cfg::kSmem, stream>>>(p, total, pair_store);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

// returns 1 when the shape is not covered by the fast path (0 / negative: launched / error)
template <typename T>
int try_launch_fast(const UpfirdnParams& p, int n_sm, cudaStream_t stream) {
    if (sizeof(T) > 4) return 1;
    if (p.xs[3] != 1 || p.ys[3] != 1 || p.fH != 4 || p.fW != 4 || p.upx != p.upy || p.downx != p.downy) return 1;
    if (p.padx0 % p.upx != 0 || p.pady0 % p.upy != 0) return 1;
    if ((long long)p.outW * p.downx + 8 > 0x3fffffffll || (long long)p.outH * p.downy + 8 > 0x3fffffffll) return 1;
    if (p.upx == 1 && p.downx == 1) return launch_fast<T, 1, 1, 4>(p, n_sm, stream);
    if (p.upx == 2 && p.downx == 1) return launch_fast<T, 2, 1, 4>(p, n_sm, stream);
    if (p.upx == 1 && p.downx == 2) return launch_fast<T, 1, 2, 4>(p, n_sm, stream);
    return 1;
}
template <> int try_launch_fast<double>(const UpfirdnParams&, int, cudaStream_t) { return 1; }

// ------------------------------------------------------------------------------------------ strided (any layout)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_upfirdn2d_strided(const UpfirdnParams p, int c_fastest) {
    using S = typename Px<T>::acc;
    __shared__ float s_w[kMaxTaps];
    const int taps = p.fH * p.fW;
    for (int i = threadIdx.x; i < taps; i += kThreads) {
        const int ky = i / p.fW, kx = i - ky * p.fW;
        const int sy = p.flip ? ky : p.fH - 1 - ky, sx = p.flip ? kx : p.fW - 1 - kx;
        s_w[i] = p.f[sy * p.fsh + sx * p.fsw] * p.gain;
    }
    __syncthreads();
    const T* xp = reinterpret_cast<const T*>(p.x);
    T* yp = reinterpret_cast<T*>(p.y);
    const long long total = (long long)p.N * p.C * p.outH * p.outW;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        int n, c, oy, ox;
        long long r = idx;
        if (c_fastest) { c = (int)(r % p.C); r /= p.C; ox = (int)(r % p.outW); r /= p.outW; oy = (int)(r % p.outH); n = (int)(r / p.outH); }
        else { ox = (int)(r % p.outW); r /= p.outW; oy = (int)(r % p.outH); r /= p.outH; c = (int)(r % p.C); n = (int)(r / p.C); }
        const int ux0 = ox * p.downx - p.padx0, uy0 = oy * p.downy - p.pady0;
        const int kx0 = pos_mod(-ux0, p.upx), ky0 = pos_mod(-uy0, p.upy);
        const T* xi = xp + n * p.xs[0] + c * p.xs[1];
        S acc = (S)0;
        for (int ky = ky0; ky < p.fH; ky += p.upy) {
            const int iy = (uy0 + ky) / p.upy;
            if (iy < 0 || iy >= p.inH) continue;
            for (int kx = kx0; kx < p.fW; kx += p.upx) {
                const int ix = (ux0 + kx) / p.upx;
                if (ix < 0 || ix >= p.inW) continue;
                acc += Px<T>::ld(xi[iy * p.xs[2] + ix * p.xs[3]]) * (S)s_w[ky * p.fW + kx];
            }
        }
        yp[n * p.ys[0] + c * p.ys[1] + oy * p.ys[2] + ox * p.ys[3]] = Px<T>::st(acc);
    }
}

template <typename T>
int launch_upfirdn2d(UpfirdnParams p, cudaStream_t stream) {
    using S = typename Px<T>::acc;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        P3D_CUDA_TRY(cudaGetDevice(&dev));
        P3D_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    {
        const int rc = try_launch_fast<T>(p, n_sm, stream);
        if (rc <= 0) return rc;
    }
    // footprint of a 64x16 output tile in input pixels (+1 for the floor of the first index)
    p.inTileW = (kTileW * p.downx + p.fW - 1 + p.upx - 1) / p.upx + 1;
    p.inTileH = (kTileH * p.downy + p.fH - 1 + p.upy - 1) / p.upy + 1;
    const size_t smem = ((size_t)p.fH * p.fW * 4 + 15 & ~(size_t)15) + (size_t)p.inTileW * p.inTileH * sizeof(S);
    const bool tiled_ok = p.xs[3] == 1 && smem <= 96 * 1024;
    if (tiled_ok) {
        p.tilesX = (p.outW + kTileW - 1) / kTileW;
        p.tilesY = (p.outH + kTileH - 1) / kTileH;
        const long long nc = (long long)p.N * p.C;
        dim3 grid((unsigned)(p.tilesX * p.tilesY), (unsigned)(nc < 32768 ? nc : 32768));
        if (smem > 48 * 1024) P3D_CUDA_TRY(cudaFuncSetAttribute(k_upfirdn2d_tiled<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_upfirdn2d_tiled<T><<<grid, kThreads, smem, stream>>>(p);
        P3D_LAUNCH_CHECK();
        return P3D_OK;
    }
    P3D_REQUIRE(p.fH * p.fW <= kMaxTaps, "filter too large for the strided kernel (%dx%d)", p.fH, p.fW);
    const long long total = (long long)p.N * p.C * p.outH * p.outW;
    const long long blocks = (total + kThreads - 1) / kThreads;
    const int grid = (int)(blocks < (long long)n_sm * 16 ? blocks : (long long)n_sm * 16);
    k_upfirdn2d_strided<T><<<grid, kThreads, 0, stream>>>(p, p.xs[1] == 1 ? 1 : 0);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

}  // namespace
}  // namespace p3d

using namespace p3d;

extern "C" int p3d_upfirdn2d(const void* x, const float* f, void* y, int32_t dtype, int32_t n, int32_t c, int32_t in_h,
                             int32_t in_w, const int64_t* x_stride, int32_t f_h, int32_t f_w, int64_t f_stride_h,
                             int64_t f_stride_w, int32_t out_h, int32_t out_w, const int64_t* y_stride, int32_t upx,
                             int32_t upy, int32_t downx, int32_t downy, int32_t padx0, int32_t padx1, int32_t pady0,
                             int32_t pady1, int32_t flip, float gain, void* stream) {
    P3D_REQUIRE(x && f && y && x_stride && y_stride, "null pointer");
    P3D_REQUIRE(n > 0 && c > 0 && in_h > 0 && in_w > 0, "x has zero size");
    P3D_REQUIRE(f_h >= 1 && f_w >= 1, "f must be at least 1x1");
    P3D_REQUIRE(upx >= 1 && upy >= 1, "upsampling factor must be at least 1");
    P3D_REQUIRE(downx >= 1 && downy >= 1, "downsampling factor must be at least 1");
    const int ow = (in_w * upx + padx0 + padx1 - f_w + downx) / downx;
    const int oh = (in_h * upy + pady0 + pady1 - f_h + downy) / downy;
    P3D_REQUIRE(ow >= 1 && oh >= 1, "output must be at least 1x1");
    P3D_REQUIRE(ow == out_w && oh == out_h, "output size mismatch: expected %dx%d, got %dx%d", oh, ow, out_h, out_w);
    UpfirdnParams p{};
    p.x = x; p.f = f; p.y = y; p.N = n; p.C = c; p.inH = in_h; p.inW = in_w; p.outH = out_h; p.outW = out_w; p.fH = f_h; p.fW = f_w;
    for (int i = 0; i < 4; ++i) { p.xs[i] = x_stride[i]; p.ys[i] = y_stride[i]; }
    p.fsh = f_stride_h; p.fsw = f_stride_w;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip ? 1 : 0; p.gain = gain;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case P3D_F32: return launch_upfirdn2d<float>(p, st);
        case P3D_F16: return launch_upfirdn2d<__half>(p, st);
        case P3D_BF16: return launch_upfirdn2d<__nv_bfloat16>(p, st);
        case P3D_F64: return launch_upfirdn2d<double>(p, st);
    }
    set_error("unsupported dtype %d", dtype);
    return P3D_EINVAL;
}
