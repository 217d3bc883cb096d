// v1 renderer: multi-kernel SIMT pipeline, fp32 throughout.
//
// This is the parity kernel set (P3D_MLP_FP32_SIMT): every arithmetic step is plain fp32 in
// the same order class as the reference's ATen ops, so it tracks the CPU oracle to ~1e-6.
// Kernels (R = N*M rays):
//   k_ray_limits      'auto' mode only: ray/AABB slab test            math_utils.py:46-98
//   k_sample_decode   stratified depth -> xyz -> tri-plane gather -> OSGDecoder -> masks
//                     one CTA = 128 consecutive samples; gather is cooperative (8 lanes x 16 B
//                     = one 128 B texel per tap), MLP is thread-per-sample FFMA from smem weights
//   k_ray_importance  warp per ray: coarse weights -> smoothed pdf -> inverse CDF -> sorted fine depths
//   k_ray_composite   warp per ray: merge (coarse|fine) by rank, alpha/T scan, weighted colour sum
//   k_depth_finalize  nan_to_num(inf) + clamp to the batch-global [min,max] depth  ray_marcher.py:49-50
#include "render_internal.cuh"

namespace p3d {

namespace {

constexpr int kTile = 128;          // samples per CTA in k_sample_decode
constexpr int kFeatStride = 33;     // padded feature row (bank-conflict free for thread-per-row reads)
constexpr int kW2Stride = 36;       // W2^T row: 33 outputs padded to 36 floats (float4 broadcast reads)

struct SampleArgs {
    Geom g;
    const void* planes;
    const float *w1, *b1, *w2, *b2;
    const float *ro, *rd;       // (R,3)
    const float* u;             // coarse jitter (R*S) or nullptr
    const float* depth_in;      // fine pass: depths (R*Sf)
    const float* coords;        // points mode: (N,K,3)
    const float *ray_t0, *ray_t1;
    const unsigned int* bounds;
    float* depth_out;           // coarse pass: (R*S)
    float* sigma_out;           // (total)
    float* rgb_out;             // (total,32)
    long long total;            // samples in this launch
    long long per_view;         // samples per view (M*S, M*Sf or K)
    int per_ray;                // S, Sf (or 1 in points mode)
};

enum { MODE_COARSE = 0, MODE_FINE = 1, MODE_POINTS = 2 };

// -------------------------------------------------------------------------------------------
// depth of coarse sample s of a ray.  renderer.py:303-326 (+ math_utils.linspace :101-118)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float coarse_depth(const Geom& g, int s, float u, float t0, float t1) {
    const int S = g.S;
    if (g.ray_mode == P3D_RAYS_AUTOBOX) {
        float step = __fdiv_rn((float)s, (float)(S - 1));
        float span = __fsub_rn(t1, t0);
        float base = __fadd_rn(t0, __fmul_rn(step, span));
        float delta = __fdiv_rn(span, (float)(S - 1));
        return __fadd_rn(base, __fmul_rn(u, delta));
    }
    if (g.disparity) {
        float st = __fdiv_rn(1.0f, (float)(S - 1));
        float lin = (s < S / 2) ? __fmul_rn(st, (float)s) : __fsub_rn(1.0f, __fmul_rn(st, (float)(S - 1 - s)));
        float t = __fadd_rn(lin, __fmul_rn(u, g.disp_delta));
        float den = __fadd_rn(__fmul_rn(g.inv_start, __fsub_rn(1.0f, t)), __fmul_rn(g.inv_end, t));
        return __fdiv_rn(1.0f, den);
    }
    float lin = (s < S / 2) ? __fadd_rn(g.ray_start, __fmul_rn(g.lin_step, (float)s))
                            : __fsub_rn(g.ray_end, __fmul_rn(g.lin_step, (float)(S - 1 - s)));
    return __fadd_rn(lin, __fmul_rn(u, g.depth_delta));
}

// -------------------------------------------------------------------------------------------
// Tri-plane bilinear gather for one sample, channels [4q, 4q+4).  renderer.py:52-81:
//   grid = coord * (2/box_warp); texel = ((grid+1)*size - 1)/2  (align_corners=False);
//   taps outside the plane contribute zero (padding_mode='zeros'); mean over the 3 planes.
// Every tap is one contiguous 128 B (fp32) / 64 B (bf16) texel: the 8 lanes of a sample read
// 16 B (8 B) each, so a warp-wide load touches 4 lines.
// -------------------------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ float4 load_quad(const void* planes, long long elem_off) {
    if (BF16) {
        const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(planes) + elem_off));
        float4 r;
        r.x = __uint_as_float(raw.x << 16);
        r.y = __uint_as_float(raw.x & 0xffff0000u);
        r.z = __uint_as_float(raw.y << 16);
        r.w = __uint_as_float(raw.y & 0xffff0000u);
        return r;
    } else {
        return ldg128(reinterpret_cast<const float*>(planes) + elem_off);
    }
}

struct Tap4 {
    long long off[4];
    float w[4];
};

__device__ __forceinline__ Tap4 plane_taps(const Geom& g, long long base, float ca, float cb) {
    Tap4 t;
    const float gx = __fmul_rn(ca, g.coord_scale), gy = __fmul_rn(cb, g.coord_scale);
    const float fx = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)g.W), 1.f), 0.5f);
    const float fy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)g.H), 1.f), 0.5f);
    const bool sane = (fx > -2.f) && (fx < (float)g.W + 1.f) && (fy > -2.f) && (fy < (float)g.H + 1.f);  // false for NaN too
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float wx1 = fx - x0f, wy1 = fy - y0f;
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = sane ? (int)x0f : -4, y0 = sane ? (int)y0f : -4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xi = x0 + (k & 1), yi = y0 + (k >> 1);
        const bool ok = (xi >= 0) && (xi < g.W) && (yi >= 0) && (yi < g.H);
        const float wgt = ((k & 1) ? wx1 : wx0) * ((k >> 1) ? wy1 : wy0);
        t.w[k] = ok ? wgt : 0.f;
        t.off[k] = base + (ok ? (long long)yi * g.stride_row + (long long)xi * g.stride_col : 0ll);
    }
    return t;
}

template <bool BF16>
__device__ __forceinline__ float4 gather_features(const void* planes, const Geom& g, int view, float x, float y, float z, int q) {
    const long long vbase = (long long)view * g.stride_view + 4 * q;
    const float a2 = g.plane_mode == P3D_PLANES_PANIC3D ? y : z;
    const float b2 = g.plane_mode == P3D_PLANES_PANIC3D ? z : x;
    Tap4 t0 = plane_taps(g, vbase, x, y);
    Tap4 t1 = plane_taps(g, vbase + g.stride_plane, x, z);
    Tap4 t2 = plane_taps(g, vbase + 2 * g.stride_plane, a2, b2);
    float4 v[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = load_quad<BF16>(planes, t0.off[k]);
        v[4 + k] = load_quad<BF16>(planes, t1.off[k]);
        v[8 + k] = load_quad<BF16>(planes, t2.off[k]);
    }
    float4 f[3];
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const Tap4& t = p == 0 ? t0 : (p == 1 ? t1 : t2);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a.x = fmaf(v[4 * p + k].x, t.w[k], a.x);
            a.y = fmaf(v[4 * p + k].y, t.w[k], a.y);
            a.z = fmaf(v[4 * p + k].z, t.w[k], a.z);
            a.w = fmaf(v[4 * p + k].w, t.w[k], a.w);
        }
        f[p] = a;
    }
    float4 r;
    r.x = __fdiv_rn(__fadd_rn(__fadd_rn(f[0].x, f[1].x), f[2].x), 3.f);
    r.y = __fdiv_rn(__fadd_rn(__fadd_rn(f[0].y, f[1].y), f[2].y), 3.f);
    r.z = __fdiv_rn(__fadd_rn(__fadd_rn(f[0].z, f[1].z), f[2].z), 3.f);
    r.w = __fdiv_rn(__fadd_rn(__fadd_rn(f[0].w, f[1].w), f[2].w), 3.f);
    return r;
}

// crop + cull/binarize on a raw density.  renderer.py:138-153, :187-198
__device__ __forceinline__ float apply_masks(const Geom& g, float sigma, float x, float z) {
    if (g.crop_on && !((fabsf(x) <= g.crop_limit) && (fabsf(z) <= g.crop_limit))) sigma = -1e3f;
    if (g.binarize_on || g.cull_on) {
        const float alpha = 1.f - expf(-softplus_t(sigma - 1.f));
        const bool m = alpha < g.cull_thresh;
        if (g.binarize_on) sigma = m ? -1e3f : 1e3f;
        else if (m) sigma = -1e3f;
    }
    return sigma;
}

// -------------------------------------------------------------------------------------------
// k_sample_decode
// -------------------------------------------------------------------------------------------
template <int MODE, bool BF16>
__global__ void __launch_bounds__(kTile) k_sample_decode(const SampleArgs a) {
    __shared__ __align__(16) float s_w1[kHidden * kC];          // [j][k], gain folded in
    __shared__ __align__(16) float s_w2t[kHidden * kW2Stride];  // [j][m], gain folded in
    __shared__ float s_b1[kHidden];
    __shared__ float s_b2[kW2Stride];
    __shared__ float s_feat[kTile * kFeatStride];               // features in, colours out
    __shared__ float s_xz[kTile * 2];                           // raw x,z for the crop mask

    const Geom& g = a.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long tile0 = (long long)blockIdx.x * kTile;

    for (int i = tid; i < kHidden * kC; i += kTile) s_w1[i] = __fmul_rn(a.w1[i], g.w1_gain);
    for (int i = tid; i < kHidden * kOut; i += kTile) {
        const int m = i / kHidden, j = i % kHidden;             // w2 is (out, hidden)
        s_w2t[j * kW2Stride + m] = __fmul_rn(a.w2[i], g.w2_gain);
    }
    if (tid < kHidden) s_b1[tid] = __fmul_rn(a.b1[tid], g.b1_gain);
    if (tid < kOut) s_b2[tid] = __fmul_rn(a.b2[tid], g.b2_gain);

    // ---- phase A: positions + gather; 4 samples per warp per round, 8 lanes per sample
    const int sub = lane >> 3, q = lane & 7;
#pragma unroll 1
    for (int round = 0; round < 8; ++round) {
        const int local = warp * 32 + round * 4 + sub;
        const long long gidx = tile0 + local;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        float px = 0.f, pz = 0.f;
        if (gidx < a.total) {
            const int view = (int)(gidx / a.per_view);
            float py;
            if (MODE == MODE_POINTS) {
                px = a.coords[gidx * 3 + 0]; py = a.coords[gidx * 3 + 1]; pz = a.coords[gidx * 3 + 2];
            } else {
                const long long ray = gidx / a.per_ray;
                const int s = (int)(gidx - ray * a.per_ray);
                float t;
                if (MODE == MODE_COARSE) {
                    const float u = a.u ? a.u[gidx] : philox_uniform(g.seed, (uint64_t)gidx, 0u);
                    float t0 = 0.f, t1 = 0.f;
                    if (g.ray_mode == P3D_RAYS_AUTOBOX) {
                        t0 = a.ray_t0[ray]; t1 = a.ray_t1[ray];
                        if (!(t1 > t0) && a.bounds[4]) {           // renderer.py:167-170
                            t0 = ordered_to_float(a.bounds[2]);
                            t1 = ordered_to_float(a.bounds[3]);
                        }
                    }
                    t = coarse_depth(g, s, u, t0, t1);
                    if (q == 0) a.depth_out[gidx] = t;
                } else {
                    t = a.depth_in[gidx];
                }
                const float* o = a.ro + ray * 3;
                const float* d = a.rd + ray * 3;
                px = __fadd_rn(o[0], __fmul_rn(t, d[0]));
                py = __fadd_rn(o[1], __fmul_rn(t, d[1]));
                pz = __fadd_rn(o[2], __fmul_rn(t, d[2]));
            }
            f = gather_features<BF16>(a.planes, g, view, px, py, pz, q);
        }
        float* dst = s_feat + local * kFeatStride + 4 * q;
        dst[0] = f.x; dst[1] = f.y; dst[2] = f.z; dst[3] = f.w;
        if (q == 0) { s_xz[local * 2] = px; s_xz[local * 2 + 1] = pz; }
    }
    __syncthreads();

    // ---- phase B: OSGDecoder, one sample per thread.  triplane.py:528-544
    float x[kC];
#pragma unroll
    for (int k = 0; k < kC; ++k) x[k] = s_feat[tid * kFeatStride + k];
    float o[kW2Stride];
#pragma unroll
    for (int m = 0; m < kW2Stride; ++m) o[m] = m < kOut ? s_b2[m] : 0.f;
#pragma unroll 2
    for (int j = 0; j < kHidden; ++j) {
        const float4* wrow = reinterpret_cast<const float4*>(s_w1 + j * kC);
        float h = s_b1[j];
#pragma unroll
        for (int k4 = 0; k4 < kC / 4; ++k4) {
            const float4 w = wrow[k4];
            h = fmaf(x[4 * k4 + 0], w.x, h);
            h = fmaf(x[4 * k4 + 1], w.y, h);
            h = fmaf(x[4 * k4 + 2], w.z, h);
            h = fmaf(x[4 * k4 + 3], w.w, h);
        }
        h = softplus_t(h);
        const float4* w2row = reinterpret_cast<const float4*>(s_w2t + j * kW2Stride);
#pragma unroll
        for (int m4 = 0; m4 < kW2Stride / 4; ++m4) {
            const float4 w = w2row[m4];
            o[4 * m4 + 0] = fmaf(h, w.x, o[4 * m4 + 0]);
            o[4 * m4 + 1] = fmaf(h, w.y, o[4 * m4 + 1]);
            o[4 * m4 + 2] = fmaf(h, w.z, o[4 * m4 + 2]);
            o[4 * m4 + 3] = fmaf(h, w.w, o[4 * m4 + 3]);
        }
    }
    const long long gidx = tile0 + tid;
    float sigma = o[0];
    if (MODE != MODE_POINTS) sigma = apply_masks(g, sigma, s_xz[tid * 2], s_xz[tid * 2 + 1]);
    if (gidx < a.total) a.sigma_out[gidx] = sigma;
    __syncthreads();                       // everyone has consumed its feature row
#pragma unroll
    for (int c = 0; c < kRgb; ++c) {
        float v = sigmoid_t(o[1 + c]);
        if (!g.force_sigmoid) v = __fsub_rn(__fmul_rn(v, 1.002f), 0.001f);
        s_feat[tid * kFeatStride + c] = v;
    }
    __syncthreads();
    // coalesced colour store: 128 rows x 32 floats
    for (int i = tid; i < kTile * kRgb; i += kTile) {
        const int row = i >> 5, col = i & 31;
        const long long gi = tile0 + row;
        if (gi < a.total) a.rgb_out[gi * kRgb + col] = s_feat[row * kFeatStride + col];
    }
}

// -------------------------------------------------------------------------------------------
// k_ray_limits: math_utils.get_ray_limits_box (:46-98) + validity reduction (renderer.py:165-170)
// -------------------------------------------------------------------------------------------
__global__ void k_ray_limits(const float* ro, const float* rd, long long R, float h, float* t0o, float* t1o,
                             unsigned int* bounds) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    float tmin = 0.f, tmax = 0.f;
    bool valid = true;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float o = ro[r * 3 + ax], inv = __fdiv_rn(1.f, rd[r * 3 + ax]);
        const bool neg = inv < 0.f;
        const float lo = __fmul_rn(__fsub_rn(neg ? h : -h, o), inv);
        const float hi = __fmul_rn(__fsub_rn(neg ? -h : h, o), inv);
        if (ax == 0) { tmin = lo; tmax = hi; }
        else {
            if (tmin > hi || lo > tmax) valid = false;
            tmin = fmaxf(tmin, lo);       // torch.max propagates NaN, fmaxf does not: NaN only for 0*inf rays
            tmax = fminf(tmax, hi);
        }
    }
    if (!valid) { tmin = -1.f; tmax = -2.f; }
    t0o[r] = tmin; t1o[r] = tmax;
    if (tmax > tmin) {
        atomicMin(&bounds[2], float_to_ordered(tmin));
        atomicMax(&bounds[3], float_to_ordered(tmin));
        atomicOr(&bounds[4], 1u);
    }
}

__global__ void k_init_bounds(unsigned int* bounds) {
    if (threadIdx.x == 0) {
        bounds[0] = 0xffffffffu; bounds[1] = 0u; bounds[2] = 0xffffffffu; bounds[3] = 0u; bounds[4] = 0u;
    }
}

// -------------------------------------------------------------------------------------------
// per-ray compositing weights.  ray_marcher.py:25-44
//   alpha_i = 1 - exp(-softplus((s_i+s_{i+1})/2 - 1) * (t_{i+1}-t_i)),
//   w_i = alpha_i * prod_{k<i} (1 - alpha_k + 1e-10)
// t, sg: L sorted samples in shared memory; writes w[0..L-2]; returns (sum w, sum w*t_mid) on all lanes.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_weights(const float* t, const float* sg, float* w, int L, int lane,
                                            float& wsum, float& dnum) {
    float carry = 1.f, acc_w = 0.f, acc_d = 0.f;
    for (int base = 0; base < L - 1; base += 32) {
        const int i = base + lane;
        float alpha = 0.f, factor = 1.f, tmid = 0.f;
        if (i < L - 1) {
            const float delta = t[i + 1] - t[i];
            const float dens = softplus_t(__fsub_rn(__fmul_rn(__fadd_rn(sg[i], sg[i + 1]), 0.5f), 1.f));
            alpha = 1.f - expf(-__fmul_rn(dens, delta));
            factor = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
            tmid = __fmul_rn(__fadd_rn(t[i], t[i + 1]), 0.5f);
        }
        const float incl = warp_scan_mul(factor, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.f;
        const float T = carry * excl;
        const float wi = alpha * T;
        if (i < L - 1) { w[i] = wi; acc_w += wi; acc_d = fmaf(wi, tmid, acc_d); }
        carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    wsum = warp_sum(acc_w);
    dnum = warp_sum(acc_d);
}

// -------------------------------------------------------------------------------------------
// k_ray_importance: sample_importance + sample_pdf, renderer.py:328-387.  One warp per ray.
// -------------------------------------------------------------------------------------------
struct ImportanceArgs {
    Geom g;
    const float *depth_c, *sigma_c, *u_fine;
    float* depth_f;
    unsigned int* bounds;
    long long R;
    int sort_pow2;     // next power of two >= Sf
};

__global__ void __launch_bounds__(128) k_ray_importance(const ImportanceArgs a) {
    extern __shared__ float smem[];
    const Geom& g = a.g;
    const int S = g.S, Sf = g.Sf, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ray = (long long)blockIdx.x * 4 + warp;
    if (ray >= a.R) return;
    const int per_warp = 4 * S + a.sort_pow2;
    float* t = smem + warp * per_warp;      // S
    float* sg = t + S;                      // S
    float* w = sg + S;                      // S   (weights, S-1 used; later pooled pdf)
    float* cdf = w + S;                     // S   (S-2 used)
    float* fine = cdf + S;                  // sort_pow2

    for (int i = lane; i < S; i += 32) { t[i] = a.depth_c[ray * S + i]; sg[i] = a.sigma_c[ray * S + i]; }
    __syncwarp();
    float wsum, dnum;
    ray_weights(t, sg, w, S, lane, wsum, dnum);
    __syncwarp();
    // max_pool1d(k=2,s=1,pad=1) -> avg_pool1d(k=2,s=1) -> +0.01 ; keep [1:-1] ; +eps     (:336-343, :360)
    // pooled[i] = (max(w[i-1],w[i]) + max(w[i],w[i+1]))/2 for i in [0,S-2], -inf outside [0,S-2]
    const int nb = S - 3;                   // pdf bins
    float my[8];                            // up to 256 coarse samples per ray
    float part = 0.f;
    int cnt = 0;
    for (int k = lane; k < nb; k += 32, ++cnt) {
        const int i = k + 1;
        const float wl = w[i - 1], wc = w[i], wr = (i + 1 <= S - 2) ? w[i + 1] : -INFINITY;
        const float pooled = __fmul_rn(__fadd_rn(fmaxf(wl, wc), fmaxf(wc, wr)), 0.5f);
        const float v = __fadd_rn(__fadd_rn(pooled, 0.01f), 1e-5f);
        my[cnt] = v;
        part += v;
    }
    const float total = warp_sum(part);
    // cdf[0] = 0, cdf[k+1] = cumsum(pdf)[k]
    float carry = 0.f;
    cnt = 0;
    for (int base = 0; base < nb; base += 32, ++cnt) {
        const int k = base + lane;
        const float pdf = k < nb ? __fdiv_rn(my[cnt], total) : 0.f;
        const float incl = warp_scan_add(pdf, lane) + carry;
        if (k < nb) cdf[k + 1] = incl;
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) cdf[0] = 0.f;
    __syncwarp();
    // inverse CDF.  searchsorted(right=True): first index with cdf[idx] > u
    for (int f = lane; f < a.sort_pow2; f += 32) {
        float val = INFINITY;
        if (f < Sf) {
            const float u = a.u_fine ? a.u_fine[ray * Sf + f] : philox_uniform(g.seed, (uint64_t)(ray * Sf + f), 1u);
            int lo = 0, hi = nb + 1;            // cdf has nb+1 entries
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
            }
            const int below = max(lo - 1, 0), above = min(lo, nb);
            const float c0 = cdf[below], c1 = cdf[above];
            const float b0 = __fmul_rn(0.5f, __fadd_rn(t[below], t[below + 1]));
            const float b1 = __fmul_rn(0.5f, __fadd_rn(t[above], t[above + 1]));
            float den = __fsub_rn(c1, c0);
            if (den < 1e-5f) den = 1.f;
            val = __fadd_rn(b0, __fmul_rn(__fdiv_rn(__fsub_rn(u, c0), den), __fsub_rn(b1, b0)));
        }
        fine[f] = val;
    }
    __syncwarp();
    // bitonic sort ascending (per ray); the merged composite is order independent, sorting here
    // lets k_ray_composite merge by rank instead of running a global sort (renderer.py:289-301)
    for (int k = 2; k <= a.sort_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < a.sort_pow2; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float x = fine[i], y = fine[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { fine[i] = y; fine[ixj] = x; }
                }
            }
            __syncwarp();
        }
    }
    for (int f = lane; f < Sf; f += 32) a.depth_f[ray * Sf + f] = fine[f];
}

// -------------------------------------------------------------------------------------------
// k_ray_composite: unify_samples (renderer.py:289-301) + final MipRayMarcher2 (ray_marcher.py:25-57)
// on the 32 colour channels and, analytically, on xyz (the marcher is linear in the colours and
// xyz_j = o + t_j d, so  sum_i w_i xyz_mid_i = o*sum(w) + d*sum(w_i t_mid_i)).
// -------------------------------------------------------------------------------------------
struct CompositeArgs {
    Geom g;
    const float *depth_c, *sigma_c, *rgb_c, *depth_f, *sigma_f, *rgb_f, *ro, *rd;
    float *out_rgb, *out_depth, *out_wsum, *out_xyz;
    unsigned int* bounds;
    long long R;
};

__global__ void __launch_bounds__(128) k_ray_composite(const CompositeArgs a) {
    extern __shared__ float smem[];
    const Geom& g = a.g;
    const int S = g.S, Sf = g.Sf, L = S + Sf, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ray = (long long)blockIdx.x * 4 + warp;
    if (ray >= a.R) return;
    float* t = smem + warp * (4 * L);       // merged depths
    float* sg = t + L;                      // merged densities
    float* w = sg + L;                      // interval weights (L-1), then per-sample omegas
    int* src = reinterpret_cast<int*>(w + L);   // merged position -> colour row

    const float* tc = a.depth_c + ray * S;
    const float* tf = a.depth_f + ray * Sf;
    // merge by rank; ties keep coarse before fine (stable sort of cat([coarse, fine])).
    // Coarse depths ascend except for the degenerate 'auto' case where no ray hits the box
    // (t0=-1 > t1=-2, renderer.py:166-170 leaves them untouched): then they descend -> read reversed.
    const bool rev = tc[0] > tc[S - 1];
    for (int i = lane; i < S; i += 32) {
        const int ci = rev ? S - 1 - i : i;
        const float v = tc[ci];
        int lo = 0, hi = Sf;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (tf[mid] < v) lo = mid + 1; else hi = mid; }
        const int pos = i + lo;
        t[pos] = v; sg[pos] = a.sigma_c[ray * S + ci]; src[pos] = ci;
    }
    for (int j = lane; j < Sf; j += 32) {
        const float v = tf[j];
        int lo = 0, hi = S;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (tc[rev ? S - 1 - mid : mid] <= v) lo = mid + 1; else hi = mid; }
        const int pos = j + lo;
        t[pos] = v; sg[pos] = a.sigma_f[ray * Sf + j]; src[pos] = S + j;
    }
    __syncwarp();
    float wsum, dnum;
    ray_weights(t, sg, w, L, lane, wsum, dnum);
    __syncwarp();
    // omega_j = (w_{j-1} + w_j)/2 : sum_i w_i (c_i + c_{i+1})/2 == sum_j omega_j c_j
    float om[16];                           // up to 512 merged samples per ray
    int cnt = 0;
    for (int j = lane; j < L; j += 32, ++cnt) {
        const float wl = j > 0 ? w[j - 1] : 0.f, wr = j < L - 1 ? w[j] : 0.f;
        om[cnt] = __fmul_rn(__fadd_rn(wl, wr), 0.5f);
    }
    __syncwarp();
    cnt = 0;
    for (int j = lane; j < L; j += 32, ++cnt) w[j] = om[cnt];
    __syncwarp();
    float acc = 0.f;                        // lane = colour channel
    for (int j = 0; j < L; ++j) {
        const int s = src[j];
        const float c = s < S ? a.rgb_c[(ray * S + s) * kRgb + lane] : a.rgb_f[(ray * Sf + (s - S)) * kRgb + lane];
        acc = fmaf(w[j], c, acc);
    }
    const float back = g.white_back ? __fsub_rn(1.f, wsum) : 0.f;
    a.out_rgb[ray * kRgb + lane] = __fsub_rn(__fmul_rn(__fadd_rn(acc, back), 2.f), 1.f);
    if (lane < 3) {
        const float v = fmaf(a.ro[ray * 3 + lane], wsum, a.rd[ray * 3 + lane] * dnum);
        a.out_xyz[ray * 3 + lane] = __fsub_rn(__fmul_rn(__fadd_rn(v, back), 2.f), 1.f);
    }
    if (lane == 0) {
        a.out_depth[ray] = __fdiv_rn(dnum, wsum);      // 0/0 -> NaN, fixed up by k_depth_finalize
        a.out_wsum[ray] = wsum;
        atomicMin(&a.bounds[0], float_to_ordered(t[0]));
        atomicMax(&a.bounds[1], float_to_ordered(t[L - 1]));
    }
}

// nan_to_num(nan=inf) then clamp(min(depths), max(depths)) over the whole batch.  ray_marcher.py:49-50
__global__ void k_depth_finalize(float* depth, long long R, const unsigned int* bounds) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float lo = ordered_to_float(bounds[0]), hi = ordered_to_float(bounds[1]);
    float d = depth[r];
    if (isnan(d)) d = INFINITY;
    depth[r] = fminf(fmaxf(d, lo), hi);
}

template <int MODE>
int launch_sample_decode(const SampleArgs& a, bool bf16, cudaStream_t stream) {
    if (a.total <= 0) return P3D_OK;
    const long long blocks = (a.total + kTile - 1) / kTile;
    P3D_REQUIRE(blocks < (1ll << 31), "too many samples for one launch (%lld)", a.total);
    ProfileScope prof(PROF_SAMPLE_DECODE, stream);
    if (bf16) k_sample_decode<MODE, true><<<(unsigned)blocks, kTile, 0, stream>>>(a);
    else k_sample_decode<MODE, false><<<(unsigned)blocks, kTile, 0, stream>>>(a);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------
int render_forward_v1(const Geom& g, const p3d_render_params* p, const void* planes, const float* w1, const float* b1,
                      const float* w2, const float* b2, const float* ro, const float* rd, const float* u_c,
                      const float* u_f, const Workspace& ws, float* out_rgb, float* out_depth, float* out_wsum,
                      float* out_xyz, cudaStream_t stream) {
    const long long R = (long long)g.N * g.M;
    P3D_REQUIRE(g.S >= 2 && g.S <= 256, "depth_resolution must be in [2,256], got %d", g.S);
    P3D_REQUIRE(g.Sf == 0 || g.S >= 4, "importance sampling needs depth_resolution >= 4");
    P3D_REQUIRE(g.Sf >= 0 && g.S + g.Sf <= 512, "depth_resolution + depth_resolution_importance must be <= 512");

    k_init_bounds<<<1, 32, 0, stream>>>(ws.bounds);
    P3D_LAUNCH_CHECK();
    if (g.ray_mode == P3D_RAYS_AUTOBOX) {
        k_ray_limits<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(ro, rd, R, g.half_box, ws.ray_t0, ws.ray_t1, ws.bounds);
        P3D_LAUNCH_CHECK();
    }
    SampleArgs sa{};
    sa.g = g; sa.planes = planes; sa.w1 = w1; sa.b1 = b1; sa.w2 = w2; sa.b2 = b2; sa.ro = ro; sa.rd = rd;
    sa.ray_t0 = ws.ray_t0; sa.ray_t1 = ws.ray_t1; sa.bounds = ws.bounds;
    // coarse pass
    sa.u = u_c; sa.depth_out = ws.depth_c; sa.sigma_out = ws.sigma_c; sa.rgb_out = ws.rgb_c;
    sa.total = R * g.S; sa.per_view = (long long)g.M * g.S; sa.per_ray = g.S;
    int rc = launch_sample_decode<MODE_COARSE>(sa, p->planes_bf16 != 0, stream);
    if (rc) return rc;
    if (g.Sf > 0) {
        ImportanceArgs ia{};
        ia.g = g; ia.depth_c = ws.depth_c; ia.sigma_c = ws.sigma_c; ia.u_fine = u_f; ia.depth_f = ws.depth_f;
        ia.bounds = ws.bounds; ia.R = R;
        int p2 = 1; while (p2 < g.Sf) p2 <<= 1;
        ia.sort_pow2 = p2;
        const size_t smem = (size_t)4 * (4 * g.S + p2) * sizeof(float);
        {
            ProfileScope prof(PROF_IMPORTANCE, stream);
            k_ray_importance<<<(unsigned)((R + 3) / 4), 128, smem, stream>>>(ia);
            P3D_LAUNCH_CHECK();
        }
        sa.u = nullptr; sa.depth_in = ws.depth_f; sa.depth_out = nullptr; sa.sigma_out = ws.sigma_f; sa.rgb_out = ws.rgb_f;
        sa.total = R * g.Sf; sa.per_view = (long long)g.M * g.Sf; sa.per_ray = g.Sf;
        rc = launch_sample_decode<MODE_FINE>(sa, p->planes_bf16 != 0, stream);
        if (rc) return rc;
    }
    CompositeArgs ca{};
    ca.g = g; ca.depth_c = ws.depth_c; ca.sigma_c = ws.sigma_c; ca.rgb_c = ws.rgb_c; ca.depth_f = ws.depth_f;
    ca.sigma_f = ws.sigma_f; ca.rgb_f = ws.rgb_f; ca.ro = ro; ca.rd = rd; ca.out_rgb = out_rgb; ca.out_depth = out_depth;
    ca.out_wsum = out_wsum; ca.out_xyz = out_xyz; ca.bounds = ws.bounds; ca.R = R;
    const size_t smem = (size_t)4 * 4 * (g.S + g.Sf) * sizeof(float);
    {
        ProfileScope prof(PROF_COMPOSITE, stream);
        k_ray_composite<<<(unsigned)((R + 3) / 4), 128, smem, stream>>>(ca);
        P3D_LAUNCH_CHECK();
    }
    k_depth_finalize<<<(unsigned)((R + 255) / 256), 256, 0, stream>>>(out_depth, R, ws.bounds);
    P3D_LAUNCH_CHECK();
    return P3D_OK;
}

int decode_points_v1(const Geom& g, const p3d_render_params* p, const void* planes, const float* w1, const float* b1,
                     const float* w2, const float* b2, const float* coords, long long n_pts, float* out_rgb,
                     float* out_sigma, cudaStream_t stream) {
    SampleArgs sa{};
    sa.g = g; sa.planes = planes; sa.w1 = w1; sa.b1 = b1; sa.w2 = w2; sa.b2 = b2; sa.coords = coords;
    sa.sigma_out = out_sigma; sa.rgb_out = out_rgb;
    sa.total = (long long)g.N * n_pts; sa.per_view = n_pts; sa.per_ray = 1;
    return launch_sample_decode<MODE_POINTS>(sa, p->planes_bf16 != 0, stream);
}

}  // namespace p3d
