// PTX wrappers (mbarrier, tcgen05 MMA / TMEM, packed f32x2 arithmetic) and the tri-plane gather helpers shared by the fused
// tensor-core kernels: render_fused_ws3.cu (the renderer) and decode_tc.cu (point / volume decode).
#pragma once
#include "render_device.cuh"

namespace p3d {
namespace fused {

using namespace dev;

constexpr int kSBO = 128;                        // byte distance between 8-row groups of a K-major no-swizzle UMMA tile
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
#ifndef P3D_W3_WAIT
#define P3D_W3_WAIT 0          // 0: try_wait with a suspend-time hint, bounded (traps on a protocol bug); 2: plain try_wait, bounded (measured slower)
#endif

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#if P3D_W3_WAIT == 2
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {             // plain try_wait: the hardware's own suspend window per try; still bounded
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
#endif
    // (not unrolled: ptxas otherwise replicates the body 16x at each of the ~50 call sites - half of the kernel's code size)
#pragma unroll 1
    for (int it = 0; it < (1 << 17); ++it) {             // ~20 us per try: the cap turns a protocol bug into a trap after ~2 s
        uint32_t ok;
        // the suspend-time hint lets a waiting warp sleep in hardware instead of polling: roles that run ahead of the
        // critical path must not steal issue slots from it
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ constexpr uint32_t umma_idesc(int M, int N) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;" : "=r"(r) : "r"(taddr) : "memory");
    return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
    return d;
}
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void split1(float x, unsigned short& hi, unsigned short& lo) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
}
__device__ __forceinline__ int tile_off(int row, int k, int lbo) { return (row >> 3) * kSBO + (k >> 3) * lbo + (row & 7) * 16 + (k & 7) * 2; }

// ------------------------------------------------------------------------------------------ gather
// Tap table: one 64 B record per row - chunk 0 = {o[0], o[1], o[2], flags}, chunk 1+p = the four bilinear weights of plane p
// (w00, w01, w10, w11; rows y, y+1 x columns x, x+1) - chunk c of row r stored at position c ^ ((r >> 1) & 3), which makes
// both the lane = row writes and the 8-rows-per-phase reads of the gather conflict-free.  o[p] is the element offset of texel
// (ya, xa) of plane p inside the view's tri-plane, with (ya, xa) CLAMPED to [0, H-2] x [0, W-2]: the 2 x 2 footprint that is
// loaded is always inside the plane, and the weights are moved onto the loaded texels (grid_sample's zero padding,
// renderer.py:68-81: a tap outside the plane contributes nothing, so its weight is dropped; a footprint that hangs over the
// edge by one texel keeps the weights of its inside texels).  flags bit p = plane p has a non-zero weight; the loads of a
// plane whose bit is clear are predicated off (its registers keep older, finite texel values that meet weights of 0).
__device__ __forceinline__ uint32_t tab_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__device__ __forceinline__ bool plane_taps_rec(const Geom& g, int srow, int scol, int pbase, float ca, float cb, uint32_t& o, float4& w) {
    const float gx = __fmul_rn(ca, g.coord_scale), gy = __fmul_rn(cb, g.coord_scale);
    const float fx = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)g.W), 1.f), 0.5f);
    const float fy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)g.H), 1.f), 0.5f);
    const bool sane = (fx > -2.f) && (fx < (float)g.W + 1.f) && (fy > -2.f) && (fy < (float)g.H + 1.f);   // false for NaN too
    const float x0f = floorf(fx), y0f = floorf(fy);
    const float wx1 = fx - x0f, wy1 = fy - y0f;
    const float wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const int x0 = sane ? (int)x0f : -4, y0 = sane ? (int)y0f : -4;
    const int xa = min(max(x0, 0), g.W - 2), ya = min(max(y0, 0), g.H - 2);
    const int dx = x0 - xa, dy = y0 - ya;                     // 0 inside, -1 / +1: the footprint hangs over the low / high edge
    const float wl = dx == 0 ? wx0 : (dx == -1 ? wx1 : 0.f), wr = dx == 0 ? wx1 : (dx == 1 ? wx0 : 0.f);
    const float wt = dy == 0 ? wy0 : (dy == -1 ? wy1 : 0.f), wb = dy == 0 ? wy1 : (dy == 1 ? wy0 : 0.f);
    o = (uint32_t)(pbase + ya * srow + xa * scol);
    w = make_float4(wl * wt, wr * wt, wl * wb, wr * wb);
    return ((unsigned)(dx + 1) < 3u) && ((unsigned)(dy + 1) < 3u);
}
// packed f32x2 arithmetic (one instruction for two lanes of data)
__device__ __forceinline__ unsigned long long pk2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// lg2(1 + t) for t in [0, 1] on the FMA pipe, two values per instruction: t * q(t), q = degree-6 near-minimax fit of
// lg2(1 + t) / t (max abs error 1.4e-6 in fp32 Horner form, exact 0 at t = 0).  Takes the second MUFU of every softplus off the
// MUFU / MIO queue, which is what throttles the tile epilogue (ncu: mio_throttle on the softplus lines).
__device__ __forceinline__ unsigned long long lg2_1p_poly2(unsigned long long t) {
    unsigned long long q = fma2(t, pk2(0.020490340888500214f, 0.020490340888500214f), pk2(-0.09606623649597168f, -0.09606623649597168f));
    q = fma2(t, q, pk2(0.2155885100364685f, 0.2155885100364685f));
    q = fma2(t, q, pk2(-0.33924776315689087f, -0.33924776315689087f));
    q = fma2(t, q, pk2(0.4777059257030487f, 0.4777059257030487f));
    q = fma2(t, q, pk2(-0.721162736415863f, -0.721162736415863f));
    q = fma2(t, q, pk2(1.4426932334899902f, 1.4426932334899902f));
    return mul2(t, q);
}
// (h0, h1) -> bf16x2 hi word and bf16x2 lo word of the residuals, 4 instructions + the pack
__device__ __forceinline__ void split2p(unsigned long long h2, uint32_t& hi, uint32_t& lo) {
    float h0, h1;
    upk2(h2, h0, h1);
    hi = pack_bf16x2(h0, h1);
    const unsigned long long r2 = sub2(h2, pk2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u)));
    float r0, r1;
    upk2(r2, r0, r1);
    lo = pack_bf16x2(r0, r1);
}
// four consecutive channels of one texel (16 B fp32 / 8 B bf16) at addr + IMM bytes, predicated; v keeps its old value when !pred
template <bool BF16, int IMM>
__device__ __forceinline__ void load_quad_p(float (&v)[4], const char* addr, bool pred) {
    if (BF16) {
        uint32_t r0 = 0, r1 = 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.global.nc.v2.u32 {%0,%1}, [%2+%4];\n\t}"
                     : "+r"(r0), "+r"(r1) : "l"(addr), "r"((int)pred), "n"(IMM));
        v[0] = __uint_as_float(r0 << 16); v[1] = __uint_as_float(r0 & 0xffff0000u); v[2] = __uint_as_float(r1 << 16); v[3] = __uint_as_float(r1 & 0xffff0000u);   // zeros when !pred
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4+%6];\n\t}"
                     : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) : "l"(addr), "r"((int)pred), "n"(IMM));
    }
}
__device__ __forceinline__ void fma4(unsigned long long (&acc)[2], const float (&v)[4], float w) {
    unsigned long long ww;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        unsigned long long vv;
        asm("mov.b64 %0, {%1, %2};" : "=l"(vv) : "f"(v[2 * j]), "f"(v[2 * j + 1]));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(vv), "l"(ww));
    }
}

}  // namespace fused
}  // namespace p3d
