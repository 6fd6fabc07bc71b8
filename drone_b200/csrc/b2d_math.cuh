// b2d_math.cuh -- arithmetic building blocks for the drone step kernels (sm_100a).
//
// Two arithmetic policies share one source:
//   xf    : "exact float" -- every operator is ONE IEEE-754 binary32 operation
//           (__fadd_rn/__fmul_rn/__fdiv_rn/__fsqrt_rn are never contracted into
//           FMAs by nvcc), so code written with xf in the reference's
//           association reproduces gcc -O2 scalar SSE results bit for bit.
//   float : the fast policy; the compiler is free to fuse, reciprocals are
//           hoisted by hand.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2d {

struct xf {
    float v;
    __device__ __forceinline__ xf() {}
    __device__ __forceinline__ xf(float f) : v(f) {}
};
__device__ __forceinline__ xf operator+(xf a, xf b) { return xf(__fadd_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator-(xf a, xf b) { return xf(__fsub_rn(a.v, b.v)); }
__device__ __forceinline__ xf operator*(xf a, xf b) { return xf(__fmul_rn(a.v, b.v)); }
// IEEE division.  A zero numerator fails div.rn's range check (FCHK) and sends the warp down a ~100-instruction slow
// path with one lane active -- and a freshly reset drone is all zeros (velocity, rates, rotor speeds), so half of all
// warps took that path at dozens of divisions per strict step (swarm strict, A = 16: 357 -> 282 us).  0 / b for a finite
// non-zero b is a signed zero, exactly: those lanes divide 1 / b instead and get their zero back.
__device__ __forceinline__ float xdiv_rn(float a, float b) {
    const bool zero = a == 0.0f && b != 0.0f && fabsf(b) <= 3.402823466e38f;
    const float q = __fdiv_rn(zero ? 1.0f : a, b);
    return zero ? (b > 0.0f ? a : -a) : q;
}
// (operator/ stays the plain sequence: episode generation, whose numerators are never zero, is on the fast kernels' hot
// path; the strict step's own arithmetic and observations use sdiv<>)
__device__ __forceinline__ xf operator/(xf a, xf b) { return xf(__fdiv_rn(a.v, b.v)); }
// The same division as ONE out-of-line copy, for the race env's STRICT step (rigid-body rates, quaternion
// normalisation): it inlined ~45 division sequences and was instruction-cache bound (`no_instruction` 3.7 stall cycles
// per issue); calling one copy takes it from 182 to 149 us per 1 M envs.  (The swarm's strict step is slower with the
// call, 315 vs 282 us, and keeps the inline form: sdiv<false>.)
__device__ __noinline__ float xdiv_out(float a, float b) { return xdiv_rn(a, b); }
template <bool OUT> __device__ __forceinline__ xf sdiv(xf a, xf b) {
    if constexpr (OUT) return xf(xdiv_out(a.v, b.v));
    else return xf(xdiv_rn(a.v, b.v));
}
__device__ __forceinline__ xf operator-(xf a) { return xf(-a.v); }
__device__ __forceinline__ xf xsqrt(xf a) { return xf(__fsqrt_rn(a.v)); }
// clamp with the reference's branch order (NaN falls through), DR/dronelib.h:73-79
__device__ __forceinline__ xf xclamp(xf a, float lo, float hi) {
    return a.v < lo ? xf(lo) : (a.v > hi ? xf(hi) : a);
}

// Fast policy only: one MUFU each (rcp.approx <= 1 ulp, rsqrt.approx <= 2 ulp), no denormal /
// special-case fix-up code.  Arguments on the fast path are masses, inertias, motor constants and
// squared quaternion norms: always normal, positive numbers.  (__frcp_rn costs ~10 instructions,
// rsqrtf ~9; six reciprocals and four normalisations per env-step were 10 % of the issue count.)
__device__ __forceinline__ float approx_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float approx_rsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float fval(xf a) { return a.v; }
__device__ __forceinline__ float fval(float a) { return a; }
__device__ __forceinline__ xf tsqrt(xf a) { return xsqrt(a); }
__device__ __forceinline__ float tsqrt(float a) { return sqrtf(a); }

template <class T> struct V3 { T x, y, z; };
template <class T> struct Q4 { T w, x, y, z; };

// Hamilton product with left-to-right sums, DR/dronelib.h:112-119
template <class T>
__device__ __forceinline__ Q4<T> qmul(const Q4<T> &a, const Q4<T> &b) {
    Q4<T> r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}

// v' = q (0,v) q*, DR/dronelib.h:131-137
template <class T>
__device__ __forceinline__ V3<T> qrot(const Q4<T> &q, const V3<T> &v) {
    Q4<T> pure;
    pure.w = T(0.0f); pure.x = v.x; pure.y = v.y; pure.z = v.z;
    Q4<T> t = qmul(q, pure);
    Q4<T> qc;
    qc.w = q.w; qc.x = -q.x; qc.y = -q.y; qc.z = -q.z;
    Q4<T> r = qmul(t, qc);
    V3<T> o;
    o.x = r.x; o.y = r.y; o.z = r.z;
    return o;
}

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw; SC'11) ---------------------------
#ifndef B2D_PHILOX_UNROLL
#define B2D_PHILOX_UNROLL 10
#endif
constexpr int PHILOX_UNROLL = B2D_PHILOX_UNROLL;
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll PHILOX_UNROLL
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        uint4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k0;
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k1;
        n.w = (uint32_t)p0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

// rndf's unit sample from a 32-bit word: (float)(31-bit int) / 2^31, DR/dronelib.h:81-83
__device__ __forceinline__ xf unit_from_word(uint32_t w) {
    return xf(__fmul_rn(__int2float_rn((int)(w >> 1)), 4.656612873077392578125e-10f));
}
__device__ __forceinline__ xf lerp_u(float a, float b, xf u) { return xf(a) + u * (xf(b) - xf(a)); }

// Deterministic sin/cos for theta in [0, 2*pi], one IEEE binary32 operation at a time; must stay in
// lock-step with oracle/drone_oracle.c:orc_sincos_det (same constants, same order): quadrant
// k = floor(theta * 2/pi + 1/2) <= 4, three-term Cody-Waite reduction (k * DP1 and k * DP2 are exact),
// Cephes sinf / cosf minimax polynomials on [-pi/4, pi/4]; |error| < 1e-7.  (The round-1 version did this in
// double precision: 84 DMUL + 84 DADD per ring on the reset path of the step kernel.)
__device__ __forceinline__ void sincos_det(float theta, float &s, float &c) {
    const xf DP1(1.5703125f), DP2(4.837512969970703125e-4f), DP3(7.54978995489188216e-8f);
    const xf t(theta);
    const int k = __float2int_rz((t * xf(0.636619746685028076171875f) + xf(0.5f)).v);
    const xf kf((float)k);
    const xf r = ((t - kf * DP1) - kf * DP2) - kf * DP3;
    const xf z = r * r;
    xf ps(-1.9515295891e-4f);
    ps = ps * z + xf(8.3321608736e-3f);
    ps = ps * z + xf(-1.6666654611e-1f);
    const xf sr = ps * z * r + r;
    xf pc(2.443315711809948e-5f);
    pc = pc * z + xf(-1.388731625493765e-3f);
    pc = pc * z + xf(4.166664568298827e-2f);
    const xf cr = pc * z * z + (xf(1.0f) - xf(0.5f) * z);
    float sv, cv;
    switch (k & 3) {
    case 0: sv = sr.v; cv = cr.v; break;
    case 1: sv = cr.v; cv = -sr.v; break;
    case 2: sv = -sr.v; cv = -cr.v; break;
    default: sv = -cr.v; cv = sr.v; break;
    }
    s = sv;
    c = cv;
}

// x^3 rounded once from a double product (stands in for the reference's powf(x, 3.0f))
__device__ __forceinline__ float cube_det(float x) {
    double d = (double)x;
    return __double2float_rn(__dmul_rn(__dmul_rn(d, d), d));
}

} // namespace b2d
