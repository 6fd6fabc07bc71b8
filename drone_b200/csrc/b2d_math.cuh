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
__device__ __forceinline__ xf operator/(xf a, xf b) { return xf(__fdiv_rn(a.v, b.v)); }
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
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
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

// Deterministic sin/cos for theta in [0, 2*pi], one IEEE double op at a time; must stay in
// lock-step with oracle/drone_oracle.c:orc_sincos_det.
__device__ __forceinline__ void sincos_det(float theta, float &s, float &c) {
    const double TWO_OVER_PI = 0.63661977236758134308;
    const double PIO2 = 1.57079632679489661923;
    double t = (double)theta;
    int k = __double2int_rz(__dadd_rn(__dmul_rn(t, TWO_OVER_PI), 0.5));
    double r = __dsub_rn(t, __dmul_rn((double)k, PIO2));
    double z = __dmul_rn(r, r);
    double ps = -1.0 / 1307674368000.0;
    ps = __dadd_rn(__dmul_rn(ps, z), 1.0 / 6227020800.0);
    ps = __dadd_rn(__dmul_rn(ps, z), -1.0 / 39916800.0);
    ps = __dadd_rn(__dmul_rn(ps, z), 1.0 / 362880.0);
    ps = __dadd_rn(__dmul_rn(ps, z), -1.0 / 5040.0);
    ps = __dadd_rn(__dmul_rn(ps, z), 1.0 / 120.0);
    ps = __dadd_rn(__dmul_rn(ps, z), -1.0 / 6.0);
    ps = __dadd_rn(__dmul_rn(ps, z), 1.0);
    double sr = __dmul_rn(ps, r);
    double pc = 1.0 / 20922789888000.0;
    pc = __dadd_rn(__dmul_rn(pc, z), -1.0 / 87178291200.0);
    pc = __dadd_rn(__dmul_rn(pc, z), 1.0 / 479001600.0);
    pc = __dadd_rn(__dmul_rn(pc, z), -1.0 / 3628800.0);
    pc = __dadd_rn(__dmul_rn(pc, z), 1.0 / 40320.0);
    pc = __dadd_rn(__dmul_rn(pc, z), -1.0 / 720.0);
    pc = __dadd_rn(__dmul_rn(pc, z), 1.0 / 24.0);
    pc = __dadd_rn(__dmul_rn(pc, z), -0.5);
    double cr = __dadd_rn(__dmul_rn(pc, z), 1.0);
    double sv, cv;
    switch (k & 3) {
    case 0: sv = sr; cv = cr; break;
    case 1: sv = cr; cv = -sr; break;
    case 2: sv = -sr; cv = -cr; break;
    default: sv = -cr; cv = sr; break;
    }
    s = __double2float_rn(sv);
    c = __double2float_rn(cv);
}

// x^3 rounded once from a double product (stands in for the reference's powf(x, 3.0f))
__device__ __forceinline__ float cube_det(float x) {
    double d = (double)x;
    return __double2float_rn(__dmul_rn(__dmul_rn(d, d), d));
}

} // namespace b2d
