// physics.cuh -- quadrotor rigid-body step shared by the race and swarm kernels.
//
// What it computes follows the reference's dronelib (R = pufferlib/ocean/drone_race):
//   rates()        R/dronelib.h:302-381  compute_derivatives
//   probe()        R/dronelib.h:383-392  step (Euler probe + quaternion renormalise)
//   advance_body() R/dronelib.h:394-449  rk4_step + move_drone's clamps
// STRICT=true transcribes one IEEE op per reference op (bit-exact with the CPU
// build); STRICT=false is the production path: FMA contraction, reciprocals
// hoisted out of the four derivative evaluations, the thrust rotation reduced
// to the third column of the rotation matrix, rsqrt renormalisation.
#pragma once
#include "b2d_math.cuh"

namespace b2d {

struct DroneParams {
    float mass, ixx, iyy, izz, arm, kt, kad, kd, bd, g, mrpm, kmot, jmot;
};

template <class T> struct Body {
    V3<T> pos, vel;
    Q4<T> q;
    V3<T> w;
    T rpm[4];
};
template <class T> struct Rate { // d(pos)/dt is the probed body's own velocity
    V3<T> dvel;
    Q4<T> dq;
    V3<T> dw;
    T drpm[4];
};

#define B2D_DT 0.05f
#define B2D_MAX_VEL 50.0f
#define B2D_MAX_OMEGA 50.0f

// ------------------------------------------------------------------ strict path
template <bool OUT = true>
__device__ __forceinline__ void rates_strict(const Body<xf> &b, const DroneParams &p, const xf want[4],
                                             xf inv_kmot, Rate<xf> &k) {
    xf thrust[4];
#pragma unroll
    for (int m = 0; m < 4; m++) k.drpm[m] = inv_kmot * (want[m] - b.rpm[m]);
#pragma unroll
    for (int m = 0; m < 4; m++) thrust[m] = xf(p.kt) * (b.rpm[m] * b.rpm[m]);
    V3<xf> lift_body;
    lift_body.x = xf(0.0f); lift_body.y = xf(0.0f);
    lift_body.z = thrust[0] + thrust[1] + thrust[2] + thrust[3];
    V3<xf> lift = qrot(b.q, lift_body);
    xf nbd = xf(-p.bd);
    k.dvel.x = sdiv<OUT>(lift.x + nbd * b.vel.x, xf(p.mass));
    k.dvel.y = sdiv<OUT>(lift.y + nbd * b.vel.y, xf(p.mass));
    k.dvel.z = sdiv<OUT>(lift.z + nbd * b.vel.z, xf(p.mass)) - xf(p.g);
    Q4<xf> wq;
    wq.w = xf(0.0f); wq.x = b.w.x; wq.y = b.w.y; wq.z = b.w.z;
    Q4<xf> dq = qmul(b.q, wq);
    k.dq.w = dq.w * xf(0.5f); k.dq.x = dq.x * xf(0.5f); k.dq.y = dq.y * xf(0.5f); k.dq.z = dq.z * xf(0.5f);
    xf tpx = xf(p.arm) * (thrust[1] - thrust[3]);
    xf tpy = xf(p.arm) * (thrust[2] - thrust[0]);
    xf tpz = xf(p.kd) * (thrust[0] - thrust[1] + thrust[2] - thrust[3]);
    xf tmz = xf(p.jmot) * (k.drpm[0] - k.drpm[1] + k.drpm[2] - k.drpm[3]);
    xf nkad = xf(-p.kad);
    xf tix = (xf(p.iyy) - xf(p.izz)) * b.w.y * b.w.z;
    xf tiy = (xf(p.izz) - xf(p.ixx)) * b.w.z * b.w.x;
    xf tiz = (xf(p.ixx) - xf(p.iyy)) * b.w.x * b.w.y;
    k.dw.x = sdiv<OUT>(tpx + nkad * b.w.x + tix, xf(p.ixx));
    k.dw.y = sdiv<OUT>(tpy + nkad * b.w.y + tiy, xf(p.iyy));
    k.dw.z = sdiv<OUT>(tpz + nkad * b.w.z + tiz + tmz, xf(p.izz));
}

template <bool OUT = true>
__device__ __forceinline__ void probe_strict(const Body<xf> &b, const V3<xf> &dpos, const Rate<xf> &k, xf h,
                                             Body<xf> &o) {
    o.pos.x = b.pos.x + dpos.x * h; o.pos.y = b.pos.y + dpos.y * h; o.pos.z = b.pos.z + dpos.z * h;
    o.vel.x = b.vel.x + k.dvel.x * h; o.vel.y = b.vel.y + k.dvel.y * h; o.vel.z = b.vel.z + k.dvel.z * h;
    o.q.w = b.q.w + k.dq.w * h; o.q.x = b.q.x + k.dq.x * h;
    o.q.y = b.q.y + k.dq.y * h; o.q.z = b.q.z + k.dq.z * h;
    o.w.x = b.w.x + k.dw.x * h; o.w.y = b.w.y + k.dw.y * h; o.w.z = b.w.z + k.dw.z * h;
#pragma unroll
    for (int m = 0; m < 4; m++) o.rpm[m] = b.rpm[m] + k.drpm[m] * h;
    xf n = xsqrt(o.q.w * o.q.w + o.q.x * o.q.x + o.q.y * o.q.y + o.q.z * o.q.z);
    if (n.v > 0.0f) {
        o.q.w = sdiv<OUT>(o.q.w, n); o.q.x = sdiv<OUT>(o.q.x, n); o.q.y = sdiv<OUT>(o.q.y, n); o.q.z = sdiv<OUT>(o.q.z, n);
    }
}

// state layout: s[0:3] pos, [3:6] vel, [6:10] quat wxyz, [10:13] omega, [13:17] rpm
template <bool OUT = true>
__device__ __forceinline__ void advance_body_strict(float s[17], const DroneParams &p, const float act[4]) {
    Body<xf> b, tmp;
    b.pos.x = s[0]; b.pos.y = s[1]; b.pos.z = s[2];
    b.vel.x = s[3]; b.vel.y = s[4]; b.vel.z = s[5];
    b.q.w = s[6]; b.q.x = s[7]; b.q.y = s[8]; b.q.z = s[9];
    b.w.x = s[10]; b.w.y = s[11]; b.w.z = s[12];
#pragma unroll
    for (int m = 0; m < 4; m++) b.rpm[m] = s[13 + m];
    xf want[4];
#pragma unroll
    for (int m = 0; m < 4; m++) want[m] = (xf(act[m]) + xf(1.0f)) * xf(0.5f) * xf(p.mrpm);
    const xf inv_kmot = xf(1.0f) / xf(p.kmot);
    const xf h = xf(B2D_DT) * xf(1.0f);
    const xf hh = h * xf(0.5f);
    const xf two = xf(2.0f);

    Rate<xf> k, acc;
    V3<xf> accp;
    // k1
    rates_strict<OUT>(b, p, want, inv_kmot, k);
    acc = k;
    accp = b.vel;
    probe_strict<OUT>(b, b.vel, k, hh, tmp);
    // k2
    rates_strict<OUT>(tmp, p, want, inv_kmot, k);
    {
        V3<xf> v2 = tmp.vel;
#define B2D_ACC2(f) acc.f = acc.f + two * k.f
        accp.x = accp.x + two * v2.x; accp.y = accp.y + two * v2.y; accp.z = accp.z + two * v2.z;
        B2D_ACC2(dvel.x); B2D_ACC2(dvel.y); B2D_ACC2(dvel.z);
        B2D_ACC2(dq.w); B2D_ACC2(dq.x); B2D_ACC2(dq.y); B2D_ACC2(dq.z);
        B2D_ACC2(dw.x); B2D_ACC2(dw.y); B2D_ACC2(dw.z);
        B2D_ACC2(drpm[0]); B2D_ACC2(drpm[1]); B2D_ACC2(drpm[2]); B2D_ACC2(drpm[3]);
        probe_strict<OUT>(b, v2, k, hh, tmp);
    }
    // k3
    rates_strict<OUT>(tmp, p, want, inv_kmot, k);
    {
        V3<xf> v3 = tmp.vel;
        accp.x = accp.x + two * v3.x; accp.y = accp.y + two * v3.y; accp.z = accp.z + two * v3.z;
        B2D_ACC2(dvel.x); B2D_ACC2(dvel.y); B2D_ACC2(dvel.z);
        B2D_ACC2(dq.w); B2D_ACC2(dq.x); B2D_ACC2(dq.y); B2D_ACC2(dq.z);
        B2D_ACC2(dw.x); B2D_ACC2(dw.y); B2D_ACC2(dw.z);
        B2D_ACC2(drpm[0]); B2D_ACC2(drpm[1]); B2D_ACC2(drpm[2]); B2D_ACC2(drpm[3]);
#undef B2D_ACC2
        probe_strict<OUT>(b, v3, k, h, tmp);
    }
    // k4
    rates_strict<OUT>(tmp, p, want, inv_kmot, k);
    const xf h6 = h / xf(6.0f);
#define B2D_FIN(dst, a, kk) dst = dst + ((a) + (kk)) * h6
    B2D_FIN(b.pos.x, accp.x, tmp.vel.x); B2D_FIN(b.pos.y, accp.y, tmp.vel.y); B2D_FIN(b.pos.z, accp.z, tmp.vel.z);
    B2D_FIN(b.vel.x, acc.dvel.x, k.dvel.x); B2D_FIN(b.vel.y, acc.dvel.y, k.dvel.y); B2D_FIN(b.vel.z, acc.dvel.z, k.dvel.z);
    B2D_FIN(b.q.w, acc.dq.w, k.dq.w); B2D_FIN(b.q.x, acc.dq.x, k.dq.x);
    B2D_FIN(b.q.y, acc.dq.y, k.dq.y); B2D_FIN(b.q.z, acc.dq.z, k.dq.z);
    B2D_FIN(b.w.x, acc.dw.x, k.dw.x); B2D_FIN(b.w.y, acc.dw.y, k.dw.y); B2D_FIN(b.w.z, acc.dw.z, k.dw.z);
#pragma unroll
    for (int m = 0; m < 4; m++) B2D_FIN(b.rpm[m], acc.drpm[m], k.drpm[m]);
#undef B2D_FIN
    xf n = xsqrt(b.q.w * b.q.w + b.q.x * b.q.x + b.q.y * b.q.y + b.q.z * b.q.z);
    if (n.v > 0.0f) {
        b.q.w = sdiv<OUT>(b.q.w, n); b.q.x = sdiv<OUT>(b.q.x, n); b.q.y = sdiv<OUT>(b.q.y, n); b.q.z = sdiv<OUT>(b.q.z, n);
    }
    s[0] = b.pos.x.v; s[1] = b.pos.y.v; s[2] = b.pos.z.v;
    s[3] = xclamp(b.vel.x, -B2D_MAX_VEL, B2D_MAX_VEL).v;
    s[4] = xclamp(b.vel.y, -B2D_MAX_VEL, B2D_MAX_VEL).v;
    s[5] = xclamp(b.vel.z, -B2D_MAX_VEL, B2D_MAX_VEL).v;
    s[6] = b.q.w.v; s[7] = b.q.x.v; s[8] = b.q.y.v; s[9] = b.q.z.v;
    s[10] = xclamp(b.w.x, -B2D_MAX_OMEGA, B2D_MAX_OMEGA).v;
    s[11] = xclamp(b.w.y, -B2D_MAX_OMEGA, B2D_MAX_OMEGA).v;
    s[12] = xclamp(b.w.z, -B2D_MAX_OMEGA, B2D_MAX_OMEGA).v;
#pragma unroll
    for (int m = 0; m < 4; m++) s[13 + m] = b.rpm[m].v;
}

// ------------------------------------------------------------------ fast path
struct FastConsts {
    float inv_mass, inv_ixx, inv_iyy, inv_izz, inv_kmot;
    float d_yz, d_zx, d_xy; // inertia differences of the gyroscopic term
    float want_k[4]; // commanded rpm / k_mot
};

__device__ __forceinline__ void rates_fast(const Body<float> &b, const DroneParams &p, const FastConsts &c,
                                           Rate<float> &k) {
    float t0 = p.kt * (b.rpm[0] * b.rpm[0]);
    float t1 = p.kt * (b.rpm[1] * b.rpm[1]);
    float t2 = p.kt * (b.rpm[2] * b.rpm[2]);
    float t3 = p.kt * (b.rpm[3] * b.rpm[3]);
#pragma unroll
    for (int m = 0; m < 4; m++) k.drpm[m] = fmaf(-c.inv_kmot, b.rpm[m], c.want_k[m]); // (want - rpm) / k_mot
    const float lift = (t0 + t1) + (t2 + t3);
    const Q4<float> &q = b.q;
    // q (0,0,0,L) q* = L * third column of the (unnormalised) rotation matrix
    float ax = 2.0f * (q.x * q.z + q.w * q.y);
    float ay = 2.0f * (q.y * q.z - q.w * q.x);
    float az = (q.w * q.w - q.x * q.x) + (q.z * q.z - q.y * q.y);
    k.dvel.x = (lift * ax - p.bd * b.vel.x) * c.inv_mass;
    k.dvel.y = (lift * ay - p.bd * b.vel.y) * c.inv_mass;
    k.dvel.z = (lift * az - p.bd * b.vel.z) * c.inv_mass - p.g;
    k.dq.w = 0.5f * (-q.x * b.w.x - q.y * b.w.y - q.z * b.w.z);
    k.dq.x = 0.5f * (q.w * b.w.x + q.y * b.w.z - q.z * b.w.y);
    k.dq.y = 0.5f * (q.w * b.w.y - q.x * b.w.z + q.z * b.w.x);
    k.dq.z = 0.5f * (q.w * b.w.z + q.x * b.w.y - q.y * b.w.x);
    float tpx = p.arm * (t1 - t3);
    float tpy = p.arm * (t2 - t0);
    float tpz = p.kd * ((t0 - t1) + (t2 - t3));
    float tmz = p.jmot * ((k.drpm[0] - k.drpm[1]) + (k.drpm[2] - k.drpm[3]));
    k.dw.x = (tpx - p.kad * b.w.x + c.d_yz * b.w.y * b.w.z) * c.inv_ixx;
    k.dw.y = (tpy - p.kad * b.w.y + c.d_zx * b.w.z * b.w.x) * c.inv_iyy;
    k.dw.z = (tpz - p.kad * b.w.z + c.d_xy * b.w.x * b.w.y + tmz) * c.inv_izz;
}

__device__ __forceinline__ void qnormalize_fast(Q4<float> &q) {
    float n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    float inv = n2 > 0.0f ? approx_rsqrt(n2) : 1.0f;
    q.w *= inv; q.x *= inv; q.y *= inv; q.z *= inv;
}

__device__ __forceinline__ void probe_fast(const Body<float> &b, const V3<float> &dpos, const Rate<float> &k,
                                           float h, Body<float> &o) {
    o.pos.x = b.pos.x + dpos.x * h; o.pos.y = b.pos.y + dpos.y * h; o.pos.z = b.pos.z + dpos.z * h;
    o.vel.x = b.vel.x + k.dvel.x * h; o.vel.y = b.vel.y + k.dvel.y * h; o.vel.z = b.vel.z + k.dvel.z * h;
    o.q.w = b.q.w + k.dq.w * h; o.q.x = b.q.x + k.dq.x * h;
    o.q.y = b.q.y + k.dq.y * h; o.q.z = b.q.z + k.dq.z * h;
    o.w.x = b.w.x + k.dw.x * h; o.w.y = b.w.y + k.dw.y * h; o.w.z = b.w.z + k.dw.z * h;
#pragma unroll
    for (int m = 0; m < 4; m++) o.rpm[m] = b.rpm[m] + k.drpm[m] * h;
    qnormalize_fast(o.q);
}

__device__ __forceinline__ void advance_body_fast(float s[17], const DroneParams &p, const float act[4]) {
    Body<float> b, tmp;
    b.pos.x = s[0]; b.pos.y = s[1]; b.pos.z = s[2];
    b.vel.x = s[3]; b.vel.y = s[4]; b.vel.z = s[5];
    b.q.w = s[6]; b.q.x = s[7]; b.q.y = s[8]; b.q.z = s[9];
    b.w.x = s[10]; b.w.y = s[11]; b.w.z = s[12];
#pragma unroll
    for (int m = 0; m < 4; m++) b.rpm[m] = s[13 + m];
    FastConsts c;
    c.inv_mass = approx_rcp(p.mass);
    c.inv_ixx = approx_rcp(p.ixx);
    c.inv_iyy = approx_rcp(p.iyy);
    c.inv_izz = approx_rcp(p.izz);
    c.inv_kmot = approx_rcp(p.kmot);
    c.d_yz = p.iyy - p.izz; c.d_zx = p.izz - p.ixx; c.d_xy = p.ixx - p.iyy;
    const float half_mrpm = 0.5f * p.mrpm;
    const float half_mrpm_k = half_mrpm * c.inv_kmot;
#pragma unroll
    for (int m = 0; m < 4; m++) c.want_k[m] = fmaf(act[m], half_mrpm_k, half_mrpm_k);
    const float h = B2D_DT, hh = 0.5f * B2D_DT, h6 = B2D_DT / 6.0f;

    Rate<float> k, acc;
    V3<float> accp;
    rates_fast(b, p, c, k);
    acc = k;
    accp = b.vel;
    probe_fast(b, b.vel, k, hh, tmp);
#define B2D_ACC2(f) acc.f = fmaf(2.0f, k.f, acc.f)
#define B2D_ACCALL()                                                                     \
    B2D_ACC2(dvel.x); B2D_ACC2(dvel.y); B2D_ACC2(dvel.z);                                \
    B2D_ACC2(dq.w); B2D_ACC2(dq.x); B2D_ACC2(dq.y); B2D_ACC2(dq.z);                      \
    B2D_ACC2(dw.x); B2D_ACC2(dw.y); B2D_ACC2(dw.z);                                      \
    B2D_ACC2(drpm[0]); B2D_ACC2(drpm[1]); B2D_ACC2(drpm[2]); B2D_ACC2(drpm[3])
    rates_fast(tmp, p, c, k);
    {
        V3<float> v = tmp.vel;
        accp.x = fmaf(2.0f, v.x, accp.x); accp.y = fmaf(2.0f, v.y, accp.y); accp.z = fmaf(2.0f, v.z, accp.z);
        B2D_ACCALL();
        probe_fast(b, v, k, hh, tmp);
    }
    rates_fast(tmp, p, c, k);
    {
        V3<float> v = tmp.vel;
        accp.x = fmaf(2.0f, v.x, accp.x); accp.y = fmaf(2.0f, v.y, accp.y); accp.z = fmaf(2.0f, v.z, accp.z);
        B2D_ACCALL();
        probe_fast(b, v, k, h, tmp);
    }
#undef B2D_ACCALL
#undef B2D_ACC2
    rates_fast(tmp, p, c, k);
#define B2D_FIN(dst, a, kk) dst = fmaf((a) + (kk), h6, dst)
    B2D_FIN(b.pos.x, accp.x, tmp.vel.x); B2D_FIN(b.pos.y, accp.y, tmp.vel.y); B2D_FIN(b.pos.z, accp.z, tmp.vel.z);
    B2D_FIN(b.vel.x, acc.dvel.x, k.dvel.x); B2D_FIN(b.vel.y, acc.dvel.y, k.dvel.y); B2D_FIN(b.vel.z, acc.dvel.z, k.dvel.z);
    B2D_FIN(b.q.w, acc.dq.w, k.dq.w); B2D_FIN(b.q.x, acc.dq.x, k.dq.x);
    B2D_FIN(b.q.y, acc.dq.y, k.dq.y); B2D_FIN(b.q.z, acc.dq.z, k.dq.z);
    B2D_FIN(b.w.x, acc.dw.x, k.dw.x); B2D_FIN(b.w.y, acc.dw.y, k.dw.y); B2D_FIN(b.w.z, acc.dw.z, k.dw.z);
#pragma unroll
    for (int m = 0; m < 4; m++) B2D_FIN(b.rpm[m], acc.drpm[m], k.drpm[m]);
#undef B2D_FIN
    qnormalize_fast(b.q);
    s[0] = b.pos.x; s[1] = b.pos.y; s[2] = b.pos.z;
    s[3] = fminf(fmaxf(b.vel.x, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[4] = fminf(fmaxf(b.vel.y, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[5] = fminf(fmaxf(b.vel.z, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[6] = b.q.w; s[7] = b.q.x; s[8] = b.q.y; s[9] = b.q.z;
    s[10] = fminf(fmaxf(b.w.x, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[11] = fminf(fmaxf(b.w.y, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[12] = fminf(fmaxf(b.w.z, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
#pragma unroll
    for (int m = 0; m < 4; m++) s[13 + m] = b.rpm[m];
}

// ------------------------------------------------------------------ fast path, packed FP32x2
// The same arithmetic as advance_body_fast, operation for operation, with the 14 state words
// that enter the RK4 axpys as freshly computed rates held in register PAIRS: sm_100's
// fma/mul/add.rn.f32x2 do two IEEE operations per issue slot, and the step is close to the issue
// limit since resets are generated in the tile.  Positions stay scalar (their rate is the probed
// velocity, which already lives in another pair).  Saves ~50 issue slots of ~900 per env-step.
#ifndef B2D_PACKED_RK4
#define B2D_PACKED_RK4 1
#endif
struct BodyP {
    V3<float> pos;
    float2 v01, v2w0, w12; // (vx,vy) (vz,wx) (wy,wz)
    float2 q01, q23;       // (qw,qx) (qy,qz)
    float2 r01, r23;       // rotor speeds
};
struct RateP { // d/dt of the packed members; d(pos)/dt is the probed body's own velocity
    float2 v01, v2w0, w12, q01, q23, r01, r23;
};
__device__ __forceinline__ float2 pk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 pk1(float a) { return make_float2(a, a); }

__device__ __forceinline__ void rates_fast_p(const BodyP &b, const DroneParams &p, const FastConsts &c, RateP &k) {
    const float2 kt2 = pk1(p.kt);
    const float2 t01 = __fmul2_rn(kt2, __fmul2_rn(b.r01, b.r01));
    const float2 t23 = __fmul2_rn(kt2, __fmul2_rn(b.r23, b.r23));
    const float2 nik = pk1(-c.inv_kmot);
    k.r01 = __ffma2_rn(nik, b.r01, pk(c.want_k[0], c.want_k[1])); // (want - rpm) / k_mot
    k.r23 = __ffma2_rn(nik, b.r23, pk(c.want_k[2], c.want_k[3]));
    const float t0 = t01.x, t1 = t01.y, t2 = t23.x, t3 = t23.y;
    const float lift = (t0 + t1) + (t2 + t3);
    const float qw = b.q01.x, qx = b.q01.y, qy = b.q23.x, qz = b.q23.y;
    const float vx = b.v01.x, vy = b.v01.y, vz = b.v2w0.x;
    const float wx = b.v2w0.y, wy = b.w12.x, wz = b.w12.y;
    // q (0,0,0,L) q* = L * third column of the (unnormalised) rotation matrix
    const float ax = 2.0f * (qx * qz + qw * qy);
    const float ay = 2.0f * (qy * qz - qw * qx);
    const float az = (qw * qw - qx * qx) + (qz * qz - qy * qy);
    const float dvx = (lift * ax - p.bd * vx) * c.inv_mass;
    const float dvy = (lift * ay - p.bd * vy) * c.inv_mass;
    const float dvz = (lift * az - p.bd * vz) * c.inv_mass - p.g;
    const float dqw = 0.5f * (-qx * wx - qy * wy - qz * wz);
    const float dqx = 0.5f * (qw * wx + qy * wz - qz * wy);
    const float dqy = 0.5f * (qw * wy - qx * wz + qz * wx);
    const float dqz = 0.5f * (qw * wz + qx * wy - qy * wx);
    const float tpx = p.arm * (t1 - t3);
    const float tpy = p.arm * (t2 - t0);
    const float tpz = p.kd * ((t0 - t1) + (t2 - t3));
    const float tmz = p.jmot * ((k.r01.x - k.r01.y) + (k.r23.x - k.r23.y));
    const float dwx = (tpx - p.kad * wx + c.d_yz * wy * wz) * c.inv_ixx;
    const float dwy = (tpy - p.kad * wy + c.d_zx * wz * wx) * c.inv_iyy;
    const float dwz = (tpz - p.kad * wz + c.d_xy * wx * wy + tmz) * c.inv_izz;
    k.v01 = pk(dvx, dvy); k.v2w0 = pk(dvz, dwx); k.w12 = pk(dwy, dwz);
    k.q01 = pk(dqw, dqx); k.q23 = pk(dqy, dqz);
}

__device__ __forceinline__ void qnormalize_fast_p(float2 &q01, float2 &q23) {
    const float n2 = q01.x * q01.x + q01.y * q01.y + q23.x * q23.x + q23.y * q23.y;
    const float2 inv = pk1(n2 > 0.0f ? approx_rsqrt(n2) : 1.0f);
    q01 = __fmul2_rn(q01, inv);
    q23 = __fmul2_rn(q23, inv);
}

// o = b + h * (rates k; position rate = velocity of the body the rates were taken at)
__device__ __forceinline__ void probe_fast_p(const BodyP &b, float dpx, float dpy, float dpz, const RateP &k, float h, BodyP &o) {
    const float2 h2 = pk1(h);
    o.pos.x = b.pos.x + dpx * h; o.pos.y = b.pos.y + dpy * h; o.pos.z = b.pos.z + dpz * h;
    o.v01 = __ffma2_rn(k.v01, h2, b.v01); o.v2w0 = __ffma2_rn(k.v2w0, h2, b.v2w0); o.w12 = __ffma2_rn(k.w12, h2, b.w12);
    o.q01 = __ffma2_rn(k.q01, h2, b.q01); o.q23 = __ffma2_rn(k.q23, h2, b.q23);
    o.r01 = __ffma2_rn(k.r01, h2, b.r01); o.r23 = __ffma2_rn(k.r23, h2, b.r23);
    qnormalize_fast_p(o.q01, o.q23);
}

__device__ __forceinline__ void advance_body_fast_packed(float s[17], const DroneParams &p, const float act[4]) {
    BodyP b, tmp;
    b.pos.x = s[0]; b.pos.y = s[1]; b.pos.z = s[2];
    b.v01 = pk(s[3], s[4]); b.v2w0 = pk(s[5], s[10]); b.w12 = pk(s[11], s[12]);
    b.q01 = pk(s[6], s[7]); b.q23 = pk(s[8], s[9]);
    b.r01 = pk(s[13], s[14]); b.r23 = pk(s[15], s[16]);
    FastConsts c;
    c.inv_mass = approx_rcp(p.mass);
    c.inv_ixx = approx_rcp(p.ixx);
    c.inv_iyy = approx_rcp(p.iyy);
    c.inv_izz = approx_rcp(p.izz);
    c.inv_kmot = approx_rcp(p.kmot);
    c.d_yz = p.iyy - p.izz; c.d_zx = p.izz - p.ixx; c.d_xy = p.ixx - p.iyy;
    const float half_mrpm = 0.5f * p.mrpm;
    const float half_mrpm_k = half_mrpm * c.inv_kmot;
#pragma unroll
    for (int m = 0; m < 4; m++) c.want_k[m] = fmaf(act[m], half_mrpm_k, half_mrpm_k);
    const float h = B2D_DT, hh = 0.5f * B2D_DT;
    const float2 two = pk1(2.0f), h6 = pk1(B2D_DT / 6.0f);

    RateP k, acc;
    V3<float> accp;
    rates_fast_p(b, p, c, k);
    acc = k;
    accp.x = b.v01.x; accp.y = b.v01.y; accp.z = b.v2w0.x;
    probe_fast_p(b, b.v01.x, b.v01.y, b.v2w0.x, k, hh, tmp);
#define B2D_ACCP() \
    acc.v01 = __ffma2_rn(two, k.v01, acc.v01); acc.v2w0 = __ffma2_rn(two, k.v2w0, acc.v2w0); acc.w12 = __ffma2_rn(two, k.w12, acc.w12); \
    acc.q01 = __ffma2_rn(two, k.q01, acc.q01); acc.q23 = __ffma2_rn(two, k.q23, acc.q23);                                               \
    acc.r01 = __ffma2_rn(two, k.r01, acc.r01); acc.r23 = __ffma2_rn(two, k.r23, acc.r23)
    rates_fast_p(tmp, p, c, k);
    {
        const float vx = tmp.v01.x, vy = tmp.v01.y, vz = tmp.v2w0.x;
        accp.x = fmaf(2.0f, vx, accp.x); accp.y = fmaf(2.0f, vy, accp.y); accp.z = fmaf(2.0f, vz, accp.z);
        B2D_ACCP();
        probe_fast_p(b, vx, vy, vz, k, hh, tmp);
    }
    rates_fast_p(tmp, p, c, k);
    {
        const float vx = tmp.v01.x, vy = tmp.v01.y, vz = tmp.v2w0.x;
        accp.x = fmaf(2.0f, vx, accp.x); accp.y = fmaf(2.0f, vy, accp.y); accp.z = fmaf(2.0f, vz, accp.z);
        B2D_ACCP();
        probe_fast_p(b, vx, vy, vz, k, h, tmp);
    }
#undef B2D_ACCP
    rates_fast_p(tmp, p, c, k);
    const float h6s = B2D_DT / 6.0f;
    b.pos.x = fmaf(accp.x + tmp.v01.x, h6s, b.pos.x);
    b.pos.y = fmaf(accp.y + tmp.v01.y, h6s, b.pos.y);
    b.pos.z = fmaf(accp.z + tmp.v2w0.x, h6s, b.pos.z);
#define B2D_FINP(f) b.f = __ffma2_rn(__fadd2_rn(acc.f, k.f), h6, b.f)
    B2D_FINP(v01); B2D_FINP(v2w0); B2D_FINP(w12); B2D_FINP(q01); B2D_FINP(q23); B2D_FINP(r01); B2D_FINP(r23);
#undef B2D_FINP
    qnormalize_fast_p(b.q01, b.q23);
    s[0] = b.pos.x; s[1] = b.pos.y; s[2] = b.pos.z;
    s[3] = fminf(fmaxf(b.v01.x, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[4] = fminf(fmaxf(b.v01.y, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[5] = fminf(fmaxf(b.v2w0.x, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[6] = b.q01.x; s[7] = b.q01.y; s[8] = b.q23.x; s[9] = b.q23.y;
    s[10] = fminf(fmaxf(b.v2w0.y, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[11] = fminf(fmaxf(b.w12.x, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[12] = fminf(fmaxf(b.w12.y, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[13] = b.r01.x; s[14] = b.r01.y; s[15] = b.r23.x; s[16] = b.r23.y;
}

// The packed RK4 as a LOOP over its four stages: the same operations in the same order as
// advance_body_fast_packed (stage weights 1, 2, 2, 1 enter as fma(w, k, acc): fma(1, k, 0) = k and
// fma(1, k, acc) = acc + k are the unrolled form's `acc = k` and `acc + k` bit for bit), in a third of the
// code.  For kernels whose hot path does not fit the 32 KB instruction cache level (the swarm step: 27.5 KB hot,
// 19 % of its stall samples were instruction fetch) the four unrolled evaluations cost more than the loop's branches.
__device__ __forceinline__ void advance_body_fast_loop(float s[17], const DroneParams &p, const float act[4]) {
    BodyP b, tmp;
    b.pos.x = s[0]; b.pos.y = s[1]; b.pos.z = s[2];
    b.v01 = pk(s[3], s[4]); b.v2w0 = pk(s[5], s[10]); b.w12 = pk(s[11], s[12]);
    b.q01 = pk(s[6], s[7]); b.q23 = pk(s[8], s[9]);
    b.r01 = pk(s[13], s[14]); b.r23 = pk(s[15], s[16]);
    FastConsts c;
    c.inv_mass = approx_rcp(p.mass);
    c.inv_ixx = approx_rcp(p.ixx);
    c.inv_iyy = approx_rcp(p.iyy);
    c.inv_izz = approx_rcp(p.izz);
    c.inv_kmot = approx_rcp(p.kmot);
    c.d_yz = p.iyy - p.izz; c.d_zx = p.izz - p.ixx; c.d_xy = p.ixx - p.iyy;
    const float half_mrpm = 0.5f * p.mrpm;
    const float half_mrpm_k = half_mrpm * c.inv_kmot;
#pragma unroll
    for (int m = 0; m < 4; m++) c.want_k[m] = fmaf(act[m], half_mrpm_k, half_mrpm_k);
    const float2 h6 = pk1(B2D_DT / 6.0f), zero = pk1(0.0f);
    RateP k, acc;
    acc.v01 = acc.v2w0 = acc.w12 = acc.q01 = acc.q23 = acc.r01 = acc.r23 = zero;
    V3<float> accp;
    accp.x = accp.y = accp.z = 0.0f;
    tmp = b;
#pragma unroll 1
    for (int st = 0; st < 4; st++) {
        rates_fast_p(tmp, p, c, k);
        const float wgt = (st == 0 || st == 3) ? 1.0f : 2.0f;
        const float2 w2 = pk1(wgt);
        const float vx = tmp.v01.x, vy = tmp.v01.y, vz = tmp.v2w0.x;
        accp.x = fmaf(wgt, vx, accp.x); accp.y = fmaf(wgt, vy, accp.y); accp.z = fmaf(wgt, vz, accp.z);
        acc.v01 = __ffma2_rn(w2, k.v01, acc.v01); acc.v2w0 = __ffma2_rn(w2, k.v2w0, acc.v2w0); acc.w12 = __ffma2_rn(w2, k.w12, acc.w12);
        acc.q01 = __ffma2_rn(w2, k.q01, acc.q01); acc.q23 = __ffma2_rn(w2, k.q23, acc.q23);
        acc.r01 = __ffma2_rn(w2, k.r01, acc.r01); acc.r23 = __ffma2_rn(w2, k.r23, acc.r23);
        if (st < 3) probe_fast_p(b, vx, vy, vz, k, st == 2 ? B2D_DT : 0.5f * B2D_DT, tmp);
    }
    const float h6s = B2D_DT / 6.0f;
    b.pos.x = fmaf(accp.x, h6s, b.pos.x);
    b.pos.y = fmaf(accp.y, h6s, b.pos.y);
    b.pos.z = fmaf(accp.z, h6s, b.pos.z);
#define B2D_FINL(f) b.f = __ffma2_rn(acc.f, h6, b.f)
    B2D_FINL(v01); B2D_FINL(v2w0); B2D_FINL(w12); B2D_FINL(q01); B2D_FINL(q23); B2D_FINL(r01); B2D_FINL(r23);
#undef B2D_FINL
    qnormalize_fast_p(b.q01, b.q23);
    s[0] = b.pos.x; s[1] = b.pos.y; s[2] = b.pos.z;
    s[3] = fminf(fmaxf(b.v01.x, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[4] = fminf(fmaxf(b.v01.y, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[5] = fminf(fmaxf(b.v2w0.x, -B2D_MAX_VEL), B2D_MAX_VEL);
    s[6] = b.q01.x; s[7] = b.q01.y; s[8] = b.q23.x; s[9] = b.q23.y;
    s[10] = fminf(fmaxf(b.v2w0.y, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[11] = fminf(fmaxf(b.w12.x, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[12] = fminf(fmaxf(b.w12.y, -B2D_MAX_OMEGA), B2D_MAX_OMEGA);
    s[13] = b.r01.x; s[14] = b.r01.y; s[15] = b.r23.x; s[16] = b.r23.y;
}

// LOOP: the fast form with its four stages as a loop (bit-identical results; for kernels short of instruction cache)
template <bool STRICT, bool LOOP = false>
__device__ __forceinline__ void advance_body(float s[17], const DroneParams &p, const float act[4]) {
    if constexpr (STRICT) advance_body_strict(s, p, act);
    else if constexpr (LOOP) advance_body_fast_loop(s, p, act);
#if B2D_PACKED_RK4
    else advance_body_fast_packed(s, p, act);
#else
    else advance_body_fast(s, p, act);
#endif
}

// ------------------------------------------------------------------ gate crossing
// DR/dronelib.h:462-489.  Returns +1 (clean pass in the normal's direction), edge_value
// (rim hit: -1.0f in the race copy, -0.0f in the swarm copy) or 0.  T=xf keeps the
// reference's rounding; the decision thresholds 1.5 / 2.5 are exact in binary32.
template <class T>
__device__ __forceinline__ float gate_event(const float before[3], const float after[3], const float ring[6],
                                            float edge_value) {
    T ax = T(before[0]) - T(ring[0]), ay = T(before[1]) - T(ring[1]), az = T(before[2]) - T(ring[2]);
    T bx = T(after[0]) - T(ring[0]), by = T(after[1]) - T(ring[1]), bz = T(after[2]) - T(ring[2]);
    T nx = T(ring[3]), ny = T(ring[4]), nz = T(ring[5]);
    T d0 = ax * nx + ay * ny + az * nz;
    T d1 = bx * nx + by * ny + bz * nz;
    float f0 = fval(d0), f1 = fval(d1);
    bool forward = f0 < 0.0f && f1 > 0.0f;
    bool backward = f0 > 0.0f && f1 < 0.0f;
    if (forward || backward) {
        T dx = T(after[0]) - T(before[0]), dy = T(after[1]) - T(before[1]), dz = T(after[2]) - T(before[2]);
        T t = (-d0) / (nx * dx + ny * dy + nz * dz);
        T hx = T(before[0]) + dx * t - T(ring[0]);
        T hy = T(before[1]) + dy * t - T(ring[1]);
        T hz = T(before[2]) + dz * t - T(ring[2]);
        T r2 = hx * hx + hy * hy + hz * hz;
        float r = fval(tsqrt(r2));
        if (r < 1.5f && forward) return 1.0f;
        if (r < 2.5f) return edge_value;
    }
    return 0.0f;
}

// The fast gate test plus a verdict on how safe its decision is.  `suspect` is raised when a rounding
// difference between the fast and the reference arithmetic could change the outcome: either end of the
// segment within `eps_plane` of the ring plane (the crossing test itself), or a crossing whose hit-point
// radius lies within a band of 1.5 or 2.5 (the pass / rim / miss classification).  The band follows the
// conditioning of the intersection: a position error e moves the hit point by about e |d| / |n.d|
// (d = the step, n = the ring normal), so the band is 2e-3 + 1e-5 |d|_1 / |n.d| -- five times the effect
// of the largest fast-vs-reference position error measured (1 ulp at 10 m = 1e-6 m, profiles/parity_r02.json).
__device__ __forceinline__ float gate_event_guarded(const float before[3], const float after[3], const float ring[6],
                                                    float edge_value, float eps_plane, bool &suspect) {
    const float ax = before[0] - ring[0], ay = before[1] - ring[1], az = before[2] - ring[2];
    const float bx = after[0] - ring[0], by = after[1] - ring[1], bz = after[2] - ring[2];
    const float nx = ring[3], ny = ring[4], nz = ring[5];
    const float f0 = ax * nx + ay * ny + az * nz;
    const float f1 = bx * nx + by * ny + bz * nz;
    suspect = fminf(fabsf(f0), fabsf(f1)) < eps_plane;
    const bool forward = f0 < 0.0f && f1 > 0.0f;
    const bool backward = f0 > 0.0f && f1 < 0.0f;
    if (forward || backward) {
        const float dx = after[0] - before[0], dy = after[1] - before[1], dz = after[2] - before[2];
        const float inv = 1.0f / (nx * dx + ny * dy + nz * dz);
        const float t = (-f0) * inv;
        const float hx = before[0] + dx * t - ring[0];
        const float hy = before[1] + dy * t - ring[1];
        const float hz = before[2] + dz * t - ring[2];
        const float r = sqrtf(hx * hx + hy * hy + hz * hz);
        const float band = 2e-3f + 1e-5f * (fabsf(dx) + fabsf(dy) + fabsf(dz)) * fabsf(inv);
        suspect = suspect || fabsf(r - 1.5f) < band || fabsf(r - 2.5f) < band;
        if (r < 1.5f && forward) return 1.0f;
        if (r < 2.5f) return edge_value;
    }
    return 0.0f;
}

} // namespace b2d
