// episode_gen.cuh -- pieces of the reset path shared by the race and swarm kernels.
//
// Formulas follow the reference's reset code (R = pufferlib/ocean/drone_race):
//   drone_params_from_draws  R/dronelib.h:250-290  init_drone's scale laws + 12 jitter factors
//   ring_from_uniforms       R/dronelib.h:141-183  rndquat + rndring (normal = q . z-axis)
// One IEEE operation at a time (xf), so oracle/drone_oracle.c reproduces the results bit for bit.
#pragma once
#include "b2d_math.cuh"

namespace b2d {

// u[0] -> size in [size_lo, size_hi]; u[1..12] -> the 12 jitter factors in the reference's draw order
__device__ __forceinline__ void drone_params_from_draws(const float u[13], float size_lo, float size_hi, float params[13]) {
    const float jlo = __fsub_rn(1.0f, 0.1f), jhi = __fadd_rn(1.0f, 0.1f);
    xf size = lerp_u(size_lo, size_hi, xf(u[0]));
    xf uj[12];
#pragma unroll
    for (int k = 0; k < 12; k++) uj[k] = (k == 8) ? lerp_u(0.99f, 1.01f, xf(u[1 + k])) : lerp_u(jlo, jhi, xf(u[1 + k]));
    xf arm = size / xf(2.0f);
    xf mass_scale = xf(cube_det(arm.v)) / xf(cube_det(0.1f));
    xf mass = xf(1.0f) * mass_scale * uj[0];
    xf base_iscale = xf(1.0f) * xf(0.1f) * xf(0.1f);
    xf iscale = mass * (arm * arm) / base_iscale;
    xf ixx = xf(0.01f) * iscale * uj[1];
    xf iyy = xf(0.01f) * iscale * uj[2];
    xf izz = xf(0.02f) * iscale * uj[3];
    xf kt_scale = (mass * arm) / (xf(1.0f) * xf(0.1f));
    xf kt = xf(3e-5f) * kt_scale * uj[4];
    xf base_avg = (xf(0.01f) + xf(0.01f) + xf(0.02f)) / xf(3.0f);
    xf avg = (ixx + iyy + izz) / xf(3.0f);
    xf kad = xf(0.2f) * (avg / base_avg) * uj[5];
    xf drag_scale = (arm * arm) / (xf(0.1f) * xf(0.1f));
    xf kd = xf(1e-6f) * drag_scale * uj[6];
    xf bd = xf(0.1f) * drag_scale * uj[7];
    xf grav = xf(9.81f) * uj[8];
    xf mr = xf(750.0f) * (xf(0.1f) / arm) * uj[9];
    xf kmot = xf(0.1f) * uj[10];
    xf jmot = xf(1e-5f) * iscale * uj[11];
    params[0] = mass.v; params[1] = ixx.v; params[2] = iyy.v; params[3] = izz.v;
    params[4] = arm.v; params[5] = kt.v; params[6] = kad.v; params[7] = kd.v;
    params[8] = bd.v; params[9] = grav.v; params[10] = mr.v; params[11] = kmot.v; params[12] = jmot.v;
}

// ring centre uniform in the box (+-bx, +-by, +-bz) and a uniformly random orientation;
// a = (x, y, z, u1) words, b = (u2, u3, -, -) words
__device__ __forceinline__ void ring_from_words(uint4 a, uint4 b, float bx, float by, float bz, float g[6]) {
    xf cx = lerp_u(-bx, bx, unit_from_word(a.x));
    xf cy = lerp_u(-by, by, unit_from_word(a.y));
    xf cz = lerp_u(-bz, bz, unit_from_word(a.z));
    xf u1 = unit_from_word(a.w), u2 = unit_from_word(b.x), u3 = unit_from_word(b.y);
    xf ra = xsqrt(xf(1.0f) - u1), rb = xsqrt(u1);
    float th2 = __double2float_rn(__dmul_rn(6.283185307179586, (double)u2.v));
    float th3 = __double2float_rn(__dmul_rn(6.283185307179586, (double)u3.v));
    float s2, c2, s3, c3;
    sincos_det(th2, s2, c2);
    sincos_det(th3, s3, c3);
    Q4<xf> q;
    q.w = ra * xf(s2); q.x = ra * xf(c2); q.y = rb * xf(s3); q.z = rb * xf(c3);
    V3<xf> zax;
    zax.x = 0.0f; zax.y = 0.0f; zax.z = 1.0f;
    V3<xf> nrm = qrot(q, zax);
    g[0] = cx.v; g[1] = cy.v; g[2] = cz.v; g[3] = nrm.x.v; g[4] = nrm.y.v; g[5] = nrm.z.v;
}

__device__ __forceinline__ float dist3_exact(const float a[3], const float b[3]) {
    xf ex = xf(a[0]) - xf(b[0]), ey = xf(a[1]) - xf(b[1]), ez = xf(a[2]) - xf(b[2]);
    return xsqrt(ex * ex + ey * ey + ez * ez).v;
}

} // namespace b2d
