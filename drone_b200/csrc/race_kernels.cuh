// race_kernels.cuh -- device side of the single-drone ring-race env (sm_100a).
//
// Reference behaviour restated (R = pufferlib/ocean/drone_race):
//   race_step_kernel       R/drone_race.h:156-208 c_step  (+ EB:520-522 vec_step loop)
//   race_observe()         R/drone_race.h:72-125  compute_observations
//   race_generate_episode  R/drone_race.h:127-154 c_reset, R/dronelib.h:141-183,250-300,451-460
//   log accumulation       R/drone_race.h:61-70   add_log, EB:572-591 vec_log
//
// Layout in HBM (ld = num_envs rounded up to 256; lane i touches element i of
// every array, so each warp access is one contiguous 512-byte float4 run):
//   S  float4[5][ld]  (px,py,pz,vx) (vy,vz,qw,qx) (qy,qz,wx,wy) (wz,r0,r1,r2) (r3,tick,ring word,ep_return)
//                     ring word = ring_idx | ring-buffer parity << 30
//   P  float4[3][ld]  (mass,ixx,iyy,izz) (arm,k_thrust,k_ang_damp,k_drag) (b_drag,gravity,max_rpm,k_mot)
//   PJ float [ld]     j_mot
//   C0 float4[ld], C1 float2[ld]   the CURRENT ring (pos.xyz,n.x) (n.y,n.z): no dependent gather per step
//   G0 float4[2][R][ld], G1 float2[2][R][ld]  rings of the live episode (buffer `parity`) and of the
//                     prepared next episode (the other buffer); read on a ring pass / episode start
//   N float4[3][ld], NJ float[ld], NS float4[ld]  prepared next episode: params, j_mot, spawn position
//   EP u32[ld] live episode number, SLOT_EP u32[ld] episode number held by the prepared slot
// Per env-step the kernel reads 172 B (act 16, S 80, P 52, C 24) and writes
// 201 B (S 80, obs 116, reward 4, terminal 1) = 373 algorithmic bytes.
//
// Auto-reset without a reset on the critical path: an env that finishes ADOPTS its
// prepared episode (13 params + spawn + ring 0: a handful of loads/stores) and queues
// itself on a refill list; the next launch carries a few extra CTAs that regenerate the
// consumed slots (Philox + trig, ~3000 dependent instructions each) with every lane
// busy, overlapped with the step CTAs of that launch.
#pragma once
#include "physics.cuh"

namespace b2d {

#ifndef B2D_RACE_BLOCK
#define B2D_RACE_BLOCK 128
#endif
#ifndef B2D_EXPERIMENT_SKIP_MATH
#define B2D_EXPERIMENT_SKIP_MATH 0
#endif
#ifndef B2D_RACE_MIN_CTAS
#define B2D_RACE_MIN_CTAS 3
#endif
constexpr int RACE_BLOCK = B2D_RACE_BLOCK;
constexpr int RACE_MIN_CTAS = B2D_RACE_MIN_CTAS;
constexpr int RACE_LD_ALIGN = 256; // row padding of the SoA arrays
constexpr int RACE_OBS = 29;
constexpr int RESET_MAX_ATTEMPTS = 16;

// integer episode-statistics accumulators (all race Log fields are integer valued)
enum { ACC_N = 0, ACC_RETURN, ACC_LENGTH, ACC_RINGS, ACC_OOB, ACC_COLLISION, ACC_TIMEOUT, ACC_SPARE, ACC_COUNT };

constexpr int QUEUE_TILE_SHARDS = 16;
constexpr int QUEUE_ENV_SHARDS = 64;
struct alignas(256) PaddedCounter {
    unsigned int v;
    unsigned int pad[63];
};

struct Ctl {
    unsigned int epoch;  // vec steps completed since the last vec_reset
    unsigned int ticket; // CTAs finished in the running step
    unsigned int pad0[2];
    long long acc[ACC_COUNT];
    long long score_step[2]; // sum of score over episodes that ended in step (epoch & 1): R/drone_race.h:160
    double facc[8];          // float-valued sums (swarm)
    // Hot counters, one per 256-byte line: same-address atomics serialise in L2 at a few ns each,
    // so 32K claims per launch on ONE word would bound the kernel; sharded they vanish.
    PaddedCounter tile_next[QUEUE_TILE_SHARDS];       // tile scheduler: next unclaimed tile of each shard
    PaddedCounter queue_count[2][QUEUE_ENV_SHARDS];   // finished/refill queue (epoch & 1), sharded by tile % shards
};

struct RaceDev {
    int n, ld, max_rings, max_moves;
    float4 *S;
    float4 *P;
    float *PJ;
    float4 *C0;
    float2 *C1;
    float4 *G0;
    float2 *G1;
    float4 *N;
    float *NJ;
    float4 *NS;
    uint32_t *EP;
    uint32_t *SLOT_EP;
    uint2 *refill; // [2][QUEUE_ENV_SHARDS][queue_cap] (env | ring buffer << 31, tick or episode number)
    int queue_cap; // entries per queue shard: every env of every tile that maps to the shard
    const float *act_in; // [n][4] actions read this step
    float *act_out;      // [n][4] clamped actions written back, or nullptr
    float *obs;          // [n][29]
    float *rew;          // [n]
    unsigned char *term; // [n]
    Ctl *ctl;
    const float *payload; // [n][33+6R] next-episode blobs (inject mode)
    uint32_t key0, key1, env_id_base;
    int reset_mode; // b2d_reset_mode
};

// ---------------------------------------------------------------- observations
// s: 17-float body state; ring: pos(3) normal(3); row: 29 floats (stride 1)
template <bool STRICT>
__device__ __forceinline__ void race_observe(const float s[17], float mrpm, const float ring[6], float *row) {
    if constexpr (STRICT) {
        Q4<xf> q, qi;
        q.w = s[6]; q.x = s[7]; q.y = s[8]; q.z = s[9];
        qi.w = q.w; qi.x = -q.x; qi.y = -q.y; qi.z = -q.z;
        V3<xf> d, nrm, vel, zax;
        d.x = xf(ring[0]) - xf(s[0]); d.y = xf(ring[1]) - xf(s[1]); d.z = xf(ring[2]) - xf(s[2]);
        nrm.x = ring[3]; nrm.y = ring[4]; nrm.z = ring[5];
        vel.x = s[3]; vel.y = s[4]; vel.z = s[5];
        zax.x = 0.0f; zax.y = 0.0f; zax.z = 1.0f;
        V3<xf> to = qrot(qi, d), bn = qrot(qi, nrm), vb = qrot(qi, vel), up = qrot(q, zax);
        float o0 = (to.x / xf(10.0f)).v, o1 = (to.y / xf(10.0f)).v, o2 = (to.z / xf(10.0f)).v;
        row[0] = o0; row[1] = o1; row[2] = o2;
        row[3] = bn.x.v; row[4] = bn.y.v; row[5] = bn.z.v;
        // "next ring" = ring_buffer[ring_idx % max_rings] is the current ring again (R/drone_race.h:77)
        row[6] = o0; row[7] = o1; row[8] = o2;
        row[9] = bn.x.v; row[10] = bn.y.v; row[11] = bn.z.v;
        row[12] = (vb.x / xf(B2D_MAX_VEL)).v; row[13] = (vb.y / xf(B2D_MAX_VEL)).v; row[14] = (vb.z / xf(B2D_MAX_VEL)).v;
        row[15] = (xf(s[10]) / xf(B2D_MAX_OMEGA)).v; row[16] = (xf(s[11]) / xf(B2D_MAX_OMEGA)).v;
        row[17] = (xf(s[12]) / xf(B2D_MAX_OMEGA)).v;
        row[18] = up.x.v; row[19] = up.y.v; row[20] = up.z.v;
        row[21] = s[6]; row[22] = s[7]; row[23] = s[8]; row[24] = s[9];
#pragma unroll
        for (int m = 0; m < 4; m++) row[25 + m] = (xf(s[13 + m]) / xf(mrpm)).v;
    } else {
        const float w = s[6], x = s[7], y = s[8], z = s[9];
        // (unnormalised) rotation matrix of q; world->body is its transpose
        const float ww = w * w, xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
        const float r00 = (ww + xx) - (yy + zz), r11 = (ww - xx) + (yy - zz), r22 = (ww - xx) - (yy - zz);
        const float r01 = 2.0f * (xy - wz), r10 = 2.0f * (xy + wz);
        const float r02 = 2.0f * (xz + wy), r20 = 2.0f * (xz - wy);
        const float r12 = 2.0f * (yz - wx), r21 = 2.0f * (yz + wx);
        const float dx = ring[0] - s[0], dy = ring[1] - s[1], dz = ring[2] - s[2];
        const float o0 = 0.1f * (r00 * dx + r10 * dy + r20 * dz);
        const float o1 = 0.1f * (r01 * dx + r11 * dy + r21 * dz);
        const float o2 = 0.1f * (r02 * dx + r12 * dy + r22 * dz);
        const float n0 = r00 * ring[3] + r10 * ring[4] + r20 * ring[5];
        const float n1 = r01 * ring[3] + r11 * ring[4] + r21 * ring[5];
        const float n2 = r02 * ring[3] + r12 * ring[4] + r22 * ring[5];
        row[0] = o0; row[1] = o1; row[2] = o2; row[3] = n0; row[4] = n1; row[5] = n2;
        row[6] = o0; row[7] = o1; row[8] = o2; row[9] = n0; row[10] = n1; row[11] = n2;
        row[12] = 0.02f * (r00 * s[3] + r10 * s[4] + r20 * s[5]);
        row[13] = 0.02f * (r01 * s[3] + r11 * s[4] + r21 * s[5]);
        row[14] = 0.02f * (r02 * s[3] + r12 * s[4] + r22 * s[5]);
        row[15] = 0.02f * s[10]; row[16] = 0.02f * s[11]; row[17] = 0.02f * s[12];
        row[18] = r02; row[19] = r12; row[20] = r22;
        row[21] = w; row[22] = x; row[23] = y; row[24] = z;
        const float inv = __frcp_rn(mrpm);
#pragma unroll
        for (int m = 0; m < 4; m++) row[25 + m] = s[13 + m] * inv;
    }
}

// ---------------------------------------------------------------- state <-> SoA
__device__ __forceinline__ void race_store_state(const RaceDev &d, int i, const float s[17], int tick, int ring_idx,
                                                 float ep_ret) {
    d.S[0 * (size_t)d.ld + i] = make_float4(s[0], s[1], s[2], s[3]);
    d.S[1 * (size_t)d.ld + i] = make_float4(s[4], s[5], s[6], s[7]);
    d.S[2 * (size_t)d.ld + i] = make_float4(s[8], s[9], s[10], s[11]);
    d.S[3 * (size_t)d.ld + i] = make_float4(s[12], s[13], s[14], s[15]);
    d.S[4 * (size_t)d.ld + i] = make_float4(s[16], __int_as_float(tick), __int_as_float(ring_idx), ep_ret);
}

__device__ __forceinline__ void race_load_ring(const RaceDev &d, int i, int par, int r, float ring[6]) {
    float4 a = d.G0[((size_t)par * d.max_rings + r) * d.ld + i];
    float2 b = d.G1[((size_t)par * d.max_rings + r) * d.ld + i];
    ring[0] = a.x; ring[1] = a.y; ring[2] = a.z; ring[3] = a.w; ring[4] = b.x; ring[5] = b.y;
}

__device__ __forceinline__ void race_store_current_ring(const RaceDev &d, int i, const float ring[6]) {
    d.C0[i] = make_float4(ring[0], ring[1], ring[2], ring[3]);
    d.C1[i] = make_float2(ring[4], ring[5]);
}

// ---------------------------------------------------------------- episode generator
// Counter-based reset stream (DESIGN.md "reset stream"): Philox4x32-10 with
// key=(seed lo, seed hi) and counter=(global env id, episode number, item, attempt);
// item 2r / 2r+1 = ring r (x,y,z,u1 / u2,u3), 0x1000+k = size and the 12 jitter
// factors, 0x2000 = spawn position.  The k-th episode of an env is therefore a pure
// function of (seed, env id, k), whenever it gets generated.  Same distributions and
// formulas as the reference's c_reset (R/drone_race.h:127-151, R/dronelib.h:141-183,
// 250-300, 451-460); arithmetic is one IEEE op at a time so the CPU oracle
// (oracle/drone_oracle.c:race_fresh_episode) reproduces it bit for bit.
// Rings go straight to ring buffer `tb` of env i; params/spawn/ring 0 come back in registers.
__device__ __noinline__ void race_generate_episode(const RaceDev &d, int i, uint32_t episode, int tb,
                                                   float params[13], float spawn[3], float ring0[6]) {
    const uint32_t env = d.env_id_base + (uint32_t)i;
    const size_t gbase = (size_t)tb * d.max_rings * d.ld + i;
    // rings: R/dronelib.h:451-460 (each at least 2*radius from its predecessor)
    float px = 0.0f, py = 0.0f, pz = 0.0f;
    for (int r = 0; r < d.max_rings; r++) {
        float g[6];
        for (uint32_t t = 0; t < RESET_MAX_ATTEMPTS; t++) {
            uint4 a = philox4x32_10(make_uint4(env, episode, 2u * r, t), d.key0, d.key1);
            uint4 b = philox4x32_10(make_uint4(env, episode, 2u * r + 1u, t), d.key0, d.key1);
            xf cx = lerp_u(-6.0f, 6.0f, unit_from_word(a.x));
            xf cy = lerp_u(-6.0f, 6.0f, unit_from_word(a.y));
            xf cz = lerp_u(-6.0f, 6.0f, unit_from_word(a.z));
            xf u1 = unit_from_word(a.w), u2 = unit_from_word(b.x), u3 = unit_from_word(b.y);
            // R/dronelib.h:141-159 rndquat, :177-178 normal = q . z-axis
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
            if (r == 0) break;
            xf ex = cx - xf(px), ey = cy - xf(py), ez = cz - xf(pz);
            xf dist = xsqrt(ex * ex + ey * ey + ez * ez);
            if (!(dist.v < 4.0f)) break;
        }
        px = g[0]; py = g[1]; pz = g[2];
        d.G0[gbase + (size_t)r * d.ld] = make_float4(g[0], g[1], g[2], g[3]);
        d.G1[gbase + (size_t)r * d.ld] = make_float2(g[4], g[5]);
        if (r == 0) {
#pragma unroll
            for (int k = 0; k < 6; k++) ring0[k] = g[k];
        }
    }
    // R/dronelib.h:250-290 init_drone(size ~ U(0.05, 0.8), dr = 0.1)
    float u[16];
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        uint4 w = philox4x32_10(make_uint4(env, episode, 0x1000u + k, 0u), d.key0, d.key1);
        u[4 * k + 0] = unit_from_word(w.x).v; u[4 * k + 1] = unit_from_word(w.y).v;
        u[4 * k + 2] = unit_from_word(w.z).v; u[4 * k + 3] = unit_from_word(w.w).v;
    }
    const float jlo = __fsub_rn(1.0f, 0.1f), jhi = __fadd_rn(1.0f, 0.1f);
    xf size = lerp_u(0.05f, 0.8f, xf(u[0]));
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
    // spawn at least 2*radius from ring 0: R/drone_race.h:143-149
    for (uint32_t t = 0; t < RESET_MAX_ATTEMPTS; t++) {
        uint4 w = philox4x32_10(make_uint4(env, episode, 0x2000u, t), d.key0, d.key1);
        xf cx = lerp_u(-9.0f, 9.0f, unit_from_word(w.x));
        xf cy = lerp_u(-9.0f, 9.0f, unit_from_word(w.y));
        xf cz = lerp_u(-9.0f, 9.0f, unit_from_word(w.z));
        spawn[0] = cx.v; spawn[1] = cy.v; spawn[2] = cz.v;
        xf ex = cx - xf(ring0[0]), ey = cy - xf(ring0[1]), ez = cz - xf(ring0[2]);
        xf dist = xsqrt(ex * ex + ey * ey + ez * ez);
        if (!(dist.v < 4.0f)) break;
    }
}

__device__ __forceinline__ void race_store_params(const RaceDev &d, int i, const float p[13]) {
    d.P[0 * (size_t)d.ld + i] = make_float4(p[0], p[1], p[2], p[3]);
    d.P[1 * (size_t)d.ld + i] = make_float4(p[4], p[5], p[6], p[7]);
    d.P[2 * (size_t)d.ld + i] = make_float4(p[8], p[9], p[10], p[11]);
    d.PJ[i] = p[12];
}

// Generate episode `episode` of env i into the env's PREPARED slot (next params, next spawn,
// ring buffer tb) and publish it by tagging the slot with the episode number.
__device__ __forceinline__ void race_fill_slot(const RaceDev &d, int i, uint32_t episode, int tb) {
    float p[13], spawn[3], ring0[6];
    race_generate_episode(d, i, episode, tb, p, spawn, ring0);
    d.N[0 * (size_t)d.ld + i] = make_float4(p[0], p[1], p[2], p[3]);
    d.N[1 * (size_t)d.ld + i] = make_float4(p[4], p[5], p[6], p[7]);
    d.N[2 * (size_t)d.ld + i] = make_float4(p[8], p[9], p[10], p[11]);
    d.NJ[i] = p[12];
    d.NS[i] = make_float4(spawn[0], spawn[1], spawn[2], 0.0f);
    __threadfence();
    d.SLOT_EP[i] = episode;
}

// Generate episode `episode` in place as the LIVE episode of env i (rings into buffer tb, params
// into P, fresh state, current ring, observation row).  Only used when the prepared slot cannot
// be trusted (see race_begin_episode) and by vec_reset.
template <bool STRICT>
__device__ __noinline__ void race_begin_generated(const RaceDev &d, int i, uint32_t episode, int tb, float *obs_row) {
    float p[13], spawn[3], ring0[6], s[17];
    race_generate_episode(d, i, episode, tb, p, spawn, ring0);
#pragma unroll
    for (int k = 0; k < 17; k++) s[k] = 0.0f;
    s[6] = 1.0f;
    s[0] = spawn[0]; s[1] = spawn[1]; s[2] = spawn[2];
    race_store_params(d, i, p);
    race_store_state(d, i, s, 0, tb << 30, 0.0f);
    race_store_current_ring(d, i, ring0);
    race_observe<STRICT>(s, p[10], ring0, obs_row);
}

// Parity hook (B2D_RESET_INJECT): the next episode is the oracle's post-reset state from the payload.
template <bool STRICT>
__device__ __noinline__ void race_inject_episode(const RaceDev &d, int i, int par, float *obs_row) {
    const float *b = d.payload + (size_t)i * (33 + 6 * d.max_rings);
    float s[17];
#pragma unroll
    for (int k = 0; k < 17; k++) s[k] = b[k];
    race_store_params(d, i, b + 17);
    const int tick = (int)b[30], ring_idx = (int)b[31];
    const size_t gbase = (size_t)par * d.max_rings * d.ld + i;
    for (int r = 0; r < d.max_rings; r++) {
        const float *g = b + 33 + 6 * r;
        d.G0[gbase + (size_t)r * d.ld] = make_float4(g[0], g[1], g[2], g[3]);
        d.G1[gbase + (size_t)r * d.ld] = make_float2(g[4], g[5]);
    }
    const float *g = b + 33 + 6 * (ring_idx < d.max_rings ? ring_idx : 0);
    race_store_state(d, i, s, tick, ring_idx | (par << 30), b[32]);
    race_store_current_ring(d, i, g);
    race_observe<STRICT>(s, b[27], g, obs_row);
}

// ---------------------------------------------------------------- TMA bulk store helpers
__device__ __forceinline__ void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *ssrc, uint32_t bytes) {
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- async copy helpers
__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// all but the N most recent bulk (TMA) store groups have fully completed (writes performed)
template <int N> __device__ __forceinline__ void tma_store_wait_done() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// per-warp shared memory:
//   stage  11 float4 per lane: inputs of the NEXT tile, in flight while the current tile computes
//   obs    the 32x29 observation tile of the current tile (source of the TMA bulk store)
//   adopt  7 float4 per lane: the prepared episode of a lane whose env just finished, in
//          flight while the NEXT tile computes
constexpr int RACE_STAGE_SLOTS = 11; // act, S0..S4, P0..P2, C0, (j_mot | - | C1.x C1.y)
constexpr int RACE_ADOPT_SLOTS = 7;  // N0..N2, spawn, ring0 (pos,n.x), (n.y n.z j_mot -), (EP SLOT_EP - -)
constexpr int RACE_STAGE_BYTES = RACE_STAGE_SLOTS * 32 * 16;
constexpr int RACE_TILE_BYTES = 32 * RACE_OBS * 4;
constexpr int RACE_WARP_SMEM = RACE_STAGE_BYTES + RACE_TILE_BYTES + RACE_ADOPT_SLOTS * 32 * 16;
constexpr int RACE_SMEM_BYTES = (RACE_BLOCK / 32) * RACE_WARP_SMEM;

__device__ __forceinline__ void race_prefetch_tile(const RaceDev &d, float4 *stage, int lane, int i) {
    const size_t ld = d.ld;
    cp_async16(&stage[0 * 32 + lane], reinterpret_cast<const float4 *>(d.act_in) + i);
#pragma unroll
    for (int k = 0; k < 5; k++) cp_async16(&stage[(1 + k) * 32 + lane], &d.S[k * ld + i]);
#pragma unroll
    for (int k = 0; k < 3; k++) cp_async16(&stage[(6 + k) * 32 + lane], &d.P[k * ld + i]);
    cp_async16(&stage[9 * 32 + lane], &d.C0[i]);
    float *tail = reinterpret_cast<float *>(&stage[10 * 32 + lane]);
    cp_async4(tail, &d.PJ[i]);
    cp_async8(tail + 2, &d.C1[i]);
}

// the prepared episode of env i (ring buffer par^1 holds its rings) -> this lane's adopt slots
__device__ __forceinline__ void race_prefetch_slot(const RaceDev &d, float4 *adopt, int lane, int i, int par) {
    const size_t ld = d.ld;
    const size_t g = (size_t)(par ^ 1) * d.max_rings * ld + i;
#pragma unroll
    for (int k = 0; k < 3; k++) cp_async16(&adopt[k * 32 + lane], &d.N[k * ld + i]);
    cp_async16(&adopt[3 * 32 + lane], &d.NS[i]);
    cp_async16(&adopt[4 * 32 + lane], &d.G0[g]);
    float *t5 = reinterpret_cast<float *>(&adopt[5 * 32 + lane]);
    cp_async8(t5, &d.G1[g]);
    cp_async4(t5 + 2, &d.NJ[i]);
    uint32_t *t6 = reinterpret_cast<uint32_t *>(&adopt[6 * 32 + lane]);
    cp_async4(t6, &d.EP[i]);
    cp_async4(t6 + 1, &d.SLOT_EP[i]);
}

// A finished env starts its next episode: ADOPT the prepared slot (already copied into this
// lane's adopt slots): the two ring buffers swap roles, params / spawn / ring 0 are installed,
// the first observation row is written, and the consumed slot is described for the refill
// pass of the next launch.  The slot was restocked by an earlier launch; its episode tag is
// verified anyway and a mismatch (possible only if state was edited from outside mid-flight)
// falls back to generating the episode in place: same function of (seed, env, episode number).
template <bool STRICT>
__device__ __forceinline__ uint2 race_adopt_from_smem(const RaceDev &d, const float4 *adopt, int lane, int i, int par,
                                                      float *obs_row) {
    const size_t ld = d.ld;
    const int npar = par ^ 1;
    const float4 a = adopt[0 * 32 + lane], b = adopt[1 * 32 + lane], c = adopt[2 * 32 + lane];
    const float4 sp = adopt[3 * 32 + lane], r0 = adopt[4 * 32 + lane], t5 = adopt[5 * 32 + lane];
    const uint4 t6 = reinterpret_cast<const uint4 *>(adopt)[6 * 32 + lane];
    const uint32_t want = t6.x + 1u;
    if (t6.y == want) {
        float s[17];
#pragma unroll
        for (int k = 0; k < 17; k++) s[k] = 0.0f;
        s[6] = 1.0f;
        s[0] = sp.x; s[1] = sp.y; s[2] = sp.z;
        const float ring0[6] = {r0.x, r0.y, r0.z, r0.w, t5.x, t5.y};
        d.P[0 * ld + i] = a;
        d.P[1 * ld + i] = b;
        d.P[2 * ld + i] = c;
        d.PJ[i] = t5.z;
        race_store_state(d, i, s, 0, npar << 30, 0.0f);
        race_store_current_ring(d, i, ring0);
        race_observe<STRICT>(s, c.z, ring0, obs_row);
    } else {
        race_begin_generated<STRICT>(d, i, want, npar, obs_row);
    }
    d.EP[i] = want;
    return make_uint2((uint32_t)i | ((uint32_t)par << 31), want + 1u);
}

// ---------------------------------------------------------------- queues
__device__ __forceinline__ uint2 *race_queue(const RaceDev &d, unsigned int which, int shard) {
    return d.refill + ((size_t)which * QUEUE_ENV_SHARDS + shard) * d.queue_cap;
}
// tile shard s owns tiles [s*per, min((s+1)*per, ntiles))
__device__ __forceinline__ int race_tiles_per_shard(int ntiles) { return (ntiles + QUEUE_TILE_SHARDS - 1) / QUEUE_TILE_SHARDS; }

// Blocking claim with stealing (lane 0 only): used once the warp's home shard ran dry.
__device__ __noinline__ int race_claim_tile_slow(Ctl *ctl, int home, int ntiles) {
    const int per = race_tiles_per_shard(ntiles);
    for (int k = 1; k < QUEUE_TILE_SHARDS; k++) {
        const int s = (home + k) % QUEUE_TILE_SHARDS;
        const int end = min((s + 1) * per, ntiles);
        if ((int)*((volatile unsigned int *)&ctl->tile_next[s].v) < end) {
            const int t = (int)atomicAdd(&ctl->tile_next[s].v, 1u);
            if (t < end) return t;
        }
    }
    return ntiles;
}

// ---------------------------------------------------------------- the step kernel
// ONE launch per vec_step.  Persistent grid (one resident set of CTAs per SM), RACE_BLOCK
// threads each; every warp is independent and never waits on another.
//   prologue: the first warps regenerate the prepared slots consumed in the previous step (32
//     slots per warp, every lane busy: Philox + trig, ~3000 dependent instructions),
//     overlapped with the other warps' stepping.
//   main loop: warps pull tiles of 32 envs (one env per lane) from sharded atomic tile
//     counters, so late starters simply take fewer tiles.  Everything with memory latency is
//     software-pipelined one tile deep and costs no registers:
//       * the inputs of the warp's next tile stream into shared memory with cp.async while
//         the current tile computes (~1000 FP32 instructions per lane);
//       * the tile after that is being claimed (atomic in flight);
//       * a lane whose env finished streams the env's prepared next episode into shared
//         memory and installs it one tile later (no dependent-load stall, no second kernel).
//     Each warp stages its 32 observation rows in its own shared-memory tile and ships them
//     as one 3,712-byte TMA bulk store (row-major [N,29] rows are 116 B, not a multiple of 16,
//     so per-lane vector stores cannot be coalesced).  Episode statistics are summed per CTA
//     in shared memory and flushed by whichever warp of the CTA finishes last.
template <bool STRICT>
__global__ void __launch_bounds__(RACE_BLOCK, RACE_MIN_CTAS) race_step_kernel(const __grid_constant__ RaceDev d) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ int s_done;
    __shared__ int s_acc[8];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    if (tid < 8) s_acc[tid] = 0;
    if (tid == 8) s_done = 0;
    __syncthreads();
    float4 *stage = reinterpret_cast<float4 *>(s_dyn + warp * RACE_WARP_SMEM);
    float *tile_obs = reinterpret_cast<float *>(s_dyn + warp * RACE_WARP_SMEM + RACE_STAGE_BYTES);
    float4 *adopt = reinterpret_cast<float4 *>(s_dyn + warp * RACE_WARP_SMEM + RACE_STAGE_BYTES + RACE_TILE_BYTES);
    float *my_row = tile_obs + lane * RACE_OBS;
    const int warps_total = gridDim.x * (RACE_BLOCK / 32);
    const int gw = blockIdx.x * (RACE_BLOCK / 32) + warp;
    const int ntiles = (d.n + 31) >> 5;
    const uint32_t epoch = d.ctl->epoch + 1u;
    const bool inject = d.reset_mode == 1; // B2D_RESET_INJECT (parity hook)

    // claim the first tiles from the warp's home shard; two further claims stay in flight so a
    // claim's round trip to L2 has two whole tiles to complete
    const unsigned int src = (epoch - 1u) & 1u;
    const int home = gw % QUEUE_TILE_SHARDS;
    const int home_end = min((home + 1) * race_tiles_per_shard(ntiles), ntiles);
    int tile = 0, next = 0, claim_a = 0, claim_b = 0; // raw claims (lane 0); resolved when they become `next`
    if (lane == 0) {
        tile = (int)atomicAdd(&d.ctl->tile_next[home].v, 1u);
        next = (int)atomicAdd(&d.ctl->tile_next[home].v, 1u);
        claim_a = (int)atomicAdd(&d.ctl->tile_next[home].v, 1u);
        if (tile >= home_end) tile = race_claim_tile_slow(d.ctl, home, ntiles);
    }
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile < ntiles && tile * 32 + lane < d.n) race_prefetch_tile(d, stage, lane, tile * 32 + lane);
    cp_async_commit(); // group: inputs of the first tile
    cp_async_commit(); // group: (empty) adoption loads "of the tile before the first"

    // ---- prologue: restock the prepared slots consumed during step epoch-1
    {
        const unsigned int stride = (unsigned int)max(warps_total / QUEUE_ENV_SHARDS, 1) * 32u;
        // with fewer warps than shards a warp walks several shards; the warps beyond a whole
        // multiple of the shard count would only repeat chunks, so they skip the prologue
        const bool spare = warps_total >= QUEUE_ENV_SHARDS && gw >= (warps_total / QUEUE_ENV_SHARDS) * QUEUE_ENV_SHARDS;
        for (int sh = gw % QUEUE_ENV_SHARDS; sh < QUEUE_ENV_SHARDS && !spare; sh += warps_total) {
            const unsigned int c = d.ctl->queue_count[src][sh].v;
            const uint2 *l = race_queue(d, src, sh);
            for (unsigned int k = (unsigned int)(gw / QUEUE_ENV_SHARDS) * 32u + lane; k < ((c + 31u) & ~31u); k += stride) {
                if (k < c) {
                    uint2 e = l[k];
                    race_fill_slot(d, (int)(e.x & 0x7fffffffu), e.y, (int)(e.x >> 31));
                }
            }
        }
        __syncwarp();
    }

    // state carried from one tile to the next
    bool store_pending = false;      // a TMA store of this warp's observation tile is in flight
    bool pend_adopt = false;         // this lane's env finished in the previous tile (Philox mode)
    int pend_i = 0, pend_par = 0;
    unsigned int pend_m = 0u, pend_base = 0u; // ballot of pend_adopt, refill-queue slot reserved for them
    int pend_shard = 0;

    while (true) {
        const bool have_tile = tile < ntiles;
        if (!have_tile && pend_m == 0u) break;
        if (have_tile) {
            if (lane == 0 && next >= home_end) next = race_claim_tile_slow(d.ctl, home, ntiles);
            next = __shfl_sync(0xffffffffu, next, 0);
        }
        const int i = tile * 32 + lane;
        const bool valid = have_tile && i < d.n;
        cp_async_wait<1>(); // this tile's inputs have landed (the newest group, adoption loads, may still fly)
        const float4 a4 = stage[0 * 32 + lane];
        const float4 q0 = stage[1 * 32 + lane], q1 = stage[2 * 32 + lane], q2 = stage[3 * 32 + lane],
                     q3 = stage[4 * 32 + lane], q4 = stage[5 * 32 + lane];
        const float4 p0 = stage[6 * 32 + lane], p1 = stage[7 * 32 + lane], p2 = stage[8 * 32 + lane];
        const float4 c0 = stage[9 * 32 + lane];
        const float4 tl = stage[10 * 32 + lane]; // (j_mot, -, C1.x, C1.y)
        // the previous tile's observation store must have read the tile before it is rewritten
        if (store_pending && lane == 0) tma_store_wait_read();
        __syncwarp();
        if (have_tile && next < ntiles && next * 32 + lane < d.n) race_prefetch_tile(d, stage, lane, next * 32 + lane);
        cp_async_commit(); // group: inputs of the next tile
        if (have_tile && lane == 0) claim_b = (int)atomicAdd(&d.ctl->tile_next[home].v, 1u);

        float s[17];
        float ring[6];
        float mrpm = 1.0f, ep_ret = 0.0f;
        int tick = 0, ring_idx = 0, par = 0, cause = -1;
        if (valid) {
            s[0] = q0.x; s[1] = q0.y; s[2] = q0.z; s[3] = q0.w; s[4] = q1.x; s[5] = q1.y; s[6] = q1.z; s[7] = q1.w;
            s[8] = q2.x; s[9] = q2.y; s[10] = q2.z; s[11] = q2.w; s[12] = q3.x; s[13] = q3.y; s[14] = q3.z; s[15] = q3.w;
            s[16] = q4.x;
            tick = __float_as_int(q4.y) + 1;
            const int ring_word = __float_as_int(q4.z);
            ring_idx = ring_word & 0x3fffffff;
            par = (ring_word >> 30) & 1;
            ep_ret = q4.w;
            DroneParams p = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, tl.x};
            mrpm = p2.z;
            ring[0] = c0.x; ring[1] = c0.y; ring[2] = c0.z; ring[3] = c0.w; ring[4] = tl.z; ring[5] = tl.w;

            // clamp the action (written back only on request): R/dronelib.h:437
            float act[4];
            if constexpr (STRICT) {
                act[0] = xclamp(xf(a4.x), -1.0f, 1.0f).v; act[1] = xclamp(xf(a4.y), -1.0f, 1.0f).v;
                act[2] = xclamp(xf(a4.z), -1.0f, 1.0f).v; act[3] = xclamp(xf(a4.w), -1.0f, 1.0f).v;
            } else {
                act[0] = fminf(fmaxf(a4.x, -1.0f), 1.0f); act[1] = fminf(fmaxf(a4.y, -1.0f), 1.0f);
                act[2] = fminf(fmaxf(a4.z, -1.0f), 1.0f); act[3] = fminf(fmaxf(a4.w, -1.0f), 1.0f);
            }
            if (d.act_out) reinterpret_cast<float4 *>(d.act_out)[i] = make_float4(act[0], act[1], act[2], act[3]);

            const float before[3] = {s[0], s[1], s[2]};
            advance_body<STRICT>(s, p, act);

            // ---- episode logic: R/drone_race.h:165-203
            float reward = 0.0f;
            const bool oob = s[0] < -10.0f || s[0] > 10.0f || s[1] < -10.0f || s[1] > 10.0f || s[2] < -10.0f || s[2] > 10.0f;
            if (oob) {
                reward = -1.0f;
                ep_ret -= 1.0f;
                cause = ACC_OOB;
            } else {
                float gate;
                if constexpr (STRICT) gate = gate_event<xf>(before, s, ring, -1.0f);
                else gate = gate_event<float>(before, s, ring, -1.0f);
                reward = gate;
                ep_ret += gate;
                if (gate > 0.0f) ring_idx += 1;
                if (gate < 0.0f) {
                    cause = ACC_COLLISION;
                } else if (tick == d.max_moves) {
                    cause = ACC_TIMEOUT;
                } else if (ring_idx == d.max_rings) {
                    cause = ACC_SPARE; // course complete
                } else if (gate > 0.0f) {
                    race_load_ring(d, i, par, ring_idx, ring);
                    race_store_current_ring(d, i, ring);
                }
            }
            d.rew[i] = reward;
            d.term[i] = cause >= 0 ? 1 : 0;
        }

        // ---- finished lanes book the episode; their next episode is installed one tile later
        const bool finished = cause >= 0;
        const bool adopt_now = finished && !inject;
        const unsigned int m = __ballot_sync(0xffffffffu, adopt_now);
        const int qshard = tile % QUEUE_ENV_SHARDS;
        unsigned int base = 0;
        if (m != 0u && lane == 0) base = atomicAdd(&d.ctl->queue_count[epoch & 1u][qshard].v, (unsigned int)__popc(m));
        if (valid && !finished) {
            race_store_state(d, i, s, tick, ring_idx | (par << 30), ep_ret);
            race_observe<STRICT>(s, mrpm, ring, my_row);
        }
        if (finished) {
            // add_log: R/drone_race.h:61-70 (score == ring_idx at every call site)
            atomicAdd(&s_acc[ACC_N], 1);
            atomicAdd(&s_acc[ACC_RETURN], __float2int_rn(ep_ret));
            atomicAdd(&s_acc[ACC_LENGTH], tick);
            atomicAdd(&s_acc[ACC_RINGS], ring_idx);
            if (cause != ACC_SPARE) atomicAdd(&s_acc[cause], 1);
            if (inject) race_inject_episode<STRICT>(d, i, par, my_row);
        }
        __syncwarp();

        // ---- observations out: one TMA bulk store per full warp tile (rows of lanes that
        // finished hold stale data here; they are rewritten when the episode is installed)
        bool stored_now = false;
        if (have_tile) {
            const int rows = min(32, d.n - tile * 32);
            if (rows == 32) {
                stored_now = true;
                if (lane == 0) {
                    tma_store_fence();
                    tma_store_1d(d.obs + (size_t)tile * 32 * RACE_OBS, tile_obs, RACE_TILE_BYTES);
                }
                store_pending = true;
            } else {
                float *gobs = d.obs + (size_t)tile * 32 * RACE_OBS;
                for (int k = lane; k < rows * RACE_OBS; k += 32) gobs[k] = tile_obs[k];
                __syncwarp();
            }
        }

        // ---- install the next episode of the envs that finished in the PREVIOUS tile
        cp_async_wait<1>(); // their prepared slots have landed (only the next tile's inputs may still fly)
        if (pend_m != 0u) {
            pend_base = __shfl_sync(0xffffffffu, pend_base, 0);
            // their stale rows went out with the previous tile's bulk store: it must be complete
            if (lane == 0) {
                if (stored_now) tma_store_wait_done<1>(); // all but the store issued a moment ago
                else tma_store_wait_done<0>();
            }
            __syncwarp();
            if (pend_adopt) {
                const uint2 e = race_adopt_from_smem<STRICT>(d, adopt, lane, pend_i, pend_par, d.obs + (size_t)pend_i * RACE_OBS);
                race_queue(d, epoch & 1u, pend_shard)[pend_base + __popc(pend_m & ((1u << lane) - 1u))] = e;
            }
            __syncwarp();
        }
        // ---- and start streaming the prepared slots of the envs that finished in THIS tile
        if (adopt_now) race_prefetch_slot(d, adopt, lane, i, par);
        cp_async_commit(); // group: adoption loads of this tile (possibly empty)
        pend_adopt = adopt_now;
        pend_i = i;
        pend_par = par;
        pend_m = m;
        pend_base = base;
        pend_shard = qshard;

        if (have_tile) {
            tile = next;
            next = claim_a; // claimed two tiles ago
            claim_a = claim_b;
        }
    }
    if (store_pending && lane == 0) tma_store_wait_read();

    // ---- CTA epilogue by whichever warp finishes last
    int last = 0;
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        last = atomicAdd(&s_done, 1) == RACE_BLOCK / 32 - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence_block();
        if (lane < 7) {
            int v = s_acc[lane];
            if (v != 0) {
                atomicAdd((unsigned long long *)&d.ctl->acc[lane], (unsigned long long)(long long)v);
                if (lane == ACC_RINGS)
                    atomicAdd((unsigned long long *)&d.ctl->score_step[epoch & 1u], (unsigned long long)(long long)v);
            }
        }
        __syncwarp();
        // ---- every CTA takes a ticket; the last one closes the step.  No device-scope fence is
        // needed: everything a CTA publishes is consumed by the NEXT launch, and the closing
        // writes touch only words no CTA of this launch reads after taking its ticket.
        if (lane == 0) {
            unsigned int t = atomicAdd(&d.ctl->ticket, 1u);
            if (t == gridDim.x - 1) {
                d.ctl->score_step[(epoch + 1u) & 1u] = 0;
                // the queue this launch consumed is the one step epoch+1 appends to
                for (int q = 0; q < QUEUE_ENV_SHARDS; q++) d.ctl->queue_count[(epoch + 1u) & 1u][q].v = 0;
                const int per = race_tiles_per_shard(ntiles);
                for (int q = 0; q < QUEUE_TILE_SHARDS; q++) d.ctl->tile_next[q].v = (unsigned int)(q * per);
                d.ctl->ticket = 0;
                __threadfence();
                d.ctl->epoch = epoch;
            }
        }
    }
}

// ---------------------------------------------------------------- vec_reset / observe / blobs
// vec_reset (EB:500-504): episode 0 becomes the live episode, episode 1 the prepared one.
__global__ void __launch_bounds__(128) race_reset_kernel(const __grid_constant__ RaceDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    if (d.reset_mode == 1 /* B2D_RESET_INJECT */) {
        race_inject_episode<true>(d, i, 0, d.obs + (size_t)i * RACE_OBS);
        return;
    }
    race_begin_generated<true>(d, i, 0u, 0, d.obs + (size_t)i * RACE_OBS);
    d.EP[i] = 0u;
    race_fill_slot(d, i, 1u, 1);
}

// Refill every slot still queued from the last step now (instead of overlapped with the next
// step) -- used before state is edited from outside (put_state), so that no refill is in flight
// while the edited env may finish again.
__global__ void __launch_bounds__(128) race_drain_kernel(const __grid_constant__ RaceDev d) {
    const unsigned int src = d.ctl->epoch & 1u;
    const unsigned int cnt = d.ctl->queue_count[src][blockIdx.y].v;
    const uint2 *list = race_queue(d, src, blockIdx.y);
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += gridDim.x * blockDim.x) {
        uint2 e = list[k];
        if (e.y != 0xffffffffu) race_fill_slot(d, (int)(e.x & 0x7fffffffu), e.y, (int)(e.x >> 31));
    }
}
__global__ void race_drain_done_kernel(Ctl *ctl) {
    if (threadIdx.x < QUEUE_ENV_SHARDS) ctl->queue_count[ctl->epoch & 1u][threadIdx.x].v = 0;
}

__global__ void race_ctl_reset_kernel(Ctl *ctl, unsigned int epoch, int clear_acc, int ntiles) {
    if (threadIdx.x == 0) {
        ctl->epoch = epoch;
        ctl->ticket = 0;
        for (int q = 0; q < QUEUE_ENV_SHARDS; q++) ctl->queue_count[0][q].v = ctl->queue_count[1][q].v = 0;
        for (int q = 0; q < QUEUE_TILE_SHARDS; q++) ctl->tile_next[q].v = (unsigned int)(q * race_tiles_per_shard(ntiles));
        ctl->score_step[0] = ctl->score_step[1] = 0;
        if (clear_acc) {
            for (int k = 0; k < ACC_COUNT; k++) ctl->acc[k] = 0;
            for (int k = 0; k < 8; k++) ctl->facc[k] = 0.0;
        }
    }
}

// snapshot + clear for vec_log: out[0..7] = acc, out[8] = score of the last step
__global__ void race_log_snapshot_kernel(Ctl *ctl, long long *out) {
    if (threadIdx.x == 0) {
        for (int k = 0; k < ACC_COUNT; k++) { out[k] = ctl->acc[k]; ctl->acc[k] = 0; }
        out[ACC_COUNT] = ctl->score_step[ctl->epoch & 1u];
        ctl->score_step[0] = ctl->score_step[1] = 0;
    }
}

__global__ void __launch_bounds__(128) race_observe_kernel(const RaceDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const size_t ld = d.ld;
    float4 q0 = d.S[0 * ld + i], q1 = d.S[1 * ld + i], q2 = d.S[2 * ld + i], q3 = d.S[3 * ld + i], q4 = d.S[4 * ld + i];
    float s[17] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x};
    float4 c0 = d.C0[i];
    float2 c1 = d.C1[i];
    float ring[6] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y};
    race_observe<true>(s, d.P[2 * ld + i].z, ring, d.obs + (size_t)i * RACE_OBS);
}

// blob layout: include/b200drone.h b2d_get_state
__global__ void race_pack_kernel(const RaceDev d, const int *ids, int n, float *blobs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = ids ? ids[k] : k;
    const size_t ld = d.ld;
    float *b = blobs + (size_t)k * (33 + 6 * d.max_rings);
    float4 q0 = d.S[0 * ld + i], q1 = d.S[1 * ld + i], q2 = d.S[2 * ld + i], q3 = d.S[3 * ld + i], q4 = d.S[4 * ld + i];
    float4 p0 = d.P[0 * ld + i], p1 = d.P[1 * ld + i], p2 = d.P[2 * ld + i];
    b[0] = q0.x; b[1] = q0.y; b[2] = q0.z; b[3] = q0.w; b[4] = q1.x; b[5] = q1.y; b[6] = q1.z; b[7] = q1.w;
    b[8] = q2.x; b[9] = q2.y; b[10] = q2.z; b[11] = q2.w; b[12] = q3.x; b[13] = q3.y; b[14] = q3.z; b[15] = q3.w;
    b[16] = q4.x;
    b[17] = p0.x; b[18] = p0.y; b[19] = p0.z; b[20] = p0.w; b[21] = p1.x; b[22] = p1.y; b[23] = p1.z; b[24] = p1.w;
    b[25] = p2.x; b[26] = p2.y; b[27] = p2.z; b[28] = p2.w; b[29] = d.PJ[i];
    const int ring_word = __float_as_int(q4.z);
    b[30] = (float)__float_as_int(q4.y); b[31] = (float)(ring_word & 0x3fffffff); b[32] = q4.w;
    for (int r = 0; r < d.max_rings; r++) {
        float ring[6];
        race_load_ring(d, i, (ring_word >> 30) & 1, r, ring);
        for (int c = 0; c < 6; c++) b[33 + 6 * r + c] = ring[c];
    }
}

__global__ void race_unpack_kernel(const RaceDev d, const int *ids, int n, const float *blobs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = ids ? ids[k] : k;
    const size_t ld = d.ld;
    const float *b = blobs + (size_t)k * (33 + 6 * d.max_rings);
    float s[17];
    for (int c = 0; c < 17; c++) s[c] = b[c];
    const int ring_idx = (int)b[31];
    const int par = (__float_as_int(d.S[4 * ld + i].z) >> 30) & 1; // keep the env's ring-buffer parity
    race_store_state(d, i, s, (int)b[30], ring_idx | (par << 30), b[32]);
    race_store_params(d, i, b + 17);
    for (int r = 0; r < d.max_rings; r++) {
        const float *g = b + 33 + 6 * r;
        d.G0[((size_t)par * d.max_rings + r) * ld + i] = make_float4(g[0], g[1], g[2], g[3]);
        d.G1[((size_t)par * d.max_rings + r) * ld + i] = make_float2(g[4], g[5]);
    }
    const float *g = b + 33 + 6 * (ring_idx < d.max_rings ? ring_idx : 0);
    race_store_current_ring(d, i, g);
}

} // namespace b2d
