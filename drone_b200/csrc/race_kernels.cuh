// race_kernels.cuh -- device side of the single-drone ring-race env (sm_100a).
//
// Reference behaviour restated (R = pufferlib/ocean/drone_race):
//   race_step_kernel       R/drone_race.h:156-208 c_step  (+ EB:520-522 vec_step loop)
//   race_observe()         R/drone_race.h:72-125  compute_observations
//   race_generate_episode  R/drone_race.h:127-154 c_reset, R/dronelib.h:141-183,250-300,451-460
//   log accumulation       R/drone_race.h:61-70   add_log, EB:572-591 vec_log
//
// Layout in HBM.  The ten float4 a step reads per env are interleaved per tile of 32 envs,
// float4[tile][10][32] (race_at): lane i touches element i of every 512-byte row, and the 5 KB a
// warp reads for a tile / the 2.5 KB of state it writes back are one contiguous run each.
//   slots 0-4  S   (px,py,pz,vx) (vy,vz,qw,qx) (qy,qz,wx,wy) (wz,r0,r1,r2) (r3,tick,ring word,ep_return)
//                  ring word = ring_idx | (rings are external) << 30
//   slots 5-7  P   (mass,ixx,iyy,izz) (arm,k_thrust,k_ang_damp,k_drag) (b_drag,gravity,max_rpm,k_mot)
//   slot  8    C0  the CURRENT ring (pos.xyz, n.x): no dependent gather per step
//   slot  9    T   (j_mot, live episode number, current ring n.y, n.z)
//   X0 float4[R][ld], X1 float2[R][ld]  rings of episodes that came from outside (reset payload of the
//                  parity hook, put_state); allocated on first use only
// Per env-step the kernel reads 172 B (act 16, S 80, P 52, C 24; + 4 B episode number) and writes
// 201 B (S 80, obs 116, reward 4, terminal 1) = 373 algorithmic bytes.
//
// Auto-reset.  Episode k of env g is a pure function of (seed, g, k) (counter-based Philox), so
// nothing about the next episode is stored anywhere: the lane whose env finishes GENERATES the
// next episode right there, over the finished one (see race_step_kernel).  Rings beyond the
// current one are generated when they are reached.
#pragma once
#include "physics.cuh"

#ifndef B2D_RACE_RK4_LOOP
#define B2D_RACE_RK4_LOOP 0
#endif
#ifndef B2D_RO_RK4_LOOP
#define B2D_RO_RK4_LOOP 0
#endif
namespace b2d {

#ifndef B2D_RACE_BLOCK
#define B2D_RACE_BLOCK 128
#endif
#ifndef B2D_RACE_MIN_CTAS
#define B2D_RACE_MIN_CTAS 4
#endif
#ifndef B2D_EXPERIMENT_SKIP_MATH
#define B2D_EXPERIMENT_SKIP_MATH 0 // measurement aid: the memory pipeline without the RK4 arithmetic
#endif
#ifndef B2D_EXPERIMENT_DOUBLE_MATH
#define B2D_EXPERIMENT_DOUBLE_MATH 0 // measurement aid: RK4 twice per step, same memory traffic
#endif
#ifndef B2D_EXPERIMENT_NO_OBS_STORE
#define B2D_EXPERIMENT_NO_OBS_STORE 0 // measurement aid: no observation store (-31 % traffic), same arithmetic
#endif
#ifndef B2D_EXPERIMENT_NO_RESET
#define B2D_EXPERIMENT_NO_RESET 0 // measurement aid: episodes never end
#endif
#ifndef B2D_EXPERIMENT_NO_GUARD
#define B2D_EXPERIMENT_NO_GUARD 0 // measurement aid: the fast step without its near-threshold guard
#endif
#ifndef B2D_EXPERIMENT_TIMING
#define B2D_EXPERIMENT_TIMING 0 // measurement aid: clock64 sums per phase, printed by b2d_vec_close
#endif
#if B2D_EXPERIMENT_TIMING
#define B2D_TICK(var) const long long var = clock64()
#else
#define B2D_TICK(var)
#endif
constexpr int RACE_BLOCK = B2D_RACE_BLOCK;
constexpr int RACE_WARPS = RACE_BLOCK / 32;
constexpr int RACE_MIN_CTAS = B2D_RACE_MIN_CTAS;
constexpr int RACE_LD_ALIGN = 256; // row padding of the SoA arrays
constexpr int RACE_OBS = 29;
constexpr int RESET_MAX_ATTEMPTS = 16;

// integer episode-statistics accumulators (all race Log fields are integer valued)
enum { ACC_N = 0, ACC_RETURN, ACC_LENGTH, ACC_RINGS, ACC_OOB, ACC_COLLISION, ACC_TIMEOUT, ACC_SPARE, ACC_COUNT };

// Device-side control block.  Nothing in the step kernel reads back the result of a global
// atomic: on B200 a returning ATOMG issued by a busy SM was measured to take 5-10 us (longer
// than a whole tile of work), so every global update here is a fire-and-forget reduction.
struct Ctl {
    unsigned int ctas_done; // race: vec_steps completed since the last vec_reset (swarm: step CTAs of tile 0, grid = 1)
    unsigned int grid;      // step CTAs per launch (fixed for the life of the handle)
    unsigned long long guard_replays; // env-steps the fast kernels re-did in reference arithmetic (race_strict_replay)
    long long acc[ACC_COUNT];
    double facc[8];          // float-valued sums (swarm)
    unsigned long long dbg[12]; // B2D_EXPERIMENT_TIMING
};

struct RaceDev {
    int n, ld, max_rings, max_moves;
    float4 *S;           // the hot state, 10 float4 per env: S0..S4, P0..P2, C0, T (see race_at)
    float4 *X0;          // external rings [R][ld] (injected episodes / put_state); allocated on first use
    float2 *X1;
    unsigned int *chain; // [grid] sequence number of the last launch CTA c completed
    long long *cta_score; // [grid] sum of score over the episodes CTA c saw end in the LAST step (R/drone_race.h:160)
    int tile_begin, tile_end; // tiles [begin, end) stepped by this launch (host-buffer steps are issued in chunks)
    int max_grid;        // the largest grid any launch of this handle uses (size of cta_score / chain)
    int count_step;      // 1: this launch completes a vec_step (the last chunk): bump the step counter
    int score_add;       // 1: add to cta_score instead of overwriting it (chunks after the first)
    uint32_t seq;        // sequence number of this launch (host counter, +1 per launch)
    int chain_wait;      // 1: launched programmatically dependent on launch seq-1 of this kernel: CTA c
                         //    starts as soon as CTA c of that launch is done (see b2d_vec_step_tape)
    const float *act_in; // [n][4] actions read this step
    float *act_out;      // [n][4] clamped actions written back, or nullptr
    float *obs;          // [n][29]
    float *rew;          // [n]
    unsigned char *term; // [n]
    Ctl *ctl;
    const float *payload; // [n][33+6R] next-episode blobs (inject mode)
    float4 *bank;         // [n][RACE_BANK_SLOTS][6] prepared episodes for the rollout kernel (race_bank_fill_kernel); or nullptr
    uint32_t key0, key1, env_id_base;
    int reset_mode; // b2d_reset_mode
#if B2D_EXPERIMENT_TIMING
    unsigned long long *trace; // [2][max_grid][4] (smid, CTA entry ns, first tile ns, CTA done ns) of the last two launches (slot = seq & 1)
#endif
};

// ---------------------------------------------------------------- hot-state addressing
// The ten float4 a step reads per env (S0..S4, P0..P2, C0, T = slots 0..9).
//   tiled (default): float4[tile][10][32] -- the 5 KB a warp reads for a tile, and the 2.5 KB of
//     state it writes back, are ONE contiguous run in HBM, whatever the other warps are doing
//   planar (B2D_RACE_TILED=0): ten float4[ld] planes -- a warp touches ten 512-byte pieces per tile
#ifndef B2D_RACE_TILED
#define B2D_RACE_TILED 1
#endif
constexpr int RACE_HOT_SLOTS = 10;
enum { SLOT_S = 0, SLOT_P = 5, SLOT_C0 = 8, SLOT_T = 9 };
// address of slot 0 of env i; slot s is race_slot_stride(d) float4 further (a compile-time constant when
// tiled, so one 64-bit address computation serves all ten slots of an env)
__device__ __forceinline__ float4 *race_hot(const RaceDev &d, int i) {
#if B2D_RACE_TILED
    return d.S + ((size_t)(i >> 5) * (RACE_HOT_SLOTS * 32) + (size_t)(i & 31));
#else
    return d.S + i;
#endif
}
__device__ __forceinline__ size_t race_slot_stride(const RaceDev &d) {
#if B2D_RACE_TILED
    return 32;
#else
    return (size_t)d.ld;
#endif
}
__device__ __forceinline__ float4 *race_at(const RaceDev &d, int slot, int i) {
    return race_hot(d, i) + (size_t)slot * race_slot_stride(d);
}

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
        const float inv = approx_rcp(mrpm);
#pragma unroll
        for (int m = 0; m < 4; m++) row[25 + m] = s[13 + m] * inv;
    }
}

// ---------------------------------------------------------------- state <-> SoA
// ring word = ring_idx | RING_EXTERNAL when the episode's rings live in X0/X1 (injected / put_state)
constexpr int RING_EXTERNAL = 1 << 30;
constexpr int RING_INDEX_MASK = RING_EXTERNAL - 1;

__device__ __forceinline__ void race_store_state(const RaceDev &d, int i, const float s[17], int tick, int ring_word,
                                                 float ep_ret) {
    float4 *hot = race_hot(d, i);
    const size_t st = race_slot_stride(d);
    hot[(SLOT_S + 0) * st] = make_float4(s[0], s[1], s[2], s[3]);
    hot[(SLOT_S + 1) * st] = make_float4(s[4], s[5], s[6], s[7]);
    hot[(SLOT_S + 2) * st] = make_float4(s[8], s[9], s[10], s[11]);
    hot[(SLOT_S + 3) * st] = make_float4(s[12], s[13], s[14], s[15]);
    hot[(SLOT_S + 4) * st] = make_float4(s[16], __int_as_float(tick), __int_as_float(ring_word), ep_ret);
}

__device__ __forceinline__ void race_load_external_ring(const RaceDev &d, int i, int r, float ring[6]) {
    float4 a = __ldcg(&d.X0[(size_t)r * d.ld + i]);
    float2 b = __ldcg(&d.X1[(size_t)r * d.ld + i]);
    ring[0] = a.x; ring[1] = a.y; ring[2] = a.z; ring[3] = a.w; ring[4] = b.x; ring[5] = b.y;
}

__device__ __forceinline__ void race_store_current_ring(const RaceDev &d, int i, const float ring[6]) {
    *race_at(d, SLOT_C0, i) = make_float4(ring[0], ring[1], ring[2], ring[3]);
    reinterpret_cast<float2 *>(race_at(d, SLOT_T, i))[1] = make_float2(ring[4], ring[5]);
}

// ---------------------------------------------------------------- episode generator
// Counter-based reset stream (DESIGN.md "reset stream"): Philox4x32-10 with
// key=(seed lo, seed hi) and counter=(global env id, episode number, item, attempt);
// item 2r / 2r+1 = ring r (x,y,z,u1 / u2,u3), 0x1000+k = size and the 12 jitter
// factors, 0x2000 = spawn position.  The k-th episode of an env is therefore a pure
// function of (seed, env id, k), whenever and wherever it gets generated.  Same
// distributions and formulas as the reference's c_reset (R/drone_race.h:127-151,
// R/dronelib.h:141-183,250-300,451-460); arithmetic is one IEEE op at a time so the CPU
// oracle (oracle/drone_oracle.c:race_fresh_episode) reproduces it bit for bit.
//
// Rings are generated LAZILY: ring r depends only on its own counters and on the final
// position of ring r-1 (the reference redraws a ring until it is 2*radius away from its
// predecessor, R/dronelib.h:451-460), so an episode start needs ring 0 only and a ring pass
// generates ring r+1 from the ring just passed.  The reference's up-front loop over
// max_rings rings would be 7/8 of the reset cost for rings that are almost never reached.
__device__ __forceinline__ void race_generate_ring(const RaceDev &d, uint32_t env, uint32_t episode, int r,
                                                   const float prev[3], float g[6]) {
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
        xf ex = cx - xf(prev[0]), ey = cy - xf(prev[1]), ez = cz - xf(prev[2]);
        xf dist = xsqrt(ex * ex + ey * ey + ez * ez);
        if (!(dist.v < 4.0f)) break;
    }
}

// ring 0, the 13 drone parameters and the spawn position of episode `episode` of env i
__device__ __forceinline__ void race_generate_episode(const RaceDev &d, int i, uint32_t episode, float params[13],
                                                   float spawn[3], float ring0[6]) {
    const uint32_t env = d.env_id_base + (uint32_t)i;
    const float origin[3] = {0.0f, 0.0f, 0.0f};
    race_generate_ring(d, env, episode, 0, origin, ring0);
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

// the ring that follows `ring` (index r-1, just passed) in episode `episode` of env i
__device__ __noinline__ void race_next_ring(const RaceDev &d, int i, uint32_t episode, int r, float ring[6]) {
    const float prev[3] = {ring[0], ring[1], ring[2]};
    race_generate_ring(d, d.env_id_base + (uint32_t)i, episode, r, prev, ring);
}

__device__ __forceinline__ void race_store_params(const RaceDev &d, int i, const float p[13], uint32_t episode) {
    *race_at(d, SLOT_P + 0, i) = make_float4(p[0], p[1], p[2], p[3]);
    *race_at(d, SLOT_P + 1, i) = make_float4(p[4], p[5], p[6], p[7]);
    *race_at(d, SLOT_P + 2, i) = make_float4(p[8], p[9], p[10], p[11]);
    reinterpret_cast<float2 *>(race_at(d, SLOT_T, i))[0] = make_float2(p[12], __uint_as_float(episode));
}

// Generate episode `episode` in place as the LIVE episode of env i (params into P, fresh state,
// current ring, observation row): the generation passes of the step kernel and vec_reset.
// ROW_SHARED: `obs_row` is the env's row of the warp's shared-memory observation tile (step kernel),
// else its row of d.obs in global memory (vec_reset).
template <bool STRICT, bool ROW_SHARED = false>
__device__ __noinline__ void race_begin_generated(const RaceDev &d, int i, uint32_t episode, float *obs_row) {
    float p[13], spawn[3], ring0[6], s[17], o[RACE_OBS];
    race_generate_episode(d, i, episode, p, spawn, ring0);
#pragma unroll
    for (int k = 0; k < 17; k++) s[k] = 0.0f;
    s[6] = 1.0f;
    s[0] = spawn[0]; s[1] = spawn[1]; s[2] = spawn[2];
    race_store_params(d, i, p, episode);
    race_store_state(d, i, s, 0, 0, 0.0f);
    race_store_current_ring(d, i, ring0);
    race_observe<STRICT>(s, p[10], ring0, o);
#pragma unroll
    for (int k = 0; k < RACE_OBS; k++) {
        if constexpr (ROW_SHARED) obs_row[k] = o[k];
        else __stcg(obs_row + k, o[k]); // plain global stores, no generic-address decode
    }
}

// Parity hook (B2D_RESET_INJECT): the next episode is the oracle's post-reset state from the
// payload; its rings are kept in the external ring arrays X0/X1.
template <bool STRICT>
__device__ __noinline__ void race_inject_episode(const RaceDev &d, int i, uint32_t episode, float *obs_row) {
    const float *b = d.payload + (size_t)i * (33 + 6 * d.max_rings);
    float s[17];
#pragma unroll
    for (int k = 0; k < 17; k++) s[k] = b[k];
    race_store_params(d, i, b + 17, episode);
    const int tick = (int)b[30], ring_idx = (int)b[31];
    for (int r = 0; r < d.max_rings; r++) {
        const float *g = b + 33 + 6 * r;
        d.X0[(size_t)r * d.ld + i] = make_float4(g[0], g[1], g[2], g[3]);
        d.X1[(size_t)r * d.ld + i] = make_float2(g[4], g[5]);
    }
    const float *g = b + 33 + 6 * (ring_idx < d.max_rings ? ring_idx : 0);
    race_store_state(d, i, s, tick, ring_idx | RING_EXTERNAL, b[32]);
    race_store_current_ring(d, i, g);
    race_observe<STRICT>(s, b[27], g, obs_row);
}

// ---------------------------------------------------------------- episode bank (rollout kernel)
// Episode k of env g is a pure function of (seed, g, k), so it can be generated at any time.  The one-kernel
// rollout (rollout_kernels.cuh) steps a block of 128 envs in lock-step through CTA-wide barriers; generating
// an episode where it is needed (~1,100 instructions on one lane) would put that latency on the critical path
// of all 128 envs at almost every step (P(no reset among 128 envs) = 4 %).  Instead the next RACE_BANK_SLOTS
// episodes of every env are generated up front, all lanes busy, by this kernel (launched before every rollout
// launch; entries that are still valid are skipped), and a reset inside the rollout is six 16-byte loads.
// Entry of episode k: slot k % RACE_BANK_SLOTS of the env, 24 floats = 13 params, spawn position, ring 0
// (position + normal), k itself as the tag.  An env that finishes more than RACE_BANK_SLOTS episodes within
// one launch (0.6 % of the envs at K = 128) falls back to in-place generation for the excess.
constexpr int RACE_BANK_SLOTS = 8;
__global__ void __launch_bounds__(128) race_bank_fill_kernel(const __grid_constant__ RaceDev d) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d.n * RACE_BANK_SLOTS) return;
    const int i = j / RACE_BANK_SLOTS, s = j - i * RACE_BANK_SLOTS;
    const uint32_t live = __float_as_uint(race_at(d, SLOT_T, i)->y);
    const uint32_t episode = live + 1u + (uint32_t)s;
    float4 *b = d.bank + ((size_t)i * RACE_BANK_SLOTS + (episode % RACE_BANK_SLOTS)) * 6;
    if (__float_as_uint(b[5].z) == episode && __float_as_uint(b[5].w) == (d.key0 ^ (d.key1 * 0x9E3779B9u) ^ 0xB2D0u)) return; // still valid
    float p[13], spawn[3], ring0[6];
    race_generate_episode(d, i, episode, p, spawn, ring0);
    b[0] = make_float4(p[0], p[1], p[2], p[3]);
    b[1] = make_float4(p[4], p[5], p[6], p[7]);
    b[2] = make_float4(p[8], p[9], p[10], p[11]);
    b[3] = make_float4(p[12], spawn[0], spawn[1], spawn[2]);
    b[4] = make_float4(ring0[0], ring0[1], ring0[2], ring0[3]);
    b[5] = make_float4(ring0[4], ring0[5], __uint_as_float(episode), __uint_as_float(d.key0 ^ (d.key1 * 0x9E3779B9u) ^ 0xB2D0u));
}

// ---------------------------------------------------------------- near-threshold guard of the fast step
// The fast arithmetic (FMA contraction, approximate reciprocals) moves the drone to within ~1e-5 m of
// where the reference's arithmetic moves it.  That is inside the tolerance for every continuous output,
// but the step also takes DECISIONS on the position -- out of bounds (|coordinate| > 10,
// DR/drone_race.h:165-172), ring plane crossed, hit point inside radius -/+ 0.5 (DR/dronelib.h:462-489) --
// and a decision taken within rounding distance of its threshold could come out differently: a terminal,
// a reward of +-1 or a ring index that differs from the reference's.  So a lane whose fast result lies
// within a guard band of any threshold re-does the step in the reference's arithmetic, from the env's
// pre-step state, which is still in global memory at that point.  The bands (5e-5 m at the walls and the
// ring plane, a conditioning-scaled band around the two radii, see gate_event_guarded) are 50x the worst
// fast-vs-reference position error measured (1 ulp at 10 m, profiles/parity_r02.json), so every integer
// output of the fast kernel equals the strict kernel's; a few env-steps in 10^5 are replayed.
constexpr float RACE_GUARD_WALL = 5e-5f;
constexpr float RACE_GUARD_PLANE = 5e-5f;

// out[0:17] = state after the step, out[17] = out of bounds (0/1), out[18] = gate event; `out` is shared memory
__device__ __noinline__ void race_strict_replay(const RaceDev &d, int i, float4 a4, float *out) {
    const float4 *hot = race_hot(d, i);
    const size_t st = race_slot_stride(d);
    const float4 q0 = __ldcg(hot + 0 * st), q1 = __ldcg(hot + 1 * st), q2 = __ldcg(hot + 2 * st), q3 = __ldcg(hot + 3 * st),
                 q4 = __ldcg(hot + 4 * st);
    const float4 p0 = __ldcg(hot + 5 * st), p1 = __ldcg(hot + 6 * st), p2 = __ldcg(hot + 7 * st);
    const float4 c0 = __ldcg(hot + 8 * st), tl = __ldcg(hot + 9 * st);
    float s[17] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x};
    const DroneParams p = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, tl.x};
    const float ring[6] = {c0.x, c0.y, c0.z, c0.w, tl.z, tl.w};
    const float act[4] = {xclamp(xf(a4.x), -1.0f, 1.0f).v, xclamp(xf(a4.y), -1.0f, 1.0f).v, xclamp(xf(a4.z), -1.0f, 1.0f).v,
                          xclamp(xf(a4.w), -1.0f, 1.0f).v};
    const float before[3] = {s[0], s[1], s[2]};
    advance_body_strict(s, p, act);
    const bool oob = s[0] < -10.0f || s[0] > 10.0f || s[1] < -10.0f || s[1] > 10.0f || s[2] < -10.0f || s[2] > 10.0f;
#pragma unroll
    for (int k = 0; k < 17; k++) out[k] = s[k];
    out[17] = oob ? 1.0f : 0.0f;
    out[18] = oob ? 0.0f : gate_event<xf>(before, s, ring, -1.0f);
}

// ---------------------------------------------------------------- async copy helpers
__device__ __forceinline__ void cp_async16(void *sdst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- bulk async copies (TMA engine, no tensor map)
// A tile's inputs are two contiguous runs in global memory (5,120 B of hot state thanks to the tile-interleaved
// layout, 512 B of actions) and its observation rows one contiguous run of 3,712 B, so one lane moves each
// with a single cp.async.bulk instruction -- completion of the loads on a per-warp mbarrier, of the store
// through the bulk async-group -- instead of 11 LDGSTS per lane in and 7 LDS.128 + 7 STG.128 per lane out.
// MEASURED AND NOT THE DEFAULT (profiles/README.md, r02): same box, 4,000 steps at 1 M envs, bit-exact either way:
// bulk 69.39 / 69.34 us per step, per-lane cp.async 68.74 / 69.00 us.  The ~30 issue slots per lane-tile it saves
// are not what limits a kernel that waits on HBM; the extra mbarrier wait, proxy fences and the second
// observation buffer cost as much.  -DB2D_RACE_BULK=1 builds it (needs the tile-interleaved layout).
#ifndef B2D_RACE_BULK
#define B2D_RACE_BULK 0
#endif
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "B2D_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra B2D_DONE;\n\t"
        "bra B2D_WAIT;\n\t"
        "B2D_DONE:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
// generic-proxy accesses to shared memory ordered against the async proxy (bulk copies) and back
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all but the newest `N` bulk stores of this thread have finished READING shared memory
template <int N> __device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// per-warp shared memory:
//   stage  11 float4 per lane: inputs of the NEXT tile, in flight while the current tile computes
//   obs    the 32x29 observation tile of the current tile (staging for the coalesced store); two of them when
//          the tile leaves by bulk store, so that a tile's rows can be written while the previous store drains
constexpr int RACE_STAGE_SLOTS = 11; // act, S0..S4, P0..P2, C0, T
constexpr int RACE_STAGE_BYTES = RACE_STAGE_SLOTS * 32 * 16;
constexpr int RACE_TILE_BYTES = 32 * RACE_OBS * 4;
constexpr int RACE_OBS_BUFFERS = B2D_RACE_BULK ? 2 : 1;
constexpr int RACE_WARP_SMEM = RACE_STAGE_BYTES + RACE_OBS_BUFFERS * RACE_TILE_BYTES;
constexpr int RACE_SMEM_BYTES = RACE_WARPS * RACE_WARP_SMEM;

__device__ __forceinline__ void race_prefetch_tile(const RaceDev &d, float4 *stage, int lane, int i) {
    cp_async16(&stage[0 * 32 + lane], reinterpret_cast<const float4 *>(d.act_in) + i);
    const float4 *hot = race_hot(d, i);
    const size_t st = race_slot_stride(d);
#pragma unroll
    for (int k = 0; k < RACE_HOT_SLOTS; k++) cp_async16(&stage[(1 + k) * 32 + lane], hot + k * st); // stage order = slot order
}
#if B2D_RACE_BULK
// lane 0: the tile's actions (rows * 16 B) and hot-state block (5,120 B) land in `stage`, completion on `bar`
__device__ __forceinline__ void race_prefetch_tile_bulk(const RaceDev &d, float4 *stage, int tile, unsigned long long *bar) {
    const uint32_t act_bytes = (uint32_t)min(32, d.n - tile * 32) * 16u;
    mbar_expect_tx(bar, act_bytes + RACE_HOT_SLOTS * 32 * 16);
    bulk_load(stage, reinterpret_cast<const float4 *>(d.act_in) + (size_t)tile * 32, act_bytes, bar);
    bulk_load(stage + 32, d.S + (size_t)tile * (RACE_HOT_SLOTS * 32), RACE_HOT_SLOTS * 32 * 16, bar);
}
#endif

// ---------------------------------------------------------------- the step kernel
// ONE launch per vec_step.  Persistent grid (RACE_MIN_CTAS resident CTAs per SM), RACE_WARPS
// warps each; a warp owns one tile of 32 envs at a time (one env per lane).
//
// Tile order.  CTA c owns tiles c, c+G, c+2G, ... (G = grid size; the same CTA owns the same
// envs in every launch) and its warps draw them through a SHARED-MEMORY ticket, so a warp that
// spent time generating an episode simply takes fewer tiles.  Across the grid the warps sweep
// every array as one contiguous frontier, which is what DRAM wants (sharded dynamic claims ran
// 4% slower with 16 frontiers and 50% slower with 256).
//
// The inputs of the warp's next tile stream into shared memory (cp.async) while the current
// tile computes (~900 FP32 instructions per lane): no register cost, no dependent-load stall.
// Observation rows ([N,29] row-major, 116 B: not a multiple of 16) are staged per warp in
// shared memory and leave as lane-consecutive float4 stores, 512 B per instruction.
//
// Auto-reset touches no memory but the env's own state.  A lane whose env finished generates the
// env's next episode on the spot -- Philox draws, the reference's scale laws and rejection loops
// (race_generate_episode, ~1,100 instructions) -- writes parameters, spawn state and ring 0 over
// the finished episode and puts the first observation into its row of the warp's tile, which
// leaves with the other 31 rows.  About 2.5 % of the envs finish per step, so more than half of
// all tiles pay for one such (divergent) generation; that is affordable because the step is
// memory-bound with arithmetic to spare (running the RK4 twice per step costs 1 us per 1 M envs)
// and 16 warps per SM keep the loads and stores flowing meanwhile.
// How this kernel got here, all measured at 1 M envs (profiles/README.md):
//   * a PREPARED next episode per env (96 B in six arrays, adopted by scattered 16-byte gathers,
//     restocked through a per-CTA shared-memory ring with carry-over lists between launches):
//     78.7 us.  With the reset path compiled out (B2D_EXPERIMENT_NO_RESET) the step streams at
//     62 us; the slot traffic -- 2.5 % of the envs, but every access a lone sector in its own DRAM
//     page -- and the adoption latency at the end of each launch cost the rest;
//   * per-warp LISTS of finished envs, generated in place 32 at a time and after the warp's last
//     tile: 74.7 us, of which ~5 us was the generation pass every warp ran at the end of a launch;
//   * generation in the tile (this version): 68.7 us, and no list, queue, pass or tail to reason
//     about -- every launch is self-contained.
//
// No global atomic in this kernel returns a value (see Ctl).
template <bool STRICT>
__global__ void __launch_bounds__(RACE_BLOCK, RACE_MIN_CTAS) race_step_kernel(const __grid_constant__ RaceDev d) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ int s_done;
    __shared__ int s_acc[8];
    __shared__ unsigned int s_ticket;             // tile tickets handed out so far in this CTA

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int G = gridDim.x;
    const int ntiles = min((d.n + 31) >> 5, d.tile_end);
    const bool inject = d.reset_mode == 1; // B2D_RESET_INJECT (parity hook)
    float4 *stage = reinterpret_cast<float4 *>(s_dyn + warp * RACE_WARP_SMEM);
    float *tile_obs = reinterpret_cast<float *>(s_dyn + warp * RACE_WARP_SMEM + RACE_STAGE_BYTES);
    float *my_row = tile_obs + lane * RACE_OBS;
#if B2D_RACE_BULK
    __shared__ __align__(8) unsigned long long s_mbar[RACE_WARPS]; // one per warp: the tile inputs have landed
    uint32_t in_phase = 0;
    int obs_buf = 0;
    if (lane == 0) {
        mbar_init(&s_mbar[warp], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
#endif

    // Launch overlap (b2d_vec_step_tape): the next launch of this kernel may begin while this one
    // drains; its CTA c owns the same envs as this CTA c and waits for exactly this CTA's flag.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#if B2D_EXPERIMENT_TIMING
    unsigned long long tr_entry = 0, tr_go = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_entry));
#endif
    if (d.chain_wait) {
        if (tid == 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(d.chain + blockIdx.x) : "memory");
                if (seen != d.seq - 1u) __nanosleep(64);
            } while (seen != d.seq - 1u);
        }
        __syncthreads();
    }

#if B2D_EXPERIMENT_TIMING
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_go));
#endif
    // the first two tickets of every warp are static, so the first loads leave before any barrier
    // CTA c owns the tiles congruent to c modulo G, whatever range a launch covers
    const int first_tile = d.tile_begin + (int)((blockIdx.x + G - d.tile_begin % G) % G);
    int tile = warp * G + first_tile;
    int next = (RACE_WARPS + warp) * G + first_tile;
#if B2D_RACE_BULK
    if (tile < ntiles && lane == 0) race_prefetch_tile_bulk(d, stage, tile, &s_mbar[warp]);
#else
    if (tile < ntiles && tile * 32 + lane < d.n) race_prefetch_tile(d, stage, lane, tile * 32 + lane);
    cp_async_commit(); // group: inputs of the first tile
#endif

    if (tid < 8) s_acc[tid] = 0;
    if (tid == 0) {
        s_done = 0;
        s_ticket = 2 * RACE_WARPS;
    }
    __syncthreads();

#if B2D_EXPERIMENT_TIMING
    long long tm_wait = 0, tm_math = 0, tm_store = 0, tm_adopt = 0, tm_iters = 0, tm_inst = 0, tm_refill = 0;
    const long long t_begin = clock64();
#endif

    int claim = 0;            // lane 0: ticket for the tile after `next`

    while (true) {
        if (tile >= ntiles) break;
        const int i = tile * 32 + lane;
        const bool valid = i < d.n;
        B2D_TICK(t0);
#if B2D_RACE_BULK
        mbar_wait(&s_mbar[warp], in_phase); // this tile's inputs have landed
        in_phase ^= 1u;
#else
        cp_async_wait<0>(); // this tile's inputs have landed
#endif
        B2D_TICK(t1);
        const float4 a4 = stage[0 * 32 + lane];
        const float4 q0 = stage[1 * 32 + lane], q1 = stage[2 * 32 + lane], q2 = stage[3 * 32 + lane],
                     q3 = stage[4 * 32 + lane], q4 = stage[5 * 32 + lane];
        const float4 p0 = stage[6 * 32 + lane], p1 = stage[7 * 32 + lane], p2 = stage[8 * 32 + lane];
        const float4 c0 = stage[9 * 32 + lane];
        const float4 tl = stage[10 * 32 + lane]; // (j_mot, episode, ring n.y, ring n.z)
        __syncwarp();
#if B2D_RACE_BULK
        if (lane == 0) {
            if (next < ntiles) {
                fence_proxy_async_smem(); // the warp's reads of the stage (above) before the async proxy rewrites it
                race_prefetch_tile_bulk(d, stage, next, &s_mbar[warp]);
            }
            bulk_store_wait_read<1>(); // the observation buffer of two tiles ago has been read out: it is written below
        }
        tile_obs = reinterpret_cast<float *>(s_dyn + warp * RACE_WARP_SMEM + RACE_STAGE_BYTES + obs_buf * RACE_TILE_BYTES);
        my_row = tile_obs + lane * RACE_OBS;
        obs_buf ^= 1;
        __syncwarp();
#else
        if (next < ntiles && next * 32 + lane < d.n) race_prefetch_tile(d, stage, lane, next * 32 + lane);
        cp_async_commit(); // group: inputs of the next tile
#endif
        if (lane == 0) claim = (int)atomicAdd(&s_ticket, 1u); // shared memory: lands within the tile

        float s[17];
        float ring[6];
        float mrpm = 1.0f, ep_ret = 0.0f;
        int tick = 0, ring_idx = 0, ring_ext = 0, cause = -1;
        const uint32_t episode = __float_as_uint(tl.y);
        if (valid) {
            s[0] = q0.x; s[1] = q0.y; s[2] = q0.z; s[3] = q0.w; s[4] = q1.x; s[5] = q1.y; s[6] = q1.z; s[7] = q1.w;
            s[8] = q2.x; s[9] = q2.y; s[10] = q2.z; s[11] = q2.w; s[12] = q3.x; s[13] = q3.y; s[14] = q3.z; s[15] = q3.w;
            s[16] = q4.x;
            tick = __float_as_int(q4.y) + 1;
            const int ring_word = __float_as_int(q4.z);
            ring_idx = ring_word & RING_INDEX_MASK;
            ring_ext = ring_word & RING_EXTERNAL;
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
#if B2D_EXPERIMENT_SKIP_MATH
            s[0] = fmaf(act[0], 1e-6f, s[0]); s[4] += p.mass * 1e-9f; s[13] += act[3];
#else
            advance_body<STRICT, B2D_RACE_RK4_LOOP>(s, p, act);
#if B2D_EXPERIMENT_DOUBLE_MATH
            {   // measurement aid: the arithmetic twice, same memory traffic (the second result is folded in at 1e-30)
                float s2[17];
#pragma unroll
                for (int k = 0; k < 17; k++) s2[k] = s[k];
                advance_body<STRICT>(s2, p, act);
                s[0] = fmaf(s2[0] + s2[6] + s2[12] + s2[16], 1e-30f, s[0]);
            }
#endif
#endif

            // ---- episode logic: R/drone_race.h:165-203, as selects (the only branches left are the
            // rare ones: a plane crossing inside gate_event, a ring pass that needs the next ring)
            bool oob = s[0] < -10.0f || s[0] > 10.0f || s[1] < -10.0f || s[1] > 10.0f || s[2] < -10.0f || s[2] > 10.0f;
            float gate = 0.0f;
            if constexpr (STRICT) {
                if (!oob) gate = gate_event<xf>(before, s, ring, -1.0f);
            } else {
                // decisions within a guard band of their threshold are re-taken in the reference's arithmetic
                const float wall = fminf(fminf(fabsf(fabsf(s[0]) - 10.0f), fabsf(fabsf(s[1]) - 10.0f)), fabsf(fabsf(s[2]) - 10.0f));
                bool suspect = false;
                if (!oob) gate = gate_event_guarded(before, s, ring, -1.0f, RACE_GUARD_PLANE, suspect);
#if !B2D_EXPERIMENT_SKIP_MATH && !B2D_EXPERIMENT_NO_GUARD
                if (wall < RACE_GUARD_WALL || suspect) {
                    race_strict_replay(d, i, a4, my_row);
#pragma unroll
                    for (int k = 0; k < 17; k++) s[k] = my_row[k];
                    oob = my_row[17] != 0.0f;
                    gate = my_row[18];
                    atomicAdd(&s_acc[ACC_SPARE], 1);
                }
#endif
            }
            const float reward = oob ? -1.0f : gate;
            ep_ret += reward;
            const bool passed = gate > 0.0f;
            ring_idx += passed ? 1 : 0;
            cause = oob ? (int)ACC_OOB
                        : gate < 0.0f ? (int)ACC_COLLISION
                                      : tick == d.max_moves ? (int)ACC_TIMEOUT : ring_idx == d.max_rings ? (int)ACC_SPARE /* course complete */ : -1;
#if B2D_EXPERIMENT_NO_RESET
            cause = -1; // measurement aid: episodes never end (same arithmetic and streaming traffic, no reset path)
#endif
            if (passed && cause < 0) { // the next ring becomes the current one
                if (ring_ext) race_load_external_ring(d, i, ring_idx, ring);
                else race_next_ring(d, i, episode, ring_idx, ring);
                race_store_current_ring(d, i, ring);
            }
            d.rew[i] = reward;
            d.term[i] = cause >= 0 ? 1 : 0;
        }
        B2D_TICK(t2);

        // ---- finished lanes book the episode and start the next one
        const bool finished = cause >= 0;
        if (valid && !finished) {
            race_store_state(d, i, s, tick, ring_idx | ring_ext, ep_ret);
            race_observe<STRICT>(s, mrpm, ring, my_row);
        }
        if (finished) {
            // add_log: R/drone_race.h:61-70 (score == ring_idx at every call site)
            atomicAdd(&s_acc[ACC_N], 1);
            atomicAdd(&s_acc[ACC_RETURN], __float2int_rn(ep_ret));
            atomicAdd(&s_acc[ACC_LENGTH], tick);
            atomicAdd(&s_acc[ACC_RINGS], ring_idx);
            if (cause != ACC_SPARE) atomicAdd(&s_acc[cause], 1);
            // the next episode starts here and now: its first observation goes out with the tile's rows
            if (inject) race_inject_episode<STRICT>(d, i, episode + 1u, my_row);
            else race_begin_generated<STRICT, true>(d, i, episode + 1u, my_row);
        }
        __syncwarp();

        // ---- observations out: the warp's 3,712-byte tile as 232 lane-consecutive float4.
        {
            const int rows = min(32, d.n - tile * 32);
            float *gobs = d.obs + (size_t)tile * 32 * RACE_OBS;
            if (rows == 32) {
#if B2D_EXPERIMENT_NO_OBS_STORE
                if (tile_obs[lane] == 12345.678f) // measurement aid: the step without its 116 B/env observation store
#endif
                {
#if B2D_RACE_BULK
                    // one 3,712-byte run: generic-proxy writes of all lanes -> fence -> one bulk store by lane 0
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) bulk_store(gobs, tile_obs, RACE_TILE_BYTES);
#else
                    const float4 *src = reinterpret_cast<const float4 *>(tile_obs);
                    float4 *dst = reinterpret_cast<float4 *>(gobs);
#pragma unroll
                    for (int k = 0; k < 7; k++) __stcs(&dst[k * 32 + lane], src[k * 32 + lane]);
                    if (lane < 8) __stcs(&dst[224 + lane], src[224 + lane]);
#endif
                }
            } else {
                for (int k = lane; k < rows * RACE_OBS; k += 32) gobs[k] = tile_obs[k];
            }
            __syncwarp();
        }
        B2D_TICK(t3);

#if B2D_EXPERIMENT_TIMING
        tm_wait += t1 - t0; tm_math += t2 - t1; tm_store += t3 - t2; tm_iters += 1;
#endif
        tile = next;
        next = __shfl_sync(0xffffffffu, claim, 0) * G + first_tile;
    }
#if B2D_RACE_BULK
    if (lane == 0) bulk_store_wait_all(); // this warp's observation rows are in global memory before the CTA's completion flag
#else
    cp_async_wait<0>();
#endif
#if B2D_EXPERIMENT_TIMING
    if (lane == 0) {
        atomicAdd(&d.ctl->dbg[0], (unsigned long long)tm_wait); atomicAdd(&d.ctl->dbg[1], (unsigned long long)tm_math);
        atomicAdd(&d.ctl->dbg[2], (unsigned long long)tm_store); atomicAdd(&d.ctl->dbg[3], (unsigned long long)tm_adopt);
        atomicAdd(&d.ctl->dbg[4], (unsigned long long)tm_iters); atomicAdd(&d.ctl->dbg[5], (unsigned long long)tm_refill);
        atomicAdd(&d.ctl->dbg[6], (unsigned long long)(clock64() - t_begin)); atomicAdd(&d.ctl->dbg[7], 1ull);
        atomicAdd(&d.ctl->dbg[8], (unsigned long long)tm_inst);
        atomicMax(&d.ctl->dbg[9], (unsigned long long)(clock64() - t_begin));
    }
#endif

    // ---- CTA epilogue by whichever warp finishes last
    int last = 0;
    __syncwarp();
    __threadfence(); // this warp's state is visible device-wide before the CTA's completion flag can be
    if (lane == 0) last = atomicAdd(&s_done, 1) == RACE_WARPS - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        __threadfence_block();
        if (lane < 7) {
            const int v = s_acc[lane];
            if (v != 0) atomicAdd((unsigned long long *)&d.ctl->acc[lane], (unsigned long long)(long long)v);
            if (lane == 0 && s_acc[ACC_SPARE] != 0) atomicAdd(&d.ctl->guard_replays, (unsigned long long)s_acc[ACC_SPARE]);
            if (lane == ACC_RINGS) d.cta_score[blockIdx.x] = (long long)v + (d.score_add ? d.cta_score[blockIdx.x] : 0ll);
        }
        // launches may use fewer CTAs than the handle's largest grid: the unused score slots read zero
        if (blockIdx.x == 0 && !d.score_add)
            for (int k = (int)gridDim.x + lane; k < d.max_grid; k += 32) d.cta_score[k] = 0;
#if B2D_EXPERIMENT_TIMING
        if (lane == 0) {
            atomicAdd(&d.ctl->dbg[10], (unsigned long long)(clock64() - t_begin)); // CTA busy time
            atomicMax(&d.ctl->dbg[11], (unsigned long long)(clock64() - t_begin));
            unsigned long long tr_done;
            unsigned int smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_done));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long *tr = d.trace + ((size_t)(d.seq & 1u) * d.max_grid + blockIdx.x) * 4;
            tr[0] = smid; tr[1] = tr_entry; tr[2] = tr_go; tr[3] = tr_done;
        }
#endif
        __syncwarp();
        if (lane == 0) {
            if (d.count_step && blockIdx.x == 0) atomicAdd(&d.ctl->ctas_done, 1u); // one count per vec_step; result unused
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(d.chain + blockIdx.x), "r"(d.seq) : "memory");
        }
    }
}

// ---------------------------------------------------------------- vec_reset / observe / blobs
// vec_reset (EB:500-504): episode 0 becomes the live episode.
__global__ void __launch_bounds__(128) race_reset_kernel(const __grid_constant__ RaceDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    if (d.reset_mode == 1 /* B2D_RESET_INJECT */) {
        race_inject_episode<true>(d, i, 0u, d.obs + (size_t)i * RACE_OBS);
        return;
    }
    race_begin_generated<true>(d, i, 0u, d.obs + (size_t)i * RACE_OBS);
}

// (re)initialise the control block: step counter, optionally the statistics
__global__ void race_ctl_reset_kernel(Ctl *ctl, long long *cta_score, unsigned int steps, unsigned int grid, int clear_acc) {
    for (unsigned int k = threadIdx.x; k < grid; k += blockDim.x) cta_score[k] = 0;
    if (threadIdx.x == 0) {
        ctl->grid = grid;
        ctl->ctas_done = steps;
        if (clear_acc) {
            for (int k = 0; k < ACC_COUNT; k++) ctl->acc[k] = 0;
            for (int k = 0; k < 8; k++) ctl->facc[k] = 0.0;
            for (int k = 0; k < 12; k++) ctl->dbg[k] = 0;
            ctl->guard_replays = 0;
        }
    }
}

// snapshot + clear for vec_log: out[0..7] = acc, out[8] = score of the last step
__global__ void race_log_snapshot_kernel(Ctl *ctl, long long *cta_score, long long *out) {
    long long sc = 0;
    for (unsigned int k = threadIdx.x; k < ctl->grid; k += 32) {
        sc += cta_score[k];
        cta_score[k] = 0;
    }
    for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
    if (threadIdx.x == 0) {
        for (int k = 0; k < ACC_COUNT; k++) { out[k] = ctl->acc[k]; ctl->acc[k] = 0; }
        out[ACC_COUNT] = sc;
    }
}

__global__ void __launch_bounds__(128) race_observe_kernel(const RaceDev d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    float4 q0 = *race_at(d, SLOT_S + 0, i), q1 = *race_at(d, SLOT_S + 1, i), q2 = *race_at(d, SLOT_S + 2, i), q3 = *race_at(d, SLOT_S + 3, i), q4 = *race_at(d, SLOT_S + 4, i);
    float s[17] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x};
    float4 c0 = *race_at(d, SLOT_C0, i);
    float4 c1 = *race_at(d, SLOT_T, i);
    float ring[6] = {c0.x, c0.y, c0.z, c0.w, c1.z, c1.w};
    race_observe<true>(s, race_at(d, SLOT_P + 2, i)->z, ring, d.obs + (size_t)i * RACE_OBS);
}

// blob layout: include/b200drone.h b2d_get_state
__global__ void race_pack_kernel(const RaceDev d, const int *ids, int n, float *blobs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = ids ? ids[k] : k;
    float *b = blobs + (size_t)k * (33 + 6 * d.max_rings);
    float4 q0 = *race_at(d, SLOT_S + 0, i), q1 = *race_at(d, SLOT_S + 1, i), q2 = *race_at(d, SLOT_S + 2, i), q3 = *race_at(d, SLOT_S + 3, i), q4 = *race_at(d, SLOT_S + 4, i);
    float4 p0 = *race_at(d, SLOT_P + 0, i), p1 = *race_at(d, SLOT_P + 1, i), p2 = *race_at(d, SLOT_P + 2, i);
    b[0] = q0.x; b[1] = q0.y; b[2] = q0.z; b[3] = q0.w; b[4] = q1.x; b[5] = q1.y; b[6] = q1.z; b[7] = q1.w;
    b[8] = q2.x; b[9] = q2.y; b[10] = q2.z; b[11] = q2.w; b[12] = q3.x; b[13] = q3.y; b[14] = q3.z; b[15] = q3.w;
    b[16] = q4.x;
    b[17] = p0.x; b[18] = p0.y; b[19] = p0.z; b[20] = p0.w; b[21] = p1.x; b[22] = p1.y; b[23] = p1.z; b[24] = p1.w;
    const float4 pj = *race_at(d, SLOT_T, i);
    b[25] = p2.x; b[26] = p2.y; b[27] = p2.z; b[28] = p2.w; b[29] = pj.x;
    const int ring_word = __float_as_int(q4.z);
    b[30] = (float)__float_as_int(q4.y); b[31] = (float)(ring_word & RING_INDEX_MASK); b[32] = q4.w;
    // all rings of the live episode: stored ones, or the lazily generated chain replayed from ring 0
    float ring[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    for (int r = 0; r < d.max_rings; r++) {
        if (ring_word & RING_EXTERNAL) race_load_external_ring(d, i, r, ring);
        else {
            const float prev[3] = {ring[0], ring[1], ring[2]};
            race_generate_ring(d, d.env_id_base + (uint32_t)i, __float_as_uint(pj.y), r, prev, ring);
        }
        for (int c = 0; c < 6; c++) b[33 + 6 * r + c] = ring[c];
    }
}

// put_state: the blob's rings become external rings of the env's live episode (X0/X1 must exist)
__global__ void race_unpack_kernel(const RaceDev d, const int *ids, int n, const float *blobs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = ids ? ids[k] : k;
    const size_t ld = d.ld;
    const float *b = blobs + (size_t)k * (33 + 6 * d.max_rings);
    float s[17];
    for (int c = 0; c < 17; c++) s[c] = b[c];
    const int ring_idx = (int)b[31];
    race_store_state(d, i, s, (int)b[30], ring_idx | RING_EXTERNAL, b[32]);
    race_store_params(d, i, b + 17, __float_as_uint(race_at(d, SLOT_T, i)->y)); // the episode number is kept
    for (int r = 0; r < d.max_rings; r++) {
        const float *g = b + 33 + 6 * r;
        d.X0[(size_t)r * ld + i] = make_float4(g[0], g[1], g[2], g[3]);
        d.X1[(size_t)r * ld + i] = make_float2(g[4], g[5]);
    }
    const float *g = b + 33 + 6 * (ring_idx < d.max_rings ? ring_idx : 0);
    race_store_current_ring(d, i, g);
}

} // namespace b2d
