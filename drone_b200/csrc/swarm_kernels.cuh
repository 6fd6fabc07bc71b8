// swarm_kernels.cuh -- device side of the multi-drone swarm env (sm_100a).
//
// Reference behaviour restated (R = pufferlib/ocean/drone_swarm):
//   swarm_kernel            R/drone_swarm.h:445-497 c_step, :401-443 c_reset (+ EB:520-522 vec_step loop)
//   sw_reward()             R/drone_swarm.h:335-376 compute_reward, :107-129 nearest_drone
//   sw_observe()            R/drone_swarm.h:131-217 compute_observations
//   targets                 R/drone_swarm.h:219-333 move_target, set_target_*
//   respawn / env reset     R/drone_swarm.h:378-443 reset_agent, c_reset
//   log                     R/drone_swarm.h:91-105  add_log, EB:572-591 vec_log
//
// One thread per drone, an env = A consecutive threads of a 128-thread CTA (128/A envs per CTA),
// so agent rows are contiguous and every state access is a coalesced float4 run.
//
// The reference steps the agents of an env one after the other, so what agent i sees of agent j
// depends on their order: j < i has already moved this tick (and may have been re-spawned after
// leaving the arena), j > i still sits at its previous-tick position.  Nothing else couples the
// agents, so the loop is parallelised by publishing three position arrays per env in shared
// memory -- previous tick (old), moved-and-possibly-respawned (fin), after an env-wide reset
// (rst / pos) -- and selecting per neighbour by index.  Ties in the nearest-neighbour search go
// to the lowest index exactly as the reference's strict `<` scan does.
//
// Layout in HBM, ld = rows rounded up to 256 (row = env * A + agent):
//   S float4[5][ld]  pos, vel, quat, omega, rpm (17 f32) + episode_length + ring_idx + episode_return
//   P float4[3][ld]  12 params;  T float4[ld] (j_mot, respawn count, collisions, score)
//   U float4[ld] (spawn.xyz, last_abs_reward)  V float4[ld] (target_pos.xyz, last_target_reward)
//   W float4[ld] (target_vel.xyz, last_collision_reward)
//   E int4[n] (tick, task, env-reset count, -)   G0 float4[R][n], G1 float2[R][n] rings per env
// Algorithmic bytes per drone-step: reads 228 (act 16, S 80, P+T 64 of which 52 used, U/V/W 48,
// ring 24 shared per env) + writes 293 (S 80, V/W/T 48 of which 36 used, obs 164, rew 4, term 1)
// = 521 B.
#pragma once
#include "episode_gen.cuh"
#include "race_kernels.cuh"

namespace b2d {

#ifndef B2D_SWARM_EXPERIMENT_DOUBLE_MATH
#define B2D_SWARM_EXPERIMENT_DOUBLE_MATH 0
#endif
#ifndef B2D_SWARM_EXPERIMENT_NO_STATS
#define B2D_SWARM_EXPERIMENT_NO_STATS 0
#endif
#ifndef B2D_SWARM_EXPERIMENT_NO_RESPAWN
#define B2D_SWARM_EXPERIMENT_NO_RESPAWN 0
#endif
#ifndef B2D_SW_OVERLAP
#define B2D_SW_OVERLAP 1 // consecutive step launches overlap at their edges (PDL + per-CTA completion flags)
#endif
#ifndef B2D_SW_RS_POOL
#define B2D_SW_RS_POOL 8 // per warp: prepared respawn parameters fetched by cp.async a phase ahead (0: plain loads)
#endif
#ifndef B2D_SW_OBS_BULK
#define B2D_SW_OBS_BULK 1 // observation rows leave shared memory as one asynchronous bulk copy per group (cp.async.bulk)
#endif
#ifndef B2D_SW_RK4_LOOP
#define B2D_SW_RK4_LOOP 1 // RK4 stages as a loop: the hot path shrinks by ~4.5 KB of code
#endif
constexpr int SWARM_BLOCK = 128;
constexpr int SWARM_OBS = 41;
constexpr int SWARM_AGENT_BLOB = 47;
constexpr int SWARM_AGENT_PAYLOAD = 41; // oracle/drone_oracle.c "swarm env": respawn [0:16], env reset [16:41]
constexpr int SWARM_HORIZON = 1024;
constexpr int SWARM_TASK_RACE = 7;
#define SW_GX 30.0f
#define SW_GY 30.0f
#define SW_GZ 10.0f

// float-valued statistics (Ctl::facc): Log field order of R/dronelib.h:52-63
enum { FACC_RETURN = 0, FACC_LENGTH, FACC_RINGS, FACC_COLLISION, FACC_OOB, FACC_SCORE, FACC_PERF, FACC_N };

struct SwarmDev {
    int n, A, R, rows, ld, epc, max_rings;
    float4 *S, *P, *T, *U, *V, *W;
    int4 *E;
    float4 *G0;
    float2 *G1;
    const float *form; // [3][A][3] closed-form formation targets (orbit, cube, flag), computed on the host
    const float *act_in;
    float *act_out;
    float *obs;
    float *rew;
    unsigned char *term;
    Ctl *ctl;
    const float *payload; // [n][A*41 + 2 + 6R] draw results (inject mode)
    float4 *RS;           // [4][ld] the PREPARED next respawn of every drone: 13 params + position (see sw_refill_pass)
    uint32_t key0, key1, env_id_base;
    int reset_mode;
    // launch overlap (see race_step_kernel): CTA c of step launch `seq` may start while launch seq - 1 drains and
    // waits for chain[c] == seq - 1 only; CTA c owns the same tiles in every launch
    unsigned int *chain; // [grid]
    unsigned int seq;
    int chain_wait;
    int tile_begin, tile_end; // tiles [begin, end) stepped by this launch (host-buffer steps are issued in chunks); end <= 0: all
};

struct SwarmAgent {
    float s[17];
    float p[13];
    float spawn[3], tpos[3], tvel[3];
    float last_abs, last_tgt, last_col, ep_ret, collisions, score;
    int ep_len, ring_idx;
    uint32_t respawns;
};

__device__ __forceinline__ void sw_load(const SwarmDev &d, int k, SwarmAgent &g) {
    const size_t ld = d.ld;
    const float4 q0 = d.S[0 * ld + k], q1 = d.S[1 * ld + k], q2 = d.S[2 * ld + k], q3 = d.S[3 * ld + k], q4 = d.S[4 * ld + k];
    const float4 p0 = d.P[0 * ld + k], p1 = d.P[1 * ld + k], p2 = d.P[2 * ld + k];
    const float4 tt = d.T[k], u = d.U[k], v = d.V[k], w = d.W[k];
    g.s[0] = q0.x; g.s[1] = q0.y; g.s[2] = q0.z; g.s[3] = q0.w; g.s[4] = q1.x; g.s[5] = q1.y; g.s[6] = q1.z; g.s[7] = q1.w;
    g.s[8] = q2.x; g.s[9] = q2.y; g.s[10] = q2.z; g.s[11] = q2.w; g.s[12] = q3.x; g.s[13] = q3.y; g.s[14] = q3.z; g.s[15] = q3.w;
    g.s[16] = q4.x;
    g.ep_len = __float_as_int(q4.y); g.ring_idx = __float_as_int(q4.z); g.ep_ret = q4.w;
    g.p[0] = p0.x; g.p[1] = p0.y; g.p[2] = p0.z; g.p[3] = p0.w; g.p[4] = p1.x; g.p[5] = p1.y; g.p[6] = p1.z; g.p[7] = p1.w;
    g.p[8] = p2.x; g.p[9] = p2.y; g.p[10] = p2.z; g.p[11] = p2.w; g.p[12] = tt.x;
    g.respawns = __float_as_uint(tt.y); g.collisions = tt.z; g.score = tt.w;
    g.spawn[0] = u.x; g.spawn[1] = u.y; g.spawn[2] = u.z; g.last_abs = u.w;
    g.tpos[0] = v.x; g.tpos[1] = v.y; g.tpos[2] = v.z; g.last_tgt = v.w;
    g.tvel[0] = w.x; g.tvel[1] = w.y; g.tvel[2] = w.z; g.last_col = w.w;
}

// Inputs of one agent for one step, staged in shared memory by cp.async one tile ahead of their
// use (slot s of thread t at stage[s * SWARM_BLOCK + t]: conflict-free, thread-private).
constexpr int SW_STAGE_SLOTS = 15; // act, S0..S4, P0..P2, T, U, V, W, E, RS3 (the prepared respawn position)
constexpr int SW_STAGE_BYTES = SW_STAGE_SLOTS * SWARM_BLOCK * 16;
constexpr int SW_DYN_SMEM = SW_STAGE_BYTES + SWARM_BLOCK * SWARM_OBS * 4;
// The rings of a tile's envs ride along with the prefetch (two buffers of [epc][R][8 floats]) when they fit in
// SW_RING_STAGE_MAX bytes, so that the ring test, the race target and the observation never wait on a dependent
// global load; larger configurations (tiny A or hundreds of rings) read rings from global memory as before.
constexpr int SW_RING_STAGE_MAX = 8192;
__host__ __device__ inline int sw_ring_stage_bytes(int epc, int R) {
    const int b = 2 * epc * R * 32;
    return b <= SW_RING_STAGE_MAX ? b : 0;
}
__device__ __forceinline__ void cp_async8(void *sdst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void sw_prefetch(const SwarmDev &d, float4 *stage, int t, int e, int k) {
    const size_t ld = d.ld;
    cp_async16(&stage[0 * SWARM_BLOCK + t], reinterpret_cast<const float4 *>(d.act_in) + k);
#pragma unroll
    for (int m = 0; m < 5; m++) cp_async16(&stage[(1 + m) * SWARM_BLOCK + t], &d.S[m * ld + k]);
#pragma unroll
    for (int m = 0; m < 3; m++) cp_async16(&stage[(6 + m) * SWARM_BLOCK + t], &d.P[m * ld + k]);
    cp_async16(&stage[9 * SWARM_BLOCK + t], &d.T[k]);
    cp_async16(&stage[10 * SWARM_BLOCK + t], &d.U[k]);
    cp_async16(&stage[11 * SWARM_BLOCK + t], &d.V[k]);
    cp_async16(&stage[12 * SWARM_BLOCK + t], &d.W[k]);
    cp_async16(&stage[13 * SWARM_BLOCK + t], &d.E[e]);
    cp_async16(&stage[14 * SWARM_BLOCK + t], &d.RS[3 * ld + k]);
}
// the env's rings into ring-stage buffer `buf` (threads a = 0..A-1 of env slot le share the R rings)
__device__ __forceinline__ void sw_prefetch_rings(const SwarmDev &d, float *rstage, int buf, int le, int a, int e) {
    float *dst = rstage + ((size_t)(buf * d.epc + le) * d.R) * 8;
    for (int r = a; r < d.R; r += d.A) {
        cp_async16(dst + r * 8, &d.G0[(size_t)r * d.n + e]);
        cp_async8(dst + r * 8 + 4, &d.G1[(size_t)r * d.n + e]);
    }
}

__device__ __forceinline__ void sw_load_staged(const float4 *stage, int t, SwarmAgent &g, float4 &a4, int4 &ev) {
    a4 = stage[0 * SWARM_BLOCK + t];
    const float4 q0 = stage[1 * SWARM_BLOCK + t], q1 = stage[2 * SWARM_BLOCK + t], q2 = stage[3 * SWARM_BLOCK + t],
                 q3 = stage[4 * SWARM_BLOCK + t], q4 = stage[5 * SWARM_BLOCK + t];
    const float4 p0 = stage[6 * SWARM_BLOCK + t], p1 = stage[7 * SWARM_BLOCK + t], p2 = stage[8 * SWARM_BLOCK + t];
    const float4 tt = stage[9 * SWARM_BLOCK + t], u = stage[10 * SWARM_BLOCK + t], v = stage[11 * SWARM_BLOCK + t],
                 w = stage[12 * SWARM_BLOCK + t];
    const float4 e4 = stage[13 * SWARM_BLOCK + t];
    ev = make_int4(__float_as_int(e4.x), __float_as_int(e4.y), __float_as_int(e4.z), __float_as_int(e4.w));
    g.s[0] = q0.x; g.s[1] = q0.y; g.s[2] = q0.z; g.s[3] = q0.w; g.s[4] = q1.x; g.s[5] = q1.y; g.s[6] = q1.z; g.s[7] = q1.w;
    g.s[8] = q2.x; g.s[9] = q2.y; g.s[10] = q2.z; g.s[11] = q2.w; g.s[12] = q3.x; g.s[13] = q3.y; g.s[14] = q3.z; g.s[15] = q3.w;
    g.s[16] = q4.x;
    g.ep_len = __float_as_int(q4.y); g.ring_idx = __float_as_int(q4.z); g.ep_ret = q4.w;
    g.p[0] = p0.x; g.p[1] = p0.y; g.p[2] = p0.z; g.p[3] = p0.w; g.p[4] = p1.x; g.p[5] = p1.y; g.p[6] = p1.z; g.p[7] = p1.w;
    g.p[8] = p2.x; g.p[9] = p2.y; g.p[10] = p2.z; g.p[11] = p2.w; g.p[12] = tt.x;
    g.respawns = __float_as_uint(tt.y); g.collisions = tt.z; g.score = tt.w;
    g.spawn[0] = u.x; g.spawn[1] = u.y; g.spawn[2] = u.z; g.last_abs = u.w;
    g.tpos[0] = v.x; g.tpos[1] = v.y; g.tpos[2] = v.z; g.last_tgt = v.w;
    g.tvel[0] = w.x; g.tvel[1] = w.y; g.tvel[2] = w.z; g.last_col = w.w;
}

__device__ __forceinline__ void sw_store(const SwarmDev &d, int k, const SwarmAgent &g, bool params_too) {
    const size_t ld = d.ld;
    d.S[0 * ld + k] = make_float4(g.s[0], g.s[1], g.s[2], g.s[3]);
    d.S[1 * ld + k] = make_float4(g.s[4], g.s[5], g.s[6], g.s[7]);
    d.S[2 * ld + k] = make_float4(g.s[8], g.s[9], g.s[10], g.s[11]);
    d.S[3 * ld + k] = make_float4(g.s[12], g.s[13], g.s[14], g.s[15]);
    d.S[4 * ld + k] = make_float4(g.s[16], __int_as_float(g.ep_len), __int_as_float(g.ring_idx), g.ep_ret);
    d.T[k] = make_float4(g.p[12], __uint_as_float(g.respawns), g.collisions, g.score);
    d.V[k] = make_float4(g.tpos[0], g.tpos[1], g.tpos[2], g.last_tgt);
    d.W[k] = make_float4(g.tvel[0], g.tvel[1], g.tvel[2], g.last_col);
    d.U[k] = make_float4(g.spawn[0], g.spawn[1], g.spawn[2], g.last_abs);
    if (params_too) {
        d.P[0 * ld + k] = make_float4(g.p[0], g.p[1], g.p[2], g.p[3]);
        d.P[1 * ld + k] = make_float4(g.p[4], g.p[5], g.p[6], g.p[7]);
        d.P[2 * ld + k] = make_float4(g.p[8], g.p[9], g.p[10], g.p[11]);
    }
}

__device__ __forceinline__ void sw_load_ring(const SwarmDev &d, int e, int r, float ring[6]) {
    const float4 a = __ldcg(&d.G0[(size_t)r * d.n + e]);
    const float2 b = __ldcg(&d.G1[(size_t)r * d.n + e]);
    ring[0] = a.x; ring[1] = a.y; ring[2] = a.z; ring[3] = a.w; ring[4] = b.x; ring[5] = b.y;
}

// ---- random draws: Philox counters (global env id, who, ordinal, item << 8 | attempt); see
// oracle/drone_oracle.c sw_words.  who = agent (env-wide reset), agent | 0x10000 (respawn),
// 0xFFFF0000 (env-level draws).
__device__ __forceinline__ uint4 sw_words(const SwarmDev &d, uint32_t env, uint32_t who, uint32_t ordinal, uint32_t item,
                                          uint32_t attempt) {
    return philox4x32_10(make_uint4(env, who, ordinal, (item << 8) | attempt), d.key0, d.key1);
}

__device__ __noinline__ void sw_draw_params(const SwarmDev &d, uint32_t env, uint32_t who, uint32_t ordinal, float p[13]) {
    float u[16];
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        const uint4 w = sw_words(d, env, who, ordinal, k, 0u);
        u[4 * k + 0] = unit_from_word(w.x).v; u[4 * k + 1] = unit_from_word(w.y).v;
        u[4 * k + 2] = unit_from_word(w.z).v; u[4 * k + 3] = unit_from_word(w.w).v;
    }
    drone_params_from_draws(u, 0.1f, 0.4f, p); // size ~ U(0.1, 0.4): R/drone_swarm.h:387
}

__device__ __noinline__ void sw_draw_box(const SwarmDev &d, uint32_t env, uint32_t who, uint32_t ordinal, uint32_t item,
                                            uint32_t attempt, float bx, float by, float bz, float out[3]) {
    const uint4 w = sw_words(d, env, who, ordinal, item, attempt);
    out[0] = lerp_u(-bx, bx, unit_from_word(w.x)).v;
    out[1] = lerp_u(-by, by, unit_from_word(w.y)).v;
    out[2] = lerp_u(-bz, bz, unit_from_word(w.z)).v;
}

// R/drone_swarm.h:219-232
__device__ __forceinline__ void sw_move_target(float tp[3], float tv[3]) {
    tp[0] = __fadd_rn(tp[0], tv[0]); tp[1] = __fadd_rn(tp[1], tv[1]); tp[2] = __fadd_rn(tp[2], tv[2]);
    if (tp[0] < -SW_GX || tp[0] > SW_GX) tv[0] = -tv[0];
    if (tp[1] < -SW_GY || tp[1] > SW_GY) tv[1] = -tv[1];
    if (tp[2] < -SW_GZ || tp[2] > SW_GZ) tv[2] = -tv[2];
}

// Positions of the CTA's agents in shared memory, one array per coordinate so that two
// neighbouring candidates load as one 64-bit word per coordinate and their squared distances
// come out of packed FP32x2 instructions (sm_100: add/mul/fma.f32x2, each half an IEEE
// round-to-nearest op, so the strict path stays bit-exact).
//
// WINDOW LAYOUT.  The reference steps the agents of an env in index order, so agent a sees agents
// j < a where they stand NOW (array `lo`) and agents j > a where they stood BEFORE (array `hi`).
// Each env keeps its position arrays back to back, hi first: [hi_0 .. hi_{A-1} | lo_0 .. lo_{A-1}],
// so the A - 1 candidates of agent a are the CONTIGUOUS window positions a+1 .. a+A-1 (hi_{a+1} ..
// hi_{A-1}, then lo_0 .. lo_{a-1}): one linear sweep, no per-candidate array select, no self test.
// Three arrays in a row [old | fin | rst] serve the step scan (hi = old, lo = fin: window at 0) and
// the env-reset scan (hi = fin, lo = rst: window at A); the observation scan uses [pos | pos].
// Envs sit 3A floats apart, which also spreads the envs of one warp over different banks.
constexpr int SW_WIN = 3 * SWARM_BLOCK;
struct SwarmWin {
    float x[SW_WIN], y[SW_WIN], z[SW_WIN];
    __device__ __forceinline__ void put(int i, float px, float py, float pz) { x[i] = px; y[i] = py; z[i] = pz; }
};

// Nearest other agent of agent a standing at `self` (R/drone_swarm.h:107-129); w = &win.x[window
// start of the env].  Returns the distance the reference computes; `other` = that agent's position.
//
// STRICT: ascending index, strict '<' on the reference's sqrtf values (lowest index wins ties);
// the root is taken only for a candidate whose SQUARED distance beats the best so far (sqrtf is
// monotonic, nothing else can win): a handful of roots per sweep instead of A - 1.
__device__ __forceinline__ float sw_nearest_strict(const float *w, int A, int a, const float self[3], float other[3]) {
    other[0] = other[1] = other[2] = 0.0f;
    float best2 = __int_as_float(0x7f800000), bestd = 999999.0f;
    int idx = -1;
    auto consider = [&](int j, float v) {
        if (j != a && v < best2) {
            const float r = xsqrt(xf(v)).v;
            if (r < bestd) { bestd = r; best2 = v; idx = j; }
        }
    };
    if ((A & 1) == 0) {
#pragma unroll 2
        for (int j = 0; j < A; j += 2) {
            const float *src = w + (j < a ? A : 0) + j; // a pair never straddles a: the element at a is skipped
            const float2 ox = *reinterpret_cast<const float2 *>(src);
            const float2 oy = *reinterpret_cast<const float2 *>(src + SW_WIN);
            const float2 oz = *reinterpret_cast<const float2 *>(src + 2 * SW_WIN);
            const float2 dx = __fadd2_rn(ox, make_float2(-self[0], -self[0])), dy = __fadd2_rn(oy, make_float2(-self[1], -self[1])),
                         dz = __fadd2_rn(oz, make_float2(-self[2], -self[2]));
            const float2 d2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
            consider(j, d2.x);
            consider(j + 1, d2.y);
        }
    } else {
        for (int j = 0; j < A; j++) {
            const float *src = w + (j < a ? A : 0) + j;
            const float dx = __fsub_rn(src[0], self[0]), dy = __fsub_rn(src[SW_WIN], self[1]), dz = __fsub_rn(src[2 * SW_WIN], self[2]);
            consider(j, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        }
    }
    if (idx >= 0) {
        const float *src = w + (idx < a ? A : 0) + idx;
        other[0] = src[0]; other[1] = src[SW_WIN]; other[2] = src[2 * SW_WIN];
    }
    return bestd;
}
// out of line for the fast kernel's guard (a decision within rounding distance of its threshold is re-taken
// in the reference's arithmetic): rare, and it must not cost the hot loop registers
__device__ __noinline__ float sw_nearest_strict_call(const float *w, int A, int a, float sx, float sy, float sz, float *other) {
    const float self[3] = {sx, sy, sz};
    float o[3];
    const float r = sw_nearest_strict(w, A, a, self, o);
    other[0] = o[0]; other[1] = o[1]; other[2] = o[2];
    return r;
}

// odd A (not a performance target): one candidate at a time, window offsets from a itself; out of line to keep
// the step kernel's hot loop short.  keys = false: x = smallest squared-distance bits; keys = true: (smallest,
// second smallest) key = squared distance with the window offset in its 7 low mantissa bits.
__device__ __noinline__ uint2 sw_scan_odd(const float *w, int A, int a, float sx, float sy, float sz, bool keys) {
    unsigned int best = 0x7f800000u, second = 0x7f800000u;
    for (int c = 1; c < A; c++) {
        const float *src = w + a + c;
        const float dx = src[0] - sx, dy = src[SW_WIN] - sy, dz = src[2 * SW_WIN] - sz;
        const unsigned int bits = __float_as_uint(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        const unsigned int k0 = keys ? ((bits & ~127u) | (unsigned int)c) : bits;
        second = min(second, max(best, k0));
        best = min(best, k0);
    }
    return make_uint2(best, second);
}

// Fast, distance only (the reward needs no identity, R/drone_swarm.h:347-352): the window sweep above,
// two candidates per iteration on packed FP32x2, the minimum of the squared distances as one 3-input
// integer min per pair (non-negative floats order like their bit patterns).  Exact minimum of the
// squared distances as computed; the collision decision `distance < 1` is guarded by the caller.
template <int UNROLL>
__device__ __forceinline__ float sw_nearest_dist_fast(const float *w, int A, int a, const float self[3]) {
    unsigned int best = 0x7f800000u;
    if ((A & 1) == 0) {
        const int a2 = a & ~1;
        {   // the other agent of a's own pair
            const float *src = w + ((a & 1) ? A + a - 1 : a + 1);
            const float dx = src[0] - self[0], dy = src[SW_WIN] - self[1], dz = src[2 * SW_WIN] - self[2];
            best = min(best, __float_as_uint(fmaf(dz, dz, fmaf(dy, dy, dx * dx))));
        }
        const float2 nx = make_float2(-self[0], -self[0]), ny = make_float2(-self[1], -self[1]), nz = make_float2(-self[2], -self[2]);
        const float *src = w + a2;
#pragma unroll UNROLL
        for (int c = 2; c < A; c += 2) {
            const float2 ox = *reinterpret_cast<const float2 *>(src + c);
            const float2 oy = *reinterpret_cast<const float2 *>(src + c + SW_WIN);
            const float2 oz = *reinterpret_cast<const float2 *>(src + c + 2 * SW_WIN);
            const float2 dx = __fadd2_rn(ox, nx), dy = __fadd2_rn(oy, ny), dz = __fadd2_rn(oz, nz);
            const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            best = __vimin3_u32(best, __float_as_uint(d2.x), __float_as_uint(d2.y));
        }
    } else {
        best = sw_scan_odd(w, A, a, self[0], self[1], self[2], false).x;
    }
    return best >= 0x7f800000u ? 999999.0f : sqrtf(__uint_as_float(best));
}

// Fast, with the neighbour's identity (the observation needs its position, R/drone_swarm.h:187-196).  Keys =
// squared distance with its 7 low mantissa bits replaced by the window offset; the sweep keeps the smallest
// AND the second smallest key.  When the two lie within two key buckets (3e-5 relative) the order of the
// candidates is not safe against rounding -- `ambiguous` -- and the caller re-takes the decision with
// sw_nearest_strict_call; otherwise the winner is the reference's.
template <int UNROLL>
__device__ __forceinline__ void sw_nearest_fast(const float *w, int A, int a, const float self[3], float other[3], bool &ambiguous) {
    constexpr unsigned int KEY_NONE = 0x7f800000u, CODE_MASK = 127u;
    other[0] = other[1] = other[2] = 0.0f;
    unsigned int best = KEY_NONE, second = KEY_NONE;
    int a2 = a;
    if ((A & 1) == 0) {
        a2 = a & ~1;
        {   // the other agent of a's own pair: code 1
            const float *src = w + ((a & 1) ? A + a - 1 : a + 1);
            const float dx = src[0] - self[0], dy = src[SW_WIN] - self[1], dz = src[2 * SW_WIN] - self[2];
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            best = (__float_as_uint(d2) & ~CODE_MASK) | 1u;
        }
        const float2 nx = make_float2(-self[0], -self[0]), ny = make_float2(-self[1], -self[1]), nz = make_float2(-self[2], -self[2]);
        const float *src = w + a2;
#pragma unroll UNROLL
        for (int c = 2; c < A; c += 2) { // window offsets c, c + 1 from a's pair
            const float2 ox = *reinterpret_cast<const float2 *>(src + c);
            const float2 oy = *reinterpret_cast<const float2 *>(src + c + SW_WIN);
            const float2 oz = *reinterpret_cast<const float2 *>(src + c + 2 * SW_WIN);
            const float2 dx = __fadd2_rn(ox, nx), dy = __fadd2_rn(oy, ny), dz = __fadd2_rn(oz, nz);
            const float2 d2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            const unsigned int k0 = (__float_as_uint(d2.x) & ~CODE_MASK) | (unsigned int)c;
            const unsigned int k1 = (__float_as_uint(d2.y) & ~CODE_MASK) | (unsigned int)(c + 1);
            const unsigned int lo = min(k0, k1), hi = max(k0, k1);
            second = __vimin3_u32(second, hi, max(best, lo)); // second smallest of {best, second, k0, k1}
            best = min(best, lo);
        }
    } else {
        const uint2 bs = sw_scan_odd(w, A, a, self[0], self[1], self[2], true);
        best = bs.x;
        second = bs.y;
    }
    ambiguous = second < KEY_NONE && (second & ~CODE_MASK) - (best & ~CODE_MASK) <= 2u * (CODE_MASK + 1u);
    if (best >= KEY_NONE) return;
    const int code = (int)(best & CODE_MASK);
    const float *src = ((A & 1) == 0 && code == 1) ? w + ((a & 1) ? A + a - 1 : a + 1) : w + a2 + code;
    other[0] = src[0]; other[1] = src[SW_WIN]; other[2] = src[2 * SW_WIN];
}

// Guard bands of the fast kernel (see race_strict_replay for the idea): decisions within these distances of
// their threshold are re-taken in the reference's arithmetic.  Arena half-widths are 30 / 30 / 10 m (1 ulp =
// 2e-6 m), the largest fast-vs-reference position error measured is 1 ulp (profiles/parity_r02.json).
constexpr float SW_GUARD_WALL = 1e-4f;   // out of bounds: | |coordinate| - half width |
constexpr float SW_GUARD_PLANE = 1e-4f;  // ring plane crossing
constexpr float SW_GUARD_DIST = 1e-4f;   // collision: | nearest distance - 1 |

// nearest-neighbour DISTANCE for compute_reward (R/drone_swarm.h:347-352)
template <bool STRICT, int UNROLL = 2>
__device__ __forceinline__ float sw_reward_distance(const float *w, int A, int a, const float self[3], int &guard_hits) {
    if constexpr (STRICT) {
        float other[3];
        return sw_nearest_strict(w, A, a, self, other);
    } else {
        float nd = sw_nearest_dist_fast<UNROLL>(w, A, a, self);
        if (fabsf(nd - 1.0f) < SW_GUARD_DIST) {
            float other[3];
            nd = sw_nearest_strict_call(w, A, a, self[0], self[1], self[2], other);
            guard_hits += 1;
        }
        return nd;
    }
}

// the same, out of line: the scans of the rare paths (respawn, env-wide reset)
template <bool STRICT>
__device__ __noinline__ float sw_reward_distance_cold(const float *w, int A, int a, float sx, float sy, float sz, int *guard_hits) {
    const float self[3] = {sx, sy, sz};
    int hits = 0;
    const float nd = sw_reward_distance<STRICT>(w, A, a, self, hits);
    *guard_hits += hits;
    return nd;
}

// nearest neighbour's POSITION for compute_observations (R/drone_swarm.h:187-196)
template <bool STRICT, int UNROLL>
__device__ __forceinline__ void sw_obs_neighbour(const float *w, int A, int a, const float self[3], float near[3], int &guard_hits) {
    if constexpr (STRICT) {
        sw_nearest_strict(w, A, a, self, near);
    } else {
        bool ambiguous;
        sw_nearest_fast<UNROLL>(w, A, a, self, near, ambiguous);
        if (ambiguous) {
            float other[3];
            sw_nearest_strict_call(w, A, a, self[0], self[1], self[2], other);
            near[0] = other[0]; near[1] = other[1]; near[2] = other[2];
            guard_hits += 1;
        }
    }
}

// R/drone_swarm.h:335-376.  Side effects on the agent exactly as the reference's.
template <bool STRICT>
__device__ __forceinline__ float sw_reward(SwarmAgent &g, const float self[3], bool collision, int A, float nearest_dist) {
    float dist_reward;
    if constexpr (STRICT) {
        const xf dx = xf(self[0]) - xf(g.tpos[0]), dy = xf(self[1]) - xf(g.tpos[1]), dz = xf(self[2]) - xf(g.tpos[2]);
        const xf dist = xsqrt(dx * dx + dy * dy + dz * dz);
        const xf maxd = xsqrt(xf(7600.0f)); // sqrtf(60^2 + 60^2 + 20^2)
        // (float)(1.0 - (double)(dist / MAX_DIST)) == 1.0f - dist / MAX_DIST: the double difference is exact
        dist_reward = (xf(1.0f) - sdiv<false>(dist, maxd)).v;
    } else {
        const float dx = self[0] - g.tpos[0], dy = self[1] - g.tpos[1], dz = self[2] - g.tpos[2];
        dist_reward = 1.0f - sqrtf(dx * dx + dy * dy + dz * dz) * 0.011470787f; // 1 / sqrt(7600)
    }
    float density_reward = 0.0f;
    if (collision && A > 1 && nearest_dist < 1.0f) {
        density_reward = -1.0f;
        g.collisions = __fadd_rn(g.collisions, 1.0f);
    }
    float abs_reward = __fadd_rn(dist_reward, density_reward);
    if (dist_reward < 0.0f && density_reward < 0.0f) abs_reward = -abs_reward;
    const float delta = __fsub_rn(abs_reward, g.last_abs);
    g.last_col = density_reward;
    g.last_tgt = dist_reward;
    g.last_abs = abs_reward;
    g.ep_len += 1;
    g.score = __fadd_rn(g.score, abs_reward);
    return delta;
}

// R/drone_swarm.h:131-217, one agent; `near` = its nearest neighbour's position
template <bool STRICT>
__device__ __forceinline__ void sw_observe(const SwarmAgent &g, int A, const float near[3], bool race, const float ring[6],
                                           float *row, int stride) {
    float o[SWARM_OBS];
    if constexpr (STRICT) {
        Q4<xf> q, qi;
        q.w = g.s[6]; q.x = g.s[7]; q.y = g.s[8]; q.z = g.s[9];
        qi.w = q.w; qi.x = -q.x; qi.y = -q.y; qi.z = -q.z;
        V3<xf> vel, zax;
        vel.x = g.s[3]; vel.y = g.s[4]; vel.z = g.s[5];
        zax.x = 0.0f; zax.y = 0.0f; zax.z = 1.0f;
        const V3<xf> vb = qrot(qi, vel), up = qrot(q, zax);
        o[0] = sdiv<false>(vb.x, xf(B2D_MAX_VEL)).v; o[1] = sdiv<false>(vb.y, xf(B2D_MAX_VEL)).v; o[2] = sdiv<false>(vb.z, xf(B2D_MAX_VEL)).v;
        o[3] = sdiv<false>(xf(g.s[10]), xf(B2D_MAX_OMEGA)).v; o[4] = sdiv<false>(xf(g.s[11]), xf(B2D_MAX_OMEGA)).v;
        o[5] = sdiv<false>(xf(g.s[12]), xf(B2D_MAX_OMEGA)).v;
        o[6] = up.x.v; o[7] = up.y.v; o[8] = up.z.v;
        o[9] = g.s[6]; o[10] = g.s[7]; o[11] = g.s[8]; o[12] = g.s[9];
#pragma unroll
        for (int m = 0; m < 4; m++) o[13 + m] = sdiv<false>(xf(g.s[13 + m]), xf(g.p[10])).v;
        o[17] = sdiv<false>(xf(g.s[0]), xf(SW_GX)).v; o[18] = sdiv<false>(xf(g.s[1]), xf(SW_GY)).v; o[19] = sdiv<false>(xf(g.s[2]), xf(SW_GZ)).v;
        o[20] = sdiv<false>(xf(g.spawn[0]), xf(SW_GX)).v; o[21] = sdiv<false>(xf(g.spawn[1]), xf(SW_GY)).v; o[22] = sdiv<false>(xf(g.spawn[2]), xf(SW_GZ)).v;
        const xf dx = xf(g.tpos[0]) - xf(g.s[0]), dy = xf(g.tpos[1]) - xf(g.s[1]), dz = xf(g.tpos[2]) - xf(g.s[2]);
        o[23] = xclamp(dx, -1.0f, 1.0f).v; o[24] = xclamp(dy, -1.0f, 1.0f).v; o[25] = xclamp(dz, -1.0f, 1.0f).v;
        o[26] = sdiv<false>(dx, xf(SW_GX)).v; o[27] = sdiv<false>(dy, xf(SW_GY)).v; o[28] = sdiv<false>(dz, xf(SW_GZ)).v;
        o[29] = g.last_col; o[30] = g.last_tgt; o[31] = g.last_abs;
        if (A > 1) {
            o[32] = xclamp(xf(near[0]) - xf(g.s[0]), -1.0f, 1.0f).v;
            o[33] = xclamp(xf(near[1]) - xf(g.s[1]), -1.0f, 1.0f).v;
            o[34] = xclamp(xf(near[2]) - xf(g.s[2]), -1.0f, 1.0f).v;
        } else {
            o[32] = o[33] = o[34] = 0.0f;
        }
        if (race) {
            V3<xf> dd, nn;
            dd.x = xf(ring[0]) - xf(g.s[0]); dd.y = xf(ring[1]) - xf(g.s[1]); dd.z = xf(ring[2]) - xf(g.s[2]);
            nn.x = ring[3]; nn.y = ring[4]; nn.z = ring[5];
            const V3<xf> to = qrot(qi, dd), bn = qrot(qi, nn);
            o[35] = sdiv<false>(to.x, xf(SW_GX)).v; o[36] = sdiv<false>(to.y, xf(SW_GY)).v; o[37] = sdiv<false>(to.z, xf(SW_GZ)).v;
            o[38] = bn.x.v; o[39] = bn.y.v; o[40] = bn.z.v;
        } else {
#pragma unroll
            for (int m = 35; m < 41; m++) o[m] = 0.0f;
        }
    } else {
        const float w = g.s[6], x = g.s[7], y = g.s[8], z = g.s[9];
        const float ww = w * w, xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
        const float r00 = (ww + xx) - (yy + zz), r11 = (ww - xx) + (yy - zz), r22 = (ww - xx) - (yy - zz);
        const float r01 = 2.0f * (xy - wz), r10 = 2.0f * (xy + wz);
        const float r02 = 2.0f * (xz + wy), r20 = 2.0f * (xz - wy);
        const float r12 = 2.0f * (yz - wx), r21 = 2.0f * (yz + wx);
        o[0] = 0.02f * (r00 * g.s[3] + r10 * g.s[4] + r20 * g.s[5]);
        o[1] = 0.02f * (r01 * g.s[3] + r11 * g.s[4] + r21 * g.s[5]);
        o[2] = 0.02f * (r02 * g.s[3] + r12 * g.s[4] + r22 * g.s[5]);
        o[3] = 0.02f * g.s[10]; o[4] = 0.02f * g.s[11]; o[5] = 0.02f * g.s[12];
        o[6] = r02; o[7] = r12; o[8] = r22;
        o[9] = w; o[10] = x; o[11] = y; o[12] = z;
        const float inv = __frcp_rn(g.p[10]);
#pragma unroll
        for (int m = 0; m < 4; m++) o[13 + m] = g.s[13 + m] * inv;
        const float igx = 1.0f / SW_GX, igz = 1.0f / SW_GZ;
        o[17] = g.s[0] * igx; o[18] = g.s[1] * igx; o[19] = g.s[2] * igz;
        o[20] = g.spawn[0] * igx; o[21] = g.spawn[1] * igx; o[22] = g.spawn[2] * igz;
        const float dx = g.tpos[0] - g.s[0], dy = g.tpos[1] - g.s[1], dz = g.tpos[2] - g.s[2];
        o[23] = fminf(fmaxf(dx, -1.0f), 1.0f); o[24] = fminf(fmaxf(dy, -1.0f), 1.0f); o[25] = fminf(fmaxf(dz, -1.0f), 1.0f);
        o[26] = dx * igx; o[27] = dy * igx; o[28] = dz * igz;
        o[29] = g.last_col; o[30] = g.last_tgt; o[31] = g.last_abs;
        if (A > 1) {
            o[32] = fminf(fmaxf(near[0] - g.s[0], -1.0f), 1.0f);
            o[33] = fminf(fmaxf(near[1] - g.s[1], -1.0f), 1.0f);
            o[34] = fminf(fmaxf(near[2] - g.s[2], -1.0f), 1.0f);
        } else {
            o[32] = o[33] = o[34] = 0.0f;
        }
        if (race) {
            const float ex = ring[0] - g.s[0], ey = ring[1] - g.s[1], ez = ring[2] - g.s[2];
            o[35] = igx * (r00 * ex + r10 * ey + r20 * ez);
            o[36] = igx * (r01 * ex + r11 * ey + r21 * ez);
            o[37] = igz * (r02 * ex + r12 * ey + r22 * ez);
            o[38] = r00 * ring[3] + r10 * ring[4] + r20 * ring[5];
            o[39] = r01 * ring[3] + r11 * ring[4] + r21 * ring[5];
            o[40] = r02 * ring[3] + r12 * ring[4] + r22 * ring[5];
        } else {
#pragma unroll
            for (int m = 35; m < 41; m++) o[m] = 0.0f;
        }
    }
#pragma unroll
    for (int m = 0; m < SWARM_OBS; m++) row[m * stride] = o[m];
}

// R/drone_swarm.h:378-390 minus the reward: statistics zeroed, new drone, at `pos`
__device__ __forceinline__ void sw_respawn_state(SwarmAgent &g, const float p[13], const float pos[3]) {
    g.ep_ret = 0.0f;
    g.ep_len = 0;
    g.collisions = 0.0f;
    g.score = 0.0f;
    g.ring_idx = 0;
#pragma unroll
    for (int k = 0; k < 13; k++) g.p[k] = p[k];
#pragma unroll
    for (int k = 0; k < 17; k++) g.s[k] = 0.0f;
    g.s[6] = 1.0f;
    g.s[0] = pos[0]; g.s[1] = pos[1]; g.s[2] = pos[2];
    g.spawn[0] = pos[0]; g.spawn[1] = pos[1]; g.spawn[2] = pos[2];
}

// The fast kernel's guard for the move itself: the drone's step re-done in the reference's arithmetic from its
// pre-step state (still in global memory: the step stores state in its last phase), plus the ring test.
// out[0:17] = state, out[17] = gate event
__device__ __noinline__ void sw_strict_move(const SwarmDev &d, int k, float4 a4, const float *ring /* nullptr: no ring test */, float *out) {
    const size_t ld = d.ld;
    const float4 q0 = __ldcg(&d.S[0 * ld + k]), q1 = __ldcg(&d.S[1 * ld + k]), q2 = __ldcg(&d.S[2 * ld + k]), q3 = __ldcg(&d.S[3 * ld + k]),
                 q4 = __ldcg(&d.S[4 * ld + k]);
    const float4 p0 = __ldcg(&d.P[0 * ld + k]), p1 = __ldcg(&d.P[1 * ld + k]), p2 = __ldcg(&d.P[2 * ld + k]), tt = __ldcg(&d.T[k]);
    float s[17] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x};
    const DroneParams p = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, tt.x};
    const float act[4] = {xclamp(xf(a4.x), -1.0f, 1.0f).v, xclamp(xf(a4.y), -1.0f, 1.0f).v, xclamp(xf(a4.z), -1.0f, 1.0f).v,
                          xclamp(xf(a4.w), -1.0f, 1.0f).v};
    const float before[3] = {s[0], s[1], s[2]};
    advance_body_strict(s, p, act);
#pragma unroll
    for (int m = 0; m < 17; m++) out[m] = s[m];
    float gate = 0.0f;
    if (ring) {
        const float rg[6] = {ring[0], ring[1], ring[2], ring[3], ring[4], ring[5]};
        gate = gate_event<xf>(before, s, rg, -0.0f);
    }
    out[17] = gate;
}

// ---- prepared respawns.  The respawn of agent a of env g after r earlier respawns is a pure function of
// (seed, g, a, r + 1) (sw_draw_params / sw_draw_box).  Drawing it where it is needed puts ~700 instructions
// behind a branch 2.4 % of the drones take per step: more than half of all warps ran them with one lane
// active, a quarter of the whole step (profiles/r01h_swarm_variants.txt).  So every drone's NEXT respawn
// sits ready in d.RS (64 B per drone); a drone that leaves the arena reads it and leaves its row id in its
// warp's list; a warp regenerates slots 32 at a time -- all lanes busy -- once its list holds 32 rows, and
// flushes the rest at the end of the launch.  A slot is read at most once per launch, by the warp that
// later rewrites it, so every launch is self-contained.
__device__ __forceinline__ void sw_generate_slot(const SwarmDev &d, int k, uint32_t ordinal) {
    const int env = k / d.A, a = k - env * d.A;
    const uint32_t genv = d.env_id_base + (uint32_t)env;
    float rp[13], rpos[3];
    sw_draw_params(d, genv, (uint32_t)a | 0x10000u, ordinal, rp);
    sw_draw_box(d, genv, (uint32_t)a | 0x10000u, ordinal, 4u, 0u, 29.0f, 29.0f, 9.0f, rpos);
    const size_t ld = d.ld;
    d.RS[0 * ld + k] = make_float4(rp[0], rp[1], rp[2], rp[3]);
    d.RS[1 * ld + k] = make_float4(rp[4], rp[5], rp[6], rp[7]);
    d.RS[2 * ld + k] = make_float4(rp[8], rp[9], rp[10], rp[11]);
    d.RS[3 * ld + k] = make_float4(rp[12], rpos[0], rpos[1], rpos[2]);
}
__device__ __noinline__ void sw_refill_pass(const SwarmDev &d, const int2 *list, int count, int lane) {
    if (lane < count) sw_generate_slot(d, list[lane].x, (uint32_t)list[lane].y);
}
// every drone's slot from its current respawn count (create, vec_reset)
__global__ void __launch_bounds__(128) swarm_fill_slots_kernel(const __grid_constant__ SwarmDev d) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < d.rows) sw_generate_slot(d, k, __float_as_uint(d.T[k].y) + 1u);
}

// barrier over the threads of one env (see swarm_kernel): warp / named barrier / CTA
__device__ __forceinline__ void sw_env_sync(int sync_mode, int le, int grp_size) {
    if (sync_mode == 1) __syncwarp();
    else if (sync_mode == 2) asm volatile("bar.sync %0, %1;" ::"r"(1 + le), "r"(grp_size) : "memory");
    else __syncthreads();
}

struct SwResetCtx {
    SwarmAgent g;
    int tick, task;
    uint32_t env_episode;
    bool do_reset, params_dirty;
    int guard_hits;
};

// c_reset of the swarm env (R/drone_swarm.h:401-443) for the envs of this barrier group whose horizon is up
// (do_reset); called by every thread of the group.  rings_le: the env's staged rings in shared memory, or nullptr.
template <bool STRICT, bool ONLY_RESET>
__device__ __noinline__ void sw_reset_phase(const SwarmDev &d, SwResetCtx *c, SwarmWin *s_trail, float (*s_ring0)[3], float *rings_le,
                                            int le, int a, int w0, int e, const float *pay_agent, const float *pay_env,
                                            int sync_mode, int grp_size) {
    const int A = d.A;
    const bool inject = d.reset_mode == 1;
    const uint32_t genv = d.env_id_base + (uint32_t)e;
    SwarmAgent g = c->g;
    int tick = c->tick, task = c->task, guard_hits = 0;
    uint32_t env_episode = c->env_episode;
    const bool do_reset = c->do_reset;
    bool params_dirty = c->params_dirty;
    float first[3] = {0.0f, 0.0f, 0.0f}, np[13];
    if (do_reset) {
        tick = 0;
        env_episode = ONLY_RESET ? 0u : env_episode + 1u;
        if (inject) {
            task = (int)pay_env[1];
#pragma unroll
            for (int m = 0; m < 13; m++) np[m] = pay_agent[16 + m];
            first[0] = pay_agent[29]; first[1] = pay_agent[30]; first[2] = pay_agent[31];
        } else {
            const uint4 w = sw_words(d, genv, 0xFFFF0000u, env_episode, 0u, 0u);
            task = ((w.x >> 1) % 4u) ? SWARM_TASK_RACE : (int)((w.y >> 1) % 7u);
            sw_draw_params(d, genv, (uint32_t)a, env_episode, np);
            sw_draw_box(d, genv, (uint32_t)a, env_episode, 4u, 0u, 29.0f, 29.0f, 9.0f, first);
        }
        s_trail->put(w0 + 2 * A + a, first[0], first[1], first[2]);
    }
    sw_env_sync(sync_mode, le, grp_size);
    if (do_reset) {
        // reset_agent: the reward is computed against the STALE target and half-reset neighbours
        sw_respawn_state(g, np, first);
        params_dirty = true;
        float nd = 0.0f;
        if (A > 1 && task != SWARM_TASK_RACE) nd = sw_reward_distance<STRICT>(&s_trail->x[w0 + A], A, a, first, guard_hits);
        sw_reward<STRICT>(g, first, task != SWARM_TASK_RACE, A, nd);
        // set_target: R/drone_swarm.h:234-333
        if (inject) {
#pragma unroll
            for (int m = 0; m < 3; m++) { g.tpos[m] = pay_agent[32 + m]; g.tvel[m] = pay_agent[35 + m]; }
        } else if (task == 0 || ((task == 3 || task == 5) && a == 0)) {
            sw_draw_box(d, genv, (uint32_t)a, env_episode, 5u, 0u, 29.0f, 29.0f, 9.0f, g.tpos);
            sw_draw_box(d, genv, (uint32_t)a, env_episode, 6u, 0u, 0.05f, 0.05f, 0.05f, g.tvel);
        } else if (task == 3 || task == 5) {
            // follow: agent 0's idle target; congo: the same target advanced 40 moves per link of the chain
            sw_draw_box(d, genv, 0u, env_episode, 5u, 0u, 29.0f, 29.0f, 9.0f, g.tpos);
            sw_draw_box(d, genv, 0u, env_episode, 6u, 0u, 0.05f, 0.05f, 0.05f, g.tvel);
            if (task == 5)
                for (int m = 0; m < 40 * a; m++) sw_move_target(g.tpos, g.tvel);
        } else if (task == 1) {
            g.tpos[0] = first[0]; g.tpos[1] = first[1]; g.tpos[2] = first[2];
            g.tvel[0] = g.tvel[1] = g.tvel[2] = 0.0f;
        } else if (task == SWARM_TASK_RACE) {
            // rings are regenerated AFTER the targets are set: the target is the previous episode's ring 0
            float ring[6];
            if (rings_le) { for (int m = 0; m < 6; m++) ring[m] = rings_le[m]; } else sw_load_ring(d, e, 0, ring);
            g.tpos[0] = ring[0]; g.tpos[1] = ring[1]; g.tpos[2] = ring[2];
            g.tvel[0] = g.tvel[1] = g.tvel[2] = 0.0f;
        } else {
            const float *f = d.form + ((size_t)(task == 2 ? 0 : (task == 4 ? 1 : 2)) * A + a) * 3;
            g.tpos[0] = f[0]; g.tpos[1] = f[1]; g.tpos[2] = f[2];
            g.tvel[0] = g.tvel[1] = g.tvel[2] = 0.0f;
        }
    }
    sw_env_sync(sync_mode, le, grp_size); // every agent has read the old ring 0 before the rings are rewritten
    if (do_reset && a == 0) {
        float prev[3] = {0.0f, 0.0f, 0.0f};
        for (int r = 0; r < d.R; r++) {
            float ring[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (task == SWARM_TASK_RACE) {
                if (inject) {
#pragma unroll
                    for (int m = 0; m < 6; m++) ring[m] = pay_env[2 + 6 * r + m];
                } else {
                    for (uint32_t at = 0; at < RESET_MAX_ATTEMPTS; at++) {
                        const uint4 wa = sw_words(d, genv, 0xFFFF0000u, env_episode, 0x10u + 2u * r, at);
                        const uint4 wb = sw_words(d, genv, 0xFFFF0000u, env_episode, 0x11u + 2u * r, at);
                        ring_from_words(wa, wb, 26.0f, 26.0f, 6.0f, ring);
                        if (r == 0 || !(dist3_exact(ring, prev) < 4.0f)) break;
                    }
                }
            }
            prev[0] = ring[0]; prev[1] = ring[1]; prev[2] = ring[2];
            d.G0[(size_t)r * d.n + e] = make_float4(ring[0], ring[1], ring[2], ring[3]);
            d.G1[(size_t)r * d.n + e] = make_float2(ring[4], ring[5]);
            if (rings_le) { // the staged copy of this tile follows (the observation below reads ring 0)
                float *dst = rings_le + r * 8;
#pragma unroll
                for (int m = 0; m < 6; m++) dst[m] = ring[m];
            }
            if (r == 0) { s_ring0[le][0] = ring[0]; s_ring0[le][1] = ring[1]; s_ring0[le][2] = ring[2]; }
        }
    }
    sw_env_sync(sync_mode, le, grp_size);
    if (do_reset && task == SWARM_TASK_RACE) {
        // start at least 2*radius from the first ring; spawn_pos / prev_pos keep the first draw (R/drone_swarm.h:429-439)
        const float r0[3] = {s_ring0[le][0], s_ring0[le][1], s_ring0[le][2]};
        float c[3];
        if (inject) {
            c[0] = pay_agent[38]; c[1] = pay_agent[39]; c[2] = pay_agent[40];
        } else {
            for (uint32_t at = 0; at < RESET_MAX_ATTEMPTS; at++) {
                sw_draw_box(d, genv, (uint32_t)a, env_episode, 7u, at, 29.0f, 29.0f, 9.0f, c);
                if (!(dist3_exact(c, r0) < 4.0f)) break;
            }
        }
        g.s[0] = c[0]; g.s[1] = c[1]; g.s[2] = c[2];
    }

    c->g = g; c->tick = tick; c->task = task; c->env_episode = env_episode; c->params_dirty = params_dirty; c->guard_hits = guard_hits;
}

// ---------------------------------------------------------------- the kernel
// ONLY_RESET = false: one vec_step.  ONLY_RESET = true: vec_reset (every env runs c_reset).
//
// A tile = the SWARM_BLOCK / A envs one CTA steps together.  The step launch is persistent (a few
// resident CTAs per SM, CTA c takes tiles c, c + grid, ...): while a tile computes (~2-3 k
// instructions per agent) the 14 input words of the CTA's next tile stream into shared memory
// with cp.async, so no warp ever waits on a global load at the top of a tile.
// SCAN_UNROLL: unroll factor of the fast neighbour sweeps.  Two builds of the fast step: for A <= 32 the compact
// sweep (2) wins -- the step's hot code is what costs there (32 KB instruction cache level): A = 16 113 vs 119 us, A = 32
// 244 vs 250 us -- for longer sweeps the unrolled one (4): A = 64 544 vs 553 us (profiles/r02_ab/r02x_ab.txt, r02y_ab.txt).
template <bool STRICT, bool ONLY_RESET, int SCAN_UNROLL = 4>
__global__ void __launch_bounds__(SWARM_BLOCK, 3) swarm_kernel(const __grid_constant__ SwarmDev d) {
    // per env (3A floats apart): [old | fin | rst] = before this tick's move | after the move, or the respawn
    // position of an agent that left the arena | first position drawn by an env-wide reset
    __shared__ __align__(16) SwarmWin s_trail;
    __shared__ __align__(16) SwarmWin s_now; // per env [pos | pos]: final positions of the tick (what the observations see)
    __shared__ float s_ring0[SWARM_BLOCK][3];
    // episode statistics in 2^-20 fixed point, one set per warp, updated by the warp's lane 0 with plain loads and
    // stores (64-bit shared-memory atomics are compare-and-swap loops: seven of them per finished episode on seven
    // words shared by the whole CTA cost 15 % of the step at A = 16)
    __shared__ long long s_wacc[SWARM_BLOCK / 32][8];
    __shared__ int2 s_rlist[SWARM_BLOCK / 32][64]; // per warp: (row, ordinal) of the respawn slots to regenerate
    __shared__ int s_rcnt[SWARM_BLOCK / 32];
#if B2D_SW_RS_POOL
    // per warp: the prepared respawn parameters (3 float4) of the first few drones that leave the arena in this
    // tile, fetched by cp.async in phase 1 and read in phase 2 (a dependent DRAM load there cost 8 % of the step)
    __shared__ __align__(16) float4 s_rspool[SWARM_BLOCK / 32][B2D_SW_RS_POOL][3];
#endif
    __shared__ int s_guard;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    float4 *stage = reinterpret_cast<float4 *>(s_dyn);                       // step launches only
    float *s_obs = reinterpret_cast<float *>(s_dyn + (ONLY_RESET ? 0 : SW_STAGE_BYTES));
    float *rstage = reinterpret_cast<float *>(s_dyn + SW_DYN_SMEM); // step launches only, when ring_staged
    const bool ring_staged = !ONLY_RESET && sw_ring_stage_bytes(d.epc, d.R) > 0;
    int rbuf = 0; // ring-stage buffer of the current tile

    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int A = d.A;
    const int le = t / A;
    const int a = t - le * A;
    const int w0 = le * 3 * A; // this env's windows
    const bool inject = d.reset_mode == 1;
    const int ntiles_all = (d.n + d.epc - 1) / d.epc;
    const int ntiles = d.tile_end > 0 && d.tile_end < ntiles_all ? d.tile_end : ntiles_all;
    // CTA c owns the tiles congruent to c modulo the grid, whatever range a launch covers
    const int first_tile = d.tile_begin + (int)((blockIdx.x + gridDim.x - (unsigned)d.tile_begin % gridDim.x) % gridDim.x);
    const size_t pay_stride = (size_t)A * SWARM_AGENT_PAYLOAD + 2 + 6 * d.R;

    // Everything the phases exchange through shared memory stays inside one env, so the barriers
    // between the phases only need the env's own threads: a warp when A divides 32 (one or more whole
    // envs per warp), a named barrier over the env's warps when A is a multiple of 32, the whole CTA
    // otherwise.  A drone that respawns (Philox draws, a second neighbour scan) then delays its own
    // env, not the other envs of the CTA: with CTA-wide barriers the respawn path of 2.4 % of the
    // drones cost a quarter of the step (profiles/r01h_swarm_variants.txt).
    const int sync_mode = (32 % A == 0) ? 1 : (A % 32 == 0 ? 2 : 0);
    const int grp_first = sync_mode == 1 ? (t & ~31) : sync_mode == 2 ? le * A : 0;                       // first thread of my group
    const int grp_size = sync_mode == 1 ? 32 : sync_mode == 2 ? min(A, SWARM_BLOCK - le * A) : SWARM_BLOCK; // threads in it
    auto env_sync = [&]() {
        if (sync_mode == 1) __syncwarp();
        else if (sync_mode == 2) asm volatile("bar.sync %0, %1;" ::"r"(1 + le), "r"(grp_size) : "memory");
        else __syncthreads();
    };
    auto env_any = [&](bool p) -> int {
        if (sync_mode == 1) return __any_sync(0xffffffffu, p);
        if (sync_mode == 2) {
            unsigned int r;
            asm volatile("{ .reg .pred p, q; setp.ne.u32 p, %1, 0; bar.red.or.pred q, %2, %3, p; selp.u32 %0, 1, 0, q; }"
                         : "=r"(r)
                         : "r"((unsigned int)p), "r"(1 + le), "r"(grp_size)
                         : "memory");
            return (int)r;
        }
        return __syncthreads_or(p ? 1 : 0);
    };
    if (t < (SWARM_BLOCK / 32) * 8) s_wacc[t >> 3][t & 7] = 0ll; // all this CTA's tiles; flushed once at the end
    auto load_ring = [&](int e_, int r, float ring[6]) {
        if (ring_staged) {
            const float *src = rstage + ((size_t)(rbuf * d.epc + le) * d.R + r) * 8;
            const float4 q = *reinterpret_cast<const float4 *>(src);
            const float2 h = *reinterpret_cast<const float2 *>(src + 4);
            ring[0] = q.x; ring[1] = q.y; ring[2] = q.z; ring[3] = q.w; ring[4] = h.x; ring[5] = h.y;
        } else {
            sw_load_ring(d, e_, r, ring);
        }
    };
    if (t < SWARM_BLOCK / 32) s_rcnt[t] = 0;
    if (t == 0) s_guard = 0;
    int guard_hits = 0;
    if constexpr (!ONLY_RESET) {
#if B2D_SW_OVERLAP
        // the next step launch may take CTA slots as this launch's CTAs leave; its CTA c waits for this CTA c only
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (d.chain_wait && t == 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(d.chain + blockIdx.x) : "memory");
                if (seen != d.seq - 1u) __nanosleep(64);
            } while (seen != d.seq - 1u);
        }
#endif
    }
    __syncthreads();

    if constexpr (!ONLY_RESET) {
        const int e0 = first_tile * d.epc + le;
        if (first_tile < ntiles && le < d.epc && e0 < d.n) {
            sw_prefetch(d, stage, t, e0, e0 * A + a);
            if (ring_staged) sw_prefetch_rings(d, rstage, 0, le, a, e0);
        }
        cp_async_commit();
    }

  for (int tile = first_tile; tile < ntiles; tile += gridDim.x) {
    const int e = tile * d.epc + le;
    const bool active = le < d.epc && e < d.n;
    const int k = e * A + a;
    const float *pay_agent = d.payload ? d.payload + (size_t)e * pay_stride + (size_t)a * SWARM_AGENT_PAYLOAD : nullptr;
    const float *pay_env = d.payload ? d.payload + (size_t)e * pay_stride + (size_t)A * SWARM_AGENT_PAYLOAD : nullptr;

    SwarmAgent g;
    int tick = 0, task = 0;
    uint32_t env_episode = 0;
    float reward = 0.0f;
    int terminal = 0;
    bool params_dirty = false;
    bool do_reset = false;
    float4 a4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if constexpr (ONLY_RESET) {
        if (active) {
            sw_load(d, k, g);
            const int4 ev = d.E[e];
            tick = ev.x; task = ev.y; env_episode = (uint32_t)ev.z;
        }
    } else {
        cp_async_wait<0>(); // this thread's staged inputs of this tile have landed
        if (ring_staged) env_sync(); // ... and the env's rings, which its threads fetched together
        if (active) {
            int4 ev;
            sw_load_staged(stage, t, g, a4, ev);
            tick = ev.x; task = ev.y; env_episode = (uint32_t)ev.z;
        }
        // (the next tile's prefetch is issued after phase 1: until then the stage still holds this tile's inputs)
    }
    if (active) s_trail.put(w0 + a, g.s[0], g.s[1], g.s[2]);

    if constexpr (!ONLY_RESET) {
        // ---- phase 1: every drone moves (R/drone_swarm.h:452-461); agents that leave the arena take their prepared respawn
        bool oob = false;
        float rpos[3] = {0.0f, 0.0f, 0.0f};
        float passed = 0.0f; // ring test of this move (race task), R/drone_swarm.h:465
        float rs3x = 0.0f;   // 13th prepared respawn parameter (rides with the prepared position)
        int rs_rank = -1;    // this drone's slot in the warp's respawn-parameter pool
        if (active) {
            tick = (tick + 1) % SWARM_HORIZON;
            float act[4];
            if constexpr (STRICT) {
                act[0] = xclamp(xf(a4.x), -1.0f, 1.0f).v; act[1] = xclamp(xf(a4.y), -1.0f, 1.0f).v;
                act[2] = xclamp(xf(a4.z), -1.0f, 1.0f).v; act[3] = xclamp(xf(a4.w), -1.0f, 1.0f).v;
            } else {
                act[0] = fminf(fmaxf(a4.x, -1.0f), 1.0f); act[1] = fminf(fmaxf(a4.y, -1.0f), 1.0f);
                act[2] = fminf(fmaxf(a4.z, -1.0f), 1.0f); act[3] = fminf(fmaxf(a4.w, -1.0f), 1.0f);
            }
            if (d.act_out) reinterpret_cast<float4 *>(d.act_out)[k] = make_float4(act[0], act[1], act[2], act[3]);
            DroneParams p = {g.p[0], g.p[1], g.p[2], g.p[3], g.p[4], g.p[5], g.p[6], g.p[7], g.p[8], g.p[9], g.p[10], g.p[11], g.p[12]};
            const float before[3] = {g.s[0], g.s[1], g.s[2]};
#if B2D_SW_RK4_LOOP
            if constexpr (STRICT) advance_body_strict<false>(g.s, p, act); // (inline divisions: the out-of-line copy is slower here, 315 vs 278 us)
            else advance_body_fast_loop(g.s, p, act);
#else
            advance_body<STRICT>(g.s, p, act);
#endif
#if B2D_SWARM_EXPERIMENT_DOUBLE_MATH
            {   // measurement aid: the rigid-body arithmetic twice, same memory traffic
                float s2[17];
#pragma unroll
                for (int m = 0; m < 17; m++) s2[m] = g.s[m];
                advance_body<STRICT>(s2, p, act);
                g.s[0] = fmaf(s2[0] + s2[6] + s2[12] + s2[16], 1e-30f, g.s[0]);
            }
#endif
            const bool race = task == SWARM_TASK_RACE;
            float ring[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (race) load_ring(e, g.ring_idx, ring);
            if constexpr (STRICT) {
                if (race) passed = gate_event<xf>(before, g.s, ring, -0.0f);
            } else {
                // decisions within a guard band of their threshold are re-taken in the reference's arithmetic
                const float wall = fminf(fminf(fabsf(fabsf(g.s[0]) - SW_GX), fabsf(fabsf(g.s[1]) - SW_GY)), fabsf(fabsf(g.s[2]) - SW_GZ));
                bool suspect = false;
                if (race) passed = gate_event_guarded(before, g.s, ring, -0.0f, SW_GUARD_PLANE, suspect);
                if (__builtin_expect(wall < SW_GUARD_WALL || suspect, 0)) {
                    float out[18], rg[6]; // (copies: the hot path's ring[] must not have its address taken)
#pragma unroll
                    for (int m = 0; m < 6; m++) rg[m] = ring[m];
                    sw_strict_move(d, k, a4, race ? rg : nullptr, out);
#pragma unroll
                    for (int m = 0; m < 17; m++) g.s[m] = out[m];
                    passed = out[17];
                    guard_hits += 1;
                }
            }
            oob = g.s[0] < -SW_GX || g.s[0] > SW_GX || g.s[1] < -SW_GY || g.s[1] > SW_GY || g.s[2] < -SW_GZ || g.s[2] > SW_GZ;
#if B2D_SWARM_EXPERIMENT_NO_RESPAWN
            oob = false; // measurement aid: nobody leaves the arena (no respawn draws, no episode ends by OOB)
#endif
            sw_move_target(g.tpos, g.tvel);
            if (__builtin_expect(oob, 0)) {
                if (inject) {
                    rpos[0] = pay_agent[13]; rpos[1] = pay_agent[14]; rpos[2] = pay_agent[15];
                } else {
                    g.respawns += 1u;
                    const float4 r3 = stage[14 * SWARM_BLOCK + t]; // prefetched with the state
                    rpos[0] = r3.y; rpos[1] = r3.z; rpos[2] = r3.w;
                    rs3x = r3.x;
                }
            }
            if (oob) s_trail.put(w0 + A + a, rpos[0], rpos[1], rpos[2]);
            else s_trail.put(w0 + A + a, g.s[0], g.s[1], g.s[2]);
        }
        {   // drones that consumed their prepared respawn queue its regeneration (warp-uniform code)
            const bool want = active && oob && !inject;
            const unsigned int m = __ballot_sync(0xffffffffu, want);
            if (m) {
                const int base = s_rcnt[warp];
                const int rank = __popc(m & ((1u << lane) - 1u));
                if (want) {
                    s_rlist[warp][base + rank] = make_int2(k, (int)(g.respawns + 1u));
#if B2D_SW_RS_POOL
                    if (rank < B2D_SW_RS_POOL) {
                        rs_rank = rank;
                        const size_t ld = d.ld;
#pragma unroll
                        for (int q = 0; q < 3; q++) cp_async16(&s_rspool[warp][rank][q], &d.RS[q * ld + k]);
                    }
#endif
                }
                __syncwarp();
                if (lane == 0) s_rcnt[warp] = base + __popc(m);
                __syncwarp();
            }
        }
#if B2D_SW_RS_POOL
        cp_async_commit(); // group: this tile's respawn parameters (awaited in phase 2 by the drones that need them)
#endif
        {   // the next tile's inputs stream in while phases 2-4 of this tile run (stage slots are thread-private)
            const int tn = tile + gridDim.x, en = tn * d.epc + le;
            if (tn < ntiles && le < d.epc && en < d.n) {
                sw_prefetch(d, stage, t, en, en * A + a);
                if (ring_staged) sw_prefetch_rings(d, rstage, rbuf ^ 1, le, a, en);
            }
            cp_async_commit();
        }
        env_sync();

        // ---- phase 2: rewards, ring logic, respawn bookkeeping (R/drone_swarm.h:463-491)
        bool ended = false, horizon = false;
        if (active) {
            const float self[3] = {g.s[0], g.s[1], g.s[2]};
            float nd = 0.0f;
            if (A > 1) nd = sw_reward_distance<STRICT, SCAN_UNROLL>(&s_trail.x[w0], A, a, self, guard_hits);
            if (task == SWARM_TASK_RACE) {
                reward = sw_reward<STRICT>(g, self, true, A, nd);
                if (passed > 0.0f) {
                    float ring[6];
                    g.ring_idx = (g.ring_idx + 1) % d.R;
                    load_ring(e, g.ring_idx, ring);
                    g.tpos[0] = ring[0]; g.tpos[1] = ring[1]; g.tpos[2] = ring[2];
                    g.tvel[0] = g.tvel[1] = g.tvel[2] = 0.0f;
                    sw_reward<STRICT>(g, self, true, A, nd);
                }
                reward = __fadd_rn(reward, passed);
            } else {
                reward = sw_reward<STRICT>(g, self, true, A, nd);
            }
            g.ep_ret = __fadd_rn(g.ep_ret, reward);
            horizon = tick >= SWARM_HORIZON - 1;
            ended = oob || horizon;
        }
#if !B2D_SWARM_EXPERIMENT_NO_STATS
        {   // add_log (R/drone_swarm.h:91-105) for the warp's finished episodes, as warp-uniform code: every sum is an
            // exact integer sum of __float2ll_rn(x * 2^20) terms, so neither the order nor the grouping matters
            const unsigned int m_end = __ballot_sync(0xffffffffu, ended);
            const unsigned int m_ring = __ballot_sync(0xffffffffu, active && task == SWARM_TASK_RACE && passed > 0.0f);
            if ((m_end | m_ring) != 0u) { // every other warp-tile at A = 16
                const unsigned int m_oob = __ballot_sync(0xffffffffu, ended && oob);
                long long f[4] = {0ll, 0ll, 0ll, 0ll};
                int ilen = 0;
                if (ended) {
                    const float len = (float)g.ep_len;
                    ilen = g.ep_len;
                    f[0] = __float2ll_rn(g.score * 1048576.0f);
                    f[1] = __float2ll_rn(g.ep_ret * 1048576.0f);
                    // (a zero numerator sends the IEEE division down its slow path: most episodes have no collision)
                    f[2] = g.collisions == 0.0f ? 0ll : __float2ll_rn((g.collisions / len) * 1048576.0f);
                    f[3] = g.score == 0.0f ? 0ll : __float2ll_rn((g.score / len) * 1048576.0f);
                }
                long long sum[4] = {0ll, 0ll, 0ll, 0ll};
                int slen = 0;
                for (unsigned int m = m_end; m; m &= m - 1u) { // one finished episode at a time: there are rarely more than two
                    const int src = __ffs((int)m) - 1;
#pragma unroll
                    for (int q = 0; q < 4; q++) sum[q] += __shfl_sync(0xffffffffu, f[q], src);
                    slen += __shfl_sync(0xffffffffu, ilen, src);
                }
                if (lane == 0) {
                    long long *acc = s_wacc[warp];
                    acc[FACC_SCORE] += sum[0];
                    acc[FACC_RETURN] += sum[1];
                    acc[FACC_COLLISION] += sum[2];
                    acc[FACC_PERF] += sum[3];
                    acc[FACC_LENGTH] += (long long)slen << 20;
                    acc[FACC_N] += (long long)__popc(m_end) << 20;
                    acc[FACC_OOB] += (long long)__popc(m_oob) << 20;
                    acc[FACC_RINGS] += (long long)__popc(m_ring) << 20;
                }
            }
        }
#endif
        if (active) {
            if (__builtin_expect(ended, 0)) {
                terminal = 1;
                g.ep_len = 0;
                g.ep_ret = 0.0f;
            }
            if (__builtin_expect(oob, 0)) {
                reward = __fsub_rn(reward, 1.0f);
                float rp[13];
                if (inject) {
#pragma unroll
                    for (int m = 0; m < 13; m++) rp[m] = pay_agent[m];
                } else {
                    const size_t ld = d.ld;
                    float4 r0, r1, r2;
#if B2D_SW_RS_POOL
                    if (rs_rank >= 0) {
                        cp_async_wait<1>(); // all but the next tile's inputs: this drone's three pool words have landed
                        r0 = s_rspool[warp][rs_rank][0]; r1 = s_rspool[warp][rs_rank][1]; r2 = s_rspool[warp][rs_rank][2];
                    } else
#endif
                    {
                        r0 = __ldcg(&d.RS[0 * ld + k]); r1 = __ldcg(&d.RS[1 * ld + k]); r2 = __ldcg(&d.RS[2 * ld + k]);
                    }
                    rp[0] = r0.x; rp[1] = r0.y; rp[2] = r0.z; rp[3] = r0.w; rp[4] = r1.x; rp[5] = r1.y; rp[6] = r1.z; rp[7] = r1.w;
                    rp[8] = r2.x; rp[9] = r2.y; rp[10] = r2.z; rp[11] = r2.w;
                    rp[12] = rs3x;
                }
                sw_respawn_state(g, rp, rpos);
                params_dirty = true;
                float nd2 = 0.0f;
                if (A > 1 && task != SWARM_TASK_RACE) {
                    int hits = 0;
                    nd2 = sw_reward_distance_cold<STRICT>(&s_trail.x[w0], A, a, rpos[0], rpos[1], rpos[2], &hits);
                    guard_hits += hits;
                }
                sw_reward<STRICT>(g, rpos, task != SWARM_TASK_RACE, A, nd2);
            }
            do_reset = horizon;
        }
    } else {
        if (active) {
            s_trail.put(w0 + A + a, g.s[0], g.s[1], g.s[2]);
            g.respawns = 0u;
            do_reset = true;
        }
    }

    // ---- phase 3: env-wide reset (R/drone_swarm.h:401-443), every 1023 ticks for all agents of the env at once.
    // Out of line (sw_reset_phase): it is a third of the kernel's code and runs once per 1023 steps; the agent
    // travels through a local copy so that the hot path's registers never have their address taken.
    const int cta_reset = env_any(do_reset);
    if (__builtin_expect(cta_reset, 0)) {
        SwResetCtx c;
        c.g = g; c.tick = tick; c.task = task; c.env_episode = env_episode; c.do_reset = do_reset; c.params_dirty = params_dirty;
        c.guard_hits = 0;
        sw_reset_phase<STRICT, ONLY_RESET>(d, &c, &s_trail, s_ring0, ring_staged ? rstage + ((size_t)(rbuf * d.epc + le) * d.R) * 8 : nullptr,
                                           le, a, w0, e, pay_agent, pay_env, sync_mode, grp_size);
        g = c.g; tick = c.tick; task = c.task; env_episode = c.env_episode; params_dirty = c.params_dirty;
        guard_hits += c.guard_hits;
    }
    if (active) {
        s_now.put(w0 + a, g.s[0], g.s[1], g.s[2]);
        s_now.put(w0 + A + a, g.s[0], g.s[1], g.s[2]);
    }
#if B2D_SW_OBS_BULK
    if (t == grp_first) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // the previous tile's rows have left the staging area
#endif
    env_sync();

    // ---- phase 4: state out, observations (R/drone_swarm.h:131-217) staged through shared memory
    if (active) {
        sw_store(d, k, g, params_dirty);
        if (a == 0) d.E[e] = make_int4(tick, task, (int)env_episode, 0);
        if constexpr (!ONLY_RESET) {
            d.rew[k] = reward;
            d.term[k] = (unsigned char)terminal;
        }
        const float self[3] = {g.s[0], g.s[1], g.s[2]};
        float near[3] = {0.0f, 0.0f, 0.0f};
        if (A > 1) sw_obs_neighbour<STRICT, SCAN_UNROLL>(&s_now.x[w0], A, a, self, near, guard_hits);
        float ring[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (task == SWARM_TASK_RACE) load_ring(e, g.ring_idx, ring);
        sw_observe<STRICT>(g, A, near, task == SWARM_TASK_RACE, ring, s_obs + t * SWARM_OBS, 1);
    }
#if B2D_SW_OBS_BULK
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // this thread's row becomes visible to the bulk-copy engine
#endif
    env_sync();
    {   // each group stores its own rows of the tile (contiguous in shared and in global memory)
        const int rows_here = min(d.epc, d.n - tile * d.epc) * A;
        const int rows_mine = max(0, min(grp_size, rows_here - grp_first));
        const size_t row0 = (size_t)tile * d.epc * A + grp_first;
        float *gobs = d.obs + row0 * SWARM_OBS;
        const float *sobs = s_obs + (size_t)grp_first * SWARM_OBS;
        const int tg = t - grp_first;
#if B2D_SW_OBS_BULK
        // One asynchronous bulk copy, shared -> global, issued by the group's first thread: the copy engine moves the
        // 164 B x rows while the threads go on (the per-lane LDS.128 -> STG loop it replaces was 4 % of the step's
        // instructions and 9 % of its stall samples at A = 16: every iteration waited for its shared-memory load).  The
        // staging area is next written in the next tile's phase 4; the issuing thread waits for the engine to have READ
        // it before the env's barrier that precedes those writes.
        if (rows_mine > 0 && (rows_mine & 3) == 0 && (row0 & 3) == 0 && (grp_first & 3) == 0 && (reinterpret_cast<uintptr_t>(gobs) & 15) == 0) {
            if (tg == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gobs),
                             "r"((uint32_t)__cvta_generic_to_shared(sobs)), "r"((uint32_t)(rows_mine * SWARM_OBS * 4))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else
#endif
        if ((rows_mine & 3) == 0 && (row0 & 3) == 0 && (grp_first & 3) == 0) { // 16-byte aligned: float4 stores
            const float4 *src = reinterpret_cast<const float4 *>(sobs);
            float4 *dst = reinterpret_cast<float4 *>(gobs);
            const int n4 = rows_mine * SWARM_OBS / 4; // <= 10.25 * grp_size
            // (hoisting the loads of several iterations, or unrolling: measured slower, profiles/README.md)
#pragma unroll 1
            for (int m = tg; m < n4; m += grp_size) __stcs(&dst[m], src[m]);
        } else {
            for (int m = tg; m < rows_mine * SWARM_OBS; m += grp_size) __stcs(&gobs[m], sobs[m]);
        }
    }
    if constexpr (!ONLY_RESET) {
        if (t == 0 && tile == 0) atomicAdd(&d.ctl->ctas_done, 1u);
        rbuf ^= 1;
        // a full warp's worth of consumed respawn slots: regenerate them with every lane busy
        __syncwarp();
        const int cnt = s_rcnt[warp];
        if (cnt >= 32) {
            sw_refill_pass(d, s_rlist[warp], 32, lane);
            __syncwarp();
            const int2 keep = lane + 32 < cnt ? s_rlist[warp][lane + 32] : make_int2(0, 0);
            __syncwarp();
            s_rlist[warp][lane] = keep;
            if (lane == 0) s_rcnt[warp] = cnt - 32;
            __syncwarp();
        }
    }
    // the next tile's first barrier (after its phase 1) separates this tile's readers of the
    // shared arrays from their next writers, except the old / fin arrays, which are last read before this
    // tile's final barriers
  }
#if B2D_SW_OBS_BULK
  if (t == grp_first) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // the last tile's rows are in global memory
#endif
  if constexpr (!ONLY_RESET) {
      cp_async_wait<0>();
      __syncwarp();
      sw_refill_pass(d, s_rlist[warp], s_rcnt[warp], lane); // the rest of this warp's list
      if (guard_hits) atomicAdd(&s_guard, guard_hits);
      __syncthreads(); // every warp's statistics are in
      if (t < 8) {
          long long sum = 0ll;
#pragma unroll
          for (int w = 0; w < SWARM_BLOCK / 32; w++) sum += s_wacc[w][t];
          if (sum != 0ll) atomicAdd(&d.ctl->facc[t], (double)sum * (1.0 / 1048576.0));

      }
      if (t == 0 && s_guard != 0) atomicAdd(&d.ctl->guard_replays, (unsigned long long)s_guard);
#if B2D_SW_OVERLAP
      if (t == 0 && d.chain) { // everything this CTA wrote (the barrier above) is visible before the flag
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(d.chain + blockIdx.x), "r"(d.seq) : "memory");
      }
#endif
  }
}

// snapshot + clear for vec_log: out[0..7] = the float sums in 2^-20 fixed point, so that the
// cross-rank reduction is the same integer all-reduce the race env uses
__global__ void swarm_log_snapshot_kernel(Ctl *ctl, long long *out) {
    if (threadIdx.x < 8) {
        out[threadIdx.x] = __double2ll_rn(ctl->facc[threadIdx.x] * 1048576.0);
        ctl->facc[threadIdx.x] = 0.0;
    } else if (threadIdx.x < 16) {
        out[threadIdx.x] = 0;
    }
}

// state blobs: per env [A][47] agent blobs (oracle/ref_shim_swarm.c layout) then [2 + 6R] env blob
__global__ void swarm_pack_kernel(const SwarmDev d, const int *ids, int n, float *blobs) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n * d.A) return;
    const int slot = j / d.A, a = j - slot * d.A;
    const int e = ids ? ids[slot] : slot;
    const int blob = d.A * SWARM_AGENT_BLOB + 2 + 6 * d.R;
    SwarmAgent g;
    sw_load(d, e * d.A + a, g);
    float *b = blobs + (size_t)slot * blob + (size_t)a * SWARM_AGENT_BLOB;
    for (int m = 0; m < 17; m++) b[m] = g.s[m];
    for (int m = 0; m < 13; m++) b[17 + m] = g.p[m];
    for (int m = 0; m < 3; m++) { b[30 + m] = g.spawn[m]; b[33 + m] = g.tpos[m]; b[36 + m] = g.tvel[m]; }
    b[39] = g.last_abs; b[40] = g.last_tgt; b[41] = g.last_col; b[42] = g.ep_ret; b[43] = g.collisions;
    b[44] = (float)g.ep_len; b[45] = g.score; b[46] = (float)g.ring_idx;
    if (a == 0) {
        float *eb = blobs + (size_t)slot * blob + (size_t)d.A * SWARM_AGENT_BLOB;
        const int4 ev = d.E[e];
        eb[0] = (float)ev.x; eb[1] = (float)ev.y;
        for (int r = 0; r < d.R; r++) sw_load_ring(d, e, r, eb + 2 + 6 * r);
    }
}

__global__ void swarm_unpack_kernel(const SwarmDev d, const int *ids, int n, const float *blobs) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n * d.A) return;
    const int slot = j / d.A, a = j - slot * d.A;
    const int e = ids ? ids[slot] : slot;
    const int blob = d.A * SWARM_AGENT_BLOB + 2 + 6 * d.R;
    const float *b = blobs + (size_t)slot * blob + (size_t)a * SWARM_AGENT_BLOB;
    SwarmAgent g;
    const int k = e * d.A + a;
    g.respawns = __float_as_uint(d.T[k].y);
    for (int m = 0; m < 17; m++) g.s[m] = b[m];
    for (int m = 0; m < 13; m++) g.p[m] = b[17 + m];
    for (int m = 0; m < 3; m++) { g.spawn[m] = b[30 + m]; g.tpos[m] = b[33 + m]; g.tvel[m] = b[36 + m]; }
    g.last_abs = b[39]; g.last_tgt = b[40]; g.last_col = b[41]; g.ep_ret = b[42]; g.collisions = b[43];
    g.ep_len = (int)b[44]; g.score = b[45]; g.ring_idx = (int)b[46];
    sw_store(d, k, g, true);
    if (a == 0) {
        const float *eb = blobs + (size_t)slot * blob + (size_t)d.A * SWARM_AGENT_BLOB;
        d.E[e] = make_int4((int)eb[0], (int)eb[1], d.E[e].z, 0);
        for (int r = 0; r < d.R; r++) {
            const float *gg = eb + 2 + 6 * r;
            d.G0[(size_t)r * d.n + e] = make_float4(gg[0], gg[1], gg[2], gg[3]);
            d.G1[(size_t)r * d.n + e] = make_float2(gg[4], gg[5]);
        }
    }
}

// observations recomputed from the current state (after put_state)
template <bool STRICT>
__global__ void __launch_bounds__(SWARM_BLOCK) swarm_observe_kernel(const __grid_constant__ SwarmDev d) {
    __shared__ __align__(16) SwarmWin s_now;
    const int t = threadIdx.x, A = d.A;
    const int le = t / A, a = t - le * A, e = blockIdx.x * d.epc + le, w0 = le * 3 * A;
    const bool active = le < d.epc && e < d.n;
    SwarmAgent g;
    if (active) {
        sw_load(d, e * A + a, g);
        s_now.put(w0 + a, g.s[0], g.s[1], g.s[2]);
        s_now.put(w0 + A + a, g.s[0], g.s[1], g.s[2]);
    }
    __syncthreads();
    if (!active) return;
    const int task = d.E[e].y;
    const float self[3] = {g.s[0], g.s[1], g.s[2]};
    float near[3] = {0.0f, 0.0f, 0.0f};
    if (A > 1) sw_nearest_strict(&s_now.x[w0], A, a, self, near);
    float ring[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (task == SWARM_TASK_RACE) sw_load_ring(d, e, g.ring_idx, ring);
    sw_observe<STRICT>(g, A, near, task == SWARM_TASK_RACE, ring, d.obs + (size_t)(e * A + a) * SWARM_OBS, 1);
}

// ---------------------------------------------------------------- host-side launch helpers
static inline int swarm_grid(const SwarmDev &d) { return (d.n + d.epc - 1) / d.epc; }

static inline void swarm_vec_reset(SwarmDev &d, uint64_t seed, cudaStream_t st, long long *launches) {
    d.key0 = (uint32_t)seed;
    d.key1 = (uint32_t)(seed >> 32);
    swarm_kernel<true, true><<<swarm_grid(d), SWARM_BLOCK, SWARM_BLOCK * SWARM_OBS * 4, st>>>(d);
    swarm_fill_slots_kernel<<<(d.rows + 127) / 128, 128, 0, st>>>(d); // respawn counts are zero again: slot = first respawn
    *launches += 2;
}

// persistent step grid: every resident CTA slot of the handle's device, or one CTA per tile when there
// are fewer tiles.  Called once per handle at create time with the handle's device current: the
// dynamic shared memory opt-in is per device, and so are the SM count and the occupancy.
static inline int swarm_step_smem(const SwarmDev &d) { return SW_DYN_SMEM + sw_ring_stage_bytes(d.epc, d.R); }

constexpr int SW_COMPACT_SCAN_MAX_A = 32; // see swarm_kernel: SCAN_UNROLL
typedef void (*swarm_step_fn)(SwarmDev);
static inline swarm_step_fn swarm_step_kernel_for(const SwarmDev &d, int math) {
    if (math == 1) return swarm_kernel<true, false>;
    return d.A <= SW_COMPACT_SCAN_MAX_A ? swarm_kernel<false, false, 2> : swarm_kernel<false, false, 4>;
}

static inline int swarm_step_setup(const SwarmDev &d, int device, int grid_out[2]) {
    int sms = 0, per_sm[2] = {0, 0};
    const int smem = swarm_step_smem(d);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    for (int m = 0; m < 2; m++) {
        const swarm_step_fn fn = swarm_step_kernel_for(d, m);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SW_DYN_SMEM + SW_RING_STAGE_MAX) != cudaSuccess ||
            cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[m], fn, SWARM_BLOCK, smem) != cudaSuccess)
            return -1;
    }
    const int tiles = swarm_grid(d);
    for (int m = 0; m < 2; m++) {
        if (per_sm[m] < 1) return -1;
        const int slots = sms * per_sm[m];
        grid_out[m] = tiles < slots ? tiles : slots;
    }
    return 0;
}

// overlap: this launch directly follows step launch seq - 1 of the same handle on the same stream (api.cu step_impl)
static inline cudaError_t swarm_vec_step(SwarmDev &dev, const float *actions, int math, int grid, cudaStream_t st, long long *launches,
                                         unsigned int seq, bool overlap, int tile_begin = 0, int tile_end = -1) {
    SwarmDev d = dev;
    if (actions) d.act_in = actions;
    d.seq = seq;
    d.chain_wait = overlap ? 1 : 0;
    d.tile_begin = tile_begin;
    d.tile_end = tile_end;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(SWARM_BLOCK);
    cfg.dynamicSmemBytes = (size_t)swarm_step_smem(d);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    *launches += 1;
    return cudaLaunchKernelEx(&cfg, swarm_step_kernel_for(d, math), d);
}

static inline void swarm_observe_launch(SwarmDev &d, cudaStream_t st) {
    swarm_observe_kernel<true><<<swarm_grid(d), SWARM_BLOCK, 0, st>>>(d);
}

// averaging of vec_log for the swarm (EB:588-591 + DS/binding.c:13-23): sums arrive as doubles
static inline void swarm_log_finish(const long long *a, float *out) {
    double f[8];
    for (int k = 0; k < 8; k++) f[k] = (double)a[k] / 1048576.0;
    const double n = f[FACC_N];
    for (int k = 0; k < 9; k++) out[k] = 0.0f;
    if (n == 0.0) return;
    out[0] = (float)(f[FACC_RETURN] / n);
    out[1] = (float)(f[FACC_LENGTH] / n);
    out[2] = (float)(f[FACC_RINGS] / n);
    out[3] = (float)(f[FACC_COLLISION] / n);
    out[4] = (float)(f[FACC_OOB] / n);
    out[5] = 0.0f;
    out[6] = (float)(f[FACC_SCORE] / n);
    out[7] = (float)(f[FACC_PERF] / n);
    out[8] = (float)n;
}

} // namespace b2d
