// rollout_kernels.cuh -- the on-device rollout as ONE kernel for sm_100a (BASELINE configs[3]).
//
// Reference: PuffeRL.evaluate (pufferlib/pufferl.py:214-314) stepping a vectorised DroneRace env: per step
// recv (obs, reward, done), `policy.forward_eval` of pufferlib.models.Default (models.py:41-98), `sample_logits`
// on Normal(mean, exp(logstd)) (pytorch.py:189-199), the experience stores (pufferl.py:258-281), the action clip
// (pufferl.py:292-294), send -> vec_step -> c_step (drone_race.h:156-208).  Here the whole K-step loop of a block
// of 128 envs runs inside one CTA:
//
//   * one thread per env; the env's state (17 floats, 13 parameters, current ring, counters) lives in REGISTERS for
//     all K steps -- it is read from and written to HBM once per rollout, not once per step;
//   * the policy's two Linear layers run on the 5th-generation tensor cores: per step the CTA's 128 observation
//     rows are written to shared memory in the canonical K-major UMMA layout (8x16-byte core matrices), one elected
//     thread issues `tcgen05.mma.kind::tf32` (M = 128 envs, N = 128 hidden units, K = 32: 29 observations, the
//     bias as two TF32 halves against constant-1 columns, one zero) into a 128-column TMEM accumulator, every thread
//     reads ITS OWN row back with `tcgen05.ld.32x32b` (TMEM lane = env = thread), applies the exact-erf GELU in
//     registers and stores the activations back into the same TMEM columns (`tcgen05.st`), where they are the
//     A operand of the second GEMM (`tcgen05.mma` with A in TMEM: M = 128, N = 16 = 4 means + value + padding,
//     K = 128); weights are staged in shared memory once per CTA;
//   * sampling (Philox4x32-10 + Box-Muller keyed by (seed; global row, call number)), log-prob, the action clip and
//     the env step follow in registers; per env-step only the experience row leaves the SM: observation 116 B,
//     action 16 B, log-prob, value, reward, terminal 4 B each = 148 B (the two-kernel form moves 373 + 285 B).
//
// TF32 semantics: the reference runs its policy under torch.set_float32_matmul_precision('high') (pufferl.py:55),
// i.e. TF32 GEMMs with float32 accumulation.  Observations and weights are rounded to TF32 to nearest (ties away,
// cvt.rna); the hidden activations enter the second GEMM by truncation (the tensor core ignores the low 13
// mantissa bits of a 32-bit TF32 operand) -- oracle/policy_oracle.py restates exactly this (hidden="truncate").
//
// The env arithmetic is the same code the step kernel runs (advance_body, gate_event, race_observe, the Philox
// episode generator, the near-threshold guard), so a K-step launch equals K launches of one step bit for bit.
#pragma once
#include "policy_kernels.cuh"
#include "race_kernels.cuh"

namespace b2d {

#ifndef B2D_RO_TIMING
#define B2D_RO_TIMING 0 // measurement aid: clock64 sums per phase of the step loop, printed by b2d_vec_close
#endif
#if B2D_RO_TIMING
#define RO_TICK(n) { const long long now_ = clock64(); tm[n] += now_ - tprev; tprev = now_; }
#else
#define RO_TICK(n)
#endif

constexpr int RO_THREADS = 128;     // = envs per CTA = rows of the MMA tile = TMEM lanes
#ifndef B2D_RO_CTAS_PER_SM
#define B2D_RO_CTAS_PER_SM 3
#endif
constexpr int RO_CTAS_PER_SM = B2D_RO_CTAS_PER_SM;   // 3 x (128 + 32) TMEM columns of the SM's 512
constexpr int RO_HIDDEN = 128;
constexpr int RO_K1 = 32;           // encoder K: 29 observations + bias hi + bias lo + 0
constexpr int RO_N2 = 16;           // head outputs padded to the smallest N of an M = 128 MMA
constexpr int RO_B1_BYTES = (RO_K1 / 4) * RO_HIDDEN * 16;   // [8 chunks][128 n] float4
constexpr int RO_B2_BYTES = (RO_HIDDEN / 4) * RO_N2 * 16;   // [32 chunks][16 n] float4
constexpr int RO_A1_BYTES = (RO_K1 / 4) * RO_THREADS * 16;  // [8 chunks][128 rows] float4
constexpr int RO_TILE_BYTES = 32 * RACE_OBS * 4;            // per-warp staging of observation rows
#ifndef B2D_RO_GELU_DEG4
#define B2D_RO_GELU_DEG4 1
#endif
#ifndef B2D_RO_BANK_PREFETCH
#define B2D_RO_BANK_PREFETCH 1
#endif
// per lane: the bank entry of the env's NEXT episode (6 x 16 B), fetched by cp.async when the current episode
// begins, so that an episode end -- some lane of the CTA has one in 96 % of the steps, and the whole CTA waits for
// it at the step's next barrier -- never waits for DRAM (it was 6 % long-scoreboard + 12 % barrier stall samples)
constexpr int RO_NEXT_BYTES = B2D_RO_BANK_PREFETCH ? 6 * RO_THREADS * 16 : 0;
constexpr int RO_SMEM_USED = RO_B1_BYTES + RO_B2_BYTES + RO_A1_BYTES + 4 * RO_TILE_BYTES + RO_NEXT_BYTES;
// request enough shared memory that exactly RO_CTAS_PER_SM CTAs fit on an SM: a fourth CTA would find no TMEM
// columns left and sit in tcgen05.alloc until another CTA exits
constexpr int RO_SMEM_BYTES = RO_CTAS_PER_SM >= 4 ? 56 * 1024 : 72 * 1024;
static_assert(RO_SMEM_USED <= RO_SMEM_BYTES, "rollout kernel shared memory");

struct RolloutArgs {
    RaceDev d;
    const float *enc_w, *enc_b, *mean_w, *mean_b, *logstd, *value_w, *value_b; // torch nn.Linear layouts, hidden = 128
    float *st_obs, *st_act, *st_logp, *st_rew, *st_term, *st_val;              // experience, time-major [K][n][...]
    float *env_act;        // [n][4] the env's action buffer: receives the clipped action of the last step
    int horizon;
    uint32_t row_id_base, seed_lo, seed_hi;
    unsigned int *counter; // [2]: policy calls completed, CTAs arrived (same protocol as b2d_policy_act)
    int deterministic;
};

// ---------------------------------------------------------------- PTX wrappers (tcgen05 / TMEM / mbarrier)
__device__ __forceinline__ uint32_t ro_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ro_mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ro_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ro_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void ro_mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "RO_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra RO_DONE;\n\t"
        "bra RO_WAIT;\n\t"
        "RO_DONE:\n\t"
        "}" ::"r"(ro_smem_u32(bar)), "r"(parity)
        : "memory");
}
// generic-proxy writes to shared memory (the operand tiles) become visible to the async proxy (the tensor core)
__device__ __forceinline__ void ro_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void ro_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ro_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ro_tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ro_tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ro_tmem_alloc(uint32_t *slot, int cols) { // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ro_smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void ro_tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ro_tmem_free(uint32_t addr, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// completion of all tcgen05.mma issued so far by this thread -> one arrival on the mbarrier
__device__ __forceinline__ void ro_tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ro_smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, layout type INTERLEAVE): the
// operand is a grid of 8-row x 16-byte core matrices, each 128 contiguous bytes; SBO = distance between core
// matrices adjacent in M/N (8-row groups), LBO = distance between the two 16-byte K chunks one K = 8 (tf32)
// instruction consumes.  Addresses and offsets are in units of 16 bytes; bit 46 = descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t ro_umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2),
// both K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28
__host__ __device__ constexpr uint32_t ro_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void ro_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: lane = row, 8 consecutive 32-bit columns per K = 8 instruction)
__device__ __forceinline__ void ro_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// this thread's lane (= row), 32 consecutive 32-bit columns starting at taddr
__device__ __forceinline__ void ro_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ro_tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void ro_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ro_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void ro_gelu16(uint32_t (&r)[16]) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        #if B2D_RO_GELU_DEG4
        const float2 g = gelu2_tf32(make_float2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1])));
#else
        const float2 g = gelu2(make_float2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1])));
#endif
        r[2 * q] = __float_as_uint(g.x);
        r[2 * q + 1] = __float_as_uint(g.y);
    }
}
__device__ __forceinline__ void ro_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// ---------------------------------------------------------------- the env in registers
struct RaceRegs {
    float s[17];
    DroneParams p;
    float ring[6];
    int tick, ring_idx, ring_ext;
    float ep_ret;
    uint32_t episode;
};

// rare paths, out of line, exchanging through small local arrays so that the env registers never have their
// address taken
__device__ __noinline__ void ro_generate_episode(const RaceDev &d, int i, uint32_t episode, float *out /*[22]*/) {
    float p[13], spawn[3], ring0[6];
    race_generate_episode(d, i, episode, p, spawn, ring0);
#pragma unroll
    for (int k = 0; k < 13; k++) out[k] = p[k];
#pragma unroll
    for (int k = 0; k < 3; k++) out[13 + k] = spawn[k];
#pragma unroll
    for (int k = 0; k < 6; k++) out[16 + k] = ring0[k];
}

// in[0:17] state, [17:30] params, [30:36] ring, [36:40] raw action; out[0:17] state, [17] oob, [18] gate
__device__ __noinline__ void ro_strict_replay(const float *in, float *out) {
    float s[17];
#pragma unroll
    for (int k = 0; k < 17; k++) s[k] = in[k];
    const DroneParams p = {in[17], in[18], in[19], in[20], in[21], in[22], in[23], in[24], in[25], in[26], in[27], in[28], in[29]};
    const float ring[6] = {in[30], in[31], in[32], in[33], in[34], in[35]};
    const float act[4] = {xclamp(xf(in[36]), -1.0f, 1.0f).v, xclamp(xf(in[37]), -1.0f, 1.0f).v, xclamp(xf(in[38]), -1.0f, 1.0f).v,
                          xclamp(xf(in[39]), -1.0f, 1.0f).v};
    const float before[3] = {s[0], s[1], s[2]};
    advance_body_strict(s, p, act);
    const bool oob = s[0] < -10.0f || s[0] > 10.0f || s[1] < -10.0f || s[1] > 10.0f || s[2] < -10.0f || s[2] > 10.0f;
#pragma unroll
    for (int k = 0; k < 17; k++) out[k] = s[k];
    out[17] = oob ? 1.0f : 0.0f;
    out[18] = oob ? 0.0f : gate_event<xf>(before, s, ring, -1.0f);
}

// One c_step (DR/drone_race.h:156-208) of the env in `e` on the raw action `a4`; leaves the observation of the
// (possibly new) episode in o[29].  Same decisions, same arithmetic as race_step_kernel.
// the bank entry of episode `episode` of env i -> this lane's six staging words (slot q at nxt[q * RO_THREADS])
__device__ __forceinline__ void ro_prefetch_bank(const RaceDev &d, int i, uint32_t episode, float4 *nxt) {
    const float4 *b = d.bank + ((size_t)i * RACE_BANK_SLOTS + (episode % RACE_BANK_SLOTS)) * 6;
#pragma unroll
    for (int q = 0; q < 6; q++) cp_async16(&nxt[q * RO_THREADS], b + q);
    cp_async_commit();
}

template <bool STRICT>
__device__ __forceinline__ void ro_env_step(const RaceDev &d, int i, RaceRegs &e, float4 a4, float &reward, int &terminal,
                                            float (&o)[RACE_OBS], int *s_acc, bool last_step, int &score_last, float4 *nxt) {
    float act[4];
    if constexpr (STRICT) {
        act[0] = xclamp(xf(a4.x), -1.0f, 1.0f).v; act[1] = xclamp(xf(a4.y), -1.0f, 1.0f).v;
        act[2] = xclamp(xf(a4.z), -1.0f, 1.0f).v; act[3] = xclamp(xf(a4.w), -1.0f, 1.0f).v;
    } else {
        act[0] = fminf(fmaxf(a4.x, -1.0f), 1.0f); act[1] = fminf(fmaxf(a4.y, -1.0f), 1.0f);
        act[2] = fminf(fmaxf(a4.z, -1.0f), 1.0f); act[3] = fminf(fmaxf(a4.w, -1.0f), 1.0f);
    }
    e.tick += 1;
    float s0[17];
#pragma unroll
    for (int k = 0; k < 17; k++) s0[k] = e.s[k];
    advance_body<STRICT, B2D_RO_RK4_LOOP>(e.s, e.p, act);
    bool oob = e.s[0] < -10.0f || e.s[0] > 10.0f || e.s[1] < -10.0f || e.s[1] > 10.0f || e.s[2] < -10.0f || e.s[2] > 10.0f;
    float gate = 0.0f;
    if constexpr (STRICT) {
        if (!oob) gate = gate_event<xf>(s0, e.s, e.ring, -1.0f);
    } else {
        const float wall = fminf(fminf(fabsf(fabsf(e.s[0]) - 10.0f), fabsf(fabsf(e.s[1]) - 10.0f)), fabsf(fabsf(e.s[2]) - 10.0f));
        bool suspect = false;
        if (!oob) gate = gate_event_guarded(s0, e.s, e.ring, -1.0f, RACE_GUARD_PLANE, suspect);
        if (wall < RACE_GUARD_WALL || suspect) { // near-threshold guard: see race_strict_replay
            float in[40], out[19];
#pragma unroll
            for (int k = 0; k < 17; k++) in[k] = s0[k];
            in[17] = e.p.mass; in[18] = e.p.ixx; in[19] = e.p.iyy; in[20] = e.p.izz; in[21] = e.p.arm; in[22] = e.p.kt;
            in[23] = e.p.kad; in[24] = e.p.kd; in[25] = e.p.bd; in[26] = e.p.g; in[27] = e.p.mrpm; in[28] = e.p.kmot; in[29] = e.p.jmot;
#pragma unroll
            for (int k = 0; k < 6; k++) in[30 + k] = e.ring[k];
            in[36] = a4.x; in[37] = a4.y; in[38] = a4.z; in[39] = a4.w;
            ro_strict_replay(in, out);
#pragma unroll
            for (int k = 0; k < 17; k++) e.s[k] = out[k];
            oob = out[17] != 0.0f;
            gate = out[18];
            atomicAdd(&s_acc[ACC_SPARE], 1);
        }
    }
    reward = oob ? -1.0f : gate;
    e.ep_ret += reward;
    const bool passed = gate > 0.0f;
    e.ring_idx += passed ? 1 : 0;
    int cause = oob ? (int)ACC_OOB
                    : gate < 0.0f ? (int)ACC_COLLISION
                                  : e.tick == d.max_moves ? (int)ACC_TIMEOUT : e.ring_idx == d.max_rings ? (int)ACC_SPARE : -1;
#if B2D_EXPERIMENT_NO_RESET
    cause = -1; // measurement aid: episodes never end
#endif
    terminal = cause >= 0 ? 1 : 0;
    if (passed && cause < 0) { // the next ring becomes the current one
        float ring[6];
#pragma unroll
        for (int k = 0; k < 6; k++) ring[k] = e.ring[k];
        if (e.ring_ext) race_load_external_ring(d, i, e.ring_idx, ring);
        else race_next_ring(d, i, e.episode, e.ring_idx, ring);
#pragma unroll
        for (int k = 0; k < 6; k++) e.ring[k] = ring[k];
    }
    if (cause >= 0) {
        // add_log: DR/drone_race.h:61-70 (score == ring_idx at every call site)
        atomicAdd(&s_acc[ACC_N], 1);
        atomicAdd(&s_acc[ACC_RETURN], __float2int_rn(e.ep_ret));
        atomicAdd(&s_acc[ACC_LENGTH], e.tick);
        atomicAdd(&s_acc[ACC_RINGS], e.ring_idx);
        if (cause != ACC_SPARE) atomicAdd(&s_acc[cause], 1);
        if (last_step) score_last += e.ring_idx;
        // c_reset: the next episode of this env (Philox stream of the step kernel): from the episode bank when its
        // entry is there (race_bank_fill_kernel), generated in place otherwise
        float g[22];
        e.episode += 1u;
        bool banked = false;
        if (d.bank) {
#if B2D_RO_BANK_PREFETCH
            cp_async_wait<0>(); // this lane's staged entry (fetched when the episode that just ended began)
            const float4 b5 = nxt[5 * RO_THREADS];
            const float4 b0 = nxt[0 * RO_THREADS], b1 = nxt[1 * RO_THREADS], b2 = nxt[2 * RO_THREADS], b3 = nxt[3 * RO_THREADS],
                         b4 = nxt[4 * RO_THREADS];
#else
            const float4 *b = d.bank + ((size_t)i * RACE_BANK_SLOTS + (e.episode % RACE_BANK_SLOTS)) * 6;
            // all six words at once: the tag check must not put a second DRAM round trip behind the first (the whole
            // CTA waits for this lane at the step's next barrier)
            const float4 b5 = __ldcg(b + 5);
            const float4 b0 = __ldcg(b + 0), b1 = __ldcg(b + 1), b2 = __ldcg(b + 2), b3 = __ldcg(b + 3), b4 = __ldcg(b + 4);
#endif
            if (__float_as_uint(b5.z) == e.episode && __float_as_uint(b5.w) == (d.key0 ^ (d.key1 * 0x9E3779B9u) ^ 0xB2D0u)) {
                g[0] = b0.x; g[1] = b0.y; g[2] = b0.z; g[3] = b0.w; g[4] = b1.x; g[5] = b1.y; g[6] = b1.z; g[7] = b1.w;
                g[8] = b2.x; g[9] = b2.y; g[10] = b2.z; g[11] = b2.w; g[12] = b3.x; g[13] = b3.y; g[14] = b3.z; g[15] = b3.w;
                g[16] = b4.x; g[17] = b4.y; g[18] = b4.z; g[19] = b4.w; g[20] = b5.x; g[21] = b5.y;
                banked = true;
            }
        }
        if (!banked) ro_generate_episode(d, i, e.episode, g);
#if B2D_RO_BANK_PREFETCH
        if (d.bank) ro_prefetch_bank(d, i, e.episode + 1u, nxt); // the staged words have been consumed into g[]
#endif
        e.p = {g[0], g[1], g[2], g[3], g[4], g[5], g[6], g[7], g[8], g[9], g[10], g[11], g[12]};
#pragma unroll
        for (int k = 0; k < 17; k++) e.s[k] = 0.0f;
        e.s[6] = 1.0f;
        e.s[0] = g[13]; e.s[1] = g[14]; e.s[2] = g[15];
#pragma unroll
        for (int k = 0; k < 6; k++) e.ring[k] = g[16 + k];
        e.tick = 0;
        e.ring_idx = 0;
        e.ring_ext = 0;
        e.ep_ret = 0.0f;
    }
    race_observe<STRICT>(e.s, e.p.mrpm, e.ring, o);
}

// ---------------------------------------------------------------- the kernel
template <bool STRICT>
__global__ void __launch_bounds__(RO_THREADS, RO_CTAS_PER_SM) race_rollout_kernel(const __grid_constant__ RolloutArgs a) {
    extern __shared__ __align__(128) unsigned char ro_smem[];
    __shared__ __align__(8) unsigned long long s_mbar[2];
    __shared__ uint32_t s_tmem[2];
    __shared__ int s_acc[8];
    __shared__ int s_score;

    float4 *B1 = reinterpret_cast<float4 *>(ro_smem);
    float4 *B2 = reinterpret_cast<float4 *>(ro_smem + RO_B1_BYTES);
    float4 *A1 = reinterpret_cast<float4 *>(ro_smem + RO_B1_BYTES + RO_B2_BYTES);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *tile = reinterpret_cast<float *>(ro_smem + RO_B1_BYTES + RO_B2_BYTES + RO_A1_BYTES + warp * RO_TILE_BYTES);
    float *my_row = tile + lane * RACE_OBS;
    float4 *nxt = reinterpret_cast<float4 *>(ro_smem + RO_B1_BYTES + RO_B2_BYTES + RO_A1_BYTES + 4 * RO_TILE_BYTES) + tid;
    const RaceDev &d = a.d;

    // ---- weights -> shared memory, canonical K-major layout, TF32 (round to nearest, ties away)
    for (int idx = tid; idx < (RO_K1 / 4) * RO_HIDDEN; idx += RO_THREADS) {
        const int c = idx >> 7, n = idx & (RO_HIDDEN - 1);
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = 4 * c + q;
            float v = 0.0f;
            if (k < RACE_OBS) v = __uint_as_float(to_tf32(__ldg(&a.enc_w[n * RACE_OBS + k])));
            else if (k == RACE_OBS) v = __uint_as_float(to_tf32(__ldg(&a.enc_b[n])));               // bias, high half
            else if (k == RACE_OBS + 1) {                                                          // bias, low half
                const float b = __ldg(&a.enc_b[n]);
                v = __uint_as_float(to_tf32(b - __uint_as_float(to_tf32(b))));
            }
            w[q] = v;
        }
        B1[c * RO_HIDDEN + n] = make_float4(w[0], w[1], w[2], w[3]);
    }
    for (int idx = tid; idx < (RO_HIDDEN / 4) * RO_N2; idx += RO_THREADS) {
        const int c = idx >> 4, n = idx & (RO_N2 - 1);
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int h = 4 * c + q;
            const float v = n < 4 ? __ldg(&a.mean_w[n * RO_HIDDEN + h]) : (n == 4 ? __ldg(&a.value_w[h]) : 0.0f);
            w[q] = __uint_as_float(to_tf32(v));
        }
        B2[c * RO_N2 + n] = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (tid < 8) s_acc[tid] = 0;
    if (tid == 0) {
        s_score = 0;
        ro_mbar_init(&s_mbar[0], 1);
        ro_mbar_init(&s_mbar[1], 1);
        ro_fence_mbar_init();
    }
    if (warp == 0) { // TMEM: 128 columns (encoder accumulator, then the GELU activations) + 32 (head accumulator)
        ro_tmem_alloc(&s_tmem[0], 128);
        ro_tmem_alloc(&s_tmem[1], 32);
        ro_tmem_relinquish();
    }
    ro_fence_proxy_async();
    ro_tc_fence_before();
    __syncthreads();
    ro_tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16; // a warp reaches the 32 TMEM lanes of its quarter
    const uint32_t tm_h = s_tmem[0], tm_o = s_tmem[1];
    const uint32_t sA1 = ro_smem_u32(A1), sB1 = ro_smem_u32(B1), sB2 = ro_smem_u32(B2);
    constexpr uint32_t IDESC1 = ro_idesc_tf32(128, RO_HIDDEN), IDESC2 = ro_idesc_tf32(128, RO_N2);

    const unsigned int call0 = *reinterpret_cast<volatile unsigned int *>(a.counter);
    float bm[4], sd[4], ls[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        bm[c] = __ldg(&a.mean_b[c]);
        ls[c] = __ldg(&a.logstd[c]);
        sd[c] = expf(ls[c]);
    }
    const float bv = __ldg(&a.value_b[0]);
    const float lp0 = -(ls[0] + ls[1] + ls[2] + ls[3]) - 4.0f * 0.91893853320467274f;

    const int K = a.horizon;
    const int nchunks = (d.n + RO_THREADS - 1) / RO_THREADS;
    uint32_t phase = 0;
    int score_last = 0;
#if B2D_RO_TIMING
    long long tm[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif

    for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int i = chunk * RO_THREADS + tid;
        const bool valid = i < d.n;
        const int rows_w = min(32, d.n - (chunk * RO_THREADS + warp * 32)); // rows of this warp's tile that exist (may be <= 0)
        // ---- the env comes on chip: state, parameters, current ring, counters; last reward / terminal / observation
        RaceRegs e;
        float prev_rew = 0.0f, prev_term = 0.0f;
        float o[RACE_OBS];
        {
            float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0, q4 = q0, p0 = make_float4(1, 1, 1, 1), p1 = p0, p2 = p0,
                   c0 = q0, tl = make_float4(1, 0, 0, 0);
            q1.z = 1.0f;
            if (valid) {
                const float4 *hot = race_hot(d, i);
                const size_t st = race_slot_stride(d);
                q0 = hot[0 * st]; q1 = hot[1 * st]; q2 = hot[2 * st]; q3 = hot[3 * st]; q4 = hot[4 * st];
                p0 = hot[5 * st]; p1 = hot[6 * st]; p2 = hot[7 * st]; c0 = hot[8 * st]; tl = hot[9 * st];
                prev_rew = d.rew[i];
                prev_term = (float)d.term[i];
            }
            e.s[0] = q0.x; e.s[1] = q0.y; e.s[2] = q0.z; e.s[3] = q0.w; e.s[4] = q1.x; e.s[5] = q1.y; e.s[6] = q1.z; e.s[7] = q1.w;
            e.s[8] = q2.x; e.s[9] = q2.y; e.s[10] = q2.z; e.s[11] = q2.w; e.s[12] = q3.x; e.s[13] = q3.y; e.s[14] = q3.z; e.s[15] = q3.w;
            e.s[16] = q4.x;
            e.tick = __float_as_int(q4.y);
            const int ring_word = __float_as_int(q4.z);
            e.ring_idx = ring_word & RING_INDEX_MASK;
            e.ring_ext = ring_word & RING_EXTERNAL;
            e.ep_ret = q4.w;
            e.p = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y, p2.z, p2.w, tl.x};
            e.episode = __float_as_uint(tl.y);
            e.ring[0] = c0.x; e.ring[1] = c0.y; e.ring[2] = c0.z; e.ring[3] = c0.w; e.ring[4] = tl.z; e.ring[5] = tl.w;
#if B2D_RO_BANK_PREFETCH
            if (valid && d.bank) ro_prefetch_bank(d, i, e.episode + 1u, nxt);
#endif
            // current observation rows of the warp's 32 envs: coalesced into the tile, then one row per lane
            const float *gobs = d.obs + (size_t)(chunk * RO_THREADS + warp * 32) * RACE_OBS;
            __syncwarp();
            for (int m = lane; m < rows_w * RACE_OBS; m += 32) tile[m] = __ldcs(&gobs[m]);
            __syncwarp();
#pragma unroll
            for (int m = 0; m < RACE_OBS; m++) o[m] = valid ? my_row[m] : 0.0f;
            __syncwarp();
        }

#if B2D_RO_TIMING
        tprev = clock64();
#endif
        for (int k = 0; k < K; k++) {
            __syncwarp(); // lanes reconverge after the divergent episode logic before the warp-collective tcgen05 ops
            // ---- 1. observation row -> A operand of the encoder GEMM (TF32, rounded) and -> the warp's staging tile
            {
                const uint32_t R = 0x1000u; // half an ulp of the 10-bit mantissa: the tensor core drops the low 13 bits
                auto t = [&](float v) { return __uint_as_float(__float_as_uint(v) + R); };
#pragma unroll
                for (int c = 0; c < 7; c++)
                    A1[c * RO_THREADS + tid] = make_float4(t(o[4 * c]), t(o[4 * c + 1]), t(o[4 * c + 2]), t(o[4 * c + 3]));
                A1[7 * RO_THREADS + tid] = make_float4(t(o[28]), 1.0f, 1.0f, 0.0f); // x bias_hi, x bias_lo, x 0
#pragma unroll
                for (int m = 0; m < RACE_OBS; m++) my_row[m] = o[m];
            }
            RO_TICK(0) // observation row to shared memory
            ro_fence_proxy_async();
            ro_tc_fence_before();
            __syncthreads();
            RO_TICK(1) // barrier: wait for the slowest warp of the CTA
            if (tid == 0) {
                ro_tc_fence_after();
#pragma unroll
                for (int j = 0; j < RO_K1 / 8; j++) // K = 8 per instruction = two 16-byte chunks
                    ro_mma_ss(tm_h, ro_umma_desc(sA1 + j * 2 * (RO_THREADS * 16), RO_THREADS * 16, 128),
                              ro_umma_desc(sB1 + j * 2 * (RO_HIDDEN * 16), RO_HIDDEN * 16, 128), IDESC1, j > 0 ? 1u : 0u);
                ro_tc_commit(&s_mbar[0]);
            }
            // ---- meanwhile: the experience row of this step that does not depend on the policy output
            {
                const size_t row0 = (size_t)k * d.n + (size_t)chunk * RO_THREADS + warp * 32;
                if (a.st_obs && rows_w > 0) {
                    float *dst = a.st_obs + row0 * RACE_OBS;
                    if (rows_w == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                        const float4 *src4 = reinterpret_cast<const float4 *>(tile);
                        float4 *dst4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
                        for (int m = 0; m < 7; m++) __stcs(&dst4[m * 32 + lane], src4[m * 32 + lane]);
                        if (lane < 8) __stcs(&dst4[224 + lane], src4[224 + lane]);
                    } else {
                        for (int m = lane; m < rows_w * RACE_OBS; m += 32) __stcs(&dst[m], tile[m]);
                    }
                }
                if (valid) {
                    if (a.st_rew) __stcs(&a.st_rew[(size_t)k * d.n + i], fminf(fmaxf(prev_rew, -1.0f), 1.0f)); // pufferl.py:260
                    if (a.st_term) __stcs(&a.st_term[(size_t)k * d.n + i], prev_term);
                }
            }
            RO_TICK(2) // experience stores (overlap the encoder GEMM)
            // ---- 2. hidden = GELU(encoder(obs)): accumulator row -> registers -> activations back into the same columns
            ro_mbar_wait(&s_mbar[0], phase);
            ro_tc_fence_after();
            RO_TICK(3) // wait for the encoder GEMM
            {   // 16 columns at a time, the next chunk's tcgen05.ld in flight while this one is computed
                uint32_t ra[16], rb[16];
                ro_tmem_ld16(tm_h + lane_base, ra);
                ro_tc_wait_ld();
#pragma unroll 1
                for (int c = 0; c < RO_HIDDEN / 32; c++) {
                    ro_tmem_ld16(tm_h + lane_base + c * 32 + 16, rb);
                    ro_gelu16(ra);
                    ro_tmem_st16(tm_h + lane_base + c * 32, ra);
                    ro_tc_wait_ld();
                    if (c + 1 < RO_HIDDEN / 32) ro_tmem_ld16(tm_h + lane_base + c * 32 + 32, ra);
                    ro_gelu16(rb);
                    ro_tmem_st16(tm_h + lane_base + c * 32 + 16, rb);
                    ro_tc_wait_ld();
                }
            }
            RO_TICK(4) // GELU
            ro_tc_wait_st();
            ro_tc_fence_before();
            __syncthreads();
            RO_TICK(5) // barrier
            // ---- 3. heads: [means | value] = hidden . W2^T, A from TMEM
            if (tid == 0) {
                ro_tc_fence_after();
#pragma unroll
                for (int j = 0; j < RO_HIDDEN / 8; j++)
                    ro_mma_ts(tm_o, tm_h + j * 8, ro_umma_desc(sB2 + j * 2 * (RO_N2 * 16), RO_N2 * 16, 128), IDESC2, j > 0 ? 1u : 0u);
                ro_tc_commit(&s_mbar[1]);
            }
            // noise of this (row, call): independent of the GEMMs, computed while the head GEMM runs
            float z[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (!a.deterministic) policy_noise(a.row_id_base + (uint32_t)i, call0 + (unsigned int)k, a.seed_lo, a.seed_hi, z);
            ro_mbar_wait(&s_mbar[1], phase);
            ro_tc_fence_after();
            phase ^= 1u;
            uint32_t hd[8];
            ro_tmem_ld8(tm_o + lane_base, hd);
            ro_tc_wait_ld();
            ro_tc_fence_before(); // the next step's MMAs overwrite these columns after the next barrier
            RO_TICK(6) // head GEMM round trip

            // ---- 4. sample, log-prob, experience stores, clip (pufferl.py:258-294)
            float act[4], lp = lp0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float mean = __uint_as_float(hd[c]) + bm[c];
                act[c] = a.deterministic ? mean : fmaf(sd[c], z[c], mean);
                if (!a.deterministic) lp = fmaf(-0.5f * z[c], z[c], lp);
            }
            const float value = __uint_as_float(hd[4]) + bv;
            if (valid) {
                const size_t r = (size_t)k * d.n + i;
                if (a.st_act) __stcs(reinterpret_cast<float4 *>(a.st_act) + r, make_float4(act[0], act[1], act[2], act[3]));
                if (a.st_logp) __stcs(&a.st_logp[r], lp);
                if (a.st_val) __stcs(&a.st_val[r], value);
            }
            const float4 a4 = make_float4(fminf(fmaxf(act[0], -1.0f), 1.0f), fminf(fmaxf(act[1], -1.0f), 1.0f),
                                          fminf(fmaxf(act[2], -1.0f), 1.0f), fminf(fmaxf(act[3], -1.0f), 1.0f));

            // ---- 5. the env step, in registers
            float reward = 0.0f;
            int terminal = 0;
            if (valid) ro_env_step<STRICT>(d, i, e, a4, reward, terminal, o, s_acc, k == K - 1, score_last, nxt);
            prev_rew = reward;
            prev_term = (float)terminal;
            RO_TICK(7) // sampling, stores, env step
            if (k == K - 1 && valid) reinterpret_cast<float4 *>(a.env_act)[i] = a4;
        }

        // ---- the env goes back to HBM: state, current ring, parameters (episodes may have changed), contract buffers
        if (valid) {
            race_store_state(d, i, e.s, e.tick, e.ring_idx | e.ring_ext, e.ep_ret);
            const float pp[13] = {e.p.mass, e.p.ixx, e.p.iyy, e.p.izz, e.p.arm, e.p.kt, e.p.kad, e.p.kd, e.p.bd, e.p.g, e.p.mrpm, e.p.kmot, e.p.jmot};
            race_store_params(d, i, pp, e.episode);
            race_store_current_ring(d, i, e.ring);
            d.rew[i] = prev_rew;
            d.term[i] = (unsigned char)(prev_term != 0.0f);
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < RACE_OBS; m++) my_row[m] = o[m];
        __syncwarp();
        {
            float *gobs = d.obs + (size_t)(chunk * RO_THREADS + warp * 32) * RACE_OBS;
            for (int m = lane; m < rows_w * RACE_OBS; m += 32) gobs[m] = tile[m];
        }
        __syncwarp();
    }

#if B2D_RO_TIMING
    if (lane == 0) {
        for (int m = 0; m < 8; m++) atomicAdd(&d.ctl->dbg[m], (unsigned long long)tm[m]);
        atomicAdd(&d.ctl->dbg[8], 1ull);
    }
#endif
    // ---- teardown: TMEM back, statistics out, call counter advanced by K
    ro_tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ro_tc_fence_after();
        ro_tmem_free(tm_h, 128);
        ro_tmem_free(tm_o, 32);
    }
    if (score_last != 0) atomicAdd(&s_score, score_last);
    __syncthreads();
    if (tid < 7) {
        const int v = s_acc[tid];
        if (v != 0) atomicAdd((unsigned long long *)&d.ctl->acc[tid], (unsigned long long)(long long)v);
    }
    if (tid == 0) {
        if (s_acc[ACC_SPARE] != 0) atomicAdd(&d.ctl->guard_replays, (unsigned long long)s_acc[ACC_SPARE]);
        // log.score covers the episodes that ended in the LAST step only (DR/drone_race.h:160)
        if ((int)blockIdx.x < d.max_grid) d.cta_score[blockIdx.x] = (long long)s_score;
        if (blockIdx.x == 0)
            for (int m = (int)gridDim.x; m < d.max_grid; m++) d.cta_score[m] = 0;
        __threadfence();
        const unsigned int arrived = atomicAdd(&a.counter[1], 1u);
        if (arrived == gridDim.x - 1) {
            a.counter[1] = 0;
            atomicAdd(&d.ctl->ctas_done, (unsigned int)K); // K vec_steps completed
            __threadfence();
            atomicAdd(&a.counter[0], (unsigned int)K);
        }
    }
}

} // namespace b2d
