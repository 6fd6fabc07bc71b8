/* b200drone.h -- C ABI of the B200-native drone env step (libb200drone.so).
 *
 * Drop-in boundary for the batched environment step of PufferLib's Ocean drone
 * envs.  Each entry point names the reference interface it replaces
 * (EB = pufferlib/ocean/env_binding.h, DR = pufferlib/ocean/drone_race,
 * DS = pufferlib/ocean/drone_swarm).  Plain pointers and sizes only: no torch,
 * no Python, no C++ types.  All functions return B2D_OK (0) or a negative
 * b2d_status and never throw; b2d_last_error() describes the last failure of
 * the calling thread.  One host thread per handle; everything that takes a
 * stream is stream-ordered, asynchronous and CUDA-graph capturable unless
 * stated otherwise (no allocation, no sync, no host-side state mutation that
 * a graph replay would miss).
 *
 * Env state lives on the device in SoA float4 arrays owned by the handle; the
 * contract buffers (observations/actions/rewards/terminals/truncations) keep
 * the reference's flat row-major layout (PL/pufferlib.py:22-43) and may be
 * caller-owned device memory (zero-copy: torch tensors, DLPack) or host
 * memory mirrored by b2d_vec_step_host().
 */
#ifndef B200DRONE_H
#define B200DRONE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2D_VERSION 1
#define B2D_RACE_OBS 29      /* DR/drone_race.py:17-22 */
#define B2D_SWARM_OBS 41     /* DS/drone_swarm.py:19-24 */
#define B2D_ACT 4
#define B2D_LOG_FIELDS 9     /* DR/dronelib.h:52-63 */
#define B2D_RACE_BLOB 33     /* + 6*max_rings floats, see b2d_get_state */
#define B2D_SWARM_AGENT_BLOB 47
#define B2D_SWARM_AGENT_PAYLOAD 41 /* per-agent draw results of the swarm reset payload */

typedef struct b2d_vec b2d_vec; /* replaces VecEnv, EB:262-265 */

typedef enum b2d_status {
    B2D_OK = 0,
    B2D_EINVAL = -1, /* bad argument / layout   (reference: TypeError / ValueError) */
    B2D_ENOMEM = -2, /* allocation failed       (reference: MemoryError, EB:57-60) */
    B2D_ECUDA = -3,  /* CUDA runtime error */
    B2D_ESTATE = -4  /* call not valid in the handle's current state */
} b2d_status;

typedef enum b2d_math {
    B2D_MATH_FAST = 0,  /* FMA contraction, hoisted reciprocals; within 1e-5 rel of the reference per step */
    B2D_MATH_STRICT = 1 /* one IEEE-754 binary32 op per reference op, no FMA: bit-exact with the reference C */
} b2d_math;

typedef enum b2d_reset_mode {
    B2D_RESET_PHILOX = 0, /* counter-based Philox4x32-10 stream keyed by (seed, env id, step, item, attempt) */
    B2D_RESET_INJECT = 1  /* parity hook: an env that terminates takes its next episode from the payload buffer */
} b2d_reset_mode;

typedef enum b2d_mem { B2D_MEM_DEVICE = 0, B2D_MEM_HOST = 1 } b2d_mem;

/* The flat buffer contract of PufferEnv.set_buffers (PL/pufferlib.py:22-43).
 * Row counts are num_agents = num_envs (race) or num_envs*num_agents (swarm).
 * A NULL member is allocated by the library (device memory). */
typedef struct b2d_buffers {
    float *observations;        /* [num_agents, obs_dim] f32 row-major, 16-byte aligned */
    float *actions;             /* [num_agents, 4] f32, 16-byte aligned */
    float *rewards;             /* [num_agents] f32 */
    unsigned char *terminals;   /* [num_agents] u8 */
    unsigned char *truncations; /* [num_agents] u8, never written (EB:136, EB:421) */
    int location;               /* b2d_mem: where the non-NULL pointers live */
} b2d_buffers;

/* kwargs of DR/binding.c:6-11 (my_init) + placement */
typedef struct b2d_race_cfg {
    int num_envs;
    int max_rings;           /* DR/drone_race.py:15, default 10 */
    int max_moves;           /* DR/drone_race.py:16, default 1000 */
    int device;              /* CUDA device ordinal */
    uint64_t seed;           /* Philox key until the first b2d_vec_reset */
    uint32_t env_id_base;    /* global id of env 0 (multi-GPU shards keep results invariant to the split) */
    int math;                /* b2d_math */
    int write_clamped_actions; /* DR/dronelib.h:437 clamps the shared action buffer in place.  1: the kernel stores
                                  clamp(action,-1,1) back into the device action buffer (and the host-buffer step copies
                                  it down); -1: host-buffer steps clamp the caller's host array on the host (what the
                                  reference's caller observes, no PCIe traffic), device buffers untouched; 0: never */
} b2d_race_cfg;

/* kwargs of DS/binding.c:6-11 + placement */
typedef struct b2d_swarm_cfg {
    int num_envs;
    int num_agents;          /* drones per env, DS/drone_swarm.py:11 */
    int max_rings;
    int device;
    uint64_t seed;
    uint32_t env_id_base;
    int math;
    int write_clamped_actions;
} b2d_swarm_cfg;

/* ---- construction / destruction ------------------------------------------
 * b2d_race_create replaces env_init x num_envs + vectorize (EB:50-174,450-480;
 * the path DR/drone_race.py:37-51 takes) and vec_init (EB:288-446).  `ext` may
 * be NULL (library allocates device buffers) or name caller-owned buffers; the
 * library never frees caller memory (ownership as in the reference: buffers
 * belong to the caller, env structs to the binding). Synchronous. */
int b2d_race_create(b2d_vec **out, const b2d_race_cfg *cfg, const b2d_buffers *ext);
int b2d_swarm_create(b2d_vec **out, const b2d_swarm_cfg *cfg, const b2d_buffers *ext);
/* replaces vec_close (EB:600-613). Synchronous. */
int b2d_vec_close(b2d_vec *vec);

/* ---- the hot path ----------------------------------------------------------
 * replaces vec_reset (EB:482-506): every env starts a fresh episode and its
 * observation row is written.  seed re-keys the reset stream. */
int b2d_vec_reset(b2d_vec *vec, uint64_t seed, void *cuda_stream);
/* replaces vec_step (EB:508-524): reads actions, writes observations, rewards,
 * terminals in place (device buffers), auto-resets finished envs. */
int b2d_vec_step(b2d_vec *vec, void *cuda_stream);
/* same, reading this step's actions from another device buffer of the same
 * shape (a policy's output tensor, or one slice of an action tape) */
int b2d_vec_step_from(b2d_vec *vec, const float *device_actions, void *cuda_stream);
/* `steps` consecutive vec_steps whose actions come from a device-resident tape
 * [tape_len][num_agents][4] (step k reads slice (first + k) % tape_len) -- the reference's own
 * cached-action loop (DR/drone_race.py:83-90) without a host round trip per step.  One kernel
 * launch per step, each reading and writing the contract buffers like b2d_vec_step; successive
 * launches may overlap at their edges (env shards are independent between steps). */
int b2d_vec_step_tape(b2d_vec *vec, const float *device_tape, int tape_len, int first, int steps, void *cuda_stream);
/* host-buffer form of vec_step for callers that keep the reference's NumPy
 * buffers: copies actions H2D, steps, copies observations/rewards/terminals
 * D2H (chunked and overlapped), then synchronises.  Host pointers default to
 * the b2d_buffers given at create time (location == B2D_MEM_HOST); pinned
 * memory is fastest.  Not capturable. */
int b2d_vec_step_host(b2d_vec *vec, void *cuda_stream);
/* same, taking this step's actions from another host array of the same shape: they end up in the
 * caller-visible action buffer (DR/drone_race.py:59 `self.actions[:] = actions`), clamped like the
 * reference leaves them (DR/dronelib.h:437).  A PAGE-LOCKED `host_actions` (cudaHostAlloc /
 * cudaHostRegister by the caller) is uploaded by DMA from where it is and that copy is made while the
 * results come down; pageable memory is copied first, chunk by chunk.  `host_actions` is only read. */
int b2d_vec_step_host_from(b2d_vec *vec, const float *host_actions, void *cuda_stream);
/* host-buffer form of vec_reset: reset + observations D2H + sync */
int b2d_vec_reset_host(b2d_vec *vec, uint64_t seed, void *cuda_stream);

/* ---- episode statistics ----------------------------------------------------
 * replaces vec_log (EB:564-598).  out[0..8] follow the Log field order
 * (episode_return, episode_length, rings_passed, collision_rate, oob, timeout,
 * score, perf, n); fields are divided by n like EB:588-591 and out[8] is the
 * raw episode count; all zeros when no episode finished (reference returns
 * {}).  Accumulators are cleared.  Synchronises the stream. */
int b2d_vec_log(b2d_vec *vec, float out[B2D_LOG_FIELDS], void *cuda_stream);
/* split form for multi-GPU: begin snapshots+clears the accumulators into a
 * device array of `*count` int64 sums (stream-ordered, no sync) that the
 * caller may all-reduce (NCCL sum) in place; end synchronises and averages. */
int b2d_vec_log_begin(b2d_vec *vec, void *cuda_stream, long long **device_sums, int *count);
int b2d_vec_log_end(b2d_vec *vec, float out[B2D_LOG_FIELDS], void *cuda_stream);
/* vec_log across the ranks of a multi-GPU job in one call, for hosts without torch.distributed: snapshot, NCCL
 * all-reduce (sum, int64, in place, on `cuda_stream`) over `nccl_comm` -- an ncclComm_t of the caller's NCCL,
 * passed as void* so that this header needs no nccl.h -- then the averaging of EB:588-591 on the global sums:
 * every rank receives the same averages and the global episode count in out[8] (the semantics of EB:564-598
 * over all shards).  The library does not link NCCL: ncclAllReduce is resolved at run time from the NCCL the
 * process has loaded (dlsym, falling back to dlopen of libnccl.so.2); B2D_ESTATE when none is found.
 * nccl_comm == NULL reduces nothing (single rank).  Synchronises the stream. */
int b2d_vec_log_reduce(b2d_vec *vec, float out[B2D_LOG_FIELDS], void *nccl_comm, void *cuda_stream);
/* the averaging step of vec_log_end on host sums (kind 0 = race, 1 = swarm): pure host
 * arithmetic, EB:588-591 + my_log; used after a cross-rank reduction of the sums */
int b2d_log_average(int kind, int max_rings, const long long *sums, int count, float out[B2D_LOG_FIELDS]);

/* ---- introspection ------------------------------------------------------------ */
int b2d_get_buffers(const b2d_vec *vec, b2d_buffers *device_buffers); /* raw device pointers (DLPack / torch views) */
int b2d_num_agents(const b2d_vec *vec); /* rows of the contract buffers */
int b2d_obs_dim(const b2d_vec *vec);
int b2d_state_blob_floats(const b2d_vec *vec); /* floats per env in get/put_state blobs */
long long b2d_kernel_launches(const b2d_vec *vec); /* kernels launched by this handle so far */
int b2d_step_count(b2d_vec *vec, uint32_t *steps, void *cuda_stream); /* steps since the last reset (sync) */
/* B2D_MATH_FAST only: agent-steps (race) / env-steps (swarm) the step kernel re-did in the reference's
 * arithmetic since creation because a decision (out of bounds DR/drone_race.h:165-172, ring test
 * DR/dronelib.h:462-489, swarm collision / nearest neighbour DS/drone_swarm.h:107-129,347-352) fell within
 * rounding distance of its threshold; this guard is what makes the integer outputs of the fast kernels
 * identical to the strict ones.  Synchronises the stream. */
int b2d_guard_replays(b2d_vec *vec, unsigned long long *count, void *cuda_stream);

/* ---- state hooks (env_get / env_put, EB:228-260; also the env checkpoint) ---
 * Race blob per env, float32[33 + 6*max_rings], ints stored as exact floats:
 *   [0:3] pos [3:6] vel [6:10] quat(w,x,y,z) [10:13] omega [13:17] rpms
 *   [17:30] mass,ixx,iyy,izz,arm_len,k_thrust,k_ang_damp,k_drag,b_drag,gravity,max_rpm,k_mot,j_mot
 *   [30] tick [31] ring_idx [32] episodic_return, then per ring pos(3), normal(3).
 * env_ids == NULL means envs 0..n-1.  Host blobs; synchronous. */
int b2d_get_state(b2d_vec *vec, const int *env_ids, int n, float *host_blobs);
int b2d_put_state(b2d_vec *vec, const int *env_ids, int n, const float *host_blobs);
/* recompute observation rows from the current state (after put_state) */
int b2d_observe(b2d_vec *vec, void *cuda_stream);

/* ---- render / checkpoint bridge ----------------------------------------
 * Host structs with the memory layout of the reference's `Drone` (DR/dronelib.h:191-247: State, Params,
 * spawn/prev/target vectors, reward bookkeeping; 208 bytes) and `Ring` (DR/dronelib.h:161-166; 44 bytes), so
 * that a device env can be handed to code written against the reference's structs -- its viewer
 * (c_render, DR/drone_race.h:331-462; the raylib client is outside this library), compute_observations, a
 * debugger.  Fields the device does not keep are filled consistently: prev_pos = pos, max_vel = max_omega =
 * 50 (DR/dronelib.h:286-287), ring radius = 2 (DR/drone_race.h:135), ring orientation = the shortest rotation
 * that turns +z into the ring normal (the reference's is a uniformly random quaternion with the same normal;
 * it only matters to the viewer).  race spawn_pos / target_pos / target_vel are unused by the reference: zero. */
typedef struct b2d_ref_drone {
    float pos[3], vel[3], quat[4] /* w,x,y,z */, omega[3], rpms[4];                             /* State  */
    float mass, ixx, iyy, izz, arm_len, k_thrust, k_ang_damp, k_drag, b_drag, gravity, max_rpm,
          max_vel, max_omega, k_mot, j_mot;                                                    /* Params */
    float spawn_pos[3], prev_pos[3], target_pos[3], target_vel[3];
    float last_abs_reward, last_target_reward, last_collision_reward, episode_return, collisions;
    int episode_length;
    float score;
    int ring_idx;
} b2d_ref_drone;
typedef struct b2d_ref_ring {
    float pos[3], orientation[4] /* w,x,y,z */, normal[3], radius;
} b2d_ref_ring;
/* state blob (b2d_get_state layout) -> reference structs; pure host arithmetic, no device needed.
 * race: one drone, `max_rings` rings, *tick / *ring_idx / *episodic_return = the DroneRace scalars
 * (score == ring_idx, moves_left == max_moves - tick, DR/drone_race.h:31-53).
 * swarm: `num_agents` drones, `max_rings` rings (radius 0 for a zeroed ring of a non-race task), *tick, *task. */
int b2d_race_blob_to_ref(const float *blob, int max_rings, b2d_ref_drone *drone, b2d_ref_ring *rings, int *tick,
                         int *ring_idx, float *episodic_return);
int b2d_swarm_blob_to_ref(const float *blob, int num_agents, int max_rings, b2d_ref_drone *drones, b2d_ref_ring *rings,
                          int *tick, int *task);
/* b2d_get_state of one env + the conversion above (synchronous).  aux = ring_idx (race) / task (swarm);
 * episodic_return may be NULL. */
int b2d_export_ref(b2d_vec *vec, int env_id, b2d_ref_drone *drones, b2d_ref_ring *rings, int *tick, int *aux,
                   float *episodic_return);

/* ---- parity / configuration hooks --------------------------------------------- */
int b2d_set_math(b2d_vec *vec, int math);
int b2d_set_reset_mode(b2d_vec *vec, int reset_mode);
/* host_payload for B2D_RESET_INJECT (copied; synchronous).
 *   race:  [num_envs][state blob]: the post-reset state an env takes when it terminates this step.
 *   swarm: [num_envs][num_agents*41 + 2 + 6*max_rings]: the RESULTS of the random draws of this step:
 *          per agent [0:13] params [13:16] pos of an out-of-bounds respawn (DS/drone_swarm.h:378-399),
 *          [16:29] params [29:32] pos [32:35] target_pos [35:38] target_vel [38:41] race start pos of an
 *          env-wide reset (DS/drone_swarm.h:401-443); then per env [0] unused [1] task, rings pos(3) normal(3).
 * Swarm state blob (b2d_get_state / b2d_put_state) per env: [num_agents][47] (pos vel quat omega rpms |
 * 13 params | spawn_pos target_pos target_vel | last_abs last_target last_collision reward |
 * episode_return collisions episode_length score ring_idx) then [2 + 6*max_rings] (tick, task, rings). */
int b2d_set_reset_payload(b2d_vec *vec, const float *host_payload);
int b2d_set_step_count(b2d_vec *vec, uint32_t steps);

/* Per-kernel timing with CUDA events: enable=1 starts timing the kernels of every later step,
 * enable=0 synchronises and returns mean microseconds {step kernel, 0 (reserved), steps timed}. */
int b2d_profile_kernels(b2d_vec *vec, int enable, float out_us[3]);

/* ---- trainer-side helper -------------------------------------------
 * replaces torch.ops.pufferlib.compute_puff_advantage (extensions/cuda/pufferlib.cu:41-82 and its
 * CPU twin extensions/pufferlib.cpp:28-41,63-72): GAE with V-trace clipping, backwards in time per
 * row.  Device pointers to float32 tensors addressed as x[row * row_stride + t * t_stride]:
 * row-major [segments, horizon] (row_stride = horizon, t_stride = 1, the reference's layout) or
 * time-major [horizon, num_agents] (row_stride = 1, t_stride = num_agents, coalesced).
 * advantages[., horizon-1] is not written (as in the reference).  abs_sum (optional, [num_rows]) gets
 * sum_t |advantage| per row, the priority PuffeRL.train computes next (pufferl.py:342). */
int b2d_puff_advantage(const float *values, const float *rewards, const float *dones, const float *importance,
                       float *advantages, float *abs_sum, int num_rows, int horizon, long long row_stride,
                       long long t_stride, float gamma, float lambda, float rho_clip, float c_clip, int math,
                       void *cuda_stream);

/* ---- rollout-side helper -------------------------------------------
 * One kernel for everything PuffeRL.evaluate does per step between vecenv.recv() and
 * vecenv.send() for a Box action space (pufferl.py:229-296): forward_eval of
 * pufferlib.models.Default (models.py:41-98: Linear(obs_dim, hidden) + GELU, mean head,
 * state-independent log-std, value head), sample_logits on Normal(mean, exp(logstd))
 * (pytorch.py:189-199), reward clamp to [-1, 1] (pufferl.py:260), the experience stores
 * (pufferl.py:270-281) and the clip of the action to the action space (pufferl.py:292-294).
 * All pointers are device memory, float32 unless stated, torch nn.Linear layouts. */
typedef enum b2d_policy_precision {
    B2D_POLICY_FP32 = 0, /* both Linear layers in float32 FMAs on the CUDA cores (|error| ~1e-6 vs float64) */
    B2D_POLICY_TF32 = 1  /* both Linear layers as TF32 tensor-core GEMMs with float32 accumulation: what the reference
                            itself runs on a GPU (torch.set_float32_matmul_precision('high'), pufferl.py:55) */
} b2d_policy_precision;

typedef struct b2d_policy_weights {
    const float *encoder_weight;      /* [hidden, obs_dim] */
    const float *encoder_bias;        /* [hidden] */
    const float *decoder_mean_weight; /* [4, hidden] */
    const float *decoder_mean_bias;   /* [4] */
    const float *decoder_logstd;      /* [4] */
    const float *value_weight;        /* [1, hidden] */
    const float *value_bias;          /* [1] */
    int hidden;                       /* multiple of 8, <= 256 */
    int precision;                    /* b2d_policy_precision */
} b2d_policy_weights;

typedef struct b2d_policy_io {
    const float *observations;      /* [rows, obs_dim] the env's observation buffer */
    const float *rewards;           /* [rows] */
    const unsigned char *terminals; /* [rows] u8 */
    float *env_actions;             /* [rows, 4] out: clip(action, -1, 1), the env's action buffer */
    /* experience row of this step, each optional (NULL = not stored) */
    float *store_observations;      /* [rows, obs_dim] */
    float *store_actions;           /* [rows, 4] the unclipped sample */
    float *store_logprobs;          /* [rows] */
    float *store_rewards;           /* [rows] clamped to [-1, 1] */
    float *store_terminals;         /* [rows] as float */
    float *store_values;            /* [rows] */
    int rows;
    int obs_dim;                    /* B2D_RACE_OBS or B2D_SWARM_OBS */
    uint32_t row_id_base;           /* global id of row 0 (multi-GPU shards draw the noise of their global rows) */
} b2d_policy_io;

/* noise: Philox4x32-10 keyed by noise_seed, counter (global row, call number), Box-Muller.  The call
 * number lives in `device_counter` (two zero-initialised uint32 owned by the caller; word 0 = calls
 * completed) and is advanced by the kernel, so a CUDA-graph replay draws fresh noise.
 * deterministic != 0: action = mean.  Stream-ordered, capturable. */
int b2d_policy_act(const b2d_policy_weights *weights, const b2d_policy_io *io, uint64_t noise_seed,
                   unsigned int *device_counter, int deterministic, void *cuda_stream);

/* ---- the whole rollout as one kernel (BASELINE.json configs[3]) -------------------
 * replaces PuffeRL.evaluate's per-step loop (pufferl.py:214-314) for a race vec: `horizon` consecutive
 * (policy step, vec_step) pairs.  Each CTA keeps 128 envs in registers for all steps; both Linear layers of
 * the Default policy (hidden must be 128, precision is TF32 as under torch.set_float32_matmul_precision('high'),
 * pufferl.py:55) run as tcgen05 tensor-core GEMMs with TMEM accumulators; per env-step only the experience row
 * is written.  Step k stores, time-major: observations[k] = the observation the policy saw, rewards[k] /
 * terminals[k] = what the env returned for step k-1 (the contract buffers at entry for k = 0; reward clamped to
 * [-1, 1]), actions[k] (unclipped sample), logprobs[k], values[k]; the env receives clip(action, -1, 1).
 * On return (stream-ordered) the contract buffers hold the results of the last step, exactly as after
 * `horizon` calls of b2d_policy_act + b2d_vec_step, and the step counter has advanced by `horizon`.
 * Noise as in b2d_policy_act: call number = device_counter[0] + k, advanced by `horizon` by the kernel.
 * Race handles in B2D_RESET_PHILOX mode only.  Capturable. */
typedef struct b2d_rollout_store { /* device pointers, float32, [horizon][num_agents][...]; each may be NULL */
    float *observations; /* [horizon, num_agents, 29] */
    float *actions;      /* [horizon, num_agents, 4] */
    float *logprobs;     /* [horizon, num_agents] */
    float *rewards;      /* [horizon, num_agents] */
    float *terminals;    /* [horizon, num_agents] */
    float *values;       /* [horizon, num_agents] */
} b2d_rollout_store;
int b2d_race_rollout(b2d_vec *vec, const b2d_policy_weights *weights, const b2d_rollout_store *store, int horizon,
                     uint64_t noise_seed, unsigned int *device_counter, int deterministic, void *cuda_stream);

const char *b2d_last_error(void);
int b2d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200DRONE_H */
