/* drone_oracle.h -- CPU restatement of the drone env step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product path (drone_b200/) never
 * does.  Parity pin: tests/test_oracle_vs_ref.py checks this restatement
 * bit-for-bit against the unmodified reference compiled into oracle/_ref, and
 * tests/golden/ holds vectors generated from that reference build.
 *
 * Reference files restated (R = /root/reference/pufferlib/pufferlib/ocean):
 *   R/drone_race/dronelib.h:73-139,250-489   math helpers, init_drone, RK4, rings
 *   R/drone_race/drone_race.h:55-208         add_log, observations, c_reset, c_step
 *   R/drone_swarm/drone_swarm.h:84-497       swarm env (see orc_swarm_* below)
 *   R/env_binding.h:482-598                  vec_reset / vec_step / vec_log loops
 */
#ifndef DRONE_ORACLE_H
#define DRONE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_RACE_OBS 29
#define ORC_RACE_BLOB 33 /* + 6*max_rings; same layout as oracle/ref_shim_race.c */
#define ORC_SWARM_OBS 41
#define ORC_SWARM_AGENT 47 /* same layout as oracle/ref_shim_swarm.c */

/* how an env that terminates gets its next episode */
enum {
    ORC_RESET_LIBC = 0,   /* reference draw order from libc rand() (+1 discarded rand() per drone-step) */
    ORC_RESET_PHILOX = 1, /* the device's counter-based stream (DESIGN.md "reset stream") */
    ORC_RESET_INJECT = 2, /* post-reset state supplied by the caller (payload blob per env) */
};

/* per-env event bits reported by a step */
enum {
    ORC_EV_OOB = 1,
    ORC_EV_COLLISION = 2,
    ORC_EV_TIMEOUT = 4,
    ORC_EV_COMPLETE = 8,
    ORC_EV_RING_PASS = 16,
};

typedef struct OrcRace OrcRace;

OrcRace *orc_race_create(int n, int max_rings, int max_moves);
void orc_race_close(OrcRace *o);
/* env_id_base offsets the Philox env counter (multi-GPU shards) */
void orc_race_set_philox(OrcRace *o, uint64_t seed, uint32_t env_id_base);
void orc_race_reset(OrcRace *o, int mode, int seed, const float *payload, float *obs);
/* actions are clamped in place like the reference; events may be NULL */
void orc_race_step(OrcRace *o, int mode, float *actions, const float *payload, float *obs,
                   float *rew, unsigned char *term, unsigned char *events);
void orc_race_step_range(OrcRace *o, int lo, int hi, int mode, float *actions, const float *payload,
                         float *obs, float *rew, unsigned char *term, unsigned char *events);
void orc_race_log(OrcRace *o, float out[9]); /* sums (not averaged) + zeroing, EB:572-580 */
void orc_race_get_state(const OrcRace *o, int i, float *blob);
void orc_race_put_state(OrcRace *o, int i, const float *blob);
void orc_race_observe(const OrcRace *o, int i, float *obs_row);
uint32_t orc_race_epoch(const OrcRace *o);
void orc_race_set_epoch(OrcRace *o, uint32_t epoch);

/* Philox4x32-10 (Salmon et al., SC'11), exposed for known-answer tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* deterministic sin/cos of a float angle in [0, 2*pi] shared bit-for-bit with the device reset */
void orc_sincos_det(float theta, float *s, float *c);

typedef struct OrcSwarm OrcSwarm;
#define ORC_SWARM_AGENT_PAYLOAD 41 /* see drone_oracle.c "swarm env" for the row layouts */
OrcSwarm *orc_swarm_create(int n, int num_agents, int max_rings);
void orc_swarm_close(OrcSwarm *o);
void orc_swarm_set_philox(OrcSwarm *o, uint64_t seed, uint32_t env_id_base);
/* rows/flags the draws of each reset/step are recorded into (LIBC/PHILOX) or read from (INJECT) */
void orc_swarm_set_payload(OrcSwarm *o, float *agent_rows, unsigned char *agent_flags, float *env_rows,
                           unsigned char *env_flags);
void orc_swarm_reset(OrcSwarm *o, int mode, int seed, float *obs);
void orc_swarm_step(OrcSwarm *o, int mode, float *actions, float *obs, float *rew, unsigned char *term);
void orc_swarm_observe(const OrcSwarm *o, int e, float *obs);
void orc_swarm_log(OrcSwarm *o, float out[9]);
void orc_swarm_get_env(const OrcSwarm *o, int i, float *blob);
void orc_swarm_put_env(OrcSwarm *o, int i, const float *blob);
void orc_swarm_get_agent(const OrcSwarm *o, int i, int a, float *blob);
void orc_swarm_put_agent(OrcSwarm *o, int i, int a, const float *blob);
/* closed-form formation targets (task 2 orbit, 4 cube, 6 flag) exactly as the reference computes them */
void orc_swarm_formation_target(int task, int idx, int num_agents, float out[3]);

/* pufferlib/extensions/pufferlib.cpp:28-41,63-72 with explicit strides (x[row*row_stride + t*t_stride]) */
void orc_puff_advantage(const float *values, const float *rewards, const float *dones, const float *importance,
                        float *advantages, float *abs_sum, int num_rows, int horizon, long long row_stride,
                        long long t_stride, float gamma, float lambda, float rho_clip, float c_clip);

#ifdef __cplusplus
}
#endif
#endif
