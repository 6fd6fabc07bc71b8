/* oracle/_ref shim for drone_race -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Compiles the UNMODIFIED reference headers where they lie under
 * /root/reference (pufferlib/ocean/drone_race/{drone_race.h,dronelib.h}) and
 * exposes them through a plain C ABI so tests / bench.py's cpu_baseline can
 * drive the reference's own c_reset / c_step and peek at its internal state.
 * Nothing of the reference is copied here: this file only *calls* it.
 *
 * The vector loop mirrors pufferlib/ocean/env_binding.h:
 *   vec_reset  EB:500-504  (srand(i + seed*num_envs); c_reset(env))
 *   vec_step   EB:520-522  (for i: c_step(envs[i]))
 *   vec_log    EB:572-580  (float-wise sum over envs, zeroing each env's Log)
 *
 * State blob (float32[REF_RACE_BLOB + 6*max_rings]); ints stored as exact floats:
 *   [0:3] pos  [3:6] vel  [6:10] quat(w,x,y,z)  [10:13] omega  [13:17] rpms
 *   [17:30] mass,ixx,iyy,izz,arm_len,k_thrust,k_ang_damp,k_drag,b_drag,gravity,
 *           max_rpm,k_mot,j_mot
 *   [30] tick  [31] ring_idx  [32] episodic_return
 *   [33 + 6*r : 33 + 6*r + 6] ring r pos(3), normal(3)
 */
#include "drone_race.h"
#include <stdint.h>

#define REF_RACE_BLOB 33

typedef struct {
    DroneRace *envs;
    int n;
} RefRaceVec;

void *refrace_create(int n, int max_rings, int max_moves, float *obs, float *act, float *rew,
                     unsigned char *term) {
    RefRaceVec *v = (RefRaceVec *)calloc(1, sizeof(RefRaceVec));
    v->envs = (DroneRace *)calloc((size_t)n, sizeof(DroneRace));
    v->n = n;
    for (int i = 0; i < n; i++) {
        DroneRace *e = &v->envs[i];
        e->observations = obs + (size_t)i * 29;
        e->actions = act + (size_t)i * 4;
        e->rewards = rew + i;
        e->terminals = term + i;
        e->max_rings = max_rings;
        e->max_moves = max_moves;
        init(e);
    }
    return v;
}

void refrace_reset(void *vp, int seed) {
    RefRaceVec *v = (RefRaceVec *)vp;
    for (int i = 0; i < v->n; i++) {
        srand(i + seed * v->n);
        c_reset(&v->envs[i]);
    }
}

void refrace_reset_one(void *vp, int i) { c_reset(&((RefRaceVec *)vp)->envs[i]); }

void refrace_step(void *vp) {
    RefRaceVec *v = (RefRaceVec *)vp;
    for (int i = 0; i < v->n; i++) c_step(&v->envs[i]);
}

/* steps [lo,hi) only: lets a multi-process baseline own disjoint env ranges */
void refrace_step_range(void *vp, int lo, int hi) {
    RefRaceVec *v = (RefRaceVec *)vp;
    for (int i = lo; i < hi; i++) c_step(&v->envs[i]);
}

void refrace_log(void *vp, float out[9]) {
    RefRaceVec *v = (RefRaceVec *)vp;
    for (int j = 0; j < 9; j++) out[j] = 0.0f;
    for (int i = 0; i < v->n; i++) {
        float *l = (float *)&v->envs[i].log;
        for (int j = 0; j < 9; j++) {
            out[j] += l[j];
            l[j] = 0.0f;
        }
    }
}

void refrace_get_state(void *vp, int i, float *b) {
    DroneRace *e = &((RefRaceVec *)vp)->envs[i];
    State *s = &e->drone.state;
    Params *p = &e->drone.params;
    b[0] = s->pos.x; b[1] = s->pos.y; b[2] = s->pos.z;
    b[3] = s->vel.x; b[4] = s->vel.y; b[5] = s->vel.z;
    b[6] = s->quat.w; b[7] = s->quat.x; b[8] = s->quat.y; b[9] = s->quat.z;
    b[10] = s->omega.x; b[11] = s->omega.y; b[12] = s->omega.z;
    for (int k = 0; k < 4; k++) b[13 + k] = s->rpms[k];
    b[17] = p->mass; b[18] = p->ixx; b[19] = p->iyy; b[20] = p->izz;
    b[21] = p->arm_len; b[22] = p->k_thrust; b[23] = p->k_ang_damp; b[24] = p->k_drag;
    b[25] = p->b_drag; b[26] = p->gravity; b[27] = p->max_rpm; b[28] = p->k_mot; b[29] = p->j_mot;
    b[30] = (float)e->tick; b[31] = (float)e->ring_idx; b[32] = e->episodic_return;
    for (int r = 0; r < e->max_rings; r++) {
        Ring *g = &e->ring_buffer[r];
        float *o = b + REF_RACE_BLOB + 6 * r;
        o[0] = g->pos.x; o[1] = g->pos.y; o[2] = g->pos.z;
        o[3] = g->normal.x; o[4] = g->normal.y; o[5] = g->normal.z;
    }
}

void refrace_put_state(void *vp, int i, const float *b) {
    DroneRace *e = &((RefRaceVec *)vp)->envs[i];
    State *s = &e->drone.state;
    Params *p = &e->drone.params;
    s->pos = (Vec3){b[0], b[1], b[2]};
    s->vel = (Vec3){b[3], b[4], b[5]};
    s->quat = (Quat){b[6], b[7], b[8], b[9]};
    s->omega = (Vec3){b[10], b[11], b[12]};
    for (int k = 0; k < 4; k++) s->rpms[k] = b[13 + k];
    p->mass = b[17]; p->ixx = b[18]; p->iyy = b[19]; p->izz = b[20];
    p->arm_len = b[21]; p->k_thrust = b[22]; p->k_ang_damp = b[23]; p->k_drag = b[24];
    p->b_drag = b[25]; p->gravity = b[26]; p->max_rpm = b[27]; p->k_mot = b[28]; p->j_mot = b[29];
    p->max_vel = BASE_MAX_VEL; p->max_omega = BASE_MAX_OMEGA;
    e->tick = (int)b[30]; e->ring_idx = (int)b[31]; e->episodic_return = b[32];
    e->score = e->ring_idx;
    e->moves_left = e->max_moves - e->tick;
    e->drone.prev_pos = s->pos;
    for (int r = 0; r < e->max_rings; r++) {
        Ring *g = &e->ring_buffer[r];
        const float *o = b + REF_RACE_BLOB + 6 * r;
        g->pos = (Vec3){o[0], o[1], o[2]};
        g->normal = (Vec3){o[3], o[4], o[5]};
        g->radius = 2.0f;
    }
}

/* recompute the observation row from the current internal state (after put_state) */
void refrace_observe(void *vp, int i) { compute_observations(&((RefRaceVec *)vp)->envs[i]); }

void refrace_close(void *vp) {
    RefRaceVec *v = (RefRaceVec *)vp;
    for (int i = 0; i < v->n; i++) free(v->envs[i].ring_buffer);
    free(v->envs);
    free(v);
}

int refrace_sizeof_env(void) { return (int)sizeof(DroneRace); }
int refrace_sizeof_drone(void) { return (int)sizeof(Drone); }
int refrace_sizeof_ring(void) { return (int)sizeof(Ring); }

/* the reference's compute_observations on caller-provided reference structs (the render / checkpoint bridge of
 * include/b200drone.h hands out exactly these): a DroneRace is assembled around them, nothing else is touched */
void refrace_observe_structs(const Drone *drone, Ring *rings, int max_rings, int ring_idx, float *obs29) {
    DroneRace env;
    memset(&env, 0, sizeof(env));
    env.observations = obs29;
    env.max_rings = max_rings;
    env.ring_idx = ring_idx;
    env.ring_buffer = rings;
    env.drone = *drone;
    compute_observations(&env);
}
