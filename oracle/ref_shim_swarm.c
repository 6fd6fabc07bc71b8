/* oracle/_ref shim for drone_swarm -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Compiles the UNMODIFIED reference headers where they lie under
 * /root/reference (pufferlib/ocean/drone_swarm/{drone_swarm.h,dronelib.h}) and
 * exposes c_reset / c_step plus state get/put through a plain C ABI.
 * The vector loop mirrors pufferlib/ocean/env_binding.h:500-504,520-522,572-580.
 *
 * Env blob  (float32[2 + 6*max_rings]): [0] tick [1] task, then per ring pos(3), normal(3).
 * Agent blob (float32[REF_SWARM_AGENT]) per agent; ints stored as exact floats:
 *   [0:3] pos [3:6] vel [6:10] quat(w,x,y,z) [10:13] omega [13:17] rpms
 *   [17:30] mass,ixx,iyy,izz,arm_len,k_thrust,k_ang_damp,k_drag,b_drag,gravity,max_rpm,k_mot,j_mot
 *   [30:33] spawn_pos [33:36] target_pos [36:39] target_vel
 *   [39] last_abs_reward [40] last_target_reward [41] last_collision_reward
 *   [42] episode_return [43] collisions [44] episode_length [45] score [46] ring_idx
 */
#include "drone_swarm.h"
#include <stdint.h>

#define REF_SWARM_AGENT 47

typedef struct {
    DroneSwarm *envs;
    int n;
} RefSwarmVec;

void *refswarm_create(int n, int num_agents, int max_rings, float *obs, float *act, float *rew,
                      unsigned char *term) {
    RefSwarmVec *v = (RefSwarmVec *)calloc(1, sizeof(RefSwarmVec));
    v->envs = (DroneSwarm *)calloc((size_t)n, sizeof(DroneSwarm));
    v->n = n;
    for (int i = 0; i < n; i++) {
        DroneSwarm *e = &v->envs[i];
        size_t a0 = (size_t)i * num_agents;
        e->observations = obs + a0 * 41;
        e->actions = act + a0 * 4;
        e->rewards = rew + a0;
        e->terminals = term + a0;
        e->num_agents = num_agents;
        e->max_rings = max_rings;
        init(e);
    }
    return v;
}

void refswarm_reset(void *vp, int seed) {
    RefSwarmVec *v = (RefSwarmVec *)vp;
    for (int i = 0; i < v->n; i++) {
        srand(i + seed * v->n);
        c_reset(&v->envs[i]);
    }
}

void refswarm_step(void *vp) {
    RefSwarmVec *v = (RefSwarmVec *)vp;
    for (int i = 0; i < v->n; i++) c_step(&v->envs[i]);
}

void refswarm_step_range(void *vp, int lo, int hi) {
    RefSwarmVec *v = (RefSwarmVec *)vp;
    for (int i = lo; i < hi; i++) c_step(&v->envs[i]);
}

void refswarm_log(void *vp, float out[9]) {
    RefSwarmVec *v = (RefSwarmVec *)vp;
    for (int j = 0; j < 9; j++) out[j] = 0.0f;
    for (int i = 0; i < v->n; i++) {
        float *l = (float *)&v->envs[i].log;
        for (int j = 0; j < 9; j++) {
            out[j] += l[j];
            l[j] = 0.0f;
        }
    }
}

void refswarm_get_env(void *vp, int i, float *b) {
    DroneSwarm *e = &((RefSwarmVec *)vp)->envs[i];
    b[0] = (float)e->tick;
    b[1] = (float)e->task;
    for (int r = 0; r < e->max_rings; r++) {
        Ring *g = &e->ring_buffer[r];
        float *o = b + 2 + 6 * r;
        o[0] = g->pos.x; o[1] = g->pos.y; o[2] = g->pos.z;
        o[3] = g->normal.x; o[4] = g->normal.y; o[5] = g->normal.z;
    }
}

void refswarm_put_env(void *vp, int i, const float *b) {
    DroneSwarm *e = &((RefSwarmVec *)vp)->envs[i];
    e->tick = (int)b[0];
    e->task = (int)b[1];
    for (int r = 0; r < e->max_rings; r++) {
        Ring *g = &e->ring_buffer[r];
        const float *o = b + 2 + 6 * r;
        g->pos = (Vec3){o[0], o[1], o[2]};
        g->normal = (Vec3){o[3], o[4], o[5]};
        g->radius = (o[3] == 0.0f && o[4] == 0.0f && o[5] == 0.0f) ? 0.0f : 2.0f;
    }
}

void refswarm_get_agent(void *vp, int i, int a, float *b) {
    Drone *d = &((RefSwarmVec *)vp)->envs[i].agents[a];
    State *s = &d->state;
    Params *p = &d->params;
    b[0] = s->pos.x; b[1] = s->pos.y; b[2] = s->pos.z;
    b[3] = s->vel.x; b[4] = s->vel.y; b[5] = s->vel.z;
    b[6] = s->quat.w; b[7] = s->quat.x; b[8] = s->quat.y; b[9] = s->quat.z;
    b[10] = s->omega.x; b[11] = s->omega.y; b[12] = s->omega.z;
    for (int k = 0; k < 4; k++) b[13 + k] = s->rpms[k];
    b[17] = p->mass; b[18] = p->ixx; b[19] = p->iyy; b[20] = p->izz;
    b[21] = p->arm_len; b[22] = p->k_thrust; b[23] = p->k_ang_damp; b[24] = p->k_drag;
    b[25] = p->b_drag; b[26] = p->gravity; b[27] = p->max_rpm; b[28] = p->k_mot; b[29] = p->j_mot;
    b[30] = d->spawn_pos.x; b[31] = d->spawn_pos.y; b[32] = d->spawn_pos.z;
    b[33] = d->target_pos.x; b[34] = d->target_pos.y; b[35] = d->target_pos.z;
    b[36] = d->target_vel.x; b[37] = d->target_vel.y; b[38] = d->target_vel.z;
    b[39] = d->last_abs_reward; b[40] = d->last_target_reward; b[41] = d->last_collision_reward;
    b[42] = d->episode_return; b[43] = d->collisions; b[44] = (float)d->episode_length;
    b[45] = d->score; b[46] = (float)d->ring_idx;
}

void refswarm_put_agent(void *vp, int i, int a, const float *b) {
    Drone *d = &((RefSwarmVec *)vp)->envs[i].agents[a];
    State *s = &d->state;
    Params *p = &d->params;
    s->pos = (Vec3){b[0], b[1], b[2]};
    s->vel = (Vec3){b[3], b[4], b[5]};
    s->quat = (Quat){b[6], b[7], b[8], b[9]};
    s->omega = (Vec3){b[10], b[11], b[12]};
    for (int k = 0; k < 4; k++) s->rpms[k] = b[13 + k];
    p->mass = b[17]; p->ixx = b[18]; p->iyy = b[19]; p->izz = b[20];
    p->arm_len = b[21]; p->k_thrust = b[22]; p->k_ang_damp = b[23]; p->k_drag = b[24];
    p->b_drag = b[25]; p->gravity = b[26]; p->max_rpm = b[27]; p->k_mot = b[28]; p->j_mot = b[29];
    p->max_vel = BASE_MAX_VEL; p->max_omega = BASE_MAX_OMEGA;
    d->spawn_pos = (Vec3){b[30], b[31], b[32]};
    d->prev_pos = s->pos;
    d->target_pos = (Vec3){b[33], b[34], b[35]};
    d->target_vel = (Vec3){b[36], b[37], b[38]};
    d->last_abs_reward = b[39]; d->last_target_reward = b[40]; d->last_collision_reward = b[41];
    d->episode_return = b[42]; d->collisions = b[43]; d->episode_length = (int)b[44];
    d->score = b[45]; d->ring_idx = (int)b[46];
}

void refswarm_observe(void *vp, int i) { compute_observations(&((RefSwarmVec *)vp)->envs[i]); }

void refswarm_close(void *vp) {
    RefSwarmVec *v = (RefSwarmVec *)vp;
    for (int i = 0; i < v->n; i++) {
        free(v->envs[i].agents);
        free(v->envs[i].ring_buffer);
    }
    free(v->envs);
    free(v);
}

int refswarm_sizeof_env(void) { return (int)sizeof(DroneSwarm); }
int refswarm_sizeof_drone(void) { return (int)sizeof(Drone); }
int refswarm_sizeof_ring(void) { return (int)sizeof(Ring); }

/* the reference's compute_observations on caller-provided reference structs (see ref_shim_race.c) */
void refswarm_observe_structs(Drone *agents, int num_agents, Ring *rings, int max_rings, int task, float *obs) {
    DroneSwarm env;
    memset(&env, 0, sizeof(env));
    env.observations = obs;
    env.num_agents = num_agents;
    env.agents = agents;
    env.max_rings = max_rings;
    env.ring_buffer = rings;
    env.task = task;
    compute_observations(&env);
}
