/* drone_oracle.c -- CPU restatement of the drone env step.  TEST INFRASTRUCTURE ONLY.
 *
 * See drone_oracle.h for the rules (who may load this) and the parity pin.
 * Every float operation below is a single IEEE-754 binary32 operation in the
 * association the reference source writes; build with -O2 -ffp-contract=off
 * (oracle/Makefile) so nothing is fused or reassociated.
 *
 * R = /root/reference/pufferlib/pufferlib/ocean
 */
#include "drone_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ constants */
/* R/drone_race/dronelib.h:22-47 */
#define K_MASS 1.0f
#define K_IXX 0.01f
#define K_IYY 0.01f
#define K_IZZ 0.02f
#define K_ARM 0.1f
#define K_THRUST 3e-5f
#define K_ANG_DAMP 0.2f
#define K_DRAG 1e-6f
#define K_BDRAG 0.1f
#define K_GRAV 9.81f
#define K_MAX_RPM 750.0f
#define K_MAX_VEL 50.0f
#define K_MAX_OMEGA 50.0f
#define K_KMOT 0.1f
#define K_JMOT 1e-5f
#define K_DT 0.05f
#define RING_RADIUS 2.0f
#define MAX_ATTEMPTS 16

typedef struct { float x, y, z; } v3;
typedef struct { float w, x, y, z; } q4;

/* indices into the 13-float parameter row */
enum { P_MASS, P_IXX, P_IYY, P_IZZ, P_ARM, P_KT, P_KAD, P_KD, P_BD, P_G, P_MRPM, P_KMOT, P_JMOT, P_N };

/* ------------------------------------------------------------------ small math */
/* R/drone_race/dronelib.h:73-79 (NaN falls through both tests) */
static inline float clampf_(float v, float lo, float hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}

/* R/drone_race/dronelib.h:112-119: Hamilton product, left-to-right sums */
static inline q4 qmul(q4 a, q4 b) {
    q4 r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return r;
}

/* R/drone_race/dronelib.h:131-137: v' = q (0,v) q*  */
static inline v3 qrot(q4 q, v3 v) {
    q4 pure = {0.0f, v.x, v.y, v.z};
    q4 t = qmul(q, pure);
    q4 qc = {q.w, -q.x, -q.y, -q.z};
    q4 r = qmul(t, qc);
    v3 o = {r.x, r.y, r.z};
    return o;
}

/* R/drone_race/dronelib.h:121-129 */
static inline void qnormalize(q4 *q) {
    float n = sqrtf(q->w * q->w + q->x * q->x + q->y * q->y + q->z * q->z);
    if (n > 0.0f) {
        q->w /= n;
        q->x /= n;
        q->y /= n;
        q->z /= n;
    }
}

static inline float dist3(v3 a, v3 b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return sqrtf(dx * dx + dy * dy + dz * dz);
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Deterministic sin/cos for theta in [0, 2*pi]: double arithmetic, one IEEE op
 * at a time (no FMA), so the device reproduces it bit-for-bit with
 * __dmul_rn/__dadd_rn.  Quadrant reduction + Taylor polynomials on |r|<=pi/4
 * (truncation error < 5e-17), result rounded once to float. */
void orc_sincos_det(float theta, float *s, float *c) {
    /* theta in [0, 2 pi].  Single-precision throughout, one IEEE operation at a time (this file is built with
     * -ffp-contract=off; the device twin b2d_math.cuh:sincos_det uses __fadd_rn / __fmul_rn): quadrant
     * k = floor(theta * 2/pi + 1/2) <= 4, three-term Cody-Waite reduction r = theta - k pi/2 (the products
     * k * DP1 and k * DP2 are exact), minimax polynomials on [-pi/4, pi/4] (Cephes sinf / cosf). */
    const float DP1 = 1.5703125f, DP2 = 4.837512969970703125e-4f, DP3 = 7.54978995489188216e-8f;
    float t = theta;
    int k = (int)(t * 0.636619746685028076171875f + 0.5f);
    float kf = (float)k;
    float r = ((t - kf * DP1) - kf * DP2) - kf * DP3;
    float z = r * r;
    float ps = -1.9515295891e-4f;
    ps = ps * z + 8.3321608736e-3f;
    ps = ps * z + -1.6666654611e-1f;
    float sr = ps * z * r + r;
    float pc = 2.443315711809948e-5f;
    pc = pc * z + -1.388731625493765e-3f;
    pc = pc * z + 4.166664568298827e-2f;
    float cr = pc * z * z + (1.0f - 0.5f * z);
    float sv, cv;
    switch (k & 3) {
    case 0: sv = sr; cv = cr; break;
    case 1: sv = cr; cv = -sr; break;
    case 2: sv = -sr; cv = -cr; break;
    default: sv = -cr; cv = sr; break;
    }
    *s = sv;
    *c = cv;
}

/* ------------------------------------------------------------------ random sources */
typedef struct {
    int mode;          /* ORC_RESET_LIBC or ORC_RESET_PHILOX */
    uint32_t key[2];
    uint32_t env, episode;
} RandSrc;

/* R/drone_race/dronelib.h:81-83; (float)RAND_MAX == 2^31 */
static inline float u_from_i31(int32_t r) { return (float)r / 2147483648.0f; }
static inline float lerp_u(float a, float b, float u) { return a + u * (b - a); }
static inline float libc_rndf(float a, float b) { return lerp_u(a, b, u_from_i31(rand())); }

static void philox_words(const RandSrc *rs, uint32_t item, uint32_t attempt, uint32_t w[4]) {
    uint32_t ctr[4] = {rs->env, rs->episode, item, attempt};
    orc_philox4x32_10(ctr, rs->key, w);
}
static inline float u_from_word(uint32_t w) { return u_from_i31((int32_t)(w >> 1)); }

/* ------------------------------------------------------------------ race env */
struct OrcRace {
    int n, max_rings, max_moves;
    float *state;   /* [n][17] pos3 vel3 quat4 omega3 rpm4 */
    float *params;  /* [n][13] */
    int *tick;
    int *ring_idx;
    float *ep_ret;
    float *rings;   /* [n][max_rings][6] pos3 normal3 */
    float *logs;    /* [n][9]  R/drone_race/dronelib.h:52-63 field order */
    uint32_t key[2];
    uint32_t env_id_base;
    uint32_t epoch; /* number of vec steps taken since the last vec_reset */
    uint32_t *episode; /* [n] number of the live episode of each env (Philox counter word) */
};

OrcRace *orc_race_create(int n, int max_rings, int max_moves) {
    OrcRace *o = (OrcRace *)calloc(1, sizeof(OrcRace));
    o->n = n;
    o->max_rings = max_rings;
    o->max_moves = max_moves;
    o->state = (float *)calloc((size_t)n * 17, sizeof(float));
    o->params = (float *)calloc((size_t)n * P_N, sizeof(float));
    o->tick = (int *)calloc((size_t)n, sizeof(int));
    o->ring_idx = (int *)calloc((size_t)n, sizeof(int));
    o->ep_ret = (float *)calloc((size_t)n, sizeof(float));
    o->rings = (float *)calloc((size_t)n * max_rings * 6, sizeof(float));
    o->logs = (float *)calloc((size_t)n * 9, sizeof(float));
    o->episode = (uint32_t *)calloc((size_t)n, sizeof(uint32_t));
    for (int i = 0; i < n; i++) o->state[(size_t)i * 17 + 6] = 1.0f;
    return o;
}

void orc_race_close(OrcRace *o) {
    if (!o) return;
    free(o->state); free(o->params); free(o->tick); free(o->ring_idx);
    free(o->ep_ret); free(o->rings); free(o->logs); free(o->episode); free(o);
}

void orc_race_set_philox(OrcRace *o, uint64_t seed, uint32_t env_id_base) {
    o->key[0] = (uint32_t)seed;
    o->key[1] = (uint32_t)(seed >> 32);
    o->env_id_base = env_id_base;
}

uint32_t orc_race_epoch(const OrcRace *o) { return o->epoch; }
void orc_race_set_epoch(OrcRace *o, uint32_t epoch) { o->epoch = epoch; }

/* ring normal from the three quaternion uniforms.
 * R/drone_race/dronelib.h:141-159 (rndquat) and :177-178 (normal = q . z-axis) */
static v3 ring_normal_from_u(float u1, float u2, float u3, int det_trig) {
    float a = sqrtf(1.0f - u1);
    float b = sqrtf(u1);
    float th2 = (float)(2.0f * M_PI * u2); /* product in double, rounded once */
    float th3 = (float)(2.0f * M_PI * u3);
    float s2, c2, s3, c3;
    if (det_trig) {
        orc_sincos_det(th2, &s2, &c2);
        orc_sincos_det(th3, &s3, &c3);
    } else {
        s2 = sinf(th2); c2 = cosf(th2);
        s3 = sinf(th3); c3 = cosf(th3);
    }
    q4 q = {a * s2, a * c2, b * s3, b * c3};
    v3 z = {0.0f, 0.0f, 1.0f};
    return qrot(q, z);
}

/* x^3 the way the reference's powf(x, 3.0f) call rounds it (libm: correctly
 * rounded in practice); the Philox stream uses an explicit double product so
 * the device can reproduce it. */
static inline float cube_det(float x) { return (float)(((double)x * (double)x) * (double)x); }

/* R/drone_race/dronelib.h:250-290: scale laws + 12 jitter factors u[0..11] in draw order */
static void params_from_draws(float size, const float *uj, int det_pow, float *p) {
    float arm = size / 2.0f;
    float cube = det_pow ? cube_det(arm) : powf(arm, 3.0f);
    float base_cube = det_pow ? cube_det(K_ARM) : powf(K_ARM, 3.0f);
    float mass_scale = cube / base_cube;
    float mass = K_MASS * mass_scale * uj[0];
    float base_iscale = K_MASS * K_ARM * K_ARM;
    float iscale = mass * (arm * arm) / base_iscale;
    float ixx = K_IXX * iscale * uj[1];
    float iyy = K_IYY * iscale * uj[2];
    float izz = K_IZZ * iscale * uj[3];
    float kt_scale = (mass * arm) / (K_MASS * K_ARM);
    float kt = K_THRUST * kt_scale * uj[4];
    float base_avg = (K_IXX + K_IYY + K_IZZ) / 3.0f;
    float avg = (ixx + iyy + izz) / 3.0f;
    float avg_scale = avg / base_avg;
    float kad = K_ANG_DAMP * avg_scale * uj[5];
    float drag_scale = (arm * arm) / (K_ARM * K_ARM);
    float kd = K_DRAG * drag_scale * uj[6];
    float bd = K_BDRAG * drag_scale * uj[7];
    float g = K_GRAV * uj[8];
    float rpm_scale = K_ARM / arm;
    float mrpm = K_MAX_RPM * rpm_scale * uj[9];
    float kmot = K_KMOT * uj[10];
    float jmot = K_JMOT * iscale * uj[11];
    p[P_MASS] = mass; p[P_IXX] = ixx; p[P_IYY] = iyy; p[P_IZZ] = izz; p[P_ARM] = arm;
    p[P_KT] = kt; p[P_KAD] = kad; p[P_KD] = kd; p[P_BD] = bd; p[P_G] = g;
    p[P_MRPM] = mrpm; p[P_KMOT] = kmot; p[P_JMOT] = jmot;
}

/* jitter intervals: 11 x U(1-dr, 1+dr) with gravity (index 8) U(0.99, 1.01); dr = 0.1f */
static inline void jitter_bounds(int k, float *lo, float *hi) {
    if (k == 8) { *lo = 0.99f; *hi = 1.01f; }
    else { *lo = 1.0f - 0.1f; *hi = 1.0f + 0.1f; }
}

static void zero_motion(float *s) {
    for (int k = 0; k < 17; k++) s[k] = 0.0f;
    s[6] = 1.0f;
}

/* R/drone_race/drone_race.h:127-151 with the RNG abstracted.
 * size_lo/size_hi, the ring box and the spawn box are parameters so the swarm
 * env can reuse the pieces. */
static void race_fresh_episode(OrcRace *o, int i, const RandSrc *rs) {
    float *rings = o->rings + (size_t)i * o->max_rings * 6;
    float *s = o->state + (size_t)i * 17;
    float *p = o->params + (size_t)i * P_N;
    const float ring_lo = -10.0f + 2 * RING_RADIUS, ring_hi = 10.0f - 2 * RING_RADIUS;
    const float min_gap = 2.0f * RING_RADIUS;

    o->tick[i] = 0;
    o->ring_idx[i] = 0;
    o->ep_ret[i] = 0.0f;

    if (rs->mode == ORC_RESET_LIBC) {
        /* rings: R/drone_race/dronelib.h:451-460, draw order x,y,z,u1,u2,u3 per attempt */
        for (int r = 0; r < o->max_rings; r++) {
            float *g = rings + 6 * r;
            for (;;) {
                v3 c;
                c.x = libc_rndf(ring_lo, ring_hi);
                c.y = libc_rndf(ring_lo, ring_hi);
                c.z = libc_rndf(ring_lo, ring_hi);
                float u1 = libc_rndf(0.0f, 1.0f);
                float u2 = libc_rndf(0.0f, 1.0f);
                float u3 = libc_rndf(0.0f, 1.0f);
                v3 nrm = ring_normal_from_u(u1, u2, u3, 0);
                g[0] = c.x; g[1] = c.y; g[2] = c.z;
                g[3] = nrm.x; g[4] = nrm.y; g[5] = nrm.z;
                if (r == 0) break;
                v3 prev = {g[-6], g[-5], g[-4]};
                if (!(dist3(c, prev) < min_gap)) break;
            }
        }
        float size = libc_rndf(0.05f, 0.8f);
        float uj[12];
        for (int k = 0; k < 12; k++) {
            float lo, hi;
            jitter_bounds(k, &lo, &hi);
            uj[k] = libc_rndf(lo, hi);
        }
        params_from_draws(size, uj, 0, p);
        zero_motion(s);
        v3 r0 = {rings[0], rings[1], rings[2]};
        for (;;) {
            v3 c;
            c.x = libc_rndf(-9.0f, 9.0f);
            c.y = libc_rndf(-9.0f, 9.0f);
            c.z = libc_rndf(-9.0f, 9.0f);
            s[0] = c.x; s[1] = c.y; s[2] = c.z;
            if (!(dist3(c, r0) < min_gap)) break;
        }
    } else {
        /* device stream: counter = (env, episode number, item, attempt); see DESIGN.md "reset stream" */
        uint32_t w[4], w2[4];
        for (int r = 0; r < o->max_rings; r++) {
            float *g = rings + 6 * r;
            for (uint32_t t = 0; t < MAX_ATTEMPTS; t++) {
                philox_words(rs, 2u * r, t, w);
                philox_words(rs, 2u * r + 1u, t, w2);
                v3 c;
                c.x = lerp_u(ring_lo, ring_hi, u_from_word(w[0]));
                c.y = lerp_u(ring_lo, ring_hi, u_from_word(w[1]));
                c.z = lerp_u(ring_lo, ring_hi, u_from_word(w[2]));
                v3 nrm = ring_normal_from_u(u_from_word(w[3]), u_from_word(w2[0]), u_from_word(w2[1]), 1);
                g[0] = c.x; g[1] = c.y; g[2] = c.z;
                g[3] = nrm.x; g[4] = nrm.y; g[5] = nrm.z;
                if (r == 0) break;
                v3 prev = {g[-6], g[-5], g[-4]};
                if (!(dist3(c, prev) < min_gap)) break;
            }
        }
        float draws[16];
        for (uint32_t k = 0; k < 4; k++) {
            philox_words(rs, 0x1000u + k, 0, w);
            for (int j = 0; j < 4; j++) draws[4 * k + j] = u_from_word(w[j]);
        }
        float size = lerp_u(0.05f, 0.8f, draws[0]);
        float uj[12];
        for (int k = 0; k < 12; k++) {
            float lo, hi;
            jitter_bounds(k, &lo, &hi);
            uj[k] = lerp_u(lo, hi, draws[1 + k]);
        }
        params_from_draws(size, uj, 1, p);
        zero_motion(s);
        v3 r0 = {rings[0], rings[1], rings[2]};
        for (uint32_t t = 0; t < MAX_ATTEMPTS; t++) {
            philox_words(rs, 0x2000u, t, w);
            v3 c;
            c.x = lerp_u(-9.0f, 9.0f, u_from_word(w[0]));
            c.y = lerp_u(-9.0f, 9.0f, u_from_word(w[1]));
            c.z = lerp_u(-9.0f, 9.0f, u_from_word(w[2]));
            s[0] = c.x; s[1] = c.y; s[2] = c.z;
            if (!(dist3(c, r0) < min_gap)) break;
        }
    }
}

void orc_race_put_state(OrcRace *o, int i, const float *b) {
    memcpy(o->state + (size_t)i * 17, b, 17 * sizeof(float));
    memcpy(o->params + (size_t)i * P_N, b + 17, P_N * sizeof(float));
    o->tick[i] = (int)b[30];
    o->ring_idx[i] = (int)b[31];
    o->ep_ret[i] = b[32];
    memcpy(o->rings + (size_t)i * o->max_rings * 6, b + ORC_RACE_BLOB,
           (size_t)o->max_rings * 6 * sizeof(float));
}

void orc_race_get_state(const OrcRace *o, int i, float *b) {
    memcpy(b, o->state + (size_t)i * 17, 17 * sizeof(float));
    memcpy(b + 17, o->params + (size_t)i * P_N, P_N * sizeof(float));
    b[30] = (float)o->tick[i];
    b[31] = (float)o->ring_idx[i];
    b[32] = o->ep_ret[i];
    memcpy(b + ORC_RACE_BLOB, o->rings + (size_t)i * o->max_rings * 6,
           (size_t)o->max_rings * 6 * sizeof(float));
}

/* R/drone_race/drone_race.h:72-125 */
void orc_race_observe(const OrcRace *o, int i, float *ob) {
    const float *s = o->state + (size_t)i * 17;
    const float *p = o->params + (size_t)i * P_N;
    const float *g = o->rings + ((size_t)i * o->max_rings + o->ring_idx[i]) * 6;
    /* "next" ring = ring_buffer[ring_idx % max_rings] == the current one (:77) */
    const float *g2 = o->rings + ((size_t)i * o->max_rings + (o->ring_idx[i] % o->max_rings)) * 6;
    q4 q = {s[6], s[7], s[8], s[9]};
    q4 qi = {q.w, -q.x, -q.y, -q.z};
    v3 pos = {s[0], s[1], s[2]};
    v3 vel = {s[3], s[4], s[5]};
    v3 d1 = {g[0] - pos.x, g[1] - pos.y, g[2] - pos.z};
    v3 d2 = {g2[0] - pos.x, g2[1] - pos.y, g2[2] - pos.z};
    v3 n1 = {g[3], g[4], g[5]};
    v3 n2 = {g2[3], g2[4], g2[5]};
    v3 to1 = qrot(qi, d1), to2 = qrot(qi, d2);
    v3 bn1 = qrot(qi, n1), bn2 = qrot(qi, n2);
    v3 vb = qrot(qi, vel);
    v3 zax = {0.0f, 0.0f, 1.0f};
    v3 up = qrot(q, zax);
    ob[0] = to1.x / 10.0f; ob[1] = to1.y / 10.0f; ob[2] = to1.z / 10.0f;
    ob[3] = bn1.x; ob[4] = bn1.y; ob[5] = bn1.z;
    ob[6] = to2.x / 10.0f; ob[7] = to2.y / 10.0f; ob[8] = to2.z / 10.0f;
    ob[9] = bn2.x; ob[10] = bn2.y; ob[11] = bn2.z;
    ob[12] = vb.x / K_MAX_VEL; ob[13] = vb.y / K_MAX_VEL; ob[14] = vb.z / K_MAX_VEL;
    ob[15] = s[10] / K_MAX_OMEGA; ob[16] = s[11] / K_MAX_OMEGA; ob[17] = s[12] / K_MAX_OMEGA;
    ob[18] = up.x; ob[19] = up.y; ob[20] = up.z;
    ob[21] = q.w; ob[22] = q.x; ob[23] = q.y; ob[24] = q.z;
    for (int k = 0; k < 4; k++) ob[25 + k] = s[13 + k] / p[P_MRPM];
}

/* ---- dynamics --------------------------------------------------------------- */
typedef struct {
    v3 pos, vel;
    q4 q;
    v3 w;
    float rpm[4];
} Body;
typedef struct {
    v3 dpos, dvel;
    q4 dq;
    v3 dw;
    float drpm[4];
} Rate;

/* R/drone_race/dronelib.h:302-381 */
static void rates(const Body *b, const float *p, const float *act, Rate *k) {
    float want[4], thrust[4];
    for (int m = 0; m < 4; m++) want[m] = (act[m] + 1.0f) * 0.5f * p[P_MRPM];
    for (int m = 0; m < 4; m++) k->drpm[m] = (1.0f / p[P_KMOT]) * (want[m] - b->rpm[m]);
    for (int m = 0; m < 4; m++) thrust[m] = p[P_KT] * (b->rpm[m] * b->rpm[m]); /* powf(x,2) == x*x */

    v3 lift_body = {0.0f, 0.0f, thrust[0] + thrust[1] + thrust[2] + thrust[3]};
    v3 lift = qrot(b->q, lift_body);
    float dragx = -p[P_BD] * b->vel.x;
    float dragy = -p[P_BD] * b->vel.y;
    float dragz = -p[P_BD] * b->vel.z;
    k->dvel.x = (lift.x + dragx) / p[P_MASS];
    k->dvel.y = (lift.y + dragy) / p[P_MASS];
    k->dvel.z = ((lift.z + dragz) / p[P_MASS]) - p[P_G];

    q4 wq = {0.0f, b->w.x, b->w.y, b->w.z};
    q4 dq = qmul(b->q, wq);
    k->dq.w = dq.w * 0.5f; k->dq.x = dq.x * 0.5f; k->dq.y = dq.y * 0.5f; k->dq.z = dq.z * 0.5f;

    float tpx = p[P_ARM] * (thrust[1] - thrust[3]);
    float tpy = p[P_ARM] * (thrust[2] - thrust[0]);
    float tpz = p[P_KD] * (thrust[0] - thrust[1] + thrust[2] - thrust[3]);
    float tmz = p[P_JMOT] * (k->drpm[0] - k->drpm[1] + k->drpm[2] - k->drpm[3]);
    float tax = -p[P_KAD] * b->w.x;
    float tay = -p[P_KAD] * b->w.y;
    float taz = -p[P_KAD] * b->w.z;
    float tix = (p[P_IYY] - p[P_IZZ]) * b->w.y * b->w.z;
    float tiy = (p[P_IZZ] - p[P_IXX]) * b->w.z * b->w.x;
    float tiz = (p[P_IXX] - p[P_IYY]) * b->w.x * b->w.y;
    k->dw.x = (tpx + tax + tix) / p[P_IXX];
    k->dw.y = (tpy + tay + tiy) / p[P_IYY];
    k->dw.z = (tpz + taz + tiz + tmz) / p[P_IZZ];
    k->dpos = b->vel;
}

/* R/drone_race/dronelib.h:383-392 */
static void euler_probe(const Body *b, const Rate *k, float h, Body *o) {
    o->pos.x = b->pos.x + k->dpos.x * h; o->pos.y = b->pos.y + k->dpos.y * h; o->pos.z = b->pos.z + k->dpos.z * h;
    o->vel.x = b->vel.x + k->dvel.x * h; o->vel.y = b->vel.y + k->dvel.y * h; o->vel.z = b->vel.z + k->dvel.z * h;
    o->q.w = b->q.w + k->dq.w * h; o->q.x = b->q.x + k->dq.x * h;
    o->q.y = b->q.y + k->dq.y * h; o->q.z = b->q.z + k->dq.z * h;
    o->w.x = b->w.x + k->dw.x * h; o->w.y = b->w.y + k->dw.y * h; o->w.z = b->w.z + k->dw.z * h;
    for (int m = 0; m < 4; m++) o->rpm[m] = b->rpm[m] + k->drpm[m] * h;
    qnormalize(&o->q);
}

#define RK_MIX(a, b, c, d) (((a) + 2.0f * (b) + 2.0f * (c) + (d)) * h6)

/* R/drone_race/dronelib.h:394-449: clamp actions in place, RK4 at dt=DT, clamp vel/omega */
static void advance_body(float *s, const float *p, float *act) {
    for (int m = 0; m < 4; m++) act[m] = clampf_(act[m], -1.0f, 1.0f);
    Body b, tmp;
    Rate k1, k2, k3, k4;
    b.pos = (v3){s[0], s[1], s[2]};
    b.vel = (v3){s[3], s[4], s[5]};
    b.q = (q4){s[6], s[7], s[8], s[9]};
    b.w = (v3){s[10], s[11], s[12]};
    for (int m = 0; m < 4; m++) b.rpm[m] = s[13 + m];
    /* dt = DT * rndf(1, 1) == DT exactly (DT_RNG = 0); the draw itself is consumed by the caller */
    const float h = K_DT * 1.0f;
    rates(&b, p, act, &k1);
    euler_probe(&b, &k1, h * 0.5f, &tmp);
    rates(&tmp, p, act, &k2);
    euler_probe(&b, &k2, h * 0.5f, &tmp);
    rates(&tmp, p, act, &k3);
    euler_probe(&b, &k3, h, &tmp);
    rates(&tmp, p, act, &k4);
    const float h6 = h / 6.0f;
    b.pos.x += RK_MIX(k1.dpos.x, k2.dpos.x, k3.dpos.x, k4.dpos.x);
    b.pos.y += RK_MIX(k1.dpos.y, k2.dpos.y, k3.dpos.y, k4.dpos.y);
    b.pos.z += RK_MIX(k1.dpos.z, k2.dpos.z, k3.dpos.z, k4.dpos.z);
    b.vel.x += RK_MIX(k1.dvel.x, k2.dvel.x, k3.dvel.x, k4.dvel.x);
    b.vel.y += RK_MIX(k1.dvel.y, k2.dvel.y, k3.dvel.y, k4.dvel.y);
    b.vel.z += RK_MIX(k1.dvel.z, k2.dvel.z, k3.dvel.z, k4.dvel.z);
    b.q.w += RK_MIX(k1.dq.w, k2.dq.w, k3.dq.w, k4.dq.w);
    b.q.x += RK_MIX(k1.dq.x, k2.dq.x, k3.dq.x, k4.dq.x);
    b.q.y += RK_MIX(k1.dq.y, k2.dq.y, k3.dq.y, k4.dq.y);
    b.q.z += RK_MIX(k1.dq.z, k2.dq.z, k3.dq.z, k4.dq.z);
    b.w.x += RK_MIX(k1.dw.x, k2.dw.x, k3.dw.x, k4.dw.x);
    b.w.y += RK_MIX(k1.dw.y, k2.dw.y, k3.dw.y, k4.dw.y);
    b.w.z += RK_MIX(k1.dw.z, k2.dw.z, k3.dw.z, k4.dw.z);
    for (int m = 0; m < 4; m++) b.rpm[m] += RK_MIX(k1.drpm[m], k2.drpm[m], k3.drpm[m], k4.drpm[m]);
    qnormalize(&b.q);
    b.vel.x = clampf_(b.vel.x, -K_MAX_VEL, K_MAX_VEL);
    b.vel.y = clampf_(b.vel.y, -K_MAX_VEL, K_MAX_VEL);
    b.vel.z = clampf_(b.vel.z, -K_MAX_VEL, K_MAX_VEL);
    b.w.x = clampf_(b.w.x, -K_MAX_OMEGA, K_MAX_OMEGA);
    b.w.y = clampf_(b.w.y, -K_MAX_OMEGA, K_MAX_OMEGA);
    b.w.z = clampf_(b.w.z, -K_MAX_OMEGA, K_MAX_OMEGA);
    s[0] = b.pos.x; s[1] = b.pos.y; s[2] = b.pos.z;
    s[3] = b.vel.x; s[4] = b.vel.y; s[5] = b.vel.z;
    s[6] = b.q.w; s[7] = b.q.x; s[8] = b.q.y; s[9] = b.q.z;
    s[10] = b.w.x; s[11] = b.w.y; s[12] = b.w.z;
    for (int m = 0; m < 4; m++) s[13 + m] = b.rpm[m];
}

/* R/drone_race/dronelib.h:462-489.  edge_value is what a rim hit returns:
 * -1.0f in the race copy, -0.0f in the swarm copy (:485). */
static float gate_event(v3 before, v3 after, const float *g, float edge_value) {
    v3 c = {g[0], g[1], g[2]};
    v3 nrm = {g[3], g[4], g[5]};
    v3 a = {before.x - c.x, before.y - c.y, before.z - c.z};
    v3 b = {after.x - c.x, after.y - c.y, after.z - c.z};
    float d0 = a.x * nrm.x + a.y * nrm.y + a.z * nrm.z;
    float d1 = b.x * nrm.x + b.y * nrm.y + b.z * nrm.z;
    int forward = (d0 < 0.0f && d1 > 0.0f);
    int backward = (d0 > 0.0f && d1 < 0.0f);
    if (forward || backward) {
        v3 dir = {after.x - before.x, after.y - before.y, after.z - before.z};
        float t = -d0 / (nrm.x * dir.x + nrm.y * dir.y + nrm.z * dir.z);
        v3 hit = {before.x + dir.x * t, before.y + dir.y * t, before.z + dir.z * t};
        float r = dist3(hit, c);
        /* radius +- 0.5 evaluated in double by the reference; 1.5 and 2.5 are exact */
        if ((double)r < (double)RING_RADIUS - 0.5 && forward) return 1.0f;
        if ((double)r < (double)RING_RADIUS + 0.5) return edge_value;
    }
    return 0.0f;
}

/* R/drone_race/drone_race.h:61-70 */
static void race_log_episode(OrcRace *o, int i, float oob, float hit, float timeout) {
    float *l = o->logs + (size_t)i * 9;
    l[6] += (float)o->ring_idx[i]; /* score == ring_idx at every add_log call site */
    l[0] += o->ep_ret[i];
    l[1] += (float)o->tick[i];
    l[7] += (float)o->ring_idx[i] / (float)o->max_rings;
    l[4] += oob;
    l[3] += hit;
    l[5] += timeout;
    l[8] += 1.0f;
}

/* first = 1 for vec_reset (episode 0), 0 for an auto-reset (next episode number) */
static void race_new_episode(OrcRace *o, int i, int mode, const float *payload, int first) {
    if (mode == ORC_RESET_INJECT) {
        orc_race_put_state(o, i, payload + (size_t)i * (ORC_RACE_BLOB + 6 * o->max_rings));
    } else {
        o->episode[i] = first ? 0u : o->episode[i] + 1u;
        RandSrc rs = {mode, {o->key[0], o->key[1]}, o->env_id_base + (uint32_t)i, o->episode[i]};
        race_fresh_episode(o, i, &rs);
    }
}

void orc_race_reset(OrcRace *o, int mode, int seed, const float *payload, float *obs) {
    o->epoch = 0;
    for (int i = 0; i < o->n; i++) {
        if (mode == ORC_RESET_LIBC) srand(i + seed * o->n); /* EB:500-504 */
        race_new_episode(o, i, mode, payload, 1);
        if (obs) orc_race_observe(o, i, obs + (size_t)i * ORC_RACE_OBS);
    }
}

/* R/drone_race/drone_race.h:156-208 for env i */
static void race_step_one(OrcRace *o, int i, int mode, float *actions, const float *payload,
                          float *obs, float *rew, unsigned char *term, unsigned char *events) {
    float *s = o->state + (size_t)i * 17;
    const float *p = o->params + (size_t)i * P_N;
    unsigned char ev = 0;
    o->tick[i] += 1;
    rew[i] = 0;
    term[i] = 0;
    o->logs[(size_t)i * 9 + 6] = 0.0f; /* log.score = 0 every step (:160) */

    v3 before = {s[0], s[1], s[2]};
    if (mode == ORC_RESET_LIBC) (void)rand(); /* the dt-jitter draw, R/drone_race/dronelib.h:440 */
    advance_body(s, p, actions + (size_t)i * 4);
    v3 after = {s[0], s[1], s[2]};

    int oob = after.x < -10.0f || after.x > 10.0f || after.y < -10.0f || after.y > 10.0f ||
              after.z < -10.0f || after.z > 10.0f;
    int ended = 0;
    if (oob) {
        rew[i] -= 1;
        o->ep_ret[i] -= 1;
        term[i] = 1;
        race_log_episode(o, i, 1.0f, 0.0f, 0.0f);
        ev |= ORC_EV_OOB;
        ended = 1;
    } else {
        const float *g = o->rings + ((size_t)i * o->max_rings + o->ring_idx[i]) * 6;
        float r = gate_event(before, after, g, -1.0f);
        rew[i] += r;
        o->ep_ret[i] += r;
        if (r > 0) {
            o->ring_idx[i] += 1;
            ev |= ORC_EV_RING_PASS;
        }
        if (r < 0) {
            term[i] = 1;
            race_log_episode(o, i, 0.0f, 1.0f, 0.0f);
            ev |= ORC_EV_COLLISION;
            ended = 1;
        } else {
            int moves_left = o->max_moves - o->tick[i];
            if (moves_left == 0 || o->ring_idx[i] == o->max_rings) {
                term[i] = 1;
                race_log_episode(o, i, 0.0f, 0.0f, moves_left == 0 ? 1.0f : 0.0f);
                ev |= (moves_left == 0) ? ORC_EV_TIMEOUT : ORC_EV_COMPLETE;
                ended = 1;
            }
        }
    }
    if (ended) race_new_episode(o, i, mode, payload, 0);
    orc_race_observe(o, i, obs + (size_t)i * ORC_RACE_OBS);
    if (events) events[i] = ev;
}

void orc_race_step_range(OrcRace *o, int lo, int hi, int mode, float *actions, const float *payload,
                         float *obs, float *rew, unsigned char *term, unsigned char *events) {
    for (int i = lo; i < hi; i++) race_step_one(o, i, mode, actions, payload, obs, rew, term, events);
}

void orc_race_step(OrcRace *o, int mode, float *actions, const float *payload, float *obs,
                   float *rew, unsigned char *term, unsigned char *events) {
    o->epoch += 1;
    orc_race_step_range(o, 0, o->n, mode, actions, payload, obs, rew, term, events);
}

void orc_race_log(OrcRace *o, float out[9]) {
    for (int j = 0; j < 9; j++) out[j] = 0.0f;
    for (int i = 0; i < o->n; i++) {
        float *l = o->logs + (size_t)i * 9;
        for (int j = 0; j < 9; j++) {
            out[j] += l[j];
            l[j] = 0.0f;
        }
    }
}

/* ====================================================================== swarm env
 * R/drone_swarm/drone_swarm.h:84-497 restated.  GRID = (30, 30, 10), MARGIN = GRID - 1
 * (R/drone_swarm/dronelib.h:39-44), MAX_DIST = sqrtf(60^2 + 60^2 + 20^2) (:50), HORIZON 1024,
 * rim hits return -0.0f (:485).
 *
 * Agents are processed in index order inside c_step; neighbour queries therefore see the
 * already-moved (and possibly re-spawned) positions of lower-index agents and the previous
 * tick's positions of higher-index agents.  The restatement keeps the reference's loop
 * structure literally, so that order dependence is reproduced by construction.
 *
 * Reset sources: libc rand() in the reference's draw order (ORC_RESET_LIBC, incl. the unused
 * dt-jitter draw per drone per step), the device's Philox stream (ORC_RESET_PHILOX), or
 * injected draw results (ORC_RESET_INJECT).  In every mode the results of the random draws of a
 * step can be RECORDED into / READ from per-agent and per-env payload rows:
 *   agent row [ORC_SWARM_AGENT_PAYLOAD = 41]:
 *     [0:13] params, [13:16] pos                      of an out-of-bounds respawn (reset_agent)
 *     [16:29] params, [29:32] pos, [32:35] target_pos, [35:38] target_vel, [38:41] race start pos
 *                                                     of the env-wide reset (c_reset)
 *   env row [2 + 6*max_rings]: [0] unused, [1] task, then rings pos(3) normal(3)
 *   flags: agent bit0 = respawned this step; env bit0 = env-wide reset this step
 */
#define S_GX 30.0f
#define S_GY 30.0f
#define S_GZ 10.0f
#define S_MX (S_GX - 1)
#define S_MY (S_GY - 1)
#define S_MZ (S_GZ - 1)
#define S_VT 0.05f
#define S_HORIZON 1024
#define S_TASK_RACE 7

enum { AX_SPAWN = 0, AX_TPOS = 3, AX_TVEL = 6, AX_LAST_ABS = 9, AX_LAST_TGT = 10, AX_LAST_COL = 11,
       AX_RETURN = 12, AX_COLLISIONS = 13, AX_SCORE = 14, AX_N = 15 };

struct OrcSwarm {
    int n, A, R;
    float *st;   /* [n][A][17] */
    float *pr;   /* [n][A][13] */
    float *ax;   /* [n][A][AX_N] */
    float *prev; /* [n][A][3] prev_pos of the last move_drone */
    int *ep_len, *ring_idx; /* [n][A] */
    int *tick, *task;       /* [n] */
    float *rings;           /* [n][R][7] pos3 normal3 radius */
    float *logs;            /* [n][9] */
    uint32_t key[2], env_id_base;
    uint32_t *env_episode;  /* [n] env-wide resets so far (Philox counter word) */
    uint32_t *respawns;     /* [n][A] out-of-bounds respawns so far (Philox counter word) */
    float *pay_agent;       /* [n*A][41] or NULL */
    unsigned char *flag_agent;
    float *pay_env;         /* [n][2+6R] or NULL */
    unsigned char *flag_env;
    int seed;
};

OrcSwarm *orc_swarm_create(int n, int num_agents, int max_rings) {
    OrcSwarm *o = (OrcSwarm *)calloc(1, sizeof(OrcSwarm));
    o->n = n; o->A = num_agents; o->R = max_rings;
    size_t na = (size_t)n * num_agents;
    o->st = (float *)calloc(na * 17, sizeof(float));
    o->pr = (float *)calloc(na * P_N, sizeof(float));
    o->ax = (float *)calloc(na * AX_N, sizeof(float));
    o->prev = (float *)calloc(na * 3, sizeof(float));
    o->ep_len = (int *)calloc(na, sizeof(int));
    o->ring_idx = (int *)calloc(na, sizeof(int));
    o->tick = (int *)calloc((size_t)n, sizeof(int));
    o->task = (int *)calloc((size_t)n, sizeof(int));
    o->rings = (float *)calloc((size_t)n * max_rings * 7, sizeof(float));
    o->logs = (float *)calloc((size_t)n * 9, sizeof(float));
    o->env_episode = (uint32_t *)calloc((size_t)n, sizeof(uint32_t));
    o->respawns = (uint32_t *)calloc(na, sizeof(uint32_t));
    return o;
}

void orc_swarm_close(OrcSwarm *o) {
    if (!o) return;
    free(o->st); free(o->pr); free(o->ax); free(o->prev); free(o->ep_len); free(o->ring_idx);
    free(o->tick); free(o->task); free(o->rings); free(o->logs); free(o->env_episode); free(o->respawns);
    free(o);
}

void orc_swarm_set_philox(OrcSwarm *o, uint64_t seed, uint32_t env_id_base) {
    o->key[0] = (uint32_t)seed;
    o->key[1] = (uint32_t)(seed >> 32);
    o->env_id_base = env_id_base;
}

void orc_swarm_set_payload(OrcSwarm *o, float *agent_rows, unsigned char *agent_flags, float *env_rows,
                           unsigned char *env_flags) {
    o->pay_agent = agent_rows; o->flag_agent = agent_flags; o->pay_env = env_rows; o->flag_env = env_flags;
}

/* ---- the swarm's random draws, by source ------------------------------------------------- */
/* Philox counters: (global env id, who, ordinal, item << 8 | attempt); who = agent index for the
 * env-wide reset, agent | 0x10000 for a respawn, 0xFFFF0000 for env-level draws (task, rings). */
typedef struct {
    const OrcSwarm *o;
    int mode;
    uint32_t env, who, ordinal;
} SwRand;

static void sw_words(const SwRand *r, uint32_t item, uint32_t attempt, uint32_t w[4]) {
    uint32_t ctr[4] = {r->env, r->who, r->ordinal, (item << 8) | attempt};
    orc_philox4x32_10(ctr, r->o->key, w);
}

/* size ~ U(0.1, 0.4) + init_drone's 12 jitters (R/drone_swarm/drone_swarm.h:387-388) */
static void sw_draw_params(const SwRand *r, float *p) {
    float size, uj[12];
    if (r->mode == ORC_RESET_LIBC) {
        size = libc_rndf(0.1f, 0.4); /* the reference passes the double literal 0.4 -> (float)0.4 */
        for (int k = 0; k < 12; k++) {
            float lo, hi;
            jitter_bounds(k, &lo, &hi);
            uj[k] = libc_rndf(lo, hi);
        }
        params_from_draws(size, uj, 0, p);
    } else {
        float draws[16];
        uint32_t w[4];
        for (uint32_t k = 0; k < 4; k++) {
            sw_words(r, k, 0, w);
            for (int j = 0; j < 4; j++) draws[4 * k + j] = u_from_word(w[j]);
        }
        size = lerp_u(0.1f, 0.4f, draws[0]);
        for (int k = 0; k < 12; k++) {
            float lo, hi;
            jitter_bounds(k, &lo, &hi);
            uj[k] = lerp_u(lo, hi, draws[1 + k]);
        }
        params_from_draws(size, uj, 1, p);
    }
}

static v3 sw_draw_box(const SwRand *r, uint32_t item, uint32_t attempt, float bx, float by, float bz) {
    v3 c;
    if (r->mode == ORC_RESET_LIBC) {
        c.x = libc_rndf(-bx, bx); c.y = libc_rndf(-by, by); c.z = libc_rndf(-bz, bz);
    } else {
        uint32_t w[4];
        sw_words(r, item, attempt, w);
        c.x = lerp_u(-bx, bx, u_from_word(w[0]));
        c.y = lerp_u(-by, by, u_from_word(w[1]));
        c.z = lerp_u(-bz, bz, u_from_word(w[2]));
    }
    return c;
}

/* ---- helpers mirroring the reference functions ------------------------------------------- */
#define SW_AG(o, e, a) ((size_t)(e) * (o)->A + (a))

/* R/drone_swarm/drone_swarm.h:107-129: index of the nearest other agent, -1 if alone */
static int sw_nearest(const OrcSwarm *o, int e, int a) {
    float min_dist = 999999.0f;
    int nearest = -1;
    const float *me = o->st + SW_AG(o, e, a) * 17;
    for (int j = 0; j < o->A; j++) {
        if (j == a) continue;
        const float *ot = o->st + SW_AG(o, e, j) * 17;
        float dx = me[0] - ot[0], dy = me[1] - ot[1], dz = me[2] - ot[2];
        float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        if (dist < min_dist) {
            min_dist = dist;
            nearest = j;
        }
    }
    return nearest;
}

/* R/drone_swarm/drone_swarm.h:335-376 */
static float sw_compute_reward(OrcSwarm *o, int e, int a, int collision) {
    size_t k = SW_AG(o, e, a);
    const float *s = o->st + k * 17;
    float *x = o->ax + k * AX_N;
    float dx = s[0] - x[AX_TPOS + 0], dy = s[1] - x[AX_TPOS + 1], dz = s[2] - x[AX_TPOS + 2];
    float dist = sqrtf(dx * dx + dy * dy + dz * dz);
    const float max_dist = sqrtf((2 * S_GX) * (2 * S_GX) + (2 * S_GY) * (2 * S_GY) + (2 * S_GZ) * (2 * S_GZ));
    float dist_reward = (float)(1.0 - (double)(dist / max_dist));
    float density_reward = 0.0f;
    if (collision && o->A > 1) {
        int j = sw_nearest(o, e, a);
        const float *ot = o->st + SW_AG(o, e, j) * 17;
        dx = s[0] - ot[0]; dy = s[1] - ot[1]; dz = s[2] - ot[2];
        float min_dist = sqrtf(dx * dx + dy * dy + dz * dz);
        if (min_dist < 1.0f) {
            density_reward = -1.0f;
            x[AX_COLLISIONS] += 1.0f;
        }
    }
    float abs_reward = dist_reward + density_reward;
    if (dist_reward < 0.0f && density_reward < 0.0f) abs_reward *= -1.0f;
    float delta = abs_reward - x[AX_LAST_ABS];
    x[AX_LAST_COL] = density_reward;
    x[AX_LAST_TGT] = dist_reward;
    x[AX_LAST_ABS] = abs_reward;
    o->ep_len[k] += 1;
    x[AX_SCORE] += abs_reward;
    return delta;
}

/* R/drone_swarm/drone_swarm.h:219-232 */
static void sw_move_target(float *x) {
    x[AX_TPOS + 0] += x[AX_TVEL + 0];
    x[AX_TPOS + 1] += x[AX_TVEL + 1];
    x[AX_TPOS + 2] += x[AX_TVEL + 2];
    if (x[AX_TPOS + 0] < -S_GX || x[AX_TPOS + 0] > S_GX) x[AX_TVEL + 0] = -x[AX_TVEL + 0];
    if (x[AX_TPOS + 1] < -S_GY || x[AX_TPOS + 1] > S_GY) x[AX_TVEL + 1] = -x[AX_TVEL + 1];
    if (x[AX_TPOS + 2] < -S_GZ || x[AX_TPOS + 2] > S_GZ) x[AX_TVEL + 2] = -x[AX_TVEL + 2];
}

static void sw_target_idle(OrcSwarm *o, int e, int a, const SwRand *r) {
    float *x = o->ax + SW_AG(o, e, a) * AX_N;
    v3 p = sw_draw_box(r, 5, 0, S_MX, S_MY, S_MZ);
    v3 v = sw_draw_box(r, 6, 0, S_VT, S_VT, S_VT);
    x[AX_TPOS] = p.x; x[AX_TPOS + 1] = p.y; x[AX_TPOS + 2] = p.z;
    x[AX_TVEL] = v.x; x[AX_TVEL + 1] = v.y; x[AX_TVEL + 2] = v.z;
}

/* the closed-form formation targets: R/drone_swarm/drone_swarm.h:246-261 (orbit), :273-281 (cube),
 * :299-307 (flag).  PI is raylib's float literal; sqrt/cos/sin are the double libm calls. */
void orc_swarm_formation_target(int task, int idx, int num_agents, float out[3]) {
    if (task == 2) {
        float Rr = 8.0f;
        float phi = 3.14159265358979323846f * (sqrt(5.0f) - 1.0f);
        float y = 1.0f - 2 * ((float)idx / (float)num_agents);
        float radius = sqrtf(1.0f - y * y);
        float theta = phi * idx;
        float x = cos(theta) * radius;
        float z = sin(theta) * radius;
        out[0] = Rr * x; out[1] = Rr * z; out[2] = Rr * y;
    } else if (task == 4) {
        float z = idx / 16;
        idx = idx % 16;
        float x = (float)(idx % 4);
        float y = (float)(idx / 4);
        out[0] = 4 * x - 6; out[1] = 4 * y - 6; out[2] = 4 * z - 6;
    } else {
        float x = (float)(idx % 8);
        float y = (float)(idx / 8);
        x = 2.0f * x - 7;
        y = 5 - 1.5f * y;
        out[0] = 0.0f; out[1] = x; out[2] = y;
    }
}

/* R/drone_swarm/drone_swarm.h:234-333 */
static void sw_set_target(OrcSwarm *o, int e, int a, const SwRand *r) {
    float *x = o->ax + SW_AG(o, e, a) * AX_N;
    const float *s = o->st + SW_AG(o, e, a) * 17;
    const float *x0 = o->ax + SW_AG(o, e, 0) * AX_N;
    int task = o->task[e];
    if (task == 0) {
        sw_target_idle(o, e, a, r);
        return;
    }
    if (task == 3 || task == 5) { /* follow / congo */
        if (a == 0) {
            sw_target_idle(o, e, a, r);
            return;
        }
        const float *src = task == 3 ? x0 : o->ax + SW_AG(o, e, a - 1) * AX_N;
        for (int k = 0; k < 6; k++) x[AX_TPOS + k] = src[AX_TPOS + k];
        if (task == 5)
            for (int i = 0; i < 40; i++) sw_move_target(x);
        return;
    }
    if (task == 1) {
        x[AX_TPOS] = s[0]; x[AX_TPOS + 1] = s[1]; x[AX_TPOS + 2] = s[2];
    } else if (task == S_TASK_RACE) {
        const float *g = o->rings + ((size_t)e * o->R + o->ring_idx[SW_AG(o, e, a)]) * 7;
        x[AX_TPOS] = g[0]; x[AX_TPOS + 1] = g[1]; x[AX_TPOS + 2] = g[2];
    } else {
        orc_swarm_formation_target(task, a, o->A, x + AX_TPOS);
    }
    x[AX_TVEL] = x[AX_TVEL + 1] = x[AX_TVEL + 2] = 0.0f;
}

/* R/drone_swarm/drone_swarm.h:378-399; `row` = the 16 floats (params, pos) of the payload to
 * record into / read from, or NULL */
static void sw_reset_agent(OrcSwarm *o, int e, int a, const SwRand *r, float *row) {
    size_t k = SW_AG(o, e, a);
    float *s = o->st + k * 17, *p = o->pr + k * P_N, *x = o->ax + k * AX_N;
    x[AX_RETURN] = 0.0f;
    o->ep_len[k] = 0;
    x[AX_COLLISIONS] = 0.0f;
    x[AX_SCORE] = 0.0f;
    o->ring_idx[k] = 0;
    zero_motion(s);
    if (r->mode == ORC_RESET_INJECT) {
        memcpy(p, row, P_N * sizeof(float));
        s[0] = row[13]; s[1] = row[14]; s[2] = row[15];
    } else {
        sw_draw_params(r, p);
        v3 c = sw_draw_box(r, 4, 0, S_MX, S_MY, S_MZ);
        s[0] = c.x; s[1] = c.y; s[2] = c.z;
        if (row) {
            memcpy(row, p, P_N * sizeof(float));
            row[13] = s[0]; row[14] = s[1]; row[15] = s[2];
        }
    }
    o->prev[k * 3] = s[0]; o->prev[k * 3 + 1] = s[1]; o->prev[k * 3 + 2] = s[2];
    x[AX_SPAWN] = s[0]; x[AX_SPAWN + 1] = s[1]; x[AX_SPAWN + 2] = s[2];
    sw_compute_reward(o, e, a, o->task[e] != S_TASK_RACE);
}

/* R/drone_swarm/drone_swarm.h:91-105 */
static void sw_add_log(OrcSwarm *o, int e, int a, int oob) {
    size_t k = SW_AG(o, e, a);
    float *l = o->logs + (size_t)e * 9, *x = o->ax + k * AX_N;
    l[6] += x[AX_SCORE];
    l[0] += x[AX_RETURN];
    l[1] += o->ep_len[k];
    l[3] += x[AX_COLLISIONS] / (float)o->ep_len[k];
    l[7] += x[AX_SCORE] / (float)o->ep_len[k];
    if (oob) l[4] += 1.0f;
    l[8] += 1.0f;
    o->ep_len[k] = 0;
    x[AX_RETURN] = 0.0f;
}

void orc_swarm_observe(const OrcSwarm *o, int e, float *obs);

/* R/drone_swarm/drone_swarm.h:401-443 */
static void sw_env_reset(OrcSwarm *o, int e, int mode, int first) {
    const int A = o->A, R = o->R;
    float *erow = o->pay_env ? o->pay_env + (size_t)e * (2 + 6 * R) : NULL;
    if (o->flag_env && mode != ORC_RESET_INJECT) o->flag_env[e] |= 1;
    o->tick[e] = 0;
    o->env_episode[e] = first ? 0u : o->env_episode[e] + 1u;
    SwRand er = {o, mode, o->env_id_base + (uint32_t)e, 0xFFFF0000u, o->env_episode[e]};
    if (mode == ORC_RESET_LIBC) {
        if (rand() % 4) o->task[e] = S_TASK_RACE;
        else o->task[e] = rand() % 7;
    } else if (mode == ORC_RESET_PHILOX) {
        uint32_t w[4];
        sw_words(&er, 0, 0, w);
        if ((w[0] >> 1) % 4u) o->task[e] = S_TASK_RACE;
        else o->task[e] = (int)((w[1] >> 1) % 7u);
    } else {
        o->task[e] = (int)erow[1];
    }
    if (erow && mode != ORC_RESET_INJECT) erow[1] = (float)o->task[e];

    for (int a = 0; a < A; a++) {
        float *arow = o->pay_agent ? o->pay_agent + SW_AG(o, e, a) * ORC_SWARM_AGENT_PAYLOAD : NULL;
        SwRand ar = {o, mode, er.env, (uint32_t)a, er.ordinal};
        sw_reset_agent(o, e, a, &ar, arow ? arow + 16 : NULL);
        float *x = o->ax + SW_AG(o, e, a) * AX_N;
        if (mode == ORC_RESET_INJECT) {
            for (int k = 0; k < 6; k++) x[AX_TPOS + k] = arow[32 + k];
        } else {
            sw_set_target(o, e, a, &ar);
            if (arow) for (int k = 0; k < 6; k++) arow[32 + k] = x[AX_TPOS + k];
        }
    }
    float *rings = o->rings + (size_t)e * R * 7;
    memset(rings, 0, (size_t)R * 7 * sizeof(float));
    if (o->task[e] == S_TASK_RACE) {
        const float lox = -S_GX + 2 * RING_RADIUS, loy = -S_GY + 2 * RING_RADIUS, loz = -S_GZ + 2 * RING_RADIUS;
        const float min_gap = 2.0f * RING_RADIUS;
        for (int r = 0; r < R; r++) {
            float *g = rings + 7 * r;
            if (mode == ORC_RESET_INJECT) {
                memcpy(g, erow + 2 + 6 * r, 6 * sizeof(float));
            } else {
                for (uint32_t t = 0;; t++) {
                    v3 c, nrm;
                    if (mode == ORC_RESET_LIBC) {
                        c.x = libc_rndf(lox, -lox); c.y = libc_rndf(loy, -loy); c.z = libc_rndf(loz, -loz);
                        float u1 = libc_rndf(0.0f, 1.0f), u2 = libc_rndf(0.0f, 1.0f), u3 = libc_rndf(0.0f, 1.0f);
                        nrm = ring_normal_from_u(u1, u2, u3, 0);
                    } else {
                        uint32_t w[4], w2[4];
                        sw_words(&er, 0x10u + 2u * r, t, w);
                        sw_words(&er, 0x11u + 2u * r, t, w2);
                        c.x = lerp_u(lox, -lox, u_from_word(w[0]));
                        c.y = lerp_u(loy, -loy, u_from_word(w[1]));
                        c.z = lerp_u(loz, -loz, u_from_word(w[2]));
                        nrm = ring_normal_from_u(u_from_word(w[3]), u_from_word(w2[0]), u_from_word(w2[1]), 1);
                    }
                    g[0] = c.x; g[1] = c.y; g[2] = c.z; g[3] = nrm.x; g[4] = nrm.y; g[5] = nrm.z;
                    if (r == 0) break;
                    v3 prev = {g[-7], g[-6], g[-5]};
                    if (!(dist3(c, prev) < min_gap)) break;
                    if (mode == ORC_RESET_PHILOX && t + 1 >= MAX_ATTEMPTS) break;
                }
                if (erow) memcpy(erow + 2 + 6 * r, g, 6 * sizeof(float));
            }
            g[6] = RING_RADIUS;
        }
        v3 r0 = {rings[0], rings[1], rings[2]};
        for (int a = 0; a < A; a++) {
            float *s = o->st + SW_AG(o, e, a) * 17;
            float *arow = o->pay_agent ? o->pay_agent + SW_AG(o, e, a) * ORC_SWARM_AGENT_PAYLOAD : NULL;
            if (mode == ORC_RESET_INJECT) {
                s[0] = arow[38]; s[1] = arow[39]; s[2] = arow[40];
                continue;
            }
            SwRand ar = {o, mode, er.env, (uint32_t)a, er.ordinal};
            for (uint32_t t = 0;; t++) {
                v3 c = sw_draw_box(&ar, 7, t, S_MX, S_MY, S_MZ);
                s[0] = c.x; s[1] = c.y; s[2] = c.z;
                if (!(dist3(c, r0) < min_gap)) break;
                if (mode == ORC_RESET_PHILOX && t + 1 >= MAX_ATTEMPTS) break;
            }
            if (arow) { arow[38] = s[0]; arow[39] = s[1]; arow[40] = s[2]; }
        }
    } else if (erow && mode != ORC_RESET_INJECT) {
        memset(erow + 2, 0, (size_t)6 * R * sizeof(float));
    }
}

/* R/drone_swarm/drone_swarm.h:131-217 for every agent of env e; obs = [A][41] */
void orc_swarm_observe(const OrcSwarm *o, int e, float *obs) {
    for (int a = 0; a < o->A; a++) {
        size_t k = SW_AG(o, e, a);
        const float *s = o->st + k * 17, *p = o->pr + k * P_N, *x = o->ax + k * AX_N;
        float *ob = obs + (size_t)a * ORC_SWARM_OBS;
        q4 q = {s[6], s[7], s[8], s[9]};
        q4 qi = {q.w, -q.x, -q.y, -q.z};
        v3 vel = {s[3], s[4], s[5]};
        v3 vb = qrot(qi, vel);
        v3 zax = {0.0f, 0.0f, 1.0f};
        v3 up = qrot(q, zax);
        int c = 0;
        ob[c++] = vb.x / K_MAX_VEL; ob[c++] = vb.y / K_MAX_VEL; ob[c++] = vb.z / K_MAX_VEL;
        ob[c++] = s[10] / K_MAX_OMEGA; ob[c++] = s[11] / K_MAX_OMEGA; ob[c++] = s[12] / K_MAX_OMEGA;
        ob[c++] = up.x; ob[c++] = up.y; ob[c++] = up.z;
        ob[c++] = q.w; ob[c++] = q.x; ob[c++] = q.y; ob[c++] = q.z;
        for (int m = 0; m < 4; m++) ob[c++] = s[13 + m] / p[P_MRPM];
        ob[c++] = s[0] / S_GX; ob[c++] = s[1] / S_GY; ob[c++] = s[2] / S_GZ;
        ob[c++] = x[AX_SPAWN] / S_GX; ob[c++] = x[AX_SPAWN + 1] / S_GY; ob[c++] = x[AX_SPAWN + 2] / S_GZ;
        float dx = x[AX_TPOS] - s[0], dy = x[AX_TPOS + 1] - s[1], dz = x[AX_TPOS + 2] - s[2];
        ob[c++] = clampf_(dx, -1.0f, 1.0f); ob[c++] = clampf_(dy, -1.0f, 1.0f); ob[c++] = clampf_(dz, -1.0f, 1.0f);
        ob[c++] = dx / S_GX; ob[c++] = dy / S_GY; ob[c++] = dz / S_GZ;
        ob[c++] = x[AX_LAST_COL]; ob[c++] = x[AX_LAST_TGT]; ob[c++] = x[AX_LAST_ABS];
        if (o->A > 1) {
            const float *ot = o->st + SW_AG(o, e, sw_nearest(o, e, a)) * 17;
            ob[c++] = clampf_(ot[0] - s[0], -1.0f, 1.0f);
            ob[c++] = clampf_(ot[1] - s[1], -1.0f, 1.0f);
            ob[c++] = clampf_(ot[2] - s[2], -1.0f, 1.0f);
        } else {
            ob[c++] = 0.0f; ob[c++] = 0.0f; ob[c++] = 0.0f;
        }
        if (o->task[e] == S_TASK_RACE) {
            const float *g = o->rings + ((size_t)e * o->R + o->ring_idx[k]) * 7;
            v3 d = {g[0] - s[0], g[1] - s[1], g[2] - s[2]};
            v3 nrm = {g[3], g[4], g[5]};
            v3 to = qrot(qi, d), bn = qrot(qi, nrm);
            ob[c++] = to.x / S_GX; ob[c++] = to.y / S_GY; ob[c++] = to.z / S_GZ;
            ob[c++] = bn.x; ob[c++] = bn.y; ob[c++] = bn.z;
        } else {
            for (int m = 0; m < 6; m++) ob[c++] = 0.0f;
        }
    }
}

void orc_swarm_reset(OrcSwarm *o, int mode, int seed, float *obs) {
    o->seed = seed;
    for (int e = 0; e < o->n; e++) {
        if (mode == ORC_RESET_LIBC) srand(e + seed * o->n); /* EB:500-504 */
        for (int a = 0; a < o->A; a++) o->respawns[SW_AG(o, e, a)] = 0;
        sw_env_reset(o, e, mode, 1);
        if (obs) orc_swarm_observe(o, e, obs + SW_AG(o, e, 0) * ORC_SWARM_OBS);
    }
}

/* R/drone_swarm/drone_swarm.h:445-497 for env e */
static void sw_step_env(OrcSwarm *o, int e, int mode, float *actions, float *obs, float *rew, unsigned char *term) {
    const int A = o->A;
    o->tick[e] = (o->tick[e] + 1) % S_HORIZON;
    for (int a = 0; a < A; a++) {
        size_t k = SW_AG(o, e, a);
        float *s = o->st + k * 17, *x = o->ax + k * AX_N;
        rew[k] = 0;
        term[k] = 0;
        if (mode == ORC_RESET_LIBC) (void)rand(); /* dt-jitter draw, R/drone_swarm/dronelib.h:440 */
        o->prev[k * 3] = s[0]; o->prev[k * 3 + 1] = s[1]; o->prev[k * 3 + 2] = s[2];
        advance_body(s, o->pr + k * P_N, actions + k * 4);
        int oob = s[0] < -S_GX || s[0] > S_GX || s[1] < -S_GY || s[1] > S_GY || s[2] < -S_GZ || s[2] > S_GZ;
        sw_move_target(x);
        float reward;
        if (o->task[e] == S_TASK_RACE) {
            const float *g = o->rings + ((size_t)e * o->R + o->ring_idx[k]) * 7;
            reward = sw_compute_reward(o, e, a, 1);
            v3 before = {o->prev[k * 3], o->prev[k * 3 + 1], o->prev[k * 3 + 2]};
            v3 after = {s[0], s[1], s[2]};
            float passed = gate_event(before, after, g, -0.0f);
            if (passed > 0) {
                o->ring_idx[k] = (o->ring_idx[k] + 1) % o->R;
                o->logs[(size_t)e * 9 + 2] += 1.0f;
                SwRand none = {o, mode, 0, 0, 0};
                sw_set_target(o, e, a, &none);
                sw_compute_reward(o, e, a, 1);
            }
            reward += passed;
        } else {
            reward = sw_compute_reward(o, e, a, 1);
        }
        rew[k] += reward;
        x[AX_RETURN] += reward;
        if (oob) {
            rew[k] -= 1;
            term[k] = 1;
            sw_add_log(o, e, a, 1);
            float *arow = o->pay_agent ? o->pay_agent + k * ORC_SWARM_AGENT_PAYLOAD : NULL;
            if (o->flag_agent && mode != ORC_RESET_INJECT) o->flag_agent[k] |= 1;
            o->respawns[k] += 1;
            SwRand ar = {o, mode, o->env_id_base + (uint32_t)e, (uint32_t)a | 0x10000u, o->respawns[k]};
            sw_reset_agent(o, e, a, &ar, arow);
        } else if (o->tick[e] >= S_HORIZON - 1) {
            term[k] = 1;
            sw_add_log(o, e, a, 0);
        }
    }
    if (o->tick[e] >= S_HORIZON - 1) sw_env_reset(o, e, mode, 0);
    orc_swarm_observe(o, e, obs + SW_AG(o, e, 0) * ORC_SWARM_OBS);
}

void orc_swarm_step(OrcSwarm *o, int mode, float *actions, float *obs, float *rew, unsigned char *term) {
    if (mode != ORC_RESET_INJECT) {
        if (o->flag_agent) memset(o->flag_agent, 0, (size_t)o->n * o->A);
        if (o->flag_env) memset(o->flag_env, 0, (size_t)o->n);
    }
    for (int e = 0; e < o->n; e++) sw_step_env(o, e, mode, actions, obs, rew, term);
}

void orc_swarm_log(OrcSwarm *o, float out[9]) {
    for (int j = 0; j < 9; j++) out[j] = 0.0f;
    for (int e = 0; e < o->n; e++) {
        float *l = o->logs + (size_t)e * 9;
        for (int j = 0; j < 9; j++) {
            out[j] += l[j];
            l[j] = 0.0f;
        }
    }
}

/* blobs: same layouts as oracle/ref_shim_swarm.c */
void orc_swarm_get_env(const OrcSwarm *o, int e, float *b) {
    b[0] = (float)o->tick[e];
    b[1] = (float)o->task[e];
    for (int r = 0; r < o->R; r++) memcpy(b + 2 + 6 * r, o->rings + ((size_t)e * o->R + r) * 7, 6 * sizeof(float));
}

void orc_swarm_put_env(OrcSwarm *o, int e, const float *b) {
    o->tick[e] = (int)b[0];
    o->task[e] = (int)b[1];
    for (int r = 0; r < o->R; r++) {
        float *g = o->rings + ((size_t)e * o->R + r) * 7;
        memcpy(g, b + 2 + 6 * r, 6 * sizeof(float));
        g[6] = (g[3] == 0.0f && g[4] == 0.0f && g[5] == 0.0f) ? 0.0f : RING_RADIUS;
    }
}

void orc_swarm_get_agent(const OrcSwarm *o, int e, int a, float *b) {
    size_t k = SW_AG(o, e, a);
    const float *x = o->ax + k * AX_N;
    memcpy(b, o->st + k * 17, 17 * sizeof(float));
    memcpy(b + 17, o->pr + k * P_N, P_N * sizeof(float));
    memcpy(b + 30, x + AX_SPAWN, 3 * sizeof(float));
    memcpy(b + 33, x + AX_TPOS, 6 * sizeof(float));
    b[39] = x[AX_LAST_ABS]; b[40] = x[AX_LAST_TGT]; b[41] = x[AX_LAST_COL];
    b[42] = x[AX_RETURN]; b[43] = x[AX_COLLISIONS]; b[44] = (float)o->ep_len[k];
    b[45] = x[AX_SCORE]; b[46] = (float)o->ring_idx[k];
}

void orc_swarm_put_agent(OrcSwarm *o, int e, int a, const float *b) {
    size_t k = SW_AG(o, e, a);
    float *x = o->ax + k * AX_N;
    memcpy(o->st + k * 17, b, 17 * sizeof(float));
    memcpy(o->pr + k * P_N, b + 17, P_N * sizeof(float));
    memcpy(x + AX_SPAWN, b + 30, 3 * sizeof(float));
    memcpy(x + AX_TPOS, b + 33, 6 * sizeof(float));
    x[AX_LAST_ABS] = b[39]; x[AX_LAST_TGT] = b[40]; x[AX_LAST_COL] = b[41];
    x[AX_RETURN] = b[42]; x[AX_COLLISIONS] = b[43]; o->ep_len[k] = (int)b[44];
    x[AX_SCORE] = b[45]; o->ring_idx[k] = (int)b[46];
    memcpy(o->prev + k * 3, b, 3 * sizeof(float));
}

/* ====================================================================== advantage
 * pufferlib/extensions/pufferlib.cpp:28-41 (puff_advantage_row) and :63-72 (puff_advantage),
 * restated with the same strides interface as b2d_puff_advantage.  Built with -ffp-contract=off
 * like the reference's CPU build (no FMA on baseline x86-64). */
void orc_puff_advantage(const float *values, const float *rewards, const float *dones, const float *importance,
                        float *advantages, float *abs_sum, int num_rows, int horizon, long long row_stride,
                        long long t_stride, float gamma, float lambda, float rho_clip, float c_clip) {
    for (int row = 0; row < num_rows; row++) {
        long long base = (long long)row * row_stride;
        float lastpufferlam = 0;
        float prio = 0.0f;
        for (int t = horizon - 2; t >= 0; t--) {
            long long at = base + (long long)t * t_stride, an = at + t_stride;
            float nextnonterminal = 1.0 - dones[an];
            float rho_t = fminf(importance[at], rho_clip);
            float c_t = fminf(importance[at], c_clip);
            float delta = rho_t * (rewards[an] + gamma * values[an] * nextnonterminal - values[at]);
            lastpufferlam = delta + gamma * lambda * c_t * lastpufferlam * nextnonterminal;
            advantages[at] = lastpufferlam;
            prio += fabsf(lastpufferlam);
        }
        if (abs_sum) abs_sum[row] = prio;
    }
}
