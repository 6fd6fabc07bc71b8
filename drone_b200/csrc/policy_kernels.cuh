// policy_kernels.cuh -- the rollout's per-step policy work as ONE kernel for sm_100a.
//
// Reference: what PuffeRL.evaluate does between vecenv.recv() and vecenv.send() for a Box action
// space (pufferlib/pufferl.py:229-296): `policy.forward_eval` of pufferlib.models.Default
// (models.py:41-98: Linear(obs, hidden) + exact GELU, mean head, state-independent log-std, value
// head), `sample_logits` on Normal(mean, exp(logstd)) (pytorch.py:189-199: sample, summed log-prob),
// reward clamp to [-1, 1] (pufferl.py:260), the experience stores (pufferl.py:270-281) and the clip of
// the sampled action to the action space before it goes to the env (pufferl.py:292-294).  In torch
// that is ~20 launches moving [rows, hidden] activations through HBM four times; here the hidden
// layer never leaves registers: 285 B/row of HBM traffic (obs 116 in, obs copy 116 + action 16 +
// env action 16 + 4 scalars 16 + reward/terminal 5 in) against ~4.3 k FP32 issue slots per row, so
// the kernel is FP32-issue bound, not HBM bound.
//
// Mapping: one lane owns TWO rows (lane and lane+32 of a 64-row warp tile), so every weight fetched
// from shared memory (LDS.128, warp-uniform address = broadcast) feeds 4 packed FFMA2 (sm_100
// fma.rn.f32x2 with a scalar-broadcast operand): 1 LDS per 8 FMAs.  Observation rows arrive as
// lane-consecutive float4 (512 B per instruction), are copied to the experience row from the same
// registers and transposed through a per-warp shared-memory tile (row stride 29 or 41 words: odd,
// conflict-free).  GELU uses erf(t) = 1 - 2^(-t*R(t)) with a degree-7 minimax R on [0, 4]
// (|error| < 1e-7 absolute, the fp32 rounding floor; see tests/test_policy_gpu.py), evaluated two
// hidden units at a time on packed FMAs with one MUFU.EX2 each.  Noise is Philox4x32-10 keyed by
// (seed; global row, call number) + Box-Muller, so a rollout is reproducible and independent of
// the GPU count; the call number is a device-resident counter so CUDA-graph replays draw fresh noise.
#pragma once
#include "b2d_math.cuh"
#include "race_kernels.cuh"

namespace b2d {

constexpr int POL_THREADS = 128;
constexpr int POL_WARPS = POL_THREADS / 32;
constexpr int POL_TILE = 64; // rows per warp tile: two per lane

struct PolicyArgs {
    // weights (device, f32, torch nn.Linear layouts)
    const float *enc_w;   // [hidden, obs_dim]
    const float *enc_b;   // [hidden]
    const float *mean_w;  // [4, hidden]
    const float *mean_b;  // [4]
    const float *logstd;  // [4]
    const float *value_w; // [hidden]
    const float *value_b; // [1]
    int hidden;           // multiple of 8
    // env contract buffers (device)
    const float *obs;            // [rows, D]
    const float *rew;            // [rows]
    const unsigned char *term;   // [rows]
    float *env_act;              // [rows, 4]  clip(action, -1, 1)
    // experience row (each may be null)
    float *st_obs, *st_act, *st_logp, *st_rew, *st_term, *st_val;
    int rows;
    uint32_t row_id_base;
    uint32_t seed_lo, seed_hi;
    unsigned int *counter; // [2]: calls completed, CTAs arrived in the running call
    int deterministic;
};

__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// GELU(v) = 0.5 v (1 + erf(v / sqrt 2)) for two values, as relu(v) - 0.5 u erfc(u / sqrt 2) with u = |v|: the same
// function, no cancellation in the negative tail.  erfc(u / sqrt 2) = 2^(-u P(u)) with a degree-5 P fitted on
// [0, 4 sqrt 2] for the absolute error of GELU itself (8e-8 in exact arithmetic, 2.5e-7 max(1, |v|) in float32
// with ex2.approx: the float32 rounding floor, tests/test_policy_gpu.py); u is clamped to 4 sqrt 2, where the
// second term is below 1e-7.  Per pair: 8 instructions on the FMA pipe (5 FFMA2 Horner + 2 FMUL2 + 1 FFMA2), 6 on
// the ALU pipe (|.|, min, max), 2 MUFU.EX2.
__device__ __forceinline__ float2 gelu2(float2 v) {
    const float2 u = make_float2(fminf(fabsf(v.x), 5.656854249492381f), fminf(fabsf(v.y), 5.656854249492381f));
    float2 r = bc2(-2.992167357093618e-05f);
    r = __ffma2_rn(r, u, bc2(0.0007398558564848026f));
    r = __ffma2_rn(r, u, bc2(-0.0079774147335873f));
    r = __ffma2_rn(r, u, bc2(0.0532381323988627f));
    r = __ffma2_rn(r, u, bc2(0.45891571090498645f));
    r = __ffma2_rn(r, u, bc2(1.151147079583815f));
    const float2 q = __fmul2_rn(r, u);
    const float2 e = make_float2(ex2_approx(-q.x), ex2_approx(-q.y));
    const float2 hu = __fmul2_rn(u, bc2(-0.5f));
    return __ffma2_rn(hu, e, make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f)));
}

// The same with a degree-4 exponent polynomial: |error| <= 5.5e-7 in exact arithmetic (minimax fit over the same
// range), one packed FMA less per pair.  For consumers that truncate the activation to TF32 anyway (2^-11 relative:
// the second GEMM of race_rollout_kernel), where the degree-5 form's last 4e-7 cannot be seen.
__device__ __forceinline__ float2 gelu2_tf32(float2 v) {
    const float2 u = make_float2(fminf(fabsf(v.x), 5.656854249492381f), fminf(fabsf(v.y), 5.656854249492381f));
    float2 r = bc2(4.86890071e-04f);
    r = __ffma2_rn(r, u, bc2(-7.19119915e-03f));
    r = __ffma2_rn(r, u, bc2(5.21308913e-02f));
    r = __ffma2_rn(r, u, bc2(4.59608779e-01f));
    r = __ffma2_rn(r, u, bc2(1.15099711e+00f));
    const float2 q = __fmul2_rn(r, u);
    const float2 e = make_float2(ex2_approx(-q.x), ex2_approx(-q.y));
    const float2 hu = __fmul2_rn(u, bc2(-0.5f));
    return __ffma2_rn(hu, e, make_float2(fmaxf(v.x, 0.0f), fmaxf(v.y, 0.0f)));
}

#ifndef B2D_POLICY_FAST_SINCOS
#define B2D_POLICY_FAST_SINCOS 1
#endif
// four standard normals for (row, call): Philox4x32-10 + Box-Muller
__device__ __forceinline__ void policy_noise(uint32_t row, uint32_t call, uint32_t k0, uint32_t k1, float z[4]) {
    const uint4 w = philox4x32_10(make_uint4(row, call, 0x504f4c49u, 0u), k0, k1);
    const float S = 1.1920928955078125e-07f; // 2^-23
    const float u0 = ((float)(w.x >> 9) + 0.5f) * S, u1 = ((float)(w.y >> 9) + 0.5f) * S;
    const float u2 = ((float)(w.z >> 9) + 0.5f) * S, u3 = ((float)(w.w >> 9) + 0.5f) * S;
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
#if B2D_POLICY_FAST_SINCOS
    // sin / cos of the uniform angle on the special-function unit: the angle 2 pi u is taken as theta + pi with
    // theta = 2 pi (u - 0.5) in (-pi, pi), where sin.approx / cos.approx are good to 2^-21.4 absolute
    // (|z| error < 3e-6, inside the 2e-6 max(1, |z|) of tests/test_policy_gpu.py); sin(theta + pi) = -sin theta.
    const float t0 = 6.283185307179586f * (u1 - 0.5f), t1 = 6.283185307179586f * (u3 - 0.5f);
    const float nr0 = -r0, nr1 = -r1;
    z[0] = nr0 * __cosf(t0);
    z[1] = nr0 * __sinf(t0);
    z[2] = nr1 * __cosf(t1);
    z[3] = nr1 * __sinf(t1);
#else
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    z[0] = r0 * c0;
    z[1] = r0 * s0;
    z[2] = r1 * c1;
    z[3] = r1 * s1;
#endif
}

// sample, log-prob, experience stores and the clipped env action of one row (pufferl.py:258-294)
__device__ __forceinline__ void policy_row_epilogue(const PolicyArgs &a, int row, const float (&mean)[4], float value, unsigned int call,
                                                    const float (&sd)[4], float lp0) {
    float act[4], lp = lp0;
    if (a.deterministic) {
#pragma unroll
        for (int c = 0; c < 4; c++) act[c] = mean[c];
    } else {
        float z[4];
        policy_noise(a.row_id_base + (uint32_t)row, call, a.seed_lo, a.seed_hi, z);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            act[c] = fmaf(sd[c], z[c], mean[c]);
            lp = fmaf(-0.5f * z[c], z[c], lp);
        }
    }
    if (a.st_act) __stcs(reinterpret_cast<float4 *>(a.st_act) + row, make_float4(act[0], act[1], act[2], act[3]));
    if (a.st_logp) __stcs(&a.st_logp[row], lp);
    if (a.st_val) __stcs(&a.st_val[row], value);
    if (a.st_rew) __stcs(&a.st_rew[row], fminf(fmaxf(__ldcs(&a.rew[row]), -1.0f), 1.0f));
    if (a.st_term) __stcs(&a.st_term[row], (float)__ldcs(&a.term[row]));
    reinterpret_cast<float4 *>(a.env_act)[row] = make_float4(fminf(fmaxf(act[0], -1.0f), 1.0f), fminf(fmaxf(act[1], -1.0f), 1.0f),
                                                             fminf(fmaxf(act[2], -1.0f), 1.0f), fminf(fmaxf(act[3], -1.0f), 1.0f));
}

// the last CTA to finish publishes the call number of the next launch
__device__ __forceinline__ void policy_call_done(const PolicyArgs &a) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int arrived = atomicAdd(&a.counter[1], 1u);
        if (arrived == gridDim.x - 1) {
            a.counter[1] = 0;
            __threadfence();
            atomicAdd(&a.counter[0], 1u);
        }
    }
}

template <int D>
__global__ void __launch_bounds__(POL_THREADS) policy_act_kernel(const PolicyArgs a) {
    extern __shared__ float4 pol_smem4[];
    const int H = a.hidden;
    float *w1t = reinterpret_cast<float *>(pol_smem4); // [D][H], hidden contiguous
    float *b1 = w1t + D * H;                             // [H]
    float4 *w2m = reinterpret_cast<float4 *>(b1 + H);    // [H] mean weights of hidden j for the 4 actions
    float *w2v = reinterpret_cast<float *>(w2m + H);     // [H]
    float *tiles = w2v + H;                              // [POL_WARPS][POL_TILE * D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < D * H; i += POL_THREADS) {
        const int k = i / H, j = i - k * H;
        w1t[i] = __ldg(&a.enc_w[j * D + k]);
    }
    for (int j = threadIdx.x; j < H; j += POL_THREADS) {
        b1[j] = __ldg(&a.enc_b[j]);
        w2m[j] = make_float4(__ldg(&a.mean_w[j]), __ldg(&a.mean_w[H + j]), __ldg(&a.mean_w[2 * H + j]), __ldg(&a.mean_w[3 * H + j]));
        w2v[j] = __ldg(&a.value_w[j]);
    }
    const unsigned int call = *reinterpret_cast<volatile unsigned int *>(a.counter);
    float bm[4], ls[4], sd[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        bm[c] = __ldg(&a.mean_b[c]);
        ls[c] = __ldg(&a.logstd[c]);
        sd[c] = expf(ls[c]);
    }
    const float bv = __ldg(&a.value_b[0]);
    const float lp0 = -(ls[0] + ls[1] + ls[2] + ls[3]) - 4.0f * 0.91893853320467274f;
    __syncthreads();

    float *tile = tiles + warp * (POL_TILE * D);
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    const int ntiles = (a.rows + POL_TILE - 1) / POL_TILE;
    const bool ld_vec = (reinterpret_cast<uintptr_t>(a.obs) & 15) == 0;
    const bool st_vec = a.st_obs && (reinterpret_cast<uintptr_t>(a.st_obs) & 15) == 0;

    for (int t = blockIdx.x * POL_WARPS + warp; t < ntiles; t += gridDim.x * POL_WARPS) {
        const int r0 = t * POL_TILE;
        const int nrows = min(POL_TILE, a.rows - r0);
        const int nf = nrows * D;
        const float *g = a.obs + (size_t)r0 * D; // 16-byte aligned when a.obs is: 64 * D * 4 is a multiple of 16
        float *so = a.st_obs ? a.st_obs + (size_t)r0 * D : nullptr;
        __syncwarp();
        if (ld_vec) {
            const float4 *g4 = reinterpret_cast<const float4 *>(g);
            const int n4 = nf >> 2;
#pragma unroll 4
            for (int i = lane; i < n4; i += 32) {
                const float4 v = __ldcs(&g4[i]);
                tile4[i] = v;
                if (st_vec) __stcs(reinterpret_cast<float4 *>(so) + i, v);
            }
            for (int i = (n4 << 2) + lane; i < nf; i += 32) tile[i] = __ldcs(&g[i]);
            if (so) {
                if (!st_vec) {
                    __syncwarp();
                    for (int i = lane; i < nf; i += 32) __stcs(&so[i], tile[i]);
                } else {
                    for (int i = (n4 << 2) + lane; i < nf; i += 32) __stcs(&so[i], tile[i]);
                }
            }
        } else {
            for (int i = lane; i < nf; i += 32) {
                const float v = __ldcs(&g[i]);
                tile[i] = v;
                if (so) __stcs(&so[i], v);
            }
        }
        __syncwarp();

        float oA[D], oB[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            oA[k] = tile[lane * D + k];
            oB[k] = tile[(lane + 32) * D + k];
        }

        float2 mA01 = make_float2(bm[0], bm[1]), mA23 = make_float2(bm[2], bm[3]), vA = make_float2(bv, 0.0f);
        float2 mB01 = mA01, mB23 = mA23, vB = vA;
#pragma unroll 1
        for (int jb = 0; jb < H; jb += 8) {
            const float4 bq0 = *reinterpret_cast<const float4 *>(b1 + jb), bq1 = *reinterpret_cast<const float4 *>(b1 + jb + 4);
            float2 hA[4] = {make_float2(bq0.x, bq0.y), make_float2(bq0.z, bq0.w), make_float2(bq1.x, bq1.y), make_float2(bq1.z, bq1.w)};
            float2 hB[4] = {hA[0], hA[1], hA[2], hA[3]};
#pragma unroll
            for (int k = 0; k < D; k++) {
                const float4 w0 = *reinterpret_cast<const float4 *>(w1t + k * H + jb);
                const float4 w1 = *reinterpret_cast<const float4 *>(w1t + k * H + jb + 4);
                const float2 p0 = make_float2(w0.x, w0.y), p1 = make_float2(w0.z, w0.w), p2 = make_float2(w1.x, w1.y), p3 = make_float2(w1.z, w1.w);
                const float2 xa = bc2(oA[k]), xb = bc2(oB[k]);
                hA[0] = __ffma2_rn(xa, p0, hA[0]);
                hB[0] = __ffma2_rn(xb, p0, hB[0]);
                hA[1] = __ffma2_rn(xa, p1, hA[1]);
                hB[1] = __ffma2_rn(xb, p1, hB[1]);
                hA[2] = __ffma2_rn(xa, p2, hA[2]);
                hB[2] = __ffma2_rn(xb, p2, hB[2]);
                hA[3] = __ffma2_rn(xa, p3, hA[3]);
                hB[3] = __ffma2_rn(xb, p3, hB[3]);
            }
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const float2 gA = gelu2(hA[p]), gB = gelu2(hB[p]);
                const float4 m0 = w2m[jb + 2 * p], m1 = w2m[jb + 2 * p + 1];
                const float2 wv = *reinterpret_cast<const float2 *>(w2v + jb + 2 * p);
                mA01 = __ffma2_rn(bc2(gA.x), make_float2(m0.x, m0.y), mA01);
                mA23 = __ffma2_rn(bc2(gA.x), make_float2(m0.z, m0.w), mA23);
                mA01 = __ffma2_rn(bc2(gA.y), make_float2(m1.x, m1.y), mA01);
                mA23 = __ffma2_rn(bc2(gA.y), make_float2(m1.z, m1.w), mA23);
                vA = __ffma2_rn(gA, wv, vA);
                mB01 = __ffma2_rn(bc2(gB.x), make_float2(m0.x, m0.y), mB01);
                mB23 = __ffma2_rn(bc2(gB.x), make_float2(m0.z, m0.w), mB23);
                mB01 = __ffma2_rn(bc2(gB.y), make_float2(m1.x, m1.y), mB01);
                mB23 = __ffma2_rn(bc2(gB.y), make_float2(m1.z, m1.w), mB23);
                vB = __ffma2_rn(gB, wv, vB);
            }
        }

#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int row = r0 + lane + 32 * half;
            if (row < a.rows) {
                const float2 m01 = half ? mB01 : mA01, m23 = half ? mB23 : mA23, vv = half ? vB : vA;
                const float mean[4] = {m01.x, m01.y, m23.x, m23.y};
                policy_row_epilogue(a, row, mean, vv.x + vv.y, call, sd, lp0);
            }
        }
    }

    policy_call_done(a);
}

// ---------------------------------------------------------------------------------------------
// TF32 tensor-core form.  The reference runs its policy under torch.set_float32_matmul_precision('high')
// (pufferl.py:55), i.e. both Linear layers are TF32 GEMMs with float32 accumulation on the GPU; this
// kernel does the same with mma.sync.m16n8k8.tf32 (operands rounded to TF32 with round-to-nearest,
// ties away -- cvt.rna), which takes the two GEMMs (3712 + 640 FMA per row) off the FP32 pipe.  What is
// left there is GELU (~12 packed FMAs per hidden pair), so the kernel moves from FP32-pipe bound to
// HBM / FP32 co-bound.
//
// Per warp tile of 64 rows: the A fragments of MT 16-row m-tiles (obs, K padded to a multiple of 8 with
// zeros) stay in registers; for each group of 8 hidden units j: KS B-fragment loads (LDS.64, prepared
// per lane at kernel start) feed MT*KS MMAs, GELU runs on the accumulator fragment in place, and the
// fragment is fed straight back as the A operand of the second GEMM (hidden -> 4 means + value, N
// padded to 8): the C layout (row g / g+8, cols 2t, 2t+1) is the A layout (cols t, t+4) up to a
// permutation of the 8 hidden units inside the group, which is folded into the prepared B fragments
// of the second layer.  The next tile's observations arrive by cp.async while this one computes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// same rounding on the integer pipe: add half an ulp of the 10-bit mantissa, the tensor core drops the low 13 bits
__device__ __forceinline__ uint32_t to_tf32_alu(float x) { return __float_as_uint(x) + 0x1000u; }

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// cp_async16 / cp_async_commit / cp_async_wait<N> come from race_kernels.cuh (included first by api.cu)
__device__ __forceinline__ void cp_async_commit_wait_all() {
    cp_async_commit();
    cp_async_wait<0>();
}

template <int D, int WARPS, int MT>
__global__ void __launch_bounds__(WARPS * 32, 512 / (WARPS * 32)) policy_act_tf32_kernel(const PolicyArgs a) {
    constexpr int KS = (D + 7) / 8;          // k-steps of the first GEMM
    constexpr int PASSES = (POL_TILE / 16) / MT;
    extern __shared__ float4 pol_smem4[];
    const int H = a.hidden, NT = H >> 3;
    uint2 *bw1 = reinterpret_cast<uint2 *>(pol_smem4); // [NT][KS][32] B fragments of encoder.weight^T (tf32)
    uint2 *bw2 = bw1 + NT * KS * 32;                   // [NT][32]     B fragments of [mean | value | 0]^T (tf32)
    float *b1 = reinterpret_cast<float *>(bw2 + NT * 32); // [H]
    float *tiles = b1 + H;                                // [WARPS][POL_TILE * D]
    float *outs = tiles + WARPS * POL_TILE * D;           // [WARPS][POL_TILE * 8]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

    for (int i = threadIdx.x; i < NT * KS * 32; i += WARPS * 32) {
        const int ln = i & 31, s = (i >> 5) % KS, j = (i >> 5) / KS;
        const int n = 8 * j + (ln >> 2), k0 = 8 * s + (ln & 3), k1 = k0 + 4;
        bw1[i] = make_uint2(to_tf32(k0 < D ? __ldg(&a.enc_w[n * D + k0]) : 0.0f), to_tf32(k1 < D ? __ldg(&a.enc_w[n * D + k1]) : 0.0f));
    }
    for (int i = threadIdx.x; i < NT * 32; i += WARPS * 32) {
        const int ln = i & 31, j = i >> 5, n = ln >> 2, h0 = 8 * j + 2 * (ln & 3), h1 = h0 + 1;
        const float w0 = n < 4 ? __ldg(&a.mean_w[n * H + h0]) : (n == 4 ? __ldg(&a.value_w[h0]) : 0.0f);
        const float w1 = n < 4 ? __ldg(&a.mean_w[n * H + h1]) : (n == 4 ? __ldg(&a.value_w[h1]) : 0.0f);
        bw2[i] = make_uint2(to_tf32(w0), to_tf32(w1));
    }
    for (int j = threadIdx.x; j < H; j += WARPS * 32) b1[j] = __ldg(&a.enc_b[j]);
    const unsigned int call = *reinterpret_cast<volatile unsigned int *>(a.counter);
    float ls[4], sd[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        ls[c] = __ldg(&a.logstd[c]);
        sd[c] = expf(ls[c]);
    }
    const float lp0 = -(ls[0] + ls[1] + ls[2] + ls[3]) - 4.0f * 0.91893853320467274f;
    // second-layer bias of this lane's two output columns (2t, 2t+1): means 0..3, value, padding
    const int n0 = 2 * t;
    const float bz0 = n0 < 4 ? __ldg(&a.mean_b[n0]) : (n0 == 4 ? __ldg(&a.value_b[0]) : 0.0f);
    const float bz1 = n0 + 1 < 4 ? __ldg(&a.mean_b[n0 + 1]) : 0.0f;
    __syncthreads();

    float *tile = tiles + warp * (POL_TILE * D);
    float4 *tile4 = reinterpret_cast<float4 *>(tile);
    float *out = outs + warp * (POL_TILE * 8);
    const int ntiles = (a.rows + POL_TILE - 1) / POL_TILE;
    const bool ld_vec = (reinterpret_cast<uintptr_t>(a.obs) & 15) == 0;
    const bool st_vec = a.st_obs && (reinterpret_cast<uintptr_t>(a.st_obs) & 15) == 0;
    constexpr int N4 = POL_TILE * D / 4; // float4 per full tile
    bool prefetched = false;

    for (int tl = blockIdx.x * WARPS + warp; tl < ntiles; tl += gridDim.x * WARPS) {
        const int r0 = tl * POL_TILE;
        const int nrows = min(POL_TILE, a.rows - r0);
        const int nf = nrows * D;
        const float *gsrc = a.obs + (size_t)r0 * D;
        float *so = a.st_obs ? a.st_obs + (size_t)r0 * D : nullptr;
        if (prefetched) {
            cp_async_commit_wait_all();
        } else if (ld_vec) {
            const float4 *g4 = reinterpret_cast<const float4 *>(gsrc);
            const int n4 = nf >> 2;
#pragma unroll 4
            for (int i = lane; i < n4; i += 32) tile4[i] = __ldcs(&g4[i]);
            for (int i = (n4 << 2) + lane; i < nf; i += 32) tile[i] = __ldcs(&gsrc[i]);
        } else {
            for (int i = lane; i < nf; i += 32) tile[i] = __ldcs(&gsrc[i]);
        }
        __syncwarp();
        if (so) { // experience copy of the observation rows, from shared memory
            if (st_vec) {
                const int n4 = nf >> 2;
#pragma unroll 4
                for (int i = lane; i < n4; i += 32) __stcs(reinterpret_cast<float4 *>(so) + i, tile4[i]);
                for (int i = (n4 << 2) + lane; i < nf; i += 32) __stcs(&so[i], tile[i]);
            } else {
                for (int i = lane; i < nf; i += 32) __stcs(&so[i], tile[i]);
            }
        }

#pragma unroll 1
        for (int pass = 0; pass < PASSES; pass++) {
            const int mb = pass * MT * 16;
            uint32_t A[MT][KS][4];
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const float *r_lo = tile + (mb + 16 * m + g) * D, *r_hi = r_lo + 8 * D;
#pragma unroll
                for (int s = 0; s < KS; s++) {
                    const int k0 = 8 * s + t, k1 = k0 + 4; // k0 < D always for s < KS - 1; the last step may be padding
                    A[m][s][0] = to_tf32((8 * s + 3 < D || k0 < D) ? r_lo[k0] : 0.0f);
                    A[m][s][1] = to_tf32((8 * s + 3 < D || k0 < D) ? r_hi[k0] : 0.0f);
                    A[m][s][2] = to_tf32((8 * s + 7 < D || k1 < D) ? r_lo[k1] : 0.0f);
                    A[m][s][3] = to_tf32((8 * s + 7 < D || k1 < D) ? r_hi[k1] : 0.0f);
                }
            }
            if (pass == PASSES - 1) {
                // the tile is in registers: let the next one stream into the same shared memory
                __syncwarp();
                const int nx = tl + gridDim.x * WARPS;
                prefetched = ld_vec && nx < ntiles && a.rows - nx * POL_TILE >= POL_TILE;
                if (prefetched) {
                    const float4 *g4 = reinterpret_cast<const float4 *>(a.obs + (size_t)nx * POL_TILE * D);
#pragma unroll 4
                    for (int i = lane; i < N4; i += 32) cp_async16(&tile4[i], &g4[i]);
                }
            }
            float c2[MT][4];
#pragma unroll
            for (int m = 0; m < MT; m++) {
                c2[m][0] = bz0;
                c2[m][1] = bz1;
                c2[m][2] = bz0;
                c2[m][3] = bz1;
            }
#pragma unroll 1
            for (int j = 0; j < NT; j++) {
                const float2 bb = *reinterpret_cast<const float2 *>(b1 + 8 * j + 2 * t);
                float acc[MT][4];
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    acc[m][0] = bb.x;
                    acc[m][1] = bb.y;
                    acc[m][2] = bb.x;
                    acc[m][3] = bb.y;
                }
                const uint2 *bj = bw1 + (j * KS) * 32 + lane;
#pragma unroll
                for (int s = 0; s < KS; s++) {
                    const uint2 b = bj[s * 32];
#pragma unroll
                    for (int m = 0; m < MT; m++) mma_tf32(acc[m], A[m][s], b.x, b.y);
                }
                const uint2 b2 = bw2[j * 32 + lane];
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const float2 lo = gelu2(make_float2(acc[m][0], acc[m][1])); // row g,     hidden 8j + 2t, 8j + 2t + 1
                    const float2 hi = gelu2(make_float2(acc[m][2], acc[m][3])); // row g + 8
                    const uint32_t a2[4] = {to_tf32_alu(lo.x), to_tf32_alu(hi.x), to_tf32_alu(lo.y), to_tf32_alu(hi.y)};
                    mma_tf32(c2[m], a2, b2.x, b2.y);
                }
            }
#pragma unroll
            for (int m = 0; m < MT; m++) {
                *reinterpret_cast<float2 *>(out + (mb + 16 * m + g) * 8 + 2 * t) = make_float2(c2[m][0], c2[m][1]);
                *reinterpret_cast<float2 *>(out + (mb + 16 * m + g + 8) * 8 + 2 * t) = make_float2(c2[m][2], c2[m][3]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int lr = lane + 32 * half, row = r0 + lr;
            const float4 mq = *reinterpret_cast<const float4 *>(out + lr * 8);
            const float value = out[lr * 8 + 4];
            if (row < a.rows) {
                const float mean[4] = {mq.x, mq.y, mq.z, mq.w};
                policy_row_epilogue(a, row, mean, value, call, sd, lp0);
            }
        }
        __syncwarp(); // `out` is rewritten by the next tile
    }
    if (prefetched) cp_async_commit_wait_all(); // never true here (a prefetch implies another iteration); keeps the group accounting obvious
    policy_call_done(a);
}

} // namespace b2d
