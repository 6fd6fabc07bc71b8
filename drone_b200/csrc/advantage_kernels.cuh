// advantage_kernels.cuh -- PufferLib's advantage estimator (GAE with V-trace clipping) for sm_100a.
//
// Reference: pufferlib/extensions/cuda/pufferlib.cu:7-51 (puff_advantage_row_cuda / puff_advantage_kernel,
// one thread per [segments, horizon] row, strided row walks, -O3 only) and its CPU twin
// pufferlib/extensions/pufferlib.cpp:28-41,63-72.  For every row, backwards in time:
//     nnt   = 1 - dones[t+1]
//     delta = min(imp[t], rho_clip) * (rewards[t+1] + gamma * values[t+1] * nnt - values[t])
//     adv   = delta + gamma * lambda * min(imp[t], c_clip) * adv * nnt          -> advantages[t]
// advantages[horizon-1] is left untouched, as in the reference.
//
// Layout is given by two strides, so the same kernel serves the reference's row-major
// [segments, horizon] tensors (row_stride = horizon, t_stride = 1) and the time-major
// [horizon, num_agents] experience the on-device rollout writes (row_stride = 1, t_stride = num_agents),
// where the 32 lanes of a warp read 32 consecutive floats at every time step (fully coalesced:
// 20 B per element, HBM-bound).  Optionally fuses the priority the trainer computes right after,
// sum_t |adv| per row (pufferl.py:342).  STRICT = one IEEE op per reference op (bit-exact with the
// CPU twin); otherwise FMA contraction.
#pragma once
#include "b2d_math.cuh"

namespace b2d {

template <bool STRICT>
__global__ void __launch_bounds__(256) puff_advantage_kernel(const float *__restrict__ values, const float *__restrict__ rewards,
                                                             const float *__restrict__ dones, const float *__restrict__ importance,
                                                             float *__restrict__ advantages, float *__restrict__ abs_sum,
                                                             int num_rows, int horizon, long long row_stride, long long t_stride,
                                                             float gamma, float lambda, float rho_clip, float c_clip) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    const long long base = (long long)row * row_stride;
    float last = 0.0f, prio = 0.0f;
    float v_next = horizon > 0 ? values[base + (long long)(horizon - 1) * t_stride] : 0.0f;
#pragma unroll 4
    for (int t = horizon - 2; t >= 0; t--) {
        const long long at = base + (long long)t * t_stride, an = at + t_stride;
        const float v = values[at], r = rewards[an], dn = dones[an], imp = importance[at];
        if constexpr (STRICT) {
            // float nextnonterminal = 1.0 - dones[t_next]: a double subtraction rounded once; exact in float
            const xf nnt = xf(1.0f) - xf(dn);
            const xf rho = xf(fminf(imp, rho_clip)), c = xf(fminf(imp, c_clip));
            const xf delta = rho * (xf(r) + xf(gamma) * xf(v_next) * nnt - xf(v));
            last = (delta + xf(gamma) * xf(lambda) * c * xf(last) * nnt).v;
        } else {
            const float nnt = 1.0f - dn;
            const float rho = fminf(imp, rho_clip), c = fminf(imp, c_clip);
            const float delta = rho * (fmaf(gamma * v_next, nnt, r) - v);
            last = fmaf(gamma * lambda * c * nnt, last, delta);
        }
        advantages[at] = last;
        prio += fabsf(last);
        v_next = v;
    }
    if (abs_sum) abs_sum[row] = prio;
}

} // namespace b2d
