// swarm_kernels.cuh -- device side of the multi-drone swarm env (placeholder until the
// warp-per-env kernel lands; b2d_swarm_create reports B2D_ESTATE meanwhile).
#pragma once
#include "race_kernels.cuh"

namespace b2d {

struct SwarmDev {
    int n, num_agents, max_rings;
    Ctl *ctl;
    const float *payload;
    int reset_mode;
};

__global__ void swarm_log_snapshot_kernel(Ctl *, long long *) {}
__global__ void swarm_pack_kernel(const SwarmDev, const int *, int, float *) {}
__global__ void swarm_unpack_kernel(const SwarmDev, const int *, int, const float *) {}
static inline void swarm_vec_reset(SwarmDev &, uint64_t, cudaStream_t, long long *) {}
static inline void swarm_vec_step(SwarmDev &, const float *, int, cudaStream_t, long long *) {}
static inline void swarm_observe_launch(SwarmDev &, cudaStream_t) {}
static inline void swarm_log_finish(const long long *, float *) {}

} // namespace b2d
