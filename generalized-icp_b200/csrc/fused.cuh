// Fused registration loop for small pairs (the reference's own workloads: visualization.py's ~90-point pair,
// robot-visualization.py's 90-360-beam scans): the whole outer loop of gicp() (gicp.py:116-167) in ONE launch.
// One block per pair runs  K3a search -> K3b accumulation -> K4 solve  until the pair's stop rule fires; the
// stages are the very block functions of the multi-launch path (objective.cuh, solve.cuh), separated by
// __syncthreads() instead of kernel boundaries, so there is no launch, no poll and no host round trip inside a
// registration.  Requires every source cloud to fit one block (<= OBJ_THREADS * OBJ_MAX_PPT points).
#pragma once
#include "objective.cuh"
#include "solve.cuh"

namespace gicp {

template <int D, typename Real>
__global__ void __launch_bounds__(OBJ_THREADS) register_loop_kernel(const ObjArgs<Real> oa, const SolveArgs sa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NRED = Dim<D>::NRED;
    __shared__ double s_sum[NRED];
    __shared__ SolveScratch<D> s_sc;
    const int pair = blockIdx.x;
    WarpStage<Real> ws = obj_warp_stage<Real>(smem_raw);   // one mbarrier per warp for the whole loop
    for (int it = 0; it < sa.max_iterations; ++it) {
        correspond_block<D, Real>(oa, pair, 0, ws, smem_raw);
        __syncthreads();                                   // this block's matches are visible to the block
        accumulate_block<D, Real>(oa, pair, 0);            // -> partial[pair][0][NRED]
        __syncthreads();
        if (threadIdx.x < NRED) s_sum[threadIdx.x] = oa.partial[(size_t)pair * NRED + threadIdx.x];
        __syncthreads();
        if (threadIdx.x < 32) solve_pair<D>(sa, pair, sa.state[pair], s_sum, s_sc);   // warp 0, all lanes
        __syncthreads();                                   // the new state (global) is visible to the block
        if (sa.state[pair].status != PAIR_ACTIVE) break;
    }
}

}  // namespace gicp
