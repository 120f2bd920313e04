// composite_scan.cuh - the warp-level running sum shared by the stand-alone compositing kernel (samplers.cu) and the
// compositing warp fused into the tensor-core MLP kernel (mlp_tc3.cu): both spell the same arithmetic in the same order,
// so their weights / ray colours are bit-identical given the same sigma / rgb / depths.
#pragma once
#include "common.cuh"

namespace nerf {

// Exclusive running sum of x over 32 lanes as a 5-step shuffle scan (used by the fused composite, whose weights are
// compared with the reference to 1e-6: the summation order differs from the CPU's by ~1 ulp of the partial sums, which is
// far below what expf already contributes).  The sequential form above stays where the ORDER matters bit for bit: the
// fine sampler's cdf (bin indices) and the stand-alone nerf_weights.
__device__ __forceinline__ float chunk_exclusive_scan_tree(float x, float& running, int lane) {
    float incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
    }
    // exclusive = the previous lane's inclusive sum (NOT incl - x: the last sample's interval is 1e10 wide and would cancel
    // every smaller term)
    const float prev = __shfl_up_sync(kFull, incl, 1);
    const float excl = running + (lane ? prev : 0.f);
    running += __shfl_sync(kFull, incl, 31);
    return excl;
}

}  // namespace nerf
