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

// Sequential exclusive running sum of x over one ray, 32 elements at a time: lane k holds x[base+k]; the
// chain s += x_k runs identically in every lane (values broadcast by shuffle), lane k keeps s before x_k.
__device__ __forceinline__ float chunk_exclusive_scan(float x, float& running, int lane) {
    float mine = 0.f;
#pragma unroll
    for (int k = 0; k < kWarp; ++k) {
        const float v = __shfl_sync(kFull, x, k);
        if (lane == k) mine = running;
        running = __fadd_rn(running, v);
    }
    return mine;
}


// exclusive suffix sum inside a warp: lane i gets sum_{j > i} v_j (exactly 0 for lane 31)
__device__ __forceinline__ float warp_suffix_exclusive(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float dn = __shfl_down_sync(kFull, v, o);
        if (lane + o < 32) v += dn;
    }
    const float next = __shfl_down_sync(kFull, v, 1);
    return lane == 31 ? 0.f : next;
}

// Compositing backward of ONE ray by one warp (autograd of nerf_helpers.py:58-104).  Given g = dL/d(ray colour) it hands
// `put(i, dsigma_pre_i, drgb_pre_i0, _i1, _i2)` the gradients w.r.t. the head PRE-activations of every sample i (called by the
// lane that owns the sample):
//   d sigma_pre_i = [sigma_i > 0] * delta_i * ( T_{i+1} (c_i . g) - sum_{j>i} w_j (c_j . g) )
//   d rgb_pre_i,k = w_i g_k * c_i,k (1 - c_i,k)                      (sigmoid derivative)
// with T_{i+1} = T_i exp(-sigma_i delta_i) (derivation: SURVEY.md 7.2; checked against torch autograd in tests).  S <= 1024.
template <class Put>
__device__ __forceinline__ void composite_backward_ray(const float* __restrict__ sigma, const float* __restrict__ rgb,
                                                       const float* __restrict__ ts, const float* __restrict__ g_ray, int64_t n, int S,
                                                       int lane, Put&& put) {
    const int chunks = (S + kWarp - 1) / kWarp;          // <= 32: lane c keeps chunk c's starting prefix
    const float g0 = g_ray[n * 3], g1 = g_ray[n * 3 + 1], g2 = g_ray[n * 3 + 2];
    const float* sg = sigma + n * S;
    const float* tp = ts + n * S;
    // pass 0 (forward): running sum of -sigma*delta at every chunk start, with the fused forward's own 5-step shuffle scan (the
    // transmittances are then the ones the forward composited with; the sequential 32-step form cost 6x the instructions)
    float chunk_start = 0.f, running = 0.f;
    for (int c = 0; c < chunks; ++c) {
        if (lane == c) chunk_start = running;
        const int i = c * kWarp + lane;
        const bool in = i < S;
        const float s = in ? sg[i] : 0.f;
        const float t = in ? tp[i] : 0.f;
        float tn = __shfl_down_sync(kFull, t, 1);
        if (lane == 31 && i + 1 < S) tn = tp[i + 1];
        const float dl = (i == S - 1) ? 1e10f : __fsub_rn(tn, t);
        const float x = in ? __fmul_rn(__fmul_rn(-1.0f, s), dl) : 0.f;
        (void)chunk_exclusive_scan_tree(x, running, lane);
    }
    // pass 1 (backward over chunks): suffix sums of q_j = w_j (c_j . g) taken directly, so that the last sample,
    // whose interval is 1e10 wide, sees an exactly-zero suffix (total - prefix would leave a rounding residue)
    float later = 0.f;                                   // sum of q over all later chunks
    for (int c = chunks - 1; c >= 0; --c) {
        float run = __shfl_sync(kFull, chunk_start, c);
        const int i = c * kWarp + lane;
        const bool in = i < S;
        const float s = in ? sg[i] : 0.f;
        const float t = in ? tp[i] : 0.f;
        float tn = __shfl_down_sync(kFull, t, 1);
        if (lane == 31 && i + 1 < S) tn = tp[i + 1];
        const float dl = (i == S - 1) ? 1e10f : __fsub_rn(tn, t);
        const float x = in ? __fmul_rn(__fmul_rn(-1.0f, s), dl) : 0.f;
        const float excl = chunk_exclusive_scan_tree(x, run, lane);
        const float trans = expf(excl), e = expf(x);
        const float w = in ? __fmul_rn(__fsub_rn(1.0f, e), trans) : 0.f;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        if (in) { const float* cc = rgb + (n * S + i) * 3; c0 = cc[0]; c1 = cc[1]; c2 = cc[2]; }
        const float cg = c0 * g0 + c1 * g1 + c2 * g2;
        const float q = w * cg;
        const float suffix = later + warp_suffix_exclusive(q, lane);
        if (in) {
            const float ds = dl * (trans * e * cg - suffix);
            put(i, (s > 0.f) ? ds : 0.f, w * g0 * c0 * (1.f - c0), w * g1 * c1 * (1.f - c1), w * g2 * c2 * (1.f - c2));
        }
        later += warp_sum(q);
    }
}

}  // namespace nerf
