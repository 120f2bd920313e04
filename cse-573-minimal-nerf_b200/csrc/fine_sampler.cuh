// fine_sampler.cuh - K3 + K4 for ONE ray by one warp: inverse-CDF fine depths (nerf_helpers.py:106-156) merged with the coarse
// depths into one sorted row (nerf_model.py:116-120); the body of fine_sample_merge_kernel (samplers.cu).  NETWORK = false is the
// form that was tried INSIDE the coarse network's kernel (its compositing warps calling it for every ray they had just
// composited: bit-identical, but the producer warps became the coarse kernel's bottleneck - 32.7 ms against 27.6 + 1.6 ms per
// 640 000 rays, profiles/r02_notes.md - so the sampler stays its own launch).
#pragma once
#include "composite_scan.cuh"

namespace nerf {

__device__ __forceinline__ void cmpx(float& a, float& b, bool up) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = up ? lo : hi;
    b = up ? hi : lo;
}

// one compare stage of a bitonic network over nreg x 32 values held as v[j] of lane l = element 32 j + l: distances >= 32 are
// register-to-register, smaller ones one __shfl_xor each
template <int NREG>
__device__ __forceinline__ void bitonic_stage(float (&v)[NREG], int k, int dist, int lane) {
    if (dist >= 32) {
        const int dj = dist >> 5;
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            if ((j & dj) == 0 && (j | dj) < NREG) {
                const bool up = (((32 * j + lane) & k) == 0);
                cmpx(v[j], v[j | dj], up);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NREG; ++j) {
            const int el = 32 * j + lane;
            const float other = __shfl_xor_sync(kFull, v[j], dist);
            const bool up = ((el & k) == 0);
            const bool lower = ((lane & dist) == 0);
            v[j] = (up == lower) ? fminf(v[j], other) : fmaxf(v[j], other);
        }
    }
}

// per-warp scratch: cdf [C] | bounds [C + 2] | sorted fine depths [128] | merged row [256] (the last only when the row is staged
// for a coalesced write-back, NETWORK = true; the fused kernel has ~4 KB of shared memory left and scatters the row to global)
__host__ __device__ constexpr int fine_sampler_floats(int C, bool staged = true) { return 2 * C + 2 + 128 + (staged ? 256 : 0); }

// On entry scratch[0 .. C) holds the ray's raw coarse weights and scratch[C + 1 .. 2C + 1) its coarse depths (both written by
// this warp, __syncwarp'ed).  Writes the C + F sorted depths to `out` (global).  u_ray: the ray's F jitter uniforms; eps_n its
// query offset; q_base [F] the reference's torch.arange(0, 1, 1/F).
//
// The sort uses what is known about the two halves instead of a 256-wide network over everything: the C coarse depths
// normally arrive SORTED (stratified: t_i lies in stratum i), so only the F <= 128 fine depths go through a bitonic network
// (128 wide: 28 compare stages over 4 registers instead of 36 over 8), and the two sorted runs are merged by RANK - a fine
// depth's slot is its index plus the number of coarse depths below it, a coarse depth's slot its index plus the number of
// fine depths not above it (binary searches in shared memory; ties go fine-first, so the slots are a permutation).  A sorted
// array is unique, so the result is bit-identical to sorting all C + F values.  Rows whose coarse depths are NOT sorted (a
// caller's own depths), a NaN depth, or F > 128 / C > 128 take the general path: NETWORK = true - the 256-wide bitonic network over
// the concatenation (C + F <= 256); NETWORK = false (inside the fused kernel, where registers are scarce and the case cannot
// occur with its own stratified depths) - an enumeration rank in shared memory.
template <bool NETWORK>
__device__ __forceinline__ void fine_sample_merge_ray(float* scratch, int C, int F, float near_, float far_, float eps_n,
                                                      const float* __restrict__ u_ray, const float* __restrict__ q_base,
                                                      float* __restrict__ out, int lane) {
    float* cdf = scratch;
    float* bounds = scratch + C;
    float* fsorted = bounds + C + 2;
    float* merged = NETWORK ? fsorted + 128 : out;            // staged in shared memory, or straight into the output row
    const int S = C + F;
    const float Ff = (float)F;
    float running = 0.f;                                                          // nerf_helpers.py:137, sequential order
    for (int base = 0; base < C; base += kWarp) {
        const int i = base + lane;
        const float x = (i < C) ? cdf[i] : 0.f;
        const float excl = chunk_exclusive_scan(x, running, lane);
        if (i < C) cdf[i] = __fadd_rn(excl, x);
    }
    if (lane == 0) { bounds[0] = near_; bounds[C + 1] = far_; }
    __syncwarp();
    const float total = cdf[C - 1];
    __syncwarp();
    for (int i = lane; i < C; i += kWarp) cdf[i] = __fdiv_rn(cdf[i], total);       // nerf_helpers.py:138
    __syncwarp();
    const float e = __fdiv_rn(eps_n, Ff);                                           // nerf_helpers.py:139
    constexpr int NV = NETWORK ? 8 : 4;
    float v[NV];
    bool plain = true;                                                               // no NaN among this lane's values
    bool fast_shape = F <= 128 && C <= 128;
#pragma unroll
    for (int jj = 0; jj < NV; ++jj) {
        const int j = 32 * jj + lane;                                                 // element index in cat([fine, coarse])
        float t = __int_as_float(0x7f800000);
        if (j < F) {
            const float q = __fadd_rn(__ldg(q_base + j), e);                         // nerf_helpers.py:142
            int lo = 0, hi = C;                                                       // torch.searchsorted, right=False
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if (!(cdf[mid] >= q)) lo = mid + 1; else hi = mid;
            }
            const float b0 = bounds[lo], b1 = bounds[lo + 1];
            t = __fadd_rn(b0, __fmul_rn(__fsub_rn(b1, b0), u_ray[j]));               // nerf_helpers.py:154
            plain = plain && (t == t);
        }
        v[jj] = t;
    }
    bool run_sorted = fast_shape;
    for (int i = lane; i + 1 < C; i += kWarp) run_sorted = run_sorted && (bounds[i + 2] >= bounds[i + 1]);   // false on NaN too
    if (__all_sync(kFull, run_sorted && plain)) {
        // ---- 128-wide network over the fine depths (registers 0..3; entries >= F are +inf), then the rank merge
        float f4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
        for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
            for (int dist = k >> 1; dist >= 1; dist >>= 1) bitonic_stage<4>(f4, k, dist, lane);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) fsorted[32 * jj + lane] = f4[jj];
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int el = 32 * jj + lane;
            if (el < F) {
                const float t = f4[jj];
                int lo = 0, hi = C;                                                   // coarse depths below t
                while (lo < hi) {
                    const int mid = lo + ((hi - lo) >> 1);
                    if (bounds[mid + 1] < t) lo = mid + 1; else hi = mid;
                }
                merged[el + lo] = t;
            }
        }
        for (int i = lane; i < C; i += kWarp) {
            const float tc = bounds[i + 1];
            int lo = 0, hi = F;                                                       // fine depths not above tc
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if (fsorted[mid] <= tc) lo = mid + 1; else hi = mid;
            }
            merged[i + lo] = tc;
        }
        __syncwarp();
        if (NETWORK) {
            for (int el = lane; el < S; el += kWarp) out[el] = merged[el];
            __syncwarp();
        }
        return;
    }
    if constexpr (NETWORK) {
        // ---- general path: the coarse depths join the fine ones (nerf_model.py:117: fine first, then coarse), 256-wide network
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = 32 * jj + lane;
            if (j >= F && j < S) v[jj] = bounds[j - F + 1];
        }
#pragma unroll
        for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
            for (int dist = k >> 1; dist >= 1; dist >>= 1) bitonic_stage<8>(v, k, dist, lane);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int el = 32 * j + lane;
            if (el < S) out[el] = v[j];
        }
    } else {
        // ---- general path without registers to spare (F <= 128, C <= 128 guaranteed by the caller): enumeration rank of the
        // concatenation [fine | coarse] in shared memory; ties by index
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) fsorted[32 * jj + lane] = v[jj];
        __syncwarp();
        for (int a = lane; a < S; a += kWarp) {
            const float ta = a < F ? fsorted[a] : bounds[a - F + 1];
            int rank = 0;
            for (int b = 0; b < S; ++b) {
                const float tb = b < F ? fsorted[b] : bounds[b - F + 1];
                rank += (tb < ta || (tb == ta && b < a)) ? 1 : 0;
            }
            merged[rank] = ta;                       // (= out: scattered)
        }
        __syncwarp();
    }
}

}  // namespace nerf
