// common.cuh - shared host/device helpers for libnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nerf_b200.h"

namespace nerf {

// thread-local error text returned by nerf_last_error()
void set_error(const char* fmt, ...);
int check_launch(const char* what);     // cudaGetLastError() -> 0 / NERF_E_CUDA
int num_sms();                           // SM count of the current device (cached)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: callers keep one bit per device ordinal in a
// thread-local mask and (re)apply their attributes the first time a thread launches on a device.
bool attrs_pending(unsigned long long mask);      // true when the current device's bit is not yet set
void attrs_done(unsigned long long& mask);        // set it
template <class Kernel> inline cudaError_t allow_smem(Kernel k, size_t bytes) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

#define NERF_REQUIRE(cond, ...)                     \
    do {                                            \
        if (!(cond)) {                              \
            nerf::set_error(__VA_ARGS__);           \
            return NERF_E_ARG;                      \
        }                                           \
    } while (0)

// 16-byte / 4-byte stores of the training tensors (activations, sign words, dz): written once, read by a later kernel
#ifndef NERF_STREAMING_STORES
#define NERF_STREAMING_STORES 1
#endif
// (NERF_STREAMING_STORES: 0 plain st.global, 1 st.global.cs, 2 st.global.wt, 3 st.global.cg - A/B with tools/ab_build_flag.sh)
template <class T> __device__ __forceinline__ void store_once(T* p, T v) {
#if NERF_STREAMING_STORES == 1
    __stcs(p, v);
#elif NERF_STREAMING_STORES == 2
    __stwt(p, v);
#elif NERF_STREAMING_STORES == 3
    __stcg(p, v);
#else
    *p = v;
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

}  // namespace nerf
