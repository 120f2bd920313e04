// pack.cu - K5: fp32 nn.Linear parameters (state_dict order) -> the bf16 swizzled stage image + fp32 biases
// that mlp_tc3.cu streams with bulk copies (layout in pack_layout.cuh).
#include "common.cuh"
#include "pack_layout.cuh"
#include "umma.cuh"

namespace nerf {

__constant__ pk::Layout c_pack_layout = pk::kLayout;
__constant__ pk::LayoutT c_pack_layout_t = pk::kLayoutT;

struct PackParams { const float* p[20]; };

__device__ __forceinline__ void pack_forward_block(const PackParams& P, uint8_t* __restrict__ packed, int s) {
    if (s < pk::kStages) {
        const pk::Stage st = c_pack_layout.st[s];
        const float* W = P.p[2 * st.param];
        uint8_t* tile = packed + st.offset;
        for (int e = threadIdx.x; e < st.rows * 64; e += blockDim.x) {
            const int r = e >> 6, k = e & 63;
            float v = 0.f;
            if (r < st.valid_rows && k < st.kvalid) v = W[(size_t)(st.n0 + r) * st.in_features + st.k0 + k];
            *(__nv_bfloat16*)(tile + umma::sw128_offset(r, k)) = __float2bfloat16_rn(v);
        }
    } else {
        float* b = (float*)(packed + c_pack_layout.bias_offset);
        for (int i = threadIdx.x; i < pk::kBiasFloats; i += blockDim.x) {
            float v = 0.f;
            if (i < 1792) v = P.p[2 * (i >> 8) + 1][i & 255];                       // mlp.0 .. feature_fn.4
            else if (i < 1920) v = P.p[17][i - 1792];                               // rgb_fn.0
            else if (i == pk::kBiasSigma) v = P.p[15][0];                           // density_fn.0
            else if (i >= pk::kBiasRgb && i < pk::kBiasRgb + 3) v = P.p[19][i - pk::kBiasRgb];   // rgb_fn.2
            b[i] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(PackParams P, uint8_t* __restrict__ packed) { pack_forward_block(P, packed, blockIdx.x); }

__device__ __forceinline__ void pack_transposed_block(const PackParams& P, uint8_t* __restrict__ packed, int s) {
    if (s < pk::kStagesT) {
        const pk::Stage st = c_pack_layout_t.st[s];
        const float* W = P.p[2 * st.param];
        uint8_t* tile = packed + st.offset;
        for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {
            const int k = e >> 7, r = e & 127;          // consecutive threads walk r: coalesced reads of W rows
            const float v = W[(size_t)(st.k0 + k) * st.in_features + st.n0 + r];
            *(__nv_bfloat16*)(tile + umma::sw128_offset(r, k)) = __float2bfloat16_rn(v);
        }
    } else {
        float* c = (float*)(packed + c_pack_layout_t.const_offset);
        for (int i = threadIdx.x; i < pk::kConstFloatsT; i += blockDim.x)
            c[i] = (i < 384) ? P.p[18][i] : P.p[14][i - 384];      // rgb_fn.2.weight [3,128], density_fn.0.weight [1,256]
    }
}

__global__ void __launch_bounds__(256)
pack_weights_t_kernel(PackParams P, uint8_t* __restrict__ packed) { pack_transposed_block(P, packed, blockIdx.x); }

// Both images of both networks in one launch (after an optimiser step): blocks [0, 64) forward image of network 0,
// [64, 117) its W^T image, then the same for network 1.
struct PackAllParams { PackParams net[2]; uint8_t* fwd[2]; uint8_t* tr[2]; };
constexpr int kPackBlocksPerNet = (pk::kStages + 1) + (pk::kStagesT + 1);

__global__ void __launch_bounds__(256)
pack_weights_all_kernel(PackAllParams A) {
    const int net = blockIdx.x / kPackBlocksPerNet, b = blockIdx.x % kPackBlocksPerNet;
    if (b <= pk::kStages) pack_forward_block(A.net[net], A.fwd[net], b);
    else pack_transposed_block(A.net[net], A.tr[net], b - (pk::kStages + 1));
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_pack_weights_all(const float* const* params40_host, void* packed0, void* packed_t0, void* packed1,
                                     void* packed_t1, void* stream) {
    NERF_REQUIRE(params40_host && packed0 && packed_t0 && packed1 && packed_t1, "nerf_pack_weights_all: null pointer");
    NERF_REQUIRE(((((uintptr_t)packed0) | ((uintptr_t)packed_t0) | ((uintptr_t)packed1) | ((uintptr_t)packed_t1)) & 127) == 0,
                 "nerf_pack_weights_all: buffers must be 128-byte aligned");
    PackAllParams A;
    for (int i = 0; i < 40; ++i) {
        NERF_REQUIRE(params40_host[i], "nerf_pack_weights_all: params40_host[%d] is NULL", i);
        A.net[i / 20].p[i % 20] = params40_host[i];
    }
    A.fwd[0] = (uint8_t*)packed0; A.tr[0] = (uint8_t*)packed_t0; A.fwd[1] = (uint8_t*)packed1; A.tr[1] = (uint8_t*)packed_t1;
    pack_weights_all_kernel<<<2 * kPackBlocksPerNet, 256, 0, (cudaStream_t)stream>>>(A);
    return check_launch("nerf_pack_weights_all");
}

extern "C" size_t nerf_packed_t_bytes(void) { return pk::kLayoutT.total_bytes; }

extern "C" int nerf_pack_weights_t(const float* const* params20_host, void* packed_t, void* stream) {
    NERF_REQUIRE(params20_host && packed_t, "nerf_pack_weights_t: null pointer");
    NERF_REQUIRE(((uintptr_t)packed_t & 127) == 0, "nerf_pack_weights_t: buffer must be 128-byte aligned");
    PackParams P;
    for (int i = 0; i < 20; ++i) {
        NERF_REQUIRE(params20_host[i], "nerf_pack_weights_t: params20_host[%d] is NULL", i);
        P.p[i] = params20_host[i];
    }
    pack_weights_t_kernel<<<pk::kStagesT + 1, 256, 0, (cudaStream_t)stream>>>(P, (uint8_t*)packed_t);
    return check_launch("nerf_pack_weights_t");
}

extern "C" size_t nerf_packed_bytes(void) { return pk::kLayout.total_bytes; }

extern "C" int nerf_pack_weights(const float* const* params20_host, void* packed, void* stream) {
    NERF_REQUIRE(params20_host && packed, "nerf_pack_weights: null pointer");
    NERF_REQUIRE(((uintptr_t)packed & 127) == 0, "nerf_pack_weights: packed buffer must be 128-byte aligned");
    PackParams P;
    for (int i = 0; i < 20; ++i) {
        NERF_REQUIRE(params20_host[i], "nerf_pack_weights: params20_host[%d] is NULL", i);
        P.p[i] = params20_host[i];
    }
    pack_weights_kernel<<<pk::kStages + 1, 256, 0, (cudaStream_t)stream>>>(P, (uint8_t*)packed);
    return check_launch("nerf_pack_weights");
}
