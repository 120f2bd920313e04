// mlp_tc_common.cuh - device helpers shared by the tensor-core kernels (forward mlp_tc3.cu, dgrad mlp_tc_bwd3.cu, wgrad
// wgrad_tc.cu): weight-stage references, the range-reduced MUFU sin/cos, and the positional-encoding row encoder.
#pragma once
#include "common.cuh"
#include "pack_layout.cuh"
#include "umma.cuh"

namespace nerf {

namespace tcm {
constexpr float kPiF = 3.14159274101257324f;
constexpr float kInv2Pi = 0.15915494309189535f;
constexpr float k2PiHi = 6.28318548202514648f;
constexpr float k2PiLo = -1.7484555e-7f;
}  // namespace tcm

// per-ray outputs of the compositing fused into the two-tile kernel (mlp_tc3.cu); weights / depth / acc / stats may be null
struct CompositeOutputs {
    float* weights; float* ray_rgb; float* depth; float* acc; float* stats;
    const float* u_c; const float* t_base; float step; float* ts_gen;      // non-null u_c: stratified depths formed in the kernel
};

struct StageRef { uint32_t offset, bytes; };
// offset of density_fn.0's four [16 x 64] bf16 blocks in the packed image (row 0 of each = w7, the only row the kernels read:
// sigma is taken on the CUDA cores): the stages are laid out in consumption order, density_fn.0 sits between rgb_fn.0 and rgb_fn.2
constexpr uint32_t density_stage_offset() {
    for (int i = 0; i < pk::kStages; ++i)
        if (pk::kLayout.st[i].param == 7) return pk::kLayout.st[i].offset;
    return 0;
}

// cos / sin of a = fl32(2^i pi) * x for |a| up to a few thousand: Cody-Waite reduction by 2 pi in two FMAs,
// then the MUFU approximations on [-pi, pi] (abs error ~1e-6, far below bf16 resolution).
__device__ __forceinline__ void fast_sincos(float a, float& s, float& c) {
    const float k = rintf(a * tcm::kInv2Pi);
    float r = fmaf(-k, tcm::k2PiHi, a);
    r = fmaf(-k, tcm::k2PiLo, r);
    s = __sinf(r);
    c = __cosf(r);
}

template <int L>
__device__ __forceinline__ void encode_row(const float (&x)[3], uint32_t (&v)[32]) {
    // per frequency: [cos x, cos y, cos z, sin x, sin y, sin z] (nerf_model.py:29-31) -> 3 packed registers
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const float f = tcm::kPiF * (float)(1 << i);
        float s[3], c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) fast_sincos(__fmul_rn(f, x[k]), s[k], c[k]);
        v[3 * i + 0] = umma::pack_bf16(c[0], c[1]);
        v[3 * i + 1] = umma::pack_bf16(c[2], s[0]);
        v[3 * i + 2] = umma::pack_bf16(s[1], s[2]);
    }
#pragma unroll
    for (int j = 3 * L; j < 32; ++j) v[j] = 0u;
}


}  // namespace nerf
