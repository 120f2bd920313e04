// mlp_tc_common.cuh - device helpers shared by the 1-CTA (mlp_tc.cu) and CTA-pair (mlp_tc2.cu) tensor-core kernels:
// the weight-stage table, the range-reduced MUFU sin/cos, and the swizzled PE tile writers.
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
struct StageTable { StageRef s[pk::kStages]; };
constexpr StageTable make_stage_table() {
    StageTable t{};
    for (int i = 0; i < pk::kStages; ++i) {
        t.s[i].offset = pk::kLayout.st[i].offset;
        t.s[i].bytes = (uint32_t)pk::kLayout.st[i].rows * 128u;
    }
    return t;
}
static __constant__ StageTable c_stages = make_stage_table();

// Merged stages for the 1-CTA forward kernel: the copy engine completes about one request per ~550 clk per issuing
// lane whatever its size (profiles/r01_notes.md), so K-blocks that are consumed back to back are fetched as ONE bulk copy
// of up to 32 KB (the packed image is contiguous in consumption order; only this table changes):
//   mlp.0: 1 block per half | 256-wide layers: 2 + 2 blocks per half | feature_fn.0 / rgb_fn.0: PE block, then 2 + 2 |
//   density_fn.0: its 4 small blocks | rgb_fn.2: its 2 small blocks.
constexpr int kMergedStages = 33;
struct MergedTable { StageRef s[kMergedStages]; };
constexpr MergedTable make_merged_table() {
    MergedTable t{};
    int m = 0, i = 0;
    auto take = [&](int nblocks) {
        t.s[m].offset = pk::kLayout.st[i].offset;
        uint32_t bytes = 0;
        for (int j = 0; j < nblocks; ++j) bytes += (uint32_t)pk::kLayout.st[i + j].rows * 128u;
        t.s[m].bytes = bytes;
        i += nblocks;
        ++m;
    };
    take(1); take(1);                                            // mlp.0
    for (int l = 0; l < 3; ++l) for (int h = 0; h < 2; ++h) { take(2); take(2); }      // mlp.2/4/6
    for (int h = 0; h < 2; ++h) { take(1); take(2); take(2); }                        // feature_fn.0
    for (int l = 0; l < 2; ++l) for (int h = 0; h < 2; ++h) { take(2); take(2); }      // feature_fn.2/4
    take(1); take(2); take(2);                                                         // rgb_fn.0
    take(4);                                                                           // density_fn.0
    take(2);                                                                           // rgb_fn.2
    return t;
}
static __constant__ MergedTable c_merged = make_merged_table();
static_assert(make_merged_table().s[kMergedStages - 1].offset + make_merged_table().s[kMergedStages - 1].bytes ==
                  pk::kLayout.weight_bytes, "merged stage table must cover the whole weight image");

// cos / sin of a = fl32(2^i pi) * x for |a| up to a few thousand: Cody-Waite reduction by 2 pi in two FMAs,
// then the MUFU approximations on [-pi, pi] (abs error ~1e-6, far below bf16 resolution).
__device__ __forceinline__ void fast_sincos(float a, float& s, float& c) {
    const float k = rintf(a * tcm::kInv2Pi);
    float r = fmaf(-k, tcm::k2PiHi, a);
    r = fmaf(-k, tcm::k2PiLo, r);
    s = __sinf(r);
    c = __cosf(r);
}

// Row `r` of a [128 x 64] bf16 K-major 128B-swizzled tile <- 32 packed registers (64 bf16).
__device__ __forceinline__ void store_row_sw128(uint8_t* tile, int r, const uint32_t (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 q = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        *(uint4*)(tile + r * 128 + ((c ^ (r & 7)) << 4)) = q;
    }
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


// per-CTA cycle counters of the diagnostic builds (dbg[blockIdx.x * 16 + i]):
//   0 MMA warp total, 1 MMA wait(full = weights), 2 MMA wait(edone = epilogue), 3 MMA wait(pe_full),
//   4 producer wait(empty), 5 epilogue total, 6 epilogue wait(dfull), 7 unused, 8 tiles
#define NERF_PROF_BEGIN(var) long long var = 0; if (PROFILE) var = clock64();
#define NERF_PROF_END(var, slot) if (PROFILE) prof[slot] += clock64() - var;

}  // namespace nerf
