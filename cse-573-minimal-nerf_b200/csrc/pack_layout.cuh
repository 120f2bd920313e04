// pack_layout.cuh - layout of one network's packed weights and the per-tile step/stage schedule shared by
// the packer (pack.cu) and the fused tensor-core kernels (mlp_tc3.cu, mlp_tc_bwd3.cu, wgrad_tc.cu).
//
// One NeRFModel (nerf_model.py:331-360, position_dim 10, direction_dim 4) is evaluated per 128-sample tile
// as 17 "steps", each one accumulator region (<=128 fp32 columns of TMEM) produced by a run of K=64 weight
// "stages".  A stage is a [rows x 64] bf16 tile of nn.Linear.weight[out,in] (out = rows, in = K), stored
// K-major with the 128-byte swizzle the UMMA shared-memory descriptor expects, so the global image can be
// bulk-copied into shared memory unchanged.  Stages are stored in exactly the order the kernel consumes them.
//
//   step  layer (state_dict key)  out rows            K stages (source columns of W)
//   0,1   mlp.0            N-half 0 / 1      PE(x)[0:60]
//   2..7  mlp.2/4/6        N-half 0 / 1      h[0:64] h[64:128] h[128:192] h[192:256]
//   8,9   feature_fn.0     N-half 0 / 1      PE(x) = W[:,256:316] first, then h[0:256]
//   10..13 feature_fn.2/4  N-half 0 / 1      h[0:256]
//   14    rgb_fn.0         128               PE(dir) = W[:,256:280] first, then feat[0:256]
//   15    density_fn.0     1 (padded to 16)  feat[0:256]
//   16    rgb_fn.2         3 (padded to 16)  r[0:128]
#pragma once
#include <stdint.h>

namespace nerf {
namespace pk {

constexpr int kStages = 63;
constexpr int kSteps = 17;
constexpr int kStageBytesFull = 128 * 128;   // 128 rows x 64 bf16
constexpr int kStageBytesSmall = 16 * 128;   // 16 rows x 64 bf16

struct Stage {
    uint8_t param;     // index of the weight matrix in state_dict order (0..9)
    uint8_t rows;      // 128 or 16 rows in the tile
    uint8_t valid_rows;
    uint8_t kvalid;    // valid K columns (<= 64), rest zero
    uint16_t n0;       // first output row of W in this tile
    uint16_t k0;       // first input column of W in this tile
    uint16_t in_features;
    uint32_t offset;   // byte offset inside the packed buffer
    uint8_t transposed;  // 0: tile[r][k] = W[n0+r][k0+k] (forward);  1: tile[r][k] = W[k0+k][n0+r] (dgrad: B = W^T)
};

struct Layout {
    Stage st[kStages];
    uint32_t weight_bytes;
    uint32_t bias_offset;   // byte offset of the fp32 bias block
    uint32_t total_bytes;
};

// bias block (floats): mlp.0 0, mlp.2 256, mlp.4 512, mlp.6 768, feature_fn.0 1024, feature_fn.2 1280,
// feature_fn.4 1536, rgb_fn.0 1792 (128), density_fn.0 1920 (pad 4), rgb_fn.2 1924 (pad 4)
constexpr int kBiasFloats = 1928;
constexpr int kBiasR0 = 1792, kBiasSigma = 1920, kBiasRgb = 1924;
constexpr int kActFeatures = 1920;   // saved bf16 activations per sample (training): 7 x 256 + 128
// Training tensors (saved activations `acts`, pre-activation gradients `dz`) are stored TILED CHUNK-MAJOR:
//   element (row, feature f) of 128-row tile t lives at  ((t * kChunks + f/8) * 128 + row%128) * 8 + f%8   (bf16 elements)
// i.e. per tile, 16-byte chunks of 8 features are the slow index and the 128 rows the fast one.  A warp whose lanes are
// consecutive rows therefore reads/writes 512 contiguous bytes per instruction, and a [128 rows x 64 features] block is a
// contiguous 16 KB that is, as it stands, the UMMA no-swizzle MN-major canonical layout (core-matrix strides: 2048 B along
// the features, 128 B along the rows) that wgrad (contraction over rows) consumes.
// dz has two extra chunks (features 1920..1935): [dsigma_pre, drgb_pre x3, 0 ...] in bf16, the A/B operand of the two
// head weight gradients.
// ReLU sign bits for the dgrad kernel: one 64-bit word per (row, 64-feature block), bit i = [activation 64b+i > 0];
// word index ((row/128) * kMaskWords + b) * 128 + row%128, b = feature/64 (30 blocks: 28 hidden + 2 of rgb_fn.0's output).
constexpr int kMaskWords = kActFeatures / 64;       // 30
constexpr int kActChunks = kActFeatures / 8;        // 240
constexpr int kDzFeatures = kActFeatures + 16;      // 1936
constexpr int kDzChunks = kDzFeatures / 8;          // 242
__host__ __device__ constexpr int64_t tiled_offset(int64_t row, int feature, int chunks_per_tile) {
    return (((row >> 7) * chunks_per_tile + (feature >> 3)) * 128 + (row & 127)) * 8 + (feature & 7);
}

constexpr Layout make_layout() {
    Layout L{};
    int s = 0;
    uint32_t off = 0;
    auto add = [&](int param, int rows, int valid_rows, int n0, int k0, int kvalid, int in_features) {
        L.st[s].param = (uint8_t)param; L.st[s].rows = (uint8_t)rows; L.st[s].valid_rows = (uint8_t)valid_rows;
        L.st[s].kvalid = (uint8_t)kvalid; L.st[s].n0 = (uint16_t)n0; L.st[s].k0 = (uint16_t)k0;
        L.st[s].in_features = (uint16_t)in_features; L.st[s].offset = off;
        off += (uint32_t)rows * 128u;
        ++s;
    };
    for (int h = 0; h < 2; ++h) add(0, 128, 128, 128 * h, 0, 60, 60);                       // mlp.0
    for (int l = 1; l <= 3; ++l)                                                             // mlp.2/4/6
        for (int h = 0; h < 2; ++h)
            for (int kb = 0; kb < 4; ++kb) add(l, 128, 128, 128 * h, 64 * kb, 64, 256);
    for (int h = 0; h < 2; ++h) {                                                            // feature_fn.0
        add(4, 128, 128, 128 * h, 256, 60, 316);
        for (int kb = 0; kb < 4; ++kb) add(4, 128, 128, 128 * h, 64 * kb, 64, 316);
    }
    for (int l = 5; l <= 6; ++l)                                                             // feature_fn.2/4
        for (int h = 0; h < 2; ++h)
            for (int kb = 0; kb < 4; ++kb) add(l, 128, 128, 128 * h, 64 * kb, 64, 256);
    add(8, 128, 128, 0, 256, 24, 280);                                                       // rgb_fn.0: PE(dir) part
    for (int kb = 0; kb < 4; ++kb) add(8, 128, 128, 0, 64 * kb, 64, 280);
    for (int kb = 0; kb < 4; ++kb) add(7, 16, 1, 0, 64 * kb, 64, 256);                       // density_fn.0
    for (int kb = 0; kb < 2; ++kb) add(9, 16, 3, 0, 64 * kb, 64, 128);                       // rgb_fn.2
    L.weight_bytes = off;
    L.bias_offset = (off + 127u) & ~127u;
    L.total_bytes = L.bias_offset + kBiasFloats * 4u;
    return L;
}

constexpr Layout kLayout = make_layout();
static_assert(kLayout.weight_bytes == 57u * kStageBytesFull + 6u * kStageBytesSmall, "stage accounting");


// ---- backward (dgrad) image: B = W^T stages in the order mlp_tc_bwd3.cu consumes them, then two fp32 constant blocks
//   steps 0,1   rgb_fn.0 dgrad   d feat[:, 128h:128h+128] = dr[128 x K=128] . W8[:, 128h:...]      2 stages per half
//   steps 2..13 feature_fn.4, feature_fn.2, feature_fn.0 (h part), mlp.6, mlp.4, mlp.2              4 stages per half
constexpr int kStagesT = 52;
struct LayoutT {
    Stage st[kStagesT];
    uint32_t weight_bytes;
    uint32_t const_offset;   // fp32: rgb_fn.2.weight [3,128] (384 floats) then density_fn.0.weight [256]
    uint32_t total_bytes;
};
constexpr int kConstFloatsT = 640;

constexpr LayoutT make_layout_t() {
    LayoutT L{};
    int s = 0;
    uint32_t off = 0;
    auto add = [&](int param, int kin0, int nout0, int in_features) {
        L.st[s].param = (uint8_t)param; L.st[s].rows = 128; L.st[s].valid_rows = 128; L.st[s].kvalid = 64;
        L.st[s].n0 = (uint16_t)kin0; L.st[s].k0 = (uint16_t)nout0; L.st[s].in_features = (uint16_t)in_features;
        L.st[s].offset = off; L.st[s].transposed = 1;
        off += 128u * 128u;
        ++s;
    };
    for (int h = 0; h < 2; ++h)
        for (int kb = 0; kb < 2; ++kb) add(8, 128 * h, 64 * kb, 280);                        // rgb_fn.0
    const int order[6] = {6, 5, 4, 3, 2, 1};                                                 // feature_fn.4 ... mlp.2
    for (int i = 0; i < 6; ++i)
        for (int h = 0; h < 2; ++h)
            for (int kb = 0; kb < 4; ++kb) add(order[i], 128 * h, 64 * kb, order[i] == 4 ? 316 : 256);
    L.weight_bytes = off;
    L.const_offset = off;
    L.total_bytes = off + kConstFloatsT * 4u;
    return L;
}
constexpr LayoutT kLayoutT = make_layout_t();
static_assert(kLayoutT.weight_bytes == 52u * kStageBytesFull, "backward stage accounting");

}  // namespace pk
}  // namespace nerf
