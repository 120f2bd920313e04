// mlp_tc.cu - K8: NeRFModel.forward (nerf_model.py:362-389) as one persistent, warp-specialised sm_100a
// kernel: positional encoding -> 8x256 MLP with the skip concat -> sigma / rgb heads, bf16 operands on the
// 5th-gen tensor cores (tcgen05.mma, cta_group::1, M = 128 samples per tile), fp32 accumulation in TMEM.
//
// Per CTA (one per SM, persistent over 128-sample tiles):
//   warps 0,3   weight producers: stream the network's pre-swizzled bf16 weight image (pack_layout.cuh; 0.92 MB per
//               tile, L2-resident) through a 4-slot x 32 KB shared-memory ring with cp.async.bulk (TMA engine),
//               full/empty mbarriers.  K-blocks consumed back to back travel as ONE copy of up to 32 KB and the two
//               warps alternate stages: the copy engine retires ~one request per 550 clk per issuing lane whatever its
//               size, so 16 KB requests from one lane capped the stream at ~30 B/clk/SM (measured).
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.  Hidden activations never touch shared
//               memory: they live in TMEM as the A operand (bf16 pairs, lane = sample); only PE(x) / PE(dir)
//               come from swizzled shared-memory tiles (skip concat = one extra K block accumulated into the
//               same D).  Each layer is issued as two N=128 halves into the two halves of a 256-column
//               accumulator; layer l+1's first K blocks start as soon as the epilogue of layer l's first
//               half has written them, so MMA and epilogue overlap inside a single tile.
//   warp 2      TMEM allocation (512 columns: D 0..255, A0 256..383, A1 384..511).
//   warps 4-11  epilogue: tcgen05.ld the finished half, + bias, ReLU, round to bf16 (cvt.rn.relu.bf16x2),
//               tcgen05.st into the other A buffer; final steps apply ReLU / sigmoid to the sigma and rgb
//               columns and store them.  At tile start the same warps compute PE(x) (warps 4-7) and PE(dir)
//               (warps 8-11) into shared memory.
//
// Step / barrier protocol: steps are numbered globally (gs); step gs uses D region gs&1 and the barrier pair
// dfull[gs&1] (MMA -> epilogue, tcgen05.commit) / edone[gs&1] (epilogue -> MMA, 256 arrivals).
#include <stdlib.h>
#include "mlp_tc_common.cuh"

namespace nerf {

namespace tc {
constexpr int kTileM = 128;
constexpr int kSlots = 4;
constexpr uint32_t kSlotBytes = 32768;
constexpr int kThreads = 512;          // warps 0,3 weight producers, 1 MMA, 2 TMEM alloc, 4-11 epilogue, 12-15 PE
constexpr int kEpiWarps = 8;
constexpr int kPEWarps = 4;
constexpr uint32_t kColD = 0, kColA0 = 256, kColA1 = 384;

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t kOffPE = 0;             // 2 x [128 x 64] bf16 PE(x) tiles (double buffered across tiles)
constexpr uint32_t kOffPEDir = 32768;      // 2 x [128 x 64] bf16 PE(dir) tiles
constexpr uint32_t kOffRing = 65536;
constexpr uint32_t kOffBias = kOffRing + kSlots * kSlotBytes;
constexpr uint32_t kOffBars = kOffBias + ((pk::kBiasFloats * 4 + 15) / 16) * 16;
constexpr uint32_t kNumBars = 2 * kSlots + 8;
constexpr uint32_t kOffTmemHolder = kOffBars + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemHolder + 16 + 1024;   // + alignment slack
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

}  // namespace tc

template <bool PROFILE>
__global__ void __launch_bounds__(tc::kThreads, 1)
mlp_tc_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ o_rays, const float* __restrict__ d_rays,
              const float* __restrict__ ts, const float* __restrict__ samples, int64_t total, int S,
              float* __restrict__ sigma_out, float* __restrict__ rgb_out, __nv_bfloat16* __restrict__ act_out,
              unsigned long long* __restrict__ mask_out, long long* __restrict__ dbg) {
    long long prof[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem + tc::kOffRing;
    float* sBias = (float*)(smem + tc::kOffBias);
    uint64_t* bars = (uint64_t*)(smem + tc::kOffBars);
    uint64_t* full = bars;
    uint64_t* empty = bars + tc::kSlots;
    uint64_t* dfull = bars + 2 * tc::kSlots;
    uint64_t* edone = dfull + 2;
    uint64_t* pe_full = edone + 2;
    uint64_t* pe_empty = pe_full + 2;
    uint32_t* tmem_holder = (uint32_t*)(smem + tc::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (total + tc::kTileM - 1) / tc::kTileM;

    if (tid == 0) {
        for (int i = 0; i < tc::kSlots; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&dfull[i], 1);
            umma::mbar_init(&edone[i], tc::kEpiWarps);
            umma::mbar_init(&pe_full[i], tc::kPEWarps);
            umma::mbar_init(&pe_empty[i], 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 2) umma::tmem_alloc(tmem_holder, 512);
    {   // biases: resident for the whole kernel
        const float* gb = (const float*)(packed + pk::kLayout.bias_offset);
        for (int i = tid; i < pk::kBiasFloats; i += tc::kThreads) sBias[i] = gb[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 0 || warp == 3) {
        // ------------------------------------------------------------------ weight producers: two warps alternate the merged
        // stages so that two bulk copies are always in flight (one issuing lane sustains only ~1 request / 550 clk)
        const bool leader = umma::elect_one();
        const uint32_t me = (warp == 0) ? 0u : 1u;
        uint32_t cnt = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int s = 0; s < kMergedStages; ++s, ++cnt) {
                if ((cnt & 1u) != me) continue;
                const uint32_t slot = cnt % tc::kSlots, ph = (cnt / tc::kSlots) & 1;
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&empty[slot], ph ^ 1);
                NERF_PROF_END(tw, 4)
                if (leader) {
                    const StageRef st = c_merged.s[s];
                    umma::mbar_arrive_expect_tx(&full[slot], st.bytes);
                    umma::bulk_g2s(sRing + slot * tc::kSlotBytes, packed + st.offset, st.bytes, &full[slot]);
                }
                __syncwarp();
            }
        }
        if (PROFILE && lane == 0 && warp == 0) dbg[blockIdx.x * 16 + 4] = prof[4];
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform, one lane issues)
        const bool leader = umma::elect_one();
        constexpr uint32_t kI128 = umma::make_idesc_bf16(128, 128);
        constexpr uint32_t kI16 = umma::make_idesc_bf16(128, 16);
        uint32_t cnt = 0;         // weight stages consumed
        int64_t gs = 0;           // global step index
        int64_t e_waited = 0;     // epilogue steps already observed
        uint32_t tile_iter = 0;
        NERF_PROF_BEGIN(t_mma_total)

        auto ensure_e = [&](int64_t k) {       // epilogue of global step k (and all before it) is done
            NERF_PROF_BEGIN(tw)
            while (e_waited <= k) {
                umma::mbar_wait(&edone[e_waited & 1], (uint32_t)((e_waited >> 1) & 1));
                ++e_waited;
            }
            NERF_PROF_END(tw, 2)
            umma::tc_fence_after();
        };
        // one merged weight stage = nkb K=64 blocks ([rows x 64] tiles of tile_bytes each, back to back in the slot), nk16
        // K=16 slices per block; A from TMEM columns (a_col >= 0, 32 columns per block) or from a smem tile (1 block)
        auto stage = [&](uint32_t d_col, int a_col, uint64_t a_desc, uint32_t idesc, int nkb, uint32_t tile_bytes, int nk16,
                         uint32_t& acc) {
            const uint32_t slot = cnt % tc::kSlots, ph = (cnt / tc::kSlots) & 1;
            NERF_PROF_BEGIN(tw)
            umma::mbar_wait(&full[slot], ph);
            NERF_PROF_END(tw, 1)
            umma::tc_fence_after();
            if (leader) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint64_t bdesc = umma::make_desc_k_sw128(umma::smem_u32(sRing + slot * tc::kSlotBytes + kb * tile_bytes));
#pragma unroll 4
                    for (int k = 0; k < nk16; ++k) {
                        if (a_col >= 0)
                            umma::mma_ts(tmem + d_col, tmem + (uint32_t)a_col + 32u * kb + 8u * k, bdesc + 2u * k, idesc,
                                         acc | (uint32_t)(kb | k));
                        else
                            umma::mma_ss(tmem + d_col, a_desc + 2u * k, bdesc + 2u * k, idesc, acc | (uint32_t)k);
                    }
                }
                umma::mma_commit(&empty[slot]);
            }
            __syncwarp();
            acc = 1;
            ++cnt;
        };
        // a hidden layer read from A buffer `a_base` (K = 256), optionally preceded by the PE(x) K block
        auto layer = [&](int a_base, bool pe_first, uint64_t descPE) {
            for (int h = 0; h < 2; ++h) {
                const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                if (pe_first) stage(d_col, -1, descPE, kI128, 1, 16384, 4, acc);
                stage(d_col, a_base + 0, 0, kI128, 2, 16384, 4, acc);
                if (h == 0) ensure_e(gs - 1);
                stage(d_col, a_base + 64, 0, kI128, 2, 16384, 4, acc);
                if (leader) umma::mma_commit(&dfull[gs & 1]);
                __syncwarp();
                ++gs;
            }
        };

        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const uint32_t pb = tile_iter & 1;
            const uint64_t descPE = umma::make_desc_k_sw128(umma::smem_u32(smem + tc::kOffPE + pb * 16384));
            const uint64_t descPEDir = umma::make_desc_k_sw128(umma::smem_u32(smem + tc::kOffPEDir + pb * 16384));
            NERF_PROF_BEGIN(tw)
            umma::mbar_wait(&pe_full[pb], (tile_iter >> 1) & 1);
            NERF_PROF_END(tw, 3)
            umma::tc_fence_after();
            for (int h = 0; h < 2; ++h) {                                   // mlp.0: A = PE(x) tile
                uint32_t acc = 0;
                ensure_e(gs - 2);
                stage(tc::kColD + 128u * (uint32_t)(gs & 1), -1, descPE, kI128, 1, 16384, 4, acc);
                if (leader) umma::mma_commit(&dfull[gs & 1]);
                __syncwarp();
                ++gs;
            }
            layer(tc::kColA0, false, descPE);     // mlp.2        reads A0 (epilogue writes A1)
            layer(tc::kColA1, false, descPE);     // mlp.4        reads A1
            layer(tc::kColA0, false, descPE);     // mlp.6        reads A0
            layer(tc::kColA1, true, descPE);      // feature_fn.0 reads PE(x) + A1
            layer(tc::kColA0, false, descPE);     // feature_fn.2 reads A0
            layer(tc::kColA1, false, descPE);     // feature_fn.4 reads A1, feat -> A0
            {                                     // rgb_fn.0: PE(dir) + feat (A0) -> D region, r -> A1[0:64]
                const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                stage(d_col, -1, descPEDir, kI128, 1, 16384, 2, acc);
                if (leader) umma::mma_commit(&pe_empty[pb]);      // last read of this tile's PE buffers
                __syncwarp();
                stage(d_col, tc::kColA0 + 0, 0, kI128, 2, 16384, 4, acc);
                ensure_e(gs - 1);
                stage(d_col, tc::kColA0 + 64, 0, kI128, 2, 16384, 4, acc);
                if (leader) umma::mma_commit(&dfull[gs & 1]);
                __syncwarp();
                ++gs;
            }
            {                                     // density_fn.0: feat (A0) -> 16 columns
                const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                stage(d_col, tc::kColA0, 0, kI16, 4, 2048, 4, acc);
                if (leader) umma::mma_commit(&dfull[gs & 1]);
                __syncwarp();
                ++gs;
            }
            {                                     // rgb_fn.2: r (A1[0:64]) -> 16 columns
                const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                stage(d_col, tc::kColA1, 0, kI16, 2, 2048, 4, acc);
                if (leader) umma::mma_commit(&dfull[gs & 1]);
                __syncwarp();
                ++gs;
            }
        }
        NERF_PROF_END(t_mma_total, 0)
        if (PROFILE && lane == 0) { for (int i = 0; i < 4; ++i) dbg[blockIdx.x * 16 + i] = prof[i]; dbg[blockIdx.x * 16 + 8] = tile_iter; }
    } else if (warp >= 12) {
        // ------------------------------------------------------------------ PE producers (128 threads, thread = row)
        const int r = (warp - 12) * 32 + lane;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t pb = it & 1;
            const int64_t row = tile * tc::kTileM + r;
            const bool valid = row < total;
            const int64_t n = valid ? row / S : 0;
            float x[3] = {0.f, 0.f, 0.f}, u[3] = {0.f, 0.f, 0.f};
            if (valid) {
                const float dx = __ldg(d_rays + n * 3), dy = __ldg(d_rays + n * 3 + 1), dz = __ldg(d_rays + n * 3 + 2);
                if (samples) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) x[k] = samples[row * 3 + k];
                } else {
                    const float t = ts[row];                      // d * t + o (nerf_helpers.py:55)
                    x[0] = __fadd_rn(__fmul_rn(dx, t), __ldg(o_rays + n * 3 + 0));
                    x[1] = __fadd_rn(__fmul_rn(dy, t), __ldg(o_rays + n * 3 + 1));
                    x[2] = __fadd_rn(__fmul_rn(dz, t), __ldg(o_rays + n * 3 + 2));
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) x[k] = __fdiv_rn(x[k], tcm::kPiF);           // nerf_model.py:377
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);                    // nerf_model.py:373
                u[0] = __fdiv_rn(dx, nrm); u[1] = __fdiv_rn(dy, nrm); u[2] = __fdiv_rn(dz, nrm);
            }
            uint32_t v[32];
            encode_row<10>(x, v);
            umma::mbar_wait(&pe_empty[pb], ((it >> 1) & 1) ^ 1);      // MMAs of tile it-2 no longer read buffer pb
            store_row_sw128(smem + tc::kOffPE + pb * 16384, r, v);
            encode_row<4>(u, v);
            store_row_sw128(smem + tc::kOffPEDir + pb * 16384, r, v);
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&pe_full[pb]);
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int wh = (warp - 4) >> 2;         // column half handled by this warp
        const int r = q * 32 + lane;            // row (sample) inside the tile
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int64_t gs = 0;
        NERF_PROF_BEGIN(t_epi_total)

        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t row = tile * tc::kTileM + r;
            const bool valid = row < total;
            // ---- 7 hidden layers x 2 halves, then rgb_fn.0: 128 columns -> bias, (ReLU), bf16 -> A buffer
            for (int s = 0; s < 15; ++s, ++gs) {
                const int layer = s >> 1, nhalf = (s < 14) ? (s & 1) : 0;
                const bool relu = (layer != 6);                       // feature_fn.4 is linear (nerf_model.py:347)
                // writes: mlp.0 -> A0, mlp.2 -> A1, mlp.4 -> A0, mlp.6 -> A1, ff.0 -> A0, ff.2 -> A1, ff.4 -> A0, rgb_fn.0 -> A1
                const uint32_t a_dst = (layer & 1) ? tc::kColA1 : tc::kColA0;
                const float* bias = sBias + (s < 14 ? layer * 256 + nhalf * 128 : pk::kBiasR0) + wh * 64;
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                NERF_PROF_END(tw, 6)
                umma::tc_fence_after();
                const uint32_t d_addr = tmem + lane_base + tc::kColD + 128u * (uint32_t)(gs & 1) + (uint32_t)(wh * 64);
                uint32_t v0[32], v1[32];
                umma::tmem_ld32(d_addr, v0);
                umma::tmem_ld32(d_addr + 32, v1);
                umma::tmem_wait_ld();
                uint32_t p[16];
                // output feature n = nhalf*128 + wh*64 + j is K index n of the next layer: TMEM column n/2
                const uint32_t a_addr = tmem + lane_base + a_dst + (uint32_t)((nhalf * 128 + wh * 64) >> 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = *(const float4*)(bias + 4 * j);
                    const float x0 = __uint_as_float(v0[4 * j + 0]) + b4.x, x1 = __uint_as_float(v0[4 * j + 1]) + b4.y;
                    const float x2 = __uint_as_float(v0[4 * j + 2]) + b4.z, x3 = __uint_as_float(v0[4 * j + 3]) + b4.w;
                    p[2 * j + 0] = relu ? umma::pack_bf16_relu(x0, x1) : umma::pack_bf16(x0, x1);
                    p[2 * j + 1] = relu ? umma::pack_bf16_relu(x2, x3) : umma::pack_bf16(x2, x3);
                }
                umma::tmem_st16(a_addr, p);
                // training: keep what the next layer consumes (post-activation bf16) in the tiled chunk-major layout
                // (pack_layout.cuh): outputs of mlp.0..feature_fn.4 at feature 256*layer, rgb_fn.0 at 1792.  Lanes are
                // consecutive rows, so every 16-byte store of the warp lands in one contiguous 512-byte run.  Rows past
                // `total` are stored too (finite values; their dz is zero) so wgrad can read whole tiles.
                unsigned long long mbits = 0ull;
                if (act_out != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        mbits |= (unsigned long long)((p[j] & 0x00007FFFu) != 0u) << (2 * j);
                        mbits |= (unsigned long long)((p[j] & 0x7FFF0000u) != 0u) << (2 * j + 1);
                    }
                }
                uint4* act_chunk = nullptr;
                if (act_out != nullptr)
                    act_chunk = (uint4*)(act_out + pk::tiled_offset(row, (s < 14 ? layer * 256 + nhalf * 128 : 1792) + wh * 64, pk::kActChunks));
                if (act_chunk) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) act_chunk[j * 128] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b4 = *(const float4*)(bias + 32 + 4 * j);
                    const float x0 = __uint_as_float(v1[4 * j + 0]) + b4.x, x1 = __uint_as_float(v1[4 * j + 1]) + b4.y;
                    const float x2 = __uint_as_float(v1[4 * j + 2]) + b4.z, x3 = __uint_as_float(v1[4 * j + 3]) + b4.w;
                    p[2 * j + 0] = relu ? umma::pack_bf16_relu(x0, x1) : umma::pack_bf16(x0, x1);
                    p[2 * j + 1] = relu ? umma::pack_bf16_relu(x2, x3) : umma::pack_bf16(x2, x3);
                }
                umma::tmem_st16(a_addr + 16, p);
                if (act_chunk) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) act_chunk[(4 + j) * 128] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        mbits |= (unsigned long long)((p[j] & 0x00007FFFu) != 0u) << (32 + 2 * j);
                        mbits |= (unsigned long long)((p[j] & 0x7FFF0000u) != 0u) << (33 + 2 * j);
                    }
                    // sign bits of this row's 64 outputs (what the dgrad kernel masks with): 8 B instead of 128 B to re-read
                    const int blk = ((s < 14 ? layer * 256 + nhalf * 128 : 1792) + wh * 64) >> 6;
                    mask_out[((row >> 7) * pk::kMaskWords + blk) * 128 + (row & 127)] = mbits;
                }
                umma::tmem_wait_st();
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&edone[gs & 1]);
            }
            // ---- density_fn.0: column 0 -> sigma = relu(. + b) (nerf_model.py:350-353)
            {
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                NERF_PROF_END(tw, 6)
                umma::tc_fence_after();
                if (wh == 0) {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + tc::kColD + 128u * (uint32_t)(gs & 1), v);
                    umma::tmem_wait_ld();
                    if (valid) sigma_out[row] = fmaxf(__uint_as_float(v[0]) + sBias[pk::kBiasSigma], 0.f);
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&edone[gs & 1]);
                ++gs;
            }
            // ---- rgb_fn.2: columns 0..2 -> sigmoid(. + b) (nerf_model.py:358-359)
            {
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                NERF_PROF_END(tw, 6)
                umma::tc_fence_after();
                if (wh == 0) {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + tc::kColD + 128u * (uint32_t)(gs & 1), v);
                    umma::tmem_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float x = __uint_as_float(v[k]) + sBias[pk::kBiasRgb + k];
                            rgb_out[row * 3 + k] = 1.0f / (1.0f + __expf(-x));
                        }
                    }
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&edone[gs & 1]);
                ++gs;
            }
        }
        NERF_PROF_END(t_epi_total, 5)
        if (PROFILE && tid == 128) { for (int i = 5; i < 8; ++i) dbg[blockIdx.x * 16 + i] = prof[i]; }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

}  // namespace nerf

namespace nerf {
int launch_mlp_tc2(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                   int64_t total, int S, float* sigma, float* rgb, void* stream, long long* dbg);
int launch_mlp_tc3(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                   int64_t total, int S, float* sigma, float* rgb, void* act_out, void* mask_out, void* stream, long long* dbg,
                   const CompositeOutputs* comp);
bool mlp_tc3_can_composite(int S, int* group_rays, int* group_tiles);
}

using namespace nerf;

// NERF_TC_PAIR=1 selects the CTA-pair (cta_group::2) kernel of mlp_tc2.cu; the default is the 1-CTA kernel of this
// file, which measured faster on B200 (profiles/r01_notes.md): the per-half-layer MMA -> epilogue -> MMA dependency
// loop is latency-bound and the pair's cross-CTA signalling lengthens it more than the halved weight traffic saves.
static bool use_pair_kernel() {
    static const bool v = [] { const char* e = getenv("NERF_TC_PAIR"); return e && e[0] == '1'; }();
    return v;
}

// NERF_TC_ONE_TILE=1 selects the one-tile-per-CTA schedule of this file; the default is the two-tile schedule of
// mlp_tc3.cu (shared weight stages, the dependency loop of one tile hidden behind the other tile's MMAs).
static bool use_one_tile_kernel() {
    static const bool v = [] { const char* e = getenv("NERF_TC_ONE_TILE"); return e && e[0] == '1'; }();
    return v;
}

static int launch_mlp_tc(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                         int64_t N, int S, float* sigma, float* rgb, void* stream, long long* dbg = nullptr,
                         void* act_out = nullptr, void* mask_out = nullptr) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_forward_tc: bad size N=%lld S=%d", (long long)N, S);
    if (N == 0) return 0;
    NERF_REQUIRE(packed && d && sigma && rgb, "nerf_mlp_forward_tc: null pointer");
    NERF_REQUIRE(samples || (o && ts), "nerf_mlp_forward_tc: need either samples or (o, ts)");
    NERF_REQUIRE(((uintptr_t)packed & 127) == 0, "nerf_mlp_forward_tc: packed buffer must be 128-byte aligned");
    static thread_local bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        if (e != cudaSuccess) { set_error("nerf_mlp_forward_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attr_set = true;
    }
    const int64_t total = N * S;
    NERF_REQUIRE(!act_out || ((uintptr_t)act_out & 15) == 0, "nerf_mlp_forward_tc: act_out must be 16-byte aligned");
    if (use_pair_kernel() && !act_out) return launch_mlp_tc2(packed, o, d, ts, samples, total, S, sigma, rgb, stream, dbg);
    if (!use_one_tile_kernel()) return launch_mlp_tc3(packed, o, d, ts, samples, total, S, sigma, rgb, act_out, mask_out, stream, dbg, nullptr);
    const int64_t tiles = (total + tc::kTileM - 1) / tc::kTileM;
    int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    if (dbg) {                                        // diagnostic runs only: NERF_TC_MAX_CTAS limits the grid
        const char* e = getenv("NERF_TC_MAX_CTAS");
        if (e && atoi(e) > 0 && atoi(e) < grid) grid = atoi(e);
    }
    if (dbg)
        mlp_tc_kernel<true><<<grid, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>((const uint8_t*)packed, o, d, ts, samples,
                                                                                        total, S, sigma, rgb, (__nv_bfloat16*)act_out, (unsigned long long*)mask_out, dbg);
    else
        mlp_tc_kernel<false><<<grid, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>((const uint8_t*)packed, o, d, ts, samples,
                                                                                         total, S, sigma, rgb, (__nv_bfloat16*)act_out, (unsigned long long*)mask_out, nullptr);
    return check_launch("nerf_mlp_forward_tc");
}

// Diagnostic entry point (not in include/nerf_b200.h): same kernel with per-CTA cycle counters, dbg = [grid,16] int64.
extern "C" int nerf_debug_mlp_tc_profile(const void* packed, const float* o, const float* d, const float* ts,
                                         int64_t N, int S, float* sigma, float* rgb, long long* dbg, void* stream) {
    return launch_mlp_tc(packed, o, d, ts, nullptr, N, S, sigma, rgb, stream, dbg);
}

extern "C" int nerf_mlp_forward_tc(const void* packed, const float* o, const float* d, const float* ts,
                                   int64_t N, int S, float* sigma, float* rgb, void* stream) {
    return launch_mlp_tc(packed, o, d, ts, nullptr, N, S, sigma, rgb, stream);
}

// Points form: samples [N,S,3] given explicitly (the NeRFModel.forward(samples, direc) call surface).
extern "C" int nerf_mlp_forward_tc_points(const void* packed, const float* samples, const float* d,
                                          int64_t N, int S, float* sigma, float* rgb, void* stream) {
    return launch_mlp_tc(packed, nullptr, d, nullptr, samples, N, S, sigma, rgb, stream);
}

// Training form: also writes the bf16 activations every layer consumed (outputs of mlp.0, mlp.2, mlp.4, mlp.6,
// feature_fn.0, feature_fn.2, feature_fn.4 at feature 256*k, rgb_fn.0 at 1792) in the tiled chunk-major layout of
// pack_layout.cuh; act_out holds ceil(N*S/128)*128 rows x 1920 features.
extern "C" int nerf_mlp_forward_tc_train(const void* packed, const float* o, const float* d, const float* ts,
                                         int64_t N, int S, float* sigma, float* rgb, void* act_out, void* mask_out, void* stream) {
    NERF_REQUIRE(N == 0 || (act_out && mask_out), "nerf_mlp_forward_tc_train: act_out / mask_out is NULL");
    return launch_mlp_tc(packed, o, d, ts, nullptr, N, S, sigma, rgb, stream, nullptr, act_out, mask_out);
}

// ---- K8 + K2: the same network with alpha compositing (nerf_helpers.py:58-104) done inside the kernel by a dedicated warp
// (mlp_tc3.cu, COMP form).  sigma / rgb may be NULL (render: the per-sample outputs never leave the SM); act_out / mask_out
// non-NULL selects the training form, which needs sigma and rgb as well (the compositing backward reads them).
// (The diagnostic one-tile / CTA-pair schedules selected by NERF_TC_ONE_TILE / NERF_TC_PAIR have no fused form: report
// "unsupported" so that callers take the two-launch path and the matching dgrad kernel sees its own mask format.)
extern "C" int nerf_mlp_composite_tc_supported(int S) {
    return (mlp_tc3_can_composite(S, nullptr, nullptr) && !use_one_tile_kernel() && !use_pair_kernel()) ? 1 : 0;
}

extern "C" int nerf_mlp_composite_tc(const void* packed, const float* o, const float* d, const float* ts, int64_t N, int S,
                                     float* sigma, float* rgb, void* act_out, void* mask_out,
                                     float* weights, float* ray_rgb, float* depth, float* acc, float* stats4, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_composite_tc: bad size N=%lld S=%d", (long long)N, S);
    if (N == 0) return 0;
    NERF_REQUIRE(packed && o && d && ts && ray_rgb, "nerf_mlp_composite_tc: null pointer");
    NERF_REQUIRE((sigma == nullptr) == (rgb == nullptr), "nerf_mlp_composite_tc: sigma and rgb go together");
    NERF_REQUIRE((act_out == nullptr) == (mask_out == nullptr), "nerf_mlp_composite_tc: act_out and mask_out go together");
    NERF_REQUIRE(!act_out || sigma, "nerf_mlp_composite_tc: the training form also needs sigma / rgb");
    NERF_REQUIRE(((uintptr_t)packed & 127) == 0, "nerf_mlp_composite_tc: packed buffer must be 128-byte aligned");
    NERF_REQUIRE(!act_out || ((uintptr_t)act_out & 15) == 0, "nerf_mlp_composite_tc: act_out must be 16-byte aligned");
    NERF_REQUIRE(nerf_mlp_composite_tc_supported(S),
                 "nerf_mlp_composite_tc: S = %d is not supported (needs S %% 32 == 0 and a ray group of at most 6 tiles); "
                 "use nerf_mlp_forward_tc + nerf_composite", S);
    const CompositeOutputs comp{weights, ray_rgb, depth, acc, stats4, nullptr, nullptr, 0.f, nullptr};
    return launch_mlp_tc3(packed, o, d, ts, nullptr, N * S, S, sigma, rgb, act_out, mask_out, stream, nullptr, &comp);
}

// The coarse network's form: K1 (generate_coarse_samples, nerf_helpers.py:28-56) runs inside the kernel too.  u [N,S] uniforms,
// t_base [S] = the reference's torch.arange(near, far, step), ts_out [N,S] receives the depths t = t_base[i] + u * step (bit-identical
// to nerf_coarse_sample); everything else as nerf_mlp_composite_tc.
extern "C" int nerf_mlp_composite_tc_strata(const void* packed, const float* o, const float* d, const float* u, const float* t_base,
                                            float step, int64_t N, int S, float* ts_out, float* sigma, float* rgb, void* act_out,
                                            void* mask_out, float* weights, float* ray_rgb, float* depth, float* acc, float* stats4,
                                            void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_composite_tc_strata: bad size N=%lld S=%d", (long long)N, S);
    if (N == 0) return 0;
    NERF_REQUIRE(packed && o && d && u && t_base && ts_out && ray_rgb, "nerf_mlp_composite_tc_strata: null pointer");
    NERF_REQUIRE((sigma == nullptr) == (rgb == nullptr), "nerf_mlp_composite_tc_strata: sigma and rgb go together");
    NERF_REQUIRE((act_out == nullptr) == (mask_out == nullptr), "nerf_mlp_composite_tc_strata: act_out and mask_out go together");
    NERF_REQUIRE(!act_out || sigma, "nerf_mlp_composite_tc_strata: the training form also needs sigma / rgb");
    NERF_REQUIRE(((uintptr_t)packed & 127) == 0, "nerf_mlp_composite_tc_strata: packed buffer must be 128-byte aligned");
    NERF_REQUIRE(!act_out || ((uintptr_t)act_out & 15) == 0, "nerf_mlp_composite_tc_strata: act_out must be 16-byte aligned");
    NERF_REQUIRE(nerf_mlp_composite_tc_supported(S), "nerf_mlp_composite_tc_strata: S = %d is not supported", S);
    const CompositeOutputs comp{weights, ray_rgb, depth, acc, stats4, u, t_base, step, ts_out};
    return launch_mlp_tc3(packed, o, d, nullptr, nullptr, N * S, S, sigma, rgb, act_out, mask_out, stream, nullptr, &comp);
}
