// mlp_tc2.cu - the CTA-pair (tcgen05 cta_group::2) form of the fused NeRFModel kernel of mlp_tc.cu.
//
// Two CTAs of a 2-CTA cluster (one SM pair) each own a 128-sample tile; the leader CTA issues M = 256 MMAs
// that update both tiles.  The weight operand B of every MMA is split between the two CTAs' shared
// memories (each holds half of the N rows), so
//   * each SM fetches only HALF of the 0.92 MB weight stream per tile pair (L2 -> SM traffic halves), and
//   * a ring slot holds half a stage, so the same shared memory buffers twice as many stages ahead
//     (19 x 8 KB): the 1-CTA kernel was bound by exactly this prefetch depth (per-tile time identical at
//     8 and 148 CTAs, MMA warp waiting ~25 % of the time on `full`).
// Schedule, TMEM map, epilogue and PE production are those of mlp_tc.cu; what changes is the signalling:
//   full[slot]   lives in the leader: its producer expects the whole stage; both CTAs' halves arrive through
//                tensor-map TMA (cp.async.bulk.tensor ... cta_group::2), which may signal the leader's barrier
//   empty/dfull/pe_empty   tcgen05.commit multicast to both CTAs
//   edone / pe_full        live in the leader; the peer's warps arrive remotely (mapa + mbarrier.arrive.cluster)
#include <cuda.h>
#include <stdlib.h>
#include "mlp_tc_common.cuh"

namespace nerf {

namespace tc2 {
constexpr int kTileM = 128;
constexpr int kSlots = 18;
constexpr uint32_t kSlotBytes = 8192;
constexpr int kThreads = 512;          // warp 0 producer, 1 MMA (leader) / relay (peer), 2 TMEM alloc, 4-11 epilogue, 12-15 PE
constexpr int kEpiWarps = 8;
constexpr int kPEWarps = 4;
constexpr uint32_t kColD = 0, kColA0 = 256, kColA1 = 384;

constexpr uint32_t kOffPE = 0;             // 2 x [128 x 64] bf16 PE(x) tiles
constexpr uint32_t kOffPEDir = 32768;      // 2 x [128 x 64] bf16 PE(dir) tiles
constexpr uint32_t kOffRing = 65536;
constexpr uint32_t kOffBias = kOffRing + kSlots * kSlotBytes;
constexpr uint32_t kOffBars = kOffBias + ((pk::kBiasFloats * 4 + 15) / 16) * 16;
constexpr uint32_t kNumBars = 2 * kSlots + 8;
constexpr uint32_t kOffTmemHolder = kOffBars + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemHolder + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
}  // namespace tc2

template <bool PROFILE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc2::kThreads, 1)
mlp_tc2_kernel(const __grid_constant__ CUtensorMap tmap_full, const __grid_constant__ CUtensorMap tmap_small,
               const uint8_t* __restrict__ packed, const float* __restrict__ o_rays, const float* __restrict__ d_rays,
               const float* __restrict__ ts, const float* __restrict__ samples, int64_t total, int S,
               float* __restrict__ sigma_out, float* __restrict__ rgb_out, long long* __restrict__ dbg) {
    long long prof[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs: the dynamic shared window starts at the same offset in each
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem + tc2::kOffRing;
    float* sBias = (float*)(smem + tc2::kOffBias);
    uint64_t* bars = (uint64_t*)(smem + tc2::kOffBars);
    uint64_t* full = bars;
    uint64_t* empty = bars + tc2::kSlots;
    uint64_t* dfull = bars + 2 * tc2::kSlots;
    uint64_t* edone = dfull + 2;
    uint64_t* pe_full = edone + 2;
    uint64_t* pe_empty = pe_full + 2;
    uint32_t* tmem_holder = (uint32_t*)(smem + tc2::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = umma::cluster_ctarank();          // 0 = leader (issues the MMAs), 1 = peer
    const int64_t num_tiles = (total + tc2::kTileM - 1) / tc2::kTileM;
    const int64_t num_pairs_of_tiles = (num_tiles + 1) / 2;
    const int64_t pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < tc2::kSlots; ++i) {
            umma::mbar_init(&full[i], 1);
            umma::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&dfull[i], 1);
            umma::mbar_init(&edone[i], 2 * tc2::kEpiWarps);
            umma::mbar_init(&pe_full[i], 2 * tc2::kPEWarps);
            umma::mbar_init(&pe_empty[i], 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 2) umma::tmem_alloc2(tmem_holder, 512);
    {
        const float* gb = (const float*)(packed + pk::kLayout.bias_offset);
        for (int i = tid; i < pk::kBiasFloats; i += tc2::kThreads) sBias[i] = gb[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();                                // barriers of both CTAs are initialised
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 0) {
        // ------------------------------------------------------------------ weight producer: this CTA's half of every stage
        const bool leader_lane = umma::elect_one();
        if (leader_lane) { umma::prefetch_tmap(&tmap_full); umma::prefetch_tmap(&tmap_small); }
        uint32_t cnt = 0;
        for (int64_t j = pair0; j < num_pairs_of_tiles; j += pair_stride) {
            for (int s = 0; s < pk::kStages; ++s, ++cnt) {
                const uint32_t slot = cnt % tc2::kSlots, ph = (cnt / tc2::kSlots) & 1;
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&empty[slot], ph ^ 1);
                NERF_PROF_END(tw, 4)
                if (leader_lane) {
                    const StageRef st = c_stages.s[s];
                    const uint32_t half_rows = st.bytes >> 8;                    // rows of 128 B in this CTA's half
                    if (rank == 0) umma::mbar_arrive_expect_tx(&full[slot], st.bytes);   // both halves land on the leader's barrier
                    umma::tma2_load_2d(sRing + slot * tc2::kSlotBytes, half_rows == 64 ? &tmap_full : &tmap_small, 0,
                                       (int32_t)((st.offset >> 7) + rank * half_rows), &full[slot]);
                }
                __syncwarp();
            }
        }
        if (PROFILE && lane == 0) dbg[blockIdx.x * 16 + 4] = prof[4];
    } else if (warp == 1 && rank == 0) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-uniform, one lane issues)
        const bool leader_lane = umma::elect_one();
        constexpr uint32_t kI128 = umma::make_idesc_bf16(256, 128);
        constexpr uint32_t kI16 = umma::make_idesc_bf16(256, 16);
        constexpr uint16_t kBoth = 0x3;
        uint32_t cnt = 0;
        int64_t gs = 0, e_waited = 0;
        uint32_t tile_iter = 0;
        NERF_PROF_BEGIN(t_mma_total)

        auto ensure_e = [&](int64_t k) {
            NERF_PROF_BEGIN(tw)
            while (e_waited <= k) {
                umma::mbar_wait(&edone[e_waited & 1], (uint32_t)((e_waited >> 1) & 1));
                ++e_waited;
            }
            NERF_PROF_END(tw, 2)
            umma::tc_fence_after();
        };
        auto kblock = [&](uint32_t d_col, int a_col, uint64_t a_desc, uint32_t idesc, int nk16, uint32_t& acc) {
            const uint32_t slot = cnt % tc2::kSlots, ph = (cnt / tc2::kSlots) & 1;
            NERF_PROF_BEGIN(tw)
            umma::mbar_wait(&full[slot], ph);
            NERF_PROF_END(tw, 1)
            umma::tc_fence_after();
            if (leader_lane) {
                const uint64_t bdesc = umma::make_desc_k_sw128(umma::smem_u32(sRing + slot * tc2::kSlotBytes));
#pragma unroll 4
                for (int k = 0; k < nk16; ++k) {
                    if (a_col >= 0) umma::mma2_ts(tmem + d_col, tmem + (uint32_t)a_col + 8u * k, bdesc + 2u * k, idesc, acc | (uint32_t)k);
                    else            umma::mma2_ss(tmem + d_col, a_desc + 2u * k, bdesc + 2u * k, idesc, acc | (uint32_t)k);
                }
                umma::mma2_commit(&empty[slot], kBoth);
            }
            __syncwarp();
            acc = 1;
            ++cnt;
        };
        auto step_done = [&]() {
            if (leader_lane) umma::mma2_commit(&dfull[gs & 1], kBoth);
            __syncwarp();
            ++gs;
        };
        auto layer = [&](int a_base, bool pe_first, uint64_t descPE) {
            for (int h = 0; h < 2; ++h) {
                const uint32_t d_col = tc2::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                if (pe_first) kblock(d_col, -1, descPE, kI128, 4, acc);
                kblock(d_col, a_base + 0, 0, kI128, 4, acc);
                kblock(d_col, a_base + 32, 0, kI128, 4, acc);
                if (h == 0) ensure_e(gs - 1);
                kblock(d_col, a_base + 64, 0, kI128, 4, acc);
                kblock(d_col, a_base + 96, 0, kI128, 4, acc);
                step_done();
            }
        };

        for (int64_t j = pair0; j < num_pairs_of_tiles; j += pair_stride, ++tile_iter) {
            const uint32_t pb = tile_iter & 1;
            const uint64_t descPE = umma::make_desc_k_sw128(umma::smem_u32(smem + tc2::kOffPE + pb * 16384));
            const uint64_t descPEDir = umma::make_desc_k_sw128(umma::smem_u32(smem + tc2::kOffPEDir + pb * 16384));
            NERF_PROF_BEGIN(tw)
            umma::mbar_wait(&pe_full[pb], (tile_iter >> 1) & 1);
            NERF_PROF_END(tw, 3)
            umma::tc_fence_after();
            for (int h = 0; h < 2; ++h) {                                   // mlp.0
                uint32_t acc = 0;
                ensure_e(gs - 2);
                kblock(tc2::kColD + 128u * (uint32_t)(gs & 1), -1, descPE, kI128, 4, acc);
                step_done();
            }
            layer(tc2::kColA0, false, descPE);     // mlp.2
            layer(tc2::kColA1, false, descPE);     // mlp.4
            layer(tc2::kColA0, false, descPE);     // mlp.6
            layer(tc2::kColA1, true, descPE);      // feature_fn.0 (+ PE(x))
            layer(tc2::kColA0, false, descPE);     // feature_fn.2
            layer(tc2::kColA1, false, descPE);     // feature_fn.4 -> feat in A0
            {                                      // rgb_fn.0
                const uint32_t d_col = tc2::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                kblock(d_col, -1, descPEDir, kI128, 2, acc);
                if (leader_lane) umma::mma2_commit(&pe_empty[pb], kBoth);
                __syncwarp();
                kblock(d_col, tc2::kColA0 + 0, 0, kI128, 4, acc);
                kblock(d_col, tc2::kColA0 + 32, 0, kI128, 4, acc);
                ensure_e(gs - 1);
                kblock(d_col, tc2::kColA0 + 64, 0, kI128, 4, acc);
                kblock(d_col, tc2::kColA0 + 96, 0, kI128, 4, acc);
                step_done();
            }
            {                                      // density_fn.0
                const uint32_t d_col = tc2::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                for (int kb = 0; kb < 4; ++kb) kblock(d_col, tc2::kColA0 + 32 * kb, 0, kI16, 4, acc);
                step_done();
            }
            {                                      // rgb_fn.2
                const uint32_t d_col = tc2::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                for (int kb = 0; kb < 2; ++kb) kblock(d_col, tc2::kColA1 + 32 * kb, 0, kI16, 4, acc);
                step_done();
            }
        }
        NERF_PROF_END(t_mma_total, 0)
        if (PROFILE && lane == 0) { for (int i = 0; i < 4; ++i) dbg[blockIdx.x * 16 + i] = prof[i]; dbg[blockIdx.x * 16 + 8] = tile_iter; }
    } else if (warp >= 12) {
        // ------------------------------------------------------------------ PE producers (thread = row of this CTA's tile)
        const int r = (warp - 12) * 32 + lane;
        uint32_t it = 0;
        for (int64_t j = pair0; j < num_pairs_of_tiles; j += pair_stride, ++it) {
            const uint32_t pb = it & 1;
            const int64_t row = (2 * j + rank) * tc2::kTileM + r;
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
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);                     // nerf_model.py:373
                u[0] = __fdiv_rn(dx, nrm); u[1] = __fdiv_rn(dy, nrm); u[2] = __fdiv_rn(dz, nrm);
            }
            uint32_t v[32];
            encode_row<10>(x, v);
            umma::mbar_wait(&pe_empty[pb], ((it >> 1) & 1) ^ 1);
            store_row_sw128(smem + tc2::kOffPE + pb * 16384, r, v);
            encode_row<4>(u, v);
            store_row_sw128(smem + tc2::kOffPEDir + pb * 16384, r, v);
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive_cluster(&pe_full[pb], 0);
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps, this CTA's 128 rows)
        const int q = warp & 3;
        const int wh = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int64_t gs = 0;
        NERF_PROF_BEGIN(t_epi_total)

        for (int64_t j = pair0; j < num_pairs_of_tiles; j += pair_stride) {
            const int64_t row = (2 * j + rank) * tc2::kTileM + r;
            const bool valid = row < total;
            for (int s = 0; s < 15; ++s, ++gs) {
                const int layer = s >> 1, nhalf = (s < 14) ? (s & 1) : 0;
                const bool relu = (layer != 6);
                const uint32_t a_dst = (layer & 1) ? tc2::kColA1 : tc2::kColA0;
                const float* bias = sBias + (s < 14 ? layer * 256 + nhalf * 128 : pk::kBiasR0) + wh * 64;
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                NERF_PROF_END(tw, 6)
                umma::tc_fence_after();
                const uint32_t d_addr = tmem + lane_base + tc2::kColD + 128u * (uint32_t)(gs & 1) + (uint32_t)(wh * 64);
                uint32_t v0[32], v1[32];
                umma::tmem_ld32(d_addr, v0);
                umma::tmem_ld32(d_addr + 32, v1);
                umma::tmem_wait_ld();
                uint32_t p[16];
                const uint32_t a_addr = tmem + lane_base + a_dst + (uint32_t)((nhalf * 128 + wh * 64) >> 1);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 b4 = *(const float4*)(bias + 4 * jj);
                    const float x0 = __uint_as_float(v0[4 * jj + 0]) + b4.x, x1 = __uint_as_float(v0[4 * jj + 1]) + b4.y;
                    const float x2 = __uint_as_float(v0[4 * jj + 2]) + b4.z, x3 = __uint_as_float(v0[4 * jj + 3]) + b4.w;
                    p[2 * jj + 0] = relu ? umma::pack_bf16_relu(x0, x1) : umma::pack_bf16(x0, x1);
                    p[2 * jj + 1] = relu ? umma::pack_bf16_relu(x2, x3) : umma::pack_bf16(x2, x3);
                }
                umma::tmem_st16(a_addr, p);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float4 b4 = *(const float4*)(bias + 32 + 4 * jj);
                    const float x0 = __uint_as_float(v1[4 * jj + 0]) + b4.x, x1 = __uint_as_float(v1[4 * jj + 1]) + b4.y;
                    const float x2 = __uint_as_float(v1[4 * jj + 2]) + b4.z, x3 = __uint_as_float(v1[4 * jj + 3]) + b4.w;
                    p[2 * jj + 0] = relu ? umma::pack_bf16_relu(x0, x1) : umma::pack_bf16(x0, x1);
                    p[2 * jj + 1] = relu ? umma::pack_bf16_relu(x2, x3) : umma::pack_bf16(x2, x3);
                }
                umma::tmem_st16(a_addr + 16, p);
                umma::tmem_wait_st();
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_cluster(&edone[gs & 1], 0);
            }
            {   // density_fn.0 -> sigma
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                NERF_PROF_END(tw, 6)
                umma::tc_fence_after();
                if (wh == 0) {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + tc2::kColD + 128u * (uint32_t)(gs & 1), v);
                    umma::tmem_wait_ld();
                    if (valid) sigma_out[row] = fmaxf(__uint_as_float(v[0]) + sBias[pk::kBiasSigma], 0.f);
                }
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_cluster(&edone[gs & 1], 0);
                ++gs;
            }
            {   // rgb_fn.2 -> rgb
                NERF_PROF_BEGIN(tw)
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                NERF_PROF_END(tw, 6)
                umma::tc_fence_after();
                if (wh == 0) {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + tc2::kColD + 128u * (uint32_t)(gs & 1), v);
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
                if (lane == 0) umma::mbar_arrive_cluster(&edone[gs & 1], 0);
                ++gs;
            }
        }
        NERF_PROF_END(t_epi_total, 5)
        if (PROFILE && tid == 128) { for (int i = 5; i < 8; ++i) dbg[blockIdx.x * 16 + i] = prof[i]; }
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::cluster_sync_all();          // neither CTA may exit (or free TMEM) while its peer can still touch it
    if (warp == 2) umma::tmem_dealloc2(tmem, 512);
}

// Tensor maps over the packed weight image viewed as [rows of 64 bf16 (128 B)]: box = 64 rows (half of a 128-row
// stage) or 8 rows (half of a 16-row stage).  The image is already swizzled, so the maps use SWIZZLE_NONE.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_weight_maps(const void* packed, CUtensorMap* full, CUtensorMap* small) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("nerf_mlp_forward_tc: cuTensorMapEncodeTiled not available (%s)", cudaGetErrorString(e));
            return NERF_E_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t gdim[2] = {64, pk::kLayout.weight_bytes / 128};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t estride[2] = {1, 1};
    const cuuint32_t box_full[2] = {64, 64}, box_small[2] = {64, 8};
    CUresult r1 = encode(full, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), gdim, gstride, box_full, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(small, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(packed), gdim, gstride, box_small, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
        set_error("nerf_mlp_forward_tc: cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
        return NERF_E_CUDA;
    }
    return 0;
}

int launch_mlp_tc2(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                   int64_t total, int S, float* sigma, float* rgb, void* stream, long long* dbg) {
    static thread_local bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2::kSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mlp_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc2::kSmemBytes);
        if (e != cudaSuccess) { set_error("nerf_mlp_forward_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attr_set = true;
    }
    const int64_t tiles = (total + tc2::kTileM - 1) / tc2::kTileM;
    const int64_t pairs = (tiles + 1) / 2;
    int max_pairs = num_sms() / 2;
    if (dbg) {
        const char* e = getenv("NERF_TC_MAX_CTAS");
        if (e && atoi(e) >= 2 && atoi(e) / 2 < max_pairs) max_pairs = atoi(e) / 2;
    }
    const int grid = 2 * (int)(pairs < max_pairs ? pairs : max_pairs);
    CUtensorMap tm_full, tm_small;
    if (int rc = make_weight_maps(packed, &tm_full, &tm_small)) return rc;
    if (dbg)
        mlp_tc2_kernel<true><<<grid, tc2::kThreads, tc2::kSmemBytes, (cudaStream_t)stream>>>(tm_full, tm_small, (const uint8_t*)packed, o, d,
                                                                                          ts, samples, total, S, sigma, rgb, dbg);
    else
        mlp_tc2_kernel<false><<<grid, tc2::kThreads, tc2::kSmemBytes, (cudaStream_t)stream>>>(tm_full, tm_small, (const uint8_t*)packed, o, d,
                                                                                           ts, samples, total, S, sigma, rgb, nullptr);
    return check_launch("nerf_mlp_forward_tc (cta_group::2)");
}

}  // namespace nerf
