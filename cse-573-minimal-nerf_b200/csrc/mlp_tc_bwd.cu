// mlp_tc_bwd.cu - dgrad chain of one NeRFModel (autograd of nerf_model.py:362-389) as one persistent tcgen05 kernel.
//
// Same wavefront schedule, TMEM map and warp roles as the forward kernel (mlp_tc.cu), run in reverse:
//   per 128-sample tile, given the gradients w.r.t. the head pre-activations (from composite_backward_kernel)
//     dr   = (drgb_pre . W9) * [r > 0]                         CUDA cores, "producer" warps 12-15 -> smem A tile (K = 128)
//     dz6  = dr . W8[:, :256] + dsigma_pre (x) w7              steps 0,1   (A = dr tile in shared memory, SS)
//     dz5  = (dz6 . W6) * [h5 > 0]                             steps 2,3   (A = previous dz in TMEM, TS; B = W^T stages)
//     dz4 .. dz0 likewise through feature_fn.2, feature_fn.0 (h columns), mlp.6, mlp.4, mlp.2     steps 4..13
//   every dz is written to global (bf16, tiled chunk-major like the saved activations, pack_layout.cuh) for wgrad,
//   plus a 16-wide heads block [dsigma_pre, drgb_pre, 0..] at features 1920..1935.
// ReLU masks are the 64-bit sign words the forward kernel wrote per (row, 64-feature block): 8 B instead of 128 B per step.
// mlp.0's dgrad (d PE) is not needed: the inputs carry no gradient.
#include <stdlib.h>
#include "mlp_tc_common.cuh"

namespace nerf {

namespace tb {
constexpr int kTileM = 128;
constexpr int kSlots = 4;
constexpr uint32_t kSlotBytes = 32768;
constexpr int kThreads = 512;          // warps 0,3 W^T producers, 1 MMA, 2 TMEM alloc, 4-11 epilogue, 12-15 dr producers
constexpr int kEpiWarps = 8;
constexpr int kProdWarps = 4;
constexpr int kSteps = 14;
constexpr uint32_t kColD = 0, kColA0 = 256, kColA1 = 384;

constexpr uint32_t kOffDr = 0;             // 2 buffers x 2 K-blocks x [128 x 64] bf16 (dr, K = 128)
constexpr uint32_t kOffRing = 65536;
constexpr uint32_t kOffConst = kOffRing + kSlots * kSlotBytes;               // fp32 W9 [3][128], w7 [256]
constexpr uint32_t kOffBars = kOffConst + pk::kConstFloatsT * 4;
constexpr uint32_t kNumBars = 2 * kSlots + 8;
constexpr uint32_t kOffTmemHolder = kOffBars + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemHolder + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
}  // namespace tb

// All 52 K-blocks are 16 KB and consumed in pairs: fetched as 26 requests of 32 KB (see mlp_tc_common.cuh on why).
constexpr int kMergedStagesT = pk::kStagesT / 2;

__global__ void __launch_bounds__(tb::kThreads, 1)
mlp_tc_bwd_kernel(const uint8_t* __restrict__ packed_t, const unsigned long long* __restrict__ masks,
                  const float* __restrict__ dsigma_pre, const float* __restrict__ drgb_pre, int64_t total,
                  __nv_bfloat16* __restrict__ dz_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem + tb::kOffRing;
    float* sConst = (float*)(smem + tb::kOffConst);        // [0,384) W9[c][k], [384,640) w7[k]
    uint64_t* bars = (uint64_t*)(smem + tb::kOffBars);
    uint64_t* full = bars;
    uint64_t* empty = bars + tb::kSlots;
    uint64_t* dfull = bars + 2 * tb::kSlots;
    uint64_t* edone = dfull + 2;
    uint64_t* dr_full = edone + 2;
    uint64_t* dr_empty = dr_full + 2;
    uint32_t* tmem_holder = (uint32_t*)(smem + tb::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (total + tb::kTileM - 1) / tb::kTileM;

    if (tid == 0) {
        for (int i = 0; i < tb::kSlots; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&dfull[i], 1);
            umma::mbar_init(&edone[i], tb::kEpiWarps);
            umma::mbar_init(&dr_full[i], tb::kProdWarps);
            umma::mbar_init(&dr_empty[i], 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 2) umma::tmem_alloc(tmem_holder, 512);
    {
        const float* gc = (const float*)(packed_t + pk::kLayoutT.const_offset);
        for (int i = tid; i < pk::kConstFloatsT; i += tb::kThreads) sConst[i] = gc[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 0 || warp == 3) {
        // ------------------------------------------------------------------ W^T stage producers (two warps alternate 32 KB requests)
        const bool leader = umma::elect_one();
        const uint32_t me = (warp == 0) ? 0u : 1u;
        uint32_t cnt = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int s = 0; s < kMergedStagesT; ++s, ++cnt) {
                if ((cnt & 1u) != me) continue;
                const uint32_t slot = cnt % tb::kSlots, ph = (cnt / tb::kSlots) & 1;
                umma::mbar_wait(&empty[slot], ph ^ 1);
                if (leader) {
                    umma::mbar_arrive_expect_tx(&full[slot], tb::kSlotBytes);
                    umma::bulk_g2s(sRing + slot * tb::kSlotBytes, packed_t + (size_t)s * tb::kSlotBytes, tb::kSlotBytes, &full[slot]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const bool leader = umma::elect_one();
        constexpr uint32_t kI128 = umma::make_idesc_bf16(128, 128);
        uint32_t cnt = 0, tile_iter = 0;
        int64_t gs = 0, e_waited = 0;

        auto ensure_e = [&](int64_t k) {
            while (e_waited <= k) {
                umma::mbar_wait(&edone[e_waited & 1], (uint32_t)((e_waited >> 1) & 1));
                ++e_waited;
            }
            umma::tc_fence_after();
        };
        // one 32 KB request = two K=64 blocks; A block kb from TMEM columns a_col + 32 kb, or from smem tile a_desc + kb * 16 KB
        auto stage2 = [&](uint32_t d_col, int a_col, uint32_t a_smem, uint32_t& acc) {
            const uint32_t slot = cnt % tb::kSlots, ph = (cnt / tb::kSlots) & 1;
            umma::mbar_wait(&full[slot], ph);
            umma::tc_fence_after();
            if (leader) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t bdesc = umma::make_desc_k_sw128(umma::smem_u32(sRing + slot * tb::kSlotBytes + kb * 16384));
                    const uint64_t adesc = umma::make_desc_k_sw128(a_smem + kb * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (a_col >= 0)
                            umma::mma_ts(tmem + d_col, tmem + (uint32_t)a_col + 32u * kb + 8u * k, bdesc + 2u * k, kI128, acc | (uint32_t)(kb | k));
                        else
                            umma::mma_ss(tmem + d_col, adesc + 2u * k, bdesc + 2u * k, kI128, acc | (uint32_t)(kb | k));
                    }
                }
                umma::mma_commit(&empty[slot]);
            }
            __syncwarp();
            acc = 1;
            ++cnt;
        };
        auto step_done = [&]() {
            if (leader) umma::mma_commit(&dfull[gs & 1]);
            __syncwarp();
            ++gs;
        };

        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const uint32_t pb = tile_iter & 1;
            const uint32_t dr_base = umma::smem_u32(smem + tb::kOffDr + pb * 32768);
            umma::mbar_wait(&dr_full[pb], (tile_iter >> 1) & 1);
            umma::tc_fence_after();
            for (int h = 0; h < 2; ++h) {                       // rgb_fn.0 dgrad: A = dr tile (2 K blocks)
                const uint32_t d_col = tb::kColD + 128u * (uint32_t)(gs & 1);
                uint32_t acc = 0;
                ensure_e(gs - 2);
                stage2(d_col, -1, dr_base, acc);
                if (h == 1) {                                   // last read of this tile's dr buffers
                    if (leader) umma::mma_commit(&dr_empty[pb]);
                    __syncwarp();
                }
                step_done();
            }
            for (int l = 0; l < 6; ++l) {                       // feature_fn.4, .2, .0, mlp.6, mlp.4, mlp.2
                const int a_base = (l & 1) ? tb::kColA1 : tb::kColA0;      // dz6 lives in A0, dz5 in A1, ...
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d_col = tb::kColD + 128u * (uint32_t)(gs & 1);
                    uint32_t acc = 0;
                    ensure_e(gs - 2);
                    stage2(d_col, a_base + 0, 0, acc);
                    if (h == 0) ensure_e(gs - 1);
                    stage2(d_col, a_base + 64, 0, acc);
                    step_done();
                }
            }
        }
    } else if (warp >= 12) {
        // ------------------------------------------------------------------ dr producers: thread = row
        const int r = (warp - 12) * 32 + lane;
        const float* W9 = sConst;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t pb = it & 1;
            const int64_t row = tile * tb::kTileM + r;
            const bool valid = row < total;
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            if (valid) { g0 = drgb_pre[row * 3]; g1 = drgb_pre[row * 3 + 1]; g2 = drgb_pre[row * 3 + 2]; }
            umma::mbar_wait(&dr_empty[pb], ((it >> 1) & 1) ^ 1);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                uint32_t v[32];
                // sign bits of r[row, 64kb : 64kb+64] (rgb_fn.0's ReLU output), written by the forward kernel
                const unsigned long long mb = masks[((row >> 7) * pk::kMaskWords + 28 + kb) * 128 + (row & 127)];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int k = kb * 64 + 2 * j;
                    float a = g0 * W9[k] + g1 * W9[128 + k] + g2 * W9[256 + k];            // nerf_model.py:358 backward
                    float b = g0 * W9[k + 1] + g1 * W9[128 + k + 1] + g2 * W9[256 + k + 1];
                    if (!((mb >> (2 * j)) & 1ull)) a = 0.f;
                    if (!((mb >> (2 * j + 1)) & 1ull)) b = 0.f;
                    v[j] = umma::pack_bf16(a, b);
                }
                store_row_sw128(smem + tb::kOffDr + pb * 32768 + kb * 16384, r, v);
                uint4* dst = (uint4*)(dz_out + pk::tiled_offset(row, 1792 + kb * 64, pk::kDzChunks));
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j * 128] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            {   // heads block (features 1920..1935): [dsigma_pre, drgb_pre x3, 0 ...] in bf16 for the head weight gradients
                const float dsg = valid ? dsigma_pre[row] : 0.f;
                uint4* dst = (uint4*)(dz_out + pk::tiled_offset(row, 1920, pk::kDzChunks));
                dst[0] = make_uint4(umma::pack_bf16(dsg, g0), umma::pack_bf16(g1, g2), 0u, 0u);
                dst[128] = make_uint4(0u, 0u, 0u, 0u);
            }
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&dr_full[pb]);
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int q = warp & 3;
        const int wh = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float* w7 = sConst + 384;
        int64_t gs = 0;

        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t row = tile * tb::kTileM + r;
            const bool valid = row < total;
            const float dsg = valid ? dsigma_pre[row] : 0.f;
            for (int s = 0; s < tb::kSteps; ++s, ++gs) {
                const int nhalf = s & 1;
                const int j = 6 - (s >> 1);                 // output: dz_j (gradient w.r.t. the pre-activation of layer j)
                const int col0 = nhalf * 128 + wh * 64;     // first of this warp's 64 feature columns
                // ReLU mask source: saved output of layer j (post-ReLU), except dz6 (feature_fn.4 is linear)
                unsigned long long mb = 0ull;
                if (j < 6) mb = masks[((row >> 7) * pk::kMaskWords + ((j * 256 + col0) >> 6)) * 128 + (row & 127)];
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                umma::tc_fence_after();
                const uint32_t d_addr = tmem + lane_base + tb::kColD + 128u * (uint32_t)(gs & 1) + (uint32_t)(wh * 64);
                uint32_t v0[32], v1[32];
                umma::tmem_ld32(d_addr, v0);
                umma::tmem_ld32(d_addr + 32, v1);
                umma::tmem_wait_ld();
                uint32_t p[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    const uint32_t* src = (t < 16) ? v0 : v1;
                    const int c = (t & 15) * 2;
                    float a = __uint_as_float(src[c]), b = __uint_as_float(src[c + 1]);
                    if (j == 6) {                           // + dsigma_pre (x) w7  (density head, nerf_model.py:351)
                        a = fmaf(dsg, w7[col0 + 2 * t], a);
                        b = fmaf(dsg, w7[col0 + 2 * t + 1], b);
                    } else {
                        if (!((mb >> (2 * t)) & 1ull)) a = 0.f;
                        if (!((mb >> (2 * t + 1)) & 1ull)) b = 0.f;
                    }
                    p[t] = umma::pack_bf16(a, b);
                }
                if (j > 0) {                                // dz_j is the A operand of the next dgrad GEMM
                    const uint32_t a_dst = (j & 1) ? tb::kColA1 : tb::kColA0;
                    const uint32_t a_addr = tmem + lane_base + a_dst + (uint32_t)(col0 >> 1);
                    umma::tmem_st16(a_addr, *(const uint32_t(*)[16])(p));
                    umma::tmem_st16(a_addr + 16, *(const uint32_t(*)[16])(p + 16));
                }
                {   // rows past `total` carry zeros (their dsigma / drgb are zero), so wgrad can read whole tiles
                    uint4* dst = (uint4*)(dz_out + pk::tiled_offset(row, j * 256 + col0, pk::kDzChunks));
#pragma unroll
                    for (int t = 0; t < 8; ++t) dst[t * 128] = make_uint4(p[4 * t], p[4 * t + 1], p[4 * t + 2], p[4 * t + 3]);
                }
                if (j > 0) umma::tmem_wait_st();
                umma::tc_fence_before();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&edone[gs & 1]);
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

}  // namespace nerf

namespace nerf {
int launch_mlp_tc_bwd3(const void* packed_t, const void* masks, const float* dsigma_pre, const float* drgb_pre, int64_t total,
                       void* dz_out, void* stream);
}
using namespace nerf;

extern "C" int nerf_mlp_backward_tc(const void* packed_t, const void* masks, const float* dsigma_pre, const float* drgb_pre,
                                    int64_t N, int S, void* dz_out, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_backward_tc: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(packed_t && masks && dsigma_pre && drgb_pre && dz_out, "nerf_mlp_backward_tc: null pointer");
    NERF_REQUIRE(((uintptr_t)packed_t & 127) == 0 && ((uintptr_t)masks & 7) == 0 && ((uintptr_t)dz_out & 15) == 0,
                 "nerf_mlp_backward_tc: misaligned buffer");
    // NERF_TC_ONE_TILE=1 selects the one-tile-per-CTA schedule of this file; the default is the two-tile schedule (mlp_tc_bwd3.cu)
    static const bool one_tile = [] { const char* e = getenv("NERF_TC_ONE_TILE"); return e && e[0] == '1'; }();
    if (!one_tile) return launch_mlp_tc_bwd3(packed_t, masks, dsigma_pre, drgb_pre, N * S, dz_out, stream);
    static thread_local bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb::kSmemBytes);
        if (e != cudaSuccess) { set_error("nerf_mlp_backward_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attr_set = true;
    }
    const int64_t total = N * S;
    const int64_t tiles = (total + tb::kTileM - 1) / tb::kTileM;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    mlp_tc_bwd_kernel<<<grid, tb::kThreads, tb::kSmemBytes, (cudaStream_t)stream>>>(
        (const uint8_t*)packed_t, (const unsigned long long*)masks, dsigma_pre, drgb_pre, total, (__nv_bfloat16*)dz_out);
    return check_launch("nerf_mlp_backward_tc");
}
