// mlp_tc.cu - K8: NeRFModel.forward (nerf_model.py:362-389) as one persistent, warp-specialised sm_100a
// kernel: positional encoding -> 8x256 MLP with the skip concat -> sigma / rgb heads, bf16 operands on the
// 5th-gen tensor cores (tcgen05.mma, cta_group::1, M = 128 samples per tile), fp32 accumulation in TMEM.
//
// Per CTA (one per SM, persistent over 128-sample tiles):
//   warp 0      weight producer: streams the 63 pre-swizzled [rows x 64] bf16 weight stages of the network
//               (pack_layout.cuh) through an 11-slot shared-memory ring with cp.async.bulk (TMA engine),
//               full/empty mbarriers.  The stream is ~0.92 MB per tile and stays L2-resident.
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
#include "common.cuh"
#include "pack_layout.cuh"
#include "umma.cuh"

namespace nerf {

namespace tc {
constexpr int kTileM = 128;
constexpr int kSlots = 11;
constexpr int kThreads = 384;
constexpr int kEpiThreads = 256;
constexpr uint32_t kColD = 0, kColA0 = 256, kColA1 = 384;

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t kOffPE = 0;
constexpr uint32_t kOffPEDir = 16384;
constexpr uint32_t kOffRing = 32768;
constexpr uint32_t kOffBias = kOffRing + kSlots * 16384;
constexpr uint32_t kOffBars = kOffBias + ((pk::kBiasFloats * 4 + 15) / 16) * 16;
constexpr uint32_t kNumBars = 2 * kSlots + 5;
constexpr uint32_t kOffTmemHolder = kOffBars + kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemHolder + 16 + 1024;   // + alignment slack
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr float kPiF = 3.14159274101257324f;
constexpr float kInv2Pi = 0.15915494309189535f;
constexpr float k2PiHi = 6.28318548202514648f;
constexpr float k2PiLo = -1.7484555e-7f;
}  // namespace tc

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
__constant__ StageTable c_stages = make_stage_table();

// cos / sin of a = fl32(2^i pi) * x for |a| up to a few thousand: Cody-Waite reduction by 2 pi in two FMAs,
// then the MUFU approximations on [-pi, pi] (abs error ~1e-6, far below bf16 resolution).
__device__ __forceinline__ void fast_sincos(float a, float& s, float& c) {
    const float k = rintf(a * tc::kInv2Pi);
    float r = fmaf(-k, tc::k2PiHi, a);
    r = fmaf(-k, tc::k2PiLo, r);
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
        const float f = tc::kPiF * (float)(1 << i);
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

__global__ void __launch_bounds__(tc::kThreads, 1)
mlp_tc_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ o_rays, const float* __restrict__ d_rays,
              const float* __restrict__ ts, const float* __restrict__ samples, int64_t total, int S,
              float* __restrict__ sigma_out, float* __restrict__ rgb_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sPE = smem + tc::kOffPE;
    uint8_t* sPEDir = smem + tc::kOffPEDir;
    uint8_t* sRing = smem + tc::kOffRing;
    float* sBias = (float*)(smem + tc::kOffBias);
    uint64_t* bars = (uint64_t*)(smem + tc::kOffBars);
    uint64_t* full = bars;
    uint64_t* empty = bars + tc::kSlots;
    uint64_t* dfull = bars + 2 * tc::kSlots;
    uint64_t* edone = dfull + 2;
    uint64_t* pe_ready = edone + 2;
    uint32_t* tmem_holder = (uint32_t*)(smem + tc::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (total + tc::kTileM - 1) / tc::kTileM;

    if (tid == 0) {
        for (int i = 0; i < tc::kSlots; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        umma::mbar_init(&dfull[0], 1); umma::mbar_init(&dfull[1], 1);
        umma::mbar_init(&edone[0], tc::kEpiThreads); umma::mbar_init(&edone[1], tc::kEpiThreads);
        umma::mbar_init(pe_ready, tc::kEpiThreads);
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

    if (warp == 0) {
        // ------------------------------------------------------------------ weight producer
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int s = 0; s < pk::kStages; ++s, ++cnt) {
                    const uint32_t slot = cnt % tc::kSlots, ph = (cnt / tc::kSlots) & 1;
                    umma::mbar_wait(&empty[slot], ph ^ 1);
                    const StageRef st = c_stages.s[s];
                    umma::mbar_arrive_expect_tx(&full[slot], st.bytes);
                    umma::bulk_g2s(sRing + slot * 16384, packed + st.offset, st.bytes, &full[slot]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t kI128 = umma::make_idesc_bf16(128, 128);
            constexpr uint32_t kI16 = umma::make_idesc_bf16(128, 16);
            const uint64_t descPE = umma::make_desc_k_sw128(umma::smem_u32(sPE));
            const uint64_t descPEDir = umma::make_desc_k_sw128(umma::smem_u32(sPEDir));
            uint32_t cnt = 0;         // weight stages consumed
            int64_t gs = 0;           // global step index
            int64_t e_waited = 0;     // epilogue steps already observed
            uint32_t tile_iter = 0;

            auto ensure_e = [&](int64_t k) {       // epilogue of global step k (and all before it) is done
                while (e_waited <= k) {
                    umma::mbar_wait(&edone[e_waited & 1], (uint32_t)((e_waited >> 1) & 1));
                    ++e_waited;
                }
                umma::tc_fence_after();
            };
            // one K=64 weight stage: nk16 K=16 slices; A from TMEM columns (a_col >= 0) or from a smem tile
            auto kblock = [&](uint32_t d_col, int a_col, uint64_t a_desc, uint32_t idesc, int nk16, uint32_t& acc) {
                const uint32_t slot = cnt % tc::kSlots, ph = (cnt / tc::kSlots) & 1;
                umma::mbar_wait(&full[slot], ph);
                umma::tc_fence_after();
                const uint64_t bdesc = umma::make_desc_k_sw128(umma::smem_u32(sRing + slot * 16384));
                for (int k = 0; k < nk16; ++k) {
                    if (a_col >= 0) umma::mma_ts(tmem + d_col, tmem + (uint32_t)a_col + 8u * k, bdesc + 2u * k, idesc, acc);
                    else            umma::mma_ss(tmem + d_col, a_desc + 2u * k, bdesc + 2u * k, idesc, acc);
                    acc = 1;
                }
                umma::mma_commit(&empty[slot]);
                ++cnt;
            };
            // a hidden layer read from A buffer `a_base` (K = 256), optionally preceded by a smem K block
            auto layer = [&](int a_base, bool pe_first) {
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                    uint32_t acc = 0;
                    ensure_e(gs - 2);
                    if (pe_first) kblock(d_col, -1, descPE, kI128, 4, acc);
                    kblock(d_col, a_base + 0, 0, kI128, 4, acc);
                    kblock(d_col, a_base + 32, 0, kI128, 4, acc);
                    if (h == 0) ensure_e(gs - 1);
                    kblock(d_col, a_base + 64, 0, kI128, 4, acc);
                    kblock(d_col, a_base + 96, 0, kI128, 4, acc);
                    umma::mma_commit(&dfull[gs & 1]);
                    ++gs;
                }
            };

            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
                umma::mbar_wait(pe_ready, tile_iter & 1);
                umma::tc_fence_after();
                for (int h = 0; h < 2; ++h) {                                   // mlp.0: A = PE(x) tile
                    uint32_t acc = 0;
                    ensure_e(gs - 2);
                    kblock(tc::kColD + 128u * (uint32_t)(gs & 1), -1, descPE, kI128, 4, acc);
                    umma::mma_commit(&dfull[gs & 1]);
                    ++gs;
                }
                layer(tc::kColA0, false);     // mlp.2        reads A0 (epilogue writes A1)
                layer(tc::kColA1, false);     // mlp.4        reads A1
                layer(tc::kColA0, false);     // mlp.6        reads A0
                layer(tc::kColA1, true);      // feature_fn.0 reads PE(x) + A1
                layer(tc::kColA0, false);     // feature_fn.2 reads A0
                layer(tc::kColA1, false);     // feature_fn.4 reads A1, feat -> A0
                {                             // rgb_fn.0: PE(dir) + feat (A0) -> D region, r -> A1[0:64]
                    const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                    uint32_t acc = 0;
                    ensure_e(gs - 2);
                    kblock(d_col, -1, descPEDir, kI128, 2, acc);
                    kblock(d_col, tc::kColA0 + 0, 0, kI128, 4, acc);
                    kblock(d_col, tc::kColA0 + 32, 0, kI128, 4, acc);
                    ensure_e(gs - 1);
                    kblock(d_col, tc::kColA0 + 64, 0, kI128, 4, acc);
                    kblock(d_col, tc::kColA0 + 96, 0, kI128, 4, acc);
                    umma::mma_commit(&dfull[gs & 1]);
                    ++gs;
                }
                {                             // density_fn.0: feat (A0) -> 16 columns
                    const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                    uint32_t acc = 0;
                    ensure_e(gs - 2);
                    for (int kb = 0; kb < 4; ++kb) kblock(d_col, tc::kColA0 + 32 * kb, 0, kI16, 4, acc);
                    umma::mma_commit(&dfull[gs & 1]);
                    ++gs;
                }
                {                             // rgb_fn.2: r (A1[0:64]) -> 16 columns
                    const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
                    uint32_t acc = 0;
                    ensure_e(gs - 2);
                    for (int kb = 0; kb < 2; ++kb) kblock(d_col, tc::kColA1 + 32 * kb, 0, kI16, 4, acc);
                    umma::mma_commit(&dfull[gs & 1]);
                    ++gs;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (256 threads)
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int wh = (warp - 4) >> 2;         // column half handled by this warp
        const int r = q * 32 + lane;            // row (sample) inside the tile
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int64_t gs = 0;

        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t row = tile * tc::kTileM + r;
            const bool valid = row < total;
            {   // ---- positional encodings for this tile
                uint32_t v[32];
                const int64_t n = valid ? row / S : 0;
                if (wh == 0) {
                    float x[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        if (samples) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) x[k] = samples[row * 3 + k];
                        } else {
                            const float t = ts[row];
#pragma unroll
                            for (int k = 0; k < 3; ++k)      // d * t + o (nerf_helpers.py:55)
                                x[k] = __fadd_rn(__fmul_rn(__ldg(d_rays + n * 3 + k), t), __ldg(o_rays + n * 3 + k));
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) x[k] = __fdiv_rn(x[k], tc::kPiF);     // nerf_model.py:377
                    }
                    encode_row<10>(x, v);
                    store_row_sw128(sPE, r, v);
                } else {
                    float u[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        const float dx = __ldg(d_rays + n * 3), dy = __ldg(d_rays + n * 3 + 1), dz = __ldg(d_rays + n * 3 + 2);
                        const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);              // nerf_model.py:373
                        u[0] = __fdiv_rn(dx, nrm); u[1] = __fdiv_rn(dy, nrm); u[2] = __fdiv_rn(dz, nrm);
                    }
                    encode_row<4>(u, v);
                    store_row_sw128(sPEDir, r, v);
                }
                umma::fence_proxy_async_smem();
                umma::mbar_arrive(pe_ready);
            }
            // ---- 7 hidden layers x 2 halves, then rgb_fn.0: 128 columns -> bias, (ReLU), bf16 -> A buffer
            for (int s = 0; s < 15; ++s, ++gs) {
                const int layer = s >> 1, nhalf = (s < 14) ? (s & 1) : 0;
                const bool relu = (layer != 6);                       // feature_fn.4 is linear (nerf_model.py:347)
                // writes: mlp.0 -> A0, mlp.2 -> A1, mlp.4 -> A0, mlp.6 -> A1, ff.0 -> A0, ff.2 -> A1, ff.4 -> A0, rgb_fn.0 -> A1
                const uint32_t a_dst = (layer & 1) ? tc::kColA1 : tc::kColA0;
                const float* bias = sBias + (s < 14 ? layer * 256 + nhalf * 128 : pk::kBiasR0);
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                umma::tc_fence_after();
                const uint32_t d_col = tc::kColD + 128u * (uint32_t)(gs & 1);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int col = wh * 64 + c * 32;              // column inside this 128-wide half
                    uint32_t v[32];
                    umma::tmem_ld32(tmem + lane_base + d_col + col, v);
                    umma::tmem_wait_ld();
                    uint32_t p[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b4 = *(const float4*)(bias + col + 4 * j);
                        const float x0 = __uint_as_float(v[4 * j + 0]) + b4.x, x1 = __uint_as_float(v[4 * j + 1]) + b4.y;
                        const float x2 = __uint_as_float(v[4 * j + 2]) + b4.z, x3 = __uint_as_float(v[4 * j + 3]) + b4.w;
                        p[2 * j + 0] = relu ? umma::pack_bf16_relu(x0, x1) : umma::pack_bf16(x0, x1);
                        p[2 * j + 1] = relu ? umma::pack_bf16_relu(x2, x3) : umma::pack_bf16(x2, x3);
                    }
                    // output feature n = nhalf*128 + col + j is K index n of the next layer: TMEM column n/2
                    umma::tmem_st16(tmem + lane_base + a_dst + (uint32_t)((nhalf * 128 + col) >> 1), p);
                }
                umma::tmem_wait_st();
                umma::tc_fence_before();
                umma::mbar_arrive(&edone[gs & 1]);
            }
            // ---- density_fn.0: column 0 -> sigma = relu(. + b) (nerf_model.py:350-353)
            {
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
                umma::tc_fence_after();
                if (wh == 0) {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + tc::kColD + 128u * (uint32_t)(gs & 1), v);
                    umma::tmem_wait_ld();
                    if (valid) sigma_out[row] = fmaxf(__uint_as_float(v[0]) + sBias[pk::kBiasSigma], 0.f);
                }
                umma::tc_fence_before();
                umma::mbar_arrive(&edone[gs & 1]);
                ++gs;
            }
            // ---- rgb_fn.2: columns 0..2 -> sigmoid(. + b) (nerf_model.py:358-359)
            {
                umma::mbar_wait(&dfull[gs & 1], (uint32_t)((gs >> 1) & 1));
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
                umma::mbar_arrive(&edone[gs & 1]);
                ++gs;
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

}  // namespace nerf

using namespace nerf;

static int launch_mlp_tc(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                         int64_t N, int S, float* sigma, float* rgb, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_forward_tc: bad size N=%lld S=%d", (long long)N, S);
    if (N == 0) return 0;
    NERF_REQUIRE(packed && d && sigma && rgb, "nerf_mlp_forward_tc: null pointer");
    NERF_REQUIRE(samples || (o && ts), "nerf_mlp_forward_tc: need either samples or (o, ts)");
    NERF_REQUIRE(((uintptr_t)packed & 127) == 0, "nerf_mlp_forward_tc: packed buffer must be 128-byte aligned");
    static thread_local bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes);
        if (e != cudaSuccess) { set_error("nerf_mlp_forward_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attr_set = true;
    }
    const int64_t total = N * S;
    const int64_t tiles = (total + tc::kTileM - 1) / tc::kTileM;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    mlp_tc_kernel<<<grid, tc::kThreads, tc::kSmemBytes, (cudaStream_t)stream>>>((const uint8_t*)packed, o, d, ts, samples, total, S,
                                                                              sigma, rgb);
    return check_launch("nerf_mlp_forward_tc");
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
