// mlp_tc3.cu - K8, two-tile form: NeRFModel.forward (nerf_model.py:362-389) for TWO 128-sample tiles ("X" and "Y") per
// CTA at a time, bf16 weight image streamed through shared memory, activations resident in TMEM.  The schedule:
//
//   * every weight stage fetched from L2 into the shared-memory ring is used by both tiles before it is released, so the
//     L2 -> SM weight stream per sample is halved (a one-tile schedule streams 0.92 MB per tile at ~36 B/clk/SM, close to
//     the ~42 B/clk/SM the L2 can deliver to all 148 SMs at once);
//   * the MMA -> epilogue -> MMA dependency of one tile (a layer needs the previous layer's ReLU'd output) is hidden
//     behind the other tile's MMAs: steps alternate X, Y, X, Y (turn token between the two issuing warps).
//
// TMEM (512 columns): D_X 0..127, D_Y 128..255 (one fp32 accumulator of N = 128 per tile), A_X 256..383, A_Y 384..511
// (bf16 activations, K = 256, the A operand of the next layer).  There is no room for a second A buffer per tile, so
// the epilogue converts IN PLACE: the first half of a layer's output (features 0..127) is kept in registers until the
// tile's second-half MMAs - the last readers of the old activations - have completed, then written over them.
//
// Measured facts this schedule is built on (tools/probe_tmem_contention.py, profiles/r01_notes.md): an M128 x N128 x K16
// tcgen05.mma with A in TMEM retires every ~72 clk (96 clk with A in shared memory), tcgen05.commit is free, TMEM loads
// by the epilogue do not slow the MMAs down, but the MMA queue is only ~4 instructions deep: the issuing warp must come
// back with the next group within ~290 clk or the tensor pipe idles, and the TMEM -> register path delivers ~13 B/clk per
// warp.  Hence: 32-bit shared addresses and precomputed descriptors on the issue path, one issuing warp per tile, and all
// 16 epilogue warps on every task.
//
// Warps (768 threads): 0,3 weight producers | 1 MMA issuer of tile X | 2 TMEM allocation + MMA issuer of tile Y |
// 4-19 epilogue (both tiles) | 20-23 PE producers (both tiles; in the COMP form they also composite finished rays).
// Register budgets are re-balanced with setmaxnreg.
//
// Template forms: PROFILE (cycle counters), TRAIN (also stores bf16 activations + ReLU sign words), COMP (alpha compositing
// inside the kernel: CTAs own whole ray groups, see FusedComposite below), MC (COMP launched as 2-CTA clusters that share
// every weight stage through multicast bulk copies).  The render / training paths use <false, false|true, true, true>.
//
// Steps per tile (16): mlp.0 h0,h1 | mlp.2/4/6, feature_fn.0/2/4 h0,h1 | rgb_fn.0 | rgb_fn.2 (density_fn.0: CUDA cores, see below).
// Barriers per tile t: dfull[t] (MMA -> epilogue, accumulator complete), dfree[t] (accumulator read into registers),
// alo[t] / ahi[t] (K-blocks 0,1 / 2,3 of the next A operand written), pex_full/empty[t], ped_full/empty[t], turn[t].
#include <type_traits>
#include "mlp_tc3_common.cuh"
#include "composite_scan.cuh"

namespace nerf {


// Stages of this kernel: the 63 K = 64 weight blocks of the packed image in consumption order (one ring slot each) without
// density_fn.0's four, and rgb_fn.2's two 2 KB blocks fetched as one request.  sigma is a 256-long dot product per sample;
// the epilogue of feature_fn.4 takes it on the CUDA cores from the bf16 feat values it has in registers (the operands an
// N = 16 MMA step would see, fp32 accumulation), which removes one step and its accumulator hand-over per tile.
constexpr int kStages3 = pk::kStages - 4 - 1;          // 58
struct StageTable3 { StageRef s[kStages3]; };
constexpr StageTable3 make_stage_table3() {
    StageTable3 t{};
    int m = 0;
    for (int i = 0; i < pk::kStages; ++i) {
        if (pk::kLayout.st[i].param == 7) continue;                       // density_fn.0
        if (pk::kLayout.st[i].param == 9 && pk::kLayout.st[i].k0 != 0) {  // second block of rgb_fn.2: part of the previous request
            t.s[m - 1].bytes += (uint32_t)pk::kLayout.st[i].rows * 128u;
            continue;
        }
        t.s[m].offset = pk::kLayout.st[i].offset;
        t.s[m].bytes = (uint32_t)pk::kLayout.st[i].rows * 128u;
        ++m;
    }
    return t;
}
static __constant__ StageTable3 c_stages3 = make_stage_table3();
static_assert(make_stage_table3().s[kStages3 - 1].bytes == 2u * pk::kStageBytesSmall && make_stage_table3().s[kStages3 - 2].bytes == 16384u,
              "stage table of the two-tile kernel");
constexpr uint32_t kDensityStageOffset = density_stage_offset();               // 4 blocks of [16 x 64] bf16, row 0 = w7

// 32 accumulator columns (registers) -> + bias, (ReLU), 16 registers of bf16 pairs
// RELU is a template parameter (a run-time flag makes nvcc convert both ways and select) and the bias comes through a
// shared-space address (LDS.128, not a generic load): the 16 epilogue warps are instruction-issue bound.
// `signs` (training form): one funnel shift per element collects the SIGN bits of the biased pre-activations - even elements
// (low bf16 halves) into bits 15..0, odd elements into bits 31..16, pair j at bit 15-j / 31-j - i.e. the complement of the
// ReLU mask, which the dgrad kernel applies as p & ~mask.
#ifndef NERF_INTERLEAVE_STORES
#define NERF_INTERLEAVE_STORES 1
#endif
// `early` (training form, may be null): the 16-byte chunk of 8 features is stored to the saved-activation tensor as soon as it
// is packed, so that a task's stores trickle into the memory pipe instead of arriving as one 16-warp burst at its end.
template <bool RELU, bool SIGNS>
__device__ __forceinline__ uint32_t pack32(const uint32_t (&v)[32], uint32_t bias_saddr, uint32_t* p, uint4* early = nullptr) {
    uint32_t s_lo = 0u, s_hi = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        // 4 accumulators + 4 biases as two packed fp32 pairs (add.rn.f32x2: one instruction per pair on sm_100)
        unsigned long long b01, b23, x01, x23;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b01), "=l"(b23) : "r"(bias_saddr + 16u * j));
        asm("mov.b64 %0, {%1, %2};" : "=l"(x01) : "r"(v[4 * j + 0]), "r"(v[4 * j + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(x23) : "r"(v[4 * j + 2]), "r"(v[4 * j + 3]));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x01) : "l"(b01));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x23) : "l"(b23));
        float x0, x1, x2, x3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(x23));
        if (SIGNS) {
            s_lo = __funnelshift_l(__float_as_uint(x0), s_lo, 1);
            s_hi = __funnelshift_l(__float_as_uint(x1), s_hi, 1);
            s_lo = __funnelshift_l(__float_as_uint(x2), s_lo, 1);
            s_hi = __funnelshift_l(__float_as_uint(x3), s_hi, 1);
        }
        p[2 * j + 0] = RELU ? umma::pack_bf16_relu(x0, x1) : umma::pack_bf16(x0, x1);
        p[2 * j + 1] = RELU ? umma::pack_bf16_relu(x2, x3) : umma::pack_bf16(x2, x3);
        if (NERF_INTERLEAVE_STORES && early && (j & 1))
            store_once(early + (j >> 1) * 128, make_uint4(p[2 * j - 2], p[2 * j - 1], p[2 * j], p[2 * j + 1]));
    }
    return (s_hi << 16) | s_lo;
}

// Row `r` of a [128 x 64] bf16 K-major 128B-swizzled PE tile, four frequencies (12 registers = three 16-byte chunks) at a
// time so that the producer warps stay within a small register budget.  Layout per frequency as encode_row.
template <int L>
__device__ __forceinline__ void encode_store_row(uint8_t* tile, int r, const float (&x)[3]) {
    uint8_t* rowp = tile + r * 128;
    const int sw = r & 7;
#pragma unroll
    for (int g = 0; g < 8; g += 3) {              // chunk groups {0,1,2}, {3,4,5}, {6,7}
        uint32_t v[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) v[j] = 0u;
#pragma unroll
        for (int fi = 0; fi < 4; ++fi) {
            const int i = (g / 3) * 4 + fi;
            if (i < L) {
                const float f = tcm::kPiF * (float)(1 << i);
                float s[3], c[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) fast_sincos(__fmul_rn(f, x[k]), s[k], c[k]);
                v[3 * fi + 0] = umma::pack_bf16(c[0], c[1]);
                v[3 * fi + 1] = umma::pack_bf16(c[2], s[0]);
                v[3 * fi + 2] = umma::pack_bf16(s[1], s[2]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            if (g + c < 8) *(uint4*)(rowp + (((g + c) ^ sw) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
}

// sum of 64 bf16 values (32 packed registers) times 64 fp32 weights, fp32 accumulation
__device__ __forceinline__ float dot_bf16x64(const uint32_t (&p)[32], const float* __restrict__ w, float acc) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float4 w4 = *(const float4*)(w + 4 * j);
        acc = fmaf(__uint_as_float(p[2 * j] << 16), w4.x, acc);
        acc = fmaf(__uint_as_float(p[2 * j] & 0xFFFF0000u), w4.y, acc);
        acc = fmaf(__uint_as_float(p[2 * j + 1] << 16), w4.z, acc);
        acc = fmaf(__uint_as_float(p[2 * j + 1] & 0xFFFF0000u), w4.w, acc);
    }
    return acc;
}

__device__ __forceinline__ float dot_bf16x32(const uint32_t (&p)[16], const float* __restrict__ w, float acc) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 w4 = *(const float4*)(w + 4 * j);
        acc = fmaf(__uint_as_float(p[2 * j] << 16), w4.x, acc);
        acc = fmaf(__uint_as_float(p[2 * j] & 0xFFFF0000u), w4.y, acc);
        acc = fmaf(__uint_as_float(p[2 * j + 1] << 16), w4.z, acc);
        acc = fmaf(__uint_as_float(p[2 * j + 1] & 0xFFFF0000u), w4.w, acc);
    }
    return acc;
}

// training, 32 features of one row: bf16 activations (four 16-byte chunks, tiled chunk-major) + the 32-bit word of sign bits
// pack32 collected (two-tile kernels only; pair j = features (2j, 2j+1) of the group -> bit 15-j / 31-j, 1 = pre-activation
// negative = ReLU inactive).  `act` / `mask` point at this row's first feature of the group / its sign word: the callers keep
// ONE running pointer pair per thread and reach both halves of a layer and both tiles through immediate offsets (the epilogue
// warps are instruction-issue bound in the training form; a 64-bit address computation per store group was ~10 % of a task).
__device__ __forceinline__ void save_act32(__nv_bfloat16* __restrict__ act, uint32_t* __restrict__ mask,
                                           const uint32_t (&p)[16], uint32_t signs) {
    if (!NERF_INTERLEAVE_STORES) {         // (otherwise pack32 has stored the four chunks already)
        uint4* chunk = (uint4*)act;
#pragma unroll
        for (int j = 0; j < 4; ++j) store_once(chunk + j * 128, make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]));
    }
    store_once(mask, signs);
}

// dbg counters (PROFILE), per CTA x 16: 0 MMA warp total, 1 wait weights, 2 wait dfree, 3 wait alo/ahi, 4 wait PE,
// 5 producer wait empty, 6 epilogue X total, 7 epilogue X wait dfull, 8 pairs; dbg[148*16 ...] = CTA 0's wait detail
// COMP form (fused compositing, nerf_helpers.py:58-104 inside the MLP kernel): every CTA owns a CONTIGUOUS range of ray
// groups - a group is the smallest run of whole rays that is also a whole number of 128-sample tiles (S = 64: 2 rays = 1
// tile, S = 192: 2 rays = 3 tiles) - so every ray's samples are produced by one CTA.  The last step's epilogue warps put
// (sigma, r, g, b) of each sample into a shared-memory ring of tiles instead of (render) / as well as (training) global
// memory, and the four PE warps - which run a pair ahead of the MMAs and have issue slots to spare - turn finished groups into
// weights, ray colour, depth and opacity between two encodings, with the same warp-scan code as the stand-alone compositing
// kernel (composite_scan.cuh): one ray per warp, lane = sample, overlapped with the MMAs / epilogues of the following tiles.
// (A dedicated fifth warpgroup does not fit the register file: 896 threads launch with 72 registers, and the pool then
// leaves the MMA-issuing warps 48-56 registers - their issue loop starts reloading descriptors from local memory.)
struct FusedComposite {
    float* weights;      // [N,S] or null
    float* ray_rgb;      // [N,3]
    float* depth;        // [N] or null
    float* acc;          // [N] or null
    float* stats;        // [4] or null: sum sigma^2, count sigma != 0, sqrt(sum sigma^2) (published by the last CTA), ticket
    int group_rays, group_tiles;
    int64_t num_groups, N;
    // K1 inside the kernel (coarse network): when u_c is given the PE warps form the stratified depths themselves,
    // t = t_base[i] + u_c[n, i] * step (nerf_helpers.py:52-53, the arithmetic of coarse_sample_kernel), store them to ts_gen
    // [N,S] for the fine sampler, and the compositing reads them back from there
    const float* u_c;
    const float* t_base;
    float step;
    float* ts_gen;
};

__device__ __forceinline__ float ld_coherent(const float* p) {      // depths written earlier by another warp of this CTA
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// COMP form, run by each of the four PE warps: composite every ray group whose local tiles lie below `limit`, starting at
// local tile `next_lt` (advanced; `last` = number of local tiles that exist).  A group's rays are dealt round-robin to the warps; all four warps wait for the group's
// tiles and release them.  Same arithmetic, in the same order, as composite_kernel<1> (samplers.cu).
__device__ __forceinline__ void composite_groups(const FusedComposite& fc, const float* __restrict__ ts, int S, int64_t g0, uint8_t* smem,
                                              uint32_t bars, uint32_t& next_lt, uint32_t limit, uint32_t last, int w, int lane, float* stat) {
    const float4* sOut = (const float4*)(smem + t3::kOffOut);
    const uint32_t TG = (uint32_t)fc.group_tiles;
    while (next_lt + TG <= limit) {
        const uint32_t lt0 = next_lt;
        for (uint32_t j = 0; j < TG; ++j) {
            const uint32_t lt = lt0 + j;
            if (lt < last) umma::mbar_wait_u32(bars + 8u * (t3::kBarOutFull + (lt & (t3::kOutSlots - 1))), (lt >> 3) & 1u);
        }
        const int64_t n_first = (g0 + lt0 / TG) * fc.group_rays;
        for (int j = 0; j < fc.group_rays; ++j) {
            const int64_t n = n_first + j;
            if (((int)n & 3) != w || n >= fc.N) continue;                               // ray n -> warp n % 4
            const float* tp = (fc.u_c ? fc.ts_gen : ts) + n * S;
            float* wp = fc.weights ? fc.weights + n * S : nullptr;
            const uint32_t grow0 = lt0 * t3::kTileM + (uint32_t)(j * S);                // CTA-local row of the ray's first sample
            float running = 0.f;       // sum_{j<i} -sigma_j delta_j, nerf_helpers.py:86-89
            float cr = 0.f, cg = 0.f, cb = 0.f, dsum = 0.f, asum = 0.f, st_sq = 0.f, st_nz = 0.f;
            float t_n = ld_coherent(tp + lane);                                          // S is a multiple of 32
            for (int base = 0; base < S; base += kWarp) {
                const int i = base + lane;
                const float t = t_n;
                if (base + kWarp < S) t_n = ld_coherent(tp + base + kWarp + lane);       // next chunk's depths in flight
                const uint32_t grow = grow0 + (uint32_t)i;
                const float4 v = sOut[((grow >> 7) & (t3::kOutSlots - 1)) * t3::kTileM + (grow & 127u)];
                const float s = v.x;
                float tn = __shfl_down_sync(kFull, t, 1);
                const float t_first_next = __shfl_sync(kFull, t_n, 0);
                if (lane == 31 && i + 1 < S) tn = t_first_next;
                const float dl = (i == S - 1) ? 1e10f : __fsub_rn(tn, t);               // nerf_helpers.py:71-72
                const float x = __fmul_rn(__fmul_rn(-1.0f, s), dl);                      // -1 * density * deltas
                const float excl = chunk_exclusive_scan_tree(x, running, lane);
                const float trans = expf(excl);                                          // nerf_helpers.py:89
                const float wgt = __fmul_rn(__fsub_rn(1.0f, expf(x)), trans);            // nerf_helpers.py:90
                if (wp) wp[i] = wgt;
                cr = fmaf(wgt, v.y, cr); cg = fmaf(wgt, v.z, cg); cb = fmaf(wgt, v.w, cb);   // nerf_helpers.py:103
                dsum = fmaf(wgt, t, dsum); asum += wgt;
                st_sq = fmaf(s, s, st_sq); st_nz += (s != 0.f) ? 1.f : 0.f;
            }
            cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
            dsum = warp_sum(dsum); asum = warp_sum(asum);
            st_sq = warp_sum(st_sq); st_nz = warp_sum(st_nz);
            if (lane == 0) {
                fc.ray_rgb[n * 3 + 0] = cr; fc.ray_rgb[n * 3 + 1] = cg; fc.ray_rgb[n * 3 + 2] = cb;
                if (fc.depth) fc.depth[n] = dsum;
                if (fc.acc) fc.acc[n] = asum;
                stat[0] += st_sq; stat[1] += st_nz;
            }
        }
        __syncwarp();
        if (lane == 0)
            for (uint32_t j = 0; j < TG; ++j) {
                const uint32_t lt = lt0 + j;
                if (lt < last) umma::mbar_arrive_u32(bars + 8u * (t3::kBarOutEmpty + (lt & (t3::kOutSlots - 1))));
            }
        next_lt = lt0 + TG;
    }
}

// MC form (COMP only; launched as 2-CTA clusters): the two CTAs of a cluster share every weight stage - each fetches HALF of it
// from L2 and multicasts that half into both shared memories (`cp.async.bulk ... .multicast::cluster`), which halves the
// L2 -> SM weight stream once more (0.46 MB per tile pair and CTA).  Nothing else is shared: MMAs, TMEM and epilogues stay
// per CTA; the only cross-CTA dependency is the recycling of a ring slot, whose release commits go to both CTAs' `empty`
// barriers (count 4) - eight slots deep, off the MMA -> epilogue -> MMA loop (a cta_group::2 CTA-pair schedule put its
// cross-CTA signalling on that loop and measured slower, profiles/r01_notes.md).
// Both CTAs run the same number of pairs (the one with fewer tiles ends on a fully masked pair).
template <bool PROFILE, bool TRAIN, bool COMP, bool MC>
__global__ void __launch_bounds__(t3::kThreads, 1)
mlp_tc3_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ o_rays, const float* __restrict__ d_rays,
               const float* __restrict__ ts, const float* __restrict__ samples, int64_t total, int S,
               float* __restrict__ sigma_out, float* __restrict__ rgb_out, __nv_bfloat16* __restrict__ act_out,
               unsigned long long* __restrict__ mask_out, long long* __restrict__ dbg, const FusedComposite fc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t bars = sbase + t3::kOffBars;
    float* sBias = (float*)(smem + t3::kOffBias);
    float* sW7 = (float*)(smem + t3::kOffW7);
    float* sSig = (float*)(smem + t3::kOffSig);
    float* sStat = (float*)(smem + t3::kOffStat);
    uint32_t* tmem_holder = (uint32_t*)(smem + t3::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (total + t3::kTileM - 1) / t3::kTileM;
    // This CTA's tile pairs: pair k = tiles (tile0 + k * tile_stride, + 1), k < n_local; tiles >= tile_end
    // are not this CTA's.  Round-robin over all pairs, or (COMP) the contiguous tiles of ray groups [g0, g1).
    // (tile indices are 32-bit: 2^32 tiles would be 5 x 10^11 samples)
    uint32_t tile0, tile_stride, tile_end, n_local;
    int64_t g0 = 0;
    const uint32_t rank = MC ? (blockIdx.x & 1u) : 0u;          // CTA rank in its 2-CTA cluster (1-D grid: blocks 2c, 2c + 1)
    if (COMP && MC) {
        // the cluster's groups [G0, G1) are cut in the middle; both CTAs run max(own, peer) pairs
        const int64_t nc = gridDim.x >> 1, c = blockIdx.x >> 1;
        const int64_t G0 = fc.num_groups * c / nc, G1 = fc.num_groups * (c + 1) / nc, mid = G0 + (G1 - G0 + 1) / 2;
        auto tiles_of = [&](int64_t a, int64_t b) -> int64_t {
            int64_t e = b * fc.group_tiles < num_tiles ? b * fc.group_tiles : num_tiles;
            return e > a * fc.group_tiles ? e - a * fc.group_tiles : 0;
        };
        g0 = rank ? mid : G0;
        const int64_t own = tiles_of(rank ? mid : G0, rank ? G1 : mid), peer = tiles_of(rank ? G0 : mid, rank ? mid : G1);
        tile0 = (uint32_t)(g0 * fc.group_tiles);
        tile_end = tile0 + (uint32_t)own;
        tile_stride = 2;
        n_local = (uint32_t)(((own > peer ? own : peer) + 1) / 2);
    } else if (COMP) {
        g0 = fc.num_groups * blockIdx.x / gridDim.x;
        const int64_t g1 = fc.num_groups * (blockIdx.x + 1) / gridDim.x;
        tile0 = (uint32_t)(g0 * fc.group_tiles);
        tile_end = (uint32_t)(g1 * fc.group_tiles < num_tiles ? g1 * fc.group_tiles : num_tiles);
        tile_stride = 2;
        n_local = tile_end > tile0 ? (tile_end - tile0 + 1u) / 2u : 0u;
    } else {
        const int64_t num_pairs = (num_tiles + 1) / 2;
        tile0 = 2u * blockIdx.x;
        tile_stride = 2u * gridDim.x;
        tile_end = (uint32_t)num_tiles;
        n_local = (uint32_t)((num_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x);
    }

    if (tid == 0) {
        uint64_t* b = (uint64_t*)(smem + t3::kOffBars);
        for (int i = 0; i < t3::kOutSlots; ++i) { umma::mbar_init(&b[t3::kBarOutFull + i], 4); umma::mbar_init(&b[t3::kBarOutEmpty + i], t3::kPEWarps); }
        for (int i = 0; i < t3::kSlots; ++i) { umma::mbar_init(&b[t3::kBarFull + i], 1); umma::mbar_init(&b[t3::kBarEmpty + i], MC ? 4 : 2); }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&b[t3::kBarDFull + i], 1);
            umma::mbar_init(&b[t3::kBarDFree + i], t3::kEpiWarps);
            umma::mbar_init(&b[t3::kBarALo + i], t3::kEpiWarps);
            umma::mbar_init(&b[t3::kBarAHi + i], t3::kEpiWarps);
            umma::mbar_init(&b[t3::kBarPexFull + i], t3::kPEWarps);
            umma::mbar_init(&b[t3::kBarPexEmpty + i], 1);
            umma::mbar_init(&b[t3::kBarPedFull + i], t3::kPEWarps);
            umma::mbar_init(&b[t3::kBarPedEmpty + i], 1);
            umma::mbar_init(&b[t3::kBarTurn + i], 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 2) umma::tmem_alloc(tmem_holder, 512);
    {   // biases: resident for the whole kernel
        const float* gb = (const float*)(packed + pk::kLayout.bias_offset);
        for (int i = tid; i < pk::kBiasFloats; i += (int)blockDim.x) sBias[i] = gb[i];
        // density_fn.0.weight as the packer rounded it to bf16: row 0 of the four [16 x 64] blocks (row 0 is not swizzled)
        if (tid < 256) {
            const __nv_bfloat16* w7 = (const __nv_bfloat16*)(packed + kDensityStageOffset + (tid >> 6) * pk::kStageBytesSmall);
            sW7[tid] = __bfloat162float(w7[tid & 63]);
        }
        if (PROFILE && tid < 96) ((long long*)(smem + t3::kOffDetail))[tid] = 0;
    }
    umma::tc_fence_before();
    __syncthreads();
    if (MC) umma::cluster_sync_all();          // the peer's barriers exist before anything is multicast at them
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp < 4) {
        reg_dec<t3::kRegsMisc>();
        if (warp == 0 || warp == 3) {
            // -------------------------------------------------------------- weight producers (alternate stages)
            const bool leader = umma::elect_one();
            const uint32_t me = (warp == 0) ? 0u : 1u;
            const uint32_t ring = sbase + t3::kOffRing;
            long long t_wait = 0;
            uint32_t cnt = 0;
            for (uint32_t k = 0; k < n_local; ++k) {
                for (int s = 0; s < kStages3; ++s, ++cnt) {
                    if ((cnt & 1u) != me) continue;
                    const uint32_t slot = cnt & (t3::kSlots - 1), ph = (cnt >> 3) & 1u;
                    const long long t0 = PROFILE ? clock64() : 0;
                    umma::mbar_wait_u32(bars + 8u * (t3::kBarEmpty + slot), ph ^ 1u);
                    if (PROFILE) t_wait += clock64() - t0;
                    if (leader) {
                        const StageRef st = c_stages3.s[s];
                        umma::mbar_arrive_expect_tx_u32(bars + 8u * (t3::kBarFull + slot), st.bytes);
                        if (MC) {          // this CTA's half of the stage, into both CTAs' slots
                            const uint32_t half = st.bytes >> 1;
                            umma::bulk_g2s_mc_u32(ring + slot * t3::kSlotBytes + rank * half, packed + st.offset + rank * half, half,
                                                  bars + 8u * (t3::kBarFull + slot), (uint16_t)3);
                        } else {
                            umma::bulk_g2s_u32(ring + slot * t3::kSlotBytes, packed + st.offset, st.bytes, bars + 8u * (t3::kBarFull + slot));
                        }
                    }
                    __syncwarp();
                }
            }
            if (PROFILE && lane == 0 && warp == 0) dbg[blockIdx.x * 16 + 5] = t_wait;
        } else {
            // -------------------------------------------------------------- MMA issuers: warp 1 tile X, warp 2 tile Y
            const bool elected = umma::elect_one();
            const long long t_begin = PROFILE ? clock64() : 0;
            if (warp == 1) {
                MmaTile<0, PROFILE, MC> m;
                m.init(bars, sbase + t3::kOffRing, tmem, elected);
                if (PROFILE) m.detail = sbase + t3::kOffDetail;
                m.run(sbase, n_local);
                __syncwarp();
                if (PROFILE && blockIdx.x == 0) for (int i = lane; i < 64; i += 32) dbg[148 * 16 + i] = ((long long*)(smem + t3::kOffDetail))[i];
                if (PROFILE && elected) {
                    dbg[blockIdx.x * 16 + 0] = clock64() - t_begin;
                    for (int i = 1; i < 5; ++i) dbg[blockIdx.x * 16 + i] = m.prof[i];
                    dbg[blockIdx.x * 16 + 9] = m.prof[0];
                    dbg[blockIdx.x * 16 + 8] = n_local;
                }
            } else {
                MmaTile<1, PROFILE, MC> m;
                m.init(bars, sbase + t3::kOffRing, tmem, elected);
                m.run(sbase, n_local);
            }
        }
    } else if (warp >= 20) {
        // ------------------------------------------------------------------ PE producers (128 threads, thread = row)
        reg_dec<t3::kRegsPE>();
        const int r = (warp - 20) * 32 + lane;
        const uint32_t ntl = tile_end > tile0 ? (uint32_t)(tile_end - tile0) : 0u;      // COMP: this CTA's tiles
        uint32_t comp_lt = 0;                                                            // COMP: first tile of the next group to composite
        if (COMP && lane == 0) { sStat[2 * (warp - 20)] = 0.f; sStat[2 * (warp - 20) + 1] = 0.f; }
        for (uint32_t it = 0; it <= n_local; ++it) {
            const uint32_t tile_x = tile0 + it * tile_stride;
            // COMP: groups whose tiles all belong to pairs <= it - 2 are complete (pair it - 1 is in flight); after the last
            // pair, everything that is left
            if (COMP)
                composite_groups(fc, ts, S, g0, smem, bars, comp_lt,
                                 it == n_local ? ntl + (uint32_t)fc.group_tiles - 1u : (it >= 2u ? min(2u * (it - 1u), ntl) : 0u), ntl,
                                 warp - 20, lane, sStat + 2 * (warp - 20));
            if (it == n_local) break;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int64_t row = (int64_t)(tile_x + t) * t3::kTileM + r;
                float x[3] = {0.f, 0.f, 0.f};
                if (tile_x + t < tile_end && row < total) {
                    if (samples) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) x[k] = samples[row * 3 + k];
                    } else {
                        const int64_t n = row / S;
                        float tt;
                        if (COMP && fc.u_c) {                         // stratified depth formed here (nerf_helpers.py:52-53)
                            tt = __fadd_rn(__ldg(fc.t_base + (int)(row - n * S)), __fmul_rn(__ldg(fc.u_c + row), fc.step));
                            fc.ts_gen[row] = tt;
                        } else {
                            tt = ts[row];
                        }
                        // d * t + o (nerf_helpers.py:55)
#pragma unroll
                        for (int k = 0; k < 3; ++k) x[k] = __fadd_rn(__fmul_rn(__ldg(d_rays + n * 3 + k), tt), __ldg(o_rays + n * 3 + k));
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) x[k] = __fdiv_rn(x[k], tcm::kPiF);           // nerf_model.py:377
                }
                umma::mbar_wait_u32(bars + 8u * (t3::kBarPexEmpty + t), (it & 1u) ^ 1u);   // feature_fn.0 of the previous pair is done with it
                encode_store_row<10>(smem + t3::kOffPE + t * 16384, r, x);
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_u32(bars + 8u * (t3::kBarPexFull + t));
            }
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int64_t row = (int64_t)(tile_x + t) * t3::kTileM + r;
                float u[3] = {0.f, 0.f, 0.f};
                if (tile_x + t < tile_end && row < total) {
                    const int64_t n = row / S;
                    const float dx = __ldg(d_rays + n * 3), dy = __ldg(d_rays + n * 3 + 1), dz = __ldg(d_rays + n * 3 + 2);
                    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);                    // nerf_model.py:373
                    u[0] = __fdiv_rn(dx, nrm); u[1] = __fdiv_rn(dy, nrm); u[2] = __fdiv_rn(dz, nrm);
                }
                umma::mbar_wait_u32(bars + 8u * (t3::kBarPedEmpty + t), (it & 1u) ^ 1u);   // rgb_fn.0 of the previous pair is done with it
                encode_store_row<4>(smem + t3::kOffPEDir + t * 16384, r, u);
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_u32(bars + 8u * (t3::kBarPedFull + t));
            }
        }
        if (COMP) {
            if (fc.stats && lane == 0) {
                atomicAdd(fc.stats + 0, sStat[2 * (warp - 20)]);
                atomicAdd(fc.stats + 1, sStat[2 * (warp - 20) + 1]);
                // the last warp to arrive publishes the norm the reference logs (nerf_model.py:105,124): stats[2] = sqrt(stats[0])
                __threadfence();
                const unsigned ticket = atomicAdd((unsigned*)(fc.stats + 3), 1u);
                if (ticket == gridDim.x * t3::kPEWarps - 1) {
                    __threadfence();
                    fc.stats[2] = sqrtf(atomicAdd(fc.stats + 0, 0.f));
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: all 16 warps work on every task
        // (tile t, step): the TMEM -> register path delivers ~13 B/clk per warp whatever else is going on, so a task's
        // latency is set by how many warps share it.  Thread = (row r, 32 of the 128 accumulator columns).
        reg_inc<t3::kRegsEpi>();
        const int q = warp & 3;                          // TMEM lane quarter this warp may touch
        const int cq = (warp - 4) >> 2;                  // column quarter (32 accumulator columns, 32 features)
        const int r = q * 32 + lane;                     // row (sample) inside the tile
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t nd[2] = {0, 0};                         // accumulators consumed per tile
        long long t_wait = 0;
        long long tp[4] = {0, 0, 0, 0};
        const long long t_begin = PROFILE ? clock64() : 0;
        auto wait_d = [&](int t) {
            const long long t0 = PROFILE ? clock64() : 0;
            umma::mbar_wait_u32(bars + 8u * (t3::kBarDFull + t), nd[t] & 1u);
            if (PROFILE) t_wait += clock64() - t0;
            umma::tc_fence_after();
            ++nd[t];
        };

        for (uint32_t it = 0; it < n_local; ++it) {
            const uint32_t tile_x = tile0 + it * tile_stride;
            const int64_t row0 = (int64_t)tile_x * t3::kTileM + r;   // this thread's row in tile X (tile Y: + 128)
            const bool save0 = TRAIN && (tile_x < tile_end), save1 = TRAIN && (tile_x + 1 < tile_end);
            // this thread's rows in the tiled chunk-major training tensors (pack_layout.cuh): feature 0 / sign-word group 0
            // running pointers: this thread's 32-feature group of the current layer in tile X; tile Y and the layer's second
            // half (features + 128) sit at constant offsets, a layer advances them by 256 features
            constexpr int kTileActStride = pk::kActChunks * 1024;            // elements between the same row of tiles X and Y
            constexpr int kTileMaskStride = 2 * pk::kMaskWords * 128;        // 32-bit words
            constexpr int kHalfActStride = (128 / 8) * 1024, kHalfMaskStride = (128 / 32) * 128;
            __nv_bfloat16* ap = act_out + pk::tiled_offset(row0, 0, pk::kActChunks) + (size_t)(cq * 4) * 1024;
            uint32_t* mp = (uint32_t*)mask_out + ((row0 >> 7) * (2 * pk::kMaskWords)) * 128 + (row0 & 127) + cq * 128;
            float sig_part[2] = {0.f, 0.f};
            // one hidden layer (both halves, both tiles); LAST = feature_fn.4: linear (nerf_model.py:347) and feeds density_fn.0
            auto hidden_layer = [&](auto last_tag, int layer) {
                constexpr bool LAST = decltype(last_tag)::value;
                const uint32_t bias_s = sbase + t3::kOffBias + 4u * (uint32_t)(layer * 256 + cq * 32);
                uint32_t hold[2][16];                                 // first-half outputs of X and Y, kept until the tile's second half
#pragma unroll
                for (int t = 0; t < 2; ++t) {                         // ---- first halves: X then Y
                    const uint32_t d_addr = tmem + lane_base + t3::kColD + 128u * (uint32_t)t + (uint32_t)(cq * 32);
                    wait_d(t);
                    const long long t_h0 = PROFILE ? clock64() : 0;
                    uint32_t v[32];
                    umma::tmem_ld32(d_addr, v);
                    umma::tmem_wait_ld();
                    warp_arrive(bars + 8u * (t3::kBarDFree + t), lane);
                    const uint32_t sg0 = pack32<!LAST, TRAIN>(v, bias_s, hold[t], (t == 0 ? save0 : save1) ? (uint4*)(ap + t * kTileActStride) : nullptr);
                    if (PROFILE && t == 0) { asm volatile("" ::"r"(hold[0][0]), "r"(hold[0][15])); tp[3] += clock64() - t_h0; }
                    if (t == 0 ? save0 : save1) save_act32(ap + t * kTileActStride, mp + t * kTileMaskStride, hold[t], sg0);
                    if (LAST) sig_part[t] = dot_bf16x32(hold[t], sW7 + cq * 32, sig_part[t]);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {                         // ---- second halves: the held half goes in place first
                    const uint32_t d_addr = tmem + lane_base + t3::kColD + 128u * (uint32_t)t + (uint32_t)(cq * 32);
                    const uint32_t a_addr = tmem + lane_base + t3::kColA + 128u * (uint32_t)t + (uint32_t)(cq * 16);
                    wait_d(t);
                    const long long t_h1 = PROFILE ? clock64() : 0;
                    uint32_t v[32];
                    umma::tmem_ld32(d_addr, v);
                    umma::tmem_st16(a_addr, hold[t]);                 // features 0..127 -> A columns 0..63
                    umma::tmem_wait_st();
                    if (PROFILE && t == 0) tp[0] += clock64() - t_h1;
                    umma::tmem_wait_ld();
                    warp_arrive(bars + 8u * (t3::kBarDFree + t), lane);       // = accumulator free AND K blocks 0,1 of the new operand written
                    if (PROFILE && t == 0) { asm volatile("" ::"r"(v[31])); tp[1] += clock64() - t_h1; }
                    uint32_t p[16];
                    const uint32_t sg1 = pack32<!LAST, TRAIN>(v, bias_s + 512u, p, (t == 0 ? save0 : save1) ? (uint4*)(ap + t * kTileActStride + kHalfActStride) : nullptr);
                    umma::tmem_st16(a_addr + 64, p);                  // features 128..255 -> A columns 64..127
                    umma::tmem_wait_st();
                    warp_arrive(bars + 8u * (t3::kBarAHi + t), lane);
                    if (PROFILE && t == 0) tp[2] += clock64() - t_h1;
                    if (t == 0 ? save0 : save1) save_act32(ap + t * kTileActStride + kHalfActStride, mp + t * kTileMaskStride + kHalfMaskStride, p, sg1);
                    if (LAST) sig_part[t] = dot_bf16x32(p, sW7 + 128 + cq * 32, sig_part[t]);
                }
                if (TRAIN) { ap += 2 * kHalfActStride; mp += 2 * kHalfMaskStride; }       // next layer: + 256 features
            };
#pragma unroll 1
            for (int layer = 0; layer < 6; ++layer) hidden_layer(std::false_type{}, layer);
            hidden_layer(std::true_type{}, 6);
            // density_fn.0 (nerf_model.py:350-353): this warp's 64 of the 256 products; the cq = 0 warp of the same rows adds
            // the four partial sums at the last step (ordered behind this write by the alo arrive below and the MMA commit)
#pragma unroll
            for (int t = 0; t < 2; ++t) sSig[(t * 4 + cq) * 128 + r] = sig_part[t];
            // ---- rgb_fn.0: r = relu(. + b) overwrites feat (every reader of feat has completed)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const uint32_t d_addr = tmem + lane_base + t3::kColD + 128u * (uint32_t)t + (uint32_t)(cq * 32);
                const uint32_t a_addr = tmem + lane_base + t3::kColA + 128u * (uint32_t)t + (uint32_t)(cq * 16);
                wait_d(t);
                uint32_t v[32], p[16];
                umma::tmem_ld32(d_addr, v);
                umma::tmem_wait_ld();
                warp_arrive(bars + 8u * (t3::kBarDFree + t), lane);
                const uint32_t sgr = pack32<true, TRAIN>(v, sbase + t3::kOffBias + 4u * (uint32_t)(pk::kBiasR0 + cq * 32), p,
                                                         (t == 0 ? save0 : save1) ? (uint4*)(ap + t * kTileActStride) : nullptr);
                umma::tmem_st16(a_addr, p);                           // r features 0..127 -> A columns 0..63
                umma::tmem_wait_st();
                warp_arrive(bars + 8u * (t3::kBarALo + t), lane);
                if (t == 0 ? save0 : save1) save_act32(ap + t * kTileActStride, mp + t * kTileMaskStride, p, sgr);   // feature 1792 + cq * 32
            }
            // ---- rgb_fn.2: columns 0..2 -> sigmoid(. + b) (nerf_model.py:358-359); sigma = relu(feat . w7 + b7)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int64_t row = row0 + t * 128;
                wait_d(t);
                uint32_t v[4] = {0u, 0u, 0u, 0u};
                float sg = 0.f;
                if (cq == 0) {
                    umma::tmem_ld4(tmem + lane_base + t3::kColD + 128u * (uint32_t)t, v);
                    sg = (sSig[(t * 4 + 0) * 128 + r] + sSig[(t * 4 + 1) * 128 + r]) + (sSig[(t * 4 + 2) * 128 + r] + sSig[(t * 4 + 3) * 128 + r]);
                    umma::tmem_wait_ld();
                }
                warp_arrive(bars + 8u * (t3::kBarDFree + t), lane);
                if (COMP) {
                    const uint32_t lt = 2u * it + (uint32_t)t;                   // CTA-local tile index
                    if (cq == 0 && tile_x + t < tile_end) {
                        const uint32_t slot = lt & (t3::kOutSlots - 1);
                        float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < total) {
                            out.x = fmaxf(sg + sBias[pk::kBiasSigma], 0.f);
                            out.y = 1.0f / (1.0f + __expf(-(__uint_as_float(v[0]) + sBias[pk::kBiasRgb + 0])));
                            out.z = 1.0f / (1.0f + __expf(-(__uint_as_float(v[1]) + sBias[pk::kBiasRgb + 1])));
                            out.w = 1.0f / (1.0f + __expf(-(__uint_as_float(v[2]) + sBias[pk::kBiasRgb + 2])));
                            if (sigma_out) {                                         // training keeps them for the backward pass
                                sigma_out[row] = out.x;
                                rgb_out[row * 3 + 0] = out.y; rgb_out[row * 3 + 1] = out.z; rgb_out[row * 3 + 2] = out.w;
                            }
                        }
                        // the slot's previous tile (kOutSlots tiles ago) has been composited
                        umma::mbar_wait_u32(bars + 8u * (t3::kBarOutEmpty + slot), ((lt >> 3) & 1u) ^ 1u);
                        ((float4*)(smem + t3::kOffOut))[slot * t3::kTileM + r] = out;
                        __syncwarp();
                        if (lane == 0) umma::mbar_arrive_u32(bars + 8u * (t3::kBarOutFull + slot));
                    }
                } else if (cq == 0 && row < total) {
                    sigma_out[row] = fmaxf(sg + sBias[pk::kBiasSigma], 0.f);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float x = __uint_as_float(v[k]) + sBias[pk::kBiasRgb + k];
                        rgb_out[row * 3 + k] = 1.0f / (1.0f + __expf(-x));
                    }
                }
            }
        }
        if (PROFILE && tid == 128) { dbg[blockIdx.x * 16 + 6] = clock64() - t_begin; dbg[blockIdx.x * 16 + 7] = t_wait; for (int i = 0; i < 4; ++i) dbg[blockIdx.x * 16 + 10 + i] = tp[i]; }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (MC) umma::cluster_sync_all();          // neither CTA leaves while its peer can still write into its ring / signal its barriers
    if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

// Fused compositing needs whole 32-sample chunks per ray and a ray group (whole rays = whole tiles) that fits the output ring
bool mlp_tc3_can_composite(int S, int* group_rays, int* group_tiles) {
    if (S <= 0 || S % 32 != 0) return false;
    int g = 1;
    while ((g * S) % t3::kTileM != 0) ++g;              // 128 / gcd(S, 128) <= 4
    const int tg = g * S / t3::kTileM;
    if (tg > t3::kOutSlots - 2) return false;
    if (group_rays) *group_rays = g;
    if (group_tiles) *group_tiles = tg;
    return true;
}

int launch_mlp_tc3(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                   int64_t total, int S, float* sigma, float* rgb, void* act_out, void* mask_out, void* stream, long long* dbg,
                   const CompositeOutputs* comp) {
    static thread_local unsigned long long attr_mask = 0;
    if (attrs_pending(attr_mask)) {
        cudaError_t e = allow_smem(mlp_tc3_kernel<false, false, false, false>, t3::kSmemBytes);
        if (e == cudaSuccess) e = allow_smem(mlp_tc3_kernel<false, true, false, false>, t3::kSmemBytes);
        if (e == cudaSuccess) e = allow_smem(mlp_tc3_kernel<false, false, true, false>, t3::kSmemBytes);
        if (e == cudaSuccess) e = allow_smem(mlp_tc3_kernel<false, true, true, false>, t3::kSmemBytes);
        if (e == cudaSuccess) e = allow_smem(mlp_tc3_kernel<false, false, true, true>, t3::kSmemBytes);
        if (e == cudaSuccess) e = allow_smem(mlp_tc3_kernel<false, true, true, true>, t3::kSmemBytes);
#ifdef NERF_DEBUG_BUILD
        if (e == cudaSuccess) e = allow_smem(mlp_tc3_kernel<true, false, false, false>, t3::kSmemBytes);
#endif
        if (e != cudaSuccess) { set_error("nerf_mlp_forward_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attrs_done(attr_mask);
    }
    const int64_t tiles = (total + t3::kTileM - 1) / t3::kTileM;
    const int64_t pairs = (tiles + 1) / 2;
    const uint8_t* pw = (const uint8_t*)packed;
    __nv_bfloat16* ao = (__nv_bfloat16*)act_out;
    unsigned long long* mo = (unsigned long long*)mask_out;
    cudaStream_t st = (cudaStream_t)stream;
    FusedComposite fc{};
    if (comp) {                                         // fused compositing: contiguous ray groups per CTA, 25 warps
        if (!mlp_tc3_can_composite(S, &fc.group_rays, &fc.group_tiles) || !(ts || comp->u_c) || samples || dbg) {
            set_error("nerf_mlp_composite_tc: S = %d cannot be composited inside the kernel (needs S %% 32 == 0, rays + depths input)", S);
            return NERF_E_ARG;
        }
        fc.weights = comp->weights; fc.ray_rgb = comp->ray_rgb; fc.depth = comp->depth; fc.acc = comp->acc; fc.stats = comp->stats;
        fc.u_c = comp->u_c; fc.t_base = comp->t_base; fc.step = comp->step; fc.ts_gen = comp->ts_gen;
        fc.N = total / S;
        fc.num_groups = (fc.N + fc.group_rays - 1) / fc.group_rays;
        // at least two tiles per CTA where possible, so that both halves of a pair do useful work
        int64_t want = (tiles + 1) / 2;
        if (want > fc.num_groups) want = fc.num_groups;
        int grid = (int)(want < num_sms() ? want : num_sms());
        if (grid >= 2) {                                  // 2-CTA clusters sharing every weight stage by multicast (a single ray group
                                                          // runs as one plain CTA below)
            grid &= ~1;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(t3::kThreads); cfg.dynamicSmemBytes = t3::kSmemBytes; cfg.stream = st;
            cudaLaunchAttribute attr{};
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            const float* no_samples = nullptr;
            long long* no_dbg = nullptr;
            __nv_bfloat16* no_act = nullptr;
            unsigned long long* no_mask = nullptr;
            cudaError_t e = act_out
                ? cudaLaunchKernelEx(&cfg, mlp_tc3_kernel<false, true, true, true>, pw, o, d, ts, no_samples, total, S, sigma, rgb, ao, mo, no_dbg, fc)
                : cudaLaunchKernelEx(&cfg, mlp_tc3_kernel<false, false, true, true>, pw, o, d, ts, no_samples, total, S, sigma, rgb, no_act, no_mask, no_dbg, fc);
            if (e != cudaSuccess) { set_error("nerf_mlp_composite_tc: cluster launch: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
            return check_launch("nerf_mlp_composite_tc");
        }
        if (act_out)
            mlp_tc3_kernel<false, true, true, false><<<grid, t3::kThreads, t3::kSmemBytes, st>>>(pw, o, d, ts, nullptr, total, S, sigma, rgb, ao, mo, nullptr, fc);
        else
            mlp_tc3_kernel<false, false, true, false><<<grid, t3::kThreads, t3::kSmemBytes, st>>>(pw, o, d, ts, nullptr, total, S, sigma, rgb, nullptr, nullptr, nullptr, fc);
        return check_launch("nerf_mlp_composite_tc");
    }
    int grid = (int)(pairs < num_sms() ? pairs : num_sms());
#ifdef NERF_DEBUG_BUILD
    if (dbg) {                                        // diagnostic library only: per-CTA cycle counters
        mlp_tc3_kernel<true, false, false, false><<<grid, t3::kThreads, t3::kSmemBytes, st>>>(pw, o, d, ts, samples, total, S, sigma, rgb, nullptr, nullptr, dbg, fc);
        return check_launch("nerf_debug_mlp_tc_profile");
    }
#else
    if (dbg) { set_error("nerf_mlp_forward_tc: cycle counters exist in the diagnostic library only"); return NERF_E_ARG; }
#endif
    if (act_out)                                      // training form: also stores activations + sign words
        mlp_tc3_kernel<false, true, false, false><<<grid, t3::kThreads, t3::kSmemBytes, st>>>(pw, o, d, ts, samples, total, S, sigma, rgb, ao, mo, nullptr, fc);
    else
        mlp_tc3_kernel<false, false, false, false><<<grid, t3::kThreads, t3::kSmemBytes, st>>>(pw, o, d, ts, samples, total, S, sigma, rgb, nullptr, nullptr, nullptr, fc);
    return check_launch("nerf_mlp_forward_tc");
}

}  // namespace nerf

using namespace nerf;

static int launch_mlp_tc(const void* packed, const float* o, const float* d, const float* ts, const float* samples,
                         int64_t N, int S, float* sigma, float* rgb, void* stream, long long* dbg = nullptr,
                         void* act_out = nullptr, void* mask_out = nullptr) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_forward_tc: bad size N=%lld S=%d", (long long)N, S);
    if (N == 0) return 0;
    NERF_REQUIRE(packed && d && sigma && rgb, "nerf_mlp_forward_tc: null pointer");
    NERF_REQUIRE(samples || (o && ts), "nerf_mlp_forward_tc: need either samples or (o, ts)");
    NERF_REQUIRE(((uintptr_t)packed & 127) == 0, "nerf_mlp_forward_tc: packed buffer must be 128-byte aligned");
    NERF_REQUIRE(!act_out || ((uintptr_t)act_out & 15) == 0, "nerf_mlp_forward_tc: act_out must be 16-byte aligned");
    return launch_mlp_tc3(packed, o, d, ts, samples, N * S, S, sigma, rgb, act_out, mask_out, stream, dbg, nullptr);
}

#ifdef NERF_DEBUG_BUILD
// Diagnostic library only (tools/, not in include/nerf_b200.h): same kernel with per-CTA cycle counters, dbg = [grid,16] int64.
extern "C" NERF_API int nerf_debug_mlp_tc_profile(const void* packed, const float* o, const float* d, const float* ts,
                                                  int64_t N, int S, float* sigma, float* rgb, long long* dbg, void* stream) {
    return launch_mlp_tc(packed, o, d, ts, nullptr, N, S, sigma, rgb, stream, dbg);
}
#endif

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

// ---- K8 + K2: the same network with alpha compositing (nerf_helpers.py:58-104) done inside the kernel (COMP form above).
// sigma / rgb may be NULL (render: the per-sample outputs never leave the SM); act_out / mask_out non-NULL selects the
// training form, which needs sigma and rgb as well (the compositing backward reads them).
extern "C" int nerf_mlp_composite_tc_supported(int S) { return mlp_tc3_can_composite(S, nullptr, nullptr) ? 1 : 0; }

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
