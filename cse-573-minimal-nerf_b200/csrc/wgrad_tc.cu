// wgrad_tc.cu - weight and bias gradients of one NeRFModel on the tensor cores.
//
//   dW_l[n_out, k_in] = sum_samples dz_l[s, n_out] * a_l[s, k_in]          db_l[n_out] = sum_samples dz_l[s, n_out]
//
// The contraction runs over samples, so both operands are "MN-major" for the MMA.  The training path stores activations
// (`acts`) and pre-activation gradients (`dz`) tiled chunk-major (pack_layout.cuh), in which every [128 samples x 64
// features] block is a contiguous 16 KB that already IS the UMMA no-swizzle MN-major canonical layout: operands arrive by
// plain 32 / 64 KB bulk copies, no transposition anywhere.
//
// Work decomposition (one launch per network, one CTA per SM):
//   a "job" = up to four 128-row A blocks (halves of a layer's dz, or halves of feat / r for the two small heads) that all
//   multiply the SAME B operand (a 256-wide activation block and / or a recomputed PE tile); block i accumulates into its
//   own TMEM region D_i[128 x (b_cols + pe_cols)] (fp32).  9 jobs (table below); each is split-K over the 128-sample
//   tiles among a fixed group of CTAs.  Per tile a CTA reads every operand once: 1.03 MB per tile over all jobs against
//   0.96 MB of acts + dz, the minimum (dz of feature_fn.0 is read by two jobs; feat used to be, until the density head's
//   product feat^T . dsigma moved into the rgb_fn.0 job, where feat is resident as the B tile).  At the end every CTA adds its partial D (and bias sums) to the fp32 gradients with
//   atomics.
// Per CTA: warp 0 issues the A-block bulk copies (2 x 32 KB ring), warp 2 the B-tile copies (2 x 64 KB ring), warp 1 issues
// tcgen05.mma (M128 x N<=256 x K16, both operands MN-major from shared memory), warps 4-7 sum the dz blocks' columns for
// the bias gradients while they sit in shared memory and run the final TMEM -> atomics epilogue, warps 8-11 recompute the
// PE(x) / PE(dir) tile for the products whose input is a positional encoding.  The kernel is HBM-bound by design (128 KB of
// operands per 2048 clk of MMA); its roofline is bytes / HBM bandwidth.
#include "mlp_tc_common.cuh"

namespace nerf {

namespace wg {
constexpr int kThreads = 384;
constexpr int kSlots = 2;
constexpr uint32_t kABytes = 32768;        // [128 samples x 128 n_out]
constexpr uint32_t kBBytes = 65536;        // [128 samples x up to 256 k_in]
constexpr uint32_t kPEBytes = 16384;       // [128 samples x 64] recomputed encoding
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kSlots * kABytes;                       // 65536
constexpr uint32_t kOffPE = kOffB + kSlots * kBBytes;              // 196608
constexpr uint32_t kHeadsBytes = 4096;     // [128 samples x 16] dz heads block (density job folded into rgb_fn.0's)
constexpr uint32_t kOffHeads = kOffPE + kPEBytes;                  // 212992 (single PE buffer), 2 x 4 KB
constexpr uint32_t kOffBars = kOffHeads + kSlots * kHeadsBytes;    // 221184
constexpr uint32_t kOffTmemHolder = kOffBars + 16 * 8;
constexpr uint32_t kSmemBytes = kOffTmemHolder + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum { SRC_ACTS = 0, SRC_DZ = 1 };
enum { PE_NONE = 0, PE_X = 1, PE_DIR = 2 };
enum { OUT_NORMAL = 0, OUT_DENSITY = 1, OUT_RGB = 2 };

struct Block {
    uint8_t src;          // SRC_ACTS / SRC_DZ
    uint16_t feat;        // first of its 128 features
    uint8_t w_param;      // weight gradient it contributes to (state_dict order 0..9)
    uint16_t row0;        // first dW row (OUT_NORMAL) / first dW column (OUT_DENSITY)
    uint16_t in_features; // leading dimension of that dW
    uint16_t pe_col0;     // dW column where the PE part goes
    uint8_t bias;         // 1: also accumulate db[w_param][row0 : row0+128] = column sums of this block
};
struct Job {
    uint8_t nA;
    Block blk[4];
    uint8_t b_cols16;     // columns of the B tile read from memory / 16 (0 = none)
    uint8_t b_src;
    uint16_t b_feat;
    uint8_t pe;           // PE_NONE / PE_X / PE_DIR: extra (or only) B operand recomputed in the kernel
    uint8_t pe_valid;     // valid PE columns (60 / 24)
    uint8_t out_kind;
    uint8_t heads_bias;   // 1: db of the two heads = column sums of the B tile (the dz heads block)
    uint8_t ctas;         // CTAs that split this job's tiles
    uint8_t dens;         // 1: the B tile is feat: its two 128-feature halves, used as A operands against the dz heads block
                          //    (4 KB more per tile), also give dW(density_fn.0) - feat is not read a second time by a job of its own
};

// dz feature offsets: layer k's pre-activation gradient at 256 k (k = 0..6), rgb_fn.0 at 1792, heads block at 1920.
// acts feature offsets: output of layer k at 256 k, rgb_fn.0 output r at 1792.
#define NB {0, 0, 0, 0, 0, 0, 0}
__constant__ Job c_jobs[9] = {
    // dz0 and dz4 halves x PE(x): dW(mlp.0)[:, 0:60] and dW(feature_fn.0)[:, 256:316], and both layers' biases
    {4, {{SRC_DZ, 0, 0, 0, 60, 0, 1}, {SRC_DZ, 128, 0, 128, 60, 0, 1}, {SRC_DZ, 1024, 4, 0, 316, 256, 1}, {SRC_DZ, 1152, 4, 128, 316, 256, 1}},
     0, SRC_ACTS, 0, PE_X, 60, OUT_NORMAL, 0, 25, 0},
    {2, {{SRC_DZ, 256, 1, 0, 256, 0, 1}, {SRC_DZ, 384, 1, 128, 256, 0, 1}, NB, NB}, 16, SRC_ACTS, 0, PE_NONE, 0, OUT_NORMAL, 0, 17, 0},      // mlp.2
    {2, {{SRC_DZ, 512, 2, 0, 256, 0, 1}, {SRC_DZ, 640, 2, 128, 256, 0, 1}, NB, NB}, 16, SRC_ACTS, 256, PE_NONE, 0, OUT_NORMAL, 0, 17, 0},    // mlp.4
    {2, {{SRC_DZ, 768, 3, 0, 256, 0, 1}, {SRC_DZ, 896, 3, 128, 256, 0, 1}, NB, NB}, 16, SRC_ACTS, 512, PE_NONE, 0, OUT_NORMAL, 0, 18, 0},    // mlp.6
    {2, {{SRC_DZ, 1024, 4, 0, 316, 0, 0}, {SRC_DZ, 1152, 4, 128, 316, 0, 0}, NB, NB}, 16, SRC_ACTS, 768, PE_NONE, 0, OUT_NORMAL, 0, 17, 0},  // feature_fn.0 (h3 part)
    {2, {{SRC_DZ, 1280, 5, 0, 256, 0, 1}, {SRC_DZ, 1408, 5, 128, 256, 0, 1}, NB, NB}, 16, SRC_ACTS, 1024, PE_NONE, 0, OUT_NORMAL, 0, 18, 0}, // feature_fn.2
    {2, {{SRC_DZ, 1536, 6, 0, 256, 0, 1}, {SRC_DZ, 1664, 6, 128, 256, 0, 1}, NB, NB}, 16, SRC_ACTS, 1280, PE_NONE, 0, OUT_NORMAL, 0, 17, 0}, // feature_fn.4
    // rgb_fn.0: dz_r^T . [feat | PE(dir)], and with feat resident also density_fn.0: feat^T . heads (column 0 = dsigma)
    {1, {{SRC_DZ, 1792, 8, 0, 280, 256, 1}, NB, NB, NB}, 16, SRC_ACTS, 1536, PE_DIR, 24, OUT_NORMAL, 0, 12, 1},
    {1, {{SRC_ACTS, 1792, 9, 0, 128, 0, 0}, NB, NB, NB}, 1, SRC_DZ, 1920, PE_NONE, 0, OUT_RGB, 1, 7, 0},                                 // rgb_fn.2: r^T . heads (+ head biases)
};
#undef NB
constexpr int kNumJobs = 9;
constexpr int kGridCtas = 25 + (17 + 17 + 18) + (17 + 18 + 17) + 12 + 7;          // 148: CTA shares follow the measured per-tile cost of each job
                                                                 // (tools/profile_wgrad.py: the PE(x) job recomputes 60 sin/cos per row and is
                                                                 // issue-bound, the 256-wide jobs are HBM-bound), not its bytes
}  // namespace wg

struct Grads { float* p[20]; };

#ifdef NERF_DEBUG_BUILD
// diagnostic library only: per-CTA (elapsed clocks, clocks until the last MMA completed) of the most recent launch (tools/profile_wgrad.py)
__device__ long long g_wgrad_cycles[2 * 160];
#define NERF_WGRAD_CYCLES(i, v) g_wgrad_cycles[i] = (v)
#else
#define NERF_WGRAD_CYCLES(i, v) ((void)0)
#endif

__global__ void __launch_bounds__(wg::kThreads, 1)
wgrad_tc_kernel(const __nv_bfloat16* __restrict__ acts, const __nv_bfloat16* __restrict__ dz, const float* __restrict__ o_rays,
                const float* __restrict__ d_rays, const float* __restrict__ ts, int64_t total, int S, Grads G, const bool one_cta_per_job) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + wg::kOffBars);
    uint64_t* fullA = bars;            // [2]
    uint64_t* emptyA = bars + 2;       // [2]  MMA commit + 4 bias-sum warps
    uint64_t* fullB = bars + 4;        // [2]
    uint64_t* emptyB = bars + 6;       // [2]  MMA commit (+ 4 bias-sum warps for the heads job)
    uint64_t* pe_full = bars + 8;
    uint64_t* pe_empty = bars + 9;
    uint64_t* done = bars + 10;
    uint32_t* tmem_holder = (uint32_t*)(smem + wg::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (total + 127) / 128;
    const long long t_start = clock64();

    // which job, and which share of its tiles, this CTA owns
    int job_idx = 0, first = blockIdx.x;
    // deterministic form: ONE CTA per job (grid = 9) walks all tiles in order, so every gradient element receives exactly one
    // atomic add and the result is bit-reproducible (a parity-debugging mode: ~16x slower)
    if (one_cta_per_job) { job_idx = blockIdx.x; first = 0; }
    else while (job_idx < wg::kNumJobs && first >= wg::c_jobs[job_idx].ctas) { first -= wg::c_jobs[job_idx].ctas; ++job_idx; }
    if (job_idx >= wg::kNumJobs) return;                 // spare CTAs
    const wg::Job job = wg::c_jobs[job_idx];
    const int stride = one_cta_per_job ? 1 : job.ctas, nA = job.nA;
    const int b_cols = job.b_cols16 * 16;
    const bool has_pe = job.pe != wg::PE_NONE;
    const int region_cols = b_cols + (has_pe ? 64 : 0);  // TMEM columns per A block
    const int bsrc_chunks = (job.b_src == wg::SRC_DZ) ? pk::kDzChunks : pk::kActChunks;
    const __nv_bfloat16* b_base = (job.b_src == wg::SRC_DZ ? dz : acts) + (int64_t)(job.b_feat >> 3) * 1024;
    const uint32_t b_bytes = (uint32_t)b_cols * 256u;    // 128 rows x b_cols x 2 B

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&fullA[i], 1);
            umma::mbar_init(&emptyA[i], 5);
            umma::mbar_init(&fullB[i], 1);
            umma::mbar_init(&emptyB[i], job.heads_bias ? 5 : 1);
        }
        umma::mbar_init(pe_full, 4);
        umma::mbar_init(pe_empty, 1);
        umma::mbar_init(done, 1);
        umma::fence_mbar_init();
    }
    if (warp == 3) umma::tmem_alloc(tmem_holder, 512);
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 0) {
        // ------------------------------------------------------------------ A-block producer (2 x 32 KB ring)
        const bool leader = umma::elect_one();
        uint32_t ua = 0;
        for (int64_t tile = first; tile < num_tiles; tile += stride) {
            for (int i = 0; i < nA; ++i, ++ua) {
                const uint32_t slot = ua & 1, ph = (ua >> 1) & 1;
                umma::mbar_wait(&emptyA[slot], ph ^ 1);
                if (leader) {
                    const wg::Block bk = job.blk[i];
                    const int chunks = (bk.src == wg::SRC_DZ) ? pk::kDzChunks : pk::kActChunks;
                    const __nv_bfloat16* src = (bk.src == wg::SRC_DZ ? dz : acts) + (tile * chunks + (bk.feat >> 3)) * 1024;
                    umma::mbar_arrive_expect_tx(&fullA[slot], wg::kABytes);
                    umma::bulk_g2s(smem + wg::kOffA + slot * wg::kABytes, src, wg::kABytes, &fullA[slot]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------------ B-tile producer (2 x 64 KB ring)
        if (b_cols > 0) {
            const bool leader = umma::elect_one();
            uint32_t it = 0;
            for (int64_t tile = first; tile < num_tiles; tile += stride, ++it) {
                const uint32_t slot = it & 1, ph = (it >> 1) & 1;
                umma::mbar_wait(&emptyB[slot], ph ^ 1);
                if (leader) {
                    umma::mbar_arrive_expect_tx(&fullB[slot], b_bytes + (job.dens ? wg::kHeadsBytes : 0u));
                    umma::bulk_g2s(smem + wg::kOffB + slot * wg::kBBytes, b_base + tile * (int64_t)bsrc_chunks * 1024, b_bytes, &fullB[slot]);
                    if (job.dens)          // the dz heads block [128 x 16] of the same tile
                        umma::bulk_g2s(smem + wg::kOffHeads + slot * wg::kHeadsBytes, dz + (tile * pk::kDzChunks + (1920 >> 3)) * 1024,
                                       wg::kHeadsBytes, &fullB[slot]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const bool leader = umma::elect_one();
        const uint32_t kMN = umma::kIdescAMajorMN | umma::kIdescBMajorMN;
        const uint32_t idescB = umma::make_idesc_bf16(128, b_cols > 0 ? b_cols : 16) | kMN;
        const uint32_t idescPE = umma::make_idesc_bf16(128, 64) | kMN;
        uint32_t it = 0, ua = 0;
        for (int64_t tile = first; tile < num_tiles; tile += stride, ++it) {
            const uint32_t bslot = it & 1, bph = (it >> 1) & 1;
            if (b_cols > 0) umma::mbar_wait(&fullB[bslot], bph);
            if (has_pe) umma::mbar_wait(pe_full, it & 1);
            for (int i = 0; i < nA; ++i, ++ua) {
                const uint32_t slot = ua & 1, ph = (ua >> 1) & 1;
                umma::mbar_wait(&fullA[slot], ph);
                umma::tc_fence_after();
                if (leader) {
                    const uint32_t a_addr = umma::smem_u32(smem + wg::kOffA + slot * wg::kABytes);
                    const uint32_t b_addr = umma::smem_u32(smem + wg::kOffB + bslot * wg::kBBytes);
                    const uint32_t pe_addr = umma::smem_u32(smem + wg::kOffPE);
                    const uint32_t d_col = (uint32_t)(i * region_cols);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {             // 128 samples = 8 K=16 slices, 256 B apart along the rows
                        const uint64_t adesc = umma::make_desc_mn_interleave(a_addr + k * 256, 2048, 128);
                        if (b_cols > 0)
                            umma::mma_ss(tmem + d_col, adesc, umma::make_desc_mn_interleave(b_addr + k * 256, 2048, 128), idescB,
                                         (it | (uint32_t)k) != 0);
                        if (has_pe)
                            umma::mma_ss(tmem + d_col + (uint32_t)b_cols, adesc, umma::make_desc_mn_interleave(pe_addr + k * 256, 2048, 128),
                                         idescPE, (it | (uint32_t)k) != 0);
                    }
                    umma::mma_commit(&emptyA[slot]);
                }
                __syncwarp();
            }
            if (leader) {
                if (job.dens) {            // density_fn.0: the two 128-feature halves of the resident feat tile as A operands, N = 16
                    const uint32_t b_addr = umma::smem_u32(smem + wg::kOffB + bslot * wg::kBBytes);
                    const uint32_t h_addr = umma::smem_u32(smem + wg::kOffHeads + bslot * wg::kHeadsBytes);
                    const uint32_t idesc16 = umma::make_idesc_bf16(128, 16) | kMN;
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma::mma_ss(tmem + (uint32_t)(nA * region_cols + 16 * h),
                                         umma::make_desc_mn_interleave(b_addr + h * 32768 + k * 256, 2048, 128),
                                         umma::make_desc_mn_interleave(h_addr + k * 256, 2048, 128), idesc16, (it | (uint32_t)k) != 0);
                }
                if (b_cols > 0) umma::mma_commit(&emptyB[bslot]);
                if (has_pe) umma::mma_commit(pe_empty);
            }
            __syncwarp();
        }
        if (leader) umma::mma_commit(done);
        __syncwarp();
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ PE producers (thread = sample row), chunk-major tile
        if (has_pe) {
            const int r = (warp - 8) * 32 + lane;
            uint32_t it = 0;
            for (int64_t tile = first; tile < num_tiles; tile += stride, ++it) {
                const int64_t row = tile * 128 + r;
                const bool valid = row < total;
                const int64_t n = valid ? row / S : 0;
                float x[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    const float dx = __ldg(d_rays + n * 3), dy = __ldg(d_rays + n * 3 + 1), dzz = __ldg(d_rays + n * 3 + 2);
                    if (job.pe == wg::PE_X) {
                        const float t = ts[row];
                        x[0] = __fdiv_rn(__fadd_rn(__fmul_rn(dx, t), __ldg(o_rays + n * 3 + 0)), tcm::kPiF);
                        x[1] = __fdiv_rn(__fadd_rn(__fmul_rn(dy, t), __ldg(o_rays + n * 3 + 1)), tcm::kPiF);
                        x[2] = __fdiv_rn(__fadd_rn(__fmul_rn(dzz, t), __ldg(o_rays + n * 3 + 2)), tcm::kPiF);
                    } else {
                        const float nrm = sqrtf(dx * dx + dy * dy + dzz * dzz);
                        x[0] = __fdiv_rn(dx, nrm); x[1] = __fdiv_rn(dy, nrm); x[2] = __fdiv_rn(dzz, nrm);
                    }
                }
                uint32_t v[32];
                if (job.pe == wg::PE_X) encode_row<10>(x, v); else encode_row<4>(x, v);
                if (!valid) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                umma::mbar_wait(pe_empty, (it & 1) ^ 1);
                uint4* dst = (uint4*)(smem + wg::kOffPE) + r;          // chunk c of row r at c * 2048 + r * 16
#pragma unroll
                for (int c = 0; c < 8; ++c) dst[c * 128] = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(pe_full);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ bias sums while the tiles are resident, then the epilogue
        const int t = (warp - 4) * 32 + lane;          // 0..127
        const int chunk = t >> 3, rsub = t & 7;        // 16 chunks x 8 row phases
        float bsum[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[i][j] = 0.f;
        auto add_block = [&](const uint8_t* src, float (&acc)[8]) {
#pragma unroll 4
            for (int k = 0; k < 16; ++k) {
                const uint4 q = *(const uint4*)(src + chunk * 2048 + (rsub + 8 * k) * 16);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[2 * j] += __uint_as_float(w[j] << 16);
                    acc[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                }
            }
        };
        {
            uint32_t it = 0, ua = 0;
            for (int64_t tile = first; tile < num_tiles; tile += stride, ++it) {
                if (job.heads_bias) {                  // column sums of the dz heads block (the B tile): [dsigma, drgb x3, 0..]
                    const uint32_t bslot = it & 1, bph = (it >> 1) & 1;
                    umma::mbar_wait(&fullB[bslot], bph);
                    if (chunk == 0) add_block(smem + wg::kOffB + bslot * wg::kBBytes, bsum[3]);
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&emptyB[bslot]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < nA) {
                        const uint32_t slot = ua & 1, ph = (ua >> 1) & 1;
                        umma::mbar_wait(&fullA[slot], ph);
                        if (job.blk[i].bias) add_block(smem + wg::kOffA + slot * wg::kABytes, bsum[i]);
                        __syncwarp();
                        if (lane == 0) umma::mbar_arrive(&emptyA[slot]);
                        ++ua;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bsum[i][j] += __shfl_xor_sync(kFull, bsum[i][j], 1);
                bsum[i][j] += __shfl_xor_sync(kFull, bsum[i][j], 2);
                bsum[i][j] += __shfl_xor_sync(kFull, bsum[i][j], 4);
            }
        if (rsub == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < nA && job.blk[i].bias) {
                    float* db = G.p[2 * job.blk[i].w_param + 1] + job.blk[i].row0 + chunk * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) atomicAdd(db + j, bsum[i][j]);
                }
            }
            if (job.heads_bias && chunk == 0) {
                atomicAdd(G.p[15], bsum[3][0]);                        // density_fn.0.bias
                atomicAdd(G.p[19] + 0, bsum[3][1]);                    // rgb_fn.2.bias
                atomicAdd(G.p[19] + 1, bsum[3][2]);
                atomicAdd(G.p[19] + 2, bsum[3][3]);
            }
        }
        // ---- epilogue: D_i[m = lane (row of block i), n] -> atomics into dW (skipped by CTAs that had no tile)
        umma::mbar_wait(done, 0);
        umma::tc_fence_after();
        if (tid == 128) NERF_WGRAD_CYCLES(2 * blockIdx.x + 1, clock64() - t_start);
        if (first < num_tiles) {
            const int m = (warp & 3) * 32 + lane;
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            for (int i = 0; i < nA; ++i) {
                const wg::Block bk = job.blk[i];
                float* dW = G.p[2 * bk.w_param];
                const uint32_t d_col = (uint32_t)(i * region_cols);
                if (job.out_kind == wg::OUT_NORMAL) {
                    float* rowp = dW + (size_t)(bk.row0 + m) * bk.in_features;
                    for (int c0 = 0; c0 < b_cols; c0 += 32) {
                        uint32_t v[32];
                        umma::tmem_ld32(tmem + lane_base + d_col + c0, v);
                        umma::tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(rowp + c0 + j, __uint_as_float(v[j]));
                    }
                    if (has_pe) {
                        float* pep = rowp + bk.pe_col0;
                        for (int c0 = 0; c0 < 64; c0 += 32) {
                            uint32_t v[32];
                            umma::tmem_ld32(tmem + lane_base + d_col + b_cols + c0, v);
                            umma::tmem_wait_ld();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < job.pe_valid) atomicAdd(pep + c0 + j, __uint_as_float(v[j]));
                        }
                    }
                } else {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + d_col, v);
                    umma::tmem_wait_ld();
                    if (job.out_kind == wg::OUT_DENSITY) {
                        atomicAdd(dW + bk.row0 + m, __uint_as_float(v[0]));                      // density_fn.0.weight [1,256]
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) atomicAdd(dW + c * 128 + m, __uint_as_float(v[1 + c]));   // rgb_fn.2.weight [3,128]
                    }
                }
            }
            if (job.dens) {                // density_fn.0.weight [1,256]: column 0 (dsigma) of the two feat halves' products
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[4];
                    umma::tmem_ld4(tmem + lane_base + (uint32_t)(nA * region_cols + 16 * h), v);
                    umma::tmem_wait_ld();
                    atomicAdd(G.p[14] + h * 128 + m, __uint_as_float(v[0]));
                }
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (tid == 0) NERF_WGRAD_CYCLES(2 * blockIdx.x, clock64() - t_start);
    if (warp == 3) umma::tmem_dealloc(tmem, 512);
}

}  // namespace nerf

using namespace nerf;

#ifdef NERF_DEBUG_BUILD
extern "C" NERF_API int nerf_debug_wgrad_cycles(long long* host_out320) {
    return cudaMemcpyFromSymbol(host_out320, g_wgrad_cycles, sizeof(long long) * 320) == cudaSuccess ? 0 : NERF_E_CUDA;
}
#endif

extern "C" int nerf_wgrad_tc(const void* acts, const void* dz, const float* o, const float* d, const float* ts, int64_t N, int S,
                             float* const* grads20_host, int deterministic, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_wgrad_tc: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(acts && dz && o && d && ts && grads20_host, "nerf_wgrad_tc: null pointer");
    NERF_REQUIRE(((uintptr_t)acts & 127) == 0 && ((uintptr_t)dz & 127) == 0, "nerf_wgrad_tc: acts / dz must be 128-byte aligned");
    Grads G;
    for (int i = 0; i < 20; ++i) {
        NERF_REQUIRE(grads20_host[i], "nerf_wgrad_tc: grads20_host[%d] is NULL", i);
        G.p[i] = grads20_host[i];
    }
    static thread_local unsigned long long attr_mask = 0;
    if (attrs_pending(attr_mask)) {
        cudaError_t e = allow_smem(wgrad_tc_kernel, wg::kSmemBytes);
        if (e != cudaSuccess) { set_error("nerf_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attrs_done(attr_mask);
    }
    wgrad_tc_kernel<<<deterministic ? wg::kNumJobs : wg::kGridCtas, wg::kThreads, wg::kSmemBytes, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)acts, (const __nv_bfloat16*)dz, o, d, ts, N * S, S, G, deterministic != 0);
    return check_launch("nerf_wgrad_tc");
}
