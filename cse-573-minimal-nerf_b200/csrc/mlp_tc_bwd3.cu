// mlp_tc_bwd3.cu - dgrad chain of one NeRFModel (autograd of nerf_model.py:362-389), two 128-sample tiles per CTA: the
// schedule of mlp_tc3.cu (shared W^T stages, one MMA-issuing warp per tile with a turn token, in-place TMEM operands, all
// 16 epilogue warps on every task) run in reverse:
//   per tile, given the gradients w.r.t. the head pre-activations (from composite_backward_kernel)
//     dr   = (drgb_pre . W9) * [r > 0]                         CUDA cores, producer warps 20-23 -> smem A tile (K = 128)
//     dz6  = dr . W8[:, :256] + dsigma_pre (x) w7              steps 0,1   (A = dr tile in shared memory, SS)
//     dz5  = (dz6 . W6) * [h5 > 0]                             steps 2,3   (A = previous dz in TMEM, TS; B = W^T stages)
//     dz4 .. dz0 likewise through feature_fn.2, feature_fn.0 (h columns), mlp.6, mlp.4, mlp.2     steps 4..13
//   every dz is written to global (bf16, tiled chunk-major, pack_layout.cuh) for wgrad, plus the 16-wide heads block.
// ReLU masks are the 32-bit SIGN words mlp_tc3.cu wrote per (row, 32-feature group): packed pair j -> bits 15-j / 31-j,
// 1 = pre-activation negative; they are applied to the PACKED bf16 pairs (shift + PRMT sign replication + AND-NOT: three
// instructions per pair).
// mlp.0's dgrad (d PE) is not needed: the inputs carry no gradient.
//
// FUSED form (what training.py launches): the compositing backward (autograd of nerf_helpers.py:58-104) runs INSIDE this kernel.
// The four producer warps take the rays that overlap the tile pair (one ray per warp at a time, composite_backward_ray of
// composite_scan.cuh: the same routine as the stand-alone composite_backward_kernel, so dz is bit-identical), leave
// (dsigma_pre, drgb_pre) of the pair's 256 samples in shared memory (double-buffered by pair), and build the dr tile / heads
// block from there; the epilogue takes dsigma_pre from the same buffer.  Inputs are then the forward's saved sigma / rgb /
// depths and dL/d(ray colour) instead of two [N*S] gradient arrays: one launch and 32 B per sample of HBM traffic less.
#include <type_traits>
#include "mlp_tc3_common.cuh"
#include "composite_scan.cuh"

namespace nerf {

// 0xFFFF fields for the bf16 halves whose sign-word bits (bit 15 -> low half, bit 31 -> high half of `s`) are set
// (= the halves to ZERO):
// prmt in its generic mode replicates the msb of the selected byte when the selector nibble has bit 3 set
// (__byte_perm only forwards three selector bits per nibble).
__device__ __forceinline__ uint32_t drop_mask(uint32_t s) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(r) : "r"(s));
    return r;
}
// the same from bits 7 / 23 (the msbs of bytes 0 and 2): pair 8 + k of a sign word needs only the shift by k that pair k needs,
// so 16 pairs cost 7 shifts instead of 15
__device__ __forceinline__ uint32_t drop_mask_lo(uint32_t s) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(r) : "r"(s));
    return r;
}

// Storing every dz chunk as soon as it is packed (instead of the task's four chunks at its end) helps the forward kernel's
// training form by 1 % and costs this kernel 0.7 % (tools/ab_build_flag.sh): off here.
// 1: a chain step requests its four sign words up front (dgrad 0.976 -> 0.964 ms; requesting them a whole step ahead
// measured the same and is not kept)
#ifndef NERF_BWD_MASK_PREFETCH
#define NERF_BWD_MASK_PREFETCH 1
#endif
#ifndef NERF_INTERLEAVE_STORES_BWD
#define NERF_INTERLEAVE_STORES_BWD 0
#endif

namespace b3 {
constexpr uint32_t kOffDr = 0;                                    // dr tiles of X and Y: 2 x 2 K-blocks x [128 x 64] bf16
constexpr uint32_t kOffRing = t3::kOffRing;                       // 65536
constexpr uint32_t kOffConst = kOffRing + t3::kSlots * t3::kSlotBytes;      // fp32 W9 [3][128], w7 [256]
constexpr uint32_t kOffHeadGrad = kOffConst + pk::kConstFloatsT * 4;          // FUSED: 2 pairs x 256 samples x float4 (dsigma_pre, drgb_pre)
constexpr uint32_t kOffBars = kOffHeadGrad + 2 * 256 * 16;
constexpr uint32_t kOffTmemHolder = kOffBars + t3::kNumBars * 8;
constexpr uint32_t kSmemBytes = kOffTmemHolder + 16 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr int kStagesT = pk::kStagesT;                            // 52 requests of 16 KB per tile pair
}  // namespace b3

// FUSED: (dsigma_pre, drgb_pre) are not read; (sigma, rgb, ts, g_ray, S) are.
struct HeadGradSource {
    const float* dsigma_pre; const float* drgb_pre;                      // !FUSED
    const float* sigma; const float* rgb; const float* ts; const float* g_ray; int S;      // FUSED
};

template <bool FUSED>
__global__ void __launch_bounds__(t3::kThreads, 1)
mlp_tc_bwd3_kernel(const uint8_t* __restrict__ packed_t, const uint32_t* __restrict__ masks32, const HeadGradSource src, int64_t total,
                   __nv_bfloat16* __restrict__ dz_out) {
    const float* __restrict__ dsigma_pre = src.dsigma_pre;
    const float* __restrict__ drgb_pre = src.drgb_pre;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = umma::smem_u32(smem);
    const uint32_t bars = sbase + b3::kOffBars;
    float* sConst = (float*)(smem + b3::kOffConst);        // [0,384) W9[c][k], [384,640) w7[k]
    float4* sHead = (float4*)(smem + b3::kOffHeadGrad);    // FUSED: [pair parity][256] (dsigma_pre, drgb_pre x 3)
    uint32_t* tmem_holder = (uint32_t*)(smem + b3::kOffTmemHolder);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t num_tiles = (total + t3::kTileM - 1) / t3::kTileM;
    const int64_t num_pairs = (num_tiles + 1) / 2;

    if (tid == 0) {
        uint64_t* b = (uint64_t*)(smem + b3::kOffBars);
        for (int i = 0; i < t3::kSlots; ++i) { umma::mbar_init(&b[t3::kBarFull + i], 1); umma::mbar_init(&b[t3::kBarEmpty + i], 2); }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&b[t3::kBarDFull + i], 1);
            umma::mbar_init(&b[t3::kBarDFree + i], t3::kEpiWarps);
            umma::mbar_init(&b[t3::kBarALo + i], t3::kEpiWarps);
            umma::mbar_init(&b[t3::kBarAHi + i], t3::kEpiWarps);
            umma::mbar_init(&b[t3::kBarPexFull + i], t3::kPEWarps);     // dr tile written
            umma::mbar_init(&b[t3::kBarPexEmpty + i], 1);               // dr tile no longer read
            umma::mbar_init(&b[t3::kBarTurn + i], 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 2) umma::tmem_alloc(tmem_holder, 512);
    {
        const float* gc = (const float*)(packed_t + pk::kLayoutT.const_offset);
        for (int i = tid; i < pk::kConstFloatsT; i += t3::kThreads) sConst[i] = gc[i];
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp < 4) {
        reg_dec<t3::kRegsMisc>();
        if (warp == 0 || warp == 3) {
            // -------------------------------------------------------------- W^T stage producers (alternate 32 KB requests)
            const bool leader = umma::elect_one();
            const uint32_t me = (warp == 0) ? 0u : 1u;
            const uint32_t ring = sbase + b3::kOffRing;
            uint32_t cnt = 0;
            for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
                for (int s = 0; s < b3::kStagesT; ++s, ++cnt) {
                    if ((cnt & 1u) != me) continue;
                    const uint32_t slot = cnt & (t3::kSlots - 1), ph = (cnt >> 3) & 1u;
                    umma::mbar_wait_u32(bars + 8u * (t3::kBarEmpty + slot), ph ^ 1u);
                    if (leader) {
                        umma::mbar_arrive_expect_tx_u32(bars + 8u * (t3::kBarFull + slot), t3::kSlotBytes);
                        umma::bulk_g2s_u32(ring + slot * t3::kSlotBytes, packed_t + (size_t)s * t3::kSlotBytes, t3::kSlotBytes,
                                           bars + 8u * (t3::kBarFull + slot));
                    }
                    __syncwarp();
                }
            }
        } else {
            // -------------------------------------------------------------- MMA issuers: warp 1 tile X, warp 2 tile Y
            const bool elected = umma::elect_one();
            if (warp == 1) {
                MmaTile<0, false> m;
                m.init(bars, sbase + b3::kOffRing, tmem, elected);
                m.run_bwd(sbase + b3::kOffDr, num_pairs);
            } else {
                MmaTile<1, false> m;
                m.init(bars, sbase + b3::kOffRing, tmem, elected);
                m.run_bwd(sbase + b3::kOffDr, num_pairs);
            }
        }
    } else if (warp >= 20) {
        // ------------------------------------------------------------------ dr producers: thread = row, 16-byte chunks of 8 features
        reg_dec<t3::kRegsPE>();
        const int r = (warp - 20) * 32 + lane;
        const float* W9 = sConst;
        uint32_t it = 0;
        for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x, ++it) {
            float4* head = sHead + (it & 1u) * 256;               // (the buffer's last reader was pair it - 2's first chain step)
            if (FUSED) {
                // compositing backward of every ray that overlaps this pair's 256 samples, one ray per warp at a time
                const int64_t row_lo = pair * 2 * t3::kTileM, row_hi = (row_lo + 2 * t3::kTileM < total) ? row_lo + 2 * t3::kTileM : total;
                const int64_t n_hi = (row_hi - 1) / src.S;
                for (int64_t n = row_lo / src.S + (warp - 20); n <= n_hi; n += t3::kPEWarps) {
                    const int64_t first = n * src.S - row_lo;     // pair-local row of the ray's sample 0 (may be negative)
                    composite_backward_ray(src.sigma, src.rgb, src.ts, src.g_ray, n, src.S, lane, [&](int i, float ds, float d0, float d1, float d2) {
                        const int64_t local = first + i;
                        if (local >= 0 && local < 2 * t3::kTileM) head[local] = make_float4(ds, d0, d1, d2);
                    });
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");   // the four producer warps: every sample of the pair is in place
            }
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int64_t tile = pair * 2 + t;
                const int64_t row = tile * t3::kTileM + r;
                const bool valid = row < total;
                const bool store = tile < num_tiles;              // dz_out holds whole tiles only
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, dsg_row = 0.f;
                if (FUSED) {
                    if (valid) { const float4 hg = head[t * t3::kTileM + r]; dsg_row = hg.x; g0 = hg.y; g1 = hg.z; g2 = hg.w; }
                } else if (valid) {
                    g0 = drgb_pre[row * 3]; g1 = drgb_pre[row * 3 + 1]; g2 = drgb_pre[row * 3 + 2]; dsg_row = dsigma_pre[row];
                }
                umma::mbar_wait_u32(bars + 8u * (t3::kBarPexEmpty + t), (it & 1u) ^ 1u);
                uint8_t* tile_smem = smem + b3::kOffDr + t * 32768;
                // sign words of r's four 32-feature groups 56 .. 59 (r = activations 1792..1919), requested together: the chunk loop
                // below used to wait for one (L1 / L2 latency) per chunk, and the producers' time per pair bounds the kernel once the
                // compositing backward runs here too
                uint32_t mw[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
                if (store) {
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq) mw[gq] = __ldg(masks32 + ((row >> 7) * (2 * pk::kMaskWords) + 56 + gq) * 128 + (row & 127));
                }
#pragma unroll 4
                for (int c = 0; c < 16; ++c) {                    // chunk c = features 8c .. 8c+7 of r (rgb_fn.0's ReLU output)
                    // this chunk = pairs 4(c&3)..+3 of group c >> 2
                    uint32_t mb = mw[c >> 2];
                    mb <<= (c & 3) * 4;
                    uint32_t v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = c * 8 + 2 * j;
                        const float a = g0 * W9[k] + g1 * W9[128 + k] + g2 * W9[256 + k];            // nerf_model.py:358 backward
                        const float b = g0 * W9[k + 1] + g1 * W9[128 + k + 1] + g2 * W9[256 + k + 1];
                        v[j] = umma::pack_bf16(a, b) & ~drop_mask(mb << j);
                    }
                    const uint4 q4 = make_uint4(v[0], v[1], v[2], v[3]);
                    const int kb = c >> 3, cc = c & 7;            // K block, chunk inside its 128-byte row
                    *(uint4*)(tile_smem + kb * 16384 + r * 128 + ((cc ^ (r & 7)) << 4)) = q4;
                    if (store) *(uint4*)(dz_out + pk::tiled_offset(row, 1792 + c * 8, pk::kDzChunks)) = q4;
                }
                if (store) {   // heads block (features 1920..1935): [dsigma_pre, drgb_pre x3, 0 ...] in bf16 for the head weight gradients
                    const float dsg = dsg_row;
                    uint4* dst = (uint4*)(dz_out + pk::tiled_offset(row, 1920, pk::kDzChunks));
                    dst[0] = make_uint4(umma::pack_bf16(dsg, g0), umma::pack_bf16(g1, g2), 0u, 0u);
                    dst[128] = make_uint4(0u, 0u, 0u, 0u);
                }
                umma::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive_u32(bars + 8u * (t3::kBarPexFull + t));
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: all 16 warps on every task
        reg_inc<t3::kRegsEpi>();
        const int q = warp & 3;                          // TMEM lane quarter this warp may touch
        const int cq = (warp - 4) >> 2;                  // column quarter (32 accumulator columns, 32 features)
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float* w7 = sConst + 384;
        uint32_t nd[2] = {0, 0};
        auto wait_d = [&](int t) {
            umma::mbar_wait_u32(bars + 8u * (t3::kBarDFull + t), nd[t] & 1u);
            umma::tc_fence_after();
            ++nd[t];
        };
        auto store_dz = [&](__nv_bfloat16* dz_at, const uint32_t* p) {
            uint4* dst = (uint4*)dz_at;
#pragma unroll
            for (int i = 0; i < 4; ++i) store_once(dst + i * 128, make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]));
        };

        uint32_t eit = 0;
        for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x, ++eit) {
            const int64_t row0 = pair * 2 * t3::kTileM + r;
            const bool st1 = (pair * 2 + 1 < num_tiles);
            float dsg[2] = {0.f, 0.f};
            if (FUSED) {
                // the producers' shared-memory results of this pair: acquire them through the barrier they arrived on (long complete)
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    umma::mbar_wait_u32(bars + 8u * (t3::kBarPexFull + t), eit & 1u);
                    if (row0 + t * 128 < total) dsg[t] = sHead[(eit & 1u) * 256 + t * 128 + r].x;
                }
            } else {
                dsg[0] = (row0 < total) ? dsigma_pre[row0] : 0.f;
                dsg[1] = (row0 + 128 < total) ? dsigma_pre[row0 + 128] : 0.f;
            }
            // this thread's place in the tiled chunk-major tensors as ONE running pointer pair (the epilogue is instruction-issue
            // bound): its 32-feature group of layer 6 in tile X; tile Y and a layer's second half sit at constant offsets, every
            // chain step moves the pointers back by 256 features
            constexpr int kTileDzStride = pk::kDzChunks * 1024, kHalfDzStride = (128 / 8) * 1024;       // bf16 elements
            constexpr int kTileMaskStride = 2 * pk::kMaskWords * 128, kHalfMaskStride = (128 / 32) * 128;   // 32-bit words
            __nv_bfloat16* dzp = dz_out + pk::tiled_offset(row0, 0, pk::kDzChunks) + (size_t)(6 * 32 + cq * 4) * 1024;
            const uint32_t* mkp = masks32 + ((row0 >> 7) * (2 * pk::kMaskWords)) * 128 + (row0 & 127) + (6 * 8 + cq) * 128;
            // one step of the chain (both halves, both tiles).  FIRST: dz6 = accumulator + dsigma_pre (x) w7 (density head,
            // nerf_model.py:351), no mask (feature_fn.4 is linear); otherwise the ReLU mask of layer j's saved output is applied
            // to the packed pair (bits 15-i / 31-i of the sign word -> 0xFFFF fields).  j > 0: dz_j is the next A operand.
            auto chain_step = [&](auto first_tag, int j) {
                constexpr bool FIRST = decltype(first_tag)::value;
                uint32_t hold[2][16];
#if NERF_BWD_MASK_PREFETCH
                // the step's four sign words are requested up front: only the first task can still see their (HBM) latency
                uint32_t mbq[2][2] = {{0xFFFFFFFFu, 0xFFFFFFFFu}, {0xFFFFFFFFu, 0xFFFFFFFFu}};
                if (!FIRST) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int t = 0; t < 2; ++t)
                            if (t == 0 || st1) mbq[h][t] = __ldg(mkp + t * kTileMaskStride + h * kHalfMaskStride);
                }
#endif
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const bool st = (t == 0) || st1;
                        const int col0 = h * 128 + cq * 32;
#if NERF_BWD_MASK_PREFETCH
                        const uint32_t mb = mbq[h][t];
#else
                        uint32_t mb = 0xFFFFFFFFu;
                        if (!FIRST && st) mb = __ldg(mkp + t * kTileMaskStride + h * kHalfMaskStride);
#endif
                        const uint32_t d_addr = tmem + lane_base + t3::kColD + 128u * (uint32_t)t + (uint32_t)(cq * 32);
                        const uint32_t a_addr = tmem + lane_base + t3::kColA + 128u * (uint32_t)t + (uint32_t)(cq * 16);
                        wait_d(t);
                        uint32_t v[32];
                        umma::tmem_ld32(d_addr, v);
                        if (h == 1 && j > 0) {              // held first half goes in place now: every reader of the old operand is done
                            umma::tmem_st16(a_addr, hold[t]);
                            umma::tmem_wait_st();                   // signalled together with dfree below
                        }
                        umma::tmem_wait_ld();
                        warp_arrive(bars + 8u * (t3::kBarDFree + t), lane);
                        uint32_t pk_[16];
                        uint32_t* p = (h == 0) ? hold[t] : pk_;
                        if (FIRST) {
                            const uint32_t w7s = sbase + b3::kOffConst + 4u * (uint32_t)(384 + col0);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float4 w;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(w7s + 16u * i));
                                p[2 * i] = umma::pack_bf16(fmaf(dsg[t], w.x, __uint_as_float(v[4 * i])), fmaf(dsg[t], w.y, __uint_as_float(v[4 * i + 1])));
                                p[2 * i + 1] = umma::pack_bf16(fmaf(dsg[t], w.z, __uint_as_float(v[4 * i + 2])), fmaf(dsg[t], w.w, __uint_as_float(v[4 * i + 3])));
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t sh = mb << i;
                                p[i] = umma::pack_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])) & ~drop_mask(sh);
                                p[i + 8] = umma::pack_bf16(__uint_as_float(v[2 * i + 16]), __uint_as_float(v[2 * i + 17])) & ~drop_mask_lo(sh);
#if NERF_INTERLEAVE_STORES_BWD
                                if (st && (i & 3) == 3) {       // chunks i / 4 and i / 4 + 2 are complete: let them go now
                                    uint4* dst = (uint4*)(dzp + t * kTileDzStride + h * kHalfDzStride);
                                    store_once(dst + (i >> 2) * 128, make_uint4(p[i - 3], p[i - 2], p[i - 1], p[i]));
                                    store_once(dst + ((i >> 2) + 2) * 128, make_uint4(p[i + 5], p[i + 6], p[i + 7], p[i + 8]));
                                }
#endif
                            }
                        }
                        if (h == 1 && j > 0) {
                            umma::tmem_st16(a_addr + 64, pk_);
                            umma::tmem_wait_st();
                            warp_arrive(bars + 8u * (t3::kBarAHi + t), lane);
                        }
                        if (st && (FIRST || !NERF_INTERLEAVE_STORES_BWD)) store_dz(dzp + t * kTileDzStride + h * kHalfDzStride, p);     // rows past `total` carry zeros
                    }
                }
                dzp -= 2 * kHalfDzStride;
                mkp -= 2 * kHalfMaskStride;
            };
            chain_step(std::true_type{}, 6);
#pragma unroll 1
            for (int j = 5; j >= 0; --j) chain_step(std::false_type{}, j);
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

int launch_mlp_tc_bwd3(const void* packed_t, const void* masks, const HeadGradSource& src, bool fused, int64_t total, void* dz_out,
                       void* stream) {
    static thread_local unsigned long long attr_mask = 0;
    if (attrs_pending(attr_mask)) {
        cudaError_t e = allow_smem(mlp_tc_bwd3_kernel<false>, b3::kSmemBytes);
        if (e == cudaSuccess) e = allow_smem(mlp_tc_bwd3_kernel<true>, b3::kSmemBytes);
        if (e != cudaSuccess) { set_error("nerf_mlp_backward_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return NERF_E_CUDA; }
        attrs_done(attr_mask);
    }
    const int64_t tiles = (total + t3::kTileM - 1) / t3::kTileM;
    const int64_t pairs = (tiles + 1) / 2;
    const int grid = (int)(pairs < num_sms() ? pairs : num_sms());
    if (fused)
        mlp_tc_bwd3_kernel<true><<<grid, t3::kThreads, b3::kSmemBytes, (cudaStream_t)stream>>>(
            (const uint8_t*)packed_t, (const uint32_t*)masks, src, total, (__nv_bfloat16*)dz_out);
    else
        mlp_tc_bwd3_kernel<false><<<grid, t3::kThreads, b3::kSmemBytes, (cudaStream_t)stream>>>(
            (const uint8_t*)packed_t, (const uint32_t*)masks, src, total, (__nv_bfloat16*)dz_out);
    return check_launch("nerf_mlp_backward_tc");
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_mlp_backward_tc(const void* packed_t, const void* masks, const float* dsigma_pre, const float* drgb_pre,
                                    int64_t N, int S, void* dz_out, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_mlp_backward_tc: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(packed_t && masks && dsigma_pre && drgb_pre && dz_out, "nerf_mlp_backward_tc: null pointer");
    NERF_REQUIRE(((uintptr_t)packed_t & 127) == 0 && ((uintptr_t)masks & 7) == 0 && ((uintptr_t)dz_out & 15) == 0,
                 "nerf_mlp_backward_tc: misaligned buffer");
    const HeadGradSource src{dsigma_pre, drgb_pre, nullptr, nullptr, nullptr, nullptr, S};
    return launch_mlp_tc_bwd3(packed_t, masks, src, false, N * S, dz_out, stream);
}

extern "C" int nerf_mlp_backward_tc_fused(const void* packed_t, const void* masks, const float* sigma, const float* rgb, const float* ts,
                                          const float* g_ray, int64_t N, int S, void* dz_out, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0 && S <= 1024, "nerf_mlp_backward_tc_fused: bad size (S <= 1024)");
    if (N == 0) return 0;
    NERF_REQUIRE(packed_t && masks && sigma && rgb && ts && g_ray && dz_out, "nerf_mlp_backward_tc_fused: null pointer");
    NERF_REQUIRE(((uintptr_t)packed_t & 127) == 0 && ((uintptr_t)masks & 7) == 0 && ((uintptr_t)dz_out & 15) == 0,
                 "nerf_mlp_backward_tc_fused: misaligned buffer");
    const HeadGradSource src{nullptr, nullptr, sigma, rgb, ts, g_ray, S};
    return launch_mlp_tc_bwd3(packed_t, masks, src, true, N * S, dz_out, stream);
}
