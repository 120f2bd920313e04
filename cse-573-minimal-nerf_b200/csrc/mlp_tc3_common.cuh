// mlp_tc3_common.cuh - pieces shared by the two-tile tensor-core kernels (forward mlp_tc3.cu, dgrad mlp_tc_bwd3.cu):
// shared-memory / barrier map, setmaxnreg helpers and the per-tile MMA issuer.
#pragma once
#include "mlp_tc_common.cuh"

namespace nerf {

namespace t3 {
constexpr int kTileM = 128;
constexpr int kSlots = 8;               // one K = 64 weight block per slot: a PE block or mlp.0's only block wastes nothing
constexpr uint32_t kSlotBytes = 16384;
constexpr int kThreads = 768;
constexpr int kEpiWarps = 16;          // every epilogue warp takes part in every task
constexpr int kPEWarps = 4;
constexpr uint32_t kColD = 0, kColA = 256;      // + 128 * tile

constexpr uint32_t kOffPE = 0;             // PE(x) tiles of X and Y: 2 x [128 x 64] bf16
constexpr uint32_t kOffPEDir = 32768;      // PE(dir) tiles of X and Y
constexpr uint32_t kOffRing = 65536;
constexpr uint32_t kOffBias = kOffRing + kSlots * kSlotBytes;
constexpr uint32_t kOffW7 = kOffBias + ((pk::kBiasFloats * 4 + 15) / 16) * 16;     // density_fn.0 weights, fp32 [256]
constexpr uint32_t kOffSig = kOffW7 + 1024;                                          // sigma partial sums [2 tiles][4][128]
// fused compositing (mlp_tc3.cu, COMP form): ring of kOutSlots tiles of (sigma, r, g, b) per sample, written by the last
// step's epilogue warps and consumed by the PE warps, which composite finished rays between two encodings
constexpr int kOutSlots = 8;
constexpr uint32_t kOffOut = kOffSig + 4096;                                         // kOutSlots x [128] float4
constexpr uint32_t kOffStat = kOffOut + kOutSlots * kTileM * 16;                     // per PE warp: sum sigma^2, count sigma != 0
constexpr uint32_t kOffBars = kOffStat + 64;
// barrier indices (8 bytes each)
constexpr uint32_t kBarFull = 0, kBarEmpty = 8, kBarDFull = 16, kBarDFree = 18, kBarALo = 20, kBarAHi = 22, kBarPexFull = 24,
                   kBarPexEmpty = 26, kBarPedFull = 28, kBarPedEmpty = 30, kBarTurn = 32, kBarOutFull = 34, kBarOutEmpty = 42,
                   kNumBars = 50;
constexpr uint32_t kOffTmemHolder = kOffBars + kNumBars * 8;
constexpr uint32_t kOffDetail = kOffTmemHolder + 16;                                 // PROFILE builds: 96 x int64
constexpr uint32_t kSmemBytes = kOffDetail + 96 * 8 + 1024;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// setmaxnreg moves registers inside the CTA's own pool (768 threads x 80 at launch): what the two small warpgroups give
// back is exactly what the four epilogue warpgroups take
constexpr int kRegsLaunch = 80, kRegsMisc = 64, kRegsEpi = 96, kRegsPE = 32;
static_assert(kRegsMisc + 4 * kRegsEpi + kRegsPE <= 6 * kRegsLaunch, "register pool budget");
}  // namespace t3

template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {     // TMEM side effects of this warp are done
    umma::tc_fence_before();
    __syncwarp();
    if (lane == 0) umma::mbar_arrive_u32(bar);
}

// One MMA-issuing warp PER TILE (warp 1: tile X, warp 2: tile Y).  Each runs the plain one-tile program - open stage,
// wait for its own tile's dependencies, issue 8 MMAs, release the stage - and the two instruction streams meet in the
// tensor pipe's queue: while one warp is between groups (barrier waits, bookkeeping; a single warp needs 150-300 clk for
// that and the queue only covers ~290 clk) the other warp's MMAs keep the pipe busy.  A weight stage is released when
// BOTH warps have committed it (empty barriers count 2), so every stage is still fetched once per tile pair.
// Every address is a 32-bit shared-memory / TMEM address held in a register.
// PROFILE counters: prof[1] wait weights, [2] wait dfree, [3] wait alo/ahi, [4] wait PE, [0] time inside issue blocks.
// MC: the weight ring is filled by multicast copies shared with the peer CTA of a 2-CTA cluster (mlp_tc3.cu); a slot may be
// refilled only when BOTH CTAs are done with it, so the stage-release commits go to both CTAs' empty barriers.
template <int T, bool PROFILE, bool MC = false>
struct MmaTile {
    static constexpr uint32_t kI128 = umma::make_idesc_bf16(128, 128);
    static constexpr uint32_t kI16 = umma::make_idesc_bf16(128, 16);
    uint32_t bars, tmem;
    uint64_t ring_desc;                             // descriptor of ring slot 0; slot s / K block at byte offset o: + (s*32768 + o) >> 4
    bool leader;
    uint32_t cnt;                                   // weight stages opened
    uint32_t n_dfree, n_ahi, n_alo, n_step;
    uint32_t detail, stage_in_pair;                 // PROFILE: shared-memory address of the per-stage wait table (0 = off)
    long long prof[5];

    __device__ __forceinline__ void init(uint32_t bars_addr, uint32_t ring_addr, uint32_t tmem_base, bool is_leader) {
        bars = bars_addr; tmem = tmem_base; leader = is_leader;
        ring_desc = umma::make_desc_k_sw128(ring_addr);
        cnt = 0; n_dfree = n_ahi = n_alo = n_step = 0;
        detail = 0; stage_in_pair = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) prof[i] = 0;
    }
    __device__ __forceinline__ uint32_t bar(uint32_t idx) const { return bars + 8u * idx; }
    __device__ __forceinline__ void wait(uint32_t bar_addr, uint32_t parity, int slot) {
        if (PROFILE) {
            const long long t0 = clock64();
            umma::mbar_wait_u32(bar_addr, parity);
            prof[slot] += clock64() - t0;
        } else {
            umma::mbar_wait_u32(bar_addr, parity);
        }
    }
    __device__ __forceinline__ void fence() { umma::tc_fence_after(); }     // once after a run of waits, before the MMAs
    // next weight stage present in shared memory; returns its descriptor offset (slot * 32768 >> 4)
    __device__ __forceinline__ uint32_t open() {
        const uint32_t slot = cnt & (t3::kSlots - 1);
        if (PROFILE && detail != 0u) {                  // per-stage wait histogram (stage index within the pair)
            const long long t0 = clock64();
            umma::mbar_wait_u32(bar(t3::kBarFull + slot), (cnt >> 3) & 1u);
            const long long dt = clock64() - t0;
            prof[1] += dt;
            if (leader) {
                const uint32_t a = detail + 8u * (stage_in_pair & 63u);
                long long old;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(old) : "r"(a));
                asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(old + dt));
            }
        } else {
            wait(bar(t3::kBarFull + slot), (cnt >> 3) & 1u, 1);
        }
        ++stage_in_pair;
        ++cnt;
        return slot * (t3::kSlotBytes >> 4);
    }
    __device__ __forceinline__ uint32_t empty_bar(uint32_t back) const { return bar(t3::kBarEmpty + ((cnt - 1u - back) & (t3::kSlots - 1))); }
    __device__ __forceinline__ void wait_pe(uint32_t idx, uint32_t parity) { wait(bar(idx + T), parity, 4); }
    __device__ __forceinline__ void begin_step() {      // the tile's accumulator has been read by its epilogue
        wait(bar(t3::kBarDFree + T), (n_dfree & 1u) ^ 1u, 2);
        ++n_dfree;
    }
    // Ping-pong between the two issuing warps: step n of tile X is queued before step n of tile Y, which is queued before
    // step n+1 of tile X.  Left alone the two warps fall into lockstep (both tiles in the same phase, both waiting for their
    // epilogues at the same time); with the turn token one tile's MMAs always cover the other tile's epilogue latency.
    __device__ __forceinline__ void my_turn() {
        if (T == 0) wait(bar(t3::kBarTurn + 1), (n_step & 1u) ^ 1u, 2);      // Y has queued step n-1
        else wait(bar(t3::kBarTurn + 0), n_step & 1u, 2);                     // X has queued step n
        ++n_step;
    }
    __device__ __forceinline__ void pass_turn() { umma::mbar_arrive_u32(bar(t3::kBarTurn + T)); }   // leader lane, after its MMAs
    __device__ __forceinline__ void release(uint32_t empty_bar_addr) {                              // weight stage consumed
        if (MC) umma::mma_commit_mc_u32(empty_bar_addr, (uint16_t)3);
        else umma::mma_commit_u32(empty_bar_addr);
    }
    // a hidden layer's second-half task stores K blocks 0,1 of the new operand (and waits for the store) BEFORE it signals
    // dfree: after begin_step() they are known to be written; alo is only signalled by rgb_fn.0's task (for rgb_fn.2)
    __device__ __forceinline__ void lo_implied() {}
    __device__ __forceinline__ void need_lo() { wait(bar(t3::kBarALo + T), n_alo & 1u, 3); ++n_alo; }
    __device__ __forceinline__ void need_hi() { wait(bar(t3::kBarAHi + T), n_ahi & 1u, 3); ++n_ahi; }
    // ---- unguarded pieces (callers hold the leader lane)
    __device__ __forceinline__ void mma8(uint32_t a_col, uint32_t b0, uint32_t b1, uint32_t first_acc) {   // two K=64 blocks (two stages)
        const uint32_t d = tmem + t3::kColD + 128u * T;
        const uint32_t a = tmem + t3::kColA + 128u * T + a_col;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const uint64_t bdesc = ring_desc + (kb ? b1 : b0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma::mma_ts(d, a + 32u * kb + 8u * k, bdesc + (uint64_t)(2 * k), kI128, (kb | k) ? 1u : first_acc);
        }
    }
    __device__ __forceinline__ void mma4(uint32_t a_col, uint32_t b, uint32_t first_acc) {          // one K=64 block (one stage)
        const uint32_t d = tmem + t3::kColD + 128u * T;
        const uint32_t a = tmem + t3::kColA + 128u * T + a_col;
        const uint64_t bdesc = ring_desc + b;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::mma_ts(d, a + 8u * k, bdesc + (uint64_t)(2 * k), kI128, k ? 1u : first_acc);
    }
    template <int NK16> __device__ __forceinline__ void mma_pe(uint64_t a_desc, uint32_t b_off) {   // A = shared-memory PE tile
        const uint32_t d = tmem + t3::kColD + 128u * T;
        const uint64_t bdesc = ring_desc + b_off;
#pragma unroll
        for (int k = 0; k < NK16; ++k) umma::mma_ss(d, a_desc + 2u * k, bdesc + 2u * k, kI128, k ? 1u : 0u);
    }
    template <class F> __device__ __forceinline__ void issue(F&& f) {       // one group of MMAs + commits by the leader lane
        long long t0 = 0;
        if (PROFILE) t0 = clock64();
        if (leader) f();
        __syncwarp();
        if (PROFILE) prof[0] += clock64() - t0;
    }
    // mlp.0 half: one 16 KB stage, A = PE(x)
    __device__ __forceinline__ void first_layer_half(uint64_t descPE) {
        const uint32_t b = open();
        begin_step();
        my_turn();
        fence();
        issue([&] {
            mma_pe<4>(descPE, b);
            release(empty_bar(0));
            umma::mma_commit_u32(bar(t3::kBarDFull + T));
            pass_turn();
        });
    }
    // a K = 256 layer half (or rgb_fn.0): [PE stage] kb01 kb23.  A second half has no new A operand to wait for and issues
    // its 16 MMAs back to back.
    template <bool FIRST_HALF, int PE_K16>
    __device__ __forceinline__ void layer_half(uint64_t descA, uint32_t pe_done_idx) {
        constexpr bool PE = PE_K16 > 0;
        constexpr int NK = PE ? PE_K16 : 1;
        uint32_t bP = 0, e_p = 0;
        if (PE) { bP = open(); e_p = empty_bar(0); }
        const uint32_t k0 = open(), e_0 = empty_bar(0);
        const uint32_t k1 = open(), e_1 = empty_bar(0);
        uint32_t k2 = 0, k3 = 0, e_2 = 0, e_3 = 0;
        if (!FIRST_HALF) { k2 = open(); e_2 = empty_bar(0); k3 = open(); e_3 = empty_bar(0); }
        begin_step();
        if (FIRST_HALF) lo_implied();
        my_turn();
        fence();
        issue([&] {
            if (PE) {
                mma_pe<NK>(descA, bP);
                if (pe_done_idx) umma::mma_commit_u32(bar(pe_done_idx + T));
                release(e_p);
            }
            mma8(0, k0, k1, PE ? 1u : 0u);
            release(e_0);
            release(e_1);
            if (!FIRST_HALF) {
                mma4(64, k2, 1u);
                release(e_2);
                pass_turn();            // early: the other issuer needs ~200 clk to wake up; this tile's last 4 MMAs cover it
                mma4(96, k3, 1u);
                release(e_3);
                umma::mma_commit_u32(bar(t3::kBarDFull + T));
            }
        });
        if (FIRST_HALF) {               // K blocks 2,3 of the new operand are written ~500 clk after K blocks 0,1
            k2 = open(); e_2 = empty_bar(0);
            k3 = open(); e_3 = empty_bar(0);
            need_hi();
            fence();
            issue([&] {
                mma4(64, k2, 1u);
                release(e_2);
                pass_turn();
                mma4(96, k3, 1u);
                release(e_3);
                umma::mma_commit_u32(bar(t3::kBarDFull + T));
            });
        }
    }
    // rgb_fn.2: r (A columns 0..63) -> 16 columns, one 4 KB stage
    __device__ __forceinline__ void last_step() {
        const uint32_t b = open();
        begin_step();
        need_lo();
        my_turn();
        fence();
        issue([&] {
            const uint32_t d = tmem + t3::kColD + 128u * T;
            const uint32_t a = tmem + t3::kColA + 128u * T;
            const uint64_t bdesc = ring_desc + b;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma::mma_ts(d, a + 32u * kb + 8u * k, bdesc + (uint64_t)(kb * 128 + 2 * k), kI16, (kb | k) ? 1u : 0u);
            }
            release(empty_bar(0));
            umma::mma_commit_u32(bar(t3::kBarDFull + T));
            pass_turn();
        });
    }
    // ---- dgrad (mlp_tc_bwd3.cu): rgb_fn.0 half with A = this tile's dr tile in shared memory (K = 128: both K blocks of a stage)
    __device__ __forceinline__ void dr_half(uint64_t descDr, bool last) {
        const uint32_t b0 = open(), e_0 = empty_bar(0);
        const uint32_t b1 = open(), e_1 = empty_bar(0);
        begin_step();
        my_turn();
        fence();
        issue([&] {
            const uint32_t d = tmem + t3::kColD + 128u * T;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                const uint64_t bdesc = ring_desc + (kb ? b1 : b0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma::mma_ss(d, descDr + (uint64_t)(kb * 1024 + 2 * k), bdesc + (uint64_t)(2 * k), kI128, (kb | k) ? 1u : 0u);
            }
            release(e_0);
            release(e_1);
            if (last) umma::mma_commit_u32(bar(t3::kBarPexEmpty + T));          // dr tile no longer read
            umma::mma_commit_u32(bar(t3::kBarDFull + T));
            pass_turn();
        });
    }
    // dgrad program of tile T: dz6 from dr, then feature_fn.4, .2, .0 (h part), mlp.6, mlp.4, mlp.2 with W^T stages
    __device__ __forceinline__ void run_bwd(uint32_t sbase, int64_t num_pairs) {
        const uint64_t descDr = umma::make_desc_k_sw128(sbase + T * 32768);
        uint32_t it = 0;
        for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x, ++it) {
            wait_pe(t3::kBarPexFull, it & 1u);
            dr_half(descDr, false);
            dr_half(descDr, true);
#pragma unroll 1
            for (int l = 0; l < 6; ++l) {
                layer_half<true, 0>(0, 0);
                layer_half<false, 0>(0, 0);
            }
        }
    }
    // the whole per-CTA program of tile T: n_local tile pairs
    __device__ __forceinline__ void run(uint32_t sbase, uint32_t n_local) {
        const uint64_t descPE = umma::make_desc_k_sw128(sbase + t3::kOffPE + T * 16384);
        const uint64_t descPD = umma::make_desc_k_sw128(sbase + t3::kOffPEDir + T * 16384);
        for (uint32_t it = 0; it < n_local; ++it) {
            stage_in_pair = 0;
            wait_pe(t3::kBarPexFull, it & 1u);
            first_layer_half(descPE);                                  // mlp.0
            first_layer_half(descPE);
#pragma unroll 1
            for (int l = 1; l <= 6; ++l) {                             // mlp.2/4/6, feature_fn.0 (PE(x) K block first), feature_fn.2/4
                if (l == 4) {
                    layer_half<true, 4>(descPE, 0);
                    layer_half<false, 4>(descPE, t3::kBarPexEmpty);    // last reader of PE(x)
                } else {
                    layer_half<true, 0>(0, 0);
                    layer_half<false, 0>(0, 0);
                }
            }
            wait_pe(t3::kBarPedFull, it & 1u);
            layer_half<true, 2>(descPD, t3::kBarPedEmpty);             // rgb_fn.0: PE(dir) K block + feat
            last_step();                                               // rgb_fn.2
        }
    }
};

}  // namespace nerf
