// samplers.cu - the HBM-bound kernels of the NeRF hot path: ray generation, stratified sampling,
// compositing, inverse-CDF fine sampling and the per-ray merge sort.
//
// Arithmetic contract: the reference runs these as eager fp32 torch ops on CPU, i.e.
// separately rounded multiplies/adds, true divisions and *sequential* running sums.  Every operation that
// feeds a sample depth or an index is therefore spelled with __f{add,sub,mul,div}_rn so nvcc cannot
// contract it into an FMA, and running sums are taken in index order.  Given the same uniforms the
// depths, points and bin indices are bit-identical to the oracle.
//
// Mapping: one warp per ray for everything that needs a running sum, a search or a sort (coalesced
// 128-byte row loads, the serial chain broadcast through shuffles), one thread per element otherwise.
#include "common.cuh"
#include "composite_scan.cuh"
#include "fine_sampler.cuh"

namespace nerf {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * kWarp;

// ------------------------------------------------------------------------------------------ K0 raygen
struct Pose { float r[3][3]; float t[3]; };

__device__ __forceinline__ void raygen_one(const Pose& pose, int W, float half_w, float half_h, float focal,
                                           const int64_t* __restrict__ xs, const int64_t* __restrict__ ys, int64_t i, float (&dir)[3]) {
    float col, row;
    if (xs) { col = (float)xs[i]; row = (float)ys[i]; }
    else    { col = (float)(i % W); row = (float)(i / W); }
    // dataloader.py:39: [(i - W/2)/focal, -(j - H/2)/focal, -1], true fp32 divisions
    const float d0 = __fdiv_rn(__fsub_rn(col, half_w), focal);
    const float d1 = -__fdiv_rn(__fsub_rn(row, half_h), focal);
    const float d2 = -1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k)   // dataloader.py:40: sum_j dirs_j * c2w[k,j], left to right
        dir[k] = __fadd_rn(__fadd_rn(__fmul_rn(d0, pose.r[k][0]), __fmul_rn(d1, pose.r[k][1])), __fmul_rn(d2, pose.r[k][2]));
}

// One thread per FOUR consecutive rays: their 12 direction floats (and the 12 floats of the repeated origin) leave as three
// 16-byte stores each, consecutive threads writing consecutive 48-byte runs (a 4-byte store per component touched every
// sector three times: 0.6 TB/s; this form is bound by the launch, not the stores).  VEC = false: scalar tail / unaligned bases.
template <bool VEC>
__global__ void __launch_bounds__(256)
raygen_kernel(Pose pose, int H, int W, float half_w, float half_h, float focal,
              const int64_t* __restrict__ xs, const int64_t* __restrict__ ys, int64_t first, int64_t n,
              float* __restrict__ o, float* __restrict__ d) {
    if (VEC) {
        const int64_t quads = (n - first) >> 2;
        for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
            const int64_t i0 = first + 4 * q;
            float v[12];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float dir[3];
                raygen_one(pose, W, half_w, half_h, focal, xs, ys, i0 + r, dir);
                v[3 * r] = dir[0]; v[3 * r + 1] = dir[1]; v[3 * r + 2] = dir[2];
            }
            float4* dp = (float4*)(d + i0 * 3);
            float4* op = (float4*)(o + i0 * 3);
            dp[0] = make_float4(v[0], v[1], v[2], v[3]);
            dp[1] = make_float4(v[4], v[5], v[6], v[7]);
            dp[2] = make_float4(v[8], v[9], v[10], v[11]);
            op[0] = make_float4(pose.t[0], pose.t[1], pose.t[2], pose.t[0]);
            op[1] = make_float4(pose.t[1], pose.t[2], pose.t[0], pose.t[1]);
            op[2] = make_float4(pose.t[2], pose.t[0], pose.t[1], pose.t[2]);
        }
    } else {
        for (int64_t i = first + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            float dir[3];
            raygen_one(pose, W, half_w, half_h, focal, xs, ys, i, dir);
#pragma unroll
            for (int k = 0; k < 3; ++k) { d[i * 3 + k] = dir[k]; o[i * 3 + k] = pose.t[k]; }
        }
    }
}

// N1, the training batch producer in graph-safe form (dataloader.py:143-152 for a pixel list): which image the batch comes from
// is read from DEVICE memory (a CUDA graph replays this launch with a new index each step), the pose from a device table, and
// the target colours are gathered from the device-resident uint8 images: rgb = fl32(u8 / 255) evaluated in double precision
// like numpy's `imread(...) / 255` upstream (dataloader.py:148).  One thread per ray.
__global__ void __launch_bounds__(256)
batch_rays_kernel(const float* __restrict__ poses, const int64_t* __restrict__ img_idx, const uint8_t* __restrict__ images, int n_img,
                  int H, int W, float half_w, float half_h, float focal, const int64_t* __restrict__ xs, const int64_t* __restrict__ ys,
                  int64_t n, float* __restrict__ o, float* __restrict__ d, float* __restrict__ rgb) {
    int64_t im = *img_idx;
    im = im < 0 ? 0 : (im >= n_img ? n_img - 1 : im);
    Pose pose;
    const float* c2w = poses + im * 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) pose.r[r][c] = __ldg(c2w + r * 4 + c);
        pose.t[r] = __ldg(c2w + r * 4 + 3);
    }
    const uint8_t* img = images + im * (int64_t)H * W * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float dir[3];
        raygen_one(pose, W, half_w, half_h, focal, xs, ys, i, dir);
        const int64_t x = xs[i], y = ys[i];
        const bool inside = x >= 0 && x < W && y >= 0 && y < H;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            d[i * 3 + k] = dir[k];
            o[i * 3 + k] = pose.t[k];
            rgb[i * 3 + k] = inside ? (float)((double)img[(y * W + x) * 3 + k] / 255.0) : 0.f;
        }
    }
}

// --------------------------------------------------------------------------------- K1 coarse sampling
__global__ void __launch_bounds__(256)
coarse_sample_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ u,
                     const float* __restrict__ t_base, float step, int64_t total, int C,
                     float* __restrict__ samples, float* __restrict__ ts) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = e / C;
        const int i = (int)(e - n * C);
        // nerf_helpers.py:52-53: ts = arange + rand * step
        const float t = __fadd_rn(__ldg(t_base + i), __fmul_rn(u[e], step));
        ts[e] = t;
        if (samples) {
#pragma unroll
            for (int k = 0; k < 3; ++k)   // nerf_helpers.py:55: d * t + o
                samples[e * 3 + k] = __fadd_rn(__fmul_rn(__ldg(d + n * 3 + k), t), __ldg(o + n * 3 + k));
        }
    }
}

// ------------------------------------------------------------------------------------------ K2 pieces
__global__ void __launch_bounds__(256)
deltas_kernel(const float* __restrict__ ts, int64_t total, int S, float* __restrict__ deltas) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(e % S);
        deltas[e] = (i == S - 1) ? 1e10f : __fsub_rn(ts[e + 1], ts[e]);   // nerf_helpers.py:71-72
    }
}

// One warp per ray.  kind 0: weights from (sigma, deltas) [nerf_weights]; kind 1: full composite from
// (sigma, rgb, ts).
template <int KIND>
__global__ void __launch_bounds__(kThreads)
composite_kernel(const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ ts_or_deltas,
                 int64_t N, int S, float* __restrict__ deltas_out, float* __restrict__ weights_out,
                 float* __restrict__ ray_rgb, float* __restrict__ depth, float* __restrict__ acc,
                 float* __restrict__ stats2) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    float st_sq = 0.f, st_nz = 0.f;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); n < N; n += warps) {
        const float* sg = sigma + n * S;
        const float* tp = ts_or_deltas + n * S;
        const float* cp = rgb + n * S * 3;
        float running = 0.f;       // sum_{j<i} -sigma_j delta_j, nerf_helpers.py:86-89
        float cr = 0.f, cg = 0.f, cb = 0.f, dsum = 0.f, asum = 0.f;
        // software pipeline: the next chunk's sigma / t / rgb are in flight while this chunk's serial running sum is taken
        float s_n = 0.f, t_n = 0.f, c0_n = 0.f, c1_n = 0.f, c2_n = 0.f;
        auto fetch = [&](int base) {
            const int i = base + lane;
            const bool in = i < S;
            s_n = in ? __ldg(sg + i) : 0.f;
            t_n = in ? __ldg(tp + i) : 0.f;
            if (KIND == 1 && in) { c0_n = __ldg(cp + i * 3); c1_n = __ldg(cp + i * 3 + 1); c2_n = __ldg(cp + i * 3 + 2); }
        };
        fetch(0);
        for (int base = 0; base < S; base += kWarp) {
            const int i = base + lane;
            const bool in = i < S;
            const float s = s_n, tv = t_n, c0 = c0_n, c1 = c1_n, c2 = c2_n;
            if (base + kWarp < S) fetch(base + kWarp);
            float t = 0.f, dl = 0.f;
            if (KIND == 1) {
                t = tv;
                float tn = __shfl_down_sync(kFull, t, 1);
                const float t_first_next = __shfl_sync(kFull, t_n, 0);      // first depth of the next chunk (already fetched)
                if (lane == 31 && i + 1 < S) tn = t_first_next;
                dl = (i == S - 1) ? 1e10f : __fsub_rn(tn, t);             // nerf_helpers.py:71-72
            } else {
                dl = tv;
            }
            const float x = in ? __fmul_rn(__fmul_rn(-1.0f, s), dl) : 0.f;   // -1 * density * deltas
            const float excl = (KIND == 1) ? chunk_exclusive_scan_tree(x, running, lane) : chunk_exclusive_scan(x, running, lane);
            if (in) {
                const float trans = expf(excl);                               // nerf_helpers.py:89
                const float w = __fmul_rn(__fsub_rn(1.0f, expf(x)), trans);   // nerf_helpers.py:90
                if (weights_out) weights_out[n * S + i] = w;
                if (KIND == 1) {
                    if (deltas_out) deltas_out[n * S + i] = dl;
                    cr = fmaf(w, c0, cr); cg = fmaf(w, c1, cg); cb = fmaf(w, c2, cb);
                    dsum = fmaf(w, t, dsum); asum += w;
                    st_sq = fmaf(s, s, st_sq); st_nz += (s != 0.f) ? 1.f : 0.f;
                }
            }
        }
        if (KIND == 1) {
            cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
            dsum = warp_sum(dsum); asum = warp_sum(asum);
            if (lane == 0) {
                if (ray_rgb) { ray_rgb[n * 3 + 0] = cr; ray_rgb[n * 3 + 1] = cg; ray_rgb[n * 3 + 2] = cb; }
                if (depth) depth[n] = dsum;
                if (acc) acc[n] = asum;
            }
        }
    }
    if (KIND == 1 && stats2) {
        st_sq = warp_sum(st_sq); st_nz = warp_sum(st_nz);
        __shared__ float red[2][kWarpsPerBlock];
        if (lane == 0) { red[0][threadIdx.x >> 5] = st_sq; red[1][threadIdx.x >> 5] = st_nz; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int k = 0; k < kWarpsPerBlock; ++k) { a += red[0][k]; b += red[1][k]; }
            atomicAdd(stats2 + 0, a);
            atomicAdd(stats2 + 1, b);
            // the last block to arrive publishes the norm the reference logs (nerf_model.py:105,124): stats[2] = sqrt(stats[0])
            __threadfence();
            const unsigned ticket = atomicAdd((unsigned*)(stats2 + 3), 1u);
            if (ticket == gridDim.x - 1) {
                __threadfence();
                stats2[2] = sqrtf(atomicAdd(stats2 + 0, 0.f));
            }
        }
    }
}

// Thread-per-ray form of the full composite (KIND 1 above), used by nerf_composite.  The running sum is sequential in the
// sample index (the reference's CPU cumsum order), which a warp-per-ray kernel can only emulate with a 32-step shuffle chain
// per 32 samples: ~8 warp instructions per sample, issue-bound at a third of the HBM roofline.  Here a block stages a
// [64 rays x 32 samples] tile of sigma / t / rgb through shared memory with coalesced 128-byte rows, then every thread walks
// ITS ray's 32 samples in order: the same arithmetic in the same order, ~1.5 warp instructions per sample.
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
constexpr int kCompRays = 64;          // rays (= threads) per block
constexpr int kCompPad = 33;           // row stride of the staged tiles (bank-conflict free column walks)
__global__ void __launch_bounds__(kCompRays)
composite_rays_kernel(const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ ts, int64_t N, int S,
                      float* __restrict__ deltas_out, float* __restrict__ weights_out, float* __restrict__ ray_rgb,
                      float* __restrict__ depth, float* __restrict__ acc, float* __restrict__ stats2) {
    __shared__ float sS[kCompRays * kCompPad], sT[kCompRays * (kCompPad + 1)], sC[kCompRays * 3 * 32 + kCompRays];      // weights overwrite sS in place
    const int tid = threadIdx.x;
    float st_sq = 0.f, st_nz = 0.f;
    for (int64_t n0 = (int64_t)blockIdx.x * kCompRays; n0 < N; n0 += (int64_t)gridDim.x * kCompRays) {
        const int rays = (int)((N - n0 < kCompRays) ? (N - n0) : kCompRays);
        const int64_t n = n0 + tid;
        float running = 0.f;       // sum_{j<i} -sigma_j delta_j, nerf_helpers.py:86-89
        float cr = 0.f, cg = 0.f, cb = 0.f, dsum = 0.f, asum = 0.f;
        for (int base = 0; base < S; base += 32) {
            const int len = (S - base < 32) ? (S - base) : 32;
            // ---- stage the tile with cp.async (the whole 42 KB tile is in flight at once): consecutive threads read
            // consecutive samples of one ray (128-byte rows)
            for (int rr = tid >> 5; rr < rays; rr += kCompRays / 32) {      // one ray per warp iteration, no index divisions
                const int lane = tid & 31;
                const int64_t off = (n0 + rr) * S + base;
                if (lane < len) cp_async4(sS + rr * kCompPad + lane, sigma + off + lane);
                if (lane < len) cp_async4(sT + rr * (kCompPad + 1) + lane, ts + off + lane);
                if (lane == 0 && base + 32 < S) cp_async4(sT + rr * (kCompPad + 1) + 32, ts + off + 32);   // delta of the chunk's last sample
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int i = lane + 32 * k;
                    if (i < len * 3) cp_async4(sC + rr * 97 + i, rgb + off * 3 + i);
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            if (tid < rays) {
                float* ps = sS + tid * kCompPad;
                const float* pt = sT + tid * (kCompPad + 1);
                const float* pc = sC + tid * 97;
                float* pw = ps;
                for (int i = 0; i < len; ++i) {
                    const float s = ps[i], t = pt[i];
                    const float dl = (base + i == S - 1) ? 1e10f : __fsub_rn(pt[i + 1], t);      // nerf_helpers.py:71-72
                    const float x = __fmul_rn(__fmul_rn(-1.0f, s), dl);                            // -1 * density * deltas
                    const float trans = expf(running);                                             // nerf_helpers.py:89
                    const float w = __fmul_rn(__fsub_rn(1.0f, expf(x)), trans);                    // nerf_helpers.py:90
                    running = __fadd_rn(running, x);
                    pw[i] = w;
                    if (deltas_out) deltas_out[n * S + base + i] = dl;
                    cr = fmaf(w, pc[3 * i], cr); cg = fmaf(w, pc[3 * i + 1], cg); cb = fmaf(w, pc[3 * i + 2], cb);
                    dsum = fmaf(w, t, dsum); asum += w;
                    st_sq = fmaf(s, s, st_sq); st_nz += (s != 0.f) ? 1.f : 0.f;
                }
            }
            __syncthreads();
            if (weights_out) {
                for (int e = tid; e < rays * 32; e += kCompRays) {
                    const int rr = e >> 5, i = e & 31;
                    if (i < len) weights_out[(n0 + rr) * S + base + i] = sS[rr * kCompPad + i];
                }
            }
            __syncthreads();           // the next chunk's staging overwrites the tiles
        }
        if (tid < rays) {
            if (ray_rgb) { ray_rgb[n * 3 + 0] = cr; ray_rgb[n * 3 + 1] = cg; ray_rgb[n * 3 + 2] = cb; }
            if (depth) depth[n] = dsum;
            if (acc) acc[n] = asum;
        }
    }
    if (stats2) {
        st_sq = warp_sum(st_sq); st_nz = warp_sum(st_nz);
        __shared__ float red[2][kCompRays / 32];
        if ((tid & 31) == 0) { red[0][tid >> 5] = st_sq; red[1][tid >> 5] = st_nz; }
        __syncthreads();
        if (tid == 0) {
            float a = 0.f, b = 0.f;
            for (int k = 0; k < kCompRays / 32; ++k) { a += red[0][k]; b += red[1][k]; }
            atomicAdd(stats2 + 0, a);
            atomicAdd(stats2 + 1, b);
            // the last block to arrive publishes the norm the reference logs (nerf_model.py:105,124): stats[2] = sqrt(stats[0])
            __threadfence();
            const unsigned ticket = atomicAdd((unsigned*)(stats2 + 3), 1u);
            if (ticket == gridDim.x - 1) {
                __threadfence();
                stats2[2] = sqrtf(atomicAdd(stats2 + 0, 0.f));
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads)
ray_color_kernel(const float* __restrict__ w, const float* __restrict__ rgb, int64_t N, int S, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); n < N; n += warps) {
        float cr = 0.f, cg = 0.f, cb = 0.f;
        for (int i = lane; i < S; i += kWarp) {
            const float wi = w[n * S + i];
            const float* c = rgb + (n * S + i) * 3;
            cr = fmaf(wi, c[0], cr); cg = fmaf(wi, c[1], cg); cb = fmaf(wi, c[2], cb);
        }
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
        if (lane == 0) { out[n * 3 + 0] = cr; out[n * 3 + 1] = cg; out[n * 3 + 2] = cb; }
    }
}

// ----------------------------------------------------------------------------------- K3 fine sampling
// One warp per ray.  Dynamic shared memory per warp: cdf[C] then bounds[C+2].
__global__ void __launch_bounds__(kThreads)
fine_sample_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ w,
                   const float* __restrict__ ts, const float* __restrict__ eps, const float* __restrict__ u,
                   const float* __restrict__ q_base, int64_t N, int C, int F, float near_, float far_,
                   float* __restrict__ fine_samples, float* __restrict__ fine_ts, int64_t* __restrict__ idx_out) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* cdf = smem + (size_t)wib * (2 * C + 2);
    float* bounds = cdf + C;
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    const float Ff = (float)F;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + wib; n < N; n += warps) {
        // inclusive sequential cumsum of the weights (nerf_helpers.py:137)
        float running = 0.f;
        for (int base = 0; base < C; base += kWarp) {
            const int i = base + lane;
            const float x = (i < C) ? w[n * C + i] : 0.f;
            const float excl = chunk_exclusive_scan(x, running, lane);
            if (i < C) {
                cdf[i] = __fadd_rn(excl, x);
                bounds[i + 1] = ts[n * C + i];                                  // nerf_helpers.py:149
            }
        }
        if (lane == 0) { bounds[0] = near_; bounds[C + 1] = far_; }
        __syncwarp();
        const float total = cdf[C - 1];
        __syncwarp();
        for (int i = lane; i < C; i += kWarp) cdf[i] = __fdiv_rn(cdf[i], total);   // nerf_helpers.py:138
        __syncwarp();
        const float e = __fdiv_rn(eps[n], Ff);                                      // nerf_helpers.py:139
        const float ox = o[n * 3], oy = o[n * 3 + 1], oz = o[n * 3 + 2];
        const float dx = d[n * 3], dy = d[n * 3 + 1], dz = d[n * 3 + 2];
        for (int j = lane; j < F; j += kWarp) {
            const float q = __fadd_rn(__ldg(q_base + j), e);                         // nerf_helpers.py:142
            int lo = 0, hi = C;                                                       // torch.searchsorted, right=False
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                if (!(cdf[mid] >= q)) lo = mid + 1; else hi = mid;
            }
            const float b0 = bounds[lo], b1 = bounds[lo + 1];
            const float t = __fadd_rn(b0, __fmul_rn(__fsub_rn(b1, b0), u[n * F + j]));   // nerf_helpers.py:154
            fine_ts[n * F + j] = t;
            if (idx_out) idx_out[n * F + j] = lo;
            if (fine_samples) {                                                          // nerf_helpers.py:155
                float* p = fine_samples + (n * F + j) * 3;
                p[0] = __fadd_rn(ox, __fmul_rn(t, dx));
                p[1] = __fadd_rn(oy, __fmul_rn(t, dy));
                p[2] = __fadd_rn(oz, __fmul_rn(t, dz));
            }
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------- K4 merge sort
// Fast path (A + B <= 256): one warp per ray, the ray's depths padded with +inf to 256 and sorted by a bitonic network held
// in registers: element e = 32 j + lane lives in register j of that lane, so compare distances >= 32 are register-to-register
// and distances < 32 are one __shfl_xor each.  Sorting values only - equal depths are interchangeable - so the result is the
// same multiset order torch.sort produces.
__global__ void __launch_bounds__(kThreads)
merge_sort256_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ ts_a, int A,
                     const float* __restrict__ ts_b, int B, int64_t N, float* __restrict__ ts_sorted,
                     float* __restrict__ samples_sorted) {
    const int S = A + B;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); n < N; n += warps) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = 32 * j + lane;
            v[j] = (e < A) ? ts_a[n * A + e] : (e < S ? ts_b[n * B + (e - A)] : __int_as_float(0x7f800000));   // nerf_model.py:117
        }
#pragma unroll
        for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
            for (int dist = k >> 1; dist >= 1; dist >>= 1) {
                if (dist >= 32) {
                    const int dj = dist >> 5;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if ((j & dj) == 0) {
                            const bool up = (((32 * j + lane) & k) == 0);
                            cmpx(v[j], v[j | dj], up);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int e = 32 * j + lane;
                        const float other = __shfl_xor_sync(kFull, v[j], dist);
                        const bool up = ((e & k) == 0);
                        const bool lower = ((lane & dist) == 0);          // this lane holds the lower-index element of the pair
                        v[j] = (up == lower) ? fminf(v[j], other) : fmaxf(v[j], other);
                    }
                }
            }
        }
        const float ox = o ? o[n * 3] : 0.f, oy = o ? o[n * 3 + 1] : 0.f, oz = o ? o[n * 3 + 2] : 0.f;
        const float dx = d ? d[n * 3] : 0.f, dy = d ? d[n * 3 + 1] : 0.f, dz = d ? d[n * 3 + 2] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = 32 * j + lane;
            if (e < S) {
                const float t = v[j];
                ts_sorted[n * S + e] = t;
                if (samples_sorted) {
                    float* p = samples_sorted + (n * S + e) * 3;
                    p[0] = __fadd_rn(ox, __fmul_rn(t, dx));
                    p[1] = __fadd_rn(oy, __fmul_rn(t, dy));
                    p[2] = __fadd_rn(oz, __fmul_rn(t, dz));
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- K3 + K4 in one launch
// What NeRFNetwork.forward does between the two networks (nerf_model.py:114-120): inverse-CDF fine depths, concatenated
// with the coarse depths (fine first) and sorted - without the round trip of the fine depths through HBM.  One warp per
// ray; the same cdf / search / jitter arithmetic as fine_sample_kernel.
//
// The sort uses what is known about the two halves instead of a 256-wide network over everything: the C coarse depths
// arrive SORTED (stratified: t_i lies in stratum i), so only the F <= 128 fine depths go through a bitonic network (128
// wide: 28 compare stages over 4 registers instead of 36 over 8), and the two sorted runs are merged by RANK - a fine
// depth's slot is its index plus the number of coarse depths below it, a coarse depth's slot its index plus the number of
// fine depths not above it (binary searches in shared memory; ties go fine-first, so the slots are a permutation).  A sorted
// array is unique, so the result is bit-identical to sorting all C + F values (tests/test_gpu_fused_composite.py, incl. rays
// whose cdf is NaN).  Rays whose coarse depths are NOT sorted (a caller's own depths), a NaN depth, or F > 128 take the
// general path: the 256-wide network of merge_sort256_kernel over the concatenation.
__global__ void __launch_bounds__(kThreads, 8)          // 32 registers: 64 resident warps per SM (a per-ray latency chain)
fine_sample_merge_kernel(const float* __restrict__ w, const float* __restrict__ ts, const float* __restrict__ eps,
                         const float* __restrict__ u, const float* __restrict__ q_base, int64_t N, int C, int F,
                         float near_, float far_, float* __restrict__ ts_sorted) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* scratch = smem + (size_t)wib * fine_sampler_floats(C);
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + wib; n < N; n += warps) {
        for (int i = lane; i < C; i += kWarp) {
            scratch[i] = w[n * C + i];                                           // raw weights -> cdf slots
            scratch[C + 1 + i] = ts[n * C + i];                                  // coarse depths -> bounds[1 .. C] (nerf_helpers.py:149)
        }
        __syncwarp();
        fine_sample_merge_ray<true>(scratch, C, F, near_, far_, eps[n], u + n * F, q_base, ts_sorted + n * (C + F), lane);
    }
}


// General path (A + B <= 1024): one warp per ray, enumeration sort in shared memory: rank(e) = #{j : v_j < v_e or (v_j == v_e and j < e)}.
// Dynamic shared memory per warp: vals[S] then sorted[S].
__global__ void __launch_bounds__(kThreads)
merge_sort_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ ts_a, int A,
                  const float* __restrict__ ts_b, int B, int64_t N, float* __restrict__ ts_sorted,
                  float* __restrict__ samples_sorted) {
    extern __shared__ float smem[];
    const int S = A + B;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* vals = smem + (size_t)wib * 2 * S;
    float* sorted = vals + S;
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + wib; n < N; n += warps) {
        for (int i = lane; i < A; i += kWarp) vals[i] = ts_a[n * A + i];           // nerf_model.py:117 (fine first)
        for (int i = lane; i < B; i += kWarp) vals[A + i] = ts_b[n * B + i];
        __syncwarp();
        for (int e = lane; e < S; e += kWarp) {
            const float v = vals[e];
            int rank = 0;
            for (int j = 0; j < S; ++j) {
                const float x = vals[j];
                rank += (x < v || (x == v && j < e)) ? 1 : 0;
            }
            sorted[rank] = v;
        }
        __syncwarp();
        const float ox = o ? o[n * 3] : 0.f, oy = o ? o[n * 3 + 1] : 0.f, oz = o ? o[n * 3 + 2] : 0.f;
        const float dx = d ? d[n * 3] : 0.f, dy = d ? d[n * 3 + 1] : 0.f, dz = d ? d[n * 3 + 2] : 0.f;
        for (int i = lane; i < S; i += kWarp) {
            const float t = sorted[i];
            ts_sorted[n * S + i] = t;
            if (samples_sorted) {
                float* p = samples_sorted + (n * S + i) * 3;
                p[0] = __fadd_rn(ox, __fmul_rn(t, dx));
                p[1] = __fadd_rn(oy, __fmul_rn(t, dy));
                p[2] = __fadd_rn(oz, __fmul_rn(t, dz));
            }
        }
        __syncwarp();
    }
}



// ------------------------------------------------------------------------------ compositing backward
// One warp per ray (composite_backward_ray, composite_scan.cuh); the dgrad kernel runs the same routine in its producer warps
// (mlp_tc_bwd3.cu, FUSED form), this launch serves nerf_composite_backward.
__global__ void __launch_bounds__(kThreads)
composite_backward_kernel(const float* __restrict__ sigma, const float* __restrict__ rgb, const float* __restrict__ ts,
                          const float* __restrict__ g_ray, int64_t N, int S, float* __restrict__ dsigma_pre,
                          float* __restrict__ drgb_pre) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t n = blockIdx.x * (int64_t)kWarpsPerBlock + (threadIdx.x >> 5); n < N; n += warps) {
        composite_backward_ray(sigma, rgb, ts, g_ray, n, S, lane, [&](int i, float ds, float d0, float d1, float d2) {
            dsigma_pre[n * S + i] = ds;
            float* o = drgb_pre + (n * S + i) * 3;
            o[0] = d0; o[1] = d1; o[2] = d2;
        });
    }
}

static int grid_for(int64_t items, int per_block) {
    int64_t blocks = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_raygen(const float* c2w_host, int H, int W, float focal, const int64_t* xs, const int64_t* ys,
                           int64_t n, float* o, float* d, void* stream) {
    NERF_REQUIRE(c2w_host && o && d, "nerf_raygen: null pointer");
    NERF_REQUIRE(H > 0 && W > 0 && n >= 0, "nerf_raygen: bad size H=%d W=%d n=%lld", H, W, (long long)n);
    NERF_REQUIRE((xs == nullptr) == (ys == nullptr), "nerf_raygen: xs and ys must both be given or both be NULL");
    NERF_REQUIRE(xs || n == (int64_t)H * W, "nerf_raygen: full-grid mode needs n == H*W");
    if (n == 0) return 0;
    Pose p;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) p.r[r][c] = c2w_host[r * 4 + c];
        p.t[r] = c2w_host[r * 4 + 3];
    }
    // four rays per thread with 16-byte stores where the output bases allow it, the (< 4 ray) remainder one ray per thread
    const bool vec = ((((uintptr_t)o) | ((uintptr_t)d)) & 15) == 0 && n >= 4;
    const int64_t n_vec = vec ? (n & ~(int64_t)3) : 0;
    if (n_vec)
        raygen_kernel<true><<<grid_for(n_vec / 4, 256), 256, 0, (cudaStream_t)stream>>>(p, H, W, (float)(W * .5), (float)(H * .5), focal,
                                                                                     xs, ys, 0, n_vec, o, d);
    if (n_vec < n)
        raygen_kernel<false><<<grid_for(n - n_vec, 256), 256, 0, (cudaStream_t)stream>>>(p, H, W, (float)(W * .5), (float)(H * .5), focal,
                                                                                      xs, ys, n_vec, n, o, d);
    return check_launch("nerf_raygen");
}

extern "C" int nerf_batch_rays(const float* poses_dev, const int64_t* img_idx_dev, const uint8_t* images_u8, int n_img, int H, int W,
                               float focal, const int64_t* xs, const int64_t* ys, int64_t n, float* o, float* d, float* rgb, void* stream) {
    NERF_REQUIRE(poses_dev && img_idx_dev && images_u8 && xs && ys && o && d && rgb, "nerf_batch_rays: null pointer");
    NERF_REQUIRE(n_img > 0 && H > 0 && W > 0 && n >= 0, "nerf_batch_rays: bad size n_img=%d H=%d W=%d n=%lld", n_img, H, W, (long long)n);
    if (n == 0) return 0;
    batch_rays_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(poses_dev, img_idx_dev, images_u8, n_img, H, W, (float)(W * .5),
                                                                          (float)(H * .5), focal, xs, ys, n, o, d, rgb);
    return check_launch("nerf_batch_rays");
}

extern "C" int nerf_coarse_sample(const float* o, const float* d, const float* u, const float* t_base, float step,
                                  int64_t N, int C, float* samples, float* ts, void* stream) {
    NERF_REQUIRE(N >= 0 && C > 0, "nerf_coarse_sample: bad size N=%lld C=%d", (long long)N, C);
    if (N == 0) return 0;
    NERF_REQUIRE(o && d && u && t_base && ts, "nerf_coarse_sample: null pointer");
    coarse_sample_kernel<<<grid_for(N * C, 256), 256, 0, (cudaStream_t)stream>>>(o, d, u, t_base, step, N * C, C, samples, ts);
    return check_launch("nerf_coarse_sample");
}

extern "C" int nerf_deltas(const float* ts, int64_t N, int S, float* deltas, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_deltas: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(ts && deltas, "nerf_deltas: null pointer");
    deltas_kernel<<<grid_for(N * S, 256), 256, 0, (cudaStream_t)stream>>>(ts, N * S, S, deltas);
    return check_launch("nerf_deltas");
}

extern "C" int nerf_weights(const float* sigma, const float* deltas, int64_t N, int S, float* weights, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_weights: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(sigma && deltas && weights, "nerf_weights: null pointer");
    composite_kernel<0><<<grid_for(N, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(
        sigma, nullptr, deltas, N, S, nullptr, weights, nullptr, nullptr, nullptr, nullptr);
    return check_launch("nerf_weights");
}

extern "C" int nerf_ray_color(const float* weights, const float* rgb, int64_t N, int S, float* ray_rgb, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_ray_color: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(weights && rgb && ray_rgb, "nerf_ray_color: null pointer");
    ray_color_kernel<<<grid_for(N, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(weights, rgb, N, S, ray_rgb);
    return check_launch("nerf_ray_color");
}

extern "C" int nerf_composite(const float* sigma, const float* rgb, const float* ts, int64_t N, int S,
                              float* deltas, float* weights, float* ray_rgb, float* depth, float* acc,
                              float* stats2, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0, "nerf_composite: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(sigma && rgb && ts, "nerf_composite: null pointer");
    // one thread per ray needs tens of thousands of rays to fill the GPU (a ray is a serial chain of S steps); a 4096-ray
    // render / training chunk is served better by one warp per ray
    if (N >= 32768)
        composite_rays_kernel<<<grid_for(N, kCompRays), kCompRays, 0, (cudaStream_t)stream>>>(
            sigma, rgb, ts, N, S, deltas, weights, ray_rgb, depth, acc, stats2);
    else
        composite_kernel<1><<<grid_for(N, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(
            sigma, rgb, ts, N, S, deltas, weights, ray_rgb, depth, acc, stats2);
    return check_launch("nerf_composite");
}

extern "C" int nerf_fine_sample(const float* o, const float* d, const float* w, const float* ts, const float* eps,
                                const float* u, const float* q_base, int64_t N, int C, int F, float near_, float far_,
                                float* fine_samples, float* fine_ts, int64_t* idx, void* stream) {
    NERF_REQUIRE(N >= 0 && C > 0 && F > 0 && C <= 2048, "nerf_fine_sample: bad size N=%lld C=%d F=%d", (long long)N, C, F);
    if (N == 0) return 0;
    NERF_REQUIRE(o && d && w && ts && eps && u && q_base && fine_ts, "nerf_fine_sample: null pointer");
    const size_t smem = (size_t)kWarpsPerBlock * (2 * C + 2) * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(fine_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fine_sample_kernel<<<grid_for(N, kWarpsPerBlock), kThreads, smem, (cudaStream_t)stream>>>(
        o, d, w, ts, eps, u, q_base, N, C, F, near_, far_, fine_samples, fine_ts, idx);
    return check_launch("nerf_fine_sample");
}

extern "C" int nerf_merge_sort(const float* o, const float* d, const float* ts_a, int A, const float* ts_b, int B,
                               int64_t N, float* ts_sorted, float* samples_sorted, void* stream) {
    NERF_REQUIRE(N >= 0 && A >= 0 && B >= 0 && A + B > 0 && A + B <= 1024, "nerf_merge_sort: bad size A=%d B=%d", A, B);
    if (N == 0) return 0;
    NERF_REQUIRE(ts_a && ts_b && ts_sorted, "nerf_merge_sort: null pointer");
    NERF_REQUIRE(!samples_sorted || (o && d), "nerf_merge_sort: samples_sorted needs o and d");
    if (A + B <= 256) {
        merge_sort256_kernel<<<grid_for(N, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(o, d, ts_a, A, ts_b, B, N, ts_sorted,
                                                                                            samples_sorted);
        return check_launch("nerf_merge_sort");
    }
    const size_t smem = (size_t)kWarpsPerBlock * 2 * (A + B) * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merge_sort_kernel<<<grid_for(N, kWarpsPerBlock), kThreads, smem, (cudaStream_t)stream>>>(
        o, d, ts_a, A, ts_b, B, N, ts_sorted, samples_sorted);
    return check_launch("nerf_merge_sort");
}

extern "C" int nerf_fine_sample_merge(const float* w, const float* ts, const float* eps, const float* u, const float* q_base,
                                      int64_t N, int C, int F, float near_, float far_, float* ts_sorted, void* stream) {
    NERF_REQUIRE(N >= 0 && C > 0 && F > 0 && C + F <= 256, "nerf_fine_sample_merge: bad size N=%lld C=%d F=%d (C + F <= 256)", (long long)N, C, F);
    if (N == 0) return 0;
    NERF_REQUIRE(w && ts && eps && u && q_base && ts_sorted, "nerf_fine_sample_merge: null pointer");
    const size_t smem = (size_t)kWarpsPerBlock * fine_sampler_floats(C) * sizeof(float);
    fine_sample_merge_kernel<<<grid_for(N, kWarpsPerBlock), kThreads, smem, (cudaStream_t)stream>>>(
        w, ts, eps, u, q_base, N, C, F, near_, far_, ts_sorted);
    return check_launch("nerf_fine_sample_merge");
}

extern "C" int nerf_composite_backward(const float* sigma, const float* rgb, const float* ts, const float* g_ray, int64_t N, int S,
                                       float* dsigma_pre, float* drgb_pre, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0 && S <= 1024, "nerf_composite_backward: bad size (S <= 1024)");
    if (N == 0) return 0;
    NERF_REQUIRE(sigma && rgb && ts && g_ray && dsigma_pre && drgb_pre, "nerf_composite_backward: null pointer");
    composite_backward_kernel<<<grid_for(N, kWarpsPerBlock), kThreads, 0, (cudaStream_t)stream>>>(sigma, rgb, ts, g_ray, N, S,
                                                                                               dsigma_pre, drgb_pre);
    return check_launch("nerf_composite_backward");
}
