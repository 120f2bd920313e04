/* nerf_b200.h - C ABI of libnerf_b200.so: the sm_100a NeRF render-and-train hot path.
 *
 * The reference (NakuraMino/CSE-573-Minimal-NeRF) is pure Python/PyTorch and has no FFI layer; its
 * boundary is the Python call surface of nerf_helpers.py / nerf_model.py / dataloader.py.  Each entry
 * point below replaces the *inside* of one of those functions (file:line cited per function) and is
 * what the Python host modules in cse-573-minimal-nerf_b200/ bind through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (or the stated integer type) owned by the
 *     caller, except where the name ends in `_host`;
 *   - nothing is allocated or freed here except lazily created per-device function attributes;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy stream);
 *   - return value 0 = launched, negative = rejected (NERF_E_*); text via nerf_last_error() (thread-local);
 *   - N = rays, C = coarse samples per ray, F = fine samples per ray, S = samples per ray of one network.
 *   - there is no CPU implementation behind any of these: without a CUDA device they fail.
 */
#ifndef NERF_B200_H_
#define NERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERF_ABI_VERSION 3

/* libnerf_b200.so is built with -fvisibility=hidden: the functions declared here are its whole dynamic symbol table
 * (tests/test_abi.py compares `nm -D` with this header). */
#if defined(__GNUC__)
#define NERF_API __attribute__((visibility("default")))
#else
#define NERF_API
#endif

#define NERF_E_ARG    (-1)   /* null pointer / bad size / unsupported shape */
#define NERF_E_CUDA   (-2)   /* CUDA runtime error at launch */
#define NERF_E_DEVICE (-3)   /* device is not sm_100 */

NERF_API int nerf_abi_version(void);
NERF_API const char* nerf_last_error(void);

/* ---- H0: rays.  dataloader.py:36-43 (get_rays), :150-152 (pixel gather).
 * c2w_host: 12 or 16 floats on the HOST, row-major rows of the 4x4 camera-to-world matrix.
 * xs, ys: [n] int64 device arrays of pixel column / row, or both NULL for the full H*W grid in
 * row-major order (then n must be H*W).  o, d: [n,3].  d is NOT normalised (as upstream). */
NERF_API int nerf_raygen(const float* c2w_host, int H, int W, float focal, const int64_t* xs, const int64_t* ys,
                int64_t n, float* o, float* d, void* stream);

/* ---- N1: one training batch (dataloader.py:143-152: rays of a pixel list + their colours) with the image chosen ON THE DEVICE, so
 * that a CUDA graph of the training step can be replayed with a different image per step.  poses_dev [n_img,16]: row-major 4x4
 * camera-to-world matrices; img_idx_dev: one int64 (clamped to [0, n_img)); images_u8 [n_img,H,W,3].  xs, ys [n] int64 as
 * nerf_raygen.  o, d, rgb: [n,3]; rgb = fl32(u8 / 255), the division in double precision as numpy does upstream. */
NERF_API int nerf_batch_rays(const float* poses_dev, const int64_t* img_idx_dev, const uint8_t* images_u8, int n_img, int H, int W,
                             float focal, const int64_t* xs, const int64_t* ys, int64_t n, float* o, float* d, float* rgb, void* stream);

/* ---- H1: stratified depths and points.  nerf_helpers.py:28-56.
 * u: [N,C] uniforms in [0,1).  t_base: [C] strata origins (the reference's torch.arange(near, far, step),
 * evaluated by the host with torch so that its rounding is inherited).  ts[n,i] = t_base[i] + u*step,
 * samples = d*t + o (separate multiply and add).  samples may be NULL. */
NERF_API int nerf_coarse_sample(const float* o, const float* d, const float* u, const float* t_base, float step,
                       int64_t N, int C, float* samples, float* ts, void* stream);

/* ---- H2: deltas.  nerf_helpers.py:58-73.  ts, deltas: [N,S]; deltas[:, S-1] = 1e10. */
NERF_API int nerf_deltas(const float* ts, int64_t N, int S, float* deltas, void* stream);

/* ---- H3: unnormalised weights.  nerf_helpers.py:75-91.  sigma, deltas, weights: [N,S].
 * The exclusive running sum of -sigma*delta is taken sequentially in fp32 (CPU torch order). */
NERF_API int nerf_weights(const float* sigma, const float* deltas, int64_t N, int S, float* weights, void* stream);

/* ---- H4: ray colour.  nerf_helpers.py:93-104.  weights [N,S], rgb [N,S,3] -> ray_rgb [N,3]. */
NERF_API int nerf_ray_color(const float* weights, const float* rgb, int64_t N, int S, float* ray_rgb, void* stream);

/* ---- H2+H3+H4 in one pass (what NeRFNetwork.forward does at nerf_model.py:109-111 and :128-130).
 * sigma [N,S], rgb [N,S,3], ts [N,S].  Every output may be NULL: deltas [N,S], weights [N,S],
 * ray_rgb [N,3], depth [N] (sum w*t), acc [N] (sum w).  stats2 (nullable, [4], zero-initialised by the caller, ACCUMULATED with atomics):
 * [0] sum sigma^2, [1] count(sigma != 0) - the two density statistics logged at nerf_model.py:105-106 -, [2] sqrt([0]) written
 * by the last block of the launch, [3] internal block counter. */
NERF_API int nerf_composite(const float* sigma, const float* rgb, const float* ts, int64_t N, int S,
                   float* deltas, float* weights, float* ray_rgb, float* depth, float* acc,
                   float* stats2, void* stream);

/* ---- backward of H2+H3+H4 (training; autograd of nerf_model.py:109-111 / :128-130).  g_ray [N,3] = dL/d ray_rgb.
 * Outputs are the gradients w.r.t. the head PRE-activations (ReLU / sigmoid derivatives of nerf_model.py:352,359
 * folded in): dsigma_pre [N,S], drgb_pre [N,S,3]. */
NERF_API int nerf_composite_backward(const float* sigma, const float* rgb, const float* ts, const float* g_ray, int64_t N, int S,
                            float* dsigma_pre, float* drgb_pre, void* stream);

/* ---- H5: inverse-CDF fine sampling.  nerf_helpers.py:106-156.
 * w, ts: [N,C] coarse weights and depths.  eps: [N] and u: [N,F] raw uniforms.  q_base: [F] query grid
 * (the reference's torch.arange(0, 1, 1/F), evaluated by the host).  cdf = sequential fp32 cumsum / last;
 * an all-zero ray has a NaN cdf and every query falls in the last bin, as upstream.
 * fine_samples [N,F,3] (nullable), fine_ts [N,F], idx [N,F] int64 lower-bound indices (nullable). */
NERF_API int nerf_fine_sample(const float* o, const float* d, const float* w, const float* ts, const float* eps,
                     const float* u, const float* q_base, int64_t N, int C, int F, float near_, float far_,
                     float* fine_samples, float* fine_ts, int64_t* idx, void* stream);

/* ---- H6: merge + sort.  nerf_model.py:116-120.  ts_a [N,A] (fine) and ts_b [N,B] (coarse) are
 * concatenated, sorted ascending per ray; samples_sorted [N,A+B,3] = o + t*d (nullable). A+B <= 1024. */
NERF_API int nerf_merge_sort(const float* o, const float* d, const float* ts_a, int A, const float* ts_b, int B,
                    int64_t N, float* ts_sorted, float* samples_sorted, void* stream);
/* K3 + K4 in one launch: the sorted depths NeRFNetwork.forward feeds the fine network (nerf_model.py:114-120) straight from
 * the coarse weights / depths: inverse-CDF fine depths (as nerf_fine_sample), concatenated fine-first with the coarse ones
 * and sorted (as nerf_merge_sort), bit-identical to those two calls.  C + F <= 256.  ts_sorted [N, C+F]. */
NERF_API int nerf_fine_sample_merge(const float* w, const float* ts, const float* eps, const float* u, const float* q_base,
                           int64_t N, int C, int F, float near_, float far_, float* ts_sorted, void* stream);

/* ---- H7: positional encoding.  nerf_model.py:19-33.  x [n,c] -> out [n, 2*L*c];
 * per frequency i: cos(2^i pi x) for the c channels, then sin(2^i pi x). */
NERF_API int nerf_positional_encoding(const float* x, int64_t n, int c, int L, float* out, void* stream);

/* ---- H8: one NeRFModel forward.  nerf_model.py:362-389.
 * params20_host: HOST array of 20 device pointers in state_dict order
 *   (mlp.0.weight, mlp.0.bias, mlp.2.*, mlp.4.*, mlp.6.*, feature_fn.0.*, feature_fn.2.*, feature_fn.4.*,
 *    density_fn.0.*, rgb_fn.0.*, rgb_fn.2.*), fp32, nn.Linear layout [out,in].
 * samples [N,S,3], direc [N,3] -> sigma [N,S], rgb [N,S,3].
 * This is the exact-fp32 CUDA-core form (any position_dim/direction_dim, any N,S); the tensor-core form is
 * nerf_mlp_forward_tc below. */
NERF_API int nerf_mlp_forward_fp32(const float* const* params20_host, int position_dim, int direction_dim,
                          const float* samples, const float* direc, int64_t N, int S,
                          float* sigma, float* rgb, void* stream);

/* ---- K5: pack one network's weights for the tcgen05 kernels (bf16, K-major, 128B-swizzled UMMA tiles,
 * biases fp32).  position_dim must be 10 and direction_dim 4 (the only shape the tensor-core kernel is
 * specialised for).  `packed` holds nerf_packed_bytes() bytes. */
NERF_API size_t nerf_packed_bytes(void);
NERF_API int nerf_pack_weights(const float* const* params20_host, void* packed, void* stream);

/* ---- K8: NeRFModel forward on the 5th-gen tensor cores (bf16 operands, fp32 accumulation in TMEM).
 * Samples are given as rays + depths: sample (n,s) sits at o[n] + ts[n,s]*d[n].  sigma [N,S], rgb [N,S,3]. */
NERF_API int nerf_mlp_forward_tc(const void* packed, const float* o, const float* d, const float* ts,
                        int64_t N, int S, float* sigma, float* rgb, void* stream);
/* Training form: also stores the bf16 activations each layer consumed (outputs of mlp.0, mlp.2, mlp.4, mlp.6,
 * feature_fn.0, feature_fn.2, feature_fn.4 at feature 256*k; rgb_fn.0 at 1792).  act_out: ceil(N*S/128)*128 rows x 1920
 * features, TILED CHUNK-MAJOR: element (row, f) at ((row/128 * 240 + f/8) * 128 + row%128) * 8 + f%8. */
/* mask_out: ceil(N*S/128)*128 * 30 * 8 bytes of ReLU sign bits (one bit per sample and hidden feature), all the dgrad kernel
 * needs from the activations.  The encoding is private to the forward / dgrad kernel pair: one 32-bit word per (sample, 32-feature
 * group) at word ((row/128)*60 + f/32)*128 + row%128, bit 15-j / 31-j = [pre-activation of feature 32*(f/32) + 2j / 2j+1 is negative].
 * Treat it as opaque. */
NERF_API int nerf_mlp_forward_tc_train(const void* packed, const float* o, const float* d, const float* ts,
                              int64_t N, int S, float* sigma, float* rgb, void* act_out, void* mask_out, void* stream);
/* ---- backward of H8 (dgrad chain) on the tensor cores.  packed_t = nerf_pack_weights_t image (W^T stages, bf16).
 * masks: the forward's ReLU sign words (see above); dsigma_pre [N*S], drgb_pre [N*S,3] from
 * nerf_composite_backward.  dz_out: ceil(N*S/128)*128 rows x 1936 features bf16, tiled chunk-major with 242 chunks per
 * tile: gradient w.r.t. every layer's pre-activation at the same feature offsets as acts (mlp.0 .. feature_fn.4 at 256*k,
 * rgb_fn.0 at 1792) + a heads block [dsigma_pre, drgb_pre x3, 0..] at 1920; weight gradients are dz^T . (layer input). */
NERF_API size_t nerf_packed_t_bytes(void);
NERF_API int nerf_pack_weights_t(const float* const* params20_host, void* packed_t, void* stream);
NERF_API int nerf_mlp_backward_tc(const void* packed_t, const void* masks, const float* dsigma_pre, const float* drgb_pre,
                         int64_t N, int S, void* dz_out, void* stream);
/* The same dgrad chain with the compositing backward (nerf_composite_backward) INSIDE the kernel - what the training step launches:
 * the producer warps turn g_ray [N,3] = dL/d(ray colour) and the forward's saved sigma [N,S], rgb [N,S,3], ts [N,S] into the head
 * gradients of the 256 samples of each tile pair in shared memory; dz_out is bit-identical to the two-call sequence. S <= 1024. */
NERF_API int nerf_mlp_backward_tc_fused(const void* packed_t, const void* masks, const float* sigma, const float* rgb, const float* ts,
                                        const float* g_ray, int64_t N, int S, void* dz_out, void* stream);
/* ---- weight / bias gradients of one network on the tensor cores: dW_l += dz_l^T . (input of layer l), db_l += sum dz_l.
 * acts, dz: the tiled chunk-major training tensors written by nerf_mlp_forward_tc_train / nerf_mlp_backward_tc;
 * o, d, ts as in the forward (PE(x) / PE(dir) operands are recomputed).  grads20_host: HOST array of 20 device pointers
 * (state_dict order, fp32, nn.Linear layout), ACCUMULATED into with atomics - zero them first. */
NERF_API int nerf_wgrad_tc(const void* acts, const void* dz, const float* o, const float* d, const float* ts, int64_t N, int S,
                           float* const* grads20_host, int deterministic, void* stream);
/* deterministic = 0: 148 CTAs split every job's tiles (split-K) and add their partial sums with fp32 atomics - the order of those
 * adds, hence the last bits of the gradients, varies from run to run (relative spread ~1e-7).  deterministic != 0: one CTA per job
 * walks all tiles in order and every gradient element receives exactly one add - bit-reproducible, ~16x slower; for parity debugging. */
/* nerf_mlp_forward_tc with explicit sample points [N,S,3] (the NeRFModel.forward(samples, direc) call surface). */
NERF_API int nerf_mlp_forward_tc_points(const void* packed, const float* samples, const float* d,
                               int64_t N, int S, float* sigma, float* rgb, void* stream);

/* ---- K8 + K2 fused: NeRFModel forward (nerf_model.py:362-389) with deltas, weights, ray colour, depth and opacity
 * (nerf_helpers.py:58-104, generate_deltas / calculate_unnormalized_weights / estimate_ray_color) computed INSIDE the
 * tensor-core kernel: every CTA owns whole rays, the per-sample (sigma, rgb) go through a shared-memory ring and are
 * composited by warps of the same CTA while the next tiles are in the tensor pipe.  Replaces the pair
 * nerf_mlp_forward_tc[_train] + nerf_composite; same arithmetic as those two calls (bit-identical outputs).
 * Needs S % 32 == 0 and a ray group (smallest run of whole rays that is a whole number of 128-sample tiles) of at most 6
 * tiles: nerf_mlp_composite_tc_supported(S) != 0, true for 64 / 128 / 192 / 256.
 * sigma [N,S] / rgb [N,S,3]: nullable together (render does not need them); act_out / mask_out: non-NULL = training form
 * (as nerf_mlp_forward_tc_train; needs sigma / rgb too).  weights [N,S], depth [N], acc [N]: nullable.  ray_rgb [N,3].
 * stats4: nullable; as nerf_composite (zero it first). */
NERF_API int nerf_mlp_composite_tc_supported(int S);
NERF_API int nerf_mlp_composite_tc(const void* packed, const float* o, const float* d, const float* ts, int64_t N, int S,
                          float* sigma, float* rgb, void* act_out, void* mask_out,
                          float* weights, float* ray_rgb, float* depth, float* acc, float* stats4, void* stream);

/* The coarse network's form of nerf_mlp_composite_tc: K1 (generate_coarse_samples, nerf_helpers.py:28-56) runs inside the kernel
 * too.  u [N,S] uniforms, t_base [S] = the reference's torch.arange(near, far, step) on the device, ts_out [N,S] receives the depths
 * t = t_base[i] + u * step (bit-identical to nerf_coarse_sample); everything else as nerf_mlp_composite_tc. */
NERF_API int nerf_mlp_composite_tc_strata(const void* packed, const float* o, const float* d, const float* u, const float* t_base, float step,
                                 int64_t N, int S, float* ts_out, float* sigma, float* rgb, void* act_out, void* mask_out,
                                 float* weights, float* ray_rgb, float* depth, float* acc, float* stats4, void* stream);

/* ---- H9 in one call: NeRFNetwork.forward (nerf_model.py:89-132) for rendering - coarse kernel, fine sampler + merge, fine kernel
 * queued back to back on `stream`.  u_c [N,C], eps [N], u_f [N,F]: the three uniform draws of the reference (nerf_helpers.py:52,139,154);
 * t_base [C] / step: the strata (as nerf_coarse_sample), q_base [F] the query grid (as nerf_fine_sample); near_f / far_f: the fine
 * sampler's bounds (upstream never forwards them: 2.0 / 6.0).  coarse_rgb, fine_rgb [N,3]; depth, acc [N] (fine network, nullable);
 * stats8 (nullable, zeroed by the caller): the density statistics of the coarse [0..4) and fine [4..8) network as nerf_composite's
 * stats.  workspace: nerf_render_workspace_bytes(N, C, F) bytes, 256-byte aligned, contents undefined afterwards.  C and C + F must
 * be shapes nerf_mlp_composite_tc supports, C + F <= 256.  Same results as the three calls it makes. */
NERF_API size_t nerf_render_workspace_bytes(int64_t N, int C, int F);
NERF_API int nerf_render_forward(const void* packed_coarse, const void* packed_fine, const float* o, const float* d, const float* u_c,
                                 const float* t_base, float step, const float* eps, const float* u_f, const float* q_base, int64_t N,
                                 int C, int F, float near_f, float far_f, float* coarse_rgb, float* fine_rgb, float* depth, float* acc,
                                 float* stats8, void* workspace, void* stream);

/* Both weight images (nerf_pack_weights + nerf_pack_weights_t) of both networks in ONE launch - what a training step needs
 * after the optimiser has changed the parameters.  params40_host: the 40 tensors of NeRFNetwork's state_dict order (coarse
 * network's 20, then the fine network's 20). */
NERF_API int nerf_pack_weights_all(const float* const* params40_host, void* packed0, void* packed_t0, void* packed1, void* packed_t1,
                          void* stream);

/* ---- optimiser step.  nerf_model.py:134-143 (torch.optim.Adam, lr 5e-4, betas (0.9, 0.999), eps 1e-8, no weight decay)
 * over flat fp32 buffers of n elements (all parameters of both networks): params updated in place, exp_avg / exp_avg_sq
 * are the Adam moments, step >= 1 is the 1-based step count used for the bias corrections.  Same arithmetic and order as
 * torch's single-tensor Adam; buffers 16-byte aligned.  grad_scale multiplies every gradient as it is read: 1.0f, or
 * 1 / world_size when `grads` holds the all-reduced SUM of the data-parallel ranks (the mean is never materialised). */
NERF_API int nerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                            float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream);

/* The same update with every step-dependent scalar in DEVICE memory, for a training step captured in a CUDA graph:
 * state8_dev[0..4] = lr, beta1, beta2, eps, grad_scale (the host rewrites them when they change, e.g. the per-epoch LR decay),
 * state8_dev[5..6] are scratch (step size and bias correction of the current step, evaluated on the device in double precision),
 * *step_dev = steps taken so far, incremented by the call.  n must be a multiple of 4 (the flat buffers are padded). */
NERF_API int nerf_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* state8_dev,
                                int64_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H_ */
