// adam.cu - N3: the optimiser step of NeRFNetwork.configure_optimizers (nerf_model.py:134-143: torch.optim.Adam, lr 5e-4,
// betas (0.9, 0.999), eps 1e-8, no weight decay) as ONE elementwise kernel over the flat parameter / gradient / moment
// buffers of both networks (924 680 floats), instead of the library's multi-tensor kernels.  Same arithmetic, in the same
// order, as torch's single-tensor Adam:
//   m <- m + (g - m)(1 - beta1);  v <- v beta2 + (1 - beta2) g g;  p <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// with bc1 = 1 - beta1^t, bc2 = 1 - beta2^t evaluated by the host in double precision.
#include <math.h>
#include "common.cuh"

namespace nerf {

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 float one_minus_beta1, float beta2, float one_minus_beta2, float step_size, float inv_bc2_sqrt, float eps,
                 float grad_scale) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = ((float4*)p)[i], mm = ((float4*)m)[i], vv = ((float4*)v)[i];
        float4 gg = ((const float4*)g)[i];
        // data parallel: the buffer holds the SUM over ranks; grad_scale = 1 / world_size (exactly 1.0f otherwise: x * 1 == x)
        gg.x *= grad_scale; gg.y *= grad_scale; gg.z *= grad_scale; gg.w *= grad_scale;
        float* P = (float*)&pp; float* M = (float*)&mm; float* V = (float*)&vv; const float* G = (const float*)&gg;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            M[k] = fmaf(G[k] - M[k], one_minus_beta1, M[k]);                               // exp_avg.lerp_(grad, 1 - beta1)
            V[k] = fmaf(one_minus_beta2 * G[k], G[k], V[k] * beta2);                        // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = __fadd_rn(__fmul_rn(sqrtf(V[k]), inv_bc2_sqrt), eps);
            P[k] = fmaf(-step_size, __fdiv_rn(M[k], denom), P[k]);                         // addcdiv_(exp_avg, denom, -step_size)
        }
        ((float4*)p)[i] = pp; ((float4*)m)[i] = mm; ((float4*)v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * grad_scale;
        const float mi = fmaf(gi - m[i], one_minus_beta1, m[i]);
        const float vi = fmaf(one_minus_beta2 * gi, gi, v[i] * beta2);
        const float denom = __fadd_rn(__fmul_rn(sqrtf(vi), inv_bc2_sqrt), eps);
        m[i] = mi; v[i] = vi;
        p[i] = fmaf(-step_size, __fdiv_rn(mi, denom), p[i]);
    }
}

// Graph-safe form: nothing about the step comes from the host, so a CUDA graph can replay it.  state[0..4] = lr, beta1, beta2,
// eps, grad_scale (written by the host when they change), state[5..6] = step_size, 1 / sqrt(bias_correction2) (written here);
// *step is the number of steps taken so far and is incremented here.
__global__ void adam_prepare_kernel(float* __restrict__ state, long long* __restrict__ step) {
    const long long t = *step + 1;
    *step = t;
    const double bc1 = 1.0 - pow((double)state[1], (double)t), bc2 = 1.0 - pow((double)state[2], (double)t);
    state[5] = (float)((double)state[0] / bc1);
    state[6] = (float)(1.0 / sqrt(bc2));
}

__global__ void __launch_bounds__(256)
adam_step_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                     const float* __restrict__ state) {
    const float one_minus_beta1 = 1.0f - state[1], beta2 = state[2], one_minus_beta2 = 1.0f - state[2], eps = state[3], grad_scale = state[4],
                step_size = state[5], inv_bc2_sqrt = state[6];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {       // n is a multiple of 4 (flat buffers are padded)
        const float gi = g[i] * grad_scale;
        const float mi = fmaf(gi - m[i], one_minus_beta1, m[i]);
        const float vi = fmaf(one_minus_beta2 * gi, gi, v[i] * beta2);
        const float denom = __fadd_rn(__fmul_rn(sqrtf(vi), inv_bc2_sqrt), eps);
        m[i] = mi; v[i] = vi;
        p[i] = fmaf(-step_size, __fdiv_rn(mi, denom), p[i]);
    }
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* state8_dev,
                                  int64_t* step_dev, void* stream) {
    NERF_REQUIRE(n >= 0, "nerf_adam_step_dev: bad size");
    if (n == 0) return 0;
    NERF_REQUIRE(params && grads && exp_avg && exp_avg_sq && state8_dev && step_dev, "nerf_adam_step_dev: null pointer");
    adam_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state8_dev, (long long*)step_dev);
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    adam_step_dev_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, state8_dev);
    return check_launch("nerf_adam_step_dev");
}

extern "C" int nerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream) {
    NERF_REQUIRE(n >= 0 && step >= 1, "nerf_adam_step: bad size / step");
    if (n == 0) return 0;
    NERF_REQUIRE(params && grads && exp_avg && exp_avg_sq, "nerf_adam_step: null pointer");
    NERF_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
                 "nerf_adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    int64_t blocks = ((n >> 2) + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, 1.0f - beta1, beta2,
                                                                    1.0f - beta2, step_size, inv_bc2_sqrt, eps, grad_scale);
    return check_launch("nerf_adam_step");
}
