// mlp_fp32.cu - positional encoding and the exact-fp32 CUDA-core form of NeRFModel.forward.
//
// This is the tight-parity / any-shape path (arbitrary position_dim, direction_dim, N, S): it evaluates
// nerf_model.py:362-389 with fp32 FMAs and accurate sincosf, and is what the tensor-core kernel
// (mlp_tc3.cu) is cross-checked against on the device.  It is not the throughput path.
#include "common.cuh"

namespace nerf {

constexpr float kPiF = 3.14159274101257324f;   // fl32(pi): what `x / math.pi` and `2**i * torch.pi * x` see on CPU

// Layout per frequency i (nerf_model.py:29-31): [cos(a_0..a_{c-1}), sin(a_0..a_{c-1})], a = fl32(2^i pi) * x.
__device__ __forceinline__ void encode_point(const float* x, int c, int L, float* out) {
    for (int i = 0; i < L; ++i) {
        const float f = ldexpf(kPiF, i);
        for (int k = 0; k < c; ++k) {
            float s, co;
            sincosf(__fmul_rn(f, x[k]), &s, &co);
            out[i * 2 * c + k] = co;
            out[i * 2 * c + c + k] = s;
        }
    }
}

__global__ void __launch_bounds__(256)
positional_encoding_kernel(const float* __restrict__ x, int64_t total, int c, int L, float* __restrict__ out) {
    // one thread per (point, frequency, channel)
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(e % c);
        const int i = (int)((e / c) % L);
        const int64_t n = e / ((int64_t)c * L);
        float s, co;
        sincosf(__fmul_rn(ldexpf(kPiF, i), x[n * c + k]), &s, &co);
        float* row = out + n * (2 * L * c) + i * 2 * c;
        row[k] = co;
        row[c + k] = s;
    }
}

constexpr int kTile = 32;       // samples per block
constexpr int kHidden = 256;

struct NetParams { const float* p[20]; };

// out[s][n] = act(b[n] + sum_k in0[s][k] W[n][k] + sum_k in1[s][k] W[n][K0+k]); thread n owns output n.
// act: 0 none, 1 relu, 2 sigmoid.
__device__ __forceinline__ void dense(const float* in0, int ld0, int K0, const float* in1, int ld1, int K1,
                                      const float* __restrict__ W, const float* __restrict__ b, int Nout,
                                      float* out, int ldo, int act, int rows) {
    const int n = threadIdx.x;
    if (n < Nout) {
        float acc[kTile];
        const float bias = b[n];
#pragma unroll
        for (int s = 0; s < kTile; ++s) acc[s] = bias;
        const float* wr = W + (size_t)n * (K0 + K1);
        for (int k = 0; k < K0; ++k) {
            const float w = __ldg(wr + k);
#pragma unroll
            for (int s = 0; s < kTile; ++s) acc[s] = fmaf(in0[s * ld0 + k], w, acc[s]);
        }
        for (int k = 0; k < K1; ++k) {
            const float w = __ldg(wr + K0 + k);
#pragma unroll
            for (int s = 0; s < kTile; ++s) acc[s] = fmaf(in1[s * ld1 + k], w, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < kTile; ++s) {
            float v = acc[s];
            if (act == 1) v = fmaxf(v, 0.f);
            else if (act == 2) v = 1.0f / (1.0f + expf(-v));
            if (s < rows) out[s * ldo + n] = v;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kHidden)
mlp_fp32_kernel(NetParams P, int Lp, int Ld, const float* __restrict__ samples, const float* __restrict__ direc,
                int64_t total, int S, float* __restrict__ sigma, float* __restrict__ rgb) {
    extern __shared__ float smem[];
    const int pe = 6 * Lp, de = 6 * Ld;
    float* pex = smem;                       // [kTile][pe]
    float* ped = pex + kTile * pe;           // [kTile][de]
    float* h0 = ped + kTile * de;            // [kTile][256]
    float* h1 = h0 + kTile * kHidden;        // [kTile][256]
    const int64_t base = (int64_t)blockIdx.x * kTile;
    const int rows = (int)min((int64_t)kTile, total - base);

    if (threadIdx.x < kTile) {
        const int s = threadIdx.x;
        float x[3] = {0.f, 0.f, 0.f}, u[3] = {0.f, 0.f, 0.f};
        if (s < rows) {
            const int64_t e = base + s, n = e / S;
            const float dx = direc[n * 3], dy = direc[n * 3 + 1], dz = direc[n * 3 + 2];
            const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);                   // nerf_model.py:373
            u[0] = __fdiv_rn(dx, nrm); u[1] = __fdiv_rn(dy, nrm); u[2] = __fdiv_rn(dz, nrm);
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = __fdiv_rn(samples[e * 3 + k], kPiF);  // nerf_model.py:377
        }
        encode_point(x, 3, Lp, pex + s * pe);
        encode_point(u, 3, Ld, ped + s * de);
    }
    __syncthreads();

    // nerf_model.py:331-340
    dense(pex, pe, pe, nullptr, 0, 0, P.p[0], P.p[1], kHidden, h0, kHidden, 1, kTile);
    dense(h0, kHidden, kHidden, nullptr, 0, 0, P.p[2], P.p[3], kHidden, h1, kHidden, 1, kTile);
    dense(h1, kHidden, kHidden, nullptr, 0, 0, P.p[4], P.p[5], kHidden, h0, kHidden, 1, kTile);
    dense(h0, kHidden, kHidden, nullptr, 0, 0, P.p[6], P.p[7], kHidden, h1, kHidden, 1, kTile);
    // nerf_model.py:342-348, 383-384: cat(h, PE(x)) -> 256 -> 256 -> 256 (last one linear)
    dense(h1, kHidden, kHidden, pex, pe, pe, P.p[8], P.p[9], kHidden, h0, kHidden, 1, kTile);
    dense(h0, kHidden, kHidden, nullptr, 0, 0, P.p[10], P.p[11], kHidden, h1, kHidden, 1, kTile);
    dense(h1, kHidden, kHidden, nullptr, 0, 0, P.p[12], P.p[13], kHidden, h0, kHidden, 0, kTile);
    // nerf_model.py:350-353, 385: density = relu(256 -> 1)
    dense(h0, kHidden, kHidden, nullptr, 0, 0, P.p[14], P.p[15], 1, sigma + base, 1, 1, rows);
    // nerf_model.py:355-360, 387-388: rgb = sigmoid(128 -> 3 (relu(280 -> 128 (cat(feat, PE(dir))))))
    dense(h0, kHidden, kHidden, ped, de, de, P.p[16], P.p[17], 128, h1, kHidden, 1, kTile);
    dense(h1, kHidden, 128, nullptr, 0, 0, P.p[18], P.p[19], 3, rgb + base * 3, 3, 2, rows);
}

}  // namespace nerf

using namespace nerf;

extern "C" int nerf_positional_encoding(const float* x, int64_t n, int c, int L, float* out, void* stream) {
    NERF_REQUIRE(x && out, "nerf_positional_encoding: null pointer");
    NERF_REQUIRE(n >= 0 && c > 0 && L > 0 && L <= 64, "nerf_positional_encoding: bad size");
    if (n == 0) return 0;
    const int64_t total = n * c * L;
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
    positional_encoding_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, total, c, L, out);
    return check_launch("nerf_positional_encoding");
}

extern "C" int nerf_mlp_forward_fp32(const float* const* params20_host, int position_dim, int direction_dim,
                                     const float* samples, const float* direc, int64_t N, int S,
                                     float* sigma, float* rgb, void* stream) {
    NERF_REQUIRE(N >= 0 && S > 0 && position_dim > 0 && direction_dim > 0 && position_dim <= 32 && direction_dim <= 32,
                 "nerf_mlp_forward_fp32: bad size");
    if (N == 0) return 0;
    NERF_REQUIRE(params20_host && samples && direc && sigma && rgb, "nerf_mlp_forward_fp32: null pointer");
    NetParams P;
    for (int i = 0; i < 20; ++i) {
        NERF_REQUIRE(params20_host[i], "nerf_mlp_forward_fp32: params20_host[%d] is NULL", i);
        P.p[i] = params20_host[i];
    }
    const int64_t total = N * S;
    const size_t smem = (size_t)kTile * (6 * position_dim + 6 * direction_dim + 2 * kHidden) * sizeof(float);
    cudaFuncSetAttribute(mlp_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t blocks = (total + kTile - 1) / kTile;
    NERF_REQUIRE(blocks < (1ll << 31), "nerf_mlp_forward_fp32: too many samples");
    mlp_fp32_kernel<<<(int)blocks, kHidden, smem, (cudaStream_t)stream>>>(P, position_dim, direction_dim, samples, direc,
                                                                         total, S, sigma, rgb);
    return check_launch("nerf_mlp_forward_fp32");
}
