// render.cu - NeRFNetwork.forward (nerf_model.py:89-132) as ONE call of the C ABI (nerf_render_forward): the three
// launches of the render path queued back to back on the caller's stream, intermediates in a caller-provided workspace.
//   coarse network  nerf_mlp_composite_tc_strata   stratified depths + PE + MLP + heads + compositing   (nerf_model.py:103-111)
//   sampler         nerf_fine_sample_merge         inverse-CDF fine depths merged with the coarse ones  (nerf_model.py:114-120)
//   fine network    nerf_mlp_composite_tc          PE + MLP + heads + compositing                        (nerf_model.py:123-130)
// The Python host (training.forward_pass) issues the same three calls itself when it also needs the intermediates (training,
// keep_samples); a C / C++ host that only renders needs nothing else.
#include "common.cuh"

using namespace nerf;

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" size_t nerf_render_workspace_bytes(int64_t N, int C, int F) {
    if (N <= 0 || C <= 0 || F <= 0) return 0;
    // coarse depths [N,C] | coarse weights [N,C] | merged depths [N,C+F] | coarse depth / opacity [N] each
    return align256((size_t)N * C * 4) * 2 + align256((size_t)N * (C + F) * 4) + align256((size_t)N * 4) * 2;
}

extern "C" int nerf_render_forward(const void* packed_coarse, const void* packed_fine, const float* o, const float* d,
                                   const float* u_c, const float* t_base, float step, const float* eps, const float* u_f,
                                   const float* q_base, int64_t N, int C, int F, float near_f, float far_f, float* coarse_rgb,
                                   float* fine_rgb, float* depth, float* acc, float* stats8, void* workspace, void* stream) {
    NERF_REQUIRE(N >= 0 && C > 0 && F > 0, "nerf_render_forward: bad size N=%lld C=%d F=%d", (long long)N, C, F);
    if (N == 0) return 0;
    NERF_REQUIRE(packed_coarse && packed_fine && o && d && u_c && t_base && eps && u_f && q_base && coarse_rgb && fine_rgb && workspace,
                 "nerf_render_forward: null pointer");
    NERF_REQUIRE(((uintptr_t)workspace & 255) == 0, "nerf_render_forward: the workspace must be 256-byte aligned");
    NERF_REQUIRE(C + F <= 256 && nerf_mlp_composite_tc_supported(C) && nerf_mlp_composite_tc_supported(C + F),
                 "nerf_render_forward: C = %d / C + F = %d is not a shape the fused kernels support (multiples of 32, C + F <= 256)", C, C + F);
    uint8_t* ws = (uint8_t*)workspace;
    float* ts_c = (float*)ws;                ws += align256((size_t)N * C * 4);
    float* w_c = (float*)ws;                 ws += align256((size_t)N * C * 4);
    float* ts = (float*)ws;                  ws += align256((size_t)N * (C + F) * 4);
    float* depth_c = (float*)ws;             ws += align256((size_t)N * 4);
    float* acc_c = (float*)ws;
    int rc = nerf_mlp_composite_tc_strata(packed_coarse, o, d, u_c, t_base, step, N, C, ts_c, nullptr, nullptr, nullptr, nullptr, w_c,
                                          coarse_rgb, depth_c, acc_c, stats8, stream);
    if (rc) return rc;
    rc = nerf_fine_sample_merge(w_c, ts_c, eps, u_f, q_base, N, C, F, near_f, far_f, ts, stream);
    if (rc) return rc;
    return nerf_mlp_composite_tc(packed_fine, o, d, ts, N, C + F, nullptr, nullptr, nullptr, nullptr, nullptr, fine_rgb, depth, acc,
                                 stats8 ? stats8 + 4 : nullptr, stream);
}
