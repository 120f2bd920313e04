"""On-device cross-checks of the hand-written backward kernels: the same dgrad / wgrad chain as library (cuBLAS) GEMMs and torch
elementwise ops.  TEST INFRASTRUCTURE ONLY - the product path (training.py) never calls these."""
import math

import torch

import training
from training import ACT, BF, DZ, F32, composite_backward, padded_rows, untile
import _native as nat


def _pe(x, L):
    import nerf_model
    return nerf_model.positional_encoding(x, L)


def mlp_backward_library_wgrad(model, o, d, ts, sigma, rgb, acts, g_ray):
    """Same gradients with the hand-written dgrad kernel but the weight gradients as bf16 cuBLAS GEMMs on untiled copies
    (kept as an on-device cross-check of wgrad_tc.cu)."""
    acts, masks = acts
    N, S = ts.shape[0], ts.shape[1]
    M = N * S
    dsig, drgb = composite_backward(sigma, rgb, ts, g_ray)
    dz_t = torch.empty((padded_rows(M) * DZ,), device=ts.device, dtype=BF)
    with nat.timed_kernel("mlp_tc_bwd_kernel", M):
        nat.check(nat.lib().nerf_mlp_backward_tc(nat.ptr(model.packed_weights_t()), nat.ptr(masks), nat.ptr(dsig), nat.ptr(drgb),
                                                 N, S, nat.ptr(dz_t), nat.stream()), "nerf_mlp_backward_tc")
    acts, dz = untile(acts, M, ACT), untile(dz_t, M, DZ)        # interim: the library wgrad GEMMs want row-major operands
    feat, r, dr = acts[:, 1536:1792], acts[:, 1792:1920], dz[:, 1792:1920]
    pts = (d[:, None, :] * ts + o[:, None, :]).reshape(M, 3)
    pe, de = 6 * model.position_dim, 6 * model.direction_dim
    # PE operands padded to 64 / 32 columns so the library picks aligned tensor-core kernels
    pe_x = torch.zeros((M, 64), device=ts.device, dtype=BF)
    pe_x[:, :pe] = _pe(pts / math.pi, model.position_dim)
    unit = d / torch.linalg.norm(d, dim=1, keepdim=True)
    pe_d = torch.zeros((N, 32), device=ts.device, dtype=BF)
    pe_d[:, :de] = _pe(unit, model.direction_dim)
    pe_d = pe_d.repeat_interleave(S, dim=0)                                                 # [M,32]

    def wgrad(g, a):
        return torch.mm(g.t(), a, out_dtype=F32)
    # bias gradients of the 8 hidden layers + density row in ONE GEMM: [ones; dsig]^T-style 16-row left operand
    left = torch.zeros((16, M), device=ts.device, dtype=BF)
    left[0] = 1.0
    left[1] = dsig
    red = torch.mm(left, dz, out_dtype=F32)                                                 # row 0: column sums of dz
    bsum = red[0]
    grads = [None] * 20
    grads[0], grads[1] = wgrad(dz[:, 0:256], pe_x)[:, :pe], bsum[0:256]                     # mlp.0
    for li in (1, 2, 3, 5, 6):                                                              # mlp.2/4/6, feature_fn.2/4
        grads[2 * li], grads[2 * li + 1] = wgrad(dz[:, 256 * li:256 * li + 256], acts[:, 256 * (li - 1):256 * li]), \
            bsum[256 * li:256 * li + 256]
    dz4 = dz[:, 1024:1280]                                                                  # feature_fn.0: input [h3, PE(x)]
    grads[8], grads[9] = torch.cat([wgrad(dz4, acts[:, 768:1024]), wgrad(dz4, pe_x)[:, :pe]], dim=1), bsum[1024:1280]
    grads[14], grads[15] = torch.mm(left, feat, out_dtype=F32)[1:2], dsig.sum().reshape(1)  # density_fn.0
    grads[16], grads[17] = torch.cat([wgrad(dr, feat), wgrad(dr, pe_d)[:, :de]], dim=1), bsum[1792:1920]   # rgb_fn.0
    g9 = torch.zeros((M, 16), device=ts.device, dtype=BF)
    g9[:, :3] = drgb
    grads[18], grads[19] = wgrad(g9, r)[:3], drgb.sum(0)                                    # rgb_fn.2
    return grads


def mlp_backward_reference(model, o, d, ts, sigma, rgb, acts, g_ray):
    """The same gradients with the whole chain as library GEMMs + elementwise torch ops: on-device reference for the
    hand-written dgrad kernel (tests/test_gpu_training.py)."""
    N, S = ts.shape[0], ts.shape[1]
    M = N * S
    acts = untile(acts[0], M, ACT)
    dsig, drgb = composite_backward(sigma, rgb, ts, g_ray)
    P = [p.detach() for p in model.ordered_params()]
    W = [P[2 * i] for i in range(10)]
    Wb = [w.to(BF) for w in W]
    h = [acts[:, 256 * k:256 * (k + 1)] for k in range(7)]        # outputs of mlp.0,2,4,6, feature_fn.0,2,4
    feat, r = h[6], acts[:, 1792:1920]
    pts = (d[:, None, :] * ts + o[:, None, :]).reshape(M, 3)
    pe_x = _pe(pts / math.pi, model.position_dim).to(BF)                                    # [M,60]
    unit = d / torch.linalg.norm(d, dim=1, keepdim=True)
    pe_d = _pe(unit, model.direction_dim).to(BF).repeat_interleave(S, dim=0)               # [M,24]
    ones = torch.ones((1, M), device=ts.device, dtype=BF)

    def wgrad(dz, a):                      # dW[out,in] = dz^T a, fp32 accumulate/output
        return torch.mm(dz.t(), a, out_dtype=F32)

    def bgrad(dz):
        return torch.mm(ones, dz, out_dtype=F32)[0]

    grads = [None] * 20
    # rgb_fn.2 (nerf_model.py:358): rgb_pre = r W9^T + b9
    g9 = drgb.to(BF)
    grads[18], grads[19] = wgrad(g9, r), drgb.sum(0)
    dr = torch.mm(g9, Wb[9]) * (r > 0)
    # rgb_fn.0 (nerf_model.py:356, 387): r_pre = [feat, PE(dir)] W8^T + b8
    grads[16], grads[17] = torch.cat([wgrad(dr, feat), wgrad(dr, pe_d)], dim=1), bgrad(dr)
    dfeat = torch.mm(dr, Wb[8][:, :256])
    # density_fn.0 (nerf_model.py:351, 385): sigma_pre = feat W7^T + b7
    gs = dsig.to(BF)[:, None]
    grads[14], grads[15] = wgrad(gs, feat), dsig.sum().reshape(1)
    dz = dfeat + gs * Wb[7]
    # feature_fn.4 (linear), feature_fn.2, feature_fn.0 (input = [h3, PE(x)]), mlp.6, mlp.4, mlp.2, mlp.0
    for li in (6, 5, 4, 3, 2, 1):
        a = h[li - 1]
        if li == 4:
            grads[8], grads[9] = torch.cat([wgrad(dz, a), wgrad(dz, pe_x)], dim=1), bgrad(dz)
            dz = torch.mm(dz, Wb[4][:, :256]) * (a > 0)
        else:
            grads[2 * li], grads[2 * li + 1] = wgrad(dz, a), bgrad(dz)
            dz = torch.mm(dz, Wb[li]) * (a > 0)
    grads[0], grads[1] = wgrad(dz, pe_x), bgrad(dz)
    return grads
