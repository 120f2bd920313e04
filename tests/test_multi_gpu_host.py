"""Host-side logic of the N > 1 paths on CPU with the gloo backend, world_size 2: ray-slab partition + gather for
rendering, and the single flat-buffer gradient all-reduce of data-parallel training."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, ws, port, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        import multi_gpu
        from trainer import FlatGradients
        # --- render: slabs cover the frame exactly once, gather reproduces the single-process image
        H, W = 5, 7                                   # 35 rays: uneven split 18 + 17
        o = torch.arange(H * W * 3, dtype=torch.float32).reshape(H, W, 3) / (H * W * 3)
        d = o.flip(-1)

        def render_rays(oo, dd):
            return (oo * 0.5 + dd * 0.5).clamp(0, 1)
        lo, hi = multi_gpu.ray_slab(H * W, rank, ws)
        im = multi_gpu.sharded_render(render_rays, o, d)
        ref = (render_rays(o.reshape(-1, 3), d.reshape(-1, 3)) * 255).clamp(0, 255).to(torch.uint8).reshape(H, W, 3)
        ok_render = bool(torch.equal(im, ref)) and (hi - lo) in (17, 18)
        # --- train: every gradient is a view of one flat buffer; one all-reduce averages it
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        flat = FlatGradients(model.parameters())
        flat.zero()
        x = torch.full((5, 4), float(rank + 1))
        model(x).sum().backward()
        local = flat.flat.clone()
        flat.all_reduce_mean()
        gathered = [torch.empty_like(local) for _ in range(ws)]
        dist.all_gather(gathered, local)
        ok_train = bool(torch.allclose(flat.flat, sum(gathered) / ws)) and all(
            p.grad.data_ptr() >= flat.flat.data_ptr() for p in model.parameters())
        ok_views = bool(torch.equal(torch.cat([p.grad.flatten() for p in model.parameters()]), flat.flat))
        # --- the product's form (optim.FlatAdam): the optimiser owns the flat gradient buffer and applies 1 / world itself
        # (grad_scale), so the reduced buffer holds the SUM; a first slice may be reduced early (the coarse network's gradients)
        class FlatOpt:
            def __init__(self, n):
                self.flat_grads, self.grad_scale = torch.zeros(n), 1.0
        opt = FlatOpt(10)
        params = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(6))]
        folded = FlatGradients(params, opt)
        ok_fold = folded.flat is opt.flat_grads and opt.grad_scale == 1.0 / ws and folded.folded
        folded.zero()
        opt.flat_grads += torch.arange(10, dtype=torch.float32) * (rank + 1)
        mine = opt.flat_grads.clone()
        folded.reduce_async(4)                        # the first 4 values are final: their reduction starts now
        folded.all_reduce_mean()                      # the rest, wait for both, NO rescale
        total = sum((torch.arange(10, dtype=torch.float32) * (r + 1)) for r in range(ws))
        ok_fold = ok_fold and bool(torch.equal(opt.flat_grads, total)) and folded._pending == []
        ok_fold = ok_fold and bool(torch.allclose(folded.norm(), (total / ws).norm()))
        # local_only(): inside it the render helpers see a single process (rank-0-only validation frames)
        with multi_gpu.local_only():
            alone = multi_gpu.world() == (0, 1) and bool(torch.equal(multi_gpu.sharded_render(render_rays, o, d), ref))
        ok_fold = ok_fold and alone and multi_gpu.world() == (rank, ws)
        results[rank] = (ok_render, ok_train, ok_views and ok_fold)
    finally:
        dist.destroy_process_group()


def test_slab_partition():
    import multi_gpu
    for n, ws in ((640000, 8), (35, 2), (10, 3), (7, 8)):
        spans = [multi_gpu.ray_slab(n, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert multi_gpu.ray_slab(640000, 3, 8) == (240000, 320000)      # 100 image rows per GPU at 800x800


def test_world_size_2_gloo():
    ws = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(ws, port, results), nprocs=ws, join=True)
    assert len(results) == ws
    for r in range(ws):
        assert results[r] == (True, True, True), (r, results[r])
