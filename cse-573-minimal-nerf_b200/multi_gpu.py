"""Multi-GPU plumbing.  The path shards by rays with no exchange inside it:
  render: each rank renders a contiguous slab of the frame's rays; one all-gather of the uint8 slab per frame;
  train:  data parallel over ray batches; one all-reduce of the flat gradient buffer per step (trainer.FlatGradients).
One process per GPU (torchrun), torch.distributed over NCCL on GPUs (gloo in the CPU tests of the host logic)."""
import torch
import torch.distributed as dist


import contextlib

_local_only = 0


@contextlib.contextmanager
def local_only():
    """Inside this context the render helpers behave as a single process (world() == (0, 1)): used where only ONE rank renders
    (validation frames during data-parallel training), so that no all-gather waits for ranks that never arrive."""
    global _local_only
    _local_only += 1
    try:
        yield
    finally:
        _local_only -= 1


def world():
    if not _local_only and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def ray_slab(n_rays, rank, world_size):
    """[lo, hi) of the contiguous slab of rays rank `rank` renders; slabs differ by at most one ray."""
    base, rem = divmod(n_rays, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_slabs(local, n_rays, world_size=None):
    """All-gather variable-length slabs [n_local, C] (uint8 or float) into [n_rays, C] on every rank: ONE collective into one
    [world, longest, C] tensor (slabs differ by at most one ray; shorter ones are padded)."""
    rank, ws = world()
    ws = world_size or ws
    if ws == 1:
        return local
    longest = ray_slab(n_rays, 0, ws)[1]
    if local.shape[0] == longest:
        padded = local.contiguous()
    else:
        padded = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    out = torch.empty((ws,) + tuple(padded.shape), dtype=local.dtype, device=local.device)
    if out.is_cuda:
        dist.all_gather_into_tensor(out, padded)
    else:                                        # gloo (the CPU tests of this host logic) has no flat all-gather
        dist.all_gather(list(out.unbind(0)), padded)
    if n_rays % ws == 0:
        return out.reshape((n_rays,) + tuple(local.shape[1:]))
    return torch.cat([out[r, :ray_slab(n_rays, r, ws)[1] - ray_slab(n_rays, r, ws)[0]] for r in range(ws)], dim=0)


def sharded_render(render_rays, all_o, all_d):
    """render_rays(o [n,3], d [n,3]) -> [n,3] float in [0,1].  Returns the full uint8 image [H,W,3] on every rank."""
    H, W, C = all_o.shape
    rank, ws = world()
    lo, hi = ray_slab(H * W, rank, ws)
    o, d = all_o.reshape(H * W, C)[lo:hi], all_d.reshape(H * W, C)[lo:hi]
    rgb = render_rays(o, d)
    slab = (rgb * 255).clamp(0, 255).to(torch.uint8)
    return gather_slabs(slab, H * W, ws).reshape(H, W, C)
