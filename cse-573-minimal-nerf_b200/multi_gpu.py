"""Multi-GPU plumbing.  The path shards by rays with no exchange inside it:
  render: each rank renders a contiguous slab of the frame's rays; one all-gather of the uint8 slab per frame;
  train:  data parallel over ray batches; one all-reduce of the flat gradient buffer per step (trainer.FlatGradients).
One process per GPU (torchrun), torch.distributed over NCCL on GPUs (gloo in the CPU tests of the host logic)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def ray_slab(n_rays, rank, world_size):
    """[lo, hi) of the contiguous slab of rays rank `rank` renders; slabs differ by at most one ray."""
    base, rem = divmod(n_rays, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_slabs(local, n_rays, world_size=None):
    """All-gather variable-length slabs [n_local, C] (uint8 or float) into [n_rays, C] on every rank."""
    rank, ws = world()
    ws = world_size or ws
    if ws == 1:
        return local
    longest = ray_slab(n_rays, 0, ws)[1]
    padded = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(out, padded)
    return torch.cat([out[r][:ray_slab(n_rays, r, ws)[1] - ray_slab(n_rays, r, ws)[0]] for r in range(ws)], dim=0)


def sharded_render(render_rays, all_o, all_d):
    """render_rays(o [n,3], d [n,3]) -> [n,3] float in [0,1].  Returns the full uint8 image [H,W,3] on every rank."""
    H, W, C = all_o.shape
    rank, ws = world()
    lo, hi = ray_slab(H * W, rank, ws)
    o, d = all_o.reshape(H * W, C)[lo:hi], all_d.reshape(H * W, C)[lo:hi]
    rgb = render_rays(o, d)
    slab = (rgb * 255).clamp(0, 255).to(torch.uint8)
    return gather_slabs(slab, H * W, ws).reshape(H, W, C)
