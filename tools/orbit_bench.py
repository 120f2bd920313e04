"""BASELINE.json configs[4]: the full 40-frame 360-degree orbit at 800x800 (64 + 128 samples per ray, ray chunk 4096) through
the public API - pose_spherical -> get_rays -> view_reconstruction, the loop of generate_360_view_synthesis - on 1..8 GPUs
(torchrun: every frame's rays are split into contiguous slabs, one uint8 all-gather per frame).  Prints one JSON line.

    python tools/orbit_bench.py [--poses 40] [--hw 800]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/orbit_bench.py
"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=40)
    ap.add_argument("--hw", type=int, default=800)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import dataloader
    import nerf_helpers
    import nerf_model
    import synthetic
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(5, "dense"))
    net = net.to(dev)
    H = W = args.hw
    focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
    angles = np.linspace(-180, 180, args.poses + 1)[:-1]

    def frame(angle):
        o, d = dataloader.get_rays(H, W, focal, nerf_helpers.pose_spherical(angle, -30, 4.0), device=dev)
        return nerf_helpers.view_reconstruction(net, o, d, N=4096)           # uint8 [H,W,3] on the host, on every rank

    frame(angles[0])                                                          # warm-up (weight packing, allocator, NCCL)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    views = [frame(a) for a in angles]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    if rank == 0:
        t1 = time.perf_counter()
        with tempfile.TemporaryDirectory() as tmp:
            dataloader.write_gif(Path(tmp, "orbit-360.gif"), views)
            gif_bytes = Path(tmp, "orbit-360.gif").stat().st_size
        gif_s = time.perf_counter() - t1
        rays = args.poses * H * W
        print(json.dumps({"workload": f"{args.poses}-frame 360 orbit at {H}x{W}, 64 + 128 samples/ray, chunk 4096, through view_reconstruction "
                                      "(host uint8 frames on every rank)", "n_gpus": world, "seconds": dt, "frames_per_s": args.poses / dt,
                          "rays_per_s": rays / dt, "ms_per_frame": dt / args.poses * 1e3, "gif_write_s": gif_s, "gif_bytes": gif_bytes,
                          "frame_mean": float(np.mean(views[0])), "timing": "wall clock around the loop, max over ranks, synchronised on both sides"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
