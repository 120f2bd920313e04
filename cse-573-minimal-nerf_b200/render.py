"""Orbit renderer entry point - the B200 counterpart of the reference's `render.py` (render.py:14-28).

    python render.py -c CKPT_PATH -r 4096 -p 40 -s SAVE_DIR
    torchrun --nproc-per-node 8 render.py -c CKPT_PATH ...      # every frame's rays sharded over the GPUs

Same command line, same `render(ckpt, save_dir, rays, num_poses)` call, same output file (`<save_dir>/epoch=<E>-360.gif`,
the epoch tag cut out of the checkpoint's file name as upstream does); the frames come from the hand-written CUDA path
(`nerf_helpers.generate_360_view_synthesis` -> `view_reconstruction` -> `NeRFNetwork.forward`).
"""
import argparse
import os
import re
from pathlib import Path

import torch

import nerf_helpers
import nerf_model

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')

# flag, destination, type, default, help: the reference's four options
_OPTIONS = (
    ('-c', '--ckpt', str, None, 'ckpt path for model'),
    ('-r', '--rays', int, 4096, 'number of rays per batch'),
    ('-p', '--num_poses', int, 40, 'number of images in gif.'),
    ('-s', '--save_dir', Path, './recons/', 'where to save the resulting gif'),
)


def epoch_tag(ckpt_name):
    """'…epoch=1089-step=108999.ckpt' -> 'epoch=1089' (everything from 'epoch=' up to the next '-', render.py:15-16)."""
    m = re.search(r'epoch=[^-]*', str(ckpt_name))
    if m is None:
        raise ValueError(f"checkpoint name {ckpt_name!r} carries no 'epoch=' tag")
    return m.group(0)


def render(ckpt, save_dir, rays, num_poses):
    """Loads the PL-format checkpoint and writes the `num_poses`-frame orbit GIF; returns the uint8 frames."""
    if device.type != 'cuda':
        raise RuntimeError("render: a CUDA device is required (this path has no CPU implementation)")
    _maybe_join_process_group()
    network = nerf_model.NeRFNetwork.load_from_checkpoint(str(ckpt)).to(device)
    return nerf_helpers.generate_360_view_synthesis(network, Path(save_dir), epoch_tag(ckpt), N=rays, num_poses=num_poses)


def _maybe_join_process_group():
    """Under torchrun (WORLD_SIZE > 1) bind this rank to its GPU and join the NCCL group: `view_reconstruction` then renders one
    ray slab per rank and all-gathers the uint8 slabs."""
    global device
    import torch.distributed as dist
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        device = torch.device('cuda', local)
        dist.init_process_group('nccl', device_id=device)


def build_parser():
    parser = argparse.ArgumentParser(description='Render a 360 view from a NeRF Model')
    for short, long_, kind, default, text in _OPTIONS:
        parser.add_argument(short, long_, type=kind, default=default, required=default is None, help=text)
    return parser


def main(argv=None):
    opts = build_parser().parse_args(argv)
    opts.save_dir.mkdir(parents=True, exist_ok=True)
    render(opts.ckpt, opts.save_dir, opts.rays, opts.num_poses)


if __name__ == '__main__':
    main()
