"""B200 drop-in for the reference's `render.py`: renders a 360 degree orbit GIF from a checkpoint.

    python render.py -c CKPT_PATH -r 4096 -p 40 -s SAVE_DIR
"""
import argparse
from pathlib import Path

import torch

import nerf_helpers
import nerf_model

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def render(ckpt, save_dir, rays, num_poses):
    ckpt = str(ckpt)
    epoch_idx = ckpt.find('epoch=')
    epoch = ckpt[epoch_idx: epoch_idx + ckpt[epoch_idx:].find('-')]
    model = nerf_model.NeRFNetwork.load_from_checkpoint(ckpt).to(device)
    return nerf_helpers.generate_360_view_synthesis(model, Path(save_dir), epoch, N=rays, num_poses=num_poses)


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Render a 360 view from a NeRF Model')
    parser.add_argument('-c', '--ckpt', type=str, required=True, help='ckpt path for model')
    parser.add_argument('-r', '--rays', type=int, default=4096, help='number of rays per batch')
    parser.add_argument('-p', '--num_poses', type=int, default=40, help='number of images in gif.')
    parser.add_argument('-s', '--save_dir', type=Path, default='./recons/', help='where to save the resulting gif')
    args = parser.parse_args()
    args.save_dir.mkdir(parents=True, exist_ok=True)
    render(args.ckpt, args.save_dir, args.rays, args.num_poses)
