"""B200 drop-in for the reference's `train_nerf.py` (same flags and sub-commands, train_nerf.py:62-96).

    python train_nerf.py -n NAME --gpu -s STEPS -rd ROOT -r 4096 full -b ./data/nerf_synthetic/lego/ -cr 0
    torchrun --nproc-per-node 8 train_nerf.py -n NAME --gpu ... full -b ...     (data-parallel over ray batches)

Only the `full` sub-command (coarse + fine NeRF, the path BASELINE.json names) is implemented; `single` and `simple`
are the reference's debugging toys and exit with a message.
"""
import argparse
import os

import torch
import torch.distributed as dist

import dataloader
import nerf_model
from trainer import JsonLogger, Trainer


def train_full_nerf(root_dir, base_dir, logger_name, steps, pos_enc, direc_enc, use_gpu,
                    num_rays, coarse_samples, fine_samples, near, far, cropping_epochs, ckpt, args):
    """Train the full NeRF model (coarse + fine network)."""
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    logger = JsonLogger(name=logger_name, project="NeRF", save_dir=root_dir)
    logger.log_hyperparams(args)
    trainer = Trainer(gpus=int(use_gpu), default_root_dir=root_dir, max_steps=steps,
                      resume_from_checkpoint=ckpt, logger=logger,
                      check_val_every_n_epoch=10, track_grad_norm=2,
                      reload_dataloaders_every_n_epochs=cropping_epochs)
    data_module = dataloader.SyntheticDataModule(base_dir, num_rays, cropping_epochs, num_workers=2)
    model = nerf_model.NeRFNetwork(position_dim=pos_enc, direction_dim=direc_enc,
                                   coarse_samples=coarse_samples, fine_samples=fine_samples,
                                   near=near, far=far)
    trainer.fit(model, datamodule=data_module)
    return trainer


def build_parser():
    parser = argparse.ArgumentParser(description='Train a NeRF model')
    subparsers = parser.add_subparsers(dest='type', help='Training different NeRF Versions')
    parser.add_argument('-n', '--name', type=str, help='name of the model experiment', required=True)
    parser.add_argument('-s', '--steps', type=int, default=100000, help='max number of steps')
    parser.add_argument('--gpu', action='store_true')
    parser.add_argument('-p', '--position_encoding', type=int, default=10, help='position encoding length')
    parser.add_argument('-d', '--direction_encoding', type=int, default=4, help='direction encoding length')
    parser.add_argument('-rd', '--root_dir', type=str, default="./experiments/", help='directory to save models')
    parser.add_argument('-r', '--rays', type=int, default=4096, help='number of rays per batch')
    parser.add_argument('-l', '--ckpt', type=str, default=None, help='load/resume from checkpoint (path to a .ckpt)')
    subparsers.add_parser("simple")
    full_parser = subparsers.add_parser("full")
    subparsers.add_parser("single")
    full_parser.add_argument('-b', '--base_dir', type=str, default='./data/nerf_synthetic/lego/', help='directory for dataset')
    full_parser.add_argument('-c', '--coarse', type=int, default=64, help='number of coarse samples')
    full_parser.add_argument('-f', '--fine', type=int, default=128, help='number of fine samples')
    full_parser.add_argument('-nr', '--near', type=float, default=2.0, help='near bound for dataset')
    full_parser.add_argument('-fr', '--far', type=float, default=6.0, help='far bound of dataset')
    full_parser.add_argument('-cr', '--cropping_epochs', type=int, default=10, help='num. epochs to crop image for ray sampling.')
    return parser


if __name__ == '__main__':
    args = build_parser().parse_args()
    if args.type == 'full':
        train_full_nerf(args.root_dir, args.base_dir, args.name, args.steps, args.position_encoding,
                        args.direction_encoding, args.gpu, args.rays, args.coarse, args.fine, args.near,
                        args.far, args.cropping_epochs, args.ckpt, args)
    else:
        raise SystemExit(f"sub-command {args.type!r}: only 'full' (the coarse + fine NeRF hot path) is built on the B200 path")
