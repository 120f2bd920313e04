"""Training entry point - the B200 counterpart of the reference's `train_nerf.py` (flags and sub-commands of
train_nerf.py:62-96, `train_full_nerf` of train_nerf.py:20-34).

    python train_nerf.py -n NAME --gpu -s STEPS -rd ROOT -r 4096 full -b ./data/nerf_synthetic/lego/ -cr 0
    torchrun --nproc-per-node 8 train_nerf.py -n NAME --gpu ... full -b ...     (data-parallel over ray batches)

Only the `full` sub-command (coarse + fine NeRF, the path BASELINE.json names) is built; `single` and `simple` are the
reference's debugging toys: they parse, and exit with a message.

What differs from upstream under the same command line: the Lightning `Trainer` / `WandbLogger` pair is `trainer.Trainer` +
`trainer.JsonLogger` (no network, PL-format checkpoints kept), the learning-rate monitor's job is done by the per-step `lr`
entry of the metrics file, and under torchrun every rank joins one NCCL group (one all-reduce of the flat gradients per step).
"""
import argparse
import os

import torch
import torch.distributed as dist

import dataloader
import nerf_model
from trainer import JsonLogger, Trainer

# (flags, keyword arguments) of the options shared by every sub-command, then of `full`; defaults as upstream except the
# author's home directory in --root_dir
_COMMON = (
    (('-n', '--name'), dict(type=str, required=True, help='experiment name (directory under ROOT/NeRF/)')),
    (('-s', '--steps'), dict(type=int, default=100000, help='optimiser steps to run')),
    (('--gpu',), dict(action='store_true', help='kept for compatibility: this path always runs on the GPU')),
    (('-p', '--position_encoding'), dict(type=int, default=10, help='frequencies of the position encoding')),
    (('-d', '--direction_encoding'), dict(type=int, default=4, help='frequencies of the direction encoding')),
    (('-rd', '--root_dir'), dict(type=str, default='./experiments/', help='where checkpoints and metrics go')),
    (('-r', '--rays'), dict(type=int, default=4096, help='rays per training batch')),
    (('-l', '--ckpt'), dict(type=str, default=None, help='resume from this .ckpt file')),
)
_FULL = (
    (('-b', '--base_dir'), dict(type=str, default='./data/nerf_synthetic/lego/', help='Blender-synthetic scene directory')),
    (('-c', '--coarse'), dict(type=int, default=64, help='stratified samples per ray')),
    (('-f', '--fine'), dict(type=int, default=128, help='importance samples per ray')),
    (('-nr', '--near'), dict(type=float, default=2.0, help='near plane')),
    (('-fr', '--far'), dict(type=float, default=6.0, help='far plane')),
    (('-cr', '--cropping_epochs'), dict(type=int, default=10, help='epochs that sample only the image centre')),
)


def _join_process_group():
    """torchrun sets RANK / LOCAL_RANK: bind the rank to its GPU and join the NCCL group once."""
    if 'RANK' in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')


def train_full_nerf(root_dir, base_dir, logger_name, steps, pos_enc, direc_enc, use_gpu,
                    num_rays, coarse_samples, fine_samples, near, far, cropping_epochs, ckpt, args):
    """Coarse + fine NeRF on a Blender-synthetic scene; returns the trainer (global_step, last_checkpoint, ...)."""
    _join_process_group()
    rank0 = not dist.is_initialized() or dist.get_rank() == 0
    metrics = JsonLogger(name=logger_name, project='NeRF', save_dir=root_dir, enabled=rank0)   # one writer per run
    metrics.log_hyperparams(args)
    network = nerf_model.NeRFNetwork(position_dim=pos_enc, direction_dim=direc_enc, coarse_samples=coarse_samples,
                                     fine_samples=fine_samples, near=near, far=far)
    scene = dataloader.SyntheticDataModule(base_dir, num_rays, cropping_epochs, num_workers=2)
    schedule = dict(max_steps=steps, check_val_every_n_epoch=10, reload_dataloaders_every_n_epochs=cropping_epochs)
    run = Trainer(gpus=int(use_gpu), default_root_dir=root_dir, resume_from_checkpoint=ckpt, logger=metrics,
                  track_grad_norm=2, **schedule)
    run.fit(network, datamodule=scene)
    return run


def build_parser():
    parser = argparse.ArgumentParser(description='Train a NeRF model')
    kinds = parser.add_subparsers(dest='type', help='which model to train')
    for flags, kw in _COMMON:
        parser.add_argument(*flags, **kw)
    sub = {name: kinds.add_parser(name) for name in ('simple', 'full', 'single')}
    for flags, kw in _FULL:
        sub['full'].add_argument(*flags, **kw)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.type != 'full':
        raise SystemExit(f"sub-command {args.type!r}: only 'full' (the coarse + fine NeRF hot path) is built on the B200 path")
    return train_full_nerf(args.root_dir, args.base_dir, args.name, args.steps, args.position_encoding, args.direction_encoding,
                           args.gpu, args.rays, args.coarse, args.fine, args.near, args.far, args.cropping_epochs, args.ckpt, args)


if __name__ == '__main__':
    main()
