"""B200 drop-in for the reference's `score.py` (score.py:20-52): test-set PSNR / SSIM of a checkpoint.

    python score.py -c CKPT_PATH -r 4096 -b BASE_DIR

The frames come from `nerf_helpers.view_reconstruction` (the hand-written render path); the two metrics - upstream calls
scikit-image 0.18.3, which is not installable here - are evaluated on the device in float64 with skimage's definitions and
defaults: PSNR = 10 log10(255^2 / MSE); SSIM = Wang et al. with a 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance,
3-pixel border excluded, mean over the three channels (`multichannel=True`).
"""
import argparse
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

import nerf_helpers
import nerf_model
from dataloader import SyntheticDataset

device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')


def _u8_on_device(im):
    t = torch.from_numpy(np.ascontiguousarray(im)) if isinstance(im, np.ndarray) else im
    if t.dtype != torch.uint8:
        raise TypeError(f"expected a uint8 image, got {t.dtype}")
    if not t.is_cuda:
        if device.type != 'cuda':
            raise RuntimeError("score: a CUDA device is required (this path has no CPU implementation)")
        t = t.to(device)
    return t.to(torch.float64)


def peak_signal_noise_ratio(gt_im, recon):
    """skimage.metrics.peak_signal_noise_ratio for uint8 images (data range 255)."""
    a, b = _u8_on_device(gt_im), _u8_on_device(recon)
    return float(10.0 * torch.log10(255.0 ** 2 / torch.mean((a - b) ** 2)))


def structural_similarity(gt_im, recon, multichannel=True):
    """skimage.metrics.structural_similarity with its defaults for uint8 [H,W,3] (or [H,W]) images."""
    a, b = _u8_on_device(gt_im), _u8_on_device(recon)
    if a.ndim == 2 or not multichannel:
        a, b = a.reshape(1, 1, *a.shape[-2:]), b.reshape(1, 1, *b.shape[-2:])
    else:
        a, b = a.permute(2, 0, 1).unsqueeze(1), b.permute(2, 0, 1).unsqueeze(1)            # [C,1,H,W]
    win, c1, c2 = 7, (0.01 * 255.0) ** 2, (0.03 * 255.0) ** 2
    cov_norm = win * win / (win * win - 1.0)

    def mean7(x):                   # 7x7 box means of the "valid" positions = the image minus skimage's 3-pixel border
        return F.avg_pool2d(x, win, stride=1)
    ux, uy = mean7(a), mean7(b)
    vx = cov_norm * (mean7(a * a) - ux * ux)
    vy = cov_norm * (mean7(b * b) - uy * uy)
    vxy = cov_norm * (mean7(a * b) - ux * uy)
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(s.mean(dim=(1, 2, 3)).mean())


def calculate_scores(ckpt, base_dir, rays):
    """Mean PSNR / SSIM of the checkpoint's renders over the scene's test split (score.py:20-41); prints upstream's three
    lines and returns (psnr, ssim)."""
    network = nerf_model.NeRFNetwork.load_from_checkpoint(str(ckpt)).to(device)
    views = SyntheticDataset(base_dir, 'test', rays, cropping=False)
    totals = {"psnr": 0.0, "ssim": 0.0}
    for index in range(len(views)):
        view = views[index]
        # ground truth as upstream forms it: float image * 255, clipped, truncated to uint8 (score.py:31)
        truth = (view['image'] * 255).clamp(0, 255).to(torch.uint8)
        frame = nerf_helpers.view_reconstruction(network, view['all_origin'], view['all_direc'], N=rays)
        totals["psnr"] += peak_signal_noise_ratio(truth, frame)
        totals["ssim"] += structural_similarity(truth, frame, multichannel=True)
    count = max(len(views), 1)
    psnr, ssim = totals["psnr"] / count, totals["ssim"] / count
    print("==============Calculate Scores==============")
    print(f"average psnr score: {psnr}")
    print(f"average ssim score: {ssim}")
    return psnr, ssim


def build_parser():
    parser = argparse.ArgumentParser(description='Calculate score metrics for NeRF Models.')
    for flags, kw in ((('-c', '--ckpt'), dict(type=str, required=True, help='checkpoint (.ckpt) to score')),
                      (('-r', '--rays'), dict(type=int, default=4096, help='rays per rendered chunk')),
                      (('-b', '--base_dir'), dict(type=Path, default='/content/CSEP573-NeRF/data/nerf_synthetic/lego/',
                                                  help='Blender-synthetic scene directory (its test split is scored)'))):
        parser.add_argument(*flags, **kw)
    return parser


def main(argv=None):
    opts = build_parser().parse_args(argv)
    return calculate_scores(opts.ckpt, opts.base_dir, opts.rays)


if __name__ == '__main__':
    main()
