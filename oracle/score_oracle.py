"""CPU oracle for the scoring stage (score.py:20-41 of the reference).  TEST INFRASTRUCTURE ONLY - nothing in the product
path imports this file.

The reference calls two functions of a third-party dependency that is not vendored in /root/reference and not installed
here: scikit-image, pinned `scikit_image==0.18.3` (requirements.txt:4) - `skimage.metrics.peak_signal_noise_ratio(gt, im)` and
`skimage.metrics.structural_similarity(gt, im, multichannel=True)` on uint8 [H,W,3] images (score.py:33-37).  Their published
algorithms are restated below with numpy + scipy.ndimage (which is what skimage itself calls):

  PSNR  = 10 log10(R^2 / mean((gt - im)^2)), float64, R = 255 for uint8 input.
  SSIM  (Wang et al. 2004, skimage defaults): per channel, float64, 7x7 uniform window (`scipy.ndimage.uniform_filter`),
          K1 = 0.01, K2 = 0.03, R = 255, SAMPLE covariance (x NP/(NP-1), NP = 49), S = ((2 ux uy + C1)(2 vxy + C2)) /
          ((ux^2 + uy^2 + C1)(vx + vy + C2)), mean of S over the image minus a (7-1)/2 = 3 pixel border; multichannel = mean of the
          per-channel values.

Parity status: UNPINNED against skimage itself (not installable offline); pinned only by the known answers in
tests/test_oracle_golden.py (identical images -> 1, constant offset -> closed form, PSNR of a unit error = 20 log10 255) and by two
implementations that share no code with this file (OpenCV's cv2.PSNR; a brute-force per-window evaluation of the SSIM definition).
"""
import numpy as np
from scipy.ndimage import uniform_filter


def peak_signal_noise_ratio(gt, im, data_range=255.0):
    gt, im = np.asarray(gt, dtype=np.float64), np.asarray(im, dtype=np.float64)
    mse = np.mean((gt - im) ** 2)
    return 10.0 * np.log10(data_range ** 2 / mse)


def _ssim_channel(x, y, win=7, data_range=255.0, k1=0.01, k2=0.03):
    x, y = x.astype(np.float64), y.astype(np.float64)
    npix = win ** x.ndim
    cov_norm = npix / (npix - 1.0)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return s[pad:-pad, pad:-pad].mean()


def structural_similarity(gt, im, multichannel=True):
    gt, im = np.asarray(gt), np.asarray(im)
    if not multichannel:
        return _ssim_channel(gt, im)
    return float(np.mean([_ssim_channel(gt[..., c], im[..., c]) for c in range(gt.shape[-1])]))
