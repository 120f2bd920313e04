"""Stand-in for imageio==2.4.1 (imread via PIL) used only by tests/golden/make_golden.py."""
import numpy as np
from PIL import Image


def imread(path, pilmode=None):
    im = Image.open(path)
    if pilmode:
        im = im.convert(pilmode)
    return np.asarray(im)


def mimwrite(path, frames, **kw):
    ims = [Image.fromarray(f) for f in frames]
    ims[0].save(path, save_all=True, append_images=ims[1:], loop=0, duration=100)
