"""B200 drop-in for the reference's `dataloader.py`: ray generation on the device + Blender-synthetic data.

Reference counterparts (file:line): sample_random_coordinates 13-34, get_rays 36-43,
SyntheticDataModule 78-103, SyntheticDataset 105-158, getSyntheticDataloader 160-162.

The reference builds all H*W rays of an image on the CPU for every 4096-ray batch and keeps 0.6 % of them
(dataloader.py:147-152).  Here the images are decoded once into a device-resident uint8 tensor and a batch
is: randint on the device -> ray kernel for just those pixels -> colour gather.  The per-item dict keeps the
reference's keys and shapes.
"""
import json
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

import _native as nat

_cuda = torch.device('cuda') if torch.cuda.is_available() else None


def sample_random_coordinates(N, height, width, cropping=False, device=None):
    """Two [N] int64 tensors (xs in [0,width), ys in [0,height)); centre half of the image when cropping."""
    if cropping:
        ew, eh = width // 4, height // 4
        xs = torch.randint(ew, width - ew, size=(N,), device=device)
        ys = torch.randint(eh, height - eh, size=(N,), device=device)
    else:
        xs = torch.randint(0, width, size=(N,), device=device)
        ys = torch.randint(0, height, size=(N,), device=device)
    return xs, ys


def _c2w_host(c2w):
    m = torch.as_tensor(c2w, dtype=torch.float32).detach().cpu().contiguous()
    if m.shape not in ((4, 4), (3, 4)):
        raise ValueError(f"c2w must be 4x4 or 3x4, got {tuple(m.shape)}")
    return m


def get_rays(H, W, focal, c2w, device=None):
    """Pinhole rays for the whole H x W grid: (rays_o [H,W,3], rays_d [H,W,3]) on the CUDA device."""
    dv = torch.device(device) if device is not None else _cuda
    if dv is None or dv.type != 'cuda':
        raise RuntimeError("get_rays: a CUDA device is required (this path has no CPU implementation)")
    m = _c2w_host(c2w)
    o = torch.empty((H, W, 3), device=dv, dtype=torch.float32)
    d = torch.empty((H, W, 3), device=dv, dtype=torch.float32)
    with torch.cuda.device(dv):
        nat.check(nat.lib().nerf_raygen(m.data_ptr(), H, W, float(np.float32(focal)), None, None, H * W,
                                        nat.ptr(o), nat.ptr(d), nat.stream()), "nerf_raygen")
    return o, d


def get_rays_at(H, W, focal, c2w, xs, ys):
    """Rays for a pixel list only: xs, ys [n] int64 CUDA tensors -> (o [n,3], d [n,3])."""
    xs, ys = nat.dev(xs, "xs", torch.int64), nat.dev(ys, "ys", torch.int64)
    m = _c2w_host(c2w)
    n = xs.shape[0]
    o = torch.empty((n, 3), device=xs.device, dtype=torch.float32)
    d = torch.empty((n, 3), device=xs.device, dtype=torch.float32)
    with torch.cuda.device(xs.device):
        nat.check(nat.lib().nerf_raygen(m.data_ptr(), H, W, float(np.float32(focal)), nat.ptr(xs), nat.ptr(ys), n,
                                        nat.ptr(o), nat.ptr(d), nat.stream()), "nerf_raygen")
    return o, d


def batch_rays(poses, img_idx, images, focal, xs, ys):
    """Rays AND target colours of a pixel list of image `img_idx` (a 0-d int64 DEVICE tensor): poses [n_img,4,4] fp32 and images
    [n_img,H,W,3] uint8 are device-resident tables (SyntheticDataset.stacked()).  Nothing about the choice of image passes through
    the host, so the call can sit inside a CUDA graph (training.GraphedTrainStep).  Returns (o [n,3], d [n,3], rgb [n,3])."""
    xs, ys = nat.dev(xs, "xs", torch.int64), nat.dev(ys, "ys", torch.int64)
    n_img, H, W, _ = images.shape
    n = xs.shape[0]
    o = torch.empty((n, 3), device=xs.device, dtype=torch.float32)
    d = torch.empty((n, 3), device=xs.device, dtype=torch.float32)
    rgb = torch.empty((n, 3), device=xs.device, dtype=torch.float32)
    with torch.cuda.device(xs.device):
        nat.check(nat.lib().nerf_batch_rays(nat.ptr(poses), nat.ptr(img_idx), nat.ptr(images), n_img, H, W, float(np.float32(focal)),
                                            nat.ptr(xs), nat.ptr(ys), n, nat.ptr(o), nat.ptr(d), nat.ptr(rgb), nat.stream()), "nerf_batch_rays")
    return o, d, rgb


def read_image(path, pilmode="RGB"):
    from PIL import Image
    return np.asarray(Image.open(path).convert(pilmode))


def write_gif(path, frames, duration_ms=100):
    from PIL import Image
    ims = [Image.fromarray(np.asarray(f)) for f in frames]
    ims[0].save(str(path), save_all=True, append_images=ims[1:], loop=0, duration=duration_ms)


class SyntheticDataset(Dataset):
    """Blender-synthetic split: one item = `num_rays` random rays of one 800x800 image."""

    def __init__(self, base_dir, tvt, num_rays, cropping=False, device=None):
        self.H = self.W = 800        # the reference hard-codes the synthetic image size (dataloader.py:126-127)
        self.tvt, self.cropping, self.num_rays, self.base_dir = tvt, cropping, num_rays, base_dir
        self.device = torch.device(device) if device is not None else _cuda
        with open(Path(base_dir, f'transforms_{tvt}.json')) as fh:
            self.data = json.load(fh)
        self.camera_angle = self.data['camera_angle_x']
        self.focal = 0.5 * self.W / np.tan(0.5 * self.camera_angle)
        self.frames = []
        for f in self.data['frames']:
            f = dict(f)
            f['file_path'] = Path(base_dir, f"{f['file_path']}.png")
            f.pop('rotation', None)
            self.frames.append(f)
        self._images = {}

    def __len__(self):
        return len(self.frames)

    def image_u8(self, idx):
        """[H,W,3] uint8 on the device (alpha dropped, as pilmode='RGB' does upstream); decoded once."""
        if idx not in self._images:
            self._images[idx] = torch.from_numpy(read_image(self.frames[idx]['file_path'], "RGB").copy()).to(self.device)
        return self._images[idx]

    def stacked(self):
        """(images [n,H,W,3] uint8, poses [n,4,4] fp32) of the whole split on the device: the tables `batch_rays` indexes."""
        if getattr(self, "_stacked", None) is None:
            images = torch.stack([self.image_u8(i) for i in range(len(self))]).contiguous()
            poses = torch.stack([torch.tensor(f['transform_matrix'], dtype=torch.float32) for f in self.frames]).to(self.device).contiguous()
            self._stacked = (images, poses)
        return self._stacked

    def __getitem__(self, idx):
        frame = self.frames[idx]
        c2w = torch.tensor(frame['transform_matrix'], dtype=torch.float32)
        img = self.image_u8(idx)
        xs, ys = sample_random_coordinates(self.num_rays, self.H, self.W, cropping=self.cropping, device=self.device)
        origin, direction = get_rays_at(self.H, self.W, self.focal, c2w, xs, ys)
        rgb = img[ys, xs, :].to(torch.float32) / 255.0
        item = {'origin': origin, 'direc': direction, 'rgb': rgb, 'xs': xs, 'ys': ys}
        if self.tvt != 'train':
            o_all, d_all = get_rays(self.H, self.W, self.focal, c2w, device=self.device)
            item.update({'all_origin': o_all, 'all_direc': d_all, 'image': img.to(torch.float32) / 255.0})
        return item


class _BatchOfOne:
    """Iterates a dataset like DataLoader(batch_size=1): every tensor gains a leading 1 (which
    nerf_helpers.fix_batchify removes again).  Data already lives on the device, so no worker processes."""

    def __init__(self, dataset, shuffle):
        self.dataset, self.shuffle = dataset, shuffle

    def __len__(self):
        return len(self.dataset)

    def __iter__(self):
        order = torch.randperm(len(self.dataset)).tolist() if self.shuffle else range(len(self.dataset))
        for i in order:
            yield {k: v.unsqueeze(0) for k, v in self.dataset[i].items()}


def getSyntheticDataloader(base_dir, tvt, num_rays, cropping=False, num_workers=8, shuffle=True):
    return _BatchOfOne(SyntheticDataset(base_dir, tvt, num_rays, cropping=cropping), shuffle)


class SyntheticDataModule:
    """Crops to the image centre for the first `cropping_epochs` epochs, then samples the whole image."""

    def __init__(self, base_dir, num_rays, cropping_epochs, num_workers=8):
        self.num_rays, self.base_dir, self.cropping_epochs, self.num_workers = num_rays, base_dir, cropping_epochs, num_workers
        self.crop_train_ds = SyntheticDataset(base_dir, 'train', num_rays, cropping=True)
        self.train_ds = SyntheticDataset(base_dir, 'train', num_rays, cropping=False)
        self.train_ds._images = self.crop_train_ds._images          # one decoded copy of the images
        self.val_ds = SyntheticDataset(base_dir, 'val', num_rays, cropping=False)
        self.trainer = None

    def train_dataloader(self):
        epoch = self.trainer.current_epoch if self.trainer is not None else self.cropping_epochs
        return _BatchOfOne(self.crop_train_ds if epoch < self.cropping_epochs else self.train_ds, shuffle=True)

    def val_dataloader(self):
        return _BatchOfOne(self.val_ds, shuffle=False)
