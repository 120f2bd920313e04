"""Deterministic synthetic inputs: uniforms, NeRF weights, checkpoints, Blender-shaped scenes.

The reference's shipped checkpoints (`models/model=lego-epoch=1089-step=108999.ckpt`, ...) and the
Blender dataset are not available offline, so every test / benchmark input is generated here from a
counter-based integer hash.  Nothing depends on a library RNG stream, so the same arrays come out in
the container that writes `tests/golden/*.npz` and on the GPU box that checks them.

Checkpoint layout follows what PyTorch-Lightning 1.5.10 writes for the reference's `NeRFNetwork`
(`/root/reference/nerf_model.py:56-87`, loaded by `/root/reference/render.py:17`): a `torch.save`d dict
whose `state_dict` holds the 40 tensors `{coarse,fine}_network.{mlp.{0,2,4,6},feature_fn.{0,2,4},
density_fn.0,rgb_fn.{0,2}}.{weight,bias}` and no `hyper_parameters`.
"""
import json
import math
from pathlib import Path

import numpy as np
import torch

NETS = ("coarse_network", "fine_network")


def layer_table(position_dim=10, direction_dim=4):
    """(state_dict sub-key, out_features, in_features) in state_dict order (nerf_model.py:331-360)."""
    pe, de = position_dim * 6, direction_dim * 6
    return [
        ("mlp.0", 256, pe), ("mlp.2", 256, 256), ("mlp.4", 256, 256), ("mlp.6", 256, 256),
        ("feature_fn.0", 256, 256 + pe), ("feature_fn.2", 256, 256), ("feature_fn.4", 256, 256),
        ("density_fn.0", 1, 256),
        ("rgb_fn.0", 128, 256 + de), ("rgb_fn.2", 3, 128),
    ]


def state_dict_keys(position_dim=10, direction_dim=4):
    keys = []
    for net in NETS:
        for name, _, _ in layer_table(position_dim, direction_dim):
            keys += [f"{net}.{name}.weight", f"{net}.{name}.bias"]
    return keys


def hash_u32(idx, seed):
    """lowbias32 integer hash of (idx, seed); idx is a uint64 numpy array."""
    m = np.uint64(0xFFFFFFFF)
    x = (idx.astype(np.uint64) + np.uint64((seed * 0x9E3779B9 + 0x7F4A7C15) & 0xFFFFFFFF)) & m
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def uniforms(seed, shape):
    """float32 uniforms in [0,1) on a 2^-24 grid (the grid `torch.rand` uses), as a numpy array."""
    n = int(np.prod(shape)) if len(shape) else 1
    bits = hash_u32(np.arange(n, dtype=np.uint64), seed)
    return ((bits >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(shape)


def normals(seed, shape):
    u1 = np.maximum(uniforms(seed * 2 + 1, shape).astype(np.float64), 2.0 ** -24)
    u2 = uniforms(seed * 2 + 2, shape).astype(np.float64)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)


def make_state_dict(seed=0, kind="init", position_dim=10, direction_dim=4):
    """Synthetic NeRFNetwork state_dict.

    kind="init":  nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases.
    kind="dense": the same, with the density head's bias raised by 0.5 and its weights doubled, so that
                  sigma sits well away from the ReLU knife-edge at the 1e10-wide last interval
                   and bf16-vs-fp32 differences stay smooth.
    """
    sd = {}
    k = 0
    for net in NETS:
        for name, fo, fi in layer_table(position_dim, direction_dim):
            bound = 1.0 / math.sqrt(fi)
            w = (uniforms(seed * 1000 + k, (fo, fi)) * 2.0 - 1.0) * bound
            b = (uniforms(seed * 1000 + k + 1, (fo,)) * 2.0 - 1.0) * bound
            k += 2
            if kind == "dense" and name == "density_fn.0":
                w = w * 2.0
                b = b + 0.5
            elif kind not in ("init", "dense"):
                raise ValueError(kind)
            sd[f"{net}.{name}.weight"] = torch.from_numpy(w.astype(np.float32))
            sd[f"{net}.{name}.bias"] = torch.from_numpy(b.astype(np.float32))
    return sd


def make_checkpoint(path, state_dict, epoch=0, global_step=0):
    """Write a PL-1.5.10-shaped checkpoint (no hyper_parameters: the reference never saves them)."""
    ckpt = {
        "epoch": epoch,
        "global_step": global_step,
        "pytorch-lightning_version": "1.5.10",
        "state_dict": {k: v.detach().cpu().clone() for k, v in state_dict.items()},
        "callbacks": {},
        "optimizer_states": [],
        "lr_schedulers": [],
    }
    torch.save(ckpt, str(path))
    return ckpt


def orbit_pose(theta_deg, phi_deg, radius):
    """c2w of a camera on a sphere looking at the origin; same construction as the orbit the reference
    renders (nerf_helpers.py:258-284), built in float64 and rounded once per factor like the reference."""
    def f32(m):
        return np.asarray(m, dtype=np.float64).astype(np.float32)
    th, ph = theta_deg / 180.0 * np.pi, phi_deg / 180.0 * np.pi
    t = f32([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]])
    rp = f32([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1]])
    rt = f32([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]])
    flip = f32([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    c2w = torch.from_numpy(rp) @ torch.from_numpy(t)
    c2w = torch.from_numpy(rt) @ c2w
    return torch.from_numpy(flip) @ c2w


def analytic_scene_rgba(c2w, H, W, focal):
    """RGBA uint8 image of three coloured spheres on transparent black, ray-cast analytically.
    Gives the trainer something with real structure when no Blender data is available."""
    c2w = np.asarray(c2w, dtype=np.float64)
    j, i = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    dirs = np.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -np.ones_like(i)], -1)
    d = dirs @ c2w[:3, :3].T
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    o = c2w[:3, 3]
    spheres = [((0.0, 0.0, 0.0), 0.9, (0.9, 0.25, 0.2)), ((0.9, 0.6, 0.3), 0.45, (0.2, 0.8, 0.3)),
               ((-0.7, -0.6, 0.5), 0.5, (0.25, 0.35, 0.95))]
    light = np.array([0.4, 0.5, 0.77]); light /= np.linalg.norm(light)
    best = np.full((H, W), np.inf)
    rgb = np.zeros((H, W, 3))
    for c, r, col in spheres:
        oc = o - np.array(c)
        b = d @ oc
        disc = b * b - (oc @ oc - r * r)
        t = -b - np.sqrt(np.maximum(disc, 0.0))
        hit = (disc > 0) & (t > 0) & (t < best)
        p = o + t[..., None] * d
        n = (p - np.array(c)) / r
        shade = 0.25 + 0.75 * np.clip(n @ light, 0.0, 1.0)
        rgb[hit] = (shade[..., None] * np.array(col))[hit]
        best = np.where(hit, t, best)
    alpha = np.isfinite(best)
    out = np.zeros((H, W, 4), dtype=np.uint8)
    out[..., :3] = np.clip(rgb * 255.0 + 0.5, 0, 255).astype(np.uint8)
    out[..., 3] = alpha.astype(np.uint8) * 255
    return out


def write_blender_scene(base_dir, n_train=8, n_val=2, n_test=2, H=800, W=800, camera_angle_x=0.6911112070083618):
    """Blender-synthetic-shaped dataset: transforms_{train,val,test}.json + RGBA PNGs, the schema
    `SyntheticDataset` reads (dataloader.py:105-141, tests/test_data/transforms_train.json)."""
    from PIL import Image
    base = Path(base_dir)
    focal = 0.5 * W / np.tan(0.5 * camera_angle_x)
    for split, n, off in (("train", n_train, 0.0), ("val", n_val, 7.0), ("test", n_test, 13.0)):
        (base / split).mkdir(parents=True, exist_ok=True)
        frames = []
        for k in range(n):
            theta = -180.0 + 360.0 * k / max(n, 1) + off
            phi = -30.0 - 25.0 * ((k * 7) % 5) / 5.0
            c2w = orbit_pose(theta, phi, 4.0)
            Image.fromarray(analytic_scene_rgba(c2w.numpy(), H, W, focal), "RGBA").save(base / split / f"r_{k}.png")
            frames.append({"file_path": f"./{split}/r_{k}", "rotation": 0.012566370614359171,
                           "transform_matrix": [[float(v) for v in row] for row in c2w.numpy()]})
        with open(base / f"transforms_{split}.json", "w") as fh:
            json.dump({"camera_angle_x": camera_angle_x, "frames": frames}, fh, indent=1)
    return base
