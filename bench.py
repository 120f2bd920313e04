#!/usr/bin/env python
"""bench.py - rays/s of the NeRF render hot path on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hw 800]

One "step" = one H x W novel-view frame (default 800x800 = 640 000 rays) pushed through NeRFNetwork.forward by
nerf_helpers.render_rays_chunked (the reference's 4096-ray chunks grouped into one set of launches per frame): stratified
sampling -> coarse MLP -> compositing -> inverse-CDF sampling -> merge sort -> fine MLP -> compositing (64 coarse + 128 fine
samples per ray), weights from a synthetic checkpoint in the reference's format (the shipped lego checkpoint is not available
offline).

  value     device-resident rays, CUDA-event time per step (L2 flushed between steps), max over ranks
  e2e       the same frame through the public call, nerf_helpers.view_reconstruction, with HOST buffers: pinned rays H2D,
            render, uint8 conversion, image D2H inside the timed region (weak scaling at N > 1: one frame per rank)
  strong    N > 1 only: ONE frame's rays sharded over the N ranks through view_reconstruction, incl. the all-gather of the
            uint8 slabs and the D2H copy (BASELINE.json configs[4] per frame)
  roofline  fused tcgen05 MLP kernel: algorithmic FLOPs / CUDA-event time of its launches inside the timed region
  train     4096-ray training steps (configs[2]/[3]) + `roofline_train`: per-kernel time, bytes and fraction of peak
  cpu_baseline  the CPU oracle (port of the reference) on this box's host cores, bounded sample (N=1 only), with the
            PSNR of this repo's render of the same chunk against it (`parity`)

`--impl reference` times the reference's CPU path (oracle port; the Python reference cannot travel to the GPU
box) with all host threads on the same metric.  Multi-GPU (torchrun): weak scaling, one frame of the orbit per
rank per step, no data-path collective; the only exchange is the final image gather in the e2e arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

FLOP_PER_SAMPLE = 2 * 460416          # true (unpadded) MACs of one NeRFModel forward x 2
CAM_ANGLE_X = 0.6911112070083618
COARSE, FINE, CHUNK = 64, 128, 4096


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return {"tflops_burst": p["bf16_tflops"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def mark(self, wait_s=3.0):
        """Start of the timed region: wait (bounded) until nvidia-smi has delivered its first sample - with 8 ranks each starting
        one it can take longer than a short timed region - and count only the samples taken from here on."""
        t0 = time.time()
        while not self.rows and time.time() - t0 < wait_s:
            time.sleep(0.02)
        self.first = max(len(self.rows) - 1, 0)

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def frame_setup(H, W, pose_index):
    import synthetic
    focal = 0.5 * W / np.tan(0.5 * CAM_ANGLE_X)          # nerf_helpers.py:164 (product arm: nothing from oracle/)
    angle = float(np.linspace(-180, 180, 41)[:-1][pose_index % 40])
    return synthetic.orbit_pose(angle, -30.0, 4.0), focal


def run_reference(args, rank, world):
    """Reference arm: the CPU oracle (port of the reference's forward) with all host threads."""
    if rank != 0:
        return
    import synthetic
    from oracle import nerf_oracle as O
    torch.set_num_threads(os.cpu_count())
    on_gpu = args.reference_device == "cuda"
    sd = synthetic.make_state_dict(5, "dense")
    c2w, focal = frame_setup(args.hw, args.hw, 0)
    o, d = O.get_rays(args.hw, args.hw, focal, c2w)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    if on_gpu:                                         # the port's factory calls (arange, full, ...) follow the default device
        torch.set_default_device("cuda")
        sd = {k: v.cuda() for k, v in sd.items()}
        o, d = o.cuda(), d.cuda()
    n = min(CHUNK, o.shape[0])                         # bounded sample: one 4096-ray chunk per step

    def step(k):
        lo = (k * n) % max(o.shape[0] - n + 1, 1)
        rand = (torch.rand(n, COARSE), torch.rand(n, 1), torch.rand(n, FINE, 1))
        with torch.no_grad():
            O.network_forward(sd, o[lo:lo + n], d[lo:lo + n], *rand)
        if on_gpu:
            torch.cuda.synchronize()
    for k in range(args.warmup):
        step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(args.warmup + k)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = n / dt
    train = None
    if on_gpu:          # the reference's training step (forward, loss, autograd, torch.optim.Adam) in eager PyTorch on the GPU
        leaf = {k: t.clone().requires_grad_(True) for k, t in sd.items()}
        opt = torch.optim.Adam(list(leaf.values()), lr=5e-4)
        target = torch.rand(n, 3)

        def tstep():
            rand = (torch.rand(n, COARSE), torch.rand(n, 1), torch.rand(n, FINE, 1))
            loss, _ = O.training_loss(leaf, o[:n], d[:n], target, *rand)
            opt.zero_grad()
            loss.backward()
            opt.step()
            torch.cuda.synchronize()
        for _ in range(2):
            tstep()
        t1 = time.perf_counter()
        for _ in range(max(args.steps, 1)):
            tstep()
        tdt = (time.perf_counter() - t1) / max(args.steps, 1)
        train = {"metric": "rays/sec train", "value": n / tdt, "unit": "rays/s", "ms_per_step": tdt * 1e3,
                 "what": "forward + MSE losses + autograd backward + torch.optim.Adam on one 4096-ray batch, eager PyTorch fp32"}
    sample = (f"one {n}-ray chunk of the {args.hw}x{args.hw} frame per step (no_grad), oracle port of the reference "
              + ("as eager PyTorch fp32 on the GPU (torch ops, cuBLAS GEMMs; wall clock with a synchronize per step)" if on_gpu
                 else "on the CPU"))
    emit({
        "impl": "reference", "metric": "rays/sec render (device-timed)", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": 0 if on_gpu else os.cpu_count(), "kind": "port", "sample": sample},
        "reference_device": args.reference_device, "train": train,
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def workload_config(args):
    return {"workload": f"render {args.hw}x{args.hw} lego-360 orbit frame, 64 coarse + 128 fine samples/ray, ray chunk 4096 "
                        "(chunks grouped into one set of launches per frame by nerf_helpers.render_rays_chunked), "
                        "synthetic checkpoint in the reference's PL format (dense weight set, seed 5)",
            "H": args.hw, "W": args.hw, "chunk": CHUNK, "coarse": COARSE, "fine": FINE,
            "l2": "256 MB L2 flush between timed steps; per-step uniforms (0.49 GB) exceed the 126 MB L2"}


def cpu_baseline(args, net=None):
    """The CPU oracle on a bounded sample of the workload (also the one place this file checks the product against it: the
    chunks the oracle renders are rendered by `net` with the SAME uniforms and the PSNR between the two goes into `parity`)."""
    import synthetic
    from oracle import nerf_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = synthetic.make_state_dict(5, "dense")
    c2w, focal = frame_setup(args.hw, args.hw, 0)
    o, d = O.get_rays(args.hw, args.hw, focal, c2w)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    n = min(CHUNK, o.shape[0])
    chunks = 3
    t_best, mse, max_abs = [], [], 0.0
    with torch.no_grad():
        for k in range(chunks):
            lo = (k * 7919 * n) % max(o.shape[0] - n + 1, 1)
            rand = (torch.rand(n, COARSE), torch.rand(n, 1), torch.rand(n, FINE, 1))
            t0 = time.perf_counter()
            ref = O.network_forward(sd, o[lo:lo + n], d[lo:lo + n], *rand)
            t_best.append(time.perf_counter() - t0)
            if net is not None:
                dev = next(net.parameters()).device
                got = net.forward(o[lo:lo + n].contiguous().to(dev), d[lo:lo + n].contiguous().to(dev), rand=tuple(r.to(dev) for r in rand))
                diff = (got["fine_rgb_rays"].cpu() - ref["fine_rgb_rays"]).double()
                mse.append(float((diff ** 2).mean()))
                max_abs = max(max_abs, float(diff.abs().max()))
    dt = float(np.mean(t_best[1:]))                     # first chunk is warm-up
    out = {"value": n / dt, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
           "sample": f"{chunks - 1} timed 4096-ray chunks (+1 warm-up) of the same frame through the CPU oracle, no_grad, "
                     f"{os.cpu_count()} threads"}
    if mse:
        m = float(np.mean(mse))
        out["parity"] = {"what": f"fine ray colours of {chunks} x {n} rays, this repo (bf16 tensor-core path) vs the CPU oracle, same uniforms",
                         "psnr_db": float(10 * np.log10(1.0 / max(m, 1e-20))), "max_abs": max_abs}
    return out


TRAIN_FLOP_PER_SAMPLE = 2 * (460416 + 460416 + 426624)     # fwd + wgrad + needed dgrad
# algorithmic HBM bytes per sample of the training kernels: bf16 activations 1920 x 2 + sign words 60 x 4 written
# by the forward; sign words read + dz 1936 x 2 written by dgrad; acts + dz read once by wgrad
TRAIN_BYTES = {"mlp_tc_kernel(train)": 3840 + 240 + 16, "mlp_tc_bwd_kernel": 240 + 3872 + 16, "wgrad_tc_kernel": 3840 + 3872}
TRAIN_FLOPS = {"mlp_tc_kernel(train)": 2 * 460416, "mlp_tc_bwd_kernel": 2 * 426624, "wgrad_tc_kernel": 2 * 460416}


def bench_train(args, rank, world, dev):
    """BASELINE.json configs[2]/[3]: training steps on 4096-ray batches (per GPU), random-init coarse + fine MLP, Adam;
    data-parallel across ranks with one summed gradient per step (trainer.FlatGradients).  Batches are drawn from the centre
    crop of analytic-scene images (what the reference does for its first `cropping_epochs`, dataloader.py:13-34), two orbit
    poses per rank, so that the network sees the object and the loss moves."""
    import torch.distributed as dist
    import _native as nat
    import dataloader
    import nerf_model
    import synthetic
    from trainer import FlatGradients
    torch.manual_seed(1234 + rank)
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(0, "init"))
    net = net.to(dev)
    opt = net.configure_optimizers()["optimizer"]
    grads = FlatGradients(net.parameters(), opt)
    n_coarse = sum(p.numel() for p in net.coarse_network.parameters())
    if world > 1:
        net.on_coarse_grads_ready = lambda: grads.reduce_async(n_coarse)
    H = W = 800
    import training
    images, poses = [], []
    for j in range(2):
        c2w, focal = frame_setup(H, W, 7 * rank + 3 + 20 * j)
        poses.append(c2w.to(torch.float32))
        images.append(torch.from_numpy(synthetic.analytic_scene_rgba(c2w.numpy(), H, W, focal)[..., :3].copy()))
    images, poses = torch.stack(images).to(dev).contiguous(), torch.stack(poses).to(dev).contiguous()
    n = CHUNK
    # the product's training step: Trainer.fit's body captured once as a CUDA graph (training.GraphedTrainStep) and replayed
    stepper = training.GraphedTrainStep(net, opt, grads, images, poses, focal, n, cropping=True, warmup=max(args.warmup, 3))
    counter = [0]

    def step():
        counter[0] += 1
        return stepper.eager_step(counter[0] % 2) if args.eager_train else stepper.step(counter[0] % 2)
    loss_first = stepper.first_loss            # loss of the very first optimiser step (first warm-up step before the capture)
    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(args.steps, 1) * 4
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        loss = step()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    loss_last = float(loss.detach())
    nz = net.logged.get("fine_density_non_zeros") if hasattr(net, "logged") else None
    # ---- per-kernel record (a few extra steps with CUDA events around the three tensor-core kernels; not part of the timing above)
    # Each measured step is queued behind a ~25 ms device-side delay (torch.cuda._sleep), so that the host has issued the whole
    # step before its first kernel starts: the kernels then run back to back and an event pair brackets the kernel alone.  Without
    # it the eager step is host-bound, the GPU idles between launches and every interval also counts the host's launch path
    # (+0.1 ms per kernel, profiles/r02_bench_final2.json: the six intervals summed to 1.14 x the replayed step).
    # The same step is bracketed as a whole (`eager_ms`): the kernels' share is taken of THAT step, measured in the same pass.
    nat.kernel_events = []
    brackets = []
    for k in range(4):
        torch.cuda._sleep(40_000_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        stepper.eager_step(k % 2)
        b.record()
        brackets.append((a, b))
    torch.cuda.synchronize()
    eager_ms = sum(a.elapsed_time(b) for a, b in brackets) / len(brackets)
    events, nat.kernel_events = nat.kernel_events, None
    pk = peaks()
    per = {}
    for name, units, a, b in events:
        e = per.setdefault(name, {"ms": 0.0, "samples": 0})
        e["ms"] += a.elapsed_time(b) / 4
        e["samples"] += units // 4
    # DRAM bytes per kernel from the committed ncu --set full capture of the same step shape (two launches per kernel and step:
    # coarse + fine network - summed, like the algorithmic bytes beside them)
    dram, dram_src = {}, None
    for tf_ in sorted((ROOT / "profiles").glob("r*_train_traffic.json"), reverse=True):
        rec = json.loads(tf_.read_text())
        if rec.get("rays_per_step") == n and rec.get("coarse_samples") == COARSE and rec.get("fine_samples") == COARSE + FINE:
            dram = {k: sum(v["dram_bytes_per_launch"]) for k, v in rec["kernels"].items()}
            dram_src = f"profiles/{tf_.name}: {rec.get('what', '')}"
            break
    kernels = []
    for name, e in per.items():
        gb = e["samples"] * TRAIN_BYTES.get(name, 0) / 1e9
        tf = e["samples"] * TRAIN_FLOPS.get(name, 0) / 1e12
        sec = e["ms"] * 1e-3
        kernels.append({"kernel": name, "ms_per_step": e["ms"], "algorithmic_gb": gb, "hbm_gbs": gb / sec, "hbm_frac": gb / sec / pk["hbm_gbs"],
                        "tflops": tf / sec, "tensor_frac_of_sustained": tf / sec / pk["tflops_sustained"], "traffic": dram.get(name)})
    torch.cuda.synchronize()
    stepper.close()                               # destroys the CUDA graph (and the NCCL work captured in it) before the process group goes
    kms = sum(k["ms_per_step"] for k in kernels)
    tfl = n * 256 * TRAIN_FLOP_PER_SAMPLE / (ms * 1e-3) / 1e12
    dom = max(kernels, key=lambda k: k["ms_per_step"]) if kernels else None
    roofline_train = None
    if dom is not None:
        hbm_bound = dom["hbm_frac"] >= dom["tensor_frac_of_sustained"]
        roofline_train = {"bound": "hbm" if hbm_bound else "tensor", "kernel": dom["kernel"],
                          "achieved": dom["hbm_gbs"] if hbm_bound else dom["tflops"], "peak": pk["hbm_gbs"] if hbm_bound else pk["tflops_sustained"],
                          "unit": "GB/s" if hbm_bound else "TFLOP/s", "frac": dom["hbm_frac"] if hbm_bound else dom["tensor_frac_of_sustained"],
                          "traffic": dom.get("traffic"), "traffic_source": dram_src, "kernels": kernels, "kernel_ms_per_step": kms, "kernel_share_of_step": kms / eager_ms,
                          "measured_in": "4 launch-by-launch steps queued behind a device-side delay (kernels back to back), CUDA events per kernel",
                          "launch_by_launch_step_ms": eager_ms, "kernel_ms_over_replayed_step": kms / ms,
                          "step_algorithmic_tflops": tfl, "step_frac_of_sustained_peak": tfl / pk["tflops_sustained"],
                          "step_frac_of_burst_peak": tfl / pk["tflops_burst"]}
    return {"metric": "rays/sec train (device-timed)", "value": n * world / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
            "steps": steps, "rays_per_step_per_gpu": n, "loss_first": float(loss_first), "loss_last": loss_last, "final_loss": loss_last,
            "fine_density_non_zeros": float(nz) if nz is not None else None,
            "algorithmic_tflops_per_gpu": tfl, "frac_of_sustained_peak": tfl / pk["tflops_sustained"],
            "frac_of_burst_peak": tfl / pk["tflops_burst"], "roofline_train": roofline_train,
            "data": "centre-cropped pixels of two analytic-scene orbit views per rank (synthetic), random-init weights (seed 0)",
            "step": "eager (launch by launch)" if args.eager_train else "training.GraphedTrainStep: the whole step replayed as one CUDA graph",
            "backward": "hand-written: composite_backward_kernel, mlp_tc_bwd3_kernel (tcgen05 dgrad chain), "
                        "wgrad_tc_kernel (tcgen05 wgrad + bias sums); Adam = hand-written flat kernel (adam.cu), gradients accumulated straight into the flat buffer",
            "collective": ("NCCL sum of the 924 680-float flat gradient buffer per step (coarse slice overlapped with the fine backward), "
                           "1/world folded into the Adam kernel") if world > 1 else None}


def emit(line):
    """The one JSON line goes to the real stdout; everything else libraries print (e.g. NCCL's version banner, which is a
    C-level printf) has been routed to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                                # fd 1 -> stderr for the rest of the process (NCCL_DEBUG output included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hw", type=int, default=800, help="frame height = width")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--eager-train", action="store_true", help="time the training step launch by launch instead of as the replayed CUDA graph")
    ap.add_argument("--reference-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' runs the same eager-PyTorch port of the reference on the GPU (what the "
                         "reference itself does when a GPU is present); the contract arm is the default, 'cpu'")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import _native as nat
    import dataloader
    import multi_gpu
    import nerf_helpers
    import nerf_model
    import synthetic

    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(5, "dense"))
    net = net.to(dev)
    H = W = args.hw
    nrays = H * W
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = torch.empty((nrays, 3), device=dev, dtype=torch.float32)
    host_o = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
    host_d = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()

    def rays_for(step, per_rank=True):
        c2w, focal = frame_setup(H, W, step * world + rank if per_rank else step)
        return dataloader.get_rays(H, W, focal, c2w, device=dev)                # [H,W,3] each

    def render(o, d):                         # the chunk loop of view_reconstruction (public API), fine colours into `out`
        nerf_helpers.render_rays_chunked(net, o.reshape(nrays, 3), d.reshape(nrays, 3), CHUNK, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident arm ("value")
    o, d = rays_for(0)
    for _ in range(args.warmup):
        render(o, d)
    barrier()
    nat.launches = 0
    nat.kernel_events = []
    step_ms = []
    with ClockSampler(local_rank) as clocks:
        render(o, d)                          # one more untimed frame: the GPU is under load while the sampler comes up
        clocks.mark()
        barrier()
        nat.launches = 0
        nat.kernel_events = []
        for k in range(args.steps):
            o, d = rays_for(k)
            flush.fill_(k & 0xFF)
            barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            render(o, d)
            t1.record()
            barrier()
            step_ms.append(t0.elapsed_time(t1))
    launches = nat.launches
    events, nat.kernel_events = nat.kernel_events, None
    ms = torch.tensor([float(np.mean(step_ms))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item())
    value = nrays * world / (ms_per_step * 1e-3)

    mlp_ms = sum(a.elapsed_time(b) for _, _, a, b in events)
    mlp_samples = sum(u for _, u, _, _ in events)
    pk = peaks()
    achieved = mlp_samples * FLOP_PER_SAMPLE / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    traffic, traffic_src = None, None
    for tf in sorted((ROOT / "profiles").glob("r*_mlp_tc_traffic.json"), reverse=True):      # ncu --set full of the fine network's launch
        rec = json.loads(tf.read_text())
        if rec.get("rays_per_launch") == min(nrays, nerf_helpers.RAYS_PER_LAUNCH or CHUNK):       # same launch shape as this run
            traffic, traffic_src = rec.get("dram_bytes_per_launch"), f"profiles/{tf.name}: {rec.get('what', '')}"
            break
    roofline = {"bound": "tensor", "kernel": "mlp_tc3_kernel", "achieved": achieved, "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": achieved / pk["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"{pk['source']} MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step); "
                               f"burst {pk['tflops_burst']}",
                "frac_of_burst": achieved / pk["tflops_burst"], "launches_timed": len(events),
                "kernel_share_of_step": mlp_ms / (sum(step_ms) or 1.0),
                "step_frac_of_sustained": nrays * 256 * FLOP_PER_SAMPLE / (ms_per_step * 1e-3) / 1e12 / pk["tflops_sustained"],
                "step_frac_of_burst": nrays * 256 * FLOP_PER_SAMPLE / (ms_per_step * 1e-3) / 1e12 / pk["tflops_burst"],
                "flop_per_sample": FLOP_PER_SAMPLE}

    # ------------------------------------------------------------------ end-to-end arm (host buffers, the public call)
    def e2e_step(k, sharded):
        o, d = rays_for(k, per_rank=not sharded)      # stands in for the reference's CPU get_rays: produce host rays first
        host_o.copy_(o); host_d.copy_(d)
        torch.cuda.synchronize()
        flush.fill_(k & 0xFF)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        od = host_o.to(dev, non_blocking=True)           # nerf_helpers.py:184: o_rays.to(device), d_rays.to(device)
        dd = host_d.to(dev, non_blocking=True)
        if sharded:                                      # one frame over all ranks: slabs + all-gather inside view_reconstruction
            im = nerf_helpers.view_reconstruction(net, od, dd, N=CHUNK)
        else:
            with multi_gpu.local_only():                 # weak scaling: every rank renders its own frame
                im = nerf_helpers.view_reconstruction(net, od, dd, N=CHUNK)          # uint8 numpy [H,W,3] on the host
        t1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        barrier()
        assert im.shape == (H, W, 3) and im.dtype == np.uint8
        return t0.elapsed_time(t1), wall * 1e3

    def e2e_run(sharded):
        """(mean, median, per-step list) of max(device time, wall time) per step, each the max over ranks."""
        e2e_step(0, sharded)
        e2e = [e2e_step(k, sharded) for k in range(args.steps)]
        t = torch.tensor([max(a, b) for a, b in e2e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_step = [float(x) for x in t.tolist()]
        return float(np.mean(per_step)), float(np.median(per_step)), per_step
    e_ms, e_med, e_steps = e2e_run(False)
    e2e_value = nrays * world / (e_ms * 1e-3)
    strong = None
    if world > 1:
        s_ms, s_med, s_steps = e2e_run(True)
        strong = {"what": f"ONE {H}x{W} frame, rays sharded over {world} GPUs through nerf_helpers.view_reconstruction (host rays in, "
                          "uint8 all-gather, host image out; BASELINE.json configs[4] per frame)",
                  "ms_per_frame": s_ms, "rays_per_s": nrays / (s_ms * 1e-3), "speedup_vs_one_rank_e2e": e_ms / s_ms, "n_gpus": world,
                  "ms_per_frame_median": s_med, "ms_steps": s_steps, "scaling": "strong"}

    train = None
    if not args.no_train:
        try:
            train = bench_train(args, rank, world, dev)
        except Exception as exc:                 # the render line must not be lost to a failure of the extra training section
            if world > 1:
                raise                            # (ranks must not diverge: under torchrun a failure stays a failure)
            import traceback
            traceback.print_exc()
            train = {"metric": "rays/sec train (device-timed)", "value": None, "error": f"{type(exc).__name__}: {exc}"[:500]}

    if rank == 0:
        line = {
            "metric": "rays/sec render (device-timed)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(host_o.numel() * 4 * 2),
                    "d2h_bytes_per_step": int(H * W * 3), "ms_per_step": e_ms, "ms_per_step_median": e_med, "ms_steps": e_steps,
                    "call": "nerf_helpers.view_reconstruction(model, o.to(device), d.to(device), N=4096) -> host uint8 image"},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks.summary(), "train": train,
        }
        if train is not None and train.get("roofline_train") is not None:
            line["roofline_train"] = train["roofline_train"]
        if strong is not None:
            line["strong"] = strong
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, net)
            line["parity"] = line["cpu_baseline"].get("parity")
        emit(line)
    if world > 1:
        # the line is out; nothing below may keep the job alive (a stuck communicator teardown would stall the driver's run)
        watchdog = threading.Timer(60.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()


if __name__ == "__main__":
    main()
