"""End-to-end demonstration through the reference's entry points: train_nerf.py `full` on a synthetic Blender-shaped scene, then
score.py (test-set PSNR / SSIM through view_reconstruction) on the checkpoint it wrote.  Prints one JSON line.
usage: python tools/train_demo.py [--steps 6000] [--views 24]      (torchrun --nproc-per-node N ... for data-parallel training)"""
import argparse, json, os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "cse-573-minimal-nerf_b200")]
import torch
import score, synthetic, train_nerf

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6000)
ap.add_argument("--views", type=int, default=24)
ap.add_argument("--crop-epochs", type=int, default=20)
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
work = Path(tempfile.gettempdir()) / "nerf_b200_train_demo"
ready = work / f"scene_ready_{args.views}_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
if rank == 0:
    synthetic.write_blender_scene(work / "scene", n_train=args.views, n_val=1, n_test=3)
    ready.touch()
while not ready.exists():                  # the other ranks: the scene is written once, by rank 0 (no process group exists yet)
    time.sleep(0.2)
torch.manual_seed(0)
t0 = time.time()
run = train_nerf.main(["-n", "demo", "--gpu", "-s", str(args.steps), "-rd", str(work / "exp"), "-r", "4096", "full", "-b", str(work / "scene"),
                       "-cr", str(args.crop_epochs)])
torch.cuda.synchronize()
train_s = time.time() - t0
import multi_gpu
import torch.distributed as dist
if rank == 0:
    # rank 0 alone scores: view_reconstruction must not shard the frames over ranks that are not rendering (an all-gather nobody
    # else joins would hang - measured the hard way)
    with multi_gpu.local_only():
        psnr, ssim = score.calculate_scores(run.last_checkpoint, work / "scene", 4096)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    print(json.dumps({"what": "train_nerf.py full on the analytic three-sphere scene (800x800, Blender layout), then score.py on its test split",
                      "train_views": args.views, "steps": run.global_step, "rays_per_step": 4096 * world, "n_gpus": world, "train_wall_s": train_s,
                      "ms_per_step_wall": train_s / max(run.global_step, 1) * 1e3, "final_train_loss": float(run.metrics().get("train_loss", float("nan"))),
                      "test_psnr_db": psnr, "test_ssim": ssim, "checkpoint": run.last_checkpoint.name}))
if dist.is_available() and dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
