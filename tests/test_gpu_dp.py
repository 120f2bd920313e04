"""Data-parallel training on real GPUs (needs >= 2 visible devices: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`;
skipped on a single-GPU box).  SURVEY.md 8(e): identical replicas, different batches per rank, one summed gradient.

  * the product path (`train_nerf.py full` under torchrun, Trainer.fit) starts from UNSEEDED, hence different, NeRFNetwork()s on
    the two ranks: after the broadcast of Trainer.synchronize_replicas and 10 steps the flat parameter buffers of the two ranks
    are BIT-identical, although the ranks drew different pixels;
  * the same number of steps on batches that are a pure function of (rank, step) equals - up to the order of the fp32 atomics
    in wgrad - a single-process run that pushes both ranks' batches through one gradient buffer with grad_scale = 1/2.
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import synthetic

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent
STEPS = 10


def _torchrun(nproc, args, timeout=900):
    env = dict(os.environ, NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29653", str(HERE / "dp_worker.py")] + [str(a) for a in args]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.fixture(scope="module")
def two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_product_path_replicas_are_bit_identical(two_gpus, tmp_path):
    scene = tmp_path / "scene"
    synthetic.write_blender_scene(scene, n_train=4, n_val=1, n_test=1)
    _torchrun(2, ["entry", tmp_path, STEPS, scene])
    a, b = (torch.load(tmp_path / f"entry_rank{r}.pt") for r in (0, 1))
    assert not torch.equal(a["before"], b["before"]), "the two processes were expected to initialise different networks"
    assert torch.equal(a["after"], b["after"]) and torch.equal(a["after"], a["before"])      # rank 0's parameters win
    assert a["step"] == b["step"] == STEPS
    assert not torch.equal(a["draws"], b["draws"]), "ranks must draw different pixels (seed + rank)"
    assert torch.equal(a["final"], b["final"]), f"replicas diverged: max |d| {(a['final'] - b['final']).abs().max():.3e}"
    assert torch.equal(a["m"], b["m"])
    assert (a["final"] - a["after"]).abs().max() > 1e-4                                       # and they did train
    # one writer: rank 0's metrics file only
    assert (tmp_path / "exp" / "NeRF" / "dp" / "metrics.jsonl").exists()


def test_two_ranks_equal_one_process_on_the_concatenated_batch(two_gpus, tmp_path):
    _torchrun(2, ["dp", tmp_path, STEPS])
    res = subprocess.run([sys.executable, str(HERE / "dp_worker.py"), "single", str(tmp_path), str(STEPS), "2"], capture_output=True,
                         text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    r0, r1, one = (torch.load(tmp_path / f) for f in ("dp_rank0.pt", "dp_rank1.pt", "single_rank0.pt"))
    assert torch.equal(r0["final"], r1["final"])
    assert r0["grad_scale"] == 0.5 and one["grad_scale"] == 0.5
    init = torch.cat([v.flatten() for v in synthetic.make_state_dict(0, "init").values()])
    n = init.numel()
    upd_dp, upd_one = r0["final"][:n] - init, one["final"][:n] - init
    assert upd_dp.abs().max() > 1e-3
    # Adam turns each gradient into a step of about lr whatever its size, so where a gradient is ~0 the atomics' summation order
    # can flip a step: compare the update VECTORS (cosine, relative norm of the difference), not elements
    cos = torch.nn.functional.cosine_similarity(upd_dp.double(), upd_one.double(), dim=0)
    rel = (upd_dp - upd_one).double().norm() / upd_one.double().norm()
    assert cos > 0.999 and rel < 0.05, (float(cos), float(rel))
    assert abs(r0["losses"][0] - one["losses"][0]) < 1.0            # (different batches: rank 0's loss vs the last rank's)
