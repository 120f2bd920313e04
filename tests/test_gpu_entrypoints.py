"""GPU end-to-end checks of the reference's entry points on a synthetic Blender-shaped scene: train_nerf.py `full`
(Trainer shim, cropping switch, PL-format checkpoint), resume from that checkpoint, render.py (checkpoint -> orbit GIF),
and the dataset item contract of dataloader.SyntheticDataset."""
import sys

import numpy as np
import pytest
import torch

import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(tmp_path_factory):
    base = tmp_path_factory.mktemp("scene")
    synthetic.write_blender_scene(base, n_train=3, n_val=1, n_test=1)
    return base


def test_dataset_items_match_reference_contract(scene):
    import dataloader
    from oracle import nerf_oracle as O
    ds = dataloader.SyntheticDataset(scene, "train", 512)
    assert len(ds) == 3 and abs(ds.focal - O.focal_from_fov(800, 0.6911112070083618)) < 1e-9
    item = ds[1]
    assert set(item) == {"origin", "direc", "rgb", "xs", "ys"}                      # dataloader.py:155
    assert item["origin"].shape == (512, 3) and item["direc"].shape == (512, 3) and item["rgb"].shape == (512, 3)
    assert item["rgb"].min() >= 0 and item["rgb"].max() <= 1
    # rays of the sampled pixels are bit-identical to the oracle's full-grid rays gathered at [ys, xs]
    c2w = torch.tensor(ds.frames[1]["transform_matrix"], dtype=torch.float32)
    o, d = O.get_rays(800, 800, ds.focal, c2w, item["xs"].cpu(), item["ys"].cpu())
    assert torch.equal(item["direc"].cpu(), d) and torch.equal(item["origin"].cpu(), o.contiguous())
    val = dataloader.SyntheticDataset(scene, "val", 64)[0]
    assert {"all_origin", "all_direc", "image"} <= set(val) and val["all_direc"].shape == (800, 800, 3)
    crop = dataloader.SyntheticDataset(scene, "train", 2048, cropping=True)[0]
    assert crop["xs"].min() >= 200 and crop["xs"].max() < 600 and crop["ys"].min() >= 200 and crop["ys"].max() < 600


def test_train_then_render(scene, tmp_path):
    import render
    import train_nerf
    args = train_nerf.build_parser().parse_args(["-n", "unit", "--gpu", "-s", "12", "-rd", str(tmp_path), "-r", "1024", "full",
                                                 "-b", str(scene), "-cr", "1"])
    trainer = train_nerf.train_full_nerf(args.root_dir, args.base_dir, args.name, args.steps, args.position_encoding,
                                         args.direction_encoding, args.gpu, args.rays, args.coarse, args.fine, args.near,
                                         args.far, args.cropping_epochs, args.ckpt, args)
    assert trainer.global_step == 12 and trainer.current_epoch == 4                 # 3 train images per epoch
    ckpt = trainer.last_checkpoint
    assert ckpt is not None and ckpt.name == "epoch=3-step=11.ckpt"                 # PL naming; render.py parses 'epoch=...-'
    blob = torch.load(str(ckpt), map_location="cpu", weights_only=False)
    assert {"epoch", "global_step", "pytorch-lightning_version", "state_dict", "optimizer_states", "lr_schedulers"} <= set(blob)
    assert list(blob["state_dict"].keys()) == synthetic.state_dict_keys() and "hyper_parameters" not in blob
    first = [l for l in open(tmp_path / "NeRF" / "unit" / "metrics.jsonl")]
    assert any("hyperparams" in l for l in first)
    # resume (train_nerf.py -l): continues the step count and keeps training
    args2 = train_nerf.build_parser().parse_args(["-n", "unit2", "--gpu", "-s", "15", "-rd", str(tmp_path), "-r", "512", "-l", str(ckpt),
                                                  "full", "-b", str(scene), "-cr", "0"])
    t2 = train_nerf.train_full_nerf(args2.root_dir, args2.base_dir, args2.name, args2.steps, 10, 4, True, args2.rays, 64, 128, 2.0, 6.0,
                                    args2.cropping_epochs, args2.ckpt, args2)
    assert t2.global_step == 15
    # render.py: checkpoint -> 2-pose orbit at reduced resolution through the same code path
    import nerf_helpers
    import nerf_model
    out = tmp_path / "recons"
    out.mkdir()
    model = nerf_model.NeRFNetwork.load_from_checkpoint(str(ckpt)).to("cuda")
    views = nerf_helpers.generate_360_view_synthesis(model, out, "epoch=3", height=64, width=64, N=4096, num_poses=2)
    assert (out / "epoch=3-360.gif").exists() and len(views) == 2
    assert views[0].shape == (64, 64, 3) and views[0].dtype == np.uint8
    epoch = str(ckpt)[str(ckpt).find("epoch="):]
    assert epoch[:epoch.find("-")] == "epoch=3" == render.epoch_tag(ckpt)           # render.py:15-16 name parsing
    # the entry point itself: render(ckpt, save_dir, rays, num_poses) at the reference's 800 x 800 (one pose)
    out2 = tmp_path / "recons2"
    out2.mkdir()
    frames = render.render(str(ckpt), out2, 4096, 1)
    assert (out2 / "epoch=3-360.gif").exists() and len(frames) == 1 and frames[0].shape == (800, 800, 3)


def test_score_metrics_match_oracle_and_entry_point(scene, tmp_path):
    """score.py (score.py:20-41): device PSNR / SSIM against the scikit-image restatement in oracle/score_oracle.py on a random
    pair and on a rendered-vs-ground-truth pair; then the `calculate_scores(ckpt, base_dir, rays)` entry point end to end."""
    import dataloader
    import nerf_helpers
    import nerf_model
    import score
    from oracle import score_oracle as S
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=(96, 120, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int16) + rng.integers(-20, 21, size=a.shape), 0, 255).astype(np.uint8)
    assert abs(score.peak_signal_noise_ratio(a, b) - S.peak_signal_noise_ratio(a, b)) < 1e-9
    assert abs(score.structural_similarity(a, b, multichannel=True) - S.structural_similarity(a, b)) < 1e-9
    assert abs(score.structural_similarity(torch.from_numpy(a).cuda(), torch.from_numpy(a).cuda()) - 1.0) < 1e-12
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(5, "dense"))
    net = net.cuda()
    item = dataloader.SyntheticDataset(scene, "test", 1024)[0]
    o, d = item["all_origin"][300:420, 300:460].contiguous(), item["all_direc"][300:420, 300:460].contiguous()
    recon = nerf_helpers.view_reconstruction(net, o, d, N=4096)
    gt = (item["image"][300:420, 300:460] * 255).clamp(0, 255).to(torch.uint8)
    assert abs(score.peak_signal_noise_ratio(gt, recon) - S.peak_signal_noise_ratio(gt.cpu().numpy(), recon)) < 1e-9
    assert abs(score.structural_similarity(gt, recon) - S.structural_similarity(gt.cpu().numpy(), recon)) < 1e-9
    ckpt = tmp_path / "model=synthetic-epoch=0-step=0.ckpt"
    synthetic.make_checkpoint(ckpt, synthetic.make_state_dict(5, "dense"))
    psnr, ssim = score.calculate_scores(str(ckpt), scene, 8192)
    assert np.isfinite(psnr) and 0.0 < psnr < 60.0 and -1.0 <= ssim <= 1.0


def test_validation_step_and_logged_reconstruction(scene, tmp_path):
    """The every-n-epochs validation of train_nerf.py:26-29 / nerf_model.py:171-205 with n = 1: validation_step runs on the val
    image, logs val_loss / val_fine_loss / val_coarse_loss, renders the full 800x800 frame through view_reconstruction and hands it
    to the logger (recon_*.png), then training continues and the checkpoint is written."""
    import json
    import dataloader
    import nerf_model
    from trainer import JsonLogger, Trainer
    torch.manual_seed(5)
    logger = JsonLogger(name="val", project="NeRF", save_dir=tmp_path)
    net = nerf_model.NeRFNetwork()
    net.load_state_dict(synthetic.make_state_dict(4, "dense"))
    net.max_idx = 0        # upstream draws the logged view from randint(0, max_idx) with max_idx starting at 1 (nerf_model.py:171-180): with
                           # ONE validation image that logs a frame only every other time; 0 makes the draw land on the image we have
    scene_dm = dataloader.SyntheticDataModule(scene, 256, cropping_epochs=1)
    run = Trainer(gpus=1, default_root_dir=tmp_path, max_steps=6, logger=logger, check_val_every_n_epoch=1, track_grad_norm=2,
                  log_every_n_steps=1)
    run.fit(net, datamodule=scene_dm)
    assert run.global_step == 6 and run.current_epoch == 2                            # 3 train images per epoch
    pngs = sorted((tmp_path / "NeRF" / "val").glob("recon_*.png"))
    assert len(pngs) >= 1
    from PIL import Image
    im = np.asarray(Image.open(pngs[0]))
    assert im.shape == (800, 800, 3) and im.dtype == np.uint8 and im.max() > 0
    for key in ("val_loss", "val_fine_loss", "val_coarse_loss"):
        assert key in net.logged and np.isfinite(float(net.logged[key])), key
    lines = [json.loads(l) for l in open(tmp_path / "NeRF" / "val" / "metrics.jsonl")]
    assert any("val_loss" in l for l in lines) and any("grad_2.0_norm_total" in l for l in lines)
    assert net.training                                                                # validate() puts the model back in train mode


def test_non_default_encodings_are_refused_up_front(scene, tmp_path):
    """train_nerf.py -p / -d (train_nerf.py:68-69 upstream) other than 10 / 4: the B200 training kernels are specialised, so the run
    stops at construction with the reason instead of dying in loss.backward()."""
    import train_nerf
    with pytest.raises(RuntimeError, match="differentiable path exists"):
        train_nerf.main(["-n", "odd", "--gpu", "-s", "2", "-rd", str(tmp_path), "-r", "64", "-p", "6", "-d", "2", "full", "-b", str(scene)])


def test_resume_continues_the_step_count_in_the_replayed_graph(scene, tmp_path):
    """train_nerf.py -l CKPT with enough steps left for the CUDA-graph path: the optimiser's step count (host and device copies -
    Adam's bias corrections are evaluated on the device in the replayed step) continues from the checkpoint, the per-epoch LR decay
    of configure_optimizers reaches the captured kernel, and the resumed run keeps training."""
    import train_nerf
    first = train_nerf.main(["-n", "a", "--gpu", "-s", "9", "-rd", str(tmp_path), "-r", "512", "full", "-b", str(scene), "-cr", "0"])
    assert first.global_step == 9 and first.optimizer._step == 9 and int(first.optimizer.dev_step) == 9
    ckpt = first.last_checkpoint
    blob = torch.load(str(ckpt), map_location="cpu", weights_only=False)
    assert blob["global_step"] == 9 and float(blob["optimizer_states"][0]["state"][0]["step"]) == 9.0
    run = train_nerf.main(["-n", "b", "--gpu", "-s", "30", "-rd", str(tmp_path), "-r", "512", "-l", str(ckpt), "full", "-b", str(scene), "-cr", "0"])
    opt = run.optimizer
    assert run.global_step == 30 and opt._step == 30 and int(opt.dev_step) == 30
    gamma = (5e-5 / 5e-4) ** (1 / 1200)
    epochs_done = run.current_epoch                                   # 3 images per epoch: 10 epochs for 30 steps
    assert epochs_done == 10
    assert abs(opt.param_groups[0]["lr"] - 5e-4 * gamma ** epochs_done) < 1e-12
    assert abs(float(opt.dev_state[0]) - 5e-4 * gamma ** (epochs_done - 1)) < 1e-9     # the lr the last replayed step used
    moved = (opt.flat_params.cpu() - torch.cat([v.flatten() for v in blob["state_dict"].values()])[: opt.flat_params.numel()].float()).abs().max()
    assert float(moved) > 1e-4
