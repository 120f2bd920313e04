"""The reference's own unit tests for the in-scope modules (tests/nerf_helpers_test.py, tests/nerf_model_test.py,
tests/dataloader_test.py), same names and assertions, with the inputs moved to the CUDA device (this path has no CPU
implementation) - SURVEY.md section 8b."""
import pytest
import torch
import torch.testing as testing

pytestmark = pytest.mark.gpu
DEV = "cuda"


class TestNeRFHelpers:
    def test_calculate_unnormalized_weights(self):
        import nerf_helpers
        deltas = torch.full((1, 5, 1), 0.2, device=DEV)
        density = torch.Tensor([0, 50, 1, 0.3, 1]).view(deltas.shape).to(DEV)
        weights = nerf_helpers.calculate_unnormalized_weights(density, deltas)
        gt_weights = torch.Tensor([0, 0.9999546001, 8.229611e-6, 2.1646e-6, 6.34545e-6]).view(deltas.shape)
        testing.assert_close(weights.cpu(), gt_weights)

    def test_estimate_ray_color(self):
        import nerf_helpers
        weights = torch.full((1, 256, 1), 1 / 256, device=DEV)
        rgbs = torch.full((1, 256, 3), 1.0, device=DEV)
        testing.assert_close(nerf_helpers.estimate_ray_color(weights, rgbs).cpu(), torch.ones((1, 3)))

    def test_estimate_ray_color_one_weight(self):
        import nerf_helpers
        weights = torch.zeros((1, 256, 1), device=DEV)
        weights[:, 200, :] = 1.0
        rgbs = torch.full((1, 256, 3), 1.0, device=DEV)
        testing.assert_close(nerf_helpers.estimate_ray_color(weights, rgbs).cpu(), torch.ones((1, 3)))

    def test_generate_deltas(self):
        import nerf_helpers
        ts = torch.arange(2, 6, 1, device=DEV).view((1, -1, 1))
        gt_deltas = torch.ones((1, 4, 1))
        gt_deltas[:, -1, :] = 1e10
        testing.assert_close(nerf_helpers.generate_deltas(ts).cpu(), gt_deltas)

    def test_generate_random_samples(self):
        import nerf_helpers
        o_rays = torch.Tensor([[0.0, 0.0, 0.0]]).to(DEV)
        d_rays = torch.Tensor([[1.0, 1.0, 1.0]]).to(DEV)
        samples, ts = nerf_helpers.generate_coarse_samples(o_rays, d_rays, 2)
        samples, ts = samples.cpu(), ts.cpu()
        ts_bounds = torch.Tensor([[2.0, 4.0, 6.0]]).T
        assert torch.logical_and(ts_bounds[None, :-1, :] <= ts, ts < ts_bounds[None, 1:, :]).all()
        sample_bounds = torch.Tensor([[2.0, 2.0, 2.0], [4.0, 4.0, 4.0], [6.0, 6.0, 6.0]])
        assert torch.logical_and(sample_bounds[:-1, :] <= samples, samples < sample_bounds[1:, :]).all()


class TestNerfModel:
    def test_nerf_network_training_step(self):
        import nerf_model
        network = nerf_model.NeRFNetwork(position_dim=10, direction_dim=4, coarse_samples=64, fine_samples=128).to(DEV)
        batch = {"origin": torch.full((1, 1, 3), 0.5, device=DEV), "direc": torch.tensor([[[0.1, -0.2, -1.0]]], device=DEV),
                 "rgb": torch.rand(1, 1, 3, device=DEV)}
        loss = network.training_step(batch, 0)
        assert loss >= 0 and loss.requires_grad
        loss.backward()
        assert all(p.grad is not None for p in network.parameters())

    def test_positional_encoding_shape(self):
        import nerf_model
        assert nerf_model.positional_encoding(torch.Tensor([[1.0, 1.0, 1.0]]).to(DEV), dim=1).shape == (1, 6)

    def test_positional_encoding_values(self):
        import nerf_model
        enc = nerf_model.positional_encoding(torch.Tensor([[1.0, 1.0, 1.0]]).to(DEV), dim=1)
        testing.assert_close(enc.cpu(), torch.Tensor([[-1.0, -1.0, -1.0, 0.0, 0.0, 0.0]]), atol=1e-6, rtol=0)

    def test_complex_positional_encoding_values(self):
        import nerf_model
        enc = nerf_model.positional_encoding(torch.Tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]]).to(DEV), dim=1)
        assert enc.shape == (2, 6)
        expected = torch.Tensor([[1.0, 1.0, 1.0, 0.0, 0.0, 0.0], [-1.0, -1.0, -1.0, 0.0, 0.0, 0.0]])
        testing.assert_close(enc.cpu(), expected, atol=1e-6, rtol=0)

    def test_3D_positional_encoding_shape(self):
        import nerf_model
        enc = nerf_model.positional_encoding(torch.rand((4096, 64, 3), device=DEV), dim=10)
        assert enc.shape == (4096, 64, 60)

    def test_single_complex_forward_prop_shape(self):
        import nerf_model
        model = nerf_model.NeRFModel(position_dim=10, direction_dim=4).to(DEV)
        density, rgb = model(torch.rand((4, 4, 3), device=DEV), torch.rand((4, 3), device=DEV))
        assert density.shape == (4, 4, 1) and rgb.shape == (4, 4, 3)


class TestDataloader:
    def test_synthetic_focal_length_and_batch(self, tmp_path):
        import dataloader
        import synthetic
        synthetic.write_blender_scene(tmp_path, n_train=1, n_val=1, n_test=1, camera_angle_x=0.6)
        sds = dataloader.SyntheticDataset(tmp_path, "train", 1)
        assert abs(sds.focal - 1293.091257506331) < 5e-8                      # dataloader_test.py:39-41
        batch = next(iter(dataloader.getSyntheticDataloader(tmp_path, "train", 4096, num_workers=1, shuffle=True)))
        assert "origin" in batch and "direc" in batch and "rgb" in batch        # dataloader_test.py:43-47
        assert batch["origin"].shape == (1, 4096, 3)
