"""f-1: the PyTorch rebuild of the networks / model assembly / losses around the renderer
(reference: dpc/nets/*.py, dpc/models/model_pc.py, dpc/util/losses.py, dpc/util/train.py)."""
import pytest
import torch

from dpc_b200.models.model_pc import (ModelPointCloud, get_dropout_prob, get_learning_rate, get_smooth_sigma,
                                      tf_repeat_0)
from dpc_b200.train import Trainer, synthetic_batch
from dpc_b200.util.config import default_config, experiment_config
from tests.emu_support import emu  # noqa: F401


def test_parameter_counts_match_the_reference_architecture():
    # conv 1,177,216 + FC 6,294,528 + decoder 24,600,000 + scale 1,025
    m = ModelPointCloud(experiment_config("chair_camera_supervision"))
    assert sum(p.numel() for p in m.parameters()) == 1177216 + 6294528 + 24600000 + 1025
    convs = sum(p.numel() for n, p in m.encoder.named_parameters() if n.startswith("convs"))
    assert convs == 1177216
    m2 = ModelPointCloud(experiment_config("chair_unsupervised"))
    assert sum(p.numel() for p in m2.parameters()) == 33292309


def test_schedules_and_helpers():
    cfg = experiment_config("chair_camera_supervision")
    assert get_smooth_sigma(cfg, 0) == 3.0
    assert abs(get_smooth_sigma(cfg, cfg.max_number_of_steps) - 0.2) < 1e-12
    assert abs(get_dropout_prob(cfg, 0) - 0.07) < 1e-12 and get_dropout_prob(cfg, cfg.max_number_of_steps) == 1.0
    assert get_learning_rate(cfg, 0) == 1e-4
    t = torch.arange(3).reshape(3, 1)
    assert tf_repeat_0(t, 2).flatten().tolist() == [0, 0, 1, 1, 2, 2]


def test_encoder_same_padding_shapes():
    cfg = default_config(image_size=128)
    m = ModelPointCloud(cfg)
    out = m.encoder(torch.rand(2, 128, 128, 3))
    assert out["conv_features"].shape == (2, 256 * 4 * 4)
    assert out["ids"].shape == (2, 1024)


def test_pose_candidate_loss_picks_the_best_candidate():
    cfg = default_config(pose_predict_num_candidates=3, predict_pose=True)
    m = ModelPointCloud(cfg)
    gt = torch.rand(2, 4, 4, 1)
    pred = torch.rand(6, 4, 4, 1)
    pred[1] = gt[0]
    pred[5] = gt[1]
    loss, winner = m.proj_loss_pose_candidates(gt, pred)
    assert winner.tolist() == [1, 2]
    assert float(loss) == 0.0


def _tiny(name):
    over = dict(vox_size=16, image_size=32, pc_num_points=64, batch_size=2, step_size=2, z_dim=32, fc_dim=32, f_dim=4,
                pc_gauss_kernel_size=5, pc_relative_sigma=1.0, pc_point_dropout=0.5, max_number_of_steps=100)
    if name == "unsup":
        over.update(predict_pose=True, pose_predict_num_candidates=2)
    return default_config(**over)


@pytest.mark.parametrize("name", ["sup", "unsup"])
def test_train_step_runs_on_the_emulated_kernels(emu, name):  # noqa: F811
    """Whole train step (encoder -> decoder -> pose ensemble -> dropout gather -> renderer -> losses ->
    Adam) on CPU tensors with the kernels under emulation: shapes, gradient flow, loss goes down."""
    torch.manual_seed(0)
    cfg = _tiny(name)
    tr = Trainer(cfg, torch.device("cpu"), ddp=False, bf16=False)
    batch = synthetic_batch(cfg, torch.device("cpu"))
    losses = [float(tr.step(batch)) for _ in range(4)]
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]
    assert all(p.grad is not None for p in tr.model.decoder.parameters())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["chair_camera_supervision", "chair_unsupervised"])
def test_train_step_full_config_gpu(name):
    torch.manual_seed(0)
    cfg = experiment_config(name)
    dev = torch.device("cuda:0")
    tr = Trainer(cfg, dev, ddp=False, bf16=True)
    batch = synthetic_batch(cfg, dev)
    losses = [float(tr.step(batch)) for _ in range(3)]
    assert all(torch.isfinite(torch.tensor(losses)))
    k = cfg.pose_predict_num_candidates
    out = tr.model(batch, 0, True)
    assert out["projs"].shape == (cfg.batch_size * cfg.step_size * k, 64, 64, 1)
