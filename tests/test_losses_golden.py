"""The loss surface of the train step (models/model_pc.py: proj_loss_pose_candidates, add_student_loss, add_proj_loss,
get_loss; util/losses.py: add_drc_loss, add_proj_rgb_loss, add_proj_depth_loss; util/gauss_kernel.gauss_smoothen_image)
against fixtures produced by the REFERENCE'S OWN loss code run over the TF1 shim (tests/golden/make_golden_losses.py):
loss values and the gradients w.r.t. every predicted tensor.  Pure torch, runs on the CPU."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from dpc_b200.models.model_pc import ModelPointCloud, get_smooth_sigma
from dpc_b200.util import gauss_kernel as gk
from dpc_b200.util.config import default_config

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(DIR, "*.npz"))
               if not os.path.basename(p).startswith("smoothen_image"))
TINY = dict(z_dim=8, fc_dim=8, f_dim=2, pc_num_points=10)     # the networks are not exercised here


def test_fixtures_present():
    assert len(NAMES) >= 12


@pytest.mark.parametrize("name", NAMES)
def test_loss_matches_reference(name):
    z = np.load(os.path.join(DIR, name + ".npz"))
    over = json.loads(str(z["cfg_json"]))
    cfg = default_config(**over, **TINY)
    model = ModelPointCloud(cfg)
    if "alignloss_pc" in z.files:
        with torch.no_grad():
            model.pc_for_alignloss.copy_(torch.from_numpy(z["alignloss_pc"]))
    inputs = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    outputs = {k[4:]: torch.from_numpy(z[k]).clone().requires_grad_(True) for k in z.files if k.startswith("out_")}
    leaves = dict(outputs)
    outputs["sigma_rel"] = get_smooth_sigma(cfg, int(z["global_step"]))
    loss = model.get_loss(inputs, outputs)
    ref = float(z["loss"])
    assert abs(float(loss) - ref) <= 2e-6 * max(1.0, abs(ref)), (float(loss), ref)
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    for n, g in zip(names, grads):
        want = z["grad_" + n]
        got = np.zeros_like(want) if g is None else g.numpy()
        scale = max(1.0, float(np.abs(want).max()))
        assert float(np.abs(got - want).max()) <= 2e-6 * scale, (n, float(np.abs(got - want).max()))


@pytest.mark.parametrize("fsz", [7, 10])
def test_gauss_smoothen_image(fsz):
    z = np.load(os.path.join(DIR, "smoothen_image_k%d.npz" % fsz))
    cfg = default_config(pc_gauss_kernel_size=int(z["fsz"]))
    out = gk.gauss_smoothen_image(cfg, torch.from_numpy(z["img"]), torch.tensor(float(z["sigma"])))
    assert out.shape == tuple(z["out"].shape)
    assert float((out - torch.from_numpy(z["out"])).abs().max()) <= 1e-6


def test_unimplemented_keys_fail_loudly(tmp_path):
    from dpc_b200.util.config import load_config
    p = tmp_path / "c.yaml"
    p.write_text("pc_normalise_gauss: true\n")
    with pytest.raises(NotImplementedError):
        load_config(str(p))
    p.write_text("not_a_key: 1\n")
    with pytest.raises(KeyError):
        load_config(str(p))
    p.write_text("vis_size: 64\nvox_size: 32\nproj_rgb_weight: 0.5\n")
    cfg = load_config(str(p))
    assert cfg.vox_size == 32 and cfg.proj_rgb_weight == 0.5 and "vis_size" not in cfg
    with pytest.raises(NotImplementedError):
        ModelPointCloud(default_config(bicubic_gt_downsampling=True, **TINY))
