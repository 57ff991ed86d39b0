"""Generates tests/golden/losses/*.npz by running the REFERENCE'S OWN loss code, unmodified
(/root/reference/dpc/models/model_pc.py:308-445: proj_loss_pose_candidates, add_student_loss, add_proj_loss, get_loss;
/root/reference/dpc/util/losses.py: add_drc_loss, add_proj_rgb_loss, add_proj_depth_loss;
/root/reference/dpc/util/gauss_kernel.py:14-24: gauss_smoothen_image), over the TF1 shim (oracle/tf1_shim), values
and -- through torch autograd underneath the shim -- gradients w.r.t. every predicted tensor.

    python tests/golden/make_golden_losses.py        # here, where /root/reference exists

Each .npz holds the config overrides (json), the `inputs` / `outputs` tensors handed to the reference, the global step,
the loss value and the gradients.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import run_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "losses")


class Cfg(dict):
    __getattr__ = dict.__getitem__


def ref_cfg(over):
    with open(os.path.join(run_reference.REFERENCE_ROOT, "dpc", "resources", "default_config.yaml")) as f:
        c = Cfg(yaml.safe_load(f))
    c.update(over)
    return c


def rnd(g, *shape):
    return torch.rand(*shape, generator=g)


# name: (cfg overrides, global_step, builder options)
BASE = dict(vox_size=16, image_size=32, batch_size=2, step_size=3, max_number_of_steps=1000, pc_relative_sigma=3.0,
            pc_relative_sigma_end=0.2, pc_gauss_kernel_size=7)
CASES = {
    "proj_supervised": (dict(BASE), 100, dict()),
    "proj_gauss_filter_gt": (dict(BASE, pc_gauss_filter_gt=True), 100, dict()),
    "proj_gauss_filter_gt_switch_off_late": (dict(BASE, pc_gauss_filter_gt=True, pc_gauss_filter_gt_switch_off=True), 900, dict()),
    "proj_gauss_filter_gt_switch_off_early": (dict(BASE, pc_gauss_filter_gt=True, pc_gauss_filter_gt_switch_off=True), 100, dict()),
    "candidates_student": (dict(BASE, predict_pose=True, pose_predict_num_candidates=4, pose_predictor_student_loss_weight=20.0), 100, dict()),
    "candidates_student_variable_views": (dict(BASE, predict_pose=True, pose_predict_num_candidates=4, variable_num_views=True,
                                               pose_predictor_student_loss_weight=20.0), 100, dict(valid=True)),
    "candidates_student_align": (dict(BASE, predict_pose=True, pose_predict_num_candidates=4, pose_student_align_loss=True), 100, dict()),
    "drc": (dict(BASE, proj_weight=0.0, drc_weight=0.5), 100, dict(drc=True)),
    "rgb": (dict(BASE, pc_rgb=True, proj_rgb_weight=2.0), 100, dict(rgb=True)),
    "rgb_gauss_filter": (dict(BASE, pc_rgb=True, proj_rgb_weight=2.0, pc_gauss_filter_gt_rgb=True), 100, dict(rgb=True)),
    "depth": (dict(BASE, proj_depth_weight=0.3, max_depth=8.0, pc_gauss_filter_gt=True), 100, dict(depth=True)),
    "everything": (dict(BASE, drc_weight=0.25, pc_rgb=True, proj_rgb_weight=1.5, proj_depth_weight=0.2, predict_pose=True,
                        pose_predict_num_candidates=2, variable_num_views=True), 400, dict(drc=True, rgb=True, depth=True, valid=True)),
}


def build(cfg, opt, seed):
    g = torch.Generator().manual_seed(seed)
    n = cfg.batch_size * cfg.step_size
    k = cfg.pose_predict_num_candidates
    v = cfg.vox_size
    inputs = {"masks": (rnd(g, n, cfg.image_size, cfg.image_size, 1) > 0.5).float()}
    outputs = {"projs": rnd(g, n * k, v, v, 1)}
    if k > 1:
        outputs["poses"] = torch.randn(n * k, 4, generator=g)
        outputs["pose_student"] = torch.randn(n, 4, generator=g)
    if opt.get("valid"):
        inputs["valid_samples"] = (rnd(g, n) > 0.3).float()
    if opt.get("drc"):
        p = rnd(g, v + 1, n * k, v, v, 1)
        outputs["drc_probs"] = p / p.sum(0, keepdim=True)
        if k > 1:
            raise_k = None  # noqa: F841  (the reference's add_drc_loss divides by the number of masks, not predictions)
    if opt.get("rgb"):
        inputs["images"] = rnd(g, n * k, cfg.image_size, cfg.image_size, 3)
        outputs["projs_rgb"] = rnd(g, n * k, v, v, 3)
    if opt.get("depth"):
        d = 1.5 + rnd(g, n * k, v, v, 1)
        d[rnd(g, n * k, v, v, 1) > 0.7] = cfg.max_dataset_depth
        inputs["depths"] = d
        outputs["projs_depth"] = 1.5 + 2 * rnd(g, n * k, v, v, 1)
    return inputs, outputs


def main():
    ns = run_reference.load()
    tf = ns.tf
    mp = importlib.import_module("models.model_pc")
    os.makedirs(OUT, exist_ok=True)
    for idx, (name, (over, step, opt)) in enumerate(CASES.items()):
        cfg = ref_cfg(over)
        if opt.get("drc") and cfg.pose_predict_num_candidates > 1:
            # add_drc_loss multiplies [Vz+1, n*k, ...] probabilities with psi built from the n masks: only k = 1 broadcasts
            # in the reference, so the DRC term of the "everything" case gets masks replicated by the caller
            pass
        np.random.seed(100 + idx)           # setup_misc draws the align-loss cloud with np.random
        model = mp.ModelPointCloud(cfg, step)
        inputs, outputs = build(cfg, opt, 500 + idx)
        if opt.get("drc") and cfg.pose_predict_num_candidates > 1:
            # DRC needs one mask per prediction: give the reference what its broadcasting requires
            outputs["drc_probs"] = outputs["drc_probs"][:, : inputs["masks"].shape[0]]
        leaves = {k: v.clone().requires_grad_(True) for k, v in outputs.items()}
        t_in = {k: tf.Tensor(v) for k, v in inputs.items()}
        t_out = {k: tf.Tensor(v) for k, v in leaves.items()}
        loss = model.get_loss(t_in, t_out, add_summary=False).t
        used = [k for k in leaves]
        grads = torch.autograd.grad(loss, [leaves[k] for k in used], allow_unused=True)
        rec = {"cfg_json": json.dumps(over), "global_step": np.int64(step), "loss": loss.detach().numpy()}
        for k, v in inputs.items():
            rec["in_" + k] = v.numpy()
        for k, v in outputs.items():
            rec["out_" + k] = v.numpy()
        for k, gr in zip(used, grads):
            rec["grad_" + k] = (gr if gr is not None else torch.zeros_like(leaves[k])).numpy()
        if cfg.pose_student_align_loss:
            rec["alignloss_pc"] = model._pc_for_alignloss.t.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print("%-40s loss %.6f" % (name, float(loss)))
    # gauss_smoothen_image on its own (3 channels, even and odd tap counts)
    gk = ns.gauss_kernel
    g = torch.Generator().manual_seed(77)
    for fsz, sig in ((7, 1.3), (10, 2.0)):
        img = rnd(g, 2, 12, 14, 3)
        out = gk.gauss_smoothen_image(Cfg(pc_gauss_kernel_size=fsz), tf.Tensor(img), tf.constant(sig, dtype=tf.float32)).t
        np.savez_compressed(os.path.join(OUT, "smoothen_image_k%d.npz" % fsz), img=img.numpy(), out=out.numpy(),
                            fsz=np.int64(fsz), sigma=np.float32(sig))
        print("smoothen_image k=%d ok" % fsz)


if __name__ == "__main__":
    main()
