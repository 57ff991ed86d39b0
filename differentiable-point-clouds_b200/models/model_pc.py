"""Model assembly and losses (reference: dpc/models/model_pc.py:130-445, dpc/util/losses.py:6-20,
dpc/util/train.py:17-23) rebuilt in PyTorch: encoder -> point decoder (+ occupancy scale, pose
ensemble) -> replicate per view / per pose candidate -> the B200 projection path -> losses.

Everything here is library kernels (cuDNN/cuBLAS through torch) EXCEPT `compute_projection`, which
calls the hand-written renderer through the reference-shaped API.  The renderer stays fp32; the
CNN runs under bf16 autocast when the caller enables it (BASELINE config 3).
"""
import torch
import torch.nn as nn

from ..nets.img_encoder import ImgEncoder, _fc
from ..nets.pc_decoder import PcDecoder, _trunc_fc
from ..nets.pose_net import PoseNet
from ..util import gauss_kernel, point_cloud
from ..util import losses as L
from ..util.quaternion import quaternion_conjugate, quaternion_multiply, quaternion_normalise, quaternion_rotate


def tf_repeat_0(t, num):
    """[a,b,..] -> [a,a,..,b,b,..] along axis 0 (model_pc.py:23-32)."""
    return t.repeat_interleave(num, dim=0)


def get_smooth_sigma(cfg, global_step):
    """Linear schedule pc_relative_sigma -> pc_relative_sigma_end over training (model_pc.py:35-40)."""
    diff = cfg.pc_relative_sigma_end - cfg.pc_relative_sigma
    return float(cfg.pc_relative_sigma + global_step / cfg.max_number_of_steps * diff)


def get_dropout_prob(cfg, global_step):
    """Keep-probability schedule of the point dropout (model_pc.py:43-64), linear variant."""
    if not cfg.pc_point_dropout_scheduled:
        return float(cfg.pc_point_dropout)
    start, end = cfg.pc_point_dropout, 1.0
    k = (end - start) / (cfg.pc_point_dropout_end_step - cfg.pc_point_dropout_start_step)
    b = start - k * cfg.pc_point_dropout_start_step
    x = global_step / cfg.max_number_of_steps
    if cfg.pc_point_dropout_exponential_schedule:
        import math
        keep = start * math.exp(math.log(end / start) * x)
    else:
        keep = k * x + b
    return float(min(max(keep, start), end))


def get_learning_rate(cfg, global_step):
    return cfg.learning_rate if global_step < cfg.learning_rate_step * cfg.max_number_of_steps else cfg.learning_rate_2


def pool_single_view(cfg, tensor, view_idx):
    """Every step_size-th row starting at view_idx (model_base.py:7-18)."""
    return tensor[view_idx::cfg.step_size]


class ModelPointCloud(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.encoder = ImgEncoder(cfg)
        self.decoder = PcDecoder(cfg)
        self.scale_fc = _trunc_fc(cfg.z_dim, 1, 0.025) if cfg.pc_learn_occupancy_scaling else None
        self.focal_fc = _trunc_fc(cfg.z_dim, 1, 0.025) if cfg.learn_focal_length else None      # model_pc.py:111-127
        self.posenet = PoseNet(cfg) if cfg.predict_pose else None
        if cfg.pc_rgb_deep_decoder:
            raise NotImplementedError("pc_rgb_deep_decoder is not implemented (pc_decoder.py:24-31)")
        if cfg.bicubic_gt_downsampling:
            raise NotImplementedError("bicubic_gt_downsampling is not implemented (model_pc.py:394-395)")
        if cfg.pose_student_align_loss:
            # 2000 reference points ~ N(0, 1) clipped to +-3 sigma (model_pc.py:162-169).  A tf.Variable outside the
            # 'encoder' / 'decoder' scopes the optimizer trains (run/train.py:78,89), i.e. constant: a buffer here
            import numpy as np
            vals = np.clip(np.random.normal(loc=0.0, scale=1.0, size=(2000, 3)), -3.0, 3.0)
            self.register_buffer("pc_for_alignloss", torch.tensor(vals, dtype=torch.float32))

    # ------------------------------------------------------------------ prediction
    def model_predict(self, images):
        cfg = self.cfg
        enc = self.encoder(images)
        out = {"conv_features": enc["conv_features"], "ids": enc["ids"], "z_latent": enc["z_latent"]}
        ids = enc["ids"]
        if ids.shape[0] != cfg.batch_size:      # all views were encoded: keep the first view's identity
            ids = pool_single_view(cfg, ids, 0)
        out["ids_1"] = ids
        dec = self.decoder(ids)
        out["points_1"], out["rgb_1"] = dec["xyz"], dec["rgb"]
        out["scaling_factor"] = (torch.sigmoid(self.scale_fc(ids)) * cfg.pc_occupancy_scaling_maximum
                                 if self.scale_fc is not None else None)
        # model_pc.py:205: the focal length is predicted from the identity units of ALL encoded views
        out["focal_length"] = (cfg.focal_length_mean + torch.sigmoid(self.focal_fc(enc["ids"])) * cfg.focal_length_range
                               if self.focal_fc is not None else None)
        if self.posenet is not None:
            out.update(self.posenet(enc["poses"]))
        return out

    def compute_projection(self, inputs, outputs, global_step, is_training=True):
        """model_pc.py:220-259."""
        cfg = self.cfg
        all_points, all_rgb = outputs["all_points"], outputs["all_rgb"]
        camera_pose = outputs["poses"] if cfg.predict_pose else (
            inputs["camera_quaternion"] if cfg.pose_quaternion else inputs["matrices"])
        sel = None
        if is_training and cfg.pc_point_dropout != 1:
            keep = get_dropout_prob(cfg, global_step)
            if all_rgb is None and all_points.is_cuda:
                # the subset is consumed by the splat's load stage: dropped points are never read or copied
                n_keep = point_cloud.num_points_after_dropout(all_points.shape[1], keep)
                state = getattr(self, "dropout_state", None)      # device {seed, draw}: set by a Trainer (graph replays)
                sel = point_cloud.dropout_indices(all_points.shape[0], all_points.shape[1], n_keep, all_points.device,
                                                  state=state)
                if state is not None:
                    state[1] += 1
            else:
                all_points, all_rgb = point_cloud.pc_point_dropout(all_points, all_rgb, keep)
        # sigma is a function of the step count (model_pc.py:35-40): a host value, so the taps are built
        # on the CPU and the renderer receives them as launch parameters as well as a device buffer
        sigma_dev = getattr(self, "sigma_override", None)   # a device tensor holding the same value (captured steps)
        kernel = gauss_kernel.smoothing_kernel(cfg, sigma_dev if sigma_dev is not None
                                               else float(get_smooth_sigma(cfg, global_step)))
        trans = outputs.get("predicted_translation") if cfg.predict_translation else None
        with torch.autocast(device_type=all_points.device.type, enabled=False):   # the renderer is fp32
            proj_out = point_cloud.pointcloud_project_fast(
                cfg, all_points.float(), camera_pose.float(), trans, all_rgb, kernel,
                scaling_factor=None if outputs["all_scaling_factors"] is None else outputs["all_scaling_factors"].float(),
                focal_length=outputs["all_focal_length"], point_indices=sel)
        outputs["projs"] = proj_out["proj"]
        outputs["projs_rgb"] = proj_out["proj_rgb"]
        outputs["proj_out"] = proj_out      # drc_probs / proj_depth stay lazy until a loss asks for them (get_loss)
        outputs["sigma_rel"] = float(get_smooth_sigma(cfg, global_step))
        outputs["projs_1"] = proj_out["proj"][0:outputs["points_1"].shape[0]]
        return outputs

    def forward(self, inputs, global_step=0, is_training=True, run_projection=True):
        """model_pc.py:266-306 (get_model_fn)."""
        cfg = self.cfg
        outputs = self.model_predict(inputs["images"] if cfg.predict_pose else inputs["images_1"])
        hook = getattr(self, "predict_hook", None)      # test seam: lets a test pin the predictions the renderer sees
        if hook is not None:
            outputs = hook(outputs)
        if not run_projection:
            return outputs
        k = int(cfg.pose_predict_num_candidates)
        all_points = tf_repeat_0(outputs["points_1"], cfg.step_size)
        if k > 1:
            all_points = tf_repeat_0(all_points, k)
            if cfg.predict_translation:
                outputs["predicted_translation"] = tf_repeat_0(outputs["predicted_translation"], k)
        # model_pc.py:283-296: the predicted focal length only reaches the renderer with several pose candidates
        outputs["all_focal_length"] = (tf_repeat_0(outputs["focal_length"], k)
                                       if (k > 1 and outputs["focal_length"] is not None) else None)
        outputs["all_points"] = all_points
        sc = outputs["scaling_factor"]
        if sc is not None:
            sc = tf_repeat_0(sc, cfg.step_size)
            if k > 1:
                sc = tf_repeat_0(sc, k)
        outputs["all_scaling_factors"] = sc
        outputs["all_rgb"] = tf_repeat_0(outputs["rgb_1"], cfg.step_size) if cfg.pc_rgb else None
        return self.compute_projection(inputs, outputs, global_step, is_training)

    # ------------------------------------------------------------------ losses
    def proj_loss_pose_candidates(self, gt, pred, inputs=None):
        """min over pose candidates (model_pc.py:308-337) -> (loss, winning candidate per view)."""
        cfg = self.cfg
        k = int(cfg.pose_predict_num_candidates)
        gt = tf_repeat_0(gt, k)
        all_loss = ((gt - pred) ** 2).sum(dim=(1, 2, 3)).reshape(-1, k)
        winner = all_loss.argmin(dim=1)
        mask = torch.nn.functional.one_hot(winner, k).to(pred.dtype).reshape(-1, 1, 1, 1)
        loss_tensor = (gt - pred) * mask
        if cfg.variable_num_views:      # padded views carry weight 0 (model_pc.py:329-333)
            w = tf_repeat_0(inputs["valid_samples"].to(pred.dtype), k)
            loss_tensor = loss_tensor * w.reshape(-1, 1, 1, 1)
        num_samples = winner.shape[0]
        return (loss_tensor ** 2).sum() / 2 / num_samples, winner

    def add_student_loss(self, inputs, outputs, winner, add_summary=False):
        """model_pc.py:339-381: the student regresses the winning candidate (teacher, no gradient): quaternion-angle
        loss, or -- pose_student_align_loss -- the distance between a fixed cloud rotated by both.  The quaternion
        algebra runs in fp32 whatever the autocast state of the networks (1 - w^2 near w = 1 needs the mantissa)."""
        cfg = self.cfg
        k = int(cfg.pose_predict_num_candidates)
        student = outputs["pose_student"].float()
        teachers = outputs["poses"].float().reshape(-1, k, 4)
        teachers = teachers[torch.arange(teachers.shape[0], device=teachers.device), winner].detach()
        weights = inputs["valid_samples"].float() if cfg.variable_num_views else 1.0
        if cfg.pose_student_align_loss:
            ref = self.pc_for_alignloss.float().unsqueeze(0).expand(teachers.shape[0], -1, -1)
            d = quaternion_rotate(ref, teachers) - quaternion_rotate(ref, student)
            loss = (d ** 2).sum() / 2 / self.pc_for_alignloss.shape[0]
        else:
            q_diff = quaternion_normalise(quaternion_multiply(teachers, quaternion_conjugate(student)))
            loss = ((1.0 - q_diff[:, 0] ** 2) * weights).sum()
        loss = loss / winner.shape[0]
        return loss * cfg.pose_predictor_student_loss_weight

    def add_proj_loss(self, inputs, outputs, weight_scale=None, add_summary=False):
        """model_pc.py:383-423.  TF1's bilinear resize without half-pixel centres is exact [::2,::2]
        subsampling for 128 -> 64."""
        cfg = self.cfg
        weight_scale = cfg.proj_weight if weight_scale is None else weight_scale
        gt, pred = inputs["masks"], outputs["projs"].float()
        assert gt.shape[1] >= pred.shape[1], "GT size should not be higher than prediction size"
        gt = L.resize_tf1(gt, pred.shape[1], "bicubic" if cfg.bicubic_gt_downsampling else "bilinear")
        gt = L._filtered_gt(cfg, gt, outputs.get("sigma_rel"), cfg.pc_gauss_filter_gt)
        total = pred.new_zeros(())
        if int(cfg.pose_predict_num_candidates) > 1:
            proj_loss, winner = self.proj_loss_pose_candidates(gt, pred, inputs)
            if cfg.pose_predictor_student:
                total = total + self.add_student_loss(inputs, outputs, winner, add_summary)
        else:
            if pred.is_cuda:      # model_pc.py:414-415 as one kernel (value + gradient), util/losses.py
                proj_loss = L.proj_l2_loss(gt.contiguous(), pred, pred.shape[0])
            else:                 # CPU tensors only occur in the CPU test-suite
                proj_loss = ((gt - pred) ** 2).sum() / 2 / pred.shape[0]
        return (total + proj_loss) * weight_scale

    def get_loss(self, inputs, outputs, add_summary=False):
        """model_pc.py:425-445: projection loss, + DRC loss (drc_weight), + rgb projection loss (pc_rgb, weighted by
        proj_rgb_weight), + depth projection loss (proj_depth_weight).  drc_probs / projs_depth are taken from the
        projection's lazy result, so they cost nothing unless their weight is non-zero."""
        cfg = self.cfg
        loss = outputs["projs"].new_zeros((), dtype=torch.float32)
        sigma = outputs.get("sigma_rel")
        if cfg.proj_weight:
            loss = loss + self.add_proj_loss(inputs, outputs, cfg.proj_weight, add_summary)
        if cfg.drc_weight:
            if outputs.get("drc_probs") is None:
                outputs["drc_probs"] = outputs["proj_out"]["drc_probs"]
            loss = loss + L.add_drc_loss(cfg, inputs, outputs, cfg.drc_weight, add_summary)
        if cfg.pc_rgb:
            loss = loss + L.add_proj_rgb_loss(cfg, inputs, outputs, cfg.proj_rgb_weight, add_summary, sigma)
        if cfg.proj_depth_weight:
            if outputs.get("projs_depth") is None:
                outputs["projs_depth"] = outputs["proj_out"]["proj_depth"]
            loss = loss + L.add_proj_depth_loss(cfg, inputs, outputs, cfg.proj_depth_weight, sigma, add_summary)
        return loss

    def regularization_loss(self):
        """weight_decay * sum l2_loss(W) over encoder/decoder weights (util/losses.py:6-20)."""
        if self.cfg.weight_decay <= 0:
            return 0.0
        reg = 0.0
        for mod in (self.encoder, self.decoder, self.scale_fc, self.focal_fc, self.posenet):
            if mod is None:
                continue
            for name, p in mod.named_parameters():
                if name.endswith("weight"):
                    reg = reg + (p.float() ** 2).sum() / 2
        return reg * self.cfg.weight_decay
