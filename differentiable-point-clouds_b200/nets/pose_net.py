"""Pose predictor (reference: dpc/nets/pose_net.py:5-56): an ensemble of small MLPs predicting
candidate quaternions, a student branch, optional translation."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .img_encoder import _fc
from .pc_decoder import _trunc_fc


class PoseBranch(nn.Module):
    def __init__(self, cin, num_layers):
        super().__init__()
        dims = [cin] + [32] * (num_layers - 1) + [4]
        self.layers = nn.ModuleList(_fc(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, t):
        for i, layer in enumerate(self.layers):
            t = layer(t)
            if i + 1 < len(self.layers):
                t = F.leaky_relu(t, 0.2)
        return t


class PoseNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        k = int(cfg.pose_predict_num_candidates)
        if k > 1:
            self.branches = nn.ModuleList(PoseBranch(cfg.z_dim, cfg.pose_candidates_num_layers) for _ in range(k))
            self.student = PoseBranch(cfg.z_dim, cfg.pose_candidates_num_layers) if cfg.pose_predictor_student else None
        else:
            self.single = _fc(cfg.z_dim, 4)
        self.trans = _trunc_fc(cfg.z_dim, 3, cfg.predict_translation_init_stddev) if cfg.predict_translation else None

    def forward(self, x):
        out = {}
        if int(self.cfg.pose_predict_num_candidates) > 1:
            q = torch.cat([b(x) for b in self.branches], dim=1).reshape(-1, 4)   # [B*K,4], candidates adjacent
            if self.student is not None:
                out["pose_student"] = self.student(x)
        else:
            q = self.single(x)
        t = None
        if self.trans is not None:
            t = self.trans(x)
            if self.cfg.predict_translation_tanh:
                t = torch.tanh(t) * self.cfg.predict_translation_scaling_factor
        out["poses"] = q
        out["predicted_translation"] = t
        return out
