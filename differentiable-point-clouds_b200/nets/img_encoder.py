"""Image encoder (reference: dpc/nets/img_encoder.py:11-50): 5x5/2 conv, then log2(S/4)-1 blocks
of (3x3/2, 3x3/1) convs doubling the width, then three FC layers; leaky-ReLU(0.2) throughout.
TF "SAME" padding on stride-2 convs is asymmetric (extra pixel on the high side); reproduced with
explicit padding so a TF checkpoint would map 1:1.  Flatten order is NHWC like the reference."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _same_pad(k, s, size):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


class SameConv2d(nn.Module):
    def __init__(self, cin, cout, k, s):
        super().__init__()
        self.k, self.s = k, s
        self.conv = nn.Conv2d(cin, cout, k, stride=s)
        # tf.contrib.layers.variance_scaling_initializer(): factor 2, fan_in, truncated normal
        nn.init.kaiming_normal_(self.conv.weight, a=0.0, mode="fan_in", nonlinearity="relu")
        nn.init.zeros_(self.conv.bias)

    def forward(self, x):
        ph = _same_pad(self.k, self.s, x.shape[2])
        pw = _same_pad(self.k, self.s, x.shape[3])
        return self.conv(F.pad(x, (pw[0], pw[1], ph[0], ph[1])))


def _fc(cin, cout):
    fc = nn.Linear(cin, cout)
    nn.init.kaiming_normal_(fc.weight, a=0.0, mode="fan_in", nonlinearity="relu")
    nn.init.zeros_(fc.bias)
    return fc


class ImgEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        size, f = int(cfg.image_size), int(cfg.f_dim)
        layers = [SameConv2d(3, f, 5, 2)]
        for _ in range(int(math.log2(size / 4) - 1)):
            layers += [SameConv2d(f, 2 * f, 3, 2), SameConv2d(2 * f, 2 * f, 3, 1)]
            f *= 2
        self.convs = nn.ModuleList(layers)
        self.fc1 = _fc(f * 4 * 4, cfg.fc_dim)
        self.fc2 = _fc(cfg.fc_dim, cfg.fc_dim)
        self.fc3 = _fc(cfg.fc_dim, cfg.z_dim)
        self.fc_pose = _fc(cfg.fc_dim, cfg.z_dim) if cfg.predict_pose else None

    def forward(self, images):
        """images [B,H,W,3] in [0,1] (the reference's NHWC layout) -> dict(ids, z_latent, conv_features[, poses])."""
        act = lambda t: F.leaky_relu(t, 0.2)  # noqa: E731
        x = (images * 2 - 1).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        for conv in self.convs:
            x = act(conv(x))
        feat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
        fc1 = act(self.fc1(feat))
        fc2 = act(self.fc2(fc1))
        out = {"conv_features": feat, "z_latent": fc1, "ids": act(self.fc3(fc2))}
        if self.fc_pose is not None:
            out["poses"] = F.relu(self.fc_pose(fc2))  # slim.fully_connected's default activation (img_encoder.py:49)
        return out
