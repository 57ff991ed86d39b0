"""Point decoder (reference: dpc/nets/pc_decoder.py:5-42): one FC layer z -> N*3, tanh, /2
(unit cube); optional sigmoid colours."""
import torch
import torch.nn as nn


def _trunc_fc(cin, cout, std):
    fc = nn.Linear(cin, cout)
    nn.init.trunc_normal_(fc.weight, std=std, a=-2 * std, b=2 * std)
    nn.init.zeros_(fc.bias)
    return fc


class PcDecoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        n = int(cfg.pc_num_points)
        self.fc_xyz = _trunc_fc(cfg.z_dim, n * 3, cfg.pc_decoder_init_stddev)
        self.fc_rgb = _trunc_fc(cfg.z_dim, n * 3, cfg.pc_decoder_init_stddev) if cfg.pc_rgb else None

    def forward(self, z):
        n = int(self.cfg.pc_num_points)
        pts = torch.tanh(self.fc_xyz(z).reshape(z.shape[0], n, 3))
        if self.cfg.pc_unit_cube:
            pts = pts / 2.0
        rgb = torch.sigmoid(self.fc_rgb(z).reshape(z.shape[0], n, 3)) if self.fc_rgb is not None else None
        return {"xyz": pts, "rgb": rgb}
