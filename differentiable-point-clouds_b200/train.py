"""One training step of the reference's experiments on synthetic data (reference:
dpc/run/train.py:43-117): Adam(lr schedule), loss = task + L2 regularisation, DDP over ranks with
its single bucketed NCCL gradient all-reduce (the renderer itself needs no collective)."""
import torch

from .models.model_pc import ModelPointCloud, get_learning_rate


def synthetic_batch(cfg, device, seed=0):
    """Random stand-ins for one batch of ShapeNet renders: `batch_size` objects x `step_size` views."""
    g = torch.Generator().manual_seed(seed)
    n = cfg.batch_size * cfg.step_size
    s = cfg.image_size
    images = torch.rand(n, s, s, 3, generator=g)
    masks = (torch.rand(n, s, s, 1, generator=g) > 0.5).float()
    quat = torch.randn(n, 4, generator=g)
    batch = {"images": images, "masks": masks, "camera_quaternion": quat,
             "images_1": images[0::cfg.step_size].contiguous()}
    return {k: v.to(device) for k, v in batch.items()}


class Trainer:
    def __init__(self, cfg, device, ddp=False, bf16=True):
        self.cfg, self.device, self.bf16 = cfg, device, bf16
        self.model = ModelPointCloud(cfg).to(device)
        self.net = self.model
        if ddp:
            from torch.nn.parallel import DistributedDataParallel as DDP
            self.net = DDP(self.model, device_ids=[device.index] if device.type == "cuda" else None)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=cfg.learning_rate)
        self.global_step = 0

    def step(self, batch):
        cfg = self.cfg
        for grp in self.opt.param_groups:
            grp["lr"] = get_learning_rate(cfg, self.global_step)
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast(device_type=self.device.type, dtype=torch.bfloat16, enabled=self.bf16):
            outputs = self.net(batch, self.global_step, True)
        loss = self.model.get_loss(batch, outputs) + self.model.regularization_loss()
        loss.backward()
        self.opt.step()
        self.global_step += 1
        return loss.detach()
