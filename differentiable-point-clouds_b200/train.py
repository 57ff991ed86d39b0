"""One training step of the reference's experiments on synthetic data (reference: dpc/run/train.py:43-117):
loss = task loss + L2 regularisation of the weights (util/losses.py:6-20), Adam with the step-dependent learning rate
(util/train.py:17-23), data parallel over ranks.

B200 shape of it:
  * all parameters are views of ONE flat fp32 buffer and all gradients views of another, so the data-parallel exchange is
    a single NCCL all-reduce of that buffer per step ("one sum-allreduce of the CNN/decoder gradients"),
    the regulariser's gradient is one fused multiply-add over it (weight_decay * W for weights, 0 for biases) and Adam is
    one fused kernel;
  * the whole step -- networks under bf16 autocast, fp32 renderer (hand-written kernels), losses, backward, all-reduce,
    Adam -- is captured in a CUDA graph (one per number of points kept by the dropout schedule) and replayed: the eager
    step is launch-bound (~5 ms of host time for ~1.5 ms of GPU work at 8 objects).  Values that change every step
    (learning rate, sigma of the smoothing taps, the dropout draw counter) live in device tensors the graph reads.
The renderer itself needs no collective (samples are independent).
"""
import torch
import torch.distributed as dist

from .models.model_pc import ModelPointCloud, get_dropout_prob, get_learning_rate, get_smooth_sigma
from .util import point_cloud


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
    def __init__(self, cfg, device, ddp=False, bf16=True, graph=False):
        self.cfg, self.device, self.bf16 = cfg, device, bf16
        self.model = ModelPointCloud(cfg).to(device)
        self.net = self.model
        self.world = dist.get_world_size() if (ddp and dist.is_available() and dist.is_initialized()) else 1
        self.ddp = ddp and self.world > 1
        self._flatten()
        cuda = device.type == "cuda"
        # the learning rate is a device tensor on CUDA (a captured step reads it); a plain float on the CPU (tests)
        self.lr = torch.tensor(float(cfg.learning_rate), dtype=torch.float32, device=device) if cuda else None
        self.opt = (torch.optim.Adam([self.flat_p], lr=self.lr, capturable=True, fused=True) if cuda
                    else torch.optim.Adam([self.flat_p], lr=float(cfg.learning_rate)))
        self.global_step = 0
        self.graph = bool(graph) and cuda
        self._graphs = {}
        self._static = None
        # device-side values a captured step reads: sigma of the smoothing taps, {seed, draw} of the dropout subsets
        self.sigma = torch.tensor(float(get_smooth_sigma(cfg, 0)), dtype=torch.float32, device=device)
        self.model.dropout_state = (torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
                                    if cuda else None)
        self.comm_bytes = self.flat_g.numel() * 4 if self.ddp else 0

    # ------------------------------------------------------------------ flat parameter / gradient buffers
    def _flatten(self):
        params = [p for p in self.model.parameters()]
        total = sum(p.numel() for p in params)
        dev = self.device
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        mask = torch.zeros(total, dtype=torch.float32, device=dev)      # 1 where the regulariser applies (weights)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        names = {id(p): n for n, p in self.model.named_parameters()}
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view_as(p.data)
            p.grad = self.flat_g[off:off + n].view_as(p.data)
            if names[id(p)].endswith("weight"):       # 'kernel' / 'weights' variables of the reference (losses.py:9-10)
                mask[off:off + n] = 1.0
            off += n
        self.flat_p = torch.nn.Parameter(flat, requires_grad=True)
        # the Parameter wraps the same storage; keep the views pointing at it
        off = 0
        for p in params:
            n = p.numel()
            p.data = self.flat_p.data[off:off + n].view_as(p.data)
            off += n
        self.flat_p.grad = self.flat_g
        self.wd_mask = mask
        # flat order = registration order: the encoder's parameters come first
        self.n_encoder = sum(p.numel() for n, p in self.model.named_parameters() if n.startswith("encoder."))
        assert [n.startswith("encoder.") for n, _ in self.model.named_parameters()] == \
            sorted([n.startswith("encoder.") for n, _ in self.model.named_parameters()], reverse=True)

    def _reduce_slice(self, lo, hi):
        """regulariser + all-reduce + average of flat_g[lo:hi] on the current stream."""
        g = self.flat_g[lo:hi]
        if self.cfg.weight_decay > 0:
            g.addcmul_(self.flat_p.detach()[lo:hi], self.wd_mask[lo:hi], value=float(self.cfg.weight_decay))
        if self.ddp:
            dist.all_reduce(g)
            g.mul_(1.0 / self.world)

    def load_flat(self, other):
        """Copy another trainer's parameters (same architecture)."""
        with torch.no_grad():
            self.flat_p.copy_(other.flat_p)

    # ------------------------------------------------------------------ one step
    def _forward_backward(self, batch):
        """zero grads -> forward -> loss -> backward -> regulariser -> all-reduce.  Returns (task + reg loss)."""
        cfg = self.cfg
        self.flat_g.zero_()
        with torch.autocast(device_type=self.device.type, dtype=torch.bfloat16, enabled=self.bf16):
            outputs = self.net(batch, self.global_step, True)
        loss = self.model.get_loss(batch, outputs)
        loss.backward()
        if cfg.weight_decay > 0:
            reg = (self.flat_p.detach() * self.flat_p.detach() * self.wd_mask).sum() * (0.5 * float(cfg.weight_decay))
            loss = loss.detach() + reg
        # d/dW of weight_decay * sum(W^2) / 2 (weights only) and THE collective of the step: a sum over ranks of the flat
        # gradient buffer, averaged as DDP would (per-rank losses are means over the local batch).  (Reducing the heads'
        # slice on a side stream under the encoder's backward gained 5 % at 2 GPUs but the process then hung in its
        # tear-down with the NCCL work captured on a second stream -- profiles/r02_n_allreduce_overlap.md -- so the
        # collective stays one call on the main stream.)
        self._reduce_slice(0, self.flat_g.numel())
        return loss.detach()

    def _eager_step(self, batch):
        loss = self._forward_backward(batch)
        if self.cfg.clip_gradient_norm > 0:
            torch.nn.utils.clip_grad_norm_([self.flat_p], float(self.cfg.clip_gradient_norm))
        self.opt.step()
        return loss

    def _set_step_scalars(self):
        cfg = self.cfg
        lr = float(get_learning_rate(cfg, self.global_step))
        if self.lr is not None:
            self.lr.fill_(lr)
        else:
            for grp in self.opt.param_groups:
                grp["lr"] = lr
        self.sigma.fill_(float(get_smooth_sigma(cfg, self.global_step)))
        # the model builds its taps from this device tensor when a graph is being captured / replayed
        self.model.sigma_override = self.sigma if self.graph else None

    def n_keep(self):
        cfg = self.cfg
        if cfg.pc_point_dropout == 1:
            return int(cfg.pc_num_points)
        return point_cloud.num_points_after_dropout(int(cfg.pc_num_points), get_dropout_prob(cfg, self.global_step))

    def step(self, batch):
        self._set_step_scalars()
        if not self.graph:
            loss = self._eager_step(batch)
            self.global_step += 1
            return loss
        if self.cfg.pc_gauss_filter_gt or self.cfg.pc_gauss_filter_gt_rgb:
            raise NotImplementedError("graph=True: the GT filter's switch-off reads sigma on the host; use graph=False")
        key = self.n_keep()
        if self._static is None:
            self._static = {k: torch.empty_like(v) for k, v in batch.items()}
        for k, v in batch.items():
            if v.data_ptr() != self._static[k].data_ptr():
                self._static[k].copy_(v, non_blocking=True)
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(key)
        g, loss_buf = entry
        g.replay()
        self.global_step += 1
        return loss_buf

    def _capture(self, key):
        """Warm up on a side stream (cuDNN autotuning, lazy initialisation, NCCL communicator), restore the state the
        warm-up steps changed, then capture one step."""
        dev = self.device
        snap_p = self.flat_p.detach().clone()
        snap_opt = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.opt.state.get(self.flat_p, {}).items()}
        snap_draw = self.model.dropout_state.clone()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(3):
                self._eager_step(self._static)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            self.flat_p.copy_(snap_p)
            st = self.opt.state.get(self.flat_p, {})
            for k, v in snap_opt.items():
                if torch.is_tensor(v):
                    st[k].copy_(v)
            if not snap_opt:                      # first capture: Adam's state was created by the warm-up; reset it
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.zero_()
            self.model.dropout_state.copy_(snap_draw)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            loss = self._eager_step(self._static)
            loss_buf = loss.clone()
        torch.cuda.synchronize(dev)
        self._graphs[key] = (g, loss_buf)
        if len(self._graphs) > 8:                 # the schedule moves on: drop the oldest captures
            self._graphs.pop(next(iter(self._graphs)))
        return self._graphs[key]
