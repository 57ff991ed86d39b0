"""SURVEY section 4 item 4, the gloo/CPU variant: a data-parallel train step over two processes -- objects sharded over the
ranks (a rank keeps all views of its objects), the renderer local with no collective, ONE all-reduce of the flat gradient buffer -- gives
the gradients of the single-process step on the concatenated batch.  Tiny model, kernels under the CPU emulation
(tests/emu); the NCCL variant is scripts/ddp_check.py (profiles/r01_n_ddp_check_*.json)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cfg(batch, unsup=False):
    from dpc_b200.util.config import default_config
    over = dict(vox_size=16, image_size=32, pc_num_points=64, batch_size=batch, step_size=2, z_dim=32, fc_dim=32,
                f_dim=4, pc_gauss_kernel_size=5, pc_relative_sigma=1.0, pc_point_dropout=1.0, max_number_of_steps=100)
    if unsup:
        over.update(predict_pose=True, pose_predict_num_candidates=3, pose_predictor_student_loss_weight=20.0)
    return default_config(**over)


def _worker(rank, world, port, ret, unsup=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from dpc_b200 import _capi, distributed as D
    from dpc_b200.train import Trainer, synthetic_batch
    from tests.emu_support import build_emu
    _capi._LIB, _capi._REQUIRE_CUDA = _capi.load_library(build_emu()), False      # test infrastructure: emulated kernels
    D.init(backend="gloo")
    cpu = torch.device("cpu")
    per_rank = 2
    cfg, full = _cfg(per_rank, unsup), _cfg(per_rank * world, unsup)
    big = synthetic_batch(full, cpu, seed=7)
    lo, hi = D.shard_range(full.batch_size, rank, world)
    s = cfg.step_size
    mine = {k: (v[lo * s:hi * s] if k != "images_1" else v[lo:hi]).contiguous() for k, v in big.items()}
    torch.manual_seed(0)
    tr = Trainer(cfg, cpu, ddp=True, bf16=False)
    assert tr.world == world and tr.comm_bytes == tr.flat_g.numel() * 4
    # forward, loss, backward, regulariser, then THE collective of the step: one all-reduce of the flat gradient buffer.
    # The per-rank loss is normalised by the rank's sample count and the reduced sum is divided by the number of ranks:
    # together that is the global 1/num_samples of model_pc.py:415 (equal shard sizes)
    tr._forward_backward(mine)
    worst = None
    if rank == 0:
        torch.manual_seed(0)
        ref = Trainer(full, cpu, ddp=False, bf16=False)
        ref.load_flat(tr)
        ref._forward_backward(big)
        d = (tr.flat_g - ref.flat_g).double()
        worst = float((d ** 2).sum() / max(float((ref.flat_g.double() ** 2).sum()), 1e-30)) ** 0.5
        # every parameter's .grad is a view of the flat buffer
        assert all(p.grad.data_ptr() >= tr.flat_g.data_ptr() for p in tr.model.parameters())
    D.barrier()
    ret[rank] = worst
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("unsup", [False, True])
def test_two_rank_ddp_gradients_equal_the_full_batch_gradients(unsup):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, unsup), nprocs=world, join=True)
    assert ret[0] is not None and ret[0] < 1e-4, ret[0]
