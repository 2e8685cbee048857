"""GPU, NCCL, world_size 2 (skipped on a single-GPU box; run with `gpurun --gpus 2`): the data-parallel sampler.

`parallel.sample_sharded` over 2 ranks must give BIT-IDENTICAL samples to 1 rank: canvases are independent, the noise
is drawn for the global batch and sliced (parallel.global_randn), and no kernel choice that changes a summation order
depends on the batch size (split-K is decided per image, GroupNorm reductions run in a fixed order)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

import helpers
from helpers import FakeLDM, O, synthetic_inputs

pytestmark = pytest.mark.gpu

S, B, H, W = 4, 4, 16, 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, cfg_name, q):
    import leftrefill_b200 as lr
    from leftrefill_b200 import parallel as P
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    r, w = P.init_distributed(backend="nccl" if world > 1 else None)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    cfg = getattr(O, cfg_name)
    sd = O.make_state_dict(cfg, seed=0)
    m = lr.UNetModel(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    x_T, c_cat, ctx, uc = synthetic_inputs(B, h=H, w=W, ctx_dim=cfg["context_dim"], device=dev)
    noise = P.global_randn((S, B, 4, H, W), seed=11, device=dev)
    cond = {"c_concat": [c_cat], "c_crossattn": [ctx]}
    ucond = {"c_concat": [c_cat], "c_crossattn": [uc]}
    out = P.sample_sharded(lambda: lr.DDIMSampler(FakeLDM(m, dev)), S, B, (4, H, W), cond, ucond, x_T, noise, r, w,
                           eta=1.0, unconditional_guidance_scale=2.5)
    t_max = P.max_over_ranks(float(rank + 1), dev)
    if rank == 0:
        q.put((out.cpu().numpy(), t_max))  # by value: the child may exit before the parent reads the queue
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def _launch(world, cfg_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, cfg_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    arr, t_max = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return torch.from_numpy(arr), t_max


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL)")
def test_two_ranks_equal_one_rank_bit_for_bit():
    one, _ = _launch(1, "SMALL_CFG")
    two, t_max = _launch(2, "SMALL_CFG")
    assert t_max == 2.0                                   # max-over-ranks reduction went through NCCL
    assert one.shape == two.shape == (B, 4, H, W)
    assert torch.equal(one, two), (one - two).abs().max().item()
