"""CPU, gloo, world_size 2: the data-parallel sharding + single all-gather logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers  # noqa: F401  (sys.path)
from leftrefill_b200 import parallel as P


def test_shard_bounds_cover_and_respect_groups():
    for total, world, group in [(32, 8, 1), (5, 2, 1), (3, 4, 1), (8, 3, 2), (0, 2, 1), (16, 4, 4)]:
        spans = [P.shard_bounds(total, r, world, group) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        assert all((hi - lo) % group == 0 for lo, hi in spans)
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= group


def test_global_noise_is_world_size_independent():
    full = P.global_randn((6, 4, 2, 2), seed=7)
    parts = [P.global_randn((6, 4, 2, 2), seed=7, lo=lo, hi=hi) for lo, hi in
             (P.shard_bounds(6, r, 4) for r in range(4))]
    assert torch.equal(torch.cat(parts), full)


class _ToySampler:
    """Stands in for DDIMSampler on CPU: a deterministic per-sample function of (x_T, cond, noise)."""

    def __init__(self):
        self.noise_source = None

    def sample(self, S, batch_size, shape, conditioning, x_T=None, unconditional_conditioning=None, verbose=False,
               **kw):
        x = x_T.clone()
        for i in range(S):
            n = self.noise_source(x.shape, x.device, i)
            x = 0.9 * x + 0.1 * conditioning["c_concat"][0][:, :4] + 0.01 * n \
                + conditioning["c_crossattn"][0].mean(dim=(1, 2))[:, None, None, None]
        return x, {}


def _worker(rank, world, port, total, group, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w = P.init_distributed(backend="gloo")
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(total, 4, 4, 8, generator=g)
    cond = {"c_concat": [torch.randn(total, 5, 4, 8, generator=g)], "c_crossattn": [torch.randn(total, 7, 16, generator=g)]}
    noise = P.global_randn((3, total, 4, 4, 8), seed=11)
    out = P.sample_sharded(_ToySampler, 3, total, (4, 4, 8), cond, None, x_T, noise, r, w, group=group)
    t = P.max_over_ranks(float(rank + 1), "cpu")
    # ragged gather on its own
    lo, hi = P.shard_bounds(total, r, w, group)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    gathered = P.gather_outputs(local, total, r, w, group)
    if rank == 0:
        q.put((out, t, gathered))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total,group", [(5, 1), (8, 2)])
def test_two_rank_gloo_matches_single_process(total, group):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, group, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, t, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process result
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(total, 4, 4, 8, generator=g)
    cond = {"c_concat": [torch.randn(total, 5, 4, 8, generator=g)], "c_crossattn": [torch.randn(total, 7, 16, generator=g)]}
    noise = P.global_randn((3, total, 4, 4, 8), seed=11)
    ref = P.sample_sharded(_ToySampler, 3, total, (4, 4, 8), cond, None, x_T, noise, 0, 1, group=group)
    assert torch.equal(out, ref)                      # bit-identical regardless of world size
    assert t == 2.0                                   # max over ranks
    assert torch.equal(gathered[:, 0], torch.arange(total, dtype=torch.float32))
