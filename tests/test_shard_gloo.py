"""The N>1 path on CPU: two processes over gloo shard a batch of utterances, run a stand-in (deterministic, per-utterance)
sampler on their shard and all-gather the latents; the result must equal the unsharded run in utterance order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ns2vc_b200.shard import gather_latents, shard_bounds, shard_features


def _fake_sampler(x):                      # independent per utterance, like the denoiser (no cross-sample op)
    return torch.tanh(x) * 0.5 + x.mean(dim=(1, 2), keepdim=True)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.randn((6, 5, 17), generator=torch.Generator().manual_seed(3))
        lo, hi = shard_bounds(full.shape[0], world, rank)
        out = gather_latents(_fake_sampler(full[lo:hi]))
        # the feature inputs of the pipeline (ragged lengths) take the same cut: a per-utterance stand-in for encoders + sampler
        lengths = torch.tensor([17, 9, 13, 17, 5, 11])
        x_l, len_l = shard_features(world, rank, full, lengths)
        masked = lambda x, n: _fake_sampler(x * (torch.arange(x.shape[2])[None, None, :] < n[:, None, None]))
        out2 = gather_latents(masked(x_l, len_l))
        q.put((rank, torch.equal(out, _fake_sampler(full)) and torch.equal(out2, masked(full, lengths)), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_and_gather_in_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (6, 5, 17)), (1, True, (6, 5, 17))]


def test_shard_bounds_reject_uneven_batches():
    assert [shard_bounds(64, 8, r) for r in (0, 7)] == [(0, 8), (56, 64)]
    with pytest.raises(ValueError):
        shard_bounds(10, 4, 0)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
    x = torch.zeros(2, 3, 4)
    assert gather_latents(x) is x          # no process group: single rank
    a, b = shard_features(2, 1, torch.arange(8).view(4, 2), torch.arange(4))
    assert a.tolist() == [[4, 5], [6, 7]] and b.tolist() == [2, 3]
    with pytest.raises(ValueError):
        shard_features(2, 0, torch.zeros(4, 1), torch.zeros(3))
