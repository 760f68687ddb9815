"""CPU tests (gloo, world_size 2 and 3) of the clip-sharding host logic used for multi-GPU runs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from koemorph_b200.parallel import ShardedInference, gather_outputs, shard_range, shard_sizes


def test_shard_ranges_cover_everything():
    for n in (0, 1, 5, 512, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_forward(audio, egemaps):
    # stands in for the CUDA forward: any clip-wise function (the real one needs a GPU)
    return (audio.mean(dim=1, keepdim=True) + egemaps[:, :52]).unsqueeze(1)


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        audio = torch.randn(n_clips, 100, generator=g)
        eg = torch.randn(n_clips, 264, generator=g)
        out = ShardedInference(_fake_forward)(audio, eg)
        ref = _fake_forward(audio, eg)
        lo, hi = shard_range(n_clips, rank, world)
        local = ShardedInference(_fake_forward)(audio, eg, gather=False)
        ok = torch.equal(out, ref) and local.shape[0] == hi - lo and torch.equal(local, ref[lo:hi])
        try:
            gather_outputs(ref[:0], n_clips)
            ok = ok and (hi - lo == 0)
        except ValueError:
            pass
        q.put((rank, bool(ok), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clips", [(2, 7), (2, 8), (3, 10)])
def test_sharded_inference_gloo(world, n_clips):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res) and all(r[2] == (n_clips, 1, 52) for r in res)


def test_gather_without_process_group_is_identity():
    x = torch.randn(4, 1, 52)
    assert gather_outputs(x, 4) is x
    with pytest.raises(ValueError):
        gather_outputs(x, 5)
