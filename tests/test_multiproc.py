"""CPU suite, world_size 2 over gloo: the N>1 path is line-block sharding + host concatenation.
Each rank encodes its shard (with the oracle standing in for the GPU encoder: this test is about the
host logic) and rank 0 assembles; the result must equal a single-shard encode byte for byte."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_lines, n_samples, q):
    import oraclelib as O
    import vcfgen
    sharding = importlib.import_module("vcf-compression_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, data = vcfgen.random_vcf_like(n_lines, n_samples, seed=77)
    a, b = sharding.split_ranges(data, world)[rank]
    rc, enc, nl, _, offs = O.compress_block(data[a:b], want_offsets=True)
    assert rc == 0
    # timing plumbing of bench.py: max over ranks of a per-rank scalar
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    gathered = [None] * world
    dist.all_gather_object(gathered, (enc, nl, offs))
    if rank == 0:
        got, lines, goffs = sharding.gather_concat(gathered)
        rc, ref, rnl, _, roffs = O.compress_block(data, want_offsets=True)
        q.put((got == ref, lines == rnl, goffs == roffs))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_lines,n_samples", [(101, 37), (7, 2504)])
def test_two_rank_sharding_matches_single_shard(n_lines, n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_lines, n_samples, q)) for r in range(2)]
    [p.start() for p in procs]
    res = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res == (True, True, True)


def test_split_ranges_properties():
    sharding = importlib.import_module("vcf-compression_b200.sharding")
    data = b"".join(b"line%d\tx\n" % i for i in range(50))
    for g in (1, 2, 3, 8, 64):
        r = sharding.split_ranges(data, g)
        assert len(r) == g and r[0][0] == 0 and r[-1][1] == len(data)
        for (a, b), (c, d) in zip(r, r[1:]):
            assert b == c and a <= b
        for a, b in r:
            assert a == b or data[b - 1:b] == b"\n"
    assert sharding.split_ranges(b"", 4) == [(0, 0)] * 4
    assert sharding.split_ranges(b"no newline", 2) == [(0, 10), (10, 10)]
