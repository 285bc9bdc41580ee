"""N > 1 host logic on CPU (gloo, world_size 2): round-robin sharding, gather-v of encoded blocks, ordered bit-packed
commit on rank 0. The block bit strings come from the oracle here (no GPU in this test); the assembled stream must be the
oracle's own stream byte for byte, which checks the commit logic against Writer.processBlock's."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist

        from oracle import pyoracle

        par = importlib.import_module("kanzi-go_b200.parallel")
        synth = importlib.import_module("kanzi-go_b200.synth")
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        bs = 65536
        x = synth.zipf_bytes(7 * bs + 1234, 1.0, seed=11)
        nblocks = (len(x) + bs - 1) // bs
        mine = par.shard_blocks(nblocks, world, rank)
        local = []
        for b in mine:
            data, nbits = pyoracle.encode_block(x[b * bs:(b + 1) * bs], 0, pyoracle.E_ANS0)
            local.append((b, data, nbits))
        blocks = par.gather_blocks(local, nblocks, dist)
        if rank == 0:
            stream = par.assemble_stream(blocks, 0, pyoracle.E_ANS0, bs, input_size=len(x))
            want = pyoracle.compress(x, "NONE", "ANS0", block_size=bs, jobs=2, input_size=len(x))
            q.put(("ok", bool(len(stream) == len(want) and np.array_equal(stream, want))))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e)))


def test_shard_blocks(kz):
    par = importlib.import_module("kanzi-go_b200.parallel")
    assert par.shard_blocks(10, 4, 1) == [1, 5, 9]
    allb = sorted(sum((par.shard_blocks(13, 8, r) for r in range(8)), []))
    assert allb == list(range(13))


def test_gather_and_commit_world2(oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    kind, val = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert kind == "ok" and val is True, (kind, val)
