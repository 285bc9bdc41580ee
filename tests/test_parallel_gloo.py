"""N > 1 host logic on CPU (gloo, world_size 2): kanzi-go_b200/parallel.py's ShardedStream — scatter of the block ranges, fragment encode,
gather-v, ordered bit-granular commit on rank 0, and the mirror for decompression — with a CPU stand-in for the engine built around the
ORACLE (no GPU in this test). The stream rank 0 commits must be the oracle's own stream byte for byte, and the sharded decompression
must return the input."""
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


def _bits(a, nbits):
    return np.unpackbits(np.ascontiguousarray(a, np.uint8))[:nbits]


class OracleEngine:
    """The engine interface of parallel.GpuEngine on CPU tensors: a fragment = the oracle's stream of the range without header and end marker."""

    def __init__(self, torch, pyoracle, par, tname, ename):
        self.torch, self.o, self.par, self.tname, self.ename = torch, pyoracle, par, tname, ename

    def empty(self, n):
        return self.torch.zeros(int(n) + 256, dtype=self.torch.uint8)

    zeros = empty

    def compress_fragment(self, src, n, t48, et, block_size, checksum_bits):
        if n == 0:
            return self.empty(0), 0
        x = src[:n].numpy()
        s = self.o.compress(x, self.tname, self.ename, block_size=block_size, checksum_bits=checksum_bits, jobs=2, input_size=0)
        hb = self.par.stream_header_bits(0)
        bits = _bits(s, 8 * len(s))
        # the stream ends with the 8-bit end marker and zero padding: find the exact body by walking the records
        pos, end = hb, hb
        while True:
            lw = int("".join(map(str, bits[pos:pos + 5])), 2) + 3
            ln = int("".join(map(str, bits[pos + 5:pos + 5 + lw])), 2)
            if ln == 0:
                end = pos
                break
            pos += 5 + lw + ln
        body = bits[hb:end]
        out = self.empty((len(body) + 7) // 8)
        packed = np.packbits(body)
        out[: len(packed)] = self.torch.from_numpy(packed)
        return out, len(body)

    def commit(self, t48, et, block_size, checksum_bits, input_size, frags, frag_bits, cap):
        ref = self.o.compress(np.zeros(0, np.uint8), self.tname, self.ename, block_size=block_size, checksum_bits=checksum_bits, jobs=1, input_size=input_size)
        hb = self.par.stream_header_bits(input_size)
        parts = [_bits(ref, hb)] + [_bits(f.numpy(), b) for f, b in zip(frags, frag_bits)] + [np.zeros(8, np.uint8)]
        packed = np.packbits(np.concatenate(parts))
        out = self.empty(len(packed))
        out[: len(packed)] = self.torch.from_numpy(packed)
        return out, len(packed)

    def index(self, stream, nbytes, max_blocks):
        bits = _bits(stream[:nbytes].numpy(), 8 * nbytes)
        # header fields through the oracle's own parser would need an export; the test knows them
        input_size = self.input_size
        hb = self.par.stream_header_bits(input_size)
        rec, pos = [hb], hb
        while True:
            lw = int("".join(map(str, bits[pos:pos + 5])), 2) + 3
            ln = int("".join(map(str, bits[pos + 5:pos + 5 + lw])), 2)
            if ln == 0:
                break
            pos += 5 + lw + ln
            rec.append(pos)
        return {"transform48": self.t48, "entropy": self.et, "block_size": self.block_size, "checksum_bits": 0, "input_size": input_size, "rec_bit": rec,
                "nblocks": len(rec) - 1}

    def decompress_fragment(self, frag, frag_bytes, start_bit, nblocks, t48, et, block_size, checksum_bits, cap):
        bits = _bits(frag[:frag_bytes].numpy(), 8 * frag_bytes)
        pos = start_bit
        for _ in range(nblocks):
            lw = int("".join(map(str, bits[pos:pos + 5])), 2) + 3
            ln = int("".join(map(str, bits[pos + 5:pos + 5 + lw])), 2)
            pos += 5 + lw + ln
        ref = self.o.compress(np.zeros(0, np.uint8), self.tname, self.ename, block_size=block_size, jobs=1, input_size=0)
        hb = self.par.stream_header_bits(0)
        s = np.packbits(np.concatenate([_bits(ref, hb), bits[start_bit:pos], np.zeros(8, np.uint8)]))
        y = self.o.decompress(s, cap, jobs=2)
        out = self.empty(len(y))
        out[: len(y)] = self.torch.from_numpy(np.ascontiguousarray(y))
        return out, len(y)


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist

        from oracle import pyoracle
        from test_parallel_gloo import OracleEngine

        par = importlib.import_module("kanzi-go_b200.parallel")
        synth = importlib.import_module("kanzi-go_b200.synth")
        kz = importlib.import_module("kanzi-go_b200")
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        ok = True
        for tname, ename, bs, n in (("NONE", "ANS0", 65536, 7 * 65536 + 1234), ("TEXT+UTF+PACK+MM+LZX", "HUFFMAN", 32768, 5 * 32768 + 77), ("NONE", "HUFFMAN", 4096, 4096)):
            x = synth.markov_text(n, seed=11)
            t48, et = kz.transform_type(tname), kz.entropy_type(ename)
            eng = OracleEngine(torch, pyoracle, par, tname, ename)
            eng.t48, eng.et, eng.block_size, eng.input_size = t48, et, bs, n
            sh = par.ShardedStream(eng, dist, torch)
            slab = torch.from_numpy(x.copy()) if rank == 0 else None
            stream, nbytes = sh.compress(slab, n, t48, et, bs, 0, n)
            if rank == 0:
                want = pyoracle.compress(x, tname, ename, block_size=bs, jobs=2, input_size=n)
                got = stream[:nbytes].numpy()
                ok = ok and len(got) == len(want) and bool(np.array_equal(got, want))
                assert sh.exchanged_bytes > 0 or n <= bs
            back, m = sh.decompress(stream if rank == 0 else None, nbytes, n + 4096)
            if rank == 0:
                ok = ok and m == n and bool(np.array_equal(back[:m].numpy(), x))
        if rank == 0:
            q.put(("ok", ok))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put(("err", traceback.format_exc()))


def test_block_ranges(kz):
    par = importlib.import_module("kanzi-go_b200.parallel")
    assert par.block_ranges(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert par.block_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert par.block_ranges(0, 2) == [(0, 0), (0, 0)]
    for nb in (1, 7, 48, 49):
        for w in (1, 2, 4, 8):
            r = par.block_ranges(nb, w)
            assert r[0][0] == 0 and r[-1][1] == nb and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_sharded_stream_world2(oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    kind, val = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert kind == "ok" and val is True, (kind, val)
