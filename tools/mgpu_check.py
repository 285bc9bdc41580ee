"""torchrun --nproc-per-node N tools/mgpu_check.py : round-robin sharded compression over N GPUs, NCCL gather of the encoded
blocks, ordered commit on rank 0, comparison with the oracle's stream (parity of the multi-GPU path)."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kz = importlib.import_module("kanzi-go_b200")
par = importlib.import_module("kanzi-go_b200.parallel")
synth = importlib.import_module("kanzi-go_b200.synth")
ctx = kz.Context(local)
bs = 1 << 20
x = synth.zipf_bytes(21 * bs + 4321, 1.0, seed=5)  # every rank derives the same slab (stands for the scattered input)
nblocks = (len(x) + bs - 1) // bs
ok = True
for tname, ename in (("NONE", "ANS0"), ("NONE", "HUFFMAN"), ("BWT", "ANS0")):
    t48, et = kz.transform_type(tname), kz.entropy_type(ename)
    mine = par.shard_blocks(nblocks, world, rank)
    enc = ctx.encode_blocks(t48, et, [x[b * bs:(b + 1) * bs] for b in mine])
    blocks = par.gather_blocks([(b, d, n) for b, (d, n) in zip(mine, enc)], nblocks, dist, device="cuda")
    if rank == 0:
        from oracle import pyoracle

        stream = par.assemble_stream(blocks, t48, et, bs, input_size=len(x))
        want = pyoracle.compress(x, tname, ename, block_size=bs, jobs=8, input_size=len(x))
        same = len(stream) == len(want) and np.array_equal(stream, want)
        back = ctx.decompress(stream, len(x) + 64)
        print("mgpu", world, tname, ename, "stream parity:", same, "round trip:", bool(np.array_equal(back, x)))
        ok = ok and same and bool(np.array_equal(back, x))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
