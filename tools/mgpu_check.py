"""torchrun --nproc-per-node N tools/mgpu_check.py : ONE stream over N GPUs (kanzi-go_b200/parallel.py ShardedStream: NCCL scatter of the block
ranges, fragment encode on every rank, gather-v, ordered commit on rank 0; the mirror for decompression) compared with the ORACLE's stream."""
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
sh = par.ShardedStream(par.GpuEngine(ctx, torch), dist, torch)
ok = True
for tname, ename, bs, n in (("NONE", "ANS0", 1 << 20, 21 * (1 << 20) + 4321), ("TEXT+UTF+PACK+MM+LZX", "HUFFMAN", 4 << 20, 40_000_000),
                            ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0", 4 << 20, 40_000_000), ("NONE", "HUFFMAN", 1 << 16, 1 << 16)):
    x = synth.silesia_shaped(n, seed=synth.SEED + 1) if n > (1 << 24) else synth.zipf_bytes(n, 1.0, seed=5)
    t48, et = kz.transform_type(tname), kz.entropy_type(ename)
    slab = None
    if rank == 0:
        slab = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
        slab[:n].copy_(torch.from_numpy(x))
    stream, nb = sh.compress(slab, n, t48, et, bs, 0, n)
    back, m = sh.decompress(stream, nb, n + 4096)
    if rank == 0:
        from oracle import pyoracle

        want = pyoracle.compress(x, tname, ename, block_size=bs, jobs=16, input_size=n)
        got = stream[:nb].cpu().numpy()
        same = len(got) == len(want) and bool(np.array_equal(got, want))
        rt = m == n and bool(np.array_equal(back[:n].cpu().numpy(), x))
        print("mgpu", world, tname, ename, "stream == oracle stream:", same, "sharded round trip:", rt, "NCCL bytes (rank 0, last call):", sh.exchanged_bytes, flush=True)
        ok = ok and same and rt
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
