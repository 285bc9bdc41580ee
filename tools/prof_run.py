"""Tiny driver for ncu captures: N passes of compress+decompress of the bench workload through the device API."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

kz = importlib.import_module("kanzi-go_b200")
synth = importlib.import_module("kanzi-go_b200.synth")
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nblocks = int(sys.argv[2]) if len(sys.argv) > 2 else 64
BLOCK = 4 << 20
n = nblocks * BLOCK
ctx = kz.Context(0)
x = synth.zipf_bytes(n, 1.0, seed=synth.SEED + 1)
cap = int(ctx.lib.kz_max_stream_output(n, BLOCK))
d_src = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
d_cmp = torch.zeros(cap + 256, dtype=torch.uint8, device="cuda")
d_dec = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
d_src[:n].copy_(torch.from_numpy(x))
torch.cuda.synchronize()
for _ in range(passes):
    m = ctx.compress_device(d_src.data_ptr(), n, d_cmp.data_ptr(), cap, 0, kz.E_ANS0, BLOCK, 0, n)
    k = ctx.decompress_device(d_cmp.data_ptr(), m, d_dec.data_ptr(), n)
torch.cuda.synchronize()
assert k == n and torch.equal(d_dec[:n], d_src[:n])
print("ok", n, m)
