"""Localises a round-trip failure of a transform chain: GPU stream vs oracle stream, cross decoding, and prefixes of the chain."""
import importlib
import sys
import time

import numpy as np

sys.path.insert(0, ".")
kz = importlib.import_module("kanzi-go_b200")
synth = importlib.import_module("kanzi-go_b200.synth")
from oracle import pyoracle as po

po.build()
bs = 4 << 20
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = synth.markov_text(nb * bs, seed=synth.SEED + 1)
n = len(x)
gpu = kz.Context(0)
for chain, ent in [("TEXT+UTF+BWT+RANK+ZRLT", "ANS0"), ("TEXT+BWT", "NONE"), ("TEXT", "NONE"), ("TEXT+BWT+RANK", "NONE")]:
    t0 = time.time()
    try:
        g = gpu.compress(x, chain, ent, block_size=bs, input_size=n)
    except Exception as e:
        print(chain, "gpu compress failed:", e)
        continue
    w = po.compress(x, chain, ent, block_size=bs, jobs=8, input_size=n)
    same = len(g) == len(w) and np.array_equal(g, w)
    first = -1 if same else int(np.argmax(g[: min(len(g), len(w))] != w[: min(len(g), len(w))]))
    res = []
    for name, stream in (("gpu", g), ("oracle", w)):
        try:
            y = gpu.decompress(stream, n + 64)
            res.append("%s-stream->gpu %s" % (name, np.array_equal(y, x)))
        except Exception as e:
            res.append("%s-stream->gpu ERR %s" % (name, str(e)[:60]))
    try:
        y = po.decompress(g, n + 64, jobs=8)
        res.append("gpu-stream->oracle %s" % np.array_equal(y, x))
    except Exception as e:
        res.append("gpu-stream->oracle ERR %s" % str(e)[:60])
    print(chain, ent, "parity", same, "first diff", first, "len", len(g), len(w), res, "%.1fs" % (time.time() - t0), flush=True)
