"""Times one transform stage on one block (for ncu captures and variant comparisons): python tools/prof_stage.py rank_inv [random|text] [n]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

kz = importlib.import_module("kanzi-go_b200")
synth = importlib.import_module("kanzi-go_b200.synth")
what = sys.argv[1] if len(sys.argv) > 1 else "rank_inv"
kind = sys.argv[2] if len(sys.argv) > 2 else "random"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4 << 20
ctx = kz.Context(0)
T_BWT, T_RANK, T_MM = 1, 8, 15
x = synth.uniform_bytes(n, seed=5) if kind == "random" else synth.english_text(n, seed=5)
if what == "rank_inv":
    if kind == "random":
        ranks = x  # any byte string is a rank stream
    else:
        bw, _ = ctx.transform_forward(T_BWT, x)
        ranks, _ = ctx.transform_forward(T_RANK, bw)
    ctx.profile(True)
    for _ in range(3):
        ctx.profile_reset()
        y = ctx.transform_inverse(T_RANK, ranks, len(ranks) + 64)
        print(kind, len(ranks), {k: round(ctx.kernel_time(k)[1], 3) for k in ctx.profile_names()}, "zeros %.2f mean %.1f" % (float((ranks == 0).mean()), float(ranks.mean())))
elif what == "mm_inv":
    w = synth.walk16(n, seed=5)
    enc, _ = ctx.transform_forward(T_MM, w)
    ctx.profile(True)
    for _ in range(3):
        ctx.profile_reset()
        y = ctx.transform_inverse(T_MM, enc, n + 64)
        print("walk16", len(enc), {k: round(ctx.kernel_time(k)[1], 3) for k in ctx.profile_names()}, "escapes %.4f" % float((enc == 255).mean()))
    assert np.array_equal(y, w)
