"""Driver for one ncu capture across the kernel families: small batches (6 blocks of 128 KiB) through every codec / transform
so that each hand-written kernel appears at least once. Used as: ncu --set full -k regex:... python tools/prof_all.py"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

kz = importlib.import_module("kanzi-go_b200")
synth = importlib.import_module("kanzi-go_b200.synth")
ctx = kz.Context(0)
bs = 128 << 10
n = 6 * bs
text = synth.markov_text(n, seed=3)
zipf = synth.zipf_bytes(n, 1.0, seed=4)
for transform, entropy, data, ck in (("NONE", "HUFFMAN", zipf, 0), ("NONE", "ANS1", text, 0), ("NONE", "RANGE", zipf, 0), ("NONE", "ANS0", zipf, 32),
                                     ("BWT+RANK+ZRLT", "ANS0", text, 0), ("BWTS", "NONE", text, 0), ("LZX", "NONE", text, 0), ("LZ", "NONE", text, 0),
                                     ("ROLZ", "NONE", text, 64)):
    s = ctx.compress(data, transform, entropy, block_size=bs, checksum_bits=ck, input_size=n)
    y = ctx.decompress(s, n + 64)
    assert np.array_equal(y, data), (transform, entropy)
    print(transform, entropy, n, len(s))
