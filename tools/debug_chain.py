"""Localises a round-trip / parity failure of a transform chain on the GPU.

    python tools/debug_chain.py [nblocks] [device|host]

For the l5 workload of bench.py (nblocks x 4 MiB of the text generator, seed SEED + 1) and a list of chains (the full -l 5 chain and its
prefixes) it reports: stream parity GPU vs oracle, cross decoding in both directions, and — per block, through kz_encode_blocks /
kz_decode_blocks — which blocks differ from the oracle's payload or fail to decode. `device` drives kz_compress_stream_device /
kz_decompress_stream_device with torch buffers exactly like bench.py (output capacity = n), `host` the host-buffer entry points.

Open item it was written for (profiles/README.md): `bench.py --workload l5` fails its round trip with 16 blocks while 3 blocks pass."""
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
path = sys.argv[2] if len(sys.argv) > 2 else "host"
x = synth.markov_text(nb * bs, seed=synth.SEED + 1)
n = len(x)
gpu = kz.Context(0)


def gpu_compress(chain, ent):
    if path == "host":
        return gpu.compress(x, chain, ent, block_size=bs, input_size=n)
    import torch

    cap = int(gpu.lib.kz_max_stream_output(n, bs))
    d_src = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    d_cmp = torch.zeros(cap + 256, dtype=torch.uint8, device="cuda")
    d_src[:n].copy_(torch.from_numpy(x))
    torch.cuda.synchronize()
    m = gpu.compress_device(d_src.data_ptr(), n, d_cmp.data_ptr(), cap, kz.transform_type(chain), kz.entropy_type(ent), bs, 0, n)
    return d_cmp[:m].cpu().numpy(), d_cmp


def gpu_decompress(stream, d_cmp=None):
    if path == "host" or d_cmp is None:
        return gpu.decompress(stream, n + 64)
    import torch

    d_dec = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    k = gpu.decompress_device(d_cmp.data_ptr(), len(stream), d_dec.data_ptr(), n)  # capacity exactly n, like bench.py
    return d_dec[:k].cpu().numpy()


for chain, ent in [("TEXT+UTF+BWT+RANK+ZRLT", "ANS0"), ("TEXT+UTF+BWT+RANK+ZRLT", "NONE"), ("TEXT+BWT+RANK", "NONE"), ("TEXT+BWT", "NONE"), ("TEXT", "NONE")]:
    t0 = time.time()
    try:
        r = gpu_compress(chain, ent)
    except Exception as e:
        print(chain, ent, "gpu compress failed:", e, flush=True)
        continue
    g, d_cmp = r if isinstance(r, tuple) else (r, None)
    w = po.compress(x, chain, ent, block_size=bs, jobs=8, input_size=n)
    same = len(g) == len(w) and np.array_equal(g, w)
    first = -1 if same else int(np.argmax(g[: min(len(g), len(w))] != w[: min(len(g), len(w))]))
    res = []
    for name, stream, buf in (("gpu", g, d_cmp), ("oracle", w, None)):
        try:
            y = gpu_decompress(stream, buf)
            res.append("%s-stream->gpu %s" % (name, np.array_equal(y, x)))
        except Exception as e:
            res.append("%s-stream->gpu ERR %s" % (name, str(e)[:60]))
    try:
        res.append("gpu-stream->oracle %s" % np.array_equal(po.decompress(g, n + 64, jobs=8), x))
    except Exception as e:
        res.append("gpu-stream->oracle ERR %s" % str(e)[:60])
    print(chain, ent, "| stream parity", same, "first diff byte", first, "| sizes", len(g), len(w), "|", res, "| %.1fs" % (time.time() - t0), flush=True)
    # per block through the batch entry points
    blocks = [x[i * bs:(i + 1) * bs] for i in range(nb)]
    t48, et = kz.transform_type(chain), kz.entropy_type(ent)
    gpu.set_stream_block_size(bs)
    try:
        enc = gpu.encode_blocks(t48, et, blocks)
        bad_enc = []
        for i, b in enumerate(blocks):
            wp, wbits = po.encode_block(b, t48, et)  # the oracle sizes TEXT's hash map from the block itself: equal here (full blocks)
            if enc[i][1] != wbits or not np.array_equal(enc[i][0], wp):
                bad_enc.append(i)
        bad_dec = []
        for i in range(nb):
            try:
                d = gpu.decode_blocks(t48, et, [enc[i]], bs)[0]
                if not np.array_equal(d, blocks[i]):
                    bad_dec.append(i)
            except Exception as e:
                bad_dec.append((i, str(e)[:40]))
        print("   per block: payload differs from the oracle in", bad_enc, "| decode of the own payload fails in", bad_dec, flush=True)
    finally:
        gpu.set_stream_block_size(0)
