"""Two contexts, two host threads: the -l 3 and the -l 5 round trips of the metric's slab one after the other, then concurrently
(one CUDA stream each). Prints ms per step of both schedules and checks that the concurrent streams are the sequential ones."""
import importlib
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

kz = importlib.import_module("kanzi-go_b200")
synth = importlib.import_module("kanzi-go_b200.synth")


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    passes, BLOCK, n, kind, _ = bench.WORKLOADS["silesia"]
    torch.cuda.set_device(0)
    x = bench.make_data(kind, n, synth.SEED + 1)
    d_src = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    d_src[:n].copy_(torch.from_numpy(x))
    ctxs = [kz.Context(0) for _ in passes]
    cap = int(ctxs[0].lib.kz_max_stream_output(n, BLOCK))
    P = [(label, kz.transform_type(tr), kz.entropy_type(en)) for label, tr, en in passes]
    d_cmp = [torch.zeros(cap + 256, dtype=torch.uint8, device="cuda") for _ in P]
    d_ref = [None] * len(P)
    d_dec = [torch.empty(n + 256, dtype=torch.uint8, device="cuda") for _ in P]
    m = [0] * len(P)

    gate = threading.Barrier(len(passes))

    def chain(i, k, gated=False):
        label, t48, et = P[i]
        for _ in range(k):
            if gated:
                gate.wait()
            m[i] = ctxs[i].compress_device(d_src.data_ptr(), n, d_cmp[i].data_ptr(), cap, t48, et, BLOCK, 0, n)
            got = ctxs[i].decompress_device(d_cmp[i].data_ptr(), m[i], d_dec[i].data_ptr(), n)
            assert got == n

    for i in range(len(P)):
        chain(i, 2)
        torch.cuda.synchronize()
        assert torch.equal(d_dec[i][:n], d_src[:n])
        d_ref[i] = d_cmp[i][: m[i]].clone()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def table(tag):
        rows = []
        for c in ctxs:
            for name in c.profile_names():
                if not name.startswith("stage:"):
                    cnt, ms = c.kernel_time(name)
                    if cnt:
                        rows.append((ms / steps, cnt, name))
            c.profile_reset()
        rows.sort(reverse=True)
        print(tag, "  ".join("%s %.1f" % (r[2], r[0]) for r in rows[:14]))

    def run(concurrent):
        for i in range(len(P)):
            d_cmp[i].zero_()
            d_dec[i].zero_()
        torch.cuda.synchronize()
        e[0].record()
        if concurrent:
            th = [threading.Thread(target=chain, args=(i, steps, True)) for i in range(len(P))]
            for t in th:
                t.start()
            for t in th:
                t.join()
        else:
            for i in range(len(P)):
                chain(i, steps)
        torch.cuda.synchronize()
        e[1].record()
        torch.cuda.synchronize()
        same = all(torch.equal(d_cmp[i][: m[i]], d_ref[i]) and torch.equal(d_dec[i][:n], d_src[:n]) for i in range(len(P)))
        assert same, "streams differ"
        return e[0].elapsed_time(e[1]) / steps

    mb = 2 * n * len(P) / 1e6
    for prof in (False, True):
        for c in ctxs:
            c.profile(prof)
            c.profile_reset()
        for concurrent in (False, True, True):
            ms = run(concurrent)
            print("profile %s, %s: %.1f ms/step (%.0f MB/s)" % (prof, "concurrent" if concurrent else "sequential", ms, mb / ms * 1e3), flush=True)
            if prof:
                table("   ")


if __name__ == "__main__":
    main()
