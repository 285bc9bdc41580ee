#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 block-compression engine (contract: see the task statement / DESIGN.md).

Workload at N=1 (BASELINE.json configs[1]): `-e ANS0 -t NONE`, 256 MiB of Zipf(1.0) bytes, 64 x 4 MiB blocks.
A "step" is one pass of the hot path over that batch: compress the whole slab to a kanzi bitstream, then decompress it.

  value  = uncompressed bytes through (encode + decode) per second with inputs resident in HBM (device API), MB = 1e6 B
  e2e    = the same through the host-buffer C ABI (kz_compress_stream / kz_decompress_stream) from pinned host
           memory, host<->device copies inside the timed region
  roofline: dominant kernel ans0_decode, achieved = (compressed bytes read + decoded bytes written) / CUDA-event time
  cpu_baseline / --impl reference: the C++ restatement of the reference path (oracle/, "port": no Go toolchain here)
           with one worker per block on all host cores, on a bounded sample of the same workload.
With N > 1 ranks every rank processes its own 64-block slab (weak scaling, no data-path collective: blocks are
independent); time = max over ranks.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BLOCK = 4 << 20
NBLOCKS = 64
METRIC = "encode+decode MB/s (ANS0/NONE, 64x4MiB Zipf(1.0) blocks per GPU)"

# The default workload is BASELINE.json configs[1] (the configuration the roofline target is quoted on). The others are
# the GPU-covered parts of the remaining configs, run by hand with --workload and recorded under profiles/ (they are
# parity-test cases first; the driver's bench line is always the default).
#   name: (transform, entropy, block size, blocks per GPU, data generator, dominant kernel for the roofline, metric label)
WORKLOADS = {
    "ans0": ("NONE", "ANS0", 4 << 20, 64, "zipf1", "ans0_decode", METRIC),
    "huffman": ("NONE", "HUFFMAN", 4 << 20, 64, "zipf1", "huf_decode", "encode+decode MB/s (HUFFMAN/NONE, 64x4MiB Zipf(1.0) blocks per GPU)"),
    "ans1": ("NONE", "ANS1", 4 << 20, 64, "text", "ans1_decode", "encode+decode MB/s (ANS1/NONE, 64x4MiB order-1 text blocks per GPU; configs[4] at 4 MiB)"),
    "l3core": ("LZX", "HUFFMAN", 4 << 20, 48, "text", "lz_parse",
               "encode+decode MB/s (-t LZX -e HUFFMAN, 48x4MiB text blocks: the GPU covered part of kanzi -l 3, configs[2] shape)"),
    "l5core": ("BWT+RANK+ZRLT", "ANS0", 32 << 20, 4, "text", "bwt_forward",
               "encode+decode MB/s (-t BWT+RANK+ZRLT -e ANS0, 4x32MiB text blocks: the GPU covered part of kanzi -l 5, configs[3] shape)"),
    "l5core4m": ("BWT+RANK+ZRLT", "ANS0", 4 << 20, 16, "text", "bwt_forward",
                 "encode+decode MB/s (-t BWT+RANK+ZRLT -e ANS0, 16x4MiB text blocks)"),
    # complete level chains (app/BlockCompressor.go:664-700); TEXT / UTF walk their blocks with one thread each
    "l2": ("DNA+LZ", "HUFFMAN", 4 << 20, 48, "text", "lz_parse", "encode+decode MB/s (kanzi -l 2 = -t DNA+LZ -e HUFFMAN, 48x4MiB text blocks)"),
    "l3": ("TEXT+UTF+PACK+MM+LZX", "HUFFMAN", 4 << 20, 48, "text", "lz_parse",
           "encode+decode MB/s (kanzi -l 3 = -t TEXT+UTF+PACK+MM+LZX -e HUFFMAN, 48x4MiB text blocks, configs[2] shape)"),
    "l5": ("TEXT+UTF+BWT+RANK+ZRLT", "ANS0", 4 << 20, 16, "text", "bwt_forward",
           "encode+decode MB/s (kanzi -l 5 = -t TEXT+UTF+BWT+RANK+ZRLT -e ANS0, 16x4MiB text blocks)"),
}


def make_data(kind, n, seed):
    synth = importlib.import_module("kanzi-go_b200.synth")
    if kind == "zipf1":
        return synth.zipf_bytes(n, 1.0, seed=seed)
    if kind == "text":
        return synth.markov_text(n, seed=seed)
    raise ValueError(kind)


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the newest committed `ncu --set full` summary
    (profiles/*ncu*summary.json, written by tools/ncu_summary.py): per launch of the default workload, or None."""
    import glob

    best = None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu*summary*.json")), key=lambda f: ("final" in os.path.basename(f), os.path.basename(f)))
    for f in files:  # the capture named "final" wins, otherwise the highest generation
        try:
            for k in json.load(open(f)).get("kernels", []):
                if kernel in k.get("kernel", "") and k.get("dram_traffic_bytes"):
                    best = (float(k["dram_traffic_bytes"]), os.path.basename(f))
        except Exception:
            pass
    return best


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip().split("\n")[0]
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_reference(steps, warmup, sample_blocks=None, workload="ans0"):
    """Times the CPU restatement of the reference path (oracle) with one worker per block on all host cores."""
    from oracle import pyoracle

    synth = importlib.import_module("kanzi-go_b200.synth")
    transform, entropy, BLOCK, NBLOCKS, kind = WORKLOADS[workload][:5]
    cores = os.cpu_count() or 1
    jobs = min(64, cores)
    if sample_blocks is None:
        sample_blocks = NBLOCKS  # the full batch is ~1-3 s of CPU work per step on a multi-core host
    n = sample_blocks * BLOCK
    x = make_data(kind, n, synth.SEED + 1)
    times = []
    for it in range(warmup + steps):
        te, td = [], []
        s = pyoracle.compress(x, transform, entropy, block_size=BLOCK, jobs=jobs, input_size=n, timing=te)
        y = pyoracle.decompress(s, n + 64, jobs=jobs, timing=td)
        if it == 0:
            assert np.array_equal(y, x)
        if it >= warmup:
            times.append(te[0] + td[0])
    t = float(np.mean(times))
    return {"value": 2 * n / t / 1e6, "unit": "MB/s", "cores": jobs, "kind": "port",
            "sample": "%d x %d MiB %s blocks (%d MiB), -t %s -e %s, compress+decompress in memory, %d worker threads, C++ restatement of kanzi-go (no Go toolchain)"
                      % (sample_blocks, BLOCK >> 20, kind, n >> 20, transform, entropy, jobs),
            "ms_per_step": t * 1e3}


def workload_label(name):
    transform, entropy, block, nblocks, kind = WORKLOADS[name][:5]
    label = "%s/%s %dx%dMiB %s" % (entropy, transform, nblocks, block >> 20, {"zipf1": "Zipf(1.0)", "text": "order-1 text"}[kind])
    return label + (" (BASELINE.json configs[1])" if name == "ans0" else "")


def emit(line):
    """Writes the one JSON line to the real stdout (fd 1 is pointed at stderr during the run so that library chatter such
    as "NCCL version ..." cannot end up in front of it)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="ans0", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    transform, entropy, BLOCK, NBLOCKS, kind, roof_kernel, METRIC = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = min(args.steps, 5)
        r = cpu_reference(steps, min(args.warmup, 1), workload=args.workload)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MB/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "%s, CPU path" % workload_label(args.workload), "block_size": BLOCK, "blocks": NBLOCKS},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    import torch

    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kz = importlib.import_module("kanzi-go_b200")
    synth = importlib.import_module("kanzi-go_b200.synth")
    torch.cuda.set_device(local_rank)
    ctx = kz.Context(local_rank)
    n = NBLOCKS * BLOCK
    x = make_data(kind, n, synth.SEED + 1 + rank)
    t48, et = kz.transform_type(transform), kz.entropy_type(entropy)
    cap = int(ctx.lib.kz_max_stream_output(n, BLOCK))

    # ---------------- device-resident pass (value) ----------------
    d_src = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    d_cmp = torch.zeros(cap + 256, dtype=torch.uint8, device="cuda")
    d_dec = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    d_src[:n].copy_(torch.from_numpy(x))
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(ctx.cuda_stream())

    def step_device():
        m = ctx.compress_device(d_src.data_ptr(), n, d_cmp.data_ptr(), cap, t48, et, BLOCK, 0, n)
        k = ctx.decompress_device(d_cmp.data_ptr(), m, d_dec.data_ptr(), n)
        return m, k

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    m = 0
    for _ in range(args.warmup):
        m, k = step_device()
    assert k == n and torch.equal(d_dec[:n], d_src[:n]), "round trip mismatch"
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile(True)
    ctx.profile_reset()
    ctx.launch_count(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_ms = dec_ms = 0.0
    with torch.cuda.stream(ext):
        e0.record()
        for _ in range(args.steps):
            a, b, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m = ctx.compress_device(d_src.data_ptr(), n, d_cmp.data_ptr(), cap, t48, et, BLOCK, 0, n)
            b.record()
            ctx.decompress_device(d_cmp.data_ptr(), m, d_dec.data_ptr(), n)
            c.record()
            c.synchronize()
            enc_ms += a.elapsed_time(b)
            dec_ms += b.elapsed_time(c)
        e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count(reset=True)
    kern = {}
    for name in ("ans0_decode", "ans0_tables", "ans0_encode", "ans0_stats", "ans_walk", "concat", "concat_zero", "scan", "stream_walk", "block_header", "block_prefix", "extract",
                 "huf_stats", "huf_encode", "huf_walk", "huf_decode", "ans1_hist", "ans1_stats", "ans1_encode", "ans1_decode", "range_encode", "range_decode",
                 "bwt_forward", "bwt_inverse", "bwts_forward", "bwts_inverse", "lz_parse", "lz_gather", "lz_inverse", "sbrt_last2", "sbrt_carry", "sbrt_rank",
                 "sbrt_inverse", "zrlt_forward", "zrlt_inverse", "xxhash", "rolz_forward", "rolz_inverse", "alias_forward", "alias_inverse", "fsd_forward",
                 "fsd_inverse", "text_forward", "text_inverse", "utf_forward", "utf_inverse", "exe_forward", "exe_inverse", "host:encode_batch", "host:enc_tables", "host:enc_sync"):
        cnt, ms = ctx.kernel_time(name)
        if cnt:
            kern[name] = {"launches": cnt, "ms_avg": ms / cnt}
    ctx.profile(False)

    # ---------------- end-to-end pass through the host-buffer C ABI (pinned host memory) ----------------
    h_src = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_cmp = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_dec = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
    h_src.numpy()[:] = x
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        mm = ctx.compress_host(h_src.data_ptr(), n, h_cmp.data_ptr(), cap, t48, et, BLOCK, 0, n)
        kk = ctx.decompress_host(h_cmp.data_ptr(), mm, h_dec.data_ptr(), n + 64)
    assert kk == n and np.array_equal(h_dec.numpy()[:n], x), "e2e round trip mismatch"
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(ext):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(e2e_steps):
            mm = ctx.compress_host(h_src.data_ptr(), n, h_cmp.data_ptr(), cap, t48, et, BLOCK, 0, n)
            ctx.decompress_host(h_cmp.data_ptr(), mm, h_dec.data_ptr(), n + 64)
        f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---------------- max over ranks ----------------
    vals = torch.tensor([dev_ms, e2e_ms, enc_ms, dec_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, enc_ms, dec_ms = [float(v) for v in vals.tolist()]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    total_bytes = 2.0 * n * world  # uncompressed bytes through encode and through decode, all ranks
    ms_per_step = dev_ms / args.steps
    value = total_bytes * args.steps / (dev_ms / 1e3) / 1e6
    e2e_value = total_bytes * e2e_steps / (e2e_ms / 1e3) / 1e6
    peak, peak_src = peaks()
    roof = None
    if roof_kernel in kern:
        # algorithmic bytes per launch: entropy stage = compressed + uncompressed bytes of the batch; a transform
        # launched per block (BWT) = 2 x block bytes; the serial LZ parse = block bytes in + out of the batch
        per_block = roof_kernel in ("bwt_forward", "bwt_inverse")
        alg = float(2 * BLOCK) if per_block else float(n + m)
        ach = alg / (kern[roof_kernel]["ms_avg"] / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": roof_kernel + ("_device (all kernels of one block)" if per_block else "_kernel"), "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                "ms_per_launch": kern[roof_kernel]["ms_avg"]}
        if args.workload == "ans0":
            t = measured_traffic("ans0_decode")
            if t:
                roof["traffic"] = t[0]
                roof["traffic_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/" + t[1]
    line = {
        "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_label(args.workload), "block_size": BLOCK, "blocks_per_gpu": NBLOCKS,
                   "uncompressed_bytes_per_gpu": n, "compressed_bytes_per_gpu": int(m),
                   "l2": "inputs (%d MiB in, %d MiB compressed) larger than the 126 MB L2" % (n >> 20, m >> 20),
                   "encode_MBps": n * world * args.steps / (enc_ms / 1e3) / 1e6, "decode_MBps": n * world * args.steps / (dec_ms / 1e3) / 1e6,
                   "parallelism": "1 process per GPU, %d blocks per GPU, no data-path collective" % NBLOCKS},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": int(n + m), "d2h_bytes_per_step": int(m + n), "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": e2e_wall_ms / e2e_steps, "api": "kz_compress_stream + kz_decompress_stream, pinned host buffers"},
        "gpu_launches": int(launches),
        "kernels": kern,
    }
    if roof:
        line["roofline"] = roof
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference(2, 1, sample_blocks=NBLOCKS, workload=args.workload)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker is absent: report, never substitute
            line["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": 0, "kind": "port", "sample": "unavailable: %s" % e}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
