"""The one JSON line of bench.py, checked on the arm that runs without a GPU (--impl reference: the CPU restatement on the host cores) and a
2 MB workload: every key of the contract is there, exactly one line is written to stdout, and the workload description (`config`) is what
the GPU arm builds from the same function."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "tiny"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches", "schedule"):
        assert key in d, key
    assert d["unit"] == "MB/s" and d["higher_is_better"] is True and d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["levels"]) == {"l3", "l5"} and all(v["compressed_bytes"] > 0 for v in d["levels"].values())
    # the same description object as the GPU arm's (bench.static_config is the only builder of `config`)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": static_config(args.workload, passes, BLOCK, NBLOCKS, n, world)') == 2  # reference arm and b200 arm
    assert "workload" in d["config"] and d["config"]["block_size"] == 256 << 10 and d["config"]["uncompressed_bytes"] == 2_000_000
