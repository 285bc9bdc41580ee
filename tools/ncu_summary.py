"""Extracts the metrics the design notes cite from an .ncu-rep (ncu --set full) into a small JSON for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/out.json"""
import csv
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        k = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
        for w in WANT:
            if w in idx:
                k[w] = {"value": num(r[idx[w]]), "unit": units[idx[w]]}
        stalls = {}
        for h, i in idx.items():
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
                v = num(r[i])
                if isinstance(v, float) and v >= 0.05:
                    stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(v, 3)
        k["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        rd, wr = k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            k["dram_traffic_bytes"] = rd["value"] * scale.get(rd["unit"], 1) + wr["value"] * scale.get(wr["unit"], 1)
        res.append(k)
    if len(sys.argv) > 3 and sys.argv[3] == "--longest":  # one entry per kernel name: its longest launch (+ launch count)
        best = {}
        for k in res:
            t = k.get("gpu__time_duration.sum", {}).get("value", 0)
            name = k["kernel"]
            if name not in best or t > best[name]["gpu__time_duration.sum"]["value"]:
                cnt = best[name]["launches_captured"] + 1 if name in best else 1
                best[name] = k
                best[name]["launches_captured"] = cnt
            else:
                best[name]["launches_captured"] += 1
        res = list(best.values())
    json.dump({"source": rep, "kernels": res}, open(out, "w"), indent=1)
    print("wrote", out, len(res), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
