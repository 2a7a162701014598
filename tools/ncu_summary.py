"""Summarises an `ncu --set full` report of one fwd+bwd step (+ the launch list of a bench run) into
profiles/<round>_ncu_kernels.json.

  ncu -i step.ncu-rep --page raw --csv > raw.csv          (on the GPU box or here)
  python tools/ncu_summary.py raw.csv launches.csv out.json "<capture command>" "<launch-list command>"
"""
import csv
import json
import re
import sys
from collections import defaultdict

WANT = {
    "gpu__time_duration.sum": "ncu_time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "sm__inst_executed.sum": "inst_executed",
    "smsp__inst_executed.sum": "inst_executed",
    "launch__grid_size": "grid",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic_KB",
    "launch__shared_mem_per_block_static": "smem_static_KB",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0,
              "Gbyte": 1e3}


def short(name):
    name = re.sub(r"\(.*$", "", name).strip()
    name = re.sub(r"^void\s+", "", name)
    return re.sub(r"^ggrt::", "", name)


def read_raw(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    kcol = hdr.index("Kernel Name")
    out = {}
    for r in rows[hdr_i + 2:]:
        if len(r) != len(hdr):
            continue
        name = short(r[kcol])
        d = out.setdefault(name, {})
        for c, (h, u) in enumerate(zip(hdr, units)):
            if h not in WANT:
                continue
            try:
                v = float(r[c].replace(",", ""))
            except ValueError:
                continue
            key = WANT[h]
            if key == "ncu_time_us":
                v *= UNIT_SCALE.get(u, 1.0)
            elif key.endswith("_MB"):
                v *= UNIT_SCALE.get(u, 1e-6)
            elif key.endswith("_KB"):
                v *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u, 1e-3)
            d.setdefault(key, v)
    for d in out.values():
        d["dram_traffic_bytes"] = int(round((d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0)) * 1e6))
    return out


def read_launches(path):
    rows = list(csv.reader(l for l in open(path, newline="") if not l.startswith("==")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    k, m, v, u = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = defaultdict(list)
    for r in rows[hdr_i + 1:]:
        if len(r) == len(hdr) and r[m] == "gpu__time_duration.sum":
            acc[short(r[k])].append(float(r[v].replace(",", "")) * UNIT_SCALE.get(r[u], 1e-3))
    return acc


if __name__ == "__main__":
    raw, launches, out = sys.argv[1:4]
    src = sys.argv[4] if len(sys.argv) > 4 else ""
    lsrc = sys.argv[5] if len(sys.argv) > 5 else ""
    workload = sys.argv[6] if len(sys.argv) > 6 else "C2: P=300000, 1008x756, SH degree 4, N=800306 pairs"
    kernels = read_raw(raw)
    la = read_launches(launches)
    ours = {k: v for k, v in la.items() if "_kernel" in k and not k.startswith(("void at::", "at::"))}
    total = sum(sum(v) for v in ours.values())
    doc = {
        "source": src, "launch_list": lsrc, "workload": workload,  # bench.py matches the workload prefix ("C2")
        "kernels": kernels,
        "launch_list_mean_us_per_launch": {k: round(sum(v) / len(v), 2) for k, v in ours.items()},
        "launch_list_launches": {k: len(v) for k, v in ours.items()},
        "launch_list_share_of_step": {k: round(sum(v) / total, 4) for k, v in ours.items()},
    }
    json.dump(doc, open(out, "w"), indent=1)
    for k, d in kernels.items():
        print(f"{k:45s} {d.get('ncu_time_us', 0):8.2f} us  dram {d.get('dram_read_MB', 0) + d.get('dram_write_MB', 0):8.2f} MB")
