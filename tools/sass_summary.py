"""Per-kernel SASS mnemonic counts of libggrt_raster.so (cuobjdump -sass): the evidence for which hardware paths each
kernel uses (UBLKCP = TMA bulk copy, LDGSTS = cp.async, HMMA = mma.sync tensor pipe, FFMA2/FMUL2/FADD2 = packed fp32,
REDG/ATOMG = global reductions / atomics, SYNCS = mbarrier, PREEXIT / ACQBULK = programmatic dependent launch
(griddepcontrol.launch_dependents / .wait), MULTIMEM via *.MMEM*).  python tools/sass_summary.py [out.json]"""
import collections
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "ggrt_official_b200" / "lib" / "libggrt_raster.so"
KEY = ("UBLKCP", "UTMA", "PREEXIT", "ACQBULK", "LDGSTS", "HMMA", "FFMA2", "FMUL2", "FADD2", "REDG", "RED", "ATOMG", "ATOM", "SYNCS", "SHFL", "MUFU", "VOTE",
       "VIMNMX", "LDS", "STS", "LDG", "STG", "BAR", "MATCH", "REDUX")


def main(out=None):
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("ggrt::", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            cur[op.split(".")[0]] += 1
            if "MMEM" in op or "MULTIMEM" in op:
                cur["multimem"] += 1
    doc = {k: {"total": v["total"], **{m: v[m] for m in KEY + ("multimem",) if v[m]}} for k, v in sorted(kernels.items())}
    if out:
        json.dump({"library": "ggrt_official_b200/lib/libggrt_raster.so (nvcc -gencode arch=compute_100a,code=sm_100a)",
                   "how": "cuobjdump -sass, static instruction counts per kernel", "kernels": doc}, open(out, "w"), indent=1)
    for k, v in doc.items():
        print(f"{k[:60]:60s}", {a: b for a, b in v.items() if a in ("total", "UBLKCP", "LDGSTS", "HMMA", "FFMA2", "REDG", "SYNCS", "multimem")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
