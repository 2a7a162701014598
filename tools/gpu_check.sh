#!/bin/bash
# GPU session: full parity suite + default bench (quick evidence run after a kernel change).  bash tools/gpu_check.sh [tag]
tag=${1:-check}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rf > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -12 $out/pytest_gpu.log
ts "grad diag"
timeout 300 python tools/grad_diag.py > $out/grad_diag.log 2>&1
python - <<'PY' $out
import json, sys
for l in open(sys.argv[1] + "/grad_diag.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["case"], {k: (round(v["max_rel"], 6), round(v["floor1e-05"]["worst_strict"], 3)) for k, v in d.items() if isinstance(v, dict)})
PY
ts "bench default"
timeout 600 python bench.py --steps 40 --no-cpu-baseline --no-extras > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?" >> $out/timeline.log
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  ts "variant $v: bench"
  GGRT_RASTER_LIB=$lib timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_$v.json 2> $out/bench_$v.err
done
ts done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None, (d.get("clocks") or {}).get("reasons"), d.get("gpu_baseline"))
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
