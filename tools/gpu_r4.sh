#!/bin/bash
# GPU session: parity suite, default bench, the same library without programmatic dependent launch, and every variant library.
tag=${1:-r4}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -4 $out/pytest_gpu.log
B="--steps 60 --no-cpu-baseline --no-gpu-baseline --no-extras"
ts "bench default"
timeout 300 python bench.py $B > $out/bench_default.json 2> $out/bench_default.err
ts "bench PDL off"
GGRT_RASTER_PDL=0 timeout 300 python bench.py $B > $out/bench_nopdl.json 2> $out/bench_nopdl.err
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  ts "variant $v"
  GGRT_RASTER_LIB=$lib timeout 300 python bench.py $B > $out/bench_$v.json 2> $out/bench_$v.err
done
ts done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], d["details"].get("graph_capture_error"),
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None, (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
