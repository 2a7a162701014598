#!/bin/bash
# A/B of library variants under gpurun_variants/: parity subset + short bench for each.  bash tools/gpu_variants.sh [tag] [pytest-targets]
tag=${1:-var}
targets=${2:-tests/test_gpu_parity.py}
out=gpurun_out/$tag
mkdir -p $out
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  GGRT_RASTER_LIB=$lib timeout 600 python -m pytest $targets -x -q > $out/pytest_$v.log 2>&1; echo "variant $v pytest rc=$? $(tail -1 $out/pytest_$v.log)"
  GGRT_RASTER_LIB=$lib timeout 300 python bench.py --steps 60 --no-cpu-baseline --no-gpu-baseline $BENCH_ARGS > $out/bench_$v.json 2> $out/bench_$v.err
done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json','.err')).read()[-300:])
PY
