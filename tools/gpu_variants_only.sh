#!/bin/bash
# Bench the default library and every gpurun_variants/*/libggrt_raster.so, two short runs each (run-to-run noise is ~1.5 us).
tag=${1:-var}; shift
out=gpurun_out/$tag
mkdir -p $out
B="--steps 60 --no-cpu-baseline --no-gpu-baseline --no-extras"
for rep in a b; do
  timeout 300 python bench.py $B "$@" > $out/bench_default_$rep.json 2> $out/bench_default_$rep.err
  for v in $(ls gpurun_variants 2>/dev/null); do
    lib=$PWD/gpurun_variants/$v/libggrt_raster.so
    [ -f $lib ] || continue
    GGRT_RASTER_LIB=$lib timeout 300 python bench.py $B "$@" > $out/bench_${v}_$rep.json 2> $out/bench_${v}_$rep.err
  done
done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], d["details"].get("graph_capture_error"),
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items() if k.startswith("prep")} if r else None, (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
