#!/bin/bash
# Round-2 GPU session G (1 GPU): parity suite incl. device glue, bench, merge-kernel variants.
tag=${1:-r2g}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts start
ts "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rf > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -8 $out/pytest_gpu.log
ts "bench default"
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?" >> $out/timeline.log
ts "merge default"
timeout 300 python tools/bench_merge.py > $out/merge_default.json 2> $out/merge_default.err
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  ts "variant $v: merge"
  GGRT_RASTER_LIB=$lib timeout 300 python tools/bench_merge.py > $out/merge_$v.json 2> $out/merge_$v.err
done
ts done
for f in $out/merge_*.json; do echo $f; cat $f; done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None)
        for k in ("depth_pass", "through_caller"):
            if k in d: print("   ", k, json.dumps(d[k])[:900])
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
