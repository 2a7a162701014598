#!/bin/bash
# Round-2 GPU session D (1 GPU): parity suite, bench, forward-staging A/B, ncu launch list + full capture of K7 / K6.
tag=${1:-r2d}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts start
ts "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rf > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -8 $out/pytest_gpu.log
ts "bench default (graph)"
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?" >> $out/timeline.log
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  ts "variant $v: bench"
  GGRT_RASTER_LIB=$lib timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_$v.json 2> $out/bench_$v.err
done
ts "exchange kernels (virtual views)"
timeout 300 python tools/bench_merge.py > $out/bench_merge.json 2> $out/bench_merge.err
ts "ncu full: merge kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sh_gradient_merge' -s 20 -c 1 -o $out/prof_merge python tools/bench_merge.py > $out/ncu_merge.log 2>&1
ts "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches.csv python bench.py --launch eager --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/ncu_bench.log 2>&1
ts "ncu full: render kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_(backward|forward)|preprocess_backward|emit|sort_tiles' -s 40 -c 5 -o $out/prof_render python bench.py --launch eager --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/ncu_render.log 2>&1
ts done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None, (d.get("clocks") or {}).get("reasons"), (d.get("details") or {}).get("launch", "")[:20])
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
