#!/bin/bash
# One GPU-box session: parity tests, bench A/B of the colour-kernel overlap and of library variants, ncu launch list.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh [tag]
tag=${1:-r1b}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts start
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt 2>&1
ts "pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -5 $out/pytest_gpu.log
ts "bench default"
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?" >> $out/timeline.log
ts "bench overlap off"
GGRT_RASTER_OVERLAP=0 timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-gpu-baseline > $out/bench_nooverlap.json 2> $out/bench_nooverlap.err
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  ts "variant $v: parity"
  GGRT_RASTER_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q > $out/pytest_$v.log 2>&1; echo "variant $v pytest rc=$?" | tee -a $out/timeline.log
  ts "variant $v: bench"
  GGRT_RASTER_LIB=$lib timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-gpu-baseline > $out/bench_$v.json 2> $out/bench_$v.err
done
ts "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $out/ncu_bench.log 2>&1
ts done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()}, d["clocks"])
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e)
PY
