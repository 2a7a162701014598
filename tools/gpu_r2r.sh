#!/bin/bash
tag=${1:-r2r}; out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 $out/pytest_gpu.log
for n in 300000 2000000; do
  timeout 300 python bench.py --gaussians $n --steps 30 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_q_$n.json 2> $out/bench_q_$n.err
  GGRT_RASTER_LIB=$PWD/gpurun_variants/sort64/libggrt_raster.so timeout 300 python bench.py --gaussians $n --steps 30 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_s64_$n.json 2> $out/bench_s64_$n.err
done
timeout 300 python bench.py --workload c3 --steps 30 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_q_c3.json 2> $out/bench_q_c3.err
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None)
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
