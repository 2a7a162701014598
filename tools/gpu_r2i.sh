#!/bin/bash
# Round-2 GPU session I (1 GPU): MMA-based render backward -- parity suite, bench A/B against variants, ncu of K7.
tag=${1:-r2i}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts start
[ -x gpurun_variants/ubench_pipes ] && timeout 120 gpurun_variants/ubench_pipes > $out/ubench_pipes.jsonl 2>&1
ts "pytest parity"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py tests/test_gpu_edge_cases.py tests/test_gpu_golden.py tests/test_gpu_device_glue.py -m gpu -q -x -rf > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -15 $out/pytest_gpu.log
ts "grad diag"
timeout 300 python tools/grad_diag.py > $out/grad_diag.log 2>&1; tail -12 $out/grad_diag.log
ts "bench default"
timeout 600 python bench.py --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_default.json 2> $out/bench_default.err; echo "rc=$?" >> $out/timeline.log
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  ts "variant $v: bench"
  GGRT_RASTER_LIB=$lib timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_$v.json 2> $out/bench_$v.err
done
ts "ncu full: K7"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_backward' -s 2 -c 1 -f -o $out/k7 python tools/profile_step.py c2 2 > $out/ncu_k7.log 2>&1
ncu -i $out/k7.ncu-rep --page raw --csv > $out/k7_raw.csv 2> $out/raw.err
ncu -i $out/k7.ncu-rep --page source --csv > $out/k7_source.csv 2>> $out/raw.err
rm -f $out/k7.ncu-rep
ts done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None, (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
cat $out/ubench_pipes.jsonl 2>/dev/null | tail -12
