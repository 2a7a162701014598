#!/bin/bash
# Evidence session (1 GPU): parity suite, bench lines (C2 with baselines, reference arm, C3, C3+pose, C4, C5 sweep), ncu launch
# list + full capture of one step (all 9 kernels) with source pages, adapter bench + ncu.   bash tools/gpu_evidence.sh [tag]
tag=${1:-ev}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts start
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt 2>&1
ts "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -rf > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
tail -4 $out/pytest_gpu.log
ts "bench default"
timeout 600 python bench.py > $out/bench_c2.json 2> $out/bench_c2.err; echo "rc=$?" >> $out/timeline.log
ts "bench eager"
timeout 300 python bench.py --launch eager --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_c2_eager.json 2> $out/bench_c2_eager.err
ts reference-arm
timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $out/bench_c2_reference_arm.json 2> $out/ref.err
ts "ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches.csv python bench.py --launch eager --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/ncu_bench.log 2>&1
ts "ncu full: one step"
K='regex:geometry_kernel|scan_tiles|color_kernel|emit_kernel|sort_tiles|render_forward|render_backward|preprocess_backward|sh_gradient_kernel'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 18 -c 9 -f -o $out/step python tools/profile_step.py c2 2 > $out/ncu_full.log 2>&1
ncu -i $out/step.ncu-rep --page raw --csv > $out/raw.csv 2> $out/raw.err
ncu -i $out/step.ncu-rep --page source --csv -k regex:render_backward > $out/source_render_backward.csv 2>> $out/raw.err
ncu -i $out/step.ncu-rep --page source --csv -k regex:render_forward > $out/source_render_forward.csv 2>> $out/raw.err
ls -la $out/step.ncu-rep | tee -a $out/timeline.log
rm -f $out/step.ncu-rep
ts c3; timeout 300 python bench.py --workload c3 --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_c3.json 2> $out/bench_c3.err
timeout 300 python bench.py --workload c3 --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras --pose-grads > $out/bench_c3_pose.json 2> $out/bench_c3_pose.err
ts c4; timeout 300 python bench.py --workload c4 --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/bench_c4_1gpu.json 2> $out/bench_c4.err
ts sweep
for n in 50000 100000 300000 600000 1000000 2000000; do
  timeout 300 python bench.py --gaussians $n --steps 30 --no-cpu-baseline --no-gpu-baseline --no-extras > $out/sweep_$n.json 2> $out/sweep_$n.err
done
if [ "$2" != "noadapter" ]; then
ts adapter
timeout 300 python tools/bench_adapter.py > $out/bench_adapter.json 2> $out/bench_adapter.err
timeout 600 ncu --set full --clock-control none -k regex:adapter_ -s 4 -c 2 -f -o $out/adapter python tools/bench_adapter.py > $out/ncu_adapter.log 2>&1
ncu -i $out/adapter.ncu-rep --page raw --csv > $out/adapter_raw.csv 2>> $out/raw.err
rm -f $out/adapter.ncu-rep
fi
ts done
python - <<'PY' $out
import json, sys, glob, os
out = sys.argv[1]
for f in sorted(glob.glob(out + "/bench_*.json")) + sorted(glob.glob(out + "/sweep_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "ms_per_step" not in d:
            print(os.path.basename(f), json.dumps(d)[:400]); continue
        r = d.get("roofline")
        print(os.path.basename(f), "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              {k: round(v * 1e3, 1) for k, v in r["stage_ms"].items()} if r else None, (d.get("clocks") or {}).get("reasons"))
        for k in ("depth_pass", "through_caller"):
            if k in d: print("   ", k, json.dumps(d[k])[:600])
    except Exception as e:
        print(os.path.basename(f), "unreadable:", e, open(f.replace('.json', '.err')).read()[-600:])
PY
