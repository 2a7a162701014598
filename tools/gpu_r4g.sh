#!/bin/bash
out=gpurun_out/${1:-r4g}; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -rf > $out/pytest_graph.log 2>&1; echo "pytest graph rc=$?"; tail -5 $out/pytest_graph.log
bash tools/gpu_sanitize.sh ${1:-r4g}_san
timeout 300 python bench.py --workload c3 --steps 40 --no-cpu-baseline --no-gpu-baseline --no-extras --pose-grads > $out/bench_c3_pose.json 2> $out/bench_c3_pose.err
python - <<'PY' $out
import json, sys
d = json.loads(open(sys.argv[1] + "/bench_c3_pose.json").read().strip().splitlines()[-1])
print("c3 pose ms/step", d["ms_per_step"], "fps", d["value"])
PY
bash tools/gpu_final.sh ${1:-r4g}_final
