#!/bin/bash
out=gpurun_out/${1:-r4d}; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_device_glue.py -m gpu -q -rf -x > $out/pytest_glue.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_glue.log
timeout 600 python bench.py --steps 60 --no-cpu-baseline --no-gpu-baseline > $out/bench_extras.json 2> $out/bench_extras.err; echo "bench rc=$?"
python - <<'PY' $out
import json, sys
d = json.loads(open(sys.argv[1] + "/bench_extras.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
tc = d.get("through_caller", {})
print({k: (v.get("ms_per_step") if isinstance(v, dict) else v) for k, v in tc.items() if k != "multi_view_device_glue"})
print(json.dumps(tc.get("multi_view_device_glue"), indent=0))
print(d.get("depth_pass"))
PY
bash tools/gpu_variants_quick.sh ${1:-r4d}_var
