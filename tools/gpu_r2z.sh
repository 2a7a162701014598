#!/bin/bash
out=gpurun_out/${1:-r2z}; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_device_glue.py tests/test_gpu_golden.py -m gpu -q -rf > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY' $out
import json, sys
d = json.loads(open(sys.argv[1] + "/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "fps", d["value"])
tc = d.get("through_caller", {})
for k, v in tc.items():
    print(k, json.dumps(v)[:400])
PY
tail -5 $out/bench.err
