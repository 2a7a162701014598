#!/bin/bash
# Round-end validation exactly as the driver does it: smoke(), the GPU test suite, the default bench command and the reference arm.
out=gpurun_out/${1:-final}; mkdir -p $out
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_gpu.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<'PY' $out
import json, sys
for f in ("bench_reference.json", "bench.json"):
    d = json.loads(open(sys.argv[1] + "/" + f).read().strip().splitlines()[-1])
    print(f, {k: d.get(k) for k in ("impl", "value", "ms_per_step", "n_gpus", "gpu_launches")}, "e2e", d["e2e"]["value"],
          "roofline", {k: (d.get("roofline") or {}).get(k) for k in ("kernel", "frac", "traffic")}, "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
PY
