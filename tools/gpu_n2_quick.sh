#!/bin/bash
# 2-GPU session (short): multi-GPU pytest + the driver's bench command at N=2 (captured graph incl. exchange).
out=gpurun_out/${1:-n2}; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rf > $out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 50 --warmup 5 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY' $out
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        det = d.get("details", {})
        print(os.path.basename(f), "n_gpus", d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              str(det.get("launch", ""))[:40], det.get("graph_capture_error"), det.get("rank_gpu_ms_per_step"))
    except Exception as e:
        print(f, "bad", e, open(f.replace(".json", ".err")).read()[-800:])
PY
