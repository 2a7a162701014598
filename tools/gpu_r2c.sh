#!/bin/bash
# Round-2 multi-GPU session (N = number of visible GPUs): exchange variants against the arena all-reduce, the
# multi-GPU pytest, bench at N (graph and eager).   bash tools/gpu_r2c.sh [tag] [N] [workload]
tag=${1:-r2c}
N=${2:-2}
wl=${3:-c2}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
run() { timeout -k 10 $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
ts start
nvidia-smi --query-gpu=index,name --format=csv > $out/gpu.txt 2>&1
ts "multi_gpu_check"
run 420 29511 tools/multi_gpu_check.py --workload $wl --iters 20 > $out/mg.jsonl 2> $out/mg.err; echo "mg rc=$?" | tee -a $out/timeline.log
ts "pytest multi"
timeout -k 10 400 python -m pytest tests/test_gpu_multi.py -q -rf > $out/pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a $out/timeline.log
ts "bench graph"
run 400 29512 bench.py --gpus $N --steps 50 --warmup 5 --workload $wl > $out/bench_graph.json 2> $out/bench_graph.err; echo "rc=$?" | tee -a $out/timeline.log
ts "bench eager"
run 400 29513 bench.py --gpus $N --steps 50 --warmup 5 --workload $wl --launch eager > $out/bench_eager.json 2> $out/bench_eager.err; echo "rc=$?" | tee -a $out/timeline.log
ts "bench reference arm"
run 300 29514 bench.py --gpus $N --steps 3 --warmup 1 --workload $wl --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
ts done
cat $out/mg.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items() if k not in ('trace', 'bytes', 'max_rel_err_vs_arena')})
    except Exception: print(l[:300])
"
tail -3 $out/pytest_multi.log
for f in $out/bench_*.json; do python - $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"], (d.get("details") or {}).get("rank_gpu_ms_per_step"), (d.get("details") or {}).get("launch", "")[:30], d.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "unreadable", e, open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
