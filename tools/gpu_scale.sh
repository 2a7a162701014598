#!/bin/bash
# Scaling session on an N-GPU box (charged N x): the driver's bench command at 1, 2, 4, .. N GPUs, graph and eager launch at N,
# the reference arm under torchrun, BASELINE config 4 at N.    bash tools/gpu_scale.sh N [tag]
N=${1:-8}; tag=${2:-scale$N}; out=gpurun_out/$tag; mkdir -p $out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-extras --no-gpu-baseline --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
port=29600
for n in 2 4 8; do
  [ $n -le $N ] || continue
  port=$((port+1)); run $n $port bench.py --gpus $n --steps 50 --warmup 5 > $out/bench_n$n.json 2> $out/bench_n$n.err; echo "bench n=$n rc=$?"
done
port=$((port+1)); run $N $port bench.py --gpus $N --steps 50 --warmup 5 --launch graph > $out/bench_n${N}_graph.json 2> $out/bench_n${N}_graph.err
port=$((port+1)); run $N $port bench.py --gpus $N --steps 50 --warmup 5 --workload c4 > $out/bench_c4_n$N.json 2> $out/bench_c4_n$N.err
port=$((port+1)); run $N $port bench.py --gpus $N --impl reference --steps 5 --warmup 1 > $out/bench_reference_n$N.json 2> $out/bench_reference_n$N.err
port=$((port+1)); run $N $port tools/multi_gpu_check.py --workload c2 --iters 15 > $out/check_c2.log 2>&1
grep -E '^\{' $out/check_c2.log | cut -c1-200 > $out/check_c2.jsonl
python - <<'PY' $out
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        det = d.get("details", {})
        print(os.path.basename(f), "n_gpus", d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"],
              "cores", (d.get("cpu_baseline") or {}).get("cores"), str(det.get("launch", ""))[:28], det.get("rank_gpu_ms_per_step"))
    except Exception as e:
        print(f, "bad", e, open(f.replace(".json", ".err")).read()[-600:])
PY
cat $out/check_c2.jsonl | cut -c1-120
