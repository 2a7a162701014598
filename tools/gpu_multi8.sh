#!/bin/bash
# 8-GPU session (expensive: charged 8x): exchange check at C2, bench at C2 and at BASELINE config 4 (600K Gaussians, 8 views).
N=${1:-8}; out=gpurun_out/mg$N; mkdir -p $out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tools/multi_gpu_check.py --workload c2 --iters 15 > $out/check_c2.log 2>&1; echo "check rc=$?"
grep -E '^\{' $out/check_c2.log | cut -c1-160
grep -E '^\{' $out/check_c2.log | grep -o '"phase_ms_rank0.*'
run 29512 bench.py --gpus $N --steps 30 --warmup 5 > $out/bench_c2_auto.json 2> $out/bench_c2_auto.err; echo "bench c2 rc=$?"
run 29513 bench.py --gpus $N --steps 30 --warmup 5 --workload c4 > $out/bench_c4_auto.json 2> $out/bench_c4_auto.err; echo "bench c4 rc=$?"
run 29514 bench.py --gpus $N --steps 30 --warmup 5 --workload c4 --exchange arena > $out/bench_c4_arena.json 2> $out/bench_c4_arena.err; echo "bench c4 arena rc=$?"
python - <<'PY' $out
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(os.path.basename(f), "n_gpus", d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"], d["config"]["parallelism"][:70])
    except Exception as e:
        print(f, "bad", e, open(f.replace(".json", ".err")).read()[-600:])
PY
