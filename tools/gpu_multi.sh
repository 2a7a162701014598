#!/bin/bash
# Multi-GPU session: exchange check (all variants + phase times) and bench.py at N ranks.  bash tools/gpu_multi.sh N [tag]
N=${1:-2}
tag=${2:-mg$N}
out=gpurun_out/$tag
mkdir -p $out
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tools/multi_gpu_check.py --workload c2 --iters 20 > $out/check_c2.log 2>&1; echo "check rc=$?"
grep -E '^\{' $out/check_c2.log | cut -c1-900
for ex in auto compact arena; do
  run 29512 bench.py --gpus $N --steps 40 --warmup 5 --exchange $ex > $out/bench_$ex.json 2> $out/bench_$ex.err; echo "bench $ex rc=$?"
done
python - <<'PY' $out
import json, sys, glob, os
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(os.path.basename(f), "n_gpus", d["n_gpus"], "ms/step %.4f" % d["ms_per_step"], "fps %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"], d["config"]["parallelism"][:80])
    except Exception as e:
        print(f, "bad", e, open(f.replace(".json", ".err")).read()[-500:])
PY
