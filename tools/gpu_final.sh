#!/bin/bash
# Round-end validation on one GPU: full GPU test suite, smoke(), the default bench line (incl. CPU and GPU baselines),
# the adapter bench.  bash tools/gpu_final.sh [tag]
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $out/pytest_gpu.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $out/smoke.log)"
timeout 500 python bench.py > $out/bench_c2.json 2> $out/bench_c2.err; echo "bench rc=$?"
timeout 120 python tools/bench_adapter.py > $out/bench_adapter.json 2> $out/bench_adapter.err; echo "adapter rc=$?"
tail -c 2500 $out/bench_c2.json; echo; tail -1 $out/bench_adapter.json
