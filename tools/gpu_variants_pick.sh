#!/bin/bash
# One call: bench the default library and every variant twice, then run the GPU test suite on the best variant if it beats the default.
out=gpurun_out/${1:-last}; mkdir -p $out
B="--steps 60 --no-cpu-baseline --no-gpu-baseline --no-extras"
for rep in a b; do
  timeout 100 python bench.py $B > $out/bench_default_$rep.json 2> $out/bench_default_$rep.err
  for v in $(ls gpurun_variants); do
    GGRT_RASTER_LIB=$PWD/gpurun_variants/$v/libggrt_raster.so timeout 100 python bench.py $B > $out/bench_${v}_$rep.json 2> $out/bench_${v}_$rep.err
  done
done
best=$(python - <<'PY' $out
import json, sys, glob, os, collections
out = sys.argv[1]
ms = collections.defaultdict(list)
for f in sorted(glob.glob(out + "/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ms[os.path.basename(f)[6:-7]].append(d["ms_per_step"])
        print(os.path.basename(f), "%.4f" % d["ms_per_step"], {k: round(v * 1e3, 1) for k, v in d["roofline"]["stage_ms"].items() if k in ("geometry", "preprocess_backward")}, file=sys.stderr)
    except Exception as e:
        print(f, "bad", e, file=sys.stderr)
mean = {k: sum(v) / len(v) for k, v in ms.items() if v}
print({k: round(v, 5) for k, v in mean.items()}, file=sys.stderr)
cands = {k: v for k, v in mean.items() if k != "default"}
b = min(cands, key=cands.get) if cands else ""
print(b if b and cands[b] < mean.get("default", 0) - 0.0004 else "")
PY
)
echo "best variant: '$best'"
if [ -n "$best" ]; then
  GGRT_RASTER_LIB=$PWD/gpurun_variants/$best/libggrt_raster.so timeout 400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_graph.py::test_library_switches_keep_the_results > $out/pytest_$best.log 2>&1; echo "pytest($best) rc=$?"; tail -2 $out/pytest_$best.log
fi
