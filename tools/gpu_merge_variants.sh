#!/bin/bash
# tools/bench_merge.py for the default library and every gpurun_variants/*/libggrt_raster.so -> one JSON object
out=gpurun_out/${1:-merge}; mkdir -p $out
timeout 200 python tools/bench_merge.py > $out/default.json 2> $out/default.err
for v in $(ls gpurun_variants 2>/dev/null); do
  lib=$PWD/gpurun_variants/$v/libggrt_raster.so
  [ -f $lib ] || continue
  GGRT_RASTER_LIB=$lib timeout 200 python tools/bench_merge.py > $out/$v.json 2> $out/$v.err
done
python - <<'PY' $out
import json, sys, glob, os
res = {}
for f in sorted(glob.glob(sys.argv[1] + "/*.json")):
    try:
        res[os.path.basename(f)[:-5]] = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        res[os.path.basename(f)[:-5]] = {"error": open(f.replace(".json", ".err")).read()[-400:]}
for k, v in res.items():
    print(k, {a: (round(b * 1e3, 1) if isinstance(b, float) else {x: round(y * 1e3, 1) for x, y in b.items()}) for a, b in v.items()} if "error" not in v else v)
json.dump(res, open(sys.argv[1] + "/merge_variants.json", "w"), indent=1)
PY
