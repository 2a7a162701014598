#!/bin/bash
# ncu --set full of the two render kernels of one C2 step (+ source pages).  bash tools/gpu_ncu_render.sh [tag]
tag=${1:-ncu}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'render_(backward|forward)' -s 4 -c 2 -f -o $out/render python tools/profile_step.py c2 2 > $out/ncu.log 2>&1
ncu -i $out/render.ncu-rep --page raw --csv > $out/raw.csv 2> $out/raw.err
ncu -i $out/render.ncu-rep --page source --csv -k regex:render_backward > $out/source_render_backward.csv 2>> $out/raw.err
ncu -i $out/render.ncu-rep --page source --csv -k regex:render_forward > $out/source_render_forward.csv 2>> $out/raw.err
rm -f $out/render.ncu-rep
tail -3 $out/ncu.log
