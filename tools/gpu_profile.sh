#!/bin/bash
# Evidence run for profiles/: full bench line, reference arm, ncu launch list of the bench command, ncu --set full of one step,
# C3 (+pose gradients), C5 sweep.  bash tools/gpu_profile.sh [tag]
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
ts() { echo "[$(date +%H:%M:%S)] $*" | tee -a $out/timeline.log; }
ts bench; timeout 600 python bench.py > $out/bench_c2.json 2> $out/bench_c2.err
ts reference-arm; timeout 600 python bench.py --impl reference --steps 20 --warmup 1 > $out/bench_c2_reference_arm.json 2> $out/ref.err
ts launch-list
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $out/ncu_bench.log 2>&1
ts ncu-full
timeout 900 ncu --set full --clock-control none --import-source on -k regex:_kernel -s 16 -c 8 -f -o $out/step python tools/profile_step.py c2 2 > $out/ncu_full.log 2>&1
ncu -i $out/step.ncu-rep --page raw --csv > $out/raw.csv 2> $out/raw.err
ncu -i $out/step.ncu-rep --page details --csv > $out/details.csv 2>> $out/raw.err
ls -la $out/step.ncu-rep | tee -a $out/timeline.log
rm -f $out/step.ncu-rep
ts c3; timeout 300 python bench.py --workload c3 --steps 40 --no-cpu-baseline --no-gpu-baseline > $out/bench_c3.json 2> $out/bench_c3.err
timeout 300 python bench.py --workload c3 --steps 40 --no-cpu-baseline --no-gpu-baseline --pose-grads > $out/bench_c3_pose.json 2> $out/bench_c3_pose.err
ts c4; timeout 300 python bench.py --workload c4 --steps 40 --no-cpu-baseline --no-gpu-baseline > $out/bench_c4_1gpu.json 2> $out/bench_c4.err
ts sweep
for n in 50000 100000 300000 600000 1000000 2000000; do
  timeout 300 python bench.py --gaussians $n --steps 30 --no-cpu-baseline --no-gpu-baseline > $out/sweep_$n.json 2> $out/sweep_$n.err
done
ts done
