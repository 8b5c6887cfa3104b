#!/bin/bash
# round-2 GPU check, single GPU: tests, headline bench (full), other configs, reference arm, ncu captures
TAG=${1:-r2b}; STEPS=${2:-50}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi_$TAG.txt
lscpu | grep -E "Model name|^CPU\(s\)|NUMA" > gpurun_out/cpu_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}_$name.json | head -c 1500; echo; tail -2 gpurun_out/bench_${TAG}_$name.err; }
run c2 python bench.py --steps $STEPS --warmup 3
run ref python bench.py --impl reference --steps 10 --warmup 2
run c3 python bench.py --config c3 --steps 5 --warmup 3
run 4k python bench.py --config 4k --steps 20 --warmup 3
run rich python bench.py --config rich --steps 20 --warmup 3
run c4n1 python bench.py --config c4 --poses 200 --steps 2 --warmup 3
run c5n1 python bench.py --config c5 --poses 4000 --chunk 250 --steps 2
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_raster -s 3 -c 1 -o gpurun_out/prof_raster_$TAG -f $B > gpurun_out/ncu_raster_$TAG.log 2>&1; echo "ncu raster rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_walk -s 3 -c 1 -o gpurun_out/prof_walk_$TAG -f $B > gpurun_out/ncu_walk_$TAG.log 2>&1; echo "ncu walk rc=$?"
