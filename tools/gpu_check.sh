#!/bin/bash
# One gpurun call = tests + bench + ncu captures.  Usage: tools/gpu_check.sh <tag> [steps]
TAG=${1:-x}; STEPS=${2:-100}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_raster -s 3 -c 1 -o gpurun_out/prof_raster_$TAG -f $B > gpurun_out/ncu_raster_$TAG.log 2>&1; echo "ncu rc=$?"
