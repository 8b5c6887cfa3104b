#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_benchmark and not campaign" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_$name.json") if l.startswith("{")][-1]); r=d.get("roofline") or {}
    print("$name: %.0f fps  ms/step %.4f raster %s walk %s frac %s" % (d["value"], d["ms_per_step"], r.get("avg_launch_ms"), r.get("walk_avg_launch_ms"), r.get("frac")))
except Exception as e: print("$name: no result", e)
PY
tail -2 gpurun_out/bench_${TAG}_$name.err; }
Q="--no-e2e --no-cpu-baseline --steps 60 --warmup 3"
P=$PWD/rust-doom_b200/libb2d_prev.so
run c2 python bench.py $Q
run c2seq python bench.py $Q --no-pipeline
B2D_LIB=$P run c2seq_prev python bench.py $Q --no-pipeline
run c2seq_b python bench.py $Q --no-pipeline
B2D_LIB=$P run c2seq_prev_b python bench.py $Q --no-pipeline
run 4k python bench.py --config 4k --steps 20 --warmup 3
run rich python bench.py --config rich --steps 20 --warmup 3
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_raster -s 3 -c 1 -o gpurun_out/prof_raster_$TAG -f $B > gpurun_out/ncu_raster_$TAG.log 2>&1; echo "ncu raster rc=$?"
