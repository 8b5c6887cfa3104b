#!/bin/bash
# round-2 final single-GPU check: full GPU suite, campaign, all configs, sanitizer, ncu captures
TAG=${1:-r2e}; STEPS=${2:-100}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi_$TAG.txt
lscpu | grep -E "Model name|^CPU\(s\)|NUMA" > gpurun_out/cpu_$TAG.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python tools/campaign_gpu.py 300 > gpurun_out/campaign_$TAG.log 2>&1; echo "campaign rc=$?"; tail -2 gpurun_out/campaign_$TAG.log
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_$name.json") if l.startswith("{")][-1]); r=d.get("roofline") or {}; e=d.get("e2e") or {}; c=d.get("cpu_baseline") or {}
    print("$name: %.0f fps  ms/step %.4f raster %s walk %s frac %s e2e %s cpu %s/%s" % (d["value"], d["ms_per_step"], r.get("avg_launch_ms"), r.get("walk_avg_launch_ms"), r.get("frac"), e.get("value"), c.get("value"), c.get("cores")))
except Exception as e: print("$name: no result", e)
PY
tail -2 gpurun_out/bench_${TAG}_$name.err; }
run c2 python bench.py --steps $STEPS --warmup 3
run c2seq python bench.py --steps $STEPS --warmup 3 --no-pipeline --no-e2e --no-cpu-baseline
run ref python bench.py --impl reference --steps 10 --warmup 2
run c3 python bench.py --config c3 --steps 5 --warmup 3
run 4k python bench.py --config 4k --steps 20 --warmup 3
run rich python bench.py --config rich --steps 20 --warmup 3
run c4n1 python bench.py --config c4 --poses 400 --steps 2 --warmup 3
run c5n1 python bench.py --config c5 --poses 4000 --chunk 250 --steps 2
run rgba python bench.py --steps 20 --warmup 3 --rgba --no-e2e --no-cpu-baseline
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_$TAG.log
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-pipeline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_${TAG}_pipe.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_raster -s 3 -c 1 -o gpurun_out/prof_raster_$TAG -f $B > gpurun_out/ncu_raster_$TAG.log 2>&1; echo "ncu raster rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_walk -s 3 -c 1 -o gpurun_out/prof_walk_$TAG -f $B > gpurun_out/ncu_walk_$TAG.log 2>&1; echo "ncu walk rc=$?"
