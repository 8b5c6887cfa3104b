#!/bin/bash
# One gpurun call of round 2: GPU tests + A/B bench of the raster variants (B2D_TUNE) + ncu captures.
# usage: tools/gpu_r2.sh <tag> [steps] [tunes...]
TAG=${1:-x}; STEPS=${2:-60}; shift 2; TUNES=${@:-0}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
for T in $TUNES; do
  B2D_TUNE=$T timeout 300 python bench.py --steps $STEPS --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_${TAG}_t$T.json 2> gpurun_out/bench_${TAG}_t$T.err
  echo "tune $T rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_t$T.json") if l.startswith("{")][-1]); r=d["roofline"]
    print("tune $T: %.0f fps  raster %.4f ms walk %.4f ms  frac %.4f" % (d["value"], r["avg_launch_ms"], r["walk_avg_launch_ms"], r["frac"]))
except Exception as e: print("tune $T: no result", e)
PY
done
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_$TAG.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:b2d_raster -s 3 -c 1 -o gpurun_out/prof_raster_$TAG -f $B > gpurun_out/ncu_raster_$TAG.log 2>&1; echo "ncu rc=$?"
