#!/bin/bash
# round-2: GPU suite on the merged tree, then A/B of cache-policy variants of the raster (B2D_LIB builds, B2D_L2PERSIST)
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_$name.json") if l.startswith("{")][-1]); r=d.get("roofline") or {}
    print("$name: %.0f fps  ms/step %.4f raster %s walk %s frac %s" % (d["value"], d["ms_per_step"], r.get("avg_launch_ms"), r.get("walk_avg_launch_ms"), r.get("frac")))
except Exception as e: print("$name: no result", e)
PY
tail -2 gpurun_out/bench_${TAG}_$name.err; }
Q="--no-e2e --no-cpu-baseline --steps 60 --warmup 3 --no-pipeline"
P="--no-e2e --no-cpu-baseline --steps 60 --warmup 3"
D=$PWD/rust-doom_b200
for rep in a b; do
  run base_$rep python bench.py $Q
  B2D_L2PERSIST=1 run persist_$rep python bench.py $Q
  for v in stcs l2el l2elcs el; do B2D_LIB=$D/libb2d_$v.so run ${v}_$rep python bench.py $Q; done
done
run pipe_base python bench.py $P
run pipe_base_1stream python bench.py $P --raster-streams 1
run pipe_base_b python bench.py $P
run pipe_base_1stream_b python bench.py $P --raster-streams 1
B2D_L2PERSIST=1 run pipe_persist python bench.py $P
for v in stcs l2el l2elcs; do B2D_LIB=$D/libb2d_$v.so run pipe_$v python bench.py $P; done
for v in base stcs l2elcs; do
  L=$D/libb2d_$v.so; [ $v = base ] && L=$D/libb2d.so
  B2D_LIB=$L run rich_$v python bench.py --config rich --steps 20 --warmup 3
  B2D_LIB=$L run 4k_$v python bench.py --config 4k --steps 20 --warmup 3
done
B2D_L2PERSIST=1 run rich_persist python bench.py --config rich --steps 20 --warmup 3
B2D_L2PERSIST=1 run 4k_persist python bench.py --config 4k --steps 20 --warmup 3
