#!/bin/bash
# round-2: map jobs with alternating raster streams, flat-batch / occupancy variants, GPU campaign with moved sectors
TAG=${1:-r2h}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_$name.json") if l.startswith("{")][-1]); r=d.get("roofline") or {}
    print("$name: %.0f fps  ms/step %.4f raster %s walk %s frac %s alone %s" % (d["value"], d["ms_per_step"], r.get("avg_launch_ms"), r.get("walk_avg_launch_ms"), r.get("frac"), r.get("alone_frac")))
except Exception as e: print("$name: no result", e)
PY
tail -2 gpurun_out/bench_${TAG}_$name.err; }
P="--no-e2e --no-cpu-baseline --steps 60 --warmup 3"
D=$PWD/rust-doom_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python tools/campaign_gpu.py 240 > gpurun_out/campaign_$TAG.log 2>&1; echo "campaign rc=$?"; tail -2 gpurun_out/campaign_$TAG.log
for rep in a b; do
  run c2_$rep python bench.py $P
  for v in fb16 w2b20 l2el; do B2D_LIB=$D/libb2d_$v.so run c2_${v}_$rep python bench.py $P; done
done
Q="--no-e2e --no-cpu-baseline --steps 60 --warmup 3 --no-pipeline"
for rep in a b; do
  run seq_$rep python bench.py $Q
  B2D_LIB=$D/libb2d_nobulk.so run seq_nobulk_$rep python bench.py $Q
done
B2D_LIB=$D/libb2d_nobulk.so run c2_nobulk python bench.py $P
for c in 0 12 25; do B2D_CARVEOUT=$c run c2_carve$c python bench.py $P; B2D_CARVEOUT=$c run seq_carve$c python bench.py $Q; done
run c3 python bench.py --config c3 --steps 5 --warmup 3
run c3_1stream python bench.py --config c3 --steps 5 --warmup 3 --raster-streams 1
run 4k python bench.py --config 4k --steps 20 --warmup 3
run 4k_1stream python bench.py --config 4k --steps 20 --warmup 3 --raster-streams 1
run rich python bench.py --config rich --steps 20 --warmup 3
run rich_1stream python bench.py --config rich --steps 20 --warmup 3 --raster-streams 1
run c4n1 python bench.py --config c4 --steps 2 --warmup 3
for v in fb16 w2b20 l2el; do
  B2D_LIB=$D/libb2d_$v.so run 4k_$v python bench.py --config 4k --steps 20 --warmup 3
  B2D_LIB=$D/libb2d_$v.so run rich_$v python bench.py --config rich --steps 20 --warmup 3
done
run c2_full python bench.py --steps 100 --warmup 3
