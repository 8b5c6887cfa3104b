#!/bin/bash
# A/B of cache-policy variants of the raster (B2D_LIB builds): sequential step, 60 steps each, twice
TAG=${1:-r2f}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${TAG}_$name.json") if l.startswith("{")][-1]); r=d.get("roofline") or {}
    print("$name: %.0f fps  ms/step %.4f raster %s walk %s frac %s" % (d["value"], d["ms_per_step"], r.get("avg_launch_ms"), r.get("walk_avg_launch_ms"), r.get("frac")))
except Exception as e: print("$name: no result", e)
PY
tail -2 gpurun_out/bench_${TAG}_$name.err; }
Q="--no-e2e --no-cpu-baseline --steps 60 --warmup 3 --no-pipeline"
D=$PWD/rust-doom_b200
for rep in a b; do
  run base_$rep python bench.py $Q
  for v in stcs stwt el elcs; do B2D_LIB=$D/libb2d_$v.so run ${v}_$rep python bench.py $Q; done
done
B2D_LIB=$D/libb2d_stcs.so run stcs_rich python bench.py --config rich --steps 20 --warmup 3
run base_rich python bench.py --config rich --steps 20 --warmup 3
B2D_LIB=$D/libb2d_stcs.so run stcs_4k python bench.py --config 4k --steps 20 --warmup 3
run base_4k python bench.py --config 4k --steps 20 --warmup 3
