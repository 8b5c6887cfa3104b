#!/bin/bash
# multi-GPU check: usage tools/gpu_multi.sh <tag> <N> <what...>   what: c5small c5 c4 c4small c2 cli
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}_$name.json; echo; grep -v "^W0\|^\*\*\*\|OMP_NUM_THREADS" gpurun_out/bench_${TAG}_$name.err | tail -5; }
for w in "$@"; do
  case $w in
    c5small) NCCL_DEBUG=WARN run c5small $TR bench.py --gpus $N --config c5 --poses 8000 --chunk 250 --steps 1 --transports window,plain,ce ;;
    c5) run c5 $TR bench.py --gpus $N --config c5 --chunk 256 --steps 2 --transports window,plain,ce ;;
    c5chan) NCCL_MIN_NCHANNELS=32 run c5chan $TR bench.py --gpus $N --config c5 --chunk 256 --steps 2 --transports window,plain ;;
    c5big) run c5big $TR bench.py --gpus $N --config c5 --chunk 512 --steps 2 ;;
    c5noreg) B2D_NCCL_NO_REGISTER=1 run c5noreg $TR bench.py --gpus $N --config c5 --chunk 256 --steps 2 ;;
    c5nowin) B2D_NCCL_NO_WINDOW=1 run c5nowin $TR bench.py --gpus $N --config c5 --chunk 256 --steps 2 ;;
    c4small) run c4small $TR bench.py --gpus $N --config c4 --poses 200 --steps 2 --warmup 3 ;;
    c4) run c4 $TR bench.py --gpus $N --config c4 --steps 3 --warmup 3 ;;
    c2) run c2 $TR bench.py --gpus $N --steps 30 --warmup 3 ;;
    d2h) timeout 300 $TR tools/d2h_ceiling.py > gpurun_out/d2h_${TAG}.json 2> gpurun_out/d2h_${TAG}.err; cat gpurun_out/d2h_${TAG}.json
         timeout 300 $TR tools/d2h_ceiling.py --no-numa > gpurun_out/d2h_${TAG}_nonuma.json 2>> gpurun_out/d2h_${TAG}.err; cat gpurun_out/d2h_${TAG}_nonuma.json ;;
    c5ce) run c5ce $TR bench.py --gpus $N --config c5 --chunk 256 --steps 1 --transports ce ;;
    c2q) run c2q $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline ;;
    d2hb) timeout 300 $TR tools/d2h_ceiling.py > gpurun_out/d2h_${TAG}.json 2> gpurun_out/d2h_${TAG}.err; cat gpurun_out/d2h_${TAG}.json ;;
    c2e2e) run c2e2e $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline ;;
    ref) run ref $TR bench.py --gpus $N --impl reference --steps 5 --warmup 2 ;;
    test2) timeout 900 python -m pytest tests/test_cli.py -m gpu -q -k "two_ranks or world1" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_${TAG}.log ;;
    cli)
      python - <<'PY'
from rust_doom_b200 import synthwad
open("/tmp/t.wad","wb").write(synthwad.build_iwad(1,("E1M1",)))
PY
      rm -f /tmp/b2d_id
      for r in $(seq 0 $((N-1))); do rust-doom_b200/b2d --iwad /tmp/t.wad --resolution 640x400 --poses 64 --world $N --rank $r --chunk 5 --id-file /tmp/b2d_id > gpurun_out/cli_${TAG}_$r.txt 2>&1 & done
      wait; cat gpurun_out/cli_${TAG}_*.txt ;;
  esac
done
