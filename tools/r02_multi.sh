#!/bin/bash
# Multi-GPU measurements of one round (gpurun --gpus N):   bash tools/r02_multi.sh r02e 8 "train long infer"
P=${1:-r02e}
N=${2:-2}
WHAT=${3:-"train long infer"}
O=gpurun_out
mkdir -p $O
export NCCL_DEBUG=${NCCL_DEBUG:-VERSION}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for w in $WHAT; do
  case $w in
    2dev)  (timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "two_devices" 2>&1 | tail -3) > $O/${P}_pytest_2dev.log ;;
    train) timeout 240 $RUN bench.py --gpus $N --config train --steps 20 --warmup 5 > $O/${P}_train_n$N.json 2> $O/${P}_train_n$N.err ;;
    eager) timeout 240 $RUN bench.py --gpus $N --config train --steps 20 --warmup 5 --no-graph > $O/${P}_train_eager_n$N.json 2> $O/${P}_train_eager_n$N.err ;;
    long)  timeout 240 $RUN bench.py --gpus $N --config long --steps 5 --warmup 3 > $O/${P}_long_n$N.json 2> $O/${P}_long_n$N.err ;;
    infer) timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/${P}_infer_n$N.json 2> $O/${P}_infer_n$N.err ;;
  esac
done
for f in train_n$N train_eager_n$N long_n$N infer_n$N; do [ -f $O/${P}_$f.json ] && { echo == $f; cut -c1-240 $O/${P}_$f.json; }; done
grep -h -i "NCCL version\|Traceback\|Error" $O/${P}_*_n$N.err 2>/dev/null | sort | uniq -c | head -10
