#!/bin/bash
# Multi-GPU measurements of one round (gpurun --gpus N):   bash tools/r02_multi.sh r02b 2
P=${1:-r02b}
N=${2:-2}
O=gpurun_out
mkdir -p $O
export NCCL_DEBUG=${NCCL_DEBUG:-VERSION}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  (timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "two_devices" 2>&1 | tail -3) > $O/${P}_pytest_2dev.log
fi
timeout 600 $RUN bench.py --gpus $N --config train --steps 20 --warmup 5 > $O/${P}_train_n$N.json 2> $O/${P}_train_n$N.err
timeout 600 $RUN bench.py --gpus $N --config train --steps 20 --warmup 5 --no-graph > $O/${P}_train_eager_n$N.json 2>> $O/${P}_train_n$N.err
timeout 600 $RUN bench.py --gpus $N --config long --steps 5 --warmup 3 > $O/${P}_long_n$N.json 2> $O/${P}_long_n$N.err
timeout 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/${P}_infer_n$N.json 2> $O/${P}_infer_n$N.err
for f in train_n$N train_eager_n$N long_n$N infer_n$N; do echo == $f; cut -c1-260 $O/${P}_$f.json; done
grep -h -i "nranks\|NCCL version\|error\|Traceback" $O/${P}_*_n$N.err | sort | uniq -c | head -20
cat $O/${P}_pytest_2dev.log 2>/dev/null
