#!/bin/bash
# ncu launch list of the bench command (after the same command has exited 0 without ncu); numbers printed under ncu are
# never bench values.   gpurun --timeout 600 -- 'bash tools/r02_launches.sh r02w'
P=${1:-r02w}
O=gpurun_out
mkdir -p $O
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-attention-probe > $O/${P}_plain_bench.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-attention-probe > $O/${P}_ncu_bench.log 2>&1
tail -3 $O/${P}_ncu_bench.log | cut -c1-200; wc -l $O/${P}_launches.csv
