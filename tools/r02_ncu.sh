#!/bin/bash
# ncu captures of one round (run AFTER the plain commands have exited 0 in r02_round.sh); numbers printed under ncu
# are never bench values.   gpurun --timeout 1200 -- 'bash tools/r02_ncu.sh r02a'
P=${1:-r02a}
O=gpurun_out
mkdir -p $O
timeout 200 python tools/prof_long.py 8 8192 2 > $O/${P}_plain_long.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -f -o $O/${P}_prof_attn_T8192 -k regex:attention_tc -s 1 -c 1 python tools/prof_long.py 8 8192 2 > $O/${P}_ncu_attn.log 2>&1
timeout 100 python tools/prof_step.py 3 > $O/${P}_plain_step.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-attention-probe > $O/${P}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -f -o $O/${P}_prof_step -k regex:"lstm_tc|gemm_tc|knapsack|shot_pool" -s 20 -c 10 python tools/prof_step.py 3 > $O/${P}_ncu_full.log 2>&1
timeout 100 python tools/prof_train.py 3 > $O/${P}_plain_train.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -f -o $O/${P}_prof_bptt -k regex:lstm_backward_tc -s 1 -c 1 python tools/prof_train.py 3 > $O/${P}_ncu_bptt.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${P}_train_launches.csv python tools/prof_train.py 4 > $O/${P}_ncu_train.log 2>&1
ls -la $O | tail -12
