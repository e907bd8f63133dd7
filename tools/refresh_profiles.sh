#!/bin/bash
# Regenerates everything under profiles/ for one round (run under gpurun from the repo root; ~7 GPU-minutes).
# Afterwards, here: python tools/ncu_summary.py profiles/r01d_ncu_full_summary.csv gpurun_out/r01d_prof_step.ncu-rep gpurun_out/r01d_prof_attn.ncu-rep
P=gpurun_out
mkdir -p $P
(python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $P/r01d_pytest.log
python bench.py > $P/r01d_bench.json 2> $P/r01d_bench.err
python bench.py --axis temporal --no-cpu-baseline > $P/r01d_bench_temporal.json 2>> $P/r01d_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $P/r01d_bench_ref.json 2>> $P/r01d_bench.err
python tools/bench_config1.py 2>/dev/null | tail -1 > $P/r01d_config1.log
python tools/bench_config3.py 2>/dev/null | tail -2 > $P/r01d_config3.log
python tools/bench_long.py 2>/dev/null | tail -2 > $P/r01d_long.log
python tools/bench_train.py 2>/dev/null | tail -1 > $P/r01d_train.log
python tools/lstm_trace.py > $P/r01d_lstm_trace.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $P/r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $P/r01d_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -f -o $P/r01d_prof_step -k regex:"lstm_tc_kernel|gemm_tc|knapsack|shot_pool" -c 12 python tools/prof_step.py 1 > $P/r01d_ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -f -o $P/r01d_prof_attn -k regex:attention_tc -c 1 python tools/prof_step.py 1 temporal >> $P/r01d_ncu_full.log 2>&1
cat $P/r01d_pytest.log; cut -c1-200 $P/r01d_bench.json; cat $P/r01d_long.log $P/r01d_train.log $P/r01d_config1.log $P/r01d_config3.log; ls -la $P
