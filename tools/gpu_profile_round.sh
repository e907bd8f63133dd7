# One GPU call that refreshes the round's measured artefacts under gpurun_out/ (copied to profiles/ afterwards).
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile_round.sh r01c'
P=${1:-r01c}
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
python bench.py --steps 20 --warmup 5 --axis temporal --no-cpu-baseline > gpurun_out/${P}_bench_temporal.json 2> gpurun_out/${P}_bench_temporal.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err
python tools/bench_config1.py > gpurun_out/${P}_config1.log 2>&1
python tools/bench_config3.py > gpurun_out/${P}_config3.log 2>&1
python tools/bench_long.py 8 8192 > gpurun_out/${P}_long.log 2>&1
python tools/bench_train.py --cpu-baseline > gpurun_out/${P}_train.log 2>&1
python tools/lstm_trace.py > gpurun_out/${P}_lstm_trace.log 2>&1
python tools/gemm_trace.py > gpurun_out/${P}_gemm_trace.log 2>&1
python tools/e2e_trace.py 20 > gpurun_out/${P}_e2e_trace.log 2>&1
python tools/e2e_stream.py 30 > gpurun_out/${P}_e2e_stream.log 2>&1
python tools/prof_step.py 3 > gpurun_out/plain_step.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${P}_launches.csv python tools/prof_step.py 3 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|lstm_tc_kernel|knapsack|shot_pool|convert" -s 40 -c 14 -o gpurun_out/${P}_prof_step python tools/prof_step.py 3 > gpurun_out/ncu_step.log 2>&1
python tools/prof_step.py 3 temporal > gpurun_out/plain_step_t.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o gpurun_out/${P}_prof_attn python tools/prof_step.py 3 temporal > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/${P}_long.log gpurun_out/${P}_train.log
ls -la gpurun_out | tail -15
