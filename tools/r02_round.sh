#!/bin/bash
# One GPU call that refreshes a round-2 set of measured artefacts under gpurun_out/ (copied to profiles/ afterwards).
#   gpurun --timeout 1500 -- 'bash tools/r02_round.sh r02a'
P=${1:-r02a}
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15) > $O/${P}_pytest.log
timeout 300 python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
timeout 300 python bench.py --axis temporal --no-cpu-baseline --no-attention-probe > $O/${P}_bench_temporal.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_bench_ref.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --config long --steps 5 --warmup 3 --no-cpu-baseline > $O/${P}_bench_long.json 2>> $O/${P}_bench.err
timeout 300 python bench.py --config train --steps 20 --warmup 5 > $O/${P}_bench_train.json 2>> $O/${P}_bench.err
timeout 200 python tools/bench_config3.py 2>/dev/null | tail -2 > $O/${P}_config3.log
timeout 200 python tools/lstm_trace.py > $O/${P}_lstm_trace.log 2>&1
timeout 100 python tools/prof_knapsack_long.py 8 8192 5 > $O/${P}_knapsack_long.log 2>&1
(timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2) > $O/${P}_smoke.log
tail -3 $O/${P}_pytest.log; cut -c1-400 $O/${P}_bench.json; tail -5 $O/${P}_bench.err; cut -c1-300 $O/${P}_bench_long.json; cut -c1-300 $O/${P}_bench_train.json; cat $O/${P}_config3.log
