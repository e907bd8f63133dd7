#!/bin/bash
# A/B of the pipelined-tail variants (config 2, device-resident forward): tools/ab_pipe_sweep.sh
O=gpurun_out/ab_pipe_sweep.log
: > $O
run() { echo "== $*" >> $O; env "$@" python tools/ab_pipe_tail.py 30 ${AXIS:-literal_b1} 1 2>&1 | grep -v sorted >> $O; }
run AVS_PIPE_NO_MERGE=1
run AVS_X=1
run AVS_PIPE_NO_MERGE=1
run AVS_X=1
AXIS=temporal run AVS_PIPE_NO_MERGE=1
AXIS=temporal run AVS_X=1
echo "== trace" >> $O
AVS_PIPE_TRACE=1 python tools/ab_pipe_tail.py 2 literal_b1 1 2>&1 | tail -10 >> $O
python -m pytest tests -m gpu -x -q -k "pipelined_tail" 2>&1 | tail -2 >> $O
cat $O
