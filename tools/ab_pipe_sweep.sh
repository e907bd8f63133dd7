#!/bin/bash
# A/B of the pipelined-tail variants (config 2, device-resident forward): tools/ab_pipe_sweep.sh
O=gpurun_out/ab_pipe_sweep.log
: > $O
run() { echo "== $*" >> $O; env "$@" python tools/ab_pipe_tail.py 30 ${AXIS:-literal_b1} 1 2>&1 | grep -v sorted >> $O; }
run AVS_X=1
run AVS_PIPE_EXCL2=1
run AVS_X=1
run AVS_PIPE_EXCL2=1
AXIS=temporal run AVS_X=1
AXIS=temporal run AVS_PIPE_EXCL2=1
echo "== trace EXCL2" >> $O
AVS_PIPE_EXCL2=1 AVS_PIPE_TRACE=1 python tools/ab_pipe_tail.py 3 literal_b1 1 2>&1 | tail -19 >> $O
AVS_PIPE_EXCL2=1 python -m pytest tests -m gpu -x -q -k "pipelined_tail" 2>&1 | tail -2 >> $O
cat $O
