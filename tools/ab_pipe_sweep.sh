#!/bin/bash
# sweep of the pipelined-tail placement knobs (config 2, device-resident forward)
O=gpurun_out/ab_pipe_sweep.log
: > $O
run() { echo "== $*" >> $O; env "$@" python tools/ab_pipe_tail.py 30 literal_b1 1 2>&1 | grep -v sorted >> $O; }
run AVS_PIPE_TAIL=0
run AVS_X=1
run AVS_PIPE_EXCL=2
run AVS_PIPE_EXCL=2 AVS_PIPE_STAGGER_EACH=1
run AVS_PIPE_LONG=1
run AVS_PIPE_LONG=2
run AVS_PIPE_LONG=3
run AVS_PIPE_LONG=3 AVS_PIPE_STAGGER_EACH=1
run AVS_PIPE_STAGGER_EACH=1
run AVS_PIPE_EXCL=0 AVS_PIPE_LONG=4
echo "== trace EXCL=2 STAGGER_EACH" >> $O
AVS_PIPE_TRACE=1 AVS_PIPE_EXCL=2 AVS_PIPE_STAGGER_EACH=1 python tools/ab_pipe_tail.py 2 literal_b1 1 2>&1 | tail -12 >> $O
echo "== trace LONG=3" >> $O
AVS_PIPE_TRACE=1 AVS_PIPE_LONG=3 python tools/ab_pipe_tail.py 2 literal_b1 1 2>&1 | tail -12 >> $O
cat $O
