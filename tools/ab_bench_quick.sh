#!/bin/bash
# quick A/B of bench.py's device-resident leg under environment switches: tools/ab_bench_quick.sh "ENV1=.. ENV2=.." "ENV3=.." ...
O=gpurun_out/ab_bench_quick.log
: > $O
for cfg in "$@"; do
  echo "== $cfg" >> $O
  env $cfg python bench.py --no-cpu-baseline --no-attention-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step', round(d['ms_per_step'],4), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()}, 'e2e', round(d['e2e']['ms_per_step'],4), d.get('schedule',{}).get('one_launch_schedule_ms_per_step'), 'launches', d['gpu_launches'], 'roofline', round(d['roofline']['frac'],4), round(d['roofline']['share_of_step'],3))" >> $O
done
cat $O
