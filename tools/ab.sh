#!/bin/bash
# A/B of engine options on one workload: tools/ab.sh <workload> <steps> "opt1=v ..." "opt1=v ..."
wl=$1; steps=$2; shift 2
for o in "$@"; do
  args=""; for kv in $o; do args="$args --opt $kv"; done
  timeout 400 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu --e2e-steps 1 $args 2>gpurun_out/ab.err | python -c "
import sys,json
t=sys.stdin.read()
try:
    d=json.loads(t)
    print('$o', '%.1fM/s'%(d['value']/1e6), '%.2f ms/step'%d['ms_per_step'], 'passes', d['passes_per_step'], d['move_pass'], {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
except Exception as ex:
    print('$o', 'FAILED', ex, t[-300:]); print(open('gpurun_out/ab.err').read()[-600:])
"
done
