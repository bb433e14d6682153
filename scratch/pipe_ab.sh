#!/bin/bash
# A/B runs of the pipelined fused-stage kernel's ring depths (env overrides read by pipe_config in csrc/ops.cu)
run() {
  env $1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras $2 > /tmp/b.json 2>/tmp/b.err || { echo "$1 $2 FAILED"; tail -3 /tmp/b.err; return; }
  python -c "
import json; d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1]); print('$1 $2', round(d['ms_per_step'],3), 'ms; fused', round(d['kernels']['fused_stage']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"
}
for cfg in "$@"; do
  run "$cfg" ""
done
