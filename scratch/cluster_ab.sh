# A/B of the cluster (loader CTA + solver CTA) fused stage against the default kernel
run() {
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > /tmp/b.json 2>/tmp/b.err || { echo "FAILED $1"; tail -5 /tmp/b.err; return; }
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$1', round(d['ms_per_step'],3), 'ms; fused', round(d['kernels']['fused_stage']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"
}
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "pipelined_fused_stage_kernel_is_bit_identical and float64" 2>&1 | tail -15
for i in 1 2; do
run default
CTS_FUSED_CLUSTER=1 run cluster
done
CTS_FUSED_CLUSTER=1 CTS_FUSED_TIMELINE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep -A3 "cluster timeline" | tail -16
