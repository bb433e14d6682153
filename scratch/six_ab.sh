run() {
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > /tmp/b.json 2>/tmp/b.err || { echo "FAILED $1"; tail -5 /tmp/b.err; return; }
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$1', round(d['ms_per_step'],3), 'ms; fused', round(d['kernels']['fused_stage']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])"
}
for i in 1 2; do
run default
for v in "$@"; do CTS_LIB_PATH=$PWD/scratch/variants/lib_$v.so run $v; done
done
