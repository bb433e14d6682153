run() {
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --rows $2 > /tmp/b.json 2>/tmp/b.err || { echo "FAILED $1"; tail -5 /tmp/b.err; return; }
  python -c "
import json; d=json.load(open('/tmp/b.json')); k=d['kernels']; print('$1 rows=$2', round(d['ms_per_step'],4), 'ms; fused', round(k['fused_stage']['ms_per_step'],4), 'texp', round(k['T_exp']['ms_per_step'],4), 'final', round(k['final_update']['ms_per_step'],4), 'clk', d['clocks']['sm_mhz'])"
}
for rows in 128 256 512; do
for i in 1 2; do
CTS_LIB_PATH=$PWD/scratch/variants/lib_rows8.so run rows8 $rows
run adaptive $rows
done
done
