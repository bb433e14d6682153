run() {
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras $2 > /tmp/b.json 2>/tmp/b.err || { tail -3 /tmp/b.err; return; }
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$1 $2', round(d['value']/1e9,2), 'G DOF-stage/s', round(d['ms_per_step'],3), 'ms; fused', round(d['kernels']['fused_stage']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3))"
}
for i in 1 2; do
run default ""
CTS_LIB_PATH=$PWD/scratch/variants/lib_nopark.so run nopark ""
done
