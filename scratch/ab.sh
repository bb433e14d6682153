for v in mb3 mb4; do
  for mf in "" "--matrix-free"; do
    CTS_LIB_PATH=$PWD/scratch/variants/lib_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras $mf > /tmp/b.json 2>/dev/null
    python -c "
import json; d=json.load(open('/tmp/b.json')); print('$v mf=$mf', round(d['value']/1e9,2), 'G DOF-stage/s', round(d['ms_per_step'],3), 'ms; fused', round(d['kernels']['fused_stage']['ms_per_step'],3), 'roof', round(d['roofline']['frac'],3))"
  done
done
