# The measurement recipe of /opt/skills/guides/B200_PROFILING.md for this repo: plain run first, then the launch list of the same command,
# then ONE --set full capture of the fused-stage launches of a two-step program (scratch/ncu_kernels.py).  Outputs -> gpurun_out/$1_*
tag=${1:-r2b}
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
python scratch/ncu_kernels.py > /dev/null 2>&1 || { echo "ncu_kernels.py failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:column_fused_stage_reg_kernel -s 3 -c 3 -o /tmp/${tag}_fused python scratch/ncu_kernels.py > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_fused.ncu-rep --page raw --csv > gpurun_out/${tag}_fused_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_fused.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_fused_source.csv.gz
ls -la gpurun_out/${tag}_*
