# NOTE: compute-sanitizer is refused on the shared B200 pool (gpurun answers rc 86: 'closed on this pool'), so this script has
# only ever run its first line -- the plain pass over every kernel family -- there.  Kept for a box that allows the tools.
# compute-sanitizer over CI-sized invocations of every kernel family (scratch/sanitize_target.py); logs -> gpurun_out/sanitizer_*.log
S=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
python scratch/sanitize_target.py > gpurun_out/sanitizer_plain.log 2>&1 || { echo "target fails without the sanitizer"; tail -20 gpurun_out/sanitizer_plain.log; exit 1; }
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  [ $tool = initcheck ] && extra="--track-unused-memory no"
  timeout 200 $S --tool $tool $extra --print-limit 30 --log-file gpurun_out/sanitizer_$tool.log python scratch/sanitize_target.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
  tail -2 gpurun_out/sanitizer_$tool.out
done
# the opt-in variants of the fused stage use mbarrier / named-barrier / DSMEM hand-overs: racecheck + synccheck them too
for v in CTS_FUSED_PIPE CTS_FUSED_SP CTS_FUSED_CLUSTER; do
  for tool in racecheck synccheck; do
    env $v=1 timeout 90 $S --tool $tool --print-limit 30 --log-file gpurun_out/sanitizer_${tool}_$v.log python scratch/sanitize_target.py "ars343 fused" > gpurun_out/sanitizer_${tool}_$v.out 2>&1
    echo "== $v $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$v.log | tail -1)"
  done
done
