# Round-2 evidence: bench line (no profiler), then the ncu launch list of the same program, then --set full captures of
# the MLP kernel at the benchmarked launch size.  Outputs under gpurun_out/ (reports are exported to CSV on the box and
# the second report deleted: gpurun copies back at most 64 MiB); summaries are copied into profiles/ afterwards.
#   bash tools/refresh_profiles.sh [bench]     ("bench" also re-runs the plain bench lines first)
set -x
if [ "${1:-}" = bench ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
  timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
fi
timeout 300 python bench.py --steps 1 --warmup 1 --views 256 --no-cpu-baseline > gpurun_out/r2_bench_256views.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_bench_256views.csv python bench.py --steps 1 --warmup 1 --views 256 --no-cpu-baseline > gpurun_out/r2_ncu_launch.log 2>&1
timeout 200 python tools/profile_room.py fp16 3 > gpurun_out/r2_profile_room_plain.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -c 2 -f -o gpurun_out/r2_prof_tc_fp16 python tools/profile_room.py fp16 > gpurun_out/r2_ncu_full_fp16.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:mlp_tc -s 1 -c 1 -f -o gpurun_out/r2_prof_tc_fp16x3 python tools/profile_room.py fp16x3 > gpurun_out/r2_ncu_full_fp16x3.log 2>&1
for p in fp16 fp16x3; do
  ncu -i gpurun_out/r2_prof_tc_$p.ncu-rep --page raw --csv > gpurun_out/r2_tc_${p}_raw.csv 2> gpurun_out/r2_tc_${p}_raw.err
done
ncu -i gpurun_out/r2_prof_tc_fp16.ncu-rep --page source --csv > gpurun_out/r2_tc_fp16_source.csv 2> gpurun_out/r2_tc_fp16_source.err
rm -f gpurun_out/r2_prof_tc_fp16x3.ncu-rep
ls -la gpurun_out/
