# Final-build evidence: bench line (no profiler), then the ncu launch list of the same command, then one --set full
# capture of the MLP kernel.  Outputs under gpurun_out/; summaries are copied into profiles/ afterwards.
set -e
timeout 400 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 3 -c 1 -f -o gpurun_out/prof_tc_final python tools/profile_tc.py fp16 32 > gpurun_out/ncu_full_final.log 2>&1
grep -h '^{' gpurun_out/final_bench_default.json | cut -c1-200
