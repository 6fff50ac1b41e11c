# Re-capture after a change to the full-search prefilter only: launch list, one full capture of fs8_prefilter_kernel, the
# default bench line.  The other kernels' captures (profiles/capture.sh) stay valid.  Usage as capture.sh.
set -x
CMD="timeout 300 python bench.py --steps 1 --warmup 1 --pairs 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fs8_prefilter -s 1 -c 2 -o gpurun_out/prof_fs8 -f $CMD > gpurun_out/ncu_fs8.log 2>&1
tail -3 gpurun_out/ncu_fs8.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 300 gpurun_out/bench_default.json
