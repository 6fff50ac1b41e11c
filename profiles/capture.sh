set -x
CMD="python bench.py --steps 1 --warmup 1 --pairs 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fs_tiled -s 2 -c 2 -o gpurun_out/prof_fs $CMD > gpurun_out/ncu_fs.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mci_splat -s 1 -c 2 -o gpurun_out/prof_mci $CMD > gpurun_out/ncu_mci.log 2>&1
tail -3 gpurun_out/ncu_fs.log gpurun_out/ncu_mci.log
ls -la gpurun_out
