# Capture the ncu evidence for a round: launch list + one full capture each of the dominant kernel
# (fs8_prefilter_kernel; and fs_tiled_kernel, the exact kernel, with the prefilter switched off) and the HBM-side kernel (mci_splat_kernel).  Every ncu pass runs only after the same
# command exited 0 without ncu.  Usage (on the GPU box): bash profiles/capture.sh ; then, in the build
# container:  python profiles/summarise.py rNN
set -x
CMD="timeout 300 python bench.py --steps 1 --warmup 1 --pairs 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fs8_prefilter -s 1 -c 2 -o gpurun_out/prof_fs8 -f $CMD > gpurun_out/ncu_fs8.log 2>&1
MEMCI_FS_PREFILTER=0 $CMD > gpurun_out/plain2b.log 2>&1 && \
MEMCI_FS_PREFILTER=0 ncu --set full --clock-control none --import-source on -k regex:fs_tiled -s 1 -c 2 -o gpurun_out/prof_fs -f $CMD > gpurun_out/ncu_fs.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mci_splat -s 1 -c 2 -o gpurun_out/prof_mci -f $CMD > gpurun_out/ncu_mci.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mci_holefill -s 1 -c 1 -o gpurun_out/prof_hf -f $CMD > gpurun_out/ncu_hf.log 2>&1
tail -3 gpurun_out/ncu_fs8.log gpurun_out/ncu_fs.log gpurun_out/ncu_mci.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 400 gpurun_out/bench_default.json
