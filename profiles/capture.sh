# Capture the ncu evidence for a round: launch list + one full capture each of the dominant kernel (fs8_prefilter_kernel), the
# kernels that follow it in the two-phase full search (fs8_scan_kernel, fs8_refine_kernel, fs_redo_kernel), the exact kernel
# (fs_tiled_kernel, prefilter switched off) and the MCI kernels (mci_splat_kernel, mci_holefill_kernel).  Every ncu pass runs
# only after the same command exited 0 without ncu.  Usage (on the GPU box): bash profiles/capture.sh ; then, in the build
# container:  python profiles/summarise.py rNN
set -x
mkdir -p gpurun_out
CMD="timeout 300 python bench.py --steps 1 --warmup 1 --pairs 4 --passes 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
cap() {   # name, kernel regex, launches, extra env
  env $4 $CMD > gpurun_out/plain_$1.log 2>&1 && \
  env $4 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c $3 -o gpurun_out/prof_$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -n 1 gpurun_out/ncu_$1.log
}
cap fs8 fs8_prefilter 2 X=1
cap scan fs8_scan 1 X=1
cap refine fs8_refine 1 X=1
cap redo fs_redo 1 X=1
cap fs fs_tiled 2 MEMCI_FS_PREFILTER=0
cap mci mci_splat 2 X=1
cap hf mci_holefill 1 X=1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for wl in cfg3 cfg4 cfg5; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
tail -c 300 gpurun_out/bench_cfg2.json
