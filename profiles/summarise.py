"""Turn gpurun_out/{launches.csv,prof_*.ncu-rep} into small text summaries under profiles/."""
import collections, csv, subprocess, sys, os
tag = sys.argv[1]
assert tag.startswith("r") and tag[1:].isdigit(), "usage: python profiles/summarise.py rNN"
out = open(f"profiles/{tag}_summary.md", "w")
def w(s=""):
    out.write(s + "\n")
w(f"# ncu summaries, {tag}")
w()
w("Command profiled: `python bench.py --steps 1 --warmup 1 --pairs 4 --passes 1 --no-cpu-baseline --no-extras` (cfg2: 1080p, fs 8x8 +-16, bidir),")
w("each ncu pass run only after the same command exited 0 without ncu (profiles/capture.sh).")
w()
w("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`) -- cold-cache, serialised: compare SHARES")
w()
lines = [l for l in open("gpurun_out/launches.csv") if l.startswith('"')]
open(f"profiles/{tag}_launches.csv", "w").writelines(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0].replace('void ', '')
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(u, 1)
    agg[name][0] += 1; agg[name][1] += v
step = {k: v for k, v in agg.items() if not k.startswith('ub_kernel')}
tot = sum(v[1] for v in step.values())
w("| kernel | launches | total us | avg us | share of hot-path kernels |")
w("|---|---:|---:|---:|---:|")
for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
    w(f"| {k} | {v[0]} | {v[1]/1e3:.1f} | {v[1]/v[0]/1e3:.1f} | {v[1]/tot:.3f} |")
w()
w("Note: the profiled command also runs bench.py's exact-kernel pass (`roofline_fs_exact` / `value_prefilter_off`: the same pairs with")
w("MEMCI_FS_PREFILTER=0), so `fs_tiled_kernel` appears with real launches (~270 us each) on top of the gated launches of")
w("the two-phase path, which return at once (~3 us: the kernel only runs when the survivor list overflowed), and `fs_redo_kernel`")
w("with the exact path's redo passes.  Split by duration below.")
w()
tiled = [float(r['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r['Metric Unit'], 1e-3)
         for r in csv.DictReader(lines) if 'fs_tiled_kernel' in r['Kernel Name']]
gated = [t for t in tiled if t < 30]
real = [t for t in tiled if t >= 30]
if tiled:
    w(f"`fs_tiled_kernel`: {len(gated)} gated launches, avg {sum(gated) / max(len(gated), 1):.1f} us; {len(real)} real launches, avg {sum(real) / max(len(real), 1):.1f} us.")
    n_pf = step.get([k for k in step if k.startswith('fs8_prefilter_kernel')][0], [0, 0])[0] if any(k.startswith('fs8_prefilter_kernel') for k in step) else 0
    two_phase = {k: v for k, v in step.items() if not k.startswith('fs_tiled_kernel')}
    tot2 = sum(v[1] / v[0] * min(v[0], n_pf) if k.startswith(('fs_redo', 'luma', 'mci')) else v[1] for k, v in two_phase.items()) / 1e3 + sum(gated)
    w()
    w(f"Two-phase step proper (per pair: one launch of each fs8_* kernel, the gated launch, one redo pass, one splat / hole fill):")
    w()
    w("| kernel | avg us | share of the two-phase pair |")
    w("|---|---:|---:|")
    per_pair = {k: v[1] / v[0] / 1e3 for k, v in two_phase.items()}
    per_pair['fs_tiled_kernel (gated)'] = sum(gated) / max(len(gated), 1)
    tp = sum(per_pair.values())
    for k, v in sorted(per_pair.items(), key=lambda kv: -kv[1]):
        w(f"| {k} | {v:.1f} | {v / tp:.3f} |")
    w(f"| sum | {tp:.1f} | 1 |")
w()
for k, v in agg.items():
    if k.startswith('ub_kernel'):
        w(f"| ({k}: roofline microbenchmark, not part of a step) | {v[0]} | {v[1]/1e3:.1f} | {v[1]/v[0]/1e3:.1f} | - |")
w()
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
for rep, title in (("prof_fs8", "fs8_prefilter_kernel (dominant kernel: 8-bit prefilter of the two-phase full search)"),
                   ("prof_scan", "fs8_scan_kernel (column minima against block thresholds, survivor list)"),
                   ("prof_refine", "fs8_refine_kernel (exact luma-SAD of the survivors)"),
                   ("prof_redo", "fs_redo_kernel (exact block search of the listed blocks)"),
                   ("prof_fs", "fs_tiled_kernel (the exact kernel, MEMCI_FS_PREFILTER=0)"), ("prof_mci", "mci_splat_kernel"), ("prof_hf", "mci_holefill_kernel")):
    path = f"gpurun_out/{rep}.ncu-rep"
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    w(f"## `ncu --set full --clock-control none --import-source on`: {title}")
    w()
    w("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
    w("|---|---|" + "---:|" * len(data))
    for m in ['Kernel Name'] + want:
        if m in hdr:
            i = hdr.index(m)
            w(f"| {m} | {units[i]} | " + " | ".join(r[i][:70] for r in data) + " |")
    w()
import json
traffic = {}
for rep, kname in (("prof_fs8", "fs8_prefilter_kernel"), ("prof_scan", "fs8_scan_kernel"), ("prof_refine", "fs8_refine_kernel"), ("prof_fs", "fs_tiled_kernel"),
                   ("prof_mci", "mci_splat_kernel"), ("prof_hf", "mci_holefill_kernel")):
    path = f"gpurun_out/{rep}.ncu-rep"
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    def col(m):
        i = hdr.index(m)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return [float(r[i]) * mult for r in data]
    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    traffic[kname] = {"dram_bytes_per_launch": sum(rd) / len(rd) + sum(wr) / len(wr), "launches_captured": len(rd),
                      "source": f"profiles/{tag}_summary.md (ncu --set full --clock-control none)"}
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
out.close()
print(open(f"profiles/{tag}_summary.md").read()[:6000])
