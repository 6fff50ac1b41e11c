#!/usr/bin/env python
"""bench.py -- interpolated frames/s of the MEMCI hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

Workload (BASELINE.json configs[1], "cfg2"): 1080p, full search 8x8 blocks +-16, 30 -> 60 fps, bidirectional MCI,
synthetic frames (interpolatory_b200.synth.make_sequence).  ONE clip of N * P + 1 source frames is shared by all
ranks and partitioned by frame-pair index (multigpu.rank_pairs: rank r takes the contiguous run of P pairs
[r P, (r + 1) P), no collective on the data path); weak scaling: the clip grows with N.  A STEP = `passes` passes over
the rank's P pairs: per pair 2 full-search motion fields (forward + backward) and one interpolated frame; `passes`
is chosen so that the K timed steps take >= 2 s (run.passes_per_step).  The pass-through source frames of the
60 fps output involve no computation and are not counted in `value` (interpolated frames/s);
`output_frames_per_s` adds them back.

value    : inputs resident in HBM before the timed region, outputs left in HBM; the step's pairs are spread over
           C contexts (= CUDA streams, --contexts) so that one pair's small kernels overlap another pair's search;
           timed with CUDA events on a timing stream that fans out to / in from the C streams; max over ranks.
e2e      : the same pairs through the public pipeline API (pipeline.PairPipeline, the engine under
           BaseInterpolator.interpolate_video and multigpu.MultiGpuPipeline) with page-locked HOST buffers: the H2D
           copy of every source frame and the D2H copy of every interpolated frame are inside the timed region.
           `e2e.host_link` is the duplex host<->device copy rate measured in this run on the same buffers (all
           ranks at once); `e2e.host_link_frac` = bytes e2e moves per second / that ceiling.
           `e2e.class_api` is the drop-in call itself, MEMCI(...).interpolate_video(), on the same clip.
roofline : the workload's dominant kernel, from per-launch CUDA events of a single-stream pass of the same step
           (`single_stream`: a launch timed next to another stream's persistent grid would include queueing):
             cfg2 / cfg4 / cfg5  fs8_prefilter_kernel, SAD pixel-ops per launch (exact candidate count of the
                                 reference's own bounds x block pixels x 2 directions) against the VABSDIFF4.U8.ACC
                                 issue rate measured by a microbenchmark in this run: integer-ALU bound, not HBM
                                 bound; `roofline_fs_exact` = fs_tiled_kernel (prefilter off) against
                                 the 32-bit VABSDIFF rate
             cfg3                the HBMA search kernels (coarse search + densifications) against the 32-bit
                                 VABSDIFF rate; `roofline_hbm_small` = pyramid + smoothing kernels against HBM
           `roofline_mci` = mci_splat_kernel against the measured HBM copy bandwidth, everywhere.
natural  : (N = 1) the reference's own 12 Middlebury pairs tiled to the workload's frame size, device resident:
           interpolated frames/s with the prefilter chosen by the library's content feedback, pinned on, and off
           (`value_prefilter_off` is also given for the synthetic clip: the floor when nothing can be pruned).
cpu_baseline / --impl reference: the C restatement of the reference (oracle/, validated bit-exact
           against the reference) on the host cores, OpenMP over block rows, one full-size pair per sample.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, settings, fps_in, fps_out, default pairs per pass)
    "cfg2": (1080, 1920, dict(me_mode='fs', block_size='8', target_region='16', mci_mode='bidir'), 30, 60, 16),
    "cfg1": (360, 640, dict(me_mode='fs', block_size='8', target_region='7', mci_mode='bidir'), 24, 60, 64),
    "cfg3": (1080, 1920, dict(me_mode='hbma', block_size='8', target_region='7', mci_mode='bidir',
                              filter_mode='median', filter_size='4'), 24, 120, 16),
    "cfg4": (2160, 3840, dict(me_mode='fs', block_size='16', target_region='32', mci_mode='bidir'), 30, 60, 4),
    "cfg5": (4320, 7680, dict(me_mode='fs', block_size='16', target_region='32', mci_mode='bidir'), 30, 60, 1),
    # not a BASELINE.json config: the third mci_mode on the cfg2 / cfg3 motion estimators
    "cfg2_bidir2": (1080, 1920, dict(me_mode='fs', block_size='8', target_region='16', mci_mode='bidir2'), 30, 60, 16),
    "cfg3_bidir2": (1080, 1920, dict(me_mode='hbma', block_size='8', target_region='7', mci_mode='bidir2'), 24, 120, 16),
}
WORKLOAD_TEXT = {
    "cfg2": "cfg2: 1080p full-search 8x8 blocks +-16, 30->60 fps, bidir MCI, synthetic frames",
    "cfg1": "cfg1: 640x360 full-search 8x8 +-7, 24->60 fps, bidir MCI, synthetic frames",
    "cfg3": "cfg3: 1080p HBMA (auto 4-level pyramid via class rule) + median MV smoothing, 24->120 fps, bidir MCI",
    "cfg4": "cfg4: 4K full-search 16x16 +-32, 30->60 fps, bidir MCI, synthetic frames",
    "cfg2_bidir2": "cfg2 motion estimation (1080p full-search 8x8 +-16, 30->60 fps) with bidir2 MCI (IEWMC + EAC + BDHI)",
    "cfg3_bidir2": "cfg3 motion estimation (1080p HBMA, auto 4-level pyramid, block-level vectors) with bidir2 MCI, 24->120 fps",
    "cfg5": "cfg5: 8K (7680x4320) single frame pair, full-search 16x16 +-32, bidir MCI, split into one horizontal band per GPU "
            "with frame-row + MV-row halo exchange (NCCL send/recv)",
}
MIN_TIMED_S = 2.2                   # the K timed steps together take at least this long (clock samples mean something)


def workload_config(name, world, pairs, passes=None, extra=None):
    """`config` of the JSON line: identical keys in both arms (the driver compares them)."""
    H, W, st, fps_in, fps_out, _ = WORKLOADS[name]
    frame_bytes = H * W * 3
    return {"workload": WORKLOAD_TEXT[name], "H": H, "W": W, "settings": st, "fps_in": fps_in, "fps_out": fps_out,
            "pairs_per_pass_per_gpu": pairs,
            "sharding": "one shared clip, contiguous runs of frame pairs per rank (multigpu.rank_pairs), no collective",
            "l2_policy": f"inputs larger than L2: {((pairs + 1) * frame_bytes + 3 * H * W * 4) / 1e6:.0f} MB of frames + luma planes "
                         "streamed per pass vs 126 MB L2"}


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions (NVML in a thread, every ~5 ms;
    falls back to `nvidia-smi -lms` when pynvml is unavailable)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.gpu, self.sm, self.pw, self.mask, self.max_sm = gpu_index, [], [], 0, None
        self.stop_flag, self.thread, self.proc, self.rows = False, None, None, []

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                        self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            try:
                q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                              "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=lambda: [self.rows.append([x.strip() for x in l.split(",")]) for l in self.proc.stdout],
                                               daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            for r in self.rows:
                if len(r) >= 7 and r[0].replace('.', '').isdigit():
                    self.sm.append(float(r[0])); self.max_sm = float(r[1]); self.pw.append(float(r[2]))
                    for i, n in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                        if r[3 + i].lower().startswith("active"):
                            self.mask |= self.REASONS[n]
        elif self.thread:
            self.thread.join(timeout=1)
        reasons = [n for n, bit in self.REASONS.items() if self.mask & bit]
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max(self.pw) if self.pw else None, "samples": len(self.sm), "reasons": reasons,
                "window": "device-resident timed region + e2e timed region"}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPU cores NVML reports as local to its GPU, BEFORE any page-locked buffer is allocated:
    the buffers then live on that NUMA node and the H2D / D2H DMA does not cross the socket interconnect (the e2e
    leg at 8 ranks is host-bound).  Best effort: silently keeps the inherited affinity if NVML has no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            bind_to_gpu_numa_node.original = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def shared_clip_segment(H, W, P, rank):
    """Source frames [rank * P, (rank + 1) * P] of the ONE synthetic clip every rank works on (seed 100): the frames
    rank `rank` needs for its run of P pairs.  make_sequence draws its noise frame after frame from one generator, so a
    rank generates the clip up to its last frame and keeps its own slice."""
    from interpolatory_b200.synth import make_sequence
    return np.ascontiguousarray(make_sequence(H, W, (rank + 1) * P + 1, seed=100)[rank * P:])


# ------------------------------------------------------------------------------------------
def cpu_reference_sample(H, W, st, dists, frames, threads=None):
    """One full-size pair through the CPU restatement of the reference; returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would pin the port to one thread)
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))
    a, b = frames[0], frames[1]
    bs, R = int(st['block_size']), int(st['target_region'])
    t0 = time.perf_counter()
    if st['me_mode'] == 'fs':
        fwd = O.fs(bs, R, a, b)
        bwd = O.fs(bs, R, b, a)
    else:
        steps = O.hbma_auto_steps(H, W)
        fwd = O.hbma(bs, R, 1, steps, 4, a, b)
        bwd = O.hbma(bs, R, 1, steps, 4, b, a)
    mode = st.get('filter_mode', 'none')
    fwd = O.smooth(mode, fwd, int(st.get('filter_size', 4)))
    bwd = O.smooth(mode, bwd, int(st.get('filter_size', 4)))
    for d in dists:
        O.bidir_helper(a, b, fwd, bwd, np.float32(d))
    return time.perf_counter() - t0, O.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from interpolatory_b200.synth import make_sequence
    from interpolatory_b200.pipeline import interpolation_dists
    H, W, st, fps_in, fps_out, P_default = WORKLOADS[args.workload]
    dists = interpolation_dists(fps_out / fps_in)
    frames = make_sequence(H, W, 2, seed=100)                   # the first pair of the shared clip
    for _ in range(args.warmup):
        cpu_reference_sample(H, W, st, dists, frames)
    t = 0.0
    cores = 1
    for _ in range(args.steps):
        dt, cores = cpu_reference_sample(H, W, st, dists, frames)
        t += dt
    value = args.steps * len(dists) / t
    sample = (f"{args.steps} x 1 full-size frame pair of the shared clip (2 motion fields + {len(dists)} interpolated frame(s)) per step; "
              "the GPU arm's step is `pairs_per_pass_per_gpu x passes_per_step` such pairs per GPU")
    line = {
        "impl": "reference", "metric": "interpolated_frames_per_s", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world, args.pairs or P_default),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Python/numba (single-threaded); this arm times its C restatement (oracle/, bit-exact vs the "
                "reference) with OpenMP over block rows on all host cores; a step here is ONE pair of the workload (bounded sample)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def microbench_peaks(ctx):
    """(32-bit VABSDIFF, VABSDIFF4.U8.ACC) issue rates in T SAD px-op/s, measured now on this GPU."""
    ub_ms, ub_ops = min((ctx.microbench(0, 8192) for _ in range(3)), key=lambda x: x[0])
    ub4_ms, ub4_ops = min((ctx.microbench(15, 8192) for _ in range(3)), key=lambda x: x[0])
    return ub_ops / (ub_ms * 1e-3) / 1e12, ub4_ops / (ub4_ms * 1e-3) / 1e12


def run_cfg5(args):
    """8K single-pair mode: STRONG scaling -- one pair, one horizontal band per rank, two halo exchanges."""
    import torch
    import torch.distributed as dist
    from interpolatory_b200._native import lib
    from interpolatory_b200.device import DeviceContext
    from interpolatory_b200.synth import make_pair_tiled
    from interpolatory_b200.tiling import BandPlan, BandWorker, exchange_distributed
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lib()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    H, W, st, fps_in, fps_out, _ = WORKLOADS["cfg5"]
    bs, R = int(st['block_size']), int(st['target_region'])
    a, b = make_pair_tiled(H, W, 5)                                 # the pair tests/golden/fullsize_sha.npz pins against the oracle
    from interpolatory_b200.device import PinnedBuffer
    pin = PinnedBuffer((3, H, W, 3))                                # e2e copies go from / to page-locked host memory
    pin.array[0] = a; pin.array[1] = b
    a, b = pin.array[0], pin.array[1]
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ctx = DeviceContext(H, W, device=local_rank, stream=stream.cuda_stream)
    plan = BandPlan(H, W, bs, R, world)
    wk = BandWorker(ctx, plan, rank, local_rank)
    out = pin.array[2]; out[...] = 0
    dist05 = np.float32(0.5)
    halo_ev = []

    def step(e2e, time_halo=False):
        if e2e:
            wk.upload(a, b)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if time_halo else None
        if world > 1:
            if ev: ev[0].record(stream)
            exchange_distributed(wk, "frame")
            if ev: ev[1].record(stream)
        wk.frames_ready()
        wk.estimate()
        if world > 1:
            if ev: ev[2].record(stream)
            exchange_distributed(wk, "mv")
            if ev: ev[3].record(stream)
        wk.interpolate(dist05, True)
        if e2e:
            wk.download(out)
        if ev and world > 1:
            halo_ev.append(ev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wk.upload(a, b)
    for _ in range(max(args.warmup, 1)):
        step(False)
    barrier()
    # single-stream profile pass: per-launch time of the band's prefilter launches (2 per step: one per direction)
    ctx.profile_start(64)
    for _ in range(2):
        step(False, time_halo=True)
    barrier()
    kinds = ctx.profile_collect_kinds()
    n_surv, pf_overflow = ctx.fs_prefilter_stats()
    halo_ms = float(np.mean([e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]) for e in halo_ev])) if halo_ev else 0.0
    res = {}
    steps = max(args.steps, 1)
    for name, e2e in (("value", False), ("e2e", True)):
        # >= MIN_TIMED_S in total: repeat the K steps as a block (the pair stays the same; inputs are 2 x 99.5 MB + planes >> L2)
        step(e2e); barrier()
        t0 = time.perf_counter(); step(e2e); barrier()
        one = time.perf_counter() - t0
        reps = max(1, int(math.ceil(MIN_TIMED_S / (steps * one))))
        rt = torch.tensor([reps], device="cuda")
        if world > 1:
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        reps = int(rt[0])
        ctx.reset_launch_count()
        barrier(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps * reps):
            step(e2e)
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([e0.elapsed_time(e1) if not e2e else wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = (float(t[0]), ctx.launch_count(), reps)
    # every rank re-does the whole pair on its own GPU (untimed) and checks its band bit-for-bit
    ctx.upload_frame(2, a); ctx.upload_frame(3, b)
    ctx.me_fs(2, 3, bs, R, 2); ctx.me_fs(3, 2, bs, R, 3)
    ctx.mci_bidir(2, 3, 2, 3, dist05, 6)
    full = ctx.download_frame(6)
    y0r, y1r = plan.rows[rank]
    ok = torch.tensor([1 if np.array_equal(full[y0r:y1r], out[y0r:y1r]) else 0], device="cuda")
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok[0]) != 1:
        raise SystemExit("band-split output differs from the single-GPU result")
    digest = int(np.asarray(out[y0r:y1r], np.int64).sum())
    peak_tops, peak4_tops = microbench_peaks(ctx)
    if rank == 0:
        fbytes, mbytes = plan.bytes_moved()
        y0, y1 = plan.rows[0]
        fs_ms, fs_n, fs_work = kinds["fs"]
        mci_ms, mci_n, mci_work = kinds["mci"]
        peaks, peak_src = load_peaks()
        used_pf = n_surv > 0 or pf_overflow > 0
        ach = fs_work / (fs_ms * 1e-3) / 1e12 if fs_ms > 0 else None
        pk = peak4_tops if used_pf else peak_tops
        n_steps_v, n_steps_e = steps * res["value"][2], steps * res["e2e"][2]
        line = {"metric": "interpolated_frames_per_s", "value": n_steps_v / (res["value"][0] * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": res["value"][0] / steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
                "config": {"workload": WORKLOAD_TEXT["cfg5"], "H": H, "W": W, "settings": st, "bands": world,
                           "pairs_per_step": res["value"][2], "halo_bytes_per_pair": {"frame_rows": fbytes, "mv_rows": mbytes},
                           "l2_policy": "inputs larger than L2: two 99.5 MB frames + 133 MB luma planes per pair"},
                "e2e": {"value": n_steps_e / (res["e2e"][0] * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": res["e2e"][2] * 2 * (y1 - y0) * W * 3, "d2h_bytes_per_step": res["e2e"][2] * (y1 - y0) * W * 3,
                        "note": "bytes are per rank (its own band rows of both frames in, its band of the output frame out), page-locked host memory; upload, search, interpolation and download of a pair run back to back (single-pair latency mode)"},
                "gpu_launches": int(res["value"][1]), "band0_output_checksum": digest,
                "roofline": {"kernel": ("fs8_prefilter_kernel" if used_pf else "fs_tiled_kernel") + " over rank 0's band (two launches per pair: one per direction)",
                             "bound": "int_alu", "achieved": ach, "peak": pk, "unit": "T SAD px-op/s", "frac": ach / pk if ach else None,
                             "peak_source": "VABSDIFF4.U8.ACC issue-rate microbenchmark measured in this run" if used_pf else "VABSDIFF issue-rate microbenchmark measured in this run",
                             "ops_per_launch": fs_work / max(fs_n, 1), "launches_timed": fs_n, "avg_launch_ms": fs_ms / max(fs_n, 1),
                             "survivors_last_launch": n_surv, "traffic": None},
                "roofline_mci": {"kernel": "mci_splat_kernel over rank 0's band (+ 10 hole-fill halo rows)", "bound": "hbm",
                                 "achieved": mci_work / (mci_ms * 1e-3) / 1e9 if mci_ms > 0 else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": (mci_work / (mci_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if mci_ms > 0 else None, "peak_source": peak_src,
                                 "bytes_per_launch": mci_work / max(mci_n, 1), "launches_timed": mci_n, "avg_launch_ms": mci_ms / max(mci_n, 1), "traffic": None},
                "halo": {"bytes_per_pair_all_ranks": fbytes + mbytes, "ms_per_pair_rank0": halo_ms,
                         "achieved_gbs_rank0": ((fbytes + mbytes) / max(world, 1)) / (halo_ms * 1e-3) / 1e9 if halo_ms > 0 else None,
                         "link": "NVLink 5 / NVSwitch through NCCL send/recv (900 GB/s per direction nominal): the exchange is latency-, not bandwidth-bound",
                         "share_of_step": halo_ms / (res["value"][0] / n_steps_v) if halo_ms > 0 else 0.0},
                "verified": "every rank's band equals the same rows of a single-GPU run of the whole pair (bit-exact); the single-GPU "
                            "result is pinned against the oracle's digests by tests/test_gpu_parity.py::test_cfg5_8k_full_size_vs_oracle_digests"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def load_natural_pairs(H, W):
    """The reference's 12 Middlebury pairs (tests/golden/middlebury_pairs.npz), each tiled to H x W (both frames of a
    pair alike, so the motion is preserved)."""
    import io
    from PIL import Image
    z = np.load(os.path.join(ROOT, "tests", "golden", "middlebury_pairs.npz"), allow_pickle=False)
    out = []
    for n in z["names"]:
        fr = []
        for f in ("frame10", "frame11"):
            im = np.array(Image.open(io.BytesIO(z[f"{n}/{f}"].tobytes())).convert("RGB"))
            reps = (-(-H // im.shape[0]), -(-W // im.shape[1]), 1)
            fr.append(np.ascontiguousarray(np.tile(im, reps)[:H, :W]))
        out.append((str(n), fr[0], fr[1]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per pass per GPU (default: per workload)")
    ap.add_argument("--passes", type=int, default=0, help="passes over the pairs per step (default: enough for >= 2.2 s of timed steps)")
    ap.add_argument("--contexts", type=int, default=8, help="device contexts (CUDA streams) of the resident and the e2e pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the natural-content arm, the class-API line and the host-link measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.steps = max(args.steps, 1)

    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cfg5":
        return run_cfg5(args)

    import torch
    import torch.distributed as dist
    from interpolatory_b200 import util as U
    from interpolatory_b200._native import lib
    from interpolatory_b200.device import DeviceContext, PinnedBuffer
    from interpolatory_b200.pipeline import PairPipeline, interpolation_dists, pair_params_from_settings

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MEMCI path has no CPU fallback")
    lib()                                         # fail loudly if the CUDA library is missing
    bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    H, W, st, fps_in, fps_out, P_default = WORKLOADS[args.workload]
    P = args.pairs or P_default
    dists = interpolation_dists(fps_out / fps_in)
    nd = len(dists)
    params = pair_params_from_settings((H, W), **st)
    frame_bytes = H * W * 3
    is_fs = st['me_mode'] == 'fs'

    # ---- inputs: this rank's run of the shared clip (pair-index sharding, no collective) ----
    frames = shared_clip_segment(H, W, P, rank)
    dev_frames = torch.from_numpy(frames).cuda()
    stream = torch.cuda.Stream()              # a real (non-null) stream: the library enqueues on it, torch events time it
    torch.cuda.set_stream(stream)
    ctx = DeviceContext(H, W, device=local_rank, stream=stream.cuda_stream)
    base_ptr = dev_frames.data_ptr()

    bidir2 = st.get('mci_mode') == 'bidir2'
    if bidir2:
        from interpolatory_b200.MCI.Bidirectional_2 import get_weight_kern
        b2_w = get_weight_kern(8)

    def enqueue_pair(c, k):
        """pair k (frames k, k+1 of the resident run) on context c: 2 motion fields + nd interpolated frames"""
        a, b = k % 3, (k + 1) % 3
        if c.first:                                        # first pair of this context's run: both frames are new
            c.set_frame_device(a, base_ptr + k * frame_bytes)
            c.first = False
        c.set_frame_device(b, base_ptr + (k + 1) * frame_bytes)
        c.estimate_pair(params, a, b, 0, 1)
        for j, d in enumerate(dists):
            if bidir2:
                c.mci_bidir2(a, b, 0, 1, np.float32(d), np.float32(1 - d), b2_w, 6 + (j & 1))
            else:
                c.mci_bidir(a, b, 0, 1, d, 6 + (j & 1))

    def pass_resident(ctxs):
        """One pass = P pairs.  Each context (one CUDA stream) takes a contiguous run of pairs; the runs are enqueued
        round-robin so that the small kernels and the tail of one pair overlap another pair's full search."""
        C = len(ctxs)
        bounds = [round(i * P / C) for i in range(C + 1)]
        pos = [bounds[i] for i in range(C)]
        for c in ctxs:
            c.first = True
        live = True
        while live:
            live = False
            for i, c in enumerate(ctxs):
                if pos[i] < bounds[i + 1]:
                    enqueue_pair(c, pos[i])
                    pos[i] += 1
                    live = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def agree_max(v):
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- pass 1, ONE stream: per-launch device times by kernel family (roofline numerators).  A launch timed
    #      while another stream's persistent grid drains would include queueing, so the kernels are timed alone. ----
    for _ in range(max(args.warmup, 1)):
        pass_resident([ctx])
    barrier()
    n_cand, px_ops, _ = ctx.fs_stats() if is_fs else (0, 0, 0)
    prof_passes = 2
    ctx.profile_start(prof_passes * P * (12 + 4 * nd) + 64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(prof_passes):
        pass_resident([ctx])
    e1.record(stream)
    barrier()
    ms_single = e0.elapsed_time(e1) / prof_passes
    kinds = ctx.profile_collect_kinds()
    n_surv, pf_overflow = ctx.fs_prefilter_stats() if is_fs else (0, 0)
    n_redo = ctx.fs_stats()[2] if is_fs else 0
    used_prefilter = n_surv > 0 or pf_overflow > 0

    # ---- pass 2, the measured value: the same P pairs per pass over C contexts / streams (inputs resident in HBM) ----
    n_res = max(1, min(args.contexts, P))
    streams = [stream] + [torch.cuda.Stream() for _ in range(n_res - 1)]
    ctxs = [ctx] + [DeviceContext(H, W, device=local_rank, stream=s_.cuda_stream) for s_ in streams[1:]]
    tstream = torch.cuda.Stream()

    def timed_resident(n_pass):
        """device time (ms) of n_pass passes over the C streams"""
        barrier()
        e0.record(tstream)
        for s_ in streams:
            s_.wait_event(e0)                      # every stream starts after e0 ...
        for _ in range(n_pass):
            pass_resident(ctxs)
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            tstream.wait_event(ev)                 # ... and e1 completes after the last one finishes
        e1.record(tstream)
        barrier()
        return e0.elapsed_time(e1)

    for _ in range(max(args.warmup, 1)):
        pass_resident(ctxs)
    t_pass = agree_max(timed_resident(8) / 8) * 1e-3          # (short runs pay the pipeline fill: margin below)
    passes = args.passes or max(1, int(math.ceil(1.2 * MIN_TIMED_S / (args.steps * t_pass))))
    passes = int(agree_max(passes))
    for _ in range(args.warmup):                           # W untimed warm-up steps of the real size
        for _ in range(passes):
            pass_resident(ctxs)
    clocks = ClockSampler(local_rank)
    clocks.start()
    for c in ctxs:
        c.reset_launch_count()
    ms = timed_resident(args.steps * passes)
    launches = sum(c.launch_count() for c in ctxs)
    n_holes = ctxs[-1].mci_stats()
    resident_last = ctxs[-1].download_frame(6 + ((nd - 1) & 1)).copy()

    # the floor: the same passes with the prefilter pinned off (nothing pruned: the exact kernel searches every pair)
    ms_floor = exact = None
    if is_fs and used_prefilter:
        prev = os.environ.get("MEMCI_FS_PREFILTER")
        os.environ["MEMCI_FS_PREFILTER"] = "0"
        pass_resident(ctxs)
        n_floor = max(2, min(passes, int(math.ceil(0.4 / t_pass))))
        ms_floor = agree_max(timed_resident(n_floor) / n_floor)
        pass_resident([ctx]); barrier()
        ctx.profile_start(P * (12 + 4 * nd) + 64)
        pass_resident([ctx]); barrier()
        exact = ctx.profile_collect_kinds()["fs"]
        if prev is None:
            del os.environ["MEMCI_FS_PREFILTER"]
        else:
            os.environ["MEMCI_FS_PREFILTER"] = prev
        pass_resident(ctxs); barrier()

    # ---- e2e: host buffers, H2D + D2H inside the timed region, through the public pipeline API ----
    pin_in = PinnedBuffer((P + 1, H, W, 3))
    pin_in.array[...] = frames
    pin_out = PinnedBuffer((P * nd, H, W, 3))
    pipe = PairPipeline(H, W, device=local_rank, n_contexts=args.contexts, **st)
    for _ in range(max(args.warmup, 1)):
        pipe.run(pin_in.array, dists, pin_out.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(6):
        pipe.run(pin_in.array, dists, pin_out.array, sync=False)
    pipe.sync(); barrier()
    t_pass_e = agree_max((time.perf_counter() - t0) / 6)
    passes_e = args.passes or int(agree_max(max(1, int(math.ceil(1.2 * MIN_TIMED_S / (args.steps * t_pass_e))))))
    for _ in range(min(args.warmup, 2) * passes_e):
        pipe.run(pin_in.array, dists, pin_out.array, sync=False)
    pipe.sync(); barrier()
    l0 = pipe.launch_count()
    up0 = pipe.uploads
    t0 = time.perf_counter()
    for _ in range(args.steps * passes_e):
        pipe.run(pin_in.array, dists, pin_out.array, sync=False)      # every pass: H2D of its P + C frames, D2H of its P * nd frames
    pipe.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    e2e_launches = pipe.launch_count() - l0
    e2e_uploads = pipe.uploads - up0
    if not np.array_equal(pin_out.array[P * nd - 1], resident_last):
        raise SystemExit("e2e pipeline output differs from the device-resident path")
    n_ctx = min(args.contexts, P)
    pipe.close()

    # ---- host link: duplex copy rate of the same page-locked buffers, all ranks at once ----
    host_link = None
    if not args.no_extras:
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        src_t = torch.from_numpy(pin_in.array)
        dst_t = torch.from_numpy(pin_out.array)
        n_f = min(P, 16, P * nd)
        dev_a = torch.empty((n_f, H, W, 3), dtype=torch.uint8, device="cuda")
        dev_b = torch.empty((n_f, H, W, 3), dtype=torch.uint8, device="cuda")

        def link_pass(do_in, do_out, reps):
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                if do_in:
                    with torch.cuda.stream(s_in):
                        dev_a.copy_(src_t[:n_f], non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s_out):
                        dst_t[:n_f].copy_(dev_b, non_blocking=True)
            barrier()
            return n_f * frame_bytes * reps / agree_max(time.perf_counter() - t0) / 1e9
        link_pass(True, True, 2)
        reps = max(4, int(0.3 / (n_f * frame_bytes / 20e9)))
        host_link = {"h2d_alone_gbs": link_pass(True, False, reps), "d2h_alone_gbs": link_pass(False, True, reps),
                     "duplex_each_way_gbs": link_pass(True, True, reps), "ranks_copying_at_once": world,
                     "how": "torch copy_ between the bench's own page-locked buffers and device tensors on two streams, per rank; max time over ranks"}
        del dev_a, dev_b

    # ---- class API: the drop-in call, MEMCI(...).interpolate_video(), on the same clip ----
    class_api = None
    if not args.no_extras and world == 1:
        from interpolatory_b200 import MEMCI, ArrayVideoStream
        it = MEMCI(fps_out, max_cache_size=P + 1, **st)
        it.sink_keeps_no_reference = True                  # like the reference's video writer: a frame is consumed before the next is asked for
        it.set_video_stream(ArrayVideoStream([pin_in.array[i] for i in range(P + 1)], fps_in, page_locked=True))
        it.max_out_frames = int(P * fps_out / fps_in) + 1
        cnt = [0]

        def sink(f):
            cnt[0] += 1
        old_flag = U.debug_flags['debug_progress']
        U.debug_flags['debug_progress'] = False
        it.interpolate_video(sink=sink)
        reps = max(1, int(math.ceil(1.0 / max(t_pass_e, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(reps):
            it.interpolate_video(sink=sink)
        dt = time.perf_counter() - t0
        U.debug_flags['debug_progress'] = old_flag
        it.close()
        class_api = {"value": reps * P * nd / dt, "unit": "interpolated frames/s", "output_frames_per_call": cnt[0] // (reps + 1),
                     "how": f"MEMCI({fps_out}, max_cache_size={P + 1}, ...).interpolate_video(sink) over the {P + 1}-frame clip, {reps} calls; "
                            "source frames handed over by an ArrayVideoStream on page-locked memory, the sink drops the frame (like a video "
                            "writer it keeps no reference); every call starts and drains its own pipeline"}

    # ---- natural content (N = 1): the reference's Middlebury pairs tiled to the frame size ----
    natural = None
    if not args.no_extras and world == 1 and is_fs and not bidir2:
        nat = load_natural_pairs(H, W)
        nat_dev = torch.from_numpy(np.stack([f for _, a_, b_ in nat for f in (a_, b_)])).cuda()
        nptr = nat_dev.data_ptr()

        def nat_pass():
            for i in range(len(nat)):
                c = ctxs[i % len(ctxs)]
                c.set_frame_device(0, nptr + (2 * i) * frame_bytes)
                c.set_frame_device(1, nptr + (2 * i + 1) * frame_bytes)
                c.estimate_pair(params, 0, 1, 0, 1)
                for j, d in enumerate(dists):
                    c.mci_bidir(0, 1, 0, 1, d, 6 + (j & 1))

        def nat_rate(mode):
            prev = os.environ.get("MEMCI_FS_PREFILTER")
            if mode is None:
                os.environ.pop("MEMCI_FS_PREFILTER", None)
            else:
                os.environ["MEMCI_FS_PREFILTER"] = mode
            for _ in range(3):
                nat_pass()                                 # also lets the content feedback settle
            barrier()
            n_pass = 6
            t0 = time.perf_counter()
            for _ in range(n_pass):
                nat_pass()
            barrier()
            dt = time.perf_counter() - t0
            if prev is None:
                os.environ.pop("MEMCI_FS_PREFILTER", None)
            else:
                os.environ["MEMCI_FS_PREFILTER"] = prev
            return n_pass * len(nat) * nd / dt
        # survivor fraction per pair with the prefilter pinned on
        prev = os.environ.get("MEMCI_FS_PREFILTER")
        os.environ["MEMCI_FS_PREFILTER"] = "1"
        fracs = {}
        for i, (name, _, _) in enumerate(nat):
            ctx.set_frame_device(0, nptr + (2 * i) * frame_bytes); ctx.set_frame_device(1, nptr + (2 * i + 1) * frame_bytes)
            ctx.estimate_pair(params, 0, 1, 0, 1)
            sv, ov = ctx.fs_prefilter_stats()
            fracs[name] = None if ov else sv / max(ctx.fs_stats()[0], 1)
        if prev is None:
            os.environ.pop("MEMCI_FS_PREFILTER", None)
        else:
            os.environ["MEMCI_FS_PREFILTER"] = prev
        natural = {"value": nat_rate(None), "value_prefilter_on": nat_rate("1"), "value_prefilter_off": nat_rate("0"), "unit": "interpolated frames/s",
                   "survivor_fraction": fracs, "pairs": len(nat),
                   "content": "benchmarks/middlebury frame10 / frame11 of the reference (tests/golden/middlebury_pairs.npz), each pair tiled to the frame size; "
                              "device resident, same streams as `value`; `value` lets the library's content feedback choose (a null fraction = that pair fell "
                              "back to the exact kernel)"}
        del nat_dev

    peak_tops, peak4_tops = microbench_peaks(ctx)

    t_max = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(t_max[0]), float(t_max[1])

    if rank == 0:
        peaks, peak_src = load_peaks()
        total_interp = world * P * nd * args.steps * passes
        total_interp_e = world * P * nd * args.steps * passes_e
        value = total_interp / (ms_all * 1e-3)
        e2e_value = total_interp_e / (e2e_ms_all * 1e-3)
        out_per_interp = (fps_out / fps_in) / nd            # output frames (incl. pass-through) per interpolated frame

        def roof(kind, unit_scale, peak, unit, kernel, bound, peak_source, extra=None):
            ms_k, n_k, work_k = kinds[kind]
            if n_k == 0 or ms_k <= 0:
                return None
            ach = work_k / (ms_k * 1e-3) / unit_scale
            r = {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                 "peak_source": peak_source, "work_per_launch": work_k / n_k, "launches_timed": n_k, "avg_launch_ms": ms_k / n_k,
                 "share_of_single_stream_pass": ms_k / (ms_single * prof_passes)}
            if extra:
                r.update(extra)
            return r

        src4 = "VABSDIFF4.U8.ACC issue-rate microbenchmark measured in this run (memci_microbench mode 15)"
        src1 = "VABSDIFF issue-rate microbenchmark measured in this run (memci_microbench mode 0)"
        if is_fs and used_prefilter:
            roofline = roof("fs", 1e12, peak4_tops, "T SAD px-op/s",
                            "fs8_prefilter_kernel (8-bit luma-SAD of every candidate, block minima; phase 1 of the exact two-phase full search -- the scan, refine, replay, output and redo kernels that follow are the `refine` entry of single_stream.kernel_ms_per_pass)",
                            "int_alu", src4,
                            {"ops_per_launch": px_ops, "candidates_per_launch": n_cand, "survivors_per_launch": n_surv,
                             "survivor_fraction": n_surv / max(n_cand, 1), "redo_blocks_per_launch": n_redo, "list_overflowed": bool(pf_overflow),
                             "traffic": load_traffic("fs8_prefilter_kernel") if args.workload == "cfg2" else None,
                             "vs_exact_arithmetic_peak": {"peak": peak_tops, "note": "same algorithmic px-ops against the 32-bit VABSDIFF rate the exact kernel is bound by"}})
            if roofline:
                roofline["vs_exact_arithmetic_peak"]["frac"] = roofline["achieved"] / peak_tops
        elif is_fs:
            roofline = roof("fs", 1e12, peak_tops, "T SAD px-op/s", "fs_tiled_kernel (full-search luma-SAD + argmin)", "int_alu", src1,
                            {"ops_per_launch": px_ops, "candidates_per_launch": n_cand,
                             "traffic": load_traffic("fs_tiled_kernel") if args.workload == "cfg2" else None})
        else:
            roofline = roof("hbma", 1e12, peak_tops, "T cost px-op/s",
                            "hbma_fs_kernel + hbma_densify_parent_kernel<8|4> (one entry per HBMA call: coarse search + every densification; 2 calls per pair)",
                            "int_alu", src1,
                            {"work": "nominal cost pixel-ops: (2 win + 1)^2 candidates per coarse block, 3 vectors x (2 sub + 1)^2 per child block (hbma.py:100-140)",
                             "traffic": None})
        line = {
            "metric": "interpolated_frames_per_s", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": workload_config(args.workload, world, P),
            "run": {"passes_per_step": passes, "passes_per_step_e2e": passes_e, "interpolated_per_pair": nd, "streams": n_res,
                    "timed_region_s": ms_all * 1e-3, "pairs_per_step_all_gpus": world * P * passes},
            "output_frames_per_s": value * out_per_interp,
            "sad_gops": ((px_ops * world * P * args.steps * passes) / (ms_all * 1e-3) / 1e9) if is_fs else None,
            "value_prefilter_off": (world * P * nd / (ms_floor * 1e-3)) if ms_floor else None,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": (e2e_uploads // args.steps) * frame_bytes,
                    "d2h_bytes_per_step": passes_e * P * nd * frame_bytes, "contexts": n_ctx, "timed_region_s": e2e_ms_all * 1e-3,
                    "output_frames_per_s": e2e_value * out_per_interp, "host_link": host_link, "class_api": class_api},
            "gpu_launches": int(launches),
            "gpu_launches_e2e": int(e2e_launches),
            "clocks": clk,
            "roofline": roofline,
            "roofline_fs_exact": ({
                "kernel": "fs_tiled_kernel (the exact kernel: every candidate at 18-bit luma; runs when the prefilter is off or its survivor list overflows)",
                "bound": "int_alu", "achieved": exact[2] / (exact[0] * 1e-3) / 1e12, "peak": peak_tops, "unit": "T SAD px-op/s",
                "frac": exact[2] / (exact[0] * 1e-3) / 1e12 / peak_tops, "avg_launch_ms": exact[0] / exact[1], "launches_timed": exact[1],
                "peak_source": src1, "traffic": load_traffic("fs_tiled_kernel") if args.workload == "cfg2" else None,
            } if exact and exact[1] else None),
            "roofline_mci": roof("mci", 1e9, peaks["hbm_gbs"], "GB/s", "mci_splat_kernel (warp/blend, order-preserving gather)", "hbm", peak_src,
                                 {"holes_last_frame": n_holes, "traffic": load_traffic("mci_splat_kernel") if args.workload == "cfg2" else None}),
            "roofline_hbm_small": roof("hbm_small", 1e9, peaks["hbm_gbs"], "GB/s",
                                       "pyr_down_kernel + luma_kernel per pyramid level, smooth_kernel (one entry per launch group)", "hbm", peak_src, {"traffic": None}),
            "single_stream": {"value": world * P * nd / (ms_single * 1e-3), "ms_per_pass": ms_single, "passes": prof_passes,
                              "kernel_ms_per_pass": {k: v[0] / prof_passes for k, v in kinds.items() if v[1]},
                              "note": "same pass on one CUDA stream; the per-launch kernel times of the roofline lines come from this pass "
                                      "(a launch timed next to another stream's persistent grid would include queueing)"},
            "natural": natural,
        }
        if host_link and host_link.get("duplex_each_way_gbs"):
            moved_in = e2e_uploads * frame_bytes / e2e_s / 1e9
            moved_out = args.steps * passes_e * P * nd * frame_bytes / e2e_s / 1e9
            line["e2e"]["host_link_frac"] = max(moved_in, moved_out) / host_link["duplex_each_way_gbs"]
            line["e2e"]["moved_gbs"] = {"h2d": moved_in, "d2h": moved_out, "note": "rank 0's own traffic against the per-rank duplex rate with all ranks copying"}
        if world == 1 and not args.no_cpu_baseline:
            if getattr(bind_to_gpu_numa_node, "original", None):
                os.sched_setaffinity(0, bind_to_gpu_numa_node.original)        # the CPU baseline may use every host core
            dt, cores = cpu_reference_sample(H, W, st, dists, frames)
            line["cpu_baseline"] = {"value": nd / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"1 full-size frame pair of the same workload ({dt:.2f} s): C restatement of the "
                                              "reference (oracle/), OpenMP over block rows"}
        print(json.dumps(line), flush=True)

    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
