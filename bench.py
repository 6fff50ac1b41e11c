#!/usr/bin/env python
"""bench.py -- interpolated frames/s of the MEMCI hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

Workload (BASELINE.json configs[1], "cfg2"): 1080p, full search 8x8 blocks +-16, 30 -> 60 fps,
bidirectional MCI.  A step = one pass of the hot path over a batch of P independent frame pairs:
2 full-search motion fields (forward + backward) and one interpolated frame per pair.  The
pass-through source frames of the 60 fps output involve no computation and are not counted in
`value` (interpolated frames/s); `output_frames_per_s` adds them back.

value    : inputs resident in HBM before the timed region, outputs left in HBM; the step's P pairs are spread over
           C contexts (= CUDA streams, --contexts) so that one pair's small kernels overlap another pair's search;
           timed with CUDA events on a timing stream that fans out to / in from the C streams.
e2e      : same work through the public pipeline API with page-locked HOST buffers, the H2D copy
           of every source frame and the D2H copy of every interpolated frame inside the timed region.
roofline : dominant kernel (fs8_prefilter_kernel, phase 1 of the two-phase exact full search): SAD pixel-ops per launch
           (exact candidate count of the reference's own bounds x 64 x 2 directions) / mean launch time from CUDA events
           recorded around every launch of a single-stream pass of the same step (`single_stream`: timed next to
           another stream's persistent grid a launch would include queueing), against the VABSDIFF4.U8.ACC issue rate
           measured by a microbenchmark in this run.  Integer-ALU bound, not HBM bound.
           `roofline_fs_exact` is the exact kernel (fs_tiled_kernel, prefilter off) against the 32-bit VABSDIFF rate,
           `roofline_mci` the HBM-bound MCI splat kernel.
cpu_baseline / --impl reference: the C restatement of the reference (oracle/, validated bit-exact
           against the reference) on the host cores, OpenMP over block rows, one full-size pair per sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, settings, fps_in, fps_out, default pairs per step)
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
    """SM clock / throttle reasons sampled DURING the timed regions (NVML in a thread, every ~2 ms;
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
                    time.sleep(0.002)
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
    if rank != 0:
        return
    from interpolatory_b200.synth import make_sequence
    from interpolatory_b200.pipeline import interpolation_dists
    H, W, st, fps_in, fps_out, _ = WORKLOADS[args.workload]
    dists = interpolation_dists(fps_out / fps_in)
    frames = make_sequence(H, W, 2, seed=0)
    for _ in range(args.warmup):
        cpu_reference_sample(H, W, st, dists, frames)
    t = 0.0
    cores = 1
    for _ in range(args.steps):
        dt, cores = cpu_reference_sample(H, W, st, dists, frames)
        t += dt
    value = args.steps * len(dists) / t
    sample = f"{args.steps} x 1 full-size frame pair (2 motion fields + {len(dists)} interpolated frame(s)) per step"
    line = {
        "impl": "reference", "metric": "interpolated_frames_per_s", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "pairs_per_step": 1},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Python/numba (single-threaded); this arm times its C restatement (oracle/, bit-exact vs the "
                "reference) with OpenMP over block rows on all host cores",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_cfg5(args):
    """8K single-pair mode: STRONG scaling -- one pair, one horizontal band per rank, two halo exchanges."""
    import torch
    import torch.distributed as dist
    from interpolatory_b200._native import lib
    from interpolatory_b200.device import DeviceContext
    from interpolatory_b200.synth import make_sequence
    from interpolatory_b200.tiling import BandPlan, BandWorker, exchange_distributed
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lib()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    H, W, st, fps_in, fps_out, _ = WORKLOADS["cfg5"]
    bs, R = int(st['block_size']), int(st['target_region'])
    base = make_sequence(H // 4, W // 4, 2, seed=7)                 # 1080p pair tiled 4x4 (cheap to build on every rank)
    a = np.ascontiguousarray(np.tile(base[0], (4, 4, 1))); b = np.ascontiguousarray(np.tile(base[1], (4, 4, 1)))
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    ctx = DeviceContext(H, W, device=local_rank, stream=stream.cuda_stream)
    plan = BandPlan(H, W, bs, R, world)
    wk = BandWorker(ctx, plan, rank, local_rank)
    out = np.zeros_like(a)
    dist05 = np.float32(0.5)

    def step(e2e):
        if e2e:
            wk.upload(a, b)
        if world > 1:
            exchange_distributed(wk, "frame")
        wk.frames_ready()
        wk.estimate()
        if world > 1:
            exchange_distributed(wk, "mv")
        wk.interpolate(dist05, True)
        if e2e:
            wk.download(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wk.upload(a, b)
    for _ in range(max(args.warmup, 1)):
        step(False)
    barrier()
    res = {}
    for name, e2e in (("value", False), ("e2e", True)):
        ctx.reset_launch_count()
        barrier(); t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step(e2e)
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([e0.elapsed_time(e1) if not e2e else wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = (float(t[0]), ctx.launch_count())
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
    if rank == 0:
        fbytes, mbytes = plan.bytes_moved()
        y0, y1 = plan.rows[0]
        line = {"metric": "interpolated_frames_per_s", "value": args.steps / (res["value"][0] * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["value"][0] / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
                "config": {"workload": WORKLOAD_TEXT["cfg5"], "H": H, "W": W, "settings": st, "bands": world,
                           "halo_bytes_per_pair": {"frame_rows": fbytes, "mv_rows": mbytes},
                           "l2_policy": "inputs larger than L2: two 99.5 MB frames + 133 MB luma planes per pair"},
                "e2e": {"value": args.steps / (res["e2e"][0] * 1e-3), "unit": "frames/s",
                        "h2d_bytes_per_step": 2 * (y1 - y0) * W * 3, "d2h_bytes_per_step": (y1 - y0) * W * 3,
                        "note": "bytes are per rank (its own band rows of both frames in, its band of the output frame out)"},
                "gpu_launches": int(res["value"][1]), "band0_output_checksum": digest,
                "verified": "every rank's band equals the same rows of a single-GPU run of the whole pair (bit-exact)"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per step per GPU (default: per workload)")
    ap.add_argument("--contexts", type=int, default=8, help="device contexts (CUDA streams) of the resident and the e2e pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "cfg5":
        return run_cfg5(args)

    import torch
    import torch.distributed as dist
    from interpolatory_b200 import build as B
    from interpolatory_b200._native import lib, MemciError
    from interpolatory_b200.device import DeviceContext, PinnedBuffer
    from interpolatory_b200.pipeline import PairPipeline, interpolation_dists, pair_params_from_settings
    from interpolatory_b200.synth import make_sequence

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

    # ---- inputs: every rank owns its own P pairs (weak scaling; pair-index sharding, no collective) ----
    frames = make_sequence(H, W, P + 1, seed=100 + rank)
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
        """pair k (frames k, k+1 of the resident clip) on context c: 2 motion fields + nd interpolated frames"""
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

    def step_resident(ctxs):
        """One step = P pairs.  Each context (one CUDA stream) takes a contiguous run of pairs; the runs are enqueued
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

    # ---- pass 1, ONE stream: per-launch device times of the two hot kernels (roofline numerators).  A launch timed
    #      while another stream's persistent grid drains would include queueing, so the kernels are timed alone. ----
    for _ in range(max(args.warmup, 1)):
        step_resident([ctx])
    barrier()
    n_cand, px_ops, _ = ctx.fs_stats()
    prof_steps = max(1, min(args.steps, 2))
    ctx.profile_start(prof_steps * P * (2 + nd) + 16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(prof_steps):
        step_resident([ctx])
    e1.record(stream)
    barrier()
    ms_single = e0.elapsed_time(e1) / prof_steps
    fs_ms_sum, fs_n, mci_ms_sum, mci_n = ctx.profile_collect()
    n_surv, pf_overflow = ctx.fs_prefilter_stats()
    n_redo = ctx.fs_stats()[2]
    used_prefilter = n_surv > 0 or pf_overflow > 0

    # ---- pass 1b: the exact full-search kernel alone (prefilter off), for the second roofline line ----
    exact_ms_sum = exact_n = 0
    if used_prefilter:
        prev = os.environ.get("MEMCI_FS_PREFILTER")
        os.environ["MEMCI_FS_PREFILTER"] = "0"
        step_resident([ctx]); barrier()
        ctx.profile_start(P * (2 + nd) + 16)
        step_resident([ctx]); barrier()
        exact_ms_sum, exact_n, _, _ = ctx.profile_collect()
        if prev is None:
            del os.environ["MEMCI_FS_PREFILTER"]
        else:
            os.environ["MEMCI_FS_PREFILTER"] = prev
        step_resident([ctx]); barrier()

    # ---- pass 2, the measured value: the same P pairs per step over C contexts / streams (inputs resident in HBM) ----
    n_res = max(1, min(args.contexts, P))
    streams = [stream] + [torch.cuda.Stream() for _ in range(n_res - 1)]
    ctxs = [ctx] + [DeviceContext(H, W, device=local_rank, stream=s_.cuda_stream) for s_ in streams[1:]]
    tstream = torch.cuda.Stream()
    for _ in range(max(args.warmup, 1)):
        step_resident(ctxs)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    for c in ctxs:
        c.reset_launch_count()
    barrier()
    e0.record(tstream)
    for s_ in streams:
        s_.wait_event(e0)                      # every stream starts after e0 ...
    for _ in range(args.steps):
        step_resident(ctxs)
    for s_ in streams:
        ev = torch.cuda.Event()
        ev.record(s_)
        tstream.wait_event(ev)                 # ... and e1 completes after the last one finishes
    e1.record(tstream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(c.launch_count() for c in ctxs)
    n_holes = ctxs[-1].mci_stats()
    resident_last = ctxs[-1].download_frame(6 + ((nd - 1) & 1)).copy()

    # ---- e2e: host buffers, H2D + D2H inside the timed region, through the public pipeline API ----
    pin_in = PinnedBuffer((P + 1, H, W, 3))
    pin_in.array[...] = frames
    pin_out = PinnedBuffer((P * nd, H, W, 3))
    pipe = PairPipeline(H, W, device=local_rank, n_contexts=args.contexts, **st)
    for _ in range(max(args.warmup, 1)):
        pipe.run(pin_in.array, dists, pin_out.array)
    barrier()
    l0 = pipe.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.run(pin_in.array, dists, pin_out.array, sync=False)      # every step: H2D of its P+C frames, D2H of its P*nd frames
    pipe.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    e2e_launches = pipe.launch_count() - l0
    if not np.array_equal(pin_out.array[P * nd - 1], resident_last):
        raise SystemExit("e2e pipeline output differs from the device-resident path")
    n_ctx = min(args.contexts, P)

    # ---- roofline denominator: VABSDIFF issue rate measured now, on this GPU ----
    ub_ms, ub_ops = min((ctx.microbench(0, 8192) for _ in range(3)), key=lambda x: x[0])
    peak_tops = ub_ops / (ub_ms * 1e-3) / 1e12
    ub4_ms, ub4_ops = min((ctx.microbench(15, 8192) for _ in range(3)), key=lambda x: x[0])
    peak4_tops = ub4_ops / (ub4_ms * 1e-3) / 1e12               # VABSDIFF4.U8.ACC: four pixel pairs per instruction

    t_max = torch.tensor([ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(t_max[0]), float(t_max[1])

    if rank == 0:
        peaks, peak_src = load_peaks()
        total_interp = world * P * nd * args.steps
        value = total_interp / (ms_all * 1e-3)
        e2e_value = total_interp / (e2e_ms_all * 1e-3)
        out_per_interp = (fps_out / fps_in) / nd            # output frames (incl. pass-through) per interpolated frame
        fs_avg_s = (fs_ms_sum / max(fs_n, 1)) * 1e-3
        achieved_tops = px_ops / fs_avg_s / 1e12 if fs_n else None
        mci_avg_s = (mci_ms_sum / max(mci_n, 1)) * 1e-3
        mci_bytes = 9.0 * H * W                               # 2 frame reads + 1 frame write
        luma_bytes = 2.0 * H * ((W + 3) // 4 * 4) * 4 + (-(-H // int(st['block_size']))) * (-(-W // int(st['block_size']))) * 12
        line = {
            "metric": "interpolated_frames_per_s", "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "pairs_per_step_per_gpu": P, "H": H, "W": W,
                       "settings": st, "interpolated_per_pair": nd, "streams": n_res,
                       "sharding": "pair index, independent per rank, no collective",
                       "l2_policy": f"inputs larger than L2: {((P + 1) * frame_bytes + 3 * H * W * 4) / 1e6:.0f} MB of frames + luma planes "
                                    "streamed per step vs 126 MB L2"},
            "output_frames_per_s": value * out_per_interp,
            "sad_gops": (px_ops * world * P * args.steps) / (ms_all * 1e-3) / 1e9,    # px_ops already covers both directions of a pair
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": (P + 1) * frame_bytes,
                    "d2h_bytes_per_step": P * nd * frame_bytes, "contexts": n_ctx,
                    "output_frames_per_s": e2e_value * out_per_interp},
            "gpu_launches": int(launches),
            "gpu_launches_e2e": int(e2e_launches),
            "clocks": clk,
            "roofline": ({
                "kernel": "fs8_prefilter_kernel (8-bit luma-SAD of every candidate + survivor selection; phase 1 of the exact two-phase full search)",
                "bound": "int_alu",
                "achieved": achieved_tops, "peak": peak4_tops, "unit": "T SAD px-op/s",
                "frac": (achieved_tops / peak4_tops) if achieved_tops else None,
                "peak_source": "VABSDIFF4.U8.ACC issue-rate microbenchmark measured in this run (memci_microbench mode 15)",
                "ops_per_launch": px_ops, "candidates_per_launch": n_cand, "launches_timed": fs_n,
                "avg_launch_ms": fs_avg_s * 1e3,
                "survivors_per_launch": n_surv, "survivor_fraction": n_surv / max(n_cand, 1), "redo_blocks_per_launch": n_redo,
                "list_overflowed": bool(pf_overflow),
                "traffic": load_traffic("fs8_prefilter_kernel") if args.workload == "cfg2" else None,
                "vs_exact_arithmetic_peak": {"peak": peak_tops, "frac": (achieved_tops / peak_tops) if achieved_tops else None,
                                             "note": "same algorithmic px-ops against the 32-bit VABSDIFF rate the exact kernel is bound by"},
            } if used_prefilter else {
                "kernel": "fs_tiled_kernel (full-search luma-SAD + argmin)", "bound": "int_alu",
                "achieved": achieved_tops, "peak": peak_tops, "unit": "T SAD px-op/s",
                "frac": (achieved_tops / peak_tops) if achieved_tops else None,
                "peak_source": "VABSDIFF issue-rate microbenchmark measured in this run (memci_microbench mode 0)",
                "ops_per_launch": px_ops, "candidates_per_launch": n_cand, "launches_timed": fs_n,
                "avg_launch_ms": fs_avg_s * 1e3,
                "traffic": load_traffic("fs_tiled_kernel") if args.workload == "cfg2" else None,
                "hbm_view": {"algorithmic_bytes_per_launch": luma_bytes, "achieved_gbs": luma_bytes / fs_avg_s / 1e9 if fs_n else None,
                             "peak_gbs": peaks["hbm_gbs"], "frac": (luma_bytes / fs_avg_s / 1e9) / peaks["hbm_gbs"] if fs_n else None,
                             "peak_source": peak_src},
            }),
            "roofline_fs_exact": ({
                "kernel": "fs_tiled_kernel (the exact kernel: every candidate at 18-bit luma; runs when the prefilter is off or its survivor list overflows)",
                "bound": "int_alu", "achieved": px_ops / ((exact_ms_sum / exact_n) * 1e-3) / 1e12, "peak": peak_tops, "unit": "T SAD px-op/s",
                "frac": px_ops / ((exact_ms_sum / exact_n) * 1e-3) / 1e12 / peak_tops, "avg_launch_ms": exact_ms_sum / exact_n, "launches_timed": exact_n,
                "peak_source": "VABSDIFF issue-rate microbenchmark measured in this run (memci_microbench mode 0)",
                "traffic": load_traffic("fs_tiled_kernel") if args.workload == "cfg2" else None,
            } if exact_n else None),
            "roofline_mci": {
                "kernel": "mci_splat_kernel (warp/blend, order-preserving gather)", "bound": "hbm",
                "achieved": mci_bytes / mci_avg_s / 1e9 if mci_n else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": (mci_bytes / mci_avg_s / 1e9) / peaks["hbm_gbs"] if mci_n else None, "peak_source": peak_src,
                "bytes_per_launch": mci_bytes, "launches_timed": mci_n, "avg_launch_ms": mci_avg_s * 1e3,
                "holes_last_frame": n_holes, "traffic": load_traffic("mci_splat_kernel") if args.workload == "cfg2" else None,
            },
            "single_stream": {"value": world * P * nd / (ms_single * 1e-3), "ms_per_step": ms_single, "steps": prof_steps,
                              "note": "same step on one CUDA stream; the per-launch kernel times of `roofline` / `roofline_mci` "
                                      "come from this pass (a launch timed next to another stream's persistent grid would include queueing)"},
            "kernel_share_of_step": {("fs8_prefilter" if used_prefilter else "fs_tiled"): fs_ms_sum / (ms_single * prof_steps), "mci_splat": mci_ms_sum / (ms_single * prof_steps),
                                     "of": "single_stream step"},
        }
        if world == 1 and not args.no_cpu_baseline:
            if getattr(bind_to_gpu_numa_node, "original", None):
                os.sched_setaffinity(0, bind_to_gpu_numa_node.original)        # the CPU baseline may use every host core
            dt, cores = cpu_reference_sample(H, W, st, dists, frames)
            line["cpu_baseline"] = {"value": nd / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"1 full-size frame pair of the same workload ({dt:.2f} s): C restatement of the "
                                              "reference (oracle/), OpenMP over block rows"}
        print(json.dumps(line), flush=True)

    pipe.close()
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
