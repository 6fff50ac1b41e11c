"""One clip, N GPUs (north_star: "work is partitioned across the GPUs of one box by frame-pair index,
with no collectives").  Output frames of source pair k depend on frames k and k + 1 only
(reference: MCI/Bidirectional.py:128-157), so the pairs of a clip are cut into one contiguous run per GPU
(sharding.runs_of_rank: every source frame is then uploaded to at most two GPUs) and each run goes through that
GPU's own PairPipeline -- staged uploads, several CUDA streams, outputs drained straight into the caller's array.

Two ways to drive it:
  * MultiGpuPipeline: one process, one host thread per GPU (every library call is a ctypes call, which releases the
    GIL, and sets its own CUDA device), the drop-in for a caller that holds the whole clip -- what
    BaseInterpolator.interpolate_video uses when it is given more than one device;
  * under torchrun (one process per GPU): rank_pairs() gives a rank its run of the same partition and the rank feeds
    it to an ordinary PairPipeline -- what bench.py does.
Nothing is exchanged between the GPUs in either mode."""
import threading

from .pipeline import PairPipeline
from .sharding import runs_of_rank


def rank_pairs(pairs, rank, world_size):
    """The contiguous run of `pairs` (a sorted list of pair indices that have work) that belongs to `rank`."""
    lo, hi = runs_of_rank(len(pairs), rank, world_size)
    return pairs[lo:hi]


class MultiGpuPipeline:
    def __init__(self, H, W, devices=(0,), n_contexts=3, **settings):
        self.H, self.W = int(H), int(W)
        self.devices = [int(d) for d in devices]
        if not self.devices:
            raise ValueError("MultiGpuPipeline needs at least one device")
        self.pipes = [PairPipeline(H, W, device=d, n_contexts=n_contexts, **settings) for d in self.devices]

    # -- the PairPipeline interface -------------------------------------------------------------------------
    def plan_clip(self, n_frames, fps_in, fps_out):
        return self.pipes[0].plan_clip(n_frames, fps_in, fps_out)

    def run_clip(self, frames, fps_in, fps_out, out, sync=True):
        """Same contract as PairPipeline.run_clip: fills out[i] for every output frame i of plan_clip(); returns the plan."""
        plan = self.plan_clip(len(frames), fps_in, fps_out)
        per_pair = {}
        for i, (k, d) in enumerate(plan):
            if d == 0.0:
                out[i] = frames[k]
            else:
                per_pair.setdefault(k, []).append((i, d))
        self._dispatch(frames, per_pair, out)
        if sync:
            self.sync()
        return plan

    def run(self, frames, dists, out, sync=True):
        nd = len(dists)
        per_pair = {k: [(k * nd + j, d) for j, d in enumerate(dists)] for k in range(len(frames) - 1)}
        self._dispatch(frames, per_pair, out)
        if sync:
            self.sync()
        return out

    def _dispatch(self, frames, per_pair, out):
        pairs = sorted(per_pair)
        n = len(self.pipes)
        shards = [rank_pairs(pairs, r, n) for r in range(n)]
        self.last_shards = shards
        errors = [None] * n

        def work(r):
            try:
                self.pipes[r]._run_pairs(frames, {k: per_pair[k] for k in shards[r]}, out)
            except BaseException as e:                   # re-raised in the caller's thread
                errors[r] = e

        if n == 1:
            work(0)
        else:
            threads = [threading.Thread(target=work, args=(r,), name=f"memci-gpu{self.devices[r]}") for r in range(n)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        for e in errors:
            if e is not None:
                raise e

    def sync(self):
        for p in self.pipes:
            p.sync()

    def launch_count(self):
        return sum(p.launch_count() for p in self.pipes)

    def close(self):
        for p in self.pipes:
            p.close()
        self.pipes = []
