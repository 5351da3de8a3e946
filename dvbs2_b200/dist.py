"""Multi-GPU plumbing: frames are independent, so a job shards by contiguous frame batches with no collective on the data
path; one collective remains, the SUM of the monitor counters {frames, bit errors, frame errors} — the
NVLink/NCCL replacement for aff3ct::tools::Monitor_reduction{,_MPI} (lib/aff3ct/include/Tools/Monitor/
Monitor_reduction.hxx:96-121, Monitor_reduction_MPI.hxx:93).  Works on any torch.distributed backend (nccl on the GPUs,
gloo in the CPU tests)."""


def shard_frames(n_frames, rank, world_size):
    """Contiguous, balanced split of `n_frames` frames: returns (first_frame, count) of `rank` (sizes differ by <= 1)."""
    if world_size <= 0 or not 0 <= rank < world_size or n_frames < 0:
        raise ValueError("bad shard request: n_frames=%d rank=%d world_size=%d" % (n_frames, rank, world_size))
    base, extra = divmod(n_frames, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def reduce_counters(counters, device=None, group=None):
    """All-reduce (SUM) a dict of integer counters over the ranks; returns the global dict on every rank.
    With no initialised process group (single process) the input is returned unchanged."""
    import torch
    import torch.distributed as dist

    keys = sorted(counters)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {k: int(counters[k]) for k in keys}
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(keys, (int(x) for x in t.tolist())))


def reduce_max(value, device=None, group=None):
    """MAX over ranks of a float (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class Monitor_BFER_reduction:
    """Stop criterion of the Monte-Carlo loop, reduced over ranks: done when the global frame-error count reaches
    `max_fe` (Monitor_BFER::fe_limit_achieved, lib/aff3ct/src/Module/Monitor/BFER/Monitor_BFER.cpp:190-200) or
    `max_frames` frames were simulated."""

    def __init__(self, max_fe=100, max_frames=0, device=None, group=None):
        self.max_fe, self.max_frames, self.device, self.group = int(max_fe), int(max_frames), device, group
        self.total = {"n_frames": 0, "n_be": 0, "n_fe": 0}

    def update(self, local):
        """`local` = this rank's cumulative counters (Transmitter_DVBS2_cuda.monitor()); returns the global ones."""
        self.total = reduce_counters({k: local[k] for k in ("n_frames", "n_be", "n_fe")}, self.device, self.group)
        return self.total

    def is_done(self):
        t = self.total
        return t["n_fe"] >= self.max_fe or (self.max_frames > 0 and t["n_frames"] >= self.max_frames)

    def rates(self, k_bits):
        t = self.total
        n = max(1, t["n_frames"])
        return {"fer": t["n_fe"] / n, "ber": t["n_be"] / (n * k_bits)}
