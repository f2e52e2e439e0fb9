"""Multi-GPU plumbing: chains / members / Monte-Carlo samples are independent units,
so ranks only differ in WHICH units they own; collectives are used for diagnostics and gathers only.

One process per GPU (torchrun), backend nccl on GPUs; the same functions run on CPU tensors over gloo, which
is how tests/test_dist_cpu.py exercises them without a GPU."""
import os

import torch
import torch.distributed as td


def env_rank_world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not td.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        td.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_units, rank, world):
    """Block partition [lo, hi) of n_units over `world` ranks; sizes differ by at most one."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _allreduce(t, op=td.ReduceOp.SUM):
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        td.all_reduce(t, op=op)
    return t


def rhat(chain_mean, chain_var, n_steps):
    """Gelman-Rubin R-hat of a scalar (or vector of scalars) monitored in every chain.

    chain_mean / chain_var: [K_local, ...] per-chain mean and variance (ddof=1) over n_steps draws of the
    chains THIS rank owns.  The cross-chain sums (count, sum m, sum m^2, sum s^2) are all-reduced, so every
    rank returns the same value: 4 small vectors per call."""
    cm = chain_mean.double()
    cv = chain_var.double()
    k = torch.tensor([float(cm.shape[0])], dtype=torch.float64, device=cm.device)
    stats = torch.stack([cm.sum(0), (cm * cm).sum(0), cv.sum(0)])
    _allreduce(k)
    _allreduce(stats)
    K = k.item()
    mean_of_means = stats[0] / K
    B_over_n = (stats[1] - K * mean_of_means ** 2) / max(K - 1.0, 1.0)     # variance of the chain means
    W = stats[2] / K
    var_plus = (n_steps - 1.0) / n_steps * W + B_over_n
    return torch.sqrt(var_plus / W)


def reduce_predictive_moments(sum_y, sum_y2, count):
    """Combine per-rank sums over ensemble members into mean / variance (ddof=1) on every rank."""
    c = torch.tensor([float(count)], dtype=torch.float64, device=sum_y.device)
    s1, s2 = sum_y.double().clone(), sum_y2.double().clone()
    _allreduce(c)
    _allreduce(s1)
    _allreduce(s2)
    M = c.item()
    mean = s1 / M
    var = (s2 - M * mean ** 2) / max(M - 1.0, 1.0)
    return mean, var


def gather_to_rank0(t):
    """Concatenate a per-rank tensor along dim 0 on rank 0 (returns None elsewhere).  Shapes may differ in dim 0."""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return t
    world, rank = td.get_world_size(), td.get_rank()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    td.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    td.all_gather(bufs, pad)
    if rank != 0:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def max_over_ranks(value):
    dev = 'cuda' if (torch.cuda.is_available() and td.is_initialized() and td.get_backend() == 'nccl') else 'cpu'
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    return _allreduce(t, td.ReduceOp.MAX).item()


def barrier():
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        td.barrier()


# --------------------------------------------------------------------------------------------------------------------
# running diagnostics of a sharded chain run (north_star: "NCCL ... only to gather samples and reduce cross-chain
# diagnostics (R-hat, predictive mean and variance)")
# --------------------------------------------------------------------------------------------------------------------
class RunningDiagnostics:
    """Per-segment cross-chain diagnostics, reduced over all ranks on a SIDE stream so that the chain kernels of the next
    segment are not held up: the Gelman-Rubin R-hat of the log-posterior over the segment and the mean acceptance rate.
    Use as the `on_segment` hook of MCMCBase.run (quinn_b200/mcmc/mcmc.py); `finish()` returns the history."""

    def __init__(self, device=None):
        self.device = device
        self.stream = torch.cuda.Stream(device=device) if (device is not None and torch.device(device).type == 'cuda') else None
        self.rows = []
        self._prev_acc = None

    def __call__(self, state, rec, t_end):
        lp, acc = rec.logpost, rec.accepted
        n = lp.shape[1]
        if self.stream is None:
            self._reduce(lp, acc, n, t_end)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(lp.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self._reduce(lp, acc, n, t_end)

    def _reduce(self, lp, acc, n, t_end):
        if n >= 2:
            r = rhat(lp.mean(1), lp.var(1, unbiased=True), n)
        else:
            r = torch.full((1,), float('nan'), dtype=torch.float64, device=lp.device)
        s = torch.stack([acc.double().sum(), torch.tensor(float(acc.numel()), dtype=torch.float64, device=lp.device)])
        _allreduce(s)
        self.rows.append((int(t_end), r.reshape(-1)[:1], s))

    def finish(self):
        if self.stream is not None:
            self.stream.synchronize()
        return [dict(step=t, rhat_logpost=float(r[0].item()), accept_rate=float((s[0] / s[1]).item())) for t, r, s in self.rows]


def allreduce_sum_(t):
    """In-place SUM over ranks (no-op for a single process); returns t."""
    return _allreduce(t)


def broadcast_from_rank0(t):
    if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
        td.broadcast(t, src=0)
    return t


def world():
    """(rank, world_size) of the initialised process group, (0, 1) otherwise."""
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1
