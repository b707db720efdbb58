"""Multi-GPU sharding of the trajectory batch: one process per GPU, one small all-reduce.

The reference is single-process.  Trajectories are i.i.d., so rank r owns the global
trajectory ids [r K/G, (r+1) K/G); the Philox counter uses the GLOBAL id, so per-trajectory results
are bitwise independent of the number of GPUs.  The only exchange step is one ``all_reduce(sum)`` of
the packed ``[gradient | statistics]`` buffer per iteration -- NCCL over NVLink on GPUs,
gloo in the CPU tests.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_bounds(K_global, rank, world_size):
    """Contiguous shard [begin, end) of K_global trajectories for ``rank``; sizes differ by at most 1."""
    base, rem = divmod(int(K_global), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@dataclass
class Shard:
    """What the rollout wrappers need to run one rank's share of a global batch."""
    K_global: int
    rank: int = 0
    world_size: int = 1
    group: object = None

    @property
    def traj_offset(self):
        return shard_bounds(self.K_global, self.rank, self.world_size)[0]

    @property
    def K_local(self):
        b, e = shard_bounds(self.K_global, self.rank, self.world_size)
        return e - b

    def all_reduce_sum(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_stats(self, stats):
        """Statistics record of the global batch: every entry is a sum over trajectories except the maximum."""
        from ._lib import ST_MAX_T
        if self.world_size > 1:
            mx = stats[ST_MAX_T].clone()
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            stats[ST_MAX_T] = mx
        return stats

    def all_gather_rows(self, row):
        """``[world_size, n]`` tensor holding every rank's ``row`` (ONE collective).  All-gather rather than all-reduce:
        the rows are then added in rank order by every rank itself, so all ranks form bit-identical sums whatever
        algorithm NCCL picks, and the maximum entry of the statistics record survives the exchange."""
        if self.world_size <= 1:
            return row.reshape(1, -1)
        out = torch.empty(self.world_size * row.numel(), dtype=row.dtype, device=row.device)
        dist.all_gather_into_tensor(out, row.contiguous().reshape(-1), group=self.group)
        return out.view(self.world_size, row.numel())

    @classmethod
    def from_env(cls, K_global, group=None):
        if dist.is_available() and dist.is_initialized():
            return cls(K_global, dist.get_rank(group), dist.get_world_size(group), group)
        return cls(K_global)


def pack_grad_and_stats(grad, stats):
    """One float64 buffer [P gradient entries | RLSDE_NSTATS statistics] so an iteration costs one collective."""
    return torch.cat([grad.to(torch.float64).reshape(-1), stats.to(torch.float64).reshape(-1)])


def unpack_grad_and_stats(buf, n_params):
    return buf[:n_params].to(torch.float32), buf[n_params:]


def reduce_gathered(rows, n_params):
    """Rank-ordered sum of all-gathered ``[gradient | statistics]`` rows -> ``(grad float32[P], stats float64[NSTATS])``."""
    from ._lib import ST_MAX_T
    acc = rows[0].clone()
    for r in range(1, rows.shape[0]):
        acc = acc + rows[r]
    acc[n_params + ST_MAX_T] = rows[:, n_params + ST_MAX_T].max()
    return acc[:n_params].to(torch.float32), acc[n_params:]


def merge_stats(records):
    """Combine statistics records of disjoint shards (sums add; the max entry takes the max)."""
    from ._lib import ST_MAX_T
    out = records[0].clone()
    for r in records[1:]:
        mx = torch.maximum(out[ST_MAX_T], r[ST_MAX_T])
        out = out + r
        out[ST_MAX_T] = mx
    return out
