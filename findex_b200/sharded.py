"""
findex_b200.sharded — multi-GPU plumbing of the search path: the index is replicated on every GPU, the query batch is
cut into contiguous shards [r*m/G, (r+1)*m/G) and the only exchange step is an all-gather of the results
(SURVEY.md §8e): fixed-size counts / (sp,ep) pairs, or counts-then-values for the variable-length outputs of locate and
regex search.  One process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests); there is
no collective inside the search loop.  The compute callable is injected so the same code runs with the GPU searcher in
production and with a stand-in in the world-size-2 gloo tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(m, rank, world):
    """Contiguous shard of rank `rank`: [lo, hi)."""
    return (m * rank) // world, (m * (rank + 1)) // world


def _world(group):
    return (dist.get_world_size(group), dist.get_rank(group)) if dist.is_initialized() else (1, 0)


def allgather_fixed(local, m_global, group=None):
    """local: 1-D tensor holding this rank's shard (length hi-lo).  Returns the m_global-long gathered tensor on every rank.
    Shards may differ by one element; they are padded to the largest shard for the collective."""
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(m_global, r, world)[1] - shard_bounds(m_global, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[:local.numel()] = local
    out = torch.empty(pad * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    if all(s == pad for s in sizes):
        return out[:m_global]
    return torch.cat([out[r * pad:r * pad + sizes[r]] for r in range(world)])


def allgather_variable(local_counts, local_values, m_global, group=None):
    """Variable-length outputs (locate positions, regex triples): local_counts[q] values belong to local query q.
    Returns (off[m_global+1], values) in global query order on every rank: all-gather the counts, exclusive-scan them,
    then all-gather the value slabs padded to the largest slab."""
    world, rank = _world(group)
    counts = allgather_fixed(local_counts, m_global, group)
    off = torch.zeros(m_global + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts.to(torch.int64), 0, out=off[1:])
    if world == 1:
        return off, local_values
    slab = [int(off[shard_bounds(m_global, r, world)[1]] - off[shard_bounds(m_global, r, world)[0]]) for r in range(world)]
    pad = max(max(slab), 1)
    shape = (pad,) + tuple(local_values.shape[1:])
    buf = torch.zeros(shape, dtype=local_values.dtype, device=local_values.device)
    buf[:local_values.shape[0]] = local_values
    out = torch.empty((pad * world,) + tuple(local_values.shape[1:]), dtype=local_values.dtype, device=local_values.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return off, torch.cat([out[r * pad:r * pad + slab[r]] for r in range(world)])


class ShardedSearcher:
    """Query-sharded front end over one replicated index per rank.

    count_fn(pats2d)            -> (sp, ep) numpy int64 arrays for the local shard
    locate_fn(sp, ep)           -> (off, pos) for the local shard
    regex_fn(list_of_regexes)   -> list (per regex) of sorted (len, sp, ep) triples
    """

    def __init__(self, count_fn=None, locate_fn=None, regex_fn=None, device="cpu", group=None):
        self.count_fn, self.locate_fn, self.regex_fn = count_fn, locate_fn, regex_fn
        self.device, self.group = device, group

    @classmethod
    def for_gpu_searcher(cls, g, device, group=None):
        from . import fmindex as fx
        return cls(count_fn=g.count_fixed, locate_fn=g.locate_batch,
                   regex_fn=lambda rxs: g.regex_search_batch([fx.ReTree(r) for r in rxs]), device=device, group=group)

    def count(self, pats2d):
        """pats2d: the GLOBAL uint8 [m, len] batch (identical on every rank).  Returns global (sp, ep) on every rank."""
        world, rank = _world(self.group)
        m = pats2d.shape[0]
        lo, hi = shard_bounds(m, rank, world)
        sp, ep = self.count_fn(pats2d[lo:hi])
        both = torch.from_numpy(np.stack([sp, ep], 1).reshape(-1).astype(np.int64)).to(self.device)
        out = allgather_fixed_pairs(both, m, self.group)
        return out[:, 0].cpu().numpy(), out[:, 1].cpu().numpy()

    def locate(self, sp, ep):
        world, rank = _world(self.group)
        m = len(sp)
        lo, hi = shard_bounds(m, rank, world)
        off, pos = self.locate_fn(sp[lo:hi], ep[lo:hi])
        cnt = torch.from_numpy(np.diff(off).astype(np.int64)).to(self.device)
        goff, gpos = allgather_variable(cnt, torch.from_numpy(np.asarray(pos, np.int64)).to(self.device), m, self.group)
        return goff.cpu().numpy(), gpos.cpu().numpy()

    def regex_search(self, regexes):
        world, rank = _world(self.group)
        m = len(regexes)
        lo, hi = shard_bounds(m, rank, world)
        res = self.regex_fn(regexes[lo:hi])
        cnt = torch.tensor([len(r) for r in res], dtype=torch.int64, device=self.device)
        flat = torch.tensor([t for r in res for t in r], dtype=torch.int64, device=self.device).reshape(-1, 3)
        goff, gval = allgather_variable(cnt, flat, m, self.group)
        goff, gval = goff.cpu().numpy(), gval.cpu().numpy()
        return [[tuple(int(x) for x in row) for row in gval[goff[i]:goff[i + 1]]] for i in range(m)]


def allgather_fixed_pairs(local_flat, m_global, group=None):
    """local_flat: interleaved (sp,ep) pairs of the local shard.  Returns an [m_global, 2] tensor."""
    world, rank = _world(group)
    if world == 1:
        return local_flat.reshape(-1, 2)
    sizes = [shard_bounds(m_global, r, world)[1] - shard_bounds(m_global, r, world)[0] for r in range(world)]
    pad = max(sizes) * 2
    buf = torch.zeros(pad, dtype=local_flat.dtype, device=local_flat.device)
    buf[:local_flat.numel()] = local_flat
    out = torch.empty(pad * world, dtype=local_flat.dtype, device=local_flat.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * pad:r * pad + 2 * sizes[r]] for r in range(world)]).reshape(-1, 2)
