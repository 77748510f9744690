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


# ------------------------------------------------------------------------------------------------ device-resident exchange (GPU)
class GpuExchange:
    """The exchange step of the multi-GPU path with everything resident in HBM: hit counts, located positions and regex
    result records of this rank's shard are STORED by kernels straight into every rank's gathered buffers (cudaMalloc'ed, shared
    through CUDA IPC, i.e. peer memory over NVLink/NVSwitch) at their scanned offsets; NCCL is left with one 4-byte all-reduce per
    exchange as the "every rank's stores have landed" barrier.  No value ever visits the host.

        counts : uint32[M]        per-query hit counts / per-regex result counts of the whole batch
        values : uint32[cap]      located positions (1 word each) or regex records {regex, len, sp, ep} (4 words each)

    `peers` may be given explicitly (lists of device pointers, this rank's own buffers included, in rank order) so that one GPU can
    play several ranks in tests; otherwise the handles are exchanged over torch.distributed."""

    def __init__(self, g, rank, world, m_global, value_words_cap, device, group=None, peers=None, barrier=None):
        from . import fmindex as fx
        self.fx, self.g, self.rank, self.world, self.M, self.cap = fx, g, rank, world, m_global, value_words_cap
        self.device, self.group = device, group
        self.counts = fx.SharedDeviceBuffer(max(m_global, 1))
        self.values = fx.SharedDeviceBuffer(max(value_words_cap, 4))
        if peers is None and world > 1:
            mine = (self.counts.export_handle(), self.values.export_handle())
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            self.count_sinks = [self.counts.ptr if r == rank else self.counts.import_peer(r, everyone[r][0]) for r in range(world)]
            self.value_sinks = [self.values.ptr if r == rank else self.values.import_peer(r, everyone[r][1]) for r in range(world)]
        elif peers is None:
            self.count_sinks, self.value_sinks = [self.counts.ptr], [self.values.ptr]
        else:
            self.count_sinks, self.value_sinks = peers
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._barrier = barrier

    def set_peers(self, count_sinks, value_sinks):
        self.count_sinks, self.value_sinks = list(count_sinks), list(value_sinks)

    def barrier(self):
        """after this, every rank's stores issued before it have landed everywhere (stream-ordered on the current stream)"""
        if self._barrier is not None:
            return self._barrier()
        if self.world > 1:
            dist.all_reduce(self.flag, group=self.group)

    def _view(self, buf, n, dtype=torch.int32):
        """torch view over a SharedDeviceBuffer (no copy)"""
        class _Ptr:
            pass
        holder = _Ptr()
        holder.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (buf.ptr, False), "version": 3}
        return torch.as_tensor(holder, device=self.device)

    def offsets(self):
        """exclusive scan of the gathered counts -> int64[M+1] on the device"""
        c = self._view(self.counts, self.M).to(torch.int64) & 0xFFFFFFFF
        off = torch.zeros(self.M + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(c, 0, out=off[1:])
        return off

    # ---- locate: count (fused gather of the counts) -> scan -> LF walks + sort -> peer stores of the position slab
    def locate_count(self, d_pat, ln, lo, hi):
        """phase 1: count this rank's shard; the hit counts land in every rank's gathered count buffer (fused into the count kernel)"""
        m = hi - lo
        st = torch.cuda.current_stream().cuda_stream
        self._sp = torch.empty(max(m, 1), dtype=torch.int32, device=self.device)
        self._ep = torch.empty(max(m, 1), dtype=torch.int32, device=self.device)
        self.g.count_fixed_dev_gather(d_pat.data_ptr(), ln, m, self._sp.data_ptr(), self._ep.data_ptr(), self.count_sinks, lo, st)

    def locate_values(self, lo, hi, scratch_cap):
        """phase 2 (after the barrier): scan of all counts, LF walks + per-query sort of this shard, peer stores of its positions at the
        scanned offset.  Returns (off[M+1] int64 on the device, number of positions of this shard)."""
        m = hi - lo
        st = torch.cuda.current_stream().cuda_stream
        off = self.offsets()
        d_off_local = torch.empty(m + 1, dtype=torch.int64, device=self.device)
        d_pos_local = torch.empty(max(scratch_cap, 1), dtype=torch.int32, device=self.device)
        total_local = self.g.locate_dev(self._sp.data_ptr(), self._ep.data_ptr(), m, d_off_local.data_ptr(), d_pos_local.data_ptr(), scratch_cap, st)
        self.fx.scatter_dev(d_pos_local.data_ptr(), total_local, self.value_sinks, 0, off.data_ptr() + 8 * lo, 1, st)
        self._keep = (off, d_pos_local)                      # alive until the stores have been issued and waited for
        return off, total_local

    def locate(self, d_pat, ln, lo, hi, scratch_cap):
        """d_pat: device uint8 tensor [hi-lo, ln] = this rank's shard of the batch.  Returns (off[M+1], number of positions of this shard);
        the positions of the WHOLE batch are then in gathered_values(off[M]) on every rank."""
        self.locate_count(d_pat, ln, lo, hi)
        self.barrier()
        out = self.locate_values(lo, hi, scratch_cap)
        self.barrier()
        return out

    def gathered_values(self, n_words):
        return self._view(self.values, n_words)

    # ---- regex: traversal + ordering on the device -> peer stores of the per-regex counts -> scan -> peer stores of the records
    def regex_count(self, rset, lo, hi, scratch_cap):
        m = hi - lo
        st = torch.cuda.current_stream().cuda_stream
        if getattr(self, "_res", None) is None or self._res.shape[0] < max(scratch_cap, 1):
            self._res = torch.empty((max(scratch_cap, 1), 4), dtype=torch.int32, device=self.device)
        if getattr(self, "_doff", None) is None or self._doff.numel() != m + 1:
            self._doff = torch.zeros(m + 1, dtype=torch.int64, device=self.device)
        d_off = self._doff
        # the library searches on a private stream of its own and returns when the results are in place: everything this stream still
        # has in flight on the scratch buffers (the previous step's peer stores read _res; the fill of a fresh d_off) must be over first
        torch.cuda.current_stream().synchronize()
        self._total = rset.search_dev(self.g, self._res.data_ptr(), scratch_cap, d_off.data_ptr())
        self._cnt = (d_off[1:] - d_off[:-1]).to(torch.int32).contiguous()
        self._res[:self._total, 0] += lo                     # batch-wide regex ids
        self.fx.scatter_dev(self._cnt.data_ptr(), m, self.count_sinks, lo, 0, 1, st)

    def regex_values(self, lo, hi):
        st = torch.cuda.current_stream().cuda_stream
        off = self.offsets()
        self.fx.scatter_dev(self._res.data_ptr(), 4 * self._total, self.value_sinks, 0, off.data_ptr() + 8 * lo, 4, st)
        self._keep = (off,)
        return off, self._total

    def regex(self, rset, lo, hi, scratch_cap):
        """this rank's shard [lo, hi) of the regex batch (rset holds exactly those).  Returns (off[M+1], results of this shard); the
        {regex, len, sp, ep} records of the WHOLE batch are then in gathered_values(4 * off[M]) on every rank."""
        self.regex_count(rset, lo, hi, scratch_cap)
        self.barrier()
        out = self.regex_values(lo, hi)
        self.barrier()
        return out

    def close_peers(self):
        """unmap the peers' buffers (every rank does this, then a barrier, before anyone frees its own)"""
        for buf in (self.counts, self.values):
            for p in list(buf.peers.values()):
                self.fx.lib().fmx_ipc_close(p)
            buf.peers = {}

    def close(self):
        self.close_peers()
        self.counts.close()
        self.values.close()

    def retire(self):
        """Done with this exchange, but keep its mappings until the process group goes away (close_retired).  Closing the last CUDA-IPC
        mapping of a peer tears down the lazily enabled peer access between the two devices, which NCCL in the same process still
        relies on: a later collective then dies with cudaErrorInvalidAddressSpace (seen on 2 x B200 when an exchange was closed between
        two legs of a run).  Exchanges therefore live as long as the process group."""
        _RETIRED.append(self)


_RETIRED = []


def close_retired(barrier=None):
    """at the very end of a run: every rank unmaps its peers' buffers, a barrier, then everybody frees its own"""
    for ex in _RETIRED:
        ex.close_peers()
    if barrier is not None:
        barrier()
    for ex in list(_RETIRED):
        (getattr(ex, "really_close", None) or ex.close)()
    del _RETIRED[:]
