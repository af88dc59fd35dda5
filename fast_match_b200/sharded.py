"""Target-sharded exact top-2 across GPUs (config 5: 1M x 1M).

cv2.BFMatcher cannot even hold >= 2^18 train rows; here the target set is split row-wise
into one contiguous shard per rank, the queries are replicated, every rank runs the dense
kernel on its shard with `t_index_base` = first global row of the shard, and the per-shard
candidates -- two packed keys (d2 << 32 | global index) per query, 16 bytes -- are exchanged
over NCCL / NVLink and reduced by fm_merge_top2.  Unsigned order on the packed key is the
lexicographic (d2, index) order, so the merged result is bit-identical to a single-GPU run.
The ratio test runs after the merge (it needs the global second-best).

Two exchanges:
  * `sharded_top2`        one all-gather; every rank merges all M queries and holds the full
                          result (S x M x 16 bytes arrive at every rank);
  * `sharded_top2_sliced` one all-to-all; rank r merges (and ratio-tests) only the queries of
                          its slice `shard_range(M, r, S)`, so M x 16 bytes arrive per rank --
                          S times less -- and the merge work is split S ways.  `gather_full`
                          re-assembles the finished slices (16 B/query) where a caller wants
                          the whole answer on every rank.

One process per GPU; torch.distributed is plumbing only.  `local_top2` / `merge` are
parameters so the host logic can be exercised on CPU (gloo) with oracle stand-ins.
"""
import torch
import torch.distributed as dist

from . import backend


def shard_range(n_rows, rank, world):
    """Contiguous, balanced row range of shard `rank` (first `n_rows % world` shards get +1)."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None):
    if not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _local_top2_keys(q, t_shard, t_index_base):
    _, _, keys = backend.top2(q, t_shard, t_index_base=t_index_base, want_keys=True)
    return keys


def _merge(gathered):
    return backend.merge_top2(gathered)


def sharded_top2(q, t_shard, t_index_base, group=None, local_top2=_local_top2_keys, merge=_merge,
                 gather_buf=None):
    """Exact global top-2 of every query row; every rank returns the full (keys, d2, idx).

    q            [M,128] u8, replicated on every rank
    t_shard      this rank's rows of the target set
    t_index_base global row index of t_shard[0]
    """
    world, _ = _world(group)
    keys = local_top2(q, t_shard, t_index_base)            # int64 [M,2] (uint64 bit patterns)
    if world == 1:
        return merge(keys.unsqueeze(0))
    M = keys.shape[0]
    if gather_buf is None or gather_buf.shape != (world, M, 2):
        gather_buf = torch.empty((world, M, 2), dtype=torch.int64, device=keys.device)
    if keys.is_cuda:
        dist.all_gather_into_tensor(gather_buf.view(world * M, 2), keys.contiguous(), group=group)
    else:  # gloo
        parts = [gather_buf[r] for r in range(world)]
        dist.all_gather(parts, keys.contiguous(), group=group)
    return merge(gather_buf)


def sharded_top2_sliced(q, t_shard, t_index_base, group=None, local_top2=_local_top2_keys, merge=_merge,
                        recv_buf=None, world_rank=None):
    """Exact global top-2 of THIS RANK'S SLICE of the query rows.

    Returns ((q_lo, q_hi), merged) with merged = merge(keys [S, q_hi - q_lo, 2]) for the rows
    shard_range(M, rank, S): rank p receives from every rank the candidates of p's rows only
    (all_to_all_single with row splits), i.e. M x 16 bytes in total instead of S x M x 16.
    """
    world, rank = world_rank if world_rank is not None else _world(group)
    keys = local_top2(q, t_shard, t_index_base)            # int64 [M,2]
    M = keys.shape[0]
    if world == 1:
        return (0, M), merge(keys.unsqueeze(0))
    splits = [shard_range(M, r, world) for r in range(world)]
    sizes = [hi - lo for lo, hi in splits]
    mine = sizes[rank]
    if recv_buf is None or recv_buf.shape != (world * mine, 2):
        recv_buf = torch.empty((world * mine, 2), dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(recv_buf, keys.contiguous(), output_split_sizes=[mine] * world,
                           input_split_sizes=sizes, group=group)
    return splits[rank], merge(recv_buf.view(world, mine, 2))


class Grid2D(object):
    """A (query groups x target shards) arrangement of the ranks: rank r = g * St + s matches the
    queries of group g (rows shard_range(M, g, Sq)) against target shard s of St (rows
    shard_range(N, s, St)); the packed-key exchange and the merge stay inside the St ranks of a
    group.  Sq = 1 is plain target sharding.  More target shards mean less memory per GPU but also
    more rows swept from no bound (every shard pays its rows' ~2 ln(shard size) exact updates) and a
    wider exchange; at 1M x 1M on 8 GPUs a 2 x 4 grid does the same work per rank as 1 x 8 with
    half-as-deep cold starts.  Build it on every rank (new_group is collective)."""

    def __init__(self, query_groups=1, backend=None):
        self.world, self.rank = _world(None)
        if self.world % query_groups:
            raise ValueError("query_groups must divide the world size")
        self.Sq, self.St = int(query_groups), self.world // int(query_groups)
        self.g, self.s = divmod(self.rank, self.St)
        self.group = None
        if self.world > 1 and self.Sq > 1:
            for g in range(self.Sq):                      # every rank creates every group, in the same order
                grp = dist.new_group(list(range(g * self.St, (g + 1) * self.St)), backend=backend)
                if g == self.g:
                    self.group = grp

    def target_range(self, N):
        return shard_range(N, self.s, self.St)

    def query_range(self, M):
        return shard_range(M, self.g, self.Sq)

    def my_rows(self, M):
        """Global query rows whose final result this rank holds after ratio_match."""
        q_lo, q_hi = self.query_range(M)
        lo, hi = shard_range(q_hi - q_lo, self.s, self.St)
        return q_lo + lo, q_lo + hi

    def top2_sliced(self, q, t_shard, N, local_top2=_local_top2_keys, merge=_merge):
        """q: ALL queries (replicated); t_shard: this rank's target rows target_range(N).
        Returns (my_rows(M), merged top-2 of those rows)."""
        M = q.shape[0]
        q_lo, q_hi = self.query_range(M)
        (lo, hi), merged = sharded_top2_sliced(q[q_lo:q_hi], t_shard, self.target_range(N)[0], group=self.group,
                                               local_top2=local_top2, merge=merge, world_rank=(self.St, self.s))
        return (q_lo + lo, q_lo + hi), merged

    def ratio_match(self, q, t_shard, N, tau, want_ratio=False):
        """(idx, d2, ratio | None, mask) of the query rows my_rows(M)."""
        _, (_, d2, idx) = self.top2_sliced(q, t_shard, N)
        ratio, mask = backend.ratio(d2[:, 0], den_d2=d2[:, 1], tau=tau, want_ratio=want_ratio)
        return idx, d2, ratio, mask


def gather_full(x_slice, M, group=None):
    """Re-assemble per-rank row slices (rows shard_range(M, r, S) of a [M, ...] tensor) on every rank."""
    world, _ = _world(group)
    if world == 1:
        return x_slice
    sizes = [shard_range(M, r, world)[1] - shard_range(M, r, world)[0] for r in range(world)]
    pad = max(sizes)
    as_bool = x_slice.dtype == torch.bool
    src = x_slice.view(torch.uint8) if as_bool else x_slice
    buf = torch.zeros((pad,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    buf[:src.shape[0]] = src
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    full = torch.cat([o[:n] for o, n in zip(out, sizes)])
    return full.view(torch.bool) if as_bool else full


def ratio_match_sharded(q, t_shard, t_index_base, tau, group=None, want_ratio=True, full=False, world_rank=None):
    """Ratio-Match (Classic Matching.ipynb cell 3) over a sharded target set.

    Returns (idx [m,2] global target rows, d2 [m,2], ratio float64 [m] | None, mask bool [m]) for
    this rank's slice of the queries (rows shard_range(M, rank, S); m = its length), or for all
    M queries on every rank when full=True."""
    M = q.shape[0]
    _, (_, d2, idx) = sharded_top2_sliced(q, t_shard, t_index_base, group=group, world_rank=world_rank)
    ratio, mask = backend.ratio(d2[:, 0], den_d2=d2[:, 1], tau=tau, want_ratio=want_ratio)
    if full:
        idx, d2, mask = gather_full(idx, M, group), gather_full(d2, M, group), gather_full(mask, M, group)
        ratio = gather_full(ratio, M, group) if ratio is not None else None
    return idx, d2, ratio, mask


# ---- reporting helpers (bench.py, tests) ------------------------------------------------------
def exchange_description(world):
    if world == 1:
        return "none (one shard)"
    return ("all_to_all_single of packed keys: every rank receives only its M/%d query slice from each shard "
            "(M x 16 B per rank), merges and ratio-tests that slice" % world)


def count_true(mask_slice, group=None):
    """Number of set entries over all ranks' slices."""
    n = mask_slice.sum().to(torch.int64)
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(n, group=group)
    return int(n.item())


def rows_of(x_slice, rows, device, M=None, group=None):
    """Rows `rows` (global query indices) of a result that is distributed in query slices."""
    world, _ = _world(group)
    if world == 1:
        return x_slice[rows.to(x_slice.device)]
    if M is None:
        n = torch.tensor([x_slice.shape[0]], dtype=torch.int64, device=x_slice.device)
        dist.all_reduce(n, group=group)
        M = int(n.item())
    return gather_full(x_slice, M, group)[rows.to(x_slice.device)]


def xor_checksum(idx_slice, group=None, distributed=True):
    """XOR of all (global row * 2 + slot, index) words of the idx result -- a cheap fingerprint that
    is independent of the number of shards.  distributed=False: `idx_slice` already is the full
    result (no collective is issued)."""
    import numpy as np
    world, _ = _world(group)
    if world > 1 and distributed:
        n = torch.tensor([idx_slice.shape[0]], dtype=torch.int64, device=idx_slice.device)
        dist.all_reduce(n, group=group)
        idx_slice = gather_full(idx_slice, int(n.item()), group)
    a = idx_slice.cpu().numpy().astype(np.int64).reshape(-1)
    mix = (a * np.int64(2654435761)) ^ (np.arange(len(a), dtype=np.int64) * np.int64(40503))
    return int(np.bitwise_xor.reduce(mix)) if len(a) else 0
