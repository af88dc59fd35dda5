"""Target-sharded exact top-2 across GPUs (config 5: 1M x 1M).

cv2.BFMatcher cannot even hold >= 2^18 train rows; here the target set is split row-wise
into one contiguous shard per rank, the queries are replicated, every rank runs the dense
kernel on its shard with `t_index_base` = first global row of the shard, and the per-shard
candidates -- two packed keys (d2 << 32 | global index) per query, 16 bytes -- are exchanged
with ONE all-gather (NCCL over NVLink) and reduced by fm_merge_top2.  Unsigned order on the
packed key is the lexicographic (d2, index) order, so the merged result is bit-identical to a
single-GPU run.  The ratio test runs after the merge (it needs the global second-best).

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
    world = dist.get_world_size(group) if dist.is_initialized() else 1
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


def ratio_match_sharded(q, t_shard, t_index_base, tau, group=None):
    """Ratio-Match (Classic Matching.ipynb cell 3) over a sharded target set:
    (idx [M,2] global, d2 [M,2], ratio float64 [M], mask bool [M])."""
    _, d2, idx = sharded_top2(q, t_shard, t_index_base, group=group)
    ratio, mask = backend.ratio(d2[:, 0], den_d2=d2[:, 1], tau=tau)
    return idx, d2, ratio, mask
