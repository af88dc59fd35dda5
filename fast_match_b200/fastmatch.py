"""Fast-Match driver: thumbnail match, then a flood fill over target grid cells.

Public surface and results are those of the reference's fastmatch.pyx:32-180:

    get_matches = match(query_cache, target_img, options)
    matches = get_matches(tau)   # [(query_index, {"positions": 2x2 float64, "ratio": float})]

with the same options (thumb_size, grid_size, thumb_strategy, log, grid_margin, radius)
and the same log records (the dicts figures.visualize_log consumes).

What changed is how the per-round matcher is fed.  The reference runs one
cv2.BFMatcher(crossCheck=True).knnMatch per flood-fill round (fastmatch.pyx:161-162),
thousands of tiny launches if done naively on a GPU.  A round is a pure function of
(int query position, target cell), so this driver evaluates every round that is currently
*pending* in the depth-first iterator as one wave -- a single grouped launch
(fm_grouped_mutual_u8: query rows gathered by index from the resident descriptor pool,
target cells resident in a device pool) -- memoises the results and lets the sequential
depth-first replay consume them in exactly the reference's order.  Emission order,
de-duplication, neighbour pushes and logs are therefore identical to a round-by-round run.
"""
import collections

import numpy
import torch

from . import backend, imaging, matchutil
from .cache import Grid_Cache, Metric_Cache  # noqa: F401  (re-exported like the reference)

_EMPTY = (numpy.array([]), numpy.array([]), numpy.array([]))


def match(query_cache, target_img, options={}):
    thumb_x, thumb_y = options.get("thumb_size", (400, 400))
    grid_x, grid_y = options.get("grid_size", (50, 50))
    thumb_strategy = options.get("thumb_strategy", lambda n: n)
    log = options.get("log", None)
    grid_margin = options.get("grid_margin", 25)
    radius = options.get("radius", 100)
    features = options.get("features", matchutil.get_features)
    target_cache = Grid_Cache(target_img, (grid_x, grid_y), features, margin=grid_margin)
    thumb_positions, thumb_ratios = match_thumbs(target_img, query_cache, thumb_x=thumb_x,
                                                 thumb_y=thumb_y, features=features)
    rounds = _Rounds(query_cache, target_cache, int(radius), options.get("stats"))

    def get_matches(tau):
        thumb_tau = thumb_strategy(tau)
        seeds = thumb_positions[thumb_ratios < thumb_tau]
        return do_iter(seeds, rounds, tau=tau, log=log)

    return get_matches


# ---------------------------------------------------------------------------------------
# thumbnail round (fastmatch.pyx:107-141)
# ---------------------------------------------------------------------------------------
def _mutual_pairs(q_dev, t_dev):
    """crossCheck=True, k=1: (query rows, target rows, float32 distances), by query row."""
    M, N = q_dev.shape[0], t_dev.shape[0]
    if M == 0 or N == 0:
        return numpy.zeros(0, numpy.int64), numpy.zeros(0, numpy.int64), numpy.zeros(0, numpy.float32)
    off = torch.tensor([[0, M], [0, N]], dtype=torch.int64, device=q_dev.device)
    d2, idx, _, mutual = backend.grouped_mutual(q_dev, off[0], t_dev, off[1], max_nq=M,
                                                total_q=M, total_t=N)
    keep = torch.nonzero(mutual).flatten()
    qi = keep.cpu().numpy()
    ti = idx[keep, 0].cpu().numpy().astype(numpy.int64)
    dist = numpy.sqrt(d2[keep, 0].cpu().numpy().view(numpy.uint32).astype(numpy.float32))
    return qi, ti, dist


def match_thumbs(img, query_cache, thumb_x=400, thumb_y=400, features=matchutil.get_features):
    target = imaging.get_thumbnail(img, (thumb_x, thumb_y))
    t_orig_x, t_orig_y = imaging.get_size(img)
    t_keypoints, t_descriptors = features(target)
    q_distances = query_cache.thumb["distances"]
    q_dev = query_cache.thumb["descriptors"]
    t_dev = _to_pool(matchutil.to_u8(t_descriptors), q_dev)
    qi, ti, dist = _mutual_pairs(q_dev, t_dev)
    with numpy.errstate(divide="ignore", invalid="ignore"):
        ratios = dist.astype(numpy.float64) / q_distances[qi]
    t_pts = numpy.array([k.pt for k in t_keypoints], dtype=numpy.float64).reshape(-1, 2)
    t_ratio = numpy.array([t_orig_x / float(target.shape[1]), t_orig_y / float(target.shape[0])])
    q_ratio = numpy.array([query_cache.original["size"][0] / float(query_cache.thumb["size"][0]),
                           query_cache.original["size"][1] / float(query_cache.thumb["size"][1])])
    pos_scaled = numpy.stack([query_cache.thumb["positions"][qi] * q_ratio, t_pts[ti] * t_ratio],
                             axis=1) if len(qi) else numpy.zeros((0, 2, 2))
    order = numpy.argsort(ratios)
    return pos_scaled[order], ratios[order]


# ---------------------------------------------------------------------------------------
# rounds: memoised, evaluated a wave at a time
# ---------------------------------------------------------------------------------------
class _Rounds(object):
    """match_position (fastmatch.pyx:145-169) for many (query position, cell) pairs at once."""

    def __init__(self, query_cache, target_grid, radius, stats=None):
        self.cache = query_cache
        self.grid = target_grid
        self.radius = radius
        self.memo = {}
        self.cells = {}        # (col, row) -> (start row in pool, count, positions float64 [n,2])
        self.pool = None       # device uint8 [rows, 128]: descriptors of every fetched cell
        self.visited = set()   # cells in first-visit order of the depth-first replay
        self.last = None       # what Grid_Cache.last would be in a round-by-round run
        self.stats = stats if stats is not None else {}
        for k in ("waves", "rounds_evaluated", "launches"):
            self.stats.setdefault(k, 0)

    @staticmethod
    def key(query_pos, col, row):
        return (int(query_pos[0]), int(query_pos[1]), col, row)

    def _fetch_cells(self, wanted):
        new = [c for c in wanted if c not in self.cells]
        if not new:
            return
        like = self.cache.original["descriptors"]
        start = 0 if self.pool is None else self.pool.shape[0]
        chunks = []
        for (col, row) in new:
            kp, ds = self.grid.get_cell(col, row)
            u8 = matchutil.to_u8(ds)
            off_x = row * self.grid.cell_width - self.grid.margin    # Grid_Cache.offset
            off_y = col * self.grid.cell_height - self.grid.margin
            pos = numpy.array([[k.pt[0] + off_x, k.pt[1] + off_y] for k in kp],
                              dtype=numpy.float64).reshape(-1, 2)[:len(u8)]
            self.cells[(col, row)] = (start, len(u8), pos)
            start += len(u8)
            chunks.append(u8)
        flat = numpy.concatenate(chunks) if chunks else numpy.zeros((0, 128), numpy.uint8)
        if len(flat) or self.pool is None:
            up = _to_pool(flat, like)
            self.pool = up if self.pool is None else _pool_cat(self.pool, up)

    def evaluate(self, keys):
        """Run every round in `keys` (not yet memoised) as one grouped launch."""
        keys = [k for k in dict.fromkeys(keys) if k not in self.memo]
        if not keys:
            return
        self._fetch_cells(dict.fromkeys((k[2], k[3]) for k in keys))
        q_lists = [self.cache.get_indices(k[0], k[1], self.radius) for k in keys]
        cells = [self.cells[(k[2], k[3])] for k in keys]
        results = _run_groups(self.cache.original["descriptors"], q_lists, self.pool,
                              [c[0] for c in cells], [c[1] for c in cells])
        o = self.cache.original
        for k, q_idx, cell, (qi, ti, dist) in zip(keys, q_lists, cells, results):
            if cell[1] == 0 or len(qi) == 0:
                # reference: `target_ds == None` -> three empty arrays; no mutual pairs -> same shapes
                self.memo[k] = _EMPTY
                continue
            sel = q_idx[qi]
            with numpy.errstate(divide="ignore", invalid="ignore"):
                ratios = dist.astype(numpy.float64) / o["distances"][sel]
            positions = numpy.stack([o["positions"][sel], cell[2][ti]], axis=1)
            self.memo[k] = (positions, ratios, sel)
        self.stats["waves"] += 1
        self.stats["launches"] += 1
        self.stats["rounds_evaluated"] += len(keys)


def _to_pool(u8, like):
    """Host u8 descriptors -> resident device rows next to `like`."""
    return torch.from_numpy(numpy.ascontiguousarray(u8)).to(like.device)


def _pool_cat(pool, rows):
    return torch.cat([pool, rows])


def _run_groups(q_dev, q_lists, t_pool, t_starts, t_counts):
    """One grouped launch: for each group the mutual pairs (local query i, local target j,
    float32 distance), ordered by local query -- BFMatcher(crossCheck=True).knnMatch(k=1)."""
    dev = q_dev.device
    nq = numpy.array([len(x) for x in q_lists], dtype=numpy.int64)
    nt = numpy.asarray(t_counts, dtype=numpy.int64)
    G = len(q_lists)
    q_off = numpy.zeros(G + 1, numpy.int64)
    t_off = numpy.zeros(G + 1, numpy.int64)
    numpy.cumsum(nq, out=q_off[1:])
    numpy.cumsum(nt, out=t_off[1:])
    total_q, total_t = int(q_off[-1]), int(t_off[-1])
    gather = (numpy.concatenate(q_lists) if total_q else numpy.zeros(0)).astype(numpy.int32)
    meta = torch.from_numpy(numpy.concatenate([q_off, t_off, numpy.asarray(t_starts, numpy.int64)])).to(dev)
    d2, idx, _, mutual = backend.grouped_mutual(
        q_dev, meta[:G + 1], t_pool, meta[G + 1:2 * G + 2], q_gather=torch.from_numpy(gather).to(dev),
        t_base=meta[2 * G + 2:], max_nq=int(nq.max()) if G else 0, total_q=total_q, total_t=total_t)
    packed = torch.stack([d2[:, 0], idx[:, 0], mutual.to(torch.int32)]).cpu().numpy()
    d2h, idxh, muth = packed[0].view(numpy.uint32), packed[1], packed[2].astype(bool)
    out = []
    for g in range(G):
        sl = slice(q_off[g], q_off[g + 1])
        qi = numpy.nonzero(muth[sl])[0]
        ti = idxh[sl][qi].astype(numpy.int64)
        dist = numpy.sqrt(d2h[sl][qi].astype(numpy.float32))
        out.append((qi, ti, dist))
    return out


# ---------------------------------------------------------------------------------------
# flood fill (fastmatch.pyx:56-103, 172-180)
# ---------------------------------------------------------------------------------------
def do_iter(seeds, rounds, tau, log=None):
    grid = rounds.grid
    pending = collections.deque((numpy.asarray(p[0], dtype=numpy.float64),
                                 numpy.asarray(p[1], dtype=numpy.float64)) for p in seeds)
    matches = []
    has_matched = set()
    found_matches = {}

    def cell_key(item):
        query_pos, target_pos = item
        col, row = grid.block(target_pos[0], target_pos[1])
        query_col, query_row = grid.block(query_pos[0], query_pos[1])
        return (col, row, query_col, query_row)

    while pending:
        item = pending.popleft()
        query_pos, target_pos = item
        ck = cell_key(item)
        if ck in has_matched:
            continue
        has_matched.add(ck)
        col, row = ck[0], ck[1]
        rk = rounds.key(query_pos, col, row)
        if rk not in rounds.memo:
            # Speculate: every pending position whose (cell, query cell) key is still free will
            # either be evaluated or be pre-empted by a round on the same cell, so its cell's
            # features are needed either way; evaluate the first position per free key now.
            wave, claimed = [rk], {ck}
            for other in pending:
                ok = cell_key(other)
                if ok in has_matched or ok in claimed:
                    continue
                claimed.add(ok)
                wave.append(rounds.key(other[0], ok[0], ok[1]))
            rounds.evaluate(wave)
        result_pos, ratios, query_idx = rounds.memo[rk]
        if (col, row) not in rounds.visited:
            # Grid_Cache.last only moves when a cell is cached for the first time (cache.pyx:105);
            # cells are prefetched by waves here, so replay that bookkeeping in visit order.
            rounds.visited.add((col, row))
            rounds.last = grid.rect(col, row)
        grid.last = rounds.last
        accepted = ratios < tau
        acc_pos = result_pos[accepted]
        neighbors = get_neighbors(target_pos, acc_pos, grid)
        if neighbors:
            pending.extendleft(reversed(neighbors))
        if log is not None:
            log.append(log_round(query_pos, target_pos, result_pos, grid, ratios, tau, rounds.radius))
        for p, r, index in zip(acc_pos, ratios[accepted], query_idx[accepted]):
            p_tuple = [int(p[0, 0]), int(p[0, 1]), int(p[1, 0]), int(p[1, 1])]
            seen = found_matches.setdefault(r, [])
            if p_tuple not in seen:
                seen.append(p_tuple)
                matches.append((index, {"positions": p, "ratio": r}))
    return matches


def get_neighbors(target_pos, result_pos, target_grid):
    col, row = target_grid.block(target_pos[0], target_pos[1])
    neighbors = []
    for p_query, p_target in result_pos:
        neighbor_pos = target_grid.get_neighbor(col, row, p_target[0], p_target[1])
        if neighbor_pos[0] != -1:
            neighbors.append((numpy.asarray(p_query, dtype=numpy.float64),
                              neighbor_pos.astype(numpy.float64)))
    return neighbors


def log_round(query_pos, target_pos, result_pos, target_grid, ratios, tau, radius):
    return {"query_pos": query_pos, "target_pos": target_pos, "target_grid": target_grid.last,
            "matches": result_pos[ratios < tau], "radius": radius, "ratios": ratios[ratios < tau],
            "margin": target_grid.margin}
