"""Fast-Match driver: thumbnail match, then a flood fill over target grid cells.

Public surface and results are those of the reference's fastmatch.pyx:32-180:

    get_matches = match(query_cache, target_img, options)
    matches = get_matches(tau)   # [(query_index, {"positions": 2x2 float64, "ratio": float})]

with the same options (thumb_size, grid_size, thumb_strategy, log, grid_margin, radius)
and the same log records (the dicts figures.visualize_log consumes).

What changed is how the per-round matcher is fed.  The reference runs one
cv2.BFMatcher(crossCheck=True).knnMatch per flood-fill round (fastmatch.pyx:161-162),
thousands of tiny launches if done naively on a GPU.  A round is a pure function of
(int query position, target cell), so this driver

  * simulates the depth-first iterator ahead of the real one whenever it meets a round it has no
    result for: rounds whose results are already known are expanded exactly as the iterator
    would expand them, unknown rounds are collected (and treated as childless for the rest of
    the simulation).  Everything collected is evaluated as ONE wave -- a single grouped launch
    (fm_grouped_mutual_u8: query rows gathered by index from the resident descriptor pool,
    target cells resident in a device pool);
  * memoises the results and lets the sequential depth-first replay consume them in exactly the
    reference's order.  The first unknown round of a wave is always the one the replay needs;
    the others are speculation that is right unless an unknown round's children claim their
    (cell, query cell) key first.  Emission order, de-duplication, neighbour pushes and logs are
    therefore identical to a round-by-round run, whatever the speculation does;
  * `match_many` runs the flood fills of several image pairs in lock step and puts all their
    pending waves into the same grouped launch (query descriptors of every cache in one resident
    pool, the cells of every target image in one growing pool).

Memoised rounds, the cell pool and the cells' features stay resident across `get_matches(tau)`
calls of one `match()`.
"""
import collections
import concurrent.futures
import os
import time

import numpy
import torch

from . import backend, imaging, matchutil
from .cache import Grid_Cache, Metric_Cache  # noqa: F401  (re-exported like the reference)

_EMPTY = (numpy.array([]), numpy.array([]), numpy.array([]))
MAX_WAVE = 1024       # rounds per grouped launch (bounds the speculation, not the result)
_EXECUTORS = {}


def _executor(threads):
    """Shared thread pool for the per-cell feature extraction of a wave (None = run inline)."""
    if threads is None:
        threads = min(16, os.cpu_count() or 1)
    threads = int(threads)
    if threads <= 1:
        return None
    ex = _EXECUTORS.get(threads)
    if ex is None:
        ex = _EXECUTORS[threads] = concurrent.futures.ThreadPoolExecutor(max_workers=threads, thread_name_prefix="fm-sift")
    return ex


def match(query_cache, target_img, options={}):
    job = _Job(query_cache, target_img, options, _Pools())

    def get_matches(tau):
        return _drive([job], tau)[0]

    return get_matches


def match_many(query_caches, target_imgs, options={}):
    """Fast-Match for several (query cache, target image) pairs at once: pair i is
    query_caches[i] vs target_imgs[i] (pass the same cache object several times to match one
    query against many targets).  Returns get_matches(tau) -> [matches of pair 0, matches of
    pair 1, ...], each list identical to match(query_caches[i], target_imgs[i], options)(tau).
    `options["log"]`, if given, must be a list of lists (one per pair).  All pairs' pending rounds
    travel in the same grouped launches."""
    if len(query_caches) != len(target_imgs):
        raise ValueError("match_many: %d caches for %d target images" % (len(query_caches), len(target_imgs)))
    pools = _Pools()
    logs = options.get("log", None)
    jobs = []
    for i, (qc, img) in enumerate(zip(query_caches, target_imgs)):
        opts = dict(options)
        opts["log"] = None if logs is None else logs[i]
        jobs.append(_Job(qc, img, opts, pools))

    def get_matches(tau):
        return _drive(jobs, tau)

    return get_matches


# ---------------------------------------------------------------------------------------
# thumbnail round (fastmatch.pyx:107-141)
# ---------------------------------------------------------------------------------------
def _mutual_pairs(q_dev, t_dev):
    """crossCheck=True, k=1: (query rows, target rows, float32 distances), by query row."""
    M, N = q_dev.shape[0], t_dev.shape[0]
    if M == 0 or N == 0:
        return numpy.zeros(0, numpy.int64), numpy.zeros(0, numpy.int64), numpy.zeros(0, numpy.float32)
    d2, idx, mutual = backend.mutual_single(q_dev, t_dev)
    keep = torch.nonzero(mutual).flatten()
    qi = keep.cpu().numpy()
    ti = idx[keep].cpu().numpy().astype(numpy.int64)
    dist = numpy.sqrt(d2[keep].cpu().numpy().view(numpy.uint32).astype(numpy.float32))
    return qi, ti, dist


def match_thumbs(img, query_cache, thumb_x=400, thumb_y=400, features=matchutil.get_features, stats=None):
    target = imaging.get_thumbnail(img, (thumb_x, thumb_y))
    t_orig_x, t_orig_y = imaging.get_size(img)
    t0 = time.perf_counter()
    t_keypoints, t_descriptors = features(target)
    t1 = time.perf_counter()
    q_distances = query_cache.thumb["distances"]
    q_dev = query_cache.thumb["descriptors"]
    t_dev = _to_pool(matchutil.to_u8(t_descriptors), q_dev)
    qi, ti, dist = _mutual_pairs(q_dev, t_dev)
    if stats is not None:
        stats["sift_s"] = stats.get("sift_s", 0.0) + (t1 - t0)
        stats["matcher_s"] = stats.get("matcher_s", 0.0) + (time.perf_counter() - t1)
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
# resident pools shared by the pairs of one match() / match_many()
# ---------------------------------------------------------------------------------------
class _Pools(object):
    """Query descriptors of every cache (one tensor; rows of cache c start at q_base[c]) and the
    descriptors of every fetched target cell (one growing tensor)."""

    def __init__(self):
        self.q_pool = None
        self.q_base = {}          # id(cache) -> first row
        self.t_pool = None        # device uint8 [rows, 128]
        self.t_rows = 0

    def add_cache(self, cache):
        if id(cache) in self.q_base:
            return
        desc = cache.original["descriptors"]
        if self.q_pool is None:
            self.q_base[id(cache)] = 0
            self.q_pool = desc                       # single pair: the cache's own tensor, no copy
        else:
            self.q_base[id(cache)] = self.q_pool.shape[0]
            self.q_pool = _pool_cat(self.q_pool, desc)

    def add_cells(self, flat_u8):
        """Append host u8 rows; returns the first row they got."""
        start = self.t_rows
        if len(flat_u8) or self.t_pool is None:
            up = _to_pool(flat_u8, self.q_pool)
            self.t_pool = up if self.t_pool is None else _pool_cat(self.t_pool, up)
        self.t_rows += len(flat_u8)
        return start


def _to_pool(u8, like):
    """Host u8 descriptors -> resident device rows next to `like`."""
    return torch.from_numpy(numpy.ascontiguousarray(u8)).to(like.device)


def _pool_cat(pool, rows):
    return torch.cat([pool, rows])


def _to_dev(arr, dev):
    return torch.from_numpy(arr).to(dev)


# ---------------------------------------------------------------------------------------
# one image pair: cells, memoised rounds
# ---------------------------------------------------------------------------------------
class _Job(object):
    """State of one (query cache, target image) pair: match_position (fastmatch.pyx:145-169) results
    memoised per (int query position, cell), the target's grid cells, the thumbnail seeds."""

    def __init__(self, query_cache, target_img, options, pools):
        thumb_x, thumb_y = options.get("thumb_size", (400, 400))
        grid_x, grid_y = options.get("grid_size", (50, 50))
        self.thumb_strategy = options.get("thumb_strategy", lambda n: n)
        self.log = options.get("log", None)
        grid_margin = options.get("grid_margin", 25)
        self.radius = int(options.get("radius", 100))
        self.features = options.get("features", matchutil.get_features)
        # a wave knows all the cells it needs before it needs them, so their SIFT can run on several
        # host threads at once (the reference extracts one cell per round, fastmatch.pyx:156)
        self.executor = _executor(options.get("sift_threads"))
        self.stats = options.get("stats") if options.get("stats") is not None else {}
        for k in ("waves", "rounds_evaluated", "launches"):
            self.stats.setdefault(k, 0)
        for k in ("sift_s", "matcher_s"):
            self.stats.setdefault(k, 0.0)
        self.cache = query_cache
        self.pools = pools
        pools.add_cache(query_cache)
        self.grid = Grid_Cache(target_img, (grid_x, grid_y), self.features, margin=grid_margin)
        self.thumb_positions, self.thumb_ratios = match_thumbs(
            target_img, query_cache, thumb_x=thumb_x, thumb_y=thumb_y, features=self.features, stats=self.stats)
        self.memo = {}
        self.cells = {}        # (col, row) -> (start row in the shared pool, count, positions float64 [n,2])
        self.visited = set()   # cells in first-visit order of the depth-first replay
        self.last = None       # what Grid_Cache.last would be in a round-by-round run

    @staticmethod
    def key(query_pos, col, row):
        return (int(query_pos[0]), int(query_pos[1]), col, row)

    def fetch_cells(self, wanted):
        new = [c for c in wanted if c not in self.cells]
        if not new:
            return
        t0 = time.perf_counter()
        chunks, meta = [], []
        self.grid.cache_many(new, self.executor)
        for (col, row) in new:
            kp, ds = self.grid.get_cell(col, row)
            u8 = matchutil.to_u8(ds)
            off_x = row * self.grid.cell_width - self.grid.margin    # Grid_Cache.offset
            off_y = col * self.grid.cell_height - self.grid.margin
            pos = numpy.array([[k.pt[0] + off_x, k.pt[1] + off_y] for k in kp],
                              dtype=numpy.float64).reshape(-1, 2)[:len(u8)]
            chunks.append(u8)
            meta.append(((col, row), len(u8), pos))
        self.stats["sift_s"] += time.perf_counter() - t0
        flat = numpy.concatenate(chunks) if chunks else numpy.zeros((0, 128), numpy.uint8)
        start = self.pools.add_cells(flat)
        for cell, n, pos in meta:
            self.cells[cell] = (start, n, pos)
            start += n

    def query_lists(self, keys):
        """Indices of the query features within `radius` of each round's query position, nearest first."""
        many = getattr(self.cache, "get_indices_many", None)
        if many is not None:
            return many([(k[0], k[1]) for k in keys], self.radius)
        return [self.cache.get_indices(k[0], k[1], self.radius) for k in keys]

    def store(self, keys, q_lists, results):
        o = self.cache.original
        for k, q_idx, (qi, ti, dist) in zip(keys, q_lists, results):
            cell = self.cells[(k[2], k[3])]
            if cell[1] == 0 or len(qi) == 0:
                # reference: `target_ds == None` -> three empty arrays; no mutual pairs -> same shapes
                self.memo[k] = _EMPTY
                continue
            sel = q_idx[qi]
            with numpy.errstate(divide="ignore", invalid="ignore"):
                ratios = dist.astype(numpy.float64) / o["distances"][sel]
            positions = numpy.stack([o["positions"][sel], cell[2][ti]], axis=1)
            self.memo[k] = (positions, ratios, sel)


def _evaluate(requests):
    """requests: [(job, [round keys not memoised yet])] -> ONE grouped launch over all of them."""
    requests = [(job, [k for k in dict.fromkeys(keys) if k not in job.memo]) for job, keys in requests]
    requests = [(job, keys) for job, keys in requests if keys]
    if not requests:
        return
    for job, keys in requests:
        job.fetch_cells(dict.fromkeys((k[2], k[3]) for k in keys))      # (SIFT of new cells: host, timed as sift_s)
    t0 = time.perf_counter()
    pools = requests[0][0].pools
    q_all, starts, counts, lists_per_job = [], [], [], []
    for job, keys in requests:
        q_lists = job.query_lists(keys)
        lists_per_job.append(q_lists)
        base = pools.q_base[id(job.cache)]
        q_all.extend(q_lists if base == 0 else [q + base for q in q_lists])
        for k in keys:
            cell = job.cells[(k[2], k[3])]
            starts.append(cell[0])
            counts.append(cell[1])
    results = _run_groups(pools.q_pool, q_all, pools.t_pool, starts, counts)
    pos = 0
    dt = time.perf_counter() - t0
    n_all = sum(len(keys) for _, keys in requests)
    for (job, keys), q_lists in zip(requests, lists_per_job):
        job.store(keys, q_lists, results[pos:pos + len(keys)])
        pos += len(keys)
        job.stats["waves"] += 1
        job.stats["launches"] += 1          # (one launch serves every job of the wave)
        job.stats["rounds_evaluated"] += len(keys)
        job.stats["matcher_s"] += dt * len(keys) / n_all


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
    # one H2D copy for all the metadata of the wave: offsets, cell starts and (as int32 pairs) the gather list
    gpad = numpy.zeros(total_q + (total_q & 1), numpy.int32)
    gpad[:total_q] = gather
    meta = _to_dev(numpy.concatenate([q_off, t_off, numpy.asarray(t_starts, numpy.int64), gpad.view(numpy.int64)]), dev)
    d2, idx, _, mutual = backend.grouped_mutual(
        q_dev, meta[:G + 1], t_pool, meta[G + 1:2 * G + 2], q_gather=meta[3 * G + 2:].view(torch.int32)[:total_q],
        t_base=meta[2 * G + 2:3 * G + 2], max_nq=int(nq.max()) if G else 0, total_q=total_q, total_t=total_t)
    packed = torch.stack([d2[:, 0], idx[:, 0], mutual.to(torch.int32)]).cpu().numpy()
    d2h, idxh, muth = packed[0].view(numpy.uint32), packed[1], packed[2].astype(bool)
    dist_all = numpy.sqrt(d2h.astype(numpy.float32))
    out = []
    for g in range(G):
        lo, hi = q_off[g], q_off[g + 1]
        qi = numpy.nonzero(muth[lo:hi])[0]
        out.append((qi, idxh[lo:hi][qi].astype(numpy.int64), dist_all[lo:hi][qi]))
    return out


# ---------------------------------------------------------------------------------------
# flood fill (fastmatch.pyx:56-103, 172-180)
# ---------------------------------------------------------------------------------------
def _cell_key(grid, query_pos, target_pos):
    col, row = grid.block(target_pos[0], target_pos[1])
    query_col, query_row = grid.block(query_pos[0], query_pos[1])
    return (col, row, query_col, query_row)


def _item(grid, query_pos, target_pos):
    """A pending position with its de-duplication key and round key computed once."""
    q = numpy.asarray(query_pos, dtype=numpy.float64)
    t = numpy.asarray(target_pos, dtype=numpy.float64)
    ck = _cell_key(grid, q, t)
    return (q, t, ck, (int(q[0]), int(q[1]), ck[0], ck[1]))


def _children(job, item, result_pos, ratios, tau):
    """Neighbour pushes of a round, in the iterator's order (get_neighbors, fastmatch.pyx:92-103)."""
    acc_pos = result_pos[ratios < tau]
    col, row = item[2][0], item[2][1]
    out = []
    for p_query, p_target in acc_pos:
        neighbor_pos = job.grid.get_neighbor(col, row, p_target[0], p_target[1])
        if neighbor_pos[0] != -1:
            out.append(_item(job.grid, p_query, neighbor_pos.astype(numpy.float64)))
    return out


def _speculate(job, first, pending, has_matched, tau, child_cache):
    """The rounds the iterator will need, as far as the memoised results can tell: simulate it from
    the current state; known rounds are expanded exactly, unknown ones are collected and treated
    as childless.  `first` (an unknown round) always comes first."""
    wave = [first[3]]
    seen = set(has_matched)
    seen.add(first[2])
    stack = list(pending)              # index 0 = next to pop, like the deque
    stack.reverse()                    # pop() from the end = popleft()
    while stack and len(wave) < MAX_WAVE:
        it = stack.pop()
        ck, rk = it[2], it[3]
        if ck in seen:
            continue
        seen.add(ck)
        res = job.memo.get(rk)
        if res is None:
            wave.append(rk)
            continue
        kids = child_cache.get(rk)
        if kids is None:
            kids = child_cache[rk] = _children(job, it, res[0], res[1], tau)
        stack.extend(reversed(kids))
    return wave


def _do_iter(job, seeds, tau):
    """Generator form of do_iter: yields the wave of round keys it wants evaluated whenever the
    replay meets an unknown round; returns the match list."""
    grid, log = job.grid, job.log
    pending = collections.deque(_item(grid, p[0], p[1]) for p in seeds)
    matches = []
    has_matched = set()
    found_matches = {}
    child_cache = {}
    while pending:
        item = pending.popleft()
        query_pos, target_pos, ck, rk = item
        if ck in has_matched:
            continue
        has_matched.add(ck)
        col, row = ck[0], ck[1]
        if rk not in job.memo:
            yield _speculate(job, item, pending, has_matched, tau, child_cache)
        result_pos, ratios, query_idx = job.memo[rk]
        if (col, row) not in job.visited:
            # Grid_Cache.last only moves when a cell is cached for the first time (cache.pyx:105);
            # cells are prefetched by waves here, so replay that bookkeeping in visit order.
            job.visited.add((col, row))
            job.last = grid.rect(col, row)
        grid.last = job.last
        accepted = ratios < tau
        acc_pos = result_pos[accepted]
        neighbors = child_cache.get(rk)
        if neighbors is None:
            neighbors = child_cache[rk] = _children(job, item, result_pos, ratios, tau)
        if neighbors:
            pending.extendleft(reversed(neighbors))
        if log is not None:
            log.append(log_round(query_pos, target_pos, result_pos, grid, ratios, tau, job.radius))
        for p, r, index in zip(acc_pos, ratios[accepted], query_idx[accepted]):
            p_tuple = [int(p[0, 0]), int(p[0, 1]), int(p[1, 0]), int(p[1, 1])]
            seen = found_matches.setdefault(r, [])
            if p_tuple not in seen:
                seen.append(p_tuple)
                matches.append((index, {"positions": p, "ratio": r}))
    return matches


def _drive(jobs, tau):
    """Run the flood fills of all jobs in lock step; every step's waves share one launch."""
    gens, results = [], [None] * len(jobs)
    for i, job in enumerate(jobs):
        thumb_tau = job.thumb_strategy(tau)
        seeds = job.thumb_positions[job.thumb_ratios < thumb_tau]
        gens.append(_do_iter(job, seeds, tau))
    waiting = {}
    for i, g in enumerate(gens):
        try:
            waiting[i] = next(g)
        except StopIteration as stop:
            results[i] = stop.value
    while waiting:
        _evaluate([(jobs[i], keys) for i, keys in waiting.items()])
        nxt = {}
        for i in waiting:
            try:
                nxt[i] = next(gens[i])
            except StopIteration as stop:
                results[i] = stop.value
        waiting = nxt
    return results


def do_iter(seeds, job, tau, log=None):
    """fastmatch.pyx:56-89 for one pair (kept for callers of the reference's function)."""
    job.log = log
    g = _do_iter(job, seeds, tau)
    try:
        wave = next(g)
        while True:
            _evaluate([(job, wave)])
            wave = next(g)
    except StopIteration as stop:
        return stop.value


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
