"""Sequential CPU restatement of the Fast-Match driver -- TEST INFRASTRUCTURE ONLY.

Round-by-round, exactly as the reference runs it (one matcher call per flood-fill round,
no batching), so it can check the wave-batched product driver in
fast_match_b200/fastmatch.py.  Follows
    fastmatch.pyx:32-53   match / get_matches
    fastmatch.pyx:56-89   do_iter (depth-first iterator, (cell, query cell) de-duplication)
    fastmatch.pyx:92-103  get_neighbors
    fastmatch.pyx:107-141 match_thumbs
    fastmatch.pyx:145-169 match_position
    fastmatch.pyx:172-180 log_round
    cache.pyx:31-138      Grid_Cache geometry (incl. the offset()/cache() margin asymmetry and
                          the `last` rectangle that only moves on a first visit)
The matcher is a parameter: `mutual(q_u8, t_u8) -> (query rows, target rows, float32 dist)`
with cv2.BFMatcher(NORM_L2, crossCheck=True).knnMatch(k=1) semantics; the default is the
integer oracle, `cv2_mutual` is OpenCV itself (used to generate tests/golden).
Python-2-isms of the reference are restated for Python 3 (next(), int() truncation of the
`cdef int` position variables, `is None`).
"""
import numpy as np

from . import oracle as _o


def oracle_mutual(q_u8, t_u8):
    d2, idx, t2q = _o.np_mutual(q_u8, t_u8)
    keep = _o.mutual_pairs(idx, t2q)
    return keep, idx[keep, 0].astype(np.int64), np.sqrt(d2[keep, 0].astype(np.float32))


def cv2_mutual(q_u8, t_u8):
    import cv2
    if len(q_u8) == 0 or len(t_u8) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
    ms = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).knnMatch(q_u8.astype(np.float32), t_u8.astype(np.float32), k=1)
    ms = [m[0] for m in ms if len(m) > 0]
    return (np.array([m.queryIdx for m in ms], np.int64), np.array([m.trainIdx for m in ms], np.int64),
            np.array([m.distance for m in ms], np.float32))


class RefGrid(object):
    """cache.pyx:31-138, restated."""

    def __init__(self, data, cell_size, fun, margin):
        self.width, self.height = int(data.shape[1]), int(data.shape[0])
        self.cell_width, self.cell_height = int(cell_size[0]), int(cell_size[1])
        self.rows = int(self.width / self.cell_width) + 1
        self.cols = int(self.height / self.cell_height) + 1
        self.data, self.fun, self.margin = data, fun, int(margin)
        self.last = None
        self.grid = {}

    def block(self, x, y):
        return int(y / self.cell_height), int(x / self.cell_width)   # (col, row)

    def offset(self, x, y):
        col, row = self.block(x, y)
        return row * self.cell_width - self.margin, col * self.cell_height - self.margin

    def center(self, col, row):
        x = int((row + 0.5) * self.cell_width)
        y = int((col + 0.5) * self.cell_height)
        return np.array((x if x < self.width - 1 else self.width - 1,
                         y if y < self.height - 1 else self.height - 1), dtype=np.int64)

    def get_neighbor(self, col, row, pos_x, pos_y):
        err = np.array((-1, -1), dtype=np.int64)
        x, y = self.center(col, row)
        x_diff, y_diff = int(pos_x) - x, int(pos_y) - y
        if y_diff < x_diff and y_diff < -1 * x_diff:
            return self.center(col - 1, row) if col - 1 >= 0 else err
        elif x_diff > y_diff:
            return self.center(col, row + 1) if row + 1 < self.rows else err
        elif y_diff > -1 * x_diff:
            return self.center(col + 1, row) if col + 1 < self.cols else err
        return self.center(col, row - 1) if row - 1 >= 0 else err

    def get(self, x, y):
        if x > self.width or y > self.height:
            raise Exception("outside data bounds")
        col, row = self.block(x, y)
        if (col, row) not in self.grid:
            x_min = row * self.cell_width - (self.margin * (row > 0))
            x_max = x_min + self.cell_width + self.margin * 2 if row + 1 < self.rows else self.width
            y_min = col * self.cell_height - (self.margin * (col > 0))
            y_max = y_min + self.cell_height + self.margin * 2 if col + 1 < self.cols else self.height
            self.grid[(col, row)] = self.fun(self.data[y_min:y_max, x_min:x_max, :])
            self.last = ((x_min, x_max), (y_min, y_max))
        return self.grid[(col, row)]


def _u8(desc):
    if desc is None:
        return np.zeros((0, 128), np.uint8)
    if hasattr(desc, "cpu"):
        desc = desc.cpu().numpy()
    return np.ascontiguousarray(np.asarray(desc).astype(np.uint8))


def match_thumbs(img, query_cache, thumb_x, thumb_y, features, get_thumbnail, get_size, mutual):
    target = get_thumbnail(img, (thumb_x, thumb_y))
    t_orig_x, t_orig_y = get_size(img)
    t_keypoints, t_descriptors = features(target)
    q_distances = query_cache.thumb["distances"]
    qi, ti, dist = mutual(_u8(query_cache.thumb["descriptors"]), _u8(t_descriptors))
    with np.errstate(divide="ignore", invalid="ignore"):
        ratios = np.array([float(d) for d in dist]) / q_distances[qi] if len(qi) else np.zeros(0)
    t_pos = np.array([t_keypoints[j].pt for j in ti]).reshape(-1, 2)
    t_ratio = np.array([t_orig_x / float(target.shape[1]), t_orig_y / float(target.shape[0])])
    q_pos = query_cache.thumb["positions"][qi].reshape(-1, 2)
    q_ratio = np.array([query_cache.original["size"][0] / float(query_cache.thumb["size"][0]),
                        query_cache.original["size"][1] / float(query_cache.thumb["size"][1])])
    indices = np.argsort(ratios)
    pos_scaled = np.array([(q_p * q_ratio, t_p * t_ratio) for q_p, t_p in zip(q_pos, t_pos)]).reshape(-1, 2, 2)
    return pos_scaled[indices], ratios[indices]


def match_position(pos, query_cache, grid, radius, mutual):
    query_x, query_y = int(pos[0][0]), int(pos[0][1])      # `cdef int` truncation
    target_x, target_y = int(pos[1][0]), int(pos[1][1])
    query_idx = query_cache.get_indices(query_x, query_y, radius)
    query_ds = _u8(query_cache.original["descriptors"])[query_idx]
    query_pos = query_cache.original["positions"][query_idx]
    query_dis = query_cache.original["distances"][query_idx]
    target_kp, target_ds = grid.get(target_x, target_y)
    if target_ds is None:
        return np.array([]), np.array([]), np.array([])
    offset_x, offset_y = grid.offset(target_x, target_y)
    target_pos = [np.array([k.pt[0] + offset_x, k.pt[1] + offset_y]) for k in target_kp]
    qi, ti, dist = mutual(query_ds, _u8(target_ds))
    if len(qi) == 0:
        return np.array([]), np.array([]), np.array([])
    with np.errstate(divide="ignore", invalid="ignore"):
        ratios = np.array([float(d) for d in dist]) / query_dis[qi]
    positions = np.array([(query_pos[i], target_pos[j]) for i, j in zip(qi, ti)])
    return positions, ratios, query_idx[qi]


def match(query_cache, target_img, options={}, mutual=oracle_mutual, features=None,
          get_thumbnail=None, get_size=None):
    """fastmatch.match restated; returns get_matches(tau)."""
    from fast_match_b200 import imaging, matchutil   # host glue only (PIL / cv2 SIFT), not the matcher
    features = features or options.get("features", matchutil.get_features)
    get_thumbnail = get_thumbnail or imaging.get_thumbnail
    get_size = get_size or imaging.get_size
    thumb_x, thumb_y = options.get("thumb_size", (400, 400))
    grid_x, grid_y = options.get("grid_size", (50, 50))
    thumb_strategy = options.get("thumb_strategy", lambda n: n)
    log = options.get("log", None)
    grid_margin = options.get("grid_margin", 25)
    radius = options.get("radius", 100)
    grid = RefGrid(target_img, (grid_x, grid_y), features, grid_margin)
    thumb_positions, thumb_ratios = match_thumbs(target_img, query_cache, thumb_x, thumb_y, features,
                                                 get_thumbnail, get_size, mutual)

    def get_matches(tau):
        thumb_tau = thumb_strategy(tau)
        positions = list(thumb_positions[thumb_ratios < thumb_tau])
        matches, has_matched, found = [], {}, {}
        rounds = 0
        while positions:
            query_pos, target_pos = positions.pop(0)
            col, row = grid.block(target_pos[0], target_pos[1])
            qcol, qrow = grid.block(query_pos[0], query_pos[1])
            if has_matched.get((col, row, qcol, qrow), False):
                continue
            has_matched[(col, row, qcol, qrow)] = True
            rounds += 1
            result_pos, ratios, query_idx = match_position((query_pos, target_pos), query_cache, grid, radius, mutual)
            acc = ratios < tau
            neighbors = []
            for p_query, p_target in result_pos[acc]:
                n_pos = grid.get_neighbor(col, row, p_target[0], p_target[1])
                if n_pos[0] != -1:
                    neighbors.append(np.array((p_query, n_pos)))
            positions = neighbors + positions
            if log is not None:
                log.append({"query_pos": query_pos, "target_pos": target_pos, "target_grid": grid.last,
                            "matches": result_pos[acc], "radius": radius, "ratios": ratios[acc],
                            "margin": grid.margin})
            for p, r, index in zip(result_pos[acc], ratios[acc], query_idx[acc]):
                p_tuple = [int(p[0, 0]), int(p[0, 1]), int(p[1, 0]), int(p[1, 1])]
                if p_tuple not in found.get(r, []):
                    found[r] = found.get(r, []) + [p_tuple]
                    matches.append((index, {"positions": p, "ratio": r}))
        get_matches.rounds = rounds
        return matches

    return get_matches


class RefMetricCache(object):
    """CPU stand-in for Metric_Cache (cache.pyx:151-284) in exact-denominator mode: numpy u8
    descriptors, self distances from the integer oracle (slot 1 of the exact self top-2,
    cache.pyx:250-252), BallTree radius lookup sorted by distance (cache.pyx:173-186)."""

    def __init__(self, thumb_desc, thumb_pos, thumb_size, desc, pos, size):
        from sklearn.neighbors import BallTree
        self.thumb = self._slot(thumb_desc, thumb_pos, thumb_size)
        self.original = self._slot(desc, pos, size)
        self.original["position_tree"] = BallTree(self.original["positions"], metric="minkowski")

    @staticmethod
    def _slot(desc, pos, size):
        u8 = _u8(desc)
        d2, _ = _o.np_top2(u8, u8) if len(u8) else (np.zeros((0, 2), np.uint32), None)
        dist = np.sqrt(d2[:, 1].astype(np.float32)).astype(np.float64) if len(u8) else np.zeros(0)
        if len(u8):
            dist[d2[:, 1] == _o.NONE_D2] = np.inf
        return {"descriptors": u8, "positions": np.asarray(pos, np.float64).reshape(-1, 2),
                "distances": dist, "size": (int(size[0]), int(size[1]))}

    @classmethod
    def from_image(cls, path_or_img, thumb_size=(600, 600)):
        import cv2
        from fast_match_b200 import imaging, matchutil
        img = cv2.imread(path_or_img) if isinstance(path_or_img, str) else path_or_img
        thumb = imaging.get_thumbnail(path_or_img, thumb_size)
        tk, td = matchutil.get_features(thumb)
        k, d = matchutil.get_features(img)
        return cls(td, [p.pt for p in tk], (thumb.shape[1], thumb.shape[0]), d, [p.pt for p in k],
                   (img.shape[1], img.shape[0]))

    def get_indices(self, x, y, radius, options={}):
        ind = self.original["position_tree"].query_radius(np.array([[x, y]], dtype=np.float64), r=radius,
                                                          return_distance=True, sort_results=True)[0]
        return ind[0]
