"""Feature stores on either side of the matcher.

Metric_Cache (query side) and Grid_Cache (target side) keep the reference's public
surface -- cache.pyx:31-138 and :151-284, cache.pxd:8-39 -- on Python 3, with the two
changes BASELINE.json's north_star asks for:
  * descriptors are held as torch uint8 tensors resident on the GPU (SIFT values are
    integers 0..255, so nothing is lost);
  * the self-match that produces every feature's distance to its nearest *other*
    feature (the denominator of all Fast-Match ratios) is the exact CUDA top-2 for both
    the thumbnail (cache.pyx:250-252) and the full image (cache.pyx:271-273, where the
    reference used approximate, non-deterministic FLANN).
Geometry, persistence and the BallTree radius lookup are host glue and stay on the CPU.
"""
import hashlib
import os
import struct

import numpy
import torch
from sklearn.neighbors import BallTree

from . import backend, imaging, matchutil


#########################################
#              Grid Cache               #
#########################################
class Grid_Cache(object):
    """Lazy per-cell store over the target image (cache.pyx:31-138).

    Note the reference's naming: `row` indexes x (columns of pixels) and `col` indexes y;
    block(x, y) returns (col, row).  Kept as is so logs and neighbour order match.
    """

    def __init__(self, data, cell_size, caching_function=None, margin=25):
        self.width = int(data.shape[1])
        self.height = int(data.shape[0])
        self.cell_width = int(cell_size[0])
        self.cell_height = int(cell_size[1])
        self.rows = int(self.width / self.cell_width) + 1
        self.cols = int(self.height / self.cell_height) + 1
        self.data = data
        self.fun = caching_function
        self.last = None
        self.margin = int(margin)
        self.grid = {n: {} for n in range(self.cols)}

    # -- geometry ---------------------------------------------------------------
    def block(self, x, y):
        row = int(x / self.cell_width)
        col = int(y / self.cell_height)
        return col, row

    def offset(self, x, y):
        """Top-left of the cell's padded crop -- unconditional -margin (cache.pyx:64-69),
        which differs from rect() on the first row/column; kept for parity."""
        col, row = self.block(x, y)
        return (row * self.cell_width - self.margin, col * self.cell_height - self.margin)

    def center(self, col, row):
        x = int((row + 0.5) * self.cell_width)
        y = int((col + 0.5) * self.cell_height)
        return numpy.array((min(x, self.width - 1), min(y, self.height - 1)), dtype=numpy.int64)

    def get_neighbor(self, col, row, pos_x, pos_y):
        """Centre of the neighbouring cell across the border `pos` is closest to, or
        (-1, -1) outside the image (cache.pyx:72-92)."""
        none = numpy.array((-1, -1), dtype=numpy.int64)
        x, y = self.center(col, row)
        x_diff = int(pos_x) - int(x)
        y_diff = int(pos_y) - int(y)
        if y_diff < x_diff and y_diff < -x_diff:
            return self.center(col - 1, row) if col - 1 >= 0 else none
        if x_diff > y_diff:
            return self.center(col, row + 1) if row + 1 < self.rows else none
        if y_diff > -x_diff:
            return self.center(col + 1, row) if col + 1 < self.cols else none
        return self.center(col, row - 1) if row - 1 >= 0 else none

    def rect(self, col, row):
        """((x_min, x_max), (y_min, y_max)) of the padded crop of a cell (cache.pyx:127-131)."""
        x_min = row * self.cell_width - (self.margin * (row > 0))
        x_max = x_min + self.cell_width + self.margin * 2 if row + 1 < self.rows else self.width
        y_min = col * self.cell_height - (self.margin * (col > 0))
        y_max = y_min + self.cell_height + self.margin * 2 if col + 1 < self.cols else self.height
        return ((x_min, x_max), (y_min, y_max))

    # -- lazy store ---------------------------------------------------------------
    def is_cached(self, x, y):
        col, row = self.block(x, y)
        return row in self.grid[col]

    def cache(self, col, row):
        (x_min, x_max), (y_min, y_max) = r = self.rect(col, row)
        cell = self.data[y_min:y_max, x_min:x_max, :]
        self.grid[col][row] = cell if self.fun is None else self.fun(cell)
        return r

    def get_cell(self, col, row):
        if row not in self.grid[col]:
            self.last = self.cache(col, row)
        return self.grid[col][row]

    def cache_many(self, cells, executor=None):
        """Cache several (col, row) cells at once; with an executor the caching function (per-cell
        SIFT, cache.pyx:124-138) runs on its threads -- cv2 releases the GIL, every call builds its
        own detector, and the result per cell is the same as a one-by-one visit.  `last` is left
        alone: a batched prefetch has no visiting order (the flood fill replays it itself)."""
        todo = [(col, row) for (col, row) in dict.fromkeys(cells) if row not in self.grid[col]]
        crops = []
        for col, row in todo:
            (x_min, x_max), (y_min, y_max) = self.rect(col, row)
            crops.append(self.data[y_min:y_max, x_min:x_max, :])
        if self.fun is None:
            results = crops
        elif executor is None or len(crops) < 2:
            results = [self.fun(c) for c in crops]
        else:
            results = list(executor.map(self.fun, crops))
        for (col, row), res in zip(todo, results):
            self.grid[col][row] = res

    def get(self, x, y):
        if x > self.width or y > self.height:
            raise Exception("(%i,%i) is outside data bounds of (%i,%i)" % (x, y, self.width, self.height))
        col, row = self.block(x, y)
        return self.get_cell(col, row)


#########################################
#             Metric Cache              #
#########################################
def _ripemd160(data):
    """RIPEMD-160 (the reference names its cache files RIPEMD-160(path), cache.pyx:193, 216).
    OpenSSL 3 builds often ship without it, and a different digest would silently miss every cache
    file the reference wrote, so a pure-Python fallback is used when hashlib has none."""
    try:
        return hashlib.new("ripemd160", data).hexdigest()
    except ValueError:
        pass
    rol = lambda x, n: ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF
    f = [lambda x, y, z: x ^ y ^ z, lambda x, y, z: (x & y) | (~x & 0xFFFFFFFF & z),
         lambda x, y, z: (x | (~y & 0xFFFFFFFF)) ^ z, lambda x, y, z: (x & z) | (y & ~z & 0xFFFFFFFF),
         lambda x, y, z: x ^ (y | (~z & 0xFFFFFFFF))]
    KL = [0x00000000, 0x5A827999, 0x6ED9EBA1, 0x8F1BBCDC, 0xA953FD4E]
    KR = [0x50A28BE6, 0x5C4DD124, 0x6D703EF3, 0x7A6D76E9, 0x00000000]
    RL = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 7, 4, 13, 1, 10, 6, 15, 3, 12, 0, 9, 5, 2, 14, 11, 8,
          3, 10, 14, 4, 9, 15, 8, 1, 2, 7, 0, 6, 13, 11, 5, 12, 1, 9, 11, 10, 0, 8, 12, 4, 13, 3, 7, 15, 14, 5, 6, 2,
          4, 0, 5, 9, 7, 12, 2, 10, 14, 1, 3, 8, 11, 6, 15, 13]
    RR = [5, 14, 7, 0, 9, 2, 11, 4, 13, 6, 15, 8, 1, 10, 3, 12, 6, 11, 3, 7, 0, 13, 5, 10, 14, 15, 8, 12, 4, 9, 1, 2,
          15, 5, 1, 3, 7, 14, 6, 9, 11, 8, 12, 2, 10, 0, 4, 13, 8, 6, 4, 1, 3, 11, 15, 0, 5, 12, 2, 13, 9, 7, 10, 14,
          12, 15, 10, 4, 1, 5, 8, 7, 6, 2, 13, 14, 0, 3, 9, 11]
    SL = [11, 14, 15, 12, 5, 8, 7, 9, 11, 13, 14, 15, 6, 7, 9, 8, 7, 6, 8, 13, 11, 9, 7, 15, 7, 12, 15, 9, 11, 7, 13, 12,
          11, 13, 6, 7, 14, 9, 13, 15, 14, 8, 13, 6, 5, 12, 7, 5, 11, 12, 14, 15, 14, 15, 9, 8, 9, 14, 5, 6, 8, 6, 5, 12,
          9, 15, 5, 11, 6, 8, 13, 12, 5, 12, 13, 14, 11, 8, 5, 6]
    SR = [8, 9, 9, 11, 13, 15, 15, 5, 7, 7, 8, 11, 14, 14, 12, 6, 9, 13, 15, 7, 12, 8, 9, 11, 7, 7, 12, 7, 6, 15, 13, 11,
          9, 7, 15, 11, 8, 6, 6, 14, 12, 13, 5, 14, 13, 13, 7, 5, 15, 5, 8, 11, 14, 14, 6, 14, 6, 9, 12, 9, 12, 5, 15, 8,
          8, 5, 12, 9, 12, 5, 14, 6, 8, 13, 6, 5, 15, 13, 11, 11]
    h = [0x67452301, 0xEFCDAB89, 0x98BADCFE, 0x10325476, 0xC3D2E1F0]
    msg = data + b"\x80" + b"\x00" * ((55 - len(data)) % 64) + struct.pack("<Q", 8 * len(data))
    for off in range(0, len(msg), 64):
        X = struct.unpack("<16I", msg[off:off + 64])
        al, bl, cl, dl, el = h
        ar, br, cr, dr, er = h
        for j in range(80):
            r = j // 16
            t = (rol((al + f[r](bl, cl, dl) + X[RL[j]] + KL[r]) & 0xFFFFFFFF, SL[j]) + el) & 0xFFFFFFFF
            al, el, dl, cl, bl = el, dl, rol(cl, 10), bl, t
            t = (rol((ar + f[4 - r](br, cr, dr) + X[RR[j]] + KR[r]) & 0xFFFFFFFF, SR[j]) + er) & 0xFFFFFFFF
            ar, er, dr, cr, br = er, dr, rol(cr, 10), br, t
        t = (h[1] + cl + dr) & 0xFFFFFFFF
        h[1] = (h[2] + dl + er) & 0xFFFFFFFF
        h[2] = (h[3] + el + ar) & 0xFFFFFFFF
        h[3] = (h[4] + al + br) & 0xFFFFFFFF
        h[4] = (h[0] + bl + cr) & 0xFFFFFFFF
        h[0] = t
    return struct.pack("<5I", *h).hex()



def self_distances(desc_dev):
    """Distance from each descriptor to its nearest *other* descriptor: slot 1 of the exact
    self top-2, as float64 holding float32 values (what `r[1].distance` gives at
    cache.pyx:252, :273).  Fewer than two features -> +inf."""
    n = desc_dev.shape[0]
    if n == 0:
        return numpy.zeros(0, numpy.float64)
    d2, _ = backend.top2(desc_dev, desc_dev)
    d2 = d2[:, 1].cpu().numpy().view(numpy.uint32)
    with numpy.errstate(invalid="ignore"):
        dist = numpy.sqrt(d2.astype(numpy.float32)).astype(numpy.float64)
    dist[d2 == 0xFFFFFFFF] = numpy.inf
    return dist


class Metric_Cache(object):
    """Query-side cache: thumbnail + full-image features, self-match distances, position
    tree (cache.pyx:151-284).  `.thumb` / `.original` keep the reference's keys; the
    `descriptors` entries are CUDA uint8 tensors."""

    def __init__(self, path, options={}):
        self.path = path
        self.thumb = {}
        self.original = {}
        self.device = matchutil._device(options.get("device"))
        if path is None:
            return
        force_reload = options.get("force_reload", False)
        max_size = options.get("max_size", -1)
        metric = options.get("metric", "minkowski")
        thumb_x, thumb_y = options.get("thumb_size", (600, 600))
        self.cache_dir = options.get("cache_dir", "data/image_data")
        if not force_reload and self.load(self.cache_dir, metric):
            return
        self.create_thumbnail(path, thumb_x, thumb_y)
        self.create_image(path, max_size, metric)
        if options.get("save", True):
            self.save(self.cache_dir)

    # -- construction from features (used by tests and by callers with their own SIFT) ---
    @classmethod
    def from_features(cls, thumb_desc, thumb_pos, thumb_size, desc, pos, size, options={}):
        self = cls(None, options)
        self._fill(self.thumb, thumb_desc, thumb_pos, thumb_size)
        self._fill(self.original, desc, pos, size)
        self.original["position_tree"] = BallTree(self.original["positions"].reshape(-1, 2),
                                                  metric=options.get("metric", "minkowski"))
        return self

    def _fill(self, slot, desc, pos, size):
        dev = matchutil.to_device(desc, self.device)
        slot["descriptors"] = dev
        slot["positions"] = numpy.asarray(pos, dtype=numpy.float64).reshape(-1, 2)
        slot["distances"] = self_distances(dev)
        slot["size"] = (int(size[0]), int(size[1]))

    def create_thumbnail(self, path, thumb_x, thumb_y):
        thumbnail = imaging.get_thumbnail(path, (thumb_x, thumb_y))
        keypoints, descriptors = matchutil.get_features(thumbnail)
        self._fill(self.thumb, descriptors, [k.pt for k in keypoints],
                   (thumbnail.shape[1], thumbnail.shape[0]))

    def create_image(self, path, max_size, metric):
        img = imaging.open_img(path, max_size)
        keypoints, descriptors = matchutil.get_features(img)
        self._fill(self.original, descriptors, [k.pt for k in keypoints], (img.shape[1], img.shape[0]))
        self.original["position_tree"] = BallTree(self.original["positions"], metric=metric)

    # -- lookups ------------------------------------------------------------------
    def get_indices(self, x, y, radius, options={}):
        """Indices of the features within `radius` px of (x, y), nearest first (the order
        decides crossCheck tie-breaks downstream) -- cache.pyx:173-186."""
        tree = self.original["position_tree"]
        indices = tree.query_radius(numpy.array([[x, y]], dtype=numpy.float64), r=radius,
                                    return_distance=True,
                                    sort_results=options.get("sort_results", True))[0]
        return indices[0]

    def get_indices_many(self, points, radius, options={}):
        """get_indices for many (x, y) at once: one tree query for a whole wave of rounds (the
        per-call validation of scikit-learn costs more than the search itself)."""
        if len(points) == 0:
            return []
        tree = self.original["position_tree"]
        ind, _ = tree.query_radius(numpy.asarray(points, dtype=numpy.float64).reshape(-1, 2), r=radius,
                                   return_distance=True, sort_results=options.get("sort_results", True))
        return list(ind)

    def get(self, x, y, radius, options={}):
        """(descriptors[idx], positions[idx], distances[idx], idx) -- cache.pyx:173-188."""
        idx = self.get_indices(x, y, radius, options)
        sel = torch.from_numpy(numpy.ascontiguousarray(idx)).to(self.device)
        return (self.original["descriptors"][sel], self.original["positions"][idx],
                self.original["distances"][idx], idx)

    # -- persistence (cache.pyx:191-239) ----------------------------------------------
    def _key(self):
        return _ripemd160(self.path.encode("utf-8") if isinstance(self.path, str) else self.path)

    def save(self, dir="data/image_data"):
        """Same two files as the reference (<key>.npz, <key>_thumb.npz), with u8 descriptors, the
        EXACT self-match distances and a flag that says so.  The position tree is not stored: it is
        rebuilt from the positions on load (milliseconds), so loading never unpickles anything."""
        key = self._key()
        if not os.path.exists(dir):
            os.makedirs(dir)
        o, t = self.original, self.thumb
        numpy.savez("%s/%s" % (dir, key), descriptors=o["descriptors"].cpu().numpy(),
                    positions=o["positions"], distances=o["distances"], size=o["size"], exact_self_match=True)
        numpy.savez("%s/%s_thumb" % (dir, key), positions=t["positions"],
                    descriptors=t["descriptors"].cpu().numpy(), distances=t["distances"], size=t["size"],
                    exact_self_match=True)
        return key

    def load(self, dir="data/image_data", metric="minkowski"):
        """Reads this class's files and the reference's layout (cache.pyx:200-210: float32
        integer-valued descriptors, `distances` from the approximate and non-deterministic FLANN
        self-match, a BallTree pickled into a 0-d object array).  Files without the
        `exact_self_match` flag get their distances recomputed by the exact CUDA self-match (one
        top-2 call each), so ratios never mix approximate denominators with exact numerators; the
        pickled tree is ignored (never unpickled) and rebuilt from the positions."""
        key = self._key()
        full, thumb = "%s/%s.npz" % (dir, key), "%s/%s_thumb.npz" % (dir, key)
        if not (os.path.isfile(full) and os.path.isfile(thumb)):
            return False
        slots = []
        for fname in (thumb, full):
            with numpy.load(fname, allow_pickle=False) as data:      # object entries are never touched
                dev = matchutil.to_device(data["descriptors"], self.device)
                exact = "exact_self_match" in data.files and bool(data["exact_self_match"])
                slots.append({"positions": numpy.asarray(data["positions"], dtype=numpy.float64).reshape(-1, 2),
                              "descriptors": dev,
                              "distances": numpy.asarray(data["distances"], dtype=numpy.float64) if exact else self_distances(dev),
                              "size": tuple(int(v) for v in data["size"])})
        self.thumb, self.original = slots
        self.original["position_tree"] = BallTree(self.original["positions"], metric=metric)
        return True
