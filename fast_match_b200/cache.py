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
import pickle

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

    def get(self, x, y):
        if x > self.width or y > self.height:
            raise Exception("(%i,%i) is outside data bounds of (%i,%i)" % (x, y, self.width, self.height))
        col, row = self.block(x, y)
        return self.get_cell(col, row)


#########################################
#             Metric Cache              #
#########################################
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
        if not force_reload and self.load(self.cache_dir):
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

    @staticmethod
    def _load_tree(data):
        try:
            raw = data["position_tree"]
            raw = raw.tobytes() if raw.dtype == numpy.uint8 else raw.item()
            tree = pickle.loads(raw)
            tree.query_radius(numpy.zeros((1, 2)), r=1.0)
            return tree
        except Exception:  # noqa: BLE001
            return BallTree(numpy.asarray(data["positions"], dtype=numpy.float64).reshape(-1, 2), metric="minkowski")

    # -- lookups ------------------------------------------------------------------
    def get_indices(self, x, y, radius, options={}):
        """Indices of the features within `radius` px of (x, y), nearest first (the order
        decides crossCheck tie-breaks downstream) -- cache.pyx:173-186."""
        tree = self.original["position_tree"]
        indices = tree.query_radius(numpy.array([[x, y]], dtype=numpy.float64), r=radius,
                                    return_distance=True,
                                    sort_results=options.get("sort_results", True))[0]
        return indices[0]

    def get(self, x, y, radius, options={}):
        """(descriptors[idx], positions[idx], distances[idx], idx) -- cache.pyx:173-188."""
        idx = self.get_indices(x, y, radius, options)
        sel = torch.from_numpy(numpy.ascontiguousarray(idx)).to(self.device)
        return (self.original["descriptors"][sel], self.original["positions"][idx],
                self.original["distances"][idx], idx)

    # -- persistence (cache.pyx:191-239) ----------------------------------------------
    def _key(self):
        try:
            h = hashlib.new("ripemd160")
        except ValueError:  # OpenSSL builds without the legacy provider
            h = hashlib.sha1()
        h.update(self.path.encode("utf-8") if isinstance(self.path, str) else self.path)
        return h.hexdigest()

    def save(self, dir="data/image_data"):
        key = self._key()
        if not os.path.exists(dir):
            os.makedirs(dir)
        o, t = self.original, self.thumb
        numpy.savez("%s/%s" % (dir, key), descriptors=o["descriptors"].cpu().numpy(),
                    positions=o["positions"], distances=o["distances"],
                    position_tree=numpy.frombuffer(pickle.dumps(o["position_tree"]), dtype=numpy.uint8),
                    size=o["size"], exact_self_match=True)
        numpy.savez("%s/%s_thumb" % (dir, key), positions=t["positions"],
                    descriptors=t["descriptors"].cpu().numpy(), distances=t["distances"], size=t["size"])
        return key

    def load(self, dir="data/image_data"):
        key = self._key()
        full, thumb = "%s/%s.npz" % (dir, key), "%s/%s_thumb.npz" % (dir, key)
        if not (os.path.isfile(full) and os.path.isfile(thumb)):
            return False
        # also reads the reference's layout (float32 integer-valued descriptors, FLANN distances,
        # BallTree pickled by an older scikit-learn): descriptors are converted exactly, a tree
        # that does not unpickle is rebuilt from the positions
        data, data_thumb = numpy.load(full, allow_pickle=True), numpy.load(thumb, allow_pickle=True)
        self.thumb = {"positions": data_thumb["positions"],
                      "descriptors": matchutil.to_device(data_thumb["descriptors"], self.device),
                      "distances": data_thumb["distances"], "size": tuple(int(v) for v in data_thumb["size"])}
        self.original = {"descriptors": matchutil.to_device(data["descriptors"], self.device),
                         "positions": data["positions"], "distances": data["distances"],
                         "position_tree": self._load_tree(data),
                         "size": tuple(int(v) for v in data["size"])}
        return True
