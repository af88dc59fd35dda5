"""Feature extraction and nearest-neighbour matching wrappers.

Same names and argument meaning as the reference's matchutil.py:22-67; the bodies of
bf_match / flann_match call the CUDA matcher (fm_top2_u8 / fm_grouped_mutual_u8) instead
of cv2.BFMatcher / cv2.FlannBasedMatcher.  Results come back as a MatchList that behaves
like cv2's list-of-lists of DMatch and also exposes the underlying arrays.
"""
import collections

import numpy
import torch

from . import backend

DMatch = collections.namedtuple("DMatch", ["queryIdx", "trainIdx", "imgIdx", "distance"])


# ---------------------------------------------------------------------------------
# features (host, OpenCV) -- out of the hot path
# ---------------------------------------------------------------------------------
def sift():
    import cv2
    if hasattr(cv2, "SIFT_create"):
        return cv2.SIFT_create()
    if hasattr(cv2, "xfeatures2d"):
        return cv2.xfeatures2d.SIFT_create()
    raise Exception("Can't find SIFT")


def get_features(data, feature_type="SIFT"):
    """(keypoints, descriptors float32 [n,128] or None) -- matchutil.py:31-33."""
    return sift().detectAndCompute(data, None)


def get_keypoints(data, feature_type="SIFT"):
    return sift().detect(data)


# ---------------------------------------------------------------------------------
# descriptor containers
# ---------------------------------------------------------------------------------
def to_u8(desc):
    """numpy descriptors -> contiguous uint8 [n,128].  SIFT descriptors are integer valued in
    0..255, so the conversion is exact; anything else is rejected rather than rounded."""
    if desc is None:
        return numpy.zeros((0, 128), numpy.uint8)
    a = numpy.asarray(desc)
    if a.ndim != 2 or a.shape[1] != 128:
        raise ValueError("descriptors must have shape [n, 128], got %r" % (a.shape,))
    if a.dtype == numpy.uint8:
        return numpy.ascontiguousarray(a)
    u = a.astype(numpy.uint8)
    if not numpy.array_equal(u.astype(a.dtype), a):
        raise ValueError("descriptors are not integers in 0..255; the exact u8 matcher cannot take them")
    return numpy.ascontiguousarray(u)


def to_device(desc, device=None):
    """numpy / torch descriptors -> contiguous CUDA uint8 tensor [n,128]."""
    if isinstance(desc, torch.Tensor):
        if desc.dtype != torch.uint8:
            raise ValueError("torch descriptors must be uint8")
        t = desc
        if not t.is_cuda:
            t = t.to(_device(device))
        return t.contiguous()
    return torch.from_numpy(to_u8(desc)).to(_device(device))


def _device(device):
    if device is None:
        device = "cuda:%d" % torch.cuda.current_device() if torch.cuda.is_available() else None
    if device is None:
        raise backend.FastMatchError("no CUDA device: fast_match_b200 has no CPU fallback")
    return torch.device(device)


class MatchList(collections.abc.Sequence):
    """cv2.knnMatch-shaped result: matches[i] is a list of <= k DMatch for query i.

    .indices  int32 [M, k]   (-1 = no match),  .d2 uint32-as-int64 [M, k],
    .distances float32 [M, k] = sqrt_f32(d2) (DMatch.distance),  .valid bool [M, k]
    """

    def __init__(self, indices, d2, valid):
        self.indices = indices
        self.d2 = d2
        self.valid = valid
        with numpy.errstate(invalid="ignore"):
            self.distances = numpy.sqrt(d2.astype(numpy.float32))

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return [DMatch(int(i), int(self.indices[i, c]), 0, float(self.distances[i, c]))
                for c in range(self.indices.shape[1]) if self.valid[i, c]]


def _knn(dt1, dt2, k, cross_check, device=None):
    q = to_device(dt1, device)
    t = q if dt2 is dt1 else to_device(dt2, q.device)
    M, N = q.shape[0], t.shape[0]
    if k not in (1, 2):
        raise ValueError("only k = 1 or 2 is supported by the top-2 matcher")
    if cross_check:
        d2, idx, mutual = backend.mutual_single(q, t)
        d2 = d2[:, None].cpu().numpy().view(numpy.uint32).astype(numpy.int64)
        idx = idx[:, None].cpu().numpy()
        valid = mutual.cpu().numpy()[:, None]
    else:
        d2, idx = backend.top2(q, t)
        d2 = d2[:, :k].cpu().numpy().view(numpy.uint32).astype(numpy.int64)
        idx = idx[:, :k].cpu().numpy()
        valid = idx >= 0
    return MatchList(numpy.ascontiguousarray(idx), numpy.ascontiguousarray(d2), valid)


def bf_match(dt1, dt2, k=1, options={}):
    """Exact k-NN under L2; crossCheck honoured only when k == 1 (matchutil.py:39-43).

    Same arguments as the reference wrapper, narrower domain: the CUDA matcher is a u8 top-2
    kernel, so descriptors must be 128-d with integer values 0..255 (SIFT as OpenCV emits it; float32
    arrays holding such values are converted exactly) and k must be 1 or 2.  Anything else --
    RootSIFT / normalised SIFT, SURF, ORB, k > 2 -- raises ValueError instead of being rounded or
    silently routed to a CPU matcher; keep cv2.BFMatcher for those."""
    cross_check = k == 1 and options.get("crossCheck", False) == True  # noqa: E712
    return _knn(dt1, dt2, k, cross_check, options.get("device"))


def flann_match(dt1, dt2, k=1, options={}):
    """Same signature as the reference's FLANN wrapper (matchutil.py:46-67).  The kd-forest
    parameters (algorithm / trees / checks) are accepted and ignored: the exact kernel is
    faster than the approximate index it replaces, so the result is the exact k-NN (FLANN's own
    answers differ from it on 3-15 % of the rows: see bench.py's flann_recall leg).  Same domain
    restriction as bf_match (128-d integer-valued 0..255 descriptors, k <= 2)."""
    return _knn(dt1, dt2, k, False, options.get("device"))
