"""Geometric evaluation of a match list against a ground-truth homography.

The reference ships the Oxford "graffiti" homographies (images/graf/H1to{2..6}p) but no code
reads them; its own evaluation (turntable.py:27-80) needs a
data set that is not shipped.  This is the same precision-style measure on what is shipped:
a match (p_query, p_target) is an inlier when the target point, mapped through H into the query
image, lands within `tol` pixels of the query point.
"""
import numpy


def load_homography(path):
    return numpy.loadtxt(path, dtype=numpy.float64).reshape(3, 3)


def project(H, pts):
    pts = numpy.asarray(pts, dtype=numpy.float64).reshape(-1, 2)
    ph = numpy.concatenate([pts, numpy.ones((len(pts), 1))], axis=1) @ H.T
    return ph[:, :2] / ph[:, 2:3]


def inlier_fraction(matches, H_target_to_query, tol=5.0):
    """matches: the list fastmatch.match(...)(tau) returns, or an [n,2,2] array of
    [[qx,qy],[tx,ty]].  Returns (inliers, total, fraction)."""
    if len(matches) == 0:
        return 0, 0, float("nan")
    if isinstance(matches, (list, tuple)):
        pos = numpy.array([m[1]["positions"] for m in matches], dtype=numpy.float64)
    else:
        pos = numpy.asarray(matches, dtype=numpy.float64)
    err = numpy.linalg.norm(project(H_target_to_query, pos[:, 1]) - pos[:, 0], axis=1)
    ok = int((err <= tol).sum())
    return ok, len(pos), ok / float(len(pos))


def ratio_match_positions(q_desc, q_pos, t_desc, t_pos, tau, device=None):
    """Ratio-Match (Classic Matching.ipynb cell 3) on the CUDA matcher -> [n,2,2] positions of
    the matches passing d1/d2 < tau, sorted by ratio."""
    from . import matchutil
    ml = matchutil.bf_match(q_desc, t_desc, k=2, options={"device": device})
    with numpy.errstate(divide="ignore", invalid="ignore"):
        ratios = ml.distances[:, 0].astype(numpy.float64) / ml.distances[:, 1].astype(numpy.float64)
    keep = numpy.nonzero(ml.valid[:, 1] & (ratios < tau))[0]
    keep = keep[numpy.argsort(ratios[keep])]
    q_pos, t_pos = numpy.asarray(q_pos, numpy.float64), numpy.asarray(t_pos, numpy.float64)
    return numpy.stack([q_pos[keep], t_pos[ml.indices[keep, 0]]], axis=1), ratios[keep]
