"""fast_match_b200 -- B200-native descriptor matching for Fast-Match (arnfred/Fast-Match).

Same Python surface as the reference for the matching path (fastmatch.match,
Metric_Cache / Grid_Cache, matchutil.bf_match / flann_match), with the OpenCV
brute-force / FLANN calls replaced by hand-written sm_100a kernels behind the
C-ABI in include/fastmatch_b200.h.  There is no CPU fallback.
"""
__version__ = "0.1.0"
