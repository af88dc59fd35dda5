"""ctypes binding of libfmatch.so (include/fastmatch_b200.h) over torch CUDA tensors.

PyTorch is plumbing here: it owns device memory and streams; every matcher call goes
through the C-ABI.  There is NO CPU fallback: a missing library or a non-CUDA tensor
raises.
"""
import ctypes
import os

import torch

from . import build as _build

FM_ALGO_AUTO, FM_ALGO_MMA_SYNC, FM_ALGO_TCGEN05 = 0, 1, 2
NONE_D2 = -1  # 0xFFFFFFFF viewed as int32

_lib = None


class FastMatchError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("FM_LIB", _build.LIB)   # FM_LIB: instrumented builds (tools/ only)
        if not os.path.exists(path):
            raise FastMatchError(
                "libfmatch.so is not built (%s); run `python -m fast_match_b200.build` -- "
                "there is no CPU fallback" % path)
        L = ctypes.CDLL(path)
        vp, i64, i32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_size_t
        L.fm_version.restype = ctypes.c_int
        L.fm_last_error.restype = ctypes.c_char_p
        L.fm_device_caps.argtypes = [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 4
        L.fm_top2_workspace_bytes.argtypes = [i64, i64]
        L.fm_top2_workspace_bytes.restype = sz
        L.fm_top2_u8.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, vp, sz, ctypes.c_int, vp]
        L.fm_ratio_match_u8.argtypes = [vp, i64, vp, i64, ctypes.c_double, vp, vp, vp, vp, vp, sz, ctypes.c_int, vp]
        L.fm_ratio_f32sqrt.argtypes = [vp, i64, vp, i64, vp, i64, ctypes.c_double, vp, vp, vp]
        L.fm_grouped_workspace_bytes.argtypes = [i64, i64, i64, i32]
        L.fm_grouped_workspace_bytes.restype = sz
        L.fm_grouped_mutual_u8.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, i64, i64, i32, vp, vp, vp,
                                           vp, vp, sz, ctypes.c_int, vp]
        L.fm_merge_top2.argtypes = [vp, i32, i64, vp, vp, vp, vp]
        L.fm_top2_host_u8.argtypes = [vp, i64, vp, i64, vp, vp, vp, ctypes.c_double, vp, ctypes.c_int]
        L.fm_launch_count.restype = ctypes.c_longlong
        L.fm_profile_enable.argtypes = [ctypes.c_int]
        L.fm_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise FastMatchError("%s failed (%d): %s" % (what, rc, lib().fm_last_error().decode()))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _desc(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 2
            and t.shape[1] == 128 and t.is_contiguous()):
        raise FastMatchError("%s must be a contiguous CUDA uint8 tensor of shape [n, 128]" % name)
    return t


def device_caps(device=0):
    v = [ctypes.c_int() for _ in range(4)]
    _check(lib().fm_device_caps(int(device), *[ctypes.byref(x) for x in v]), "fm_device_caps")
    return dict(sm_major=v[0].value, sm_minor=v[1].value, sm_count=v[2].value,
                has_tcgen05=bool(v[3].value))


_ws_cache = {}


_WS_MAX_ENTRIES = 16


def _workspace(device, nbytes):
    """Grow-only scratch per (device, stream): reuse is stream-ordered.  The cache is bounded: when
    more than _WS_MAX_ENTRIES (device, stream) pairs have been seen, the oldest entry is dropped
    (torch's caching allocator keeps the block alive until the work queued on it has run)."""
    nbytes = max(int(nbytes), 16)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            ws.record_stream(torch.cuda.current_stream(device))
        ws = torch.empty(nbytes + nbytes // 4, dtype=torch.uint8, device=device)
    _ws_cache[key] = ws                      # re-inserted last: dict order = least recently used first
    while len(_ws_cache) > _WS_MAX_ENTRIES:
        _ws_cache.pop(next(iter(_ws_cache)))
    return ws


def release_workspaces():
    """Drop every cached scratch buffer (they are re-created on demand)."""
    _ws_cache.clear()


def top2(q, t, t_index_base=0, algo=FM_ALGO_AUTO, want_keys=False, out=None):
    """Exact top-2 of each row of q among rows of t.

    Returns (d2 int32 [M,2] (bit pattern of uint32; -1 = missing), idx int32 [M,2])
    and, if want_keys, keys int64 [M,2] (bit pattern of uint64 d2<<32|idx).
    """
    q, t = _desc(q, "q"), _desc(t, "t")
    M, N = q.shape[0], t.shape[0]
    dev = q.device
    if out is None:
        d2 = torch.empty((M, 2), dtype=torch.int32, device=dev)
        idx = torch.empty((M, 2), dtype=torch.int32, device=dev)
        keys = torch.empty((M, 2), dtype=torch.int64, device=dev) if want_keys else None
    else:
        d2, idx, keys = out
    L = lib()
    with torch.cuda.device(dev):         # (the plan behind the workspace size is per device)
        ws = _workspace(dev, L.fm_top2_workspace_bytes(M, N))
        _check(L.fm_top2_u8(_ptr(q), M, _ptr(t), N, int(t_index_base), _ptr(d2), _ptr(idx),
                            _ptr(keys), _ptr(ws), ws.numel(), int(algo), _stream(dev)),
               "fm_top2_u8")
    return (d2, idx, keys) if want_keys else (d2, idx)


def ratio_match(q, t, tau, algo=FM_ALGO_AUTO, want_ratio=False, out=None):
    """Ratio-Match in one call: exact top-2 with the ratio test fused into the kernel write-out.
    Returns (d2, idx, ratio float64 | None, mask bool)."""
    q, t = _desc(q, "q"), _desc(t, "t")
    M, N = q.shape[0], t.shape[0]
    dev = q.device
    if out is None:
        d2 = torch.empty((M, 2), dtype=torch.int32, device=dev)
        idx = torch.empty((M, 2), dtype=torch.int32, device=dev)
        mask = torch.empty(M, dtype=torch.uint8, device=dev)
    else:
        d2, idx, mask = out
    r = torch.empty(M, dtype=torch.float64, device=dev) if want_ratio else None
    L = lib()
    with torch.cuda.device(dev):
        ws = _workspace(dev, L.fm_top2_workspace_bytes(M, N))
        _check(L.fm_ratio_match_u8(_ptr(q), M, _ptr(t), N, float(tau), _ptr(d2), _ptr(idx), _ptr(r),
                                   _ptr(mask), _ptr(ws), ws.numel(), int(algo), _stream(dev)),
               "fm_ratio_match_u8")
    return d2, idx, r, mask.view(torch.bool)     # 0/1 bytes: a view, no kernel


def ratio(num_d2, den_d2=None, den_f32=None, tau=0.7, want_ratio=True):
    """ratio = sqrt_f32(num)/den (float64), mask = ratio < tau.  num/den may be strided views."""
    dev = num_d2.device
    M = num_d2.shape[0]
    assert num_d2.dtype == torch.int32 and num_d2.dim() == 1
    r = torch.empty(M, dtype=torch.float64, device=dev) if want_ratio else None
    m = torch.empty(M, dtype=torch.uint8, device=dev)
    if den_f32 is not None:
        assert den_f32.dtype == torch.float32 and den_f32.is_contiguous() and den_f32.shape[0] == M
        den_ptr, den_stride = None, 0
    else:
        assert den_d2.dtype == torch.int32 and den_d2.dim() == 1 and den_d2.shape[0] == M
        den_ptr, den_stride = _ptr(den_d2), (den_d2.stride(0) if M else 1)
    with torch.cuda.device(dev):
        _check(lib().fm_ratio_f32sqrt(_ptr(num_d2), num_d2.stride(0) if M else 1, den_ptr,
                                      den_stride, _ptr(den_f32), M, float(tau), _ptr(r), _ptr(m),
                                      _stream(dev)), "fm_ratio_f32sqrt")
    return r, m.view(torch.bool)


def grouped_mutual(qpool, q_off, tpool, t_off, q_gather=None, t_base=None, max_nq=None, total_q=None,
                   total_t=None, want_mutual=True, algo=FM_ALGO_AUTO):
    """G independent mutual-NN rounds in one launch (see fm_grouped_mutual_u8)."""
    qpool, tpool = _desc(qpool, "qpool"), _desc(tpool, "tpool")
    dev = qpool.device
    for o in (q_off, t_off):
        if not (o.is_cuda and o.dtype == torch.int64 and o.is_contiguous()):
            raise FastMatchError("q_off/t_off must be contiguous CUDA int64 tensors")
    G = q_off.numel() - 1
    if total_q is None or total_t is None or max_nq is None:
        qo = q_off.cpu()
        total_q, total_t = int(qo[-1]), int(t_off[-1].item())
        max_nq = int((qo[1:] - qo[:-1]).max()) if G else 0
    if q_gather is not None and not (q_gather.is_cuda and q_gather.dtype == torch.int32
                                     and q_gather.is_contiguous()):
        raise FastMatchError("q_gather must be a contiguous CUDA int32 tensor")
    if t_base is not None and not (t_base.is_cuda and t_base.dtype == torch.int64
                                   and t_base.is_contiguous() and t_base.numel() == G):
        raise FastMatchError("t_base must be a contiguous CUDA int64 tensor of G entries")
    d2 = torch.empty((total_q, 2), dtype=torch.int32, device=dev)
    idx = torch.empty((total_q, 2), dtype=torch.int32, device=dev)
    t2q = torch.empty(total_t, dtype=torch.int32, device=dev)
    mutual = torch.empty(total_q, dtype=torch.uint8, device=dev) if want_mutual else None
    L = lib()
    tpool_rows = tpool.shape[0]
    ws = _workspace(dev, L.fm_grouped_workspace_bytes(total_q, total_t, tpool_rows, G))
    with torch.cuda.device(dev):
        _check(L.fm_grouped_mutual_u8(_ptr(qpool), _ptr(q_gather), _ptr(q_off), _ptr(tpool),
                                      _ptr(t_off), _ptr(t_base), G, total_q, total_t, tpool_rows, int(max_nq), _ptr(d2),
                                      _ptr(idx), _ptr(t2q), _ptr(mutual), _ptr(ws), ws.numel(), int(algo),
                                      _stream(dev)), "fm_grouped_mutual_u8")
    return d2, idx, t2q, (mutual.view(torch.bool) if want_mutual else None)


def mutual_single(q, t):
    """crossCheck=True, k=1 for ONE (possibly large) pair of sets: (d2 int32 [M] of the nearest
    target, idx int32 [M], mutual bool [M]).  The grouped kernel gives a group to a single CTA, which
    is right for thousands of small rounds but slow for one big group (match_thumbs: 2426 x 1058 on
    one SM), so a big pair runs the dense kernel in both directions instead; ties go to the lowest
    index both ways, as in the grouped kernel and in cv2's crossCheck."""
    q, t = _desc(q, "q"), _desc(t, "t")
    M, N = q.shape[0], t.shape[0]
    dev = q.device
    if M == 0 or N == 0:
        return (torch.full((M,), NONE_D2, dtype=torch.int32, device=dev), torch.full((M,), -1, dtype=torch.int32, device=dev),
                torch.zeros(M, dtype=torch.bool, device=dev))
    if M * N < (1 << 16):
        off = torch.tensor([[0, M], [0, N]], dtype=torch.int64, device=dev)
        d2, idx, _, mutual = grouped_mutual(q, off[0], t, off[1], max_nq=M, total_q=M, total_t=N)
        return d2[:, 0], idx[:, 0], mutual
    d2, idx = top2(q, t)
    _, back = top2(t, q)
    fwd = idx[:, 0].long()
    mutual = back[fwd, 0] == torch.arange(M, device=dev, dtype=torch.int32)
    return d2[:, 0], idx[:, 0], mutual


def merge_top2(keys, want_unpacked=True):
    """keys int64 [S, M, 2] (uint64 bit patterns) -> (out_keys [M,2], d2 [M,2], idx [M,2])."""
    assert keys.is_cuda and keys.dtype == torch.int64 and keys.dim() == 3 and keys.shape[2] == 2
    keys = keys.contiguous()
    S, M, _ = keys.shape
    dev = keys.device
    out = torch.empty((M, 2), dtype=torch.int64, device=dev)
    d2 = torch.empty((M, 2), dtype=torch.int32, device=dev) if want_unpacked else None
    idx = torch.empty((M, 2), dtype=torch.int32, device=dev) if want_unpacked else None
    with torch.cuda.device(dev):
        _check(lib().fm_merge_top2(_ptr(keys), S, M, _ptr(out), _ptr(d2), _ptr(idx), _stream(dev)),
               "fm_merge_top2")
    return out, d2, idx


def top2_host(q_np, t_np, device=0, want_dist=True, tau=None, out=None):
    """numpy in / numpy out through fm_top2_host_u8 (H2D + kernels + D2H inside the call).
    tau: also return the Lowe ratio-test mask.  out: (d2, idx, dist|None, mask|None) buffers to
    reuse (pinned buffers are used directly by the library)."""
    import numpy as np
    q_np = np.ascontiguousarray(q_np, dtype=np.uint8)
    t_np = np.ascontiguousarray(t_np, dtype=np.uint8)
    if q_np.ndim != 2 or q_np.shape[1] != 128 or t_np.ndim != 2 or t_np.shape[1] != 128:
        raise FastMatchError("descriptors must have shape [n, 128]")
    M, N = len(q_np), len(t_np)
    if out is None:
        d2 = np.empty((M, 2), np.uint32)
        idx = np.empty((M, 2), np.int32)
        dist = np.empty((M, 2), np.float32) if want_dist else None
        mask = np.empty(M, np.uint8) if tau is not None else None
    else:
        d2, idx, dist, mask = out
    vp = ctypes.c_void_p
    _check(lib().fm_top2_host_u8(q_np.ctypes.data_as(vp), M, t_np.ctypes.data_as(vp), N,
                                 d2.ctypes.data_as(vp), idx.ctypes.data_as(vp),
                                 None if dist is None else dist.ctypes.data_as(vp),
                                 float(tau if tau is not None else 0.0),
                                 None if mask is None else mask.ctypes.data_as(vp), int(device)),
           "fm_top2_host_u8")
    if tau is not None:
        return d2, idx, dist, mask
    return d2, idx, dist


def launch_count():
    return int(lib().fm_launch_count())


def profile_enable(on=True):
    lib().fm_profile_enable(1 if on else 0)


def profile_read(reset=True):
    ms, n = ctypes.c_double(), ctypes.c_int()
    _check(lib().fm_profile_read(ctypes.byref(ms), ctypes.byref(n), 1 if reset else 0), "fm_profile_read")
    return ms.value, n.value
