#!/usr/bin/env python
"""bench.py -- headline benchmark of the descriptor-matching hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the reference's CPU matcher)

Workload (BASELINE.json configs[2], "c3"): synthetic 10-MP pair, 50 000 x 50 000 SIFT-like u8
128-d descriptors, exact top-2 + Lowe ratio test (tau 0.7).  A step = one full pass over one
pair.  With N GPUs every rank matches its own pair (independent image pairs: no data-path
collective, weak scaling); `value` = query descriptors matched per second over all ranks.
Extra legs on the same JSON line: "grouped" (configs[3], 10k Fast-Match cell rounds in one
launch, HBM roofline) and "sharded" (configs[4], 1M x 1M with the target set sharded over the
N ranks + NCCL all-gather + merge; strong scaling).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_C3 = N_C3 = 50000
TAU = 0.7
METRIC = "ratio-matched query descriptors/sec"
UNIT = "query descriptors/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the grouped / sharded legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sharded-n", type=int, default=1000000)
    ap.add_argument("--groups", type=int, default=10000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if isinstance(uuid, str) else uuid)
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------
# the reference's CPU matcher (cv2.BFMatcher, as matchutil.py:39-43 calls it) or the C port
# ------------------------------------------------------------------------------------------
def cpu_matcher():
    """Returns (kind, cores, fn(q_u8, t_u8) -> None)."""
    try:
        import cv2
        import numpy as np
        cv2.setNumThreads(os.cpu_count() or 1)
        cores = cv2.getNumThreads()

        def run(q, t):
            qf, tf = q.astype(np.float32), t.astype(np.float32)   # the reference's dtype
            t0 = time.perf_counter()
            m = cv2.BFMatcher(cv2.NORM_L2, crossCheck=False)
            parts = []
            for lo in range(0, len(tf), 200000):                   # BFMatcher asserts < 2^18 train rows
                parts.append(m.knnMatch(qf, tf[lo:lo + 200000], k=2))
            return time.perf_counter() - t0
        return "reference", cores, run
    except Exception:
        import oracle

        def run(q, t):
            t0 = time.perf_counter()
            oracle.c_top2(q, t)
            return time.perf_counter() - t0
        return "port", os.cpu_count() or 1, run


def cpu_baseline(sample_queries=16384):
    from fast_match_b200 import synth
    kind, cores, run = cpu_matcher()
    q, t = synth.make_pair(sample_queries, N_C3, seed=1237)
    run(q[:512], t[:4096])                       # warm the thread pool
    dt = run(q, t)
    return {"value": sample_queries / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d of the %d queries x all %d targets, cv2.BFMatcher(NORM_L2).knnMatch(k=2) on float32 "
                      "(matcher only, no DMatch unpacking), %.2f s" % (sample_queries, M_C3, N_C3, dt)
            if kind == "reference" else
            "%d of the %d queries x all %d targets, oracle.c top-2 (OpenMP), %.2f s" % (sample_queries, M_C3, N_C3, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fast_match_b200 import synth
    kind, cores, run = cpu_matcher()
    sample = 8192
    q, t = synth.make_pair(sample, N_C3, seed=1237)
    for _ in range(max(args.warmup, 1)):
        run(q[:1024], t)
    times = [run(q, t) for _ in range(args.steps)]
    total = sum(times)
    value = sample * args.steps / total
    what = ("cv2.BFMatcher(NORM_L2).knnMatch(k=2), float32 descriptors" if kind == "reference"
            else "oracle.c top-2 (OpenMP)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c3: 50000x50000 u8 128-d, exact top-2 + ratio 0.7; each step = %d-query sample "
                                   "x all 50000 targets on the host CPU (%s)" % (sample, what)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d queries x 50000 targets per step" % sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# synthetic data on the device (large legs; deterministic per 65536-row chunk)
# ------------------------------------------------------------------------------------------
def siftlike_torch(lo, hi, seed, device):
    import torch
    out = torch.empty((hi - lo, 128), dtype=torch.uint8, device=device)
    chunk = 65536
    pos = 0
    shape = torch.full((chunk, 128), 0.6, device=device)
    while lo + pos < hi:
        c = (lo + pos) // chunk
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + c)
        x = torch._standard_gamma(shape, generator=g)     # Gamma(0.6) magnitudes ~ SIFT statistics
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        x = x.clamp_max(0.2)
        x = x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        rows = (512.0 * x).round().clamp(0, 255).to(torch.uint8)
        a = (lo + pos) - c * chunk
        n = min(chunk - a, hi - lo - pos)
        out[pos:pos + n] = rows[a:a + n]
        pos += n
    return out


def plant_torch(q, t, t_lo, seed):
    """Overwrite ~half of the target rows with noisy copies of pseudo-randomly chosen queries.
    Deterministic per 65536-row chunk of the *global* target index, so the data set does not
    depend on how it is sharded."""
    import torch
    chunk = 65536
    n = t.shape[0]
    pos = 0
    while pos < n:
        c = (t_lo + pos) // chunk
        a = (t_lo + pos) - c * chunk
        m = min(chunk - a, n - pos)
        g = torch.Generator(device=t.device)
        g.manual_seed(seed * 7919 + c)
        pick = torch.rand(chunk, generator=g, device=t.device) < 0.5
        src = torch.randint(0, q.shape[0], (chunk,), generator=g, device=t.device)
        sigma = torch.rand((chunk, 1), generator=g, device=t.device) * 36.0 + 4.0
        noise = torch.randn((chunk, 128), generator=g, device=t.device) * sigma
        planted = (q[src].float() + noise).round().clamp(0, 255).to(torch.uint8)
        view = t[pos:pos + m]
        sel = pick[a:a + m]
        view[sel] = planted[a:a + m][sel]
        pos += m
    return t


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fast_match_b200 import backend, sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    backend.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- headline: c3, one pair per rank, inputs resident in HBM ----------------
    q_np, t_np = synth.make_pair(M_C3, N_C3, seed=1237 + rank)
    q_dev, t_dev = torch.from_numpy(q_np).to(dev), torch.from_numpy(t_np).to(dev)
    out = (torch.empty((M_C3, 2), dtype=torch.int32, device=dev), torch.empty((M_C3, 2), dtype=torch.int32, device=dev),
           torch.empty(M_C3, dtype=torch.uint8, device=dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def step():   # exact top-2 + Lowe ratio test, fused (fm_ratio_match_u8)
        return backend.ratio_match(q_dev, t_dev, TAU, out=out)[3]

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    backend.profile_enable(True)
    backend.profile_read(reset=True)
    launches0 = backend.launch_count()
    evs = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()                                   # untimed: evict the inputs from L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mask = step()
        b.record()
        evs.append((a, b))
    barrier()
    wall = time.perf_counter() - wall0
    launches = backend.launch_count() - launches0
    kern_ms, kern_n = backend.profile_read(reset=True)
    backend.profile_enable(False)
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    total_ms = max_over_ranks(dev_ms)
    value = world * M_C3 * args.steps / (total_ms * 1e-3)
    matched = int(mask.sum().item())

    # ---------------- end to end: host buffers through the C-ABI host entry point ----------------
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
    q_pin, t_pin = pin((M_C3, 128), torch.uint8), pin((N_C3, 128), torch.uint8)
    q_pin.copy_(torch.from_numpy(q_np)); t_pin.copy_(torch.from_numpy(t_np))
    o_d2, o_idx, o_mask = pin((M_C3, 2), torch.int32), pin((M_C3, 2), torch.int32), pin((M_C3,), torch.uint8)
    host_out = (o_d2.numpy().view(np.uint32), o_idx.numpy(), None, o_mask.numpy())
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        backend.top2_host(q_pin.numpy(), t_pin.numpy(), device=local, want_dist=False, tau=TAU, out=host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        backend.top2_host(q_pin.numpy(), t_pin.numpy(), device=local, want_dist=False, tau=TAU, out=host_out)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert int(o_mask.numpy().sum()) == matched, "host path and device path disagree"
    e2e = {"value": world * M_C3 * e2e_steps / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": (M_C3 + N_C3) * 128, "d2h_bytes_per_step": M_C3 * (8 + 8 + 1),
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
           "api": "fm_top2_host_u8 (pinned host buffers -> H2D -> kernels -> D2H -> sync)"}

    # ---------------- roofline of the dominant kernel ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops")
    peak_tops = 2.0 * bf16 if bf16 else 2.0 * 1590.0
    ops = 2.0 * M_C3 * N_C3 * 128
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    k_ms = kern_ms / max(kern_n, 1)
    achieved = ops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "k_top2_tc_pair", "achieved": achieved, "peak": peak_tops, "unit": "TOP/s",
                "frac": achieved / peak_tops,
                "traffic": (traffic.get("k_top2_tc_c3_50k") or {}).get("bytes"),
                "traffic_note": "DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json); "
                                "the kernel is tensor-bound, its algorithmic HBM bytes are (M+N)*128 + 16*M = 13.6 MB",
                "peak_source": ("2 x bf16_tflops of MEASURED_PEAKS.json (u8 tensor rate is twice bf16; the file has no "
                                "int8 entry)" if bf16 else "2 x 1.59 PFLOP/s fallback"),
                "frac_of_datasheet_4500": achieved / 4500.0, "kernel_ms": k_ms, "kernel_launches_timed": kern_n,
                "algorithmic_ops_per_launch": ops}
    try:   # same-run measured int8 GEMM throughput, for context
        a8 = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev)
        b8 = torch.randint(-8, 8, (8192, 8192), dtype=torch.int8, device=dev).t()
        for _ in range(3):
            torch._int_mm(a8, b8)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(10):
            s.record(); torch._int_mm(a8, b8); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        roofline["int8_gemm_8192_tops_same_run"] = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        del a8, b8
    except Exception as ex:  # noqa: BLE001
        roofline["int8_gemm_8192_tops_same_run"] = None

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "c3: synthetic 10-MP pair, 50000x50000 u8 128-d SIFT-like descriptors, exact top-2 + "
                                   "Lowe ratio test tau=0.7, one pair per GPU (BASELINE.json configs[2])",
                       "M": M_C3, "N": N_C3, "tau": TAU, "pairs_per_step": world,
                       "l2": "256 MiB buffer written between timed steps (inputs are 12.8 MB < L2)",
                       "matched_queries": matched},
            "pairs_per_s": world * args.steps / (total_ms * 1e-3),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "wall_s_timed_region": wall}

    # ---------------- extra leg: grouped launch (configs[3]) ----------------
    if not args.no_extra:
        try:
            line["grouped"] = grouped_leg(args, dev, peaks, backend)
        except Exception as ex:  # noqa: BLE001
            line["grouped"] = {"error": repr(ex)}
        try:
            line["sharded"] = sharded_leg(args, dev, world, rank, backend, sharded, barrier, max_over_ranks)
        except Exception as ex:  # noqa: BLE001
            line["sharded"] = {"error": repr(ex)}
        if rank == 0 and world == 1:
            try:
                line["fastmatch_readme"] = fastmatch_leg(dev)
            except Exception as ex:  # noqa: BLE001
                line["fastmatch_readme"] = {"error": repr(ex)}

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _traffic(key):
    try:
        return (json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(key) or {}).get("bytes")
    except Exception:
        return None


def grouped_leg(args, dev, peaks, backend):
    import torch
    G = args.groups
    g = torch.Generator(device="cpu"); g.manual_seed(1238)
    nq = torch.randint(32, 513, (G,), generator=g)
    nt = torch.randint(32, 513, (G,), generator=g)
    q_off = torch.zeros(G + 1, dtype=torch.int64); q_off[1:] = torch.cumsum(nq, 0)
    t_off = torch.zeros(G + 1, dtype=torch.int64); t_off[1:] = torch.cumsum(nt, 0)
    Q, T = int(q_off[-1]), int(t_off[-1])
    qpool = siftlike_torch(0, Q, 11, dev)
    tpool = siftlike_torch(0, T, 12, dev)
    # plant: half of every group's targets are noisy copies of queries of the same group
    gid = torch.repeat_interleave(torch.arange(G), nt).to(dev)
    gg = torch.Generator(device=dev); gg.manual_seed(1238)
    src = (q_off.to(dev)[gid] + (torch.rand(T, generator=gg, device=dev) * nq.to(dev)[gid]).long()).clamp_max(Q - 1)
    pick = torch.rand(T, generator=gg, device=dev) < 0.5
    sigma = torch.rand((T, 1), generator=gg, device=dev) * 36.0 + 4.0
    planted = (qpool[src].float() + torch.randn((T, 128), generator=gg, device=dev) * sigma).round().clamp(0, 255).to(torch.uint8)
    tpool[pick] = planted[pick]
    del planted
    qo, to = q_off.to(dev), t_off.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = lambda: backend.grouped_mutual(qpool, qo, tpool, to, max_nq=int(nq.max()), total_q=Q, total_t=T)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    backend.profile_enable(True); backend.profile_read(reset=True)
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    kms, kn = backend.profile_read(reset=True)
    backend.profile_enable(False)
    ms = sorted(ts)[len(ts) // 2]
    byts = 128.0 * (Q + T) + 16.0 * Q + 4.0 * T + 16.0 * (G + 1)
    hbm = peaks.get("hbm_gbs", 6650.0)
    kms1 = kms / max(kn, 1)
    return {"workload": "c4: %d groups, n_q,n_t ~ U[32,512], packed, one grouped launch" % G, "ms": ms,
            "groups_per_s": G / (ms * 1e-3), "query_descriptors_per_s": Q / (ms * 1e-3),
            "total_q": Q, "total_t": T, "mutual_matches": None,
            "roofline": {"bound": "hbm", "kernel": "k_grouped_tc", "achieved": byts / (kms1 * 1e-3) / 1e9 if kms1 else None, "peak": hbm,
                         "unit": "GB/s", "frac": (byts / (kms1 * 1e-3) / 1e9 / hbm) if kms1 else None,
                         "traffic": _traffic("k_grouped_tc_c4"),
                         "algorithmic_bytes": byts, "kernel_ms": kms1,
                         "tensor_ops": float((2 * 128 * nq.double() * nt.double()).sum())}}


def fastmatch_leg(dev):
    """configs[0]: the README example, Fast-Match graf img4 (query) -> img1 (target), through the
    drop-in API fastmatch.match(query_cache, target_img, options)(tau).  SIFT (OpenCV, host) is part
    of both arms; the matcher is the wave-batched grouped CUDA launch vs one cv2.BFMatcher call per
    round (the reference's loop, restated in oracle/fastmatch_ref.py)."""
    import cv2
    import torch
    from fast_match_b200 import cache as fm_cache, fastmatch
    from oracle import fastmatch_ref
    gold = os.path.join(ROOT, "tests", "golden")
    img1 = cv2.imread(os.path.join(gold, "graf1.png"))
    ref_cache = fastmatch_ref.RefMetricCache.from_image(os.path.join(gold, "graf4.png"))
    o, th = ref_cache.original, ref_cache.thumb
    mc = fm_cache.Metric_Cache.from_features(th["descriptors"], th["positions"], th["size"],
                                             o["descriptors"], o["positions"], o["size"], {"device": str(dev)})
    out = {"workload": "c1: README example, graf img4 -> img1, Metric_Cache, defaults (grid 50, margin 25, radius 100)"}
    for tau in (0.7, 0.9):
        res = {}
        for name in ("ours", "reference_loop_cv2"):
            best = None
            for _ in range(2):
                stats = {}
                t0 = time.perf_counter()
                if name == "ours":
                    ms = fastmatch.match(mc, img1, {"stats": stats})(tau)
                    torch.cuda.synchronize()
                else:
                    gm = fastmatch_ref.match(ref_cache, img1, {}, mutual=fastmatch_ref.cv2_mutual)
                    ms = gm(tau)
                    stats = {"rounds_evaluated": gm.rounds, "launches": gm.rounds}
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, len(ms), stats)
            res[name] = {"s_per_pair": best[0], "pairs_per_s": 1.0 / best[0], "matches": best[1],
                         "rounds": best[2].get("rounds_evaluated"), "matcher_calls": best[2].get("launches")}
        res["identical_match_count"] = res["ours"]["matches"] == res["reference_loop_cv2"]["matches"]
        try:   # precision under the shipped ground-truth homography (evaluate.py)
            from fast_match_b200 import evaluate
            H = evaluate.load_homography(os.path.join(gold, "graf_H1to4p.txt"))
            res["fastmatch_inliers_5px"] = list(evaluate.inlier_fraction(fastmatch.match(mc, img1, {})(tau), H))
            rp, _ = evaluate.ratio_match_positions(o["descriptors"], o["positions"], *_graf1_features(gold), tau)
            res["ratiomatch_inliers_5px"] = list(evaluate.inlier_fraction(rp, H))
        except Exception as ex:  # noqa: BLE001
            res["inliers_error"] = repr(ex)
        out["tau_%.1f" % tau] = res
    return out


def _graf1_features(gold):
    import numpy as np
    g = np.load(os.path.join(gold, "bf_golden.npz"))
    return g["graf1_desc"], g["graf1_pos"]


def sharded_leg(args, dev, world, rank, backend, sharded, barrier, max_over_ranks):
    import torch
    Ntot = Mtot = args.sharded_n
    lo, hi = sharded.shard_range(Ntot, rank, world)
    q = siftlike_torch(0, Mtot, 21, dev)
    t = plant_torch(q, siftlike_torch(lo, hi, 22, dev), lo, 23)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        _, d2, idx = sharded.sharded_top2(q, t, lo)
        _, mask = backend.ratio(d2[:, 0], den_d2=d2[:, 1], tau=TAU, want_ratio=False)
        return mask
    step(); barrier()
    ts = []
    for _ in range(3):
        flush.zero_()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mask = step(); b.record()
        barrier()
        ts.append(max_over_ranks(a.elapsed_time(b)))
    ms = sorted(ts)[len(ts) // 2]
    ops = 2.0 * Mtot * Ntot * 128
    return {"workload": "c5: %dx%d, target set sharded over %d GPU(s), all-gather of packed top-2 keys + merge" % (Mtot, Ntot, world),
            "ms": ms, "value": Mtot / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
            "tops_aggregate": ops / (ms * 1e-3) / 1e12, "matched_queries": int(mask.sum().item())}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
