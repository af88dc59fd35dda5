// fm_api.cu -- the extern "C" surface of libfmatch.so (see include/fastmatch_b200.h)
// plus the small element-wise kernels (ratio test, shard merge).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "fm_common.cuh"

namespace fm {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_events;   // start/stop pairs
// the bracket a launcher has open: begin and end run on the same thread inside one call, so the
// slot is per thread and concurrent calls on other threads / streams cannot pair up each other's events
static thread_local cudaEvent_t g_prof_open = nullptr;

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void prof_begin(cudaStream_t s) {
    {
        std::lock_guard<std::mutex> lock(g_prof_mu);
        if (!g_prof_on) return;
    }
    if (g_prof_open) { cudaEventDestroy(g_prof_open); g_prof_open = nullptr; }   // an aborted call
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    cudaEventRecord(a, s);
    g_prof_open = a;
}

void prof_end(cudaStream_t s) {
    if (!g_prof_open) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(g_prof_open); g_prof_open = nullptr; return; }
    cudaEventRecord(b, s);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_events.push_back(g_prof_open);
    g_prof_events.push_back(b);
    g_prof_open = nullptr;
}

namespace {

constexpr unsigned long long NONE = 0xFFFFFFFFFFFFFFFFull;

// Ratio test: fastmatch.pyx:124,165 / Classic Matching cell 3 (float32 sqrt, float64 divide).
__global__ void k_ratio(const uint32_t *__restrict__ num_d2, int64_t num_stride,
                        const uint32_t *__restrict__ den_d2, int64_t den_stride,
                        const float *__restrict__ den_f32, int64_t M, double tau,
                        double *__restrict__ ratio, uint8_t *__restrict__ mask) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t n = num_d2[i * num_stride];
        uint32_t dd = den_f32 ? 0u : den_d2[i * den_stride];
        double r;
        if (!den_f32) r = ratio_f32sqrt(n, dd);
        else if (n == FM_NONE_D2) r = __longlong_as_double(0x7FF0000000000000ll);
        else r = __ddiv_rn((double)__fsqrt_rn((float)n), (double)den_f32[i]);
        if (ratio) ratio[i] = r;
        if (mask) mask[i] = r < tau ? 1 : 0;
    }
}

// Two smallest packed keys per query over S shards.
__global__ void k_merge(const unsigned long long *__restrict__ keys, int32_t S, int64_t M,
                        unsigned long long *__restrict__ out, uint32_t *__restrict__ d2,
                        int32_t *__restrict__ idx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M;
         i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long a = NONE, b = NONE;
        for (int32_t s = 0; s < S; ++s) {
            const ulonglong2 v = *(const ulonglong2 *)(keys + ((int64_t)s * M + i) * 2);
            insert2(v.x, a, b);
            insert2(v.y, a, b);
        }
        if (out) { out[2 * i] = a; out[2 * i + 1] = b; }
        if (d2) { d2[2 * i] = (uint32_t)(a >> 32); d2[2 * i + 1] = (uint32_t)(b >> 32); }
        if (idx) {
            idx[2 * i] = a == NONE ? -1 : (int32_t)(uint32_t)a;
            idx[2 * i + 1] = b == NONE ? -1 : (int32_t)(uint32_t)b;
        }
    }
}

__global__ void k_dist(const uint32_t *__restrict__ d2, float *__restrict__ dist, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = d2[i];
        dist[i] = v == FM_NONE_D2 ? __int_as_float(0x7F800000) : __fsqrt_rn((float)v);
    }
}

inline int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return (int)g;
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace
}  // namespace fm

using namespace fm;

extern "C" {

int fm_version(void) { return 100; }

long long fm_launch_count(void) { return fm::g_launches.load(); }

int fm_profile_enable(int on) {
    std::lock_guard<std::mutex> lock(fm::g_prof_mu);
    fm::g_prof_on = on != 0;
    return FM_OK;
}

int fm_profile_read(double *total_ms, int *launches, int reset) {
    std::lock_guard<std::mutex> lock(fm::g_prof_mu);
    double tot = 0;
    int n = 0;
    for (size_t i = 0; i + 1 < fm::g_prof_events.size(); i += 2) {
        float ms = 0;
        FM_CUDA_TRY(cudaEventSynchronize(fm::g_prof_events[i + 1]));
        FM_CUDA_TRY(cudaEventElapsedTime(&ms, fm::g_prof_events[i], fm::g_prof_events[i + 1]));
        tot += ms;
        ++n;
    }
    if (reset) {
        for (cudaEvent_t e : fm::g_prof_events) cudaEventDestroy(e);
        fm::g_prof_events.clear();
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return FM_OK;
}

const char *fm_last_error(void) { return fm::g_err; }

int fm_device_caps(int device, int *sm_major, int *sm_minor, int *sm_count, int *has_tcgen05) {
    cudaDeviceProp p;
    FM_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (sm_major) *sm_major = p.major;
    if (sm_minor) *sm_minor = p.minor;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (has_tcgen05) *has_tcgen05 = (p.major == 10) ? 1 : 0;
    return FM_OK;
}

size_t fm_top2_workspace_bytes(int64_t M, int64_t N) { return fm::top2_tc_workspace_bytes(M, N); }

static int top2_impl(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, int32_t t_index_base,
                     uint32_t *d2, int32_t *idx, uint64_t *keys, fm::RatioOut rout, void *ws,
                     size_t ws_bytes, int algo, void *stream, const char *who) {
    if (M < 0 || N < 0 || (M > 0 && (!q || !d2 || !idx)) || (N > 0 && !t)) {
        set_error("%s: bad argument (M=%lld N=%lld q=%p t=%p d2=%p idx=%p)", who, (long long)M,
                  (long long)N, (const void *)q, (const void *)t, (void *)d2, (void *)idx);
        return FM_EINVAL;
    }
    if (!aligned16(q) || !aligned16(t)) {
        set_error("%s: descriptor pointers must be 16-byte aligned", who);
        return FM_EINVAL;
    }
    if (N + (int64_t)t_index_base > 0x7FFFFFFFll || M > 0x7FFFFFFFll) {
        set_error("%s: index range exceeds int32", who);
        return FM_EINVAL;
    }
    if (M == 0) return FM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    bool use_tc;
    if (algo == FM_ALGO_TCGEN05) {
        if (!fm::tc_supported()) {
            set_error("%s: FM_ALGO_TCGEN05 requested but the device is not sm_100", who);
            return FM_EUNSUPPORTED;
        }
        use_tc = true;
    } else if (algo == FM_ALGO_MMA_SYNC) {
        use_tc = false;
    } else if (algo == FM_ALGO_AUTO) {
        // Measured (tools/quick_perf.py, round 2): with its three launches chained by programmatic
        // dependent launch the tensor-core path is at least as fast as the warp-MMA kernel from
        // 300 x 300 up (13 vs 15 us; 2426 x 1058, the README example's thumbnail shape: 17 vs 37 us;
        // 3668 x 3668, its Metric_Cache self-match: 26 vs 107 us).  Only really tiny problems, where
        // one launch beats three, stay on the warp-MMA kernel.
        use_tc = fm::tc_supported() && M * N >= (int64_t)1 << 16;
    } else {
        set_error("%s: unknown algo %d", who, algo);
        return FM_EINVAL;
    }
    if (use_tc) {
        if (ws_bytes < fm::top2_tc_workspace_bytes(M, N) || (!ws && fm::top2_tc_workspace_bytes(M, N))) {
            set_error("%s: workspace too small (%zu < %zu)", who, ws_bytes, fm::top2_tc_workspace_bytes(M, N));
            return FM_ENOSPACE;
        }
        return fm::launch_top2_tc(q, M, t, N, t_index_base, d2, idx, keys, rout, ws, ws_bytes, s);
    }
    int rc = fm::launch_sweep_mma_dense(q, M, t, N, t_index_base, d2, idx, keys, s);
    if (rc != FM_OK || (!rout.ratio && !rout.mask)) return rc;
    return fm_ratio_f32sqrt(d2, 2, d2 + 1, 2, nullptr, M, rout.tau, rout.ratio, rout.mask, stream);
}

int fm_top2_u8(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, int32_t t_index_base,
               uint32_t *d2, int32_t *idx, uint64_t *keys, void *ws, size_t ws_bytes, int algo,
               void *stream) {
    return top2_impl(q, M, t, N, t_index_base, d2, idx, keys, fm::RatioOut{0.0, nullptr, nullptr}, ws,
                     ws_bytes, algo, stream, "fm_top2_u8");
}

int fm_ratio_match_u8(const uint8_t *q, int64_t M, const uint8_t *t, int64_t N, double tau,
                      uint32_t *d2, int32_t *idx, double *ratio, uint8_t *mask, void *ws,
                      size_t ws_bytes, int algo, void *stream) {
    if (M > 0 && !ratio && !mask) {
        set_error("fm_ratio_match_u8: give ratio and/or mask");
        return FM_EINVAL;
    }
    return top2_impl(q, M, t, N, 0, d2, idx, nullptr, fm::RatioOut{tau, ratio, mask}, ws, ws_bytes, algo,
                     stream, "fm_ratio_match_u8");
}

int fm_ratio_f32sqrt(const uint32_t *num_d2, int64_t num_stride, const uint32_t *den_d2,
                     int64_t den_stride, const float *den_f32, int64_t M, double tau,
                     double *ratio, uint8_t *mask, void *stream) {
    if (M < 0 || (M > 0 && (!num_d2 || (!den_d2 && !den_f32) || (!ratio && !mask)))) {
        set_error("fm_ratio_f32sqrt: bad argument");
        return FM_EINVAL;
    }
    if (M == 0) return FM_OK;
    k_ratio<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(num_d2, num_stride, den_d2,
                                                               den_stride, den_f32, M, tau, ratio,
                                                               mask);
    FM_CUDA_TRY(cudaGetLastError());
    fm::count_launch();
    return FM_OK;
}

size_t fm_grouped_workspace_bytes(int64_t total_q, int64_t total_t, int64_t tpool_rows, int32_t G) {
    (void)G;
    const size_t a = (size_t)(total_t > 0 ? total_t : 0) * sizeof(unsigned long long) + 16;
    const size_t b = fm::grouped_tc_workspace_bytes(total_q > 0 ? total_q : 0,
                                                    tpool_rows > 0 ? tpool_rows : 0, true);
    return a > b ? a : b;
}

int fm_grouped_mutual_u8(const uint8_t *qpool, const int32_t *q_gather, const int64_t *q_off,
                         const uint8_t *tpool, const int64_t *t_off, const int64_t *t_base,
                         int32_t G, int64_t total_q, int64_t total_t, int64_t tpool_rows,
                         int32_t max_nq, uint32_t *q2t_d2, int32_t *q2t_idx, int32_t *t2q_idx,
                         uint8_t *mutual, void *ws, size_t ws_bytes, int algo, void *stream) {
    if (G < 0 || total_q < 0 || total_t < 0 || max_nq < 0 || tpool_rows < 0 ||
        (G > 0 && (!q_off || !t_off)) || (total_q > 0 && (!qpool || !q2t_d2 || !q2t_idx)) ||
        (total_t > 0 && (!tpool || !t2q_idx))) {
        set_error("fm_grouped_mutual_u8: bad argument");
        return FM_EINVAL;
    }
    if (!t_base) tpool_rows = total_t;
    if (!aligned16(qpool) || !aligned16(tpool)) {
        set_error("fm_grouped_mutual_u8: descriptor pointers must be 16-byte aligned");
        return FM_EINVAL;
    }
    if (total_q > 0x7FFFFFFFll || tpool_rows > 0x7FFFFFFFll) {
        set_error("fm_grouped_mutual_u8: row counts exceed int32");
        return FM_EINVAL;
    }
    if (ws_bytes < fm_grouped_workspace_bytes(total_q, total_t, tpool_rows, G) || !ws) {
        set_error("fm_grouped_mutual_u8: workspace too small (%zu < %zu)", ws_bytes,
                  fm_grouped_workspace_bytes(total_q, total_t, tpool_rows, G));
        return FM_ENOSPACE;
    }
    if (G == 0) return FM_OK;
    bool use_tc;
    if (algo == FM_ALGO_TCGEN05) {
        if (!fm::tc_supported()) {
            set_error("fm_grouped_mutual_u8: FM_ALGO_TCGEN05 requested but the device is not sm_100");
            return FM_EUNSUPPORTED;
        }
        use_tc = true;
    } else if (algo == FM_ALGO_MMA_SYNC) {
        use_tc = false;
    } else if (algo == FM_ALGO_AUTO) {
        use_tc = fm::tc_supported();
    } else {
        set_error("fm_grouped_mutual_u8: unknown algo %d", algo);
        return FM_EINVAL;
    }
    if (use_tc)
        return fm::launch_grouped_tc(qpool, q_gather, q_off, tpool, t_off, t_base, G, total_q, total_t,
                                     tpool_rows, q2t_d2, q2t_idx, t2q_idx, mutual, ws, ws_bytes,
                                     (cudaStream_t)stream);
    return fm::launch_sweep_mma_grouped(qpool, q_gather, q_off, tpool, t_off, t_base, G, total_q, total_t,
                                        max_nq, q2t_d2, q2t_idx, t2q_idx, mutual,
                                        (unsigned long long *)ws, (cudaStream_t)stream);
}

int fm_merge_top2(const uint64_t *keys, int32_t S, int64_t M, uint64_t *out_keys, uint32_t *d2,
                  int32_t *idx, void *stream) {
    if (S < 0 || M < 0 || (M > 0 && S > 0 && !keys) || (!out_keys && !d2 && !idx)) {
        set_error("fm_merge_top2: bad argument");
        return FM_EINVAL;
    }
    if (M == 0) return FM_OK;
    k_merge<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)keys, S, M, (unsigned long long *)out_keys, d2, idx);
    FM_CUDA_TRY(cudaGetLastError());
    fm::count_launch();
    return FM_OK;
}

// ---- host-buffer convenience --------------------------------------------------------
namespace {
constexpr int HOST_MAX_DEVICES = 64;
struct HostCtx {                 // one per device: staging buffers, two streams, events
    std::mutex mu;
    bool ready = false;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t in_ready[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr};
    cudaEvent_t t_begin = nullptr, t_up = nullptr;      // bracket the first upload of a call (timing enabled)
    double h2d_bytes_per_ms = 0.0;                       // measured on the previous calls of this context
    uint8_t *pin_in = nullptr; size_t pin_in_cap = 0;
    uint8_t *pin_out = nullptr; size_t pin_out_cap = 0;
    uint8_t *dev = nullptr; size_t dev_cap = 0;
};
HostCtx g_host[HOST_MAX_DEVICES];

int ensure(uint8_t **p, size_t *cap, size_t need, bool pinned) {
    if (*cap >= need) return FM_OK;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t n = need + need / 4 + 4096;
    if (pinned) FM_CUDA_TRY(cudaMallocHost((void **)p, n));
    else FM_CUDA_TRY(cudaMalloc((void **)p, n));
    *cap = n;
    return FM_OK;
}
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

bool is_device_accessible_host(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

struct DeviceGuard {             // the caller's current device is restored on every exit path
    int prev = -1;
    bool switched = false;
    int enter(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        FM_CUDA_TRY(cudaSetDevice(device));
        switched = true;
        return FM_OK;
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
};

// Body of fm_top2_host_u8 (runs with the context locked and the device selected).
// Copies run on a copy stream, kernels on a compute stream.  On a slow host link the queries go up
// in two halves: the compute stream matches the first half against the targets while the second is
// still being copied, and each half's results go back while the other half computes (row halves
// are independent, so nothing has to be merged); on a fast link everything goes up in one piece.
int top2_host_locked(HostCtx &c, const uint8_t *q_host, int64_t M, const uint8_t *t_host, int64_t N,
                     uint32_t *d2_host, int32_t *idx_host, float *dist_host, double tau, uint8_t *mask_host,
                     bool *enqueued) {
    if (!c.ready) {
        FM_CUDA_TRY(cudaStreamCreateWithFlags(&c.compute, cudaStreamNonBlocking));
        FM_CUDA_TRY(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
        for (int h = 0; h < 2; ++h) {
            FM_CUDA_TRY(cudaEventCreateWithFlags(&c.in_ready[h], cudaEventDisableTiming));
            FM_CUDA_TRY(cudaEventCreateWithFlags(&c.done[h], cudaEventDisableTiming));
        }
        FM_CUDA_TRY(cudaEventCreate(&c.t_begin));
        FM_CUDA_TRY(cudaEventCreate(&c.t_up));
        c.ready = true;
    }
    // Two halves pay one more launch sequence (~0.05 ms of fixed cost and cold rows at 50k x 50k) to
    // hide the upload of the second half of the queries (a quarter of the input bytes): worth it when
    // this box's PCIe needs more than ~0.36 ms for the whole input (measured: 0.29 ms -> one piece is
    // 0.05 ms faster, 0.85 ms -> two halves are 0.25 ms faster).  The bandwidth is taken from the
    // previous calls on this context; the first call assumes a slow link.
    // (FM_HOST_HALVES=1|2 forces one form, for A/B timing.)
    static const int halves_override = [] { const char *e = getenv("FM_HOST_HALVES"); return e && *e ? atoi(e) : 0; }();
    const double in_bytes = (double)(M + N) * FM_DIM;
    const bool slow_link = c.h2d_bytes_per_ms <= 0.0 || in_bytes / c.h2d_bytes_per_ms > 0.36;
    const int halves = halves_override == 1 ? 1 : (halves_override == 2 && M >= 1024) ? 2
                       : (M >= 16384 && N >= 4096 && slow_link) ? 2 : 1;
    const int64_t m0 = halves == 2 ? ((M / 2 + 511) / 512) * 512 : M;      // whole M-blocks in the first half
    const int64_t hm[2] = {m0, M - m0}, hbeg[2] = {0, m0};
    const size_t qb = (size_t)M * FM_DIM, tb = (size_t)N * FM_DIM;
    const size_t ob_d2 = (size_t)M * 2 * 4, ob_idx = ob_d2, ob_dist = dist_host ? ob_d2 : 0;
    const size_t ob_mask = mask_host ? (size_t)M : 0;
    size_t wsb = up256(fm_top2_workspace_bytes(hm[0], N));
    if (hm[1] > 0 && up256(fm_top2_workspace_bytes(hm[1], N)) > wsb) wsb = up256(fm_top2_workspace_bytes(hm[1], N));
    // device layout: q | t | d2 | idx | dist | mask | ws
    const size_t o_q = 0, o_t = up256(qb), o_d2 = o_t + up256(tb), o_idx = o_d2 + up256(ob_d2),
                 o_dist = o_idx + up256(ob_idx), o_mask = o_dist + up256(ob_dist),
                 o_ws = o_mask + up256(ob_mask), total = o_ws + wsb;
    int rc;
    if ((rc = ensure(&c.dev, &c.dev_cap, total, false)) != FM_OK) return rc;
    // inputs: pinned / registered host memory is copied straight from the caller's buffer,
    // pageable memory goes through the library's pinned staging area (chunk by chunk, so the
    // staging memcpy of one piece overlaps the DMA of the previous one)
    const bool q_pinned = is_device_accessible_host(q_host);
    const bool t_pinned = tb == 0 || is_device_accessible_host(t_host);
    if (!q_pinned || !t_pinned) {
        if ((rc = ensure(&c.pin_in, &c.pin_in_cap, o_d2, true)) != FM_OK) return rc;
    }
    const bool out_pinned = is_device_accessible_host(d2_host) && is_device_accessible_host(idx_host) &&
                            (!dist_host || is_device_accessible_host(dist_host)) &&
                            (!mask_host || is_device_accessible_host(mask_host));
    if (!out_pinned && (rc = ensure(&c.pin_out, &c.pin_out_cap, o_ws - o_d2, true)) != FM_OK) return rc;

    static const bool trace = [] { const char *e = getenv("FM_HOST_TRACE"); return e && *e && atoi(e) != 0; }();
    cudaEvent_t tr[8] = {};
    auto stamp = [&](int i, cudaStream_t st) { if (trace) { cudaEventCreate(&tr[i]); cudaEventRecord(tr[i], st); } };
    *enqueued = true;
    stamp(0, c.copy);
    FM_CUDA_TRY(cudaEventRecord(c.t_begin, c.copy));
    if (tb) {
        const uint8_t *src = t_host;
        if (!t_pinned) { memcpy(c.pin_in + o_t, t_host, tb); src = c.pin_in + o_t; }
        FM_CUDA_TRY(cudaMemcpyAsync(c.dev + o_t, src, tb, cudaMemcpyHostToDevice, c.copy));
    }
    for (int h = 0; h < halves; ++h) {
        const size_t off = (size_t)hbeg[h] * FM_DIM, nb = (size_t)hm[h] * FM_DIM;
        const uint8_t *src = q_host + off;
        if (!q_pinned) { memcpy(c.pin_in + o_q + off, q_host + off, nb); src = c.pin_in + o_q + off; }
        FM_CUDA_TRY(cudaMemcpyAsync(c.dev + o_q + off, src, nb, cudaMemcpyHostToDevice, c.copy));
        FM_CUDA_TRY(cudaEventRecord(c.in_ready[h], c.copy));
        if (h == 0) FM_CUDA_TRY(cudaEventRecord(c.t_up, c.copy));
        stamp(1 + h, c.copy);
    }
    uint8_t *out_base = out_pinned ? nullptr : c.pin_out;
    auto dst = [&](void *user, size_t dev_off) -> uint8_t * {
        return out_pinned ? (uint8_t *)user : out_base + (dev_off - o_d2);
    };
    for (int h = 0; h < halves; ++h) {
        if (hm[h] == 0) continue;
        const int64_t r0 = hbeg[h], m = hm[h];
        FM_CUDA_TRY(cudaStreamWaitEvent(c.compute, c.in_ready[h], 0));
        stamp(3 + 2 * h, c.compute);
        uint32_t *d2_dev = (uint32_t *)(c.dev + o_d2) + r0 * 2;
        int32_t *idx_dev = (int32_t *)(c.dev + o_idx) + r0 * 2;
        if (mask_host)     // Lowe ratio test d1/d2 < tau fused into the same launch sequence
            rc = fm_ratio_match_u8(c.dev + o_q + r0 * FM_DIM, m, c.dev + o_t, N, tau, d2_dev, idx_dev, nullptr,
                                   c.dev + o_mask + r0, c.dev + o_ws, wsb, FM_ALGO_AUTO, c.compute);
        else
            rc = fm_top2_u8(c.dev + o_q + r0 * FM_DIM, m, c.dev + o_t, N, 0, d2_dev, idx_dev, nullptr,
                            c.dev + o_ws, wsb, FM_ALGO_AUTO, c.compute);
        if (rc != FM_OK) return rc;
        if (dist_host) {
            k_dist<<<grid_for(m * 2, 256), 256, 0, c.compute>>>(d2_dev, (float *)(c.dev + o_dist) + r0 * 2, m * 2);
            FM_CUDA_TRY(cudaGetLastError());
            fm::count_launch();
        }
        FM_CUDA_TRY(cudaEventRecord(c.done[h], c.compute));
        stamp(4 + 2 * h, c.compute);
        // results of this half go back on the copy stream (which has nothing else left to do once
        // the inputs are up) while the compute stream works on the other half
        FM_CUDA_TRY(cudaStreamWaitEvent(c.copy, c.done[h], 0));
        const size_t r8 = (size_t)r0 * 8, m8 = (size_t)m * 8;
        FM_CUDA_TRY(cudaMemcpyAsync(dst(d2_host, o_d2) + r8, c.dev + o_d2 + r8, m8, cudaMemcpyDeviceToHost, c.copy));
        FM_CUDA_TRY(cudaMemcpyAsync(dst(idx_host, o_idx) + r8, c.dev + o_idx + r8, m8, cudaMemcpyDeviceToHost, c.copy));
        if (dist_host)
            FM_CUDA_TRY(cudaMemcpyAsync(dst(dist_host, o_dist) + r8, c.dev + o_dist + r8, m8, cudaMemcpyDeviceToHost, c.copy));
        if (mask_host)
            FM_CUDA_TRY(cudaMemcpyAsync(dst(mask_host, o_mask) + r0, c.dev + o_mask + r0, (size_t)m, cudaMemcpyDeviceToHost, c.copy));
    }
    stamp(7, c.copy);
    FM_CUDA_TRY(cudaStreamSynchronize(c.copy));
    FM_CUDA_TRY(cudaStreamSynchronize(c.compute));
    *enqueued = false;
    {   // this box's host-to-device bandwidth, for the next call's choice
        float ms = 0;
        const double up_bytes = (double)tb + (double)hm[0] * FM_DIM;
        if (cudaEventElapsedTime(&ms, c.t_begin, c.t_up) == cudaSuccess && ms > 0 && up_bytes >= (1 << 20)) {
            const double bw = up_bytes / ms;
            c.h2d_bytes_per_ms = c.h2d_bytes_per_ms > 0 ? 0.5 * (c.h2d_bytes_per_ms + bw) : bw;
        } else {
            (void)cudaGetLastError();
        }
    }
    if (trace) {
        const char *nm[8] = {"start", "in0 up", "in1 up", "k0 begin", "k0 done", "k1 begin", "k1 done", "copies done"};
        for (int i = 1; i < 8; ++i) {
            if (!tr[i]) continue;
            float ms = 0;
            cudaEventElapsedTime(&ms, tr[0], tr[i]);
            fprintf(stderr, "  [host trace] %-12s %.3f ms\n", nm[i], ms);
        }
        for (int i = 0; i < 8; ++i) if (tr[i]) cudaEventDestroy(tr[i]);
    }
    if (!out_pinned) {
        memcpy(d2_host, c.pin_out, ob_d2);
        memcpy(idx_host, c.pin_out + (o_idx - o_d2), ob_idx);
        if (dist_host) memcpy(dist_host, c.pin_out + (o_dist - o_d2), ob_dist);
        if (mask_host) memcpy(mask_host, c.pin_out + (o_mask - o_d2), ob_mask);
    }
    return FM_OK;
}
}  // namespace

int fm_top2_host_u8(const uint8_t *q_host, int64_t M, const uint8_t *t_host, int64_t N,
                    uint32_t *d2_host, int32_t *idx_host, float *dist_host, double tau,
                    uint8_t *mask_host, int device) {
    if (M < 0 || N < 0 || (M > 0 && (!q_host || !d2_host || !idx_host)) || (N > 0 && !t_host)) {
        set_error("fm_top2_host_u8: bad argument");
        return FM_EINVAL;
    }
    if (device < 0 || device >= HOST_MAX_DEVICES) {
        set_error("fm_top2_host_u8: device %d out of range", device);
        return FM_EINVAL;
    }
    if (M == 0) return FM_OK;
    HostCtx &c = g_host[device];                       // per device: calls on different GPUs do not serialise
    std::lock_guard<std::mutex> lock(c.mu);
    DeviceGuard guard;
    int rc = guard.enter(device);
    if (rc != FM_OK) return rc;
    bool enqueued = false;
    rc = top2_host_locked(c, q_host, M, t_host, N, d2_host, idx_host, dist_host, tau, mask_host, &enqueued);
    if (enqueued) {
        // an error after work was queued: drain both streams so that the staging buffers are not
        // reused (or the caller's buffers freed) under a copy that is still in flight
        cudaStreamSynchronize(c.copy);
        cudaStreamSynchronize(c.compute);
        (void)cudaGetLastError();
    }
    return rc;
}

}  // extern "C"
