// stencil_b200_capi.cu -- the C-ABI declared in include/stencil_b200.h (contexts, device groups, launches, staging).
//
// Host side of the drop-in boundary: owns device tables, pinned staging buffers, the partial-sum
// slots and one stream per context.  There is no CPU fallback anywhere in this file: when no CUDA
// device is usable every compute entry point returns SB200_ECUDA and says why.
#include "../../include/stencil_b200.h"
#include "dispersion_kernels.cuh"
#include "objective_fa.cuh"

// 1: regrouped sqrtarg in the hot loop, reference order only where bit-identity matters (objective_fa.cuh);
// 0: reference order at every k-point (objective_kernel).  Same results contract either way.
#ifndef SB200_FAST_A
#define SB200_FAST_A 1
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace sb200;

namespace {
thread_local char g_create_error[512] = "";
thread_local char g_group_error[512] = "";

// ---- host-side helper threads for the last hop of a dense export (pinned staging -> the caller's pageable
// buffer).  One memcpy thread moves ~10 GB/s into freshly mapped pages; PCIe delivers ~50 GB/s.
class CopyPool {
public:
    static CopyPool& get() {
        static CopyPool pool;
        return pool;
    }
    int threads() const { return (int)workers_.size() + 1; }
    // dst[0:bytes) = src[0:bytes), striped over the pool (the calling thread takes a stripe too)
    void copy(void* dst, const void* src, size_t bytes) {
        const int T = threads();
        if (bytes < ((size_t)1 << 20) || T == 1) {
            memcpy(dst, src, bytes);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes;
            pending_ = T - 1;
            ++epoch_;
        }
        cv_.notify_all();
        stripe(T - 1, T);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int T = (int)std::min(8u, std::max(1u, hw / 2));
        if (const char* e = getenv("SB200_COPY_THREADS")) T = std::max(1, std::min(64, atoi(e)));
        for (int t = 0; t < T - 1; ++t) workers_.emplace_back([this, t] { run(t); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    void stripe(int t, int T) {
        const size_t per = ((bytes_ + T - 1) / T + 4095) & ~(size_t)4095;
        const size_t lo = std::min(bytes_, per * (size_t)t), hi = std::min(bytes_, lo + per);
        if (hi > lo) memcpy(dst_ + lo, src_ + lo, hi - lo);
    }
    void run(int t) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            stripe(t, threads());
            {
                std::lock_guard<std::mutex> lk(m_);
                --pending_;
            }
            done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    char* dst_ = nullptr;
    const char* src_ = nullptr;
    size_t bytes_ = 0;
    int pending_ = 0;
    unsigned long long epoch_ = 0;
    bool stop_ = false;
};
}  // namespace

struct sb200_ctx {
    int dim = 0;               // caller's dimensionality (2 or 3)
    int n_user[3] = {0, 0, 0};
    int ncoef = 13;
    int device = 0;
    GridDev g{};               // internal 3-axis view
    bool unit_cell = true;     // Y2 == 1 && Z2 == 1 -> no divisions in sqrtarg
    int wkind = 0;             // internal: 0 unit, 1 plane, 2 dense
    double* d_tables = nullptr;
    double* d_w = nullptr;
    size_t w_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_user = nullptr, ev_x[2] = {nullptr, nullptr};
    bool user_pending = false;  // a *_device call was issued on a caller's stream since the last host-path call
    // staging
    double* h_pin = nullptr;   size_t h_pin_cap = 0;   // pinned host, doubles
    double* d_coeffs = nullptr; size_t d_coeffs_cap = 0;
    double* d_out = nullptr;    size_t d_out_cap = 0;
    void* d_part = nullptr;     size_t d_part_cap = 0;  // bytes
    unsigned int* d_tickets = nullptr; size_t d_tickets_cap = 0;
    bool tickets_dirty = false; // a launch failed or its completion is unknown: zero the tickets before the next one
    unsigned int* d_xticket = nullptr;
    double* d_xstats = nullptr;
    double* d_xbuf = nullptr;   size_t d_xbuf_cap = 0;
    unsigned char* h_xpin[2] = {nullptr, nullptr};      // pinned staging of the chunked D2H
    size_t h_xpin_bytes = 0;
    int sm_count = 148;
    int force_c = 0, force_ty = 0, force_tx = 0;
    int env_force_c = 0, env_block_threads = 0;         // experiment knobs, read once at create
    long long launches = 0;
    double last_ms = 0.0;
    double x_bytes = 0.0, x_seconds = 0.0;
    // a batch in flight between objective_enqueue and objective_collect
    long long pend_B = 0;
    double* pend_out = nullptr;
    bool pend_timed = false;
    char err[512] = "";
};

static int fail(sb200_ctx* ctx, int code, const char* fmt, ...) {
    char* dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(ctx, call)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? SB200_ENOMEM : SB200_ECUDA,         \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <typename T>
static int grow(sb200_ctx* ctx, T** p, size_t* cap, size_t need) {
    if (*cap >= need) return SB200_OK;
    if (*p) CUDA_TRY(ctx, cudaFree(*p));
    *p = nullptr; *cap = 0;
    size_t want = std::max(need, (size_t)256);
    CUDA_TRY(ctx, cudaMalloc((void**)p, want * sizeof(T)));
    *cap = want;
    return SB200_OK;
}

static int grow_pinned(sb200_ctx* ctx, size_t need) {
    if (ctx->h_pin_cap >= need) return SB200_OK;
    if (ctx->h_pin) CUDA_TRY(ctx, cudaFreeHost(ctx->h_pin));
    ctx->h_pin = nullptr; ctx->h_pin_cap = 0;
    size_t want = std::max(need + need / 2, (size_t)4096);
    CUDA_TRY(ctx, cudaMallocHost((void**)&ctx->h_pin, want * sizeof(double)));
    ctx->h_pin_cap = want;
    return SB200_OK;
}

extern "C" int sb200_abi_version(void) { return SB200_ABI_VERSION; }

extern "C" int sb200_device_count(int32_t* count) {
    if (!count) return SB200_EINVAL;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, SB200_ECUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    *count = n;
    return SB200_OK;
}

extern "C" const char* sb200_last_error(const sb200_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

static int create_impl(sb200_ctx* ctx, const sb200_grid_desc* d) {
    ctx->dim = d->dim;
    ctx->ncoef = d->dim == 2 ? 7 : 13;
    ctx->device = d->device;
    for (int a = 0; a < d->dim; ++a) ctx->n_user[a] = d->n[a];
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, SB200_ECUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (d->device < 0 || d->device >= ndev) return fail(ctx, SB200_EINVAL, "device %d out of range [0,%d)", d->device, ndev);
    CUDA_TRY(ctx, cudaSetDevice(d->device));
    cudaDeviceProp prop;
    CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, d->device));
    ctx->sm_count = prop.multiProcessorCount;
    {
        // The library carries sm_100a machine code only (no PTX): ask the runtime whether this device can load it
        // instead of guessing from the compute capability (sm_103 / sm_110 / sm_120 parts cannot).
        cudaFuncAttributes fa;
        cudaError_t le = cudaFuncGetAttributes(&fa, combine_slabs_kernel);
        if (le != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, SB200_ECUDA, "device %d (%s, sm_%d%d) cannot run this library's sm_100a kernels: %s",
                        d->device, prop.name, prop.major, prop.minor, cudaGetErrorString(le));
        }
    }

    // Internal 3-axis view.  dim 3: as given.  dim 2: (degenerate, x, y) -- see DESIGN.md 3.1.
    const int off = d->dim == 2 ? 1 : 0;
    int n3[3] = {1, 1, 1};
    for (int a = 0; a < d->dim; ++a) n3[a + off] = d->n[a];
    std::vector<double> host;
    size_t offs[9];
    double amax[2] = {1.0, 1.0};      // max(1, max|cos|), max(1, max|s2|); NaN-sticky
    for (int t = 0; t < 3; ++t)       // 0 cos, 1 s2, 2 k^2
        for (int a = 0; a < 3; ++a) {
            offs[t * 3 + a] = host.size();
            for (int i = 0; i < n3[a]; ++i) {
                double v;
                if (a < off) v = (t == 0) ? 1.0 : 0.0;  // kappa = 0 on the degenerate axis
                else {
                    const int ua = a - off;
                    if (t == 0) v = d->cos_kappa[ua][i];
                    else if (t == 1) v = d->s2[ua][i];
                    else v = d->kcomp[ua][i] * d->kcomp[ua][i];  // dispersion.py:276, 350 (x**2)
                }
                if (t < 2 && amax[t] == amax[t] && !(std::fabs(v) <= amax[t])) amax[t] = std::fabs(v);   // |NaN| = NaN sticks
                host.push_back(v);
            }
            while (host.size() % 2) host.push_back(0.0);  // keep 16-byte alignment per table
        }
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_tables, host.size() * sizeof(double)));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_tables, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
    GridDev& g = ctx->g;
    g.nx = n3[0]; g.ny = n3[1]; g.nz = n3[2];
    const double* T = ctx->d_tables;
    g.cx = T + offs[0]; g.cy = T + offs[1]; g.cz = T + offs[2];
    g.sx2 = T + offs[3]; g.sy2 = T + offs[4]; g.sz2 = T + offs[5];
    g.kx2 = T + offs[6]; g.ky2 = T + offs[7]; g.kz2 = T + offs[8];
    // (Y*dx)**2 with dx = 1.0 (dispersion.py:319, 398), preferably the caller's own bits (see the header).
    const double Y2 = d->Y2 > 0 ? d->Y2 : (d->Y * 1.0) * (d->Y * 1.0);
    const double Z2 = d->Z2 > 0 ? d->Z2 : (d->Z * 1.0) * (d->Z * 1.0);
    if (d->dim == 3) { g.Y2 = Y2; g.Z2 = Z2; }
    else             { g.Y2 = 1.0; g.Z2 = Y2; }
    g.rY2 = 1.0 / g.Y2;   // IEEE division on the host: the correctly rounded reciprocals div_invariant needs
    g.rZ2 = 1.0 / g.Z2;
    ctx->unit_cell = (g.Y2 == 1.0 && g.Z2 == 1.0);
    g.tab_scale = amax[0] * amax[1];   // 1 for genuine cos / s2 tables; widens the regrouping error bound otherwise
    g.w = nullptr;
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaEventCreate(&ctx->ev0));
    CUDA_TRY(ctx, cudaEventCreate(&ctx->ev1));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_user, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_x[0], cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_x[1], cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_xticket, sizeof(unsigned int)));
    CUDA_TRY(ctx, cudaMemset(ctx->d_xticket, 0, sizeof(unsigned int)));
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_xstats, 3 * sizeof(double)));
    // experiment knobs (tools/sweep.py), read once: never on the launch path
    if (const char* ev = getenv("SB200_FORCE_C")) {
        const int v = atoi(ev);
        if (v == 1 || v == 2 || v == 3 || v == 4 || v == 8) ctx->env_force_c = v;
    }
    if (const char* ev = getenv("SB200_BLOCK_THREADS")) {
        const int v = atoi(ev);
        if (v >= 32 && v <= 256 && v % 32 == 0) ctx->env_block_threads = v;
    }
    return SB200_OK;
}

static int check_desc(const sb200_grid_desc* d, char* errbuf) {
    auto bad = [&](const char* fmt, int a, int b) {
        snprintf(errbuf, 512, fmt, a, b);
        return SB200_EINVAL;
    };
    if (!d) return bad("desc is NULL", 0, 0);
    if (d->dim != 2 && d->dim != 3) return bad("dim must be 2 or 3, got %d", d->dim, 0);
    for (int a = 0; a < d->dim; ++a) {
        if (d->n[a] < 1) return bad("n[%d] = %d must be >= 1", a, d->n[a]);
        if (!d->cos_kappa[a] || !d->s2[a] || !d->kcomp[a]) return bad("table %d is NULL", a, 0);
    }
    if (!(d->Y > 0) || (d->dim == 3 && !(d->Z > 0))) return bad("Y and Z must be positive", 0, 0);
    return SB200_OK;
}

extern "C" int sb200_create(sb200_ctx** out, const sb200_grid_desc* d) {
    if (!out) return fail(nullptr, SB200_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = check_desc(d, g_create_error);
    if (rc) return rc;
    sb200_ctx* ctx = new (std::nothrow) sb200_ctx();
    if (!ctx) return fail(nullptr, SB200_ENOMEM, "out of host memory");
    try {                                   // std::vector / std::thread inside: no exception may cross the C boundary
        rc = create_impl(ctx, d);
    } catch (const std::bad_alloc&) {
        rc = fail(ctx, SB200_ENOMEM, "out of host memory");
    } catch (...) {
        rc = fail(ctx, SB200_ECUDA, "unexpected C++ exception while creating the context");
    }
    if (rc != SB200_OK) {
        snprintf(g_create_error, sizeof g_create_error, "%s", ctx->err);
        sb200_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SB200_OK;
}

extern "C" void sb200_destroy(sb200_ctx* ctx) {
    if (!ctx) return;
    if (ctx->d_tables || ctx->stream) cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_tables); cudaFree(ctx->d_w); cudaFree(ctx->d_coeffs); cudaFree(ctx->d_out);
    cudaFree(ctx->d_part); cudaFree(ctx->d_tickets); cudaFree(ctx->d_xticket); cudaFree(ctx->d_xstats);
    cudaFree(ctx->d_xbuf);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (int i = 0; i < 2; ++i) if (ctx->h_xpin[i]) cudaFreeHost(ctx->h_xpin[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_user) cudaEventDestroy(ctx->ev_user);
    for (int i = 0; i < 2; ++i) if (ctx->ev_x[i]) cudaEventDestroy(ctx->ev_x[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// Host-path calls run on ctx->stream; make that stream wait for whatever a *_device call left on a caller's stream
// (shared scratch: coefficient staging, partial slots, tickets).
static int join_user_stream(sb200_ctx* ctx) {
    if (ctx->user_pending) {
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_user, 0));
        ctx->user_pending = false;
    }
    return SB200_OK;
}

static int mark_user_stream(sb200_ctx* ctx, cudaStream_t st) {
    if (st != ctx->stream) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_user, st));
        ctx->user_pending = true;
    }
    return SB200_OK;
}

extern "C" int sb200_set_weight(sb200_ctx* ctx, int32_t kind, const double* w, int64_t count) {
    if (!ctx) return SB200_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const GridDev& g = ctx->g;
    if (kind == SB200_W_UNIT) {
        ctx->wkind = 0; ctx->g.w = nullptr;
        return SB200_OK;
    }
    size_t need;
    int internal;
    if (kind == SB200_W_PLANE) {
        if (ctx->dim != 3) return fail(ctx, SB200_EINVAL, "SB200_W_PLANE is for dim 3; pass a dense (nx*ny) table in 2D");
        need = (size_t)g.ny * g.nz; internal = 1;
    } else if (kind == SB200_W_DENSE) {
        need = (size_t)g.nx * g.ny * g.nz;
        internal = ctx->dim == 2 ? 1 : 2;  // the 2-D grid IS the (y,z) plane of the internal view
    } else {
        return fail(ctx, SB200_EINVAL, "unknown weight kind %d", kind);
    }
    if (!w || count != (int64_t)need) return fail(ctx, SB200_EINVAL, "weight kind %d needs %zu doubles, got %lld", kind, need, (long long)count);
    int rc = join_user_stream(ctx);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    rc = grow(ctx, &ctx->d_w, &ctx->w_count, need);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_w, w, need * sizeof(double), cudaMemcpyHostToDevice));
    ctx->wkind = internal; ctx->g.w = ctx->d_w;
    return SB200_OK;
}

extern "C" int sb200_set_tiling(sb200_ctx* ctx, int32_t c, int32_t ty, int32_t tx) {
    if (!ctx) return SB200_EINVAL;
    if (!(c == 0 || c == 1 || c == 2 || c == 3 || c == 4 || c == 8)) return fail(ctx, SB200_EINVAL, "cand_per_thread must be 0,1,2,3,4,8");
    if (ty < 0 || tx < 0 || ty > 1024) return fail(ctx, SB200_EINVAL, "bad tile");
    ctx->force_c = c; ctx->force_ty = ty; ctx->force_tx = tx;
    return SB200_OK;
}

namespace {
// A resolved request: which planes, how the batch is segmented, which kernel variant.
struct Request {
    int xb = 0, stride = 1, nplanes = 0;   // INTERNAL first axis
    long long seg = 0;                     // candidates per segment (== B when un-segmented)
    long long nseg = 1;
    bool scan = false;
    bool fast = false;                     // the shorter asin series (throughput launches); screening implies it
};

struct Plan {
    int C, TZ, TY, TX, n_ztiles, n_xchunks, n_ychunks, P;
    long long chunks_per_seg, n_chunks;
    size_t smem;
};

size_t smem_bytes(int C, int TZ) {
    const int nwarps = (TZ + 31) / 32;       // tables are per warp (32 rows), not per y tile
#if SB200_FAST_A
    const int row = 2 + 2 * C;               // RowLayoutFA<C>::ROW
    return (size_t)C * sizeof(Cand) + (size_t)nwarps * 32 * row * 8 +
           (size_t)C * 8 * 8 * 4 + (size_t)C * TZ * 8 * 2;
#else
    const int row = ((2 + 3 * C) + 1) & ~1;  // RowLayout<C>::ROW
    return (size_t)C * sizeof(Cand) + (size_t)nwarps * 32 * row * 8 + (size_t)C * 8 * 8 * 4 + (size_t)C * TZ * 8;
#endif
}

// The tiling is a function of (grid, planes per launch, candidates per SEGMENT) only -- never of the number of
// segments -- which is what makes a segment's results those of its stand-alone launch.
Plan make_plan(const sb200_ctx* ctx, const Request& r) {
    const GridDev& g = ctx->g;
    const long long B = r.seg;
    const int nxr = r.nplanes;
    Plan p;
    // Candidates per thread: 4 chains in flight per thread saturate the FP64 pipe (8-cycle dependent issue, one
    // instruction per 2 cycles).  Small batches take the C in {3, 4} that pads the last chunk least (an SLSQP
    // iterate of a 2-parameter stencil is 3 rows: C = 4 would spend a quarter of the launch on a padded copy).
    if (ctx->force_c) p.C = ctx->force_c;
    else if (B >= 3) p.C = (B <= 64 && ((B + 2) / 3) * 3 < ((B + 3) / 4) * 4) ? 3 : 4;
    else p.C = B >= 2 ? 2 : 1;
    if (ctx->env_force_c) p.C = ctx->env_force_c;
    const int tz_max = ctx->env_block_threads ? ctx->env_block_threads : 256;
    p.TZ = std::min(tz_max, ((g.nz + 31) / 32) * 32);
    p.n_ztiles = (g.nz + p.TZ - 1) / p.TZ;
    p.chunks_per_seg = (B + p.C - 1) / p.C;
    p.n_chunks = p.chunks_per_seg * r.nseg;
    p.TY = ctx->force_ty ? std::min(ctx->force_ty, g.ny) : g.ny;
    p.TX = ctx->force_tx ? std::min(ctx->force_tx, nxr) : nxr;
    // Enough CTAs for ~16 waves of (2 CTAs x SMs) when the problem is that large -- the cost of a k-point depends
    // on its class, many small CTAs balance that -- but never fewer than 128 rows per CTA: below that the
    // prologue (candidate constants, corner probe, three barriers) and the epilogue cost more than the sweep
    // (ncu of 32-row CTAs: 3.7 warps per issue slot stalled on barriers, FP64 pipe 38 % active).
    const long long target = (long long)ctx->sm_count * 2 * 16;
    const int min_rows = 128;
    auto ctas = [&]() {
        return p.chunks_per_seg * p.n_ztiles * ((nxr + p.TX - 1) / p.TX) * ((g.ny + p.TY - 1) / p.TY);
    };
    if (!ctx->force_tx)
        while (ctas() < target && p.TX > 1 && (long long)((p.TX + 1) / 2) * p.TY >= min_rows) p.TX = (p.TX + 1) / 2;
    if (!ctx->force_ty)
        while (ctas() < target && p.TY > 32 && p.TX * (((p.TY + 1) / 2 + 31) / 32 * 32) >= min_rows)
            p.TY = ((p.TY + 1) / 2 + 31) / 32 * 32;   // whole 32-row tiles
    p.n_xchunks = (nxr + p.TX - 1) / p.TX;
    p.n_ychunks = (g.ny + p.TY - 1) / p.TY;
    p.P = p.n_ztiles * p.n_xchunks * p.n_ychunks;
    p.smem = smem_bytes(p.C, p.TZ);
    return p;
}

template <int C>
cudaError_t launch_objective(const sb200_ctx* ctx, const Plan& p, const BatchArgs& a, const Request& r, cudaStream_t st) {
    dim3 grid((unsigned)(p.n_chunks * p.n_ztiles), (unsigned)p.n_xchunks, (unsigned)p.n_ychunks), block(p.TZ);
#if SB200_FAST_A
#define SB200_KERNEL(UNIT, WK, SCAN, PREC) objective_kernel_fa<C, UNIT, WK, SCAN, PREC>
#else
#define SB200_KERNEL(UNIT, WK, SCAN, PREC) objective_kernel<C, UNIT, WK>
#endif
#define SB200_LAUNCH(UNIT, WK, SCAN, PREC)                                                                  \
    do {                                                                                                    \
        auto kern = SB200_KERNEL(UNIT, WK, SCAN, PREC);                                                     \
        if (p.smem > 48 * 1024) {                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem); \
            if (e != cudaSuccess) return e;                                                                 \
        }                                                                                                   \
        kern<<<grid, block, p.smem, st>>>(ctx->g, a);                                                       \
    } while (0)
#define SB200_LAUNCH_W(UNIT, SCAN, PREC)                                                                    \
    do {                                                                                                    \
        if (ctx->wkind == 0) SB200_LAUNCH(UNIT, 0, SCAN, PREC);                                             \
        else if (ctx->wkind == 1) SB200_LAUNCH(UNIT, 1, SCAN, PREC);                                        \
        else SB200_LAUNCH(UNIT, 2, SCAN, PREC);                                                             \
    } while (0)
    // three variants per (C, cell, weight): screening (always the fast series), fast, precise
#define SB200_LAUNCH_V(UNIT)                                                                                \
    do {                                                                                                    \
        if (r.scan) SB200_LAUNCH_W(UNIT, true, 0);                                                          \
        else if (r.fast) SB200_LAUNCH_W(UNIT, false, 0);                                                    \
        else SB200_LAUNCH_W(UNIT, false, 1);                                                                \
    } while (0)
    if (ctx->unit_cell) SB200_LAUNCH_V(true); else SB200_LAUNCH_V(false);
#undef SB200_LAUNCH_V
#undef SB200_LAUNCH_W
#undef SB200_LAUNCH
#undef SB200_KERNEL
    return cudaGetLastError();
}
}  // namespace

// opts (may be NULL) + nbatch -> Request in the INTERNAL axis convention.  For dim 2 the caller's first axis is
// the internal second axis, which the kernels do not slice: slabs are a 3-D feature.
static int resolve_request(sb200_ctx* ctx, int64_t B, const sb200_batch_opts* o, Request* r) {
    const int nx_user = ctx->n_user[0];
    int xb = 0, stride = 1, np = 0;
    long long seg = 0;
    int flags = 0;
    if (o) { xb = o->x_begin; stride = o->x_stride ? o->x_stride : 1; np = o->n_planes; seg = o->seg_size; flags = o->flags; }
    if (stride < 1) return fail(ctx, SB200_EINVAL, "x_stride %d must be >= 1", stride);
    if (xb < 0 || xb >= nx_user) return fail(ctx, SB200_EINVAL, "x_begin %d outside [0,%d)", xb, nx_user);
    if (np == 0) np = (nx_user - xb + stride - 1) / stride;
    if (np < 1 || (long long)xb + (long long)(np - 1) * stride >= nx_user)
        return fail(ctx, SB200_EINVAL, "plane list (begin %d, stride %d, count %d) leaves [0,%d)", xb, stride, np, nx_user);
    if (seg < 0 || (seg > 0 && B % seg != 0))
        return fail(ctx, SB200_EINVAL, "nbatch %lld is not a multiple of seg_size %lld", (long long)B, seg);
    if (ctx->dim == 2) {
        if (xb != 0 || stride != 1 || np != nx_user)
            return fail(ctx, SB200_EINVAL, "x-plane lists (k-slabs) are supported for dim 3 only");
        r->xb = 0; r->stride = 1; r->nplanes = 1;
    } else {
        r->xb = xb; r->stride = stride; r->nplanes = np;
    }
    r->seg = seg > 0 ? seg : B;
    r->nseg = B / r->seg;
    r->scan = (flags & SB200_F_SCAN) != 0;
    r->fast = (flags & SB200_F_FAST) != 0;
    return SB200_OK;
}

static int objective_device_impl(sb200_ctx* ctx, const double* d_coeffs, int64_t B, const Request& r, double* d_out,
                                 cudaStream_t st) {
    const Plan p = make_plan(ctx, r);
    if (p.n_chunks * p.n_ztiles > 0x7fffffffLL || p.n_xchunks > 65535 || p.n_ychunks > 65535)
        return fail(ctx, SB200_EINVAL, "batch of %lld candidates is too large for one launch; split it", (long long)B);
    const size_t bpad = (size_t)p.n_chunks * p.C;
    const size_t part_bytes = bpad * p.P * 8 * 4;
    int rc = grow(ctx, (unsigned char**)&ctx->d_part, &ctx->d_part_cap, part_bytes);
    if (rc) return rc;
    if (ctx->d_tickets_cap < (size_t)p.n_chunks) {
        rc = grow(ctx, &ctx->d_tickets, &ctx->d_tickets_cap, (size_t)p.n_chunks + (size_t)p.n_chunks / 2);
        if (rc) return rc;
        ctx->tickets_dirty = true;
    }
    if (ctx->tickets_dirty) {
        // first use, growth, or a previous launch whose completion is unknown: the tickets must read zero
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_tickets, 0, ctx->d_tickets_cap * sizeof(unsigned int), st));
        ctx->tickets_dirty = false;
    }
    BatchArgs a;
    a.coeffs = d_coeffs; a.ncoef = ctx->ncoef; a.B = B;
    a.seg_size = r.seg; a.chunks_per_seg = (int)p.chunks_per_seg;
    a.x_begin = r.xb; a.x_stride = r.stride; a.n_planes = r.nplanes;
    a.tile_x = p.TX; a.tile_y = p.TY; a.n_ztiles = p.n_ztiles; a.n_xchunks = p.n_xchunks; a.n_ychunks = p.n_ychunks;
    a.P = p.P;
    unsigned char* base = (unsigned char*)ctx->d_part;
    a.part_S = (double*)base;
    a.part_min = (unsigned long long*)(base + bpad * p.P * 8);
    a.part_max = (unsigned long long*)(base + bpad * p.P * 16);
    a.part_nan = (unsigned long long*)(base + bpad * p.P * 24);
    a.tickets = ctx->d_tickets;
    a.out = d_out;
    cudaError_t e;
    switch (p.C) {
        case 1: e = launch_objective<1>(ctx, p, a, r, st); break;
        case 2: e = launch_objective<2>(ctx, p, a, r, st); break;
        case 3: e = launch_objective<3>(ctx, p, a, r, st); break;
        case 4: e = launch_objective<4>(ctx, p, a, r, st); break;
        default: e = launch_objective<8>(ctx, p, a, r, st); break;
    }
    if (e != cudaSuccess) {
        ctx->tickets_dirty = true;
        return fail(ctx, SB200_ECUDA, "objective kernel launch failed: %s", cudaGetErrorString(e));
    }
    ctx->launches += 1;
    return SB200_OK;
}

extern "C" int sb200_objective_batch_device_ex(sb200_ctx* ctx, const double* d_coeffs, int64_t nbatch,
                                               const sb200_batch_opts* opts, double* d_out, void* stream) {
    if (!ctx) return SB200_EINVAL;
    if (!d_coeffs || !d_out || nbatch < 1) return fail(ctx, SB200_EINVAL, "bad arguments (nbatch=%lld)", (long long)nbatch);
    Request r;
    int rc = resolve_request(ctx, nbatch, opts, &r);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = objective_device_impl(ctx, d_coeffs, nbatch, r, d_out, (cudaStream_t)stream))) return rc;
    return mark_user_stream(ctx, (cudaStream_t)stream);
}

static void range_opts(sb200_batch_opts* o, int32_t x_begin, int32_t x_end) {
    memset(o, 0, sizeof *o);
    o->x_begin = x_begin; o->x_stride = 1; o->n_planes = x_end - x_begin;
}

extern "C" int sb200_objective_batch_device(sb200_ctx* ctx, const double* d_coeffs, int64_t nbatch, int32_t x_begin,
                                            int32_t x_end, double* d_out, void* stream) {
    if (!ctx) return SB200_EINVAL;
    if (x_begin >= x_end) return fail(ctx, SB200_EINVAL, "x range [%d,%d) is empty", x_begin, x_end);
    sb200_batch_opts o;
    range_opts(&o, x_begin, x_end);
    return sb200_objective_batch_device_ex(ctx, d_coeffs, nbatch, &o, d_out, stream);
}

extern "C" int sb200_combine_slabs_device(sb200_ctx* ctx, const double* d_parts, int32_t world, int64_t nbatch,
                                          double* d_out, void* stream) {
    if (!ctx) return SB200_EINVAL;
    if (!d_parts || !d_out || world < 1 || nbatch < 1) return fail(ctx, SB200_EINVAL, "bad arguments");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    combine_slabs_kernel<<<(unsigned)((nbatch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_parts, world, nbatch, d_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SB200_ECUDA, "combine kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches += 1;
    return SB200_OK;
}

// SB200_F_AUTO_SCAN: is this a start-grid-like batch?  A candidate whose sqrtarg at one of the far corners of the
// grid (closed forms of dispersion.py:294-297, 369-375) already gives s^2 = dt^2 A > 1.05 has NaN omegas over a
// good part of the grid; the screening variant pays off when at least a quarter of the batch is like that.
static bool looks_like_a_scan(const sb200_ctx* ctx, const double* c, int64_t B) {
    if (B < 64) return false;
    const double rY = 1.0 / ctx->g.Y2, rZ = 1.0 / ctx->g.Z2;
    int64_t beyond = 0;
    const int64_t step = B > 4096 ? B / 4096 : 1;   // a sample is enough
    int64_t seen = 0;
    for (int64_t b = 0; b < B; b += step, ++seen) {
        const double* r = c + b * ctx->ncoef;
        double top;
        if (ctx->ncoef == 13) {
            const double ax = r[1], ay = r[2], az = r[3], bxy = r[4], bxz = r[5], byx = r[6], byz = r[7], bzx = r[8],
                         bzy = r[9], dlx = r[10], dly = r[11], dlz = r[12];
            const double X = ax - dlx, Yv = (ay - dly) * rY, Zv = (az - dlz) * rZ;
            top = std::max({X + 2 * bxy + 2 * bxz, Yv + 2 * (byx + byz) * rY, Zv + 2 * (bzx + bzy) * rZ,
                            X - 2 * bxy + 2 * bxz + Yv - 2 * (byx - byz) * rY,
                            X + 2 * bxy - 2 * bxz + Zv - 2 * (bzx - bzy) * rZ,
                            Yv + 2 * (byx - byz) * rY + Zv + 2 * (bzx - bzy) * rZ,
                            X - 2 * bxy - 2 * bxz + Yv - 2 * (byx + byz) * rY + Zv - 2 * (bzx + bzy) * rZ});
        } else {
            const double ax = r[1], ay = r[2], bxy = r[3], byx = r[4], dlx = r[5], dly = r[6];
            // the internal view divides the caller's second axis by g.Z2 (create_impl)
            top = std::max({(ay + 2 * byx - dly) * rZ, ax - 2 * bxy - dlx + (ay - 2 * byx - dly) * rZ, ax + 2 * bxy - dlx});
        }
        if (r[0] * r[0] * top > 1.05) ++beyond;
    }
    return 4 * beyond >= seen;
}

// ---- host-buffer path, split into "enqueue" and "collect" so that a device group can have every device busy before
// it waits for the first one ---------------------------------------------------------------------------------------
static int objective_enqueue(sb200_ctx* ctx, const double* coeffs, int64_t B, Request r, bool auto_scan) {
    int rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = join_user_stream(ctx))) return rc;
    const size_t nc = (size_t)B * ctx->ncoef;
    if ((rc = grow_pinned(ctx, nc + 4 * (size_t)B))) return rc;
    double* h_out = ctx->h_pin + nc;
    memcpy(ctx->h_pin, coeffs, nc * sizeof(double));
    if (auto_scan && !r.scan) r.scan = looks_like_a_scan(ctx, ctx->h_pin, B);
    const double evals = (double)B * (double)r.nplanes * ctx->g.ny * ctx->g.nz;
    ctx->pend_B = B;
    ctx->pend_out = h_out;
    if (B <= 64 && evals < 67108864.0) {
        // Latency path (SciPy's SLSQP asks for one iterate and its finite-difference neighbours at a
        // time): no event records and no D2H copy call -- the last CTA of each candidate chunk writes
        // the 4 results straight into pinned host memory (unified addressing).  With few CTAs the
        // coefficients are read from pinned host memory as well (one launch + one sync in total);
        // with many CTAs each of them would cross PCIe, so they are copied to the device first.
        const Plan p = make_plan(ctx, r);
        const double* d_in = ctx->h_pin;
        if (p.n_chunks * p.n_ztiles * p.n_xchunks * p.n_ychunks > 64) {
            if ((rc = grow(ctx, &ctx->d_coeffs, &ctx->d_coeffs_cap, nc))) return rc;
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coeffs, ctx->h_pin, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            d_in = ctx->d_coeffs;
        }
        if ((rc = objective_device_impl(ctx, d_in, B, r, h_out, ctx->stream))) return rc;
        ctx->pend_timed = false;
    } else {
        if ((rc = grow(ctx, &ctx->d_coeffs, &ctx->d_coeffs_cap, nc))) return rc;
        if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, 4 * (size_t)B))) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coeffs, ctx->h_pin, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        if ((rc = objective_device_impl(ctx, ctx->d_coeffs, B, r, ctx->d_out, ctx->stream))) return rc;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(h_out, ctx->d_out, 4 * (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->pend_timed = true;
    }
    return SB200_OK;
}

static int objective_collect(sb200_ctx* ctx, double* S, double* Amin, double* Amax, int64_t* nan) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        ctx->tickets_dirty = true;   // the launch may have died midway: never trust the tickets again without zeroing them
        return fail(ctx, SB200_ECUDA, "objective batch failed on device %d: %s", ctx->device, cudaGetErrorString(e));
    }
    ctx->last_ms = 0.0;
    if (ctx->pend_timed) {
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_ms = ms;
    }
    const long long B = ctx->pend_B;
    const double* h_out = ctx->pend_out;
    memcpy(S, h_out, B * sizeof(double));
    memcpy(Amin, h_out + B, B * sizeof(double));
    memcpy(Amax, h_out + 2 * B, B * sizeof(double));
    for (long long b = 0; b < B; ++b) nan[b] = (int64_t)h_out[3 * B + b];
    ctx->pend_B = 0;
    return SB200_OK;
}

extern "C" int sb200_objective_batch_ex(sb200_ctx* ctx, const double* coeffs, int64_t B, const sb200_batch_opts* opts,
                                        double* S, double* Amin, double* Amax, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || B < 1 || !S || !Amin || !Amax || !nan) return fail(ctx, SB200_EINVAL, "bad arguments (nbatch=%lld)", (long long)B);
    Request r;
    int rc = resolve_request(ctx, B, opts, &r);
    if (rc) return rc;
    if ((rc = objective_enqueue(ctx, coeffs, B, r, opts && (opts->flags & SB200_F_AUTO_SCAN)))) return rc;
    return objective_collect(ctx, S, Amin, Amax, nan);
}

extern "C" int sb200_objective_batch(sb200_ctx* ctx, const double* coeffs, int64_t B, int32_t x_begin, int32_t x_end,
                                     double* S, double* Amin, double* Amax, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (x_begin >= x_end) return fail(ctx, SB200_EINVAL, "x range [%d,%d) is empty", x_begin, x_end);
    sb200_batch_opts o;
    range_opts(&o, x_begin, x_end);
    return sb200_objective_batch_ex(ctx, coeffs, B, &o, S, Amin, Amax, nan);
}

// ---- dense exports ------------------------------------------------------------------------------------------------
static int check_export_range(sb200_ctx* ctx, int32_t x_begin, int32_t x_end, int* xb, int* xe) {
    const int nx_user = ctx->n_user[0];
    if (x_begin < 0 || x_end > nx_user || x_begin >= x_end)
        return fail(ctx, SB200_EINVAL, "x range [%d,%d) outside [0,%d) or empty", x_begin, x_end, nx_user);
    if (ctx->dim == 2) {
        if (x_begin != 0 || x_end != nx_user) return fail(ctx, SB200_EINVAL, "x-slab ranges are supported for dim 3 only");
        *xb = 0; *xe = 1;
    } else {
        *xb = x_begin; *xe = x_end;
    }
    return SB200_OK;
}

static int export_device_impl(sb200_ctx* ctx, const double* coeffs_host, int what, int xb, int xe, double* d_out,
                              double* d_stats, cudaStream_t st) {
    int rc;
    const int threads = std::min(256, ((ctx->g.nz + 31) / 32) * 32);
    const int n_ztiles = (ctx->g.nz + threads - 1) / threads;
    const int n_ytiles = (ctx->g.ny + EXPORT_ROWS - 1) / EXPORT_ROWS;
    const long long ntiles = (long long)(xe - xb) * n_ytiles * n_ztiles;
    const int blocks = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 8);   // persistent: 8 CTAs per SM
    if ((rc = grow(ctx, (unsigned char**)&ctx->d_part, &ctx->d_part_cap, (size_t)blocks * 3 * 8))) return rc;
    ExportArgs a;
    // the coefficient vector travels in the kernel's parameter block: no staging buffer, nothing to synchronise
    for (int i = 0; i < 13; ++i) a.coeffs[i] = i < ctx->ncoef ? coeffs_host[i] : 0.0;
    a.ncoef = ctx->ncoef; a.what = what; a.x_begin = xb; a.x_end = xe;
    a.n_ytiles = n_ytiles; a.n_ztiles = n_ztiles;
    a.out = d_out; a.part = (unsigned long long*)ctx->d_part; a.ticket = ctx->d_xticket; a.stats = d_stats;
#define SB200_EXPORT(UNIT)                                                                  \
    do {                                                                                    \
        if (what == SB200_X_OMEGA) export_kernel<UNIT, 0><<<blocks, threads, 0, st>>>(ctx->g, a);        \
        else if (what == SB200_X_SQRTARG) export_kernel<UNIT, 1><<<blocks, threads, 0, st>>>(ctx->g, a); \
        else export_kernel<UNIT, 2><<<blocks, threads, 0, st>>>(ctx->g, a);                 \
    } while (0)
    if (ctx->unit_cell) SB200_EXPORT(true); else SB200_EXPORT(false);
#undef SB200_EXPORT
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SB200_ECUDA, "export kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches += 1;
    return SB200_OK;
}

extern "C" int sb200_export_device(sb200_ctx* ctx, const double* coeffs, int32_t what, int32_t x_begin, int32_t x_end,
                                   double* d_out, double* d_stats, void* stream) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || !d_out || what < 0 || what > 2) return fail(ctx, SB200_EINVAL, "bad arguments");
    int xb, xe, rc;
    if ((rc = check_export_range(ctx, x_begin, x_end, &xb, &xe))) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = export_device_impl(ctx, coeffs, what, xb, xe, d_out, d_stats, (cudaStream_t)stream))) return rc;
    return mark_user_stream(ctx, (cudaStream_t)stream);
}

// Device -> the caller's (pageable) host buffer at PCIe speed: chunks go to two pinned staging buffers by
// cudaMemcpyAsync, and while chunk k+1 crosses the bus the helper threads move chunk k from the staging buffer into
// the caller's array.  A plain cudaMemcpy into pageable memory stages through the driver's own small buffers on one
// thread and reaches about a quarter of the link rate.
namespace {
constexpr size_t X_CHUNK = (size_t)16 << 20;

struct CopyJob { void* dst; const void* src; size_t bytes; };
}

static int d2h_chunked(sb200_ctx* ctx, const std::vector<CopyJob>& jobs) {
    if (!ctx->h_xpin[0]) {
        for (int i = 0; i < 2; ++i) CUDA_TRY(ctx, cudaMallocHost((void**)&ctx->h_xpin[i], X_CHUNK));
        ctx->h_xpin_bytes = X_CHUNK;
    }
    struct Piece { char* dst; const char* src; size_t bytes; };
    std::vector<Piece> pieces;
    for (const CopyJob& j : jobs)
        for (size_t off = 0; off < j.bytes; off += X_CHUNK)
            pieces.push_back({(char*)j.dst + off, (const char*)j.src + off, std::min(X_CHUNK, j.bytes - off)});
    CopyPool& pool = CopyPool::get();
    const size_t n = pieces.size();
    for (size_t k = 0; k <= n; ++k) {
        if (k < n) {
            const int b = (int)(k & 1);
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_xpin[b], pieces[k].src, pieces[k].bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_x[b], ctx->stream));
        }
        if (k >= 1) {
            const int b = (int)((k - 1) & 1);
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_x[b]));
            pool.copy(pieces[k - 1].dst, ctx->h_xpin[b], pieces[k - 1].bytes);
        }
    }
    return SB200_OK;
}

extern "C" int sb200_export(sb200_ctx* ctx, const double* coeffs, int32_t what, int32_t x_begin, int32_t x_end,
                            double* out, double* Amin, double* Amax, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || !out || what < 0 || what > 2) return fail(ctx, SB200_EINVAL, "bad arguments");
    int xb, xe, rc;
    if ((rc = check_export_range(ctx, x_begin, x_end, &xb, &xe))) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if ((rc = join_user_stream(ctx))) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    const size_t total = (size_t)(xe - xb) * ctx->g.ny * ctx->g.nz;
    if ((rc = grow(ctx, &ctx->d_xbuf, &ctx->d_xbuf_cap, total))) return rc;
    if ((rc = export_device_impl(ctx, coeffs, what, xb, xe, ctx->d_xbuf, ctx->d_xstats, ctx->stream))) return rc;
    if ((rc = d2h_chunked(ctx, {{out, ctx->d_xbuf, total * sizeof(double)}}))) return rc;
    if ((rc = grow_pinned(ctx, 4))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->d_xstats, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->x_bytes = (double)total * sizeof(double);
    ctx->x_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (Amin) *Amin = ctx->h_pin[0];
    if (Amax) *Amax = ctx->h_pin[1];
    if (nan) *nan = (int64_t)ctx->h_pin[2];
    return SB200_OK;
}

extern "C" int sb200_group_velocity(sb200_ctx* ctx, const double* coeffs, const double* h, double* omega, double* vgc,
                                    double* vg, double* vph, double* k, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || !h) return fail(ctx, SB200_EINVAL, "coeffs and h must not be NULL");
    for (int a = 0; a < ctx->dim; ++a)
        if (ctx->n_user[a] < 3) return fail(ctx, SB200_EINVAL, "numpy.gradient(edge_order=2) needs >= 3 points per axis, axis %d has %d", a, ctx->n_user[a]);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc;
    if ((rc = join_user_stream(ctx))) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    const GridDev& g = ctx->g;
    const size_t total = (size_t)g.nx * g.ny * g.nz;
    const int off = ctx->dim == 2 ? 1 : 0;
    // device scratch: omega | comps (dim) | vg | vph | k
    const int nfields = 1 + ctx->dim + 3;
    if ((rc = grow(ctx, &ctx->d_xbuf, &ctx->d_xbuf_cap, total * nfields))) return rc;
    double* d_omega = ctx->d_xbuf;
    if ((rc = export_device_impl(ctx, coeffs, SB200_X_OMEGA, 0, g.nx, d_omega, ctx->d_xstats, ctx->stream))) return rc;
    GradArgs ga;
    ga.omega = d_omega;
    ga.naxes = ctx->dim;
    for (int a = 0; a < 3; ++a) { ga.comp[a] = nullptr; ga.ax[a] = AxisDiff{1, 1, 0, 0, 0, 0, 0, 0}; }
    for (int ua = 0; ua < ctx->dim; ++ua) {
        const int a = ua + off;
        const double hh = h[ua];
        AxisDiff d;
        d.den = 2. * hh; d.rden = 1.0 / d.den;                       // numpy: / (2. * ax_dx)
        d.a0 = -1.5 / hh; d.b0 = 2. / hh; d.c0 = -0.5 / hh;          // numpy: a, b, c of the near edge
        d.a1 = 0.5 / hh; d.b1 = -2. / hh; d.c1 = 1.5 / hh;           // ... and of the far edge
        ga.ax[a] = d;
        ga.comp[a] = vgc ? d_omega + total * (1 + ua) : nullptr;
    }
    ga.vg = vg ? d_omega + total * (1 + ctx->dim) : nullptr;
    ga.vph = vph ? d_omega + total * (2 + ctx->dim) : nullptr;
    ga.k = k ? d_omega + total * (3 + ctx->dim) : nullptr;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->sm_count * 16);
    gradient_kernel<<<blocks, 256, 0, ctx->stream>>>(g, ga);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SB200_ECUDA, "gradient kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches += 1;
    const size_t bytes = total * sizeof(double);
    std::vector<CopyJob> jobs;
    if (omega) jobs.push_back({omega, d_omega, bytes});
    if (vgc) jobs.push_back({vgc, d_omega + total, bytes * ctx->dim});
    if (vg) jobs.push_back({vg, ga.vg, bytes});
    if (vph) jobs.push_back({vph, ga.vph, bytes});
    if (k) jobs.push_back({k, ga.k, bytes});
    if ((rc = d2h_chunked(ctx, jobs))) return rc;
    if ((rc = grow_pinned(ctx, 4))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pin, ctx->d_xstats, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double moved = 0.0;
    for (const CopyJob& j : jobs) moved += (double)j.bytes;
    ctx->x_bytes = moved;
    ctx->x_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (nan) *nan = (int64_t)ctx->h_pin[2];
    return SB200_OK;
}

extern "C" int sb200_last_export_stats(const sb200_ctx* ctx, double* bytes, double* seconds) {
    if (!ctx || !bytes || !seconds) return SB200_EINVAL;
    *bytes = ctx->x_bytes;
    *seconds = ctx->x_seconds;
    return SB200_OK;
}

extern "C" int sb200_get_stream(const sb200_ctx* ctx, void** stream) {
    if (!ctx || !stream) return SB200_EINVAL;
    *stream = (void*)ctx->stream;
    return SB200_OK;
}

extern "C" int sb200_launch_count(const sb200_ctx* ctx, int64_t* launches) {
    if (!ctx || !launches) return SB200_EINVAL;
    *launches = ctx->launches;
    return SB200_OK;
}

extern "C" int sb200_last_batch_ms(const sb200_ctx* ctx, double* ms) {
    if (!ctx || !ms) return SB200_EINVAL;
    *ms = ctx->last_ms;
    return SB200_OK;
}

// ---- device groups ------------------------------------------------------------------------------------------------
struct sb200_group {
    std::vector<sb200_ctx*> members;
    std::vector<double> tmp;      // slab partials of members 1..G-1
    std::vector<int64_t> tmpn;
    char err[512] = "";
};

static int gfail(sb200_group* grp, int code, const char* fmt, ...) {
    char* dst = grp ? grp->err : g_group_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* sb200_group_last_error(const sb200_group* grp) { return grp ? grp->err : g_group_error; }

extern "C" void sb200_group_destroy(sb200_group* grp) {
    if (!grp) return;
    for (sb200_ctx* c : grp->members) sb200_destroy(c);
    delete grp;
}

static int group_create_impl(sb200_group** out, const sb200_grid_desc* desc, const int32_t* devices, int32_t ndevices);
static int group_objective_batch_impl(sb200_group* grp, const double* coeffs, int64_t B, int32_t shard,
                                      const sb200_batch_opts* opts, double* S, double* Amin, double* Amax, int64_t* nan,
                                      int32_t* used);

// The group functions use std::vector and std::thread: no exception may cross the C boundary.
extern "C" int sb200_group_create(sb200_group** out, const sb200_grid_desc* desc, const int32_t* devices, int32_t ndevices) {
    try {
        return group_create_impl(out, desc, devices, ndevices);
    } catch (const std::bad_alloc&) {
        return gfail(nullptr, SB200_ENOMEM, "out of host memory");
    } catch (...) {
        return gfail(nullptr, SB200_ECUDA, "unexpected C++ exception while creating the device group");
    }
}

extern "C" int sb200_group_objective_batch(sb200_group* grp, const double* coeffs, int64_t B, int32_t shard,
                                           const sb200_batch_opts* opts, double* S, double* Amin, double* Amax,
                                           int64_t* nan, int32_t* used) {
    try {
        return group_objective_batch_impl(grp, coeffs, B, shard, opts, S, Amin, Amax, nan, used);
    } catch (const std::bad_alloc&) {
        return gfail(grp, SB200_ENOMEM, "out of host memory");
    } catch (...) {
        return gfail(grp, SB200_ECUDA, "unexpected C++ exception in the device group");
    }
}

static int group_create_impl(sb200_group** out, const sb200_grid_desc* desc, const int32_t* devices, int32_t ndevices) {
    if (!out) return gfail(nullptr, SB200_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = check_desc(desc, g_group_error);
    if (rc) return rc;
    std::vector<int> devs;
    if (!devices) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            return gfail(nullptr, SB200_ECUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        for (int i = 0; i < n; ++i) devs.push_back(i);
    } else {
        if (ndevices < 1) return gfail(nullptr, SB200_EINVAL, "ndevices must be >= 1");
        devs.assign(devices, devices + ndevices);
        for (size_t i = 0; i < devs.size(); ++i)
            for (size_t j = 0; j < i; ++j)
                if (devs[i] == devs[j]) return gfail(nullptr, SB200_EINVAL, "device %d is listed twice", devs[i]);
    }
    sb200_group* grp = new (std::nothrow) sb200_group();
    if (!grp) return gfail(nullptr, SB200_ENOMEM, "out of host memory");
    // One thread per device: creating a context (primary context, stream, pinned staging, table upload) takes 0.15-0.2 s
    // each, and the devices do not wait for one another.
    const size_t G = devs.size();
    grp->members.assign(G, nullptr);
    std::vector<int> rcs(G, SB200_OK);
    std::vector<std::string> msgs(G);
    auto create_one = [&](size_t i) {
        sb200_grid_desc d = *desc;
        d.device = devs[i];
        rcs[i] = sb200_create(&grp->members[i], &d);
        if (rcs[i]) msgs[i] = sb200_last_error(nullptr);       // the creating thread's own message buffer
    };
    {
        std::vector<std::thread> pool;
        std::vector<size_t> inline_jobs(1, 0);
        for (size_t i = 1; i < G; ++i) {
            try {
                pool.emplace_back(create_one, i);
            } catch (...) {               // no thread to be had: this one is created by the caller's thread
                inline_jobs.push_back(i);
            }
        }
        for (size_t i : inline_jobs) create_one(i);
        for (std::thread& t : pool) t.join();
    }
    for (size_t i = 0; i < G; ++i)
        if (rcs[i]) {
            rc = rcs[i];
            gfail(nullptr, rc, "device %d: %s", devs[i], msgs[i].c_str());
            std::vector<sb200_ctx*> made;
            for (sb200_ctx* c : grp->members)
                if (c) made.push_back(c);
            grp->members.swap(made);
            sb200_group_destroy(grp);
            return rc;
        }
    *out = grp;
    return SB200_OK;
}

extern "C" int sb200_group_size(const sb200_group* grp, int32_t* n) {
    if (!grp || !n) return SB200_EINVAL;
    *n = (int32_t)grp->members.size();
    return SB200_OK;
}

extern "C" int sb200_group_member(sb200_group* grp, int32_t i, sb200_ctx** ctx) {
    if (!grp || !ctx) return SB200_EINVAL;
    if (i < 0 || i >= (int32_t)grp->members.size()) return gfail(grp, SB200_EINVAL, "member %d out of range", i);
    *ctx = grp->members[i];
    return SB200_OK;
}

extern "C" int sb200_group_set_weight(sb200_group* grp, int32_t kind, const double* w, int64_t count) {
    if (!grp) return SB200_EINVAL;
    for (sb200_ctx* c : grp->members) {
        int rc = sb200_set_weight(c, kind, w, count);
        if (rc) return gfail(grp, rc, "device %d: %s", c->device, c->err);
    }
    return SB200_OK;
}

static int group_objective_batch_impl(sb200_group* grp, const double* coeffs, int64_t B, int32_t shard,
                                      const sb200_batch_opts* opts, double* S, double* Amin, double* Amax, int64_t* nan,
                                      int32_t* used) {
    if (!grp) return SB200_EINVAL;
    if (!coeffs || B < 1 || !S || !Amin || !Amax || !nan) return gfail(grp, SB200_EINVAL, "bad arguments (nbatch=%lld)", (long long)B);
    sb200_ctx* c0 = grp->members[0];
    const int G = (int)grp->members.size();
    Request r;
    int rc = resolve_request(c0, B, opts, &r);
    if (rc) return gfail(grp, rc, "%s", c0->err);
    const bool auto_scan = opts && (opts->flags & SB200_F_AUTO_SCAN);
    const double plane = (double)c0->g.ny * c0->g.nz;
    const double evals = (double)B * r.nplanes * plane;
    if (shard == SB200_SHARD_AUTO) {
        // a device is worth waking for >= 2^27 evaluations (~0.2 ms of kernel time); below that one device is faster
        const double per_dev = 134217728.0;
        if (G == 1 || evals < 2 * per_dev) shard = SB200_SHARD_NONE;
        else if (r.nseg >= 2 || B >= 8LL * G) shard = SB200_SHARD_CANDIDATES;
        else if (c0->dim == 3 && r.nplanes >= 8 * G) shard = SB200_SHARD_SLABS;
        else shard = SB200_SHARD_NONE;
    }
    if ((shard == SB200_SHARD_SLABS || shard == SB200_SHARD_SLABS_CONTIGUOUS) && c0->dim != 3)
        return gfail(grp, SB200_EINVAL, "k-slab sharding is a 3-D feature");
    int ndev = 1;
    if (shard == SB200_SHARD_NONE || G == 1) {
        shard = SB200_SHARD_NONE;
        if ((rc = objective_enqueue(c0, coeffs, B, r, auto_scan))) return gfail(grp, rc, "device %d: %s", c0->device, c0->err);
        if ((rc = objective_collect(c0, S, Amin, Amax, nan))) return gfail(grp, rc, "device %d: %s", c0->device, c0->err);
    } else if (shard == SB200_SHARD_CANDIDATES) {
        // contiguous blocks of whole segments, first blocks larger (sharding.split_range)
        const long long units = r.nseg > 1 ? r.nseg : B, unit = r.nseg > 1 ? r.seg : 1;
        ndev = (int)std::min<long long>(G, units);
        std::vector<long long> lo(ndev + 1, 0);
        for (int g = 0; g < ndev; ++g) lo[g + 1] = lo[g] + (units / ndev + (g < units % ndev ? 1 : 0)) * unit;
        for (int g = 0; g < ndev; ++g) {
            sb200_ctx* c = grp->members[g];
            Request rg = r;
            const long long n = lo[g + 1] - lo[g];
            if (r.nseg > 1) rg.nseg = n / r.seg; else { rg.seg = n; rg.nseg = 1; }
            if ((rc = objective_enqueue(c, coeffs + lo[g] * c->ncoef, n, rg, auto_scan)))
                return gfail(grp, rc, "device %d: %s", c->device, c->err);
        }
        int first = SB200_OK;
        for (int g = 0; g < ndev; ++g) {   // collect from every device even after a failure: nothing stays in flight
            sb200_ctx* c = grp->members[g];
            rc = objective_collect(c, S + lo[g], Amin + lo[g], Amax + lo[g], nan + lo[g]);
            if (rc && !first) first = gfail(grp, rc, "device %d: %s", c->device, c->err);
        }
        if (first) return first;
    } else if (shard == SB200_SHARD_SLABS || shard == SB200_SHARD_SLABS_CONTIGUOUS) {
        ndev = std::min(G, r.nplanes);
        std::vector<Request> rq(ndev, r);
        int done = 0;
        for (int g = 0; g < ndev; ++g) {
            if (shard == SB200_SHARD_SLABS) {      // planes g, g + G, ... of the list
                rq[g].xb = r.xb + g * r.stride;
                rq[g].stride = r.stride * ndev;
                rq[g].nplanes = (r.nplanes - g + ndev - 1) / ndev;
            } else {
                const int n = r.nplanes / ndev + (g < r.nplanes % ndev ? 1 : 0);
                rq[g].xb = r.xb + done * r.stride;
                rq[g].nplanes = n;
                done += n;
            }
            sb200_ctx* c = grp->members[g];
            if ((rc = objective_enqueue(c, coeffs, B, rq[g], auto_scan))) return gfail(grp, rc, "device %d: %s", c->device, c->err);
        }
        grp->tmp.resize(3 * (size_t)B);
        grp->tmpn.resize((size_t)B);
        int first = SB200_OK;
        for (int g = 0; g < ndev; ++g) {
            sb200_ctx* c = grp->members[g];
            if (g == 0) {
                rc = objective_collect(c, S, Amin, Amax, nan);
            } else {
                double* tS = grp->tmp.data(), *tlo = tS + B, *thi = tlo + B;
                rc = objective_collect(c, tS, tlo, thi, grp->tmpn.data());
                if (!rc) {
                    // fixed device order: sum / min / max / sum, a NaN in either extremum poisons both
                    for (long long b = 0; b < B; ++b) {
                        S[b] += tS[b];
                        nan[b] += grp->tmpn[b];
                        const bool bad = Amin[b] != Amin[b] || Amax[b] != Amax[b] || tlo[b] != tlo[b] || thi[b] != thi[b];
                        Amin[b] = bad ? NAN : std::min(Amin[b], tlo[b]);
                        Amax[b] = bad ? NAN : std::max(Amax[b], thi[b]);
                    }
                }
            }
            if (rc && !first) first = gfail(grp, rc, "device %d: %s", c->device, c->err);
        }
        if (first) return first;
    } else {
        return gfail(grp, SB200_EINVAL, "unknown shard mode %d", shard);
    }
    if (used) { used[0] = shard; used[1] = ndev; }
    return SB200_OK;
}

extern "C" int sb200_fp64_peak(int32_t device, double seconds, double* tflops) {
    if (!tflops) return SB200_EINVAL;
    *tflops = 0.0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev)
        return fail(nullptr, SB200_ECUDA, "no usable CUDA device %d (%s)", device, e != cudaSuccess ? cudaGetErrorString(e) : "out of range");
#define PK(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) return fail(nullptr, SB200_ECUDA, "%s: %s", #call, cudaGetErrorString(e2)); } while (0)
    PK(cudaSetDevice(device));
    cudaDeviceProp prop;
    PK(cudaGetDeviceProperties(&prop, device));
    const int threads = 256;
    double* d = nullptr;
    PK(cudaMalloc((void**)&d, (size_t)prop.multiProcessorCount * 8 * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    PK(cudaEventCreate(&e0)); PK(cudaEventCreate(&e1));
    if (!(seconds > 0)) seconds = 0.2;
    // shapes: (chains per thread, CTAs of 256 threads per SM).  4 chains x 4 warps per sub-partition and
    // 8 chains x 2 or 4 warps reach one warp instruction per 2.00 cycles in tools/ubench_latency.cu.
    // Few, long launches (one warm-up and three timed ones per shape, ~seconds/20 each): a launch list of a
    // program that calls this stays readable.
    struct Shape { int chains, ctas_per_sm; };
    const Shape shapes[] = {{8, 2}, {4, 2}, {8, 1}, {4, 4}, {8, 4}};
    const bool verbose = getenv("SB200_PEAK_VERBOSE") != nullptr;
    double best = 0.0;
    for (const Shape& s : shapes) {
        const int blocks = prop.multiProcessorCount * s.ctas_per_sm;
        // iterations for about seconds/20 per launch at the nominal rate (64 DFMA per thread and iteration)
        const double per_iter = 2.0 * 64.0 * (double)blocks * threads;
        const int iters = (int)std::max(1024.0, std::min(4.0e6, seconds / 20.0 * 37.0e12 / per_iter));
        const double flop_per_launch = per_iter * iters;
        auto launch = [&]() {
            if (s.chains == 8) dfma_peak_kernel<8><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
            else dfma_peak_kernel<4><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        };
        launch();
        PK(cudaDeviceSynchronize());
        double here = 0.0;
        for (int rep = 0; rep < 3; ++rep) {
            PK(cudaEventRecord(e0));
            launch();
            PK(cudaEventRecord(e1));
            PK(cudaEventSynchronize(e1));
            float ms = 0.f;
            PK(cudaEventElapsedTime(&ms, e0, e1));
            here = std::max(here, flop_per_launch / (ms * 1e-3) / 1e12);
        }
        if (verbose) fprintf(stderr, "sb200_fp64_peak: %d chains x %d CTAs/SM: %.2f TFLOP/s\n", s.chains, s.ctas_per_sm, here);
        best = std::max(best, here);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
#undef PK
    *tflops = best;
    return SB200_OK;
}
