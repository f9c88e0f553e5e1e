// stencil_b200_capi.cu -- the C-ABI declared in include/stencil_b200.h (context, launches, staging).
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
#if SB200_FAST_A
#define SB200_OBJECTIVE_KERNEL objective_kernel_fa
#else
#define SB200_OBJECTIVE_KERNEL objective_kernel
#endif

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

using namespace sb200;

namespace {
thread_local char g_create_error[512] = "";
}

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
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // staging
    double* h_pin = nullptr;   size_t h_pin_cap = 0;   // pinned host, doubles
    double* d_coeffs = nullptr; size_t d_coeffs_cap = 0;
    double* d_out = nullptr;    size_t d_out_cap = 0;
    void* d_part = nullptr;     size_t d_part_cap = 0;  // bytes
    unsigned int* d_tickets = nullptr; size_t d_tickets_cap = 0;
    unsigned int* d_xticket = nullptr;
    double* d_xstats = nullptr;
    double* d_xbuf = nullptr;   size_t d_xbuf_cap = 0;
    int sm_count = 148;
    int force_c = 0, force_ty = 0, force_tx = 0;
    long long launches = 0;
    double last_ms = 0.0;
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
    size_t want = std::max(need, (size_t)1024);
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
    if (prop.major < 10)
        return fail(ctx, SB200_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", d->device,
                    prop.major, prop.minor);

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
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_xticket, sizeof(unsigned int)));
    CUDA_TRY(ctx, cudaMemset(ctx->d_xticket, 0, sizeof(unsigned int)));
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_xstats, 3 * sizeof(double)));
    return SB200_OK;
}

extern "C" int sb200_create(sb200_ctx** out, const sb200_grid_desc* d) {
    if (!out) return fail(nullptr, SB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (!d) return fail(nullptr, SB200_EINVAL, "desc is NULL");
    if (d->dim != 2 && d->dim != 3) return fail(nullptr, SB200_EINVAL, "dim must be 2 or 3, got %d", d->dim);
    for (int a = 0; a < d->dim; ++a) {
        if (d->n[a] < 1) return fail(nullptr, SB200_EINVAL, "n[%d] = %d must be >= 1", a, d->n[a]);
        if (!d->cos_kappa[a] || !d->s2[a] || !d->kcomp[a]) return fail(nullptr, SB200_EINVAL, "table %d is NULL", a);
    }
    if (!(d->Y > 0) || (d->dim == 3 && !(d->Z > 0))) return fail(nullptr, SB200_EINVAL, "Y and Z must be positive");
    sb200_ctx* ctx = new (std::nothrow) sb200_ctx();
    if (!ctx) return fail(nullptr, SB200_ENOMEM, "out of host memory");
    int rc = create_impl(ctx, d);
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
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
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
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int rc = grow(ctx, &ctx->d_w, &ctx->w_count, need);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_w, w, need * sizeof(double), cudaMemcpyHostToDevice));
    ctx->wkind = internal; ctx->g.w = ctx->d_w;
    return SB200_OK;
}

extern "C" int sb200_set_tiling(sb200_ctx* ctx, int32_t c, int32_t ty, int32_t tx) {
    if (!ctx) return SB200_EINVAL;
    if (!(c == 0 || c == 1 || c == 2 || c == 4 || c == 8)) return fail(ctx, SB200_EINVAL, "cand_per_thread must be 0,1,2,4,8");
    if (ty < 0 || tx < 0 || ty > 1024) return fail(ctx, SB200_EINVAL, "bad tile");
    ctx->force_c = c; ctx->force_ty = ty; ctx->force_tx = tx;
    return SB200_OK;
}

namespace {
struct Plan {
    int C, TZ, TY, TX, n_ztiles, n_xchunks, n_ychunks, P;
    long long n_chunks;
    size_t smem;
};

size_t smem_bytes(int C, int TY, int TZ) {
    const int nwarps = (TZ + 31) / 32;
    (void)TY;                                // tables are per warp (32 rows), not per y tile
#if SB200_FAST_A
    const int row = 2 + 2 * C;               // RowLayoutFA<C>::ROW
    return (size_t)C * sizeof(Cand) + (size_t)nwarps * 32 * row * 8 +
           (size_t)C * 8 * 8 * 4 + (size_t)C * TZ * 8 * 2;
#else
    const int row = ((2 + 3 * C) + 1) & ~1;  // RowLayout<C>::ROW
    return (size_t)C * sizeof(Cand) + (size_t)nwarps * 32 * row * 8 + (size_t)C * 8 * 8 * 4 + (size_t)C * TZ * 8;
#endif
}

Plan make_plan(const sb200_ctx* ctx, long long B, int nxr) {
    const GridDev& g = ctx->g;
    Plan p;
    // candidates per thread: 4 from B = 3 on (a padded fourth candidate costs less than the 2 x 2 layout's second
    // pass over the tables: 52 vs 59 us for 3 x 128^3)
    p.C = ctx->force_c ? ctx->force_c : (B >= 3 ? 4 : (B >= 2 ? 2 : 1));
    if (const char* e = getenv("SB200_FORCE_C")) {         // experiments only
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) p.C = v;
    }
    int tz_max = 256;
    if (const char* e = getenv("SB200_BLOCK_THREADS")) {   // experiments only (tools/sweep.py)
        const int v = atoi(e);
        if (v >= 32 && v <= 256 && v % 32 == 0) tz_max = v;
    }
    p.TZ = std::min(tz_max, ((g.nz + 31) / 32) * 32);
    p.n_ztiles = (g.nz + p.TZ - 1) / p.TZ;
    p.n_chunks = (B + p.C - 1) / p.C;
    p.TY = ctx->force_ty ? std::min(ctx->force_ty, g.ny) : g.ny;
    p.TX = ctx->force_tx ? std::min(ctx->force_tx, nxr) : nxr;
    // Enough CTAs for ~16 waves of (2 CTAs x SMs) when the problem is that large; never below 8
    // k-points of work per thread and candidate.
    const long long target = (long long)ctx->sm_count * 2 * 16;
    auto ctas = [&]() {
        return p.n_chunks * p.n_ztiles * ((nxr + p.TX - 1) / p.TX) * ((g.ny + p.TY - 1) / p.TY);
    };
    if (!ctx->force_tx)
        while (ctas() < target && p.TX > 1) p.TX = (p.TX + 1) / 2;
    if (!ctx->force_ty)
        while (ctas() < target && p.TY > 32) p.TY = ((p.TY + 1) / 2 + 31) / 32 * 32;   // whole 32-row tiles
    p.n_xchunks = (nxr + p.TX - 1) / p.TX;
    p.n_ychunks = (g.ny + p.TY - 1) / p.TY;
    p.P = p.n_ztiles * p.n_xchunks * p.n_ychunks;
    p.smem = smem_bytes(p.C, p.TY, p.TZ);
    return p;
}

template <int C>
cudaError_t launch_objective(const sb200_ctx* ctx, const Plan& p, const BatchArgs& a, cudaStream_t st) {
    dim3 grid((unsigned)(p.n_chunks * p.n_ztiles), (unsigned)p.n_xchunks, (unsigned)p.n_ychunks), block(p.TZ);
#define SB200_LAUNCH(UNIT, WK)                                                                              \
    do {                                                                                                    \
        auto kern = SB200_OBJECTIVE_KERNEL<C, UNIT, WK>;                                                    \
        if (p.smem > 48 * 1024) {                                                                           \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem); \
            if (e != cudaSuccess) return e;                                                                 \
        }                                                                                                   \
        kern<<<grid, block, p.smem, st>>>(ctx->g, a);                                                       \
    } while (0)
    if (ctx->unit_cell) {
        if (ctx->wkind == 0) SB200_LAUNCH(true, 0); else if (ctx->wkind == 1) SB200_LAUNCH(true, 1); else SB200_LAUNCH(true, 2);
    } else {
        if (ctx->wkind == 0) SB200_LAUNCH(false, 0); else if (ctx->wkind == 1) SB200_LAUNCH(false, 1); else SB200_LAUNCH(false, 2);
    }
#undef SB200_LAUNCH
    return cudaGetLastError();
}
}  // namespace

static int check_range(sb200_ctx* ctx, int32_t x_begin, int32_t x_end) {
    const int nx_user = ctx->n_user[0];
    if (x_begin < 0 || x_end > nx_user || x_begin >= x_end)
        return fail(ctx, SB200_EINVAL, "x range [%d,%d) outside [0,%d) or empty", x_begin, x_end, nx_user);
    return SB200_OK;
}

// For dim 2 the caller's first axis is the internal second axis: an x-range there is a y-range of
// the internal view, which the kernels do not slice.  Slabs are a 3-D feature.
static int internal_range(sb200_ctx* ctx, int32_t x_begin, int32_t x_end, int* xb, int* xe) {
    int rc = check_range(ctx, x_begin, x_end);
    if (rc) return rc;
    if (ctx->dim == 2) {
        if (x_begin != 0 || x_end != ctx->n_user[0])
            return fail(ctx, SB200_EINVAL, "x-slab ranges are supported for dim 3 only");
        *xb = 0; *xe = 1;
    } else {
        *xb = x_begin; *xe = x_end;
    }
    return SB200_OK;
}

static int objective_device_impl(sb200_ctx* ctx, const double* d_coeffs, int64_t B, int xb, int xe, double* d_out,
                                 cudaStream_t st) {
    const long long max_chunks = 1 << 20;
    Plan p = make_plan(ctx, B, xe - xb);
    if (p.n_chunks * p.n_ztiles > 0x7fffffffLL || p.n_xchunks > 65535 || p.n_ychunks > 65535 || p.n_chunks > max_chunks * 64)
        return fail(ctx, SB200_EINVAL, "batch of %lld candidates is too large for one launch; split it", (long long)B);
    const size_t bpad = (size_t)p.n_chunks * p.C;
    const size_t part_bytes = bpad * p.P * 8 * 4;
    int rc = grow(ctx, (unsigned char**)&ctx->d_part, &ctx->d_part_cap, part_bytes);
    if (rc) return rc;
    if (ctx->d_tickets_cap < (size_t)p.n_chunks) {
        rc = grow(ctx, &ctx->d_tickets, &ctx->d_tickets_cap, (size_t)p.n_chunks);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_tickets, 0, ctx->d_tickets_cap * sizeof(unsigned int), st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));   // rare (growth only): later calls may come on another stream
    }
    BatchArgs a;
    a.coeffs = d_coeffs; a.ncoef = ctx->ncoef; a.B = B; a.x_begin = xb; a.x_end = xe;
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
        case 1: e = launch_objective<1>(ctx, p, a, st); break;
        case 2: e = launch_objective<2>(ctx, p, a, st); break;
        case 4: e = launch_objective<4>(ctx, p, a, st); break;
        default: e = launch_objective<8>(ctx, p, a, st); break;
    }
    if (e != cudaSuccess) return fail(ctx, SB200_ECUDA, "objective kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches += 1;
    return SB200_OK;
}

extern "C" int sb200_objective_batch_device(sb200_ctx* ctx, const double* d_coeffs, int64_t nbatch, int32_t x_begin,
                                            int32_t x_end, double* d_out, void* stream) {
    if (!ctx) return SB200_EINVAL;
    if (!d_coeffs || !d_out || nbatch < 1) return fail(ctx, SB200_EINVAL, "bad arguments (nbatch=%lld)", (long long)nbatch);
    int xb, xe;
    int rc = internal_range(ctx, x_begin, x_end, &xb, &xe);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return objective_device_impl(ctx, d_coeffs, nbatch, xb, xe, d_out, (cudaStream_t)stream);
}

extern "C" int sb200_objective_batch(sb200_ctx* ctx, const double* coeffs, int64_t B, int32_t x_begin, int32_t x_end,
                                     double* S, double* Amin, double* Amax, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || B < 1 || !S || !Amin || !Amax || !nan) return fail(ctx, SB200_EINVAL, "bad arguments (nbatch=%lld)", (long long)B);
    int xb, xe;
    int rc = internal_range(ctx, x_begin, x_end, &xb, &xe);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t nc = (size_t)B * ctx->ncoef;
    if ((rc = grow_pinned(ctx, nc + 4 * (size_t)B))) return rc;
    double* h_out = ctx->h_pin + nc;
    memcpy(ctx->h_pin, coeffs, nc * sizeof(double));
    const double evals = (double)B * (double)(xe - xb) * ctx->g.ny * ctx->g.nz;
    if (B <= 64 && evals < 67108864.0) {
        // Latency path (SciPy's SLSQP asks for one iterate and its finite-difference neighbours at a
        // time): no event records and no D2H copy call -- the last CTA of each candidate chunk writes
        // the 4 results straight into pinned host memory (unified addressing).  With few CTAs the
        // coefficients are read from pinned host memory as well (one launch + one sync in total);
        // with many CTAs each of them would cross PCIe, so they are copied to the device first.
        const Plan p = make_plan(ctx, B, xe - xb);
        const double* d_in = ctx->h_pin;
        if (p.n_chunks * p.n_ztiles * p.n_xchunks * p.n_ychunks > 64) {
            if ((rc = grow(ctx, &ctx->d_coeffs, &ctx->d_coeffs_cap, nc))) return rc;
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coeffs, ctx->h_pin, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            d_in = ctx->d_coeffs;
        }
        if ((rc = objective_device_impl(ctx, d_in, B, xb, xe, h_out, ctx->stream))) return rc;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->last_ms = 0.0;
    } else {
        if ((rc = grow(ctx, &ctx->d_coeffs, &ctx->d_coeffs_cap, nc))) return rc;
        if ((rc = grow(ctx, &ctx->d_out, &ctx->d_out_cap, 4 * (size_t)B))) return rc;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coeffs, ctx->h_pin, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        if ((rc = objective_device_impl(ctx, ctx->d_coeffs, B, xb, xe, ctx->d_out, ctx->stream))) return rc;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(h_out, ctx->d_out, 4 * (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->last_ms = ms;
    }
    memcpy(S, h_out, B * sizeof(double));
    memcpy(Amin, h_out + B, B * sizeof(double));
    memcpy(Amax, h_out + 2 * B, B * sizeof(double));
    for (int64_t b = 0; b < B; ++b) nan[b] = (int64_t)h_out[3 * B + b];
    return SB200_OK;
}

static int export_device_impl(sb200_ctx* ctx, const double* coeffs_host, int what, int xb, int xe, double* d_out,
                              double* d_stats, cudaStream_t st) {
    int rc;
    if ((rc = grow_pinned(ctx, 16))) return rc;
    if ((rc = grow(ctx, &ctx->d_coeffs, &ctx->d_coeffs_cap, (size_t)16))) return rc;
    CUDA_TRY(ctx, cudaStreamSynchronize(st));  // the pinned staging row may still be in flight
    memcpy(ctx->h_pin, coeffs_host, ctx->ncoef * sizeof(double));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coeffs, ctx->h_pin, ctx->ncoef * sizeof(double), cudaMemcpyHostToDevice, st));
    const int threads = std::min(256, ((ctx->g.nz + 31) / 32) * 32);
    const int n_ztiles = (ctx->g.nz + threads - 1) / threads;
    const int n_ytiles = (ctx->g.ny + EXPORT_ROWS - 1) / EXPORT_ROWS;
    const long long ntiles = (long long)(xe - xb) * n_ytiles * n_ztiles;
    const int blocks = (int)std::min<long long>(ntiles, (long long)ctx->sm_count * 8);   // persistent: 8 CTAs per SM
    if ((rc = grow(ctx, (unsigned char**)&ctx->d_part, &ctx->d_part_cap, (size_t)blocks * 3 * 8))) return rc;
    ExportArgs a;
    a.coeffs = ctx->d_coeffs; a.ncoef = ctx->ncoef; a.what = what; a.x_begin = xb; a.x_end = xe;
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
    if ((rc = internal_range(ctx, x_begin, x_end, &xb, &xe))) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return export_device_impl(ctx, coeffs, what, xb, xe, d_out, d_stats, (cudaStream_t)stream);
}

extern "C" int sb200_export(sb200_ctx* ctx, const double* coeffs, int32_t what, int32_t x_begin, int32_t x_end,
                            double* out, double* Amin, double* Amax, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || !out || what < 0 || what > 2) return fail(ctx, SB200_EINVAL, "bad arguments");
    int xb, xe, rc;
    if ((rc = internal_range(ctx, x_begin, x_end, &xb, &xe))) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t total = (size_t)(xe - xb) * ctx->g.ny * ctx->g.nz;
    if ((rc = grow(ctx, &ctx->d_xbuf, &ctx->d_xbuf_cap, total))) return rc;
    if ((rc = export_device_impl(ctx, coeffs, what, xb, xe, ctx->d_xbuf, ctx->d_xstats, ctx->stream))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->d_xbuf, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    double st[3];
    CUDA_TRY(ctx, cudaMemcpyAsync(st, ctx->d_xstats, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (Amin) *Amin = st[0];
    if (Amax) *Amax = st[1];
    if (nan) *nan = (int64_t)st[2];
    return SB200_OK;
}

extern "C" int sb200_group_velocity(sb200_ctx* ctx, const double* coeffs, const double* h, double* omega, double* vgc,
                                    double* vg, double* vph, double* k, int64_t* nan) {
    if (!ctx) return SB200_EINVAL;
    if (!coeffs || !h) return fail(ctx, SB200_EINVAL, "coeffs and h must not be NULL");
    for (int a = 0; a < ctx->dim; ++a)
        if (ctx->n_user[a] < 3) return fail(ctx, SB200_EINVAL, "numpy.gradient(edge_order=2) needs >= 3 points per axis, axis %d has %d", a, ctx->n_user[a]);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const GridDev& g = ctx->g;
    const size_t total = (size_t)g.nx * g.ny * g.nz;
    const int off = ctx->dim == 2 ? 1 : 0;
    int rc;
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
    if (omega) CUDA_TRY(ctx, cudaMemcpyAsync(omega, d_omega, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (vgc) CUDA_TRY(ctx, cudaMemcpyAsync(vgc, d_omega + total, bytes * ctx->dim, cudaMemcpyDeviceToHost, ctx->stream));
    if (vg) CUDA_TRY(ctx, cudaMemcpyAsync(vg, ga.vg, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (vph) CUDA_TRY(ctx, cudaMemcpyAsync(vph, ga.vph, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (k) CUDA_TRY(ctx, cudaMemcpyAsync(k, ga.k, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    double st[3];
    CUDA_TRY(ctx, cudaMemcpyAsync(st, ctx->d_xstats, sizeof st, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (nan) *nan = (int64_t)st[2];
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
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* d = nullptr;
    PK(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    PK(cudaEventCreate(&e0)); PK(cudaEventCreate(&e1));
    const double flop_per_launch = 2.0 * 64.0 * iters * (double)blocks * threads;
    for (int i = 0; i < 3; ++i) dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    PK(cudaDeviceSynchronize());
    double best = 0.0, spent = 0.0;
    if (!(seconds > 0)) seconds = 0.2;
    while (spent < seconds) {
        PK(cudaEventRecord(e0));
        for (int i = 0; i < 4; ++i) dfma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        PK(cudaEventRecord(e1));
        PK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::max(best, 4.0 * flop_per_launch / (ms * 1e-3) / 1e12);
        spent += ms * 1e-3;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
#undef PK
    *tflops = best;
    return SB200_OK;
}
