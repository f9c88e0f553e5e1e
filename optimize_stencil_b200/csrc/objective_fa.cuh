// objective_fa.cuh -- the batched objective with a REGROUPED sqrtarg in the hot loop ("fast A").
//
// objective_kernel (dispersion_kernels.cuh) evaluates sqrtarg in the reference's own operation order at
// every k-point: 8 FP64 instructions (18 for a non-unit cell) whose only purpose is that min/max and the
// on-CFL classification are bit-identical to NumPy's.  But bit-identity is needed at very few points:
//   (1) where sqrtarg can be the grid maximum            (dt_ok, dispersion.py:149)
//   (2) where sqrtarg can be negative                    (stencil_ok, dispersion.py:292-302)
//   (3) where s = dt*sqrt(sqrtarg) is within 2^-10 of 1  (asin is ill-conditioned there; NaN iff s > 1)
// Everywhere else omega only has to agree to 1e-13, and sqrtarg regroups by what its terms depend on:
//
//   A = R(x,y) + T(x,z) + tzy(z) * sy2r(y) + bzy(y) * sz2r(z)                        3 FP64 instructions
//       R = axy*sx2 + ayx*sy2r      (row table, built once per (x, y) by one lane)
//       T = tzx*sx2 + tzzx*sz2r     (registers, rebuilt once per x-plane)
//       sy2r = sy2/Y2, sz2r = sz2/Z2 (folded into the tables: a non-unit cell costs nothing extra)
//
// A' (regrouped) and A (reference order) both approximate the real-number value with an error below
// 8u * M, M = sum of the magnitudes of all products <= Mbound(candidate); |A' - A| <= 2^-49 Mbound.
// A lane asks for the exact value when the high word of A' is
//     >= thr   = min(high word of the thread's running exact maximum - 1, high word of the CFL threshold)
//  or <  LO    = high word of 2^-40        (also true for every negative or NaN A': sign bit / all-ones)
// in ONE unsigned compare of (hi(A') - LO) against (thr - LO).  If any lane of the warp asks, the whole
// iteration is redone from the tables in the reference's order (sqrtarg_exact_at) and only those exact
// values feed the extrema.  With rows walked from high x, y to low this happens about once per plane.
//
// The guards are sufficient when  Mbound <= 64  (2^-46 Mbound <= 2^-40),
//                                 Mbound <= 2^23 * A(last visited corner)  (one high-word unit of margin
//                                        at the maximum exceeds 2 * 2^-46 Mbound),
//                                 dt^2 Mbound <= 2^30  (the CFL threshold's 2^-12 slack exceeds the error of v).
// A candidate that fails any of them (huge or cancelling coefficients, NaNs, negative corner) makes its
// CTA take the exact path at every point: same results as objective_kernel, at about half the speed.
#pragma once
#include "dispersion_kernels.cuh"

namespace sb200 {

#define SB200_FA_LO 0x3D700000u   // high word of 2^-40

// sqrtarg at one k-point from the per-axis tables, reference order (dispersion.py:394-398).  Rare path.
template <bool UNIT>
__device__ __noinline__ double sqrtarg_exact_at(const Cand* c, double cx, double cy, double cz, double sx2,
                                                double sy2, double sz2, double Y2, double rY2, double Z2,
                                                double rZ2) {
    const double opx = __dadd_rn(1.0, __dmul_rn(2.0, cx));
    const double opy = __dadd_rn(1.0, __dmul_rn(2.0, cy));
    const double opz = __dadd_rn(1.0, __dmul_rn(2.0, cz));
    const double axy = __dadd_rn(__dadd_rn(c->ax, __dmul_rn(c->dlx, opx)), __dmul_rn(c->bxy2, cy));
    const double ayx = __dadd_rn(__dadd_rn(c->ay, __dmul_rn(c->dly, opy)), __dmul_rn(c->byx2, cx));
    const double tzzx = __dadd_rn(__dadd_rn(c->az, __dmul_rn(c->dlz, opz)), __dmul_rn(c->bzx2, cx));
    const CellScale cs = {Y2, rY2, Z2, rZ2};
    return sqrtarg_point<UNIT>(axy, ayx, __dmul_rn(c->bzy2, cy), __dmul_rn(c->bxz2, cz), __dmul_rn(c->byz2, cz),
                               tzzx, sx2, sy2, sz2, cs);
}

// Loads the compiler must not hoist out of the rare branch (they would cost registers in the hot loop).
__device__ __forceinline__ double ld_rare(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int C>
struct RowLayoutFA {
    static constexpr int ROW = 2 + 2 * C;   // (sy2r, kx2+ky2) then (R, bzy) per candidate
    static constexpr int ROWS = 32;
};

// Shared memory: Cand[C] | guards[C] (uint2: cfl_hi, force) | rows | s_red | s_redu | s_umax | s_smax
template <int C, bool UNIT, int WK>
__global__ void __launch_bounds__(SB200_MAXTHREADS, (C <= 4 ? SB200_MINBLOCKS_C4 : SB200_MINBLOCKS_C8))
objective_kernel_fa(const GridDev g, const BatchArgs a) {
    constexpr int ROW = RowLayoutFA<C>::ROW;
    constexpr int ROWS = RowLayoutFA<C>::ROWS;
    constexpr int U = (C >= 4) ? 1 : 4 / C;   // rows per iteration
    constexpr int J = U * C;                  // independent chains per iteration
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cand* cand = reinterpret_cast<Cand*>(smem_raw);
    uint2* guard = reinterpret_cast<uint2*>(cand + C);                                   // [C] (+ pad to 16 B)
    double* s_rows = reinterpret_cast<double*>(guard + ((C + 1) & ~1));                  // [nwarps][ROWS][ROW]
    const int nwarps = (blockDim.x + 31) >> 5;
    double* s_red = s_rows + (size_t)nwarps * ROWS * ROW;                                // [C][8]
    unsigned long long* s_redu = reinterpret_cast<unsigned long long*>(s_red + C * 8);   // [3][C][8]
    unsigned long long* s_umax = s_redu + 3 * C * 8;   // [C][blockDim.x] most negative exact value (raw bits)
    long long* s_smax = reinterpret_cast<long long*>(s_umax + (size_t)C * blockDim.x);   // [C][blockDim.x] max(exact, +0)
    __shared__ unsigned int s_ticket;
    __shared__ int s_force;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x / a.n_ztiles;
    const int ztile = blockIdx.x - chunk * a.n_ztiles;
    const int xs = a.x_begin + blockIdx.y * a.tile_x;
    const int xe = min(xs + a.tile_x, a.x_end);
    const int ys = blockIdx.z * a.tile_y;
    const int ye = min(ys + a.tile_y, g.ny);
    const long long b0 = (long long)chunk * C;
    const double rY = UNIT ? 1.0 : g.rY2, rZ = UNIT ? 1.0 : g.rZ2;

    if (tid == 0) s_force = 0;
    __syncthreads();
    if (tid < C) {
        long long b = b0 + tid;
        if (b >= a.B) b = a.B - 1;  // pad the last chunk with a copy; its results are not written
        Cand& c = cand[tid];
        load_candidate(c, a.coeffs + b * a.ncoef, a.ncoef);
        // Are the guards of the header sufficient for this candidate?  (every test is false for NaNs)
        const double mb = (fabs(c.ax) + 3. * fabs(c.dlx) + fabs(c.bxy2) + fabs(c.bxz2)) +
                          (fabs(c.ay) + 3. * fabs(c.dly) + fabs(c.byx2) + fabs(c.byz2)) * rY +
                          (fabs(c.az) + 3. * fabs(c.dlz) + fabs(c.bzx2) + fabs(c.bzy2)) * rZ;
        const int xc = a.x_end - 1, yc = g.ny - 1, zc = g.nz - 1;
        const double corner = sqrtarg_exact_at<UNIT>(&c, g.cx[xc], g.cy[yc], g.cz[zc], g.sx2[xc], g.sy2[yc],
                                                     g.sz2[zc], g.Y2, g.rY2, g.Z2, g.rZ2);
        const bool sane = (mb <= 64.) && (mb <= 8388608. * corner) && (c.dt2 * mb <= 1073741824.);
        // sqrtarg from which v = 2 dt^2 A - 1 may reach 1 - 2^-9 (the EXACT class), with 2^-12 of slack
        const double acfl = (2. - 0x1p-9 - 0x1p-12) / c.c2dt.x;
        unsigned int cfl_hi = 0x7ff00000u;
        if (acfl >= 0. && acfl < __longlong_as_double(0x7ff0000000000000ll)) cfl_hi = (unsigned int)__double2hiint(acfl);
        guard[tid] = make_uint2(cfl_hi, sane ? 0u : 1u);
        if (!sane) atomicOr(&s_force, 1);
    }
    __syncthreads();
    const bool force = s_force != 0;   // CTA-uniform

    // Per-thread state.  T and tzy hold the z-dependent parts for the thread's current z; span is the
    // filter threshold minus LO (0 = "always ask", the initial state).
    double T[C], tzy[C], S[C];
    unsigned int span[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        S[c] = 0.; span[c] = 0u;
        s_umax[c * blockDim.x + tid] = 0ull;   // own slots only: no barrier needed
        s_smax[c * blockDim.x + tid] = 0ll;
    }

    double* wrows = s_rows + (size_t)warp * ROWS * ROW;   // this warp's private table
    for (int x = xe - 1; x >= xs; --x) {
        // Load balance: the cost of a k-point depends on its class, i.e. on |k|, i.e. on z.  Warp w
        // therefore takes z slice (w + x) mod nwarps of the CTA's z tile for plane x, so that over an
        // x range every warp sees every z slice and all warps of the CTA finish together.
        const int zslice = (warp + (x - a.x_begin)) % nwarps;
        const int z = ztile * blockDim.x + zslice * 32 + lane;
        const bool active = z < g.nz;
        const unsigned int amask = __ballot_sync(0xffffffffu, active);   // lanes that own a z index
        if (amask == 0u) continue;                                        // warp-uniform
        const double cx = g.cx[x], sx2 = g.sx2[x], kx2 = g.kx2[x];
        const double opx = __dadd_rn(1.0, __dmul_rn(2.0, cx));
        double sz2r = 0., kz2 = 0.;
        if (active) {
            const double cz = g.cz[z];
            sz2r = __dmul_rn(g.sz2[z], rZ);
            kz2 = g.kz2[z];
            const double opz = __dadd_rn(1.0, __dmul_rn(2.0, cz));
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double tzx = __dmul_rn(cand[c].bxz2, cz);
                const double tzzx = fma(cand[c].bzx2, cx, fma(cand[c].dlz, opz, cand[c].az));
                tzy[c] = __dmul_rn(cand[c].byz2, cz);
                T[c] = fma(tzx, sx2, __dmul_rn(tzzx, sz2r));
            }
        }
        for (int y1 = ye; y1 > ys; y1 -= ROWS) {          // y tiles from the top of the range downwards
            const int y0 = max(ys, y1 - ROWS);
            const int nrows = y1 - y0;
            __syncwarp();                      // every lane is done reading the previous tile
            {                                  // lane l builds row l (all 32 lanes, active or not);
                const int y = min(y0 + lane, y1 - 1);   // rows past the tile's end repeat its last row
                const double cy = g.cy[y];
                const double opy = __dadd_rn(1.0, __dmul_rn(2.0, cy));
                const double sy2r = __dmul_rn(g.sy2[y], rY);
                double* row = wrows + lane * ROW;
                *reinterpret_cast<double2*>(row) = make_double2(sy2r, __dadd_rn(kx2, g.ky2[y]));
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double axy = fma(cand[c].bxy2, cy, fma(cand[c].dlx, opx, cand[c].ax));
                    const double ayx = fma(cand[c].byx2, cx, fma(cand[c].dly, opy, cand[c].ay));
                    double2 v;
                    v.x = fma(axy, sx2, __dmul_rn(ayx, sy2r));
                    v.y = __dmul_rn(cand[c].bzy2, cy);
                    *reinterpret_cast<double2*>(row + 2 + 2 * c) = v;
                }
            }
            __syncwarp();
            if (!active) continue;
            const double* wrow = nullptr;
            if (WK == 1) wrow = g.w + (size_t)y0 * g.nz + z;
            if (WK == 2) wrow = g.w + ((size_t)x * g.ny + y0) * g.nz + z;
            // U rows x C candidates = J independent evaluation chains per iteration (J = 4 for C <= 4): the
            // FP64 pipe has an 8-cycle dependent-issue latency at one instruction per 2 cycles.
            const int nit = (nrows + U - 1) / U;
            const double* row = wrows + (size_t)(nit - 1) * U * ROW;
#pragma unroll 1
            for (int yy = (nit - 1) * U; yy >= 0; yy -= U, row -= U * ROW) {
                double A[J], v[J], e[J], hk[J], w[U];
                bool need = force, all_mid = true;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double* r = row + u * ROW;
                    const double2 yk = *reinterpret_cast<const double2*>(r);   // (sy2r, kx2+ky2)
                    w[u] = 1.0;
                    if (WK != 0) w[u] = __ldg(wrow + (size_t)min(yy + u, nrows - 1) * g.nz);
                    const double k = sqrt_fast(__dadd_rn(yk.y, kz2));           // dispersion.py:350
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = u * C + c;
                        const double2 t = *reinterpret_cast<const double2*>(r + 2 + 2 * c);   // (R, bzy)
                        A[j] = fma(tzy[c], yk.x, fma(t.y, sz2r, __dadd_rn(t.x, T[c])));
                        need = need || ((unsigned int)__double2hiint(A[j]) - SB200_FA_LO >= span[c]);
                        const double2 cd = cand[c].c2dt;                        // (2 dt^2, dt), warp-broadcast
                        v[j] = fma(cd.x, A[j], -1.0);                           // 2 s^2 - 1
                        hk[j] = fma(-k, cd.y, SB200_PI_2);                      // pi/2 - k dt
                        all_mid = all_mid && v_is_mid(v[j]);
                    }
                }
                if (__any_sync(amask, need)) {
                    // Rare: redo the iteration in the reference's operation order and let the exact values
                    // (and only those) into the extrema.
                    const double czr = ld_rare(g.cz + z), sz2 = ld_rare(g.sz2 + z);
                    all_mid = true;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int y = y0 + min(yy + u, nrows - 1);
                        const double cyr = ld_rare(g.cy + y), sy2 = ld_rare(g.sy2 + y);
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int j = u * C + c;
                            const double Ae = sqrtarg_exact_at<UNIT>(cand + c, cx, cyr, czr, sx2, sy2, sz2, g.Y2, g.rY2,
                                                                     g.Z2, g.rZ2);
                            const long long bits = __double_as_longlong(Ae);
                            long long* sm = s_smax + c * blockDim.x + tid;
                            *sm = max(*sm, bits);                               // max(sqrtarg, +0)
                            if (bits < 0) {                                     // negative or sign-bit NaN
                                unsigned long long* um = s_umax + c * blockDim.x + tid;
                                *um = max(*um, (unsigned long long)bits);
                            }
                            A[j] = Ae;
                            v[j] = fma(cand[c].c2dt.x, Ae, -1.0);
                            all_mid = all_mid && v_is_mid(v[j]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const unsigned int hm = (unsigned int)(s_smax[c * blockDim.x + tid] >> 32);
                        const unsigned int thr = min(max(hm, 1u) - 1u, guard[c].x);
                        span[c] = max(thr, SB200_FA_LO) - SB200_FA_LO;
                    }
                }
                scaled_dev_chains<C, J>(A, v, hk, e, cand, amask, all_mid);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (U == 1 || yy + u < nrows) {       // rows past the end are repeats: skip their sum only
#pragma unroll
                        for (int c = 0; c < C; ++c) {     // dispersion.py:251, times dt^2
                            const int j = u * C + c;
                            S[c] = (WK == 0) ? fma(e[j], e[j], S[c]) : fma(__dmul_rn(w[u], e[j]), e[j], S[c]);
                        }
                    }
                }
            }
        }
    }

    // ---- CTA reduction: fixed shuffle tree, then warp leaders in order -------------------------
#pragma unroll
    for (int c = 0; c < C; ++c) {
        double s = S[c];
        unsigned long long um = s_umax[c * blockDim.x + tid];
        long long sm = s_smax[c * blockDim.x + tid];
        // A NaN omega makes the thread's sum NaN (w*(NaN-k)^2 accumulates by fma); the per-candidate
        // NaN indicator is therefore taken from the partial sums instead of being counted per point.
        unsigned long long nn = (s != s) ? 1ull : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, off));
            um = max(um, __shfl_down_sync(0xffffffffu, um, off));
            sm = max(sm, __shfl_down_sync(0xffffffffu, sm, off));
            nn += __shfl_down_sync(0xffffffffu, nn, off);
        }
        if (lane == 0) {
            s_red[c * 8 + warp] = s;
            s_redu[(0 * C + c) * 8 + warp] = um;
            s_redu[(1 * C + c) * 8 + warp] = (unsigned long long)sm;
            s_redu[(2 * C + c) * 8 + warp] = nn;
        }
    }
    __syncthreads();
    const int slot = (ztile * a.n_xchunks + blockIdx.y) * a.n_ychunks + blockIdx.z;
    if (tid < C) {
        double s = s_red[tid * 8];
        unsigned long long um = s_redu[(0 * C + tid) * 8], nn = s_redu[(2 * C + tid) * 8];
        long long sm = (long long)s_redu[(1 * C + tid) * 8];
        for (int wq = 1; wq < nwarps; ++wq) {
            s = __dadd_rn(s, s_red[tid * 8 + wq]);
            um = max(um, s_redu[(0 * C + tid) * 8 + wq]);
            sm = max(sm, (long long)s_redu[(1 * C + tid) * 8 + wq]);
            nn += s_redu[(2 * C + tid) * 8 + wq];
        }
        const size_t o = (size_t)(b0 + tid) * a.P + slot;
        a.part_S[o] = s; a.part_min[o] = um; a.part_max[o] = (unsigned long long)sm; a.part_nan[o] = nn;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(&a.tickets[chunk], 1u);
    __syncthreads();
    if (s_ticket != (unsigned int)(a.P - 1)) return;

    // ---- last CTA of this candidate chunk: fixed-order reduction over the P slots ---------------
    __threadfence();
    for (int c = warp; c < C; c += nwarps) {
        const long long b = b0 + c;
        if (b >= a.B) continue;
        const size_t o = (size_t)b * a.P;
        double s = 0.;
        Extrema e;
        ext_init(e);
        unsigned long long nn = 0ull;
        for (int p = lane; p < a.P; p += 32) {
            s = __dadd_rn(s, __ldcg(a.part_S + o + p));
            ext_merge(e, __ldcg(a.part_min + o + p), (long long)__ldcg(a.part_max + o + p));
            nn += __ldcg(a.part_nan + o + p);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, off));
            ext_merge(e, __shfl_down_sync(0xffffffffu, e.umax, off), __shfl_down_sync(0xffffffffu, e.smax, off));
            nn += __shfl_down_sync(0xffffffffu, nn, off);
        }
        if (lane == 0) {
            double amin, amax;
            ext_final(e, amin, amax);
            a.out[0 * a.B + b] = __ddiv_rn(s, cand[c].dt2);   // the sum was taken over (dt*(omega-k))^2
            a.out[1 * a.B + b] = amin;
            a.out[2 * a.B + b] = amax;
            a.out[3 * a.B + b] = (double)nn;
        }
    }
    if (tid == 0) a.tickets[chunk] = 0u;  // re-arm for the next launch
}

}  // namespace sb200
