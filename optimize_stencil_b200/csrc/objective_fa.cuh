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
// A' (regrouped) and A (reference order) both approximate the real-number value; each goes through at most 11
// roundings of quantities bounded by M = sum of the magnitudes of all products <= Mbound(candidate), so
// |A' - A| <= 22u * Mbound < 2^-48 Mbound (u = 2^-53); a plain-rounding emulation over 5 million points
// peaks at 2^-50.4.  The guards below are sized for 2^-47 Mbound.
// All four tables carry the factor 2 dt^2, so the three instructions give W' = 2 dt^2 A' and v = W' - 1.
// A lane asks for the exact value
//   (1) when the high word of W' reaches thr = high word of (2 dt^2 * the thread's running exact maximum) - 2
//       -- ONE signed integer compare per evaluation, the only extra work in the common (MID) path;
//   (2,3) outside the MID class only: v <= -1 + 2^-21 (or a sign-bit NaN), 1 - 2^-9 - 2^-12 <= v < 1 + 2^-12
//       (beyond that s > 1 and omega is NaN whatever the last bits of sqrtarg are).
// If any lane of the warp asks, the whole iteration is redone from the tables in the reference's order
// (sqrtarg_exact_at) and only those exact values feed the extrema.  With rows walked from high x, y to low
// this happens about once per plane.
//
// The guards are sufficient when  Mbound <= 2^24 * Amax       (two high-word units at the maximum, >= 2^-21
//                                        relative, exceed twice the error 2^-47 Mbound plus the rounding of
//                                        2 dt^2 * maximum)
//                                 2 dt^2 Mbound <= 2^24     (the error of W' stays below 2^-23: a negative
//                                        sqrtarg gives W' < 2^-22, well inside the v <= -1 + 2^-21 guard, and
//                                        the CFL band keeps 2^-12 of slack on either side).
// The kernel asks for much more, Mbound <= 16 * P with P = the largest sqrtarg on the 8 corners of the
// visited slab (a lower bound of Amax): that also keeps the relative error of A' (and so of omega, 1e-13
// being the bar) at a few 1e-15 -- real stencils have Mbound / P of 1 to 3.
// A candidate that fails (huge or cancelling coefficients, NaNs, non-positive corners) makes its
// CTA take the exact path at every point: same results as objective_kernel, at about a quarter of the speed.
#pragma once
#include "dispersion_kernels.cuh"
#include <climits>

#ifndef SB200_ALLMID_VIMAX
#define SB200_ALLMID_VIMAX (SB200_INNER != 0)   // "every chain is MID" from three-input integer maxima of the high words of v^2
#endif

namespace sb200 {

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

// ------------------------------------------------------------------------------------------------------------
// Extrema-only sweep for a chunk whose candidates ALL violate the CFL limit at one of the 8 corners of the visited
// slab (2 dt^2 sqrtarg >= 2.001 there, i.e. s > 1, i.e. omega is NaN at that k-point).  The sum of such a candidate is
// NaN whatever the rest of the grid holds, so only min / max of sqrtarg remain to be found: the regrouped value and
// the two filters per k-point (3 FP64 + 2 integer compares instead of ~20 FP64), the reference-order redo where a
// lane asks -- the same exact values reach the extrema as in the full sweep, and S = NaN, nan > 0 are what the full
// sweep returns too.  Start grids (optimize.py:118-134: most starts lie beyond the CFL limit) and the first rounds of
// an SLSQP scan consist of such candidates.  A separate function (not inlined) so that its register allocation and
// code motion stay out of the hot loop's.
// ------------------------------------------------------------------------------------------------------------
template <int C, bool UNIT>
__device__ __noinline__ void extrema_only_sweep(const GridDev& g, const BatchArgs& a, const Cand* cand, const double* s_mb,
                                                const double (*s_corner)[8], double* wrows, unsigned long long* s_umax,
                                                long long* s_smax, int ztile, int is, int ie, int ys, int ye) {
    constexpr int ROW = 2 + 2 * C, ROWS = 32;
    constexpr int U = (C >= 8) ? 1 : 8 / C;      // rows per iteration: 6 to 8 independent chains in flight per thread
    constexpr unsigned int FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (blockDim.x + 31) >> 5;
    const double rY = UNIT ? 1.0 : g.rY2, rZ = UNIT ? 1.0 : g.rZ2;
    double Ts[C], tzys[C];
    // Both filters compare doubles here (two DSETPs per point, chained on one predicate: no branches, no integer
    // work): a point is passed over iff  tmin < 2 dt^2 A' < tmax,  anything else -- NaNs included -- asks for the exact
    // value.  With M the regrouping error bound 2^-47 Mbound, and 2 dt^2 Mbound <= 2^24 as in the full sweep:
    //   tmax = (1 - 2^-20) * 2 dt^2 * (running exact maximum) once that maximum is >= 2^-20 Mbound (the slack then
    //          exceeds the error by 2^7); until then every value that may be positive asks: tmax = -2^-40 * 2 dt^2
    //          Mbound, so that the all-negative part of the grid of a stencil gone negative (78 % of it for the
    //          iterates an SLSQP line search throws at the delta = 1/4 bound) is passed over without a single exact
    //          evaluation.  Unlike the full sweep's CTA-wide `force` this is decided as the sweep goes: such a stencil
    //          has no positive corner to vouch for its maximum beforehand, but the first positive values it meets do.
    //   tmin = (1 - 2^-20) * 2 dt^2 * (running exact minimum) once that minimum is <= -2^-20 Mbound; until then every
    //          value that may be negative asks: tmin = 2^-21 (the error is below 2^-23 on this scale).
    // -inf / +inf = "always ask" (NaN or absurdly scaled coefficients).
    double tmax[C], tmin[C];
    const double minus_inf = __longlong_as_double(0xfff0000000000000ll);
#pragma unroll
    for (int c = 0; c < C; ++c) { tmax[c] = minus_inf; tmin[c] = -minus_inf; }
    auto refresh = [&]() {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double c2 = cand[c].c2dt.x, tiny = 9.5367431640625e-07 * s_mb[c];          // 2^-20 Mbound
            const bool scale_ok = c2 * s_mb[c] <= 16777216.;                                  // false for NaNs
            const double amax = __longlong_as_double(s_smax[c * blockDim.x + tid]);
            const double amin = __longlong_as_double((long long)s_umax[c * blockDim.x + tid]);   // +0 if none yet
            const double wmax = __dmul_rn(c2, amax), wmin = __dmul_rn(c2, amin);
            tmax[c] = !scale_ok ? minus_inf
                                : (amax >= tiny ? fma(-9.5367431640625e-07, wmax, wmax) : -9.094947017729282e-13 * (c2 * s_mb[c]));
            tmin[c] = !scale_ok ? -minus_inf : (amin <= -tiny ? fma(-9.5367431640625e-07, wmin, wmin) : 4.76837158203125e-07);
        }
    };
    // The 8 corners of the visited slab are grid points, evaluated exactly in the prologue: they seed every thread's
    // running extrema (for a Yee-like stencil the far corner IS the maximum, for a stencil gone negative it is the
    // minimum), and the corner with the largest value tells from which end to walk x and y -- the filters fire while
    // the values still rise along the walk, so the walk starts where sqrtarg is largest.
    int qmax = 0;
    {
        double best = 0.;
#pragma unroll
        for (int q = 1; q < 8; ++q)
            if (s_corner[0][q] > best) { best = s_corner[0][q]; qmax = q; }
    }
    const bool xup = !(qmax & 1), yup = !(qmax & 2);          // the largest corner sits at the low end: walk upwards
#pragma unroll
    for (int c = 0; c < C; ++c) {
        long long sm = 0ll;
        unsigned long long um = 0ull;
#pragma unroll
        for (int q = 1; q < 8; ++q) {
            const long long bits = __double_as_longlong(s_corner[c][q]);
            sm = max(sm, bits);
            if (bits < 0) um = max(um, (unsigned long long)bits);
        }
        s_smax[c * blockDim.x + tid] = sm;
        s_umax[c * blockDim.x + tid] = um;
    }
    refresh();
    for (int ii = 0; ii < ie - is; ++ii) {
        const int i = xup ? is + ii : ie - 1 - ii;
        const int x = a.x_begin + i * a.x_stride;
        const int zslice = (warp + i) % nwarps;
        const int zraw = ztile * blockDim.x + zslice * 32 + lane;
        if (__ballot_sync(FULL, zraw < g.nz) == 0u) continue;
        const int z = min(zraw, g.nz - 1);                 // lanes past the end repeat the last point (extrema unaffected)
        const double cz = g.cz[z], sz2 = g.sz2[z];
        const double sz2r = __dmul_rn(sz2, rZ);
        const double opz = __dadd_rn(1.0, __dmul_rn(2.0, cz));
        const double cx = g.cx[x], sx2 = g.sx2[x];
        const double opx = __dadd_rn(1.0, __dmul_rn(2.0, cx));
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double c2 = cand[c].c2dt.x;
            const double tzx = __dmul_rn(cand[c].bxz2, cz);
            const double tzzx = fma(cand[c].bzx2, cx, fma(cand[c].dlz, opz, cand[c].az));
            tzys[c] = __dmul_rn(c2, __dmul_rn(cand[c].byz2, cz));
            Ts[c] = __dmul_rn(c2, fma(tzx, sx2, __dmul_rn(tzzx, sz2r)));
        }
        for (int t = 0; t * ROWS < ye - ys; ++t) {
            const int y0 = yup ? ys + t * ROWS : max(ys, ye - (t + 1) * ROWS);
            const int y1 = yup ? min(ye, y0 + ROWS) : ye - t * ROWS;
            const int nrows = y1 - y0;
            __syncwarp();
            {
                const int y = min(y0 + lane, y1 - 1);
                const double cy = g.cy[y];
                const double opy = __dadd_rn(1.0, __dmul_rn(2.0, cy));
                const double sy2r = __dmul_rn(g.sy2[y], rY);
                double* row = wrows + lane * ROW;
                *reinterpret_cast<double2*>(row) = make_double2(sy2r, 0.0);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double c2 = cand[c].c2dt.x;
                    const double axy = fma(cand[c].bxy2, cy, fma(cand[c].dlx, opx, cand[c].ax));
                    const double ayx = fma(cand[c].byx2, cx, fma(cand[c].dly, opy, cand[c].ay));
                    *reinterpret_cast<double2*>(row + 2 + 2 * c) =
                        make_double2(__dmul_rn(c2, fma(axy, sx2, __dmul_rn(ayx, sy2r))), __dmul_rn(c2, __dmul_rn(cand[c].bzy2, cy)));
                }
            }
            __syncwarp();
            const int nit = (nrows + U - 1) / U;
#pragma unroll 1
            for (int it = 0; it < nit; ++it) {
                const int yy = (yup ? it : nit - 1 - it) * U;
                // one predicate per chain (two compares each), combined afterwards: a single running predicate would
                // string all 2 U C compares into one dependent chain (ncu: "wait" stalls at 2.6 warps per issue slot)
                bool ok[U * C];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double* r = wrows + (size_t)min(yy + u, nrows - 1) * ROW;
                    const double sy2r = r[0];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const double2 q = *reinterpret_cast<const double2*>(r + 2 + 2 * c);
                        const double wv = fma(tzys[c], sy2r, fma(q.y, sz2r, __dadd_rn(q.x, Ts[c])));
                        ok[u * C + c] = (wv < tmax[c]) & (wv > tmin[c]);
                    }
                }
#pragma unroll
                for (int w = 1; w < U * C; w *= 2)
#pragma unroll
                    for (int j = 0; j + w < U * C; j += 2 * w) ok[j] = ok[j] & ok[j + w];
                const bool need = !ok[0];
                if (!__any_sync(FULL, need)) continue;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int y = y0 + min(yy + u, nrows - 1);
                    const double cyr = g.cy[y], sy2 = g.sy2[y];
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const double Ae = sqrtarg_exact_at<UNIT>(cand + c, cx, cyr, cz, sx2, sy2, sz2, g.Y2, g.rY2, g.Z2, g.rZ2);
                        const long long bits = __double_as_longlong(Ae);
                        long long* sm = s_smax + c * blockDim.x + tid;
                        *sm = max(*sm, bits);
                        if (bits < 0) {
                            unsigned long long* um = s_umax + c * blockDim.x + tid;
                            *um = max(*um, (unsigned long long)bits);
                        }
                    }
                }
                refresh();
            }
        }
    }
}

// Shared memory: Cand[C] | rows | s_red | s_redu | s_umax | s_smax
//
// SCAN: the variant for start-grid screening (Optimize's first pass over the whole start grid, optimize.py:118-134
// of the reference, where most candidates violate the CFL limit at most k-points).  A warp whose 32 lanes x J chains
// all have v >= 1 + 2^-12 -- s > 1, omega is NaN whatever the last bits of sqrtarg are -- skips the evaluation and
// takes NaN directly; the extrema were already dealt with by then.  Results are bit-identical to the plain variant;
// it is a separate instantiation because the extra merge point costs the plain MID loop two register moves per
// iteration (measured: 682 -> 672 G evals/s on the 4096 x 512^3 batch).
//
// PREC: 1 the precise series (search launches), 0 the fast one (throughput launches) -- dispersion_kernels.cuh.
template <int C, bool UNIT, int WK, bool SCAN, int PREC>
__global__ void __launch_bounds__(SB200_MAXTHREADS, (C <= 4 ? SB200_MINBLOCKS_C4 : SB200_MINBLOCKS_C8))
objective_kernel_fa(const GridDev g, const BatchArgs a) {
    constexpr int ROW = RowLayoutFA<C>::ROW;
    constexpr int ROWS = RowLayoutFA<C>::ROWS;
    constexpr int U = (C >= 4) ? 1 : 4 / C;   // rows per iteration
    constexpr int J = U * C;                  // independent chains per iteration
    constexpr unsigned int FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cand* cand = reinterpret_cast<Cand*>(smem_raw);
    double* s_rows = reinterpret_cast<double*>(cand + C);                                // [nwarps][ROWS][ROW]
    const int nwarps = (blockDim.x + 31) >> 5;
    double* s_red = s_rows + (size_t)nwarps * ROWS * ROW;                                // [C][8]
    unsigned long long* s_redu = reinterpret_cast<unsigned long long*>(s_red + C * 8);   // [3][C][8]
    unsigned long long* s_umax = s_redu + 3 * C * 8;   // [C][blockDim.x] most negative exact value (raw bits)
    long long* s_smax = reinterpret_cast<long long*>(s_umax + (size_t)C * blockDim.x);   // [C][blockDim.x] max(exact, +0)
    __shared__ int s_force;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x / a.n_ztiles;
    const int ztile = blockIdx.x - chunk * a.n_ztiles;
    const int is = blockIdx.y * a.tile_x;                 // plane indices [is, ie) of the visited plane list
    const int ie = min(is + a.tile_x, a.n_planes);
    const int ys = blockIdx.z * a.tile_y;
    const double rY = UNIT ? 1.0 : g.rY2, rZ = UNIT ? 1.0 : g.rZ2;

    __shared__ long long s_probe[8];
    __shared__ int s_alive, s_negcorner;
    __shared__ double s_mb[8];
    __shared__ double s_corner[8][8];        // [candidate][corner q]: sqrtarg on the corners of the visited slab
    if (tid == 0) { s_force = 0; s_alive = 0; s_negcorner = 0; }
    if (tid < C) {
        const int seg = chunk / a.chunks_per_seg;
        const long long bend = (long long)(seg + 1) * a.seg_size;
        long long b = (long long)seg * a.seg_size + (long long)(chunk - seg * a.chunks_per_seg) * C + tid;
        if (b >= bend) b = bend - 1;  // pad the segment's last chunk with a copy; its results are not written
        load_candidate(cand[tid], a.coeffs + b * a.ncoef, a.ncoef);
        s_probe[tid] = 0ll;
    }
    __syncthreads();
    // Are the guards of the header sufficient for these candidates?  A lower bound of the grid maximum first: the
    // largest sqrtarg on the 8 corners of the visited slab (the maximum sits on the far corner for Yee-like stencils,
    // on an axis end for smoothed ones like Lehe's) -- one thread per (candidate, corner), the maximum taken on the
    // raw bits (non-negative doubles order like integers; a negative value never beats the initial +0, a positive NaN
    // beats everything and then fails the test below, which is the safe side).
    for (int t = tid; t < 7 * C; t += blockDim.x) {
        const int c = t / 7, q = t - 7 * c + 1;
        const int xc = (q & 1) ? a.x_begin + (a.n_planes - 1) * a.x_stride : a.x_begin, yc = (q & 2) ? g.ny - 1 : 0, zc = (q & 4) ? g.nz - 1 : 0;
        const double val = sqrtarg_exact_at<UNIT>(cand + c, g.cx[xc], g.cy[yc], g.cz[zc], g.sx2[xc], g.sy2[yc], g.sz2[zc],
                                                  g.Y2, g.rY2, g.Z2, g.rZ2);
        atomicMax(&s_probe[c], __double_as_longlong(val));
        if (val < 0.) atomicOr(&s_negcorner, 1 << c);   // sqrt of a negative sqrtarg: omega is NaN at that corner
        if (SCAN || PREC == 1) s_corner[c][q] = val;    // (only the variants with the extrema-only path read them)
    }
    __syncthreads();
    if (tid < C) {
        const Cand& c = cand[tid];
        // Mbound: |cos| <= 1, |1 + 2 cos| <= 3, 0 <= s2 <= 1 for the reference's tables; tab_scale covers a
        // caller who hands over anything else (it is 1 otherwise, NaN for NaN tables).  Every test is false for NaNs.
        const double mb = ((fabs(c.ax) + 3. * fabs(c.dlx) + fabs(c.bxy2) + fabs(c.bxz2)) +
                           (fabs(c.ay) + 3. * fabs(c.dly) + fabs(c.byx2) + fabs(c.byz2)) * rY +
                           (fabs(c.az) + 3. * fabs(c.dlz) + fabs(c.bzx2) + fabs(c.bzy2)) * rZ) * g.tab_scale;
        const double probe = __longlong_as_double(s_probe[tid]);
        const bool sane = (mb <= 16. * probe) && (c.c2dt.x * mb <= 16777216.);
        if (!sane) atomicOr(&s_force, 1);
        // s > 1 at a corner for sure (2 dt^2 A >= 2.001; false for NaNs), or sqrtarg < 0 at a corner: omega is NaN
        // there, so the sum is NaN whatever the rest of the grid holds
        const bool dead = (__dmul_rn(c.c2dt.x, probe) >= 2.001) || ((s_negcorner >> tid) & 1);
        if (!dead) atomicOr(&s_alive, 1);
        s_mb[tid] = mb;
    }
    __syncthreads();
    const bool force = s_force != 0;   // CTA-uniform

    // Per-thread state.  Ts and tzys hold the z-dependent parts for the thread's current z (scaled by
    // 2 dt^2); thr is the high-word threshold of the maximum filter (INT_MIN = "always ask").
    double Ts[C], tzys[C], dtc[C], S[C];
    int thr[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        S[c] = 0.; thr[c] = INT_MIN; dtc[c] = cand[c].dt;
        s_umax[c * blockDim.x + tid] = 0ull;   // own slots only: no barrier needed
        s_smax[c * blockDim.x + tid] = 0ll;
    }

    // The extrema-only path is compiled into the search (PREC 1) and screening (SCAN) variants only: the call costs the
    // plain throughput kernel 1.3 % (measured on the same GPU: 729.6 vs 738.9 G evals/s on 512 x 512^3) even when it
    // is never taken, and large batches that may hold hopeless candidates arrive as screening launches anyway.
    if ((SCAN || PREC == 1) && s_alive == 0) {
        // every candidate of this chunk is beyond the CFL limit: extrema only, the sums are NaN (CTA-uniform branch)
        int bz = blockIdx.z;
        extrema_only_sweep<C, UNIT>(g, a, cand, s_mb, s_corner, s_rows + (size_t)warp * ROWS * ROW, s_umax, s_smax, ztile,
                                    is, ie, ys, min(bz * a.tile_y + a.tile_y, g.ny));
        unsigned long long um[C];
        long long sm[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            S[c] = __longlong_as_double(0x7ff8000000000000ll);
            um[c] = s_umax[c * blockDim.x + tid];
            sm[c] = s_smax[c * blockDim.x + tid];
        }
        const int seg_d = chunk / a.chunks_per_seg;
        finish_candidates<C>(a, cand, S, um, sm, s_red, s_redu, chunk, ztile,
                             (long long)seg_d * a.seg_size + (long long)(chunk - seg_d * a.chunks_per_seg) * C,
                             (long long)(seg_d + 1) * a.seg_size);
        return;
    }

    // 0.375 (the cubic step of |k|'s square root) as an opaque register value: left to itself the compiler sometimes
    // re-loads it from the constant bank in every iteration of the hot loop (2.5 % of the throughput)
    // (built from a launch argument so that it cannot be re-materialised from the constant bank: a.P < 2^30 always)
    const double c375 = __hiloint2double(0x3fd80000 + (a.P >> 30), 0);
    double* wrows = s_rows + (size_t)warp * ROWS * ROW;   // this warp's private table
    for (int i = ie - 1; i >= is; --i) {
        const int x = a.x_begin + i * a.x_stride;
        // Load balance: the cost of a k-point depends on its class, i.e. on |k|, i.e. on z.  Warp w
        // therefore takes z slice (w + i) mod nwarps of the CTA's z tile for plane i, so that over an
        // x range every warp sees every z slice and all warps of the CTA finish together.
        const int zslice = (warp + i) % nwarps;
        const int zraw = ztile * blockDim.x + zslice * 32 + lane;
        // Lanes past the end of the z axis repeat its last point: all 32 lanes stay converged (votes with
        // a constant full mask), the extrema are unaffected and their sums are simply not accumulated.
        const bool active = zraw < g.nz;
        if (__ballot_sync(FULL, active) == 0u) continue;                  // warp-uniform
        const int z = min(zraw, g.nz - 1);
        const double cz = g.cz[z];
        const double sz2r = __dmul_rn(g.sz2[z], rZ);
        const double kz2 = g.kz2[z];
        const double opz = __dadd_rn(1.0, __dmul_rn(2.0, cz));
        {
            const double cx = g.cx[x], sx2 = g.sx2[x];   // needed here only: the row builder re-reads them
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double c2 = cand[c].c2dt.x;
                const double tzx = __dmul_rn(cand[c].bxz2, cz);
                const double tzzx = fma(cand[c].bzx2, cx, fma(cand[c].dlz, opz, cand[c].az));
                tzys[c] = __dmul_rn(c2, __dmul_rn(cand[c].byz2, cz));
                Ts[c] = __dmul_rn(c2, fma(tzx, sx2, __dmul_rn(tzzx, sz2r)));
            }
        }
        int bz = blockIdx.z;                               // (opaque copy: the end of the y range is re-derived per
        asm volatile("" : "+r"(bz));                       //  plane instead of being held, or spilled, across the sweep)
        const int ye = min(bz * a.tile_y + a.tile_y, g.ny);
        for (int y1 = ye; y1 > ys; y1 -= ROWS) {          // y tiles from the top of the range downwards
            const int y0 = max(ys, y1 - ROWS);
            const int nrows = y1 - y0;
            __syncwarp();                      // every lane is done reading the previous tile
            {                                  // lane l builds row l
                // the x-dependent table entries are re-read here (L1 hits) through an opaque index rather
                // than kept in 8 registers across the sweep of the previous tile
                int xq = x;
                asm volatile("" : "+r"(xq));
                const double cx = g.cx[xq], sx2 = g.sx2[xq], kx2 = g.kx2[xq];
                const double opx = __dadd_rn(1.0, __dmul_rn(2.0, cx));
                const int y = min(y0 + lane, y1 - 1);   // rows past the tile's end repeat its last row
                const double cy = g.cy[y];
                const double opy = __dadd_rn(1.0, __dmul_rn(2.0, cy));
                const double sy2r = __dmul_rn(g.sy2[y], rY);
                double* row = wrows + lane * ROW;
                *reinterpret_cast<double2*>(row) = make_double2(sy2r, __dadd_rn(kx2, g.ky2[y]));
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const double c2 = cand[c].c2dt.x;
                    const double axy = fma(cand[c].bxy2, cy, fma(cand[c].dlx, opx, cand[c].ax));
                    const double ayx = fma(cand[c].byx2, cx, fma(cand[c].dly, opy, cand[c].ay));
                    double2 v;
                    v.x = __dmul_rn(c2, fma(axy, sx2, __dmul_rn(ayx, sy2r)));
                    v.y = __dmul_rn(c2, __dmul_rn(cand[c].bzy2, cy));
                    *reinterpret_cast<double2*>(row + 2 + 2 * c) = v;
                }
            }
            __syncwarp();
            const double* wrow = nullptr;
            if (WK == 1) wrow = g.w + (size_t)y0 * g.nz + z;
            if (WK == 2) wrow = g.w + ((size_t)x * g.ny + y0) * g.nz + z;
            // U rows x C candidates = J independent evaluation chains per iteration (J = 4 for C <= 4): the
            // FP64 pipe has an 8-cycle dependent-issue latency at one instruction per 2 cycles.
            const int nit = (nrows + U - 1) / U;
            const double* row = wrows + (size_t)(nit - 1) * U * ROW;
#pragma unroll 1
            for (int yy = (nit - 1) * U; yy >= 0; yy -= U, row -= U * ROW) {
                // W = 2 dt^2 A (regrouped), v = W - 1 = 2 s^2 - 1, t = v^2, hk = pi/2 - k dt
                double W[J], A[J], v[J], t[J], e[J], hk[J], w[U], kr[U];
                bool need = false, all_mid = true;
#if SB200_INNER
                bool all_inner = false;
#endif
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double* r = row + u * ROW;
                    const double2 yk = *reinterpret_cast<const double2*>(r);   // (sy2r, kx2+ky2)
                    w[u] = 1.0;
                    if (WK != 0) w[u] = __ldg(wrow + (size_t)min(yy + u, nrows - 1) * g.nz);
                    // |k| (dispersion.py:350) to 1 ulp: its rounding enters (omega - k) like omega's own
                    // (2.8e-15) and far below it; the correctly rounded version costs 4 more instructions per row
                    const double k = sqrt_approx(__dadd_rn(yk.y, kz2), c375);
                    kr[u] = k;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int j = u * C + c;
                        const double2 q = *reinterpret_cast<const double2*>(r + 2 + 2 * c);   // 2dt^2 * (R, bzy)
                        W[j] = fma(tzys[c], yk.x, fma(q.y, sz2r, __dadd_rn(q.x, Ts[c])));
                        need = need || (__double2hiint(W[j]) >= thr[c]);
                        v[j] = __dadd_rn(W[j], -1.0);
                        hk[j] = fma(-k, dtc[c], kPi2);
                        t[j] = __dmul_rn(v[j], v[j]);
#if !SB200_ALLMID_VIMAX
                        all_mid = all_mid && ((unsigned int)__double2hiint(t[j]) <= 0x3fd00000u);   // |v| <= 1/2, not NaN
#endif
                    }
                }
#if SB200_ALLMID_VIMAX
                {   // |v| <= 1/2 on every chain <=> the largest high word of v^2 (as unsigned; a NaN is larger still) <= 1/4:
                    // three-input integer maxima instead of one compare per chain
                    unsigned int hmax = (unsigned int)__double2hiint(t[0]);
#pragma unroll
                    for (int j = 1; j + 1 < J; j += 2)
                        hmax = __vimax3_u32(hmax, (unsigned int)__double2hiint(t[j]), (unsigned int)__double2hiint(t[j + 1]));
                    if (J % 2 == 0) hmax = max(hmax, (unsigned int)__double2hiint(t[J - 1]));
                    all_mid = hmax <= 0x3fd00000u;
#if SB200_INNER
                    all_inner = hmax <= SB200_INNER_HI;
#endif
                }
#endif
                // Rare: redo the iteration in the reference's operation order and let the exact values (and
                // only those) into the extrema.  The opaque copies keep the address arithmetic in here.
                auto redo_exact = [&]() {
                    int zq = z, yq = yy, xq = x;
                    asm volatile("" : "+r"(zq), "+r"(yq), "+r"(xq));
                    const double czr = ld_rare(g.cz + zq), sz2 = ld_rare(g.sz2 + zq);
                    const double cx = ld_rare(g.cx + xq), sx2 = ld_rare(g.sx2 + xq);
                    all_mid = true;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int y = y0 + min(yq + u, nrows - 1);
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
                            const double c2 = cand[c].c2dt.x;
                            A[j] = Ae;
                            W[j] = __dmul_rn(c2, Ae);
                            v[j] = fma(c2, Ae, -1.0);
                            t[j] = __dmul_rn(v[j], v[j]);
                            all_mid = all_mid && ((unsigned int)__double2hiint(t[j]) <= 0x3fd00000u);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const double wm = __dmul_rn(cand[c].c2dt.x, __longlong_as_double(s_smax[c * blockDim.x + tid]));
                        thr[c] = force ? INT_MIN : max(__double2hiint(wm), 2) - 2;
                    }
                };
                const bool redone = __any_sync(FULL, need);
                if (redone) redo_exact();
                if (__all_sync(FULL, all_mid)) {
                    // |v| <= 1/2 for every lane and chain: dt*omega = pi/2 + v f(v^2), no square root
#if SB200_INNER
                    // ... and with the shorter series when v^2 is small on every lane and chain (after a redo the
                    // flag is stale, i.e. false: the full series is always valid)
                    if (__all_sync(FULL, all_inner && !redone)) {
#pragma unroll
                        for (int j = 0; j < J; ++j) e[j] = fma(v[j], asin_f_inner<PREC>(t[j]), hk[j]);
                    } else
#endif
                    {
#pragma unroll
                        for (int j = 0; j < J; ++j) e[j] = fma(v[j], asin_f<PREC>(t[j]), hk[j]);
                    }
                } else {
                    if (SCAN) {
                        // v >= 1 + 2^-12 as a signed compare of high words.  +inf and a positive NaN pass as well: omega
                        // is NaN for those too, and a positive NaN was already caught by the maximum filter above (its
                        // high word exceeds every threshold), i.e. the iteration was redone and the extrema are poisoned;
                        // a sign-bit NaN fails the test and goes on to the guards below.
                        bool beyond = true;
#pragma unroll
                        for (int j = 0; j < J; ++j) beyond = beyond && (__double2hiint(v[j]) >= 0x3ff00100);
                        if (__all_sync(FULL, beyond)) {
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                if (active) S[c] = __longlong_as_double(0x7ff8000000000000ll);   // NaN is sticky in the sum
                            continue;
                        }
                    }
                    double z2[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) z2[j] = __dmul_rn(W[j], 0.5);             // dt^2 A
                    // |v| <= sqrt(3)/2, or -1 + 2^-21 < v < -1/2, on every lane and chain: out of reach of every guard
                    if (!dev_try_nosqrt<C, J, PREC>(v, hk, e, FULL) && !dev_try_low<C, J, true, PREC>(z2, v, kr, e, cand, FULL, c375)) {
                        if (!redone) {
                            // v <= -1 + 2^-21 (sqrtarg may be negative or a negative NaN) or v within
                            // [1 - 2^-9 - 2^-12, 1 + 2^-12): the EXACT class needs the reference's bits there.
                            // Beyond 1 + 2^-12 it does not: s > 1 and omega is NaN whatever the last bits are.
                            need = false;
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                const unsigned int hv = (unsigned int)__double2hiint(v[j]);
                                need = need || (hv >= 0xbfefffffu) || (hv - 0x3fefee00u < 0x1300u);
                            }
                            if (__any_sync(FULL, need)) {
                                redo_exact();
#pragma unroll
                                for (int j = 0; j < J; ++j) z2[j] = __dmul_rn(W[j], 0.5);
                            } else {
#pragma unroll
                                for (int j = 0; j < J; ++j) A[j] = __dmul_rn(W[j], cand[j % C].ic2);   // read where s > 1 only
                            }
                        }
                        dev_high_or_any<C, J, PREC>(A, z2, v, hk, kr, e, cand, FULL, c375);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (active && (U == 1 || yy + u < nrows)) {   // rows past the end are repeats: skip their sum only
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

    unsigned long long um[C];
    long long sm[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        um[c] = s_umax[c * blockDim.x + tid];
        sm[c] = s_smax[c * blockDim.x + tid];
    }
    int bx = blockIdx.x;                                   // re-derived from an opaque copy: chunk / ztile / b0 need not stay alive
    asm volatile("" : "+r"(bx));
    const int chunk_e = bx / a.n_ztiles;
    const int seg_e = chunk_e / a.chunks_per_seg;
    const long long b0_e = (long long)seg_e * a.seg_size + (long long)(chunk_e - seg_e * a.chunks_per_seg) * C;
    finish_candidates<C>(a, cand, S, um, sm, s_red, s_redu, chunk_e, bx - chunk_e * a.n_ztiles, b0_e,
                         (long long)(seg_e + 1) * a.seg_size);
}

}  // namespace sb200
