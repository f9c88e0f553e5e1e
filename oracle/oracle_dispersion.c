/*
 * oracle_dispersion.c -- TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference objective.
 *
 * A second, NumPy-independent statement of what /root/reference/extended_stencil/dispersion.py computes
 * (sqrtarg :309-320 / :387-399, sqrtres :111-121, omega :186-191, the weighted sum :251-252, min/max
 * :292 / :367 / :149), in the reference's operation order, with libm sqrt/asin.  It exists so that
 *   - parity can be checked at BASELINE.json's full sizes (512^3 per candidate in well under a second on
 *     a few cores; 2048^3 slab by slab) where the NumPy oracle needs gigabytes of temporaries, and
 *   - the sum has an accuracy reference: it is accumulated with Neumaier compensation, i.e. to ~1 ulp,
 *     whereas NumPy's pairwise sum and the GPU's tree are both only good to ~1e-15.
 * PARITY PINNED: tests/test_c_oracle.py checks it against tests/golden/objective_cases.npz (produced by
 * the unmodified reference) -- extrema bitwise, omega and the sum to the stated tolerances.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC; no -march flags: the
 * operation order of sqrtarg must not be contracted into FMAs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    int dim;
    int n[3];            /* n[2] = 1 for dim 2 */
    const double* cosk[3];
    const double* s2[3];
    const double* kcomp[3];
    double Y, Z;
    double Y2, Z2;       /* (Y*dx)**2, (Z*dx)**2 as the HOST LANGUAGE evaluated them (dispersion.py:319, 398):
                            Python's / NumPy's x**2 is libm pow, not always the correctly rounded x*x */
    int wkind;           /* 0: w == 1, 1: w[ny*nz] independent of x (dim 3), 2: dense C-order */
    const double* w;
} oracle_grid;

static inline double sqrtarg3(const double* c, double cx, double cy, double cz, double sx2, double sy2, double sz2,
                              double Y2, double Z2) {
    /* c: dt, ax, ay, az, bxy, bxz, byx, byz, bzx, bzy, dlx, dly, dlz   (stencil.py:275) */
    const double dx = 1.0;
    double Ax = c[1] + c[10] * (1.0 + 2.0 * cx) + 2. * c[4] * cy + 2. * c[5] * cz;   /* dispersion.py:394 */
    double Ay = c[2] + c[11] * (1.0 + 2.0 * cy) + 2. * c[6] * cx + 2. * c[7] * cz;   /* :395 */
    double Az = c[3] + c[12] * (1.0 + 2.0 * cz) + 2. * c[8] * cx + 2. * c[9] * cy;   /* :396 */
    return Ax * sx2 / (dx * dx) + Ay * sy2 / Y2 + Az * sz2 / Z2;                     /* :398 */
}

static inline double sqrtarg2(const double* c, double cx, double cy, double sx2, double sy2, double Y2) {
    /* c: dt, ax, ay, bxy, byx, dlx, dly   (stencil.py:209) */
    const double dx = 1.0;
    double Ax = c[1] + c[5] * (1.0 + 2.0 * cx) + 2. * c[3] * cy;                     /* dispersion.py:316 */
    double Ay = c[2] + c[6] * (1.0 + 2.0 * cy) + 2. * c[4] * cx;                     /* :317 */
    return Ax * sx2 / (dx * dx) + Ay * sy2 / Y2;                                     /* :319 */
}

static inline void neumaier(double* sum, double* comp, double v) {
    double t = *sum + v;
    if (fabs(*sum) >= fabs(v)) *comp += (*sum - t) + v; else *comp += (v - t) + *sum;
    *sum = t;
}

/* Un-gated reductions over x-planes [x0, x1) for nbatch coefficient vectors (row-major, 7 or 13 doubles):
 * S = sum w (omega - k)^2, Amin = min sqrtarg, Amax = max sqrtarg (true min / max, NaN if any is NaN),
 * nan = number of NaN omegas.  Returns 0. */
int oracle_objective(const oracle_grid* g, const double* coeffs, int64_t nbatch, int x0, int x1,
                     double* S, double* Amin, double* Amax, int64_t* nan) {
    const int dim = g->dim, ncoef = dim == 2 ? 7 : 13;
    const int ny = g->n[1], nz = dim == 3 ? g->n[2] : 1;
    for (int64_t b = 0; b < nbatch; ++b) {
        const double* c = coeffs + b * ncoef;
        const double dt = c[0], dx = 1.0;
        double tot = 0., totc = 0., amin = INFINITY, amax = -INFINITY;
        int64_t nn = 0;
        int anynan = 0;
#pragma omp parallel for schedule(static) reduction(+ : nn) reduction(| : anynan)
        for (int x = x0; x < x1; ++x) {
            double s = 0., sc = 0., lo = INFINITY, hi = -INFINITY;
            for (int y = 0; y < ny; ++y)
                for (int z = 0; z < nz; ++z) {
                    double A, k;
                    if (dim == 3) {
                        A = sqrtarg3(c, g->cosk[0][x], g->cosk[1][y], g->cosk[2][z], g->s2[0][x], g->s2[1][y],
                                     g->s2[2][z], g->Y2, g->Z2);
                        k = sqrt(g->kcomp[0][x] * g->kcomp[0][x] + g->kcomp[1][y] * g->kcomp[1][y] +
                                 g->kcomp[2][z] * g->kcomp[2][z]);                    /* dispersion.py:350 */
                    } else {
                        A = sqrtarg2(c, g->cosk[0][x], g->cosk[1][y], g->s2[0][x], g->s2[1][y], g->Y2);
                        k = sqrt(g->kcomp[0][x] * g->kcomp[0][x] + g->kcomp[1][y] * g->kcomp[1][y]);  /* :276 */
                    }
                    if (A != A) anynan = 1;
                    if (A < lo) lo = A;
                    if (A > hi) hi = A;
                    const double om = (2. / (dt * dx)) * asin(dt * dx * sqrt(A));     /* :116, :191 */
                    if (om != om) nn++;
                    double w = 1.0;
                    if (g->wkind == 1) w = g->w[(size_t)y * nz + z];
                    else if (g->wkind == 2) w = g->w[((size_t)x * ny + y) * nz + z];
                    neumaier(&s, &sc, w * ((om - k) * (om - k)));                     /* :251-252 */
                }
#pragma omp critical
            {
                neumaier(&tot, &totc, s);
                neumaier(&tot, &totc, sc);
                if (lo < amin) amin = lo;
                if (hi > amax) amax = hi;
            }
        }
        S[b] = tot + totc;
        Amin[b] = anynan ? NAN : amin;
        Amax[b] = anynan ? NAN : amax;
        nan[b] = nn;
    }
    return 0;
}

/* Dense field over x-planes [x0, x1): what = 0 omega, 1 sqrtarg, 2 sqrtres (C order). */
int oracle_export(const oracle_grid* g, const double* c, int what, int x0, int x1, double* out) {
    const int dim = g->dim;
    const int ny = g->n[1], nz = dim == 3 ? g->n[2] : 1;
    const double dt = c[0], dx = 1.0;
#pragma omp parallel for schedule(static)
    for (int x = x0; x < x1; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                double A = dim == 3 ? sqrtarg3(c, g->cosk[0][x], g->cosk[1][y], g->cosk[2][z], g->s2[0][x],
                                               g->s2[1][y], g->s2[2][z], g->Y2, g->Z2)
                                    : sqrtarg2(c, g->cosk[0][x], g->cosk[1][y], g->s2[0][x], g->s2[1][y], g->Y2);
                double v = A;
                if (what != 1) {
                    v = sqrt(A);
                    if (what == 0) v = (2. / (dt * dx)) * asin(dt * dx * v);
                }
                out[((size_t)(x - x0) * ny + y) * nz + z] = v;
            }
    return 0;
}

int oracle_abi_version(void) { return 1; }
