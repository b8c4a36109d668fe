/*
 * synth.c -- deterministic synthetic black-oil Jacobians (benchmark/test INPUT generator).
 *
 * Produces the 3x3-block BSR systems of BASELINE.json's configs: a structured
 * nx*ny*nz grid in natural ordering cell = i + nx*(j + ny*k), 7-point TPFA pattern plus optional
 * fault-like non-neighbour connections, log-normal permeability from box-smoothed Gaussian noise,
 * harmonic-mean transmissibilities, non-symmetric 3x3 blocks whose first column is scaled 1e-7
 * like the reference fixture tests/matr33.txt:4-66.  Every random number is a pure hash of
 * (seed, stream, index), so any k-slab [k0,k1) can be generated independently (one slab per GPU
 * rank) and bit-identically to the same rows of the full system.
 *
 * Host-side harness code: not on the solve path and not part of the oracle.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int nx, ny, nz;
    double sigma;        /* std of log-permeability */
    double kvkh;         /* vertical/horizontal transmissibility ratio */
    double acc_frac;     /* accumulation term relative to the mean face transmissibility */
    double offdiag_rand; /* 0.25: blocks are T * (I + offdiag_rand * U(-1,1)) */
    uint64_t seed;
    int nfaults;         /* fault planes: cells (fi-1,j,k) <-> (fi,j,k+throw) */
    int fault_i[4];
    int fault_throw[4];
    double fault_mult;   /* transmissibility multiplier of the NNC */
} synth_cfg;

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t seed, uint64_t stream, uint64_t idx)
{
    uint64_t h = mix64(mix64(seed ^ (stream * 0xD1B54A32D192ED03ULL)) + idx);
    return ((double) (h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double upm1(uint64_t seed, uint64_t stream, uint64_t idx) { return 2.0 * u01(seed, stream, idx) - 1.0; }
static inline double gauss(uint64_t seed, uint64_t idx)
{
    double a = u01(seed, 1, idx), b = u01(seed, 2, idx);
    return sqrt(-2.0 * log(a)) * cos(6.283185307179586 * b);
}

static const double CS[3] = {1e-7, 1.0, 1.0};

/* log-permeability field on k in [ka, kb): box filter radius 2 (clipped), unit variance per cell */
static double *perm_field(const synth_cfg *c, int ka, int kb)
{
    const int R = 2;
    int nx = c->nx, ny = c->ny, nz = c->nz;
    int wa = ka - R < 0 ? 0 : ka - R, wb = kb + R > nz ? nz : kb + R;
    size_t plane = (size_t) nx * ny;
    size_t nw = plane * (size_t) (wb - wa);
    double *w = (double *) malloc(nw * sizeof(double));
    double *t = (double *) malloc(nw * sizeof(double));
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < nw; ++q) w[q] = gauss(c->seed, (uint64_t) wa * plane + q);
    /* x pass: w -> t */
#pragma omp parallel for schedule(static)
    for (size_t row = 0; row < (size_t) ny * (wb - wa); ++row) {
        const double *src = w + row * nx; double *dst = t + row * nx;
        for (int i = 0; i < nx; ++i) {
            double s = 0.0; int lo = i - R < 0 ? 0 : i - R, hi = i + R >= nx ? nx - 1 : i + R;
            for (int q = lo; q <= hi; ++q) s += src[q];
            dst[i] = s;
        }
    }
    /* y pass: t -> w */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < wb - wa; ++k)
        for (int j = 0; j < ny; ++j) {
            int lo = j - R < 0 ? 0 : j - R, hi = j + R >= ny ? ny - 1 : j + R;
            for (int i = 0; i < nx; ++i) {
                double s = 0.0;
                for (int q = lo; q <= hi; ++q) s += t[(size_t) k * plane + (size_t) q * nx + i];
                w[(size_t) k * plane + (size_t) j * nx + i] = s;
            }
        }
    /* z pass + normalisation + exp: w -> out (k in [ka,kb)) */
    double *out = (double *) malloc(plane * (size_t) (kb - ka) * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int k = ka; k < kb; ++k) {
        int lo = k - R < 0 ? 0 : k - R, hi = k + R >= nz ? nz - 1 : k + R;
        for (int j = 0; j < ny; ++j) {
            int cy = (j + R >= ny ? ny - 1 : j + R) - (j - R < 0 ? 0 : j - R) + 1;
            for (int i = 0; i < nx; ++i) {
                int cx = (i + R >= nx ? nx - 1 : i + R) - (i - R < 0 ? 0 : i - R) + 1;
                double s = 0.0;
                for (int q = lo; q <= hi; ++q) s += w[(size_t) (q - wa) * plane + (size_t) j * nx + i];
                double n = (double) cx * cy * (hi - lo + 1);
                out[(size_t) (k - ka) * plane + (size_t) j * nx + i] = exp(c->sigma * s / sqrt(n));
            }
        }
    }
    free(w); free(t);
    return out;
}

typedef struct { int64_t col; double T; } nb_t;

/* neighbours of cell (i,j,k), ascending by column, diagonal excluded; returns count */
static int neighbours(const synth_cfg *c, int i, int j, int k, const double *K, int ka, nb_t *nb)
{
    int nx = c->nx, ny = c->ny, nz = c->nz;
    size_t plane = (size_t) nx * ny;
    int64_t me = (int64_t) i + (int64_t) nx * (j + (int64_t) ny * k);
    double Kc = K ? K[(size_t) (k - ka) * plane + (size_t) j * nx + i] : 1.0;
    int n = 0;
#define ADD(ii, jj, kk, a)                                                                  \
    do {                                                                                    \
        nb[n].col = (int64_t) (ii) + (int64_t) nx * ((jj) + (int64_t) ny * (kk));           \
        if (K) { double Kd = K[(size_t) ((kk) - ka) * plane + (size_t) (jj) * nx + (ii)];  \
                 nb[n].T = 2.0 * Kc * Kd / (Kc + Kd) * (a); }                               \
        ++n;                                                                                \
    } while (0)
    if (k > 0) ADD(i, j, k - 1, c->kvkh);
    if (j > 0) ADD(i, j - 1, k, 1.0);
    if (i > 0) ADD(i - 1, j, k, 1.0);
    if (i < nx - 1) ADD(i + 1, j, k, 1.0);
    if (j < ny - 1) ADD(i, j + 1, k, 1.0);
    if (k < nz - 1) ADD(i, j, k + 1, c->kvkh);
    for (int f = 0; f < c->nfaults; ++f) {
        int fi = c->fault_i[f], s = c->fault_throw[f];
        if (fi <= 0 || fi >= nx || s <= 0) continue;
        if (i == fi - 1 && k + s < nz) ADD(fi, j, k + s, c->fault_mult);
        if (i == fi && k - s >= 0) ADD(fi - 1, j, k - s, c->fault_mult);
    }
#undef ADD
    for (int a = 1; a < n; ++a) {          /* insertion sort by column */
        nb_t t = nb[a]; int b = a - 1;
        while (b >= 0 && nb[b].col > t.col) { nb[b + 1] = nb[b]; --b; }
        nb[b + 1] = t;
    }
    (void) me;
    return n;
}

/* largest |k offset| any connection can have (halo depth in planes) */
int synth_halo_planes(const synth_cfg *c)
{
    int h = 1;
    for (int f = 0; f < c->nfaults; ++f) if (c->fault_throw[f] > h) h = c->fault_throw[f];
    return h;
}

/* Pass 1: row pointer (local, rowptr[0] = 0) for block rows of planes [k0,k1). Returns nnzb. */
int64_t synth_count(const synth_cfg *c, int k0, int k1, int64_t *rowptr)
{
    int nx = c->nx, ny = c->ny;
    size_t plane = (size_t) nx * ny;
    size_t nrows = plane * (size_t) (k1 - k0);
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nrows; ++r) {
        int i = (int) (r % nx), j = (int) ((r / nx) % ny), k = k0 + (int) (r / plane);
        nb_t nb[16];
        rowptr[r + 1] = 1 + neighbours(c, i, j, k, NULL, 0, nb);
    }
    rowptr[0] = 0;
    for (size_t r = 0; r < nrows; ++r) rowptr[r + 1] += rowptr[r];
    return rowptr[nrows];
}

static inline void block_T(double T, double rnd, uint64_t seed, uint64_t stream, uint64_t idx, double *blk, int accumulate)
{
    for (int q = 0; q < 9; ++q) {
        double id = (q == 0 || q == 4 || q == 8) ? 1.0 : 0.0;
        double v = T * (id + rnd * upm1(seed, stream, idx * 9 + q)) * CS[q % 3];
        if (accumulate) blk[q] += v; else blk[q] = v;
    }
}

/* x_true(cell, comp): U(-1,1) scaled by (1e5, 1, 1) */
static inline double xtrue(const synth_cfg *c, int64_t cell, int comp)
{
    static const double XS[3] = {1e5, 1.0, 1.0};
    return upm1(c->seed, 7, (uint64_t) cell * 3 + comp) * XS[comp];
}

void synth_xtrue(const synth_cfg *c, int64_t cell0, int64_t ncells, double *x)
{
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < ncells; ++q)
        for (int comp = 0; comp < 3; ++comp) x[q * 3 + comp] = xtrue(c, cell0 + q, comp);
}

/* Pass 2: global column ids (int64), values, and b = A x_true for the rows of planes [k0,k1). */
void synth_fill(const synth_cfg *c, int k0, int k1, const int64_t *rowptr, int64_t *cols, double *vals, double *b)
{
    int nx = c->nx, ny = c->ny, nz = c->nz;
    size_t plane = (size_t) nx * ny;
    size_t nrows = plane * (size_t) (k1 - k0);
    int h = synth_halo_planes(c);
    int ka = k0 - h < 0 ? 0 : k0 - h, kb = k1 + h > nz ? nz : k1 + h;
    double *K = perm_field(c, ka, kb);
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nrows; ++r) {
        int i = (int) (r % nx), j = (int) ((r / nx) % ny), k = k0 + (int) (r / plane);
        int64_t me = (int64_t) k0 * plane + (int64_t) r;
        nb_t nb[16];
        int n = neighbours(c, i, j, k, K, ka, nb);
        double diag[9], Tsum = 0.0;
        for (int q = 0; q < 9; ++q) diag[q] = 0.0;
        for (int a = 0; a < n; ++a) Tsum += nb[a].T;
        double acc = c->acc_frac * (n > 0 ? Tsum / n : 1.0);
        block_T(acc, c->offdiag_rand, c->seed, 3, (uint64_t) me, diag, 1);
        int64_t pos = rowptr[r];
        int placed_diag = 0;
        double bb[3] = {0, 0, 0};
        for (int a = 0; a <= n; ++a) {
            if (!placed_diag && (a == n || nb[a].col > me)) {
                /* diagonal = accumulation + sum_d T_cd M_cd with the SAME M_cd as the off-diagonal block of
                 * this row, so the flux part of every block row annihilates constants (weak block
                 * diagonal dominance, like a TPFA flux Jacobian) */
                for (int aa = 0; aa < n; ++aa)
                    block_T(nb[aa].T, c->offdiag_rand, c->seed, 5, (uint64_t) me * 0x9E3779B97F4A7C15ULL + (uint64_t) nb[aa].col, diag, 1);
                cols[pos] = me;
                memcpy(vals + pos * 9, diag, sizeof diag);
                for (int rr = 0; rr < 3; ++rr)
                    for (int cc = 0; cc < 3; ++cc) bb[rr] += diag[rr * 3 + cc] * xtrue(c, me, cc);
                ++pos; placed_diag = 1;
            }
            if (a == n) break;
            double blk[9];
            /* pair hash is ordered: (me,col) and (col,me) draw independently (upwind-like asymmetry) */
            block_T(-nb[a].T, c->offdiag_rand, c->seed, 5, (uint64_t) me * 0x9E3779B97F4A7C15ULL + (uint64_t) nb[a].col, blk, 0);
            cols[pos] = nb[a].col;
            memcpy(vals + pos * 9, blk, sizeof blk);
            for (int rr = 0; rr < 3; ++rr)
                for (int cc = 0; cc < 3; ++cc) bb[rr] += blk[rr * 3 + cc] * xtrue(c, nb[a].col, cc);
            ++pos;
        }
        if (b) { b[r * 3] = bb[0]; b[r * 3 + 1] = bb[1]; b[r * 3 + 2] = bb[2]; }
    }
    free(K);
}

/* uniform helper exported for the well generator in synth.py */
double synth_u01(uint64_t seed, uint64_t stream, uint64_t idx) { return u01(seed, stream, idx); }
void synth_u01_array(uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, double *out)
{
    for (int64_t q = 0; q < n; ++q) out[q] = u01(seed, stream, idx0 + (uint64_t) q);
}
