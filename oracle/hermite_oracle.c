/*
 * hermite_oracle.c — CPU restatement of MrMustard's Gaussian-to-Fock recurrence (TEST INFRASTRUCTURE).
 *
 * This file is the parity ORACLE of the repository.  It is test infrastructure, not product code:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it.  The product path (mrmustard_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against golden
 * vectors produced by the unmodified reference (numba strategies imported from /root/reference by
 * tests/golden/gen_golden.py) and against the reference's own known-answer tests
 * (tests/test_math/test_special.py:24-36 Hermite polynomials).
 *
 * Each function restates, in plain C with strict IEEE-754 double arithmetic (compile with
 * -ffp-contract=off; no FMA, no reassociation), the algorithm of the reference function cited in
 * its header comment.  Complex arithmetic mirrors what numba lowers complex128 operators to
 * (numba/cpython/numbers.py: complex_mul_impl, complex_div_impl), including the promotion of a
 * float64 operand to complex128 (re, +0.0) before a mixed complex*float / complex/float operation.
 *
 * Layout: all tensors are C-contiguous complex128 stored as interleaved (re, im) doubles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct { double re, im; } cplx;

/* numba complex_mul_impl: real = a*c - b*d ; imag = a*d + b*c  (x = a+ib, y = c+id) */
static inline cplx cmul(cplx x, cplx y) {
    double ac = x.re * y.re, bd = x.im * y.im, ad = x.re * y.im, bc = x.im * y.re;
    cplx r = { ac - bd, ad + bc };
    return r;
}
static inline cplx cadd(cplx x, cplx y) { cplx r = { x.re + y.re, x.im + y.im }; return r; }
static inline cplx cfromreal(double s) { cplx r = { s, 0.0 }; return r; }
/* complex * float64: numba promotes the float to complex128 first */
static inline cplx cmulr(cplx x, double s) { return cmul(x, cfromreal(s)); }
static inline cplx rmulc(double s, cplx x) { return cmul(cfromreal(s), x); }

/* numba complex_div_impl (Smith's algorithm as in CPython's c_quot) */
static inline cplx cdiv(cplx a, cplx b) {
    cplx r;
    if (b.re == 0.0 && b.im == 0.0) { r.re = NAN; r.im = NAN; return r; } /* reference raises ZeroDivisionError */
    if (fabs(b.re) >= fabs(b.im)) {
        if (b.re == 0.0) { r.re = NAN; r.im = NAN; return r; }
        double ratio = b.im / b.re;
        double denom = b.re + b.im * ratio;
        r.re = (a.re + a.im * ratio) / denom;
        r.im = (a.im - a.re * ratio) / denom;
    } else {
        double ratio = b.re / b.im;
        double denom = b.re * ratio + b.im;
        r.re = (a.re * ratio + a.im) / denom;
        r.im = (a.im * ratio - a.re) / denom;
    }
    return r;
}
static inline cplx cdivr(cplx a, double s) { return cdiv(a, cfromreal(s)); }

/* SQRT = np.sqrt(np.arange(100000)) (vanilla/core.py:22) — IEEE sqrt of the integer, computed on demand */
static inline double SQRT_(int64_t n) { return sqrt((double)n); }

static void make_strides(int D, const int64_t *shape, int64_t *strides) {
    /* vanilla/core.py:68-70 */
    for (int i = 0; i < D; i++) strides[i] = 1;
    for (int i = D - 1; i > 0; i--) strides[i - 1] = strides[i] * shape[i];
}

static int next_index(int D, const int64_t *shape, int64_t *idx) {
    /* np.ndindex successor (row-major odometer) */
    for (int d = D - 1; d >= 0; d--) {
        if (++idx[d] < shape[d]) return 1;
        idx[d] = 0;
    }
    return 0;
}

/* read with numpy/numba negative-index wraparound on the FLAT array (core.py:101-103 reads
 * G[pivot - strides[j]] which may be negative when k_j == 0; the value is then multiplied by SQRT[0]) */
static inline cplx gread(const cplx *G, int64_t N, int64_t p) {
    if (p < 0) p += N;
    if (p < 0 || p >= N) { cplx z = { 0.0, 0.0 }; return z; } /* numba would read out of bounds; unreachable for valid shapes */
    return G[p];
}

/* ---------------------------------------------------------------------------------------------
 * vanilla_numba  — mrmustard/math/lattice/strategies/vanilla/core.py:25-124
 * G must hold N = prod(shape) entries. If zero_init != 0 the buffer is zeroed first (out=None case,
 * core.py:73); otherwise it is used as the caller's `out` (written in place).
 * ------------------------------------------------------------------------------------------- */
void mmo_vanilla(int D, const int64_t *shape, const double *A_, const double *b_, const double *c_,
                 double *G_, int zero_init) {
    const cplx *A = (const cplx *)A_, *b = (const cplx *)b_;
    cplx *G = (cplx *)G_;
    int64_t *strides = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
    int64_t *idx = (int64_t *)calloc(D + 1, sizeof(int64_t));
    make_strides(D, shape, strides);
    int64_t N = 1;
    for (int i = 0; i < D; i++) N *= shape[i];
    if (zero_init) memset(G, 0, sizeof(cplx) * N);
    if (N == 0) { free(strides); free(idx); return; }
    G[0].re = c_[0]; G[0].im = c_[1];
    for (int64_t flat = 1; flat < N; flat++) {
        next_index(D, shape, idx);
        int i = 0;
        int64_t pivot = 0;
        if (flat < strides[0]) {               /* core.py:85-94: first stride with pivot >= 0 */
            for (i = 0; i < D; i++) { pivot = flat - strides[i]; if (pivot >= 0) break; }
        } else {                               /* core.py:108-111 */
            i = 0; pivot = flat - strides[0];
        }
        cplx v = cmul(b[i], G[pivot]);                                                   /* :97  */
        v = cadd(v, cmul(cmulr(A[i * D + i], SQRT_(idx[i] - 1)), gread(G, N, pivot - strides[i]))); /* :101 */
        for (int j = i + 1; j < D; j++)
            v = cadd(v, cmul(cmulr(A[i * D + j], SQRT_(idx[j])), gread(G, N, pivot - strides[j]))); /* :103 */
        G[flat] = cdivr(v, SQRT_(idx[i]));                                               /* :104 */
    }
    free(strides); free(idx);
}

/* ---------------------------------------------------------------------------------------------
 * stable_numba — vanilla/core.py:127-213 (average over all valid pivots)
 * ------------------------------------------------------------------------------------------- */
void mmo_stable(int D, const int64_t *shape, const double *A_, const double *b_, const double *c_,
                double *G_, int zero_init) {
    const cplx *A = (const cplx *)A_, *b = (const cplx *)b_;
    cplx *G = (cplx *)G_;
    int64_t *strides = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
    int64_t *idx = (int64_t *)calloc(D + 1, sizeof(int64_t));
    make_strides(D, shape, strides);
    int64_t N = 1;
    for (int i = 0; i < D; i++) N *= shape[i];
    if (zero_init) memset(G, 0, sizeof(cplx) * N);
    if (N == 0) { free(strides); free(idx); return; }
    G[0].re = c_[0]; G[0].im = c_[1];
    for (int64_t flat = 1; flat < N; flat++) {
        next_index(D, shape, idx);
        int num_pivots = 0;
        cplx vals = { 0.0, 0.0 };                       /* `vals = 0` (:185) */
        for (int i = 0; i < D; i++) {
            if (idx[i] == 0) continue;                  /* :187 */
            num_pivots++;
            int64_t pivot = flat - strides[i];
            cplx v = cmul(b[i], G[pivot]);              /* :193 */
            for (int j = 0; j < i; j++)                 /* :198 */
                v = cadd(v, cmul(cmulr(A[i * D + j], SQRT_(idx[j])), gread(G, N, pivot - strides[j])));
            v = cadd(v, cmul(cmulr(A[i * D + i], SQRT_(idx[i] - 1)), gread(G, N, pivot - strides[i]))); /* :201 */
            for (int j = i + 1; j < D; j++)             /* :203 */
                v = cadd(v, cmul(cmulr(A[i * D + j], SQRT_(idx[j])), gread(G, N, pivot - strides[j])));
            vals = cadd(vals, cdivr(v, SQRT_(idx[i]))); /* :207 */
        }
        G[flat] = cdivr(vals, (double)num_pivots);      /* :210 */
    }
    free(strides); free(idx);
}

/* ---------------------------------------------------------------------------------------------
 * vanilla_batch_numba — vanilla/batch.py:27-61 (prange over the batch; here pthreads)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int D; const int64_t *shape; const double *A, *b, *c; double *G; int stable;
    int64_t B, N; int tid, nthreads;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    int D = j->D;
    for (int64_t k = j->tid; k < j->B; k += j->nthreads) {
        const double *A = j->A + 2 * (int64_t)D * D * k, *b = j->b + 2 * (int64_t)D * k, *c = j->c + 2 * k;
        double *G = j->G + 2 * j->N * k;
        /* G[k] = vanilla_numba(shape, A[k], b[k], c[k]) — a fresh zero tensor assigned into G[k] */
        if (j->stable) mmo_stable(D, j->shape, A, b, c, G, 1);
        else mmo_vanilla(D, j->shape, A, b, c, G, 1);
    }
    return NULL;
}

void mmo_vanilla_batch(int64_t B, int D, const int64_t *shape, const double *A, const double *b,
                       const double *c, int stable, double *G, int nthreads) {
    int64_t N = 1;
    for (int i = 0; i < D; i++) N *= shape[i];
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    batch_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        batch_job jb = { D, shape, A, b, c, G, stable, B, N, t, nthreads };
        jobs[t] = jb;
        if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    batch_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------------
 * vanilla_vjp_numba — vanilla/gradients.py:25-82
 * outputs: dLdA[D*D], dLdb[D], dLdc[1] (complex128 interleaved)
 * ------------------------------------------------------------------------------------------- */
void mmo_vanilla_vjp(int D, const int64_t *shape, const double *G_, const double *c_,
                     const double *dLdG_, double *dLdA_, double *dLdb_, double *dLdc_) {
    const cplx *G = (const cplx *)G_, *g = (const cplx *)dLdG_;
    cplx *outA = (cplx *)dLdA_, *outb = (cplx *)dLdb_;
    int64_t *strides = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
    int64_t *idx = (int64_t *)calloc(D + 1, sizeof(int64_t));
    make_strides(D, shape, strides);
    int64_t N = 1;
    for (int i = 0; i < D; i++) N *= shape[i];
    cplx *dA = (cplx *)calloc((size_t)D * D + 1, sizeof(cplx));
    cplx *db = (cplx *)calloc((size_t)D + 1, sizeof(cplx));
    cplx *accA = (cplx *)calloc((size_t)D * D + 1, sizeof(cplx));
    cplx *accb = (cplx *)calloc((size_t)D + 1, sizeof(cplx));
    for (int64_t flat = 1; flat < N; flat++) {               /* :64 */
        next_index(D, shape, idx);
        for (int i = 0; i < D; i++) {
            int64_t pivot = flat - strides[i];               /* :67 */
            db[i] = rmulc(SQRT_(idx[i]), gread(G, N, pivot)); /* :68 */
            if (idx[i] > 1) {                                /* :69-73 */
                double f = 0.5 * SQRT_(idx[i]) * SQRT_(idx[i] - 1);
                dA[i * D + i] = rmulc(f, gread(G, N, pivot - strides[i]));
            } else { dA[i * D + i].re = 0.0; dA[i * D + i].im = 0.0; }
            for (int j = i + 1; j < D; j++) {                /* :74-75 */
                double f = SQRT_(idx[i]) * SQRT_(idx[j]);
                dA[i * D + j] = rmulc(f, gread(G, N, pivot - strides[j]));
            }
        }
        for (int q = 0; q < D * D; q++) accA[q] = cadd(accA[q], cmul(dA[q], g[flat])); /* :77 */
        for (int q = 0; q < D; q++) accb[q] = cadd(accb[q], cmul(db[q], g[flat]));     /* :78 */
    }
    cplx s = { 0.0, 0.0 };                                   /* :80 np.sum(G * dLdG) / c */
    for (int64_t f = 0; f < N; f++) s = cadd(s, cmul(G[f], g[f]));
    cplx c = { c_[0], c_[1] };
    cplx dc = cdiv(s, c);
    dLdc_[0] = dc.re; dLdc_[1] = dc.im;
    for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++)                          /* :82 (dLdA + dLdA.T) / 2 */
            outA[i * D + j] = cdiv(cadd(accA[i * D + j], accA[j * D + i]), cfromreal(2.0));
    for (int i = 0; i < D; i++) outb[i] = accb[i];
    free(strides); free(idx); free(dA); free(db); free(accA); free(accb);
}

/* vanilla_batch_vjp_numba — vanilla/gradients.py:85-116 */
typedef struct {
    int D; const int64_t *shape; const double *G, *c, *g; double *dA, *db, *dc;
    int64_t B, N; int tid, nthreads;
} vjp_job;

static void *vjp_worker(void *arg) {
    vjp_job *j = (vjp_job *)arg;
    int D = j->D;
    for (int64_t k = j->tid; k < j->B; k += j->nthreads)
        mmo_vanilla_vjp(D, j->shape, j->G + 2 * j->N * k, j->c + 2 * k, j->g + 2 * j->N * k,
                        j->dA + 2 * (int64_t)D * D * k, j->db + 2 * (int64_t)D * k, j->dc + 2 * k);
    return NULL;
}

void mmo_vanilla_batch_vjp(int64_t B, int D, const int64_t *shape, const double *G, const double *c,
                           const double *dLdG, double *dLdA, double *dLdb, double *dLdc, int nthreads) {
    int64_t N = 1;
    for (int i = 0; i < D; i++) N *= shape[i];
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    vjp_job jobs[256];
    for (int t = 0; t < nthreads; t++) {
        vjp_job jb = { D, shape, G, c, dLdG, dLdA, dLdb, dLdc, B, N, t, nthreads };
        jobs[t] = jb;
        if (t > 0) pthread_create(&th[t], NULL, vjp_worker, &jobs[t]);
    }
    vjp_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------------
 * binomial — strategies/binomial.py:30-72, steps.binomial_step (steps.py:208-235),
 * steps.vanilla_step (steps.py:37-67), pivots.first_available_pivot (pivots.py:21-34),
 * neighbors.lower_neighbors (neighbors.py:42-46), paths.binomial_subspace_basis (paths.py:24-72).
 * Returns the L2 norm accumulated (binomial.py:58) through norm_out; G is zero-initialised.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int D; const int64_t *shape; const int64_t *strides; const cplx *A, *b; cplx *G; int64_t N;
    double norm;
} binom_ctx;

static void binom_point(binom_ctx *x, const int64_t *idx) {
    int D = x->D;
    int i = 0;
    while (i < D && idx[i] == 0) i++;                 /* first_available_pivot */
    int64_t flat = 0;
    for (int d = 0; d < D; d++) flat += idx[d] * x->strides[d];
    int64_t pivot = flat - x->strides[i];
    cplx v = cmul(x->b[i], x->G[pivot]);              /* steps.py:60 */
    for (int j = 0; j < D; j++) {                     /* steps.py:63-64, all lower neighbours of the pivot */
        int64_t pj = idx[j] - (j == i ? 1 : 0);       /* pivot[j] */
        /* neighbour index pivot[j]-1 wraps to shape[j]-1 when pivot[j]==0 (numpy negative index) */
        int64_t nb = (pj > 0) ? pivot - x->strides[j] : pivot + (x->shape[j] - 1) * x->strides[j];
        v = cadd(v, cmul(cmulr(x->A[i * D + j], SQRT_(pj)), x->G[nb]));
    }
    cplx val = cdivr(v, SQRT_(idx[i]));               /* steps.py:66 */
    x->G[flat] = val;
    double a = hypot(val.re, val.im);                 /* np.abs(value) ** 2 (steps.py:233) */
    x->norm = x->norm + a * a;
}

static void binom_enum(binom_ctx *x, int64_t weight, int mode, int64_t *idx) {
    /* paths.py:24-54: lexicographic enumeration of all idx with sum == weight, idx[m] < cutoffs[m] */
    if (mode == x->D) { if (weight == 0) binom_point(x, idx); return; }
    for (int64_t ph = 0; ph < x->shape[mode]; ph++) {
        if (weight - ph >= 0) { idx[mode] = ph; binom_enum(x, weight - ph, mode + 1, idx); }
    }
}

void mmo_binomial(int D, const int64_t *shape, const double *A_, const double *b_, const double *c_,
                  double max_l2, int64_t global_cutoff, double *G_, double *norm_out) {
    int64_t *strides = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
    int64_t *idx = (int64_t *)calloc(D + 1, sizeof(int64_t));
    make_strides(D, shape, strides);
    int64_t N = 1;
    for (int i = 0; i < D; i++) N *= shape[i];
    memset(G_, 0, sizeof(cplx) * N);
    cplx *G = (cplx *)G_;
    G[0].re = c_[0]; G[0].im = c_[1];
    double a0 = hypot(c_[0], c_[1]);
    double norm = a0 * a0;                              /* binomial.py:55 */
    binom_ctx x = { D, shape, strides, (const cplx *)A_, (const cplx *)b_, G, N, 0.0 };
    for (int64_t photons = 1; photons < global_cutoff; photons++) {  /* :58 */
        x.norm = 0.0;
        binom_enum(&x, photons, 0, idx);
        norm += x.norm;                                 /* :65 */
        if (norm > max_l2) break;                       /* :67 */
    }
    *norm_out = norm;
    free(strides); free(idx);
}
