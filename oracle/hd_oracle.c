/*
 * hd_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the linear-analysis hot path of Konstantin8105/hd.
 * Every routine cites the reference file:line it follows (paths relative to the
 * reference repository root).  Arithmetic is evaluated in the reference's order,
 * in FP64, without FMA contraction (build with -ffp-contract=off): Go on amd64
 * never fuses a*b+c.
 *
 * Third-party arithmetic that is NOT in the reference tree is restated from the
 * published algorithms it ports:
 *   - github.com/Konstantin8105/sparse v0.0.0-20240530141733-b591fd4a631c
 *     (go.mod:10) = CSparse (T. Davis): cs_compress, cs_fkeep, cs_dupl, cs_reach,
 *     cs_spsolve, cs_lu (natural order, partial pivoting tol = 1), cs_lsolve,
 *     cs_usolve.  [ordering/tolerance recalled, unverifiable here]
 *   - gonum.org/v1/gonum v0.17.0 (go.mod:11): mat.Dense.Mul -> native Dgemm
 *     (accumulate over the inner index ascending, skip exact-zero multipliers,
 *     mul then add), mat.Eigen -> Dgeev, restated with the EISPACK sequence
 *     balanc / orthes / ortran / hqr2 / balbak that LAPACK's dgebal / dgehrd /
 *     dhseqr / dtrevc / dgebak descend from, plus dgeev's unit-2-norm scaling.
 *   - github.com/Konstantin8105/pow v0.1.0 (go.mod:9): E2(x)=x*x, E3(x)=x*x*x.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against all 14 golden
 * reports of the reference (tests/golden/golden.json, 6 significant digits) and
 * tests/test_oracle_lapack.py against LAPACK.  See hd_oracle.h.
 */
#include "hd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GRAVITY 9.80665 /* model.go:713 */

/* ------------------------------------------------------------------------- */
/* small helpers                                                               */
/* ------------------------------------------------------------------------- */

void hdo_free(void *p) { free(p); }

/* Go math.Hypot (pure-Go and amd64 asm use the same scaling formula). */
static double go_hypot(double p, double q)
{
    p = fabs(p);
    q = fabs(q);
    if (isinf(p) || isinf(q)) return INFINITY;
    if (isnan(p) || isnan(q)) return NAN;
    if (p < q) { double t = p; p = q; q = t; }
    if (p == 0) return 0;
    q = q / p;
    return p * sqrt(1 + q * q);
}

/* feBeam2d.go:11-19 — start minus end inside Hypot */
double hdo_distance(const double xy[4]) { return go_hypot(xy[0] - xy[2], xy[1] - xy[3]); }

/* gonum native Dgemm, C = A(m x k) * B(k x n), beta = 0, alpha = 1:
 * for each (i,l): tmp = A[i,l]; if tmp != 0: C[i,:] += tmp * B[l,:]   (mul, then add) */
static void go_dgemm(int m, int n, int k, const double *a, int lda, const double *b, int ldb,
                     double *c, int ldc)
{
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) c[i * ldc + j] = 0.0;
    for (int i = 0; i < m; i++) {
        for (int l = 0; l < k; l++) {
            double tmp = a[i * lda + l];
            if (tmp != 0) {
                for (int j = 0; j < n; j++) {
                    double prod = tmp * b[l * ldb + j];
                    c[i * ldc + j] = c[i * ldc + j] + prod;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* element routines                                                            */
/* ------------------------------------------------------------------------- */

/* feBeam2d.go:22-140 */
void hdo_stiff_beam2d(const double xy[4], double A, double J, double E, const uint8_t pins[6],
                      double kr[36])
{
#define K(i, j) kr[(i) * 6 + (j)]
    memset(kr, 0, 36 * sizeof(double));
    double length = hdo_distance(xy);

    double EFL = E * A / length; /* :28 */
    K(0, 0) = +EFL;
    K(0, 3) = -EFL;
    K(3, 0) = -EFL;
    K(3, 3) = +EFL;

    double EJL3 = 12.0 * E * J / (length * length * length); /* :34, pow.E3 */
    K(1, 1) = +EJL3;
    K(4, 4) = +EJL3;
    K(1, 4) = -EJL3;
    K(4, 1) = -EJL3;

    double EJL2 = 6.0 * E * J / (length * length); /* :40, pow.E2 */
    K(1, 2) = +EJL2;
    K(2, 1) = +EJL2;
    K(1, 5) = +EJL2;
    K(5, 1) = +EJL2;
    K(2, 4) = -EJL2;
    K(4, 2) = -EJL2;
    K(4, 5) = -EJL2;
    K(5, 4) = -EJL2;

    double EJL = 2.0 * E * J / length; /* :50 */
    K(2, 5) = EJL;
    K(5, 2) = EJL;

    EJL = 2 * EJL; /* :54 */
    K(2, 2) = EJL;
    K(5, 5) = EJL;

    /* :58-137 pin condensation, literal including the write-back index rule */
    for (int fr = 0; fr < 6; fr++) {
        if (!pins[fr]) continue;
        double KBB[25], KBB2[25], KS[25], KBA[5], KAB[5];
        double KAA = K(fr, fr); /* :91 */
        for (int i = 0; i < 6; i++) {
            int ia = i;
            if (i == fr) continue;
            else if (i > fr) ia--;
            KBA[ia] = K(i, fr);
            KAB[ia] = K(fr, i);
            for (int j = 0; j < 6; j++) {
                int ja = j;
                if (j == fr) continue;
                else if (j > fr) ja--;
                KBB[ia * 5 + ja] = K(i, j);
            }
        }
        go_dgemm(5, 5, 1, KBA, 1, KAB, 5, KBB2, 5); /* :112 KBB2.Mul(KBA, KAB) */
        double f = 1 / KAA;
        for (int q = 0; q < 25; q++) KBB2[q] = f * KBB2[q]; /* :113 Scale */
        for (int q = 0; q < 25; q++) KS[q] = KBB[q] - KBB2[q]; /* :114 Sub */

        for (int i = 0; i < 5; i++) { /* :116-132 */
            int ia = i;
            if (i == fr) continue;
            else if (i > fr) ia++;
            for (int j = 0; j < 5; j++) {
                int ja = j;
                if (j == fr) continue;
                else if (j > fr) ja++;
                K(ia, ja) = KS[i * 5 + j];
            }
        }
        for (int i = 0; i < 6; i++) { /* :133-136 */
            K(i, fr) = 0.0;
            K(fr, i) = 0.0;
        }
    }
#undef K
}

/* feBeam2d.go:142-172 — end minus start for the direction cosines */
void hdo_trans_beam2d(const double xy[4], double tr[36])
{
    memset(tr, 0, 36 * sizeof(double));
    double length = hdo_distance(xy);
    double lambdaXX = (xy[2] - xy[0]) / length;
    double lambdaXY = (xy[3] - xy[1]) / length;
    double lambdaYX = -lambdaXY;
    double lambdaYY = +lambdaXX;
    tr[0 * 6 + 0] = lambdaXX;
    tr[3 * 6 + 3] = lambdaXX;
    tr[0 * 6 + 1] = lambdaXY;
    tr[3 * 6 + 4] = lambdaXY;
    tr[1 * 6 + 0] = lambdaYX;
    tr[4 * 6 + 3] = lambdaYX;
    tr[1 * 6 + 1] = lambdaYY;
    tr[4 * 6 + 4] = lambdaYY;
    tr[2 * 6 + 2] = 1.0;
    tr[5 * 6 + 5] = 1.0;
}

/* feBeam2d.go:175-275 */
void hdo_geometric_beam2d(const double xy[4], const uint8_t pins[6], const double bf[6], double kr[36])
{
#define K(i, j) kr[(i) * 6 + (j)]
    memset(kr, 0, 36 * sizeof(double));
    double axialForce = 0; /* :180-188 */
    if (bf[0] > 0) axialForce = bf[0];
    if (-bf[3] > axialForce) axialForce = -bf[3];
    double length = hdo_distance(xy);
    if (!pins[2] && !pins[5]) { /* :197 */
        double G1 = 6. / (5. * length);
        K(1, 1) = +G1; K(4, 4) = +G1; K(1, 4) = -G1; K(4, 1) = -G1;
        K(1, 2) = +0.1; K(2, 1) = +0.1; K(1, 5) = +0.1; K(5, 1) = +0.1;
        K(2, 4) = -0.1; K(4, 2) = -0.1; K(5, 4) = -0.1; K(4, 5) = -0.1;
        double G2 = 2. * length / 15.;
        K(2, 2) = +G2; K(5, 5) = +G2;
        double G3 = -length / 30.;
        K(5, 2) = +G3; K(2, 5) = +G3;
    } else if (pins[2] && !pins[5]) { /* :226 */
        double G1 = 6. / (5. * length);
        K(1, 1) = +G1; K(4, 4) = +G1; K(1, 4) = -G1; K(4, 1) = -G1;
        K(1, 5) = +1. / 5.; K(5, 1) = +1. / 5.;
        K(4, 5) = -1. / 5.; K(5, 4) = -1. / 5.;
        K(5, 5) = length / 5.0;
    } else if (!pins[2] && pins[5]) { /* :245 */
        double G1 = 6. / (5. * length);
        K(1, 1) = +G1; K(4, 4) = +G1; K(1, 4) = -G1; K(4, 1) = -G1;
        K(1, 2) = +1. / 5.; K(2, 1) = +1. / 5.;
        K(4, 2) = -1. / 5.; K(2, 4) = -1. / 5.;
        K(2, 2) = length / 5.0;
    } else { /* :264 */
        K(2, 2) = +1.0 / length; K(5, 5) = +1.0 / length;
        K(2, 5) = -1.0 / length; K(5, 2) = -1.0 / length;
    }
    for (int q = 0; q < 36; q++) kr[q] = axialForce * kr[q]; /* :190-192 deferred Scale */
#undef K
}

/* model.go:619-620: kr.Mul(tr.T(), kr); kr.Mul(kr, tr) */
void hdo_rotate_to_global(const double k[36], const double t[36], double kg[36])
{
    double tt[36], tmp[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) tt[i * 6 + j] = t[j * 6 + i];
    go_dgemm(6, 6, 6, tt, 6, k, 6, tmp, 6);
    go_dgemm(6, 6, 6, tmp, 6, t, 6, kg, 6);
}

static void beam_xy(const hdo_model *m, int b, double xy[4])
{
    int64_t st = m->beamN[2 * b], en = m->beamN[2 * b + 1];
    xy[0] = m->points[2 * st];
    xy[1] = m->points[2 * st + 1];
    xy[2] = m->points[2 * en];
    xy[3] = m->points[2 * en + 1];
}

/* ------------------------------------------------------------------------- */
/* CSparse restatement (the parts model.go uses through Konstantin8105/sparse) */
/* ------------------------------------------------------------------------- */

typedef struct {
    int n, nzmax;
    int *p, *i;
    double *x;
} csc;

static void csc_free(csc *A)
{
    if (!A) return;
    free(A->p);
    free(A->i);
    free(A->x);
    A->p = A->i = NULL;
    A->x = NULL;
}

static int csc_alloc(csc *A, int n, int nzmax)
{
    A->n = n;
    A->nzmax = nzmax > 0 ? nzmax : 1;
    A->p = (int *)calloc((size_t)n + 1, sizeof(int));
    A->i = (int *)malloc((size_t)A->nzmax * sizeof(int));
    A->x = (double *)malloc((size_t)A->nzmax * sizeof(double));
    return (A->p && A->i && A->x) ? 0 : -1;
}

static int csc_grow(csc *A, int nzmax)
{
    int *ni = (int *)realloc(A->i, (size_t)nzmax * sizeof(int));
    if (!ni) return -1;
    A->i = ni;
    double *nx = (double *)realloc(A->x, (size_t)nzmax * sizeof(double));
    if (!nx) return -1;
    A->x = nx;
    A->nzmax = nzmax;
    return 0;
}

/* cs_compress: triplet -> CSC, stable inside a column (model.go:639) */
static int cs_compress(int n, int nz, const int *Ti, const int *Tj, const double *Tx, csc *C)
{
    if (csc_alloc(C, n, nz)) return -1;
    int *w = (int *)calloc((size_t)n + 1, sizeof(int));
    if (!w) return -1;
    for (int k = 0; k < nz; k++) w[Tj[k]]++;
    int sum = 0;
    for (int j = 0; j < n; j++) {
        C->p[j] = sum;
        sum += w[j];
        w[j] = C->p[j];
    }
    C->p[n] = sum;
    for (int k = 0; k < nz; k++) {
        int p = w[Tj[k]]++;
        C->i[p] = Ti[k];
        C->x[p] = Tx[k];
    }
    free(w);
    return 0;
}

/* cs_fkeep with the predicate x != 0.0 (model.go:645-649) */
static void cs_drop_zeros(csc *A)
{
    int nz = 0;
    for (int j = 0; j < A->n; j++) {
        int p = A->p[j];
        A->p[j] = nz;
        for (; p < A->p[j + 1]; p++) {
            if (A->x[p] != 0.0) {
                A->x[nz] = A->x[p];
                A->i[nz++] = A->i[p];
            }
        }
    }
    A->p[A->n] = nz;
}

/* cs_dupl: sum duplicates, first-appearance order, summation in storage order (model.go:652) */
static int cs_dupl(csc *A)
{
    int n = A->n, nz = 0;
    int *w = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    if (!w) return -1;
    for (int i = 0; i < n; i++) w[i] = -1;
    for (int j = 0; j < n; j++) {
        int q = nz;
        for (int p = A->p[j]; p < A->p[j + 1]; p++) {
            int i = A->i[p];
            if (w[i] >= q) {
                A->x[w[i]] += A->x[p];
            } else {
                w[i] = nz;
                A->i[nz] = i;
                A->x[nz++] = A->x[p];
            }
        }
        A->p[j] = q;
    }
    A->p[n] = nz;
    free(w);
    return 0;
}

/* cs_dfs / cs_reach / cs_spsolve (lower-triangular G, unit or not, with pinv) */
#define CS_FLIP(i) (-(i) - 2)
#define CS_UNFLIP(i) (((i) < 0) ? CS_FLIP(i) : (i))
#define CS_MARKED(w, j) (w[j] < 0)
#define CS_MARK(w, j) { w[j] = CS_FLIP(w[j]); }

static int cs_dfs(int j, csc *G, int top, int *xi, int *pstack, const int *pinv)
{
    int head = 0, *Gp = G->p, *Gi = G->i;
    xi[0] = j;
    while (head >= 0) {
        j = xi[head];
        int jnew = pinv ? pinv[j] : j;
        if (!CS_MARKED(Gp, j)) {
            CS_MARK(Gp, j);
            pstack[head] = (jnew < 0) ? 0 : CS_UNFLIP(Gp[jnew]);
        }
        int done = 1;
        int p2 = (jnew < 0) ? 0 : CS_UNFLIP(Gp[jnew + 1]);
        for (int p = pstack[head]; p < p2; p++) {
            int i = Gi[p];
            if (CS_MARKED(Gp, i)) continue;
            pstack[head] = p;
            xi[++head] = i;
            done = 0;
            break;
        }
        if (done) {
            head--;
            xi[--top] = j;
        }
    }
    return top;
}

static int cs_reach(csc *G, const csc *B, int k, int *xi, const int *pinv)
{
    int n = G->n, top = n;
    for (int p = B->p[k]; p < B->p[k + 1]; p++)
        if (!CS_MARKED(G->p, B->i[p])) top = cs_dfs(B->i[p], G, top, xi, xi + n, pinv);
    for (int p = top; p < n; p++) CS_MARK(G->p, xi[p]);
    return top;
}

static int cs_spsolve_lower(csc *G, const csc *B, int k, int *xi, double *x, const int *pinv)
{
    int n = G->n;
    int top = cs_reach(G, B, k, xi, pinv);
    for (int p = top; p < n; p++) x[xi[p]] = 0;
    for (int p = B->p[k]; p < B->p[k + 1]; p++) x[B->i[p]] = B->x[p];
    for (int px = top; px < n; px++) {
        int j = xi[px];
        int J = pinv ? pinv[j] : j;
        if (J < 0) continue;
        x[j] /= G->x[G->p[J]];
        int p = G->p[J] + 1, q = G->p[J + 1];
        for (; p < q; p++) x[G->i[p]] -= G->x[p] * x[j];
    }
    return top;
}

typedef struct {
    int n;
    csc L, U;
    int *pinv;
    /* elimination bookkeeping of sparse.LU.Factorize(A, ignore...) */
    int nFull;
    int *freeOf; /* [nFull] reduced index or -1 */
    double *work; /* [n] */
} lu_t;

static void lu_free(lu_t *F)
{
    csc_free(&F->L);
    csc_free(&F->U);
    free(F->pinv);
    free(F->freeOf);
    free(F->work);
    memset(F, 0, sizeof(*F));
}

/* cs_lu, natural column order, partial pivoting with tol (CSparse) */
static int cs_lu(const csc *A, double tol, lu_t *F)
{
    int n = A->n;
    int guess = 4 * A->p[n] + n;
    F->n = n;
    if (csc_alloc(&F->L, n, guess) || csc_alloc(&F->U, n, guess)) return HDO_ERR_ALLOC;
    F->pinv = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    double *x = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    int *xi = (int *)malloc((size_t)(2 * n > 0 ? 2 * n : 1) * sizeof(int));
    if (!F->pinv || !x || !xi) { free(x); free(xi); return HDO_ERR_ALLOC; }
    csc *L = &F->L, *U = &F->U;
    int *pinv = F->pinv;
    for (int i = 0; i < n; i++) pinv[i] = -1;
    for (int k = 0; k <= n; k++) L->p[k] = 0;
    int lnz = 0, unz = 0, rc = HDO_OK;
    for (int k = 0; k < n; k++) {
        L->p[k] = lnz;
        U->p[k] = unz;
        if (lnz + n > L->nzmax && csc_grow(L, 2 * L->nzmax + n)) { rc = HDO_ERR_ALLOC; break; }
        if (unz + n > U->nzmax && csc_grow(U, 2 * U->nzmax + n)) { rc = HDO_ERR_ALLOC; break; }
        int col = k;
        int top = cs_spsolve_lower(L, A, col, xi, x, pinv);
        int ipiv = -1;
        double a = -1;
        for (int p = top; p < n; p++) {
            int i = xi[p];
            if (pinv[i] < 0) {
                double t = fabs(x[i]);
                if (t > a) { a = t; ipiv = i; }
            } else {
                U->i[unz] = pinv[i];
                U->x[unz++] = x[i];
            }
        }
        if (ipiv == -1 || a <= 0) { rc = HDO_ERR_SINGULAR; break; }
        if (pinv[col] < 0 && fabs(x[col]) >= a * tol) ipiv = col;
        double pivot = x[ipiv];
        U->i[unz] = k;
        U->x[unz++] = pivot;
        pinv[ipiv] = k;
        L->i[lnz] = ipiv;
        L->x[lnz++] = 1;
        for (int p = top; p < n; p++) {
            int i = xi[p];
            if (pinv[i] < 0) {
                L->i[lnz] = i;
                L->x[lnz++] = x[i] / pivot;
            }
            x[i] = 0;
        }
    }
    if (rc == HDO_OK) {
        L->p[n] = lnz;
        U->p[n] = unz;
        for (int p = 0; p < lnz; p++) L->i[p] = pinv[L->i[p]];
    }
    free(x);
    free(xi);
    return rc;
}

/* sparse.LU.Factorize(A, ignore...): drop the listed rows and columns (sorted,
 * unique), renumber, factorize.  model.go:326-328.  [recalled semantics] */
static int lu_factorize(const csc *K, int nElim, const int *elim, lu_t *F)
{
    memset(F, 0, sizeof(*F));
    int N = K->n;
    F->nFull = N;
    F->freeOf = (int *)malloc((size_t)(N > 0 ? N : 1) * sizeof(int));
    if (!F->freeOf) return HDO_ERR_ALLOC;
    for (int i = 0; i < N; i++) F->freeOf[i] = 0;
    for (int e = 0; e < nElim; e++) F->freeOf[elim[e]] = -1;
    int n = 0;
    for (int i = 0; i < N; i++) F->freeOf[i] = (F->freeOf[i] < 0) ? -1 : n++;
    csc A;
    if (csc_alloc(&A, n, K->p[N])) return HDO_ERR_ALLOC;
    int nz = 0;
    for (int j = 0; j < N; j++) {
        int jj = F->freeOf[j];
        if (jj < 0) continue;
        A.p[jj] = nz;
        for (int p = K->p[j]; p < K->p[j + 1]; p++) {
            int ii = F->freeOf[K->i[p]];
            if (ii < 0) continue;
            A.i[nz] = ii;
            A.x[nz++] = K->x[p];
        }
    }
    A.p[n] = nz;
    int rc = cs_lu(&A, 1.0, F);
    csc_free(&A);
    if (rc == HDO_OK) {
        F->work = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
        if (!F->work) rc = HDO_ERR_ALLOC;
    }
    return rc;
}

/* sparse.LU.Solve(b): compact, x = P b, L solve, U solve, expand with zeros (model.go:440) */
static void lu_solve(const lu_t *F, const double *b, double *d)
{
    int n = F->n;
    double *x = F->work;
    for (int i = 0; i < F->nFull; i++) {
        int r = F->freeOf[i];
        if (r >= 0) x[F->pinv[r]] = b[i];
    }
    const csc *L = &F->L, *U = &F->U;
    for (int j = 0; j < n; j++) { /* cs_lsolve */
        x[j] /= L->x[L->p[j]];
        for (int p = L->p[j] + 1; p < L->p[j + 1]; p++) x[L->i[p]] -= L->x[p] * x[j];
    }
    for (int j = n - 1; j >= 0; j--) { /* cs_usolve */
        x[j] /= U->x[U->p[j + 1] - 1];
        for (int p = U->p[j]; p < U->p[j + 1] - 1; p++) x[U->i[p]] -= U->x[p] * x[j];
    }
    for (int i = 0; i < F->nFull; i++) {
        int r = F->freeOf[i];
        d[i] = (r >= 0) ? x[r] : 0.0;
    }
}

/* ------------------------------------------------------------------------- */
/* assembly (model.go:607-681)                                                 */
/* ------------------------------------------------------------------------- */

/* which: 0 getStiffBeam2d, 1 getGeometricBeam2d(pos, lc) */
static int assembly_k(const hdo_model *m, int which, const double *beamForces, csc *K,
                      int *nIgnore, int **ignoreOut)
{
    int nB = m->nB, dof = 3 * m->nP;
    int nz = 36 * nB;
    int *Ti = (int *)malloc((size_t)(nz > 0 ? nz : 1) * sizeof(int));
    int *Tj = (int *)malloc((size_t)(nz > 0 ? nz : 1) * sizeof(int));
    double *Tx = (double *)malloc((size_t)(nz > 0 ? nz : 1) * sizeof(double));
    if (!Ti || !Tj || !Tx) { free(Ti); free(Tj); free(Tx); return HDO_ERR_ALLOC; }
    int q = 0, rc = HDO_OK;
    for (int b = 0; b < nB; b++) { /* :615-636 */
        double xy[4], kr[36], tr[36], kg[36];
        beam_xy(m, b, xy);
        if (which == 0)
            hdo_stiff_beam2d(xy, m->sections[3 * b], m->sections[3 * b + 1], m->sections[3 * b + 2],
                             m->pins + 6 * b, kr);
        else
            hdo_geometric_beam2d(xy, m->pins + 6 * b, beamForces + 6 * b, kr);
        hdo_trans_beam2d(xy, tr);
        hdo_rotate_to_global(kr, tr, kg);
        for (int p1 = 0; p1 < 2; p1++)
            for (int p2 = 0; p2 < 2; p2++)
                for (int r1 = 0; r1 < 3; r1++)
                    for (int r2 = 0; r2 < 3; r2++) {
                        double val = kg[(p1 * 3 + r1) * 6 + (p2 * 3 + r2)];
                        /* sparse.Entry rejects NaN/Inf values [recalled] */
                        if (isnan(val) || isinf(val)) rc = HDO_ERR_INPUT;
                        Ti[q] = (int)(m->beamN[2 * b + p1] * 3 + r1);
                        Tj[q] = (int)(m->beamN[2 * b + p2] * 3 + r2);
                        Tx[q] = val;
                        q++;
                    }
    }
    if (rc == HDO_OK && cs_compress(dof, nz, Ti, Tj, Tx, K)) rc = HDO_ERR_ALLOC;
    free(Ti);
    free(Tj);
    free(Tx);
    if (rc != HDO_OK) return rc;
    cs_drop_zeros(K);          /* :645 */
    if (cs_dupl(K)) return HDO_ERR_ALLOC; /* :652 */
    /* :657-678 zero diagonals -> ignore */
    uint8_t *bz = (uint8_t *)calloc((size_t)(dof > 0 ? dof : 1), 1);
    int *ign = (int *)malloc((size_t)(dof > 0 ? dof : 1) * sizeof(int));
    if (!bz || !ign) { free(bz); free(ign); return HDO_ERR_ALLOC; }
    for (int j = 0; j < dof; j++)
        for (int p = K->p[j]; p < K->p[j + 1]; p++)
            if (K->i[p] == j) bz[j] = 1;
    int ni = 0;
    for (int i = 0; i < dof; i++)
        if (!bz[i]) ign[ni++] = i;
    free(bz);
    *nIgnore = ni;
    *ignoreOut = ign;
    return HDO_OK;
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* append(ignore, m.addSupport()...) then sort-unique (model.go:326-328, 700-710) */
static int eliminated_list(const hdo_model *m, int nIgnore, const int *ignore, int *nElim, int **elimOut)
{
    int dof = 3 * m->nP;
    int *e = (int *)malloc((size_t)(dof + nIgnore + 1) * sizeof(int));
    if (!e) return HDO_ERR_ALLOC;
    int n = 0;
    for (int i = 0; i < nIgnore; i++) e[n++] = ignore[i];
    for (int p = 0; p < m->nP; p++)
        for (int i = 0; i < 3; i++)
            if (m->supports[3 * p + i]) e[n++] = 3 * p + i;
    qsort(e, (size_t)n, sizeof(int), cmp_int);
    int u = 0;
    for (int i = 0; i < n; i++)
        if (u == 0 || e[u - 1] != e[i]) e[u++] = e[i];
    *nElim = u;
    *elimOut = e;
    return HDO_OK;
}

/* getK (model.go:318-333) */
static int get_k(const hdo_model *m, csc *K, lu_t *F, int *nIgnore, int **ignore)
{
    memset(K, 0, sizeof(*K));
    memset(F, 0, sizeof(*F));
    *ignore = NULL;
    int rc = assembly_k(m, 0, NULL, K, nIgnore, ignore);
    if (rc != HDO_OK) return rc;
    int nElim, *elim;
    rc = eliminated_list(m, *nIgnore, *ignore, &nElim, &elim);
    if (rc != HDO_OK) return rc;
    rc = lu_factorize(K, nElim, elim, F);
    free(elim);
    return rc;
}

int hdo_assemble(const hdo_model *m, int which, const double *beamForces, int32_t *dof, int32_t *nnz,
                 int32_t **cscColptr, int32_t **cscRowind, double **cscVal, int32_t **csrRowptr,
                 int32_t **csrColind, double **csrVal, int32_t *nIgnore, int32_t **ignore)
{
    csc K;
    memset(&K, 0, sizeof(K));
    int ni, *ign = NULL;
    int rc = assembly_k(m, which, beamForces, &K, &ni, &ign);
    if (rc != HDO_OK) { csc_free(&K); free(ign); return rc; }
    int n = K.n, nz = K.p[n];
    *dof = n;
    *nnz = nz;
    *nIgnore = ni;
    *ignore = ign;
    *cscColptr = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    *cscRowind = (int32_t *)malloc((size_t)(nz + 1) * sizeof(int32_t));
    *cscVal = (double *)malloc((size_t)(nz + 1) * sizeof(double));
    *csrRowptr = (int32_t *)calloc((size_t)n + 1, sizeof(int32_t));
    *csrColind = (int32_t *)malloc((size_t)(nz + 1) * sizeof(int32_t));
    *csrVal = (double *)malloc((size_t)(nz + 1) * sizeof(double));
    memcpy(*cscColptr, K.p, ((size_t)n + 1) * sizeof(int));
    memcpy(*cscRowind, K.i, (size_t)nz * sizeof(int));
    memcpy(*cscVal, K.x, (size_t)nz * sizeof(double));
    /* canonical CSR, ascending columns: counting pass over CSC columns in order */
    int32_t *rp = *csrRowptr;
    for (int p = 0; p < nz; p++) rp[K.i[p] + 1]++;
    for (int i = 0; i < n; i++) rp[i + 1] += rp[i];
    int *w = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; i++) w[i] = rp[i];
    for (int j = 0; j < n; j++)
        for (int p = K.p[j]; p < K.p[j + 1]; p++) {
            int q = w[K.i[p]]++;
            (*csrColind)[q] = j;
            (*csrVal)[q] = K.x[p];
        }
    free(w);
    csc_free(&K);
    return HDO_OK;
}

int hdo_eliminated(const hdo_model *m, int32_t *nElim, int32_t **elim, int32_t *nIgnore, int32_t **ignore)
{
    csc K;
    memset(&K, 0, sizeof(K));
    int ni, *ign = NULL;
    int rc = assembly_k(m, 0, NULL, &K, &ni, &ign);
    csc_free(&K);
    if (rc != HDO_OK) { free(ign); return rc; }
    int ne, *e;
    rc = eliminated_list(m, ni, ign, &ne, &e);
    if (rc != HDO_OK) { free(ign); return rc; }
    *nElim = ne;
    *elim = e;
    *nIgnore = ni;
    *ignore = ign;
    return HDO_OK;
}

/* ------------------------------------------------------------------------- */
/* LinearStatic (model.go:358-490)                                             */
/* ------------------------------------------------------------------------- */

static int g_literal_scans = 1; /* 1: O(nSup*nnz) Fkeep scan as model.go:481; 0: same sums via row lists */

void hdo_set_literal_scans(int on) { g_literal_scans = on; }

static int static_with_factor(const hdo_model *m, const csc *K, const lu_t *F, int nIgnore,
                              const int *ignore, int nLC, const double *loads, double *disp,
                              double *forces, double *react, int32_t *failCase)
{
    int nP = m->nP, nB = m->nB, dof = 3 * nP;
    double *d = (double *)malloc((size_t)(dof > 0 ? dof : 1) * sizeof(double));
    /* row lists for the non-literal reaction path (same summation order: columns ascending) */
    int *rptr = NULL, *rcol = NULL, *rpos = NULL;
    if (!d) return HDO_ERR_ALLOC;
    if (!g_literal_scans && react) {
        int nz = K->p[dof];
        rptr = (int *)calloc((size_t)dof + 1, sizeof(int));
        rcol = (int *)malloc((size_t)(nz + 1) * sizeof(int));
        rpos = (int *)malloc((size_t)(nz + 1) * sizeof(int));
        int *w = (int *)malloc((size_t)(dof + 1) * sizeof(int));
        for (int p = 0; p < nz; p++) rptr[K->i[p] + 1]++;
        for (int i = 0; i < dof; i++) rptr[i + 1] += rptr[i];
        for (int i = 0; i < dof; i++) w[i] = rptr[i];
        for (int j = 0; j < dof; j++)
            for (int p = K->p[j]; p < K->p[j + 1]; p++) {
                int q = w[K->i[p]]++;
                rcol[q] = j;
                rpos[q] = p;
            }
        free(w);
    }
    int rc = HDO_OK;
    for (int lc = 0; lc < nLC; lc++) { /* :415 */
        const double *p = loads + (size_t)lc * dof;
        for (int i = 0; i < dof; i++) /* :692-696 */
            if (isnan(p[i]) || isinf(p[i])) { rc = HDO_ERR_INPUT; break; }
        if (rc != HDO_OK) { if (failCase) *failCase = lc; break; }
        for (int q = 0; q < nIgnore; q++) /* :424-437 */
            if (p[ignore[q]] != 0.0) { rc = HDO_ERR_IGNORE_LOAD; break; }
        if (rc != HDO_OK) { if (failCase) *failCase = lc; break; }

        lu_solve(F, p, d); /* :440 */

        if (disp) /* :446-451 */
            for (int i = 0; i < dof; i++) disp[(size_t)lc * dof + i] = d[i];

        if (forces) { /* :455-468, recomputes tr and kr for every beam of every case */
            for (int b = 0; b < nB; b++) {
                double Zo[6], z[6], xy[4], tr[36], kr[36];
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 2; j++) Zo[j * 3 + i] = d[m->beamN[2 * b + j] * 3 + i];
                beam_xy(m, b, xy);
                hdo_trans_beam2d(xy, tr);
                go_dgemm(6, 1, 6, tr, 6, Zo, 1, z, 1);
                hdo_stiff_beam2d(xy, m->sections[3 * b], m->sections[3 * b + 1], m->sections[3 * b + 2],
                                 m->pins + 6 * b, kr);
                go_dgemm(6, 1, 6, kr, 6, z, 1, forces + ((size_t)lc * nB + b) * 6, 1);
            }
        }

        if (react) { /* :470-489 */
            double *R = react + (size_t)lc * dof;
            for (int i = 0; i < dof; i++) R[i] = 0.0;
            for (int pt = 0; pt < nP; pt++)
                for (int i = 0; i < 3; i++) {
                    if (!m->supports[3 * pt + i]) continue;
                    int row = 3 * pt + i;
                    double r = -p[row];
                    if (g_literal_scans) {
                        for (int j = 0; j < dof; j++)
                            for (int q = K->p[j]; q < K->p[j + 1]; q++)
                                if (K->i[q] == row) r += K->x[q] * d[j];
                    } else {
                        for (int q = rptr[row]; q < rptr[row + 1]; q++) r += K->x[rpos[q]] * d[rcol[q]];
                    }
                    R[row] = r;
                }
        }
    }
    free(d);
    free(rptr);
    free(rcol);
    free(rpos);
    return rc;
}

int hdo_linear_static(const hdo_model *m, int32_t nLC, const double *loads, double *disp,
                      double *forces, double *react, int32_t *failCase)
{
    csc K;
    lu_t F;
    int nIgnore = 0, *ignore = NULL;
    if (failCase) *failCase = -1;
    int rc = get_k(m, &K, &F, &nIgnore, &ignore); /* :410 */
    if (rc == HDO_OK)
        rc = static_with_factor(m, &K, &F, nIgnore, ignore, nLC, loads, disp, forces, react, failCase);
    csc_free(&K);
    lu_free(&F);
    free(ignore);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* dense non-symmetric eigen solver: balanc, orthes, ortran, hqr2, balbak       */
/* (EISPACK sequence; stand-in for gonum mat.Eigen / LAPACK dgeev)              */
/* ------------------------------------------------------------------------- */

#define A_(i, j) a[(size_t)(i) * n + (j)]
#define V_(i, j) v[(size_t)(i) * n + (j)]

static void eig_exchange(int n, double *a, double *scale, int j, int m, int k, int l)
{
    scale[m] = (double)j;
    if (j == m) return;
    for (int i = 0; i <= l; i++) { double f = A_(i, j); A_(i, j) = A_(i, m); A_(i, m) = f; }
    for (int i = k; i < n; i++) { double f = A_(j, i); A_(j, i) = A_(m, i); A_(m, i) = f; }
}

static void eig_balance(int n, double *a, int *low, int *igh, double *scale)
{
    const double radix = 2.0, b2 = 4.0;
    int k = 0, l = n - 1;
    /* rows isolating an eigenvalue -> push down */
    for (;;) {
        int found = 0;
        for (int j = l; j >= 0; j--) {
            int iso = 1;
            for (int i = 0; i <= l; i++) {
                if (i == j) continue;
                if (A_(j, i) != 0.0) { iso = 0; break; }
            }
            if (iso) {
                eig_exchange(n, a, scale, j, l, k, l);
                found = 1;
                break;
            }
        }
        if (!found) break;
        if (l == 0) { *low = 0; *igh = 0; return; }
        l--;
    }
    /* columns isolating an eigenvalue -> push left */
    for (;;) {
        int found = 0;
        for (int j = k; j <= l; j++) {
            int iso = 1;
            for (int i = k; i <= l; i++) {
                if (i == j) continue;
                if (A_(i, j) != 0.0) { iso = 0; break; }
            }
            if (iso) {
                eig_exchange(n, a, scale, j, k, k, l);
                found = 1;
                break;
            }
        }
        if (!found) break;
        k++;
    }
    for (int i = k; i <= l; i++) scale[i] = 1.0;
    int noconv = 1;
    while (noconv) {
        noconv = 0;
        for (int i = k; i <= l; i++) {
            double c = 0.0, r = 0.0;
            for (int j = k; j <= l; j++) {
                if (j == i) continue;
                c += fabs(A_(j, i));
                r += fabs(A_(i, j));
            }
            if (c == 0.0 || r == 0.0) continue;
            double g = r / radix, f = 1.0, s = c + r;
            while (c < g) { f *= radix; c *= b2; }
            g = r * radix;
            while (c >= g) { f /= radix; c /= b2; }
            if ((c + r) / f >= 0.95 * s) continue;
            g = 1.0 / f;
            scale[i] *= f;
            noconv = 1;
            for (int j = k; j < n; j++) A_(i, j) *= g;
            for (int j = 0; j <= l; j++) A_(j, i) *= f;
        }
    }
    *low = k;
    *igh = l;
}

static void eig_orthes(int n, int low, int high, double *a, double *v, double *ort)
{
    for (int m = low + 1; m <= high - 1; m++) {
        double scale = 0.0;
        for (int i = m; i <= high; i++) scale += fabs(A_(i, m - 1));
        if (scale != 0.0) {
            double h = 0.0;
            for (int i = high; i >= m; i--) {
                ort[i] = A_(i, m - 1) / scale;
                h += ort[i] * ort[i];
            }
            double g = sqrt(h);
            if (ort[m] > 0) g = -g;
            h = h - ort[m] * g;
            ort[m] = ort[m] - g;
            for (int j = m; j < n; j++) {
                double f = 0.0;
                for (int i = high; i >= m; i--) f += ort[i] * A_(i, j);
                f = f / h;
                for (int i = m; i <= high; i++) A_(i, j) -= f * ort[i];
            }
            for (int i = 0; i <= high; i++) {
                double f = 0.0;
                for (int j = high; j >= m; j--) f += ort[j] * A_(i, j);
                f = f / h;
                for (int j = m; j <= high; j++) A_(i, j) -= f * ort[j];
            }
            ort[m] = scale * ort[m];
            A_(m, m - 1) = scale * g;
        }
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V_(i, j) = (i == j) ? 1.0 : 0.0;
    for (int m = high - 1; m >= low + 1; m--) {
        if (A_(m, m - 1) != 0.0) {
            for (int i = m + 1; i <= high; i++) ort[i] = A_(i, m - 1);
            for (int j = m; j <= high; j++) {
                double g = 0.0;
                for (int i = m; i <= high; i++) g += ort[i] * V_(i, j);
                g = (g / ort[m]) / A_(m, m - 1);
                for (int i = m; i <= high; i++) V_(i, j) += g * ort[i];
            }
        }
    }
}

static void cdiv(double xr, double xi, double yr, double yi, double *cr, double *ci)
{
    double r, d;
    if (fabs(yr) > fabs(yi)) {
        r = yi / yr;
        d = yr + r * yi;
        *cr = (xr + r * xi) / d;
        *ci = (xi - r * xr) / d;
    } else {
        r = yr / yi;
        d = yi + r * yr;
        *cr = (r * xr + xi) / d;
        *ci = (r * xi - xr) / d;
    }
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* returns 0 ok, 1 no convergence */
static int eig_hqr2(int nn, int low, int high, double *a, double *v, double *d, double *e)
{
    const int n0 = nn;
    int n = high; /* EISPACK: en = igh; roots outside low..igh are already isolated */
#undef A_
#undef V_
#define A_(i, j) a[(size_t)(i) * n0 + (j)]
#define V_(i, j) v[(size_t)(i) * n0 + (j)]
    const double eps = 2.220446049250313e-16;
    double exshift = 0.0;
    double p = 0, q = 0, r = 0, s = 0, z = 0, t, w, x, y;
    double norm = 0.0;
    for (int i = 0; i < nn; i++) {
        if (i < low || i > high) { d[i] = A_(i, i); e[i] = 0.0; }
        for (int j = imax(i - 1, 0); j < nn; j++) norm += fabs(A_(i, j));
    }
    int iter = 0;
    long total_iter = 0;
    while (n >= low) {
        int l = n;
        while (l > low) {
            s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
            if (s == 0.0) s = norm;
            if (fabs(A_(l, l - 1)) < eps * s) break;
            l--;
        }
        if (l == n) { /* one root */
            A_(n, n) = A_(n, n) + exshift;
            d[n] = A_(n, n);
            e[n] = 0.0;
            n--;
            iter = 0;
        } else if (l == n - 1) { /* two roots */
            w = A_(n, n - 1) * A_(n - 1, n);
            p = (A_(n - 1, n - 1) - A_(n, n)) / 2.0;
            q = p * p + w;
            z = sqrt(fabs(q));
            A_(n, n) = A_(n, n) + exshift;
            A_(n - 1, n - 1) = A_(n - 1, n - 1) + exshift;
            x = A_(n, n);
            if (q >= 0) { /* real pair */
                z = (p >= 0) ? p + z : p - z;
                d[n - 1] = x + z;
                d[n] = d[n - 1];
                if (z != 0.0) d[n] = x - w / z;
                e[n - 1] = 0.0;
                e[n] = 0.0;
                x = A_(n, n - 1);
                s = fabs(x) + fabs(z);
                p = x / s;
                q = z / s;
                r = sqrt(p * p + q * q);
                p = p / r;
                q = q / r;
                for (int j = n - 1; j < nn; j++) {
                    z = A_(n - 1, j);
                    A_(n - 1, j) = q * z + p * A_(n, j);
                    A_(n, j) = q * A_(n, j) - p * z;
                }
                for (int i = 0; i <= n; i++) {
                    z = A_(i, n - 1);
                    A_(i, n - 1) = q * z + p * A_(i, n);
                    A_(i, n) = q * A_(i, n) - p * z;
                }
                for (int i = low; i <= high; i++) {
                    z = V_(i, n - 1);
                    V_(i, n - 1) = q * z + p * V_(i, n);
                    V_(i, n) = q * V_(i, n) - p * z;
                }
            } else { /* complex pair */
                d[n - 1] = x + p;
                d[n] = x + p;
                e[n - 1] = z;
                e[n] = -z;
            }
            n = n - 2;
            iter = 0;
        } else { /* no convergence yet */
            x = A_(n, n);
            y = 0.0;
            w = 0.0;
            if (l < n) {
                y = A_(n - 1, n - 1);
                w = A_(n, n - 1) * A_(n - 1, n);
            }
            if (iter == 10) { /* Wilkinson's original ad hoc shift */
                exshift += x;
                for (int i = low; i <= n; i++) A_(i, i) -= x;
                s = fabs(A_(n, n - 1)) + fabs(A_(n - 1, n - 2));
                x = y = 0.75 * s;
                w = -0.4375 * s * s;
            }
            if (iter == 30) { /* MATLAB's ad hoc shift */
                s = (y - x) / 2.0;
                s = s * s + w;
                if (s > 0) {
                    s = sqrt(s);
                    if (y < x) s = -s;
                    s = x - w / ((y - x) / 2.0 + s);
                    for (int i = low; i <= n; i++) A_(i, i) -= s;
                    exshift += s;
                    x = y = w = 0.964;
                }
            }
            iter = iter + 1;
            if (++total_iter > 60L * nn + 1000) return 1;
            int m = n - 2;
            while (m >= l) {
                z = A_(m, m);
                r = x - z;
                s = y - z;
                p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
                q = A_(m + 1, m + 1) - z - r - s;
                r = A_(m + 2, m + 1);
                s = fabs(p) + fabs(q) + fabs(r);
                p = p / s;
                q = q / s;
                r = r / s;
                if (m == l) break;
                if (fabs(A_(m, m - 1)) * (fabs(q) + fabs(r)) <
                    eps * (fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) + fabs(A_(m + 1, m + 1)))))
                    break;
                m--;
            }
            for (int i = m + 2; i <= n; i++) {
                A_(i, i - 2) = 0.0;
                if (i > m + 2) A_(i, i - 3) = 0.0;
            }
            for (int k = m; k <= n - 1; k++) { /* double QR step */
                int notlast = (k != n - 1);
                if (k != m) {
                    p = A_(k, k - 1);
                    q = A_(k + 1, k - 1);
                    r = notlast ? A_(k + 2, k - 1) : 0.0;
                    x = fabs(p) + fabs(q) + fabs(r);
                    if (x == 0.0) continue;
                    p = p / x;
                    q = q / x;
                    r = r / x;
                }
                s = sqrt(p * p + q * q + r * r);
                if (p < 0) s = -s;
                if (s != 0) {
                    if (k != m) A_(k, k - 1) = -s * x;
                    else if (l != m) A_(k, k - 1) = -A_(k, k - 1);
                    p = p + s;
                    x = p / s;
                    y = q / s;
                    z = r / s;
                    q = q / p;
                    r = r / p;
                    for (int j = k; j < nn; j++) {
                        p = A_(k, j) + q * A_(k + 1, j);
                        if (notlast) {
                            p = p + r * A_(k + 2, j);
                            A_(k + 2, j) = A_(k + 2, j) - p * z;
                        }
                        A_(k, j) = A_(k, j) - p * x;
                        A_(k + 1, j) = A_(k + 1, j) - p * y;
                    }
                    for (int i = 0; i <= imin(n, k + 3); i++) {
                        p = x * A_(i, k) + y * A_(i, k + 1);
                        if (notlast) {
                            p = p + z * A_(i, k + 2);
                            A_(i, k + 2) = A_(i, k + 2) - p * r;
                        }
                        A_(i, k) = A_(i, k) - p;
                        A_(i, k + 1) = A_(i, k + 1) - p * q;
                    }
                    for (int i = low; i <= high; i++) {
                        p = x * V_(i, k) + y * V_(i, k + 1);
                        if (notlast) {
                            p = p + z * V_(i, k + 2);
                            V_(i, k + 2) = V_(i, k + 2) - p * r;
                        }
                        V_(i, k) = V_(i, k) - p;
                        V_(i, k + 1) = V_(i, k + 1) - p * q;
                    }
                }
            }
        }
    }

    if (norm == 0.0) return 0;
    /* back-substitute for the vectors of the upper triangular form */
    for (n = nn - 1; n >= 0; n--) {
        p = d[n];
        q = e[n];
        if (q == 0) { /* real vector */
            int l = n;
            A_(n, n) = 1.0;
            for (int i = n - 1; i >= 0; i--) {
                w = A_(i, i) - p;
                r = 0.0;
                for (int j = l; j <= n; j++) r = r + A_(i, j) * A_(j, n);
                if (e[i] < 0.0) {
                    z = w;
                    s = r;
                } else {
                    l = i;
                    if (e[i] == 0.0) {
                        if (w != 0.0) A_(i, n) = -r / w;
                        else A_(i, n) = -r / (eps * norm);
                    } else {
                        x = A_(i, i + 1);
                        y = A_(i + 1, i);
                        q = (d[i] - p) * (d[i] - p) + e[i] * e[i];
                        t = (x * s - z * r) / q;
                        A_(i, n) = t;
                        if (fabs(x) > fabs(z)) A_(i + 1, n) = (-r - w * t) / x;
                        else A_(i + 1, n) = (-s - y * t) / z;
                    }
                    t = fabs(A_(i, n));
                    if ((eps * t) * t > 1)
                        for (int j = i; j <= n; j++) A_(j, n) = A_(j, n) / t;
                }
            }
        } else if (q < 0) { /* complex vector, last component imaginary */
            int l = n - 1;
            if (fabs(A_(n, n - 1)) > fabs(A_(n - 1, n))) {
                A_(n - 1, n - 1) = q / A_(n, n - 1);
                A_(n - 1, n) = -(A_(n, n) - p) / A_(n, n - 1);
            } else {
                cdiv(0.0, -A_(n - 1, n), A_(n - 1, n - 1) - p, q, &A_(n - 1, n - 1), &A_(n - 1, n));
            }
            A_(n, n - 1) = 0.0;
            A_(n, n) = 1.0;
            for (int i = n - 2; i >= 0; i--) {
                double ra = 0.0, sa = 0.0, vr, vi;
                for (int j = l; j <= n; j++) {
                    ra = ra + A_(i, j) * A_(j, n - 1);
                    sa = sa + A_(i, j) * A_(j, n);
                }
                w = A_(i, i) - p;
                if (e[i] < 0.0) {
                    z = w;
                    r = ra;
                    s = sa;
                } else {
                    l = i;
                    if (e[i] == 0) {
                        cdiv(-ra, -sa, w, q, &A_(i, n - 1), &A_(i, n));
                    } else {
                        x = A_(i, i + 1);
                        y = A_(i + 1, i);
                        vr = (d[i] - p) * (d[i] - p) + e[i] * e[i] - q * q;
                        vi = (d[i] - p) * 2.0 * q;
                        if (vr == 0.0 && vi == 0.0)
                            vr = eps * norm * (fabs(w) + fabs(q) + fabs(x) + fabs(y) + fabs(z));
                        cdiv(x * r - z * ra + q * sa, x * s - z * sa - q * ra, vr, vi, &A_(i, n - 1),
                             &A_(i, n));
                        if (fabs(x) > (fabs(z) + fabs(q))) {
                            A_(i + 1, n - 1) = (-ra - w * A_(i, n - 1) + q * A_(i, n)) / x;
                            A_(i + 1, n) = (-sa - w * A_(i, n) - q * A_(i, n - 1)) / x;
                        } else {
                            cdiv(-r - y * A_(i, n - 1), -s - y * A_(i, n), z, q, &A_(i + 1, n - 1),
                                 &A_(i + 1, n));
                        }
                    }
                    t = fmax(fabs(A_(i, n - 1)), fabs(A_(i, n)));
                    if ((eps * t) * t > 1)
                        for (int j = i; j <= n; j++) {
                            A_(j, n - 1) = A_(j, n - 1) / t;
                            A_(j, n) = A_(j, n) / t;
                        }
                }
            }
        }
    }
    /* vectors of isolated roots */
    for (int i = 0; i < nn; i++)
        if (i < low || i > high)
            for (int j = i; j < nn; j++) V_(i, j) = A_(i, j);
    /* back transformation */
    for (int j = nn - 1; j >= low; j--)
        for (int i = low; i <= high; i++) {
            z = 0.0;
            for (int k = low; k <= imin(j, high); k++) z = z + V_(i, k) * A_(k, j);
            V_(i, j) = z;
        }
    return 0;
#undef A_
#undef V_
}

int hdo_eig_general(int32_t n, double *a, double *wr, double *wi, double *v)
{
    if (n <= 0) return 0;
    double *scale = (double *)malloc((size_t)n * sizeof(double));
    double *ort = (double *)calloc((size_t)n, sizeof(double));
    if (!scale || !ort) { free(scale); free(ort); return HDO_ERR_ALLOC; }
    int low, igh;
    eig_balance(n, a, &low, &igh, scale);
    eig_orthes(n, low, igh, a, v, ort);
    int rc = eig_hqr2(n, low, igh, a, v, wr, wi);
    if (rc == 0) {
        /* balbak (scaling is skipped when igh == low, as in EISPACK) */
        for (int i = low; i <= igh && igh != low; i++) {
            double s = scale[i];
            for (int j = 0; j < n; j++) v[(size_t)i * n + j] *= s;
        }
        for (int ii = 0; ii < n; ii++) {
            int i = ii;
            if (i >= low && i <= igh) continue;
            if (i < low) i = low - 1 - ii;
            int k = (int)scale[i];
            if (k == i) continue;
            for (int j = 0; j < n; j++) {
                double s = v[(size_t)i * n + j];
                v[(size_t)i * n + j] = v[(size_t)k * n + j];
                v[(size_t)k * n + j] = s;
            }
        }
        /* dgeev: real eigenvectors scaled to unit Euclidean norm */
        for (int j = 0; j < n; j++) {
            if (wi[j] != 0.0) continue;
            double nrm = 0.0, amax = 0.0;
            for (int i = 0; i < n; i++) amax = fmax(amax, fabs(v[(size_t)i * n + j]));
            if (amax == 0.0 || !isfinite(amax)) continue;
            for (int i = 0; i < n; i++) {
                double tq = v[(size_t)i * n + j] / amax;
                nrm += tq * tq;
            }
            nrm = amax * sqrt(nrm);
            double scl = 1.0 / nrm;
            for (int i = 0; i < n; i++) v[(size_t)i * n + j] *= scl;
        }
    }
    free(scale);
    free(ort);
    return rc ? HDO_ERR_EIGEN : HDO_OK;
}

/* ------------------------------------------------------------------------- */
/* Modal (model.go:719-851) and the buckling block (model.go:491-601)           */
/* ------------------------------------------------------------------------- */

typedef struct {
    double key;
    int idx;
} sort_item;

static void stable_sort_items(sort_item *it, int n)
{ /* insertion sort: stable, n is small */
    for (int i = 1; i < n; i++) {
        sort_item t = it[i];
        int j = i - 1;
        while (j >= 0 && it[j].key > t.key) { it[j + 1] = it[j]; j--; }
        it[j + 1] = t;
    }
}

/* shared tail: eigen of dense h (dof x dof, row-major), collect real non-zero eigenvalues.
 * mode 0: modal Hz = 1/(sqrt(|val|)*2*pi) (model.go:830); mode 1: buckling factor 1/|val| (:578) */
static int eigen_collect(int dof, int nP, double *h, int mode, int amount, int maxOut, int32_t *nOut,
                         double *keyOut, double *shapes)
{
    double *wr = (double *)malloc((size_t)dof * sizeof(double));
    double *wi = (double *)malloc((size_t)dof * sizeof(double));
    double *v = (double *)malloc((size_t)dof * dof * sizeof(double));
    sort_item *items = (sort_item *)malloc((size_t)dof * sizeof(sort_item));
    if (!wr || !wi || !v || !items) { free(wr); free(wi); free(v); free(items); return HDO_ERR_ALLOC; }
    int rc = hdo_eig_general(dof, h, wr, wi, v);
    int cnt = 0;
    if (rc == HDO_OK) {
        for (int i = 0; i < dof; i++) {
            if (fabs(wi[i]) > 0 || wr[i] == 0) continue; /* :818 / :566 */
            double val = fabs(wr[i]);
            if (val == 0.0) continue;
            double key = (mode == 0) ? 1. / (sqrt(val) * 2.0 * M_PI) : 1. / val;
            items[cnt].key = key;
            items[cnt].idx = i;
            cnt++;
        }
        stable_sort_items(items, cnt); /* :846 / :593 */
        if (mode == 1 && cnt > amount) cnt = amount; /* :598 */
        int out = cnt < maxOut ? cnt : maxOut;
        for (int r = 0; r < out; r++) {
            keyOut[r] = items[r].key;
            int col = items[r].idx;
            for (int p = 0; p < nP; p++)
                for (int c = 0; c < 3; c++)
                    shapes[((size_t)r * nP + p) * 3 + c] = v[(size_t)(3 * p + c) * dof + col];
        }
        *nOut = cnt;
    }
    free(wr);
    free(wi);
    free(v);
    free(items);
    return rc;
}

int hdo_modal(const hdo_model *m, int32_t nMass, const int64_t *massN, const double *mass, int32_t maxOut,
              int32_t *nOut, double *hz, double *shapes)
{
    csc K;
    lu_t F;
    int nIgnore = 0, *ignore = NULL;
    *nOut = 0;
    int rc = get_k(m, &K, &F, &nIgnore, &ignore); /* :758 */
    int dof = 3 * m->nP;
    double *h = NULL, *mv = NULL, *hh = NULL;
    if (rc == HDO_OK) {
        h = (double *)calloc((size_t)dof * dof, sizeof(double));
        mv = (double *)malloc((size_t)dof * sizeof(double));
        hh = (double *)malloc((size_t)dof * sizeof(double));
        if (!h || !mv || !hh) rc = HDO_ERR_ALLOC;
    }
    if (rc == HDO_OK) {
        for (int col = 0; col < dof; col++) { /* :772-805 */
            for (int i = 0; i < dof; i++) mv[i] = 0.0;
            int isZero = 1;
            for (int dir = 0; dir < 2; dir++)
                for (int q = 0; q < nMass; q++) {
                    int64_t index = massN[q] * 3 + dir;
                    if (index != col) continue;
                    isZero = 0;
                    mv[col] += mass[q] / GRAVITY;
                }
            if (isZero) continue;
            lu_solve(&F, mv, hh);
            for (int i = 0; i < dof; i++) h[(size_t)i * dof + col] = hh[i];
        }
        rc = eigen_collect(dof, m->nP, h, 0, 0, maxOut, nOut, hz, shapes);
    }
    free(h);
    free(mv);
    free(hh);
    csc_free(&K);
    lu_free(&F);
    free(ignore);
    return rc;
}

int hdo_buckling(const hdo_model *m, const double *beamForces, int32_t amount, int32_t maxOut,
                 int32_t *nOut, double *factors, double *shapes)
{
    csc K, G;
    lu_t F;
    int nIgnore = 0, *ignore = NULL, nIgG = 0, *ignG = NULL;
    *nOut = 0;
    memset(&G, 0, sizeof(G));
    int rc = get_k(m, &K, &F, &nIgnore, &ignore);
    int dof = 3 * m->nP;
    double *M = NULL, *h = NULL, *MS = NULL, *hh = NULL;
    if (rc == HDO_OK) rc = assembly_k(m, 1, beamForces, &G, &nIgG, &ignG); /* :494 */
    if (rc == HDO_OK) {
        M = (double *)calloc((size_t)dof * dof, sizeof(double));
        h = (double *)calloc((size_t)dof * dof, sizeof(double));
        MS = (double *)malloc((size_t)dof * sizeof(double));
        hh = (double *)malloc((size_t)dof * sizeof(double));
        if (!M || !h || !MS || !hh) rc = HDO_ERR_ALLOC;
    }
    if (rc == HDO_OK) {
        for (int j = 0; j < dof; j++) /* :510-516 */
            for (int p = G.p[j]; p < G.p[j + 1]; p++) M[(size_t)G.i[p] * dof + j] += G.x[p];
        for (int col = 0; col < dof; col++) { /* :526-552 */
            int isZero = 1;
            for (int i = 0; i < dof; i++) {
                if (M[(size_t)i * dof + col] != 0) isZero = 0;
                MS[i] = M[(size_t)i * dof + col];
            }
            if (isZero) continue;
            lu_solve(&F, MS, hh);
            for (int i = 0; i < dof; i++) h[(size_t)i * dof + col] = hh[i];
        }
        rc = eigen_collect(dof, m->nP, h, 1, amount, maxOut, nOut, factors, shapes);
    }
    free(M);
    free(h);
    free(MS);
    free(hh);
    csc_free(&K);
    csc_free(&G);
    lu_free(&F);
    free(ignore);
    free(ignG);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* validation (check.go)                                                       */
/* ------------------------------------------------------------------------- */

static int bad_float(double f) { return isnan(f) || isinf(f); }

/* check.go:47-109 */
static int graph_check(const hdo_model *m)
{
    int nB = m->nB, nP = m->nP;
    if (nB == 0) return HDO_OK;
    uint8_t *mark = (uint8_t *)calloc((size_t)(nP > 0 ? nP : 1), 1);
    uint8_t *beamMark = (uint8_t *)calloc((size_t)nB, 1);
    int *present = (int *)malloc((size_t)nB * sizeof(int));
    int *next = (int *)malloc((size_t)nB * sizeof(int));
    uint8_t *queued = (uint8_t *)calloc((size_t)nB, 1);
    if (!mark || !beamMark || !present || !next || !queued) {
        free(mark); free(beamMark); free(present); free(next); free(queued);
        return HDO_ERR_ALLOC;
    }
    int np = 0;
    present[np++] = 0;
    while (np > 0) {
        int nn = 0;
        memset(queued, 0, (size_t)nB);
        for (int q = 0; q < np; q++) {
            int pr = present[q];
            int64_t n0 = m->beamN[2 * pr], n1 = m->beamN[2 * pr + 1];
            mark[n0] = 1;
            mark[n1] = 1;
            for (int idx = 0; idx < nB; idx++) {
                int add = 0;
                for (int i = 0; i < 2; i++)
                    if (m->beamN[2 * idx + i] == n0 || m->beamN[2 * idx + i] == n1) add = 1;
                if (add && !queued[idx]) { queued[idx] = 1; next[nn++] = idx; }
            }
        }
        for (int q = 0; q < np; q++) beamMark[present[q]] = 1;
        np = 0;
        for (int q = 0; q < nn; q++)
            if (!beamMark[next[q]]) present[np++] = next[q];
    }
    int unmarked = 0;
    for (int i = 0; i < nP; i++)
        if (!mark[i]) unmarked++;
    free(mark); free(beamMark); free(present); free(next); free(queued);
    return unmarked ? HDO_ERR_INPUT : HDO_OK; /* ErrorAmountModel */
}

int hdo_check_model(const hdo_model *m, int32_t nPins, int32_t nSupports)
{
    int err = 0;
    for (int i = 0; i < m->nP; i++) /* checkPoints :223-244 */
        for (int j = 0; j < 2; j++)
            if (bad_float(m->points[2 * i + j])) err = 1;
    if (m->nP > 0 && m->nB == 0) return HDO_ERR_INPUT; /* :266 ErrorBeamAmount */
    for (int b = 0; b < m->nB; b++) { /* checkBeams :270-325 */
        int idxOk = 1;
        for (int j = 0; j < 2; j++) {
            int64_t n = m->beamN[2 * b + j];
            if (n < 0 || n >= m->nP) { err = 1; idxOk = 0; }
        }
        if (m->beamN[2 * b] == m->beamN[2 * b + 1]) err = 1;
        if (idxOk) {
            double xy[4];
            beam_xy(m, b, xy);
            if (hdo_distance(xy) <= 0) err = 1;
        } /* else: distance() panics -> NaN -> `NaN <= 0` false, index error already recorded */
        for (int c = 0; c < 3; c++) {
            double v = m->sections[3 * b + c];
            if (bad_float(v) || v < 0 || v == 0) err = 1;
        }
    }
    if (m->nP != nSupports) err = 1; /* checkSupports :332-337 */
    if (m->nB != nPins) err = 1;     /* checkPins :353 */
    for (int b = 0; b < nPins && b < m->nB; b++) { /* :357-391 */
        const uint8_t *pn = m->pins + 6 * b;
        int cnt = 0;
        for (int j = 0; j < 6; j++) cnt += pn[j] ? 1 : 0;
        if (cnt > 4) err = 1;
        if (pn[0] && pn[3]) err = 1;
        if (pn[2] && pn[5] && !(pn[1] && pn[4])) err = 1;
    }
    if (err) return HDO_ERR_INPUT;
    return graph_check(m); /* only when everything else passed, :31-36 */
}

int hdo_check_loads(int32_t nP, int32_t nLoad, const int64_t *loadN, const double *forces)
{
    int err = 0;
    for (int j = 0; j < nLoad; j++) { /* check.go:131-155 */
        if (loadN[j] < 0 || loadN[j] >= nP) err = 1;
        for (int k = 0; k < 3; k++)
            if (bad_float(forces[3 * j + k])) err = 1;
    }
    return err ? HDO_ERR_INPUT : HDO_OK;
}

int hdo_check_modal(const hdo_model *m, int32_t nMass, const int64_t *massN, const double *mass)
{
    int err = 0;
    for (int j = 0; j < nMass; j++) { /* check.go:413-447 */
        if (massN[j] < 0 || massN[j] >= m->nP) err = 1;
        double v = mass[j];
        if (bad_float(v) || v < 0 || v == 0) err = 1;
        if (massN[j] >= 0 && massN[j] < m->nP) {
            const uint8_t *s = m->supports + 3 * massN[j];
            if (s[0] || s[1] || s[2]) err = 1;
        }
    }
    return err ? HDO_ERR_INPUT : HDO_OK;
}

/* ------------------------------------------------------------------------- */
/* batch drivers for the CPU baseline: independent hd.Run calls over models     */
/* ------------------------------------------------------------------------- */

int hdo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int hdo_static_batch(const hdo_model *topo, int32_t nModels, const double *sections, int32_t nLC,
                     const double *loads, double *disp, double *forces, double *react, int32_t *status,
                     int32_t nThreads)
{
    int nP = topo->nP, nB = topo->nB, dof = 3 * nP;
#ifdef _OPENMP
    if (nThreads <= 0) nThreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nThreads)
#endif
    for (int mi = 0; mi < nModels; mi++) {
        hdo_model m = *topo;
        m.sections = sections + (size_t)mi * nB * 3;
        int32_t fc;
        int rc = hdo_linear_static(&m, nLC, loads + (size_t)mi * nLC * dof,
                                   disp ? disp + (size_t)mi * nLC * dof : NULL,
                                   forces ? forces + (size_t)mi * nLC * nB * 6 : NULL,
                                   react ? react + (size_t)mi * nLC * dof : NULL, &fc);
        if (status) status[mi] = rc;
    }
    (void)nThreads;
    return HDO_OK;
}

int hdo_modal_batch(const hdo_model *topo, int32_t nModels, const double *sections, int32_t nMass,
                    const int64_t *massN, const double *mass, int32_t maxOut, int32_t *nOut, double *hz,
                    double *shapes, int32_t *status, int32_t nThreads)
{
    int nP = topo->nP, nB = topo->nB;
#ifdef _OPENMP
    if (nThreads <= 0) nThreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nThreads)
#endif
    for (int mi = 0; mi < nModels; mi++) {
        hdo_model m = *topo;
        m.sections = sections + (size_t)mi * nB * 3;
        int rc = hdo_modal(&m, nMass, massN, mass + (size_t)mi * nMass, maxOut, nOut + mi,
                           hz + (size_t)mi * maxOut, shapes + (size_t)mi * maxOut * nP * 3);
        if (status) status[mi] = rc;
    }
    (void)nThreads;
    return HDO_OK;
}
