/*
 * hd_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the linear-analysis hot path of Konstantin8105/hd
 * (feBeam2d.go, model.go, check.go) and of the published algorithms of its
 * un-vendored dependencies (Konstantin8105/sparse = CSparse port, gonum/mat
 * Dgemm + Dgeev).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this library.  The product (libhdb200.so) never does.
 *
 * Parity pin: checked against the reference's 14 golden reports
 * (example/testdata/ files, 6 significant digits) via tests/golden/golden.json and
 * against LAPACK (SciPy) in tests/test_oracle_*.py.  Above 6 digits and at the
 * BASELINE.json sizes the reference itself holds no vectors: "parity unpinned"
 * beyond the goldens.
 */
#ifndef HD_ORACLE_H
#define HD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened hd.Model (model.go:18-77). pins may not be NULL here: the caller
 * applies model.go:344-346 (empty Pins => all false) before calling. */
typedef struct hdo_model {
    int32_t nP;              /* len(Points)                     */
    int32_t nB;              /* len(Beams)                      */
    const double *points;    /* [nP][2]                         */
    const int64_t *beamN;    /* [nB][2]  BeamProp.N             */
    const double *sections;  /* [nB][3]  BeamProp.A, .J, .E     */
    const uint8_t *pins;     /* [nB][6]                         */
    const uint8_t *supports; /* [nP][3]                         */
} hdo_model;

/* return codes */
enum {
    HDO_OK = 0,
    HDO_ERR_INPUT = 1,      /* check.go rejects / NaN,Inf in load or matrix entry */
    HDO_ERR_SINGULAR = 2,   /* "LU error factorization" model.go:330            */
    HDO_ERR_IGNORE_LOAD = 3,/* load on an `ignore` DOF, model.go:424-437        */
    HDO_ERR_EIGEN = 4,      /* "Eigen factorization is not ok" model.go:558,810 */
    HDO_ERR_ALLOC = 5
};

/* ---- element routines (feBeam2d.go) ---- */
/* xy = {x_start, y_start, x_end, y_end} */
double hdo_distance(const double xy[4]);                                   /* feBeam2d.go:11-19  */
void hdo_stiff_beam2d(const double xy[4], double A, double J, double E,
                      const uint8_t pins[6], double k[36]);                /* feBeam2d.go:22-140 */
void hdo_trans_beam2d(const double xy[4], double t[36]);                   /* feBeam2d.go:142-172*/
void hdo_geometric_beam2d(const double xy[4], const uint8_t pins[6],
                          const double beamForces[6], double g[36]);       /* feBeam2d.go:175-275*/
/* kg = (T^T * k) * T with gonum Dgemm semantics, model.go:619-620 */
void hdo_rotate_to_global(const double k[36], const double t[36], double kg[36]);

/* ---- assembly (model.go:607-681) ----
 * Returns the assembled K both in the reference's internal CSC form
 * (first-appearance order inside each column) and, for pattern comparison,
 * as canonical CSR with ascending columns.  Buffers are malloc'ed; free with
 * hdo_free().  which: 0 = stiffness, 1 = geometric (needs beamForces[nB][6]). */
int hdo_assemble(const hdo_model *m, int which, const double *beamForces,
                 int32_t *dof, int32_t *nnz,
                 int32_t **cscColptr, int32_t **cscRowind, double **cscVal,
                 int32_t **csrRowptr, int32_t **csrColind, double **csrVal,
                 int32_t *nIgnore, int32_t **ignore);
/* eliminated set = sort-unique(ignore ++ addSupport()), model.go:326-328, 700-710 */
int hdo_eliminated(const hdo_model *m, int32_t *nElim, int32_t **elim,
                   int32_t *nIgnore, int32_t **ignore);
void hdo_free(void *p);

/* ---- LinearStatic (model.go:358-490) ----
 * loads[nLC][dof] are the dense vectors produced by assemblyNodeLoad
 * (model.go:683-698).  Outputs: disp[nLC][nP][3], forces[nLC][nB][6],
 * react[nLC][nP][3].  On HDO_ERR_IGNORE_LOAD *failCase is the offending case;
 * cases before it are filled (the reference keeps them), later ones untouched. */
int hdo_linear_static(const hdo_model *m, int32_t nLC, const double *loads,
                      double *disp, double *forces, double *react, int32_t *failCase);

/* ---- linear buckling block (model.go:491-601) for one solved load case ----
 * beamForces[nB][6] from hdo_linear_static. factors[maxOut], shapes[maxOut][nP][3]
 * sorted ascending by factor, truncated to `amount` (and to maxOut). */
int hdo_buckling(const hdo_model *m, const double *beamForces, int32_t amount,
                 int32_t maxOut, int32_t *nOut, double *factors, double *shapes);

/* ---- Modal (model.go:719-851) ----
 * hz[maxOut], shapes[maxOut][nP][3]; ascending by Hz (stable). */
int hdo_modal(const hdo_model *m, int32_t nMass, const int64_t *massN, const double *mass,
              int32_t maxOut, int32_t *nOut, double *hz, double *shapes);

/* ---- validation (check.go) ---- returns HDO_OK or HDO_ERR_INPUT.
 * nPins/nSupports are the slice lengths as the caller holds them (check.go:333,353). */
int hdo_check_model(const hdo_model *m, int32_t nPins, int32_t nSupports);          /* check.go:12-109  */
int hdo_check_loads(int32_t nP, int32_t nLoad, const int64_t *loadN, const double *forces); /* check.go:127-163 */
int hdo_check_modal(const hdo_model *m, int32_t nMass, const int64_t *massN, const double *mass); /* check.go:411-452 */

/* ---- dense non-symmetric eigen solver (stand-in for gonum mat.Eigen = Dgeev) ----
 * a[n][n] row-major (destroyed). wr, wi[n]; v[n][n] row-major, column j is the
 * eigenvector of eigenvalue j (real eigenvalues: unit 2-norm). returns 0 on success. */
int hdo_eig_general(int32_t n, double *a, double *wr, double *wi, double *v);

/* ---- batch driver for the CPU baseline (independent models, OpenMP) ----
 * Same topology, per-model sections[nModels][nB][3] and loads[nModels][nLC][dof];
 * any output pointer may be NULL. status[nModels]. nThreads<=0 => all cores. */
int hdo_static_batch(const hdo_model *topo, int32_t nModels, const double *sections,
                     int32_t nLC, const double *loads,
                     double *disp, double *forces, double *react, int32_t *status,
                     int32_t nThreads);
int hdo_modal_batch(const hdo_model *topo, int32_t nModels, const double *sections,
                    int32_t nMass, const int64_t *massN, const double *mass /*[nModels][nMass]*/,
                    int32_t maxOut, int32_t *nOut /*[nModels]*/, double *hz, double *shapes,
                    int32_t *status, int32_t nThreads);
int hdo_max_threads(void);
/* 1 (default): reactions by the reference's O(nSup*nnz) Fkeep scan (model.go:481-486);
 * 0: identical sums (columns ascending) through per-row lists. Results are bit-identical. */
void hdo_set_literal_scans(int on);

#ifdef __cplusplus
}
#endif
#endif
