/*
 * hdb200.h — C ABI of libhdb200.so: the B200-native batched replacement of the
 * linear-analysis hot path of Konstantin8105/hd.
 *
 * Boundary (reference side, paths relative to the reference repository root):
 *   hd.Run(out, *Model, []LoadCase, []ModalCase)            model.go:926-969
 *   hd.LinearStatic(out, *Model, ...*LoadCase)              model.go:358-605
 *   hd.Modal(out, *Model, *ModalCase)                       model.go:719-851
 * A Go shim keeps those three signatures and the exported structs, keeps check.go and the
 * text reports in Go, flattens the structs and makes ONE cgo call per phase into the entry
 * points below (see INTEGRATION.md for the cgo binding).  Everything the reference delegates
 * to Konstantin8105/sparse and gonum/mat on this path is computed on the GPU; there is no CPU
 * fallback: without a CUDA device every compute entry returns HDB_ERR_NO_DEVICE.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are dense, row-major, C order;
 *   - data pointers of the *_batch calls may be HOST (pageable or pinned, see hdb_host_alloc)
 *     or DEVICE memory of the topology's device — detected per pointer; host data is staged
 *     through pinned buffers in chunks with copies overlapped with compute;
 *   - no Go pointer is retained after a call returns; calls are synchronous unless a stream is
 *     passed to a *_dev variant;
 *   - return value: 0 ok, <0 API / CUDA error (hdb_error_string). Per-model numerical outcomes
 *     are reported in status[model] (HDB_STATUS_*), so one bad model never poisons a batch;
 *   - thread safety: a topology handle is single-owner (one call at a time per handle);
 *     distinct handles may be used from distinct threads. Every entry sets its device.
 */
#ifndef HDB200_H
#define HDB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HDB_VERSION 100

/* ---- return codes ---- */
enum {
    HDB_OK = 0,
    HDB_ERR_ARG = -1,         /* NULL / negative size / inconsistent argument          */
    HDB_ERR_NO_DEVICE = -2,   /* no CUDA device: the product has no CPU path           */
    HDB_ERR_CUDA = -3,        /* CUDA runtime error (hdb_last_error_detail)            */
    HDB_ERR_NOMEM = -4,
    HDB_ERR_UNSUPPORTED = -5, /* too large for a shared-memory resident kernel: the modal eigen solve above ~160 mass DOFs
                               * (DESIGN.md section 8), LinearStatic only above ~28 k padded DOFs per load case; also
                               * hdb_multi_gather_summary without NCCL */
    HDB_ERR_INPUT = -6,       /* topology rejected (what check.go rejects for shared data) */
    HDB_ERR_SMALL_BUFFER = -7
};

/* ---- per-model status (int32) ---- */
enum {
    HDB_STATUS_OK = 0,
    HDB_STATUS_NONFINITE = 1,   /* NaN/Inf/non-positive section or NaN/Inf matrix entry or load:
                                   check.go:301-324, model.go:692-696                          */
    HDB_STATUS_SINGULAR = 2,    /* exactly zero or non-finite pivot: "LU error factorization",
                                   model.go:326-331 (round-off sized pivots pass, as in the reference) */
    HDB_STATUS_IGNORE_LOAD = 3, /* non-zero load on an `ignore` DOF, model.go:424-437; the
                                   offending case index is written to fail_case[model]          */
    HDB_STATUS_EIGEN = 4,       /* eigen iteration did not converge, model.go:809-811           */
    HDB_STATUS_PATTERN = 5      /* this model has a non-zero element entry at a matrix position that
                                   is absent from the pattern of the topology's representative
                                   model (exact zeros dropped, model.go:645-649): its sparsity
                                   pattern differs, solve it under its own topology handle     */
};

/* Go hd.BeamProp memory layout (model.go:58-77): N [2]int, A, J, E float64 = 40 bytes. */
typedef struct hdb_beam {
    int64_t n0, n1;
    double A, J, E;
} hdb_beam;

/* Go hd.LoadNode (model.go:304-315): N int, Forces [3]float64 = 32 bytes. */
typedef struct hdb_load_node {
    int64_t n;
    double forces[3];
} hdb_load_node;

/* Go hd.ModalMass (model.go:294-301): N int, Mass float64 = 16 bytes. */
typedef struct hdb_modal_mass {
    int64_t n;
    double mass;
} hdb_modal_mass;

typedef struct hdb_options {
    int32_t device;    /* CUDA device ordinal; -1 = current device                         */
    int32_t ordering;  /* 0 natural (identity permutation: what the reference's LU uses),
                          1 reverse Cuthill-McKee on the free-DOF node graph,
                          2 automatic: whichever of the two gives the smaller factor (multi-storey
                            frames need it: natural order fills a band as wide as a storey)      */
    int32_t reserved[6];
} hdb_options;

typedef struct hdb_topology hdb_topology; /* opaque */

typedef struct hdb_topology_info {
    int32_t nP, nB, dof;
    int32_t nFree, nElim, nIgnore, nSupported;
    int32_t nnzK;      /* entries of the assembled K after model.go:645-652 (zeros dropped, duplicates summed) */
    int32_t nnzKff;    /* of which both row and column are free                                               */
    int32_t nnzL;      /* strictly-lower entries of the LDL^T factor incl. fill                               */
    int32_t nnzR;      /* entries of the supported rows x free columns block used for reactions               */
    int32_t nUpdates;  /* multiply-subtract operations of one numeric factorization                           */
    int32_t device;
    int32_t reserved[3];
} hdb_topology_info;

/* integer structures, for bit-exact comparison against the reference / oracle */
enum {
    HDB_Q_ELIMINATED = 0, /* int32[nElim]   sort-unique(ignore ++ supports), model.go:326-328, 700-710 */
    HDB_Q_IGNORE = 1,     /* int32[nIgnore] DOFs without a diagonal entry, model.go:657-678            */
    HDB_Q_CSR_ROWPTR = 2, /* int32[dof+1]   pattern of K, canonical CSR, ascending columns             */
    HDB_Q_CSR_COLIND = 3, /* int32[nnzK]                                                               */
    HDB_Q_SCATTER = 4,    /* int32[nB*36]   triplet (beam, p1,p2,r1,r2 order of model.go:622-635) ->
                                            CSR slot, -1 where the entry is exactly zero (dropped)     */
    HDB_Q_FREE_INDEX = 5, /* int32[dof]     reduced index of a DOF or -1 (eliminated)                  */
    HDB_Q_PERM = 6,       /* int32[nFree]   elimination order: position -> reduced index               */
    HDB_Q_L_COLPTR = 7,   /* int32[nFree+1] pattern of L (strictly lower, CSC, permuted numbering)     */
    HDB_Q_L_ROWIND = 8    /* int32[nnzL]                                                               */
};

/* ---- symbolic phase: replaces the structural half of getK/assemblyK (model.go:318-333, 607-681) ----
 * points[nP][2], beams[nB] (Go layout; A,J,E of this call are the REPRESENTATIVE sections that
 * define the value-dependent pattern), pins[nB][6] or NULL (= all rigid, model.go:344-346),
 * supports[nP][3].  Runs entirely on the host; needs no GPU until the first compute call. */
int hdb_topology_create(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams,
                        const uint8_t *pins, const uint8_t *supports, const hdb_options *opts,
                        hdb_topology **out);
void hdb_topology_destroy(hdb_topology *t);
int hdb_topology_get_info(const hdb_topology *t, hdb_topology_info *info);
/* copies structure `which` into buf (capacity in bytes); returns the element count or <0 */
int64_t hdb_topology_query(const hdb_topology *t, int32_t which, void *buf, int64_t capacity);

/* ---- input validation of check.go, for host shims that do not carry their own ----
 * Returns HDB_OK or HDB_ERR_INPUT; a human-readable list is written to msg (may be NULL).
 * nPins / nSupports are the lengths of the caller's slices (check.go:333, 353). */
int hdb_check_model(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams,
                    int32_t nPins, const uint8_t *pins, int32_t nSupports, const uint8_t *supports,
                    char *msg, int32_t msgLen);                                   /* check.go:12-109, 223-397 */
int hdb_check_loads(int32_t nP, int32_t nLoads, const hdb_load_node *loads,
                    char *msg, int32_t msgLen);                                   /* check.go:127-163 */
int hdb_check_modal(int32_t nP, int32_t nSupports, const uint8_t *supports, int32_t nMasses,
                    const hdb_modal_mass *masses, char *msg, int32_t msgLen);     /* check.go:411-452 */
/* assemblyNodeLoad (model.go:683-698): LoadNodes -> dense p[dof], accumulated in order. */
int hdb_assemble_node_loads(int32_t nP, int32_t nLoads, const hdb_load_node *loads, double *p);

/* ---- LinearStatic for a batch of models sharing the topology (model.go:401-490) ----
 * sections[nModels][nB][3] (A,J,E per beam), loads[nModels][nLC][dof] (assemblyNodeLoad output),
 * disp[nModels][nLC][nP][3], forces[nModels][nLC][nB][6], react[nModels][nLC][nP][3] (zero on
 * unsupported DOFs), status[nModels], failCase[nModels] (may be NULL).  Any of disp / forces /
 * react may be NULL (not materialised). */
int hdb_static_batch(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC,
                     const double *loads, double *disp, double *forces, double *react,
                     int32_t *status, int32_t *failCase);
/* all pointers are device memory; enqueues on `stream` (a cudaStream_t, NULL = default stream)
 * and returns without synchronising.  Workspace lives in the handle and grows on demand. */
int hdb_static_batch_dev(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC,
                         const double *loads, double *disp, double *forces, double *react,
                         int32_t *status, int32_t *failCase, void *stream);

/* The same with per-model coordinates (shape sweeps): points[nModels][nP][2]; beams, pins and supports are the
 * topology's.  This is where the geometry part of check.go runs for a batch: a NaN / Inf coordinate
 * (checkPoints, check.go:223-245) or a beam of length <= 0 (checkBeams, check.go:294-300) gives
 * HDB_STATUS_NONFINITE for that model only; graphCheck and the index / pin / support checks (check.go:47-109,
 * 263-397) depend on the topology alone and ran once in hdb_topology_create.  A member that is axis-aligned in the
 * topology's own coordinates (exact zeros dropped from the pattern, model.go:645-649) and skewed in a batch member
 * gives HDB_STATUS_PATTERN, as for sections.  hdb_static_batch_geom takes host or device pointers and returns
 * when the results are in place; the _dev form enqueues on `stream` like hdb_static_batch_dev. */
int hdb_static_batch_geom(hdb_topology *t, int32_t nModels, const double *points, const double *sections,
                          int32_t nLC, const double *loads, double *disp, double *forces, double *react,
                          int32_t *status, int32_t *failCase);
int hdb_static_batch_geom_dev(hdb_topology *t, int32_t nModels, const double *points, const double *sections,
                              int32_t nLC, const double *loads, double *disp, double *forces, double *react,
                              int32_t *status, int32_t *failCase, void *stream);

/* assembled stiffness values only (tests): kval[nModels][nnzK] in HDB_Q_CSR_* order, bit-exact
 * against the reference's summation order (beam order, model.go:652). Host or device pointers. */
int hdb_assemble_batch(hdb_topology *t, int32_t nModels, const double *sections, double *kval,
                       int32_t *status);
/* element matrices only (tests): klocal[nModels][nB][36] (getStiffBeam2d, feBeam2d.go:22-140),
 * tr[nModels][nB][36] (getCoordTransStiffBeam2d, feBeam2d.go:142-172), kglobal[nModels][nB][36]
 * ((T^T k) T, model.go:619-620); any may be NULL. Host or device pointers. */
int hdb_element_batch(hdb_topology *t, int32_t nModels, const double *sections, double *klocal,
                      double *tr, double *kglobal);

/* ---- Modal for a batch (model.go:719-851) ----
 * masses: massN[nMass] (shared), mass[nModels][nMass] in newtons (divided by Gravity on device,
 * model.go:787).  Outputs ascending in Hz: nModes[nModels], hz[nModels][maxModes],
 * shapes[nModels][maxModes][nP][3] (unit 2-norm over 3*nP entries; sign is implementation-
 * defined as in LAPACK), status[nModels].  maxModes >= number of distinct free mass DOFs
 * (hdb_modal_mode_count).  hdb_modal_batch_dev synchronises `stream` once after uploading its mass tables (and when
 * its scratch has to grow); the kernels themselves are enqueued without a further synchronisation. */
int hdb_modal_mode_count(const hdb_topology *t, int32_t nMass, const int64_t *massN);
int hdb_modal_batch(hdb_topology *t, int32_t nModels, const double *sections, int32_t nMass,
                    const int64_t *massN, const double *mass, int32_t maxModes, int32_t *nModes,
                    double *hz, double *shapes, int32_t *status);
int hdb_modal_batch_dev(hdb_topology *t, int32_t nModels, const double *sections, int32_t nMass,
                        const int64_t *massN /*host*/, const double *mass, int32_t maxModes,
                        int32_t *nModes, double *hz, double *shapes, int32_t *status, void *stream);

/* ---- linear buckling: replaces the buckling block of LinearStatic (model.go:491-601) and getGeometricBeam2d
 * (feBeam2d.go:175-275) for one value of AmountLinearBuckling.
 * forces[nModels][nLC][nB][6] are the beam end forces of the load cases (output of hdb_static_batch).  Per
 * (model, case): nOut = number of buckling factors found (<= amount), factors[maxOut] ascending = 1/|eigenvalue| of
 * K^-1 G (model.go:578), shapes[maxOut][nP][3] (unit 2-norm, sign free) or NULL.  The reference takes the real
 * eigenvalues of the dense non-symmetric H = K^-1 G; here the same eigenvalues come from the symmetric
 * C^T G C with (K^-1)_ff = C C^T, which needs a stable (positive definite) model: otherwise HDB_STATUS_EIGEN.
 * Up to hdb_buckling_max_free() free DOFs the eigen solver keeps its work matrices in shared memory; larger models
 * (uniform_frame(128): 387) run the same code on a global scratch record per (model, case): slower, no size limit
 * other than memory. */
int hdb_buckling_max_free(hdb_topology *t);
int hdb_buckling_batch(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *forces,
                       int32_t amount, int32_t maxOut, int32_t *nOut, double *factors, double *shapes, int32_t *status);
int hdb_buckling_batch_dev(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *forces,
                           int32_t amount, int32_t maxOut, int32_t *nOut, double *factors, double *shapes,
                           int32_t *status, void *stream);

/* ---- load cases built on the device: replaces mod.LoadUniform (mod/uniform.go:46-211) followed by assemblyNodeLoad
 * (model.go:683-698) for sweeps in which every case of every model loads the same beams / nodes with different
 * intensities.  ubeam[nU], uproj[nU] (host): uniformly loaded beams and their projection flags; q[nModels][nLC][nU][2]
 * (device): intensities {qx, qy}; nodes[nN] (host), f[nModels][nLC][nN][3] (device, may be NULL when nN == 0): node
 * loads appended after the uniform loads.  loads[nModels][nLC][dof] (device) is bit-identical to flattening the
 * LoadNode list the reference would build (same order of additions).  Moves 16 bytes per loaded beam and case over
 * PCIe instead of 8*dof bytes per case.  Unlike the other *_dev entries this one SYNCHRONISES `stream` before it
 * returns (its per-call index tables are freed on return). */
int hdb_uniform_loads_dev(hdb_topology *t, int32_t nModels, int32_t nLC, int32_t nU, const int32_t *ubeam,
                          const uint8_t *uproj, const double *q, int32_t nN, const int64_t *nodes, const double *f,
                          double *loads, void *stream);

/* ---- several GPUs of one box from ONE process (a Go program is one process) ----
 * hdb_multi_create runs the symbolic phase once per device (devices[nDev] = CUDA ordinals, NULL = 0 .. nDev-1; opts->device
 * is ignored).  A batch is cut into contiguous model ranges, shard d = hdb_multi_shard(m, nModels, d) on device d; no
 * collective sits on the data path (models are independent, model.go:926 has no cross-model state).
 *   hdb_multi_static_batch       host arrays in the layout of hdb_static_batch; one host thread per device runs its range
 *                                through that device's chunked H2D / kernels / D2H pipeline and writes straight into the
 *                                caller's arrays.  Results are bit-identical to hdb_static_batch on one device.
 *   hdb_multi_static_batch_dev   shards already resident: entry d of every pointer array is the array of shard d on
 *                                device d (count[d] models); enqueues on streams[d] (NULL = default streams), no sync.
 *   hdb_multi_gather_summary     the only NCCL use: all-gathers per-shard REDUCED results over NVLink — summary[nDev][4] =
 *                                {failed models, largest |displacement| of the healthy ones, models, 0} — and returns them
 *                                on the host.  Full results stay sharded: gathering them costs ~20x the solve.  NCCL is
 *                                bound at run time (libnccl.so.2); HDB_ERR_UNSUPPORTED without it or with repeated devices. */
typedef struct hdb_multi hdb_multi; /* opaque */
int hdb_multi_create(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, const uint8_t *pins,
                     const uint8_t *supports, const hdb_options *opts, int32_t nDev, const int32_t *devices, hdb_multi **out);
void hdb_multi_destroy(hdb_multi *m);
int32_t hdb_multi_device_count(const hdb_multi *m);
int hdb_multi_shard(const hdb_multi *m, int32_t nModels, int32_t shard, int32_t *first, int32_t *count);
/* the same rule without a handle (no device needed): shard d of nShards gets a contiguous range, the first nModels % nShards
 * shards one model more; what a one-rank-per-GPU host (bench.py under torchrun) uses for its own range */
int hdb_shard_range(int32_t nModels, int32_t nShards, int32_t shard, int32_t *first, int32_t *count);
hdb_topology *hdb_multi_topology(hdb_multi *m, int32_t shard); /* the per-device handle (owned by m) */
int hdb_multi_static_batch(hdb_multi *m, int32_t nModels, const double *sections, int32_t nLC, const double *loads,
                           double *disp, double *forces, double *react, int32_t *status, int32_t *failCase);
int hdb_multi_static_batch_dev(hdb_multi *m, const int32_t *count, const double *const *sections, int32_t nLC,
                               const double *const *loads, double *const *disp, double *const *forces, double *const *react,
                               int32_t *const *status, int32_t *const *failCase, void *const *streams);
int hdb_multi_synchronize(hdb_multi *m);
int hdb_multi_gather_summary(hdb_multi *m, const int32_t *count, const int32_t *const *status, const double *const *disp,
                             int32_t nLC, double *summary);

/* ---- pinned host memory for asynchronous batches (cgo: wrap with unsafe.Slice) ---- */
int hdb_host_alloc(void **ptr, int64_t bytes);
int hdb_host_free(void *ptr);

/* ---- measurement: CUDA-event time of each kernel of the last *_dev / *_batch call ----
 * ms[0..n): static: {assemble, factor, solve (k_solve_pipe)}; modal: {assemble, factor, reduce, eigen,
 * shapes}.  Blocks until the events have completed.  Returns the count written.
 * hdb_set_profiling: 0 off, 1 time the last call only, 2 accumulate over calls until set again. */
int hdb_set_profiling(hdb_topology *t, int32_t on);
int hdb_last_kernel_ms(hdb_topology *t, float *ms, int32_t capacity);
/* number of kernels this library has launched in the process (the bench's gpu_launches). */
int64_t hdb_launch_count(void);

int hdb_device_count(void);
const char *hdb_error_string(int code);
const char *hdb_last_error_detail(void);
int hdb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HDB200_H */
