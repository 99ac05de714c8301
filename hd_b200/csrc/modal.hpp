// modal.hpp — host interface of the batched Modal path (kernels in kernels_modal.cu).
//
// Reference: Modal (model.go:719-851) builds H = K^-1 M column by column and takes all eigenpairs of the
// dense non-symmetric dof x dof H with gonum's Dgeev.  M is diagonal (lumped masses on X and Y of the
// mass points, model.go:780-789), so the same spectrum comes from the symmetric s x s matrix
//     B = M_S^1/2 (K_ff^-1)_SS M_S^1/2,   s = number of distinct free mass DOFs,
// and the shape of mode k is K_ff^-1[:,S] M_S^1/2 y_k (zeros on eliminated DOFs) scaled to unit 2-norm
// (SURVEY.md A.8).  Per model: s unit-load solves through the LinearStatic kernels, Householder
// tridiagonalisation + implicit QL in shared memory, one dense FP64 contraction for the shapes.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "device_topo.hpp"
#include "hdb_internal.hpp"

namespace hdb {

// distinct free mass DOFs (X and Y of every mass point that is not eliminated), ascending DOF;
// rows = their global DOF numbers, massOf (optional) lists for every row the masses that add to it.
int modal_mass_rows(const Symbolic &S, int nMass, const int64_t *massN, std::vector<int32_t> &dofs,
                    std::vector<std::vector<int32_t>> *massOf);

struct ModalLaunch {
    int nModels, s, dof, nMass;
    const int32_t *massDof;   // device [s] global DOF of mass row i
    const int32_t *mptr;      // device [s+1]
    const int32_t *midx;      // device [mptr[s]] indices into the mass list
    const double *mass;       // device [nModels][nMass] (newtons)
    double *msqrt;            // device [nModels][s] sqrt(sum mass / Gravity)
    double *unitLoads;        // device [nModels][s][dof]
    const double *sol;        // device [nModels][s][dof] K^-1 applied to the unit loads (displacement layout)
    int maxModes;
    int32_t *nModes;          // device [nModels]
    double *hz;               // device [nModels][maxModes]
    double *shapes;           // device [nModels][maxModes][dof] or NULL
    int32_t *status;          // device [nModels]
    // scratch of the split pipeline (launch_modal_split), NULL for the single-kernel path
    double *qScratch;         // device [nModels][s * LD] the orthogonal Q of the tridiagonalisation
    double *deScratch;        // device [nModels][2 s] diagonal | subdiagonal of T; the eigenvalues replace the diagonal
    double2 *rotLog;          // device [nModels][rotCap] (cos, sin) of every QL rotation, sweep after sweep
    int4 *rotHdr;             // device [nModels][rotMaxSweeps] {l, m, offset into rotLog, 0} of every sweep
    int32_t *rotSweeps;       // device [nModels] number of sweeps, -1 = QL did not converge / log full
    int rotCap, rotMaxSweeps;
};

// ---- linear buckling (model.go:491-601): eigenvalues of (K^-1)_ff G_ff per (model, load case) ----
struct BucklingLaunch {
    int nModels, nLC, s, dof, nB, amount, maxOut;
    const int32_t *freeDof;   // device [s] global DOF of free unknown i (ascending)
    const int32_t *gPtr;      // device [s*s+1] contributions to G(i, j), position i + j*s
    const int32_t *gIdx;      // device [gPtr[s*s]] beam*36 + local index into gUnit
    const double *gUnit;      // device [nB][36] geometric element matrices in global axes for unit axial force
    const double *sol;        // device [nModels][s][dof] K^-1 applied to the unit loads of the free DOFs
    const double *forces;     // device [nModels][nLC][nB][6] beam end forces of the load cases
    int32_t *nOut;            // device [nModels][nLC]
    double *factors;          // device [nModels][nLC][maxOut]
    double *shapes;           // device [nModels][nLC][maxOut][dof] or NULL
    int32_t *status;          // device [nModels]
};
// loads[nModels][s][dof] = unit load on free DOF i in case i
cudaError_t launch_unit_loads(double *loads, int nModels, int s, int dof, const int32_t *freeDof, cudaStream_t st);
// gscr: NULL (work matrices in shared memory) or nModels * nLC records of buckling_scratch_doubles() doubles
cudaError_t launch_buckling_eigen(const BucklingLaunch &B, int smemOptin, cudaStream_t st, double *gscr);
size_t buckling_scratch_doubles(int n, int nB, int smemOptin);
int buckling_max_s(int smemOptin);

cudaError_t launch_modal_loads(const ModalLaunch &M, cudaStream_t st);
// returns cudaErrorInvalidConfiguration when s is too large for the shared-memory resident eigen solver
cudaError_t launch_modal_eigen(const ModalLaunch &M, int smemOptin, cudaStream_t st);
int modal_max_s(int smemOptin);
// the split pipeline: tridiagonalisation | QL recurrences of all models at once | rotations + tensor-core shapes
bool modal_split_fits(int n, int smemOptin);
void modal_split_scratch(int n, size_t &qDoubles, size_t &rotCap, size_t &maxSweeps);
cudaError_t launch_modal_split(const ModalLaunch &M, int smemOptin, cudaStream_t st);

} // namespace hdb
