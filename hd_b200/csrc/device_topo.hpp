// device_topo.hpp — device-side view of the symbolic phase and kernel launch interface.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hdb200.h"

namespace hdb {

struct DevTopo {
    int nP, nB, dof, nFree, nnzL, nSlots, nSup, nnzR, nnzK, maxColCount;
    // geometry / topology (shared by every model of a batch)
    const double *points;      // [nP][2]
    const int32_t *beamN;      // [nB][2]
    const uint32_t *pinMask;   // [nB]
    const uint8_t *dropped;    // [nB*36] 1 = entry is exactly zero in the representative model
    // assembly gather lists (codes are padded smem indices beam*KE_LD + local)
    const int32_t *gptr, *gsrc;   // factor slots
    const int32_t *rgptr, *rgsrc; // reaction entries
    const int32_t *kptr, *ksrc;   // full CSR pattern (tests)
    // factorization schedule
    const int32_t *Lp, *Li, *uptr, *utgt;
    // solve / recovery
    const int32_t *rLp, *rLcol, *rLslot;
    const int32_t *xrow;       // [dof]
    const int32_t *beamRows;   // [nB*6]
    const int32_t *supIndex;   // [dof]
    const int32_t *rptr, *rcol;
};

constexpr int KE_LD = 37; // padded leading dimension of one element matrix in shared memory

struct StaticLaunch {
    int nModels, nLC, ld;      // ld = leading dimension (models) of the interleaved factor workspace
    const double *sections;    // [nModels][nB][3]
    const double *loads;       // [nModels][nLC][dof]
    double *disp, *forces, *react; // outputs, any may be NULL
    int32_t *status, *failCase;    // failCase may be NULL
    double *W;                 // [nSlots][ld] factor workspace, interleaved over models
    double *Ksf;               // [nModels][nnzR]
    double *kval;              // [nModels][nnzK] or NULL
    double *keScratch;         // [nModels][nB*KE_LD] or NULL (NULL = element matrices live in shared memory)
};

// each returns cudaGetLastError() of its launch
cudaError_t launch_assemble(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_factor(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_solve_recover(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_elements(const DevTopo &T, int nModels, const double *sections, double *klocal, double *tr,
                            double *kglobal, cudaStream_t st);
// shared-memory requirement checks (host); return false when the model does not fit
bool assemble_fits_smem(const DevTopo &T, int smemOptin);
bool solve_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes);

} // namespace hdb
