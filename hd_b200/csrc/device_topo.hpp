// device_topo.hpp — device-side view of the symbolic phase and kernel launch interface.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/hdb200.h"

namespace hdb {

// substitution op lists (one warp walks them, one lane per load case; built at the end of build_symbolic):
// runs of int4 records, each run led by a header {2 * count, 0, kind, 0}; an op is {slot of the off block L_IJ,
// header slot of the target block, tgt | src << 14 | flags, 0} and means acc(tgt) -= L x(src) resp. L^T x(src)
enum : int { RUN_GENERIC = 0, RUN_CHAIN = 1 };
constexpr int OPF_FIRST = 1 << 28;   // first op of a target block: the accumulator starts from its right-hand side
constexpr int OPF_PREV = 1 << 29;    // source block is the block stored by the op just before (values are in registers)
constexpr int OPF_LAST = 1 << 30;    // last op of a target block: finish with the diagonal block and store x(tgt)
constexpr int OPF_NOPULL = (int)(1u << 31); // block without off blocks in this direction
constexpr int OP_BLK_MASK = (1 << 14) - 1;
constexpr int OP_SRC_SHIFT = 14;

struct DevTopo {
    int nP, nB, dof, nFree, nBlk, nPad, nSlots, nSup, nnzR, nnzK, nDummy, maxBlkCount, winNeed;
    int nFwdOps, nBwdOps; // records incl. run headers and terminators
    int gWidth, rgWidth;  // widths of the ELL gather lists, 0 = use the CSR lists
    int rWidth;           // widest reaction row K[r, free columns] (k_solve_bulk holds the rows as fixed-width lists)
    int rChunk, rNCh;     // k_solve_bulk: a [dof] output row leaves in rNCh chunks of rChunk doubles (even, <= BULK_CHUNK)
    const int32_t *rsPtr, *rsList; // supported DOFs of row chunk j: entries [rsPtr[j], rsPtr[j+1]) = position in the chunk | index among the supported DOFs << 16
    // geometry / topology (shared by every model of a batch)
    const double *points;      // [nP][2]
    const int32_t *beamN;      // [nB][2]
    const uint32_t *pinMask;   // [nB]
    const uint8_t *dropped;    // [nB*36] 1 = matrix position absent from the representative pattern
    const unsigned long long *droppedMask; // [nB] the same, one bit per position
    // assembly gather lists (codes are padded smem indices beam*KE_LD + local)
    const int32_t *gptr, *gsrc;   // factor slots
    const int32_t *gEll, *rgEll;  // [gWidth][nSlots], [rgWidth][nnzR]: the same lists, fixed width, -1 = none
    // the same two lists over the compact element store of k_assemble_quad: code = (keOff[beam] + value) * 2 + negate,
    // -1 = none, -2 = unit diagonal of a dummy unknown; a beam without pins keeps its 8 distinct values
    // (element.cuh: rigid_global_values), a beam with pins all 36 entries
    const int32_t *cgEll, *crgEll;
    const int32_t *keOff;         // [nB] first value of a beam in the compact store
    int keLen;                    // values per model in the compact store; 0 = the compact lists were not built
    const double *geom;           // [nB][3] cos, sin, length of every beam (element.cuh routines, evaluated once on the host)
    const uint8_t *slotInit;      // [nSlots]
    const int32_t *rgptr, *rgsrc; // reaction entries
    const int32_t *kptr, *ksrc;   // full CSR pattern (tests)
    // block factorization schedule
    const int32_t *BLp, *BLi, *colOff, *blkSlot, *pairPtr, *pairTgt, *needUpto;
    const int4 *colDesc;       // [nBlk] {first slot, off-block count, prefetch bound, first Schur target}
    const int4 *fwdOps, *bwdOps; // substitution op lists
    // recovery
    const int32_t *xrow;       // [dof]
    const int32_t *beamRows;   // [nB*6]
    const int32_t *supIndex;   // [dof]
    const int32_t *supDofs;    // [nSup] supported DOFs ascending
    const int32_t *rptr, *rcol;
    const int32_t *dummyPos;   // [nDummy]
};

constexpr int BULK_CHUNK = 192; // doubles per staging slot of k_solve_bulk = one bulk store of 1536 bytes
constexpr int KE_LD = 37; // padded leading dimension of one element matrix in shared memory

struct StaticLaunch {
    int nModels, nLC, ld;      // ld = leading dimension (models) of the interleaved factor workspace
    const double *sections;    // [nModels][nB][3]
    const double *points;      // [nModels][nP][2] per-model coordinates, or NULL: every model has the topology's geometry
    double *geom;              // [nModels][nB][3] cos, sin, length per model (k_geometry), NULL with points == NULL
    const double *loads;       // [nModels][nLC][dof]
    double *disp, *forces, *react; // outputs, any may be NULL
    int32_t *status, *failCase;    // failCase may be NULL
    double *W;                 // [ld/32][nSlots][32] factor workspace, interleaved over the 32 models of a tile (assembly -> factorization)
    double *Lm;                // [ld][nSlots] finished factor, model-major (factorization -> solve)
    double *Ksf;               // [nModels][nnzR]
    double *kval;              // [nModels][nnzK] or NULL
    double *keScratch;         // [nModels][nB*KE_LD] or NULL (NULL = element matrices live in shared memory)
    long long *dbgClocks;      // [blocks][32] phase cycle counters of k_solve_pipe (HDB_DEBUG_CLOCKS=1) or NULL
};

// mod.LoadUniform + assemblyNodeLoad on the device (mod/uniform.go:46-211, model.go:683-698)
struct UniformLaunch {
    int nModels, nLC, nU, nN, dof;
    const int32_t *uPtr, *uIdx;   // per DOF: contributions of the uniform loads, uIdx = item * 8 + code
    const double *uGeo;           // [nU][4] dx, dy, projected/true lengths used for Fx and Fy
    const int32_t *nPtr, *nIdx;   // per DOF: contributions of the node loads, nIdx = node item * 3 + component
    const double *q;              // [nModels][nLC][nU][2]
    const double *f;              // [nModels][nLC][nN][3] or NULL
    double *loads;                // [nModels][nLC][dof]
};
cudaError_t launch_uniform_loads(const UniformLaunch &U, cudaStream_t st);

// each returns cudaGetLastError() of its launch
cudaError_t launch_geometry(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_assemble(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_factor(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_solve_pipe(const DevTopo &T, const StaticLaunch &L, cudaStream_t st);
cudaError_t launch_solve_bulk(const DevTopo &T, const StaticLaunch &L, cudaStream_t st, int smemOptin, int nSM, bool &taken); // kernels_solve_bulk.cu
bool solve_bulk_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes);
bool solve_general_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes);
cudaError_t launch_elements(const DevTopo &T, int nModels, const double *sections, double *klocal, double *tr,
                            double *kglobal, cudaStream_t st);
// shared-memory requirement checks (host); return false when the model does not fit
bool assemble_fits_smem(const DevTopo &T, int smemOptin);
bool solve_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes);

} // namespace hdb
