// hdb_internal.hpp — internal structures of libhdb200.so (not part of the C ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hdb200.h"

namespace hdb {

constexpr int BLK_HDR = 6;        // 1/d (3) + strictly lower part of the inverse of the unit diagonal block (3)
constexpr int BLK_STRIDE = 10;    // 9 values of an off block + 1 pad (keeps every block 16-byte aligned)
constexpr int FACTOR_PREFETCH = 6; // block columns the factor kernel prefetches ahead

// Result of the host symbolic phase.  All index arrays are int32 and identical for every model
// of a batch (same topology, geometry, pins, supports).
struct Symbolic {
    int nP = 0, nB = 0, dof = 0;
    std::vector<double> points;     // [nP][2]
    std::vector<int64_t> beamN;     // [nB][2]
    std::vector<double> repSections;// [nB][3] representative A,J,E
    std::vector<uint8_t> pins;      // [nB][6]
    std::vector<uint8_t> supports;  // [nP][3]
    int ordering = 0;

    // ---- pattern of K exactly as model.go:607-681 leaves it ----
    std::vector<int32_t> scatter;   // [nB*36] -> CSR slot or -1 (exact zero, dropped by model.go:645)
    std::vector<int32_t> posSlot;   // [nB*36] -> CSR slot of the triplet's position or -1 (position absent)
    std::vector<int32_t> rowptr, colind; // canonical CSR (ascending columns)
    std::vector<int32_t> ignore;    // DOFs without diagonal entry (model.go:657-678)
    std::vector<int32_t> elim;      // sort-unique(ignore ++ supports) (model.go:326-328)
    std::vector<int32_t> freeIndex; // [dof] reduced index or -1
    std::vector<int32_t> freeDofs;  // [nFree] reduced index -> DOF
    int nFree = 0, nSup = 0, nnzK = 0, nnzKff = 0;

    // ---- elimination order and (informational) scalar factor pattern ----
    std::vector<int32_t> perm;      // [nFree] position -> reduced index
    std::vector<int32_t> iperm;     // [nFree] reduced index -> position
    std::vector<int32_t> Lp, Li;    // strictly lower CSC pattern of the scalar L (minimal fill), exported
    int nnzL = 0;
    int maxColCount = 0;

    // ---- node-block structure used by the kernels ----
    // every point with a free DOF is a 3x3 block; non-free components are dummy unknowns (unit diagonal)
    int nBlk = 0, nPad = 0;
    std::vector<int32_t> blkOfPoint; // [nP] block (in elimination order) or -1
    std::vector<int32_t> pointOfBlk; // [nBlk]
    std::vector<int32_t> BLp, BLi;   // strictly lower block pattern of L incl. fill, CSC over blocks
    int maxBlkCount = 0;
    // value slots: block column J occupies [colOff[J], colOff[J+1]):
    //   [1/d0 1/d1 1/d2 | l10 l20 l21 | per off block (ascending row block): 9 values row-major + 1 pad]
    std::vector<int32_t> colOff;     // [nBlk+1]
    std::vector<int32_t> blkSlot;    // [nOff] first slot of off block q
    int nSlots = 0;
    std::vector<uint8_t> slotInit;   // [nSlots] 1 = the slot starts at 1.0 (diagonal of a dummy unknown)
    std::vector<int32_t> dummyPos;   // padded positions that are not free DOFs

    // ---- numeric schedules ----
    // gather lists: factor slot s sums element entries gsrc[gptr[s]..gptr[s+1]) (code = beam*36+local),
    // in beam order = the reference's duplicate-summation order (cs_dupl, model.go:652)
    std::vector<int32_t> gptr, gsrc;
    // same for the full CSR pattern (hdb_assemble_batch, tests)
    std::vector<int32_t> kptr, ksrc;
    // reactions: supported DOF list (ascending), rows restricted to free columns (ascending DOF)
    std::vector<int32_t> supDofs;   // [nSup]
    std::vector<int32_t> supIndex;  // [dof] index into supDofs or -1
    std::vector<int32_t> rptr;      // [nSup+1]
    std::vector<int32_t> rcol;      // [nnzR] padded position of the column
    std::vector<int32_t> rgptr, rgsrc; // gather lists of the K[r][j] values
    int nnzR = 0;
    // block factorization: for column J the Schur target of every pair (a >= b) of its off blocks,
    // a outer, b inner: first slot of block (I_a, I_b), or the header of diagonal block I_a when a == b
    std::vector<int32_t> pairPtr, pairTgt;
    std::vector<int32_t> needUpto;  // [nBlk] slots [0, needUpto[J]) must be resident to eliminate column J
    int winNeed = 0;                // shared-memory window (slots) the factor kernel needs incl. prefetch
    // forward substitution: row block I pulls from (J, slot) ascending J
    std::vector<int32_t> fwdPtr, fwdJ, fwdSlot, fwdQ;
    // substitution op lists walked by the solve kernel: records {slot of L_IJ, header slot of tgt, tgt | src << 14 | flags, 0}
    std::vector<int32_t> fwdOps, bwdOps;
    // DOF -> x row (padded position) for free DOFs, -1 supported/eliminated, -2 ignore (load must be zero)
    std::vector<int32_t> xrow;      // [dof]
    std::vector<int32_t> beamRows;  // [nB*6] x rows of a beam's 6 global DOFs or -1
    std::vector<uint32_t> pinMask;  // [nB]
    std::vector<int32_t> beamN32;   // [nB*2]
    int64_t nUpdates = 0;
};

// builds everything above; returns HDB_OK / HDB_ERR_*; detail gets a message on failure
int build_symbolic(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams,
                   const uint8_t *pins, const uint8_t *supports, int ordering, Symbolic &S,
                   std::string &detail);

// check.go restated for host shims
int check_model_impl(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, int32_t nPins,
                     const uint8_t *pins, int32_t nSupports, const uint8_t *supports, std::string &msg);

void set_error_detail(const std::string &s);

} // namespace hdb
