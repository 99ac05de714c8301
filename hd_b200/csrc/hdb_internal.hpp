// hdb_internal.hpp — internal structures of libhdb200.so (not part of the C ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/hdb200.h"

namespace hdb {

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

    // ---- elimination order and factor pattern ----
    std::vector<int32_t> perm;      // [nFree] position -> reduced index
    std::vector<int32_t> iperm;     // [nFree] reduced index -> position
    std::vector<int32_t> Lp, Li;    // strictly lower CSC pattern of L in position numbering
    int nnzL = 0;
    int nSlots = 0;                 // nnzL + nFree (L entries, then the diagonal)

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
    std::vector<int32_t> rcol;      // [nnzR] position (permuted reduced index) of the column
    std::vector<int32_t> rgptr, rgsrc; // gather lists of the K[r][j] values
    int nnzR = 0;
    // factorization: column j updates targets utgt[uptr[j]..uptr[j+1]) enumerated a outer, b<=a inner
    std::vector<int32_t> uptr, utgt;
    // row access of L for the forward solve: row i has entries (column rLcol, value slot rLslot)
    std::vector<int32_t> rLp, rLcol, rLslot;
    // DOF -> x row (position) for free DOFs, -1 supported/eliminated, -2 ignore (load must be zero)
    std::vector<int32_t> xrow;      // [dof]
    // per beam: x rows of its 6 global DOFs
    std::vector<int32_t> beamRows;  // [nB*6]
    // per beam: pin mask
    std::vector<uint32_t> pinMask;  // [nB]
    std::vector<int32_t> beamN32;   // [nB*2]
    int maxColCount = 0;            // largest |struct(L_j)|
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
