// symbolic.cpp — host symbolic phase of libhdb200.so.
//
// Replaces the structural half of getK / assemblyK (reference model.go:318-333, 607-681) and
// of sparse.LU.Factorize's analysis: DOF numbering (model.go:402, 626-627), the value-dependent
// sparsity pattern (exact zeros dropped before duplicates are summed, model.go:645-652), the
// `ignore` list (model.go:657-678), the eliminated set (model.go:326-328, 700-710), and — new —
// the elimination order, the fill pattern of the LDL^T factor, and the index schedules the
// CUDA kernels execute for every model of a batch.  Pure host C++; runs without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>

#include "element.cuh"
#include "hdb_internal.hpp"

namespace hdb {

static bool bad_float(double f) { return std::isnan(f) || std::isinf(f); }

// ------------------------------------------------------------------------------------------
// check.go restated (accept / reject behaviour of hd for the model data)
// ------------------------------------------------------------------------------------------
static void add(std::string &msg, const char *fmt, long a = 0, long b = 0)
{
    char buf[160];
    std::snprintf(buf, sizeof buf, fmt, a, b);
    if (msg.size() < 4000) {
        msg += buf;
        msg += "\n";
    }
}

// check.go:47-109
static bool graph_is_single(int32_t nP, int32_t nB, const hdb_beam *beams)
{
    if (nB == 0) return true;
    std::vector<std::vector<int>> adj(nP); // point -> beams
    for (int b = 0; b < nB; b++) {
        adj[beams[b].n0].push_back(b);
        adj[beams[b].n1].push_back(b);
    }
    std::vector<char> mark(nP, 0), bmark(nB, 0);
    std::queue<int> q;
    q.push(0);
    bmark[0] = 1;
    while (!q.empty()) {
        int b = q.front();
        q.pop();
        for (int64_t n : {beams[b].n0, beams[b].n1}) {
            mark[n] = 1;
            for (int nb : adj[n])
                if (!bmark[nb]) {
                    bmark[nb] = 1;
                    q.push(nb);
                }
        }
    }
    for (int i = 0; i < nP; i++)
        if (!mark[i]) return false;
    return true;
}

int check_model_impl(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, int32_t nPins,
                     const uint8_t *pins, int32_t nSupports, const uint8_t *supports, std::string &msg)
{
    (void)supports;
    bool err = false;
    for (int i = 0; i < nP; i++) // checkPoints, check.go:223-244
        for (int j = 0; j < 2; j++)
            if (bad_float(points[2 * i + j])) {
                add(msg, "Error in point %ld, coordinate %ld : not finite", i, j);
                err = true;
            }
    if (nP > 0 && nB == 0) { // check.go:266 ErrorBeamAmount
        add(msg, "Zero beams with non-zero amount of points");
        return HDB_ERR_INPUT;
    }
    for (int b = 0; b < nB; b++) { // checkBeams, check.go:270-325
        bool idxOk = true;
        for (int j = 0; j < 2; j++) {
            int64_t n = j ? beams[b].n1 : beams[b].n0;
            if (n < 0 || n >= nP) {
                add(msg, "Error in beam %ld, checking `Point %ld` : outside index of point", b, j);
                err = true;
                idxOk = false;
            }
        }
        if (beams[b].n0 == beams[b].n1) {
            add(msg, "Error in beam %ld, checking `Point indexes` : is same", b);
            err = true;
        }
        if (idxOk) {
            double L = beam_length(points[2 * beams[b].n0], points[2 * beams[b].n0 + 1], points[2 * beams[b].n1],
                                   points[2 * beams[b].n1 + 1]);
            if (L <= 0) {
                add(msg, "Error in beam %ld, checking `distance between points` : is less or equal zero", b);
                err = true;
            }
        }
        const double v[3] = {beams[b].A, beams[b].J, beams[b].E};
        for (int c = 0; c < 3; c++)
            if (bad_float(v[c]) || v[c] < 0 || v[c] == 0) {
                add(msg, "Error in beam %ld, section value %ld : NaN, infinite, negative or zero", b, c);
                err = true;
            }
    }
    if (nP != nSupports) { // checkSupports, check.go:332-337
        add(msg, "Amount of supports is not same amount of points");
        err = true;
    }
    if (nB != nPins) { // checkPins, check.go:353-356
        add(msg, "Amount of pins is not same beams, %ld != %ld", nB, nPins);
        err = true;
    }
    for (int b = 0; b < nPins && pins; b++) { // check.go:357-391
        const uint8_t *pn = pins + 6 * b;
        int cnt = 0;
        for (int j = 0; j < 6; j++) cnt += pn[j] ? 1 : 0;
        if (cnt > 4) {
            add(msg, "Error in pin of beam %ld: Amount of pins of point is %ld more 4", b, cnt);
            err = true;
        }
        if (pn[0] && pn[3]) {
            add(msg, "Error in pin of beam %ld: Not acceptable X direction is free", b);
            err = true;
        }
        if (pn[2] && pn[5] && !(pn[1] && pn[4])) {
            add(msg, "Error in pin of beam %ld: Pins [1] and [4] in direction Y must be true", b);
            err = true;
        }
    }
    if (err) return HDB_ERR_INPUT;
    if (!graph_is_single(nP, nB, beams)) { // check.go:31-36, 47-109
        add(msg, "Model isn`t single calculation model");
        return HDB_ERR_INPUT;
    }
    return HDB_OK;
}

// ------------------------------------------------------------------------------------------
// ordering
// ------------------------------------------------------------------------------------------
static void rcm_order(int n, const std::vector<std::vector<int>> &adj, std::vector<int32_t> &perm)
{
    std::vector<char> seen(n, 0);
    std::vector<int> order;
    order.reserve(n);
    std::vector<int> deg(n);
    for (int i = 0; i < n; i++) deg[i] = (int)adj[i].size();
    for (;;) {
        int start = -1;
        for (int i = 0; i < n; i++)
            if (!seen[i] && (start < 0 || deg[i] < deg[start])) start = i;
        if (start < 0) break;
        std::queue<int> q;
        q.push(start);
        seen[start] = 1;
        while (!q.empty()) {
            int u = q.front();
            q.pop();
            order.push_back(u);
            std::vector<int> nb;
            for (int v : adj[u])
                if (!seen[v]) {
                    seen[v] = 1;
                    nb.push_back(v);
                }
            std::stable_sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] < deg[b]; });
            for (int v : nb) q.push(v);
        }
    }
    perm.resize(n);
    for (int i = 0; i < n; i++) perm[i] = order[n - 1 - i];
}

// ------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------
int build_symbolic(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, const uint8_t *pins,
                   const uint8_t *supports, int ordering, Symbolic &S, std::string &detail)
{
    if (nP <= 0 || nB <= 0 || !points || !beams || !supports) {
        detail = "hdb_topology_create: empty model or NULL argument";
        return HDB_ERR_ARG;
    }
    if ((int64_t)nB * 36 > INT32_MAX / 2 || (int64_t)nP * 3 > INT32_MAX / 4) {
        detail = "model too large for int32 indexing";
        return HDB_ERR_UNSUPPORTED;
    }
    S = Symbolic();
    S.nP = nP;
    S.nB = nB;
    S.dof = 3 * nP;
    S.ordering = ordering;
    S.points.assign(points, points + 2 * (size_t)nP);
    S.supports.assign(supports, supports + 3 * (size_t)nP);
    S.pins.assign(6 * (size_t)nB, 0); // model.go:344-346: empty Pins => all rigid
    if (pins) S.pins.assign(pins, pins + 6 * (size_t)nB);
    {
        std::string msg;
        int rc = check_model_impl(nP, points, nB, beams, nB, S.pins.data(), nP, supports, msg);
        if (rc != HDB_OK) {
            detail = "hdb_topology_create: model rejected (check.go rules):\n" + msg;
            return HDB_ERR_INPUT;
        }
    }
    S.beamN.resize(2 * (size_t)nB);
    S.beamN32.resize(2 * (size_t)nB);
    S.repSections.resize(3 * (size_t)nB);
    S.pinMask.resize(nB);
    for (int b = 0; b < nB; b++) {
        S.beamN[2 * b] = beams[b].n0;
        S.beamN[2 * b + 1] = beams[b].n1;
        S.beamN32[2 * b] = (int32_t)beams[b].n0;
        S.beamN32[2 * b + 1] = (int32_t)beams[b].n1;
        S.repSections[3 * b] = beams[b].A;
        S.repSections[3 * b + 1] = beams[b].J;
        S.repSections[3 * b + 2] = beams[b].E;
        uint32_t mask = 0;
        for (int j = 0; j < 6; j++)
            if (S.pins[6 * b + j]) mask |= 1u << j;
        S.pinMask[b] = mask;
    }
    const int dof = S.dof;

    // ---- representative element matrices, literal rule; triplets in model.go:622-635 order ----
    struct Trip {
        int32_t row, col, code; // code = beam*36 + local row-major index
        bool zero;
    };
    std::vector<Trip> trips;
    trips.reserve((size_t)nB * 36);
    for (int b = 0; b < nB; b++) {
        const int64_t n0 = beams[b].n0, n1 = beams[b].n1;
        const double x0 = points[2 * n0], y0 = points[2 * n0 + 1], x1 = points[2 * n1], y1 = points[2 * n1 + 1];
        const double L = beam_length(x0, y0, x1, y1);
        BeamCoef cf = beam_coef(beams[b].A, beams[b].J, beams[b].E, L);
        double k[36], kg[36], c, s;
        stiff_beam2d(cf, S.pinMask[b], k);
        beam_cs(x0, y0, x1, y1, L, c, s);
        rotate_to_global(k, c, s, kg);
        const int64_t N[2] = {n0, n1};
        for (int p1 = 0; p1 < 2; p1++)
            for (int p2 = 0; p2 < 2; p2++)
                for (int r1 = 0; r1 < 3; r1++)
                    for (int r2 = 0; r2 < 3; r2++) {
                        const int local = (p1 * 3 + r1) * 6 + (p2 * 3 + r2);
                        const double val = kg[local];
                        if (bad_float(val)) {
                            detail = "hdb_topology_create: representative model has a NaN/Inf stiffness entry";
                            return HDB_ERR_INPUT;
                        }
                        trips.push_back({(int32_t)(N[p1] * 3 + r1), (int32_t)(N[p2] * 3 + r2), b * 36 + local,
                                         !(val != 0.0)});
                    }
    }

    // ---- pattern: entries kept by model.go:645-652, canonical CSR ----
    std::vector<std::vector<int32_t>> rowCols(dof);
    for (const Trip &t : trips)
        if (!t.zero) rowCols[t.row].push_back(t.col);
    S.rowptr.assign(dof + 1, 0);
    for (int r = 0; r < dof; r++) {
        auto &v = rowCols[r];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        S.rowptr[r + 1] = S.rowptr[r] + (int32_t)v.size();
    }
    S.nnzK = S.rowptr[dof];
    S.colind.resize(S.nnzK);
    for (int r = 0; r < dof; r++) std::copy(rowCols[r].begin(), rowCols[r].end(), S.colind.begin() + S.rowptr[r]);
    auto csr_slot = [&](int r, int c) -> int32_t {
        auto b = S.colind.begin() + S.rowptr[r], e = S.colind.begin() + S.rowptr[r + 1];
        auto it = std::lower_bound(b, e, c);
        return (it != e && *it == c) ? (int32_t)(it - S.colind.begin()) : -1;
    };
    // scatter (API): per triplet, -1 where the representative value is exactly zero (dropped by the
    // reference).  posSlot (kernels): the slot of the triplet's POSITION whenever that position exists in
    // the pattern, even if this beam's own contribution is an exact zero in the representative model —
    // another model of the batch may carry a round-off residue there (pin condensation, feBeam2d.go:58-137)
    // which the reference would add; adding the representative's exact 0.0 changes no sum.
    S.scatter.assign((size_t)nB * 36, -1);
    S.posSlot.assign((size_t)nB * 36, -1);
    S.kptr.assign(S.nnzK + 1, 0);
    for (size_t q = 0; q < trips.size(); q++) {
        const int32_t slot = csr_slot(trips[q].row, trips[q].col);
        S.posSlot[q] = slot;
        if (!trips[q].zero) S.scatter[q] = slot;
        if (slot >= 0) S.kptr[slot + 1]++;
    }
    for (int s = 0; s < S.nnzK; s++) S.kptr[s + 1] += S.kptr[s];
    S.ksrc.resize(S.kptr[S.nnzK]);
    {
        std::vector<int32_t> w(S.kptr.begin(), S.kptr.end() - 1);
        for (size_t q = 0; q < trips.size(); q++) // beam order = summation order of cs_dupl
            if (S.posSlot[q] >= 0) S.ksrc[w[S.posSlot[q]]++] = trips[q].code;
    }

    // ---- ignore, eliminated, free numbering ----
    for (int i = 0; i < dof; i++)
        if (csr_slot(i, i) < 0) S.ignore.push_back(i);
    {
        std::vector<int32_t> e(S.ignore);
        for (int p = 0; p < nP; p++)
            for (int i = 0; i < 3; i++)
                if (supports[3 * p + i]) e.push_back(3 * p + i);
        std::sort(e.begin(), e.end());
        e.erase(std::unique(e.begin(), e.end()), e.end());
        S.elim = e;
    }
    S.freeIndex.assign(dof, 0);
    for (int32_t e : S.elim) S.freeIndex[e] = -1;
    S.nFree = 0;
    for (int i = 0; i < dof; i++)
        if (S.freeIndex[i] == 0) {
            S.freeIndex[i] = S.nFree++;
            S.freeDofs.push_back(i);
        }
    const int n = S.nFree;
    if (n == 0) {
        detail = "hdb_topology_create: no free degree of freedom";
        return HDB_ERR_INPUT;
    }
    S.supIndex.assign(dof, -1);
    for (int p = 0; p < nP; p++)
        for (int i = 0; i < 3; i++)
            if (supports[3 * p + i]) {
                S.supIndex[3 * p + i] = (int32_t)S.supDofs.size();
                S.supDofs.push_back(3 * p + i);
            }
    S.nSup = (int)S.supDofs.size();

    // ---- adjacency of K_ff (reduced numbering) ----
    std::vector<std::vector<int>> adj(n);
    S.nnzKff = 0;
    for (int r = 0; r < dof; r++) {
        const int i = S.freeIndex[r];
        if (i < 0) continue;
        for (int q = S.rowptr[r]; q < S.rowptr[r + 1]; q++) {
            const int j = S.freeIndex[S.colind[q]];
            if (j < 0) continue;
            S.nnzKff++;
            if (i != j) {
                adj[i].push_back(j);
                adj[j].push_back(i);
            }
        }
    }
    for (auto &v : adj) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }

    // ---- node blocks: every point with at least one free DOF is one 3x3 block (X, Y, M) ----
    // Positions 3*blk + component; components that are supported or ignored are DUMMY unknowns
    // (unit diagonal, zero right-hand side), so every block of the factor is a dense 3x3.
    S.blkOfPoint.assign(nP, -1);
    std::vector<int> blockPoints;
    for (int p = 0; p < nP; p++)
        if (S.freeIndex[3 * p] >= 0 || S.freeIndex[3 * p + 1] >= 0 || S.freeIndex[3 * p + 2] >= 0) blockPoints.push_back(p);
    const int nBlk = (int)blockPoints.size();
    // block adjacency from K_ff
    std::vector<int> rankOfPoint(nP, -1);
    for (int i = 0; i < nBlk; i++) rankOfPoint[blockPoints[i]] = i;
    std::vector<std::vector<int>> badj(nBlk);
    for (int i = 0; i < n; i++)
        for (int j : adj[i]) {
            const int bi = rankOfPoint[S.freeDofs[i] / 3], bj = rankOfPoint[S.freeDofs[j] / 3];
            if (bi != bj) badj[bi].push_back(bj);
        }
    for (auto &v : badj) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
    // elimination order of the blocks
    std::vector<int32_t> border; // position -> rank in blockPoints
    if (ordering == 1) {
        rcm_order(nBlk, badj, border);
    } else { // natural: ascending point index = ascending DOF = identity permutation of the free DOFs,
             // which is what the reference's sparse.LU uses
        border.resize(nBlk);
        std::iota(border.begin(), border.end(), 0);
    }
    std::vector<int> bpos(nBlk);
    S.pointOfBlk.resize(nBlk);
    for (int pos = 0; pos < nBlk; pos++) {
        bpos[border[pos]] = pos;
        S.pointOfBlk[pos] = blockPoints[border[pos]];
        S.blkOfPoint[blockPoints[border[pos]]] = pos;
    }
    S.nBlk = nBlk;
    S.nPad = 3 * nBlk;
    // scalar permutation implied by the block order (exported; identity for natural ordering)
    S.perm.clear();
    for (int b = 0; b < nBlk; b++)
        for (int c = 0; c < 3; c++) {
            const int fi = S.freeIndex[3 * S.pointOfBlk[b] + c];
            if (fi >= 0) S.perm.push_back(fi);
        }
    S.iperm.assign(n, 0);
    for (int pos = 0; pos < n; pos++) S.iperm[S.perm[pos]] = pos;

    // ---- scalar symbolic LDL^T (informational: the minimal fill pattern, exported for tests) ----
    {
        std::vector<std::vector<int32_t>> cols(n);
        for (int i = 0; i < n; i++) {
            const int pi = S.iperm[i];
            for (int j : adj[i]) {
                const int pj = S.iperm[j];
                if (pj > pi) cols[pi].push_back(pj);
            }
        }
        S.Lp.assign(n + 1, 0);
        for (int j = 0; j < n; j++) {
            auto &v = cols[j];
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
            if (!v.empty()) {
                const int parent = v[0];
                cols[parent].insert(cols[parent].end(), v.begin() + 1, v.end());
            }
            S.Lp[j + 1] = S.Lp[j] + (int32_t)v.size();
            S.maxColCount = std::max(S.maxColCount, (int)v.size());
        }
        S.nnzL = S.Lp[n];
        S.Li.resize(S.nnzL);
        for (int j = 0; j < n; j++) std::copy(cols[j].begin(), cols[j].end(), S.Li.begin() + S.Lp[j]);
    }

    // ---- block symbolic LDL^T ----
    std::vector<std::vector<int32_t>> bcols(nBlk);
    for (int i = 0; i < nBlk; i++)
        for (int j : badj[i])
            if (bpos[j] > bpos[i]) bcols[bpos[i]].push_back(bpos[j]);
    S.BLp.assign(nBlk + 1, 0);
    for (int J = 0; J < nBlk; J++) {
        auto &v = bcols[J];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        if (!v.empty()) bcols[v[0]].insert(bcols[v[0]].end(), v.begin() + 1, v.end());
        S.BLp[J + 1] = S.BLp[J] + (int32_t)v.size();
        S.maxBlkCount = std::max(S.maxBlkCount, (int)v.size());
    }
    const int nOff = S.BLp[nBlk];
    S.BLi.resize(nOff);
    for (int J = 0; J < nBlk; J++) std::copy(bcols[J].begin(), bcols[J].end(), S.BLi.begin() + S.BLp[J]);
    // slot layout: block column J = [d0 d1 d2 | l10 l20 l21 | off blocks: 9 values row-major + 1 pad each]
    S.colOff.assign(nBlk + 1, 0);
    S.blkSlot.resize(nOff);
    for (int J = 0; J < nBlk; J++) {
        for (int q = S.BLp[J]; q < S.BLp[J + 1]; q++) S.blkSlot[q] = S.colOff[J] + BLK_HDR + BLK_STRIDE * (q - S.BLp[J]);
        S.colOff[J + 1] = S.colOff[J] + BLK_HDR + BLK_STRIDE * (S.BLp[J + 1] - S.BLp[J]);
    }
    S.nSlots = S.colOff[nBlk];
    auto off_block = [&](int I, int J) -> int32_t { // I > J
        auto b = S.BLi.begin() + S.BLp[J], e = S.BLi.begin() + S.BLp[J + 1];
        auto it = std::lower_bound(b, e, I);
        return (it != e && *it == I) ? (int32_t)(it - S.BLi.begin()) : -1;
    };
    auto slot_of = [&](int pr, int pc) -> int32_t { // pr >= pc, padded positions
        const int I = pr / 3, r = pr % 3, J = pc / 3, c = pc % 3;
        if (pr == pc) return S.colOff[J] + c;
        if (I == J) return S.colOff[J] + 3 + (r == 1 ? 0 : (c == 0 ? 1 : 2));
        const int32_t q = off_block(I, J);
        return q < 0 ? -1 : S.blkSlot[q] + r * 3 + c;
    };

    // ---- DOF -> x row (padded position) ----
    S.xrow.assign(dof, -1);
    S.slotInit.assign(S.nSlots, 0);
    for (int b = 0; b < nBlk; b++)
        for (int c = 0; c < 3; c++) {
            const int g = 3 * S.pointOfBlk[b] + c;
            if (S.freeIndex[g] >= 0) S.xrow[g] = 3 * b + c;
            else {
                S.dummyPos.push_back(3 * b + c);
                S.slotInit[S.colOff[b] + c] = 1; // unit diagonal of a dummy unknown
            }
        }
    for (int32_t g : S.ignore) S.xrow[g] = -2;
    S.beamRows.resize((size_t)nB * 6);
    for (int b = 0; b < nB; b++)
        for (int p = 0; p < 2; p++)
            for (int r = 0; r < 3; r++) {
                int32_t xr = S.xrow[S.beamN[2 * b + p] * 3 + r];
                S.beamRows[(size_t)b * 6 + p * 3 + r] = xr < 0 ? -1 : xr;
            }

    // ---- gather lists: factor slots and reaction rows, contributions in beam order ----
    std::vector<std::vector<int32_t>> slotSrc(S.nSlots);
    std::vector<std::map<int32_t, std::vector<int32_t>>> rrow(S.nSup); // column DOF -> codes
    for (size_t q = 0; q < trips.size(); q++) {
        if (S.posSlot[q] < 0) continue;
        const Trip &t = trips[q];
        const int xi = S.xrow[t.row], xj = S.xrow[t.col];
        if (xi >= 0 && xj >= 0 && xi >= xj) { // the factor takes the lower triangle (in elimination order)
            const int32_t sl = slot_of(xi, xj);
            if (sl < 0) {
                detail = "internal: factor slot missing";
                return HDB_ERR_ARG;
            }
            slotSrc[sl].push_back(t.code);
        }
        const int si = S.supIndex[t.row];
        if (si >= 0 && xj >= 0) rrow[si][t.col].push_back(t.code);
    }
    S.gptr.assign(S.nSlots + 1, 0);
    for (int s2 = 0; s2 < S.nSlots; s2++) S.gptr[s2 + 1] = S.gptr[s2] + (int32_t)slotSrc[s2].size();
    S.gsrc.resize(S.gptr[S.nSlots]);
    for (int s2 = 0; s2 < S.nSlots; s2++) std::copy(slotSrc[s2].begin(), slotSrc[s2].end(), S.gsrc.begin() + S.gptr[s2]);
    S.rptr.assign(S.nSup + 1, 0);
    S.rgptr.assign(1, 0);
    for (int si = 0; si < S.nSup; si++) {
        for (auto &kv : rrow[si]) { // std::map: ascending column DOF = the CSC scan order of model.go:481
            S.rcol.push_back(S.xrow[kv.first]);
            S.rgsrc.insert(S.rgsrc.end(), kv.second.begin(), kv.second.end());
            S.rgptr.push_back((int32_t)S.rgsrc.size());
        }
        S.rptr[si + 1] = (int32_t)S.rcol.size();
    }
    S.nnzR = (int)S.rcol.size();

    // ---- block factorization schedule: Schur targets of every pair of off blocks of a column ----
    S.pairPtr.assign(nBlk + 1, 0);
    for (int J = 0; J < nBlk; J++) {
        const int nb = S.BLp[J + 1] - S.BLp[J];
        S.pairPtr[J + 1] = S.pairPtr[J] + nb * (nb + 1) / 2;
        S.nUpdates += 3 * (3 + 3 * nb) * (4 + 3 * nb) / 2; // scalar multiply-subtracts, upper bound
    }
    S.pairTgt.resize(S.pairPtr[nBlk]);
    S.needUpto.assign(nBlk, 0);
    for (int J = 0; J < nBlk; J++) {
        int32_t k = S.pairPtr[J];
        int maxI = J;
        for (int a = S.BLp[J]; a < S.BLp[J + 1]; a++) {
            maxI = std::max(maxI, (int)S.BLi[a]);
            for (int b = S.BLp[J]; b <= a; b++) {
                int32_t tg;
                if (a == b) tg = S.colOff[S.BLi[a]]; // diagonal block header of row block I
                else {
                    const int32_t q = off_block(S.BLi[a], S.BLi[b]);
                    if (q < 0) {
                        detail = "internal: fill block missing";
                        return HDB_ERR_ARG;
                    }
                    tg = S.blkSlot[q];
                }
                S.pairTgt[k++] = tg;
            }
        }
        S.needUpto[J] = S.colOff[maxI + 1];
    }
    for (int J = 1; J < nBlk; J++) S.needUpto[J] = std::max(S.needUpto[J], S.needUpto[J - 1]);
    S.winNeed = 0;
    for (int J = 0; J < nBlk; J++)
        S.winNeed = std::max(S.winNeed, S.needUpto[std::min(J + FACTOR_PREFETCH, nBlk - 1)] - S.colOff[J]);

    // ---- row access of the off blocks for the forward substitution ----
    S.fwdPtr.assign(nBlk + 1, 0);
    for (int q = 0; q < nOff; q++) S.fwdPtr[S.BLi[q] + 1]++;
    for (int I = 0; I < nBlk; I++) S.fwdPtr[I + 1] += S.fwdPtr[I];
    S.fwdJ.resize(nOff);
    S.fwdSlot.resize(nOff);
    S.fwdQ.resize(nOff);
    {
        std::vector<int32_t> w(S.fwdPtr.begin(), S.fwdPtr.end() - 1);
        for (int J = 0; J < nBlk; J++) // ascending J inside every row
            for (int q = S.BLp[J]; q < S.BLp[J + 1]; q++) {
                const int e = w[S.BLi[q]]++;
                S.fwdJ[e] = J;
                S.fwdSlot[e] = S.blkSlot[q];
                S.fwdQ[e] = q;
            }
    }

    // ---- substitution op lists (see device_topo.hpp) ----
    // The factor kernel leaves every block column as header [1/d0 1/d1 1/d2 | v10 v20 v21] (V = Ldiag^-1) plus
    // its off blocks L_IJ.  One op of the solver warp is acc(tgt) -= L x(src) (forward) or L^T x(src)
    // (backward) on 3-vectors, with flags FIRST (acc starts from the right-hand side of tgt, scaled by D^-1 in
    // the backward pass), LAST (x(tgt) = V acc resp. V^T acc is stored), PREV (src is the block stored by the
    // op just before: still in registers) and NOPULL (block without off blocks in this direction).
    // Ops are grouped into runs, each led by a header record {2 * count, 0, kind, 0}.  A CHAIN run is a
    // sequence of ops that each make up a whole target block (FIRST and LAST) and pull from the target of the
    // op before (the first one from any finished block): the common banded case, walked by a branch-free
    // loop.  All other ops form GENERIC runs with per-op flags; PREV is relative to the run.
    if (nBlk >= (1 << 14)) {
        detail = "too many node blocks for the substitution op encoding";
        return HDB_ERR_ARG;
    }
    const int FFIRST = 1 << 28, FPREV = 1 << 29, FLAST = 1 << 30, FNOPULL = (int)(1u << 31);
    struct Op { int slot, src, tgt, flags; };
    auto segment = [&](const std::vector<Op> &ops, std::vector<int32_t> &out) {
        out.clear();
        auto eligible = [&](const Op &o) { return (o.flags & FFIRST) && (o.flags & FLAST) && !(o.flags & FNOPULL); };
        size_t k = 0;
        while (k < ops.size()) {
            size_t e = k + 1;
            const bool chain = eligible(ops[k]);
            if (chain) while (e < ops.size() && eligible(ops[e]) && ops[e].src == ops[e - 1].tgt) e++;
            else while (e < ops.size() && !eligible(ops[e])) e++;
            // header: an even first field and small other fields, because the chain loop's one-op-ahead prefetch
            // may read (never use) a header as if it were an op
            out.insert(out.end(), {(int32_t)(2 * (e - k)), 0, chain ? 1 : 0, 0});
            int lastDone = -1;
            for (size_t i = k; i < e; i++) {
                int f = ops[i].flags;
                if (!chain && !(f & FNOPULL) && ops[i].src == lastDone) f |= FPREV;
                out.insert(out.end(), {ops[i].slot, S.colOff[ops[i].tgt], ops[i].tgt | (ops[i].src << 14) | f, 0});
                if (f & FLAST) lastDone = ops[i].tgt;
            }
            k = e;
        }
        // two zero records end the list: a zero header stops the walk, and the prefetch may read past a run
        out.insert(out.end(), 8, 0);
    };
    std::vector<Op> fo, bo;
    for (int I = 0; I < nBlk; I++) {
        const int e0 = S.fwdPtr[I], e1 = S.fwdPtr[I + 1];
        if (e0 == e1) fo.push_back({S.colOff[I] + BLK_HDR, 0, I, FFIRST | FLAST | FNOPULL});
        for (int e = e0; e < e1; e++)
            fo.push_back({S.fwdSlot[e], S.fwdJ[e], I, (e == e0 ? FFIRST : 0) | (e == e1 - 1 ? FLAST : 0)});
    }
    for (int J = nBlk - 1; J >= 0; J--) {
        const int q0 = S.BLp[J], q1 = S.BLp[J + 1];
        if (q0 == q1) bo.push_back({S.colOff[J] + BLK_HDR, 0, J, FFIRST | FLAST | FNOPULL});
        for (int q = q0; q < q1; q++)
            bo.push_back({S.blkSlot[q], S.BLi[q], J, (q == q0 ? FFIRST : 0) | (q == q1 - 1 ? FLAST : 0)});
    }
    segment(fo, S.fwdOps);
    segment(bo, S.bwdOps);
    return HDB_OK;
}

} // namespace hdb
