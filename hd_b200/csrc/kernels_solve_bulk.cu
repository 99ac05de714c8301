// kernels_solve_bulk.cu — k_solve_bulk: the solve / recovery step of LinearStatic (model.go:415-489) as a persistent
// block per SM whose bytes move through the bulk-copy engine (cp.async.bulk, SASS UBLKCP) wherever that is the cheaper path.
//
// Same three-stage pipeline over two shared-memory buffers as k_solve_pipe (kernels_static.cu): in iteration i the
// solver warp (warp 0; it shares SM sub-partition 0 with two helper warps only) runs both block substitutions of item
// i-1, the nine worker warps and the helpers turn item i-2's solution into its three outputs, then the workers pull item
// i in.  384 threads at 168 registers: every loop of this kernel is kept free of spills, which cost it more than
// anything else that was measured (a spilled value in a unit loop: 5 to 10 % of the kernel).
//   * Outputs.  Reactions and beam forces are built as 1536-byte chunks in a warp's staging slots with 16-byte shared
//     stores and handed to the bulk-copy engine (fence.proxy.async, one elected lane, two slots per warp so that a pair
//     drains while the next is built); displacements go out as 16-byte streaming stores (a second path into L2).
//   * Reactions are written ONCE: R = -p + K[r,:] d (model.go:470-489) is formed in the slot.  Non-zero loads on supported
//     DOFs (almost always none) travel in a side list from the load stage; p is read again only if that list overflows.
//     A warp keeps to one row chunk: its supported DOFs, their positions and reaction-row entries are read once per item
//     and the slots need their zeros once (a unit overwrites exactly the positions the unit before it wrote).
//   * Inputs.  The load cases arrive by bulk copies global -> shared into a ring of three buffers per worker warp (its
//     two staging slots, idle in that stage, and a landing slot; one mbarrier each): no load instruction and no register
//     holds data in flight, the scatter into the right-hand sides is straight-line integer code with one vote per chunk
//     for the rare cases.  Factor and reaction block arrive by one bulk copy each, issued by a helper thread.
//   * No L2 prefetch and no start offset between blocks: with the copies above both measured slower or neutral.
// Units of work are flat: chunk u of an output covers elements [192 u, 192 u + 192) of the item's run (forces) or
// (case, row chunk) pairs; a unit handles a case of the first half of the load cases and its partner from the second
// half, so that every table entry, coefficient and address is fetched once for two chunks.
//
// Needs dof even and 16-byte aligned arrays (the chunks must start on 16 bytes); otherwise, or when the model does not
// fit, launch_solve_pipe keeps k_solve_pipe / k_solve_general.  profiles/experiments/ has the variants that were
// measured and dropped.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "bulk_ops.cuh"

namespace hdb {

constexpr int BULK_THREADS = 384;                  // warp 0 = solver, warps 4 and 8 = helpers (its sub-partition), nine workers
constexpr int NW = BULK_THREADS / 32 / 4 * 3;      // worker warps
constexpr int NH_WARPS = BULK_THREADS / 32 / 4 - 1; // helper warps (the solver's sub-partition)
constexpr int NWT = NW * 32;                       // worker threads
constexpr int CHUNK = BULK_CHUNK;                  // doubles per staging slot = 1536 bytes = 32 beams x 6 = 64 points x 3
constexpr unsigned CHUNK_BYTES = CHUNK * 8u;
constexpr unsigned SLOT_BYTES = 2u * CHUNK_BYTES;  // two slots per worker warp
constexpr unsigned NOFAIL = NO_FAIL_CASE;
constexpr int NU = NW + NH_WARPS;                    // unit warps: workers + helpers
constexpr int PCAP = 64;                             // non-zero loads on supported DOFs kept per item
constexpr int NH = (BULK_THREADS / 32 / 4 - 1) * 32; // helper threads: warps 4, 8
__device__ __forceinline__ void bar_tables() { asm volatile("bar.sync 3, %0;" ::"r"(NWT + NH) : "memory"); } // workers + helpers

struct BulkCfg {
    SolveCfg s;
    unsigned magicNB; // floor(2^32 / nB) + 1: q = umulhi(e, magic) is e / nB for e * nB < 2^32; 0 = divide by one
    // byte offsets of the shared-memory regions from the start of dynamic shared memory, computed on the host: as kernel
    // parameters they are constant-bank operands, not registers that live through every loop
    unsigned oDt, oBb, oRp, oRl, oEl, oPm, oGm, oLand, oSlot, oLvm, oXs, lvmBytes, xsBytes, BS, ZB;
    unsigned landBytes; // one landing slot
    unsigned magicCh; // the same for the row chunks of a case; 0 = one chunk
    int dbg; // HDB_BULK_DBG (debug instantiation only): 1 = no bulk stores, 2 = no fence, 4 = forces without arithmetic
};

static size_t align16b(size_t b) { return (b + 15) & ~(size_t)15; }

// one xs buffer: nBlk blocks of [cases][3] (+ pad) and a block of zeros, [cases][3], that the tables point non-free
// DOFs at (their displacement is 0 and they contribute nothing to a beam's end forces)
static size_t bulk_xs_doubles(const DevTopo &T, int cp) { return (size_t)T.nBlk * xs_block_doubles(cp) + (((size_t)3 * cp + 1) & ~(size_t)1); }

static size_t bulk_idx_bytes(const DevTopo &T)
{
    size_t b = (size_t)(T.nFwdOps + T.nBwdOps) * 16;
    b += align16b((size_t)T.dof * 4);                              // DOF table
    b += align16b((size_t)T.nB * 24);                              // beam rows (6 x int32)
    b += align16b((size_t)(T.rNCh + 1) * 4);                       // supported DOFs per row chunk: pointers,
    b += align16b((size_t)std::max(T.nSup, 1) * 4);                //   entries,
    b += align16b((size_t)std::max(T.nSup * T.rWidth, 1) * 4);     //   fixed-width reaction rows
    b += align16b((size_t)T.nB * 4);                               // pin masks
    b += align16b((size_t)T.nB * 24);                              // geometry
    b += (size_t)NW * align16b((size_t)T.rChunk * 8);              // landing slots: the load stage's third buffer per worker warp, one row chunk
    b += (size_t)NU * SLOT_BYTES;                                  // staging slots
    return b;
}

static size_t bulk_lvm_doubles(const DevTopo &T)
{
    size_t d = std::max<size_t>((size_t)T.nSlots, (size_t)T.nB * 4 + (size_t)T.nnzR + 4); // + the odd doubles of the 16-byte bulk copy and the padding zero
    return (d + 1) & ~(size_t)1;
}

static size_t bulk_smem_bytes(const DevTopo &T, int cp)
{
    return bulk_idx_bytes(T) + 2 * (bulk_lvm_doubles(T) + bulk_xs_doubles(T, cp)) * sizeof(double);
}

bool solve_bulk_config(const DevTopo &T, int nLC, int smemOptin, int &casesPerPass, size_t &smemBytes)
{
    if (T.dof % 2 || T.dof >= 65536 || T.rWidth > 8 || T.nnzR >= 65534 || T.nSlots % 2 || T.nSup >= 65535 || T.rChunk <= 0) return false;
    if (bulk_xs_doubles(T, 32) >= 65535) return false; // reaction rows keep xs offsets in 16 bits (units of 8 bytes)
    int cp = nLC < 32 ? nLC : 32;
    if (const char *e = std::getenv("HDB_CASES_PER_PASS")) cp = std::min(cp, atoi(e));
    if (cp < 1) cp = 1;
    for (; cp >= 1; cp--) {
        const size_t bytes = bulk_smem_bytes(T, cp);
        if (bytes + 512 <= (size_t)smemOptin) {
            if ((unsigned long long)cp * T.nB * (unsigned long long)T.nB >= (1ull << 32)) return false; // multiply-high division
            casesPerPass = cp;
            smemBytes = bytes;
            return true;
        }
    }
    return false;
}

// Work of the worker warps comes in PAIRS of chunks: the load cases of an item are split into a first half A and a
// second half B, and a warp always handles the same rows / beams of case c (slot A) and case c + halfC (slot B), so
// every table entry, coefficient and address is fetched once for two chunks.
template <bool DBG>
__global__ void __launch_bounds__(BULK_THREADS, 1) k_solve_bulk(const DevTopo T, const StaticLaunch L, const BulkCfg cfg)
{
    extern __shared__ __align__(16) double smem_d[];
    __shared__ unsigned s_failNF[2], s_failIG[2];
    // loads on SUPPORTED DOFs enter the reactions only (model.go:478) and are almost always zero: the load stage records
    // the non-zero ones here, (case << 16 | DOF, value), so that the recovery stage need not read p again.  More than PCAP
    // of them in one item: its reactions read p from global memory a second time.
    __shared__ __align__(8) unsigned long long s_mbar[NW][3]; // a worker warp's three landing buffers: the bulk copy has arrived
    __shared__ __align__(8) unsigned long long s_mbarKF[2];   // the reaction block / the factor of an item has arrived
    __shared__ unsigned s_pcnt[2], s_pkey[2][PCAP];
    __shared__ double s_pval[2][PCAP];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0); // broadcast: the compiler then knows it is warp-uniform (bulk-copy operands live in uniform registers)
    const bool worker = (warp & 3) != 0;
    const int sw = warp - (warp >> 2) - 1;        // worker warp index 0..8
    const int wt = sw * 32 + lane;                // worker thread index
    const int hx = ((warp >> 2) - 1) * 32 + lane; // helper thread index (warps 4, 8)
    const int dof = T.dof, nB = T.nB, CSP = cfg.s.casesPerPass, RW = T.rWidth, nCh = T.rNCh, chD = T.rChunk;
    const unsigned smA = (unsigned)__cvta_generic_to_shared(smem_d);
    const unsigned opsA = smA;
    const unsigned dtA = smA + cfg.oDt;     // [dof] DOF table, see below
    const unsigned bbA = smA + cfg.oBb;     // [nB][6] xs offset of a beam's DOFs
    const unsigned rpA = smA + cfg.oRp;     // [nCh + 1] supported DOFs of a row chunk
    const unsigned rlA = smA + cfg.oRl;     // [nSup] position in the chunk | index << 16
    const unsigned elA = smA + cfg.oEl;     // [nSup][RW] xs offset / 8 | value index << 16
    const unsigned pmA = smA + cfg.oPm;     // [nB] pin masks
    const unsigned gmA = smA + cfg.oGm;     // [3][nB] cos, sin, length
    const unsigned slot0 = smA + cfg.oSlot; // staging slots, two per worker warp
    const unsigned lvm0 = smA + cfg.oLvm;   // two factor / table buffers
    const unsigned xs0 = smA + cfg.oXs;     // two right-hand side / solution buffers
    const unsigned lvmBytes = cfg.lvmBytes, xsBytes = cfg.xsBytes;
    const unsigned BS = cfg.BS;             // bytes of one block of xs: [cases][3] (+ pad)
    const unsigned ZB = cfg.ZB;             // the block of zeros of an xs buffer

    // ---- once per block: integer tables -> shared memory ----
    auto xoff = [&](int r) { return r >= 0 ? (int)((unsigned)(r / 3) * BS + (unsigned)(r % 3) * 8u) : (int)ZB; };
    for (int k = tid; k < T.nFwdOps + T.nBwdOps; k += blockDim.x) {
        const int4 r = k < T.nFwdOps ? __ldg(T.fwdOps + k) : __ldg(T.bwdOps + (k - T.nFwdOps));
        asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(opsA + 16u * k), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
    }
    // DOF table: xs byte offset of the DOF's value (a multiple of 8; the zero block for a DOF that is not free), bit 0 =
    // not free, bit 1 = no stiffness (its load must be zero, model.go:424-437), bit 2 = supported
    for (int g = tid; g < dof; g += blockDim.x) {
        const int xr = __ldg(T.xrow + g);
        sts_s32(dtA + 4u * g, xoff(xr) | (xr >= 0 ? 0 : 1) | (xr == -2 ? 2 : 0) | (__ldg(T.supIndex + g) >= 0 ? 4 : 0));
    }
    for (int k = tid; k < nB * 6; k += blockDim.x) sts_s32(bbA + 4u * k, xoff(__ldg(T.beamRows + k)));
    for (int k = tid; k <= nCh; k += blockDim.x) sts_s32(rpA + 4u * k, __ldg(T.rsPtr + k));
    // reaction row of supported DOF si: RW entries {xs offset of the column / 8, index of K[r, col] in the model's reaction
    // block}, columns ascending (model.go:470-489); the padding multiplies a zero value with the zero block
    for (int si = tid; si < T.nSup; si += blockDim.x) {
        sts_s32(rlA + 4u * si, __ldg(T.rsList + si));
        const int q0 = __ldg(T.rptr + si), q1 = __ldg(T.rptr + si + 1);
        for (int w = 0; w < RW; w++) {
            const int q = q0 + w;
            const unsigned e = q < q1 ? (((unsigned)xoff(__ldg(T.rcol + q)) >> 3) | ((unsigned)q << 16)) : ((ZB >> 3) | ((unsigned)(T.nnzR + 1) << 16)); // value nnzR + 1: a zero behind the block's bulk copy
            sts_s32(elA + 4u * (unsigned)(si * RW + w), (int)e);
        }
    }
    for (int b = tid; b < nB; b += blockDim.x) {
        const BeamGeom g = beam_geom(T, b);
        sts_f64(gmA + (unsigned)b * 8u, g.c);
        sts_f64(gmA + (unsigned)(nB + b) * 8u, g.s);
        sts_f64(gmA + (unsigned)(2 * nB + b) * 8u, g.length);
        sts_s32(pmA + 4u * b, (int)T.pinMask[b]);
    }
    for (int k = tid; k < 2 * ((3 * CSP + 1) & ~1); k += blockDim.x) { // the zero blocks of both xs buffers
        const int zb = (3 * CSP + 1) & ~1;
        sts_f64(xs0 + (unsigned)(k / zb) * xsBytes + ZB + 8u * (unsigned)(k % zb), 0.0);
    }
    if (tid < 3 * NW) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[0][0]) + 8u * (unsigned)tid) : "memory");
    if (tid < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbarKF[tid])) : "memory");
    if (tid < 2) {
        s_failNF[tid] = NOFAIL;
        s_failIG[tid] = NOFAIL;
        s_pcnt[tid] = 0;
    }
    __syncthreads();

    const int nMy = (L.nModels - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nIt = nMy * cfg.s.nPass;
    long long tSolve = 0, tAll = DBG ? clock64() : 0;
    long long tPh[DBG ? 16 : 1] = {0};
#define TICK(k)                                  \
    do {                                         \
        if (DBG) {                               \
            const long long t1_ = clock64();     \
            tPh[k] += t1_ - t0;                  \
            t0 = t1_;                            \
        }                                        \
    } while (0)

    const int uw = worker ? sw : NW + (warp >> 2) - 1; // unit-warp index: workers, then helpers
    const unsigned slotA = slot0 + (unsigned)uw * SLOT_BYTES, slotB = slotA + CHUNK_BYTES;
    // hand both staging slots of this warp to the bulk-copy engine: generic-proxy writes ordered before the async-proxy
    // reads, then one elected lane issues the copies
    auto commit_pair = [&](double *gA, unsigned bytesA, double *gB, unsigned bytesB) {
        if (!(DBG && (cfg.dbg & 2))) fence_async_smem();
        __syncwarp();
        if (elect_one() && !(DBG && (cfg.dbg & 1))) {
            const unsigned long long polOut = l2_policy_evict_first(); // (made here: two registers less across every loop of the kernel)
            bulk_store(gA, slotA, bytesA, polOut);
            if (bytesB) bulk_store(gB, slotB, bytesB, polOut);
            bulk_commit();
        }
    };
    auto slots_free = [&]() { // the copies of the previous pair have read their slots
        if (elect_one()) bulk_wait_read<0>();
        __syncwarp();
    };

    unsigned landPhase = 0; // parity bits of this warp's three landing buffers
    const unsigned mbarK = (unsigned)__cvta_generic_to_shared(&s_mbarKF[0]), mbarF = mbarK + 8u;
    // reactions: unit warp uw keeps to row chunk rJ0 (chunks rJ0, rJ0 + NU, ... when there are more chunks than unit warps) and
    // takes the case pairs rC0, rC0 + rCStep, ... there
    const bool rSpread = NU >= nCh;
    const int rJ0 = rSpread ? uw % nCh : uw, rJStep = rSpread ? nCh : NU;
    const int rC0 = rSpread ? uw / nCh : 0, rCStep = rSpread ? (NU - rJ0 + nCh - 1) / nCh : 1;
    for (int i = 0; i < nIt + 2; i++) {
        const unsigned buf = (unsigned)(i & 1);
        const unsigned lvB = lvm0 + buf * lvmBytes, xsB = xs0 + buf * xsBytes;
        const unsigned cs = (unsigned)nB * 8u;
        if (tid < 32) {
            // ================= solver warp: item i-1 =================
            if (i >= 1 && i <= nIt) {
                const int it = i - 1, pass = it % cfg.s.nPass;
                const int nc = min(CSP, L.nLC - pass * CSP);
                const unsigned sbuf = (unsigned)(it & 1);
                if (tid < nc) {
                    mbar_wait_sleep(mbarF, (unsigned)it & 1u); // the factor's bulk copy (issued one iteration ago) has landed
                    const long long t0 = DBG ? clock64() : 0;
                    const unsigned tabA = lvm0 + sbuf * lvmBytes, xA = xs0 + sbuf * xsBytes + (unsigned)tid * 24u;
                    walk_ops<true>(opsA, tabA, xA, BS);
                    walk_ops<false>(opsA + 16u * (unsigned)T.nFwdOps, tabA, xA, BS);
                    if (DBG) tSolve += clock64() - t0;
                }
            }
        } else {
            // ================= unit warps: the workers and, for the reactions and the forces, the helper warps =================
            // (the helpers sit in the solver's sub-partition: they get shared-memory work only, apart from fetching the
            // reaction block and the factor)
            long long t0 = DBG ? clock64() : 0;
            if (i >= 2) {
                // ---------- recover item i-2 ----------
                const int it = i - 2, pass = it % cfg.s.nPass;
                const int model = (int)blockIdx.x + (it / cfg.s.nPass) * (int)gridDim.x;
                const int c0 = pass * CSP, nc = min(CSP, L.nLC - c0);
                const int halfC = (nc + 1) >> 1;
                const unsigned offB = (unsigned)halfC * 24u; // slot B's case in xs
                const unsigned ksShift = (unsigned)((reinterpret_cast<uintptr_t>(L.Ksf + (size_t)model * T.nnzR) >> 3) & 1);
                const unsigned ksA = lvB + 4u * cs + 8u * ksShift; // the reaction block in the dead factor buffer, behind the coefficient table
                const int nRowU = halfC * nCh;               // row-chunk pairs of a [cases][dof] output
                const size_t rowBase = ((size_t)model * L.nLC + c0) * dof;
                const int stepC = NW / nCh, stepJ = NW - stepC * nCh, cStart = sw / nCh, jStart = sw - cStart * nCh; // unit u = (case u / nCh, chunk u % nCh), walked without divisions
                const int stepCU = NU / nCh, stepJU = NU - stepCU * nCh, cStartU = uw / nCh, jStartU = uw - cStartU * nCh; // the same walk over all unit warps
                // this thread's beams (one for a frame of up to 32 * NW beams): section loads issued now, used before the forces
                double sc0 = 1.0, sc1 = 1.0, sc2 = 1.0;
                if (worker && L.forces && wt < nB) {
                    const double *sec = L.sections + ((size_t)model * nB + wt) * 3;
                    sc0 = __ldg(sec); sc1 = __ldg(sec + 1); sc2 = __ldg(sec + 2);
                }
                if (!worker) {
                    // The factor buffer is dead once the solve is over: it takes the reaction block K[sup, free] of the item (helpers,
                    // while the workers write the displacements) and its beam coefficients [4][nB] (workers, before the forces).
                    const double *ksf = L.Ksf + (size_t)model * T.nnzR;
                    // the reaction block: one bulk copy (it starts on 16 bytes, so the block may sit one double into its buffer);
                    // the zero that the padding of the reaction rows points at lies behind the copied range
                    if (L.react && hx == 0) {
                        bulk_load(ksA & ~15u, ksf - ksShift, (((unsigned)T.nnzR + ksShift + 1u) & ~1u) * 8u, mbarK);
                        sts_f64(ksA + 8u * (unsigned)(T.nnzR + 1), 0.0);
                    }
                    if (DBG && hx == 0) tPh[0] += clock64() - t0;
                }
                const unsigned pCnt = s_pcnt[buf]; // non-zero loads on supported DOFs of this item (read before bar_tables, reset after it)
                // ---- displacements (model.go:446-451) ----
                if (L.disp) // (the helper warps too: their copy of the reaction block is under way)
                    for (int u = uw, cA = cStartU, j = jStartU; u < nRowU; u += NU, cA += stepCU, j += stepJU) {
                        if (j >= nCh) { j -= nCh; cA++; }
                        const int g0 = j * chD, nE = min(chD, dof - g0);
                        const bool hasB = cA + halfC < nc;
                        const unsigned xA = xsB + (unsigned)cA * 24u, oB = hasB ? offB : 0u; // no second case: read the first again
                        double dA[3][2], dB[3][2];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int p2 = 2 * (lane + 32 * k);
                            int t0_, t1_;
                            lds_v2s32(dtA + 4u * (unsigned)(g0 + (p2 < nE ? p2 : 0)), t0_, t1_);
                            const unsigned a0 = xA + ((unsigned)t0_ & ~7u), a1 = xA + ((unsigned)t1_ & ~7u);
                            dA[k][0] = lds_f64(a0); dA[k][1] = lds_f64(a1);
                            dB[k][0] = lds_f64(a0 + oB); dB[k][1] = lds_f64(a1 + oB);
                        }
                        // a lane's pair of doubles is 16 contiguous bytes of the row and the warp's 32 pairs are 512 contiguous bytes:
                        // these go out as plain 16-byte streaming stores, through the load/store unit, next to the bulk copies of the
                        // other outputs (two paths into L2 instead of one)
                        double *gA = L.disp + rowBase + (size_t)cA * dof + g0 + 2 * lane, *gB = gA + (size_t)halfC * dof;
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            if (2 * (lane + 32 * k) < nE) {
                                __stcs(reinterpret_cast<double2 *>(gA + 64 * k), make_double2(dA[k][0], dA[k][1]));
                                if (hasB) __stcs(reinterpret_cast<double2 *>(gB + 64 * k), make_double2(dB[k][0], dB[k][1]));
                            }
                    }
                if (L.forces) {
                    // beam coefficients of the model (element.cuh; the divisions are why they are tabulated once per item)
                    for (int b = worker ? wt : nB; b < nB; b += NWT) {
                        if (L.geom) { // coordinates vary per model: this model's row replaces the shared geometry table
                            const BeamGeom g = beam_geom(T, L, model, b);
                            sts_f64(gmA + (unsigned)b * 8u, g.c);
                            sts_f64(gmA + (unsigned)(nB + b) * 8u, g.s);
                            sts_f64(gmA + (unsigned)(2 * nB + b) * 8u, g.length);
                        }
                        const double *sec = L.sections + ((size_t)model * nB + b) * 3;
                        const BeamCoef cf = b == wt ? beam_coef(sc0, sc1, sc2, lds_f64(gmA + (unsigned)(2 * nB + b) * 8u))
                                                    : beam_coef(__ldg(sec), __ldg(sec + 1), __ldg(sec + 2), lds_f64(gmA + (unsigned)(2 * nB + b) * 8u));
                        const unsigned a = lvB + (unsigned)b * 8u; // component-major: conflict-free
                        sts_f64(a, cf.EFL); sts_f64(a + cs, cf.EJL3); sts_f64(a + 2u * cs, cf.EJL2); sts_f64(a + 3u * cs, cf.EJL);
                    }
                }
                TICK(0);
                if (L.react) mbar_wait_sleep(mbarK, (unsigned)i & 1u); // the reaction block has landed (every thread observes the copy itself)
                bar_tables(); // the helpers' reaction block and the workers' coefficient table are complete; everybody has read the side list's count
                TICK(1);
                if (!worker && hx == 0) s_pcnt[buf] = 0; // the load stage of item i starts a new list after the next barrier
                // ---- reactions R = -p + K[r, free] d (model.go:470-489), written once: zero rows, then K d on the supported DOFs
                //      of the chunk, then -p where p is not zero (side list; read from global memory only if it overflowed) ----
                if (L.react) {
                    // A warp keeps to ONE row chunk j (rJ0; several when there are more chunks than warps) and takes its share of
                    // the case pairs there: the chunk's supported DOFs, their positions and the entries of their reaction rows
                    // are read once per item instead of once per unit, and the staging slots need their zeros only once (a unit
                    // overwrites exactly the positions the unit before it wrote).
                    const bool reread = pCnt > (unsigned)PCAP;
                    for (int j = rJ0; j < nCh; j += rJStep) {
                        const int g0 = j * chD, nE = min(chD, dof - g0);
                        const int tBeg = lds_s32(rpA + 4u * (unsigned)j), tEnd = lds_s32(rpA + 4u * (unsigned)j + 4u);
                        unsigned le[3], rp_[3];
#pragma unroll
                        for (int r = 0; r < 3; r++) {
                            const int t = tBeg + lane + 32 * r;
                            le[r] = (unsigned)lds_s32(rlA + 4u * (unsigned)(t < tEnd ? t : tBeg));
                        }
#pragma unroll
                        for (int r = 0; r < 3; r++) rp_[r] = tBeg + lane + 32 * r < tEnd ? (le[r] & 0xffffu) : 0xffffffffu;
                        unsigned row_[3]; // first entry of the DOF's reaction row
#pragma unroll
                        for (int r = 0; r < 3; r++) row_[r] = elA + 4u * (le[r] >> 16) * (unsigned)RW;
                        const bool hoist = RW <= 3; // rows of up to three entries (a column or a girder end): the entries stay in registers
                        unsigned en_[3][3];
#pragma unroll
                        for (int w = 0; w < 3; w++)
#pragma unroll
                            for (int r = 0; r < 3; r++) en_[w][r] = (hoist && w < RW) ? (unsigned)lds_s32(row_[r] + 4u * (unsigned)w) : 0u;
                        bool clean = false; // the slots hold zeros wherever this chunk has no supported DOF
                        for (int cA = rC0; cA < halfC; cA += rCStep) {
                            const bool hasB = cA + halfC < nc;
                            const double *pA = L.loads + rowBase + (size_t)cA * dof + g0;
                            const double *pB = pA + (size_t)(hasB ? halfC : 0) * dof;
                            const unsigned xA = xsB + (unsigned)cA * 24u, oB = hasB ? offB : 0u;
                            // K d of up to 96 supported DOFs of the chunk first, into registers: the previous pair's copies drain meanwhile
                            double rA[3] = {0.0, 0.0, 0.0}, rB[3] = {0.0, 0.0, 0.0};
                            auto react_item = [&](int t, double &oa, double &ob) {
                                const unsigned e = (unsigned)lds_s32(rlA + 4u * (unsigned)t);
                                const unsigned pos = e & 0xffffu, si = e >> 16;
                                double vA = 0.0, vB = 0.0;
                                if (reread) { vA = __ldcs(pA + pos); vB = __ldcs(pB + pos); }
                                double accA = 0.0, accB = 0.0;
                                for (int w = 0; w < RW; w++) {
                                    const unsigned en = (unsigned)lds_s32(elA + 4u * (si * (unsigned)RW + (unsigned)w));
                                    const double kv = lds_f64(ksA + (en >> 16) * 8u);
                                    const unsigned xa = xA + (en & 0xffffu) * 8u;
                                    accA = fma(kv, lds_f64(xa), accA);
                                    accB = fma(kv, lds_f64(xa + oB), accB);
                                }
                                oa = -vA + accA; // -(+0) + acc is acc, bit for bit
                                ob = -vB + accB;
                                return pos;
                            };
                            if (tEnd > tBeg && hoist) {
#pragma unroll
                                for (int w = 0; w < 3; w++)
                                    if (w < RW) {
                                        double kv[3], x0[3], x1[3];
#pragma unroll
                                        for (int r = 0; r < 3; r++) {
                                            kv[r] = lds_f64(ksA + (en_[w][r] >> 16) * 8u);
                                            const unsigned xa = xA + (en_[w][r] & 0xffffu) * 8u;
                                            x0[r] = lds_f64(xa);
                                            x1[r] = lds_f64(xa + oB);
                                        }
#pragma unroll
                                        for (int r = 0; r < 3; r++) {
                                            rA[r] = fma(kv[r], x0[r], rA[r]);
                                            rB[r] = fma(kv[r], x1[r], rB[r]);
                                        }
                                    }
                            } else if (tEnd > tBeg)
                                for (int w = 0; w < RW; w++) { // the three rounds level by level (the shared-memory loads are ordered asm statements)
                                    unsigned en[3];
#pragma unroll
                                    for (int r = 0; r < 3; r++) en[r] = (unsigned)lds_s32(row_[r] + 4u * (unsigned)w);
                                    double kv[3], x0[3], x1[3];
#pragma unroll
                                    for (int r = 0; r < 3; r++) {
                                        kv[r] = lds_f64(ksA + (en[r] >> 16) * 8u);
                                        const unsigned xa = xA + (en[r] & 0xffffu) * 8u;
                                        x0[r] = lds_f64(xa);
                                        x1[r] = lds_f64(xa + oB);
                                    }
#pragma unroll
                                    for (int r = 0; r < 3; r++) {
                                        rA[r] = fma(kv[r], x0[r], rA[r]);
                                        rB[r] = fma(kv[r], x1[r], rB[r]);
                                    }
                                }
                            if (reread) { // the side list overflowed: p is read a second time (without a load the sum is the result as it stands)
#pragma unroll
                                for (int r = 0; r < 3; r++)
                                    if (rp_[r] != 0xffffffffu) {
                                        rA[r] = -__ldcs(pA + rp_[r]) + rA[r];
                                        rB[r] = -__ldcs(pB + rp_[r]) + rB[r];
                                    }
                            }
                            slots_free();
                            if (!clean) {
#pragma unroll
                                for (int k = 0; k < 3; k++) {
                                    sts_v2f64(slotA + 16u * (unsigned)(lane + 32 * k), 0.0, 0.0);
                                    sts_v2f64(slotB + 16u * (unsigned)(lane + 32 * k), 0.0, 0.0);
                                }
                                __syncwarp();
                                clean = true;
                            }
#pragma unroll
                            for (int r = 0; r < 3; r++)
                                if (rp_[r] != 0xffffffffu) {
                                    sts_f64(slotA + 8u * rp_[r], rA[r]);
                                    sts_f64(slotB + 8u * rp_[r], rB[r]);
                                }
                            for (int t = tBeg + lane + 96; t < tEnd; t += 32) { // a chunk with more than 96 supported DOFs
                                double oa, ob;
                                const unsigned pos = react_item(t, oa, ob);
                                sts_f64(slotA + 8u * pos, oa);
                                sts_f64(slotB + 8u * pos, ob);
                            }
                            if (pCnt != 0u && !reread) { // rare: non-zero loads on supported DOFs of this item
                                __syncwarp();
                                for (unsigned q = (unsigned)lane; q < pCnt; q += 32u) {
                                    const unsigned key = s_pkey[buf][q];
                                    const int c = (int)(key >> 16), g = (int)(key & 0xffffu) - g0;
                                    if (g >= 0 && g < nE && (c == cA || (hasB && c == cA + halfC))) {
                                        const unsigned a = (c == cA ? slotA : slotB) + 8u * (unsigned)g;
                                        sts_f64(a, -s_pval[buf][q] + lds_f64(a));
                                    }
                                }
                            }
                            double *gA = L.react + rowBase + (size_t)cA * dof + g0;
                            commit_pair(gA, (unsigned)nE * 8u, gA + (size_t)halfC * dof, hasB ? (unsigned)nE * 8u : 0u);
                        }
                    }
                }
                TICK(2);
                // ---- beam end forces s = k (T Zo), model.go:455-468: lane = (case, beam) item of the first half of the cases
                //      and the same beam in the second half, 48 bytes each ----
                if (L.forces) {
                    double *out = L.forces + ((size_t)model * L.nLC + c0) * nB * 6;
                    const int nItA = halfC * nB, nItB = (nc - halfC) * nB, nFU = (nItA + 31) / 32;
                    for (int u = uw; u < nFU; u += NU) {
                        const int f = u * 32 + lane;
                        const int fl = f < nItA ? f : 0;
                        const int c = cfg.magicNB ? (int)__umulhi((unsigned)fl, cfg.magicNB) : fl, b = fl - c * nB; // magic 0: one beam
                        const unsigned a = lvB + (unsigned)b * 8u;
                        BeamCoef cf;
                        cf.EFL = lds_f64(a); cf.EJL3 = lds_f64(a + cs); cf.EJL2 = lds_f64(a + 2u * cs); cf.EJL = lds_f64(a + 3u * cs);
                        const unsigned ga = gmA + (unsigned)b * 8u;
                        const double gc = lds_f64(ga), gs = lds_f64(ga + cs);
                        const uint32_t pm = (uint32_t)lds_s32(pmA + 4u * (unsigned)b);
                        int r[6];
                        lds_v2s32(bbA + 24u * (unsigned)b, r[0], r[1]);
                        lds_v2s32(bbA + 24u * (unsigned)b + 8u, r[2], r[3]);
                        lds_v2s32(bbA + 24u * (unsigned)b + 16u, r[4], r[5]);
                        const unsigned xA = xsB + (unsigned)c * 24u, oBc = c + halfC < nc ? offB : 0u; // no second case: read the first again
                        double uA[6], uB[6];
#pragma unroll
                        for (int e = 0; e < 6; e++) {
                            uA[e] = lds_f64(xA + (unsigned)r[e]);
                            uB[e] = lds_f64(xA + (unsigned)r[e] + oBc);
                        }
                        cf.EJL4 = __dmul_rn(2.0, cf.EJL); // as beam_coef (element.cuh)
                        slots_free();
                        const unsigned sl = 48u * (unsigned)lane;
#pragma unroll
                        for (int h = 0; h < 2; h++) { // one half at a time: its six results leave for the slot before the other half's exist
                            double o[6];
                            if (DBG && (cfg.dbg & 4)) {
#pragma unroll
                                for (int e = 0; e < 6; e++) o[e] = h ? uB[e] : uA[e];
                            } else
                                beam_forces_rigid(cf, gc, gs, h ? uB : uA, o);
                            const unsigned a_ = (h ? slotB : slotA) + sl;
                            sts_v2f64(a_, o[0], o[1]);
                            sts_v2f64(a_ + 16u, o[2], o[3]);
                            sts_v2f64(a_ + 32u, o[4], o[5]);
                        }
                        // released ends are rare: only a warp that has one enters the out-of-line general element, and only its
                        // released beams take its result (a beam's forces do not depend on which warp it fell into)
                        if (__any_sync(0xffffffffu, pm != 0u)) {
                            if (pm != 0u) {
                                pinned_forces_to_slot(cf, pm, gc, gs, uA[0], uA[1], uA[2], uA[3], uA[4], uA[5], slotA + sl);
                                pinned_forces_to_slot(cf, pm, gc, gs, uB[0], uB[1], uB[2], uB[3], uB[4], uB[5], slotB + sl);
                            }
                        }
                        const int liveA = min(32, nItA - u * 32), liveB = max(0, min(32, nItB - u * 32));
                        commit_pair(out + (size_t)u * CHUNK, (unsigned)liveA * 48u, out + (size_t)nItA * 6 + (size_t)u * CHUNK, (unsigned)liveB * 48u);
                    }
                }
                TICK(3);
                slots_free();
                TICK(4);
                bar_stream(NWT + NH); // every reader of this buffer is done: the load stage may overwrite it
                TICK(5);
            }
            if (i < nIt && !worker) {
                // factor of item i -> shared memory, dummy unknowns of its right-hand sides = 0, while the workers scatter its loads
                const int pass = i % cfg.s.nPass;
                const int model = (int)blockIdx.x + (i / cfg.s.nPass) * (int)gridDim.x;
                const int nc = min(CSP, L.nLC - pass * CSP);
                if (hx == 0) bulk_load(lvB, L.Lm + (size_t)model * T.nSlots, (unsigned)T.nSlots * 8u, mbarF); // one bulk copy (block columns are 6 + 10 k slots: whole 16-byte pairs)
                for (int idx = hx; idx < T.nDummy * nc; idx += NH) {
                    const int q = idx / nc, c = idx - q * nc;
                    const int r = __ldg(T.dummyPos + q);
                    sts_f64(xsB + (unsigned)(r / 3) * BS + (unsigned)(r % 3) * 8u + (unsigned)c * 24u, 0.0);
                }
            }
            if (i < nIt && worker) {
                // ---------- load item i: load cases -> xs (free rows); checks of model.go:692-696 and 424-437 ----------
                const int it = i, pass = it % cfg.s.nPass;
                const int model = (int)blockIdx.x + (it / cfg.s.nPass) * (int)gridDim.x;
                const int c0 = pass * CSP, nc = min(CSP, L.nLC - c0);
                const int halfC = (nc + 1) >> 1;
                const int nRowU = halfC * nCh;
                int sOld = 0;
                if (wt == 0) asm volatile("ld.global.s32 %0, [%1];" : "=r"(sOld) : "l"(L.status + model) : "memory");
                const double *pb = L.loads + ((size_t)model * L.nLC + c0) * dof;
                unsigned failNF = NOFAIL, failIG = NOFAIL;
                // Row chunk u = (case u / nCh, chunk u % nCh), warp sw takes u = sw, sw + NW, ...  The chunks arrive by bulk copies
                // global -> shared (no load instruction, no register: the load/store unit's queue of outstanding misses is what
                // limits a block's plain loads) into a ring of three buffers per warp - its two staging slots, idle in this
                // stage, and a landing slot - so that two or three chunks per warp are in flight while one is scattered.
                const int nSU = nc * nCh;
                const int stepC = NW / nCh, stepJ = NW - stepC * nCh;
                const unsigned mb0 = (unsigned)__cvta_generic_to_shared(&s_mbar[sw][0]);
                const unsigned ring[3] = {slotA, slotB, smA + cfg.oLand + (unsigned)sw * cfg.landBytes};
                int uI = sw, cI = cfg.magicCh ? (int)__umulhi((unsigned)sw, cfg.magicCh) : sw, jI = sw - cI * nCh; // next chunk to request
                auto request = [&](int r) {
                    if (uI < nSU && elect_one()) {
                        const int g0 = jI * chD;
                        bulk_load(ring[r], pb + (size_t)cI * dof + g0, (unsigned)min(chD, dof - g0) * 8u, mb0 + 8u * (unsigned)r);
                    }
                    uI += NW; cI += stepC; jI += stepJ;
                    if (jI >= nCh) { jI -= nCh; cI++; }
                };
                auto not_free = [&](int t, double v, int c, int g) { // rare: a DOF that is not free carries a load
                    if ((t & 2) && v != 0.0) failIG = min(failIG, (unsigned)(c0 + c)); // no stiffness (model.go:424-437)
                    if ((t & 4) && __double_as_longlong(v) != 0ll) {                  // supported: enters the reaction only
                        const unsigned q = atomicAdd(&s_pcnt[buf], 1u);
                        if (q < (unsigned)PCAP) {
                            s_pkey[buf][q] = ((unsigned)c << 16) | (unsigned)g;
                            s_pval[buf][q] = v;
                        }
                    }
                };
                int cS = cI, jS = jI; // chunk being scattered
#pragma unroll
                for (int r = 0; r < 3; r++) request(r);
                for (int u = sw; u < nSU;) {
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        if (u < nSU) {
                            mbar_wait_sleep(mb0 + 8u * (unsigned)r, (landPhase >> r) & 1u); // (try_wait: the warp sleeps until the copy has landed)
                            landPhase ^= 1u << r;
                            const int g0 = jS * chD, nE = min(chD, dof - g0);
                            const unsigned xA = xsB + (unsigned)cS * 24u;
                            // All six loads first, then straight-line integer code: a warp on its own runs dependent instructions one
                            // latency after the other.  Non-finite values are found by the largest |high word| of the lane's six values
                            // (exponent all ones), loads on DOFs that are not free by OR-ing the value bits under the DOF's flag
                            // mask; one compare and one vote per chunk decide whether the rare paths run at all.
                            int tt[3][2];
                            double a[3][2];
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const int p2 = 2 * (lane + 32 * k), pc = p2 < nE ? p2 : 0;
                                lds_v2o<0>(ring[r] + 8u * (unsigned)pc, a[k][0], a[k][1]);
                                lds_v2s32(dtA + 4u * (unsigned)(g0 + pc), tt[k][0], tt[k][1]);
                            }
                            unsigned hmax = 0u, rare = 0u;
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const bool in = 2 * (lane + 32 * k) < nE;
#pragma unroll
                                for (int h = 0; h < 2; h++) {
                                    const int t = in ? tt[k][h] : 1; // past the end of the row: not free, nothing to check
                                    const unsigned hi = (unsigned)__double2hiint(a[k][h]), lo = (unsigned)__double2loint(a[k][h]);
                                    hmax = max(hmax, in ? (hi & 0x7fffffffu) : 0u);
                                    rare |= (hi | lo) & ((t & 6) ? 0xffffffffu : 0u);
                                    if (!(t & 1)) sts_f64(xA + (unsigned)t, a[k][h]);
                                }
                            }
                            if (hmax >= 0x7ff00000u) failNF = min(failNF, (unsigned)(c0 + cS));
                            if (__any_sync(0xffffffffu, rare != 0u)) { // a DOF that is not free carries a load
#pragma unroll
                                for (int k = 0; k < 3; k++) {
                                    const int p2 = 2 * (lane + 32 * k);
#pragma unroll
                                    for (int h = 0; h < 2; h++)
                                        if (p2 < nE && (tt[k][h] & 6)) not_free(tt[k][h], a[k][h], cS, g0 + p2 + h);
                                }
                            }
                            __syncwarp(); // every lane has read the buffer: the next chunk may land in it
                            request(r);
                            u += NW; cS += stepC; jS += stepJ;
                            if (jS >= nCh) { jS -= nCh; cS++; }
                        }
                    }
                }
                if (failNF != NOFAIL) atomicMin(&s_failNF[buf], failNF);
                if (failIG != NOFAIL) atomicMin(&s_failIG[buf], failIG);
                TICK(6);
                bar_load(NWT);
                TICK(7);
                if (wt == 0) { // status of this pass; passes of a model arrive in ascending case order
                    const unsigned nf = s_failNF[buf], ig = s_failIG[buf];
                    merge_load_status(L, model, pass, sOld, nf, ig);
                    s_failNF[buf] = NOFAIL;
                    s_failIG[buf] = NOFAIL;
                }
            }
        }
        __syncthreads();
    }
    if (tid >= 32 && elect_one()) bulk_wait_all(); // every bulk store of this warp has been written
    if (DBG) {
        long long *d = L.dbgClocks + (size_t)blockIdx.x * 32;
        if (tid == 0) { d[0] = clock64() - tAll; d[1] = tSolve; d[2] = nMy; }
        if (tid == 32)
            for (int k = 0; k < 16; k++) d[8 + k] = tPh[k];
        if (tid == 128) d[24] = tPh[0]; // helper warp 4: tables
    }
#undef TICK
}

cudaError_t launch_solve_bulk(const DevTopo &T, const StaticLaunch &L, cudaStream_t st, int optin, int nSM, bool &taken)
{
    taken = false;
    if (const char *e = std::getenv("HDB_SOLVE_LEGACY"))
        if (atoi(e)) return cudaSuccess; // diagnostics / tests: k_solve_pipe
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!(al16(L.loads) && al16(L.disp) && al16(L.react) && al16(L.forces) && al16(L.Lm))) return cudaSuccess;
    BulkCfg cfg;
    size_t smem = 0;
    if (!solve_bulk_config(T, L.nLC, optin, cfg.s.casesPerPass, smem)) return cudaSuccess;
    cfg.s.nPass = (L.nLC + cfg.s.casesPerPass - 1) / cfg.s.casesPerPass;
    if (cfg.s.nPass < 1) cfg.s.nPass = 1;
    cfg.s.lvmDoubles = (int)bulk_lvm_doubles(T);
    cfg.s.vec2 = 1;
    cfg.s.xsDoubles = (int)bulk_xs_doubles(T, cfg.s.casesPerPass);
    {
        auto al = [](size_t b) { return (unsigned)((b + 15) & ~(size_t)15); };
        unsigned o = (unsigned)(T.nFwdOps + T.nBwdOps) * 16u;
        cfg.oDt = o; o += al((size_t)T.dof * 4);
        cfg.oBb = o; o += al((size_t)T.nB * 24);
        cfg.oRp = o; o += al((size_t)(T.rNCh + 1) * 4);
        cfg.oRl = o; o += al((size_t)std::max(T.nSup, 1) * 4);
        cfg.oEl = o; o += al((size_t)std::max(T.nSup * T.rWidth, 1) * 4);
        cfg.oPm = o; o += al((size_t)T.nB * 4);
        cfg.oGm = o; o += al((size_t)T.nB * 24);
        cfg.landBytes = al((size_t)T.rChunk * 8);
        cfg.oLand = o; o += (unsigned)NW * cfg.landBytes;
        cfg.oSlot = o; o += (unsigned)NU * SLOT_BYTES;
        cfg.lvmBytes = (unsigned)cfg.s.lvmDoubles * 8u;
        cfg.xsBytes = (unsigned)cfg.s.xsDoubles * 8u;
        cfg.oLvm = o; o += 2u * cfg.lvmBytes;
        cfg.oXs = o; o += 2u * cfg.xsBytes;
        cfg.BS = 8u * (unsigned)xs_block_doubles(cfg.s.casesPerPass);
        cfg.ZB = (unsigned)T.nBlk * cfg.BS;
        if ((size_t)o != smem) return cudaErrorInvalidConfiguration; // the two layouts (bulk_idx_bytes) disagree
    }
    cfg.magicNB = T.nB > 1 ? (unsigned)((1ull << 32) / (unsigned)T.nB) + 1u : 0u;
    cfg.magicCh = T.rNCh > 1 ? (unsigned)((1ull << 32) / (unsigned)T.rNCh) + 1u : 0u;
    cfg.dbg = 0;
    if (const char *e = std::getenv("HDB_BULK_DBG")) cfg.dbg = atoi(e);
    auto kern = L.dbgClocks ? k_solve_bulk<true> : k_solve_bulk<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int grid = L.nModels < nSM ? L.nModels : nSM;
    if (const char *e = std::getenv("HDB_SOLVE_BLOCKS")) grid = std::max(1, std::min(grid, atoi(e)));
    if (L.dbgClocks) std::fprintf(stderr, "[hdb] k_solve_bulk: %d blocks, %zu bytes of shared memory, %d cases per pass\n", grid, smem, cfg.s.casesPerPass);
    taken = true;
    kern<<<grid, BULK_THREADS, smem, st>>>(T, L, cfg);
    return cudaGetLastError();
}

} // namespace hdb
