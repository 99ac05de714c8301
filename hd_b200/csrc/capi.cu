// capi.cu — implementation of the C ABI declared in include/hdb200.h.
//
// Host side of the batched LinearStatic / Modal path: the topology handle (symbolic phase +
// device copies of its schedules), workspace management, the chunked host<->device pipeline
// for callers that pass host memory, and CUDA-event timing of every kernel.  No torch types,
// no CPU compute path: without a device the compute entries return HDB_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include "device_topo.hpp"
#include "element.cuh"
#include "hdb_internal.hpp"
#include "modal.hpp"

namespace hdb {

static thread_local std::string g_detail;
static std::atomic<int64_t> g_launches{0};

void set_error_detail(const std::string &s) { g_detail = s; }
void count_launch(int n) { g_launches += n; }

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            char buf__[512];                                                                  \
            std::snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,  \
                          cudaGetErrorString(e__));                                           \
            set_error_detail(buf__);                                                          \
            return (e__ == cudaErrorMemoryAllocation) ? HDB_ERR_NOMEM : HDB_ERR_CUDA;          \
        }                                                                                     \
    } while (0)

// NVTX range around a phase of a C-ABI call: shows up in Nsight Systems timelines, costs nothing without a tool
struct Nvtx {
    explicit Nvtx(const char *name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
};

struct HostChunk {
    double *sec = nullptr, *loads = nullptr, *disp = nullptr, *forces = nullptr, *react = nullptr;
    int32_t *status = nullptr, *fail = nullptr;
    cudaEvent_t evIn = nullptr, evK = nullptr, evOut = nullptr;
    size_t bytes = 0;
    void *base = nullptr;
};
constexpr int NBUF = 3;

} // namespace hdb

using namespace hdb;

struct hdb_topology {
    Symbolic S;
    int device = 0;
    bool devReady = false;
    int smemOptin = 0, smCount = 0;
    DevTopo D{};
    std::vector<void *> devAllocs;
    // workspace of the device path
    // two workspace sets: the assemble/factor stream prepares sub-batch i+1 while the solve stream works on i
    double *Wbuf[2] = {nullptr, nullptr}, *Lmbuf[2] = {nullptr, nullptr}, *Ksfbuf[2] = {nullptr, nullptr}, *keScratch = nullptr;
    double *geomBuf = nullptr; // [geomModels][nB][3] per-model beam geometry (hdb_static_batch_geom*)
    size_t geomModels = 0;
    cudaStream_t prepStream = nullptr;
    cudaEvent_t evFork = nullptr, evPrep[2] = {nullptr, nullptr}, evSolve[2] = {nullptr, nullptr}, evJoin = nullptr;
    long long *dbgClocks = nullptr; // HDB_DEBUG_CLOCKS=1: per-block phase cycle counters of k_solve_pipe
    size_t wsModels = 0;
    int subBatch = 8192; // the factor kernel runs one thread per model: fewer than ~2 warps per SM leave it pure latency
    // host path
    HostChunk chunk[NBUF];
    cudaStream_t sIn = nullptr, sK = nullptr, sOut = nullptr; // copy-in, compute, copy-out
    int chunkModels = 0, chunkLC = 0;
    // linear buckling tables (built at the first hdb_buckling_batch call)
    bool bkReady = false;
    const int32_t *bkFreeDof = nullptr, *bkGPtr = nullptr, *bkGIdx = nullptr;
    const double *bkGUnit = nullptr;
    // profiling
    int profiling = 0; // 0 off, 1 per call (reset at every call), 2 accumulate until hdb_set_profiling is called again
    std::vector<cudaEvent_t> events;
    std::vector<int> evKind; // kernel kind of interval [2i, 2i+1]
    int lastKinds = 0;
    // modal scratch (unit loads, their solutions, mass tables)
    void *modalBuf = nullptr;
    size_t modalBytes = 0;
};

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
template <typename T>
static int upload(hdb_topology *t, const std::vector<T> &v, const T **out)
{
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc(&p, bytes));
    t->devAllocs.push_back(p);
    if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T *)p;
    return HDB_OK;
}

static int ensure_device(hdb_topology *t)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        set_error_detail("no CUDA device visible: libhdb200 has no CPU path");
        return HDB_ERR_NO_DEVICE;
    }
    if (t->device < 0 || t->device >= n) {
        set_error_detail("device ordinal out of range");
        return HDB_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(t->device));
    if (t->devReady) return HDB_OK;
    for (void *p : t->devAllocs) cudaFree(p); // leftovers of an upload that failed half-way
    t->devAllocs.clear();
    CUDA_TRY(cudaDeviceGetAttribute(&t->smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, t->device));
    CUDA_TRY(cudaDeviceGetAttribute(&t->smCount, cudaDevAttrMultiProcessorCount, t->device));
    const Symbolic &S = t->S;
    DevTopo &D = t->D;
    D.nP = S.nP; D.nB = S.nB; D.dof = S.dof; D.nFree = S.nFree; D.nBlk = S.nBlk; D.nPad = S.nPad; D.nSlots = S.nSlots;
    D.nSup = S.nSup; D.nnzR = S.nnzR; D.nnzK = S.nnzK; D.nDummy = (int)S.dummyPos.size(); D.maxBlkCount = S.maxBlkCount;
    D.rWidth = 0;
    for (int i = 0; i < S.nSup; i++) D.rWidth = std::max(D.rWidth, S.rptr[i + 1] - S.rptr[i]);
    D.winNeed = S.winNeed; D.nFwdOps = (int)S.fwdOps.size() / 4; D.nBwdOps = (int)S.bwdOps.size() / 4;
    // gather codes -> padded shared-memory indices
    auto pad = [](const std::vector<int32_t> &src) {
        std::vector<int32_t> o(src.size());
        for (size_t i = 0; i < src.size(); i++) o[i] = (src[i] / 36) * KE_LD + (src[i] % 36);
        return o;
    };
    std::vector<uint8_t> dropped((size_t)S.nB * 36, 0);
    { // posSlot is in triplet order (p1,p2,r1,r2); dropped[] is indexed by beam*36 + local row-major
        size_t q = 0;
        for (int b = 0; b < S.nB; b++)
            for (int p1 = 0; p1 < 2; p1++)
                for (int p2 = 0; p2 < 2; p2++)
                    for (int r1 = 0; r1 < 3; r1++)
                        for (int r2 = 0; r2 < 3; r2++, q++)
                            if (S.posSlot[q] < 0) dropped[(size_t)b * 36 + (p1 * 3 + r1) * 6 + (p2 * 3 + r2)] = 1;
    }
    int rc;
#define UP(vec, field)                                  \
    if ((rc = upload(t, vec, &D.field)) != HDB_OK) return rc
    UP(S.points, points);
    UP(S.beamN32, beamN);
    UP(S.pinMask, pinMask);
    UP(dropped, dropped);
    {
        std::vector<unsigned long long> mask((size_t)S.nB, 0ull);
        for (int b = 0; b < S.nB; b++)
            for (int e = 0; e < 36; e++)
                if (dropped[(size_t)b * 36 + e]) mask[b] |= 1ull << e;
        UP(mask, droppedMask);
    }
    UP(S.gptr, gptr);
    UP(pad(S.gsrc), gsrc);
    { // fixed-width (ELL) copies of the two gather lists the solve path uses: [width][n], -1 = no contribution.
      // Every index load of a slot is then independent and coalesced; widths above 8 keep the CSR walk.
        auto ell = [](const std::vector<int32_t> &ptr, const std::vector<int32_t> &src, int n, int &width) {
            width = 0;
            for (int s2 = 0; s2 < n; s2++) width = std::max(width, ptr[s2 + 1] - ptr[s2]);
            if (width > 8) width = 0;
            std::vector<int32_t> o((size_t)width * n, -1);
            for (int s2 = 0; s2 < n && width; s2++)
                for (int k = ptr[s2]; k < ptr[s2 + 1]; k++) o[(size_t)(k - ptr[s2]) * n + s2] = src[k];
            return o;
        };
        {
            std::vector<int32_t> g = ell(S.gptr, pad(S.gsrc), S.nSlots, D.gWidth);
            bool ok = D.gWidth > 0;
            for (int s2 = 0; s2 < S.nSlots && ok; s2++) // unit diagonal of a dummy unknown: marker in row 0
                if (S.slotInit[s2]) {
                    if (S.gptr[s2 + 1] != S.gptr[s2]) ok = false; // never the case: such a slot has no stiffness entry
                    else g[s2] = -2;
                }
            if (!ok) D.gWidth = 0;
            UP(g, gEll);
        }
        UP(ell(S.rgptr, pad(S.rgsrc), S.nnzR, D.rgWidth), rgEll);
        { // compact element store (k_assemble_quad): per beam 8 values (+1 pad) without pins, 36 (+1) with
            std::vector<int32_t> keOff((size_t)S.nB);
            int len = 0;
            for (int b = 0; b < S.nB; b++) {
                keOff[b] = len;
                len += S.pinMask[b] ? 37 : 9;
            }
            const unsigned long long negMask = HD_RIGID_NEG;
            auto compact = [&](const std::vector<int32_t> &src) { // codes beam*36 + position
                std::vector<int32_t> o(src.size());
                for (size_t i = 0; i < src.size(); i++) {
                    const int b = src[i] / 36, e = src[i] % 36;
                    if (S.pinMask[b]) o[i] = (keOff[b] + e) * 2;
                    else {
                        int v = 0;
                        while (!((HD_RIGID_POS(v) >> e) & 1ull)) v++;
                        o[i] = (keOff[b] + v) * 2 + (int)((negMask >> e) & 1ull);
                    }
                }
                return o;
            };
            int w1 = 0, w2 = 0;
            std::vector<int32_t> cg = ell(S.gptr, compact(S.gsrc), S.nSlots, w1);
            for (int s2 = 0; s2 < S.nSlots && w1; s2++)
                if (S.slotInit[s2]) cg[s2] = -2;
            std::vector<int32_t> crg = ell(S.rgptr, compact(S.rgsrc), S.nnzR, w2);
            D.keLen = (D.gWidth > 0 && D.rgWidth > 0 && w1 == D.gWidth && w2 == D.rgWidth) ? len : 0;
            UP(cg, cgEll);
            UP(crg, crgEll);
            UP(keOff, keOff);
        }
        std::vector<double> geom((size_t)S.nB * 3);
        for (int b = 0; b < S.nB; b++) { // same routines, same rounding as on the device (element.cuh)
            const int n0 = S.beamN32[2 * b], n1 = S.beamN32[2 * b + 1];
            const double x0 = S.points[2 * n0], y0 = S.points[2 * n0 + 1], x1 = S.points[2 * n1], y1 = S.points[2 * n1 + 1];
            const double len = beam_length(x0, y0, x1, y1);
            double c, sn;
            beam_cs(x0, y0, x1, y1, len, c, sn);
            geom[3 * b] = c; geom[3 * b + 1] = sn; geom[3 * b + 2] = len;
        }
        UP(geom, geom);
    }
    UP(S.slotInit, slotInit);
    UP(S.rgptr, rgptr);
    UP(pad(S.rgsrc), rgsrc);
    UP(S.kptr, kptr);
    UP(pad(S.ksrc), ksrc);
    UP(S.BLp, BLp);
    UP(S.BLi, BLi);
    UP(S.colOff, colOff);
    UP(S.blkSlot, blkSlot);
    UP(S.pairPtr, pairPtr);
    UP(S.pairTgt, pairTgt);
    UP(S.needUpto, needUpto);
    {
        std::vector<int32_t> desc((size_t)S.nBlk * 4);
        for (int J = 0; J < S.nBlk; J++) {
            const int nb = S.BLp[J + 1] - S.BLp[J];
            desc[4 * J + 0] = S.colOff[J];
            desc[4 * J + 1] = nb;
            desc[4 * J + 2] = S.needUpto[std::min(J + FACTOR_PREFETCH, S.nBlk - 1)];
            desc[4 * J + 3] = nb > 0 ? S.pairTgt[S.pairPtr[J]] : -1;
        }
        const int32_t *dp = nullptr;
        if ((rc = upload(t, desc, &dp)) != HDB_OK) return rc;
        D.colDesc = reinterpret_cast<const int4 *>(dp);
    }
    {
        const int32_t *f = nullptr, *b = nullptr;
        if ((rc = upload(t, S.fwdOps, &f)) != HDB_OK) return rc;
        if ((rc = upload(t, S.bwdOps, &b)) != HDB_OK) return rc;
        D.fwdOps = reinterpret_cast<const int4 *>(f);
        D.bwdOps = reinterpret_cast<const int4 *>(b);
    }
    UP(S.xrow, xrow);
    UP(S.beamRows, beamRows);
    UP(S.supIndex, supIndex);
    UP(S.supDofs, supDofs);
    UP(S.rptr, rptr);
    UP(S.rcol, rcol);
    UP(S.dummyPos, dummyPos);
    { // row chunks of k_solve_bulk and the supported DOFs inside each
        D.rNCh = (S.dof + BULK_CHUNK - 1) / BULK_CHUNK;
        D.rChunk = D.rNCh > 0 ? (((S.dof + D.rNCh - 1) / D.rNCh + 1) & ~1) : 0;
        std::vector<int32_t> ptr(1, 0), list;
        for (int j = 0; j < D.rNCh; j++) {
            for (int g = j * D.rChunk; g < std::min(S.dof, (j + 1) * D.rChunk); g++)
                if (S.supIndex[g] >= 0) list.push_back((g - j * D.rChunk) | (S.supIndex[g] << 16));
            ptr.push_back((int32_t)list.size());
        }
        UP(ptr, rsPtr);
        UP(list, rsList);
    }
#undef UP
    if (const char *e = std::getenv("HDB_SUB_BATCH")) {
        int v = std::atoi(e);
        if (v >= 32) t->subBatch = v;
    }
    t->devReady = true;
    return HDB_OK;
}

static int ensure_workspace(hdb_topology *t, size_t models)
{
    if (!t->prepStream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&t->prepStream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&t->evFork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t->evJoin, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaEventCreateWithFlags(&t->evPrep[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&t->evSolve[i], cudaEventDisableTiming));
        }
    }
    if (models <= t->wsModels) return HDB_OK;
    CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < 2; i++) {
        if (t->Wbuf[i]) cudaFree(t->Wbuf[i]);
        if (t->Lmbuf[i]) cudaFree(t->Lmbuf[i]);
        if (t->Ksfbuf[i]) cudaFree(t->Ksfbuf[i]);
        t->Wbuf[i] = t->Lmbuf[i] = t->Ksfbuf[i] = nullptr;
    }
    if (t->keScratch) cudaFree(t->keScratch);
    t->keScratch = nullptr;
    t->wsModels = 0;
    const size_t ld = (models + 31) / 32 * 32;
    if (std::getenv("HDB_DEBUG_CLOCKS")) {
        if (t->dbgClocks) cudaFree(t->dbgClocks);
        CUDA_TRY(cudaMalloc(&t->dbgClocks, ld * 32 * sizeof(long long)));
        CUDA_TRY(cudaMemset(t->dbgClocks, 0, ld * 32 * sizeof(long long)));
    }
    for (int i = 0; i < 2; i++) {
        CUDA_TRY(cudaMalloc(&t->Wbuf[i], (size_t)t->S.nSlots * ld * sizeof(double)));
        CUDA_TRY(cudaMalloc(&t->Lmbuf[i], (size_t)t->S.nSlots * ld * sizeof(double)));
        CUDA_TRY(cudaMalloc(&t->Ksfbuf[i], std::max<size_t>(1, ld * (size_t)t->S.nnzR) * sizeof(double) + 16)); // + 16: the solve kernel's bulk copy of a model's block covers whole 16-byte pairs
    }
    if (!assemble_fits_smem(t->D, t->smemOptin))
        CUDA_TRY(cudaMalloc(&t->keScratch, ld * (size_t)t->S.nB * KE_LD * sizeof(double)));
    t->wsModels = ld; // every buffer above is sized for ld models: a later, larger batch within ld reuses them
    return HDB_OK;
}

static bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

enum { KIND_ASSEMBLE = 0, KIND_FACTOR = 1, KIND_SOLVE = 2, KIND_MODAL0 = 3 };
constexpr size_t PROF_MAX_INTERVALS = 1 << 16;

static void prof_reset(hdb_topology *t)
{
    if (t->profiling != 2) t->evKind.clear();
}

static cudaEvent_t prof_event(hdb_topology *t, size_t i)
{
    while (t->events.size() <= i) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        t->events.push_back(e);
    }
    return t->events[i];
}

struct ProfScope {
    hdb_topology *t;
    cudaStream_t st;
    size_t slot;
    bool on;
    ProfScope(hdb_topology *t_, int kind, cudaStream_t s) : t(t_), st(s), slot(0), on(t_->profiling != 0)
    {
        if (on && t->evKind.size() >= PROF_MAX_INTERVALS) on = false; // accumulate mode: bounded, the rest goes untimed
        if (!on) return;
        slot = t->evKind.size();
        t->evKind.push_back(kind);
        cudaEventRecord(prof_event(t, 2 * slot), st);
    }
    ~ProfScope()
    {
        if (on) cudaEventRecord(prof_event(t, 2 * slot + 1), st);
    }
};

// ------------------------------------------------------------------------------------------------
// C ABI: topology
// ------------------------------------------------------------------------------------------------
extern "C" {

int hdb_version(void) { return HDB_VERSION; }

const char *hdb_error_string(int code)
{
    switch (code) {
    case HDB_OK: return "ok";
    case HDB_ERR_ARG: return "invalid argument";
    case HDB_ERR_NO_DEVICE: return "no CUDA device (libhdb200 has no CPU path)";
    case HDB_ERR_CUDA: return "CUDA runtime error";
    case HDB_ERR_NOMEM: return "out of memory";
    case HDB_ERR_UNSUPPORTED: return "model too large for the shared-memory resident kernels";
    case HDB_ERR_INPUT: return "input data rejected";
    case HDB_ERR_SMALL_BUFFER: return "output buffer too small";
    default: return "unknown error";
    }
}

const char *hdb_last_error_detail(void) { return g_detail.c_str(); }

int64_t hdb_launch_count(void) { return g_launches.load(); }

int hdb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int hdb_topology_create(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, const uint8_t *pins,
                        const uint8_t *supports, const hdb_options *opts, hdb_topology **out)
{
    if (!out) return HDB_ERR_ARG;
    *out = nullptr;
    hdb_topology *t = new (std::nothrow) hdb_topology();
    if (!t) return HDB_ERR_NOMEM;
    std::string detail;
    const int ordering = opts ? opts->ordering : 0;
    Nvtx nv("hdb_topology_create (symbolic phase)");
    int rc = build_symbolic(nP, points, nB, beams, pins, supports, ordering == 2 ? 0 : ordering, t->S, detail);
    if (rc == HDB_OK && ordering == 2) { // automatic: keep the ordering whose factor is smaller
        Symbolic rcm;
        std::string d2;
        if (build_symbolic(nP, points, nB, beams, pins, supports, 1, rcm, d2) == HDB_OK && rcm.nSlots < t->S.nSlots) t->S = std::move(rcm);
    }
    if (rc != HDB_OK) {
        set_error_detail(detail);
        delete t;
        return rc;
    }
    t->device = (opts && opts->device >= 0) ? opts->device : -1;
    if (t->device < 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) {
            (void)cudaGetLastError();
            cur = 0;
        }
        t->device = cur;
    }
    *out = t;
    return HDB_OK;
}

void hdb_topology_destroy(hdb_topology *t)
{
    if (!t) return;
    if (!t->devReady && !t->devAllocs.empty()) { // an upload failed half-way and was never retried
        cudaSetDevice(t->device);
        for (void *p : t->devAllocs) cudaFree(p);
    }
    if (t->devReady) {
        cudaSetDevice(t->device);
        cudaDeviceSynchronize();
        for (void *p : t->devAllocs) cudaFree(p);
        for (int i = 0; i < 2; i++) {
            if (t->Wbuf[i]) cudaFree(t->Wbuf[i]);
            if (t->Lmbuf[i]) cudaFree(t->Lmbuf[i]);
            if (t->Ksfbuf[i]) cudaFree(t->Ksfbuf[i]);
            if (t->evPrep[i]) cudaEventDestroy(t->evPrep[i]);
            if (t->evSolve[i]) cudaEventDestroy(t->evSolve[i]);
        }
        if (t->evFork) cudaEventDestroy(t->evFork);
        if (t->evJoin) cudaEventDestroy(t->evJoin);
        if (t->prepStream) cudaStreamDestroy(t->prepStream);
        if (t->keScratch) cudaFree(t->keScratch);
        if (t->geomBuf) cudaFree(t->geomBuf);
        for (auto &c : t->chunk) {
            if (c.base) cudaFree(c.base);
            if (c.evIn) cudaEventDestroy(c.evIn);
            if (c.evK) cudaEventDestroy(c.evK);
            if (c.evOut) cudaEventDestroy(c.evOut);
        }
        for (cudaStream_t st : {t->sIn, t->sK, t->sOut})
            if (st) cudaStreamDestroy(st);
        for (cudaEvent_t e : t->events) cudaEventDestroy(e);
        if (t->modalBuf) cudaFree(t->modalBuf);
    }
    delete t;
}

int hdb_topology_get_info(const hdb_topology *t, hdb_topology_info *info)
{
    if (!t || !info) return HDB_ERR_ARG;
    std::memset(info, 0, sizeof(*info));
    const Symbolic &S = t->S;
    info->nP = S.nP; info->nB = S.nB; info->dof = S.dof; info->nFree = S.nFree;
    info->nElim = (int32_t)S.elim.size(); info->nIgnore = (int32_t)S.ignore.size(); info->nSupported = S.nSup;
    info->nnzK = S.nnzK; info->nnzKff = S.nnzKff; info->nnzL = S.nnzL; info->nnzR = S.nnzR;
    info->nUpdates = (int32_t)S.nUpdates; info->device = t->device;
    return HDB_OK;
}

int64_t hdb_topology_query(const hdb_topology *t, int32_t which, void *buf, int64_t capacity)
{
    if (!t) return HDB_ERR_ARG;
    const std::vector<int32_t> *v = nullptr;
    const Symbolic &S = t->S;
    switch (which) {
    case HDB_Q_ELIMINATED: v = &S.elim; break;
    case HDB_Q_IGNORE: v = &S.ignore; break;
    case HDB_Q_CSR_ROWPTR: v = &S.rowptr; break;
    case HDB_Q_CSR_COLIND: v = &S.colind; break;
    case HDB_Q_SCATTER: v = &S.scatter; break;
    case HDB_Q_FREE_INDEX: v = &S.freeIndex; break;
    case HDB_Q_PERM: v = &S.perm; break;
    case HDB_Q_L_COLPTR: v = &S.Lp; break;
    case HDB_Q_L_ROWIND: v = &S.Li; break;
    default: return HDB_ERR_ARG;
    }
    const int64_t bytes = (int64_t)v->size() * (int64_t)sizeof(int32_t);
    if (buf) {
        if (capacity < bytes) return HDB_ERR_SMALL_BUFFER;
        if (bytes) std::memcpy(buf, v->data(), (size_t)bytes);
    }
    return (int64_t)v->size();
}

// ------------------------------------------------------------------------------------------------
// C ABI: validation helpers (check.go) and assemblyNodeLoad
// ------------------------------------------------------------------------------------------------
static void copy_msg(const std::string &s, char *msg, int32_t len)
{
    if (!msg || len <= 0) return;
    const size_t n = std::min<size_t>(s.size(), (size_t)len - 1);
    std::memcpy(msg, s.data(), n);
    msg[n] = 0;
}

int hdb_check_model(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, int32_t nPins,
                    const uint8_t *pins, int32_t nSupports, const uint8_t *supports, char *msg, int32_t msgLen)
{
    if (nP < 0 || nB < 0 || nPins < 0 || nSupports < 0 || (nP > 0 && !points) || (nB > 0 && !beams)) return HDB_ERR_ARG;
    std::string m;
    int rc = check_model_impl(nP, points, nB, beams, nPins, pins, nSupports, supports, m);
    copy_msg(m, msg, msgLen);
    return rc;
}

int hdb_check_loads(int32_t nP, int32_t nLoads, const hdb_load_node *loads, char *msg, int32_t msgLen)
{
    if (nLoads < 0 || (nLoads > 0 && !loads)) return HDB_ERR_ARG;
    std::string m;
    bool err = false;
    for (int j = 0; j < nLoads; j++) { // check.go:131-155
        if (loads[j].n < 0 || loads[j].n >= nP) {
            m += "Error in load case, load position " + std::to_string(j) + " : outside index of point\n";
            err = true;
        }
        for (int k = 0; k < 3; k++)
            if (std::isnan(loads[j].forces[k]) || std::isinf(loads[j].forces[k])) {
                m += "Error in load case, load position " + std::to_string(j) + " : not finite force\n";
                err = true;
            }
    }
    copy_msg(m, msg, msgLen);
    return err ? HDB_ERR_INPUT : HDB_OK;
}

int hdb_check_modal(int32_t nP, int32_t nSupports, const uint8_t *supports, int32_t nMasses,
                    const hdb_modal_mass *masses, char *msg, int32_t msgLen)
{
    if (nMasses < 0 || (nMasses > 0 && !masses)) return HDB_ERR_ARG;
    std::string m;
    bool err = false;
    for (int j = 0; j < nMasses; j++) { // check.go:413-447
        const int64_t n = masses[j].n;
        const double v = masses[j].mass;
        if (n < 0 || n >= nP) {
            m += "Error in modal case, mass position " + std::to_string(j) + " : outside index of point\n";
            err = true;
        }
        if (std::isnan(v) || std::isinf(v) || v < 0 || v == 0) {
            m += "Error in modal case, mass position " + std::to_string(j) + " : not valid mass\n";
            err = true;
        }
        if (supports && n >= 0 && n < nSupports && (supports[3 * n] || supports[3 * n + 1] || supports[3 * n + 2])) {
            m += "Error in modal case, mass position " + std::to_string(j) + " : Modal mass on support\n";
            err = true;
        }
    }
    copy_msg(m, msg, msgLen);
    return err ? HDB_ERR_INPUT : HDB_OK;
}

int hdb_assemble_node_loads(int32_t nP, int32_t nLoads, const hdb_load_node *loads, double *p)
{
    if (nP < 0 || nLoads < 0 || !p || (nLoads > 0 && !loads)) return HDB_ERR_ARG;
    for (int i = 0; i < 3 * nP; i++) p[i] = 0.0;
    for (int j = 0; j < nLoads; j++) { // model.go:687-691
        if (loads[j].n < 0 || loads[j].n >= nP) return HDB_ERR_INPUT;
        for (int i = 0; i < 3; i++) p[loads[j].n * 3 + i] += loads[j].forces[i];
    }
    return HDB_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: pinned memory, profiling
// ------------------------------------------------------------------------------------------------
int hdb_host_alloc(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return HDB_ERR_ARG;
    if (hdb_device_count() <= 0) {
        set_error_detail("hdb_host_alloc: no CUDA device");
        return HDB_ERR_NO_DEVICE;
    }
    CUDA_TRY(cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return HDB_OK;
}

int hdb_host_free(void *ptr)
{
    if (!ptr) return HDB_OK;
    CUDA_TRY(cudaFreeHost(ptr));
    return HDB_OK;
}

int hdb_set_profiling(hdb_topology *t, int32_t on)
{
    if (!t) return HDB_ERR_ARG;
    t->profiling = on;
    t->evKind.clear();
    return HDB_OK;
}

int hdb_last_kernel_ms(hdb_topology *t, float *ms, int32_t capacity)
{
    if (!t || !ms || capacity <= 0) return HDB_ERR_ARG;
    int kinds = 0;
    for (int k : t->evKind) kinds = std::max(kinds, k + 1);
    kinds = std::min<int>(kinds, capacity);
    for (int i = 0; i < kinds; i++) ms[i] = 0.f;
    for (size_t i = 0; i < t->evKind.size(); i++) {
        float v = 0.f;
        if (cudaEventSynchronize(t->events[2 * i + 1]) != cudaSuccess) return HDB_ERR_CUDA;
        if (cudaEventElapsedTime(&v, t->events[2 * i], t->events[2 * i + 1]) != cudaSuccess) return HDB_ERR_CUDA;
        if (t->evKind[i] < kinds) ms[t->evKind[i]] += v;
    }
    return kinds;
}

// ------------------------------------------------------------------------------------------------
// C ABI: LinearStatic batch
// ------------------------------------------------------------------------------------------------
static int static_enqueue(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *loads,
                          double *disp, double *forces, double *react, int32_t *status, int32_t *failCase,
                          double *kval, cudaStream_t st, bool solve, const double *points = nullptr)
{
    const Symbolic &S = t->S;
    int cp;
    size_t smem;
    if (solve && !solve_config(t->D, nLC, t->smemOptin, cp, smem) && !solve_general_config(t->D, nLC, t->smemOptin, cp, smem)) {
        set_error_detail("model does not fit the solve kernels: the right-hand side of one load case exceeds shared memory");
        return HDB_ERR_UNSUPPORTED;
    }
    const int sb = std::min<int>(nModels, t->subBatch);
    int rc = ensure_workspace(t, (size_t)sb);
    if (rc != HDB_OK) return rc;
    if (points && (size_t)sb > t->geomModels) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (t->geomBuf) cudaFree(t->geomBuf);
        t->geomBuf = nullptr;
        t->geomModels = 0;
        CUDA_TRY(cudaMalloc(&t->geomBuf, (size_t)sb * S.nB * 3 * sizeof(double)));
        t->geomModels = (size_t)sb;
    }
    const int nSub = (nModels + sb - 1) / sb;
    // Overlapping the assembly/factorization of sub-batch i+1 with the solve of sub-batch i on a second stream
    // stopped paying once the solve kernel became persistent (one block per SM holding ~210 kB of shared
    // memory: the other kernels' blocks only get an SM when a solve block retires).  Opt-in for experiments.
    static const bool wantOverlap = std::getenv("HDB_OVERLAP_PREP") != nullptr;
    const bool pipelined = wantOverlap && solve && nSub > 1 && !t->keScratch && !t->dbgClocks && !points;
    cudaStream_t ps = pipelined ? t->prepStream : st;
    if (pipelined) {
        CUDA_TRY(cudaEventRecord(t->evFork, st));
        CUDA_TRY(cudaStreamWaitEvent(ps, t->evFork, 0));
    }
    for (int i = 0, m0 = 0; m0 < nModels; m0 += sb, i++) {
        const int nm = std::min(sb, nModels - m0);
        const int buf = pipelined ? (i & 1) : 0;
        StaticLaunch L{};
        L.nModels = nm;
        L.nLC = nLC;
        L.ld = (int)t->wsModels;
        L.sections = sections + (size_t)m0 * S.nB * 3;
        L.loads = loads ? loads + (size_t)m0 * nLC * S.dof : nullptr;
        L.disp = disp ? disp + (size_t)m0 * nLC * S.dof : nullptr;
        L.forces = forces ? forces + (size_t)m0 * nLC * S.nB * 6 : nullptr;
        L.react = react ? react + (size_t)m0 * nLC * S.dof : nullptr;
        L.status = status + m0;
        L.failCase = failCase ? failCase + m0 : nullptr;
        L.W = t->Wbuf[buf];
        L.Lm = t->Lmbuf[buf];
        L.Ksf = t->Ksfbuf[buf];
        L.kval = kval ? kval + (size_t)m0 * S.nnzK : nullptr;
        L.keScratch = t->keScratch;
        L.dbgClocks = t->dbgClocks;
        if (pipelined && i >= 2) CUDA_TRY(cudaStreamWaitEvent(ps, t->evSolve[buf], 0)); // workspace set is free again
        if (points) { // per-model coordinates: beam geometry of this sub-batch first
            L.points = points + (size_t)m0 * S.nP * 2;
            L.geom = t->geomBuf;
            CUDA_TRY(launch_geometry(t->D, L, ps));
            count_launch(1);
        }
        {
            Nvtx nv("hdb:assemble");
            ProfScope pr(t, KIND_ASSEMBLE, ps);
            CUDA_TRY(launch_assemble(t->D, L, ps));
        }
        count_launch(1);
        if (!solve) continue;
        {
            Nvtx nv("hdb:factor");
            ProfScope pr(t, KIND_FACTOR, ps);
            CUDA_TRY(launch_factor(t->D, L, ps));
        }
        if (pipelined) {
            CUDA_TRY(cudaEventRecord(t->evPrep[buf], ps));
            CUDA_TRY(cudaStreamWaitEvent(st, t->evPrep[buf], 0));
        }
        {
            Nvtx nv("hdb:solve+recover");
            ProfScope pr(t, KIND_SOLVE, st);
            CUDA_TRY(launch_solve_pipe(t->D, L, st));
        }
        if (pipelined) CUDA_TRY(cudaEventRecord(t->evSolve[buf], st));
        count_launch(2);
    }
    return HDB_OK;
}

static void debug_report_clocks(hdb_topology *t, int nModels)
{
    if (!t->dbgClocks) return;
    cudaDeviceSynchronize();
    const int n = std::min<int>(std::min<int>(nModels, (int)t->wsModels), t->smCount);
    std::vector<long long> h((size_t)n * 32);
    cudaMemcpy(h.data(), t->dbgClocks, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double blk = 0, sol = 0, models = 0, ph[24] = {};
    for (int b = 0; b < n; b++) {
        const long long *d = h.data() + (size_t)b * 32;
        blk += (double)d[0];
        sol += (double)d[1];
        models += (double)d[2];
        for (int k = 0; k < 24; k++) ph[k] += (double)d[8 + k];
    }
    if (models <= 0) return;
    std::fprintf(stderr, "[hdb] solve kernel cycles/model (last launch, %d blocks): block %.0f | solver warp %.0f | worker warp 1:\n", n,
                 blk / models, sol / models);
    static const char *namesPipe[11] = {"Ksf/section loads issued", "-", "disp + Ksf, coefficients -> smem", "(barrier)", "react", "forces", "(barrier)", "factor gather issue",
                                        "load cases", "(barrier)", "status"};
    static const char *namesBulk[11] = {"displacements + coefficient table", "(reaction block has landed)", "reactions", "beam forces", "(last copies read)", "(barrier)", "load cases", "(barrier)", "-", "-", "-"};
    const char *leg = std::getenv("HDB_SOLVE_LEGACY");
    const char **names = (leg && atoi(leg)) ? namesPipe : namesBulk;
    std::fprintf(stderr, "[hdb]  ");
    for (int k = 0; k < 11; k++) std::fprintf(stderr, " %s %.0f |", names[k], ph[k] / models);
    std::fprintf(stderr, "\n");
    if (!(leg && atoi(leg))) std::fprintf(stderr, "[hdb]   helper warp: bulk copy of the reaction block issued after %.0f\n", ph[16] / models);
}

int hdb_static_batch_dev(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *loads,
                         double *disp, double *forces, double *react, int32_t *status, int32_t *failCase, void *stream)
{
    if (!t || nModels < 0 || nLC < 0 || !sections || !status || (nLC > 0 && !loads)) return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0) return HDB_OK;
    prof_reset(t);
    rc = static_enqueue(t, nModels, sections, nLC, loads, disp, forces, react, status, failCase, nullptr,
                        (cudaStream_t)stream, true);
    if (rc == HDB_OK) debug_report_clocks(t, nModels);
    return rc;
}

int hdb_static_batch_geom_dev(hdb_topology *t, int32_t nModels, const double *points, const double *sections, int32_t nLC,
                              const double *loads, double *disp, double *forces, double *react, int32_t *status,
                              int32_t *failCase, void *stream)
{
    if (!t || nModels < 0 || nLC < 0 || !points || !sections || !status || (nLC > 0 && !loads)) return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0) return HDB_OK;
    prof_reset(t);
    return static_enqueue(t, nModels, sections, nLC, loads, disp, forces, react, status, failCase, nullptr, (cudaStream_t)stream,
                          true, points);
}

static int ensure_chunks(hdb_topology *t, int chunkModels, int nLC)
{
    if (t->chunkModels >= chunkModels && t->chunkLC >= nLC && t->chunk[0].base) return HDB_OK;
    const Symbolic &S = t->S;
    for (auto &c : t->chunk) {
        if (c.base) cudaFree(c.base);
        c.base = nullptr;
    }
    const size_t m = (size_t)chunkModels;
    auto al = [](size_t doubles) { return (doubles + 31) / 32 * 32; }; // keep every sub-buffer 256-byte aligned
    const size_t nSec = al(m * S.nB * 3), nLoad = al(m * nLC * S.dof), nForce = al(m * nLC * S.nB * 6);
    const size_t doubles = nSec + 3 * nLoad + nForce;
    const size_t bytes = doubles * sizeof(double) + 2 * al(m) * sizeof(int32_t) + 256;
    for (auto &c : t->chunk) {
        CUDA_TRY(cudaMalloc(&c.base, bytes));
        double *p = (double *)c.base;
        c.sec = p; p += nSec;
        c.loads = p; p += nLoad;
        c.disp = p; p += nLoad;
        c.react = p; p += nLoad;
        c.forces = p; p += nForce;
        c.status = (int32_t *)p;
        c.fail = c.status + al(m);
        c.bytes = bytes;
        if (!c.evIn) {
            CUDA_TRY(cudaEventCreateWithFlags(&c.evIn, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c.evK, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c.evOut, cudaEventDisableTiming));
        }
    }
    if (!t->sIn) {
        CUDA_TRY(cudaStreamCreateWithFlags(&t->sIn, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&t->sK, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&t->sOut, cudaStreamNonBlocking));
    }
    t->chunkModels = chunkModels;
    t->chunkLC = nLC;
    return HDB_OK;
}

int hdb_static_batch(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *loads,
                     double *disp, double *forces, double *react, int32_t *status, int32_t *failCase)
{
    if (!t || nModels < 0 || nLC < 0 || !sections || !status || (nLC > 0 && !loads)) return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0) return HDB_OK;
    const bool dev = is_device_ptr(sections);
    const void *ptrs[] = {loads, disp, forces, react, status, failCase};
    for (const void *p : ptrs)
        if (p && is_device_ptr(p) != dev) {
            set_error_detail("hdb_static_batch: host and device pointers mixed in one call");
            return HDB_ERR_ARG;
        }
    if (dev) {
        rc = hdb_static_batch_dev(t, nModels, sections, nLC, loads, disp, forces, react, status, failCase, nullptr);
        if (rc != HDB_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        return HDB_OK;
    }
    Nvtx nvAll("hdb_static_batch (host arrays)");
    // ---- host memory: chunked three-stage pipeline. copy-in, compute and copy-out each own a stream;
    //      NBUF device buffers rotate, so H2D(i+1), kernels(i) and D2H(i-1) overlap. ----
    const Symbolic &S = t->S;
    const size_t perModel = ((size_t)S.nB * 3 + 3 * (size_t)nLC * S.dof + (size_t)nLC * S.nB * 6) * sizeof(double);
    int cm = (int)(((size_t)192 << 20) / std::max<size_t>(perModel, 1));
    cm = std::max(32, std::min(cm, 1024));
    cm = std::min(cm, nModels);
    cm = std::min(cm, t->subBatch);
    rc = ensure_chunks(t, cm, nLC);
    if (rc != HDB_OK) return rc;
    prof_reset(t);
    int it = 0;
    // an error exit must not leave copies into the caller's buffers in flight
    struct Drain {
        hdb_topology *t;
        bool armed = true;
        ~Drain()
        {
            if (!armed) return;
            for (cudaStream_t s : {t->sIn, t->sK, t->sOut})
                if (s) cudaStreamSynchronize(s);
        }
    } drain{t};
    for (int m0 = 0; m0 < nModels; m0 += t->chunkModels, it++) {
        HostChunk &c = t->chunk[it % NBUF];
        const int nm = std::min(t->chunkModels, nModels - m0);
        Nvtx nvChunk("hdb:chunk H2D | kernels | D2H");
        if (it >= NBUF) CUDA_TRY(cudaStreamWaitEvent(t->sIn, c.evOut, 0)); // buffer is free again
        CUDA_TRY(cudaMemcpyAsync(c.sec, sections + (size_t)m0 * S.nB * 3, (size_t)nm * S.nB * 3 * sizeof(double),
                                 cudaMemcpyHostToDevice, t->sIn));
        CUDA_TRY(cudaMemcpyAsync(c.loads, loads + (size_t)m0 * nLC * S.dof, (size_t)nm * nLC * S.dof * sizeof(double),
                                 cudaMemcpyHostToDevice, t->sIn));
        CUDA_TRY(cudaEventRecord(c.evIn, t->sIn));
        CUDA_TRY(cudaStreamWaitEvent(t->sK, c.evIn, 0));
        rc = static_enqueue(t, nm, c.sec, nLC, c.loads, disp ? c.disp : nullptr, forces ? c.forces : nullptr,
                            react ? c.react : nullptr, c.status, c.fail, nullptr, t->sK, true);
        if (rc != HDB_OK) return rc;
        CUDA_TRY(cudaEventRecord(c.evK, t->sK));
        CUDA_TRY(cudaStreamWaitEvent(t->sOut, c.evK, 0));
        if (disp)
            CUDA_TRY(cudaMemcpyAsync(disp + (size_t)m0 * nLC * S.dof, c.disp, (size_t)nm * nLC * S.dof * sizeof(double),
                                     cudaMemcpyDeviceToHost, t->sOut));
        if (forces)
            CUDA_TRY(cudaMemcpyAsync(forces + (size_t)m0 * nLC * S.nB * 6, c.forces,
                                     (size_t)nm * nLC * S.nB * 6 * sizeof(double), cudaMemcpyDeviceToHost, t->sOut));
        if (react)
            CUDA_TRY(cudaMemcpyAsync(react + (size_t)m0 * nLC * S.dof, c.react, (size_t)nm * nLC * S.dof * sizeof(double),
                                     cudaMemcpyDeviceToHost, t->sOut));
        CUDA_TRY(cudaMemcpyAsync(status + m0, c.status, (size_t)nm * sizeof(int32_t), cudaMemcpyDeviceToHost, t->sOut));
        if (failCase)
            CUDA_TRY(cudaMemcpyAsync(failCase + m0, c.fail, (size_t)nm * sizeof(int32_t), cudaMemcpyDeviceToHost, t->sOut));
        CUDA_TRY(cudaEventRecord(c.evOut, t->sOut));
    }
    CUDA_TRY(cudaStreamSynchronize(t->sOut));
    drain.armed = false;
    return HDB_OK;
}

// stage a host array on the device for the small test entry points
struct Staged {
    void *dev = nullptr;
    bool owned = false;
    ~Staged()
    {
        if (owned && dev) cudaFree(dev);
    }
};

static int stage_in(const void *p, size_t bytes, Staged &s)
{
    if (is_device_ptr(p)) {
        s.dev = const_cast<void *>(p);
        return HDB_OK;
    }
    CUDA_TRY(cudaMalloc(&s.dev, std::max<size_t>(bytes, 1)));
    s.owned = true;
    if (p && bytes) CUDA_TRY(cudaMemcpy(s.dev, p, bytes, cudaMemcpyHostToDevice));
    return HDB_OK;
}

static int stage_out(void *p, size_t bytes, Staged &s)
{
    if (!p) return HDB_OK;
    if (is_device_ptr(p)) {
        s.dev = p;
        return HDB_OK;
    }
    CUDA_TRY(cudaMalloc(&s.dev, std::max<size_t>(bytes, 1)));
    s.owned = true;
    return HDB_OK;
}

static int fetch_out(void *p, size_t bytes, Staged &s)
{
    if (p && s.owned && bytes) CUDA_TRY(cudaMemcpy(p, s.dev, bytes, cudaMemcpyDeviceToHost));
    return HDB_OK;
}

int hdb_static_batch_geom(hdb_topology *t, int32_t nModels, const double *points, const double *sections, int32_t nLC,
                          const double *loads, double *disp, double *forces, double *react, int32_t *status, int32_t *failCase)
{
    if (!t || nModels < 0 || nLC < 0 || !points || !sections || !status || (nLC > 0 && !loads)) return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0) return HDB_OK;
    const Symbolic &S = t->S;
    Staged pts, sec, ld, dd, ff, rr, stt, fc;
    const size_t nM = (size_t)nModels, rows = nM * nLC;
    const size_t bPts = nM * S.nP * 2 * sizeof(double), bSec = nM * S.nB * 3 * sizeof(double), bDof = rows * S.dof * sizeof(double),
                 bFor = rows * S.nB * 6 * sizeof(double), bSt = nM * sizeof(int32_t);
    if ((rc = stage_in(points, bPts, pts)) != HDB_OK) return rc;
    if ((rc = stage_in(sections, bSec, sec)) != HDB_OK) return rc;
    if ((rc = stage_in(loads, bDof, ld)) != HDB_OK) return rc;
    if ((rc = stage_out(disp, bDof, dd)) != HDB_OK) return rc;
    if ((rc = stage_out(forces, bFor, ff)) != HDB_OK) return rc;
    if ((rc = stage_out(react, bDof, rr)) != HDB_OK) return rc;
    if ((rc = stage_out(status, bSt, stt)) != HDB_OK) return rc;
    if ((rc = stage_out(failCase, bSt, fc)) != HDB_OK) return rc;
    rc = hdb_static_batch_geom_dev(t, nModels, (const double *)pts.dev, (const double *)sec.dev, nLC, (const double *)ld.dev,
                                   (double *)dd.dev, (double *)ff.dev, (double *)rr.dev, (int32_t *)stt.dev, (int32_t *)fc.dev, nullptr);
    if (rc != HDB_OK) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    if ((rc = fetch_out(disp, bDof, dd)) != HDB_OK) return rc;
    if ((rc = fetch_out(forces, bFor, ff)) != HDB_OK) return rc;
    if ((rc = fetch_out(react, bDof, rr)) != HDB_OK) return rc;
    if ((rc = fetch_out(status, bSt, stt)) != HDB_OK) return rc;
    return fetch_out(failCase, bSt, fc);
}

int hdb_assemble_batch(hdb_topology *t, int32_t nModels, const double *sections, double *kval, int32_t *status)
{
    if (!t || nModels <= 0 || !sections || !kval || !status) return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    const Symbolic &S = t->S;
    Staged sec, kv, stt;
    const size_t bSec = (size_t)nModels * S.nB * 3 * sizeof(double), bK = (size_t)nModels * S.nnzK * sizeof(double),
                 bSt = (size_t)nModels * sizeof(int32_t);
    if ((rc = stage_in(sections, bSec, sec)) != HDB_OK) return rc;
    if ((rc = stage_out(kval, bK, kv)) != HDB_OK) return rc;
    if ((rc = stage_out(status, bSt, stt)) != HDB_OK) return rc;
    prof_reset(t);
    rc = static_enqueue(t, nModels, (const double *)sec.dev, 0, nullptr, nullptr, nullptr, nullptr, (int32_t *)stt.dev,
                        nullptr, (double *)kv.dev, nullptr, false);
    if (rc != HDB_OK) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    if ((rc = fetch_out(kval, bK, kv)) != HDB_OK) return rc;
    return fetch_out(status, bSt, stt);
}

int hdb_element_batch(hdb_topology *t, int32_t nModels, const double *sections, double *klocal, double *tr,
                      double *kglobal)
{
    if (!t || nModels <= 0 || !sections) return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    const Symbolic &S = t->S;
    Staged sec, a, b, c;
    const size_t bSec = (size_t)nModels * S.nB * 3 * sizeof(double), bM = (size_t)nModels * S.nB * 36 * sizeof(double);
    if ((rc = stage_in(sections, bSec, sec)) != HDB_OK) return rc;
    if ((rc = stage_out(klocal, bM, a)) != HDB_OK) return rc;
    if ((rc = stage_out(tr, bM, b)) != HDB_OK) return rc;
    if ((rc = stage_out(kglobal, bM, c)) != HDB_OK) return rc;
    CUDA_TRY(launch_elements(t->D, nModels, (const double *)sec.dev, (double *)a.dev, (double *)b.dev, (double *)c.dev,
                             nullptr));
    count_launch(1);
    CUDA_TRY(cudaDeviceSynchronize());
    if ((rc = fetch_out(klocal, bM, a)) != HDB_OK) return rc;
    if ((rc = fetch_out(tr, bM, b)) != HDB_OK) return rc;
    return fetch_out(kglobal, bM, c);
}

// ------------------------------------------------------------------------------------------------
// C ABI: Modal batch (kernels in kernels_modal.cu)
// ------------------------------------------------------------------------------------------------
int hdb_modal_mode_count(const hdb_topology *t, int32_t nMass, const int64_t *massN)
{
    if (!t || nMass < 0 || (nMass > 0 && !massN)) return HDB_ERR_ARG;
    std::vector<int32_t> dofs;
    return modal_mass_rows(t->S, nMass, massN, dofs, nullptr);
}

int hdb_modal_batch_dev(hdb_topology *t, int32_t nModels, const double *sections, int32_t nMass, const int64_t *massN,
                        const double *mass, int32_t maxModes, int32_t *nModes, double *hz, double *shapes,
                        int32_t *status, void *stream)
{
    if (!t || nModels < 0 || !sections || nMass < 0 || (nMass > 0 && (!massN || !mass)) || !nModes || !hz || !status ||
        maxModes < 0)
        return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0) return HDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Nvtx nvAll("hdb_modal_batch_dev");
    prof_reset(t);
    const Symbolic &S = t->S;
    std::vector<int32_t> dofs;
    std::vector<std::vector<int32_t>> massOf;
    const int s = modal_mass_rows(S, nMass, massN, dofs, &massOf);
    if (s < 0) {
        set_error_detail("hdb_modal_batch: mass point index out of range");
        return HDB_ERR_INPUT;
    }
    if (s == 0) { // empty mass list: H = 0, no results and no error (model.go:772-805)
        CUDA_TRY(cudaMemsetAsync(nModes, 0, (size_t)nModels * sizeof(int32_t), st));
        CUDA_TRY(cudaMemsetAsync(status, 0, (size_t)nModels * sizeof(int32_t), st));
        return HDB_OK;
    }
    if (maxModes < s) {
        set_error_detail("hdb_modal_batch: maxModes is smaller than hdb_modal_mode_count()");
        return HDB_ERR_SMALL_BUFFER;
    }
    if (s > modal_max_s(t->smemOptin)) {
        set_error_detail("hdb_modal_batch: too many mass DOFs for the shared-memory resident eigen solver");
        return HDB_ERR_UNSUPPORTED;
    }
    // mass-row tables (small; rebuilt per call)
    std::vector<int32_t> mptr(1, 0), midx;
    for (auto &l : massOf) {
        midx.insert(midx.end(), l.begin(), l.end());
        mptr.push_back((int32_t)midx.size());
    }
    const size_t per = (size_t)s * S.dof * sizeof(double);
    // the split pipeline (tridiagonalisation | all QL recurrences at once | rotations + shapes) needs per model: Q, the
    // tridiagonal, and the log of the QL rotations
    const bool split = modal_split_fits(s, t->smemOptin) && !std::getenv("HDB_MODAL_SINGLE");
    size_t qDoubles = 0, rotCap = 0, maxSweeps = 0;
    if (split) modal_split_scratch(s, qDoubles, rotCap, maxSweeps);
    auto al256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t perSplit = split ? al256(qDoubles * 8) + al256((size_t)2 * s * 8) + al256(rotCap * 16) + al256(maxSweeps * 16) + 256 : 0;
    const size_t perAll = 2 * per + (size_t)s * sizeof(double) + perSplit;
    int sbm = (int)std::max<size_t>(1, std::min<size_t>((size_t)nModels, ((size_t)6 << 30) / perAll));
    sbm = std::min(sbm, t->subBatch);
    const size_t tabBytes = (dofs.size() + mptr.size() + std::max<size_t>(midx.size(), 1)) * sizeof(int32_t);
    const size_t need = al256(2 * per * sbm) + al256((size_t)sbm * s * sizeof(double)) + al256(tabBytes) + perSplit * sbm + 4096;
    if (need > t->modalBytes) {
        if (t->modalBuf) {
            CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(t->modalBuf);
        }
        t->modalBuf = nullptr;
        t->modalBytes = 0;
        CUDA_TRY(cudaMalloc(&t->modalBuf, need));
        t->modalBytes = need;
    }
    char *base = (char *)t->modalBuf;
    double *unitLoads = (double *)base;
    double *sol = (double *)(base + per * sbm);
    char *cur = base + al256(2 * per * sbm);
    double *msqrt = (double *)cur;
    cur += al256((size_t)sbm * s * sizeof(double));
    int32_t *dMassDof = (int32_t *)cur;
    int32_t *dMptr = dMassDof + dofs.size();
    int32_t *dMidx = dMptr + mptr.size();
    cur += al256(tabBytes);
    double *qScratch = nullptr, *deScratch = nullptr;
    double2 *rotLog = nullptr;
    int4 *rotHdr = nullptr;
    int32_t *rotSweeps = nullptr;
    if (split) {
        qScratch = (double *)cur; cur += al256(qDoubles * 8) * sbm;
        deScratch = (double *)cur; cur += al256((size_t)2 * s * 8) * sbm;
        rotLog = (double2 *)cur; cur += al256(rotCap * 16) * sbm;
        rotHdr = (int4 *)cur; cur += al256(maxSweeps * 16) * sbm;
        rotSweeps = (int32_t *)cur;
    }
    CUDA_TRY(cudaMemcpyAsync(dMassDof, dofs.data(), dofs.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(dMptr, mptr.data(), mptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (!midx.empty())
        CUDA_TRY(cudaMemcpyAsync(dMidx, midx.data(), midx.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st)); // the host vectors above go out of scope

    for (int m0 = 0; m0 < nModels; m0 += sbm) {
        const int nm = std::min(sbm, nModels - m0);
        ModalLaunch M{};
        M.nModels = nm; M.s = s; M.dof = S.dof; M.nMass = nMass;
        M.massDof = dMassDof; M.mptr = dMptr; M.midx = dMidx;
        M.mass = mass + (size_t)m0 * nMass;
        M.msqrt = msqrt; M.unitLoads = unitLoads; M.sol = sol;
        M.maxModes = maxModes;
        M.nModes = nModes + m0;
        M.hz = hz + (size_t)m0 * maxModes;
        M.shapes = shapes ? shapes + (size_t)m0 * maxModes * S.dof : nullptr;
        M.status = status + m0;
        if (split) { // strides of the per-model scratch records are the aligned sizes above
            M.qScratch = qScratch; M.deScratch = deScratch; M.rotLog = rotLog; M.rotHdr = rotHdr; M.rotSweeps = rotSweeps;
            M.rotCap = (int)(al256(rotCap * 16) / 16); M.rotMaxSweeps = (int)(al256(maxSweeps * 16) / 16);
        }
        {
            ProfScope ps(t, KIND_MODAL0, st);
            CUDA_TRY(launch_modal_loads(M, st));
        }
        count_launch(1);
        // K^-1 applied to the s unit loads of every model: the LinearStatic kernels with s load cases
        rc = static_enqueue(t, nm, sections + (size_t)m0 * S.nB * 3, s, unitLoads, sol, nullptr, nullptr, status + m0,
                            nullptr, nullptr, st, true);
        if (rc != HDB_OK) return rc;
        {
            Nvtx nv("hdb:modal eigen");
            ProfScope ps(t, KIND_MODAL0 + 1, st);
            cudaError_t e = split ? launch_modal_split(M, t->smemOptin, st) : launch_modal_eigen(M, t->smemOptin, st);
            if (e == cudaErrorInvalidConfiguration) {
                set_error_detail("modal eigen kernel does not fit shared memory");
                return HDB_ERR_UNSUPPORTED;
            }
            CUDA_TRY(e);
        }
        count_launch(split ? 3 : 1);
    }
    return HDB_OK;
}

int hdb_modal_batch(hdb_topology *t, int32_t nModels, const double *sections, int32_t nMass, const int64_t *massN,
                    const double *mass, int32_t maxModes, int32_t *nModes, double *hz, double *shapes, int32_t *status)
{
    if (!t || nModels < 0 || !sections || nMass < 0 || (nMass > 0 && (!massN || !mass)) || !nModes || !hz || !status ||
        maxModes < 0)
        return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0) return HDB_OK;
    if (maxModes == 0 && hdb_modal_mode_count(t, nMass, massN) == 0) {
        for (int i = 0; i < nModels; i++) nModes[i] = status[i] = 0;
        return HDB_OK;
    }
    if (is_device_ptr(sections)) {
        rc = hdb_modal_batch_dev(t, nModels, sections, nMass, massN, mass, maxModes, nModes, hz, shapes, status, nullptr);
        if (rc != HDB_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        return HDB_OK;
    }
    const Symbolic &S = t->S;
    Staged sec, ms, nm, hzS, sh, stt;
    const size_t bSec = (size_t)nModels * S.nB * 3 * sizeof(double), bMass = (size_t)nModels * nMass * sizeof(double);
    const size_t bN = (size_t)nModels * sizeof(int32_t), bHz = (size_t)nModels * maxModes * sizeof(double);
    const size_t bSh = (size_t)nModels * maxModes * S.dof * sizeof(double);
    if ((rc = stage_in(sections, bSec, sec)) != HDB_OK) return rc;
    if ((rc = stage_in(mass, bMass, ms)) != HDB_OK) return rc;
    if ((rc = stage_out(nModes, bN, nm)) != HDB_OK) return rc;
    if ((rc = stage_out(hz, bHz, hzS)) != HDB_OK) return rc;
    if ((rc = stage_out(shapes, bSh, sh)) != HDB_OK) return rc;
    if ((rc = stage_out(status, bN, stt)) != HDB_OK) return rc;
    rc = hdb_modal_batch_dev(t, nModels, (const double *)sec.dev, nMass, massN, (const double *)ms.dev, maxModes,
                             (int32_t *)nm.dev, (double *)hzS.dev, (double *)sh.dev, (int32_t *)stt.dev, nullptr);
    if (rc != HDB_OK) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    if ((rc = fetch_out(nModes, bN, nm)) != HDB_OK) return rc;
    if ((rc = fetch_out(hz, bHz, hzS)) != HDB_OK) return rc;
    if ((rc = fetch_out(shapes, bSh, sh)) != HDB_OK) return rc;
    return fetch_out(status, bN, stt);
}

static int ensure_buckling_tables(hdb_topology *t)
{
    if (t->bkReady) return HDB_OK;
    const Symbolic &S = t->S;
    const int s = S.nFree, nB = S.nB;
    std::vector<double> gUnit((size_t)nB * 36);
    for (int b = 0; b < nB; b++) { // feBeam2d.go:175-275 for unit axial force, rotated like model.go:619-620
        const int n0 = S.beamN32[2 * b], n1 = S.beamN32[2 * b + 1];
        const double x0 = S.points[2 * n0], y0 = S.points[2 * n0 + 1], x1 = S.points[2 * n1], y1 = S.points[2 * n1 + 1];
        const double len = beam_length(x0, y0, x1, y1);
        double c, sn, k[36];
        beam_cs(x0, y0, x1, y1, len, c, sn);
        geom_unit_beam2d(len, S.pinMask[b], k);
        rotate_to_global(k, c, sn, gUnit.data() + (size_t)b * 36);
    }
    // contributions to the dense G(i, j) over free unknowns, position i + j*s, in beam order
    std::vector<std::vector<int32_t>> lists((size_t)s * s);
    for (int b = 0; b < nB; b++)
        for (int a = 0; a < 6; a++)
            for (int c2 = 0; c2 < 6; c2++) {
                if (gUnit[(size_t)b * 36 + a * 6 + c2] == 0.0) continue;
                const int ga = 3 * S.beamN32[2 * b + a / 3] + a % 3, gc = 3 * S.beamN32[2 * b + c2 / 3] + c2 % 3;
                const int i = S.freeIndex[ga], j = S.freeIndex[gc];
                if (i < 0 || j < 0) continue;
                lists[(size_t)i + (size_t)j * s].push_back(b * 36 + a * 6 + c2);
            }
    std::vector<int32_t> gPtr(1, 0), gIdx;
    for (auto &l : lists) {
        gIdx.insert(gIdx.end(), l.begin(), l.end());
        gPtr.push_back((int32_t)gIdx.size());
    }
    int rc;
    if ((rc = upload(t, S.freeDofs, &t->bkFreeDof)) != HDB_OK) return rc;
    if ((rc = upload(t, gPtr, &t->bkGPtr)) != HDB_OK) return rc;
    if ((rc = upload(t, gIdx, &t->bkGIdx)) != HDB_OK) return rc;
    if ((rc = upload(t, gUnit, &t->bkGUnit)) != HDB_OK) return rc;
    t->bkReady = true;
    return HDB_OK;
}

int hdb_buckling_max_free(hdb_topology *t)
{
    if (!t || ensure_device(t) != HDB_OK) return 0;
    return buckling_max_s(t->smemOptin);
}

int hdb_buckling_batch_dev(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *forces,
                           int32_t amount, int32_t maxOut, int32_t *nOut, double *factors, double *shapes, int32_t *status,
                           void *stream)
{
    if (!t || nModels < 0 || nLC < 0 || !sections || (nLC > 0 && !forces) || !nOut || !factors || !status || amount < 0 ||
        maxOut < 0)
        return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0 || nLC == 0) return HDB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    prof_reset(t);
    const Symbolic &S = t->S;
    const int s = S.nFree;
    if (s == 0 || amount == 0) {
        CUDA_TRY(cudaMemsetAsync(nOut, 0, (size_t)nModels * nLC * sizeof(int32_t), st));
        CUDA_TRY(cudaMemsetAsync(status, 0, (size_t)nModels * sizeof(int32_t), st));
        return HDB_OK;
    }
    if ((rc = ensure_buckling_tables(t)) != HDB_OK) return rc;
    const size_t per = (size_t)s * S.dof * sizeof(double);
    // models whose three s x s work matrices do not fit shared memory keep them in a global scratch record per (model, case)
    const size_t scrD = buckling_scratch_doubles(s, S.nB, t->smemOptin);
    const size_t perAll = 2 * per + scrD * sizeof(double) * (size_t)nLC;
    int sbm = (int)std::max<size_t>(1, std::min<size_t>((size_t)nModels, ((size_t)4 << 30) / perAll));
    sbm = std::min(sbm, t->subBatch);
    const size_t need = perAll * sbm + 1024;
    if (need > t->modalBytes) {
        if (t->modalBuf) {
            CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(t->modalBuf);
        }
        t->modalBuf = nullptr;
        t->modalBytes = 0;
        CUDA_TRY(cudaMalloc(&t->modalBuf, need));
        t->modalBytes = need;
    }
    double *unitLoads = (double *)t->modalBuf;
    double *sol = (double *)((char *)t->modalBuf + per * sbm);
    double *gscr = scrD ? (double *)((char *)t->modalBuf + 2 * per * sbm) : nullptr;
    for (int m0 = 0; m0 < nModels; m0 += sbm) {
        const int nm = std::min(sbm, nModels - m0);
        CUDA_TRY(launch_unit_loads(unitLoads, nm, s, S.dof, t->bkFreeDof, st));
        count_launch(1);
        // (K^-1)_ff column by column: the LinearStatic kernels with one unit load case per free DOF
        rc = static_enqueue(t, nm, sections + (size_t)m0 * S.nB * 3, s, unitLoads, sol, nullptr, nullptr, status + m0,
                            nullptr, nullptr, st, true);
        if (rc != HDB_OK) return rc;
        BucklingLaunch B{};
        B.nModels = nm; B.nLC = nLC; B.s = s; B.dof = S.dof; B.nB = S.nB; B.amount = amount; B.maxOut = maxOut;
        B.freeDof = t->bkFreeDof; B.gPtr = t->bkGPtr; B.gIdx = t->bkGIdx; B.gUnit = t->bkGUnit;
        B.sol = sol;
        B.forces = forces + (size_t)m0 * nLC * S.nB * 6;
        B.nOut = nOut + (size_t)m0 * nLC;
        B.factors = factors + (size_t)m0 * nLC * maxOut;
        B.shapes = shapes ? shapes + (size_t)m0 * nLC * maxOut * S.dof : nullptr;
        B.status = status + m0;
        {
            ProfScope ps(t, KIND_MODAL0 + 1, st);
            cudaError_t e = launch_buckling_eigen(B, t->smemOptin, st, gscr);
            if (e == cudaErrorInvalidConfiguration) {
                set_error_detail("buckling eigen kernel does not fit shared memory");
                return HDB_ERR_UNSUPPORTED;
            }
            CUDA_TRY(e);
        }
        count_launch(1);
    }
    return HDB_OK;
}

int hdb_buckling_batch(hdb_topology *t, int32_t nModels, const double *sections, int32_t nLC, const double *forces,
                       int32_t amount, int32_t maxOut, int32_t *nOut, double *factors, double *shapes, int32_t *status)
{
    if (!t || nModels < 0 || nLC < 0 || !sections || (nLC > 0 && !forces) || !nOut || !factors || !status || amount < 0 ||
        maxOut < 0)
        return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0 || nLC == 0) return HDB_OK;
    if (is_device_ptr(sections)) {
        rc = hdb_buckling_batch_dev(t, nModels, sections, nLC, forces, amount, maxOut, nOut, factors, shapes, status, nullptr);
        if (rc != HDB_OK) return rc;
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        return HDB_OK;
    }
    const Symbolic &S = t->S;
    Staged sec, fr, no, fa, sh, stt;
    const size_t bSec = (size_t)nModels * S.nB * 3 * sizeof(double), bFr = (size_t)nModels * nLC * S.nB * 6 * sizeof(double);
    const size_t bN = (size_t)nModels * nLC * sizeof(int32_t), bFa = (size_t)nModels * nLC * maxOut * sizeof(double);
    const size_t bSh = bFa * S.dof, bSt = (size_t)nModels * sizeof(int32_t);
    if ((rc = stage_in(sections, bSec, sec)) != HDB_OK) return rc;
    if ((rc = stage_in(forces, bFr, fr)) != HDB_OK) return rc;
    if ((rc = stage_out(nOut, bN, no)) != HDB_OK) return rc;
    if ((rc = stage_out(factors, bFa, fa)) != HDB_OK) return rc;
    if ((rc = stage_out(shapes, bSh, sh)) != HDB_OK) return rc;
    if ((rc = stage_out(status, bSt, stt)) != HDB_OK) return rc;
    rc = hdb_buckling_batch_dev(t, nModels, (const double *)sec.dev, nLC, (const double *)fr.dev, amount, maxOut,
                                (int32_t *)no.dev, (double *)fa.dev, (double *)sh.dev, (int32_t *)stt.dev, nullptr);
    if (rc != HDB_OK) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    if ((rc = fetch_out(nOut, bN, no)) != HDB_OK) return rc;
    if ((rc = fetch_out(factors, bFa, fa)) != HDB_OK) return rc;
    if ((rc = fetch_out(shapes, bSh, sh)) != HDB_OK) return rc;
    return fetch_out(status, bSt, stt);
}

int hdb_uniform_loads_dev(hdb_topology *t, int32_t nModels, int32_t nLC, int32_t nU, const int32_t *ubeam, const uint8_t *uproj,
                          const double *q, int32_t nN, const int64_t *nodes, const double *f, double *loads, void *stream)
{
    if (!t || nModels < 0 || nLC < 0 || nU < 0 || nN < 0 || (nU > 0 && (!ubeam || !uproj || !q)) || (nN > 0 && (!nodes || !f)) ||
        !loads)
        return HDB_ERR_ARG;
    int rc = ensure_device(t);
    if (rc != HDB_OK) return rc;
    if (nModels == 0 || nLC == 0) return HDB_OK;
    const Symbolic &S = t->S;
    cudaStream_t st = (cudaStream_t)stream;
    // per-DOF contribution lists in LoadNodes order (mod/uniform.go:108-210, then the node loads)
    std::vector<std::vector<int32_t>> ul((size_t)S.dof), nl((size_t)S.dof);
    std::vector<double> geo((size_t)nU * 4);
    for (int u = 0; u < nU; u++) {
        const int b = ubeam[u];
        if (b < 0 || b >= S.nB) {
            set_error_detail("hdb_uniform_loads: index of beam is outside of model slice");
            return HDB_ERR_INPUT;
        }
        const int n0 = S.beamN32[2 * b], n1 = S.beamN32[2 * b + 1];
        const double dx = S.points[2 * n0] - S.points[2 * n1], dy = S.points[2 * n0 + 1] - S.points[2 * n1 + 1];
        double lx, ly;
        if (uproj[u]) { lx = std::fabs(dx); ly = std::fabs(dy); }
        else { lx = std::sqrt(dx * dx + dy * dy); ly = lx; }
        geo[4 * u] = dx; geo[4 * u + 1] = dy; geo[4 * u + 2] = lx; geo[4 * u + 3] = ly;
        auto add = [&](int node, int comp, int code) { ul[(size_t)3 * node + comp].push_back(u * 8 + code); };
        if (dx > 0) { add(n0, 2, 1); add(n1, 2, 0); }
        else if (dx < 0) { add(n0, 2, 0); add(n1, 2, 1); }
        if (dy > 0) { add(n0, 2, 2); add(n1, 2, 3); }
        else if (dy < 0) { add(n0, 2, 3); add(n1, 2, 2); }
        for (int n : {n0, n1}) { add(n, 0, 4); add(n, 1, 5); }
    }
    for (int j = 0; j < nN; j++) {
        if (nodes[j] < 0 || nodes[j] >= S.nP) {
            set_error_detail("hdb_uniform_loads: index of point is outside of model slice");
            return HDB_ERR_INPUT;
        }
        for (int c = 0; c < 3; c++) nl[(size_t)3 * nodes[j] + c].push_back(j * 3 + c);
    }
    std::vector<int32_t> uPtr(1, 0), uIdx, nPtr(1, 0), nIdx;
    for (int g = 0; g < S.dof; g++) {
        uIdx.insert(uIdx.end(), ul[g].begin(), ul[g].end());
        uPtr.push_back((int32_t)uIdx.size());
        nIdx.insert(nIdx.end(), nl[g].begin(), nl[g].end());
        nPtr.push_back((int32_t)nIdx.size());
    }
    const size_t bI = (uPtr.size() + nPtr.size() + std::max<size_t>(uIdx.size(), 1) + std::max<size_t>(nIdx.size(), 1)) * sizeof(int32_t);
    const size_t bG = std::max<size_t>(geo.size(), 1) * sizeof(double);
    char *tab = nullptr;
    CUDA_TRY(cudaMalloc(&tab, bG + bI + 64));
    double *dGeo = (double *)tab;
    int32_t *dUPtr = (int32_t *)(tab + bG), *dNPtr = dUPtr + uPtr.size(), *dUIdx = dNPtr + nPtr.size();
    int32_t *dNIdx = dUIdx + std::max<size_t>(uIdx.size(), 1);
    auto cp = [&](void *d, const void *h, size_t b) { return b ? cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, st) : cudaSuccess; };
    cudaError_t e = cp(dGeo, geo.data(), geo.size() * sizeof(double));
    if (e == cudaSuccess) e = cp(dUPtr, uPtr.data(), uPtr.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cp(dNPtr, nPtr.data(), nPtr.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cp(dUIdx, uIdx.data(), uIdx.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cp(dNIdx, nIdx.data(), nIdx.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaStreamSynchronize(st); // the host vectors go out of scope
    if (e == cudaSuccess) {
        UniformLaunch U{};
        U.nModels = nModels; U.nLC = nLC; U.nU = nU; U.nN = nN; U.dof = S.dof;
        U.uPtr = dUPtr; U.uIdx = dUIdx; U.uGeo = dGeo; U.nPtr = dNPtr; U.nIdx = dNIdx;
        U.q = q; U.f = nN > 0 ? f : nullptr; U.loads = loads;
        e = launch_uniform_loads(U, st);
        count_launch(1);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st); // the tables are freed below
    cudaFree(tab);
    CUDA_TRY(e);
    return HDB_OK;
}

} // extern "C"
