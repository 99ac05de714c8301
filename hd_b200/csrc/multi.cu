// multi.cu — one process, several GPUs of one box: the batch is cut into contiguous model ranges, one
// per device; every device has its own topology handle (the symbolic phase is a few kB and is simply repeated), its own
// host thread for calls that take host memory, and its own streams.  No collective sits on the data path: models are
// independent.  NCCL is used only to gather REDUCED results (per-device failure counts and the largest displacement
// magnitude): gathering the full result arrays would cost twenty times the solve.  NCCL is bound at run time
// (dlopen of libnccl.so.2: the copy the host program already loaded, if any), so the library itself has no link-time
// dependency on it; without it hdb_multi_gather_summary returns HDB_ERR_UNSUPPORTED and everything else works.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include "../../include/hdb200.h"

namespace hdb {
void set_error_detail(const std::string &s);

// ---- the handful of NCCL entry points this file uses, bound at run time ----
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
struct Nccl {
    void *so = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int /*ncclDataType_t*/, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
constexpr int NCCL_FLOAT64 = 8; // ncclDouble (nccl.h, stable since 2.0)

static Nccl &nccl()
{
    static Nccl n = [] {
        Nccl q;
        q.so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!q.so) q.so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!q.so) return q;
        q.CommInitAll = (decltype(q.CommInitAll))dlsym(q.so, "ncclCommInitAll");
        q.CommDestroy = (decltype(q.CommDestroy))dlsym(q.so, "ncclCommDestroy");
        q.AllGather = (decltype(q.AllGather))dlsym(q.so, "ncclAllGather");
        q.GroupStart = (decltype(q.GroupStart))dlsym(q.so, "ncclGroupStart");
        q.GroupEnd = (decltype(q.GroupEnd))dlsym(q.so, "ncclGroupEnd");
        q.GetErrorString = (decltype(q.GetErrorString))dlsym(q.so, "ncclGetErrorString");
        q.ok = q.CommInitAll && q.CommDestroy && q.AllGather && q.GroupStart && q.GroupEnd;
        return q;
    }();
    return n;
}

// per-device summary of a shard: {models that failed, largest |displacement| over the healthy models, models, 0}
__global__ void k_shard_summary(const int32_t *status, const double *disp, int nModels, size_t perModel, double *out4)
{
    __shared__ double s_bad[256], s_max[256];
    double bad = 0.0, mx = 0.0;
    for (int m = blockIdx.x; m < nModels; m += gridDim.x) {
        const bool ok = status[m] == HDB_STATUS_OK;
        if (!ok && threadIdx.x == 0) bad += 1.0;
        if (ok && disp)
            for (size_t i = threadIdx.x; i < perModel; i += blockDim.x) mx = fmax(mx, fabs(disp[(size_t)m * perModel + i]));
    }
    s_bad[threadIdx.x] = bad;
    s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            s_bad[threadIdx.x] += s_bad[threadIdx.x + o];
            s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { // one contributor per address and a commutative max: order does not matter for either
        atomicAdd(out4, s_bad[0]);
        unsigned long long *p = reinterpret_cast<unsigned long long *>(out4 + 1);
        atomicMax(p, (unsigned long long)__double_as_longlong(s_max[0])); // non-negative doubles order like their bit patterns
    }
}

} // namespace hdb

using namespace hdb;

struct hdb_multi {
    std::vector<hdb_topology *> topo;
    std::vector<int32_t> dev;
    std::vector<double *> summary;   // per device: [4] this shard | [4 * nDev] gathered
    std::vector<cudaStream_t> stream;
    std::vector<ncclComm_t> comm;
    bool commTried = false;
    int32_t nP = 0, nB = 0, dof = 0;
};

static void shard_range(int32_t nModels, int nDev, int d, int32_t &m0, int32_t &nm)
{ // contiguous ranges, sizes differ by at most one
    const int32_t base = nModels / nDev, rem = nModels % nDev;
    m0 = d * base + std::min<int32_t>(d, rem);
    nm = base + (d < rem ? 1 : 0);
}

extern "C" {

int hdb_multi_create(int32_t nP, const double *points, int32_t nB, const hdb_beam *beams, const uint8_t *pins,
                     const uint8_t *supports, const hdb_options *opts, int32_t nDev, const int32_t *devices, hdb_multi **out)
{
    if (!out || nDev <= 0) return HDB_ERR_ARG;
    *out = nullptr;
    const int visible = hdb_device_count();
    if (visible <= 0) {
        set_error_detail("hdb_multi_create: no CUDA device visible: libhdb200 has no CPU path");
        return HDB_ERR_NO_DEVICE;
    }
    hdb_multi *m = new (std::nothrow) hdb_multi();
    if (!m) return HDB_ERR_NOMEM;
    for (int d = 0; d < nDev; d++) {
        const int32_t id = devices ? devices[d] : d;
        if (id < 0 || id >= visible) {
            set_error_detail("hdb_multi_create: device ordinal out of range");
            hdb_multi_destroy(m);
            return HDB_ERR_ARG;
        }
        hdb_options o{};
        if (opts) o = *opts;
        o.device = id;
        hdb_topology *t = nullptr;
        const int rc = hdb_topology_create(nP, points, nB, beams, pins, supports, &o, &t);
        if (rc != HDB_OK) {
            hdb_multi_destroy(m);
            return rc;
        }
        m->topo.push_back(t);
        m->dev.push_back(id);
    }
    m->nP = nP; m->nB = nB; m->dof = 3 * nP;
    m->summary.assign(nDev, nullptr);
    m->stream.assign(nDev, nullptr);
    *out = m;
    return HDB_OK;
}

void hdb_multi_destroy(hdb_multi *m)
{
    if (!m) return;
    for (size_t d = 0; d < m->comm.size(); d++)
        if (m->comm[d] && nccl().ok) nccl().CommDestroy(m->comm[d]);
    for (size_t d = 0; d < m->topo.size(); d++) {
        if (d < m->dev.size()) cudaSetDevice(m->dev[d]);
        if (d < m->summary.size() && m->summary[d]) cudaFree(m->summary[d]);
        if (d < m->stream.size() && m->stream[d]) cudaStreamDestroy(m->stream[d]);
        hdb_topology_destroy(m->topo[d]);
    }
    delete m;
}

int32_t hdb_multi_device_count(const hdb_multi *m) { return m ? (int32_t)m->topo.size() : 0; }

int hdb_shard_range(int32_t nModels, int32_t nShards, int32_t shard, int32_t *first, int32_t *count)
{
    if (nModels < 0 || nShards < 1 || shard < 0 || shard >= nShards || !first || !count) return HDB_ERR_ARG;
    shard_range(nModels, nShards, shard, *first, *count);
    return HDB_OK;
}

int hdb_multi_shard(const hdb_multi *m, int32_t nModels, int32_t shard, int32_t *first, int32_t *count)
{
    if (!m || nModels < 0 || shard < 0 || shard >= (int32_t)m->topo.size() || !first || !count) return HDB_ERR_ARG;
    shard_range(nModels, (int)m->topo.size(), shard, *first, *count);
    return HDB_OK;
}

hdb_topology *hdb_multi_topology(hdb_multi *m, int32_t shard)
{
    return (m && shard >= 0 && shard < (int32_t)m->topo.size()) ? m->topo[shard] : nullptr;
}

// host memory in, host memory out: one host thread per device runs hdb_static_batch on its contiguous range of models
// and writes straight into the caller's arrays (the ranges are disjoint)
int hdb_multi_static_batch(hdb_multi *m, int32_t nModels, const double *sections, int32_t nLC, const double *loads, double *disp,
                           double *forces, double *react, int32_t *status, int32_t *failCase)
{
    if (!m || nModels < 0 || nLC < 0 || !sections || !status || (nLC > 0 && !loads)) return HDB_ERR_ARG;
    const int nDev = (int)m->topo.size();
    std::vector<int> rc(nDev, HDB_OK);
    std::vector<std::string> detail(nDev);
    std::vector<std::thread> th;
    nvtxRangePushA("hdb_multi_static_batch");
    for (int d = 0; d < nDev; d++) {
        th.emplace_back([&, d] {
            int32_t m0, nm;
            shard_range(nModels, nDev, d, m0, nm);
            if (nm == 0) return;
            const size_t rows = (size_t)m0 * nLC;
            rc[d] = hdb_static_batch(m->topo[d], nm, sections + (size_t)m0 * m->nB * 3, nLC, loads ? loads + rows * m->dof : nullptr,
                                     disp ? disp + rows * m->dof : nullptr, forces ? forces + rows * m->nB * 6 : nullptr,
                                     react ? react + rows * m->dof : nullptr, status + m0, failCase ? failCase + m0 : nullptr);
            if (rc[d] != HDB_OK) detail[d] = hdb_last_error_detail(); // the detail string is per thread
        });
    }
    for (auto &t : th) t.join();
    nvtxRangePop();
    for (int d = 0; d < nDev; d++)
        if (rc[d] != HDB_OK) {
            set_error_detail("device " + std::to_string(m->dev[d]) + ": " + detail[d]);
            return rc[d];
        }
    return HDB_OK;
}

// shards already resident: arrays of per-device pointers (entry d = the arrays of shard d on device d, count[d] models);
// every device's kernels are enqueued on that device's stream[d] (NULL array = default streams) without synchronising
int hdb_multi_static_batch_dev(hdb_multi *m, const int32_t *count, const double *const *sections, int32_t nLC, const double *const *loads,
                               double *const *disp, double *const *forces, double *const *react, int32_t *const *status,
                               int32_t *const *failCase, void *const *streams)
{
    if (!m || !count || !sections || !status || nLC < 0 || (nLC > 0 && !loads)) return HDB_ERR_ARG;
    const int nDev = (int)m->topo.size();
    std::vector<int> rc(nDev, HDB_OK);
    std::vector<std::string> detail(nDev);
    auto enqueue = [&](int d) {
        if (count[d] <= 0) return;
        rc[d] = hdb_static_batch_dev(m->topo[d], count[d], sections[d], nLC, loads ? loads[d] : nullptr, disp ? disp[d] : nullptr,
                                     forces ? forces[d] : nullptr, react ? react[d] : nullptr, status[d],
                                     failCase ? failCase[d] : nullptr, streams ? streams[d] : nullptr);
        if (rc[d] != HDB_OK) detail[d] = hdb_last_error_detail();
    };
    // one host thread per device: enqueueing a step takes a few hundred microseconds of host time per device, which one
    // thread would serialise in front of eight GPUs that need 1.5 ms for it
    if (nDev == 1) enqueue(0);
    else {
        std::vector<std::thread> th;
        for (int d = 0; d < nDev; d++) th.emplace_back(enqueue, d);
        for (auto &t : th) t.join();
    }
    for (int d = 0; d < nDev; d++)
        if (rc[d] != HDB_OK) {
            set_error_detail("device " + std::to_string(m->dev[d]) + ": " + detail[d]);
            return rc[d];
        }
    return HDB_OK;
}

int hdb_multi_synchronize(hdb_multi *m)
{
    if (!m) return HDB_ERR_ARG;
    for (size_t d = 0; d < m->topo.size(); d++) {
        if (cudaSetDevice(m->dev[d]) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            set_error_detail(std::string("hdb_multi_synchronize: ") + cudaGetErrorString(cudaGetLastError()));
            return HDB_ERR_CUDA;
        }
    }
    return HDB_OK;
}

// reduced results of the resident shards, gathered over NCCL (NVLink / NVSwitch): summary[nDev][4] =
// {failed models, largest |displacement| of the healthy models, models, 0} of every shard, on the host.  disp may be NULL.
int hdb_multi_gather_summary(hdb_multi *m, const int32_t *count, const int32_t *const *status, const double *const *disp, int32_t nLC,
                             double *summary)
{
    if (!m || !count || !status || !summary) return HDB_ERR_ARG;
    const int nDev = (int)m->topo.size();
    for (int a = 0; a < nDev; a++)
        for (int b = a + 1; b < nDev; b++)
            if (m->dev[a] == m->dev[b]) {
                set_error_detail("hdb_multi_gather_summary: NCCL needs distinct devices");
                return HDB_ERR_UNSUPPORTED;
            }
    Nccl &N = nccl();
    if (!N.ok) {
        set_error_detail("hdb_multi_gather_summary: libnccl.so.2 not found");
        return HDB_ERR_UNSUPPORTED;
    }
    if (!m->commTried) {
        m->commTried = true;
        m->comm.assign(nDev, nullptr);
        const ncclResult_t r = N.CommInitAll(m->comm.data(), nDev, m->dev.data());
        if (r != 0) {
            m->comm.clear();
            set_error_detail(std::string("ncclCommInitAll: ") + (N.GetErrorString ? N.GetErrorString(r) : "error"));
            return HDB_ERR_CUDA;
        }
    }
    if (m->comm.empty()) return HDB_ERR_UNSUPPORTED;
    for (int d = 0; d < nDev; d++) {
        if (cudaSetDevice(m->dev[d]) != cudaSuccess) return HDB_ERR_CUDA;
        if (!m->summary[d] && cudaMalloc(&m->summary[d], (size_t)4 * (nDev + 1) * sizeof(double)) != cudaSuccess) return HDB_ERR_NOMEM;
        if (!m->stream[d] && cudaStreamCreateWithFlags(&m->stream[d], cudaStreamNonBlocking) != cudaSuccess) return HDB_ERR_CUDA;
        cudaDeviceSynchronize(); // the shard's results are complete (they may have been produced on any stream)
        const double init[4] = {0.0, 0.0, (double)count[d], 0.0};
        cudaMemcpyAsync(m->summary[d], init, sizeof init, cudaMemcpyHostToDevice, m->stream[d]);
        if (count[d] > 0)
            k_shard_summary<<<std::min(count[d], 592), 256, 0, m->stream[d]>>>(status[d], disp ? disp[d] : nullptr, count[d],
                                                                            (size_t)nLC * m->dof, m->summary[d]);
    }
    N.GroupStart();
    for (int d = 0; d < nDev; d++) N.AllGather(m->summary[d], m->summary[d] + 4, 4, NCCL_FLOAT64, m->comm[d], m->stream[d]);
    const ncclResult_t r = N.GroupEnd();
    if (r != 0) {
        set_error_detail(std::string("ncclAllGather: ") + (N.GetErrorString ? N.GetErrorString(r) : "error"));
        return HDB_ERR_CUDA;
    }
    for (int d = 0; d < nDev; d++) {
        cudaSetDevice(m->dev[d]);
        if (cudaStreamSynchronize(m->stream[d]) != cudaSuccess) return HDB_ERR_CUDA;
    }
    cudaSetDevice(m->dev[0]);
    if (cudaMemcpy(summary, m->summary[0] + 4, (size_t)4 * nDev * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return HDB_ERR_CUDA;
    return HDB_OK;
}

} // extern "C"
