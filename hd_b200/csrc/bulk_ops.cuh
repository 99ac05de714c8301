// bulk_ops.cuh — PTX wrappers of k_solve_bulk: bulk copies shared -> global and global -> shared (SASS UBLKCP), mbarrier
// waits, L2 cache policy, proxy fence, single-lane election.
#pragma once
#include "solve_ops.cuh"

namespace hdb {

constexpr int BULK_CHUNK_BYTES = BULK_CHUNK * 8; // one staging slot = one bulk store

// ---- PTX wrappers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// results are written once and not read again by this kernel: evict-first keeps them from displacing the inputs in L2
__device__ __forceinline__ void bulk_store(void *gdst, unsigned ssrc, unsigned bytes, unsigned long long pol)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(ssrc), "r"(bytes), "l"(pol) : "memory");
}
// bulk copy global -> shared; its bytes complete a phase of the mbarrier (one arrival expected per phase)
__device__ __forceinline__ void bulk_load(unsigned sdst, const void *gsrc, unsigned bytes, unsigned mbar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(unsigned mbar, unsigned parity) // try_wait may put the warp to sleep until the phase completes
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void lds_v2s32(unsigned a, int &x, int &y) { asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory"); }
__device__ __forceinline__ bool elect_one()
{ // one lane of the (converged) warp, as a warp-uniform predicate
    unsigned pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0u;
}

// a beam with released ends: the condensed element is dense (feBeam2d.go:58-137).  Out of line, arguments by value and
// the result stored straight to the staging slot, so that the caller keeps no local array
static __device__ __noinline__ void pinned_forces_to_slot(const BeamCoef cf, const uint32_t pm, const double gc, const double gs, const double u0,
                                                          const double u1, const double u2, const double u3, const double u4, const double u5,
                                                          const unsigned sl)
{
    double uu[6] = {u0, u1, u2, u3, u4, u5}, oo[6];
    beam_forces_pinned(cf, pm, gc, gs, uu, oo);
    sts_v2f64(sl, oo[0], oo[1]);
    sts_v2f64(sl + 16u, oo[2], oo[3]);
    sts_v2f64(sl + 32u, oo[4], oo[5]);
}

} // namespace hdb
