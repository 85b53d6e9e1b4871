/*
 * tma_ptx.cuh -- inline-PTX wrappers for the sm_100a asynchronous-copy machinery used by
 * the tiled kernel: mbarrier, TMA tensor load / store (SASS: UTMALDG / UTMASTG), proxy fence.
 */
#pragma once

#include <cuda.h>
#include <stdint.h>

namespace hdb {

/* ------------------------------------------------------------------ */
/* PTX: mbarrier, TMA, proxy fence                                     */
/* ------------------------------------------------------------------ */
__device__ __forceinline__ uint32_t SmemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void MbarInit(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void FenceMbarInit()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void MbarArriveExpectTx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void MbarWait(uint32_t bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}

/* generic-proxy writes to shared memory -> visible to the async proxy (TMA store) */
__device__ __forceinline__ void FenceProxyAsyncShared()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void TmaLoad3D(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
		" [%0], [%1, {%2, %3, %4}], [%5];"
		:: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

__device__ __forceinline__ void TmaStore3D(const CUtensorMap *map, uint32_t src, int x, int y, int z)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
		:: "l"(map), "r"(src), "r"(x), "r"(y), "r"(z) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int kPending>
__device__ __forceinline__ void TmaStoreWaitRead()
{
	asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(kPending) : "memory");
}

/* ------------------------------------------------------------------ */
/* cp.async (SASS: LDGSTS) and peer loads                              */
/* ------------------------------------------------------------------ */
/* 4 bytes global -> shared, asynchronously; src_bytes = 0 writes zeros instead (the address is not read) */
__device__ __forceinline__ void CpAsync4(uint32_t dst, const void *src, uint32_t src_bytes)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void CpAsync8(uint32_t dst, const void *src, uint32_t src_bytes)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void CpAsyncCommit()
{
	asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int kPending>
__device__ __forceinline__ void CpAsyncWait()
{
	asm volatile("cp.async.wait_group %0;" :: "n"(kPending) : "memory");
}

/* one word from another GPU's memory (a peer mapping: the load travels over NVLink); L2 only -- the line
 * belongs to a buffer the neighbour refills between launches */
__device__ __forceinline__ uint32_t LdPeer32(const void *p)
{
	uint32_t v;
	asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

} // namespace hdb
