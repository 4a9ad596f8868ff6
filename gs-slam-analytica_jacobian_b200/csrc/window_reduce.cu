// Sum of the keyframe window's packed gradient buffer over the GPUs of one NVSwitch domain, IN the switch.
// Replaces the reduction autograd performs when the window loss is back-propagated (utils/slam_backend.py:160-232; the
// reference is single-GPU) -- the per-iteration ncclAllReduce of window.py -- with one kernel over peer memory:
//
//   every rank holds its gradient buffer in a symmetric allocation that is also mapped through a MULTICAST address (the
//   same offset reaches the buffer of every GPU through the switch).  Rank r owns slice r of the buffer:
//     1. barrier over the ranks (system-scope release / acquire on flags in the peers' signal pads): every rank's local
//        accumulation -- the kernels queued in front of this one on its stream -- is complete and visible;
//     2. multimem.ld_reduce.add.v4.f32 on the slice: the switch reads the 16 bytes from all GPUs, adds them and returns
//        ONE result (1/N of the buffer arrives per GPU instead of (N-1)/N of it in a ring);
//        multimem.st.v4.f32 of the sum: the switch writes it into the buffer of every GPU;
//     3. barrier: all slices are in place everywhere.
//   Per GPU and direction ~1x the buffer crosses NVLink (a ring all-reduce moves 2 (N-1)/N x, in 2 (N-1) latency steps).
//
// No NCCL, no host involvement, graph-capturable; the flags reset themselves (compare-and-swap 0 -> 1 by the sender,
// 1 -> 0 by the receiver), one flag per (CTA, peer).  Waits are bounded (%globaltimer): a peer that never arrives sets
// status[0] = 1 instead of hanging the GPU.
#include "gsr_params.h"

namespace gsr {

namespace {

__device__ __forceinline__ unsigned long long now_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// sender: claim the (empty) flag in the PEER's pad; release: everything this CTA wrote before its barrier is visible to
// whoever acquires the flag
__device__ __forceinline__ bool put_flag(uint32_t* addr, unsigned long long deadline)
{
	unsigned old;
	do {
		asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
		if (old == 0u) return true;
	} while (now_ns() < deadline);
	return false;
}

// receiver: consume the flag a peer set in MY pad
__device__ __forceinline__ bool take_flag(uint32_t* addr, unsigned long long deadline)
{
	unsigned old;
	do {
		asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
		if (old == 1u) return true;
	} while (now_ns() < deadline);
	return false;
}

// CTA b of every rank meets CTA b of every other rank.  pads[p] = signal pad of rank p (peer-mapped), slot (b, sender).
__device__ __forceinline__ void rank_barrier(uint32_t* const* pads, int rank, int world, int* status)
{
	__syncthreads();
	if ((int)threadIdx.x < world) {
		const int peer = threadIdx.x;
		const unsigned long long deadline = now_ns() + 2000000000ull;      // 2 s
		bool ok = put_flag(pads[peer] + (size_t)blockIdx.x * world + rank, deadline);
		ok = take_flag(pads[rank] + (size_t)blockIdx.x * world + peer, deadline) && ok;
		if (!ok) *status = 1;
	}
	__syncthreads();
}

constexpr int kReduceThreads = 512;

template <int kReduceUnroll>
__global__ void __launch_bounds__(kReduceThreads, 1)
window_allreduce_kernel(float4* mc, uint32_t* const* pads, int rank, int world, size_t n4, int* status)
{
	rank_barrier(pads, rank, world, status);
	// slice of this rank, split over the CTAs; consecutive threads -> consecutive 16-byte words
	const size_t per = (n4 + world - 1) / world;
	const size_t lo = min(n4, per * rank), hi = min(n4, lo + per);
	const size_t stride = (size_t)gridDim.x * kReduceThreads;
	size_t i = lo + (size_t)blockIdx.x * kReduceThreads + threadIdx.x;
	for (; i + (kReduceUnroll - 1) * stride < hi; i += kReduceUnroll * stride) {
		float4 v[kReduceUnroll];
#pragma unroll
		for (int u = 0; u < kReduceUnroll; u++)
			asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
			             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + i + u * stride) : "memory");
#pragma unroll
		for (int u = 0; u < kReduceUnroll; u++)
			asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
			             :: "l"(mc + i + u * stride), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
	}
	for (; i < hi; i += stride) {
		float4 v;
		asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
		             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc + i) : "memory");
		asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
		             :: "l"(mc + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
	}
	rank_barrier(pads, rank, world, status);
}

}  // namespace

void launch_window_allreduce(float* multicast, const void* signal_pads, int rank, int world, size_t n_float4, int ctas, int* status,
                             cudaStream_t stream)
{
	// (2, 4 and 8 loads in flight per thread and 16 .. 148 CTAs per rank were measured on 8 GPUs: 0.101 .. 0.113 ms for 28 MB,
	// 0.431 .. 0.473 ms for 168 MB -- the switch, not the issue rate, sets the pace; profiles/r2_switch_reduce_n8.jsonl)
	window_allreduce_kernel<4><<<ctas, kReduceThreads, 0, stream>>>(reinterpret_cast<float4*>(multicast),
	                                                                reinterpret_cast<uint32_t* const*>(signal_pads), rank, world,
	                                                                n_float4, status);
}

}  // namespace gsr
