// gi_comm.cuh -- multi-GPU read-out kernels: snapshot + reduce-scatter of the RGBA32F accumulators over NVLink peer memory (sm_100a).
//
// the path shards by iteration (rank r of W renders iterations k = r mod W into its own accumulator); the only exchange is the sum
// of the accumulators at read-out (the reference's glGetTextureImage of ONE backbuffer, application.cpp:411).  One process per GPU; every rank
// owns an EXCHANGE BLOCK in device memory, exported to its peers as a CUDA IPC handle (gi_comm_create / gi_comm_connect):
//
//     [ flags: ready, done | 256 B ][ snapshot 0 : W x H x RGBA32F ][ snapshot 1 ]
//
// Read-out number e (every rank issues the same sequence of reduced reads):
//   k_comm_snapshot   on the render stream, in order behind the waves submitted so far: waits until every peer has finished reading snapshot
//                     e & 1 at read-out e - 2 (peer.done >= e - 1), copies the accumulator into it and publishes ready = e + 1 with a
//                     system-scope release store.  The renderer goes on at once.
//   k_comm_reduce     on the copy stream, concurrently with the next waves: waits until every peer has published snapshot e (acquire loads
//                     from peer memory), then sums THIS rank's rows of all W snapshots -- peer loads over NVLink, in rank order, so the result
//                     is the same on every run and equal to the sequential sum of the per-rank images -- into a local buffer that one
//                     cudaMemcpyAsync takes to the host, and publishes done = e + 1.
// Each rank therefore moves 1/W of the summed image to the host (33 MB per read-out in aggregate at 1080p instead of W x 33 MB of partial
// images nobody sums), and no NCCL call sits on the render stream.  Spins are bounded (GI_COMM_TIMEOUT_NS): a peer that never arrives sets
// the status word instead of hanging the GPU.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace gik {

#define GI_COMM_MAX_WORLD 16
#define GI_COMM_FLAG_BYTES 256
#define GI_COMM_TIMEOUT_NS 4000000000ull

struct CommPeers
{
	unsigned char* block[GI_COMM_MAX_WORLD]; // exchange blocks in rank order (own block at [rank])
	int world, rank;
};

__device__ __forceinline__ uint32_t commLoadAcquire(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void commStoreRelease(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long commNow()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// lanes 0 .. world-1 of every block wait for flag word `word` (0 ready, 1 done) of their peer to reach `target`
__device__ __forceinline__ void commWaitPeers(const CommPeers& peers, int word, uint32_t target, uint32_t* status)
{
	if (int(threadIdx.x) < peers.world && int(threadIdx.x) != peers.rank)
	{
		const uint32_t* flag = reinterpret_cast<const uint32_t*>(peers.block[threadIdx.x]) + word;
		const unsigned long long start = commNow();
		while (commLoadAcquire(flag) < target)
		{
			if (commNow() - start > GI_COMM_TIMEOUT_NS) { atomicExch(status, 1u + uint32_t(word)); break; }
			__nanosleep(200);
		}
	}
	__syncthreads();
}

// the last block to arrive publishes `value` in the own flag word; all blocks' prior writes are visible system-wide before it
__device__ __forceinline__ void commPublish(uint32_t* ownFlag, uint32_t value, uint32_t* arrive)
{
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0)
	{
		const uint32_t prev = atomicAdd(arrive, 1u);
		if (prev == gridDim.x - 1u)
		{
			*arrive = 0u;
			__threadfence_system();
			commStoreRelease(ownFlag, value);
		}
	}
}

__global__ void __launch_bounds__(256) k_comm_snapshot(const float4* __restrict__ accum, CommPeers peers, uint32_t epoch, uint32_t numPixels, size_t imageBytes, uint32_t* arrive,
                                                        uint32_t* status)
{
	if (epoch >= 2u) commWaitPeers(peers, 1, epoch - 1u, status);
	unsigned char* own = peers.block[peers.rank];
	float4* snap = reinterpret_cast<float4*>(own + GI_COMM_FLAG_BYTES + size_t(epoch & 1u) * imageBytes);
	for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < numPixels; p += gridDim.x * blockDim.x) snap[p] = accum[p];
	commPublish(reinterpret_cast<uint32_t*>(own), epoch + 1u, arrive);
}

__global__ void __launch_bounds__(256) k_comm_reduce(CommPeers peers, uint32_t epoch, uint32_t firstPixel, uint32_t numPixels, size_t imageBytes, float4* __restrict__ out,
                                                      uint32_t* arrive, uint32_t* status)
{
	commWaitPeers(peers, 0, epoch + 1u, status);
	const size_t off = GI_COMM_FLAG_BYTES + size_t(epoch & 1u) * imageBytes;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numPixels; i += gridDim.x * blockDim.x)
	{
		// rank order: the sum is the same on every rank count's every run (fp32 addition is not associative)
		float4 s = reinterpret_cast<const float4*>(peers.block[0] + off)[firstPixel + i];
		for (int r = 1; r < peers.world; ++r)
		{
			const float4 v = reinterpret_cast<const float4*>(peers.block[r] + off)[firstPixel + i];
			s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
		}
		out[i] = s;
	}
	commPublish(reinterpret_cast<uint32_t*>(peers.block[peers.rank]) + 1, epoch + 1u, arrive);
}

} // namespace gik
