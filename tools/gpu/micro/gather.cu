// Micro-benchmark: cost of a random 32-byte gather per lane (what one BVH inner-node step does), L2-resident table.
//   mode 0: 2 x LDG.128      mode 1: 1 x LDG.256      mode 2: 4 x LDG.128 (64-byte record)
//   mode 3: 2 x LDS.128 from a 64 KB shared-memory table     mode 4: 1 x LDG.128 (16 bytes)
// Each lane chases its own pseudo-random chain (dependent loads, like a traversal); many warps hide the latency.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(32) V8 { unsigned x[8]; };
__device__ __forceinline__ V8 ld256(const void* p) {
  V8 r;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
    : "=r"(r.x[0]),"=r"(r.x[1]),"=r"(r.x[2]),"=r"(r.x[3]),"=r"(r.x[4]),"=r"(r.x[5]),"=r"(r.x[6]),"=r"(r.x[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <int MODE>
__global__ void __launch_bounds__(128) k(const unsigned char* table, uint32_t mask, int steps, unsigned* out)
{
  extern __shared__ uint4 sm[];
  if (MODE == 3) { for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = ld128(table + 16 * size_t(i)); __syncthreads(); }
  uint32_t idx = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  unsigned acc = 0;
  for (int s = 0; s < steps; ++s)
  {
    if (MODE == 0) { const unsigned char* p = table + 32 * size_t(idx & mask); uint4 a = ld128(p), b = ld128(p + 16); acc += a.x ^ b.y; idx = idx * 1664525u + 1013904223u + (a.w & 1u) + (b.w & 1u); }
    if (MODE == 1) { const unsigned char* p = table + 32 * size_t(idx & mask); V8 v = ld256(p); acc += v.x[0] ^ v.x[5]; idx = idx * 1664525u + 1013904223u + (v.x[3] & 1u) + (v.x[7] & 1u); }
    if (MODE == 2) { const unsigned char* p = table + 64 * size_t(idx & (mask >> 1)); uint4 a = ld128(p), b = ld128(p + 16), c = ld128(p + 32), d = ld128(p + 48); acc += a.x ^ b.y ^ c.z ^ d.w; idx = idx * 1664525u + 1013904223u + (a.w & 1u) + (d.w & 1u); }
    if (MODE == 3) { const uint4* p = sm + 2 * (idx & 2047u); uint4 a = p[0], b = p[1]; acc += a.x ^ b.y; idx = idx * 1664525u + 1013904223u + (a.w & 1u) + (b.w & 1u); }
    if (MODE == 4) { const unsigned char* p = table + 32 * size_t(idx & mask); uint4 a = ld128(p); acc += a.x; idx = idx * 1664525u + 1013904223u + (a.w & 1u); }
    if (MODE == 5) { const unsigned char* p = table + 64 * size_t(idx & (mask >> 1)); V8 a = ld256(p), b = ld256(p + 32); acc += a.x[0] ^ b.x[1]; idx = idx * 1664525u + 1013904223u + (a.x[3] & 1u) + (b.x[7] & 1u); }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char* name, const unsigned char* table, uint32_t mask, unsigned* out, int blocksPerSM)
{
  const int steps = 2000, grid = 148 * blocksPerSM;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t smem = MODE == 3 ? 65536 : 0;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE><<<grid, 128, smem>>>(table, mask, 100, out);
  cudaEventRecord(e0);
  k<MODE><<<grid, 128, smem>>>(table, mask, steps, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double lanesteps = double(grid) * 128 * steps;
  printf("%-34s blocks/SM %2d  table %6.1f MB  %8.3f ms  %7.2f G lane-steps/s  %6.2f cyc/warp-step/SM @1.9GHz  err=%d\n", name, blocksPerSM, (mask + 1) * 32.0 / 1e6, ms,
         lanesteps / ms / 1e6, 1.9e9 * 148 / (lanesteps / 32 / (ms * 1e-3)), (int)cudaGetLastError());
}

int main()
{
  for (uint32_t entries : {1u << 12, 1u << 18, 1u << 20})
  {
    unsigned char* table; unsigned* out;
    cudaMalloc(&table, size_t(entries) * 32); cudaMalloc(&out, 148 * 16 * 128 * 4);
    cudaMemset(table, 1, size_t(entries) * 32);
    for (int b : {4, 8, 12})
    {
      run<4>("1 x LDG.128 (16 B)", table, entries - 1, out, b);
      run<0>("2 x LDG.128 (32 B)", table, entries - 1, out, b);
      run<1>("1 x LDG.256 (32 B)", table, entries - 1, out, b);
      run<2>("4 x LDG.128 (64 B)", table, entries - 1, out, b);
      run<5>("2 x LDG.256 (64 B)", table, entries - 1, out, b);
      if (b <= 3 * 1 + 0 || b == 4) run<3>("2 x LDS.128 (32 B, 64 KB table)", table, entries - 1, out, b > 3 ? 3 : b);
    }
    cudaFree(table); cudaFree(out);
  }
  return 0;
}
