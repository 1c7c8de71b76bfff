// Micro-measurements behind the resident layer-stack kernel (stack_sm100.cuh): what do a grid barrier and a read of
// data another SM just wrote cost on this part?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o
// gpurun_in/grid_sync_latency tools/micro/grid_sync_latency.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    for (;;) {
      unsigned v;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (static_cast<int>(v - target) >= 0) break;
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  __syncthreads();
}

// out[0] = cycles of `iters` back-to-back barriers (CTA 0); out[1] = dependent ld.cg chain over L2-resident data
// (per load); out[2] = cycles for one warp to read 16 independent float4 that ANOTHER CTA wrote before the barrier.
__global__ void probe(unsigned* ctr, float4* buf, const unsigned* chase, long long* out, int iters) {
  unsigned target = 0;
  grid_barrier(ctr, target);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) grid_barrier(ctr, target);
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (t1 - t0) / iters;
  // pointer chase
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned p = 0;
    for (int i = 0; i < 64; ++i) p = __ldcg(chase + p);     // warm
    long long a = clock64();
    for (int i = 0; i < 256; ++i) p = __ldcg(chase + p);
    long long b = clock64();
    out[1] = (b - a) / 256;
    out[5] = p;
  }
  // producer/consumer across CTAs
  for (int rep = 0; rep < 4; ++rep) {
    const int src = (blockIdx.x + 1) % gridDim.x;            // I read what my neighbour wrote
    float4* mine = buf + (static_cast<size_t>(blockIdx.x) * 4 + rep) * 16 * 32;
    if (threadIdx.x < 32)
      for (int k = 0; k < 16; ++k) mine[k * 32 + threadIdx.x] = make_float4(rep, k, blockIdx.x, threadIdx.x);
    grid_barrier(ctr, target);
    if (threadIdx.x < 32) {
      const float4* theirs = buf + (static_cast<size_t>(src) * 4 + rep) * 16 * 32;
      long long a = clock64();
      float4 v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = __ldcg(theirs + k * 32 + threadIdx.x);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;
      long long b;
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(b) : "f"(s) : "memory");
      if (blockIdx.x == 0 && threadIdx.x == 0) { out[2] = b - a; out[6] = (long long)s; }
      // one dependent load
      a = clock64();
      float4 w = __ldcg(theirs + threadIdx.x);
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(b) : "f"(w.x) : "memory");
      if (blockIdx.x == 0 && threadIdx.x == 0) out[3] = b - a;
    }
    grid_barrier(ctr, target);
  }
}

int main() {
  unsigned* ctr; float4* buf; unsigned* chase; long long* out;
  cudaMalloc(&ctr, 4); cudaMalloc(&buf, 148 * 4 * 16 * 32 * sizeof(float4)); cudaMalloc(&out, 64);
  const int n = 1 << 20;                      // 4 MB chase table (L2 resident, beyond L1)
  unsigned* h = new unsigned[n];
  for (int i = 0; i < n; ++i) h[i] = (i + 4099 * 33) % n;
  cudaMalloc(&chase, n * 4); cudaMemcpy(chase, h, n * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 2, 16, 32, 80, 148}) {
    for (int threads : {32, 256}) {
      cudaMemset(ctr, 0, 4);
      cudaMemset(out, 0, 64);
      probe<<<grid, threads, 200 * 1024>>>(ctr, buf, chase, out, 64);
      cudaError_t e = cudaDeviceSynchronize();
      long long r[8];
      cudaMemcpy(r, out, 64, cudaMemcpyDeviceToHost);
      printf("grid %3d x %3d thr: barrier %5lld cyc | L2 chase %4lld cyc/load | 16 x float4 written by neighbour %5lld cyc | one such load %5lld cyc  (%s)\n",
             grid, threads, r[0], r[1], r[2], r[3], cudaGetErrorString(e));
    }
  }
  return 0;
}
