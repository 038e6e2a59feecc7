// microbenchmark: how do the warps of co-resident CTAs map to the 4 SM sub-partitions (schedulers)?
// Every working warp runs the same FFMA2-bound loop; run time is proportional to the largest number of working warps on
// one sub-partition.  Also dumps (%smid, %warpid) per warp to tabulate the mapping, and times a warp-uniform
// constant-bank load (LDC) against a shared-memory broadcast load (LDS.128).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pfma(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__global__ void work(float* out, int* map, int iters, int working_warps, int shift) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5;
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) {
    map[(blockIdx.x * (blockDim.x >> 5) + warp) * 2] = smid;
    map[(blockIdx.x * (blockDim.x >> 5) + warp) * 2 + 1] = wid;
  }
  const int role = (warp + shift * (blockIdx.x & 1)) % (blockDim.x >> 5);
  if (role >= working_warps) return;
  u64 v[8];
  for (int i = 0; i < 8; ++i) v[i] = ((u64)__float_as_uint(threadIdx.x * 0.001f + i) << 32) | __float_as_uint(1.f + i);
  const u64 pa = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(1.0001f), pb = ((u64)__float_as_uint(1e-4f) << 32) | __float_as_uint(1e-4f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = pfma(v[i], pa, pb);
  }
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static float time_work(const char* name, int grid, int threads, size_t smem, int working, int shift, bool dump) {
  float* out; int* map;
  cudaMalloc(&out, 4 * 1024 * 1024); cudaMalloc(&map, 8 * 1024 * 64);
  cudaFuncSetAttribute(work, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  work<<<grid, threads, smem>>>(out, map, 20000, working, shift);
  cudaEventRecord(a);
  work<<<grid, threads, smem>>>(out, map, 20000, working, shift);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-34s grid=%4d threads=%4d working=%2d shift=%d  %.3f ms  (%s)\n", name, grid, threads, working, shift, ms, cudaGetErrorString(cudaGetLastError()));
  if (dump) {
    const int nw = grid * (threads / 32);
    int* h = (int*)malloc(8 * nw);
    cudaMemcpy(h, map, 8 * nw, cudaMemcpyDeviceToHost);
    // SM of block 0: list (block, local warp, hw warpid) of every warp resident there
    const int sm0 = h[0];
    printf("  warps on SM %d:", sm0);
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < nw; ++i)
      if (h[2 * i] == sm0) {
        printf(" b%d.w%d->%d", i / (threads / 32), i % (threads / 32), h[2 * i + 1]);
        const int lw = i % (threads / 32), role = (lw + shift * ((i / (threads / 32)) & 1)) % (threads / 32);
        if (role < working) cnt[h[2 * i + 1] & 3]++;
      }
    printf("\n  working warps per (hw warpid %% 4): %d %d %d %d\n", cnt[0], cnt[1], cnt[2], cnt[3]);
    free(h);
  }
  cudaFree(out); cudaFree(map);
  return ms;
}

// ---- constant-bank vs shared-memory broadcast loads -----------------------------------------------------------------
struct Tab { float4 t[256]; };  // 4 KB by value in the kernel parameter bank
__constant__ float4 c_tab[256];

template <int MODE>
__global__ void bcast(const Tab tab, float* out, int iters) {
  __shared__ float4 s_tab[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tab[i] = tab.t[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = warp * 20;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      float4 w;
      if (MODE == 0) w = s_tab[(idx + j) & 255];
      else if (MODE == 1) w = tab.t[(idx + j) & 255];
      else w = c_tab[(idx + j) & 255];
      acc.x = fmaf(w.x, acc.y, acc.x); acc.y = fmaf(w.y, acc.z, acc.y); acc.z = fmaf(w.z, acc.w, acc.z); acc.w = fmaf(w.w, acc.x, acc.w);
    }
    idx = (idx + 7) & 255;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>
static void time_bcast(const char* name, int warps) {
  Tab h;
  for (int i = 0; i < 256; ++i) h.t[i] = make_float4(1e-3f * i, 1.f, 0.5f, 0.25f);
  cudaMemcpyToSymbol(c_tab, h.t, sizeof(h.t));
  float* out; cudaMalloc(&out, 4 * 1024 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  bcast<MODE><<<148, warps * 32>>>(h, out, iters);
  cudaEventRecord(a);
  bcast<MODE><<<148, warps * 32>>>(h, out, iters);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  // loads per clock per SM
  const double loads = (double)warps * iters * 10;
  printf("%-28s warps/SM=%2d  %.3f ms  -> %.3f 128-bit broadcast loads / clk / SM at 1.9 GHz (%s)\n", name, warps, ms, loads / (ms * 1e-3 * 1.9e9), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  const size_t big = 100 * 1024;  // 2 CTAs / SM
  time_work("2 CTA x 8 warps", 296, 256, big, 8, 0, true);
  time_work("2 CTA x 10 warps", 296, 320, big, 10, 0, true);
  time_work("2 CTA x 12 warps, 10 working", 296, 384, big, 10, 0, true);
  time_work("2 CTA x 12 warps, 10 working, shift 2", 296, 384, big, 10, 2, true);
  time_work("2 CTA x 10 warps, odd CTAs shift 2", 296, 320, big, 10, 2, true);
  time_work("1 CTA x 20 warps", 148, 640, 2 * big, 20, 0, true);
  time_work("2 CTA x 12 warps", 296, 384, big, 12, 0, false);
  time_bcast<0>("LDS.128 broadcast", 8); time_bcast<1>("LDC param bank", 8); time_bcast<2>("LDC __constant__", 8);
  time_bcast<0>("LDS.128 broadcast", 20); time_bcast<1>("LDC param bank", 20); time_bcast<2>("LDC __constant__", 20);
  return 0;
}
