// microbenchmark: latency / throughput of packed f32x2 vs scalar FFMA on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pfma(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 padd(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
template <int ILP, int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float s[ILP]; u64 v[ILP];
  for (int i = 0; i < ILP; ++i) { s[i] = threadIdx.x * 0.001f + i; v[i] = ((u64)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] + 1.f); }
  const float a = 1.0001f, b = 0.0001f;
  const u64 pa = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a), pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) s[i] = fmaf(s[i], a, b);
      else if (MODE == 1) v[i] = pfma(v[i], pa, pb);
      else v[i] = padd(v[i], pb);
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < ILP; ++i) acc += s[i] + __uint_as_float((unsigned)v[i]) + __uint_as_float((unsigned)(v[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int MODE>
void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<ILP, MODE><<<148, warps * 32>>>(out, cyc, iters);
  k<ILP, MODE><<<148, warps * 32>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters / ILP;  // cycles per instruction per warp
  // per-SMSP issue rate: warps/4 warps per scheduler each issuing ILP instr per `h/iters` cycles
  printf("%-8s ILP=%2d warps/SM=%2d cycles/instr/warp=%6.2f  instr/clk/SMSP=%5.2f\n", name, ILP, warps, per, (warps / 4.0) / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 0>("FFMA", 4); run<1, 1>("FFMA2", 4); run<1, 2>("FADD2", 4);
  run<4, 0>("FFMA", 4); run<4, 1>("FFMA2", 4); run<4, 2>("FADD2", 4);
  run<8, 0>("FFMA", 4); run<8, 1>("FFMA2", 4); run<8, 2>("FADD2", 4);
  run<8, 0>("FFMA", 8); run<8, 1>("FFMA2", 8); run<8, 2>("FADD2", 8);
  run<8, 0>("FFMA", 16); run<8, 1>("FFMA2", 16); run<8, 2>("FADD2", 16);
  run<2, 0>("FFMA", 20); run<2, 1>("FFMA2", 20);
  run<4, 0>("FFMA", 20); run<4, 1>("FFMA2", 20);
  return 0;
}
