// Does a packed FFMA2 occupy the issue port for 2 cycles, or only the FMA pipe?  Mix FFMA2 with ALU-pipe integer ops.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pfma(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int NF, int NI, int NL>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i;
  __syncthreads();
  u64 v[8]; unsigned q[8]; float l[8];
  for (int i = 0; i < 8; ++i) { v[i] = ((u64)__float_as_uint(1.f + i) << 32) | __float_as_uint(2.f + threadIdx.x); q[i] = threadIdx.x * 7 + i; l[i] = 0; }
  const u64 pa = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(1.0001f), pb = ((u64)__float_as_uint(1e-4f) << 32) | __float_as_uint(1e-4f);
  const float* sp = sm + (threadIdx.x & 31);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NF) v[i] = pfma(v[i], pa, pb);
      if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(it));
      if (i < NL) { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"((unsigned)__cvta_generic_to_shared(sp + 32 * i + (it & 31) * 33))); l[i] += x; }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += __uint_as_float((unsigned)v[i]) + q[i] + l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NI, int NL>
void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 4 * 1024 * 148); cudaMalloc(&cyc, 8);
  const int iters = 4000;
  k<NF, NI, NL><<<148, warps * 32>>>(out, cyc, iters);
  k<NF, NI, NL><<<148, warps * 32>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("FFMA2=%d LOP3=%d LDS=%d warps/SM=%2d : %6.2f cycles/iter/warp-group(SMSP)  [instr/iter=%d]\n", NF, NI, NL, warps, (double)h / iters / 1.0, NF + NI + NL);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<8, 0, 0>(4); run<0, 8, 0>(4); run<8, 8, 0>(4); run<8, 4, 0>(4); run<0, 0, 8>(4); run<8, 0, 8>(4); run<8, 8, 8>(4);
  run<8, 0, 0>(16); run<8, 8, 0>(16); run<8, 8, 8>(16); run<0, 8, 8>(16);
  return 0;
}
