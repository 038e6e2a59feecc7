// Microbenchmark (round 2): warp-uniform constant tables from TMEM (tcgen05.ld, per-lane replicated) vs shared-memory
// broadcast (LDS.128), alone and mixed with lane-strided LDS.64 traffic.  nvcc -arch=sm_100a -O3 -o ubench_tmem ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int WARPS = 20, ITERS = 2048, COLS = 64;  // per-warp table: COLS floats

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float& a, float& b, float& c, float& d) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: LDS.128 broadcast; 1: tcgen05.ld x4; 2: LDS.64 strided only; 3: LDS.64 strided + LDS.128 bcast; 4: LDS.64 strided + tcgen05.ld
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float* out, long long* cycles) {
  extern __shared__ float4 sm4[];
  __shared__ uint32_t s_taddr;
  float* tab = reinterpret_cast<float*>(sm4);                   // [WARPS][COLS]
  float2* strided = reinterpret_cast<float2*>(sm4 + WARPS * COLS / 4);  // [32 lanes][pitch 81 float2] per... shared by all warps
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < WARPS * COLS; i += blockDim.x) tab[i] = 1.0f + 1e-6f * i;
  for (int i = threadIdx.x; i < 32 * 81; i += blockDim.x) strided[i] = make_float2(1.0f, 1e-3f * i);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_taddr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = s_taddr;
  // this warp's columns: quadrant = warp % 4 (lanes 32*(warp%4) ..), columns (warp / 4) * COLS ..
  const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * COLS);
  for (int c = 0; c < COLS; c += 4) {
    const float* t = tab + warp * COLS + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr + c), "f"(t[0]), "f"(t[1]), "f"(t[2]), "f"(t[3]));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const unsigned tab_sa = (unsigned)__cvta_generic_to_shared(tab + warp * COLS);
  const float2* sp = strided + lane * 81;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < COLS; c += 4) {
      float a = 1.f, b = 1.f, cc = 1.f, d = 1.f;
      if (MODE == 0 || MODE == 3) {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(cc), "=f"(d) : "r"(tab_sa + 4 * c));
      }
      if (MODE == 1 || MODE == 4) {
        tmem_ld4(taddr + c, a, b, cc, d);
        tmem_wait_ld();
      }
      float2 s0 = make_float2(1.f, 1.f), s1 = s0;
      if (MODE >= 2) {
        s0 = sp[(c / 4) * 2];
        s1 = sp[(c / 4) * 2 + 1];
      }
      acc0 = fmaf(a, s0.x, acc0);
      acc1 = fmaf(b, s0.y, acc1);
      acc2 = fmaf(cc, s1.x, acc2);
      acc3 = fmaf(d, s1.y, acc3);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + acc2 + acc3;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512));
}

template <int MODE>
int run(const char* name, float* out, long long* cyc) {
  const size_t smem = WARPS * COLS * 4 + 32 * 81 * 8 + 64;
  CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<MODE><<<148, WARPS * 32, smem>>>(out, cyc);
  CK(cudaDeviceSynchronize());
  k<MODE><<<148, WARPS * 32, smem>>>(out, cyc);
  CK(cudaDeviceSynchronize());
  long long h[148];
  CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  s /= 148;
  const double groups = (double)ITERS * (COLS / 4);  // (const x4 [+ 2 x LDS.64]) groups per warp
  printf("%-38s cycles/group/warp %.2f   SM cycles per group over %d warps: %.2f\n", name, s / groups, WARPS, s / groups / WARPS);
  return 0;
}

int main() {
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, 148 * WARPS * 32 * 4));
  CK(cudaMalloc(&cyc, 148 * 8));
  if (run<0>("LDS.128 broadcast", out, cyc)) return 1;
  if (run<1>("tcgen05.ld.32x32b.x4 + wait", out, cyc)) return 1;
  if (run<2>("2 x LDS.64 strided", out, cyc)) return 1;
  if (run<3>("2 x LDS.64 strided + LDS.128 bcast", out, cyc)) return 1;
  if (run<4>("2 x LDS.64 strided + tcgen05.ld.x4", out, cyc)) return 1;
  return 0;
}
