// tmem_probe.cu -- how much operand bandwidth does TMEM add next to shared memory on sm_100a?
// Every warp reads its own 32-lane quarter of a TMEM table (tcgen05.ld.32x32b) and/or a lane-private
// shared-memory table (LDS.128), as the SSV kernel does for its emission scores.  Prints bytes/clk/SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS>
__global__ void __launch_bounds__(256) probe(int mode, int iters, uint32_t *out, unsigned long long *cycles)
{
  extern __shared__ __align__(128) uint32_t s_tab[];          // [32 rows][4 words][32 lanes] = 16 KB
  __shared__ uint32_t s_taddr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 128; i += blockDim.x) s_tab[i] = i * 2654435761u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_taddr)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = s_taddr + ((uint32_t)(32 * (warp & 3)) << 16);
  // fill this warp's quarter: COLS columns, 4 at a time
  for (int c = 0; c < COLS; c += 4) {
    const uint32_t v0 = c * 131 + lane, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(tbase + c), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  __syncthreads();

  uint32_t acc = 0;
  const uint32_t tab_lane = smem_u32(s_tab) + lane * 16;
  const long long t0 = clock64();
  uint32_t x = (uint32_t)(warp * 5 + blockIdx.x);
  if (mode == 0) {                                            // shared memory only: one LDS.128 per step
    for (int i = 0; i < iters; i++) {
      x = (x * 13 + 7) & 31;
      uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tab_lane + x * 512));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  } else if (mode == 1) {                                     // TMEM only: one tcgen05.ld.x4 per step (same bytes)
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int i = 0; i < iters; i++) {
      x = (x * 13 + 7) & 31;
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc += a0 ^ a1 ^ a2 ^ a3;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(tbase + ((x * 4) & (COLS - 1))));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    acc += a0 ^ a1 ^ a2 ^ a3;
  } else if (mode == 2) {                                     // both: LDS.128 + tcgen05.ld.x2 per step (1.5x the bytes)
    uint32_t a0 = 0, a1 = 0;
    for (int i = 0; i < iters; i++) {
      x = (x * 13 + 7) & 31;
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc += a0 ^ a1;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(tbase + ((x * 2) & (COLS - 1))));
      uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tab_lane + x * 512));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    acc += a0 ^ a1;
  } else {                                                    // TMEM x1 per step
    uint32_t a0 = 0;
    for (int i = 0; i < iters; i++) {
      x = (x * 13 + 7) & 31;
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc += a0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a0) : "r"(tbase + (x & (COLS - 1))));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    acc += a0;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_taddr), "n"(COLS));
}

int main()
{
  int dev = 0; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  uint32_t *d_out; unsigned long long *d_cyc;
  const int maxblocks = sms * 8;
  CK(cudaMalloc(&d_out, (size_t)maxblocks * 256 * 4)); CK(cudaMalloc(&d_cyc, (size_t)maxblocks * 8));
  CK(cudaFuncSetAttribute(probe<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  const int iters = 20000;
  const char *names[4] = {"LDS.128 only (512 B/warp/step)", "tcgen05.ld.x4 only (512 B/warp/step)", "LDS.128 + tcgen05.ld.x2 (768 B/warp/step)", "tcgen05.ld.x1 only (128 B/warp/step)"};
  const int bytes[4] = {512, 512, 768, 128};
  for (int per_sm = 1; per_sm <= 4; per_sm *= 2)
    for (int mode = 0; mode < 4; mode++) {
      const int blocks = sms * per_sm;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      probe<128><<<blocks, 256, 16384>>>(mode, 200, d_out, d_cyc);       // warm-up
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      probe<128><<<blocks, 256, 16384>>>(mode, iters, d_out, d_cyc);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      unsigned long long cyc0; CK(cudaMemcpy(&cyc0, d_cyc, 8, cudaMemcpyDeviceToHost));
      const double bytes_per_sm = (double)per_sm * 8 * iters * bytes[mode];
      printf("%d CTA/SM  %-46s %8.3f ms  %7.1f B/clk/SM (CTA0 clock64: %llu cycles)\n", per_sm, names[mode], ms, bytes_per_sm / (double)cyc0, cyc0);
    }
  return 0;
}
