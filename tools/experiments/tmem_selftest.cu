// Stand-alone check of the tensor-memory-as-scratchpad pattern used by k_solve_tmem (no MMA involved):
// one CTA of 352 threads allocates all 512 TMEM columns; warp w owns lane quarter 32*(w%4) and the column
// range 160*(w/4); every thread writes 20 x 8 words into its own lane with tcgen05.st.32x32b.x8 and reads
// them back with tcgen05.ld.32x32b.x8.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_selftest tmem_selftest.cu && ./tmem_selftest
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 352, kStages = 20;

__global__ void __launch_bounds__(kThreads, 1) k_selftest(int* errors) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&s_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t taddr = s_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(8 * kStages * (warp >> 2));
  for (int k = 0; k < kStages; ++k) {
    uint32_t w[8];
    for (int c = 0; c < 8; ++c) w[c] = (threadIdx.x << 16) | (k << 8) | c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 :: "r"(taddr + 8u * k), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  int bad = 0;
  for (int k = kStages - 1; k >= 0; --k) {
    uint32_t w[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(taddr + 8u * k) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int c = 0; c < 8; ++c) bad += w[c] != ((threadIdx.x << 16) | (k << 8) | c);
  }
  if (bad) atomicAdd(errors, bad);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" :: "r"(s_base) : "memory");
}

int main() {
  int* d;
  cudaMalloc(&d, sizeof(int));
  cudaMemset(d, 0, sizeof(int));
  k_selftest<<<148, kThreads>>>(d);
  int h = -1;
  cudaError_t e = cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost);
  printf("%s, errors = %d\n", cudaGetErrorString(e), h);
  return (e == cudaSuccess && h == 0) ? 0 : 1;
}
