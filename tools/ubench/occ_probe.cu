// What limits CTAs per SM for a kernel that allocates tensor memory?  (diagnostic)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 2) k_plain(float* out) { out[threadIdx.x] = 1.f; }
__global__ void __launch_bounds__(384, 2) k_tmem(float* out) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  out[threadIdx.x] = (float)slot;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(256u) : "memory");
}
__global__ void __launch_bounds__(384, 2) k_mbar(float* out) {
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  out[threadIdx.x] = 2.f;
}
__global__ void __launch_bounds__(384, 2) k_namedbar(float* out) {
  asm volatile("bar.sync 1, 256;" ::: "memory");
  out[threadIdx.x] = 3.f;
}
int main() {
  int nb;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_plain, 288, 0); printf("plain: %d\n", nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tmem, 288, 0); printf("tmem alloc: %d\n", nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_mbar, 288, 0); printf("mbarrier+cluster fence: %d\n", nb);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_namedbar, 288, 0); printf("named barrier (thread 0-255 only call it): %d\n", nb);
  return 0;
}
