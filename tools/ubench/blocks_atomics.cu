// Microbenchmarks behind the block-count / atomic choices of the column kernels (diagnostic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/blocks_atomics tools/ubench/blocks_atomics.cu
// Prints us per launch (back-to-back launches, CUDA events) for:
//   empty      : grid of N blocks x 256 threads doing nothing           -> block dispatch rate
//   store      : every thread stores one float4                          -> baseline with memory traffic
//   atom64/32  : every block ends with K atomics (fp64 / fp32) on 512 addresses -> atomic throughput
//   redv4      : the same bytes as fp32 atomics but red.global.add.v4.f32
//   ticket     : __threadfence + one same-address atomicAdd per block (last-block ticket)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_empty(int) {}
__global__ void k_store(float4* out) { out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f); }
__global__ void k_atom64(double* acc, int k, int spread) {
  if ((int)threadIdx.x < k) atomicAdd(acc + ((blockIdx.x % spread) * 512 + threadIdx.x % 512), 1.0);
}
__global__ void k_atom32(float* acc, int k, int spread) {
  if ((int)threadIdx.x < k) atomicAdd(acc + ((blockIdx.x % spread) * 512 + threadIdx.x % 512), 1.0f);
}
__global__ void k_redv4(float* acc, int k, int spread) {
  if ((int)threadIdx.x < k / 4) {
    float* p = acc + ((blockIdx.x % spread) * 512 + (threadIdx.x * 4) % 512);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.f), "f"(1.f), "f"(1.f), "f"(1.f) : "memory");
  }
}
__global__ void k_ticket(unsigned int* counter, float4* out) {
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f);
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    last = t == gridDim.x - 1;
    if (last) *counter = 0;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) out[0].x = 5.f;
}

template <class F>
float time_us(F f, int iters = 200) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 20; ++i) f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms * 1e3f / iters;
}

int main() {
  float4* out; double* acc64; float* acc32; unsigned int* ctr;
  cudaMalloc(&out, (size_t)148 * 64 * 256 * 16);
  cudaMalloc(&acc64, 64 * 512 * 8); cudaMemset(acc64, 0, 64 * 512 * 8);
  cudaMalloc(&acc32, 64 * 512 * 4); cudaMemset(acc32, 0, 64 * 512 * 4);
  cudaMalloc(&ctr, 256); cudaMemset(ctr, 0, 256);
  printf("%8s %8s %8s %8s | K=128: %8s %8s %8s %8s %8s | K=512 spread8: %8s %8s %8s\n", "blocks", "empty", "store", "ticket", "atom64", "a64 sp8", "atom32", "a32 sp8",
         "redv4", "atom64", "atom32", "redv4");
  for (int mult : {1, 2, 4, 8, 16}) {
    const int n = 148 * mult;
    printf("%8d %8.2f %8.2f %8.2f |        %8.2f %8.2f %8.2f %8.2f %8.2f |               %8.2f %8.2f %8.2f\n", n,
           time_us([&] { k_empty<<<n, 256>>>(0); }), time_us([&] { k_store<<<n, 256>>>(out); }),
           time_us([&] { k_ticket<<<n, 256>>>(ctr, out); }),
           time_us([&] { k_atom64<<<n, 256>>>(acc64, 128, 1); }), time_us([&] { k_atom64<<<n, 256>>>(acc64, 128, 8); }),
           time_us([&] { k_atom32<<<n, 256>>>(acc32, 128, 1); }), time_us([&] { k_atom32<<<n, 256>>>(acc32, 128, 8); }),
           time_us([&] { k_redv4<<<n, 256>>>(acc32, 128, 8); }),
           time_us([&] { k_atom64<<<n, 512>>>(acc64, 512, 8); }), time_us([&] { k_atom32<<<n, 512>>>(acc32, 512, 8); }),
           time_us([&] { k_redv4<<<n, 512>>>(acc32, 512, 8); }));
  }
  return 0;
}
