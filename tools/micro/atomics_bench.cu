// How expensive are same-address fp32 atomics at the end of a column-reduction kernel?
// nblocks blocks each add C values to out[0..C): time vs nblocks (contention per address = nblocks).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int C, float v) {
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, v);
}
__global__ void k_red(float* out, int C, float v) {  // red.global (no return) is what atomicAdd compiles to anyway
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(out + i, v * (float)(blockIdx.x & 3));
}
int main() {
  float* d;
  cudaMalloc(&d, 1 << 20);
  cudaMemset(d, 0, 1 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int Cs[] = {480, 1440, 2304};
  const int Bs[] = {37, 74, 148, 296, 592, 1184};
  for (int C : Cs)
    for (int nb : Bs) {
      for (int i = 0; i < 3; ++i) k<<<nb, 256>>>(d, C, 1.f);
      cudaEventRecord(e0);
      for (int i = 0; i < 20; ++i) k<<<nb, 256>>>(d, C, 1.f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("C=%d blocks=%d: %.2f us per launch (%.1f ns per same-address atomic)\n", C, nb, ms * 1e3 / 20,
             ms * 1e6 / 20 / nb);
    }
  return 0;
}
