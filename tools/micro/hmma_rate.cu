// Microbenchmark: legacy mma.sync m16n8k16 bf16 throughput per SM on sm_100a (how fast can the BiLSTM recurrence go
// if its matrix-vector product moves from FFMA to mma.sync?).  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, int ilp) {
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < ilp)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (float)(t1 - t0); }
  if (s == 123.456f) out[1] = s;
}
int main() {
  float* d; cudaMalloc(&d, 64);
  for (int warps : {4, 8, 16}) for (int ilp : {1, 2, 4, 8}) {
    const int iters = 2000;
    k<<<148, warps * 32>>>(d, iters, ilp);
    cudaDeviceSynchronize();
    k<<<148, warps * 32>>>(d, iters, ilp);
    float h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    double mmas = (double)iters * ilp * warps;          // per SM
    double fma_per_clk = mmas * 16 * 8 * 16 / h[0];
    printf("warps/SM %2d ilp %d: %.0f cycles, %.1f cycles per mma per SM, %.0f bf16 FMA/clk/SM\n", warps, ilp, h[0], h[0] / mmas, fma_per_clk);
  }
  return 0;
}
