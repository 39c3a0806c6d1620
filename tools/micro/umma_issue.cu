// Microbenchmark: what does ONE thread pay to issue tcgen05.mma / tcgen05.commit / mbarrier waits?  (sm_100a)
// The GEMM mainloop of gemm2_kernel runs at ~670 cycles per 64-wide k-block whatever the tile width and even without any
// TMA load (tools/mainloop_probe.py), against 512 (N = 256) or 256 (N = 128) cycles of tensor-pipe time: the issuing
// thread's instruction stream is the pace.  This kernel times that stream in isolation:
//   per iteration: n_mma x tcgen05.mma (M 128, N, K 16) [+ commit on a private mbarrier] [+ try_wait on a completed barrier]
// and reports issue cycles per iteration (clock64 around the loop) and cycles until everything completed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I styletts-zs_b200/csrc -o tools/micro/umma_issue tools/micro/umma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace stz;

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int n_mma, int do_commit, int do_wait, int uniform) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[8], done_bar, ready_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (uint32_t o = threadIdx.x * 16; o < 16384 + 32768; o += 128 * 16) st_shared_v4(base + o, 0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(&done_bar, 1);
    mbar_init(&ready_bar, 1);
    fence_barrier_init();
    mbar_arrive(&ready_bar);            // phase 0 of ready_bar is complete: try_wait(parity 0) succeeds immediately
  }
  if (threadIdx.x < 32) tmem_alloc<256>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (uniform && threadIdx.x < 32) {
    // warp-uniform form: the whole warp runs the loop (addresses / descriptors provably uniform -> uniform registers), one
    // elected lane issues — no ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall around every tcgen05 instruction
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 16384);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (do_wait) mbar_wait(&ready_bar, 0);
      if (do_wait > 1) tc_fence_after();
      if (elect_one()) {
        for (int m = 0; m < n_mma; ++m) umma_bf16(tmem, da + 2 * (m & 3), db + 2 * (m & 3), idesc, 1u);
        if (do_commit) umma_commit(&bars[it & 7]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(&done_bar);
    __syncwarp();
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (!uniform && threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 16384);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (do_wait) mbar_wait(&ready_bar, 0);
      if (do_wait > 1) tc_fence_after();
      for (int m = 0; m < n_mma; ++m) umma_bf16(tmem, da + 2 * (m & 3), db + 2 * (m & 3), idesc, 1u);
      if (do_commit) umma_commit(&bars[it & 7]);
    }
    const long long t1 = clock64();
    umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<256>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  const int smem = 16384 + 32768 + 1024, iters = 2000;
  cudaFuncSetAttribute(k<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int uniform = 0; uniform < 2; ++uniform)
  for (int N : {256, 128})
    for (int n_mma : {1, 4, 8})
      for (int mode = 0; mode < 4; ++mode) {
        const int commit = mode >= 1, wait = mode == 2 ? 1 : (mode == 3 ? 2 : 0);
        for (int rep = 0; rep < 2; ++rep) {
          if (N == 256) k<256><<<148, 128, smem>>>(d, iters, n_mma, commit, wait, uniform);
          else k<128><<<148, 128, smem>>>(d, iters, n_mma, commit, wait, uniform);
          cudaDeviceSynchronize();
        }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s N %3d  %d MMA/iter  %-28s issue %6.0f cycles/iter   complete %6.0f cycles/iter   (tensor-pipe floor %4d)\n",
               uniform ? "warp-uniform + elect" : "single thread       ", N, n_mma,
               mode == 0 ? "no commit" : mode == 1 ? "+ commit" : mode == 2 ? "+ commit + try_wait" : "+ commit + try_wait + fence",
               (double)h[0] / iters, (double)h[1] / iters, n_mma * N / 2);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
