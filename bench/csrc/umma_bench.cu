// Microbenchmark: issue rate of tcgen05.mma (kind::f16, SMEM x SMEM -> TMEM) and the cost of tcgen05.commit.
// Operands sit in shared memory (garbage values, 128 B-swizzled K-major tiles), nothing is loaded in the timed loop.
// The issuing warp runs converged and issues under elect.sync (uniform registers), exactly like the real kernels.
//   mode 0 : groups of 4 MMAs, no commit            -> tensor-pipe cost of the instruction shape
//   mode 1 : commit after every group (nobody waits) -> does tcgen05.commit throttle issue?
//   mode 2 : commit after every 2nd group
//   mode 3 : commit after every group AND wait for the commit issued 4 groups earlier (the kernels' 4-deep ring)
//   mode 4 : commit + wait immediately (issue -> complete -> mbarrier latency, serialized)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tcgen05.cuh"

namespace {

template <int kMode>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int M, int N, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  constexpr int kblocks = 4;
  uint8_t* a = smem;                       // kblocks x [128 rows x 128 B]
  uint8_t* b = smem + kblocks * 16384;     // kblocks x [256 rows x 128 B]
  for (int i = threadIdx.x; i < kblocks * (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) tc::mbar_init(&bars[i], 1);
    tc::mbar_init(&done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) { tc::tmem_alloc(&tmem_slot, 256); tc::tmem_relinquish(); }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t d = tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = tc::make_idesc_bf16_f32((uint32_t)M, (uint32_t)N);
    const uint64_t da0 = tc::desc_kmajor_sw128(tc::smem_u32(a));
    const uint64_t db0 = tc::desc_kmajor_sw128(tc::smem_u32(b));
    const uint32_t bar0 = tc::smem_u32(&bars[0]);
    uint32_t slot = 0, phase = 0;
    long long t0 = clock64();
    for (int grp = 0; grp < iters; ++grp) {
      if (kMode == 3 && grp >= 4) { while (!tc::mbar_try_wait_u32(bar0 + 8 * slot, phase ^ 1)) {} }
      if (tc::elect_one()) {
        const uint64_t da = da0 + (uint64_t)(slot * 1024), db = db0 + (uint64_t)(slot * 2048);
        tc::mma_bf16_ss_acc(d, da, db, idesc);
        tc::mma_bf16_ss_acc(d, da + 2, db + 2, idesc);
        tc::mma_bf16_ss_acc(d, da + 4, db + 4, idesc);
        tc::mma_bf16_ss_acc(d, da + 6, db + 6, idesc);
        if (kMode == 1 || kMode == 3 || kMode == 4 || (kMode == 2 && (grp & 1))) tc::mma_commit_u32(bar0 + 8 * slot);
      }
      __syncwarp();
      if (kMode == 4) { while (!tc::mbar_try_wait_u32(bar0 + 8 * slot, phase)) {} }
      if (++slot == 4) { slot = 0; phase ^= 1; }
    }
    if (tc::elect_one()) tc::mma_commit(&done);
    __syncwarp();
    while (!tc::mbar_try_wait(&done, 0)) {}
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = (long long)iters * 4; }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(d, 256);
}

}  // namespace

extern "C" int ts_umma_bench(int M, int N, int iters, int mode, long long* out, cudaStream_t st) {
  const int smem = 4 * (16384 + 32768) + 1024;
#define TS_LAUNCH(MODE)                                                                                               \
  {                                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(umma_bench_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e != cudaSuccess) return (int)e;                                                                              \
    umma_bench_kernel<MODE><<<1, 128, smem, st>>>(M, N, iters, out);                                                  \
  }
  switch (mode) {
    case 0: TS_LAUNCH(0); break;
    case 1: TS_LAUNCH(1); break;
    case 2: TS_LAUNCH(2); break;
    case 3: TS_LAUNCH(3); break;
    default: TS_LAUNCH(4); break;
  }
#undef TS_LAUNCH
  return (int)cudaGetLastError();
}
