// Persistent, warp-specialised tcgen05 GEMM:  C[M,N] = A[M,K] · B[N,K]^T (+ bias[N]),  bf16 in, fp32 accumulate.
// Used for the hoisted LSTM input projection  Gx = X[T·B, D] · Wx[4H, D]^T  (K2 in SURVEY §2.5: the reference's
// per-gate tf.matmul(input_data, W_x), /root/reference/src/models/recurrent/lstm.py:90, batched over all T and all
// four gates) and the backward dX = dG · Wx.
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2-D, 128 B swizzle, kStages-deep mbarrier ring)
//   warp 1      : MMA issuer     (one elected thread: tcgen05.mma cta_group::1 kind::f16, M=128, N=BLOCK_N, K=16)
//   warp 2      : TMEM allocator (2 accumulator stages of BLOCK_N fp32 columns: epilogue(i) overlaps mainloop(i+1))
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b -> +bias -> bf16/fp32 -> 16 B global stores)
// Grid = min(#tiles, #SMs); static round-robin tile schedule.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "tcgen05.cuh"
#include "tmap.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // 64 bf16 = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

template <int BLOCK_N> struct Cfg {
  static constexpr int kStages = BLOCK_N == 256 ? 4 : 6;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;     // 256 or 512 (power of two)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BLOCK_N, bool kOutF32>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    void* __restrict__ Cout, const float* __restrict__ bias, int M, int N, int K) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, C::kTmemCols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // whole warp in uniform control flow, one elected lane issues (keeps addresses in uniform registers)
    const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    const uint32_t sa0 = tc::smem_u32(smem_a), sb0 = tc::smem_u32(smem_b);
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
      for (int kb = 0, k0 = 0; kb < num_kb; ++kb, k0 += BLOCK_K) {
        const uint32_t eb = empty0 + 8 * stage, fb = full0 + 8 * stage;
        while (!tc::mbar_try_wait_u32(eb, phase ^ 1)) {}
        if (tc::elect_one()) {
          tc::mbar_expect_tx_u32(fb, C::kStageBytes);
          tc::tma_load_2d_u32(sa0 + stage * C::kABytes, &tmap_a, fb, k0, m0);
          tc::tma_load_2d_u32(sb0 + stage * C::kBBytes, &tmap_b, fb, k0, n0);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp in uniform control flow; one elected lane issues every MMA; descriptors advance by adds
    constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BLOCK_M, BLOCK_N);
    const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    const uint32_t tfull0 = tc::smem_u32(tmem_full), tempty0 = tc::smem_u32(tmem_empty);
    const uint64_t da0 = tc::desc_kmajor_sw128(tc::smem_u32(smem_a));
    const uint64_t db0 = tc::desc_kmajor_sw128(tc::smem_u32(smem_b));
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      while (!tc::mbar_try_wait_u32(tempty0 + 8 * acc, acc_phase ^ 1)) {}
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        while (!tc::mbar_try_wait_u32(full0 + 8 * stage, phase)) {}
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint64_t da = da0 + (uint64_t)(stage * (C::kABytes >> 4));
          const uint64_t db = db0 + (uint64_t)(stage * (C::kBBytes >> 4));
          if (kb == 0) tc::mma_bf16_ss_first(d_tmem, da, db, idesc); else tc::mma_bf16_ss_acc(d_tmem, da, db, idesc);
          tc::mma_bf16_ss_acc(d_tmem, da + 2, db + 2, idesc);
          tc::mma_bf16_ss_acc(d_tmem, da + 4, db + 4, idesc);
          tc::mma_bf16_ss_acc(d_tmem, da + 6, db + 6, idesc);
          tc::mma_commit_u32(empty0 + 8 * stage);        // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) tc::mma_commit_u32(tfull0 + 8 * acc);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    const int ew = warp - kEpiWarp0;                     // == warp % 4 : TMEM lane quarter this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
      tc::mbar_wait(&tmem_full[acc], acc_phase);
      tc::fence_after_sync();
      const int row = m0 + ew * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N + c, v);
        tc::tmem_ld_wait();
        const int col = n0 + c;
        if (row < M && col < N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (col + i < N) {
                float4 b4 = *reinterpret_cast<const float4*>(bias + col + i);
                f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
              }
            }
          }
          if (kOutF32) {
            float* dst = reinterpret_cast<float*>(Cout) + (size_t)row * N + col;
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (col + i < N) *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)row * N + col;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __nv_bfloat162 p2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&p2);
            }
            if ((N & 15) == 0 && col + 32 <= N) {
              // whole 32 B sectors, 256-bit stores: half the L2 write requests of 16 B stores
#pragma unroll
              for (int i = 0; i < 2; ++i)
                asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 16 * i), "r"(pk[8 * i]), "r"(pk[8 * i + 1]),
                             "r"(pk[8 * i + 2]), "r"(pk[8 * i + 3]), "r"(pk[8 * i + 4]), "r"(pk[8 * i + 5]), "r"(pk[8 * i + 6]), "r"(pk[8 * i + 7]) : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 8)
                if (col + i < N) *reinterpret_cast<uint4*>(dst + i) = make_uint4(pk[i / 2], pk[i / 2 + 1], pk[i / 2 + 2], pk[i / 2 + 3]);
            }
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int BLOCK_N, bool kOutF32>
int launch(const void* A, const void* B, void* Cc, const float* bias, int M, int N, int K, int dev, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  CUtensorMap ta, tb;
  if (int rc = ts::make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, BLOCK_K, BLOCK_M)) return rc;
  if (int rc = ts::make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)K, BLOCK_K, BLOCK_N)) return rc;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, kOutF32>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  int sms = ts::sm_count(dev);
  int grid = tiles < sms ? tiles : sms;
  kern<<<grid, kThreads, C::kSmemBytes, st>>>(ta, tb, Cc, bias, M, N, K);
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" int ts_gemm_bf16_tn(const void* A, const void* B, void* C, const float* bias, int M, int N, int K,
                               int out_fp32, int variant, int dev, cudaStream_t st) {
  if (K % 8 != 0 || N % 8 != 0) { ts::set_last_error("gemm_bf16_tn: K and N must be multiples of 8"); return -2; }
  if (variant == 1) {
    return out_fp32 ? launch<256, true>(A, B, C, bias, M, N, K, dev, st) : launch<256, false>(A, B, C, bias, M, N, K, dev, st);
  }
  return out_fp32 ? launch<128, true>(A, B, C, bias, M, N, K, dev, st) : launch<128, false>(A, B, C, bias, M, N, K, dev, st);
}
