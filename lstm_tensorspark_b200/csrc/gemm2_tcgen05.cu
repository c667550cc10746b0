// General tcgen05 GEMM for every hoisted (non-recurrent) matrix product of the training step:
//
//     C[M,N] (=|+=)  op(A)[M,K] · op(B)[K,N]  (+ bias[N]),   bf16 operands, fp32 accumulation in TMEM, bf16 or fp32 output
//
//   * K-major or MN-major operands, chosen per operand: the weight-gradient products dW = dG^T · X contract over the
//     T·B rows of two row-major activations, i.e. BOTH operands are MN-major ("NT" GEMM); they are read straight out of
//     the [T·B, 4H] / [T·B, D] tensors the recurrence kernels wrote - no transposed copies.  (Reference: the autodiff of
//     tf.matmul(ht, W_h) / tf.matmul(x, W_x), /root/reference/src/models/recurrent/lstm.py:89-90 via rnn.py:224.)
//     dX = dG · W_x reads W_x [4H, D] as an MN-major B operand (no W^T copy), the input projection is the plain TN case.
//   * cta_group::2: a thread-block cluster of 2 CTAs (one TPC) computes a 256 x 256 tile; each CTA TMA-loads its 128 rows of
//     A and HALF of B (128 columns), the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which reads both halves of B
//     out of both CTAs' shared memory: 32 KB of SMEM ingest per k-block and CTA instead of 48 KB (the limit of the
//     1-CTA kernel), six pipeline stages instead of four.
//   * fp32 output can ACCUMULATE into C (beta = 1): weight gradients go straight into the flat gradient buffer.
//   * Optional dataflow gate: tile rows [m0, m0 + 256) are only loaded once gate[m0 / gate_rows] >= gate_target
//     (the producer of A is a concurrently running persistent kernel - the layer wavefront).
//
//   warp 0 : TMA producer (both CTAs)     warp 1 : MMA issuer (leader CTA)     warp 2 : TMEM allocator
//   warps 4..7 : epilogue (tcgen05.ld 32x32b.x32 -> bias / accumulate -> 256-bit global stores)
// Two accumulator stages of 256 TMEM columns: the epilogue of tile i overlaps the mainloop of tile i+1.  Persistent,
// static round-robin tile schedule over min(#tiles, #SM pairs) clusters.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "tcgen05.cuh"
#include "tmap.h"

namespace {

constexpr int BM = 128;                     // rows per CTA (UMMA M per CTA)
constexpr int BK = 64;                      // 64 bf16 = 128 B = one swizzle atom
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

enum OutMode { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ACC = 2 };

template <int kCtas, int BN> struct Cfg2 {
  static constexpr int kBNCta = BN / kCtas;                      // B columns this CTA loads
  static constexpr int kABytes = BM * BK * 2;                    // 16 KB
  static constexpr int kBBytes = kBNCta * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 512 /*barriers*/;
};

TC_DEVICE uint32_t cluster_ctarank2() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TC_DEVICE void cluster_sync2() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TC_DEVICE uint32_t mapa2(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
// TMA load whose completion bytes may be accounted on the mbarrier of EITHER CTA of the pair (cluster address).
TC_DEVICE void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
TC_DEVICE void mma2_bf16_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
TC_DEVICE void mma1_bf16_ss(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
// all MMAs issued so far by this thread arrive (once complete) on the barrier at the same offset in BOTH CTAs of the pair
TC_DEVICE void mma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
TC_DEVICE void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_result)), "r"(ncols) : "memory");
}
TC_DEVICE void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
TC_DEVICE void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TC_DEVICE void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// MN-major operand tile: 64-element (128 B) atoms along M/N, 8 KB apart (one TMA box each); 8-row k groups 1024 B apart.
TC_DEVICE uint64_t desc_mnmajor_sw128(uint32_t smem_addr) { return tc::make_smem_desc(smem_addr, 8192, 1024, tc::LAYOUT_SW128); }

struct Gemm2Params {
  void* C;
  const float* bias;
  int M, N, K, ldc;
  // Dataflow gate (layer wavefront): the rows of A are written by a persistent LSTM kernel that is still running.  Rows
  // [t * gate_rows_per_step, +gate_rows_per_step) belong to time step t; a tile may be loaded once ALL gate_count arrival
  // counters gate[i * gate_stride] have reached gate_base + gate_per_step * t, with t the last (gate_use_last) or first time
  // step the tile touches.  done[(row / 128) * tiles_n + tile_n]++ publishes a finished 128-row x BN output block.
  const unsigned int* gate;
  int gate_count, gate_stride, gate_base, gate_per_step, gate_rows_per_step, gate_use_last;
  long long gate_spin_limit;     // clock64 ticks before giving up (sets *gate_err)
  int* gate_err;
  unsigned int* done;
  int pdl;                       // launched as a programmatic dependent of the previous kernel (see launch2); waits for it before exiting
  int reverse_m;                 // walk the M tiles from the last to the first (the backward recurrence runs backwards in time)
  // Folded operand (a batch-major [Bsz, T, F] array read as the time-major matrix [T * Bsz, F] without a transpose pass): the
  // tensor map describes the storage as [fold = Bsz rows][T * F columns]; logical row r lives at storage row r % fold, columns
  // (r / fold) * fold_cols + [0, F).  a_fold: the K-major A operand (rows = M);  b_fold: the MN-major B operand (rows = K).
  int a_fold, b_fold, fold_cols;
};

template <int kCtas, int BN, bool kAMN, bool kBMN, int kOut>
__global__ void __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Gemm2Params p) {
  using C = Cfg2<kCtas, BN>;
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
  // A kernel queued behind this one with the programmatic-dependent-launch attribute (a finished gradient bucket's fused
  // allreduce + update) may start once this grid is resident: it shares the SMs instead of waiting for the GEMM to drain.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t crank = kCtas == 2 ? cluster_ctarank2() : 0u;
  const bool leader = crank == 0;
  constexpr int TM = BM * kCtas;                              // tile rows per cluster
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + TM - 1) / TM;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x / kCtas, num_clusters = gridDim.x / kCtas;
  auto tile_m_of = [&](int tile) { const int tm = tile / tiles_n; return p.reverse_m ? tiles_m - 1 - tm : tm; };

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4 * kCtas); }
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    if (kCtas == 2) { tmem_alloc2(tmem_slot, C::kTmemCols); tmem_relinquish2(); }
    else { tc::tmem_alloc(tmem_slot, C::kTmemCols); tc::tmem_relinquish(); }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (kCtas == 2) cluster_sync2();               // the peer's barriers are initialised before anything arrives on them
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer (every CTA loads its own rows)
    const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    const uint32_t full0_leader = kCtas == 2 ? mapa2(full0, 0) : full0;   // completion bytes go to the leader's barrier
    const uint32_t sa0 = tc::smem_u32(smem_a), sb0 = tc::smem_u32(smem_b);
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    int gate_t_ok = -1;
    bool gate_dead = false;
    for (int tile = cluster_id; tile < num_tiles && ok; tile += num_clusters) {
      const int m0 = tile_m_of(tile) * TM + (int)crank * BM;
      const int n0 = (tile % tiles_n) * BN + (int)crank * C::kBNCta;
      if (p.gate != nullptr) {
        const int r0 = tile_m_of(tile) * TM, r1 = min(r0 + TM, p.M) - 1;
        const int t = (p.gate_use_last ? r1 : r0) / p.gate_rows_per_step;
        if (t != gate_t_ok && !gate_dead) {                    // tiles arrive in time order: poll once per time step and CTA
          const int target = p.gate_base + p.gate_per_step * t;
          const long long t0 = clock64();
          for (int g = lane; g < p.gate_count; g += 32) {
            const unsigned int* c = p.gate + (size_t)g * p.gate_stride;
            unsigned int v;
            while (true) {
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
              if ((int)v - target >= 0) break;
              if (clock64() - t0 > p.gate_spin_limit) { if (p.gate_err) atomicExch(p.gate_err, 1); ok = false; break; }
            }
          }
          ok = __all_sync(0xffffffffu, ok);
          asm volatile("fence.acq_rel.gpu;" ::: "memory");          // the producers' release -> our (TMA) reads
          asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy observation before async-proxy (TMA) reads
          if (!ok) { gate_dead = true; ok = true; }                 // producer died: error flag is set, finish the grid without gating
          gate_t_ok = t;
        }
      }
      // folded K-major A: the tile's 128 rows are 128 consecutive storage rows of ONE time step (fold % 128 == 0)
      const int a_col = (!kAMN && p.a_fold) ? (m0 / p.a_fold) * p.fold_cols : 0;
      const int a_row = (!kAMN && p.a_fold) ? m0 % p.a_fold : m0;
      for (int kb = 0, k0 = 0; kb < num_kb; ++kb, k0 += BK) {
        const uint32_t eb = empty0 + 8 * stage, fb = full0 + 8 * stage, fbl = full0_leader + 8 * stage;
        while (!tc::mbar_try_wait_u32(eb, phase ^ 1)) {}
        if (tc::elect_one()) {
          if (leader) tc::mbar_expect_tx_u32(fb, C::kStageBytes * kCtas);
          const uint32_t sa = sa0 + stage * C::kABytes, sb = sb0 + stage * C::kBBytes;
          // folded MN-major B: the k-block's 64 rows are 64 consecutive storage rows of one time step (fold % 64 == 0)
          const int b_col = (kBMN && p.b_fold) ? (k0 / p.b_fold) * p.fold_cols : 0;
          const int b_row = (kBMN && p.b_fold) ? k0 % p.b_fold : k0;
          if (kCtas == 2) {
            if (kAMN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &tmap_a, fbl, m0 + 64 * j, k0);
            } else {
              tma_load_2d_pair(sa, &tmap_a, fbl, k0 + a_col, a_row);
            }
            if (kBMN) {
#pragma unroll
              for (int j = 0; j < C::kBNCta / 64; ++j) tma_load_2d_pair(sb + j * 8192, &tmap_b, fbl, n0 + 64 * j + b_col, b_row);
            } else {
              tma_load_2d_pair(sb, &tmap_b, fbl, k0, n0);
            }
          } else {
            if (kAMN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tc::tma_load_2d_u32(sa + j * 8192, &tmap_a, fb, m0 + 64 * j, k0);
            } else {
              tc::tma_load_2d_u32(sa, &tmap_a, fb, k0 + a_col, a_row);
            }
            if (kBMN) {
#pragma unroll
              for (int j = 0; j < C::kBNCta / 64; ++j) tc::tma_load_2d_u32(sb + j * 8192, &tmap_b, fb, n0 + 64 * j + b_col, b_row);
            } else {
              tc::tma_load_2d_u32(sb, &tmap_b, fb, k0, n0);
            }
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================================================================== MMA issuer (leader CTA; one elected lane)
    constexpr uint32_t idesc = tc::make_idesc_bf16_f32(TM, BN, kAMN ? 1u : 0u, kBMN ? 1u : 0u);
    const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    const uint32_t tfull0 = tc::smem_u32(tmem_full), tempty0 = tc::smem_u32(tmem_empty);
    const uint64_t da0 = kAMN ? desc_mnmajor_sw128(tc::smem_u32(smem_a)) : tc::desc_kmajor_sw128(tc::smem_u32(smem_a));
    const uint64_t db0 = kBMN ? desc_mnmajor_sw128(tc::smem_u32(smem_b)) : tc::desc_kmajor_sw128(tc::smem_u32(smem_b));
    constexpr uint64_t kAStep = kAMN ? (2048 >> 4) : (32 >> 4);     // descriptor advance per UMMA_K = 16
    constexpr uint64_t kBStep = kBMN ? (2048 >> 4) : (32 >> 4);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      while (!tc::mbar_try_wait_u32(tempty0 + 8 * acc, acc_phase ^ 1)) {}
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        while (!tc::mbar_try_wait_u32(full0 + 8 * stage, phase)) {}
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint64_t da = da0 + (uint64_t)(stage * (C::kABytes >> 4));
          const uint64_t db = db0 + (uint64_t)(stage * (C::kBBytes >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
            if (kCtas == 2) mma2_bf16_ss(d_tmem, da + k * kAStep, db + k * kBStep, idesc, accum);
            else mma1_bf16_ss(d_tmem, da + k * kAStep, db + k * kBStep, idesc, accum);
          }
          if (kCtas == 2) {
            mma_commit_pair(empty0 + 8 * stage);             // both CTAs' smem slots are reusable once these MMAs retire
            if (kb == num_kb - 1) mma_commit_pair(tfull0 + 8 * acc);
          } else {
            tc::mma_commit_u32(empty0 + 8 * stage);
            if (kb == num_kb - 1) tc::mma_commit_u32(tfull0 + 8 * acc);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue (each CTA drains its own 128 rows)
    const int ew = warp - kEpiWarp0;                     // == warp % 4 : TMEM lane quarter this warp may read
    const uint32_t tempty_leader = kCtas == 2 ? mapa2(tc::smem_u32(tmem_empty), 0) : tc::smem_u32(tmem_empty);
    int acc = 0; uint32_t acc_phase = 0;
    const int N = p.N, M = p.M;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = tile_m_of(tile) * TM + (int)crank * BM, n0 = (tile % tiles_n) * BN;
      tc::mbar_wait(&tmem_full[acc], acc_phase);
      tc::fence_after_sync();
      const int row = m0 + ew * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + c, v);
        tc::tmem_ld_wait();
        const int col = n0 + c;
        if (row < M && col < N) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (col + i < N) {
                float4 b4 = *reinterpret_cast<const float4*>(p.bias + col + i);
                f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
              }
            }
          }
          if (kOut != OUT_BF16) {
            float* dst = reinterpret_cast<float*>(p.C) + (size_t)row * p.ldc + col;
            if (kOut == OUT_F32_ACC) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                if (col + i < N) {
                  const float4 o = *reinterpret_cast<const float4*>(dst + i);
                  f[i] += o.x; f[i + 1] += o.y; f[i + 2] += o.z; f[i + 3] += o.w;
                }
            }
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              if (col + i < N) *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + col;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              __nv_bfloat162 p2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&p2);
            }
            if ((p.ldc & 15) == 0 && col + 32 <= N) {
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
      if (lane == 0) {
        if (kCtas == 2) mbar_arrive_cluster(tempty_leader + 8 * acc);   // the leader's issuer may overwrite this accumulator stage
        else tc::mbar_arrive(&tmem_empty[acc]);
      }
      if (p.done != nullptr) {                     // publish this CTA's 128 x BN output block to the gated consumer kernel
        asm volatile("bar.sync 1, 128;" ::: "memory");           // all four epilogue warps have issued their stores
        if (ew == 0 && lane == 0)
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;"
                       ::"l"(p.done + ((size_t)tile_m_of(tile) * kCtas + crank) * tiles_n + (tile % tiles_n)) : "memory");
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (kCtas == 2) cluster_sync2();               // nobody leaves while the pair's MMAs may still read its shared memory
  if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (warp == 2) {
    if (kCtas == 2) tmem_dealloc2(tmem_base, C::kTmemCols); else tc::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int kCtas, int BN, bool kAMN, bool kBMN, int kOut>
int launch2(const void* A, const void* B, const Gemm2Params& p, int lda, int ldb, int dev, int max_ctas, cudaStream_t st) {
  using C = Cfg2<kCtas, BN>;
  CUtensorMap ta, tb;
  // K-major operand: [rows = M|N][cols = K];  MN-major operand: [rows = K][cols = M|N]
  if (kAMN) { if (int rc = ts::make_tmap_2d_bf16(&ta, A, (uint64_t)p.K, (uint64_t)p.M, (uint64_t)lda, 64, BK)) return rc; }
  else if (p.a_fold) { if (int rc = ts::make_tmap_2d_bf16(&ta, A, (uint64_t)p.a_fold, (uint64_t)(p.M / p.a_fold) * p.fold_cols, (uint64_t)lda, BK, BM)) return rc; }
  else      { if (int rc = ts::make_tmap_2d_bf16(&ta, A, (uint64_t)p.M, (uint64_t)p.K, (uint64_t)lda, BK, BM)) return rc; }
  if (kBMN && p.b_fold) { if (int rc = ts::make_tmap_2d_bf16(&tb, B, (uint64_t)p.b_fold, (uint64_t)(p.K / p.b_fold) * p.fold_cols, (uint64_t)ldb, 64, BK)) return rc; }
  else if (kBMN) { if (int rc = ts::make_tmap_2d_bf16(&tb, B, (uint64_t)p.K, (uint64_t)p.N, (uint64_t)ldb, 64, BK)) return rc; }
  else      { if (int rc = ts::make_tmap_2d_bf16(&tb, B, (uint64_t)p.N, (uint64_t)p.K, (uint64_t)ldb, BK, C::kBNCta)) return rc; }
  auto kern = gemm2_kernel<kCtas, BN, kAMN, kBMN, kOut>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    // same L1 / shared-memory split as the fused allreduce kernels: an SM only runs CTAs of two kernels at once when they
    // agree on it, and a gradient bucket's allreduce is launched (programmatic dependent) to run next to this GEMM
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  const int TM = BM * kCtas;
  const int tiles = ((p.M + TM - 1) / TM) * ((p.N + BN - 1) / BN);
  int sms = ts::sm_count(dev);
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  int clusters = sms / kCtas;
  if (clusters < 1) clusters = 1;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCtas); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::kSmemBytes; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (p.pdl) {
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  return (int)cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
}

template <int kCtas, int BN, bool kAMN, bool kBMN>
int launch_out(const void* A, const void* B, const Gemm2Params& p, int lda, int ldb, int out_mode, int dev, int max_ctas, cudaStream_t st) {
  switch (out_mode) {
    case OUT_BF16: return launch2<kCtas, BN, kAMN, kBMN, OUT_BF16>(A, B, p, lda, ldb, dev, max_ctas, st);
    case OUT_F32: return launch2<kCtas, BN, kAMN, kBMN, OUT_F32>(A, B, p, lda, ldb, dev, max_ctas, st);
    case OUT_F32_ACC: return launch2<kCtas, BN, kAMN, kBMN, OUT_F32_ACC>(A, B, p, lda, ldb, dev, max_ctas, st);
  }
  return -3;
}

template <int kCtas, int BN>
int launch_major(const void* A, const void* B, const Gemm2Params& p, int lda, int ldb, int a_mn, int b_mn, int out_mode, int dev,
                 int max_ctas, cudaStream_t st) {
  if (a_mn && b_mn) return launch_out<kCtas, BN, true, true>(A, B, p, lda, ldb, out_mode, dev, max_ctas, st);
  if (a_mn) return launch_out<kCtas, BN, true, false>(A, B, p, lda, ldb, out_mode, dev, max_ctas, st);
  if (b_mn) return launch_out<kCtas, BN, false, true>(A, B, p, lda, ldb, out_mode, dev, max_ctas, st);
  return launch_out<kCtas, BN, false, false>(A, B, p, lda, ldb, out_mode, dev, max_ctas, st);
}

}  // namespace

// A: K-major [M, K] (lda = row pitch) or MN-major [K, M];  B: K-major [N, K] or MN-major [K, N];  C [M, N] row pitch ldc.
// out_mode: 0 bf16, 1 fp32, 2 fp32 accumulate (C += A·B).  ctas: 1 or 2 (cta_group).  bn: 128 or 256.
// gate_cfg[7] = {count, stride (u32 words), base, per_step, rows_per_step, use_last, reverse_m} (see Gemm2Params); max_ctas > 0 caps the grid.
extern "C" int ts_gemm2(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int lda, int ldb, int ldc,
                        int a_mn, int b_mn, int out_mode, int ctas, int bn, int dev, int max_ctas, const unsigned int* gate,
                        const int* gate_cfg, unsigned int* done, int* gate_err, int pdl, int a_fold, int b_fold, int fold_cols,
                        cudaStream_t st) {
  if (K % 8 != 0 || lda % 8 != 0 || ldb % 8 != 0) { ts::set_last_error("gemm2: K and the operand pitches must be multiples of 8"); return -2; }
  if ((a_mn && M % 8 != 0) || (b_mn && N % 8 != 0) || N % 8 != 0) { ts::set_last_error("gemm2: M (MN-major A) / N must be multiples of 8"); return -2; }
  // folded operands (see Gemm2Params): no tile / k-block may straddle two time steps or run past a time step's columns
  if (a_fold && (a_mn || gate != nullptr || a_fold % 128 != 0 || M % a_fold != 0 || K % 64 != 0 || fold_cols < K)) {
    ts::set_last_error("gemm2: folded A needs a K-major ungated operand, fold % 128 == 0, M % fold == 0, K % 64 == 0"); return -2;
  }
  if (b_fold && (!b_mn || b_fold % 64 != 0 || K % b_fold != 0 || N % bn != 0 || fold_cols < N)) {
    ts::set_last_error("gemm2: folded B needs an MN-major operand, fold % 64 == 0, K % fold == 0, N % bn == 0"); return -2;
  }
  Gemm2Params p{};
  p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.ldc = ldc;
  p.a_fold = a_fold; p.b_fold = b_fold; p.fold_cols = fold_cols;
  p.gate = gate; p.gate_err = gate_err; p.done = done; p.pdl = pdl;
  if (gate != nullptr) {
    p.gate_count = gate_cfg[0]; p.gate_stride = gate_cfg[1]; p.gate_base = gate_cfg[2]; p.gate_per_step = gate_cfg[3];
    p.gate_rows_per_step = gate_cfg[4] > 0 ? gate_cfg[4] : 1; p.gate_use_last = gate_cfg[5]; p.reverse_m = gate_cfg[6];
  }
  p.gate_spin_limit = 6000000000LL;
  if (ctas == 2) {
    if (bn == 128) return launch_major<2, 128>(A, B, p, lda, ldb, a_mn, b_mn, out_mode, dev, max_ctas, st);
    return launch_major<2, 256>(A, B, p, lda, ldb, a_mn, b_mn, out_mode, dev, max_ctas, st);
  }
  if (bn == 128) return launch_major<1, 128>(A, B, p, lda, ldb, a_mn, b_mn, out_mode, dev, max_ctas, st);
  return launch_major<1, 256>(A, B, p, lda, ldb, a_mn, b_mn, out_mode, dev, max_ctas, st);
}
