// K-LSTM / K-LSTM-BWD: persistent tcgen05 kernels that run a whole layer's recurrence in ONE launch.
//
// What they replace (reference, per layer per time step): 4x tf.matmul(ht, W_h) + bias adds + 3 sigmoid + 2 tanh +
// the c/h update, each its own TF op (/root/reference/src/models/recurrent/lstm.py:88-109; K1,K3-K8 in SURVEY §2.5),
// and the mirrored autodiff backward (K13).  The reference only ever takes ONE step; these kernels deliver the
// multi-step unroll its fit_next API was built for (lstm.py:128-136).
//
// Design (B200-first):
//   * Every CTA keeps a [64 x H] bf16 slice of the recurrent weights RESIDENT in shared memory for all T steps
//     (loaded once by TMA; W_h is 8 MB at H=1024 = 128 CTAs x 64 KB..128 KB).  Rows are gate-interleaved (n = 4j+g) so
//     a CTA that owns a row slice owns complete (i,f,g,o) quadruples: the gate epilogue needs no cross-CTA traffic.
//   * Per step a CTA streams a 128-row batch tile of h_{t-1} (forward) / dG_{t+1} (backward) through a TMA->mbarrier
//     ring; one elected thread issues tcgen05.mma (M=128, N=64, K=16, bf16 -> fp32 accumulators in TMEM); four epilogue
//     warps read the accumulator with tcgen05.ld and do the whole cell in registers.
//   * The streamed operand is the L2-bandwidth bottleneck (every CTA of a batch tile needs all of it).  Forward:
//     CTAs that share a batch tile form a thread-block CLUSTER and each k-block is fetched from L2 once and
//     TMA-MULTICAST into all members.  Backward: the contraction runs over 4H, so a cluster of 4 CTAs splits K
//     (one gate-column quarter each, 4x less operand traffic than a single-CTA K=4H loop) and the four partial
//     [128 x 64] tiles are reduce-scattered through DISTRIBUTED SHARED MEMORY (st.shared::cluster + remote mbarrier
//     arrive); each member then owns 16 hidden columns of dh for the gate-gradient epilogue.
//   * Steps are separated by a grid-wide dataflow barrier in global memory (one monotonically increasing counter per
//     batch tile, red.release.gpu / ld.acquire.gpu, generic->async proxy fences because the consumer is TMA).
//     All CTAs are co-resident (grid <= #SMs, 1 CTA/SM, checked with cudaOccupancyMaxActiveClusters); every spin is
//     bounded and raises an error flag instead of hanging the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "tcgen05.cuh"
#include "tmap.h"
#include "ts_common.cuh"

namespace {

constexpr int BM = 128;           // batch rows per CTA (UMMA M)
constexpr int BN = 64;            // accumulator columns per CTA (UMMA N)
constexpr int BK = 64;            // K per pipeline stage (one 128 B swizzle atom of bf16)
constexpr int UK = 16;            // UMMA K
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kWBlockBytes = BN * BK * 2;     // 8 KB per 64-wide K block
constexpr int kXchgBytes = 3 * BM * 16 * 4;   // backward: 3 foreign partial chunks [128 x 16] fp32
constexpr long long kSpinLimit = 6000000000LL;   // ~3 s of SM clocks: a bug surfaces as an error, not a hung GPU

struct SeqSmem {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t w_full;
  uint64_t tmem_full;
  uint64_t xchg_full;
  uint32_t tmem_slot;
  int abort_flag;
  float bias[64];
};

TC_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
TC_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TC_DEVICE uint32_t mapa(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
TC_DEVICE void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
TC_DEVICE void mbar_arrive_remote(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
TC_DEVICE bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
TC_DEVICE void tma_load_3d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(tc::smem_u32(smem_dst)), "l"((uint64_t)map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
TC_DEVICE void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}

template <bool kClusterScope>
TC_DEVICE bool wait_bar(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
  auto probe = [&]() { return kClusterScope ? mbar_try_wait_cluster(bar, parity) : tc::mbar_try_wait(bar, parity); };
  if (probe()) return true;
  long long t0 = clock64();
  int n = 0;
  while (!probe()) {
    if ((++n & 255) == 0) {
      if (*abort_flag) return false;
      if (clock64() - t0 > kSpinLimit) { *abort_flag = 1; return false; }
    }
  }
  return true;
}

TC_DEVICE bool wait_counter(const unsigned int* ctr, unsigned int target, volatile int* abort_flag) {
  long long t0 = clock64();
  int n = 0;
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if ((int)(v - target) >= 0) return true;
    if ((++n & 63) == 0) {
      if (*abort_flag) return false;
      if (clock64() - t0 > kSpinLimit) { *abort_flag = 1; return false; }
    }
  }
}
TC_DEVICE void signal_counter(unsigned int* ctr) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
TC_DEVICE unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

TC_DEVICE uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
TC_DEVICE void stg16(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TC_DEVICE float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
TC_DEVICE float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
TC_DEVICE uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

struct SeqParams {
  // forward
  const __nv_bfloat16* gx;     // [T,B,4H]
  const float* bias;           // [4H]
  __nv_bfloat16* h_seq;        // [T+1,B,H]   (row 0 = h0)
  float* c_seq;                // [T+1,B,H]   (row 0 = c0)
  __nv_bfloat16* act;          // [T,B,4H]
  // backward
  const __nv_bfloat16* dh_seq; // [T,B,H]
  __nv_bfloat16* dpre;         // [T,B,4H]
  float* dh0;                  // [B,H] in: dL/dh_T extra, out: dL/dh_0
  float* dc0;                  // [B,H] in: dL/dc_T, out: dL/dc_0
  unsigned int* sync;          // [tiles_m] step counters; [63] error flag
  unsigned long long* dbg;     // optional [steps][4] timestamps of CTA 0 (ns)
  int T, B, H;
  int tiles_n;                 // CTAs per batch tile
};

// Forward : CTA (mb, nb)      -> gate columns [64 nb, +64) = hidden [16 nb, +16);   K = H;  cluster = kCluster CTAs along nb (multicast)
// Backward: CTA (mb, nb2, ks) -> partial dh columns [64 nb2, +64) over gate-column quarter ks (K = H); cluster = 4 (ks);
//           after the DSMEM reduce-scatter member ks owns hidden [64 nb2 + 16 ks, +16).
template <bool kBwd, int kCluster>
__global__ void __launch_bounds__(kThreads, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                const SeqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int num_kb = p.H / BK;
  uint8_t* smem_w = smem;                                    // resident weight slice: num_kb blocks of [64 x 64]
  uint8_t* smem_a = smem + (size_t)num_kb * kWBlockBytes;    // kStages x 16 KB
  uint8_t* smem_x = smem_a + kStages * kABytes;              // backward: DSMEM exchange buffer
  SeqSmem* ss = reinterpret_cast<SeqSmem*>(smem_x + (kBwd ? kXchgBytes : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x / p.tiles_n;
  const int in_mb = blockIdx.x % p.tiles_n;
  const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0;
  // forward: nb = in_mb.  backward: cluster of 4 = the 4 K-quarters of one (mb, nb2)
  const int nb = kBwd ? in_mb / 4 : in_mb;
  const int ks = kBwd ? (int)crank : 0;
  unsigned int* counter = p.sync + mb;
  volatile int* abort_flag = &ss->abort_flag;
  const int steps = kBwd ? p.T + 1 : p.T;       // backward runs one extra GEMM to produce dh_0
  const bool mcast = !kBwd && kCluster > 1;
  const uint16_t cmask = (uint16_t)((1u << kCluster) - 1);

  if (threadIdx.x == 0) {
    ss->abort_flag = 0;
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&ss->full[s], 1); tc::mbar_init(&ss->empty[s], mcast ? kCluster : 1); }
    tc::mbar_init(&ss->w_full, 1);
    tc::mbar_init(&ss->tmem_full, 1);
    tc::mbar_init(&ss->xchg_full, 3);
    tc::fence_barrier_init();
  }
  if (!kBwd && threadIdx.x >= 64 && threadIdx.x < 128) ss->bias[threadIdx.x - 64] = p.bias[nb * BN + threadIdx.x - 64];
  if (warp == 2) { tc::tmem_alloc(&ss->tmem_slot, 64); tc::tmem_relinquish(); }
  tc::fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();        // peers' mbarriers are initialised before anyone multicasts / arrives remotely
  tc::fence_after_sync();
  const uint32_t tmem_d = ss->tmem_slot;

  if (warp == 0) {
    // ======================================================================== TMA producer
    if (lane == 0) {
      tc::mbar_expect_tx(&ss->w_full, (uint32_t)(num_kb * kWBlockBytes));
      // forward: rows = gate columns [64 nb, +64) of W_h [4H, H].  backward: rows = hidden columns [64 nb, +64) of
      // W_h^T [H, 4H], K offset = quarter ks.
      for (int kb = 0; kb < num_kb; ++kb)
        tc::tma_load_2d(smem_w + (size_t)kb * kWBlockBytes, &tmap_w, &ss->w_full, (kBwd ? ks * p.H : 0) + kb * BK, nb * BN);
      int stage = 0; uint32_t phase = 0;
      bool ok = true;
      for (int s = kBwd ? 1 : 0; s < steps && ok; ++s) {
        // forward step s consumes h_seq[s] (rows written by step s-1); backward iteration s consumes dG[T-s]
        if (s > 0) ok = wait_counter(counter, (unsigned)s * p.tiles_n, abort_flag);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (p.dbg && blockIdx.x == 0) p.dbg[4 * s + 0] = gtime();
        const int tsl = kBwd ? p.T - s : s;
        for (int kb = 0; kb < num_kb && ok; ++kb) {
          ok = wait_bar<false>(&ss->empty[stage], phase ^ 1, abort_flag);
          if (!ok) break;
          tc::mbar_expect_tx(&ss->full[stage], kABytes);
          const int kcol = (kBwd ? ks * p.H : 0) + kb * BK;
          if (!mcast) {
            tc::tma_load_3d(smem_a + stage * kABytes, &tmap_a, &ss->full[stage], kcol, mb * BM, tsl);
          } else if ((kb % kCluster) == (int)crank) {
            tma_load_3d_mc(smem_a + stage * kABytes, &tmap_a, &ss->full[stage], kcol, mb * BM, tsl, cmask);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM, BN);
      bool ok = wait_bar<false>(&ss->w_full, 0, abort_flag);
      int stage = 0; uint32_t phase = 0;
      for (int s = kBwd ? 1 : 0; s < steps && ok; ++s) {
        for (int kb = 0; kb < num_kb; ++kb) {
          ok = wait_bar<false>(&ss->full[stage], phase, abort_flag);
          if (!ok) break;
          tc::fence_after_sync();
          const uint64_t da = tc::desc_kmajor_sw128(tc::smem_u32(smem_a + stage * kABytes));
          const uint64_t db = tc::desc_kmajor_sw128(tc::smem_u32(smem_w + (size_t)kb * kWBlockBytes));
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            tc::mma_bf16_ss(tmem_d, tc::desc_advance(da, k * UK * 2), tc::desc_advance(db, k * UK * 2), idesc, (kb | k) != 0);
          if (mcast) mma_commit_mc(&ss->empty[stage], cmask); else tc::mma_commit(&ss->empty[stage]);
          if (kb == num_kb - 1) tc::mma_commit(&ss->tmem_full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ======================================================================== epilogue: one thread = one batch row
    const int ew = warp - kEpiWarp0;
    const int rloc = ew * 32 + lane;
    const int row = mb * BM + rloc;
    const bool valid = row < p.B;
    const int H = p.H, B = p.B;
    const uint32_t taddr = tmem_d + ((uint32_t)(ew * 32) << 16);
    uint32_t tphase = 0;
    bool ok = true;

    if (!kBwd) {
      const int j0 = nb * 16;
      for (int t = 0; t < p.T && ok; ++t) {
        // operands that do not depend on the GEMM: issue their loads before waiting on the accumulator
        uint4 gxv[8];
        float4 cv[4];
        if (valid) {
          const __nv_bfloat16* gp = p.gx + ((size_t)t * B + row) * (4 * H) + nb * 64;
#pragma unroll
          for (int i = 0; i < 8; ++i) gxv[i] = ldg_nc16(gp + 8 * i);
          const float* cp = p.c_seq + ((size_t)t * B + row) * H + j0;
#pragma unroll
          for (int i = 0; i < 4; ++i) cv[i] = *reinterpret_cast<const float4*>(cp + 4 * i);
        }
        ok = wait_bar<false>(&ss->tmem_full, tphase, abort_flag);
        tphase ^= 1;
        if (!ok) break;
        tc::fence_after_sync();
        if (p.dbg && blockIdx.x == 0 && rloc == 0) p.dbg[4 * t + 1] = gtime();
        uint32_t v0[32], v1[32];
        tc::tmem_ld32(taddr, v0);
        tc::tmem_ld32(taddr + 32, v1);
        tc::tmem_ld_wait();
        tc::fence_before_sync();
        uint32_t hpk[8], apk[32];
        float cn[16];
        if (valid) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const uint32_t* vv = jj < 8 ? v0 : v1;
            const int q = (jj & 7) * 4;
            const uint4 g4 = gxv[jj >> 1];
            const uint32_t ga = (jj & 1) ? g4.z : g4.x, gb = (jj & 1) ? g4.w : g4.y;
            float pi = __uint_as_float(vv[q + 0]) + bf_lo(ga) + ss->bias[4 * jj + 0];
            float pf = __uint_as_float(vv[q + 1]) + bf_hi(ga) + ss->bias[4 * jj + 1];
            float pg = __uint_as_float(vv[q + 2]) + bf_lo(gb) + ss->bias[4 * jj + 2];
            float po = __uint_as_float(vv[q + 3]) + bf_hi(gb) + ss->bias[4 * jj + 3];
            float ig = ts::sigmoidf_fast(pi), fg = ts::sigmoidf_fast(pf), gg = ts::tanhf_fast(pg), og = ts::sigmoidf_fast(po);
            float cprev = reinterpret_cast<const float*>(cv)[jj];
            float c = fg * cprev + ig * gg;
            float h = og * ts::tanhf_fast(c);
            cn[jj] = c;
            apk[2 * jj] = pack_bf2(ig, fg);
            apk[2 * jj + 1] = pack_bf2(gg, og);
            if (jj & 1) hpk[jj >> 1] = pack_bf2(__uint_as_float(hpk[jj >> 1]), h); else hpk[jj >> 1] = __float_as_uint(h);
          }
          // h first: it is the only thing other CTAs wait for
          __nv_bfloat16* hp = p.h_seq + ((size_t)(t + 1) * B + row) * H + j0;
          stg16(hp, make_uint4(hpk[0], hpk[1], hpk[2], hpk[3]));
          stg16(hp + 8, make_uint4(hpk[4], hpk[5], hpk[6], hpk[7]));
        }
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");     // these rows are read next by TMA (async proxy)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (rloc == 0) {
          signal_counter(counter);
          if (p.dbg && blockIdx.x == 0) p.dbg[4 * t + 2] = gtime();
        }
        if (valid) {
          float* cp = p.c_seq + ((size_t)(t + 1) * B + row) * H + j0;
#pragma unroll
          for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(cp + 4 * i) = make_float4(cn[4 * i], cn[4 * i + 1], cn[4 * i + 2], cn[4 * i + 3]);
          __nv_bfloat16* ap = p.act + ((size_t)t * B + row) * (4 * H) + nb * 64;
#pragma unroll
          for (int i = 0; i < 8; ++i) stg16(ap + 8 * i, make_uint4(apk[4 * i], apk[4 * i + 1], apk[4 * i + 2], apk[4 * i + 3]));
        }
      }
    } else {
      const int j0 = nb * 64 + ks * 16;            // hidden columns this cluster member owns after the reduce-scatter
      const uint32_t xbase = tc::smem_u32(smem_x);
      const uint32_t xbar = tc::smem_u32(&ss->xchg_full);
      uint32_t xphase = 0;
      float dc[16], dh[16];
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 a = *reinterpret_cast<const float4*>(p.dc0 + (size_t)row * H + j0 + 4 * i);
          float4 b = *reinterpret_cast<const float4*>(p.dh0 + (size_t)row * H + j0 + 4 * i);
          dc[4 * i] = a.x; dc[4 * i + 1] = a.y; dc[4 * i + 2] = a.z; dc[4 * i + 3] = a.w;
          dh[4 * i] = b.x; dh[4 * i + 1] = b.y; dh[4 * i + 2] = b.z; dh[4 * i + 3] = b.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { dc[i] = 0.f; dh[i] = 0.f; }
      }
      for (int s = 0; s <= p.T && ok; ++s) {
        const int t = p.T - 1 - s;
        uint4 av[8], dhv[2];
        float4 cpv[4], cnv[4];
        if (valid && s < p.T) {
          const __nv_bfloat16* ap = p.act + ((size_t)t * B + row) * (4 * H) + 4 * j0;
#pragma unroll
          for (int i = 0; i < 8; ++i) av[i] = ldg_nc16(ap + 8 * i);
          const __nv_bfloat16* dp = p.dh_seq + ((size_t)t * B + row) * H + j0;
          dhv[0] = ldg_nc16(dp); dhv[1] = ldg_nc16(dp + 8);
          const float* c0p = p.c_seq + ((size_t)t * B + row) * H + j0;
          const float* c1p = p.c_seq + ((size_t)(t + 1) * B + row) * H + j0;
#pragma unroll
          for (int i = 0; i < 4; ++i) { cpv[i] = *reinterpret_cast<const float4*>(c0p + 4 * i); cnv[i] = *reinterpret_cast<const float4*>(c1p + 4 * i); }
        }
        if (s > 0) {
          ok = wait_bar<false>(&ss->tmem_full, tphase, abort_flag);
          tphase ^= 1;
          if (!ok) break;
          tc::fence_after_sync();
          if (p.dbg && blockIdx.x == 0 && rloc == 0) p.dbg[4 * s + 1] = gtime();
          uint32_t v0[32], v1[32];
          tc::tmem_ld32(taddr, v0);
          tc::tmem_ld32(taddr + 32, v1);
          tc::tmem_ld_wait();
          tc::fence_before_sync();
          // reduce-scatter over the 4 K-quarters: chunk q (16 columns) belongs to cluster member q
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t* src = q < 2 ? v0 + 16 * q : v1 + 16 * (q - 2);
            if (q == ks) {
#pragma unroll
              for (int i = 0; i < 16; ++i) dh[i] = __uint_as_float(src[i]);
            } else {
              const int slot = ks < q ? ks : ks - 1;                      // my index among q's three foreign sources
              const uint32_t dst = mapa(xbase + (uint32_t)((slot * BM + rloc) * 64), (uint32_t)q);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                st_cluster_f4(dst + 16 * i, make_float4(__uint_as_float(src[4 * i]), __uint_as_float(src[4 * i + 1]),
                                                       __uint_as_float(src[4 * i + 2]), __uint_as_float(src[4 * i + 3])));
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (rloc < 4 && rloc != ks) mbar_arrive_remote(mapa(xbar, (uint32_t)rloc));
          ok = wait_bar<true>(&ss->xchg_full, xphase, abort_flag);
          xphase ^= 1;
          if (!ok) break;
#pragma unroll
          for (int src = 0; src < 3; ++src) {
            const float4* xp = reinterpret_cast<const float4*>(smem_x + (size_t)(src * BM + rloc) * 64);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 x4 = xp[i];
              dh[4 * i] += x4.x; dh[4 * i + 1] += x4.y; dh[4 * i + 2] += x4.z; dh[4 * i + 3] += x4.w;
            }
          }
        }
        if (s == p.T) {
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              *reinterpret_cast<float4*>(p.dh0 + (size_t)row * H + j0 + 4 * i) = make_float4(dh[4 * i], dh[4 * i + 1], dh[4 * i + 2], dh[4 * i + 3]);
              *reinterpret_cast<float4*>(p.dc0 + (size_t)row * H + j0 + 4 * i) = make_float4(dc[4 * i], dc[4 * i + 1], dc[4 * i + 2], dc[4 * i + 3]);
            }
          }
          break;
        }
        uint32_t gpk[32];
        if (valid) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const uint4 a4 = av[jj >> 1];
            const uint32_t aa = (jj & 1) ? a4.z : a4.x, ab = (jj & 1) ? a4.w : a4.y;
            const float ig = bf_lo(aa), fg = bf_hi(aa), gg = bf_lo(ab), og = bf_hi(ab);
            const uint4 d4 = dhv[jj >> 3];
            const uint32_t dw = ((jj >> 1) & 3) == 0 ? d4.x : ((jj >> 1) & 3) == 1 ? d4.y : ((jj >> 1) & 3) == 2 ? d4.z : d4.w;
            const float dht = dh[jj] + ((jj & 1) ? bf_hi(dw) : bf_lo(dw));
            const float cprev = reinterpret_cast<const float*>(cpv)[jj];
            const float tcn = ts::tanhf_fast(reinterpret_cast<const float*>(cnv)[jj]);
            const float dct = dc[jj] + dht * og * (1.f - tcn * tcn);
            const float d_o = dht * tcn, d_i = dct * gg, d_f = dct * cprev, d_g = dct * ig;
            dc[jj] = dct * fg;
            gpk[2 * jj] = pack_bf2(d_i * ig * (1.f - ig), d_f * fg * (1.f - fg));
            gpk[2 * jj + 1] = pack_bf2(d_g * (1.f - gg * gg), d_o * og * (1.f - og));
            dh[jj] = 0.f;
          }
          __nv_bfloat16* gp = p.dpre + ((size_t)t * B + row) * (4 * H) + 4 * j0;
#pragma unroll
          for (int i = 0; i < 8; ++i) stg16(gp + 8 * i, make_uint4(gpk[4 * i], gpk[4 * i + 1], gpk[4 * i + 2], gpk[4 * i + 3]));
        }
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (rloc == 0) {
          signal_counter(counter);
          if (p.dbg && blockIdx.x == 0) p.dbg[4 * s + 2] = gtime();
        }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();          // nobody exits while a peer may still multicast into / arrive on its smem
  if (threadIdx.x == 0 && ss->abort_flag) atomicExch(reinterpret_cast<int*>(p.sync + 63), 1);
  if (warp == 2) tc::tmem_dealloc(tmem_d, 64);
}

size_t smem_bytes(int H, bool bwd) {
  return (size_t)(H / BK) * kWBlockBytes + kStages * kABytes + (bwd ? kXchgBytes : 0) + sizeof(SeqSmem) + 1024;
}

template <bool kBwd, int kCluster>
int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tw, const SeqParams& p, int grid, size_t smem, bool dry, cudaStream_t st) {
  auto kern = lstm_seq_kernel<kBwd, kCluster>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (kCluster > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (kCluster > 1) {
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return -20; }
    if (nclusters * kCluster < grid) return -21;              // not co-resident with this cluster size
  }
  if (dry) return 0;
  e = cudaLaunchKernelEx(&cfg, kern, ta, tw, p);
  return (int)e;
}

int g_fwd_cluster = -1;   // resolved once: largest multicast cluster that is co-resident for the current shape
int g_fwd_cluster_key = 0;

}  // namespace

// sync_ws: >= 64 u32; [0..tiles_m) step counters (zeroed by the caller before every launch), [63] sticky error flag.
// cluster: requested multicast cluster size for the forward kernel (1,2,4,8; 0 = auto: largest that is co-resident).
extern "C" int ts_lstm_seq_fwd(const void* gx, const void* w_h, const float* bias, const void* h_seq, const float* c_seq,
                               void* act, float*, void* dbg, void*, int T, int B, int H, unsigned int* sync_ws, int cluster,
                               cudaStream_t st) {
  SeqParams p{};
  p.gx = (const __nv_bfloat16*)gx; p.bias = bias; p.h_seq = (__nv_bfloat16*)h_seq; p.c_seq = (float*)c_seq;
  p.act = (__nv_bfloat16*)act; p.sync = sync_ws; p.T = T; p.B = B; p.H = H; p.dbg = (unsigned long long*)dbg;
  if (H % 64 != 0) { ts::set_last_error("lstm_seq: H must be a multiple of 64"); return -2; }
  const int tiles_m = (B + BM - 1) / BM, tiles_n = 4 * H / BN;
  int dev = 0;
  cudaGetDevice(&dev);
  if (tiles_m * tiles_n > ts::sm_count(dev)) { ts::set_last_error("lstm_seq: grid exceeds SM count (not co-resident)"); return -3; }
  const size_t smem = smem_bytes(H, false);
  if (smem > 227 * 1024) { ts::set_last_error("lstm_seq: weight slice does not fit in shared memory"); return -4; }
  CUtensorMap ta, tw;
  if (int rc = ts::make_tmap_3d_bf16(&ta, h_seq, (uint64_t)H, (uint64_t)B, (uint64_t)T + 1, (uint64_t)H, (uint64_t)H * B, BK, BM, 1)) return rc;
  if (int rc = ts::make_tmap_2d_bf16(&tw, w_h, (uint64_t)4 * H, (uint64_t)H, (uint64_t)H, BK, BN)) return rc;
  p.tiles_n = tiles_n;
  const int grid = tiles_m * tiles_n;
  int c = cluster;
  if (c == 0) {
    const int key = H * 1024 + tiles_m;
    if (g_fwd_cluster < 0 || g_fwd_cluster_key != key) {
      g_fwd_cluster = 1; g_fwd_cluster_key = key;
      if (tiles_n % 8 == 0 && launch_cfg<false, 8>(ta, tw, p, grid, smem, true, st) == 0) g_fwd_cluster = 8;
      else if (tiles_n % 4 == 0 && launch_cfg<false, 4>(ta, tw, p, grid, smem, true, st) == 0) g_fwd_cluster = 4;
      else if (tiles_n % 2 == 0 && launch_cfg<false, 2>(ta, tw, p, grid, smem, true, st) == 0) g_fwd_cluster = 2;
    }
    c = g_fwd_cluster;
  }
  if (c > 1 && tiles_n % c != 0) c = 1;
  switch (c) {
    case 8: return launch_cfg<false, 8>(ta, tw, p, grid, smem, false, st);
    case 4: return launch_cfg<false, 4>(ta, tw, p, grid, smem, false, st);
    case 2: return launch_cfg<false, 2>(ta, tw, p, grid, smem, false, st);
    default: return launch_cfg<false, 1>(ta, tw, p, grid, smem, false, st);
  }
}

extern "C" int ts_lstm_seq_bwd(const void* dh_seq, const void* w_hT, const void* act, const float* c_seq, const void* dpre,
                               float* dh0, float* dc0, void* dbg, void*, int T, int B, int H, unsigned int* sync_ws, int,
                               cudaStream_t st) {
  SeqParams p{};
  p.dh_seq = (const __nv_bfloat16*)dh_seq; p.act = (__nv_bfloat16*)act; p.c_seq = (float*)c_seq; p.dpre = (__nv_bfloat16*)dpre;
  p.dh0 = dh0; p.dc0 = dc0; p.sync = sync_ws; p.T = T; p.B = B; p.H = H; p.dbg = (unsigned long long*)dbg;
  if (H % 64 != 0) { ts::set_last_error("lstm_seq: H must be a multiple of 64"); return -2; }
  const int tiles_m = (B + BM - 1) / BM, tiles_n = (H / BN) * 4;        // (nb2, ks)
  int dev = 0;
  cudaGetDevice(&dev);
  if (tiles_m * tiles_n > ts::sm_count(dev)) { ts::set_last_error("lstm_seq: grid exceeds SM count (not co-resident)"); return -3; }
  const size_t smem = smem_bytes(H, true);
  if (smem > 227 * 1024) { ts::set_last_error("lstm_seq: weight slice does not fit in shared memory"); return -4; }
  CUtensorMap ta, tw;
  if (int rc = ts::make_tmap_3d_bf16(&ta, dpre, (uint64_t)4 * H, (uint64_t)B, (uint64_t)T, (uint64_t)4 * H, (uint64_t)4 * H * B, BK, BM, 1)) return rc;
  if (int rc = ts::make_tmap_2d_bf16(&tw, w_hT, (uint64_t)H, (uint64_t)4 * H, (uint64_t)4 * H, BK, BN)) return rc;
  p.tiles_n = tiles_n;
  int rc = launch_cfg<true, 4>(ta, tw, p, tiles_m * tiles_n, smem, false, st);
  if (rc == -21) ts::set_last_error("lstm_seq_bwd: clusters of 4 are not co-resident on this device");
  return rc;
}
