// K-LSTM / K-LSTM-BWD: persistent tcgen05 kernels that run a whole layer's recurrence in ONE launch.
//
// What they replace (reference, per layer per time step): 4x tf.matmul(ht, W_h) + bias adds + 3 sigmoid + 2 tanh +
// the c/h update, each its own TF op (/root/reference/src/models/recurrent/lstm.py:88-109; K1,K3-K8 in SURVEY §2.5),
// and the mirrored autodiff backward (K13).  The reference only ever takes ONE step; these kernels deliver the
// multi-step unroll its fit_next API was built for (lstm.py:128-136).
//
// Design (B200-first):
//   * The recurrent weight slice a CTA needs is loaded ONCE by TMA and stays resident in shared memory for all T steps
//     (W_h is 8 MB in bf16 at H=1024: 128 CTAs x 64 KB..128 KB).  Rows are gate-interleaved (n = 4j+g) so a CTA that
//     owns a slice of rows owns complete (i,f,g,o) quadruples: the gate epilogue needs no cross-CTA traffic.
//   * Per step a CTA streams its 128-row batch tile of h_{t-1} (forward) / dG_{t+1} (backward) through a 4-stage
//     TMA->mbarrier ring, one elected thread issues tcgen05.mma (M=128, N=64 fwd / 16 bwd, K=16, bf16 -> fp32 in TMEM),
//     and four epilogue warps read the accumulator with tcgen05.ld and do the whole cell in registers:
//       fwd: + x-projection + bias, sigmoid/tanh, c_t = f*c + i*g, h_t = o*tanh(c_t)  -> h_t (bf16, next step's operand),
//            c_t (fp32) and the activated gates (bf16, saved for backward)
//       bwd: dh = dh_above + dh_rec, gate gradients -> dG_t (bf16, next step's operand and the dW GEMM operand)
//   * Steps are separated by a grid-wide dataflow barrier in global memory (one monotonically increasing counter per
//     batch tile, red.release.gpu / ld.acquire.gpu, generic->async proxy fences around it because the consumer is TMA).
//     All CTAs are co-resident (grid <= #SMs, 1 CTA/SM) so the barrier cannot deadlock; every spin is bounded and
//     raises an error flag instead of hanging the GPU.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "tcgen05.cuh"
#include "tmap.h"
#include "ts_common.cuh"

namespace {

constexpr int BM = 128;           // batch rows per CTA (UMMA M)
constexpr int BK = 64;            // K per pipeline stage (one 128 B swizzle atom of bf16)
constexpr int UK = 16;            // UMMA K
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr long long kSpinLimit = 6000000000LL;   // ~3 s of SM clocks: a bug surfaces as an error, not a hung GPU

struct SeqSmem {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t w_full;
  uint64_t tmem_full;
  uint32_t tmem_slot;
  int abort_flag;
  float bias[64];
};

TC_DEVICE bool wait_bar(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
  if (tc::mbar_try_wait(bar, parity)) return true;
  long long t0 = clock64();
  int n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
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

TC_DEVICE uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
TC_DEVICE uint4 ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
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
  unsigned int* sync;          // [tiles_m] step counters + [1] error flag at sync[63]
  int T, B, H;
  int tiles_n;                 // CTAs per batch tile
  unsigned int sync_base;      // counter value at launch (counters are never reset)
};

// kBwd = false: N = 64 gate columns (16 hidden units), K = H
// kBwd = true : N = 16 hidden columns of dh_{t-1},     K = 4H
template <bool kBwd>
__global__ void __launch_bounds__(kThreads, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                const SeqParams p) {
  constexpr int BN = kBwd ? 16 : 64;
  constexpr int kTmemCols = kBwd ? 32 : 64;
  constexpr int kABytes = BM * BK * 2;          // 16 KB
  constexpr int kWBlockBytes = BN * BK * 2;     // 8 KB fwd / 2 KB bwd per 64-wide K block

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int K = kBwd ? 4 * p.H : p.H;
  const int num_kb = K / BK;
  uint8_t* smem_w = smem;                                    // resident weight slice: num_kb blocks
  uint8_t* smem_a = smem + (size_t)num_kb * kWBlockBytes;    // kStages x 16 KB
  SeqSmem* ss = reinterpret_cast<SeqSmem*>(smem_a + kStages * kABytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x / p.tiles_n, nb = blockIdx.x % p.tiles_n;
  unsigned int* counter = p.sync + mb;
  volatile int* abort_flag = &ss->abort_flag;
  const int steps = kBwd ? p.T + 1 : p.T;       // backward runs one extra GEMM to produce dh_0

  if (threadIdx.x == 0) {
    ss->abort_flag = 0;
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&ss->full[s], 1); tc::mbar_init(&ss->empty[s], 1); }
    tc::mbar_init(&ss->w_full, 1);
    tc::mbar_init(&ss->tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (!kBwd && threadIdx.x >= 64 && threadIdx.x < 128) ss->bias[threadIdx.x - 64] = p.bias[nb * BN + threadIdx.x - 64];
  if (warp == 2) { tc::tmem_alloc(&ss->tmem_slot, kTmemCols); tc::tmem_relinquish(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = ss->tmem_slot;

  if (warp == 0) {
    // ======================================================================== TMA producer
    if (lane == 0) {
      tc::mbar_expect_tx(&ss->w_full, (uint32_t)(num_kb * kWBlockBytes));
      for (int kb = 0; kb < num_kb; ++kb)
        tc::tma_load_2d(smem_w + (size_t)kb * kWBlockBytes, &tmap_w, &ss->w_full, kb * BK, nb * BN);
      int stage = 0; uint32_t phase = 0;
      bool ok = true;
      for (int s = kBwd ? 1 : 0; s < steps && ok; ++s) {
        // forward step s consumes h_seq[s] (rows written by step s-1); backward iteration s consumes dG[T-s]
        if (s > 0) ok = wait_counter(counter, p.sync_base + (unsigned)s * p.tiles_n, abort_flag);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        const int tsl = kBwd ? p.T - s : s;
        for (int kb = 0; kb < num_kb && ok; ++kb) {
          ok = wait_bar(&ss->empty[stage], phase ^ 1, abort_flag);
          if (!ok) break;
          tc::mbar_expect_tx(&ss->full[stage], kABytes);
          tc::tma_load_3d(smem_a + stage * kABytes, &tmap_a, &ss->full[stage], kb * BK, mb * BM, tsl);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM, BN);
      bool ok = wait_bar(&ss->w_full, 0, abort_flag);
      int stage = 0; uint32_t phase = 0;
      for (int s = kBwd ? 1 : 0; s < steps && ok; ++s) {
        for (int kb = 0; kb < num_kb; ++kb) {
          ok = wait_bar(&ss->full[stage], phase, abort_flag);
          if (!ok) break;
          tc::fence_after_sync();
          const uint64_t da = tc::desc_kmajor_sw128(tc::smem_u32(smem_a + stage * kABytes));
          const uint64_t db = tc::desc_kmajor_sw128(tc::smem_u32(smem_w + (size_t)kb * kWBlockBytes));
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            tc::mma_bf16_ss(tmem_d, tc::desc_advance(da, k * UK * 2), tc::desc_advance(db, k * UK * 2), idesc, (kb | k) != 0);
          tc::mma_commit(&ss->empty[stage]);
          if (kb == num_kb - 1) tc::mma_commit(&ss->tmem_full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ======================================================================== epilogue: one thread = one batch row
    const int ew = warp - kEpiWarp0;
    const int row = mb * BM + ew * 32 + lane;
    const bool valid = row < p.B;
    const int H = p.H, B = p.B;
    const int j0 = nb * 16;                       // 16 hidden units per CTA in both directions
    const uint32_t taddr = tmem_d + ((uint32_t)(ew * 32) << 16);
    uint32_t tphase = 0;
    bool ok = true;

    if (!kBwd) {
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
        ok = wait_bar(&ss->tmem_full, tphase, abort_flag);
        tphase ^= 1;
        if (!ok) break;
        tc::fence_after_sync();
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
        if (ew == 0 && lane == 0) signal_counter(counter);
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
          const __nv_bfloat16* ap = p.act + ((size_t)t * B + row) * (4 * H) + nb * 64;
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
          ok = wait_bar(&ss->tmem_full, tphase, abort_flag);
          tphase ^= 1;
          if (!ok) break;
          tc::fence_after_sync();
          uint32_t v[16];
          tc::tmem_ld16(taddr, v);
          tc::tmem_ld_wait();
          tc::fence_before_sync();
#pragma unroll
          for (int i = 0; i < 16; ++i) dh[i] = __uint_as_float(v[i]);       // dh_rec for time t
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
          __nv_bfloat16* gp = p.dpre + ((size_t)t * B + row) * (4 * H) + nb * 64;
#pragma unroll
          for (int i = 0; i < 8; ++i) stg16(gp + 8 * i, make_uint4(gpk[4 * i], gpk[4 * i + 1], gpk[4 * i + 2], gpk[4 * i + 3]));
        }
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");     // these rows are read next by TMA (async proxy)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ew == 0 && lane == 0) signal_counter(counter);
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0 && ss->abort_flag) atomicExch(reinterpret_cast<int*>(p.sync + 63), 1);
  if (warp == 2) tc::tmem_dealloc(tmem_d, kTmemCols);
}

template <bool kBwd>
int launch_seq(const void* a_base, uint64_t a_t, const void* w_base, const SeqParams& p0, int variant, cudaStream_t st) {
  SeqParams p = p0;
  constexpr int BN = kBwd ? 16 : 64;
  const int K = kBwd ? 4 * p.H : p.H;
  const int N = kBwd ? p.H : 4 * p.H;
  if (p.H % 64 != 0) { ts::set_last_error("lstm_seq: H must be a multiple of 64"); return -2; }
  const int tiles_m = (p.B + BM - 1) / BM;
  const int tiles_n = N / BN;
  int dev = 0;
  cudaGetDevice(&dev);
  if (tiles_m * tiles_n > ts::sm_count(dev)) { ts::set_last_error("lstm_seq: grid exceeds SM count (not co-resident)"); return -3; }
  const int num_kb = K / BK;
  const size_t smem = (size_t)num_kb * BN * BK * 2 + kStages * BM * BK * 2 + sizeof(SeqSmem) + 1024;
  if (smem > 227 * 1024) { ts::set_last_error("lstm_seq: weight slice does not fit in shared memory"); return -4; }
  CUtensorMap ta, tw;
  if (int rc = ts::make_tmap_3d_bf16(&ta, a_base, (uint64_t)K, (uint64_t)p.B, a_t, (uint64_t)K, (uint64_t)K * p.B, BK, BM, 1)) return rc;
  if (int rc = ts::make_tmap_2d_bf16(&tw, w_base, (uint64_t)N, (uint64_t)K, (uint64_t)K, BK, BN)) return rc;
  auto kern = lstm_seq_kernel<kBwd>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  p.tiles_n = tiles_n;
  (void)variant;
  kern<<<tiles_m * tiles_n, kThreads, smem, st>>>(ta, tw, p);
  return (int)cudaGetLastError();
}

}  // namespace

// sync_ws: >= 64 u32, zero-initialised once; [0..tiles_m) step counters (monotonic), [63] error flag.
// sync_base: the value the counters hold at launch (host tracks it: += steps*tiles_n per launch).
extern "C" int ts_lstm_seq_fwd(const void* gx, const void* w_h, const float* bias, const void* h_seq, const float* c_seq,
                               void* act, float*, void*, void*, int T, int B, int H, unsigned int* sync_ws, int sync_base,
                               cudaStream_t st) {
  SeqParams p{};
  p.gx = (const __nv_bfloat16*)gx; p.bias = bias; p.h_seq = (__nv_bfloat16*)h_seq; p.c_seq = (float*)c_seq;
  p.act = (__nv_bfloat16*)act; p.sync = sync_ws; p.T = T; p.B = B; p.H = H; p.sync_base = (unsigned)sync_base;
  return launch_seq<false>(h_seq, (uint64_t)T + 1, w_h, p, 0, st);
}

extern "C" int ts_lstm_seq_bwd(const void* dh_seq, const void* w_hT, const void* act, const float* c_seq, const void* dpre,
                               float* dh0, float* dc0, void*, void*, int T, int B, int H, unsigned int* sync_ws,
                               int sync_base, cudaStream_t st) {
  SeqParams p{};
  p.dh_seq = (const __nv_bfloat16*)dh_seq; p.act = (__nv_bfloat16*)act; p.c_seq = (float*)c_seq; p.dpre = (__nv_bfloat16*)dpre;
  p.dh0 = dh0; p.dc0 = dc0; p.sync = sync_ws; p.T = T; p.B = B; p.H = H; p.sync_base = (unsigned)sync_base;
  return launch_seq<true>(dpre, (uint64_t)T, w_hT, p, 0, st);
}
