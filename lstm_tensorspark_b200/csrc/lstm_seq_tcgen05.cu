// K-LSTM / K-LSTM-BWD: persistent tcgen05 kernels that run a whole layer's recurrence in ONE launch.
//
// What they replace (reference, per layer per time step): 4x tf.matmul(ht, W_h) + bias adds + 3 sigmoid + 2 tanh +
// the c/h update, each its own TF op (/root/reference/src/models/recurrent/lstm.py:88-109; K1,K3-K8 in SURVEY §2.5),
// and the mirrored autodiff backward (K13).  The reference only ever takes ONE step; these kernels deliver the
// multi-step unroll its fit_next API was built for (lstm.py:128-136).
//
// Design (B200-first):
//   * Every CTA keeps a bf16 slice of the recurrent weights RESIDENT in shared memory for all T steps (loaded once by
//     TMA; W_h is 8 MB at H=1024 = 128 CTAs x 128 KB).  Rows are gate-interleaved (n = 4j+g) so a CTA that owns a row
//     slice owns complete (i,f,g,o) quadruples: the gate epilogue needs no cross-CTA traffic.
//   * Per step a CTA streams a 128-row batch tile of h_{t-1} (forward) / dG_{t+1} (backward) through a bulk-copy ->
//     mbarrier ring (the operand is kept in global memory as ready-made 128B-swizzled tile images, 16 KB contiguous per
//     k-block); one elected thread issues tcgen05.mma (M=128, K=16, bf16 -> fp32 accumulators in TMEM); eight epilogue
//     warps read the accumulator with tcgen05.ld and do the whole cell in registers (the cell state never leaves them).
//   * The stream is bound by (bytes in flight) / (L2 latency), not by bandwidth: the ring next to a 128 KB weight slice
//     holds 5-6 of the k-blocks.  So BOTH passes split K across a thread-block cluster - forward 2 CTAs (8 k-blocks each,
//     N=128), backward 4 CTAs (one gate-column quarter each, N=64) - and reduce-scatter the partial accumulators through
//     DISTRIBUTED SHARED MEMORY with st.async (bytes are counted on the receiver's mbarrier: no release/acquire fences,
//     which the compiler lowers to MEMBAR.ALL.GPU even at cluster scope).  Every member then owns 16 hidden units.
//   * Steps are not separated by a grid barrier but by DATAFLOW: every operand k-block has an arrival counter in global
//     memory (its producer CTAs: epilogue stores -> CTA barrier -> ONE red.release.gpu); the producer warp's 32 lanes poll
//     all counters of their K slice with relaxed loads and pull the blocks into the ring in arrival order (accumulation
//     order is free; the ring stage carries its k-block id to the MMA issuer).  The consumer is the async proxy (L2),
//     so there is no acquire fence; a tmem_empty mbarrier keeps the next step's first MMA off an accumulator that is
//     still being read.
//   * What a release costs decides the step time: MEMBAR.ALL.GPU drains every outstanding store of the SM, so the
//     bookkeeping stores (h_seq / c_seq / activations for backward) are held back until the signal has left, and they are
//     256-bit (STG.256): their L2 request COUNT, not their bytes, is what used to slow the operand stream down.
//   * All CTAs are co-resident (grid <= #SMs, 1 CTA/SM, checked with cudaOccupancyMaxActiveClusters); every spin is
//     bounded and raises a sticky error flag instead of hanging the GPU.
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
constexpr int kMaxStages = 9;
constexpr int kEpiWarp0 = 4;
constexpr int kABytes = BM * BK * 2;          // 16 KB
constexpr int kWBlockBytes = BN * BK * 2;     // 8 KB per 64-wide K block
constexpr int kXchgBytes = 4 * BM * 16 * 2;   // backward: 4 source slots of [128 x 16] bf16 partial chunks
constexpr long long kSpinLimit = 6000000000LL;   // ~3 s of SM clocks: a bug surfaces as an error, not a hung GPU
constexpr long long kAbortGrace = 1000000000LL;  // ~0.5 s for the role loops to drain after an abort before the watchdog traps

// sync workspace (u32 words): [0,16) grid-barrier counters, [64,320) per-CTA step flags, [512 + 32 i) k-block arrival
// counters (one 128 B line each, i < tiles_m * 4H/64 <= 148 + ...), [kSyncWords-1] sticky error flag
constexpr int kSyncWords = 8192;
constexpr int kSyncErr = kSyncWords - 1;
constexpr int kSyncFlags = 64;
constexpr int kSyncKb = 512;

struct SeqSmem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t w_full;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];       // epilogue -> MMA issuer: the accumulator has been read (8 warp arrivals)
  uint64_t xchg_full[2];
  uint64_t xchg_free[2];        // every cluster member has consumed its exchange buffer of the previous step (4 remote arrivals)
  uint32_t tmem_slot;
  int abort_flag;
  int roles_done;               // role warps that have left their loops (producer, MMA issuer, 8 epilogue warps)
  uint32_t kb_idx[kMaxStages];  // which k-block sits in ring stage s (the producer fills stages in ARRIVAL order)
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
// 16 B into a cluster member's shared memory; the bytes are accounted on ITS mbarrier (complete_tx), so the receiver needs
// no release/acquire round trip: it arms the barrier with expect_tx and waits, exactly as for a TMA load.
TC_DEVICE void st_async_u4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar) : "memory");
}
// "Buffer consumed" notifications carry no data: the reads they order were complete (their values used) before the CTA
// barrier that precedes the arrive.  The .release form compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR - a second
// GPU-scope fence per step that competes with the one the dataflow signal needs.
TC_DEVICE void mbar_arrive_remote_relaxed(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
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

TC_DEVICE unsigned int ld_acquire_gpu(const unsigned int* ctr) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  return v;
}
TC_DEVICE uint2 ld_relaxed_gpu_v2(const unsigned int* ctr) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ctr) : "memory");
  return v;
}
TC_DEVICE void st_release_gpu(unsigned int* ctr, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ctr), "r"(v) : "memory");
}
TC_DEVICE unsigned int ld_relaxed_gpu(const unsigned int* ctr) {      // coalesces across lanes (a divergent ld.acquire does not)
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  return v;
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
// 256-bit accesses (sm_100 LDG/STG.256): half as many L2 requests as 16 B ones and only whole 32 B sectors - the
// bookkeeping stores of a step are ~460 K requests chip-wide, and it is their COUNT that slows the operand stream down.
struct alignas(32) U8 { uint32_t v[8]; };
TC_DEVICE U8 ldg_nc32(const void* p) {
  U8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
  return r;
}
TC_DEVICE U8 ldg_cg32(const void* p) {          // coherent at L2: the producer is a kernel that is still running
  U8 r;
  asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p) : "memory");
  return r;
}
TC_DEVICE uint4 ldg_cg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
// Wait (one lane polls, the warp follows) until the gated GEMM has published the 128 x 256 block that holds this warp's rows.
TC_DEVICE bool wait_in_gate(const unsigned int* ctr, int lane, volatile int* abort_flag) {
  int ok = 1;
  if (lane == 0) {
    long long t0 = clock64();
    int n = 0;
    while (ld_acquire_gpu(ctr) == 0u) {
      if ((++n & 63) == 0) {
        if (*abort_flag) { ok = 0; break; }
        if (clock64() - t0 > kSpinLimit) { *abort_flag = 1; ok = 0; break; }
      }
    }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok != 0;
}
TC_DEVICE void stg32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
TC_DEVICE void stg16(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
TC_DEVICE void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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
  const __nv_bfloat16* dh_seq; // [T,B,H] gradient into every h_t from above; null = only the final state has one (dh0 in)
  __nv_bfloat16* dpre;         // [T,B,4H]
  float* dh0;                  // [B,H] in: dL/dh_T extra, out: dL/dh_0
  float* dc0;                  // [B,H] in: dL/dc_T, out: dL/dc_0
  __nv_bfloat16* a_tiled;      // streamed operand as pre-swizzled SMEM images: [time][tiles_m][K/64][128 rows][64] (16 KB blocks)
  int tiles_m;
  int debug_mode;              // timing experiments only: 1 = skip operand loads, 2 = skip MMAs, 4 = half-size loads (garbage results), 3 = in-order stream
  unsigned int* sync;          // [63] error flag; [64 + mb * nkb + kb] arrival counter of operand k-block kb of batch tile mb
  unsigned long long* dbg;     // optional [steps][4] timestamps of CTA 0 (ns)
  int T, B, H;
  int tiles_n;                 // CTAs per batch tile
  int sync_mode;               // 0 = k-block arrival counters (dataflow), 1 = one counter per batch tile (grid barrier), 2 = per-CTA flags
  int poll_acquire;            // experiment: ld.acquire polls instead of relaxed
  // Layer wavefront (two layers' recurrences co-resident, chained through a dataflow-gated GEMM on the idle SMs):
  const unsigned int* in_gate; // completion counters of the GEMM that produces gx (fwd) / dh_seq (bwd) tile by tile while this kernel
                               // runs: [(row / 128) * in_gate_tiles_n + col / 256]; null = the operand is complete
  int in_gate_tiles_n;
  int pdl_wait;                // launched as a programmatic dependent: before exiting, wait for the grids launched before this one
                               // (completion of the chain's LAST kernel then implies completion of all of them)
  int no_trap;                 // debugging: the watchdog only records the abort (variant bit 20) instead of killing the kernel
  int extra_signal;            // one more arrival on this CTA's k-block counter after the LAST step's bookkeeping stores (a gated
                               // GEMM consumes h_seq / dpre in the natural layout, which is written after the per-step signal)
};

// Work decomposition
//   forward : CTA (mb, nb)      -> gate columns [64 nb, +64) = hidden [16 nb, +16) of batch tile mb, K = H.
//   backward: CTA (mb, nb2, ks) -> partial dh columns [64 nb2, +64) over gate-column quarter ks (K = H); the 4 ks form a
//                                 cluster; after the DSMEM reduce-scatter member ks owns hidden [64 nb2 + 16 ks, +16).
// Warps: 0 = producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4..11 = epilogue (warp % 4 = TMEM lane quarter,
// (warp-4)/4 = which 32 of the 64 accumulator columns; one thread = one batch row x 8 hidden units).
// The two role warps run CONVERGED and issue under elect.sync so descriptors / addresses stay in uniform registers
// (an `if (lane == 0)` role loop makes ptxas wrap every UTCHMMA / UBLKCP in a waterfall loop: ~590 cycles per k-block);
// k-blocks are handled in groups (overlapped mbarrier.try_waits, one fence + one elect per group).
// kTiles = 2: the CTA alternates TWO independent 128-row batch tiles (same resident weight slice): while one tile sits
// in its epilogue + grid barrier (latency), the other one streams its operand and runs its MMAs.  Half as many CTAs are
// needed (64 for B = 256, H = 1024), which leaves SMs free for the weight-gradient GEMMs that run concurrently.
// kStream = true (H too large for a resident slice, e.g. H = 2048: W_h alone is 32 MB): the weight k-block travels through
// the ring next to its operand k-block (24 KB stages, W comes out of L2 every step); everything else is unchanged.
// kFSplit (forward): a cluster of 2 CTAs splits K.  Each member contracts over HALF of h_{t-1} (8 instead of 16 operand
// k-blocks per step at H = 1024: the ring then holds most of the step's operand at once - the stream is bound by bytes in
// flight / L2 latency, not by bandwidth) against a [128 x H/2] weight slice (N = 128: 64 instead of 48 tensor-pipe cycles per
// K = 16, but half as many instructions), and the two partial [128 x 128] accumulators are reduce-scattered through DSMEM:
// every member ends up with the same 64 gate columns = 16 hidden units it owns in the unsplit kernel.
template <bool kBwd, int kStages, int kTiles, bool kStream, bool kFSplit = false>
__global__ void __launch_bounds__(384, 1)
lstm_seq_kernel(const __grid_constant__ CUtensorMap tmap_w, const SeqParams p) {
  static_assert(!(kFSplit && (kBwd || kStream || kTiles != 1)), "forward K-split: one tile, resident weights");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kSplit = kBwd ? 4 : (kFSplit ? 2 : 1);       // cluster size = K-split factor
  constexpr bool kCluster = kSplit > 1;
  constexpr int kBNm = kFSplit ? 2 * BN : BN;                // accumulator columns per CTA (UMMA N)
  constexpr int kWBlk = kBNm * BK * 2;                       // bytes of one resident weight k-block
  const int num_kb = kFSplit ? p.H / (2 * BK) : p.H / BK;    // operand k-blocks this CTA contracts over per step
  constexpr int kStageBytes = kStream ? kABytes + kWBlockBytes : kABytes;
  uint8_t* smem_w = smem;                                    // resident weight slice: num_kb blocks of [kBNm x 64] (not kStream)
  uint8_t* smem_a = smem + (kStream ? 0 : (size_t)num_kb * kWBlk);    // kStages x (16 KB [+ 8 KB weight block])
  uint8_t* smem_x = smem_a + kStages * kStageBytes;          // DSMEM exchange buffer (bf16 partial sums)
  SeqSmem* ss = reinterpret_cast<SeqSmem*>(smem_x + (kCluster ? kTiles * kXchgBytes : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Programmatic dependent launch: a kernel queued behind this one WITH the PDL attribute (the fused allreduce + update of a
  // gradient bucket that is already complete) may start once every CTA of this grid is resident - it then runs on the SMs
  // this persistent grid leaves idle instead of after it.  No effect on ordinary launches.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int mb0 = (blockIdx.x / p.tiles_n) * kTiles;          // first batch tile of this CTA
  const int in_mb = blockIdx.x % p.tiles_n;
  const uint32_t crank = kCluster ? cluster_ctarank() : 0;
  const int nb = in_mb / kSplit;                              // weight-row block of the cluster (64 rows; kFSplit: 128 rows)
  const int ks = (int)crank;                                  // K-split member
  volatile int* abort_flag = &ss->abort_flag;
  const int steps = kBwd ? p.T + 1 : p.T;                    // backward runs one extra GEMM to produce dh_0

  if (threadIdx.x == 0) {
    ss->abort_flag = 0;
    ss->roles_done = 0;
    tc::prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&ss->full[s], 1); tc::mbar_init(&ss->empty[s], 1); }
    tc::mbar_init(&ss->w_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&ss->tmem_full[i], 1); tc::mbar_init(&ss->tmem_empty[i], 8);
      tc::mbar_init(&ss->xchg_full[i], 1);                   // armed locally (expect_tx = the whole 16 KB buffer), filled by st.async
      tc::mbar_init(&ss->xchg_free[i], kFSplit ? 1 : 4);     // remote arrivals: every writer of this buffer's readers
    }
    if (kCluster)
      for (int i = 0; i < kTiles; ++i) tc::mbar_expect_tx_u32(tc::smem_u32(&ss->xchg_full[i]), (uint32_t)kXchgBytes);
    tc::fence_barrier_init();
  }
  if (!kBwd && threadIdx.x >= 64 && threadIdx.x < 128) ss->bias[threadIdx.x - 64] = p.bias[in_mb * BN + threadIdx.x - 64];
  constexpr uint32_t kTmemCols = (kTiles == 2 || kFSplit) ? 128 : 64;
  if (warp == 2) { tc::tmem_alloc(&ss->tmem_slot, kTmemCols); tc::tmem_relinquish(); }
  tc::fence_before_sync();
  __syncthreads();
  if (kCluster) cluster_sync_all();             // peers' mbarriers are initialised before anyone arrives remotely
  tc::fence_after_sync();
  const uint32_t tmem_d = ss->tmem_slot;

  if (warp == 0) {
    // ======================================================================== producer
    const uint32_t w_bar = tc::smem_u32(&ss->w_full);
    if (!kStream && tc::elect_one()) {
      tc::mbar_expect_tx_u32(w_bar, (uint32_t)(num_kb * kWBlk));
      // forward: rows = gate columns [kBNm nb, +kBNm) of W_h [4H, H] (K-split: K offset = half ks).  backward: rows = hidden
      // columns [64 nb, +64) of W_h^T [H, 4H], K offset = quarter ks.
      for (int kb = 0; kb < num_kb; ++kb)
        tc::tma_load_2d_u32(tc::smem_u32(smem_w) + kb * kWBlk, &tmap_w, w_bar, ks * num_kb * BK + kb * BK, nb * kBNm);
    }
    __syncwarp();
    const uint32_t full0 = tc::smem_u32(&ss->full[0]), empty0 = tc::smem_u32(&ss->empty[0]), a0 = tc::smem_u32(smem_a);
    const int nkb_all = kSplit * num_kb;
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    // one k-block: the operand is a contiguous 16 KB block = the 128B-swizzled K-major [128 x 64] tile image written by
    // the epilogues (no tensor map, no coordinates)
    const int wc0 = ks * num_kb * BK, wc1 = nb * kBNm;        // weight tensor-map coordinates of this CTA's slice
    uint32_t a_bytes = kABytes;                               // a partial batch tile only needs its first rows (8-row swizzle atoms)
    auto issue = [&](uint32_t st_, const __nv_bfloat16* src, int kb) {   // elected lane: fill ring stage st_ with k-block kb
      const uint32_t fb = full0 + 8 * st_;
      ss->kb_idx[st_] = (uint32_t)kb;                         // published by the release of the expect_tx arrive below
      if (p.debug_mode == 1) { tc::mbar_arrive(&ss->full[st_]); return; }
      tc::mbar_expect_tx_u32(fb, a_bytes + (kStream ? kWBlockBytes : 0));
      tc::bulk_load_1d_u32(a0 + st_ * kStageBytes, src + (size_t)kb * (BM * BK), a_bytes, fb);
      if (kStream) tc::tma_load_2d_u32(a0 + st_ * kStageBytes + kABytes, &tmap_w, fb, wc0 + kb * BK, wc1);
    };
    auto load_block = [&](const __nv_bfloat16* src, int kb) -> bool {
      if (!tc::mbar_try_wait_u32(empty0 + 8 * stage, phase ^ 1)) {
        if (!wait_bar<false>(&ss->empty[stage], phase ^ 1, abort_flag)) return false;
      }
      if (tc::elect_one()) issue(stage, src, kb);
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
      return true;
    };
    // two k-blocks per turn (both empty-barrier try_waits in flight together, two copies issued back to back)
    auto load_pair = [&](const __nv_bfloat16* src, int kb_a, int kb_b) -> bool {
      uint32_t s1 = stage + 1, ph1 = phase;
      if (s1 == kStages) { s1 = 0; ph1 ^= 1; }
      const bool r0 = tc::mbar_try_wait_u32(empty0 + 8 * stage, phase ^ 1);
      const bool r1 = tc::mbar_try_wait_u32(empty0 + 8 * s1, ph1 ^ 1);
      if (!r0 && !wait_bar<false>(&ss->empty[stage], phase ^ 1, abort_flag)) return false;
      if (!r1 && !wait_bar<false>(&ss->empty[s1], ph1 ^ 1, abort_flag)) return false;
      if (tc::elect_one()) {
        issue(stage, src, kb_a);
        issue(s1, src, kb_b);
      }
      __syncwarp();
      stage = s1 + 1; phase = ph1;
      if (stage == kStages) { stage = 0; phase ^= 1; }
      return true;
    };
    // Dataflow instead of a grid barrier: every operand k-block (64 columns of h_{t-1} / dG_{t+1}) has its own arrival
    // counter; the 32 lanes poll all of them at once and the blocks are pulled into the ring in the order in which their
    // producer CTAs finish, so the stream and the MMAs start under the stragglers' epilogues (accumulation order is free).
    const unsigned int per_step = kBwd ? 1u : 4u;             // arrivals per k-block and step (bwd: 1 CTA, fwd: 4 CTAs x 16 hidden)
    const uint64_t all_kb = num_kb >= 64 ? ~0ull : ((1ull << num_kb) - 1ull);
    for (int s = kBwd ? 1 : 0; s < steps && ok; ++s) {
      const int tsl = kBwd ? p.T - s : s;       // forward step s consumes h_seq[s]; backward iteration s consumes dG[T-s]
      for (int tile = 0; tile < kTiles && ok; ++tile) {
        const int mb = mb0 + tile;
        const int kb_base = ks * num_kb;
        const __nv_bfloat16* src = p.a_tiled + (((size_t)tsl * p.tiles_m + mb) * nkb_all + kb_base) * (BM * BK);
        const unsigned int* ctr = p.sync + kSyncKb + ((size_t)mb * nkb_all + kb_base) * 32;
        const unsigned int target = (unsigned)s * per_step;
        {
          const int rows = p.B - mb * BM;                     // rows beyond B are never read back from the accumulator
          a_bytes = rows >= BM ? kABytes : (uint32_t)(((rows + 7) / 8) * 8 * BK * 2);
          if (p.debug_mode == 4) a_bytes = kABytes / 2;       // experiment: half the operand traffic (results are garbage)
        }
        uint64_t pending = all_kb;
        const long long t0 = clock64();
        int spins = 0;
        bool stamped = false;
        while (pending && ok) {
          uint64_t ready = pending;
          if (s > 0) {
            if (p.sync_mode == 1) {
              const unsigned int v = p.poll_acquire ? ld_acquire_gpu(p.sync + mb) : ld_relaxed_gpu(p.sync + mb);
              ready = ((int)(v - (unsigned)s * (unsigned)p.tiles_n) >= 0) ? pending : 0ull;
            } else if (p.sync_mode == 2) {
              const unsigned int* fl = p.sync + kSyncFlags + (size_t)mb * p.tiles_n;
              if (kBwd) {                                       // k-block kb_base + lane is written by CTA kb_base + lane
                bool r0 = false;
                if (lane < num_kb) r0 = (int)(ld_relaxed_gpu(fl + kb_base + lane) - (unsigned)s) >= 0;
                ready = (uint64_t)__ballot_sync(0xffffffffu, r0);
              } else {                                          // k-block j is written by CTAs 4j..4j+3; lane L reads flags 2L, 2L+1
                bool r0 = false;
                if (2 * lane < p.tiles_n) {
                  const uint2 v2 = ld_relaxed_gpu_v2(fl + 2 * lane);
                  r0 = ((int)(v2.x - (unsigned)s) >= 0) && ((int)(v2.y - (unsigned)s) >= 0);
                }
                const unsigned int b = __ballot_sync(0xffffffffu, r0);
                unsigned int pr = b & (b >> 1) & 0x55555555u;   // bit 2j set <=> k-block j complete
                pr = (pr | (pr >> 1)) & 0x33333333u; pr = (pr | (pr >> 2)) & 0x0f0f0f0fu;
                pr = (pr | (pr >> 4)) & 0x00ff00ffu; pr = (pr | (pr >> 8)) & 0x0000ffffu;
                ready = pr >> kb_base;
              }
            } else {
              bool r0 = false, r1 = false;
              if (lane < num_kb) r0 = (int)(ld_relaxed_gpu(ctr + lane * 32) - target) >= 0;
              if (lane + 32 < num_kb) r1 = (int)(ld_relaxed_gpu(ctr + (lane + 32) * 32) - target) >= 0;
              ready = (uint64_t)__ballot_sync(0xffffffffu, r0) | ((uint64_t)__ballot_sync(0xffffffffu, r1) << 32);
            }
            ready &= pending;
            if (p.debug_mode == 3) {                            // experiment: in-order issue (lowest pending block first)
              const uint64_t low = pending & (~pending + 1ull);
              ready = (ready & low) ? low : 0ull;
            }
            if (!ready) {
              if ((++spins & 63) == 0) {
                if (*abort_flag) { ok = false; break; }
                if (clock64() - t0 > kSpinLimit) { *abort_flag = 1; ok = false; break; }
              }
              continue;
            }
            // The producers' release made the tile image visible at L2 before the counter moved, and the only consumer is the
            // async proxy (bulk copies read L2, issued after this control dependency): a generic-proxy acquire fence here
            // costs an L2 round trip per batch of blocks and buys nothing.  The warp barrier orders polling lanes before the
            // elected lane, the proxy fence orders generic observations before the async-proxy reads.
            __syncwarp();
            asm volatile("fence.proxy.async.global;" ::: "memory");
          }
          if (p.dbg && blockIdx.x == 0 && lane == 0 && tile == 0 && !stamped) { p.dbg[4 * s + 0] = gtime(); stamped = true; }
          pending &= ~ready;
          while (ready && ok) {
            const int ka = __ffsll((long long)ready) - 1;
            ready &= ready - 1ull;
            if (ready) {
              const int kb2 = __ffsll((long long)ready) - 1;
              ready &= ready - 1ull;
              ok = load_pair(src, ka, kb2);
            } else {
              ok = load_block(src, ka);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) atomicAdd(&ss->roles_done, 1);
  } else if (warp == 1) {
    // ======================================================================== MMA issuer
    constexpr uint32_t idesc = tc::make_idesc_bf16_f32(BM, kBNm);
    bool ok = kStream ? true : wait_bar<false>(&ss->w_full, 0, abort_flag);
    const uint32_t full0 = tc::smem_u32(&ss->full[0]), empty0 = tc::smem_u32(&ss->empty[0]);
    const uint32_t tfull0 = tc::smem_u32(&ss->tmem_full[0]);
    const uint64_t desc_a0 = tc::desc_kmajor_sw128(tc::smem_u32(smem_a));      // + stage * (kABytes >> 4)
    const uint64_t desc_w0 = tc::desc_kmajor_sw128(tc::smem_u32(smem_w));      // + kb * (kWBlockBytes >> 4)
    uint32_t stage = 0, phase = 0;
    const bool prof = p.dbg && blockIdx.x == 0;
    constexpr int kGroup = (kStages >= 5 && !kFSplit && !kBwd) ? 4 : 2;   // k-blocks per turn: all their try_waits are in flight together
    int steps_done = 0;
    for (int s = kBwd ? 1 : 0; s < steps && ok; ++s, ++steps_done) {
     long long t_wait = 0, t_begin = prof ? clock64() : 0, t_first = 0;
     for (int tile = 0; tile < kTiles && ok; ++tile) {
      const uint32_t acc = tmem_d + tile * kBNm;
      const uint32_t tfull = tfull0 + 8 * tile;
      // operand k-blocks of the next step can arrive (from faster CTAs) while this CTA's epilogue still reads the
      // accumulator of the previous one: the first MMA of a step overwrites it, so wait for the epilogue's hand-back
      if (steps_done > 0) ok = wait_bar<false>(&ss->tmem_empty[tile], (uint32_t)((steps_done - 1) & 1), abort_flag);
      if (!ok) break;
      for (int kb = 0; kb < num_kb && ok; kb += kGroup) {
        const int g = (num_kb - kb) < kGroup ? (num_kb - kb) : kGroup;
        uint32_t st[kGroup], ph[kGroup];
        bool rdy[kGroup];
        {
          uint32_t sx = stage, px = phase;
#pragma unroll
          for (int i = 0; i < kGroup; ++i) {
            st[i] = sx; ph[i] = px;
            if (++sx == kStages) { sx = 0; px ^= 1; }
          }
        }
        const long long tw0 = prof ? clock64() : 0;
#pragma unroll
        for (int i = 0; i < kGroup; ++i) rdy[i] = (i < g) ? tc::mbar_try_wait_u32(full0 + 8 * st[i], ph[i]) : true;
#pragma unroll
        for (int i = 0; i < kGroup; ++i)
          if (ok && !rdy[i]) ok = wait_bar<false>(&ss->full[st[i]], ph[i], abort_flag);
        if (!ok) break;
        if (prof) { const long long tw1 = clock64(); if (kb == 0 && tile == 0) t_first = tw1 - tw0; else t_wait += tw1 - tw0; }
        // all stages of the group have landed: issue its 4*g MMAs back to back (measured faster than issuing each
        // k-block as soon as its own stage lands: one fence + one elect per turn instead of per k-block)
        tc::fence_after_sync();
        uint32_t kbi[kGroup];                                   // stages carry k-blocks in arrival order
#pragma unroll
        for (int i = 0; i < kGroup; ++i)       // (grid-barrier mode streams in order: no shared-memory round trip)
          kbi[i] = (kStream || i >= g) ? 0u : (p.sync_mode == 1 ? (uint32_t)(kb + i) : ss->kb_idx[st[i]]);
        if (tc::elect_one()) {
#pragma unroll
          for (int i = 0; i < kGroup; ++i) {
            if (i < g) {
              if (p.debug_mode == 2) {
                tc::mbar_arrive(&ss->empty[st[i]]);
              } else {
                const uint64_t da = desc_a0 + (uint64_t)(st[i] * (kStageBytes >> 4));
                const uint64_t dbi = kStream ? da + (uint64_t)(kABytes >> 4) : desc_w0 + (uint64_t)(kbi[i] * (kWBlk >> 4));
                if (kb == 0 && i == 0) tc::mma_bf16_ss_first(acc, da, dbi, idesc); else tc::mma_bf16_ss_acc(acc, da, dbi, idesc);
                tc::mma_bf16_ss_acc(acc, da + 2, dbi + 2, idesc);
                tc::mma_bf16_ss_acc(acc, da + 4, dbi + 4, idesc);
                tc::mma_bf16_ss_acc(acc, da + 6, dbi + 6, idesc);
                tc::mma_commit_u32(empty0 + 8 * st[i]);
              }
            }
          }
          if (kb + g >= num_kb) { if (p.debug_mode == 2) tc::mbar_arrive(&ss->tmem_full[tile]); else tc::mma_commit_u32(tfull); }
        }
        __syncwarp();
        if (!ok) break;
#pragma unroll
        for (int i = 0; i < kGroup; ++i)
          if (i < g) { if (++stage == kStages) { stage = 0; phase ^= 1; } }
      }
     }
      if (prof && lane == 0) {           // [first-turn wait (incl. grid barrier + first loads), later waits, whole step] in cycles
        p.dbg[4 * s + 3] = (unsigned long long)t_first | ((unsigned long long)t_wait << 20) | ((unsigned long long)(clock64() - t_begin) << 40);
      }
    }
    __syncwarp();
    if (lane == 0) atomicAdd(&ss->roles_done, 1);
  } else if (warp == 3) {
    // ======================================================================== watchdog
    // A bounded spin that times out raises abort_flag and every role loop drains.  A thread already parked in a named
    // barrier (bar.sync cannot time out) would keep the grid alive forever: if the roles have not all left within the
    // grace period after an abort, kill the kernel (launch failure in the host process) instead of hanging the GPU.
    long long t_abort = 0;
    while (*reinterpret_cast<volatile int*>(&ss->roles_done) < 10) {
      __nanosleep(4000);
      if (*abort_flag) {
        if (t_abort == 0) t_abort = clock64();
        else if (clock64() - t_abort > kAbortGrace) {
          if (lane == 0) atomicExch(reinterpret_cast<int*>(p.sync + kSyncErr), 2);
          __threadfence_system();
          if (!p.no_trap) __trap();
          break;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ======================================================================== epilogue (8 warps, serves the tiles in turn)
    const int ewi = warp - kEpiWarp0;
    const int quarter = ewi & 3, half = ewi >> 2;
    const int rloc = quarter * 32 + lane;
    const int etid = ewi * 32 + lane;
    const int H = p.H, B = p.B;
    const uint32_t taddr0 = tmem_d + ((uint32_t)(quarter * 32) << 16) + 32 * half;
    uint32_t tphase = 0;
    bool ok = true;
    const bool dbg_thread = p.dbg && blockIdx.x == 0 && etid == 0;
    auto epi_bar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // exchange-buffer barriers: filled by st.async (async proxy, like a TMA load) / freed by relaxed arrives, so a CTA-scope
    // wait is enough; the cluster-scope acquire form (debug_mode 7, the earlier default) adds a CCTL.IVALL per wait
    const bool cluster_acquire = p.debug_mode == 7;
    auto xwait = [&](uint64_t* bar, uint32_t parity, volatile int* af) {
      return cluster_acquire ? wait_bar<true>(bar, parity, af) : wait_bar<false>(bar, parity, af);
    };
    // MEMBAR.ALL.GPU (the release of the dataflow signal) drains EVERY outstanding store of the SM, not just the signalling
    // thread's: if the other 255 threads start their 57 KB of bookkeeping stores (h_seq / c_seq / activations) meanwhile, the
    // signal - the only thing the other CTAs wait for - is held back by 1-3 us (measured).  They wait for it instead.
    auto signal_sent_bar = [&]() { asm volatile("bar.sync 2, 256;" ::: "memory"); };
    // grid-barrier arrive: the CTA barrier orders every epilogue thread's writes before this thread's release
    // (same pattern as cooperative-groups grid sync).  ONE gpu-scope release: each fence is a full L2 round trip
    // (~0.8 us) and three of them used to dominate the epilogue; the generic->async proxy fence is on the consumer side.

    if (!kBwd) {
      const int j0 = in_mb * 16 + 8 * half;         // this thread's 8 hidden units
      const int n0 = in_mb * 64 + 32 * half;        // = its 32 gate columns
      const float* bs = ss->bias + 32 * half;
      uint32_t xphase = 0;
      const uint32_t xbase = tc::smem_u32(smem_x);
      const uint32_t xbar = tc::smem_u32(&ss->xchg_full[0]), fbar = tc::smem_u32(&ss->xchg_free[0]);
      const uint32_t pbar = kFSplit ? mapa(xbar, (uint32_t)(1 - ks)) : 0u;      // the peer's exchange barrier
      float cst[kTiles][8];
#pragma unroll
      for (int tile = 0; tile < kTiles; ++tile) {
        const int row = (mb0 + tile) * BM + rloc;
#pragma unroll
        for (int i = 0; i < 8; ++i) cst[tile][i] = row < B ? p.c_seq[(size_t)row * H + j0 + i] : 0.f;     // c_0 (written by the prologue)
      }
      for (int t = 0; t < p.T && ok; ++t) {
#pragma unroll
        for (int tile = 0; tile < kTiles; ++tile) {
          const int mb = mb0 + tile;
          const int row = mb * BM + rloc;
          const bool valid = row < B;
          // operands that do not depend on the GEMM: issue their loads before waiting on the accumulator
          U8 gxw[2];
          if (p.in_gate != nullptr) {                 // wavefront: gx[t] is produced while we run (this warp's rows: one 128-row block)
            const size_t gr = (size_t)t * B + (size_t)mb * BM;
            ok = wait_in_gate(p.in_gate + (gr >> 7) * p.in_gate_tiles_n + (n0 >> 8), lane, abort_flag);
            if (!ok) break;
            if (valid) {
              const __nv_bfloat16* gp = p.gx + ((size_t)t * B + row) * (4 * H) + n0;
              gxw[0] = ldg_cg32(gp); gxw[1] = ldg_cg32(gp + 16);
            }
          } else if (valid) {
            const __nv_bfloat16* gp = p.gx + ((size_t)t * B + row) * (4 * H) + n0;
            gxw[0] = ldg_nc32(gp); gxw[1] = ldg_nc32(gp + 16);
            if (t + 2 < p.T && p.debug_mode != 6) prefetch_l2(gp + (size_t)2 * B * (4 * H));     // the x-projection comes from HBM: pull it into L2 early
          }
          ok = wait_bar<false>(&ss->tmem_full[tile], tphase, abort_flag);
          if (!ok) break;
          tc::fence_after_sync();
          unsigned long long t_acc = 0;
          if (p.dbg && t == 8 && etid == 0) t_acc = gtime();
          if (dbg_thread && tile == 0) p.dbg[4 * t + 1] = gtime();
          uint32_t v[32];
          if constexpr (kFSplit) {
            // partial sums over this member's K half: columns [64 ks, +64) are mine, the other 64 go to the peer (bf16, DSMEM)
            const uint32_t tb = tmem_d + ((uint32_t)(quarter * 32) << 16) + 32 * half;
            uint32_t u[32];
            tc::tmem_ld32(tb + 64 * (1 - ks), u);
            tc::tmem_ld32(tb + 64 * ks, v);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&ss->tmem_empty[tile]);     // the issuer may overwrite the accumulator
            if (t > 0) {                                               // the peer has consumed last step's partial
              ok = xwait(&ss->xchg_free[0], (uint32_t)((t - 1) & 1), abort_flag);
              if (!ok) break;
            }
            const uint32_t drow = mapa(xbase + (uint32_t)(rloc * 128), (uint32_t)(1 - ks));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t pk[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) pk[k] = pack_bf2(__uint_as_float(u[8 * i + 2 * k]), __uint_as_float(u[8 * i + 2 * k + 1]));
              st_async_u4(drow + (uint32_t)((((4 * half + i) ^ (rloc & 7))) * 16), make_uint4(pk[0], pk[1], pk[2], pk[3]), pbar);
            }
            ok = xwait(&ss->xchg_full[0], xphase, abort_flag);
            if (!ok) break;
            xphase ^= 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 x4 = *reinterpret_cast<const uint4*>(smem_x + (size_t)rloc * 128 + (((4 * half + i) ^ (rloc & 7)) * 16));
              const uint32_t w[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[8 * i + 2 * k] = __float_as_uint(__uint_as_float(v[8 * i + 2 * k]) + bf_lo(w[k]));
                v[8 * i + 2 * k + 1] = __float_as_uint(__uint_as_float(v[8 * i + 2 * k + 1]) + bf_hi(w[k]));
              }
            }
          } else {
            tc::tmem_ld32(taddr0 + tile * BN, v);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&ss->tmem_empty[tile]);     // the issuer may overwrite the accumulator
          }
          if (dbg_thread && t == 8 && tile == 0) p.dbg[4 * (p.T + 2) + 0] = gtime();
          float cn[8], hv[8];
          uint32_t apk[16];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint32_t ga = gxw[jj >> 2].v[2 * (jj & 3)], gb = gxw[jj >> 2].v[2 * (jj & 3) + 1];
            const float pi = __uint_as_float(v[4 * jj + 0]) + bf_lo(ga) + bs[4 * jj + 0];
            const float pf = __uint_as_float(v[4 * jj + 1]) + bf_hi(ga) + bs[4 * jj + 1];
            const float pg = __uint_as_float(v[4 * jj + 2]) + bf_lo(gb) + bs[4 * jj + 2];
            const float po = __uint_as_float(v[4 * jj + 3]) + bf_hi(gb) + bs[4 * jj + 3];
            const float ig = ts::sigmoidf_fast(pi), fg = ts::sigmoidf_fast(pf), gg = ts::tanhf_fast(pg), og = ts::sigmoidf_fast(po);
            const float c = fg * cst[tile][jj] + ig * gg;
            cst[tile][jj] = c;                        // the cell state never leaves the registers of its thread
            cn[jj] = c;
            hv[jj] = og * ts::tanhf_fast(c);
            apk[2 * jj] = pack_bf2(ig, fg);
            apk[2 * jj + 1] = pack_bf2(gg, og);
          }
          const uint4 h8 = make_uint4(pack_bf2(hv[0], hv[1]), pack_bf2(hv[2], hv[3]), pack_bf2(hv[4], hv[5]), pack_bf2(hv[6], hv[7]));
          {
            // next step's operand first (the only thing other CTAs wait for): 8 values = one 16 B chunk of the swizzled
            // tile image; chunk c of row r sits at position c ^ (r & 7)
            const size_t blk = ((size_t)(t + 1) * p.tiles_m + mb) * (H / BK) + (j0 / BK);
            const int chunk = ((j0 % BK) / 8) ^ (rloc & 7);
            stg16(p.a_tiled + blk * (BM * BK) + rloc * BK + chunk * 8, h8);
          }
          if (dbg_thread && t == 8 && tile == 0) p.dbg[4 * (p.T + 2) + 1] = gtime();
          epi_bar();
          if (etid == 0) {
            if (dbg_thread && t == 8 && tile == 0) p.dbg[4 * (p.T + 2) + 2] = gtime();
            if (p.dbg && t == 8 && tile == 0) {                // per-CTA stamps (skew study): accumulator ready / about to signal
              p.dbg[4 * (p.T + 2) + 64 + 2 * blockIdx.x] = t_acc;
              p.dbg[4 * (p.T + 2) + 64 + 2 * blockIdx.x + 1] = gtime();
            }
            // this CTA's 16 hidden units = a quarter of k-block nb/4
            if (p.sync_mode == 1) signal_counter(p.sync + mb);
            else if (p.sync_mode == 2) st_release_gpu(p.sync + kSyncFlags + (size_t)mb * p.tiles_n + in_mb, (unsigned)(t + 1));
            else signal_counter(p.sync + kSyncKb + ((size_t)mb * (H / BK) + (in_mb >> 2)) * 32);
            if (dbg_thread && tile == 0) p.dbg[4 * t + 2] = gtime();
          }
          // (after the CTA barrier every epilogue thread is done with this step's exchange buffer)
          if (kFSplit && etid == 32) {
            tc::mbar_expect_tx_u32(xbar, (uint32_t)kXchgBytes);        // arm the next phase, THEN let the peer overwrite the buffer
            mbar_arrive_remote_relaxed(mapa(fbar, (uint32_t)(1 - ks)));
          }
          signal_sent_bar();
          if (valid && p.debug_mode != 5) {              // everything below is off the critical path
            stg16(p.h_seq + ((size_t)(t + 1) * B + row) * H + j0, h8);
            float* cp = p.c_seq + ((size_t)(t + 1) * B + row) * H + j0;
            stg32(cp, __float_as_uint(cn[0]), __float_as_uint(cn[1]), __float_as_uint(cn[2]), __float_as_uint(cn[3]),
                  __float_as_uint(cn[4]), __float_as_uint(cn[5]), __float_as_uint(cn[6]), __float_as_uint(cn[7]));
            __nv_bfloat16* ap = p.act + ((size_t)t * B + row) * (4 * H) + n0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              stg32(ap + 16 * i, apk[8 * i], apk[8 * i + 1], apk[8 * i + 2], apk[8 * i + 3], apk[8 * i + 4], apk[8 * i + 5], apk[8 * i + 6], apk[8 * i + 7]);
          }
        }
        tphase ^= 1;
      }
      if (p.extra_signal && ok) {                 // the last step's h_seq rows (natural layout) are visible to the gated GEMM
#pragma unroll
        for (int tile = 0; tile < kTiles; ++tile) {
          epi_bar();
          if (etid == 0) {
            if (p.sync_mode == 1) signal_counter(p.sync + mb0 + tile);
            else signal_counter(p.sync + kSyncKb + ((size_t)(mb0 + tile) * (H / BK) + (in_mb >> 2)) * 32);
          }
        }
      }
    } else {
      // after the reduce-scatter this cluster member owns hidden [64 nb + 16 ks, +16); this thread 8 of them
      const int j0 = nb * 64 + ks * 16 + 8 * half;
      uint32_t xphase = 0;
      float dc[kTiles][8], dh[kTiles][8];
#pragma unroll
      for (int tile = 0; tile < kTiles; ++tile) {
        const int row = (mb0 + tile) * BM + rloc;
        if (row < B) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float4 a = *reinterpret_cast<const float4*>(p.dc0 + (size_t)row * H + j0 + 4 * i);
            float4 b = *reinterpret_cast<const float4*>(p.dh0 + (size_t)row * H + j0 + 4 * i);
            dc[tile][4 * i] = a.x; dc[tile][4 * i + 1] = a.y; dc[tile][4 * i + 2] = a.z; dc[tile][4 * i + 3] = a.w;
            dh[tile][4 * i] = b.x; dh[tile][4 * i + 1] = b.y; dh[tile][4 * i + 2] = b.z; dh[tile][4 * i + 3] = b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { dc[tile][i] = 0.f; dh[tile][i] = 0.f; }
        }
      }
      U8 c_carry[kTiles];                       // c_t of the previous iteration = c_{t+1} of this one (one load per step, not two)
      for (int s = 0; s <= p.T && ok; ++s) {
        const int t = p.T - 1 - s;
#pragma unroll
        for (int tile = 0; tile < kTiles; ++tile) {
          const int mb = mb0 + tile;
          const int row = mb * BM + rloc;
          const bool valid = row < B;
          uint8_t* xbuf = smem_x + tile * kXchgBytes;               // [4 src][128 rows][16 bf16]
          const uint32_t xbase = tc::smem_u32(xbuf);
          const uint32_t xbar = tc::smem_u32(&ss->xchg_full[tile]);
          U8 avw[2], cpv, cnv;
          uint4 dhv;
          if (p.in_gate != nullptr && s < p.T) {      // wavefront: dh_seq[t] (= dX of the layer above) is produced while we run
            const size_t gr = (size_t)t * B + (size_t)mb * BM;
            ok = wait_in_gate(p.in_gate + (gr >> 7) * p.in_gate_tiles_n + (j0 >> 8), lane, abort_flag);
            if (!ok) break;
          }
          if (valid && s < p.T) {
            const __nv_bfloat16* ap = p.act + ((size_t)t * B + row) * (4 * H) + 4 * j0;
            avw[0] = ldg_nc32(ap); avw[1] = ldg_nc32(ap + 16);
            dhv = p.dh_seq ? (p.in_gate ? ldg_cg16(p.dh_seq + ((size_t)t * B + row) * H + j0) : ldg_nc16(p.dh_seq + ((size_t)t * B + row) * H + j0))
                           : make_uint4(0u, 0u, 0u, 0u);
            const float* c0p = p.c_seq + ((size_t)t * B + row) * H + j0;
            const float* c1p = p.c_seq + ((size_t)(t + 1) * B + row) * H + j0;
            cpv = ldg_nc32(c0p);
            cnv = (s == 0) ? ldg_nc32(c1p) : c_carry[tile];
            c_carry[tile] = cpv;
            if (t >= 2) {                                                      // saved activations come from HBM: pull t-2 into L2 early
              prefetch_l2(ap - (size_t)2 * B * (4 * H));
              prefetch_l2(c0p - (size_t)2 * B * H);
              if (p.dh_seq && !p.in_gate) prefetch_l2(p.dh_seq + ((size_t)(t - 2) * B + row) * H + j0);
            }
          }
          if (s > 0) {
            ok = wait_bar<false>(&ss->tmem_full[tile], tphase, abort_flag);
            if (!ok) break;
            tc::fence_after_sync();
            if (dbg_thread && tile == 0) p.dbg[4 * s + 1] = gtime();
            uint32_t v[32];
            tc::tmem_ld32(taddr0 + tile * BN, v);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&ss->tmem_empty[tile]);     // the issuer may overwrite the accumulator
            // A member only needs the dG blocks of ITS K-quarter, so nothing in the dataflow stops a fast member from being a
            // whole step ahead of a slow one: explicit back-pressure before overwriting anybody's exchange buffer.
            if (s > 1) {
              ok = xwait(&ss->xchg_free[tile], (uint32_t)(s & 1), abort_flag);
              if (!ok) break;
            }
            // reduce-scatter over the 4 K-quarters: column chunk q (16 wide, bf16) goes to member q's slot [ks] (DSMEM)
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              const uint32_t dst = mapa(xbase + (uint32_t)((ks * BM + rloc) * 32), (uint32_t)(2 * half + qq));
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[i] = pack_bf2(__uint_as_float(v[16 * qq + 2 * i]), __uint_as_float(v[16 * qq + 2 * i + 1]));
              const uint32_t dbar = mapa(xbar, (uint32_t)(2 * half + qq));
              st_async_u4(dst, make_uint4(pk[0], pk[1], pk[2], pk[3]), dbar);
              st_async_u4(dst + 16, make_uint4(pk[4], pk[5], pk[6], pk[7]), dbar);
            }
            ok = xwait(&ss->xchg_full[tile], xphase, abort_flag);
            if (!ok) break;
#pragma unroll
            for (int i = 0; i < 8; ++i) dh[tile][i] = 0.f;
#pragma unroll
            for (int src = 0; src < 4; ++src) {
              const uint4 x4 = *reinterpret_cast<const uint4*>(xbuf + (size_t)(src * BM + rloc) * 32 + 16 * half);
              const uint32_t w[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) { dh[tile][2 * i] += bf_lo(w[i]); dh[tile][2 * i + 1] += bf_hi(w[i]); }
            }
          }
          if (s == p.T) {
            if (valid) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                *reinterpret_cast<float4*>(p.dh0 + (size_t)row * H + j0 + 4 * i) = make_float4(dh[tile][4 * i], dh[tile][4 * i + 1], dh[tile][4 * i + 2], dh[tile][4 * i + 3]);
                *reinterpret_cast<float4*>(p.dc0 + (size_t)row * H + j0 + 4 * i) = make_float4(dc[tile][4 * i], dc[tile][4 * i + 1], dc[tile][4 * i + 2], dc[tile][4 * i + 3]);
              }
            }
            if (p.extra_signal) {                   // dpre[0] (natural layout, written after the last per-step signal) is visible to the gated GEMM
              epi_bar();
              if (etid == 0) {
                if (p.sync_mode == 1) signal_counter(p.sync + mb);
                else signal_counter(p.sync + kSyncKb + ((size_t)mb * (4 * H / BK) + in_mb) * 32);
              }
            }
            continue;
          }
          uint32_t gpk[16];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint32_t aa = avw[jj >> 2].v[2 * (jj & 3)], ab = avw[jj >> 2].v[2 * (jj & 3) + 1];
            const float ig = bf_lo(aa), fg = bf_hi(aa), gg = bf_lo(ab), og = bf_hi(ab);
            const uint32_t dw = (jj >> 1) == 0 ? dhv.x : (jj >> 1) == 1 ? dhv.y : (jj >> 1) == 2 ? dhv.z : dhv.w;
            const float dht = dh[tile][jj] + ((jj & 1) ? bf_hi(dw) : bf_lo(dw));
            const float cprev = __uint_as_float(cpv.v[jj]);
            const float tcn = ts::tanhf_fast(__uint_as_float(cnv.v[jj]));
            const float dct = dc[tile][jj] + dht * og * (1.f - tcn * tcn);
            const float d_o = dht * tcn, d_i = dct * gg, d_f = dct * cprev, d_g = dct * ig;
            dc[tile][jj] = dct * fg;
            gpk[2 * jj] = pack_bf2(d_i * ig * (1.f - ig), d_f * fg * (1.f - fg));
            gpk[2 * jj + 1] = pack_bf2(d_g * (1.f - gg * gg), d_o * og * (1.f - og));
          }
          {
            // next iteration's operand first: this thread's 32 gate columns = 4 chunks of row rloc of k-block j0/16
            const size_t blk = ((size_t)t * p.tiles_m + mb) * (4 * H / BK) + (j0 / 16);
            __nv_bfloat16* tp = p.a_tiled + blk * (BM * BK) + rloc * BK;
            // chunk c of row r sits at c ^ (r & 7): an aligned chunk pair stays an aligned pair (32 B sector p ^ ((r & 7) >> 1)),
            // swapped when r is odd -> two whole-sector 256-bit stores instead of four 16 B ones
            const bool swp = rloc & 1;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int sp = (2 * half + k) ^ ((rloc & 7) >> 1);
              const uint32_t* a = gpk + 8 * k;
              stg32(tp + sp * 16, swp ? a[4] : a[0], swp ? a[5] : a[1], swp ? a[6] : a[2], swp ? a[7] : a[3],
                    swp ? a[0] : a[4], swp ? a[1] : a[5], swp ? a[2] : a[6], swp ? a[3] : a[7]);
            }
          }
          epi_bar();
          if (etid == 0) {
            // this CTA's 64 gate columns = dG k-block 4 nb + ks
            if (p.sync_mode == 1) signal_counter(p.sync + mb);
            else if (p.sync_mode == 2) st_release_gpu(p.sync + kSyncFlags + (size_t)mb * p.tiles_n + in_mb, (unsigned)(s + 1));
            else signal_counter(p.sync + kSyncKb + ((size_t)mb * (4 * H / BK) + in_mb) * 32);
            if (dbg_thread && tile == 0) p.dbg[4 * s + 2] = gtime();
          }
          // (the CTA barrier above also means every epilogue thread is done reading this step's exchange buffer)
          if (s > 0) {
            if (etid == 32) tc::mbar_expect_tx_u32(xbar, (uint32_t)kXchgBytes);    // arm the next phase, THEN free the buffer
            __syncwarp();
            if (etid >= 32 && etid < 36) mbar_arrive_remote_relaxed(mapa(tc::smem_u32(&ss->xchg_free[tile]), (uint32_t)(etid - 32)));
          }
          signal_sent_bar();
          if (valid) {                                   // the [T,B,4H] copy for the weight-gradient GEMMs: off the critical path
            __nv_bfloat16* gp = p.dpre + ((size_t)t * B + row) * (4 * H) + 4 * j0;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              stg32(gp + 16 * i, gpk[8 * i], gpk[8 * i + 1], gpk[8 * i + 2], gpk[8 * i + 3], gpk[8 * i + 4], gpk[8 * i + 5], gpk[8 * i + 6], gpk[8 * i + 7]);
          }
        }
        if (s > 0) { tphase ^= 1; xphase ^= 1; }
      }
    }
  }

  if (warp >= kEpiWarp0) {
    __syncwarp();
    if (lane == 0) atomicAdd(&ss->roles_done, 1);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (kCluster) cluster_sync_all();              // nobody exits while a peer may still write into / arrive on its smem
  if (p.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && ss->abort_flag) atomicExch(reinterpret_cast<int*>(p.sync + kSyncErr), 1);
  if (warp == 2) tc::tmem_dealloc(tmem_d, kTmemCols);
}

// One small launch instead of ~12 framework ops: h_seq[0] <- h0, c_seq[0] <- c0, the swizzled tile image of h0 (slot 0
// of the streamed operand, zero rows beyond B), and the step counters <- 0.
__global__ void seq_prologue_kernel(const __nv_bfloat16* __restrict__ h0, const float* __restrict__ c0,
                                    __nv_bfloat16* __restrict__ h_seq0, float* __restrict__ c_seq0,
                                    __nv_bfloat16* __restrict__ tiled0, unsigned int* __restrict__ sync, int B, int H, int tiles_m) {
  const int nchunk = H / 8, nkb = H / BK;
  const int total = tiles_m * BM * nchunk;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / nchunk, c = i % nchunk;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < B) {
      v = *reinterpret_cast<const uint4*>(h0 + (size_t)r * H + 8 * c);
      *reinterpret_cast<uint4*>(h_seq0 + (size_t)r * H + 8 * c) = v;
    }
    const int mb = r / BM, rloc = r % BM, kb = c / 8, pos = (c % 8) ^ (rloc & 7);
    *reinterpret_cast<uint4*>(tiled0 + (((size_t)mb * nkb + kb) * BM + rloc) * BK + pos * 8) = v;
  }
  const int n4 = B * H / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(c_seq0)[i] = reinterpret_cast<const float4*>(c0)[i];
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < kSyncErr; i += blockDim.x) sync[i] = 0u;
}

size_t smem_bytes(int H, bool bwd, int stages, int tiles, bool stream = false, bool fsplit = false) {
  const size_t ring = (size_t)stages * (stream ? kABytes + kWBlockBytes : kABytes);
  // resident weights: H/64 blocks of [64 x 64] (forward K-split: H/128 blocks of [128 x 64] = the same bytes)
  return (stream ? 0 : (size_t)(H / BK) * kWBlockBytes) + ring + ((bwd || fsplit) ? tiles * kXchgBytes : 0) + sizeof(SeqSmem) + 1024;
}

template <bool kBwd, int kStages, int kTiles, bool kStream = false, bool kFSplit = false>
int launch_cfg(const CUtensorMap& tw, const SeqParams& p, int grid, cudaStream_t st) {
  auto kern = lstm_seq_kernel<kBwd, kStages, kTiles, kStream, kFSplit>;
  const size_t smem = smem_bytes(p.H, kBwd, kStages, kTiles, kStream, kFSplit);
  constexpr int kClusterDim = kBwd ? 4 : (kFSplit ? 2 : 1);
  if (smem > 227 * 1024) return -4;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClusterDim; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (kClusterDim > 1) {
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return -20; }
    if (nclusters * kClusterDim < grid) return -21;    // not co-resident
  }
  if (p.pdl_wait) {                                    // start as soon as the previous kernel of the stream is fully resident
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  e = cudaLaunchKernelEx(&cfg, kern, tw, p);
  return (int)e;
}

template <bool kBwd>
int dispatch(const CUtensorMap& tw, const SeqParams& p, int grid, int stages, int tiles, bool stream, bool fsplit, cudaStream_t st) {
  if constexpr (!kBwd) {
    if (fsplit) {                                  // forward K-split (cluster of 2), one batch tile per CTA
      switch (stages) {
        case 3: return launch_cfg<false, 3, 1, false, true>(tw, p, grid, st);
        case 4: return launch_cfg<false, 4, 1, false, true>(tw, p, grid, st);
        case 5: return launch_cfg<false, 5, 1, false, true>(tw, p, grid, st);
        case 6: return launch_cfg<false, 6, 1, false, true>(tw, p, grid, st);
      }
      return -5;
    }
  }
  if (stream) {                                    // streamed weights: 24 KB stages, one batch tile per CTA
    switch (stages) {
      case 4: return launch_cfg<kBwd, 4, 1, true>(tw, p, grid, st);
      case 6: return launch_cfg<kBwd, 6, 1, true>(tw, p, grid, st);
      case 8: return launch_cfg<kBwd, 8, 1, true>(tw, p, grid, st);
    }
    return -5;
  }
  if (tiles == 2) {
    switch (stages) {
      case 2: return launch_cfg<kBwd, 2, 2>(tw, p, grid, st);
      case 3: return launch_cfg<kBwd, 3, 2>(tw, p, grid, st);
      case 4: return launch_cfg<kBwd, 4, 2>(tw, p, grid, st);
      case 5: return launch_cfg<kBwd, 5, 2>(tw, p, grid, st);
      case 6: return launch_cfg<kBwd, 6, 2>(tw, p, grid, st);
    }
  } else {
    switch (stages) {
      case 2: return launch_cfg<kBwd, 2, 1>(tw, p, grid, st);
      case 3: return launch_cfg<kBwd, 3, 1>(tw, p, grid, st);
      case 4: return launch_cfg<kBwd, 4, 1>(tw, p, grid, st);
      case 5: return launch_cfg<kBwd, 5, 1>(tw, p, grid, st);
      case 6: return launch_cfg<kBwd, 6, 1>(tw, p, grid, st);
    }
  }
  return -5;
}

int pick_stages(int H, bool bwd, int tiles) {
  for (int s = 6; s >= 2; --s)
    if (smem_bytes(H, bwd, s, tiles) <= 227 * 1024) return s;
  return 0;
}

}  // namespace

// sync_ws: kSyncWords u32 (layout above); everything but the sticky error flag in the last word is zeroed before every launch.
// variant (tuning knob, 0 = defaults) = tiles_per_cta + 16*stages + 4096*debug_mode:  tiles_per_cta 0 -> 1 (set 2 to let a
// CTA alternate two batch tiles);  stages 0 -> deepest ring that fits next to the resident weight slice.
template <bool kBwd>
static int seq_common(SeqParams& p, const void* w_base, int variant, cudaStream_t st) {
  const int H = p.H, B = p.B;
  if (H % 64 != 0) { ts::set_last_error("lstm_seq: H must be a multiple of 64"); return -2; }
  const int tiles_m = (B + BM - 1) / BM, tiles_n = kBwd ? (H / BN) * 4 : 4 * H / BN;
  int stages = (variant >> 4) & 15, tiles = variant & 15;
  p.debug_mode = (variant >> 12) & 7;
  p.sync_mode = (variant >> 16) & 3;
  p.poll_acquire = (variant >> 18) & 1;
  p.no_trap = (variant >> 20) & 1;
  if (tiles != 2 || tiles_m % 2 != 0) tiles = 1;
  // resident weight slice if it fits next to >= 4 ring stages, else stream the weights through the ring
  const bool stream = smem_bytes(H, kBwd, 4, 1) > 227 * 1024 || ((variant >> 8) & 1);
  if (stream) tiles = 1;
  // forward: 2-way K split (cluster of 2) unless disabled by variant bit 19
  const bool fsplit = !kBwd && !stream && tiles == 1 && H % 128 == 0 && !((variant >> 19) & 1) &&
                      smem_bytes(H, false, 3, 1, false, true) <= 227 * 1024;
  int dev = 0;
  cudaGetDevice(&dev);
  const int grid = (tiles_m / tiles) * tiles_n;
  if (grid > ts::sm_count(dev)) { ts::set_last_error("lstm_seq: grid exceeds SM count (not co-resident)"); return -3; }
  if (stream) {
    if (stages != 4 && stages != 6 && stages != 8) stages = smem_bytes(H, kBwd, 8, 1, true) <= 227 * 1024 ? 8 : 6;
  } else if (fsplit) {
    if (stages < 3 || stages > 6 || smem_bytes(H, false, stages, 1, false, true) > 227 * 1024) {
      stages = 6;
      while (stages > 3 && smem_bytes(H, false, stages, 1, false, true) > 227 * 1024) --stages;
    }
  } else {
    if (stages == 0) stages = pick_stages(H, kBwd, tiles);
    if (stages < 2 || smem_bytes(H, kBwd, stages, tiles) > 227 * 1024) { ts::set_last_error("lstm_seq: weight slice does not fit in shared memory"); return -4; }
  }
  const int K = kBwd ? 4 * H : H, N = kBwd ? H : 4 * H;
  CUtensorMap tw;
  if (int rc = ts::make_tmap_2d_bf16(&tw, w_base, (uint64_t)N, (uint64_t)K, (uint64_t)K, BK, fsplit ? 2 * BN : BN)) return rc;
  p.tiles_n = tiles_n;
  p.tiles_m = tiles_m;
  int rc = dispatch<kBwd>(tw, p, grid, stages, tiles, stream, fsplit, st);
  if (rc == -21) ts::set_last_error("lstm_seq: the thread-block clusters are not co-resident on this device");
  return rc;
}

extern "C" int ts_lstm_seq_fwd(const void* gx, const void* w_h, const float* bias, const void* h_seq, const float* c_seq,
                               void* act, const float* c0, void* dbg, void* a_tiled, int T, int B, int H, unsigned int* sync_ws, int variant,
                               cudaStream_t st, const void* h0, const unsigned int* in_gate, int in_gate_tiles_n, int extra_signal, int launch_flags) {
  if (!(launch_flags & 1)) {           // bit 0: the caller has run ts_lstm_seq_prologue itself (wavefront: all prologues precede the chain)
    const int tiles_m = (B + BM - 1) / BM;
    const int total = tiles_m * BM * (H / 8);
    int blocks = (total + 255) / 256;
    if (blocks > 592) blocks = 592;
    seq_prologue_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)h0, c0, (__nv_bfloat16*)h_seq, (float*)c_seq,
                                                (__nv_bfloat16*)a_tiled, sync_ws, B, H, tiles_m);
  }
  SeqParams p{};
  p.a_tiled = (__nv_bfloat16*)a_tiled;
  p.gx = (const __nv_bfloat16*)gx; p.bias = bias; p.h_seq = (__nv_bfloat16*)h_seq; p.c_seq = (float*)c_seq;
  p.act = (__nv_bfloat16*)act; p.sync = sync_ws; p.T = T; p.B = B; p.H = H; p.dbg = (unsigned long long*)dbg;
  p.in_gate = in_gate; p.in_gate_tiles_n = in_gate_tiles_n; p.extra_signal = extra_signal;
  p.pdl_wait = (launch_flags >> 1) & 1;
  return seq_common<false>(p, w_h, variant, st);
}

extern "C" int ts_lstm_seq_prologue(const void* h0, const float* c0, void* h_seq, float* c_seq, void* a_tiled, unsigned int* sync_ws,
                                    int B, int H, cudaStream_t st) {
  const int tiles_m = (B + BM - 1) / BM;
  const int total = tiles_m * BM * (H / 8);
  int blocks = (total + 255) / 256;
  if (blocks > 592) blocks = 592;
  seq_prologue_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)h0, c0, (__nv_bfloat16*)h_seq, c_seq, (__nv_bfloat16*)a_tiled, sync_ws, B, H, tiles_m);
  return (int)cudaGetLastError();
}

extern "C" int ts_lstm_seq_bwd(const void* dh_seq, const void* w_hT, const void* act, const float* c_seq, const void* dpre,
                               float* dh0, float* dc0, void* dbg, void* a_tiled, int T, int B, int H, unsigned int* sync_ws, int variant,
                               cudaStream_t st, const unsigned int* in_gate, int in_gate_tiles_n, int extra_signal, int launch_flags) {
  if (!(launch_flags & 1)) cudaMemsetAsync(sync_ws, 0, kSyncErr * sizeof(unsigned int), st);      // arrival counters restart at 0 every launch
  SeqParams p{};
  p.a_tiled = (__nv_bfloat16*)a_tiled;
  p.dh_seq = (const __nv_bfloat16*)dh_seq; p.act = (__nv_bfloat16*)act; p.c_seq = (float*)c_seq; p.dpre = (__nv_bfloat16*)dpre;
  p.dh0 = dh0; p.dc0 = dc0; p.sync = sync_ws; p.T = T; p.B = B; p.H = H; p.dbg = (unsigned long long*)dbg;
  p.in_gate = in_gate; p.in_gate_tiles_n = in_gate_tiles_n; p.extra_signal = extra_signal;
  p.pdl_wait = (launch_flags >> 1) & 1;
  return seq_common<true>(p, w_hT, variant, st);
}


// Feasibility probe: how many clusters of `cluster` CTAs of the backward kernel (1 CTA/SM, ~226 KB smem) can be co-resident.
// Measured on B200: 1 -> 148, 2 -> 74, 4 -> 33, 8 -> 15, 16 -> 7: an 8-way backward K-split (N = 128, 8 k-blocks per step,
// the mirror of the forward 2-way split) would need 16 clusters of 8 at B = 256, H = 1024 and is therefore not launchable.
extern "C" int ts_lstm_seq_cluster_probe(int cluster) {
  auto kern = lstm_seq_kernel<true, 4, 1, false, false>;
  const size_t smem = 226 * 1024;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (cluster > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return -2; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(128 / cluster * cluster); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return -3; }
  return n;
}
