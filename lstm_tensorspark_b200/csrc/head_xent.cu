// K-HEAD (generic shapes): dense head + sparse softmax cross-entropy + accuracy in ONE launch.
// Replaces reshape -> matmul -> +bias -> sparse_softmax_cross_entropy_with_logits -> reduce_mean ->
// argmax/equal/cast/reduce_mean (reference: /root/reference/src/rnn.py:214-221,55-63,84-92; K9-K11 in SURVEY §2.5).
// One warp per batch row; C <= 32 classes kept in registers.  Emits logits, sum of NLL, dlogits =
// (softmax - onehot)/B (ready for the backward GEMMs) and the number of correct rows.
#include "ts_common.cuh"

namespace {

constexpr int kMaxC = 32;

template <typename T>
__global__ void head_xent_kernel(const T* __restrict__ h, const float* __restrict__ W, const float* __restrict__ bias,
                                 const long long* __restrict__ labels, float* __restrict__ logits,
                                 float* __restrict__ dlogits, float* __restrict__ loss_sum, int* __restrict__ correct,
                                 int B, int H, int C) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B) return;
  float acc[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) acc[c] = 0.f;
  const T* hr = h + (size_t)warp * H;
  for (int k = lane; k < H; k += 32) {
    float hv = ts::Cvt<T>::to_f(hr[k]);
    const float* w = W + (size_t)k * C;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) acc[c] = fmaf(hv, w[c], acc[c]);
  }
  float mx = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    if (c < C) {
      acc[c] = ts::warp_sum(acc[c]) + bias[c];
      if (acc[c] > mx) { mx = acc[c]; arg = c; }
    }
  }
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c)
    if (c < C) se += expf(acc[c] - mx);
  float lse = mx + logf(se);
  int y = (int)labels[warp];
  float invB = 1.0f / (float)B;
  float nll = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    if (c < C) {
      float p = expf(acc[c] - lse);
      if (c == y) nll = lse - acc[c];
      if (lane == 0) {
        logits[(size_t)warp * C + c] = acc[c];
        dlogits[(size_t)warp * C + c] = (p - (c == y ? 1.f : 0.f)) * invB;
      }
    }
  }
  if (lane == 0) {
    atomicAdd(loss_sum, nll);
    if (arg == y) atomicAdd(correct, 1);
  }
}

// logits already computed (C > 32): softmax-xent only, one warp per row.
__global__ void xent_rows_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                 float* __restrict__ dlogits, float* __restrict__ loss_sum, int* __restrict__ correct,
                                 int B, int C) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* r = logits + (size_t)warp * C;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = lane; c < C; c += 32)
    if (r[c] > mx) { mx = r[c]; arg = c; }
  for (int o = 16; o > 0; o >>= 1) {
    float om = __shfl_xor_sync(0xffffffffu, mx, o);
    int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(r[c] - mx);
  se = ts::warp_sum(se);
  float lse = mx + logf(se);
  int y = (int)labels[warp];
  float invB = 1.0f / (float)B;
  for (int c = lane; c < C; c += 32)
    dlogits[(size_t)warp * C + c] = (expf(r[c] - lse) - (c == y ? 1.f : 0.f)) * invB;
  if (lane == 0) {
    atomicAdd(loss_sum, lse - r[y]);
    if (arg == y) atomicAdd(correct, 1);
  }
}

}  // namespace

extern "C" int ts_head_xent(const void* h, const float* W, const float* bias, const long long* labels, float* logits,
                            float* dlogits, float* loss_sum, int* correct, int B, int H, int C, int is_bf16,
                            cudaStream_t st) {
  if (C > kMaxC) return -1;
  int thr = 128, blk = (B * 32 + thr - 1) / thr;
  if (is_bf16)
    head_xent_kernel<__nv_bfloat16><<<blk, thr, 0, st>>>((const __nv_bfloat16*)h, W, bias, labels, logits, dlogits,
                                                         loss_sum, correct, B, H, C);
  else
    head_xent_kernel<float><<<blk, thr, 0, st>>>((const float*)h, W, bias, labels, logits, dlogits, loss_sum, correct,
                                                 B, H, C);
  return (int)cudaGetLastError();
}

extern "C" int ts_xent_rows(const float* logits, const long long* labels, float* dlogits, float* loss_sum, int* correct,
                            int B, int C, cudaStream_t st) {
  int thr = 128, blk = (B * 32 + thr - 1) / thr;
  xent_rows_kernel<<<blk, thr, 0, st>>>(logits, labels, dlogits, loss_sum, correct, B, C);
  return (int)cudaGetLastError();
}
