// Masked regression losses of the fork (train/evaluation_functions.py:32-131): one forward pass that reduces
// (sum of per-pixel loss, number of labelled pixels) and one backward pass that writes the input gradients already
// divided by the labelled-pixel count.  The reference builds each of them from 4-8 eager elementwise kernels plus a
// masked_select (stream compaction + host sync); here the mask (target != -1) is a predicate inside the pass.
// HBM-bound: 8-12 B/px (per-pixel kinds), 4*(C+1) B/px (class-probability kinds).
#include "common.cuh"

namespace erfk {

constexpr int ML_MAXC = 32;
constexpr float ML_TWO_PI = 6.283185307179586f;

__device__ __forceinline__ double block_sum_d(double v, double* s) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// per-class value of the class-probability losses and its derivatives w.r.t. cost / variance
template <int KIND>
__device__ __forceinline__ void class_term(float cost, float var, float t, float eps, float& val, float& dcost, float& dvar) {
  const float d = cost - t;
  if (KIND == ERF_LOSS_L1_CLASSPROB) {
    val = fabsf(d);
    dcost = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    dvar = 0.f;
  } else {
    const float v = var + eps;
    val = logf(ML_TWO_PI * v) + d * d / v;
    dcost = 2.f * d / v;
    dvar = 1.f / v - d * d / (v * v);
  }
}

// acc[0] += sum of per-pixel losses over labelled pixels, acc[1] += number of labelled pixels
template <int KIND>
__global__ void __launch_bounds__(256) masked_loss_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              const float* __restrict__ prob, const float* __restrict__ target,
                                                              long long n_px, long long hw, int C, float eps, double* __restrict__ acc) {
  __shared__ double red[8];
  __shared__ float s_cost[ML_MAXC], s_var[ML_MAXC];
  if (KIND == ERF_LOSS_L1_CLASSPROB || KIND == ERF_LOSS_LOGLIK_CLASSPROB) {
    if (threadIdx.x < C) {
      s_cost[threadIdx.x] = a[threadIdx.x];
      s_var[threadIdx.x] = (KIND == ERF_LOSS_LOGLIK_CLASSPROB) ? b[threadIdx.x] : 1.f;
    }
    __syncthreads();
  }
  double lsum = 0.0, cnt = 0.0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    const float t = target[p];
    if (KIND == ERF_LOSS_MSE_WEIGHTED) {
      const float d = a[p] - t;
      lsum += (double)(d * d);
      cnt += 1.0;
      continue;
    }
    if (t == -1.0f) continue;
    cnt += 1.0;
    if (KIND == ERF_LOSS_L1) {
      lsum += (double)fabsf(a[p] - t);
    } else if (KIND == ERF_LOSS_LOGLIK) {
      const float v = b[p] + eps, d = a[p] - t;
      lsum += (double)(logf(ML_TWO_PI * v) + d * d / v);
    } else {
      const size_t base = (size_t)(p / hw) * C * hw + (size_t)(p % hw);
      float s = 0.f;
      for (int c = 0; c < C; ++c) {
        float val, dc, dv;
        class_term<KIND>(s_cost[c], s_var[c], t, eps, val, dc, dv);
        s += val * prob[base + (size_t)c * hw];
      }
      lsum += (double)s;
    }
  }
  lsum = block_sum_d(lsum, red);
  cnt = block_sum_d(cnt, red);
  if (threadIdx.x == 0) {
    atomicAdd(&acc[0], lsum);
    atomicAdd(&acc[1], cnt);
  }
}

// Gradients, scaled by g = go[0] * scale / acc[1].  Per-pixel kinds write d_a (and d_b); class-probability kinds write
// d_prob and accumulate the UNSCALED per-class sums dclass[c] (cost) and dclass[C + c] (variance) in fp64.
template <int KIND>
__global__ void __launch_bounds__(256) masked_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              const float* __restrict__ prob, const float* __restrict__ target,
                                                              long long n_px, long long hw, int C, float eps, float scale,
                                                              const double* __restrict__ acc, const float* __restrict__ go,
                                                              float* __restrict__ d_a, float* __restrict__ d_b, float* __restrict__ d_prob,
                                                              double* __restrict__ dclass) {
  __shared__ float s_cost[ML_MAXC], s_var[ML_MAXC];
  __shared__ double s_dc[2 * ML_MAXC];
  constexpr bool CLS = (KIND == ERF_LOSS_L1_CLASSPROB || KIND == ERF_LOSS_LOGLIK_CLASSPROB);
  if (CLS) {
    if (threadIdx.x < C) {
      s_cost[threadIdx.x] = a[threadIdx.x];
      s_var[threadIdx.x] = (KIND == ERF_LOSS_LOGLIK_CLASSPROB) ? b[threadIdx.x] : 1.f;
    }
    if (threadIdx.x < 2 * ML_MAXC) s_dc[threadIdx.x] = 0.0;
    __syncthreads();
  }
  const float g = (float)((double)(go ? go[0] : 1.f) * (double)scale / acc[1]);
  const long long n_round = ((n_px + 31) / 32) * 32;  // warps stay converged for the shuffles of the class kinds
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += (long long)gridDim.x * blockDim.x) {
    const bool in = p < n_px;
    const float t = in ? target[p] : -1.0f;
    if (!CLS) {
      if (!in) continue;
      if (KIND == ERF_LOSS_MSE_WEIGHTED) {
        d_a[p] = 2.f * (a[p] - t) * g;
      } else if (t == -1.0f) {
        d_a[p] = 0.f;
        if (KIND == ERF_LOSS_LOGLIK) d_b[p] = 0.f;
      } else if (KIND == ERF_LOSS_L1) {
        const float d = a[p] - t;
        d_a[p] = (d > 0.f ? g : (d < 0.f ? -g : 0.f));
      } else {
        const float v = b[p] + eps, d = a[p] - t;
        d_a[p] = 2.f * d / v * g;
        d_b[p] = (1.f / v - d * d / (v * v)) * g;
      }
      continue;
    }
    const bool valid = in && t != -1.0f;
    const size_t base = in ? (size_t)(p / hw) * C * hw + (size_t)(p % hw) : 0;
    for (int c = 0; c < C; ++c) {
      float val = 0.f, dc = 0.f, dv = 0.f, pr = 0.f;
      if (valid) {
        class_term<KIND>(s_cost[c], s_var[c], t, eps, val, dc, dv);
        pr = prob[base + (size_t)c * hw];
      }
      if (in) d_prob[base + (size_t)c * hw] = valid ? val * g : 0.f;
      float sc = warp_sum(dc * pr);
      float sv = (KIND == ERF_LOSS_LOGLIK_CLASSPROB) ? warp_sum(dv * pr) : 0.f;
      if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_dc[c], (double)sc);
        if (KIND == ERF_LOSS_LOGLIK_CLASSPROB) atomicAdd(&s_dc[ML_MAXC + c], (double)sv);
      }
    }
  }
  if (CLS) {
    __syncthreads();
    if (threadIdx.x < C) {
      atomicAdd(&dclass[threadIdx.x], s_dc[threadIdx.x]);
      if (KIND == ERF_LOSS_LOGLIK_CLASSPROB) atomicAdd(&dclass[C + threadIdx.x], s_dc[ML_MAXC + threadIdx.x]);
    }
  }
}

static unsigned ml_grid(long long n) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace erfk

using namespace erfk;

static int ml_check(const char* fn, int kind, const float* a, const float* b, const float* prob, const float* target, long long n_px,
                    long long hw, int C) {
  ERF_CHECK_ARG(kind >= ERF_LOSS_L1 && kind <= ERF_LOSS_MSE_WEIGHTED, "%s: unknown loss kind %d", fn, kind);
  ERF_CHECK_ARG(a && target && n_px > 0, "%s: null pointer / empty input", fn);
  ERF_CHECK_ARG(b != nullptr || (kind != ERF_LOSS_LOGLIK && kind != ERF_LOSS_LOGLIK_CLASSPROB), "%s: variance operand missing", fn);
  if (kind == ERF_LOSS_L1_CLASSPROB || kind == ERF_LOSS_LOGLIK_CLASSPROB) {
    ERF_CHECK_ARG(prob != nullptr, "%s: class probabilities missing", fn);
    ERF_CHECK_ARG(C > 0 && C <= ML_MAXC, "%s: C=%d not in [1,%d]", fn, C, ML_MAXC);
    ERF_CHECK_ARG(hw > 0 && n_px % hw == 0, "%s: n_px=%lld is not a multiple of the image size %lld", fn, n_px, hw);
  }
  return ERF_OK;
}

extern "C" int erf_masked_loss_fwd(int kind, const float* a, const float* b, const float* prob, const float* target, long long n_px,
                                   long long hw, int C, float eps, double* acc, erf_stream_t stream) {
  if (int rc = ml_check("erf_masked_loss_fwd", kind, a, b, prob, target, n_px, hw, C)) return rc;
  ERF_CHECK_ARG(acc != nullptr, "erf_masked_loss_fwd: null accumulator");
  const unsigned grid = ml_grid(n_px);
  cudaStream_t s = (cudaStream_t)stream;
  switch (kind) {
    case ERF_LOSS_L1: masked_loss_fwd_kernel<ERF_LOSS_L1><<<grid, 256, 0, s>>>(a, b, prob, target, n_px, hw, C, eps, acc); break;
    case ERF_LOSS_LOGLIK: masked_loss_fwd_kernel<ERF_LOSS_LOGLIK><<<grid, 256, 0, s>>>(a, b, prob, target, n_px, hw, C, eps, acc); break;
    case ERF_LOSS_L1_CLASSPROB: masked_loss_fwd_kernel<ERF_LOSS_L1_CLASSPROB><<<grid, 256, 0, s>>>(a, b, prob, target, n_px, hw, C, eps, acc); break;
    case ERF_LOSS_LOGLIK_CLASSPROB: masked_loss_fwd_kernel<ERF_LOSS_LOGLIK_CLASSPROB><<<grid, 256, 0, s>>>(a, b, prob, target, n_px, hw, C, eps, acc); break;
    default: masked_loss_fwd_kernel<ERF_LOSS_MSE_WEIGHTED><<<grid, 256, 0, s>>>(a, b, prob, target, n_px, hw, C, eps, acc); break;
  }
  ERF_CHECK_LAUNCH("erf_masked_loss_fwd");
  return ERF_OK;
}

extern "C" int erf_masked_loss_bwd(int kind, const float* a, const float* b, const float* prob, const float* target, long long n_px,
                                   long long hw, int C, float eps, float scale, const double* acc, const float* go, float* d_a, float* d_b,
                                   float* d_prob, double* dclass, erf_stream_t stream) {
  if (int rc = ml_check("erf_masked_loss_bwd", kind, a, b, prob, target, n_px, hw, C)) return rc;
  ERF_CHECK_ARG(acc != nullptr, "erf_masked_loss_bwd: null accumulator");
  const bool cls = kind == ERF_LOSS_L1_CLASSPROB || kind == ERF_LOSS_LOGLIK_CLASSPROB;
  ERF_CHECK_ARG(!cls || (d_prob && dclass), "erf_masked_loss_bwd: d_prob / dclass missing");
  ERF_CHECK_ARG(cls || (d_a && (kind != ERF_LOSS_LOGLIK || d_b)), "erf_masked_loss_bwd: gradient buffer missing");
  const unsigned grid = ml_grid(n_px);
  cudaStream_t s = (cudaStream_t)stream;
#define ERF_ML_BWD(K) masked_loss_bwd_kernel<K><<<grid, 256, 0, s>>>(a, b, prob, target, n_px, hw, C, eps, scale, acc, go, d_a, d_b, d_prob, dclass)
  switch (kind) {
    case ERF_LOSS_L1: ERF_ML_BWD(ERF_LOSS_L1); break;
    case ERF_LOSS_LOGLIK: ERF_ML_BWD(ERF_LOSS_LOGLIK); break;
    case ERF_LOSS_L1_CLASSPROB: ERF_ML_BWD(ERF_LOSS_L1_CLASSPROB); break;
    case ERF_LOSS_LOGLIK_CLASSPROB: ERF_ML_BWD(ERF_LOSS_LOGLIK_CLASSPROB); break;
    default: ERF_ML_BWD(ERF_LOSS_MSE_WEIGHTED); break;
  }
#undef ERF_ML_BWD
  ERF_CHECK_LAUNCH("erf_masked_loss_bwd");
  return ERF_OK;
}

// ---- SoftMaxConv tail (train/erfnet_blocks.py:92-110): Dropout2d multiplier + softmax over the channels -----------------------------
// forward: prob[n][c][p] (NCHW fp32) = softmax_c(logits[n][p][c] * drop[n][c]) from NHWC logits with row pitch ld;
// backward: dlogits[n][p][c] = drop[n][c] * prob_c * (dprob_c - sum_k dprob_k prob_k), pad channels (c >= C) zeroed.
namespace erfk {

template <typename T>
__global__ void __launch_bounds__(256) softmax_nhwc_fwd_kernel(const T* __restrict__ logits, int ld, const float* __restrict__ drop, long long hw,
                                                               long long n_px, int C, float* __restrict__ prob) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  const long long n = p / hw, q = p - n * hw;
  float v[ML_MAXC];
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) {
    v[c] = to_f(logits[(size_t)p * ld + c]) * (drop ? drop[n * C + c] : 1.f);
    mx = fmaxf(mx, v[c]);
  }
  float se = 0.f;
  for (int c = 0; c < C; ++c) { v[c] = __expf(v[c] - mx); se += v[c]; }
  const float inv = 1.f / se;
  for (int c = 0; c < C; ++c) prob[((size_t)n * C + c) * hw + q] = v[c] * inv;
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_nhwc_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ dprob,
                                                               const float* __restrict__ drop, long long hw, long long n_px, int C,
                                                               T* __restrict__ dlogits, int ld) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  const long long n = p / hw, q = p - n * hw;
  float pr[ML_MAXC], dp[ML_MAXC];
  float dot = 0.f;
  for (int c = 0; c < C; ++c) {
    pr[c] = prob[((size_t)n * C + c) * hw + q];
    dp[c] = dprob[((size_t)n * C + c) * hw + q];
    dot += pr[c] * dp[c];
  }
  for (int c = 0; c < ld; ++c)
    dlogits[(size_t)p * ld + c] = from_f<T>(c < C ? (drop ? drop[n * C + c] : 1.f) * pr[c] * (dp[c] - dot) : 0.f);
}

}  // namespace erfk

extern "C" int erf_softmax_nhwc_fwd(const void* logits, int ld, const float* drop, int N, long long hw, int C, float* prob, int dtype,
                                    erf_stream_t stream) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_softmax_nhwc_fwd: bad dtype %d", dtype);
  ERF_CHECK_ARG(logits && prob && N > 0 && hw > 0 && C > 0 && C <= ML_MAXC && ld >= C, "erf_softmax_nhwc_fwd: bad arguments (C=%d, ld=%d)", C, ld);
  const long long n_px = (long long)N * hw;
  const unsigned grid = (unsigned)((n_px + 255) / 256);
  if (dtype == ERF_F32)
    softmax_nhwc_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)logits, ld, drop, hw, n_px, C, prob);
  else
    softmax_nhwc_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, ld, drop, hw, n_px, C, prob);
  ERF_CHECK_LAUNCH("erf_softmax_nhwc_fwd");
  return ERF_OK;
}

extern "C" int erf_softmax_nhwc_bwd(const float* prob, const float* dprob, const float* drop, int N, long long hw, int C, void* dlogits, int ld,
                                    int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_softmax_nhwc_bwd: bad dtype %d", dtype);
  ERF_CHECK_ARG(prob && dprob && dlogits && N > 0 && hw > 0 && C > 0 && C <= ML_MAXC && ld >= C, "erf_softmax_nhwc_bwd: bad arguments (C=%d, ld=%d)", C, ld);
  const long long n_px = (long long)N * hw;
  const unsigned grid = (unsigned)((n_px + 255) / 256);
  if (dtype == ERF_F32)
    softmax_nhwc_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(prob, dprob, drop, hw, n_px, C, (float*)dlogits, ld);
  else
    softmax_nhwc_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(prob, dprob, drop, hw, n_px, C, (__nv_bfloat16*)dlogits, ld);
  ERF_CHECK_LAUNCH("erf_softmax_nhwc_bwd");
  return ERF_OK;
}
