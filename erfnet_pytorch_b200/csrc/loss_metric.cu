// CrossEntropyLoss2d forward+backward in one pass over the logits, the iouEval confusion histogram, and argmax.
// All three are HBM-bound: one coalesced read of the logits / label maps, block-local reductions in shared
// memory, one atomic per block and bin.
#include "common.cuh"

namespace erfk {

__global__ void ce_weight_sum_kernel(const int64_t* __restrict__ target, long long n_px, const float* __restrict__ weight, int C, int ignore,
                                     float* __restrict__ wsum) {
  pdl_enter();
  float a = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    long long t = target[p];
    if (t != ignore && t >= 0 && t < C) a += weight ? weight[t] : 1.f;
  }
  a = warp_sum(a);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    a = warp_sum(a);
    if (threadIdx.x == 0) atomicAdd(wsum, a);
  }
}

constexpr int CE_MAXC = 32;

// NHWC bf16 fast path: one thread per pixel, the pixel's logits travel as LD8 16-byte vectors (LD8 = ld / 8, ld >= C)
template <int LD8>
__global__ void __launch_bounds__(128) ce_fwd_bwd_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ target,
                                                                   long long n_px, const float* __restrict__ weight, int C, int ignore,
                                                                   const float* __restrict__ wsum, float* __restrict__ loss_sum,
                                                                   __nv_bfloat16* __restrict__ dlogits, float* __restrict__ colsum) {
  pdl_enter();
  float lsum = 0.f;
  float csum[LD8 * 8];      // colsum: per-class sums of the STORED (bf16) dlogits = the bias gradient of the conv that produced the logits
#pragma unroll
  for (int c = 0; c < LD8 * 8; ++c) csum[c] = 0.f;
  const float inv_w = 1.f / *wsum;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    float v[LD8 * 8];
    const uint4* src = reinterpret_cast<const uint4*>(logits + (size_t)p * (LD8 * 8));
#pragma unroll
    for (int q = 0; q < LD8; ++q) {
      const uint4 t4 = src[q];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t4);
#pragma unroll
      for (int e = 0; e < 4; ++e) { v[q * 8 + 2 * e] = __low2float(h[e]); v[q * 8 + 2 * e + 1] = __high2float(h[e]); }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < LD8 * 8; ++c)
      if (c < C) mx = fmaxf(mx, v[c]);
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < LD8 * 8; ++c) {
      v[c] = c < C ? __expf(v[c] - mx) : 0.f;
      se += v[c];
    }
    const long long t = target[p];
    const bool valid = t != ignore && t >= 0 && t < C;
    const float w = valid ? (weight ? weight[t] : 1.f) : 0.f;
    const float inv_se = 1.f / se;
    if (valid) {
      float pt = 0.f;
#pragma unroll
      for (int c = 0; c < LD8 * 8; ++c)
        if (c == t) pt = v[c];
      lsum += -w * __logf(pt * inv_se);
    }
    if (dlogits) {
      const float k = w * inv_w;
      uint4* dst = reinterpret_cast<uint4*>(dlogits + (size_t)p * (LD8 * 8));
#pragma unroll
      for (int q = 0; q < LD8; ++q) {
        uint4 t4;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = q * 8 + 2 * e;
          h[e] = __floats2bfloat162_rn(k * (v[c] * inv_se - (c == t ? 1.f : 0.f)), k * (v[c + 1] * inv_se - (c + 1 == t ? 1.f : 0.f)));
          if (colsum) {
            csum[c] += __low2float(h[e]);
            csum[c + 1] += __high2float(h[e]);
          }
        }
        dst[q] = t4;   // padding channels (c >= C) get exactly 0 (v = 0, never the target)
      }
    }
  }
  lsum = warp_sum(lsum);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x < 32) {
    lsum = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    lsum = warp_sum(lsum);
    if (threadIdx.x == 0) atomicAdd(loss_sum, lsum);
  }
  if (colsum) {   // block-level class sums: warp shuffles, one shared-memory row per warp, one global atomic per class and block
    __shared__ float s_cs[4][LD8 * 8];
#pragma unroll
    for (int c = 0; c < LD8 * 8; ++c) {
      const float t = warp_sum(csum[c]);
      if ((threadIdx.x & 31) == 0) s_cs[threadIdx.x >> 5][c] = t;
    }
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(&colsum[threadIdx.x], (s_cs[0][threadIdx.x] + s_cs[1][threadIdx.x]) + (s_cs[2][threadIdx.x] + s_cs[3][threadIdx.x]));
  }
}

template <typename T>
__global__ void ce_fwd_bwd_kernel(const T* __restrict__ logits, int ld, long long hw, const int64_t* __restrict__ target, long long n_px,
                                  const float* __restrict__ weight, int C, int ignore, const float* __restrict__ wsum,
                                  float* __restrict__ loss_sum, T* __restrict__ dlogits) {
  float lsum = 0.f;
  const float inv_w = 1.f / *wsum;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    size_t base, cs;
    if (hw > 0) { base = (size_t)(p / hw) * C * hw + (p % hw); cs = (size_t)hw; } else { base = (size_t)p * ld; cs = 1; }
    float v[CE_MAXC];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CE_MAXC; ++c)
      if (c < C) { v[c] = to_f(logits[base + c * cs]); mx = fmaxf(mx, v[c]); }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < CE_MAXC; ++c)
      if (c < C) { v[c] = __expf(v[c] - mx); se += v[c]; }
    long long t = target[p];
    bool valid = t != ignore && t >= 0 && t < C;
    float w = valid ? (weight ? weight[t] : 1.f) : 0.f;
    float inv_se = 1.f / se;
    if (valid) {
      float pt = 0.f;
#pragma unroll
      for (int c = 0; c < CE_MAXC; ++c)
        if (c == t) pt = v[c];
      lsum += -w * __logf(pt * inv_se);
    }
    if (dlogits) {
      float k = w * inv_w;
#pragma unroll
      for (int c = 0; c < CE_MAXC; ++c)
        if (c < C) dlogits[base + c * cs] = from_f<T>(k * (v[c] * inv_se - (c == t ? 1.f : 0.f)));
    }
  }
  lsum = warp_sum(lsum);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x < 32) {
    lsum = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    lsum = warp_sum(lsum);
    if (threadIdx.x == 0) atomicAdd(loss_sum, lsum);
  }
}

// Confusion matrix cm[g*nc + p].  Per-warp private shared histograms (u32) remove most same-address atomic
// serialisation (label maps are piece-wise constant, so a warp mostly hits one bin: the warp first aggregates
// equal bins with match_any and issues one atomic per distinct bin).
template <typename L>
__global__ void confusion_hist_kernel(const L* __restrict__ pred, const L* __restrict__ gt, long long n, int nc,
                                      unsigned long long* __restrict__ cm) {
  extern __shared__ unsigned int hist[];  // [nc*nc]
  const int bins = nc * nc;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_round = ((n + 31) / 32) * 32;  // keep warps converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    int bin = -1;
    if (i < n) {
      long long p = (long long)pred[i], g = (long long)gt[i];
      if (p >= 0 && p < nc && g >= 0 && g < nc) bin = (int)(g * nc + p);
    }
    unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (bin >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (hist[i]) atomicAdd(&cm[i], (unsigned long long)hist[i]);
}

// tp = diag; fp[c] = sum_{g != ignore} cm[g][c] - tp; fn[c] = sum_p cm[c][p] - tp   (eval/iouEval.py:60-65 on integers)
__global__ void iou_counts_kernel(const unsigned long long* cm, int nc, int ignore, int kept, double* tp, double* fp, double* fn) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= kept) return;
  unsigned long long t = cm[c * nc + c], col = 0, row = 0;
  for (int g = 0; g < nc; ++g)
    if (g != ignore) col += cm[g * nc + c];
  for (int p = 0; p < nc; ++p) row += cm[c * nc + p];
  tp[c] = (double)t;
  fp[c] = (double)(col - t);
  fn[c] = (double)(row - t);
}

template <typename T, typename O>
__global__ void argmax_kernel(const T* __restrict__ logits, int ld, long long hw, long long n_px, int C, O* __restrict__ out) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  size_t base, cs;
  if (hw > 0) { base = (size_t)(p / hw) * C * hw + (p % hw); cs = (size_t)hw; } else { base = (size_t)p * ld; cs = 1; }
  float best = to_f(logits[base]);
  int arg = 0;
  for (int c = 1; c < C; ++c) {
    float v = to_f(logits[base + c * cs]);
    if (v > best) { best = v; arg = c; }  // first maximum wins, like torch.max(1)
  }
  out[p] = (O)arg;
}

}  // namespace erfk

using namespace erfk;

extern "C" int erf_ce_weight_sum(const int64_t* target, long long n_px, const float* weight, int C, int ignore_index, float* wsum,
                                 erf_stream_t stream) {
  ERF_CHECK_ARG(target && wsum && n_px > 0 && C > 0, "erf_ce_weight_sum: bad arguments");
  long long blocks = (n_px + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(ce_weight_sum_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), target, n_px, weight, C, ignore_index, wsum);
  ERF_CHECK_LAUNCH("erf_ce_weight_sum");
  return ERF_OK;
}

static int ce_launch(const void* logits, int ld, int nchw_hw, const int64_t* target, long long n_px, const float* weight, int C, int ignore_index,
                     const float* wsum, float* loss_sum, void* dlogits, float* colsum, int dtype, erf_stream_t stream);

extern "C" int erf_ce_fwd_bwd(const void* logits, int ld, int nchw_hw, const int64_t* target, long long n_px, const float* weight, int C,
                              int ignore_index, const float* wsum, float* loss_sum, void* dlogits, int dtype, erf_stream_t stream) {
  return ce_launch(logits, ld, nchw_hw, target, n_px, weight, C, ignore_index, wsum, loss_sum, dlogits, nullptr, dtype, stream);
}

extern "C" int erf_ce_fwd_bwd_colsum_supported(int ld, int nchw_hw, int dtype) { return dtype == ERF_BF16 && nchw_hw == 0 && (ld == 24 || ld == 32); }

extern "C" int erf_ce_fwd_bwd_colsum(const void* logits, int ld, const int64_t* target, long long n_px, const float* weight, int C, int ignore_index,
                                     const float* wsum, float* loss_sum, void* dlogits, float* colsum, erf_stream_t stream) {
  ERF_CHECK_ARG(colsum && dlogits, "erf_ce_fwd_bwd_colsum: null pointer");
  ERF_CHECK_ARG(erf_ce_fwd_bwd_colsum_supported(ld, 0, ERF_BF16) && ((uintptr_t)logits % 16) == 0 && ((uintptr_t)dlogits % 16) == 0,
                "erf_ce_fwd_bwd_colsum: needs 16-byte aligned NHWC bf16 logits with a pitch of 24 or 32");
  return ce_launch(logits, ld, 0, target, n_px, weight, C, ignore_index, wsum, loss_sum, dlogits, colsum, ERF_BF16, stream);
}

static int ce_launch(const void* logits, int ld, int nchw_hw, const int64_t* target, long long n_px, const float* weight, int C, int ignore_index,
                     const float* wsum, float* loss_sum, void* dlogits, float* colsum, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_ce_fwd_bwd: bad dtype %d", dtype);
  ERF_CHECK_ARG(logits && target && wsum && loss_sum && n_px > 0, "erf_ce_fwd_bwd: null pointer");
  ERF_CHECK_ARG(C > 0 && C <= CE_MAXC, "erf_ce_fwd_bwd: C=%d not in [1,%d]", C, CE_MAXC);
  ERF_CHECK_ARG(nchw_hw > 0 || ld >= C, "erf_ce_fwd_bwd: ld < C");
  long long blocks = (n_px + 127) / 128;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == ERF_BF16 && nchw_hw == 0 && (ld == 24 || ld == 32) &&
      ((uintptr_t)logits % 16) == 0 && (!dlogits || ((uintptr_t)dlogits % 16) == 0)) {
    if (ld == 24)
      launch_pdl(ce_fwd_bwd_nhwc_bf16_kernel<3>, dim3((unsigned)blocks), dim3(128), 0, (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)logits, target, n_px, weight, C, ignore_index, wsum, loss_sum, (__nv_bfloat16*)dlogits, colsum);
    else
      launch_pdl(ce_fwd_bwd_nhwc_bf16_kernel<4>, dim3((unsigned)blocks), dim3(128), 0, (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)logits, target, n_px, weight, C, ignore_index, wsum, loss_sum, (__nv_bfloat16*)dlogits, colsum);
    ERF_CHECK_LAUNCH("erf_ce_fwd_bwd");
    return ERF_OK;
  }
  if (dtype == ERF_F32)
    ce_fwd_bwd_kernel<float><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>((const float*)logits, ld, nchw_hw, target, n_px, weight, C, ignore_index, wsum, loss_sum, (float*)dlogits);
  else
    ce_fwd_bwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, ld, nchw_hw, target, n_px, weight, C, ignore_index, wsum, loss_sum, (__nv_bfloat16*)dlogits);
  ERF_CHECK_LAUNCH("erf_ce_fwd_bwd");
  return ERF_OK;
}

extern "C" int erf_confusion_hist(const void* pred, const void* gt, int elem_bytes, long long n, int nc, unsigned long long* cm,
                                  erf_stream_t stream) {
  ERF_CHECK_ARG(pred && gt && cm && n >= 0, "erf_confusion_hist: null pointer");
  ERF_CHECK_ARG(nc > 0 && nc <= 64, "erf_confusion_hist: nClasses=%d not in [1,64]", nc);
  ERF_CHECK_ARG(elem_bytes == 1 || elem_bytes == 8, "erf_confusion_hist: labels must be uint8 or int64 (elem_bytes=%d)", elem_bytes);
  if (n == 0) return ERF_OK;
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  size_t smem = (size_t)nc * nc * sizeof(unsigned int);
  if (elem_bytes == 1)
    confusion_hist_kernel<uint8_t><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>((const uint8_t*)pred, (const uint8_t*)gt, n, nc, cm);
  else
    confusion_hist_kernel<long long><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>((const long long*)pred, (const long long*)gt, n, nc, cm);
  ERF_CHECK_LAUNCH("erf_confusion_hist");
  return ERF_OK;
}

extern "C" int erf_iou_counts(const unsigned long long* cm, int nc, int ignore, int kept, double* tp, double* fp, double* fn, erf_stream_t stream) {
  ERF_CHECK_ARG(cm && tp && fp && fn && nc > 0 && kept > 0 && kept <= nc, "erf_iou_counts: bad arguments");
  iou_counts_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(cm, nc, ignore, kept, tp, fp, fn);
  ERF_CHECK_LAUNCH("erf_iou_counts");
  return ERF_OK;
}

extern "C" int erf_argmax(const void* logits, int ld, int nchw_hw, long long n_px, int C, void* out, int out_elem_bytes, int dtype,
                          erf_stream_t stream) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_argmax: bad dtype %d", dtype);
  ERF_CHECK_ARG(logits && out && n_px > 0 && C > 0, "erf_argmax: bad arguments");
  ERF_CHECK_ARG(out_elem_bytes == 1 || out_elem_bytes == 8, "erf_argmax: output must be uint8 or int64");
  unsigned grid = (unsigned)((n_px + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ERF_F32) {
    if (out_elem_bytes == 1) argmax_kernel<float, uint8_t><<<grid, 256, 0, s>>>((const float*)logits, ld, nchw_hw, n_px, C, (uint8_t*)out);
    else argmax_kernel<float, long long><<<grid, 256, 0, s>>>((const float*)logits, ld, nchw_hw, n_px, C, (long long*)out);
  } else {
    if (out_elem_bytes == 1) argmax_kernel<__nv_bfloat16, uint8_t><<<grid, 256, 0, s>>>((const __nv_bfloat16*)logits, ld, nchw_hw, n_px, C, (uint8_t*)out);
    else argmax_kernel<__nv_bfloat16, long long><<<grid, 256, 0, s>>>((const __nv_bfloat16*)logits, ld, nchw_hw, n_px, C, (long long*)out);
  }
  ERF_CHECK_LAUNCH("erf_argmax");
  return ERF_OK;
}
