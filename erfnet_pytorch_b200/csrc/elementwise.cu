// Bandwidth-bound pieces of the ERFNet path: layout conversion at the NCHW boundary, MaxPool half of the
// downsampler, BatchNorm statistics / apply / backward, residual+ReLU, Adam.  All coalesced along the channel
// (innermost NHWC) dimension; per-channel reductions use shared-memory partials + one atomic per block/channel.
#include "common.cuh"
#include "bn_stream.cuh"

namespace erfk {

// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int C, long long HW, long long total_px, T* __restrict__ out, int ld) {
  pdl_enter();
  // one thread per pixel: reads are coalesced per channel plane, the (small) pixel vector is written contiguously
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_px) return;
  long long n = p / HW, hw = p % HW;
  const float* ip = in + n * C * HW + hw;
  T* op = out + p * ld;
  for (int c = 0; c < ld; ++c) op[c] = from_f<T>(c < C ? ip[(long long)c * HW] : 0.f);
}

// The image of the reference's API (3 x fp32 planes) -> NHWC bf16 with the pixel padded to 4 channels (the stem's input): a thread converts FOUR
// consecutive pixels -- three 16-byte loads, two 16-byte stores -- instead of one pixel with four 2-byte stores (28 -> ~10 us for 6 x 512 x 1024; it
// is the first kernel of every training step).  Needs HW % 4 == 0 and 16-byte aligned tensors.
__global__ void nchw3_to_nhwc4_bf16_kernel(const float* __restrict__ in, unsigned HW, unsigned total_quads, __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total_quads) return;
  const unsigned p = q * 4u, n = p / HW, hw = p - n * HW;
  const float* ip = in + (size_t)n * 3 * HW + hw;
  const float4 r = *reinterpret_cast<const float4*>(ip), g = *reinterpret_cast<const float4*>(ip + HW), b = *reinterpret_cast<const float4*>(ip + 2 * (size_t)HW);
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
  uint4 o0, o1;
  __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
  __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
  h0[0] = __floats2bfloat162_rn(r.x, g.x); h0[1] = __floats2bfloat162_rn(b.x, 0.f);
  h0[2] = __floats2bfloat162_rn(r.y, g.y); h0[3] = __floats2bfloat162_rn(b.y, 0.f);
  h1[0] = __floats2bfloat162_rn(r.z, g.z); h1[1] = __floats2bfloat162_rn(b.z, 0.f);
  h1[2] = __floats2bfloat162_rn(r.w, g.w); h1[3] = __floats2bfloat162_rn(b.w, 0.f);
  (void)z;
  uint4* op = reinterpret_cast<uint4*>(out + (size_t)p * 4);
  op[0] = o0;
  op[1] = o1;
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int ld, int C, long long HW, long long total_px, float* __restrict__ out) {
  pdl_enter();
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total_px) return;
  long long n = p / HW, hw = p % HW;
  const T* ip = in + p * ld;
  float* op = out + n * C * HW + hw;
  for (int c = 0; c < C; ++c) op[(long long)c * HW] = to_f(ip[c]);
}

// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ in, int N, int H, int W, int C, int in_ld, T* __restrict__ out, int out_ld,
                                   int out_coff, const float* __restrict__ scale, const float* __restrict__ shift, int relu, double* stats,
                                   int stats_c) {
  extern __shared__ float s_stats[];  // [2*C]
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_stats[i] = 0.f;
    __syncthreads();
  }
  const int Ho = H / 2, Wo = W / 2;
  long long total = (long long)N * Ho * Wo * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    int c = (int)(i % C);
    long long p = i / C;
    int ox = (int)(p % Wo);
    long long t = p / Wo;
    int oy = (int)(t % Ho);
    int n = (int)(t / Ho);
    const T* ip = in + (((size_t)n * H + 2 * oy) * W + 2 * ox) * in_ld + c;
    float m = to_f(ip[0]);
    m = fmaxf(m, to_f(ip[in_ld]));
    m = fmaxf(m, to_f(ip[(size_t)W * in_ld]));
    m = fmaxf(m, to_f(ip[(size_t)W * in_ld + in_ld]));
    float o = m;
    if (scale) o = o * scale[c] + shift[c];
    if (relu) o = fmaxf(o, 0.f);
    out[(size_t)p * out_ld + out_coff + c] = from_f<T>(o);
    if (stats) {
      atomicAdd(&s_stats[c], m);
      atomicAdd(&s_stats[C + c], m * m);
    }
  }
  if (stats) {
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      atomicAdd(&stats[out_coff + k], (double)s_stats[k]);
      atomicAdd(&stats[stats_c + out_coff + k], (double)s_stats[C + k]);
    }
  }
}

template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ in, int N, int H, int W, int C, int in_ld, const T* __restrict__ dout, int out_ld,
                                   int out_coff, T* __restrict__ dx, int dx_ld) {
  const int Ho = H / 2, Wo = W / 2;
  long long total = (long long)N * Ho * Wo * C;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long p = i / C;
  int ox = (int)(p % Wo);
  long long t = p / Wo;
  int oy = (int)(t % Ho);
  int n = (int)(t / Ho);
  size_t base = ((size_t)n * H + 2 * oy) * W + 2 * ox;
  const T* ip = in + base * in_ld + c;
  // torch keeps the FIRST maximum in (ky, kx) scan order: update only on strictly greater
  float m = to_f(ip[0]);
  size_t arg = 0;
  float v = to_f(ip[in_ld]);
  if (v > m) { m = v; arg = 1; }
  v = to_f(ip[(size_t)W * in_ld]);
  if (v > m) { m = v; arg = W; }
  v = to_f(ip[(size_t)W * in_ld + in_ld]);
  if (v > m) { m = v; arg = (size_t)W + 1; }
  T* dp = dx + (base + arg) * dx_ld + c;
  *dp = from_f<T>(to_f(*dp) + to_f(dout[(size_t)p * out_ld + out_coff + c]));
}

// stats[c] += sum_p x[p][c], stats[C+c] += sum_p x[p][c]^2  (BN batch statistics of a stored activation)
template <typename T>
__global__ void channel_stats_kernel(const T* __restrict__ x, int ld, int C, long long n_px, int px_per_block, double* __restrict__ stats) {
  extern __shared__ float s[];  // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s[i] = 0.f;
  __syncthreads();
  const int c8 = C / 8;                       // 8-channel (16-byte for bf16) groups
  const int lanes = blockDim.x / c8;
  const int g = threadIdx.x % c8, lane = threadIdx.x / c8;
  long long p0 = (long long)blockIdx.x * px_per_block, p1 = p0 + px_per_block;
  if (p1 > n_px) p1 = n_px;
  if (lane < lanes) {
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = b[k] = 0.f;
    for (long long p = p0 + lane; p < p1; p += lanes) {
      Vec<T, 8> v;
      v.load(x + (size_t)p * ld + g * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) { a[k] += v.v[k]; b[k] += v.v[k] * v.v[k]; }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s[g * 8 + k], a[k]); atomicAdd(&s[C + g * 8 + k], b[k]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[i], (double)s[i]);
}

// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void affine_act_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ scale, const float* __restrict__ shift,
                                  int per_sample, const T* __restrict__ res, int res_ld, T* __restrict__ out, int out_ld, long long HW,
                                  int C, long long total4, int relu) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over groups of 4 channels
  if (i >= total4) return;
  int c4 = C / 4;
  int c = (int)(i % c4) * 4;
  long long p = i / c4;
  int sidx = per_sample ? (int)(p / HW) * C + c : c;
  Vec<T, 4> v;
  v.load(x + (size_t)p * x_ld + c);
  Vec<T, 4> r;
  if (res) r.load(res + (size_t)p * res_ld + c);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float f = v.v[k];
    if (scale) f = f * scale[sidx + k] + shift[sidx + k];
    if (res) f += r.v[k];
    if (relu) f = fmaxf(f, 0.f);
    v.v[k] = f;
  }
  v.store(out + (size_t)p * out_ld + c);
}

// out = g * (mask > 0)      (ReLU backward as a standalone op)
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ g, int g_ld, const T* __restrict__ mask, int mask_ld, T* __restrict__ out, int out_ld,
                                int C, long long total4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int c4 = C / 4;
  int c = (int)(i % c4) * 4;
  long long p = i / c4;
  Vec<T, 4> v, m;
  v.load(g + (size_t)p * g_ld + c);
  m.load(mask + (size_t)p * mask_ld + c);
#pragma unroll
  for (int k = 0; k < 4; ++k) v.v[k] = m.v[k] > 0.f ? v.v[k] : 0.f;
  v.store(out + (size_t)p * out_ld + c);
}

// ------------------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(double* stats, double count, const float* gamma, const float* beta, float* rm, float* rv, float momentum,
                                   float eps, const float* drop, int N, int C, float* mean, float* invstd, float* scale, float* shift) {
  pdl_enter();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = stats[c], sq = stats[C + c];
  stats[c] = 0.0;
  stats[C + c] = 0.0;
  double mu = s / count;
  double var = sq / count - mu * mu;
  if (var < 0) var = 0;
  float is = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  invstd[c] = is;
  if (rm) {
    double unb = count > 1 ? var * count / (count - 1) : var;
    rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mu;
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
  }
  float sc = gamma[c] * is;
  float sh = beta[c] - (float)mu * sc;
  if (drop) {
    for (int n = 0; n < N; ++n) {
      float d = drop[n * C + c];
      scale[n * C + c] = sc * d;
      shift[n * C + c] = sh * d;
    }
  } else {
    scale[c] = sc;
    shift[c] = sh;
  }
}

__global__ void bn_eval_affine_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps, int C, float* scale,
                                      float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

// gb = g * (mask > 0) * drop[n,c]; block-partial sums of gb and gb*xhat.  8 channels per thread (16-byte loads).
template <typename T>
__global__ void bn_bwd_reduce_kernel(const T* __restrict__ g, int g_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ mask,
                                     int mask_ld, const float* __restrict__ drop, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, long long HW, int C, long long n_px, int px_per_block,
                                     double* __restrict__ sums) {
  extern __shared__ float s[];  // [2*C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s[i] = 0.f;
  __syncthreads();
  const int c8 = C / 8;
  const int lanes = blockDim.x / c8;
  const int c0 = (threadIdx.x % c8) * 8, lane = threadIdx.x / c8;
  long long p0 = (long long)blockIdx.x * px_per_block, p1 = p0 + px_per_block;
  if (p1 > n_px) p1 = n_px;
  if (lane < lanes) {
    float a[8], b[8], mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = b[k] = 0.f; mu[k] = mean[c0 + k]; is[k] = invstd[c0 + k]; }
    for (long long p = p0 + lane; p < p1; p += lanes) {
      Vec<T, 8> gv, yv, mv;
      gv.load(g + (size_t)p * g_ld + c0);
      yv.load(y + (size_t)p * y_ld + c0);
      if (mask) mv.load(mask + (size_t)p * mask_ld + c0);
      const float* dp = drop ? drop + (p / HW) * C + c0 : nullptr;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float gg = gv.v[k];
        if (mask && !(mv.v[k] > 0.f)) gg = 0.f;
        if (dp) gg *= dp[k];
        a[k] += gg;
        b[k] += gg * (yv.v[k] - mu[k]) * is[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { atomicAdd(&s[c0 + k], a[k]); atomicAdd(&s[C + c0 + k], b[k]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], (double)s[i]);
}

template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ g, int g_ld, const T* __restrict__ y, int y_ld, const T* __restrict__ mask,
                                    int mask_ld, const float* __restrict__ drop, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ invstd, const double* __restrict__ sums,
                                    long long HW, int C, long long n_px, T* __restrict__ dy, int dy_ld, float* dgamma, float* dbeta,
                                    T* __restrict__ gmasked, int gmasked_ld) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over (pixel, 8-channel group)
  if (i < C) {
    if (dbeta) dbeta[i] = (float)sums[i];
    if (dgamma) dgamma[i] = (float)sums[C + i];
  }
  const int c8 = C / 8;
  if (i >= n_px * c8) return;
  const int c0 = (int)(i % c8) * 8;
  const long long p = i / c8;
  Vec<T, 8> gv, yv, mv;
  gv.load(g + (size_t)p * g_ld + c0);
  yv.load(y + (size_t)p * y_ld + c0);
  if (mask) mv.load(mask + (size_t)p * mask_ld + c0);
  const float* dp = drop ? drop + (p / HW) * C + c0 : nullptr;
  const float inv_m = 1.f / (float)n_px;
  if (gmasked) {  // g * (mask > 0): the ReLU backward the residual path needs, produced while g and mask are in registers
    Vec<T, 8> gm;
#pragma unroll
    for (int k = 0; k < 8; ++k) gm.v[k] = (mask && !(mv.v[k] > 0.f)) ? 0.f : (float)gv.v[k];
    gm.store(gmasked + (size_t)p * gmasked_ld + c0);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float gg = gv.v[k];
    if (mask && !(mv.v[k] > 0.f)) gg = 0.f;
    if (dp) gg *= dp[k];
    const float is = invstd[c0 + k];
    const float xh = (yv.v[k] - mean[c0 + k]) * is;
    gv.v[k] = gamma[c0 + k] * is * (gg - (float)sums[c0 + k] * inv_m - xh * (float)sums[C + c0 + k] * inv_m);
  }
  gv.store(dy + (size_t)p * dy_ld + c0);
}

// tcgen05 operand images of a 3-tap conv weight w[Co][Ci][3] (fp32): forward Wt[t][n=co][k=ci] and data-gradient
// Wt'[t][n=ci][k=co] = w[co][ci][2-t]; layout [3][C/64][C][8 chunks of 8], chunk index XOR-swizzled by (n & 7).
__device__ __forceinline__ void pack_conv1d_tc_body(const float* __restrict__ w, int C, __nv_bfloat16* __restrict__ img_fwd,
                                                    __nv_bfloat16* __restrict__ img_dgrad, int i) {
  const int ch = C < 64 ? C : 64;   // channels per K chunk (one swizzle row)
  const int cpr = ch / 8;           // 16-byte chunks per row: 8 (128B swizzle) or 2 (32B swizzle)
  const int kc_n = C / ch;
  const int total = 3 * kc_n * C * cpr;  // 16-byte chunks per image
  if (i >= 2 * total) return;
  const bool dgrad = i >= total;
  if (dgrad) i -= total;
  const int jj = i % cpr, n = (i / cpr) % C, kc = (i / (cpr * C)) % kc_n, t = i / (cpr * C * kc_n);
  // swizzle: address bits [4,7) ^= bits [7,10) (128B rows) ; bit 4 ^= bit 7 (32B rows: bit 7 of n*32 is bit 2 of n)
  const int j = cpr == 8 ? (jj ^ (n & 7)) : (jj ^ ((n >> 2) & 1));
  Vec<__nv_bfloat16, 8> v;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kc * ch + j * 8 + e;
    v.v[e] = dgrad ? w[((size_t)k * C + n) * 3 + (2 - t)] : w[((size_t)n * C + k) * 3 + t];
  }
  v.store((dgrad ? img_dgrad : img_fwd) + (size_t)i * 8);
}

__global__ void pack_conv1d_tc_kernel(const float* __restrict__ w, int C, __nv_bfloat16* __restrict__ img_fwd, __nv_bfloat16* __restrict__ img_dgrad) {
  pack_conv1d_tc_body(w, C, img_fwd, img_dgrad, blockIdx.x * blockDim.x + threadIdx.x);
}

// every 3-tap conv of the net in ONE launch (blockIdx.y = table entry): the per-step weight re-pack of a training step
__global__ void pack_conv1d_tc_batched_kernel(const erf_pack1d_entry* __restrict__ table) {
  pdl_enter();
  const erf_pack1d_entry e = table[blockIdx.y];
  pack_conv1d_tc_body(e.w, e.C, (__nv_bfloat16*)e.img_fwd, (__nv_bfloat16*)e.img_dgrad, blockIdx.x * blockDim.x + threadIdx.x);
}

// ------------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr_over_bc1, float beta1, float beta2, float eps, float wd, float inv_sqrt_bc2, float grad_scale,
                            const int* __restrict__ step_dev, float lr, const float* __restrict__ hyper_dev) {
  pdl_enter();
  if (hyper_dev) {  // CUDA-graph friendly: learning rate and weight decay live on the device (a captured launch bakes by-value arguments in)
    lr = hyper_dev[0];
    wd = hyper_dev[1];
  }
  if (step_dev) {  // CUDA-graph friendly: the step count lives on the device (incremented by adam_step_inc_kernel)
    const double st = (double)*step_dev;
    lr_over_bc1 = (float)((double)lr / (1.0 - pow((double)beta1, st)));
    inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, st)));
  }
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float pv = p[i];
  float gv = g[i] * grad_scale + wd * pv;
  float mv = beta1 * m[i] + (1.f - beta1) * gv;
  float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
  m[i] = mv;
  v[i] = vv;
  p[i] = pv - lr_over_bc1 * mv / (sqrtf(vv) * inv_sqrt_bc2 + eps);
}

__global__ void adam_step_inc_kernel(int* step_dev) {
  pdl_enter(); *step_dev += 1; }

}  // namespace erfk

using namespace erfk;
#define DISPATCH(dtype, KERNEL, grid, block, smem, stream, ...)                                   \
  do {                                                                                            \
    if ((dtype) == ERF_F32) KERNEL<float><<<grid, block, smem, stream>>>(__VA_ARGS__);            \
    else KERNEL<__nv_bfloat16><<<grid, block, smem, stream>>>(__VA_ARGS__);                       \
  } while (0)
#define CHECK_DTYPE(who) ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, who ": bad dtype %d", dtype)

extern "C" int erf_nchw_f32_to_nhwc(const float* in, int N, int C, int H, int W, void* out, int out_ld, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_nchw_f32_to_nhwc");
  ERF_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0 && out_ld >= C, "erf_nchw_f32_to_nhwc: bad arguments");
  long long HW = (long long)H * W, total = HW * N;
  if (dtype == ERF_BF16 && C == 3 && out_ld == 4 && (HW % 4) == 0 && total < 0x7fffffffLL && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0) {
    const unsigned quads = (unsigned)(total / 4);
    launch_pdl(nchw3_to_nhwc4_bf16_kernel, dim3((quads + 255) / 256), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), in, (unsigned)HW, quads, (__nv_bfloat16*)out);
    ERF_CHECK_LAUNCH("erf_nchw_f32_to_nhwc");
    return ERF_OK;
  }
  unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == ERF_F32) launch_pdl(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), in, C, HW, total, (float*)out, out_ld);
  else launch_pdl(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), in, C, HW, total, (__nv_bfloat16*)out, out_ld);
  ERF_CHECK_LAUNCH("erf_nchw_f32_to_nhwc");
  return ERF_OK;
}

extern "C" int erf_nhwc_to_nchw_f32(const void* in, int in_ld, int N, int C, int H, int W, float* out, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_nhwc_to_nchw_f32");
  ERF_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0 && in_ld >= C, "erf_nhwc_to_nchw_f32: bad arguments");
  long long HW = (long long)H * W, total = HW * N;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == ERF_F32) launch_pdl(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), (const float*)in, in_ld, C, HW, total, out);
  else launch_pdl(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)in, in_ld, C, HW, total, out);
  ERF_CHECK_LAUNCH("erf_nhwc_to_nchw_f32");
  return ERF_OK;
}

extern "C" int erf_maxpool2x2_fwd(const void* in, int N, int H, int W, int C, int in_ld, void* out, int out_ld, int out_coff, const float* scale,
                                  const float* shift, int relu, double* stats, int stats_c, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_maxpool2x2_fwd");
  ERF_CHECK_ARG(in && out && N > 0 && C > 0 && H >= 2 && W >= 2 && (H % 2) == 0 && (W % 2) == 0, "erf_maxpool2x2_fwd: H (%d) and W (%d) must be even", H, W);
  ERF_CHECK_ARG(in_ld >= C && out_ld >= out_coff + C, "erf_maxpool2x2_fwd: bad leading dimensions");
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  unsigned grid = (unsigned)((total + 255) / 256);
  size_t smem = stats ? 2 * C * sizeof(float) : 0;
  {
    const long long total8 = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (!stats && (C % 8) == 0 && (in_ld % 8) == 0 && (out_ld % 8) == 0 && (out_coff % 8) == 0 && total8 < 0x7fffffffLL && ((uintptr_t)in % 16) == 0 &&
        ((uintptr_t)out % 16) == 0) {
      const unsigned g8 = (unsigned)((total8 + 255) / 256);
      if (dtype == ERF_F32)
        launch_pdl(maxpool_fwd_v2_kernel<float>, dim3(g8), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), (const float*)in, H, W, C, in_ld, (float*)out, out_ld, out_coff, scale, shift, relu, (unsigned)total8);
      else
        launch_pdl(maxpool_fwd_v2_kernel<__nv_bfloat16>, dim3(g8), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)in, H, W, C, in_ld, (__nv_bfloat16*)out, out_ld, out_coff, scale, shift, relu, (unsigned)total8);
      ERF_CHECK_LAUNCH("erf_maxpool2x2_fwd");
      return ERF_OK;
    }
  }
  if (dtype == ERF_F32)
    maxpool_fwd_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((const float*)in, N, H, W, C, in_ld, (float*)out, out_ld, out_coff, scale, shift, relu, stats, stats_c);
  else
    maxpool_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, N, H, W, C, in_ld, (__nv_bfloat16*)out, out_ld, out_coff, scale, shift, relu, stats, stats_c);
  ERF_CHECK_LAUNCH("erf_maxpool2x2_fwd");
  return ERF_OK;
}

extern "C" int erf_maxpool2x2_bwd(const void* in, int N, int H, int W, int C, int in_ld, const void* dout, int out_ld, int out_coff, void* dx,
                                  int dx_ld, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_maxpool2x2_bwd");
  ERF_CHECK_ARG(in && dout && dx && N > 0 && C > 0 && (H % 2) == 0 && (W % 2) == 0, "erf_maxpool2x2_bwd: bad arguments");
  {
    const long long total8 = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if ((C % 8) == 0 && (in_ld % 8) == 0 && (out_ld % 8) == 0 && (out_coff % 8) == 0 && (dx_ld % 8) == 0 && total8 < 0x7fffffffLL &&
        ((uintptr_t)in % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dx % 16) == 0) {
      const unsigned g8 = (unsigned)((total8 + 255) / 256);
      if (dtype == ERF_F32)
        launch_pdl(maxpool_bwd_v2_kernel<float>, dim3(g8), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), (const float*)in, H, W, C, in_ld, (const float*)dout, out_ld, out_coff, (float*)dx, dx_ld, (unsigned)total8);
      else
        launch_pdl(maxpool_bwd_v2_kernel<__nv_bfloat16>, dim3(g8), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)in, H, W, C, in_ld, (const __nv_bfloat16*)dout, out_ld, out_coff, (__nv_bfloat16*)dx, dx_ld, (unsigned)total8);
      ERF_CHECK_LAUNCH("erf_maxpool2x2_bwd");
      return ERF_OK;
    }
  }
  long long total = (long long)N * (H / 2) * (W / 2) * C;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == ERF_F32)
    maxpool_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)in, N, H, W, C, in_ld, (const float*)dout, out_ld, out_coff, (float*)dx, dx_ld);
  else
    maxpool_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, N, H, W, C, in_ld, (const __nv_bfloat16*)dout, out_ld, out_coff, (__nv_bfloat16*)dx, dx_ld);
  ERF_CHECK_LAUNCH("erf_maxpool2x2_bwd");
  return ERF_OK;
}

extern "C" int erf_affine_act(const void* x, int x_ld, const float* scale, const float* shift, int sc_per_sample, const void* res, int res_ld,
                              void* out, int out_ld, int N, int HW, int C, int relu, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_affine_act");
  ERF_CHECK_ARG(x && out && N > 0 && HW > 0 && C > 0 && (C % 4) == 0 && (x_ld % 4) == 0 && (out_ld % 4) == 0 && (!res || (res_ld % 4) == 0),
                "erf_affine_act: C and leading dimensions must be multiples of 4");
  ERF_CHECK_ARG((scale == nullptr) == (shift == nullptr), "erf_affine_act: scale and shift must be given together");
  if (bs_fast_ok(C) && (x_ld % 8) == 0 && (out_ld % 8) == 0 && (!res || (res_ld % 8) == 0) && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
      ((uintptr_t)res % 16) == 0) {
    const long long n_px = (long long)N * HW;
    const BsLaunch l = bs_launch(C, n_px, 4);
    if (dtype == ERF_F32)
      launch_pdl(affine_act_v2_kernel<float>, dim3(l.grid), dim3(BS_THREADS), 0, (cudaStream_t)((cudaStream_t)stream), (const float*)x, x_ld, scale, shift, sc_per_sample, (const float*)res, res_ld, (float*)out, out_ld, HW, C, n_px, l.px_per_block, relu);
    else
      launch_pdl(affine_act_v2_kernel<__nv_bfloat16>, dim3(l.grid), dim3(BS_THREADS), 0, (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)x, x_ld, scale, shift, sc_per_sample, (const __nv_bfloat16*)res, res_ld, (__nv_bfloat16*)out, out_ld, HW, C, n_px, l.px_per_block, relu);
    ERF_CHECK_LAUNCH("erf_affine_act");
    return ERF_OK;
  }
  long long total4 = (long long)N * HW * (C / 4);
  unsigned grid = (unsigned)((total4 + 255) / 256);
  if (dtype == ERF_F32)
    affine_act_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, x_ld, scale, shift, sc_per_sample, (const float*)res, res_ld, (float*)out, out_ld, HW, C, total4, relu);
  else
    affine_act_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, x_ld, scale, shift, sc_per_sample, (const __nv_bfloat16*)res, res_ld, (__nv_bfloat16*)out, out_ld, HW, C, total4, relu);
  ERF_CHECK_LAUNCH("erf_affine_act");
  return ERF_OK;
}

extern "C" int erf_channel_stats(const void* x, int ld, int C, long long n_px, double* stats, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_channel_stats");
  ERF_CHECK_ARG(x && stats && n_px > 0 && C >= 8 && (C % 8) == 0 && C <= 2048 && (ld % 8) == 0, "erf_channel_stats: C and ld must be multiples of 8");
  if (bs_fast_ok(C) && ((uintptr_t)x % 16) == 0) {
    const BsLaunch l = bs_launch(C, n_px, 2);
    if (dtype == ERF_F32)
      launch_pdl(channel_reduce_v2_kernel<float, true, double>, dim3(l.grid), dim3(BS_THREADS), l.smem(C, 2), (cudaStream_t)((cudaStream_t)stream), (const float*)x, ld, C, n_px, l.px_per_block, stats);
    else
      launch_pdl(channel_reduce_v2_kernel<__nv_bfloat16, true, double>, dim3(l.grid), dim3(BS_THREADS), l.smem(C, 2), (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)x, ld, C, n_px, l.px_per_block, stats);
    ERF_CHECK_LAUNCH("erf_channel_stats");
    return ERF_OK;
  }
  long long want = 148 * 2;
  long long ppb = (n_px + want - 1) / want;
  if (ppb < 128) ppb = 128;
  unsigned grid = (unsigned)((n_px + ppb - 1) / ppb);
  size_t smem = 2 * C * sizeof(float);
  if (dtype == ERF_F32)
    channel_stats_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((const float*)x, ld, C, n_px, (int)ppb, stats);
  else
    channel_stats_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, C, n_px, (int)ppb, stats);
  ERF_CHECK_LAUNCH("erf_channel_stats");
  return ERF_OK;
}

extern "C" int erf_relu_bwd(const void* g, int g_ld, const void* mask, int mask_ld, void* out, int out_ld, long long n_px, int C, int dtype,
                            erf_stream_t stream) {
  CHECK_DTYPE("erf_relu_bwd");
  ERF_CHECK_ARG(g && mask && out && n_px > 0 && C > 0 && (C % 4) == 0 && (g_ld % 4) == 0 && (mask_ld % 4) == 0 && (out_ld % 4) == 0,
                "erf_relu_bwd: C and leading dimensions must be multiples of 4");
  long long total4 = n_px * (C / 4);
  unsigned grid = (unsigned)((total4 + 255) / 256);
  if (dtype == ERF_F32)
    relu_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)g, g_ld, (const float*)mask, mask_ld, (float*)out, out_ld, C, total4);
  else
    relu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)mask, mask_ld, (__nv_bfloat16*)out, out_ld, C, total4);
  ERF_CHECK_LAUNCH("erf_relu_bwd");
  return ERF_OK;
}

extern "C" int erf_bn_finalize(double* stats, double count, const float* gamma, const float* beta, float* running_mean, float* running_var,
                               float momentum, float eps, const float* drop, int N, int C, float* mean, float* invstd, float* scale, float* shift,
                               erf_stream_t stream) {
  ERF_CHECK_ARG(stats && gamma && beta && mean && invstd && scale && shift && C > 0 && count > 0, "erf_bn_finalize: bad arguments");
  ERF_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "erf_bn_finalize: running_mean/var must be given together");
  launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)((cudaStream_t)stream), stats, count, gamma, beta, running_mean, running_var, momentum, eps, drop, N, C, mean,
                                                                        invstd, scale, shift);
  ERF_CHECK_LAUNCH("erf_bn_finalize");
  return ERF_OK;
}

extern "C" int erf_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, int C,
                                  float* scale, float* shift, erf_stream_t stream) {
  ERF_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && C > 0, "erf_bn_eval_affine: bad arguments");
  bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps, C, scale, shift);
  ERF_CHECK_LAUNCH("erf_bn_eval_affine");
  return ERF_OK;
}

extern "C" int erf_bn_bwd_reduce(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld, const float* drop,
                                 const float* mean, const float* invstd, int N, int HW, int C, double* sums, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_bn_bwd_reduce");
  ERF_CHECK_ARG(g && y && mean && invstd && sums && N > 0 && HW > 0 && C >= 8 && (C % 8) == 0 && C <= 2048 && (g_ld % 8) == 0 && (y_ld % 8) == 0 &&
                    (!mask || (mask_ld % 8) == 0),
                "erf_bn_bwd_reduce: C and leading dimensions must be multiples of 8");
  long long n_px = (long long)N * HW;
  if (bs_fast_ok(C) && ((uintptr_t)g % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)mask % 16) == 0) {
    const BsLaunch l = bs_launch(C, n_px, 2);
    if (dtype == ERF_F32)
      launch_pdl(bn_bwd_reduce_v2_kernel<float>, dim3(l.grid), dim3(BS_THREADS), l.smem(C, 2) + 2 * C * sizeof(float), (cudaStream_t)((cudaStream_t)stream), (const float*)g, g_ld, (const float*)y, y_ld, (const float*)mask, mask_ld, drop, mean, invstd, HW, C, n_px, l.px_per_block, sums);
    else
      launch_pdl(bn_bwd_reduce_v2_kernel<__nv_bfloat16>, dim3(l.grid), dim3(BS_THREADS), l.smem(C, 2) + 2 * C * sizeof(float), (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)y, y_ld, (const __nv_bfloat16*)mask, mask_ld, drop, mean, invstd, HW, C, n_px, l.px_per_block, sums);
    ERF_CHECK_LAUNCH("erf_bn_bwd_reduce");
    return ERF_OK;
  }
  long long want = 148 * 2;
  const int thr = 256;
  long long ppb = (n_px + want - 1) / want;
  if (ppb < 64) ppb = 64;
  unsigned grid = (unsigned)((n_px + ppb - 1) / ppb);
  size_t smem = 2 * C * sizeof(float);
  if (dtype == ERF_F32)
    bn_bwd_reduce_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((const float*)g, g_ld, (const float*)y, y_ld, (const float*)mask, mask_ld, drop, mean, invstd, HW, C, n_px, (int)ppb, sums);
  else
    bn_bwd_reduce_kernel<__nv_bfloat16><<<grid, thr, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)y, y_ld, (const __nv_bfloat16*)mask, mask_ld, drop, mean, invstd, HW, C, n_px, (int)ppb, sums);
  ERF_CHECK_LAUNCH("erf_bn_bwd_reduce");
  return ERF_OK;
}

extern "C" int erf_bn_bwd_apply(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld, const float* drop,
                                const float* gamma, const float* mean, const float* invstd, const double* sums, int N, int HW, int C, void* dy,
                                int dy_ld, float* dgamma, float* dbeta, void* gmasked, int gmasked_ld, int dtype, erf_stream_t stream) {
  CHECK_DTYPE("erf_bn_bwd_apply");
  ERF_CHECK_ARG(!gmasked || (gmasked_ld % 8) == 0, "erf_bn_bwd_apply: gmasked_ld must be a multiple of 8");
  ERF_CHECK_ARG(g && y && gamma && mean && invstd && sums && dy && N > 0 && HW > 0 && C >= 8 && (C % 8) == 0 && (g_ld % 8) == 0 && (y_ld % 8) == 0 &&
                    (dy_ld % 8) == 0 && (!mask || (mask_ld % 8) == 0),
                "erf_bn_bwd_apply: C and leading dimensions must be multiples of 8");
  long long n_px = (long long)N * HW;
  if (bs_fast_ok(C) && ((uintptr_t)g % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)mask % 16) == 0 && ((uintptr_t)dy % 16) == 0 &&
      ((uintptr_t)gmasked % 16) == 0) {
    const BsLaunch l = bs_launch(C, n_px, 4);
    if (dtype == ERF_F32)
      launch_pdl(bn_bwd_apply_v2_kernel<float>, dim3(l.grid), dim3(BS_THREADS), 4 * C * sizeof(float), (cudaStream_t)((cudaStream_t)stream), (const float*)g, g_ld, (const float*)y, y_ld, (const float*)mask, mask_ld, drop, gamma, mean, invstd, sums, HW, C, n_px, l.px_per_block, (float*)dy, dy_ld, dgamma, dbeta, (float*)gmasked, gmasked_ld);
    else
      launch_pdl(bn_bwd_apply_v2_kernel<__nv_bfloat16>, dim3(l.grid), dim3(BS_THREADS), 4 * C * sizeof(float), (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)y, y_ld, (const __nv_bfloat16*)mask, mask_ld, drop, gamma, mean, invstd, sums, HW, C, n_px, l.px_per_block, (__nv_bfloat16*)dy, dy_ld, dgamma, dbeta, (__nv_bfloat16*)gmasked, gmasked_ld);
    ERF_CHECK_LAUNCH("erf_bn_bwd_apply");
    return ERF_OK;
  }
  long long work = n_px * (C / 8);
  if (work < C) work = C;
  unsigned grid = (unsigned)((work + 255) / 256);
  if (dtype == ERF_F32)
    bn_bwd_apply_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)g, g_ld, (const float*)y, y_ld, (const float*)mask, mask_ld, drop, gamma, mean, invstd, sums, HW, C, n_px, (float*)dy, dy_ld, dgamma, dbeta, (float*)gmasked, gmasked_ld);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)g, g_ld, (const __nv_bfloat16*)y, y_ld, (const __nv_bfloat16*)mask, mask_ld, drop, gamma, mean, invstd, sums, HW, C, n_px, (__nv_bfloat16*)dy, dy_ld, dgamma, dbeta, (__nv_bfloat16*)gmasked, gmasked_ld);
  ERF_CHECK_LAUNCH("erf_bn_bwd_apply");
  return ERF_OK;
}

extern "C" int erf_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                             float weight_decay, int step, float grad_scale, erf_stream_t stream) {
  ERF_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "erf_adam_step: bad arguments");
  double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  unsigned grid = (unsigned)((n + 255) / 256);
  launch_pdl(adam_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p, g, m, v, n, (float)(lr / bc1), beta1, beta2, eps, weight_decay, (float)(1.0 / sqrt(bc2)), grad_scale,
                                                      nullptr, lr, nullptr);
  ERF_CHECK_LAUNCH("erf_adam_step");
  return ERF_OK;
}

extern "C" int erf_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, int* step_dev, float grad_scale, erf_stream_t stream) {
  ERF_CHECK_ARG(p && g && m && v && n > 0 && step_dev, "erf_adam_step_dev: bad arguments");
  unsigned grid = (unsigned)((n + 255) / 256);
  launch_pdl(adam_step_inc_kernel, dim3(1), dim3(1), 0, (cudaStream_t)((cudaStream_t)stream), step_dev);
  launch_pdl(adam_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p, g, m, v, n, 0.f, beta1, beta2, eps, weight_decay, 0.f, grad_scale, step_dev, lr,
             nullptr);
  ERF_CHECK_LAUNCH("erf_adam_step_dev");
  return ERF_OK;
}

extern "C" int erf_adam_step_graph(float* p, const float* g, float* m, float* v, long long n, const float* hyper_dev, float beta1, float beta2, float eps,
                                   int* step_dev, float grad_scale, erf_stream_t stream) {
  ERF_CHECK_ARG(p && g && m && v && n > 0 && step_dev && hyper_dev, "erf_adam_step_graph: bad arguments");
  unsigned grid = (unsigned)((n + 255) / 256);
  launch_pdl(adam_step_inc_kernel, dim3(1), dim3(1), 0, (cudaStream_t)((cudaStream_t)stream), step_dev);
  launch_pdl(adam_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), p, g, m, v, n, 0.f, beta1, beta2, eps, 0.f, 0.f, grad_scale, step_dev, 0.f,
             hyper_dev);
  ERF_CHECK_LAUNCH("erf_adam_step_graph");
  return ERF_OK;
}

extern "C" int erf_pack_conv1d_tc_batched(const erf_pack1d_entry* table_dev, int n_entries, int max_C, erf_stream_t stream) {
  ERF_CHECK_ARG(table_dev && n_entries > 0 && n_entries <= 65535 && (max_C == 16 || max_C == 64 || max_C == 128),
                "erf_pack_conv1d_tc_batched: bad arguments (n=%d, max_C=%d)", n_entries, max_C);
  int total = 2 * 3 * max_C * (max_C / 8);
  launch_pdl(pack_conv1d_tc_batched_kernel, dim3(dim3((total + 255) / 256, n_entries)), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), table_dev);
  ERF_CHECK_LAUNCH("erf_pack_conv1d_tc_batched");
  return ERF_OK;
}

extern "C" int erf_pack_conv1d_tc(const float* w, int C, void* img_fwd, void* img_dgrad, erf_stream_t stream) {
  ERF_CHECK_ARG(w && img_fwd && img_dgrad && (C == 16 || C == 64 || C == 128), "erf_pack_conv1d_tc: bad arguments (C=%d)", C);
  int total = 2 * 3 * C * (C / 8);
  pack_conv1d_tc_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, C, (__nv_bfloat16*)img_fwd, (__nv_bfloat16*)img_dgrad);
  ERF_CHECK_LAUNCH("erf_pack_conv1d_tc");
  return ERF_OK;
}
