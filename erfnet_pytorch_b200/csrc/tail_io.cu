// Kernels either side of the conv stack: the ImageNet classification tail (AvgPool2d, Linear, row-wise
// cross entropy + top-k ranks, SGD with momentum), the mean-teacher weight average, and the I/O side (uint8 HWC ingest, label
// look-up table, label colouring).  All are HBM-bound element / row kernels: coalesced, vectorised where the layout allows, grids sized
// in multiples of the SM count.
#include "common.cuh"

namespace erfk {

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int grid_for(long long work, int block, int per_sm = 8) {
  long long g = (work + block - 1) / block;
  long long cap = (long long)sm_count() * per_sm;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- AvgPool2d(k, stride 1, padding 0) over NHWC (imagenet/erfnet_imagenet.py:93) ------------------------------------------------
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ in, int N, int H, int W, int ld, int C, int k, float* __restrict__ out) {
  const int Ho = H - k + 1, Wo = W - k + 1;
  const long long total = (long long)N * Ho * Wo * C;
  const float inv = 1.f / (float)(k * k);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho), n = (int)(r / Ho);
    float s = 0.f;
    for (int dy = 0; dy < k; ++dy) {
      const T* row = in + (((size_t)n * H + ho + dy) * W + wo) * ld + c;
      for (int dx = 0; dx < k; ++dx) s += to_f(row[(size_t)dx * ld]);
    }
    out[i] = s * inv;
  }
}

template <typename T>
__global__ void avgpool_bwd_kernel(const float* __restrict__ dout, int N, int H, int W, int ld, int C, int k, T* __restrict__ dx) {
  const int Ho = H - k + 1, Wo = W - k + 1;
  const long long total = (long long)N * H * W * ld;
  const float inv = 1.f / (float)(k * k);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    long long r = i / ld;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H), n = (int)(r / H);
    float s = 0.f;
    if (c < C) {
      const int h0 = max(0, h - k + 1), h1 = min(Ho - 1, h), w0 = max(0, w - k + 1), w1 = min(Wo - 1, w);
      for (int ho = h0; ho <= h1; ++ho)
        for (int wo = w0; wo <= w1; ++wo) s += dout[(((size_t)n * Ho + ho) * Wo + wo) * C + c];
    }
    dx[i] = from_f<T>(s * inv);
  }
}

// ---- small fp32 GEMM for Linear(128, num_classes) (imagenet/erfnet_imagenet.py:99-107): C[M,N] (+)= op(A)[M,K] op(B)[K,N] (+ bias[n]) ----
// op(A)[m,k] = TA ? A[k*lda+m] : A[m*lda+k];  op(B)[k,n] = TB ? B[n*ldb+k] : B[k*ldb+n].  32x32 tiles, 256 threads, 4 outputs each.
template <bool TA, bool TB>
__global__ void gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ Cm, int ldc,
                                int M, int N, int K, const float* __restrict__ bias, int accumulate) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = ty + 8 * r;
      // sa[m][k]
      if (TA) {   // consecutive tx -> consecutive m (contiguous in A)
        const int m = m0 + tx, k = k0 + row;
        sa[tx][row] = (m < M && k < K) ? A[(size_t)k * lda + m] : 0.f;
      } else {
        const int m = m0 + row, k = k0 + tx;
        sa[row][tx] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
      }
      // sb[k][n]
      if (TB) {   // consecutive tx -> consecutive k (contiguous in B)
        const int n = n0 + row, k = k0 + tx;
        sb[tx][row] = (n < N && k < K) ? B[(size_t)n * ldb + k] : 0.f;
      } else {
        const int k = k0 + row, n = n0 + tx;
        sb[row][tx] = (n < N && k < K) ? B[(size_t)k * ldb + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float b = sb[kk][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sa[ty + 8 * r][kk], b, acc[r]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < N) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = m0 + ty + 8 * r;
      if (m < M) {
        float v = acc[r] + (bias ? bias[n] : 0.f);
        float* dst = Cm + (size_t)m * ldc + n;
        *dst = accumulate ? *dst + v : v;
      }
    }
  }
}

// out[n] += sum_m x[m,n]
__global__ void colsum_f32_kernel(const float* __restrict__ x, int M, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += x[(size_t)m * N + n];
  out[n] += s;
}

// ---- row-wise cross entropy of nn.CrossEntropyLoss() (imagenet/main.py:80, 176) + rank of the target logit (accuracy(), l.288-301) ----
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  v = s[0];
  for (int i = 1; i < nw; ++i) v = is_max ? fmaxf(v, s[i]) : v + s[i];
  return v;
}

// one block per row: lse[n] = log sum exp, rank[n] = #{c : l[c] > l[t]  or (l[c] == l[t] and c < t)} (-1 for ignored rows),
// acc[0] += nll, acc[1] += 1 for rows whose target is valid.
__global__ void ce_rows_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int C, long long ignore,
                                   float* __restrict__ lse, int* __restrict__ rank, float* __restrict__ acc) {
  __shared__ float s[32];
  const int n = blockIdx.x;
  const float* row = logits + (size_t)n * C;
  const long long t = target[n];
  const bool valid = t != ignore && t >= 0 && t < C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = block_reduce(mx, true, s);
  const float lt = valid ? row[t] : 0.f;
  float se = 0.f, rk = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = row[c];
    se += __expf(v - mx);
    if (valid && (v > lt || (v == lt && c < t))) rk += 1.f;
  }
  se = block_reduce(se, false, s);
  rk = block_reduce(rk, false, s);
  if (threadIdx.x == 0) {
    const float l = mx + __logf(se);
    lse[n] = l;
    rank[n] = valid ? (int)rk : -1;
    if (valid) {
      atomicAdd(acc, l - lt);
      atomicAdd(acc + 1, 1.f);
    }
  }
}

// dlogits[n,c] = go * (softmax - onehot) / #valid  (0 for ignored rows)
__global__ void ce_rows_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, int C, long long ignore,
                                   const float* __restrict__ lse, const float* __restrict__ acc, const float* __restrict__ go,
                                   float* __restrict__ dlogits, long long total) {
  const float k = (go ? *go : 1.f) / fmaxf(acc[1], 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / C), c = (int)(i % C);
    const long long t = target[n];
    const bool valid = t != ignore && t >= 0 && t < C;
    dlogits[i] = valid ? k * (__expf(logits[i] - lse[n]) - (c == t ? 1.f : 0.f)) : 0.f;
  }
}

// ---- torch.optim.SGD(lr, momentum, weight_decay) (imagenet/main.py:83-85) over one flat fp32 buffer ---------------------------------
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n, float lr,
                                float momentum, float wd, int first, float grad_scale, const float* __restrict__ hyper_dev) {
  if (hyper_dev) {   // CUDA-graph replay: lr / weight decay are read from device memory (see erf_sgd_step_graph)
    lr = hyper_dev[0];
    wd = hyper_dev[1];
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    float d = fmaf(wd, w, g[i] * grad_scale);
    if (momentum != 0.f) {
      d = first ? d : fmaf(momentum, buf[i], d);
      buf[i] = d;
    }
    p[i] = w - lr * d;
  }
}

// ---- mean teacher (train/main.py:140-142): teacher = teacher * (1 - rate) + rate * student ----------------------------------------------
__global__ void ema_update_kernel(float* __restrict__ teacher, const float* __restrict__ student, long long n, float rate) {
  const float keep = 1.f - rate;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    teacher[i] = teacher[i] * keep + rate * student[i];
}

// ---- ingest: uint8 HWC image (what PIL / the camera delivers) -> NHWC activations with pitch ld >= 3 ----------------------------------
// = torchvision ToTensor (u8 / 255, train/dataset.py:113-114) [+ Normalize(mean, std)] + the model's NCHW fp32 -> NHWC conversion, in one pass.
template <typename T>
__global__ void ingest_u8_kernel(const uint8_t* __restrict__ in, long long n_px, int ld, const float* __restrict__ mean,
                                 const float* __restrict__ stdv, T* __restrict__ out) {
  float m[3] = {0.f, 0.f, 0.f}, s[3] = {1.f, 1.f, 1.f};
  if (mean) { m[0] = mean[0]; m[1] = mean[1]; m[2] = mean[2]; }
  if (stdv) { s[0] = stdv[0]; s[1] = stdv[1]; s[2] = stdv[2]; }
  const long long groups = n_px >> 2;   // 4 pixels = 12 bytes = 3 aligned 32-bit loads
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long long)gridDim.x * blockDim.x) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + gi * 3;
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
    const uint32_t b[12] = {w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, w0 >> 24, w1 & 255, (w1 >> 8) & 255, (w1 >> 16) & 255, w1 >> 24,
                            w2 & 255, (w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24};
    float v[4][4];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v[px][c] = __fdiv_rn((float)b[px * 3 + c], 255.f);
        if (mean || stdv) v[px][c] = __fdiv_rn(v[px][c] - m[c], s[c]);
      }
      v[px][3] = 0.f;
    }
    if (ld == 4) {   // the models' input pitch: the thread's 4 pixels are 32 (bf16) / 64 (fp32) contiguous bytes -> 16-byte stores
      T* dst = out + (size_t)gi * 16;
      if (sizeof(T) == 2) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          Vec<T, 8> o;
#pragma unroll
          for (int e = 0; e < 8; ++e) o.v[e] = v[2 * h + (e >> 2)][e & 3];
          o.store(dst + 8 * h);
        }
      } else {
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          Vec<T, 4> o;
#pragma unroll
          for (int e = 0; e < 4; ++e) o.v[e] = v[px][e];
          o.store(dst + 4 * px);
        }
      }
    } else {
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        T* dst = out + (size_t)(gi * 4 + px) * ld;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c] = from_f<T>(v[px][c]);
        for (int c = 3; c < ld; ++c) dst[c] = from_f<T>(0.f);
      }
    }
  }
  // tail pixels (n_px % 4)
  for (long long p = (groups << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (long long)gridDim.x * blockDim.x) {
    T* dst = out + (size_t)p * ld;
    for (int c = 0; c < 3; ++c) {
      float v = __fdiv_rn((float)in[p * 3 + c], 255.f);
      if (mean || stdv) v = __fdiv_rn(v - m[c], s[c]);
      dst[c] = from_f<T>(v);
    }
    for (int c = 3; c < ld; ++c) dst[c] = from_f<T>(0.f);
  }
}

// ---- labels: look-up table (ToLabel + Relabel, eval/transform.py:55-70; trainId -> id, eval_cityscapes_server.py:39-63) and colouring ----
template <typename TI>
__global__ void label_lut_kernel(const TI* __restrict__ in, long long n, const int64_t* __restrict__ lut, int lut_n, int64_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long v = (long long)in[i];
    out[i] = (v >= 0 && v < lut_n) ? lut[v] : v;
  }
}

// Colorize (eval/transform.py:73-93): out[n, ch, p] = cmap[label][ch] for label < n_colors, else 0; out is uint8 [N,3,H*W]
template <typename TI>
__global__ void colorize_kernel(const TI* __restrict__ labels, long long n_img, long long hw, const uint8_t* __restrict__ cmap, int n_colors,
                                uint8_t* __restrict__ out) {
  const long long total = n_img * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / hw, p = i % hw;
    const long long v = (long long)labels[i];
    uint8_t r = 0, g = 0, b = 0;
    if (v >= 0 && v < n_colors) { r = cmap[v * 3]; g = cmap[v * 3 + 1]; b = cmap[v * 3 + 2]; }
    uint8_t* o = out + img * 3 * hw + p;
    o[0] = r; o[hw] = g; o[2 * hw] = b;
  }
}

}  // namespace erfk

using namespace erfk;

extern "C" int erf_avgpool_fwd(const void* in, int N, int H, int W, int ld, int C, int k, float* out, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(in && out, "erf_avgpool_fwd: null pointer");
  ERF_CHECK_ARG(N > 0 && C > 0 && ld >= C && k > 0 && H >= k && W >= k, "erf_avgpool_fwd: bad shape N=%d H=%d W=%d C=%d ld=%d k=%d", N, H, W, C, ld, k);
  const long long total = (long long)N * (H - k + 1) * (W - k + 1) * C;
  const int g = grid_for(total, 256);
  if (dtype == ERF_F32) avgpool_fwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)in, N, H, W, ld, C, k, out);
  else if (dtype == ERF_BF16) avgpool_fwd_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, N, H, W, ld, C, k, out);
  else { set_error("erf_avgpool_fwd: bad dtype %d", dtype); return ERF_ERR_INVALID_ARGUMENT; }
  ERF_CHECK_LAUNCH("erf_avgpool_fwd");
  return ERF_OK;
}

extern "C" int erf_avgpool_bwd(const float* dout, int N, int H, int W, int ld, int C, int k, void* dx, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(dout && dx, "erf_avgpool_bwd: null pointer");
  ERF_CHECK_ARG(N > 0 && C > 0 && ld >= C && k > 0 && H >= k && W >= k, "erf_avgpool_bwd: bad shape N=%d H=%d W=%d C=%d ld=%d k=%d", N, H, W, C, ld, k);
  const long long total = (long long)N * H * W * ld;
  const int g = grid_for(total, 256);
  if (dtype == ERF_F32) avgpool_bwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(dout, N, H, W, ld, C, k, (float*)dx);
  else if (dtype == ERF_BF16) avgpool_bwd_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>(dout, N, H, W, ld, C, k, (__nv_bfloat16*)dx);
  else { set_error("erf_avgpool_bwd: bad dtype %d", dtype); return ERF_ERR_INVALID_ARGUMENT; }
  ERF_CHECK_LAUNCH("erf_avgpool_bwd");
  return ERF_OK;
}

static int gemm_f32(bool ta, bool tb, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                    const float* bias, int accumulate, cudaStream_t st) {
  dim3 grid(ceil_div(N, 32), ceil_div(M, 32));
  if (!ta && tb) gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate);
  else if (!ta && !tb) gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate);
  else if (ta && !tb) gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate);
  else gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, accumulate);
  return 0;
}

extern "C" int erf_linear_fwd(const float* x, const float* w, const float* bias, float* out, int M, int N, int K, erf_stream_t stream) {
  ERF_CHECK_ARG(x && w && out && M > 0 && N > 0 && K > 0, "erf_linear_fwd: bad arguments M=%d N=%d K=%d", M, N, K);
  gemm_f32(false, true, x, K, w, K, out, N, M, N, K, bias, 0, (cudaStream_t)stream);
  ERF_CHECK_LAUNCH("erf_linear_fwd");
  return ERF_OK;
}

extern "C" int erf_linear_bwd(const float* x, const float* w, const float* dout, float* dx, float* dw, float* db, int M, int N, int K,
                              erf_stream_t stream) {
  ERF_CHECK_ARG(x && w && dout && M > 0 && N > 0 && K > 0, "erf_linear_bwd: bad arguments M=%d N=%d K=%d", M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) gemm_f32(false, false, dout, N, w, K, dx, K, M, K, N, nullptr, 0, st);          // dx[m,k] = sum_n dout[m,n] w[n,k]
  if (dw) gemm_f32(true, false, dout, N, x, K, dw, K, N, K, M, nullptr, 1, st);           // dw[n,k] += sum_m dout[m,n] x[m,k]
  if (db) colsum_f32_kernel<<<ceil_div(N, 128), 128, 0, st>>>(dout, M, N, db);            // db[n] += sum_m dout[m,n]
  ERF_CHECK_LAUNCH("erf_linear_bwd");
  return ERF_OK;
}

extern "C" int erf_ce_rows_fwd(const float* logits, const long long* target, int N, int C, long long ignore_index, float* lse, int* rank,
                               float* acc, erf_stream_t stream) {
  ERF_CHECK_ARG(logits && target && lse && rank && acc && N > 0 && C > 0, "erf_ce_rows_fwd: bad arguments N=%d C=%d", N, C);
  ce_rows_fwd_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(logits, (const int64_t*)target, C, ignore_index, lse, rank, acc);
  ERF_CHECK_LAUNCH("erf_ce_rows_fwd");
  return ERF_OK;
}

extern "C" int erf_ce_rows_bwd(const float* logits, const long long* target, int N, int C, long long ignore_index, const float* lse,
                               const float* acc, const float* go, float* dlogits, erf_stream_t stream) {
  ERF_CHECK_ARG(logits && target && lse && acc && dlogits && N > 0 && C > 0, "erf_ce_rows_bwd: bad arguments N=%d C=%d", N, C);
  const long long total = (long long)N * C;
  ce_rows_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(logits, (const int64_t*)target, C, ignore_index, lse, acc, go, dlogits, total);
  ERF_CHECK_LAUNCH("erf_ce_rows_bwd");
  return ERF_OK;
}

extern "C" int erf_sgd_step(float* p, const float* g, float* buf, long long n, float lr, float momentum, float weight_decay, int first_step,
                            float grad_scale, erf_stream_t stream) {
  ERF_CHECK_ARG(p && g && n > 0 && (momentum == 0.f || buf), "erf_sgd_step: bad arguments");
  sgd_step_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, buf, n, lr, momentum, weight_decay, first_step, grad_scale, nullptr);
  ERF_CHECK_LAUNCH("erf_sgd_step");
  return ERF_OK;
}

extern "C" int erf_sgd_step_graph(float* p, const float* g, float* buf, long long n, const float* hyper_dev, float momentum, float grad_scale,
                                  erf_stream_t stream) {
  ERF_CHECK_ARG(p && g && n > 0 && hyper_dev && (momentum == 0.f || buf), "erf_sgd_step_graph: bad arguments");
  sgd_step_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, g, buf, n, 0.f, momentum, 0.f, 0, grad_scale, hyper_dev);
  ERF_CHECK_LAUNCH("erf_sgd_step_graph");
  return ERF_OK;
}

extern "C" int erf_ema_update(float* teacher, const float* student, long long n, float rate, erf_stream_t stream) {
  ERF_CHECK_ARG(teacher && student && n > 0, "erf_ema_update: bad arguments");
  ema_update_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(teacher, student, n, rate);
  ERF_CHECK_LAUNCH("erf_ema_update");
  return ERF_OK;
}

extern "C" int erf_ingest_u8(const void* in, long long n_px, int ld, const float* mean, const float* stdv, void* out, int dtype,
                             erf_stream_t stream) {
  ERF_CHECK_ARG(in && out && n_px > 0 && ld >= 3, "erf_ingest_u8: bad arguments n_px=%lld ld=%d", n_px, ld);
  ERF_CHECK_ARG(((uintptr_t)in & 3) == 0 && ((uintptr_t)out & 15) == 0, "erf_ingest_u8: the uint8 image must be 4-byte aligned and the output 16-byte aligned");
  const int g = grid_for((n_px + 3) / 4, 256);
  if (dtype == ERF_F32) ingest_u8_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in, n_px, ld, mean, stdv, (float*)out);
  else if (dtype == ERF_BF16) ingest_u8_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in, n_px, ld, mean, stdv, (__nv_bfloat16*)out);
  else { set_error("erf_ingest_u8: bad dtype %d", dtype); return ERF_ERR_INVALID_ARGUMENT; }
  ERF_CHECK_LAUNCH("erf_ingest_u8");
  return ERF_OK;
}

extern "C" int erf_label_lut(const void* in, int elem_bytes, long long n, const long long* lut, int lut_n, long long* out, erf_stream_t stream) {
  ERF_CHECK_ARG(in && lut && out && n > 0 && lut_n > 0, "erf_label_lut: bad arguments");
  ERF_CHECK_ARG(elem_bytes == 1 || elem_bytes == 8, "erf_label_lut: labels must be uint8 or int64 (elem_bytes=%d)", elem_bytes);
  const int g = grid_for(n, 256);
  if (elem_bytes == 1) label_lut_kernel<uint8_t><<<g, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)in, n, (const int64_t*)lut, lut_n, (int64_t*)out);
  else label_lut_kernel<int64_t><<<g, 256, 0, (cudaStream_t)stream>>>((const int64_t*)in, n, (const int64_t*)lut, lut_n, (int64_t*)out);
  ERF_CHECK_LAUNCH("erf_label_lut");
  return ERF_OK;
}

extern "C" int erf_colorize(const void* labels, int elem_bytes, long long n_img, long long hw, const unsigned char* cmap, int n_colors,
                            unsigned char* out, erf_stream_t stream) {
  ERF_CHECK_ARG(labels && cmap && out && n_img > 0 && hw > 0 && n_colors > 0, "erf_colorize: bad arguments");
  ERF_CHECK_ARG(elem_bytes == 1 || elem_bytes == 8, "erf_colorize: labels must be uint8 or int64 (elem_bytes=%d)", elem_bytes);
  const int g = grid_for(n_img * hw, 256);
  if (elem_bytes == 1) colorize_kernel<uint8_t><<<g, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)labels, n_img, hw, cmap, n_colors, out);
  else colorize_kernel<int64_t><<<g, 256, 0, (cudaStream_t)stream>>>((const int64_t*)labels, n_img, hw, cmap, n_colors, out);
  ERF_CHECK_LAUNCH("erf_colorize");
  return ERF_OK;
}
