// Streaming BatchNorm / per-channel kernels, second generation: a thread owns ONE group of 8 channels for its whole life and walks
// a contiguous pixel range, so per-channel coefficients live in registers (the first generation re-fetched ~200 bytes of
// parameters per 48 bytes of data), loads are issued two pixels ahead of their use, and the per-channel reductions are combined
// with warp shuffles + one shared-memory row per warp instead of same-address shared atomics (128-way conflicts at C = 16).
// Fast-path conditions (checked by the launchers): C % 8 == 0, 32 % (C/8) == 0, all leading dimensions % 8 == 0 (16-byte vectors).
#pragma once
#include "common.cuh"

namespace erfk {

constexpr int BS_THREADS = 512;

__device__ __forceinline__ long long img_of(long long q, long long HW) {
  return (q < 0x7fffffffLL) ? (long long)((unsigned)q / (unsigned)HW) : q / HW;   // 32-bit divide for any realistic tensor
}

struct PxRange {
  int c0, lane_px, lanes;
  long long p0, p1;
};

__device__ __forceinline__ PxRange px_range(int C, long long n_px, int px_per_block) {
  PxRange r;
  const int c8 = C >> 3;
  r.c0 = (threadIdx.x % c8) * 8;
  r.lane_px = threadIdx.x / c8;
  r.lanes = blockDim.x / c8;
  r.p0 = (long long)blockIdx.x * px_per_block;
  r.p1 = r.p0 + px_per_block;
  if (r.p1 > n_px) r.p1 = n_px;
  if (r.lane_px >= r.lanes) r.p1 = r.p0;   // threads beyond the last whole pixel lane (blockDim % c8 != 0) stay idle
  return r;
}

// Sum acc[v][k] over all threads of the block that share a channel group, then hand row i (< NV*C) of the block total to f(i, total).
template <int NV, typename F>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NV][8], int C, int c0, float* s_part /* [warps][NV*C] */, F&& f) {
  const int c8 = C >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  if (32 % c8 != 0) {
    // channel groups do not tile a warp (e.g. the 24-channel padded logits): per-thread rows in shared memory [threads][NV*8],
    // then thread i sums the rows of the threads that own channel i's group
    const int lanes = blockDim.x / c8;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) s_part[(size_t)threadIdx.x * NV * 8 + v * 8 + k] = acc[v][k];
    __syncthreads();
    for (int i = threadIdx.x; i < NV * C; i += blockDim.x) {
      const int v = i / C, c = i % C, g = c >> 3, k = c & 7;
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += s_part[(size_t)(l * c8 + g) * NV * 8 + v * 8 + k];
      f(i, t);
    }
    return;
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      for (int off = c8; off < 32; off <<= 1) acc[v][k] += __shfl_xor_sync(0xffffffffu, acc[v][k], off);
  if (lane < c8) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int k = 0; k < 8; ++k) s_part[(size_t)warp * NV * C + v * C + c0 + k] = acc[v][k];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV * C; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < warps; ++w) t += s_part[(size_t)w * NV * C + i];
    f(i, t);
  }
}

// 8 consecutive elements kept PACKED in registers (4 for bf16, 8 for fp32) until they are used
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 q;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { q = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = reinterpret_cast<const float4*>(p)[0]; b = reinterpret_cast<const float4*>(p)[1]; }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
};

__device__ __forceinline__ void lds8(const float* s, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(s)[0], b = reinterpret_cast<const float4*>(s)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

constexpr int UNR = 2;   // pixels in flight per thread

// ---- statistics of a stored activation: stats[c] += sum x, stats[C+c] += sum x^2 (SQ) or out[c] += sum x (!SQ) --------------------
template <typename T, bool SQ, typename OUT>
__global__ void __launch_bounds__(BS_THREADS, 2) channel_reduce_v2_kernel(const T* __restrict__ x, int ld, int C, long long n_px, int px_per_block,
                                                                          OUT* __restrict__ out) {
  pdl_enter();
  extern __shared__ __align__(16) float s_dyn[];
  const PxRange r = px_range(C, n_px, px_per_block);
  float acc[SQ ? 2 : 1][8];
#pragma unroll
  for (int v = 0; v < (SQ ? 2 : 1); ++v)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[v][k] = 0.f;
  for (long long p = r.p0 + r.lane_px; p < r.p1; p += 4 * r.lanes) {
    Raw8<T> v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * r.lanes < r.p1) v[u].load(x + (size_t)(p + u * r.lanes) * ld + r.c0);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + u * r.lanes < r.p1) {
        float f[8];
        v[u].unpack(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[0][k] += f[k];
          if (SQ) acc[SQ ? 1 : 0][k] += f[k] * f[k];
        }
      }
  }
  block_channel_reduce<SQ ? 2 : 1>(acc, C, r.c0, s_dyn, [&](int i, float t) { atomicAdd(&out[i], (OUT)t); });
}

// ---- BN backward, pass 1: gb = g * (mask > 0) * drop[n,c];  sums[c] += gb, sums[C+c] += gb * xhat -------------------------------
// dynamic smem: [2C] coefficients (mean, invstd) followed by the [warps][2C] reduction rows
template <typename T>
__global__ void __launch_bounds__(BS_THREADS, 2) bn_bwd_reduce_v2_kernel(const T* __restrict__ g, int g_ld, const T* __restrict__ y, int y_ld,
                                                                         const T* __restrict__ mask, int mask_ld, const float* __restrict__ drop,
                                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                         long long HW, int C, long long n_px, int px_per_block,
                                                                         double* __restrict__ sums) {
  pdl_enter();
  extern __shared__ __align__(16) float s_dyn[];
  float* s_coef = s_dyn;
  float* s_part = s_dyn + 2 * C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_coef[i] = mean[i]; s_coef[C + i] = invstd[i]; }
  __syncthreads();
  const PxRange r = px_range(C, n_px, px_per_block);
  float acc[2][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = 0.f;
  for (long long p = r.p0 + r.lane_px; p < r.p1; p += UNR * r.lanes) {
    Raw8<T> gv[UNR], yv[UNR], mv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        gv[u].load(g + (size_t)q * g_ld + r.c0);
        yv[u].load(y + (size_t)q * y_ld + r.c0);
        if (mask) mv[u].load(mask + (size_t)q * mask_ld + r.c0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        float gg[8], t[8];
        gv[u].unpack(gg);
        if (mask) {
          mv[u].unpack(t);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (!(t[k] > 0.f)) gg[k] = 0.f;
        }
        if (drop) {
          const float* dp = drop + img_of(q, HW) * C + r.c0;
#pragma unroll
          for (int k = 0; k < 8; ++k) gg[k] *= dp[k];
        }
        float mu[8], is[8];
        yv[u].unpack(t);
        lds8(s_coef + r.c0, mu);
        lds8(s_coef + C + r.c0, is);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          acc[0][k] += gg[k];
          acc[1][k] += gg[k] * (t[k] - mu[k]) * is[k];
        }
      }
    }
  }
  block_channel_reduce<2>(acc, C, r.c0, s_part, [&](int i, float t) { atomicAdd(&sums[i], (double)t); });
}

// ---- BN backward, pass 2: dy = k1*gb + k2*(y - mu) + k3 with per-channel coefficients in shared memory; optional gmasked = g*(mask>0)
template <typename T>
__global__ void __launch_bounds__(BS_THREADS, 2) bn_bwd_apply_v2_kernel(const T* __restrict__ g, int g_ld, const T* __restrict__ y, int y_ld,
                                                                        const T* __restrict__ mask, int mask_ld, const float* __restrict__ drop,
                                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                        const float* __restrict__ invstd, const double* __restrict__ sums,
                                                                        long long HW, int C, long long n_px, int px_per_block, T* __restrict__ dy,
                                                                        int dy_ld, float* dgamma, float* dbeta, T* __restrict__ gmasked,
                                                                        int gmasked_ld) {
  pdl_enter();
  extern __shared__ __align__(16) float s_dyn[];   // [4C]: k1, k2, k3, mu
  // dy = gamma*is*(gb - s1/M - xhat*s2/M),  xhat = (y - mu)*is   ==>   dy = k1*gb + k2*(y - mu) + k3
  const float inv_m = 1.f / (float)n_px;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float is = invstd[i], ga = gamma[i];
    const float s1 = (float)sums[i], s2 = (float)sums[C + i];
    s_dyn[i] = ga * is;
    s_dyn[C + i] = -ga * is * is * s2 * inv_m;
    s_dyn[2 * C + i] = -ga * is * s1 * inv_m;
    s_dyn[3 * C + i] = mean[i];
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[i] = s1;
      if (dgamma) dgamma[i] = s2;
    }
  }
  __syncthreads();
  const PxRange r = px_range(C, n_px, px_per_block);
  for (long long p = r.p0 + r.lane_px; p < r.p1; p += UNR * r.lanes) {
    Raw8<T> gv[UNR], yv[UNR], mv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        gv[u].load(g + (size_t)q * g_ld + r.c0);
        yv[u].load(y + (size_t)q * y_ld + r.c0);
        if (mask) mv[u].load(mask + (size_t)q * mask_ld + r.c0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        float gg[8], t[8], k[8];
        gv[u].unpack(gg);
        if (mask) {
          mv[u].unpack(t);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (!(t[e] > 0.f)) gg[e] = 0.f;
        }
        if (gmasked) Raw8<T>::store(gmasked + (size_t)q * gmasked_ld + r.c0, gg);
        if (drop) {
          const float* dp = drop + img_of(q, HW) * C + r.c0;
#pragma unroll
          for (int e = 0; e < 8; ++e) gg[e] *= dp[e];
        }
        lds8(s_dyn + r.c0, k);
#pragma unroll
        for (int e = 0; e < 8; ++e) gg[e] *= k[e];
        yv[u].unpack(t);
        lds8(s_dyn + 3 * C + r.c0, k);
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] -= k[e];
        lds8(s_dyn + C + r.c0, k);
#pragma unroll
        for (int e = 0; e < 8; ++e) gg[e] += t[e] * k[e];
        lds8(s_dyn + 2 * C + r.c0, k);
#pragma unroll
        for (int e = 0; e < 8; ++e) gg[e] += k[e];
        Raw8<T>::store(dy + (size_t)q * dy_ld + r.c0, gg);
      }
    }
  }
}

// ---- out = act(x * scale + shift (+ res)); scale/shift per channel or per (sample, channel) -------------------------------------
template <typename T>
__global__ void __launch_bounds__(BS_THREADS, 2) affine_act_v2_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, int per_sample, const T* __restrict__ res,
                                                                      int res_ld, T* __restrict__ out, int out_ld, long long HW, int C, long long n_px,
                                                                      int px_per_block, int relu) {
  pdl_enter();
  const PxRange r = px_range(C, n_px, px_per_block);
  float sc[8], sh[8];
  long long cur_n = -1;
  if (!per_sample) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale ? scale[r.c0 + k] : 1.f; sh[k] = scale ? shift[r.c0 + k] : 0.f; }
  }
  for (long long p = r.p0 + r.lane_px; p < r.p1; p += UNR * r.lanes) {
    Raw8<T> xv[UNR], rv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        xv[u].load(x + (size_t)q * x_ld + r.c0);
        if (res) rv[u].load(res + (size_t)q * res_ld + r.c0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        if (per_sample) {
          const long long n = img_of(q, HW);
          if (n != cur_n) {
            cur_n = n;
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = scale[n * C + r.c0 + k]; sh[k] = shift[n * C + r.c0 + k]; }
          }
        }
        float f[8], t[8];
        xv[u].unpack(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f[k] * sc[k] + sh[k];
        if (res) {
          rv[u].unpack(t);
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] += t[k];
        }
        if (relu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
        }
        Raw8<T>::store(out + (size_t)q * out_ld + r.c0, f);
      }
    }
  }
}

// ---- MaxPool2d(2, 2) half of a DownsamplerBlock, 8 channels (16 bytes for bf16) per thread, 32-bit index arithmetic ---------------
// forward: out[.., out_coff + c] = relu?(max2x2(in)[c] * scale[c] + shift[c])
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_v2_kernel(const T* __restrict__ in, int H, int W, int C, int in_ld, T* __restrict__ out, int out_ld,
                                                             int out_coff, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                                                             unsigned total8) {
  pdl_enter();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const unsigned c8 = (unsigned)C >> 3, Wo = (unsigned)W >> 1, Ho = (unsigned)H >> 1;
  const unsigned p = i / c8, c0 = (i - p * c8) * 8u;
  const unsigned t = p / Wo, ox = p - t * Wo, n = t / Ho, oy = t - n * Ho;
  const T* ip = in + (((size_t)n * H + 2 * oy) * W + 2 * ox) * in_ld + c0;
  Raw8<T> r[4];
  r[0].load(ip); r[1].load(ip + in_ld); r[2].load(ip + (size_t)W * in_ld); r[3].load(ip + (size_t)W * in_ld + in_ld);
  float m[8], f[8];
  r[0].unpack(m);
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    r[q].unpack(f);
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (scale) m[k] = m[k] * scale[c0 + k] + shift[c0 + k];
    if (relu) m[k] = fmaxf(m[k], 0.f);
  }
  Raw8<T>::store(out + (size_t)p * out_ld + out_coff + c0, m);
}

// backward: dx at the FIRST maximum of each window (torch's tie rule, scan order (0,0),(0,1),(1,0),(1,1)) += dout[.., out_coff + c]
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_v2_kernel(const T* __restrict__ in, int H, int W, int C, int in_ld, const T* __restrict__ dout,
                                                             int out_ld, int out_coff, T* __restrict__ dx, int dx_ld, unsigned total8) {
  pdl_enter();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const unsigned c8 = (unsigned)C >> 3, Wo = (unsigned)W >> 1, Ho = (unsigned)H >> 1;
  const unsigned p = i / c8, c0 = (i - p * c8) * 8u;
  const unsigned t = p / Wo, ox = p - t * Wo, n = t / Ho, oy = t - n * Ho;
  const size_t base = ((size_t)n * H + 2 * oy) * W + 2 * ox;
  const size_t offs[4] = {0, 1, (size_t)W, (size_t)W + 1};
  float m[8], f[8], g[8];
  int arg[8];
  Raw8<T> r;
  r.load(in + base * in_ld + c0);
  r.unpack(m);
#pragma unroll
  for (int k = 0; k < 8; ++k) arg[k] = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q) {
    r.load(in + (base + offs[q]) * in_ld + c0);
    r.unpack(f);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (f[k] > m[k]) { m[k] = f[k]; arg[k] = q; }
  }
  r.load(dout + (size_t)p * out_ld + out_coff + c0);
  r.unpack(g);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    T* dp = dx + (base + offs[q]) * dx_ld + c0;
    r.load(dp);
    r.unpack(f);
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (arg[k] == q) { f[k] += g[k]; any = true; }
    if (any) Raw8<T>::store(dp, f);
  }
}

// launch geometry shared by the kernels above
struct BsLaunch {
  unsigned grid;
  int px_per_block;
  size_t smem(int C, int nv) const {   // reduction rows: one per warp, or one per thread when the groups do not tile a warp
    const size_t per_warp = (size_t)(BS_THREADS / 32) * nv * C * sizeof(float), per_thread = (size_t)BS_THREADS * nv * 8 * sizeof(float);
    return (32 % (C / 8)) == 0 ? per_warp : (per_thread > per_warp ? per_thread : per_warp);
  }
};

static inline bool bs_fast_ok(int C) {
  const int c8 = C / 8;
  return C >= 8 && (C % 8) == 0 && c8 <= 32 && (32 % c8) == 0;
}
// reductions also take channel counts whose groups do not tile a warp (shared-memory fallback inside block_channel_reduce)
static inline bool bs_reduce_ok(int C) { return C >= 8 && (C % 8) == 0 && C <= 256; }

static inline BsLaunch bs_launch(int C, long long n_px, int blocks_per_sm) {
  BsLaunch l;
  const int lanes = BS_THREADS / (C / 8);
  long long want = 148LL * blocks_per_sm;
  long long ppb = (n_px + want - 1) / want;
  const long long min_ppb = 2LL * lanes;   // at least one two-pixel iteration per thread
  if (ppb < min_ppb) ppb = min_ppb;
  ppb = (ppb + lanes - 1) / lanes * lanes;
  l.px_per_block = (int)ppb;
  l.grid = (unsigned)((n_px + ppb - 1) / ppb);
  return l;
}

}  // namespace erfk
