// Generic gather convolution / transposed convolution / weight gradient on CUDA cores (fp32 accumulate).
// This is the fp32 "check mode" path and the fallback for shapes the tcgen05 kernels do not cover; it has no
// shared-memory tiling on purpose (weights are small enough to live in L1/L2) so that it is easy to verify.
#include <stdarg.h>

#include "common.cuh"
#include "bn_stream.cuh"

namespace erfk {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int PT = 4;   // output pixels (consecutive in x) per thread
constexpr int CT = 8;   // output channels per thread

template <typename T>
__global__ void __launch_bounds__(128) conv_gather_kernel(const T* __restrict__ in, const T* __restrict__ w, T* __restrict__ out,
                                                          erf_conv_geom g, erf_conv_epilogue e) {
  extern __shared__ float s_stats[];  // [2*w_ld] when e.stats
  const int cog = g.w_ld / CT;
  if (e.stats) {
    for (int i = threadIdx.x; i < 2 * g.w_ld; i += blockDim.x) s_stats[i] = 0.f;
    __syncthreads();
  }
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = (int)(tid % cog);
  long long pg = tid / cog;
  const int wo_groups = (g.Wo + PT - 1) / PT;
  const int oxg = (int)(pg % wo_groups);
  pg /= wo_groups;
  const int oy = (int)(pg % g.Ho);
  const int n = (int)(pg / g.Ho);
  const bool active = n < g.N;
  const int ox0 = oxg * PT;
  const int co0 = cg * CT;

  float acc[PT][CT];
#pragma unroll
  for (int j = 0; j < PT; ++j)
#pragma unroll
    for (int k = 0; k < CT; ++k) acc[j][k] = 0.f;

  if (active) {
    for (int ky = 0; ky < g.KH; ++ky) {
      int iy;
      if (!g.transposed) {
        iy = oy * g.stride - g.pad_h + ky * g.dil_h;
      } else {
        int t = oy + g.pad_h - ky * g.dil_h;
        if (t < 0 || (t % g.stride) != 0) continue;
        iy = t / g.stride;
      }
      if (iy < 0 || iy >= g.H) continue;
      for (int kx = 0; kx < g.KW; ++kx) {
        int ix[PT];
        bool any = false;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
          int ox = ox0 + j, v;
          if (!g.transposed) {
            v = ox * g.stride - g.pad_w + kx * g.dil_w;
          } else {
            int t = ox + g.pad_w - kx * g.dil_w;
            v = (t >= 0 && (t % g.stride) == 0) ? t / g.stride : -1;
          }
          if (ox >= g.Wo || v < 0 || v >= g.W) v = -1;
          ix[j] = v;
          any |= v >= 0;
        }
        if (!any) continue;
        const T* wp = w + (size_t)((ky * g.KW + kx) * g.Cin) * g.w_ld + co0;
        const T* ip = in + ((size_t)n * g.H + iy) * g.W * g.in_ld;
        for (int ci = 0; ci < g.Cin; ci += 4) {
          float xv[PT][4];
#pragma unroll
          for (int j = 0; j < PT; ++j) {
            if (ix[j] >= 0) {
              Vec<T, 4> t;
              t.load(ip + (size_t)ix[j] * g.in_ld + ci);
#pragma unroll
              for (int q = 0; q < 4; ++q) xv[j][q] = t.v[q];
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) xv[j][q] = 0.f;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vec<T, CT> wv;
            wv.load(wp + (size_t)(ci + q) * g.w_ld);
#pragma unroll
            for (int j = 0; j < PT; ++j)
#pragma unroll
              for (int k = 0; k < CT; ++k) acc[j][k] = fmaf(xv[j][q], wv.v[k], acc[j][k]);
          }
        }
      }
    }
    // epilogue
    float bsum[CT], bsq[CT];
#pragma unroll
    for (int k = 0; k < CT; ++k) bsum[k] = bsq[k] = 0.f;
    const T* maskp = (const T*)e.mask;
    const T* addp = (const T*)e.add;
    const bool vec_ok = co0 + CT <= g.Cout && (g.out_ld % CT) == 0 && (g.out_coff % CT) == 0;
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      int ox = ox0 + j;
      if (ox >= g.Wo) continue;
      size_t pix = ((size_t)n * g.Ho + oy) * g.Wo + ox;
      Vec<T, CT> o;
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        int co = co0 + k;
        float v = 0.f;
        if (co < g.Cout) {
          v = acc[j][k] + (e.bias ? e.bias[co] : 0.f);
          bsum[k] += v;
          bsq[k] += v * v;
          if (addp) v += to_f(addp[pix * e.add_ld + co]);
          if (e.relu) v = fmaxf(v, 0.f);
          if (maskp) v = to_f(maskp[pix * e.mask_ld + co]) > 0.f ? v : 0.f;
        }
        o.v[k] = v;
      }
      T* op = out + pix * g.out_ld + g.out_coff + co0;
      if (vec_ok) {
        o.store(op);
      } else {
#pragma unroll
        for (int k = 0; k < CT; ++k)
          if (co0 + k < g.Cout) op[k] = from_f<T>(o.v[k]);
      }
    }
    if (e.stats) {
#pragma unroll
      for (int k = 0; k < CT; ++k) {
        if (co0 + k < g.Cout) {
          atomicAdd(&s_stats[co0 + k], bsum[k]);
          atomicAdd(&s_stats[g.w_ld + co0 + k], bsq[k]);
        }
      }
    }
  }
  if (e.stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < g.Cout; i += blockDim.x) {
      atomicAdd(&e.stats[g.out_coff + i], (double)s_stats[i]);
      atomicAdd(&e.stats[e.stats_c + g.out_coff + i], (double)s_stats[g.w_ld + i]);
    }
  }
}

// dw[(tap*Cin + ci)*w_ld + co] += sum_p in[q(p,tap)][ci] * dout[p][co];  db[co] += sum_p dout[p][co]
template <typename T>
__global__ void conv_wgrad_kernel(const T* __restrict__ in, const T* __restrict__ dout, float* __restrict__ dw,
                                  erf_conv_geom g, int px_per_block, int ntile, int lanes) {
  const int tap = blockIdx.y;
  const int ky = tap / g.KW, kx = tap % g.KW;
  const int tile = threadIdx.x % ntile;
  const int lane = threadIdx.x / ntile;
  if (lane >= lanes) return;
  const int cog = g.w_ld / CT;
  const int co0 = (tile % cog) * CT;
  const int ci0 = (tile / cog) * 4;
  const long long total = (long long)g.N * g.Ho * g.Wo;
  long long p0 = (long long)blockIdx.x * px_per_block;
  long long p1 = p0 + px_per_block;
  if (p1 > total) p1 = total;
  float acc[4][CT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < CT; ++k) acc[q][k] = 0.f;
  const bool full = co0 + CT <= g.Cout && (g.out_coff % CT) == 0 && (g.out_ld % CT) == 0;
  for (long long p = p0 + lane; p < p1; p += lanes) {
    int ox = (int)(p % g.Wo);
    long long t = p / g.Wo;
    int oy = (int)(t % g.Ho);
    int n = (int)(t / g.Ho);
    int iy, ix;
    if (!g.transposed) {
      iy = oy * g.stride - g.pad_h + ky * g.dil_h;
      ix = ox * g.stride - g.pad_w + kx * g.dil_w;
    } else {
      int ty = oy + g.pad_h - ky * g.dil_h, tx = ox + g.pad_w - kx * g.dil_w;
      if (ty < 0 || tx < 0 || (ty % g.stride) || (tx % g.stride)) continue;
      iy = ty / g.stride;
      ix = tx / g.stride;
    }
    if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) continue;
    Vec<T, 4> xv;
    xv.load(in + (((size_t)n * g.H + iy) * g.W + ix) * g.in_ld + ci0);
    float dv[CT];
    const T* dp = dout + (size_t)p * g.out_ld + g.out_coff + co0;
    if (full) {
      Vec<T, CT> t8;
      t8.load(dp);
#pragma unroll
      for (int k = 0; k < CT; ++k) dv[k] = t8.v[k];
    } else {
#pragma unroll
      for (int k = 0; k < CT; ++k) dv[k] = (co0 + k < g.Cout) ? to_f(dp[k]) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < CT; ++k) acc[q][k] = fmaf(xv.v[q], dv[k], acc[q][k]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < CT; ++k)
      if (co0 + k < g.Cout && acc[q][k] != 0.f) atomicAdd(&dw[((size_t)tap * g.Cin + ci0 + q) * g.w_ld + co0 + k], acc[q][k]);
}

// Shared-memory tiled weight gradient (same contract as conv_wgrad_kernel): one block = one tap x one pixel chunk.
// Pixels are staged 32 at a time as fp32 ([32][CIN_S] gathered x, [32][COUT_S] dy); a 16x16 thread grid keeps a
// TI x TJ register tile of dW (ci = ty*TI + i, co = tx*TJ + j) and the block flushes it with one atomic per element.
template <typename T, int TI, int TJ>
__global__ void __launch_bounds__(256) conv_wgrad_tiled_kernel(const T* __restrict__ in, const T* __restrict__ dout, float* __restrict__ dw,
                                                               erf_conv_geom g, int px_per_block, int cin_real) {
  constexpr int CIN_S = 16 * TI, COUT_S = 16 * TJ, PXT = 32;
  __shared__ __align__(16) float xs[PXT][CIN_S];
  __shared__ __align__(16) float ds[PXT][COUT_S];
  const int tap = blockIdx.y;
  const int ky = tap / g.KW, kx = tap % g.KW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // The loop runs over the DENSE side of the gather: output pixels for a conv, INPUT pixels for a transposed conv
  // (there every (tap, input pixel) pair is valid, whereas only 1/stride^2 of the (tap, output pixel) pairs are).
  const int DW = g.transposed ? g.W : g.Wo, DH = g.transposed ? g.H : g.Ho;
  const long long total = (long long)g.N * DH * DW;
  const long long p0 = (long long)blockIdx.x * px_per_block;
  long long p1 = p0 + px_per_block;
  if (p1 > total) p1 = total;
  float acc[TI][TJ];
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[i][j] = 0.f;
  for (long long pb = p0; pb < p1; pb += PXT) {
    // ---- stage 32 pixels: 4 input channels / 4 output channels per thread-iteration
    for (int e = threadIdx.x; e < PXT * (CIN_S / 4 + COUT_S / 4); e += 256) {
      const bool is_x = e < PXT * (CIN_S / 4);
      const int e2 = is_x ? e : e - PXT * (CIN_S / 4);
      const int per = is_x ? CIN_S / 4 : COUT_S / 4;
      const int pp = e2 / per, c4 = (e2 % per) * 4;
      const long long p = pb + pp;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int cmax = is_x ? g.Cin : g.Cout;
      if (p < p1 && c4 < cmax) {
        int dx = (int)(p % DW);
        long long t = p / DW;
        int dy_ = (int)(t % DH), n = (int)(t / DH);
        // coordinates of this pixel on the input (iy, ix) and output (oy, ox) grids
        int iy, ix, oy, ox;
        bool ok = true;
        if (!g.transposed) {
          oy = dy_; ox = dx;
          iy = oy * g.stride - g.pad_h + ky * g.dil_h;
          ix = ox * g.stride - g.pad_w + kx * g.dil_w;
          ok = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
        } else {
          iy = dy_; ix = dx;
          oy = iy * g.stride - g.pad_h + ky * g.dil_h;
          ox = ix * g.stride - g.pad_w + kx * g.dil_w;
          ok = oy >= 0 && oy < g.Ho && ox >= 0 && ox < g.Wo;
        }
        if (ok) {
          if (is_x) {
            Vec<T, 4> t4;
            t4.load(in + (((size_t)n * g.H + iy) * g.W + ix) * g.in_ld + c4);
            v = make_float4(t4.v[0], t4.v[1], t4.v[2], t4.v[3]);
          } else {
            const T* dp = dout + (((size_t)n * g.Ho + oy) * g.Wo + ox) * g.out_ld + g.out_coff + c4;
            if (c4 + 4 <= g.Cout && ((g.out_ld | g.out_coff) & 3) == 0) {
              Vec<T, 4> t4;
              t4.load(dp);
              v = make_float4(t4.v[0], t4.v[1], t4.v[2], t4.v[3]);
            } else {
              float tmp[4] = {0.f, 0.f, 0.f, 0.f};
              for (int k = 0; k < 4; ++k)
                if (c4 + k < g.Cout) tmp[k] = to_f(dp[k]);
              v = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
            }
          }
        }
      }
      if (is_x) *reinterpret_cast<float4*>(&xs[pp][c4]) = v;
      else *reinterpret_cast<float4*>(&ds[pp][c4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = 0; pp < PXT; ++pp) {
      float xv[TI], dv[TJ];
#pragma unroll
      for (int i = 0; i < TI; ++i) xv[i] = xs[pp][ty * TI + i];
#pragma unroll
      for (int j = 0; j < TJ; ++j) dv[j] = ds[pp][tx * TJ + j];
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = fmaf(xv[i], dv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TI; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
      const int ci = ty * TI + i, co = tx * TJ + j;
      if (ci < g.Cin && co < g.Cout && acc[i][j] != 0.f) {
        if (cin_real <= 0) {            // packed [(tap*Cin + ci)][w_ld]
          atomicAdd(&dw[((size_t)tap * g.Cin + ci) * g.w_ld + co], acc[i][j]);
        } else if (ci < cin_real) {     // torch parameter layout
          const size_t o = g.transposed ? (((size_t)ci * g.Cout + co) * g.KH + ky) * g.KW + kx : (((size_t)co * cin_real + ci) * g.KH + ky) * g.KW + kx;
          atomicAdd(&dw[o], acc[i][j]);
        }
      }
    }
}

template <typename T>
static bool launch_wgrad_tiled(const T* in, const T* dout, float* dw, const erf_conv_geom& g, cudaStream_t s, int cin_real = 0) {
  const int ti = g.Cin <= 16 ? 1 : (g.Cin <= 32 ? 2 : (g.Cin <= 64 ? 4 : (g.Cin <= 128 ? 8 : 0)));
  const int tj = g.Cout <= 16 ? 1 : (g.Cout <= 32 ? 2 : (g.Cout <= 64 ? 4 : (g.Cout <= 128 ? 8 : 0)));
  if (!ti || !tj) return false;
  const long long total = g.transposed ? (long long)g.N * g.H * g.W : (long long)g.N * g.Ho * g.Wo;
  const int taps = g.KH * g.KW;
  long long want_blocks = 148 * 6 / taps + 1;
  long long ppb = (total + want_blocks - 1) / want_blocks;
  ppb = ((ppb + 31) / 32) * 32;
  if (ppb < 256) ppb = 256;
  dim3 grid((unsigned)((total + ppb - 1) / ppb), taps);
#define WG_CASE(TI_, TJ_) \
  if (ti == TI_ && tj == TJ_) { conv_wgrad_tiled_kernel<T, TI_, TJ_><<<grid, 256, 0, s>>>(in, dout, dw, g, (int)ppb, cin_real); return true; }
  WG_CASE(1, 1) WG_CASE(1, 2) WG_CASE(1, 4) WG_CASE(1, 8) WG_CASE(2, 1) WG_CASE(2, 2) WG_CASE(2, 4) WG_CASE(2, 8)
  WG_CASE(4, 1) WG_CASE(4, 2) WG_CASE(4, 4) WG_CASE(4, 8) WG_CASE(8, 1) WG_CASE(8, 2) WG_CASE(8, 4) WG_CASE(8, 8)
#undef WG_CASE
  return false;
}

// out[c] += sum_p x[p*ld + coff + c]
template <typename T>
__global__ void channel_sum_kernel(const T* __restrict__ x, int ld, int coff, int C, long long n_px, int px_per_block,
                                   float* __restrict__ out) {
  extern __shared__ float s[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s[i] = 0.f;
  __syncthreads();
  const int lanes = blockDim.x / C;  // host guarantees C <= blockDim.x
  long long p0 = (long long)blockIdx.x * px_per_block, p1 = p0 + px_per_block;
  if (p1 > n_px) p1 = n_px;
  const int c = threadIdx.x % C, lane = threadIdx.x / C;
  if (lane < lanes) {
    float a = 0.f;
    for (long long p = p0 + lane; p < p1; p += lanes) a += to_f(x[(size_t)p * ld + coff + c]);
    atomicAdd(&s[c], a);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&out[i], s[i]);
}

static int check_geom(const erf_conv_geom* g, const char* who) {
  ERF_CHECK_ARG(g != nullptr, "%s: null geometry", who);
  ERF_CHECK_ARG(g->N > 0 && g->H > 0 && g->W > 0 && g->Ho > 0 && g->Wo > 0, "%s: empty tensor (N=%d H=%d W=%d Ho=%d Wo=%d)", who, g->N,
                g->H, g->W, g->Ho, g->Wo);
  ERF_CHECK_ARG(g->Cin > 0 && g->Cin % 4 == 0 && g->in_ld % 4 == 0 && g->in_ld >= g->Cin, "%s: Cin (%d) and in_ld (%d) must be multiples of 4", who,
                g->Cin, g->in_ld);
  ERF_CHECK_ARG(g->Cout > 0 && g->w_ld % CT == 0 && g->w_ld >= g->Cout && g->w_ld < g->Cout + CT, "%s: w_ld (%d) must be Cout (%d) rounded up to 8", who,
                g->w_ld, g->Cout);
  ERF_CHECK_ARG(g->out_ld >= g->out_coff + g->Cout, "%s: output slice [%d,%d) exceeds out_ld %d", who, g->out_coff, g->out_coff + g->Cout, g->out_ld);
  ERF_CHECK_ARG(g->stride >= 1 && g->KH >= 1 && g->KW >= 1 && g->dil_h >= 1 && g->dil_w >= 1, "%s: bad stride/kernel/dilation", who);
  return ERF_OK;
}

}  // namespace erfk

using namespace erfk;

extern "C" const char* erf_last_error_string(void) { return erfk::g_err; }
extern "C" int erf_abi_version(void) { return 1; }

extern "C" int erf_conv_fwd(const void* in, const void* w, void* out, const erf_conv_geom* g, const erf_conv_epilogue* e, int dtype,
                            erf_stream_t stream) {
  if (int rc = check_geom(g, "erf_conv_fwd")) return rc;
  ERF_CHECK_ARG(in && w && out && e, "erf_conv_fwd: null pointer");
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_conv_fwd: bad dtype %d", dtype);
  long long threads = (long long)g->N * g->Ho * ((g->Wo + PT - 1) / PT) * (g->w_ld / CT);
  int block = 128;
  long long grid = (threads + block - 1) / block;
  size_t smem = e->stats ? 2 * g->w_ld * sizeof(float) : 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ERF_F32)
    conv_gather_kernel<float><<<(unsigned)grid, block, smem, s>>>((const float*)in, (const float*)w, (float*)out, *g, *e);
  else
    conv_gather_kernel<__nv_bfloat16><<<(unsigned)grid, block, smem, s>>>((const __nv_bfloat16*)in, (const __nv_bfloat16*)w, (__nv_bfloat16*)out, *g, *e);
  ERF_CHECK_LAUNCH("erf_conv_fwd");
  return ERF_OK;
}

extern "C" int erf_conv_wgrad(const void* in, const void* dout, float* dw, float* db, const erf_conv_geom* g, int dtype, erf_stream_t stream) {
  if (int rc = check_geom(g, "erf_conv_wgrad")) return rc;
  ERF_CHECK_ARG(in && dout && dw, "erf_conv_wgrad: null pointer");
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_conv_wgrad: bad dtype %d", dtype);
  const int ntile = (g->Cin / 4) * (g->w_ld / CT);
  ERF_CHECK_ARG(ntile <= 1024, "erf_conv_wgrad: Cin*Cout too large (%d x %d)", g->Cin, g->Cout);
  long long total = (long long)g->N * g->Ho * g->Wo;
  int taps = g->KH * g->KW;
  cudaStream_t s = (cudaStream_t)stream;
  bool done = dtype == ERF_F32 ? launch_wgrad_tiled<float>((const float*)in, (const float*)dout, dw, *g, s)
                               : launch_wgrad_tiled<__nv_bfloat16>((const __nv_bfloat16*)in, (const __nv_bfloat16*)dout, dw, *g, s);
  if (!done) {
    int block = ntile <= 256 ? 256 : (ntile <= 512 ? 512 : 1024);
    int lanes = block / ntile;
    long long want_blocks = 148 * 4 / taps + 1;
    long long ppb = (total + want_blocks - 1) / want_blocks;
    if (ppb < 64 * lanes) ppb = 64 * lanes;
    dim3 grid((unsigned)((total + ppb - 1) / ppb), taps);
    if (dtype == ERF_F32)
      conv_wgrad_kernel<float><<<grid, block, 0, s>>>((const float*)in, (const float*)dout, dw, *g, (int)ppb, ntile, lanes);
    else
      conv_wgrad_kernel<__nv_bfloat16><<<grid, block, 0, s>>>((const __nv_bfloat16*)in, (const __nv_bfloat16*)dout, dw, *g, (int)ppb, ntile, lanes);
  }
  ERF_CHECK_LAUNCH("erf_conv_wgrad");
  if (db) {
    int cblock = 256;
    long long cb = 148 * 2;
    long long cppb = (total + cb - 1) / cb;
    if (cppb < 256) cppb = 256;
    unsigned cgrid = (unsigned)((total + cppb - 1) / cppb);
    size_t smem = g->Cout * sizeof(float);
    if (dtype == ERF_F32)
      channel_sum_kernel<float><<<cgrid, cblock, smem, s>>>((const float*)dout, g->out_ld, g->out_coff, g->Cout, total, (int)cppb, db);
    else
      channel_sum_kernel<__nv_bfloat16><<<cgrid, cblock, smem, s>>>((const __nv_bfloat16*)dout, g->out_ld, g->out_coff, g->Cout, total, (int)cppb, db);
    ERF_CHECK_LAUNCH("erf_conv_wgrad(bias)");
  }
  return ERF_OK;
}

extern "C" int erf_channel_sum(const void* x, int ld, int coff, int C, long long n_px, float* out, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(x && out && n_px > 0 && C > 0 && C <= 256 && ld >= coff + C, "erf_channel_sum: bad arguments");
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_channel_sum: bad dtype %d", dtype);
  if (bs_reduce_ok(C) && (ld % 8) == 0 && (coff % 8) == 0 && ((uintptr_t)x % 16) == 0) {   // vectorised slice [coff, coff + C) of each pixel row
    const BsLaunch l = bs_launch(C, n_px, 2);
    if (dtype == ERF_F32)
      launch_pdl(channel_reduce_v2_kernel<float, false, float>, dim3(l.grid), dim3(BS_THREADS), l.smem(C, 1), (cudaStream_t)((cudaStream_t)stream), (const float*)x + coff, ld, C, n_px, l.px_per_block, out);
    else
      launch_pdl(channel_reduce_v2_kernel<__nv_bfloat16, false, float>, dim3(l.grid), dim3(BS_THREADS), l.smem(C, 1), (cudaStream_t)((cudaStream_t)stream), (const __nv_bfloat16*)x + coff, ld, C, n_px, l.px_per_block, out);
    ERF_CHECK_LAUNCH("erf_channel_sum");
    return ERF_OK;
  }
  long long cb = 148 * 2;
  long long cppb = (n_px + cb - 1) / cb;
  if (cppb < 256) cppb = 256;
  unsigned cgrid = (unsigned)((n_px + cppb - 1) / cppb);
  size_t smem = C * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == ERF_F32)
    channel_sum_kernel<float><<<cgrid, 256, smem, s>>>((const float*)x, ld, coff, C, n_px, (int)cppb, out);
  else
    channel_sum_kernel<__nv_bfloat16><<<cgrid, 256, smem, s>>>((const __nv_bfloat16*)x, ld, coff, C, n_px, (int)cppb, out);
  ERF_CHECK_LAUNCH("erf_channel_sum");
  return ERF_OK;
}

extern "C" int erf_conv_wgrad_torch(const void* in, const void* dout, float* dw, float* db, const erf_conv_geom* g, int cin_real, int dtype,
                                    erf_stream_t stream) {
  if (int rc = check_geom(g, "erf_conv_wgrad_torch")) return rc;
  ERF_CHECK_ARG(in && dout && dw && cin_real > 0 && cin_real <= g->Cin, "erf_conv_wgrad_torch: bad arguments");
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_conv_wgrad_torch: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream;
  bool done = dtype == ERF_F32 ? launch_wgrad_tiled<float>((const float*)in, (const float*)dout, dw, *g, s, cin_real)
                               : launch_wgrad_tiled<__nv_bfloat16>((const __nv_bfloat16*)in, (const __nv_bfloat16*)dout, dw, *g, s, cin_real);
  if (!done) {
    set_error("erf_conv_wgrad_torch: unsupported channel counts Cin=%d Cout=%d (<= 128 each)", g->Cin, g->Cout);
    return ERF_ERR_UNSUPPORTED;
  }
  ERF_CHECK_LAUNCH("erf_conv_wgrad_torch");
  if (db) return erf_channel_sum(dout, g->out_ld, g->out_coff, g->Cout, (long long)g->N * g->Ho * g->Wo, db, dtype, stream);
  return ERF_OK;
}
