// The 3-channel stem (encoder.initial_block, eval/erfnet.py:67 + 12-23): DownsamplerBlock(3, 16) = cat([conv3x3 s2 p1 (3 -> 13),
// maxpool2x2 (3)]) -> BN -> ReLU.  K = 27: no tensor-core shape fits, and the block is pure bandwidth (25 MB in, 25 MB out at
// batch 6), so it gets its own CUDA-core kernels:
//   stem_fwd_kernel    one thread per output pixel: 9 taps x 8 bytes in, 13 conv + 3 pooled channels (32 bytes) out, weights
//                      broadcast from shared memory, optional per-channel affine + ReLU (eval: BN folded) and fused batch statistics
//   stem_wgrad_kernel  dW[13][3][3][3] (+ db): thread = (tap, group of 4 output channels), a block walks whole output rows staged in
//                      shared memory; 12 FMAs per 2 shared loads
// Input x: NHWC [N][H][W][4] (channel 3 = padding), output y: NHWC [N][H/2][W/2][16], bf16 or fp32.
#include "common.cuh"

namespace erfk {

constexpr int STEM_CO = 13, STEM_CI = 3, STEM_C = 16;

template <typename T> struct Px4;   // one 4-channel input pixel
template <> struct Px4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x), b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    f[0] = __low2float(a); f[1] = __high2float(a); f[2] = __low2float(b); f[3] = __high2float(b);
  }
};
template <> struct Px4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
  }
};

// Two adjacent output pixels (A = 2j, B = 2j+1 of an output row) per thread: their 3x3 stride-2 windows share input column 4j+1, the
// weights are fetched from shared memory once for both, and the 13 conv channels are accumulated as 6 packed fp32 pairs (FFMA2) + 1 scalar
// per pixel -- the kernel is FMA-issue bound, not bandwidth bound (27 x 13 MACs for 32 output bytes).  The accumulation order per channel
// is bias, then (ky, kx, ci) ascending, as in a plain loop.
// NCHW_IN: x is the caller's NCHW fp32 tensor [N][3][H][W] (inference: no separate layout pass); otherwise NHWC [N][H][W][4] of type T.
template <typename T, bool NCHW_IN, bool STATS>
__global__ void __launch_bounds__(256, 2) stem_fwd_kernel(const void* __restrict__ xin, int N, int H, int W, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, int relu, T* __restrict__ y, double* __restrict__ stats) {
  pdl_enter();
  __shared__ __align__(16) float s_w[9 * STEM_CI * STEM_C];   // [tap][ci][co padded to 16]
  __shared__ float s_b[STEM_C], s_sc[STEM_C], s_sh[STEM_C];
  __shared__ float s_red[8][2 * STEM_C];
  for (int i = threadIdx.x; i < 9 * STEM_CI * STEM_C; i += blockDim.x) {
    const int co = i % STEM_C, ci = (i / STEM_C) % STEM_CI, tap = i / (STEM_C * STEM_CI);
    s_w[i] = co < STEM_CO ? w[(co * STEM_CI + ci) * 9 + tap] : 0.f;   // torch layout [co][ci][ky][kx]
  }
  if (threadIdx.x < STEM_C) {
    s_b[threadIdx.x] = (bias && threadIdx.x < STEM_CO) ? bias[threadIdx.x] : 0.f;
    s_sc[threadIdx.x] = scale ? scale[threadIdx.x] : 1.f;
    s_sh[threadIdx.x] = shift ? shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  const int Ho = H / 2, Wp = W / 4;                                   // Wp = output pixel pairs per row
  const unsigned total = (unsigned)N * (unsigned)Ho * (unsigned)Wp;   // host-checked < 2^31: 32-bit index arithmetic
  float st_sum[STEM_C], st_sq[STEM_C];
#pragma unroll
  for (int c = 0; c < STEM_C; ++c) st_sum[c] = st_sq[c] = 0.f;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const unsigned t2 = p / (unsigned)Wp;
    const int j = (int)(p - t2 * (unsigned)Wp);
    const int n = (int)(t2 / (unsigned)Ho), oy = (int)(t2 - (unsigned)n * (unsigned)Ho);
    float2 accA[6], accB[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) accA[q] = accB[q] = make_float2(s_b[2 * q], s_b[2 * q + 1]);
    float a12 = s_b[12], b12 = s_b[12];
    float mxA[3] = {-INFINITY, -INFINITY, -INFINITY}, mxB[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1   // one input row at a time: keeps the 27 x 13 weights out of the register file (they stream from shared memory)
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy - 1 + ky;
      const bool row_ok = iy >= 0;                 // with even H, W and padding 1 only the top row and the left column fall outside
      float v[5][3];                               // input columns 4j-1 .. 4j+3, channels 0..2
      if (NCHW_IN) {
        const float* x = reinterpret_cast<const float*>(xin);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* r = x + (((size_t)n * 3 + c) * H + (row_ok ? iy : 0)) * W + 4 * j;
          float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
          float l = 0.f;
          if (row_ok) {
            q4 = *reinterpret_cast<const float4*>(r);
            if (j > 0) l = r[-1];
          }
          v[0][c] = l; v[1][c] = q4.x; v[2][c] = q4.y; v[3][c] = q4.z; v[4][c] = q4.w;
        }
      } else {
        const T* r = reinterpret_cast<const T*>(xin) + (((size_t)n * H + (row_ok ? iy : 0)) * W + 4 * j) * 4;
#pragma unroll
        for (int col = 0; col < 5; ++col) {
          float f[4] = {0.f, 0.f, 0.f, 0.f};
          if (row_ok && (col > 0 || j > 0)) Px4<T>::load(r + (col - 1) * 4, f);
          v[col][0] = f[0]; v[col][1] = f[1]; v[col][2] = f[2];
        }
      }
      if (ky >= 1) {   // the 2x2 pooling windows: A = columns 4j, 4j+1 (v[1], v[2]); B = columns 4j+2, 4j+3 (v[3], v[4]); always inside
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mxA[c] = fmaxf(mxA[c], fmaxf(v[1][c], v[2][c]));
          mxB[c] = fmaxf(mxB[c], fmaxf(v[3][c], v[4][c]));
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wt = s_w + (ky * 3 + kx) * STEM_CI * STEM_C;
#pragma unroll
        for (int ci = 0; ci < STEM_CI; ++ci) {
          const float va = v[kx][ci], vb = v[kx + 2][ci];     // A reads column 4j-1+kx, B column 4j+1+kx
          const float2 va2 = make_float2(va, va), vb2 = make_float2(vb, vb);
          const float4* w4p = reinterpret_cast<const float4*>(wt + ci * STEM_C);
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const float4 w4 = w4p[q];
            accA[2 * q] = __ffma2_rn(va2, make_float2(w4.x, w4.y), accA[2 * q]);
            accA[2 * q + 1] = __ffma2_rn(va2, make_float2(w4.z, w4.w), accA[2 * q + 1]);
            accB[2 * q] = __ffma2_rn(vb2, make_float2(w4.x, w4.y), accB[2 * q]);
            accB[2 * q + 1] = __ffma2_rn(vb2, make_float2(w4.z, w4.w), accB[2 * q + 1]);
          }
          const float w12 = wt[ci * STEM_C + 12];
          a12 = fmaf(va, w12, a12);
          b12 = fmaf(vb, w12, b12);
        }
      }
    }
    // round to the storage type first: statistics and the affine epilogue see the values as stored
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      float acc[STEM_C];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        acc[2 * q] = px ? accB[q].x : accA[q].x;
        acc[2 * q + 1] = px ? accB[q].y : accA[q].y;
      }
      acc[12] = px ? b12 : a12;
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[STEM_CO + c] = px ? mxB[c] : mxA[c];
      __align__(16) T o[STEM_C];
#pragma unroll
      for (int c = 0; c < STEM_C; ++c) {
        float f = acc[c] * s_sc[c] + s_sh[c];
        if (relu) f = fmaxf(f, 0.f);
        o[c] = from_f<T>(f);
        if (STATS) {
          const float r = to_f(o[c]);
          st_sum[c] += r;
          st_sq[c] += r * r;
        }
      }
      T* dst = y + ((size_t)p * 2 + px) * STEM_C;      // output pixel index = (n*Ho + oy)*Wo + 2j + px = 2p + px
      if (sizeof(T) == 2) {
        reinterpret_cast<uint4*>(dst)[0] = *reinterpret_cast<const uint4*>(&o[0]);
        reinterpret_cast<uint4*>(dst)[1] = *reinterpret_cast<const uint4*>(&o[8]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(dst)[q] = *reinterpret_cast<const uint4*>(&o[4 * q]);
      }
    }
  }
  if (STATS) {
#pragma unroll
    for (int c = 0; c < STEM_C; ++c) {
      st_sum[c] = warp_sum(st_sum[c]);
      st_sq[c] = warp_sum(st_sq[c]);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < STEM_C; ++c) { s_red[warp][c] = st_sum[c]; s_red[warp][STEM_C + c] = st_sq[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * STEM_C) {
      float t = 0.f;
      for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += s_red[wv][threadIdx.x];
      atomicAdd(&stats[threadIdx.x], (double)t);
    }
  }
}

// dW[co][ci][ky][kx] += sum_p dy[p][co] * x[2oy-1+ky][2ox-1+kx][ci];  db[co] += sum_p dy[p][co]   (dy: first 13 channels of a 16-wide row)
constexpr int SW_LANES = 7, SW_GROUP = 36;   // 36 threads = 9 taps x 4 output-channel quads; 7 pixel lanes -> 252 active threads
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int N, int H, int W,
                                                         float* __restrict__ dw, float* __restrict__ db) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int Ho = H / 2, Wo = W / 2;
  T* s_x = reinterpret_cast<T*>(s_raw);                               // [3 rows][W][4]
  T* s_dy = s_x + (size_t)3 * W * 4;                                  // [Wo][16]
  float* s_acc = reinterpret_cast<float*>(s_raw);  // [SW_LANES][9][16][3] partial sums + [16] bias sums: ALIASES the row staging, used only after the last row
  const int tid = threadIdx.x;
  const int lane_px = tid / SW_GROUP, g = tid % SW_GROUP, tap = g / 4, cq = g % 4;
  const int ky = tap / 3, kx = tap % 3;
  const bool active = lane_px < SW_LANES;
  float acc[4][3], bsum[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { bsum[a] = 0.f; acc[a][0] = acc[a][1] = acc[a][2] = 0.f; }
  const int rows = N * Ho;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int n = r / Ho, oy = r % Ho;
    __syncthreads();   // previous row fully consumed
    // stage input rows 2oy-1 .. 2oy+1 (zero outside the image) and the dy row, 16 bytes per thread and step
    const int x_vecs = W * 4 * (int)sizeof(T) / 16, dy_vecs = Wo * STEM_C * (int)sizeof(T) / 16;
    for (int i = tid; i < 3 * x_vecs; i += blockDim.x) {
      const int rr = i / x_vecs, v = i % x_vecs, iy = 2 * oy - 1 + rr;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (iy >= 0 && iy < H) val = reinterpret_cast<const uint4*>(x + ((size_t)n * H + iy) * W * 4)[v];
      reinterpret_cast<uint4*>(s_x)[i] = val;
    }
    for (int i = tid; i < dy_vecs; i += blockDim.x) reinterpret_cast<uint4*>(s_dy)[i] = reinterpret_cast<const uint4*>(dy + (size_t)r * Wo * STEM_C)[i];
    __syncthreads();
    if (active) {
      for (int ox = lane_px; ox < Wo; ox += SW_LANES) {
        const int ix = 2 * ox - 1 + kx;
        float xv[4] = {0.f, 0.f, 0.f, 0.f};
        if (ix >= 0 && ix < W) Px4<T>::load(s_x + ((size_t)ky * W + ix) * 4, xv);
        float dv[4];
        Px4<T>::load(s_dy + (size_t)ox * STEM_C + cq * 4, dv);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          acc[a][0] += dv[a] * xv[0];
          acc[a][1] += dv[a] * xv[1];
          acc[a][2] += dv[a] * xv[2];
          bsum[a] += dv[a];
        }
      }
    }
  }
  // combine the pixel lanes of the block, then one atomic per weight and block
  __syncthreads();
  float* mine = s_acc + ((size_t)lane_px * SW_GROUP + g) * 16;   // 12 weights + 4 bias sums per thread
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      mine[a * 3] = acc[a][0]; mine[a * 3 + 1] = acc[a][1]; mine[a * 3 + 2] = acc[a][2];
      mine[12 + a] = bsum[a];
    }
  }
  __syncthreads();
  for (int i = tid; i < SW_GROUP * 16; i += blockDim.x) {
    float t = 0.f;
    for (int l = 0; l < SW_LANES; ++l) t += s_acc[(size_t)l * SW_GROUP * 16 + i];
    const int gg = i / 16, e = i % 16, tp = gg / 4, q = gg % 4;
    if (e < 12) {
      const int co = q * 4 + e / 3, ci = e % 3;
      if (co < STEM_CO) atomicAdd(&dw[(co * STEM_CI + ci) * 9 + tp], t);
    } else if (db && tp == 4) {   // every tap thread saw the same dy: take the centre tap's bias sums
      const int co = q * 4 + (e - 12);
      if (co < STEM_CO) atomicAdd(&db[co], t);
    }
  }
}

// bf16, second version.  The kernel above is FMA-ISSUE bound (12 MACs for ~35 instructions per pixel and thread) and is the LAST kernel of backward,
// fully exposed in front of the optimiser.  Here a thread owns ONE tap and ALL 13 output channels x 3 input channels: per pixel one 8-byte load of
// x, two 16-byte loads of dy, and 21 packed FFMA2 (output channels in pairs -- a bf16x2 word of dy unpacks straight into the float2 operand) = 39
// MACs for ~50 instructions.  252 threads = 28 pixel lanes x 9 taps; the lanes are combined through shared memory (aliasing the row staging) at the end.
constexpr int SW2_LANES = 28, SW2_VALS = 44;   // 7 channel pairs x 3 input channels x 2 = 42 sums per thread, padded
__global__ void __launch_bounds__(256, 3) stem_wgrad_bf16_v2_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int N, int H, int W,
                                                                 float* __restrict__ dw, float* __restrict__ db) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t s_raw[];
  const int Ho = H / 2, Wo = W / 2;
  __nv_bfloat16* s_x = reinterpret_cast<__nv_bfloat16*>(s_raw);   // [3 rows][W][4]
  __nv_bfloat16* s_dy = s_x + (size_t)3 * W * 4;                   // [Wo][16]
  float* s_acc = reinterpret_cast<float*>(s_raw);                  // [252][SW2_VALS] + [SW2_LANES][14] bias sums: ALIASES the staging, used after the last row
  const int tid = threadIdx.x;
  const int lane_px = tid / 9, tap = tid % 9, ky = tap / 3, kx = tap % 3;
  const bool active = tid < SW2_LANES * 9;
  float2 acc[7][3], bs[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    bs[k] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[k][c] = make_float2(0.f, 0.f);
  }
  const int rows = N * Ho;
  const int x_vecs = W * 4 * 2 / 16, dy_vecs = Wo * STEM_C * 2 / 16;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int n = r / Ho, oy = r % Ho;
    __syncthreads();   // previous row fully consumed
    for (int i = tid; i < 3 * x_vecs; i += blockDim.x) {
      const int rr = i / x_vecs, v = i % x_vecs, iy = 2 * oy - 1 + rr;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (iy >= 0 && iy < H) val = reinterpret_cast<const uint4*>(x + ((size_t)n * H + iy) * W * 4)[v];
      reinterpret_cast<uint4*>(s_x)[i] = val;
    }
    for (int i = tid; i < dy_vecs; i += blockDim.x) reinterpret_cast<uint4*>(s_dy)[i] = reinterpret_cast<const uint4*>(dy + (size_t)r * Wo * STEM_C)[i];
    __syncthreads();
    if (active) {
      const __nv_bfloat16* xrow = s_x + (size_t)ky * W * 4;
      for (int ox = lane_px; ox < Wo; ox += SW2_LANES) {
        const int ix = 2 * ox - 1 + kx;
        uint2 xw = make_uint2(0u, 0u);
        if (ix >= 0 && ix < W) xw = *reinterpret_cast<const uint2*>(xrow + (size_t)ix * 4);
        const float x0 = __uint_as_float(xw.x << 16), x1 = __uint_as_float(xw.x & 0xFFFF0000u), x2 = __uint_as_float(xw.y << 16);
        const float2 xs0 = make_float2(x0, x0), xs1 = make_float2(x1, x1), xs2 = make_float2(x2, x2);
        const uint4 d0 = *reinterpret_cast<const uint4*>(s_dy + (size_t)ox * STEM_C), d1 = *reinterpret_cast<const uint4*>(s_dy + (size_t)ox * STEM_C + 8);
        const uint32_t dwd[7] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z};   // channel pairs (0,1) .. (12,13)
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const float2 dv = make_float2(__uint_as_float(dwd[k] << 16), __uint_as_float(dwd[k] & 0xFFFF0000u));
          acc[k][0] = __ffma2_rn(dv, xs0, acc[k][0]);
          acc[k][1] = __ffma2_rn(dv, xs1, acc[k][1]);
          acc[k][2] = __ffma2_rn(dv, xs2, acc[k][2]);
          if (tap == 4) bs[k] = __fadd2_rn(bs[k], dv);      // every tap sees the same dy: the centre tap takes the bias sums
        }
      }
    }
  }
  __syncthreads();
  if (active) {
    float* mine = s_acc + (size_t)tid * SW2_VALS;
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) { mine[(k * 3 + c) * 2] = acc[k][c].x; mine[(k * 3 + c) * 2 + 1] = acc[k][c].y; }
    if (tap == 4) {
      float* bm = s_acc + (size_t)SW2_LANES * 9 * SW2_VALS + lane_px * 14;
#pragma unroll
      for (int k = 0; k < 7; ++k) { bm[2 * k] = bs[k].x; bm[2 * k + 1] = bs[k].y; }
    }
  }
  __syncthreads();
  for (int i = tid; i < 9 * 42; i += blockDim.x) {       // (tap, value) summed over the pixel lanes, one atomic per weight and block
    const int tp = i / 42, v = i % 42;
    float t = 0.f;
    for (int l = 0; l < SW2_LANES; ++l) t += s_acc[(size_t)(l * 9 + tp) * SW2_VALS + v];
    const int k = v / 6, c = (v % 6) / 2, co = 2 * k + (v & 1);
    if (co < STEM_CO) atomicAdd(&dw[(co * STEM_CI + c) * 9 + tp], t);
  }
  if (db && tid < STEM_CO) {
    float t = 0.f;
    for (int l = 0; l < SW2_LANES; ++l) t += s_acc[(size_t)SW2_LANES * 9 * SW2_VALS + l * 14 + tid];
    atomicAdd(&db[tid], t);
  }
}

}  // namespace erfk

using namespace erfk;

static int stem_fwd_launch(const void* x, bool nchw_in, int N, int H, int W, const float* w, const float* bias, const float* scale,
                           const float* shift, int relu, void* y, double* stats, int dtype, erf_stream_t stream, const char* name) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "%s: bad dtype %d", name, dtype);
  ERF_CHECK_ARG(x && w && y && N > 0 && H > 0 && W > 0 && (H % 2) == 0 && (W % 4) == 0, "%s: bad arguments (H must be even, W a multiple of 4)", name);
  ERF_CHECK_ARG((scale == nullptr) == (shift == nullptr), "%s: scale and shift must be given together", name);
  ERF_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0, "%s: tensors must be 16-byte aligned", name);
  const long long total = (long long)N * (H / 2) * (W / 4);
  ERF_CHECK_ARG(total * 2 < 0x7fffffffLL, "%s: %lld output pixels exceed the 32-bit index range", name, total * 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 g((unsigned)blocks), b(256);
#define STEM_LAUNCH(T_, IN_, ST_) launch_pdl(stem_fwd_kernel<T_, IN_, ST_>, g, b, 0, s, x, N, H, W, w, bias, scale, shift, relu, (T_*)y, stats)
#define STEM_DISPATCH(T_)                                                    \
  if (nchw_in) { if (stats) STEM_LAUNCH(T_, true, true); else STEM_LAUNCH(T_, true, false); } \
  else { if (stats) STEM_LAUNCH(T_, false, true); else STEM_LAUNCH(T_, false, false); }
  if (dtype == ERF_F32) { STEM_DISPATCH(float) } else { STEM_DISPATCH(__nv_bfloat16) }
#undef STEM_DISPATCH
#undef STEM_LAUNCH
  return ERF_OK;
}

extern "C" int erf_stem_fwd(const void* x, int N, int H, int W, const float* w, const float* bias, const float* scale, const float* shift,
                            int relu, void* y, double* stats, int dtype, erf_stream_t stream) {
  const int rc = stem_fwd_launch(x, false, N, H, W, w, bias, scale, shift, relu, y, stats, dtype, stream, "erf_stem_fwd");
  if (rc != ERF_OK) return rc;
  ERF_CHECK_LAUNCH("erf_stem_fwd");
  return ERF_OK;
}

extern "C" int erf_stem_fwd_nchw(const float* x, int N, int H, int W, const float* w, const float* bias, const float* scale, const float* shift,
                                 int relu, void* y, double* stats, int dtype, erf_stream_t stream) {
  const int rc = stem_fwd_launch(x, true, N, H, W, w, bias, scale, shift, relu, y, stats, dtype, stream, "erf_stem_fwd_nchw");
  if (rc != ERF_OK) return rc;
  ERF_CHECK_LAUNCH("erf_stem_fwd_nchw");
  return ERF_OK;
}

extern "C" int erf_stem_wgrad(const void* x, const void* dy, int N, int H, int W, float* dw, float* db, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_stem_wgrad: bad dtype %d", dtype);
  ERF_CHECK_ARG(x && dy && dw && N > 0 && H > 0 && W > 0 && (H % 2) == 0 && (W % 4) == 0, "erf_stem_wgrad: bad arguments (H even, W % 4 == 0)");
  ERF_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0, "erf_stem_wgrad: tensors must be 16-byte aligned");
  const size_t esz = dtype == ERF_F32 ? 4 : 2;
  size_t smem = (size_t)3 * W * 4 * esz + (size_t)(W / 2) * STEM_C * esz;
  if (smem < (size_t)SW_LANES * SW_GROUP * 16 * sizeof(float)) smem = (size_t)SW_LANES * SW_GROUP * 16 * sizeof(float);
  ERF_CHECK_ARG(smem <= 200 * 1024, "erf_stem_wgrad: image row of %d pixels does not fit in shared memory", W);
  const int rows = N * (H / 2);
  // blocks per SM: the kernel alternates staging and compute behind block-wide barriers, so co-resident blocks are what overlaps the two
  static const int per_sm_env = getenv("ERF_STEM_WGRAD_BLOCKS") ? atoi(getenv("ERF_STEM_WGRAD_BLOCKS")) : 0;
  const int per_sm = per_sm_env > 0 ? per_sm_env : 4;   // measured at the bench shape: 122 (2) / 93 (3) / 91 (4) / 96 us (5)
  int grid = rows < 148 * per_sm ? rows : 148 * per_sm;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (dtype == ERF_F32) {
    e = cudaFuncSetAttribute(stem_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) launch_pdl(stem_wgrad_kernel<float>, dim3(grid), dim3(256), smem, (cudaStream_t)(s), (const float*)x, (const float*)dy, N, H, W, dw, db);
  } else {
    static const bool v1 = getenv("ERF_STEM_WGRAD_V1") && atoi(getenv("ERF_STEM_WGRAD_V1")) != 0;
    if (v1) {
      e = cudaFuncSetAttribute(stem_wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess) launch_pdl(stem_wgrad_kernel<__nv_bfloat16>, dim3(grid), dim3(256), smem, (cudaStream_t)(s), (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, N, H, W, dw, db);
    } else {
      size_t smem2 = (size_t)(SW2_LANES * 9 * SW2_VALS + SW2_LANES * 14) * sizeof(float);
      if (smem2 < smem) smem2 = smem;
      if (per_sm_env <= 0) grid = rows < 148 * 2 ? rows : 148 * 2;   // measured at the bench shape: 55.7 us (2 per SM) / 60.0 us (3) against 91 us for the first version
      e = cudaFuncSetAttribute(stem_wgrad_bf16_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e == cudaSuccess) launch_pdl(stem_wgrad_bf16_v2_kernel, dim3(grid), dim3(256), smem2, (cudaStream_t)(s), (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, N, H, W, dw, db);
    }
  }
  if (e != cudaSuccess) {
    set_error("erf_stem_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
  ERF_CHECK_LAUNCH("erf_stem_wgrad");
  return ERF_OK;
}
