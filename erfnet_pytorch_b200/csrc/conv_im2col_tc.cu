// Generic tcgen05 implicit-GEMM convolution with SOFTWARE im2col, for every conv of the net that is not a 3-tap 1-D conv:
// the 3x3 stride-2 convs of the downsamplers (eval/erfnet.py:16), the 3x3 stride-2 transposed convs of the upsamplers
// (l.102), the 2x2 stride-2 transposed head (l.124), the 1x1 encoder head (l.85) and all their data gradients.
//
// One launch covers one "class" of output pixels that share a tap list:
//   * pixels are enumerated on a grid (EN, EH, EW); enumerated pixel (n, y, x) reads input pixel
//     (in_step*y + tap.dy, in_step*x + tap.dx) for each tap and writes output pixel (out_step*y + out_oy, out_step*x + out_ox)
//       strided conv      : grid = output grid,        in_step = stride, tap = (ky*dil - pad, kx*dil - pad), out_step = 1
//       transposed stride 2: grid = output grid / 2 per parity class (a,b), in_step = 1, taps = the 1/2/2/4 taps that hit
//                            that parity (sub-pixel form), out_step = 2, out offset = (a,b)
//   * unit = 128 consecutive enumerated pixels = one UMMA M block; per tap the A tile [128 px][Cin] is gathered by four
//     loader warps with 16-byte cp.async (zero-fill outside the image = padding) straight into the swizzled K-major layout
//     (128B swizzle for 64-channel chunks, 32B swizzle for 16-channel chunks), one ring slot per tap;
//   * the class's tap weight matrices are resident in shared memory (bulk copy of a host-packed image);
//   * warp 4 issues the MMAs (N = padded Cout, K = Cin per tap) into one of two TMEM accumulators; warps 5-8 run the
//     epilogue (bias, ReLU, bf16, per-pixel 16-byte stores into the NHWC output slice).
// These layers are 6 % of the network's FLOPs; the point of this kernel is to get them off the CUDA cores, not peak MMA rate.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace erfk {
using namespace tc;

constexpr int kImMaxSlots = 8;
constexpr int kImThreads = 288;  // warps 0-3 loaders, warp 4 MMA, warps 5-8 epilogue

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_g2s_im(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// TMA = true: a tap tile (128 enumerated pixels of ONE input row, every in_step_x-th pixel) is a strided box of a tiled tensor map -- one
// cp.async.bulk.tensor per (tap, k-chunk) issued by an elected lane of warp 0, zero fill outside the image = padding, hardware swizzle =
// the K-major operand layout -- instead of 128 threads computing addresses for 16-byte cp.async pieces.  Needs EW % 128 == 0 (a unit never
// straddles rows) and 16-byte aligned pixel rows; other shapes keep the software gather below.
template <int NP, bool NCHW = false, bool TMA = false>
__global__ void __launch_bounds__(kImThreads, 1) conv_im2col_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const erf_im2col_desc p, const int slots,
                                                                       const int slot_bytes, const int w_bytes) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + ((w_bytes + 1023) / 1024) * 1024;
  __shared__ uint64_t full_bar[kImMaxSlots], empty_bar[kImMaxSlots], tfull_bar[2], tempty_bar[2], w_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[NP];
  constexpr uint32_t TMEM_COLS = 2 * NP < 32 ? 32 : 2 * NP;
  constexpr uint32_t IDESC = make_idesc_bf16(128, NP);
  const int CH = (p.Cin % 64) == 0 ? 64 : 16;   // channels per K chunk
  const int RB = CH * 2, KC = p.Cin / CH, KSTEPS = CH / 16, CPR = RB / 16;
  const uint32_t layout = RB == 128 ? LAYOUT_SW128 : LAYOUT_SW32;
  const long long total = (long long)p.EN * p.EH * p.EW;
  // TMA loaders walk ROW units (ceil(EW / 128) per enumerated row, the last one possibly ragged: zero-filled loads, skipped stores); the
  // software gather walks the enumerated pixels linearly.  p.units is set accordingly by the host.
  const int row_segs = (p.EW + 127) / 128;
  auto decode = [&](int u, int row, int& n, int& y, int& x) -> bool {   // enumerated pixel of (unit, tile row); false = outside the grid
    if (TMA) {
      const unsigned t2 = (unsigned)u / (unsigned)row_segs;
      x = (int)((unsigned)u - t2 * (unsigned)row_segs) * 128 + row;
      n = (int)(t2 / (unsigned)p.EH);
      y = (int)(t2 - (unsigned)n * (unsigned)p.EH);
      return x < p.EW;
    }
    const unsigned q = (unsigned)u * 128u + (unsigned)row;
    if (q >= (unsigned)total) return false;
    const unsigned t2 = q / (unsigned)p.EW;
    x = (int)(q - t2 * (unsigned)p.EW);
    n = (int)(t2 / (unsigned)p.EH);
    y = (int)(t2 - (unsigned)n * (unsigned)p.EH);
    return true;
  };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NP; i += blockDim.x)
    s_bias[i] = (p.bias && i < p.n_valid) ? (p.bias_period > 0 ? p.bias[(p.out_c0 + i) % p.bias_period] : p.bias[i]) : 0.f;
  const int in_step_x = p.in_step_x > 0 ? p.in_step_x : p.in_step, out_step_x = p.out_step_x > 0 ? p.out_step_x : p.out_step;
  if (tid == 0) {
    for (int s = 0; s < slots; ++s) {
      mbar_init(&full_bar[s], TMA ? 1 : 128);   // one arrival per loader thread, or the TMA producer's expect_tx
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    mbar_init(&w_bar, 1);
    fence_barrier_init();
    if (TMA) tma_prefetch_desc(&tmap_in);
  }
  __syncthreads();
  if (warp == 4) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();

  if (TMA && warp < 4) {
    // ================= TMA producer (warp 0; uniform control flow, one elected lane issues) =================
    if (warp == 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_bar, w_bytes);
        for (int off = 0; off < w_bytes; off += 32768) {
          const int nb = w_bytes - off < 32768 ? w_bytes - off : 32768;
          bulk_g2s_im(s_w + off, reinterpret_cast<const uint8_t*>(p.w_img) + off, nb, &w_bar);
        }
      }
      __syncwarp();
      int slot = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(KC * 128 * RB);
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const unsigned t2 = (unsigned)u / (unsigned)row_segs, nn = t2 / (unsigned)p.EH;
        const int x0 = in_step_x * (int)((unsigned)u - t2 * (unsigned)row_segs) * 128, y0 = p.in_step * (int)(t2 - nn * (unsigned)p.EH);
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          if (elect_one()) {
            uint8_t* base = s_a + (size_t)slot * slot_bytes;
            mbar_arrive_expect_tx(&full_bar[slot], tx);
            for (int kc = 0; kc < KC; ++kc)
              tma_load_4d(base + kc * (128 * RB), &tmap_in, &full_bar[slot], kc * CH, x0 + p.tap_dx[t], y0 + p.tap_dy[t], (int)nn);
          }
          __syncwarp();
          if (++slot == slots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    // ================= loaders: 128 threads gather one tap tile per ring slot =================
    if (tid == 0) {
      mbar_arrive_expect_tx(&w_bar, w_bytes);
      for (int off = 0; off < w_bytes; off += 32768) {
        const int nb = w_bytes - off < 32768 ? w_bytes - off : 32768;
        bulk_g2s_im(s_w + off, reinterpret_cast<const uint8_t*>(p.w_img) + off, nb, &w_bar);
      }
    }
    // thread t copies 16-byte chunk j = t % CPR of rows r_i = i * (128 / CPR) + t / CPR
    const int rows_per_iter = 128 / CPR, iters = CPR;   // CPR = 8: 8 rows/thread; CPR = 2: 2 rows/thread
    const int j = tid % CPR, r0 = tid / CPR;
    int slot = 0;
    uint32_t phase = 0;
    // completion lags issue by `lag` tap tiles so that the gathers of consecutive taps overlap (needs slots > lag)
    const int lag = slots > 2 ? 2 : 1;
    int pend[2] = {-1, -1};
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(p.in);
    // Everything that does not depend on the tap is hoisted out of the tap loop: with ONE loader warp per scheduler the gather is bound
    // by the instruction latency of its address arithmetic (ncu: ~120 SASS instructions per cp.async before this), not by memory.
    uint32_t dst_off[8];   // shared-memory offset of this thread's chunk of row i inside a ring slot (swizzled), loop invariant
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = i * rows_per_iter + r0;
      const int key = CPR == 8 ? (row & 7) : ((row >> 2) & 1);
      dst_off[i] = (uint32_t)(row * RB + ((j ^ key) * 16));
    }
    const uint32_t s_a_addr = smem_u32(s_a);
    const uint32_t kc_stride = 128u * RB;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      // decode this thread's rows once per unit: one division pair for the first row, the others by stepping (rows_per_iter pixels apart)
      int y0[8], x0[8];                  // input coordinates of tap (0,0); y0 very negative = row outside the enumerated grid
      const __nv_bfloat16* src0[8];      // address of that pixel's chunk j
      {
        // 32-bit index arithmetic (the host checks total < 2^31): 64-bit div/mod here was the kernel's critical path
        unsigned q = (unsigned)u * 128u + (unsigned)r0;
        const unsigned t2 = q / (unsigned)p.EW;
        int x = (int)(q - t2 * (unsigned)p.EW);
        int n = (int)(t2 / (unsigned)p.EH);
        int y = (int)(t2 - (unsigned)n * (unsigned)p.EH);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < iters) {
            const bool inside = q < (unsigned)total;
            y0[i] = inside ? p.in_step * y : -(1 << 28);
            x0[i] = in_step_x * x;
            src0[i] = in + (((size_t)n * p.H + (inside ? y0[i] : 0)) * p.W + x0[i]) * p.in_ld + j * 8;
            q += (unsigned)rows_per_iter;
            x += rows_per_iter;
            while (x >= p.EW) {
              x -= p.EW;
              if (++y == p.EH) { y = 0; ++n; }
            }
          }
        }
      }
      for (int t = 0; t < p.ntaps; ++t) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        const uint32_t base = s_a_addr + (uint32_t)slot * (uint32_t)slot_bytes;
        const int dy = p.tap_dy[t], dx = p.tap_dx[t];
        const long long toff = ((long long)dy * p.W + dx) * p.in_ld;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < iters) {
            const bool ok = (unsigned)(y0[i] + dy) < (unsigned)p.H && (unsigned)(x0[i] + dx) < (unsigned)p.W;
            const __nv_bfloat16* src = ok ? src0[i] + toff : in;
            const uint32_t nb = ok ? 16u : 0u;
            uint32_t dst = base + dst_off[i];
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nb) : "memory");
            if (KC > 1) {
              const int cstep = ok ? CH : 0;
#pragma unroll 1
              for (int kc = 1; kc < KC; ++kc) {
                dst += kc_stride;
                src += cstep;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nb) : "memory");
              }
            }
          }
        }
        cp_async_commit();
        // retire the tile issued `lag` steps ago
        if (lag == 2) {
          if (pend[1] >= 0) { cp_async_wait<2>(); fence_proxy_async_smem(); mbar_arrive(&full_bar[pend[1]]); }
          pend[1] = pend[0];
          pend[0] = slot;
        } else {
          if (pend[0] >= 0) { cp_async_wait<1>(); fence_proxy_async_smem(); mbar_arrive(&full_bar[pend[0]]); }
          pend[0] = slot;
        }
        if (++slot == slots) { slot = 0; phase ^= 1; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    if (lag == 2 && pend[1] >= 0) mbar_arrive(&full_bar[pend[1]]);
    if (pend[0] >= 0) mbar_arrive(&full_bar[pend[0]]);
  } else if (warp == 4) {
    // ================= MMA issuer =================
    {   // whole warp: uniform control flow, one elected lane issues (see tc.cuh elect_one)
      int slot = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      const uint32_t w_addr = smem_u32(s_w);
      const uint32_t desc_hi = ((8u * RB) >> 4) | (1u << 14) | (layout << 29);
      mbar_wait(&w_bar, 0);
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        const uint32_t d_tmem = tmem_base + acc * NP;
        uint32_t first = 0;
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&full_bar[slot], phase);
          fence_proxy_async_smem();
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(s_a + (size_t)slot * slot_bytes);
          for (int kc = 0; kc < KC; ++kc) {
            const uint32_t a_lo = (((a_addr + kc * (128 * RB)) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t b_lo = (((w_addr + (t * KC + kc) * (NP * RB)) >> 4) & 0x3FFF) | (1u << 16);
            for (int k = 0; k < KSTEPS; ++k) {
              if (elect_one()) umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + k * 2), ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + k * 2), IDESC, first);
              first = 1;
            }
          }
          if (elect_one()) umma_commit(&empty_bar[slot]);
          if (++slot == slots) { slot = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================= epilogue: warps 5..8 (TMEM quadrant = warp % 4) =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      __nv_bfloat16* dst = nullptr;
      int pn = 0, py = 0, px = 0;
      const bool inside = decode(u, row, pn, py, px);
      if (inside) dst = outp + (((size_t)pn * p.Ho + (p.out_step * py + p.out_oy)) * p.Wo + (out_step_x * px + p.out_ox)) * p.out_ld + p.out_c0;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * NP + ((uint32_t)(quad * 32) << 16);
      if constexpr (NCHW) {
        // NCHW fp32 output of a row-parity class (N index = b * NP/2 + c): lane = input column x, so the float2 (b = 0, 1) of one channel is
        // 8 contiguous bytes per lane and 256 contiguous bytes per warp -- the logits never take the NHWC bf16 round trip
        float* o = nullptr;
        const size_t plane = (size_t)p.Ho * (size_t)(2 * p.Wo);
        if (inside) o = p.nchw_out + ((size_t)pn * p.nchw_C * p.Ho + (size_t)(p.out_step * py + p.out_oy)) * (size_t)(2 * p.Wo) + 2 * px;
#pragma unroll
        for (int c0 = 0; c0 < NP / 2; c0 += 16) {   // 16 channels at a time: columns c0.. of parity 0 and NP/2 + c0.. of parity 1
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr + c0, r0);
          tmem_ld16(taddr + NP / 2 + c0, r1);
          tmem_ld_wait();
          if (o != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (c0 + e < p.nchw_C) {
                float v0 = __uint_as_float(r0[e]) + s_bias[c0 + e], v1 = __uint_as_float(r1[e]) + s_bias[NP / 2 + c0 + e];
                if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                *reinterpret_cast<float2*>(o + (size_t)(c0 + e) * plane) = make_float2(v0, v1);
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
#pragma unroll
      for (int c0 = 0; c0 < NP; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        if (dst != nullptr && c0 < p.n_valid) {
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            v[e] = __uint_as_float(r[e]) + s_bias[c0 + e];
            if (p.relu) v[e] = fmaxf(v[e], 0.f);
          }
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            if (c0 + h8 * 8 < p.n_valid) {
              uint4 t4;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t4);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[h8 * 8 + 2 * e], v[h8 * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(dst + c0 + h8 * 8) = t4;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Weight image [ntaps][KC][NP][CH] (bf16, swizzled) from an fp32 torch weight: element (n, k, tap) = w[n*stride_n + k*stride_k + tap_off[t]]
// for n < n_src, k < k_src (zero otherwise).
__global__ void pack_im2col_weights_kernel(const float* __restrict__ w, long long stride_n, long long stride_k, int n0, int n_src, int k_src,
                                           erf_im2col_taps taps, int Cin, int NP, __nv_bfloat16* __restrict__ img) {
  pdl_enter();
  const int CH = (Cin % 64) == 0 ? 64 : 16, CPR = CH / 8, KC = Cin / CH;
  const int total = taps.ntaps * KC * NP * CPR;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int jj = i % CPR, n = (i / CPR) % NP, kc = (i / (CPR * NP)) % KC, t = i / (CPR * NP * KC);
  const int j = CPR == 8 ? (jj ^ (n & 7)) : (jj ^ ((n >> 2) & 1));
  Vec<__nv_bfloat16, 8> v;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kc * CH + j * 8 + e;
    int nn = n0 + n;
    long long off = taps.off[t];
    if (taps.n_split > 0) {   // column-folded image: n = b * n_split + c
      if (nn >= taps.n_split) { nn -= taps.n_split; off = nn < taps.n_split ? taps.off2[t] : -1; }
    }
    v.v[e] = (off >= 0 && nn < n_src && k < k_src) ? w[(long long)nn * stride_n + (long long)k * stride_k + off] : 0.f;
  }
  v.store(img + (size_t)i * 8);
}

typedef CUresult (*ImEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC bf16 input [N][H][W][ld] as (C, W, H, N); box = {ch, box_w, 1, 1} traversed with column step step_w: box_w / step_w pixels of one row
static int im_make_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int ld, int ch, int box_w, int step_w) {
  static ImEncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (ImEncodeFn)q;
  }
  if (!fn || box_w > 256 || (ld % 8) != 0) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)ch, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)step_w, 1, 1};
  return (int)fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace erfk

using namespace erfk;

extern "C" int erf_conv_im2col_tc(const erf_im2col_desc* d, erf_stream_t stream) {
  ERF_CHECK_ARG(d && d->in && d->out && d->w_img, "erf_conv_im2col_tc: null pointer");
  ERF_CHECK_ARG(d->nchw_out == nullptr || ((d->NP == 32 || d->NP == 64) && d->out_step == 2 && d->out_step_x == 1 && d->bias_period == d->NP / 2 && d->nchw_C > 0 &&
                                            d->nchw_C <= d->NP / 2 && d->n_valid == d->NP),
                "erf_conv_im2col_tc: NCHW output needs the row-parity form with n_split = NP/2 <= 32 and nchw_C <= NP/2");
  ERF_CHECK_ARG(d->Cin >= 16 && d->Cin <= 128 && (d->Cin % 16) == 0 && (d->in_ld % 8) == 0 && d->in_ld >= d->Cin,
                "erf_conv_im2col_tc: Cin=%d must be a multiple of 16 in [16,128] and in_ld=%d a multiple of 8", d->Cin, d->in_ld);
  ERF_CHECK_ARG(d->NP == 16 || d->NP == 32 || d->NP == 64 || d->NP == 128, "erf_conv_im2col_tc: padded N=%d must be 16, 32, 64 or 128", d->NP);
  ERF_CHECK_ARG(d->n_valid > 0 && d->n_valid <= d->NP && (d->n_valid % 8) == 0 && ((d->out_ld | d->out_c0) % 8) == 0,
                "erf_conv_im2col_tc: n_valid / out_ld / out_c0 must be multiples of 8");
  ERF_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= 9 && d->EN > 0 && d->EH > 0 && d->EW > 0, "erf_conv_im2col_tc: bad tap count or grid");
  const int CH = (d->Cin % 64) == 0 ? 64 : 16, RB = CH * 2, KC = d->Cin / CH;
  const int w_bytes = d->ntaps * KC * d->NP * RB;
  const int w_pad = ((w_bytes + 1023) / 1024) * 1024;
  const int slot_bytes = KC * 128 * RB;
  int slots = (200 * 1024 - w_pad) / slot_bytes;
  if (slots > kImMaxSlots) slots = kImMaxSlots;
  {
    // two CTAs per SM (twice the gathers in flight, epilogue of one unit under the gather of another) whenever a ring of >= 3 slots fits in
    // half of the shared memory; ERF_IM2COL_2CTA=0 keeps the deep single-CTA ring
    static const bool two = !(getenv("ERF_IM2COL_2CTA") && atoi(getenv("ERF_IM2COL_2CTA")) == 0);
    const int half = (100 * 1024 - 1024 - w_pad) / slot_bytes;
    if (two && 2 * d->NP <= 128 && half >= 3 && half < slots) slots = half;
  }
  if (slots < 2) {
    set_error("erf_conv_im2col_tc: %d taps x N=%d x Cin=%d of weights do not fit in shared memory; split the class", d->ntaps, d->NP, d->Cin);
    return ERF_ERR_UNSUPPORTED;
  }
  const long long total = (long long)d->EN * d->EH * d->EW;
  ERF_CHECK_ARG(total + 128 < 0x7fffffffLL, "erf_conv_im2col_tc: %lld output pixels exceed the 32-bit index range", total);
  erf_im2col_desc p = *d;
  p.units = (int)((total + 127) / 128);
  size_t smem = (size_t)w_pad + (size_t)slots * slot_bytes + 1024;
  // small classes (<= 100 KB of shared memory, <= 256 TMEM columns for two CTAs) run two CTAs per SM: their per-unit
  // gather -> MMA -> epilogue chain is latency-bound (shrinking the ring of the large classes to fit two CTAs measured neutral)
  const int per_sm = (smem <= 100 * 1024 && 2 * d->NP <= 128) ? 2 : 1;
  int grid = p.units < 148 * per_sm ? p.units : 148 * per_sm;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  // TMA-loaded tap tiles: units of 128 pixels inside one row, 16-byte aligned rows
  static const bool tma_off = getenv("ERF_IM2COL_TMA") && atoi(getenv("ERF_IM2COL_TMA")) == 0;
  const int isx = d->in_step_x > 0 ? d->in_step_x : d->in_step;
  CUtensorMap tm;
  bool tma = !tma_off && d->EW >= 16 && ((uintptr_t)d->in % 16) == 0 && isx >= 1 && isx <= 2 && d->in_step >= 1;
  if (tma) tma = im_make_map(&tm, d->in, d->EN, d->H, d->W, d->in_ld, CH, 128 * isx, isx) == 0;
  if (!tma) memset(&tm, 0, sizeof(tm));
  if (tma) {   // row units: ceil(EW / 128) per enumerated row
    const long long ru = (long long)d->EN * d->EH * ((d->EW + 127) / 128);
    ERF_CHECK_ARG(ru < 0x7fffffffLL, "erf_conv_im2col_tc: too many row units");
    p.units = (int)ru;
    grid = p.units < 148 * per_sm ? p.units : 148 * per_sm;
  }
#define IM_LAUNCH2(NP_, NCHW_, TMA_)                                                                                                 \
  e = cudaFuncSetAttribute(conv_im2col_tc_kernel<NP_, NCHW_, TMA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
  if (e == cudaSuccess && per_sm > 1) e = prefer_max_smem(conv_im2col_tc_kernel<NP_, NCHW_, TMA_>);                                   \
  if (e == cudaSuccess) launch_pdl(conv_im2col_tc_kernel<NP_, NCHW_, TMA_>, dim3(grid), dim3(kImThreads), smem, (cudaStream_t)(st), tm, p, slots, slot_bytes, w_bytes);
#define IM_LAUNCH(NP_, NCHW_) if (tma) { IM_LAUNCH2(NP_, NCHW_, true) } else { IM_LAUNCH2(NP_, NCHW_, false) }
  if (d->nchw_out != nullptr) {
    if (d->NP == 32) { IM_LAUNCH(32, true) } else { IM_LAUNCH(64, true) }
  }
  else if (d->NP == 16) { IM_LAUNCH(16, false) }
  else if (d->NP == 32) { IM_LAUNCH(32, false) }
  else if (d->NP == 64) { IM_LAUNCH(64, false) }
  else { IM_LAUNCH(128, false) }
#undef IM_LAUNCH
#undef IM_LAUNCH2
  if (e != cudaSuccess) {
    set_error("erf_conv_im2col_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
  ERF_CHECK_LAUNCH("erf_conv_im2col_tc");
  return ERF_OK;
}

extern "C" int erf_pack_im2col_weights(const float* w, long long stride_n, long long stride_k, int n0, int n_src, int k_src,
                                       const erf_im2col_taps* taps, int Cin, int NP, void* img, erf_stream_t stream) {
  ERF_CHECK_ARG(w && taps && img && taps->ntaps >= 1 && taps->ntaps <= 9 && (Cin % 16) == 0 && Cin <= 128, "erf_pack_im2col_weights: bad arguments");
  const int CH = (Cin % 64) == 0 ? 64 : 16;
  const int total = taps->ntaps * (Cin / CH) * NP * (CH / 8);
  launch_pdl(pack_im2col_weights_kernel, dim3((total + 255) / 256), dim3(256), 0, (cudaStream_t)((cudaStream_t)stream), w, stride_n, stride_k, n0, n_src, k_src, *taps, Cin, NP,
                                                                                   (__nv_bfloat16*)img);
  ERF_CHECK_LAUNCH("erf_pack_im2col_weights");
  return ERF_OK;
}
