// Generic tcgen05 weight gradient with software gather, for the convs served by conv_im2col_tc.cu (3x3 stride-2 convs,
// stride-2 transposed convs, the stem):
//
//   dW[tap][cg][cd] = sum over the DENSE-side pixels q of  G[gather(q, tap)][cg] * D[q][cd]
//
//   conv  (stride s): dense D = dy on the output grid, gathered G = x at (s*oy - pad + ky*dil, s*ox - pad + kx*dil)
//   convT (stride s): dense D = x  on the input  grid, gathered G = dy at (s*iy - pad + ky,     s*ix - pad + kx)
//   (in both cases the torch parameter index is ((cd*Cg + cg)*KH*KW + ky*KW + kx), so one reduce kernel serves both).
//
// Both operands are NHWC pixel tiles, i.e. MN-major for this product (channels contiguous, pixels = K).  The gathered
// operand sits on the M side: its tap tiles are adjacent in shared memory at a uniform pitch, so the UMMA descriptor's
// leading-byte-offset walks "M atoms" = (tap, channel sub-tile) pairs -- one M=128 MMA covers 2 taps of a 64-channel tensor
// or 8 taps of a 16-channel one at full tensor-core width (the same trick as conv1d_wgrad_tc.cu).  Each group of 128 M rows
// owns one TMEM accumulator [128 x NP]; a persistent CTA accumulates over all its 64-pixel units and dumps once.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace erfk {
using namespace tc;

constexpr int kWiThreads = 288;  // warps 0-3 loaders, warp 4 MMA, warps 5-8 epilogue
constexpr int kWiPX = 64;        // pixels (K) per unit
constexpr int kWiMaxStages = 4;

__device__ __forceinline__ void wi_cp_async(void* smem_dst, const void* gsrc, int bytes, uint32_t src_bytes) {
  if (bytes == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}

// smem stage layout: [gathered: ntaps*gsubs sub-tiles of 64 px x CHg][dense: dsubs sub-tiles of 64 px x CHd]
// MINB = CTAs per SM the launch is sized for: 2 caps the registers at 113 (small layers, latency-bound per-unit chain), 1 leaves the loaders'
// hoisted index tables in registers (large layers, one 181 KB ring per SM)
// TMA = true: the tap tiles and the dense tile arrive by cp.async.bulk.tensor from two tensor maps (built per launch by the host).  A strided
// gather is a TILED map with element strides: boxDim = 64 * step along W with elementStrides = step loads every step-th pixel, rows / columns
// outside the tensor are zero-filled (= the conv padding), and the hardware swizzle reproduces the operand layout the software gather
// wrote.  One elected lane then issues ntaps * gsubs + dsubs copies per 64-pixel unit, where 128 loader threads spent ~600 instructions per
// unit on index arithmetic (ncu: the loaders' latency-bound instruction stream set the kernel time, 94 us for the head against a 27 us
// DRAM floor).  Needs row-aligned units (EW % 64 == 0) and 16-byte aligned pixel rows; everything else keeps the cp.async gather.
template <int MINB, bool TMA>
__global__ void __launch_bounds__(kWiThreads, MINB) conv_wgrad_im2col_tc_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_d,
                                                                             const erf_wgrad_im2col_desc p, const int stages, const int stage_bytes,
                                                                             const int dense_off, const int groups, float* __restrict__ ws) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[kWiMaxStages], empty_bar[kWiMaxStages], done_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_row[kWiMaxStages][kWiPX][3];   // (n, y, x) of the unit's pixels, -1 = beyond the end
  const int CHg = (p.CgP % 64) == 0 ? 64 : 16, RBg = CHg * 2, gsubs = p.CgP / CHg;
  const int CHd = (p.NP % 64) == 0 ? 64 : 16, RBd = CHd * 2, dsubs = p.NP / CHd;
  const int g_sub_bytes = kWiPX * RBg, d_sub_bytes = kWiPX * RBd;
  const long long total = (long long)p.EN * p.EH * p.EW;
  // TMA loaders walk ROW units: ceil(EW / 64) per image row, the last one possibly ragged (zero-filled dense rows contribute nothing);
  // the software gather walks the pixels linearly
  const int row_segs = (p.EW + kWiPX - 1) / kWiPX;
  const int units = TMA ? p.EN * p.EH * row_segs : (int)((total + kWiPX - 1) / kWiPX);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < groups * p.NP) tmem_cols <<= 1;

  // zero the whole staging area once: rows shorter than their swizzle row / padded channels / garbage atoms stay finite
  for (int i = tid; i < (stages * stage_bytes + p.tail_bytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], TMA ? 1 : 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
    if (TMA) {
      tma_prefetch_desc(&tmap_g);
      tma_prefetch_desc(&tmap_d);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 4) tmem_alloc(&tmem_base_s, tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();

  if (TMA && warp < 4) {
    // ================= TMA producer: warp 0, one elected lane issues (uniform control flow) =================
    if (warp == 0) {
      const int gsx = p.g_step_x > 0 ? p.g_step_x : p.g_step;
      const uint32_t tx = (uint32_t)(p.ntaps * gsubs * g_sub_bytes + dsubs * d_sub_bytes);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const unsigned t2 = (unsigned)u / (unsigned)row_segs, nn = t2 / (unsigned)p.EH;
        const int x0 = (int)((unsigned)u - t2 * (unsigned)row_segs) * kWiPX, y0 = (int)(t2 - nn * (unsigned)p.EH);
        uint8_t* base = smem + (size_t)stage * stage_bytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          for (int t = 0; t < p.ntaps; ++t)
            for (int sub = 0; sub < gsubs; ++sub)
              tma_load_4d(base + (t * gsubs + sub) * g_sub_bytes, &tmap_g, &full_bar[stage], sub * CHg, gsx * x0 + p.tap_dx[t], p.g_step * y0 + p.tap_dy[t], (int)nn);
          for (int sub = 0; sub < dsubs; ++sub)
            tma_load_4d(base + dense_off + sub * d_sub_bytes, &tmap_d, &full_bar[stage], p.d_c0 + sub * CHd, x0, y0, (int)nn);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 4) {
    // ================= loaders =================
    const uint8_t* G = reinterpret_cast<const uint8_t*>(p.g);
    const uint8_t* D = reinterpret_cast<const uint8_t*>(p.d);
    const int g_row_bytes = p.g_copy_bytes;                 // bytes copied per gathered pixel (multiple of 16, or 8)
    const int g_chunk = g_row_bytes >= 16 ? 16 : 8;
    const int g_chunks = g_row_bytes / g_chunk;
    const int d_chunks = p.NP / 8;                          // 16-byte chunks per dense pixel
    int stage = 0;
    uint32_t phase = 0;
    // completion lag: a unit's copies are awaited (cp.async.wait_group) only after LAG later units have been issued, so LAG + 1 units of
    // gathers are in flight per CTA -- the small layers' units are 14-24 KB, far too little for one unit in flight to cover DRAM latency
    const int lag = stages >= 4 ? 2 : 1;
    int pend[2] = {-1, -1};                                  // stages of the units issued but not yet signalled (oldest first)
    const int gsx = p.g_step_x > 0 ? p.g_step_x : p.g_step;
    // Loop-invariant part of this thread's copy items, hoisted (same reasoning as conv_im2col_tc.cu: with one loader warp per scheduler the
    // gather is bound by the instruction latency of its index arithmetic).  Items (tap, row, chunk) are dealt round-robin to the 128 loader
    // threads; when one tap's items (64 * g_chunks) are a multiple of 128, a thread owns the SAME (row, chunk) pairs in every tap and unit.
    const int per_tap = kWiPX * g_chunks;
    const bool fast_g = (per_tap % 128) == 0 && per_tap / 128 <= 8;
    const int gm = fast_g ? per_tap / 128 : 0;
    int g_row[8];
    uint32_t g_dst[8], g_srcoff[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int w = tid + 128 * m, ch = w % g_chunks, row = (w / g_chunks) % kWiPX;
      const int byte0 = ch * g_chunk, c16 = byte0 / 16, sub = c16 / (CHg / 8), cj = c16 % (CHg / 8);
      const int key = RBg == 128 ? (row & 7) : ((row >> 2) & 1);
      g_row[m] = row;
      g_dst[m] = (uint32_t)(sub * g_sub_bytes + row * RBg + ((cj ^ key) * 16) + (byte0 & 15));
      g_srcoff[m] = (uint32_t)byte0;
    }
    const int dm = d_chunks / 2;                            // dense items per thread: 64 * d_chunks / 128 (d_chunks = 2, 4, 8 or 16)
    int d_row[8];
    uint32_t d_dst[8], d_srcoff[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const int w = tid + 128 * m, c16 = w % d_chunks, row = (w / d_chunks) % kWiPX;
      const int sub = c16 / (CHd / 8), cj = c16 % (CHd / 8);
      const int key = RBd == 128 ? (row & 7) : ((row >> 2) & 1);
      d_row[m] = row;
      d_dst[m] = (uint32_t)(dense_off + sub * d_sub_bytes + row * RBd + ((cj ^ key) * 16));
      d_srcoff[m] = (uint32_t)(c16 * 16);
    }
    const uint32_t tap_pitch = (uint32_t)(gsubs * g_sub_bytes);
    // Row-aligned units (EW % 64 == 0, every layer of the net): a unit is 64 consecutive pixels of ONE image row, so (n, y, x0) come from two
    // divisions per unit and a thread's pixel is x0 + its fixed row index -- no per-pixel table, no loader-group barrier.  The small layers
    // (14-24 KB per unit, two CTAs per SM) were bound by exactly that per-unit bookkeeping: ~2.4 us per unit for the head.
    const bool aligned = fast_g && (p.EW % kWiPX) == 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* base = smem + (size_t)stage * stage_bytes;
      const uint32_t base_addr = smem_u32(base);
      if (aligned) {
        const unsigned q0 = (unsigned)u * (unsigned)kWiPX;
        const unsigned t2 = q0 / (unsigned)p.EW, nn = t2 / (unsigned)p.EH;
        const int x0 = (int)(q0 - t2 * (unsigned)p.EW), y0 = (int)(t2 - nn * (unsigned)p.EH);
        const int gy = p.g_step * y0;
        const size_t img = (size_t)nn * p.GH;
        for (int t = 0; t < p.ntaps; ++t) {
          const int dy = p.tap_dy[t], dx = p.tap_dx[t];
          const bool oky = (unsigned)(gy + dy) < (unsigned)p.GH;
          const uint8_t* rowp = G + ((img + (oky ? gy + dy : 0)) * p.GW) * (size_t)p.g_ld * 2;
          const uint32_t tbase = base_addr + (uint32_t)t * tap_pitch;
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            if (m < gm) {
              const int gx = gsx * (x0 + g_row[m]) + dx;
              const bool ok = oky && (unsigned)gx < (unsigned)p.GW;
              const uint8_t* src = ok ? rowp + (size_t)gx * p.g_ld * 2 + g_srcoff[m] : G;
              if (g_chunk == 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tbase + g_dst[m]), "l"(src), "r"(ok ? 16u : 0u) : "memory");
              else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(tbase + g_dst[m]), "l"(src), "r"(ok ? 8u : 0u) : "memory");
            }
          }
        }
        const uint8_t* drow = D + ((((size_t)nn * p.EH + y0) * p.EW + x0) * p.d_ld + p.d_c0) * 2;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          if (m < dm)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base_addr + d_dst[m]), "l"(drow + (size_t)d_row[m] * p.d_ld * 2 + d_srcoff[m]), "r"(16u) : "memory");
        }
      } else {
      // per-unit pixel table (64 rows); the loader group syncs on named barrier 1
      if (tid < kWiPX) {
        const unsigned q = (unsigned)u * (unsigned)kWiPX + (unsigned)tid;   // 32-bit index arithmetic (host-checked range)
        if (q < (unsigned)total) {
          const unsigned t2 = q / (unsigned)p.EW, nn = t2 / (unsigned)p.EH;
          s_row[stage][tid][2] = (int)(q - t2 * (unsigned)p.EW);
          s_row[stage][tid][1] = (int)(t2 - nn * (unsigned)p.EH);
          s_row[stage][tid][0] = (int)nn;
        } else {
          s_row[stage][tid][0] = -1;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (fast_g) {
        int gy0[8], gx0[8];
        const uint8_t* gsrc[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          if (m < gm) {
            const int n = s_row[stage][g_row[m]][0];
            gy0[m] = n >= 0 ? p.g_step * s_row[stage][g_row[m]][1] : -(1 << 28);
            gx0[m] = gsx * s_row[stage][g_row[m]][2];
            gsrc[m] = G + ((((size_t)(n >= 0 ? n : 0) * p.GH + (n >= 0 ? gy0[m] : 0)) * p.GW + gx0[m]) * p.g_ld) * 2 + g_srcoff[m];
          }
        }
        for (int t = 0; t < p.ntaps; ++t) {
          const int dy = p.tap_dy[t], dx = p.tap_dx[t];
          const long long toff = ((long long)dy * p.GW + dx) * p.g_ld * 2;
          const uint32_t tbase = base_addr + (uint32_t)t * tap_pitch;
#pragma unroll
          for (int m = 0; m < 8; ++m) {
            if (m < gm) {
              const bool ok = (unsigned)(gy0[m] + dy) < (unsigned)p.GH && (unsigned)(gx0[m] + dx) < (unsigned)p.GW;
              const uint8_t* src = ok ? gsrc[m] + toff : G;
              if (g_chunk == 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tbase + g_dst[m]), "l"(src), "r"(ok ? 16u : 0u) : "memory");
              else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(tbase + g_dst[m]), "l"(src), "r"(ok ? 8u : 0u) : "memory");
            }
          }
        }
      } else {
        // gathered operand, general form: items (tap, row, chunk)
        const int g_items = p.ntaps * kWiPX * g_chunks;
        for (int it = tid; it < g_items; it += 128) {
          const int ch = it % g_chunks, row = (it / g_chunks) % kWiPX, t = it / (g_chunks * kWiPX);
          const int n = s_row[stage][row][0];
          const int gy = p.g_step * s_row[stage][row][1] + p.tap_dy[t], gx = gsx * s_row[stage][row][2] + p.tap_dx[t];
          const bool ok = n >= 0 && gy >= 0 && gy < p.GH && gx >= 0 && gx < p.GW;
          const uint8_t* src = ok ? G + ((((size_t)n * p.GH + gy) * p.GW + gx) * p.g_ld) * 2 + ch * g_chunk : G;
          // destination: 16-byte chunk index c16 of the pixel row -> sub-tile c16 / (CHg/8), chunk within the sub-tile row
          const int byte0 = ch * g_chunk, c16 = byte0 / 16, sub = c16 / (CHg / 8), cj = c16 % (CHg / 8);
          const int key = RBg == 128 ? (row & 7) : ((row >> 2) & 1);
          uint8_t* dst = base + (t * gsubs + sub) * g_sub_bytes + row * RBg + ((cj ^ key) * 16) + (byte0 & 15);
          wi_cp_async(dst, src, g_chunk, ok ? (uint32_t)g_chunk : 0u);
        }
      }
      // dense operand: items (row, chunk), the same (row, chunk) pairs for a thread in every unit
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        if (m < dm) {
          const int n = s_row[stage][d_row[m]][0];
          const bool ok = n >= 0;
          const uint8_t* src = ok ? D + ((((size_t)n * p.EH + s_row[stage][d_row[m]][1]) * p.EW + s_row[stage][d_row[m]][2]) * p.d_ld + p.d_c0) * 2 + d_srcoff[m] : D;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base_addr + d_dst[m]), "l"(src), "r"(ok ? 16u : 0u) : "memory");
        }
      }
      }   // !aligned
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (lag == 2) {
        if (pend[0] >= 0) {
          asm volatile("cp.async.wait_group 2;" ::: "memory");
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[pend[0]]);
        }
        pend[0] = pend[1];
        pend[1] = stage;
      } else {
        if (pend[1] >= 0) {
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[pend[1]]);
        }
        pend[1] = stage;
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async_smem();
    if (pend[0] >= 0) mbar_arrive(&full_bar[pend[0]]);
    if (pend[1] >= 0) mbar_arrive(&full_bar[pend[1]]);
  } else if (warp == 4) {
    // ================= MMA issuer =================
    {   // whole warp: uniform control flow, one elected lane issues (see tc.cuh elect_one)
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t lay_g = RBg == 128 ? LAYOUT_SW128 : LAYOUT_SW32, lay_d = RBd == 128 ? LAYOUT_SW128 : LAYOUT_SW32;
      const uint32_t hi_g = ((8u * RBg) >> 4) | (1u << 14) | (lay_g << 29);   // SBO = 8 pixel rows
      const uint32_t hi_d = ((8u * RBd) >> 4) | (1u << 14) | (lay_d << 29);
      const uint32_t idesc = make_idesc_bf16(128, p.NP) | (1u << 15) | (1u << 16);   // both operands MN-major
      const int atoms_per_m = 128 / CHg;
      bool first = true;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        fence_proxy_async_smem();
        tcgen05_fence_after();
        const uint32_t g_addr = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t d_addr = g_addr + dense_off;
        const uint32_t d_lo = ((d_addr >> 4) & 0x3FFF) | ((((uint32_t)d_sub_bytes >> 4) & 0x3FFF) << 16);   // LBO = dense sub-tile pitch
        for (int gidx = 0; gidx < groups; ++gidx) {
          const uint32_t ga = g_addr + gidx * atoms_per_m * g_sub_bytes;
          const uint32_t g_lo = ((ga >> 4) & 0x3FFF) | ((((uint32_t)g_sub_bytes >> 4) & 0x3FFF) << 16);     // LBO = (tap, sub) atom pitch
          for (int k = 0; k < kWiPX / 16; ++k)
            if (elect_one()) umma_bf16(tmem_base + gidx * p.NP, ((uint64_t)hi_g << 32) | (uint64_t)(g_lo + k * RBg), ((uint64_t)hi_d << 32) | (uint64_t)(d_lo + k * RBd),
                      idesc, (first && k == 0) ? 0u : 1u);
        }
        first = false;
        if (elect_one()) umma_commit(&empty_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&done_bar);
    }
  } else {
    // ================= epilogue (once): dump this CTA's accumulators =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    float* wsb = ws + (size_t)blockIdx.x * groups * 128 * p.NP;
    if ((int)blockIdx.x < units) {
      mbar_wait(&done_bar, 0);
      tcgen05_fence_after();
      for (int gidx = 0; gidx < groups; ++gidx) {
        for (int c0 = 0; c0 < p.NP; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + gidx * p.NP + c0 + ((uint32_t)(quad * 32) << 16), r);
          tmem_ld_wait();
          // slab layout [group][col/4][row][4]: the 32 lanes (rows) of a store instruction write 512 contiguous bytes
          float4* dst = reinterpret_cast<float4*>(wsb) + ((size_t)gidx * (p.NP / 4) + c0 / 4) * 128 + row;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[(size_t)q * 128] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dw[((n0 + col) * Cg + cg) * KHW + tap_off[tap]] += sum over CTAs; (group, row) -> atom = group*(128/CHg) + row/CHg -> (tap, sub)
__global__ void wgrad_im2col_reduce_kernel(const float* __restrict__ ws, int ctas, int groups, erf_wgrad_im2col_desc p, float* __restrict__ dw) {
  pdl_enter();
  const int CHg = (p.CgP % 64) == 0 ? 64 : 16, gsubs = p.CgP / CHg;
  const int per_cta = groups * 128 * p.NP;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_cta) return;
  const int e4 = i & 3, row = (i >> 2) & 127, c4 = (i >> 9) % (p.NP / 4), gidx = (i >> 9) / (p.NP / 4);   // slab layout [group][col/4][row][4]
  const int col = c4 * 4 + e4;
  const int atom = gidx * (128 / CHg) + row / CHg;
  const int tap = atom / gsubs, sub = atom % gsubs;
  const int cg = sub * CHg + row % CHg, cd = p.d_c0 - p.d_c_base + col;
  if (tap >= p.ntaps || cg >= p.Cg || cd >= p.Cd) return;
  size_t out_idx;
  if (p.fold > 0) {   // column-folded gathered operand: cg = b * fold + c
    const int b = cg / p.fold, c = cg - b * p.fold;
    if (c >= p.fold_c) return;
    out_idx = ((size_t)cd * p.fold_c + c) * p.KHW + p.tap_off[tap] + b;
  } else {
    out_idx = ((size_t)cd * p.Cg + cg) * p.KHW + p.tap_off[tap];
  }
  const int per = (ctas + gridDim.y - 1) / gridDim.y;  // CTA range split over blockIdx.y, one atomic per split
  const int b0 = blockIdx.y * per, b1 = min(b0 + per, ctas);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = b0;
  for (; b + 4 <= b1; b += 4) {
    s0 += ws[(size_t)b * per_cta + i];
    s1 += ws[(size_t)(b + 1) * per_cta + i];
    s2 += ws[(size_t)(b + 2) * per_cta + i];
    s3 += ws[(size_t)(b + 3) * per_cta + i];
  }
  for (; b < b1; ++b) s0 += ws[(size_t)b * per_cta + i];
  if (b1 > b0) atomicAdd(&dw[out_idx], (s0 + s1) + (s2 + s3));
}

typedef CUresult (*WiEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                               CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// NHWC bf16 tensor [N][H][W][ld] as (C = ld, W, H, N); box = {ch, box_w, 1, 1} traversed with steps (1, step_w, step_h, 1): ch channels of
// box_w / step_w pixels of one row.  The row step only matters through the start coordinate (one row per box).
static int wi_make_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int ld, int ch, int box_w, int step_w, int step_h) {
  static WiEncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (WiEncodeFn)q;
  }
  if (!fn || box_w > 256) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)ch, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)step_w, 1, 1};
  (void)step_h;
  return (int)fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace erfk

using namespace erfk;

extern "C" size_t erf_conv_wgrad_im2col_tc_workspace(void) { return (size_t)148 * 512 * 128 * sizeof(float); }   // = 296 CTAs x 256 columns

extern "C" int erf_conv_wgrad_im2col_tc(const erf_wgrad_im2col_desc* d, float* dw, void* workspace, size_t ws_bytes, erf_stream_t stream) {
  ERF_CHECK_ARG(d && d->g && d->d && dw && workspace, "erf_conv_wgrad_im2col_tc: null pointer");
  ERF_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= 9 && d->EN > 0 && d->EH > 0 && d->EW > 0, "erf_conv_wgrad_im2col_tc: bad tap count or grid");
  ERF_CHECK_ARG(d->CgP >= 16 && d->CgP <= 128 && (d->CgP % 16) == 0 && d->NP >= 16 && d->NP <= 128 && (d->NP % 16) == 0,
                "erf_conv_wgrad_im2col_tc: padded channel counts must be multiples of 16 in [16,128] (CgP=%d NP=%d)", d->CgP, d->NP);
  ERF_CHECK_ARG(d->g_copy_bytes == 8 || (d->g_copy_bytes >= 16 && (d->g_copy_bytes % 16) == 0 && d->g_copy_bytes <= d->CgP * 2),
                "erf_conv_wgrad_im2col_tc: bad g_copy_bytes %d", d->g_copy_bytes);
  ERF_CHECK_ARG((d->d_ld % 8) == 0 && (d->d_c0 % 8) == 0 && d->d_c0 + d->NP <= d->d_ld,
                "erf_conv_wgrad_im2col_tc: dense slice [%d, %d) must be 16-byte aligned and inside the pixel row of %d channels", d->d_c0,
                d->d_c0 + d->NP, d->d_ld);
  const int CHg = (d->CgP % 64) == 0 ? 64 : 16, gsubs = d->CgP / CHg, RBg = CHg * 2;
  const int CHd = (d->NP % 64) == 0 ? 64 : 16, dsubs = d->NP / CHd, RBd = CHd * 2;
  const int atoms = d->ntaps * gsubs, atoms_per_m = 128 / CHg;
  const int groups = (atoms + atoms_per_m - 1) / atoms_per_m;
  if (groups * d->NP > 512) {
    set_error("erf_conv_wgrad_im2col_tc: %d accumulator groups x N=%d exceed TMEM; split N", groups, d->NP);
    return ERF_ERR_UNSUPPORTED;
  }
  const int g_bytes = ((atoms * kWiPX * RBg + 1023) / 1024) * 1024;
  const int d_bytes = ((dsubs * kWiPX * RBd + 1023) / 1024) * 1024;
  const int stage_bytes = g_bytes + d_bytes;
  int stages = (190 * 1024) / stage_bytes;
  if (stages > kWiMaxStages) stages = kWiMaxStages;
  if (stages < 2) {
    set_error("erf_conv_wgrad_im2col_tc: stage of %d bytes does not fit twice in shared memory", stage_bytes);
    return ERF_ERR_UNSUPPORTED;
  }
  erf_wgrad_im2col_desc p = *d;
  // the last group's unused atoms read past the gathered tiles: keep those reads inside the (zeroed) allocation
  p.tail_bytes = atoms_per_m * kWiPX * RBg;
  const long long total = (long long)d->EN * d->EH * d->EW;
  ERF_CHECK_ARG(total + kWiPX < 0x7fffffffLL, "erf_conv_wgrad_im2col_tc: %lld pixels exceed the 32-bit index range", total);
  int units = (int)((total + kWiPX - 1) / kWiPX);
  // two CTAs per SM when two rings of >= 2 stages and two accumulator sets fit (latency-bound per-unit chain)
  int per_sm = 1;
  if (groups * d->NP <= 256 && (104 * 1024 - p.tail_bytes) / stage_bytes >= 2 && units > 148) {
    per_sm = 2;
    stages = (104 * 1024 - p.tail_bytes) / stage_bytes;
    if (stages > kWiMaxStages) stages = kWiMaxStages;
  }
  const int grid = units < 148 * per_sm ? units : 148 * per_sm;
  ERF_CHECK_ARG(ws_bytes >= (size_t)grid * groups * 128 * d->NP * sizeof(float), "erf_conv_wgrad_im2col_tc: workspace too small");
  size_t smem = (size_t)stages * stage_bytes + p.tail_bytes + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  // TMA loaders: row-aligned units, 16-byte aligned pixel rows and channel offsets, whole 16-byte gathered rows
  static const bool tma_off = getenv("ERF_WGRAD_IM2COL_TMA") && atoi(getenv("ERF_WGRAD_IM2COL_TMA")) == 0;
  CUtensorMap mg, md;
  bool tma = !tma_off && d->EW >= 16 && d->g_copy_bytes >= 16 && (d->g_ld % 8) == 0 && (d->d_ld % 8) == 0 && ((uintptr_t)d->g % 16) == 0 &&
             ((uintptr_t)d->d % 16) == 0 && d->g_step >= 1 && d->g_step <= 4;
  if (tma) {
    const int gsx = d->g_step_x > 0 ? d->g_step_x : d->g_step;
    tma = wi_make_map(&mg, d->g, d->EN, d->GH, d->GW, d->g_ld, CHg, kWiPX * gsx, gsx, d->g_step) == 0 &&
          wi_make_map(&md, d->d, d->EN, d->EH, d->EW, d->d_ld, CHd, kWiPX, 1, 1) == 0;
  }
  if (!tma) { memset(&mg, 0, sizeof(mg)); memset(&md, 0, sizeof(md)); }
  cudaError_t e;
#define WI_ATTR(MINB_, TMA_)                                                                                                   \
  e = cudaFuncSetAttribute(conv_wgrad_im2col_tc_kernel<MINB_, TMA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e == cudaSuccess && MINB_ == 2) e = prefer_max_smem(conv_wgrad_im2col_tc_kernel<MINB_, TMA_>);
  if (per_sm == 2) { if (tma) { WI_ATTR(2, true) } else { WI_ATTR(2, false) } }
  else { if (tma) { WI_ATTR(1, true) } else { WI_ATTR(1, false) } }
#undef WI_ATTR
  if (e != cudaSuccess) {
    set_error("erf_conv_wgrad_im2col_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
#define WI_LAUNCH(MINB_, TMA_) \
  launch_pdl(conv_wgrad_im2col_tc_kernel<MINB_, TMA_>, dim3(grid), dim3(kWiThreads), smem, (cudaStream_t)(st), mg, md, p, stages, stage_bytes, g_bytes, groups, (float*)workspace)
  if (per_sm == 2) { if (tma) WI_LAUNCH(2, true); else WI_LAUNCH(2, false); }
  else { if (tma) WI_LAUNCH(1, true); else WI_LAUNCH(1, false); }
#undef WI_LAUNCH
  const int per_cta = groups * 128 * d->NP;
  const int rblocks = (per_cta + 255) / 256;
  const int rsplit = rblocks >= 148 ? 4 : (rblocks >= 32 ? 12 : 37);
  launch_pdl(wgrad_im2col_reduce_kernel, dim3(dim3(rblocks, rsplit)), dim3(256), 0, (cudaStream_t)(st), (const float*)workspace, grid, groups, p, dw);
  ERF_CHECK_LAUNCH("erf_conv_wgrad_im2col_tc");
  return ERF_OK;
}
