// tcgen05 implicit-GEMM kernel for the factorised 3x1 / 1x3 (dilated) convolutions of non_bottleneck_1d
// (eval/erfnet.py:30-38,47-55): NHWC bf16 in/out, fp32 accumulation in TMEM, C = 64 or 128 channels.
//
//   out[p, co] = epilogue( sum_{t=0..2} sum_ci in[p + (t-1)*dil along H or W, ci] * Wt[t][co][ci] )
//   epilogue   = + bias[co] -> (+ residual[p, co]) -> relu? -> (* (mask[p, co] > 0))      (same order as erf_conv_fwd)
//
// One work unit = 128 consecutive pixels of one image row = one UMMA M-block (M=128, N=C, K=C per tap).
//   * the activation tile is staged by TMA (128B swizzle, out-of-bounds rows/columns zero-filled = conv padding):
//       vertical conv  : three 128-pixel boxes (rows y-d, y, y+d), one per tap
//       horizontal conv: one (128+2d)-pixel box; tap t is the SAME tile viewed through a descriptor whose start
//                        address is shifted by t*d pixel rows (address-based swizzle makes any row shift legal;
//                        pinned by tests/test_gpu_umma.py)
//   * the three tap weight matrices live in shared memory for the whole (persistent) CTA, pre-swizzled by the host
//   * warp 0 = TMA producer, warp 1 = MMA issuer (single thread), warps 2-5 = epilogue (TMEM -> registers -> global),
//     double-buffered TMEM accumulators so the epilogue of unit i overlaps the MMAs of unit i+1.
#include <cuda.h>

#include "common.cuh"
#include "tc.cuh"

namespace erfk {
using namespace tc;

struct Conv1dTcParams {
  const __nv_bfloat16* w_img;   // [3][C/64][C][64] pre-swizzled smem image
  const float* bias;            // [C] or null
  __nv_bfloat16* out;           // NHWC, ld = C
  const __nv_bfloat16* res;     // NHWC, ld = C, or null
  const __nv_bfloat16* mask;    // NHWC, ld = C, or null
  int N, H, W;
  int horizontal;               // 0: taps along H (3x1), 1: taps along W (1x3)
  int dil;
  int relu;
  int units;                    // N * H * (W / 128)
  int stages;                   // A-tile ring depth
  int stage_bytes;              // bytes per ring slot (multiple of 1024)
  int sub_bytes;                // bytes per (tap, k-chunk) sub-tile (vertical) / per k-chunk sub-tile (horizontal)
};

constexpr int kMaxStages = 8;

template <int C>
__global__ void __launch_bounds__(192, 1) conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Conv1dTcParams p) {
  constexpr int CH = C < 64 ? C : 64;           // channels per K chunk (one swizzle row)
  constexpr int RB = CH * 2;                    // bytes per pixel row of a chunk: 128 (128B swizzle) or 32 (32B swizzle, C = 16)
  constexpr int KC = C / CH;
  constexpr int KSTEPS = CH / 16;               // K = 16 MMAs per chunk
  constexpr uint32_t LAYOUT = RB == 128 ? LAYOUT_SW128 : LAYOUT_SW32;
  constexpr int W_BYTES = 3 * KC * C * RB;      // all tap weights
  constexpr int EW = C >= 32 ? 32 : 16;         // epilogue column chunk
  constexpr uint32_t TMEM_COLS = 2 * C < 32 ? 32 : 2 * C;
  constexpr uint32_t IDESC = make_idesc_bf16(128, C);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w = smem;
  uint8_t* s_a = smem + ((W_BYTES + 1023) / 1024) * 1024;
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[C];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- one-time setup -----------------------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < W_BYTES / 16; i += blockDim.x) dst[i] = src[i];
    for (int i = tid; i < C; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    fence_proxy_async_smem();
  }
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  __syncthreads();
  if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int segs = p.W / 128;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = p.horizontal ? (uint32_t)(KC * (128 + 2 * p.dil) * RB) : (uint32_t)(3 * KC * 128 * RB);
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int xs = u % segs, y = (u / segs) % p.H, n = u / (segs * p.H);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        uint8_t* base = s_a + (size_t)stage * p.stage_bytes;
        if (p.horizontal) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) tma_load_4d(base + kc * p.sub_bytes, &tmap, &full_bar[stage], kc * CH, xs * 128 - p.dil, y, n);
        } else {
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int kc = 0; kc < KC; ++kc)
              tma_load_4d(base + (t * KC + kc) * p.sub_bytes, &tmap, &full_bar[stage], kc * CH, xs * 128, y + (t - 1) * p.dil, n);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      const uint32_t w_addr = smem_u32(s_w);
      // descriptor high word: SBO = 1024 B, version 1, 128B swizzle; low word: start address | LBO (=16 B, ignored)
      const uint32_t desc_hi = ((8u * RB) >> 4) | (1u << 14) | (LAYOUT << 29);  // SBO = 8 rows
      const uint32_t tap_rows_bytes = p.horizontal ? (uint32_t)p.dil * RB : (uint32_t)KC * p.sub_bytes;  // A start shift per tap
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(s_a + (size_t)stage * p.stage_bytes);
        const uint32_t d_tmem = tmem_base + acc * C;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            const uint32_t a_lo = (((a_addr + t * tap_rows_bytes + kc * p.sub_bytes) >> 4) & 0x3FFF) | (1u << 16);
            const uint32_t b_lo = (((w_addr + (t * KC + kc) * (C * RB)) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + k * 2);
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + k * 2);
              umma_bf16(d_tmem, da, db, IDESC, (t | kc | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty_bar[stage]);   // A slot reusable once these MMAs have read it
        umma_commit(&tfull_bar[acc]);     // accumulator ready for the epilogue
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================= epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // pixel within the unit
    uint32_t acc = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
      const int xs = u % segs, y = (u / segs) % p.H, n = u / (segs * p.H);
      const size_t pix = ((size_t)n * p.H + y) * p.W + xs * 128 + row;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * C + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += EW) {
        uint32_t r[EW];
        if (EW == 32) tmem_ld32(taddr + c0, r); else tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        float v[EW];
#pragma unroll
        for (int j = 0; j < EW; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c0 + j];
        if (p.res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * C + c0);
#pragma unroll
          for (int q = 0; q < EW / 8; ++q) {
            uint4 t = rp[q];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[q * 8 + 2 * e] += __low2float(h[e]);
              v[q * 8 + 2 * e + 1] += __high2float(h[e]);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < EW; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.mask) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask + pix * C + c0);
#pragma unroll
          for (int q = 0; q < EW / 8; ++q) {
            uint4 t = mp[q];
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (!(__low2float(h[e]) > 0.f)) v[q * 8 + 2 * e] = 0.f;
              if (!(__high2float(h[e]) > 0.f)) v[q * 8 + 2 * e + 1] = 0.f;
            }
          }
        }
        uint4* op = reinterpret_cast<uint4*>(p.out + pix * C + c0);
#pragma unroll
        for (int q = 0; q < EW / 8; ++q) {
          uint4 t;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
          op[q] = t;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  // ---- teardown ----------------------------------------------------------------------------------------------------
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;  // benign race: every thread resolves the same pointer
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace erfk

using namespace erfk;

extern "C" int erf_conv1d_tc_supported(int C, int W, int dil) {
  return (C == 16 || C == 64 || C == 128) && W >= 128 && (W % 128) == 0 && dil >= 1 && dil <= 32;
}

extern "C" int erf_conv1d_tc(const void* in, const void* w_img, const float* bias, void* out, const void* res, const void* mask, int N, int H, int W,
                             int C, int horizontal, int dil, int relu, erf_stream_t stream) {
  ERF_CHECK_ARG(in && w_img && out, "erf_conv1d_tc: null pointer");
  ERF_CHECK_ARG(N > 0 && H > 0, "erf_conv1d_tc: empty tensor");
  if (!erf_conv1d_tc_supported(C, W, dil)) {
    set_error("erf_conv1d_tc: unsupported shape C=%d W=%d dil=%d (need C in {16,64,128}, W %% 128 == 0, dil <= 32)", C, W, dil);
    return ERF_ERR_UNSUPPORTED;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("erf_conv1d_tc: cuTensorMapEncodeTiled not available from the driver");
    return ERF_ERR_CUDA;
  }
  const int CH = C < 64 ? C : 64, RB = CH * 2, KC = C / CH;
  const int box_w = horizontal ? 128 + 2 * dil : 128;
  CUtensorMap tmap;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)CH, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    RB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("erf_conv1d_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return ERF_ERR_CUDA;
  }
  Conv1dTcParams p;
  p.w_img = (const __nv_bfloat16*)w_img;
  p.bias = bias;
  p.out = (__nv_bfloat16*)out;
  p.res = (const __nv_bfloat16*)res;
  p.mask = (const __nv_bfloat16*)mask;
  p.N = N; p.H = H; p.W = W;
  p.horizontal = horizontal; p.dil = dil; p.relu = relu;
  p.units = N * H * (W / 128);
  if (horizontal) {
    p.sub_bytes = ((((box_w + 7) / 8) * 8 * RB + 1023) / 1024) * 1024;
    p.stage_bytes = KC * p.sub_bytes;
  } else {
    p.sub_bytes = 128 * RB;
    p.stage_bytes = 3 * KC * p.sub_bytes;
  }
  const int w_bytes = ((3 * KC * C * RB + 1023) / 1024) * 1024;
  const int budget = 220 * 1024 - w_bytes;
  p.stages = budget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  if (p.stages < 1) {
    set_error("erf_conv1d_tc: tile does not fit in shared memory");
    return ERF_ERR_UNSUPPORTED;
  }
  size_t smem = (size_t)w_bytes + (size_t)p.stages * p.stage_bytes + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = p.units < sms ? p.units : sms;
  cudaError_t e;
  if (C == 16) {
    e = cudaFuncSetAttribute(conv1d_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv1d_tc_kernel<16><<<grid, 192, smem, (cudaStream_t)stream>>>(tmap, p);
  } else if (C == 64) {
    e = cudaFuncSetAttribute(conv1d_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv1d_tc_kernel<64><<<grid, 192, smem, (cudaStream_t)stream>>>(tmap, p);
  } else {
    e = cudaFuncSetAttribute(conv1d_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) conv1d_tc_kernel<128><<<grid, 192, smem, (cudaStream_t)stream>>>(tmap, p);
  }
  if (e != cudaSuccess) {
    set_error("erf_conv1d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
  ERF_CHECK_LAUNCH("erf_conv1d_tc");
  return ERF_OK;
}
