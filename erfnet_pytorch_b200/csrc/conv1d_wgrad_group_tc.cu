// Grouped tcgen05 weight gradients: the (up to) four 3-tap convolutions of one non_bottleneck_1d block (eval/erfnet.py:30-55) in ONE launch.
//
// At batch 6 a single weight-gradient launch is ~5 units of MMA work per CTA wrapped in ~6 us of fixed cost (launch, barrier / TMEM set-up,
// first TMA round trip, accumulator dump, a second launch that reduces the per-CTA partial sums).  Nothing but the optimiser waits for a
// weight gradient, so the block's backward defers all four until its data gradients are done and hands them to this kernel as a job list:
// every persistent CTA walks the jobs in order -- same pipeline as conv1d_wgrad_tc.cu (TMA-staged x / dy tiles, MN-major operands, taps
// as M atoms through the descriptor's leading-byte offset, TMEM accumulation over all units of the job) -- dumps each job's partial sums
// into that job's slab range, and ONE reduce launch (grid.z = job) adds the slabs into the four torch-layout gradients.
// 136 launches per training step become 34.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"

namespace erfk {
using namespace tc;

constexpr int kWgMaxJobs = 4;
constexpr int kWgGStages = 8;
constexpr int GPX = 64;   // pixels (K) per unit

struct WgradGroupMaps {
  CUtensorMap x[kWgMaxJobs], dy[kWgMaxJobs];
};

struct WgradGroupParams {
  float* ws;                  // [job][grid][ACCS][128][C] partial sums
  float* dw[kWgMaxJobs];
  int horizontal[kWgMaxJobs], dil[kWgMaxJobs];
  int x_sub_bytes[kWgMaxJobs], dy_off[kWgMaxJobs];
  int njobs, N, H, W, units, stages, stage_bytes;
};

template <int C>
__global__ void __launch_bounds__(192, 1) conv1d_wgrad_group_tc_kernel(const __grid_constant__ WgradGroupMaps maps, const WgradGroupParams p) {
  pdl_trigger();
  constexpr int CH = C < 64 ? C : 64;
  constexpr int RB = CH * 2;
  constexpr int KC = C / CH;
  constexpr uint32_t LAYOUT = RB == 128 ? LAYOUT_SW128 : LAYOUT_SW32;
  constexpr int ACCS = C == 128 ? 3 : (C == 64 ? 2 : 1);
  constexpr int TMEM_COLS = C == 128 ? 512 : (C == 64 ? 128 : 32);
  constexpr uint32_t IDESC = make_idesc_bf16(128, C) | (1u << 15) | (1u << 16);   // both operands MN-major
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[kWgGStages], empty_bar[kWgGStages], done_bar, acc_free_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    mbar_init(&acc_free_bar, 4);     // one arrival per epilogue warp
    fence_barrier_init();
    for (int j = 0; j < p.njobs; ++j) {
      tma_prefetch_desc(&maps.x[j]);
      tma_prefetch_desc(&maps.dy[j]);
    }
  }
  __syncthreads();
  if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int segs = (p.W + GPX - 1) / GPX;
  pdl_wait();
  const int my_units = (int)blockIdx.x < p.units ? (p.units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ================= TMA producer (whole warp in uniform control flow, one elected lane issues) =================
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < p.njobs; ++j) {
      const int horizontal = p.horizontal[j], dil = p.dil[j], xsb = p.x_sub_bytes[j], dyo = p.dy_off[j];
      const uint32_t tx = (horizontal ? (uint32_t)(KC * (GPX + 2 * dil) * RB) : (uint32_t)(3 * KC * GPX * RB)) + (uint32_t)(KC * GPX * RB);
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int xs = u % segs, y = (u / segs) % p.H, n = u / (segs * p.H);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* base = smem + (size_t)stage * p.stage_bytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          if (horizontal) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) tma_load_4d(base + kc * xsb, &maps.x[j], &full_bar[stage], kc * CH, xs * GPX - dil, y, n);
          } else {
#pragma unroll
            for (int t = 0; t < 3; ++t)
#pragma unroll
              for (int kc = 0; kc < KC; ++kc) tma_load_4d(base + (t * KC + kc) * xsb, &maps.x[j], &full_bar[stage], kc * CH, xs * GPX, y + (t - 1) * dil, n);
          }
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) tma_load_4d(base + dyo + kc * (GPX * RB), &maps.dy[j], &full_bar[stage], kc * CH, xs * GPX, y, n);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t hi_base = ((8u * RB) >> 4) | (1u << 14) | (LAYOUT << 29);
    for (int j = 0; j < p.njobs; ++j) {
      const int horizontal = p.horizontal[j], dil = p.dil[j], xsb = p.x_sub_bytes[j], dyo = p.dy_off[j];
      const uint32_t tap_pitch = horizontal ? (uint32_t)dil * RB : (uint32_t)KC * xsb;
      const uint32_t lbo_x = C == 128 ? (uint32_t)xsb : tap_pitch;
      const uint32_t lbo_dy = GPX * RB;
      if (j > 0 && my_units > 0) {   // the accumulators are free again once the epilogue has read job j - 1 out of TMEM
        mbar_wait(&acc_free_bar, (uint32_t)((j - 1) & 1));
        tcgen05_fence_after();
      }
      bool first = true;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t x_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t dy_addr = x_addr + dyo;
#pragma unroll
        for (int a = 0; a < ACCS; ++a) {
          const uint32_t xa = x_addr + a * tap_pitch;
          const uint32_t x_lo = ((xa >> 4) & 0x3FFF) | (((lbo_x >> 4) & 0x3FFF) << 16);
          const uint32_t dy_lo = ((dy_addr >> 4) & 0x3FFF) | (((lbo_dy >> 4) & 0x3FFF) << 16);
#pragma unroll
          for (int k = 0; k < GPX / 16; ++k) {
            const uint64_t da = ((uint64_t)hi_base << 32) | (uint64_t)(x_lo + k * RB);
            const uint64_t db = ((uint64_t)hi_base << 32) | (uint64_t)(dy_lo + k * RB);
            if (elect_one()) umma_bf16(tmem_base + a * C, da, db, IDESC, (first && k == 0) ? 0u : 1u);
          }
        }
        first = false;
        if (elect_one()) umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&done_bar);
      __syncwarp();
    }
  } else {
    // ================= epilogue: after each job, dump this CTA's accumulators into the job's slab =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    for (int j = 0; j < p.njobs; ++j) {
      float* wsb = p.ws + ((size_t)j * gridDim.x + blockIdx.x) * ACCS * 128 * C;
      if (my_units > 0) {
        mbar_wait(&done_bar, (uint32_t)(j & 1));
        tcgen05_fence_after();
#pragma unroll
        for (int a = 0; a < ACCS; ++a) {
          constexpr int EW = C >= 32 ? 32 : 16;
#pragma unroll
          for (int c0 = 0; c0 < C; c0 += EW) {
            uint32_t r[EW];
            if (EW == 32) tmem_ld32(tmem_base + a * C + c0 + ((uint32_t)(quad * 32) << 16), r);
            else tmem_ld16(tmem_base + a * C + c0 + ((uint32_t)(quad * 32) << 16), r);
            tmem_ld_wait();
            float4* dst = reinterpret_cast<float4*>(wsb) + ((size_t)a * (C / 4) + c0 / 4) * 128 + row;   // slab layout [acc][col/4][row][4]
#pragma unroll
            for (int q = 0; q < EW / 4; ++q)
              dst[(size_t)q * 128] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_free_bar);
      } else {
        for (int a = 0; a < ACCS; ++a)
          for (int c4 = 0; c4 < C / 4; ++c4) reinterpret_cast<float4*>(wsb)[((size_t)a * (C / 4) + c4) * 128 + row] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dw_j[co][ci][t] += sum over CTAs of job j's partial accumulators (blockIdx.z = job; see wgrad_reduce_kernel in conv1d_wgrad_tc.cu)
struct WgradGroupDw { float* dw[kWgMaxJobs]; };

template <int C>
__global__ void __launch_bounds__(256) wgrad_group_reduce_kernel(const float* __restrict__ ws, int grid_ctas, WgradGroupDw out) {
  pdl_enter();
  constexpr int ACCS = C == 128 ? 3 : (C == 64 ? 2 : 1);
  constexpr size_t SLAB = (size_t)ACCS * 128 * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int)SLAB) return;
  const int e = i & 3, row = (i >> 2) & 127, c4 = (i >> 9) % (C / 4), acc = (i >> 9) / (C / 4);
  const int co = c4 * 4 + e;
  int t, ci;
  if (C == 128) { t = acc; ci = row; }
  else if (C == 64) { t = acc + (row >> 6); ci = row & 63; if (acc == 1 && row < 64) return; }
  else { t = row >> 4; ci = row & 15; if (t > 2) return; }
  const int per = (grid_ctas + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(b0 + per, grid_ctas);
  const float* src = ws + (size_t)blockIdx.z * grid_ctas * SLAB + i;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int b = b0;
  for (; b + 8 <= b1; b += 8) {
    const float v0 = src[(size_t)b * SLAB], v1 = src[(size_t)(b + 1) * SLAB], v2 = src[(size_t)(b + 2) * SLAB], v3 = src[(size_t)(b + 3) * SLAB];
    const float v4 = src[(size_t)(b + 4) * SLAB], v5 = src[(size_t)(b + 5) * SLAB], v6 = src[(size_t)(b + 6) * SLAB], v7 = src[(size_t)(b + 7) * SLAB];
    s0 += v0 + v4;
    s1 += v1 + v5;
    s2 += v2 + v6;
    s3 += v3 + v7;
  }
  for (; b < b1; ++b) s0 += src[(size_t)b * SLAB];
  if (b1 > b0) atomicAdd(&out.dw[blockIdx.z][((size_t)co * C + ci) * 3 + t], (s0 + s1) + (s2 + s3));
}

typedef CUresult (*GEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g_make_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int box_w) {
  static GEncodeFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (GEncodeFn)q;
  }
  if (!fn) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)(C < 64 ? C : 64), (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return (int)fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 C < 64 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace erfk

using namespace erfk;

extern "C" size_t erf_conv1d_wgrad_group_tc_workspace(int C) { return (size_t)kWgMaxJobs * erf_conv1d_wgrad_tc_workspace(C); }

extern "C" int erf_conv1d_wgrad_group_tc(const erf_wgrad_job* jobs, int njobs, void* workspace, size_t ws_bytes, int N, int H, int W, int C,
                                         erf_stream_t stream) {
  ERF_CHECK_ARG(jobs && workspace && njobs >= 1 && njobs <= kWgMaxJobs, "erf_conv1d_wgrad_group_tc: 1..%d jobs", kWgMaxJobs);
  ERF_CHECK_ARG(N > 0 && H > 0, "erf_conv1d_wgrad_group_tc: empty tensor");
  if (!((C == 16 || C == 64 || C == 128) && W >= 16)) {
    set_error("erf_conv1d_wgrad_group_tc: unsupported shape C=%d W=%d", C, W);
    return ERF_ERR_UNSUPPORTED;
  }
  ERF_CHECK_ARG(ws_bytes >= erf_conv1d_wgrad_group_tc_workspace(C), "erf_conv1d_wgrad_group_tc: workspace too small (%zu < %zu)", ws_bytes,
                erf_conv1d_wgrad_group_tc_workspace(C));
  const int CH = C < 64 ? C : 64, RB = CH * 2, KC = C / CH;
  WgradGroupMaps maps;
  WgradGroupParams p;
  WgradGroupDw out;
  memset(&maps, 0, sizeof(maps));
  memset(&p, 0, sizeof(p));
  p.ws = (float*)workspace;
  p.njobs = njobs;
  p.N = N; p.H = H; p.W = W;
  p.units = N * H * ((W + GPX - 1) / GPX);
  int stage_bytes = 0;
  for (int j = 0; j < njobs; ++j) {
    const erf_wgrad_job& jb = jobs[j];
    ERF_CHECK_ARG(jb.x && jb.dy && jb.dw && jb.dil >= 1 && jb.dil <= 32, "erf_conv1d_wgrad_group_tc: bad job %d", j);
    const int box_w = jb.horizontal ? GPX + 2 * jb.dil : GPX;
    if (g_make_map(&maps.x[j], jb.x, N, H, W, C, box_w) != 0 || g_make_map(&maps.dy[j], jb.dy, N, H, W, C, GPX) != 0) {
      set_error("erf_conv1d_wgrad_group_tc: cuTensorMapEncodeTiled failed");
      return ERF_ERR_CUDA;
    }
    int x_bytes;
    if (jb.horizontal) {
      p.x_sub_bytes[j] = ((((box_w + 7) / 8) * 8 * RB + 1023) / 1024) * 1024;
      x_bytes = KC * p.x_sub_bytes[j];
    } else {
      p.x_sub_bytes[j] = GPX * RB;
      x_bytes = 3 * KC * p.x_sub_bytes[j];
    }
    x_bytes = ((x_bytes + 1023) / 1024) * 1024;
    p.dy_off[j] = x_bytes;
    const int sb = x_bytes + ((KC * GPX * RB + 1023) / 1024) * 1024;
    if (sb > stage_bytes) stage_bytes = sb;
    p.horizontal[j] = jb.horizontal;
    p.dil[j] = jb.dil;
    p.dw[j] = jb.dw;
    out.dw[j] = jb.dw;
  }
  for (int j = njobs; j < kWgMaxJobs; ++j) out.dw[j] = nullptr;
  p.stage_bytes = stage_bytes;
  p.stages = (200 * 1024) / stage_bytes;
  if (p.stages > kWgGStages) p.stages = kWgGStages;
  if (p.stages < 2) {
    set_error("erf_conv1d_wgrad_group_tc: a stage of %d bytes does not fit twice in shared memory", stage_bytes);
    return ERF_ERR_UNSUPPORTED;
  }
  const bool two_per_sm = C == 16;
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + (C == 16 ? 16 * 1024 : 0);
  const int ctas = two_per_sm ? 296 : 148;
  const int grid = p.units < ctas ? p.units : ctas;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
#define WG_GROUP(C_, RX, RY)                                                                                                               \
  e = cudaFuncSetAttribute(conv1d_wgrad_group_tc_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
  if (e == cudaSuccess && two_per_sm) e = prefer_max_smem(conv1d_wgrad_group_tc_kernel<C_>);                                                \
  if (e == cudaSuccess) {                                                                                                                  \
    launch_pdl(conv1d_wgrad_group_tc_kernel<C_>, dim3(grid), dim3(192), smem, s, maps, p);                                                 \
    launch_pdl(wgrad_group_reduce_kernel<C_>, dim3(RX, RY, njobs), dim3(256), 0, s, (const float*)p.ws, grid, out);                         \
  }
  if (C == 16) { WG_GROUP(16, (1 * 128 * 16 + 255) / 256, 37) }
  else if (C == 64) { WG_GROUP(64, (2 * 128 * 64 + 255) / 256, 12) }
  else { WG_GROUP(128, (3 * 128 * 128 + 255) / 256, 6) }
#undef WG_GROUP
  if (e != cudaSuccess) {
    set_error("erf_conv1d_wgrad_group_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
  ERF_CHECK_LAUNCH("erf_conv1d_wgrad_group_tc");
  return ERF_OK;
}
