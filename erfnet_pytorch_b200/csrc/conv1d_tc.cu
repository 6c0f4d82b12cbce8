// tcgen05 implicit-GEMM kernel for the factorised 3x1 / 1x3 (dilated) convolutions of non_bottleneck_1d
// (eval/erfnet.py:30-38,47-55) and their data gradients: NHWC bf16 in/out, fp32 accumulation in TMEM, C = 16, 64, 128.
//
//   out[p, co] = epilogue( sum_{t=0..2} sum_ci in[p + (t-1)*dil along H or W, ci] * Wt[t][co][ci] )
//   epilogue   = + bias[co] -> (+ aux[p, co] | aux_mode 1) -> relu? -> (* (aux[p, co] > 0) | aux_mode 2)
//
// One work unit = 128 consecutive pixels of one image row = one UMMA M-block (M=128, N=C, K=C per tap).
//   * activations are staged by TMA with hardware swizzle; out-of-image rows/columns are zero-filled = conv padding:
//       vertical conv  : three 128-pixel boxes (rows y-d, y, y+d) -- one ring slot and one MMA group PER TAP, so the
//                        load of tap t+1 overlaps the MMAs of tap t even when only 2-3 slots fit (C = 128)
//       horizontal conv: one (128+2d)-pixel box; tap t is the SAME tile viewed through a descriptor whose start address
//                        is shifted by t*d pixel rows (address-based swizzle makes any row shift legal; pinned by
//                        tests/test_gpu_umma.py)
//   * the three tap weight matrices (host pre-swizzled image) arrive once per persistent CTA with cp.async.bulk
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-5 and 6-9 = TWO epilogue groups that take
//     alternate units (one TMEM accumulator each): TMEM -> registers -> swizzled shared-memory staging sub-tile (one
//     64-channel k-chunk, double-buffered per group) -> TMA store (coalesced 128-byte lines instead of 32 scattered
//     16-byte stores per warp instruction); the residual / ReLU-mask operand is TMA-PREFETCHED into the staging sub-tile
//     one step ahead and combined in place
//   * the A ring works at (tap, k-chunk) granularity for vertical convs and k-chunk granularity for horizontal ones, so
//     at C = 128 (96 KB of weights) there are still 3-4 loads in flight.
#include <cuda.h>
#include <type_traits>

#include "common.cuh"
#include "tc.cuh"

namespace erfk {
using namespace tc;

struct Conv1dTcParams {
  const __nv_bfloat16* w_img;   // [3][C/CH][C][CH] pre-swizzled smem image
  const float* bias;            // [C] or null
  int N, H, W;
  int horizontal;               // 0: taps along H (3x1), 1: taps along W (1x3)
  int dil;
  int relu;
  int w_static;                 // the weight image is not written by any kernel still in flight (inference): its copy starts before pdl_wait()
  int aux_mode;                 // 0 none, 1 add (residual / gradient accumulation), 2 relu-backward mask
  int units;                    // N * H * ceil(W / 128): the last segment of a row may be ragged (TMA zero-fills its loads and clips its stores)
  int slots;                    // A ring depth
  int slot_bytes;               // bytes per ring slot (multiple of 1024)
  int sub_bytes;                // pitch of one k-chunk sub-tile inside a slot
  double* stats;                // [2C] += per-channel sum / sum of squares of the stored output (BN batch statistics), or null
  float* colsum;                // [C]  += per-channel sum of the stored output (bias gradient of the producing conv), or null
  int dbg_launch;               // timeline instrumentation: row block of this launch in the debug table
  int nb;                       // staging sub-tiles per epilogue group: 2, or 3 with a residual / mask operand (C <= 64): its tile is then prefetched TWO
                                // steps ahead (one step = ~0.6 us is less than the L2 / DRAM round trip: the in-situ trace showed those launches 2-3 us longer)
};

// Timeline instrumentation (tools/timeline_conv1d.py, tools/trace_step_conv1d.py): when a buffer is registered every CTA records SM-clock stamps
// of its phases into row [launch][blockIdx.x] of a [launches][kDbgRows][kDbgCols] table (columns 0-7 clock64, 8 / 9 = globaltimer at entry /
// exit, comparable across CTAs and launches; row kDbgRows - 1 = the launch's shape).  The launch index is handed out by the host per call.
constexpr int kDbgRows = 300, kDbgCols = 12;   // columns 10 / 11: statistics warp 10 finished its sub-tiles / issued its atomics
__device__ unsigned long long* g_conv1d_dbg = nullptr;
static bool h_conv1d_dbg = false;
static int h_conv1d_dbg_launches = 1, h_conv1d_dbg_next = 0;
__device__ __forceinline__ unsigned long long dbg_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// (nothing is evaluated unless a table is registered: reading %globaltimer takes a sizeable fraction of a microsecond)
#ifdef ERF_NO_TIMELINE
#define dbg_stamp(k) do { } while (0)
#else
#define dbg_stamp(k)                                                                                                                 \
  do {                                                                                                                                \
    unsigned long long* dbg_ = g_conv1d_dbg;                                                                                          \
    if (dbg_) dbg_[((size_t)p.dbg_launch * kDbgRows + blockIdx.x) * kDbgCols + (k)] = ((k) == 8 || (k) == 9) ? dbg_globaltimer() : (unsigned long long)clock64(); \
  } while (0)
#endif

constexpr int kMaxSlots = 8;
constexpr int kThreads = 384;  // warp 0 producer, warp 1 MMA, warps 2-5 / 6-9 epilogue groups, warps 10-11 per-channel statistics
// Statistics warps: 2, or 4 in the C = 128 statistics instantiation (448 threads) -- a 128-channel launch is 2.6 units = 5 sub-tiles per CTA, the two
// warps' serial pass over them (~0.45 us each) trailed the MMAs by 2 us (in-situ trace: tail 3.0-3.5 us against 1.3 us without statistics)
template <int C, bool RED> struct StatWarps { static constexpr int value = (RED && C == 128) ? 4 : 2; };

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

template <int C, bool RED, int NB>
__global__ void __launch_bounds__(320 + 32 * StatWarps<C, RED>::value, C == 16 ? 2 : 1) conv1d_tc_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                                                           const __grid_constant__ CUtensorMap tmap_aux, const Conv1dTcParams p) {
  constexpr int CH = C < 64 ? C : 64;           // channels per K chunk (one swizzle row)
  constexpr int RB = CH * 2;                    // bytes per pixel row of a chunk: 128 (128B swizzle) or 32 (32B swizzle, C = 16)
  constexpr int KC = C / CH;
  constexpr int KSTEPS = CH / 16;               // K = 16 MMAs per chunk
  constexpr int CPR = RB / 16;                  // 16-byte chunks per row
  constexpr uint32_t LAYOUT = RB == 128 ? LAYOUT_SW128 : LAYOUT_SW32;
  constexpr int W_BYTES = 3 * KC * C * RB;      // all tap weights
  constexpr int W_PAD = ((W_BYTES + 1023) / 1024) * 1024;
  constexpr int STG_BYTES = 128 * RB;           // one staging sub-tile: 128 px x one k-chunk
  constexpr int STG_PAD = ((STG_BYTES + 1023) / 1024) * 1024;
  // NB = staging sub-tiles per epilogue group: 2, or 3 with a residual / mask operand (see Conv1dTcParams::nb)
  constexpr int STG_TOTAL = 2 * NB * STG_PAD;
  constexpr int EW = C >= 32 ? 32 : 16;         // epilogue column chunk
  // Per-channel statistics (BN sum / sum of squares, bias-gradient column sums) of the STORED output are taken by warps 10-11 from the
  // staged bf16 sub-tiles (see their role below).
  // RED = false (no statistics requested: inference, the plain convs of a training step) compiles that role out.
  constexpr uint32_t TMEM_COLS = C == 128 ? 256 : (C == 64 ? 128 : 32);   // two accumulators of C columns
  constexpr uint32_t IDESC = make_idesc_bf16(128, C);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* s_w = smem;
  uint8_t* s_stg = smem + W_PAD;
  uint8_t* s_a = s_stg + STG_TOTAL;
  __shared__ uint64_t full_bar[kMaxSlots], empty_bar[kMaxSlots], tfull_bar[2], tempty_bar[2], w_bar, aux_bar[6];
  __shared__ uint64_t sfull_bar[6], sdone_bar[6];   // staged sub-tile ready / read by both statistics warps (per group x buffer)
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_bias[C];
  constexpr int SW = StatWarps<C, RED>::value;      // statistics warps, each takes 128 / SW rows of every sub-tile
  __shared__ float s_stat[SW][2 * C];               // per statistics warp: [C] sums, [C] sums of squares

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    dbg_stamp(0);
    dbg_stamp(8);
    if (g_conv1d_dbg && blockIdx.x == 0) {
      unsigned long long* m = g_conv1d_dbg + ((size_t)p.dbg_launch * kDbgRows + (kDbgRows - 1)) * kDbgCols;
      m[0] = C; m[1] = p.horizontal; m[2] = p.dil; m[3] = p.aux_mode; m[4] = (p.stats ? 1 : 0) | (p.colsum ? 2 : 0); m[5] = p.units; m[6] = gridDim.x;
    }
  }
  pdl_trigger();   // the set-up below overlaps the previous kernel's tail; nothing before pdl_wait() touches tensors an earlier kernel writes
  for (int i = tid; i < C; i += blockDim.x) s_bias[i] = p.bias ? p.bias[i] : 0.f;
  const bool want_red = RED && (p.stats != nullptr || p.colsum != nullptr);
  if (tid == 0) {
    for (int s = 0; s < p.slots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    mbar_init(&w_bar, 1);
    for (int a = 0; a < 6; ++a) {
      mbar_init(&aux_bar[a], 1);
      mbar_init(&sfull_bar[a], 1);
      mbar_init(&sdone_bar[a], SW);
    }
    fence_barrier_init();
    if (p.w_static) {   // static weights: fetch them while the previous kernel's last CTAs are still running
      mbar_arrive_expect_tx(&w_bar, W_BYTES);
      for (int off = 0; off < W_BYTES; off += 32768) {
        const int n = W_BYTES - off < 32768 ? W_BYTES - off : 32768;
        bulk_g2s(s_w + off, reinterpret_cast<const uint8_t*>(p.w_img) + off, n, &w_bar);
      }
    }
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_out);
    if (p.aux_mode) tma_prefetch_desc(&tmap_aux);
  }
  __syncthreads();
  if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int segs = (p.W + 127) / 128;
  pdl_wait();
  if (tid == 0) dbg_stamp(1);

  if (warp == 0) {
    // ================= TMA producer: whole warp in uniform control flow, one elected lane issues =================
    {
      // weights: one bulk copy per <= 32 KB piece, completing on w_bar
      if (!p.w_static && elect_one()) {
        mbar_arrive_expect_tx(&w_bar, W_BYTES);
        for (int off = 0; off < W_BYTES; off += 32768) {
          const int n = W_BYTES - off < 32768 ? W_BYTES - off : 32768;
          bulk_g2s(s_w + off, reinterpret_cast<const uint8_t*>(p.w_img) + off, n, &w_bar);
        }
      }
      __syncwarp();
      int slot = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const int xs = u % segs, y = (u / segs) % p.H, n = u / (segs * p.H);
        if (p.horizontal) {
#pragma unroll 1
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)((128 + 2 * p.dil) * RB));
              tma_load_4d(s_a + (size_t)slot * p.slot_bytes, &tmap_in, &full_bar[slot], kc * CH, xs * 128 - p.dil, y, n);
            }
            __syncwarp();
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
          }
        } else {
#pragma unroll 1
          for (int t = 0; t < 3; ++t) {
#pragma unroll 1
            for (int kc = 0; kc < KC; ++kc) {
              mbar_wait(&empty_bar[slot], phase ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(128 * RB));
                tma_load_4d(s_a + (size_t)slot * p.slot_bytes, &tmap_in, &full_bar[slot], kc * CH, xs * 128, y + (t - 1) * p.dil, n);
              }
              __syncwarp();
              if (++slot == p.slots) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: the whole warp runs the loop (uniform control flow), one elected lane issues =================
    {
      int slot = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      const uint32_t w_addr = smem_u32(s_w);
      const uint32_t desc_hi = ((8u * RB) >> 4) | (1u << 14) | (LAYOUT << 29);  // SBO = 8 rows, version 1, swizzle mode
      mbar_wait(&w_bar, 0);
      if (lane == 0) dbg_stamp(2);
      bool first_unit = true;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        const uint32_t d_tmem = tmem_base + acc * C;
        if (p.horizontal) {
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
            mbar_wait(&full_bar[slot], phase);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(s_a + (size_t)slot * p.slot_bytes);
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                const uint32_t a_lo = (((a_addr + t * p.dil * RB) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t b_lo = (((w_addr + (t * KC + kc) * (C * RB)) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + k * 2), ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + k * 2), IDESC,
                            (t | kc | k) != 0 ? 1u : 0u);
              }
              umma_commit(&empty_bar[slot]);
            }
            __syncwarp();
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
          }
        } else {
#pragma unroll
          for (int t = 0; t < 3; ++t) {
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
              mbar_wait(&full_bar[slot], phase);
              tcgen05_fence_after();
              const uint32_t a_addr = smem_u32(s_a + (size_t)slot * p.slot_bytes);
              const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | (1u << 16);
              const uint32_t b_lo = (((w_addr + (t * KC + kc) * (C * RB)) >> 4) & 0x3FFF) | (1u << 16);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k)
                  umma_bf16(d_tmem, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + k * 2), ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + k * 2), IDESC,
                            (t | kc | k) != 0 ? 1u : 0u);
                umma_commit(&empty_bar[slot]);   // this (tap, k-chunk) tile is reusable once its MMAs have read it
              }
              __syncwarp();
              if (++slot == p.slots) { slot = 0; phase ^= 1; }
            }
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);     // accumulator ready for the epilogue
        __syncwarp();
        if (first_unit && lane == 0) dbg_stamp(3);
        first_unit = false;
        if (lane == 0) dbg_stamp(4);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 10) {
    // ================= per-channel statistics of every staged sub-tile on the CUDA cores =================
    // (sum and sum of squares for the BN batch statistics, sum alone for a bias gradient.)  Round 1 took them as two more MMAs per
    // sub-tile (Y^T Y Gram diagonal, Y^T 1): correct, but the Gram MMAs cost as many tensor-pipe cycles as the convolution itself
    // (+5 us per launch, 34 launches per step).  Two warps reading the staged bf16 tile back -- one conflict-free 4-byte word per lane
    // and pixel row, 64 rows each -- cost ~220 issue slots per scheduler and sub-tile and leave the tensor pipe to the convolution.
    if (want_red) {
      // 16-byte loads: a lane owns ONE logical 16-byte chunk (8 channels) and reads it from 32 / CPR consecutive rows per instruction -- the
      // lanes of a row cover the whole (swizzled) row, so the access is conflict-free -- and accumulates in packed fp32x2.  (The first version
      // read one 4-byte word per lane and row: ~800 instructions per warp and sub-tile; the in-situ timeline, tools/trace_step_conv1d.py,
      // showed the statistics warps setting the pace of every launch that had them and ~4 us of extra tail.)
      const int sw = warp - 10;                       // rows [RPW*sw, RPW*sw + RPW) of every sub-tile
      constexpr int RPW = 128 / SW;
      constexpr int RPI = 32 / CPR;                   // rows per load instruction: 4 (128-byte rows) or 16 (32-byte rows)
      const int jl = lane % CPR;                      // logical 16-byte chunk = channels 8*jl .. 8*jl + 7 of the k-chunk
      const int rl = lane / CPR;
      float2 sa[KC][4], qa[KC][4];
#pragma unroll
      for (int kc = 0; kc < KC; ++kc)
#pragma unroll
        for (int e = 0; e < 4; ++e) sa[kc][e] = qa[kc][e] = make_float2(0.f, 0.f);
      int cnt[2] = {0, 0};                            // buffer of the next sub-tile of group g (rotates over NB)
      uint32_t ph = 0;                                // bit (g * 3 + b): phase of sfull_bar[g * NB + b]
      int i = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++i) {
        const int g = i & 1;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const int b = cnt[g];
          cnt[g] = b + 1 == NB ? 0 : b + 1;
          mbar_wait(&sfull_bar[g * NB + b], (ph >> (g * 3 + b)) & 1u);
          ph ^= 1u << (g * 3 + b);
          const uint8_t* tile = s_stg + (g * NB + b) * STG_PAD;
          const int vrows = min(128, p.W - (u % segs) * 128);      // pixel rows of this unit inside the image (ragged last segment)
          float2 s2[4], q2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) s2[e] = q2[e] = make_float2(0.f, 0.f);
          auto accumulate = [&](auto with_squares) {      // a bias gradient (colsum alone) needs no squares: one branch per sub-tile
#pragma unroll
            for (int it = 0; it < RPW / RPI; ++it) {
              const int r = sw * RPW + it * RPI + rl;
              const int key = CPR == 8 ? (r & 7) : ((r >> 2) & 1);
              uint4 w4 = *reinterpret_cast<const uint4*>(tile + r * RB + ((jl ^ key) * 16));
              if (r >= vrows) w4 = make_uint4(0u, 0u, 0u, 0u);
              const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = make_float2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xFFFF0000u));
                s2[e] = __fadd2_rn(s2[e], v);
                if (decltype(with_squares)::value) q2[e] = __ffma2_rn(v, v, q2[e]);
              }
            }
          };
          if (p.stats) accumulate(std::true_type{}); else accumulate(std::false_type{});
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            sa[kc][e] = __fadd2_rn(sa[kc][e], s2[e]);
            qa[kc][e] = __fadd2_rn(qa[kc][e], q2[e]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sdone_bar[g * NB + b]);     // this warp has read the sub-tile (the barrier counts both statistics warps)
        }
      }
      if (warp == 10 && lane == 0) dbg_stamp(10);
      // combine: lanes sharing a chunk (same lane % CPR), then the two warps through shared memory, one atomic per channel
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          for (int off = CPR; off < 32; off <<= 1) {
            sa[kc][e].x += __shfl_xor_sync(0xffffffffu, sa[kc][e].x, off);
            sa[kc][e].y += __shfl_xor_sync(0xffffffffu, sa[kc][e].y, off);
            qa[kc][e].x += __shfl_xor_sync(0xffffffffu, qa[kc][e].x, off);
            qa[kc][e].y += __shfl_xor_sync(0xffffffffu, qa[kc][e].y, off);
          }
          if (lane < CPR) {
            const int c = kc * CH + 8 * jl + 2 * e;
            s_stat[sw][c] = sa[kc][e].x; s_stat[sw][c + 1] = sa[kc][e].y;
            s_stat[sw][C + c] = qa[kc][e].x; s_stat[sw][C + c + 1] = qa[kc][e].y;
          }
        }
      }
      asm volatile("bar.sync 5, %0;" ::"n"(32 * SW) : "memory");   // the statistics warps
      for (int c = tid - 320; c < C; c += 32 * SW) {
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int w = 0; w < SW; ++w) { sum += s_stat[w][c]; sq += s_stat[w][C + c]; }
        if (p.colsum) atomicAdd(&p.colsum[c], sum);
        if (p.stats) {
          atomicAdd(&p.stats[c], (double)sum);
          atomicAdd(&p.stats[C + c], (double)sq);
        }
      }
      if (warp == 10 && lane == 0) dbg_stamp(11);
    }
  } else {
    // ================= epilogue: group 0 = warps 2..5, group 1 = warps 6..9; TMEM lane quadrant = warp % 4 =========
    const int grp = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // pixel within the unit = TMEM lane = staging row
    const bool leader = (((warp - 2) & 3) == 0 && lane == 0);
    // swizzled position of logical 16-byte chunk j of this row inside a staging sub-tile
    const int sw_key = CPR == 8 ? (row & 7) : ((row >> 2) & 1);
    uint8_t* stg_g = s_stg + grp * NB * STG_PAD;
    uint64_t* aux_g = &aux_bar[grp * NB];
    const uint32_t acc = grp;                  // accumulator owned by this group
    uint32_t acc_phase = 0, aux_phase = 0;     // aux_phase / sdone_phase: bit b = phase of buffer b
    int step = 0;                              // staging steps (unit, k-chunk) done by this group
    uint32_t sdone_phase = 0;
    int b = 0, bprev = NB - 1;                 // staging buffer of this step / of the previous step
    // units of this CTA alternate between the two groups
    const int u_first = blockIdx.x + grp * gridDim.x, u_stride = 2 * gridDim.x;
    // aux sub-tile (residual / mask operand) of this group's step st -> staging buffer st % NB
    auto prefetch_aux = [&](int st) {
      const int uu = u_first + (st / KC) * u_stride, kk = st % KC;
      if (uu >= p.units) return;
      const int bb = st % NB;
      const int xs2 = uu % segs, y2 = (uu / segs) % p.H, n2 = uu / (segs * p.H);
      mbar_arrive_expect_tx(&aux_g[bb], STG_BYTES);
      tma_load_4d(stg_g + bb * STG_PAD, &tmap_aux, &aux_g[bb], kk * CH, xs2 * 128, y2, n2);
    };
    if (leader && p.aux_mode) {                // the first NB - 1 steps' operands
      for (int st = 0; st < NB - 1; ++st) prefetch_aux(st);
    }
    for (int u = u_first; u < p.units; u += u_stride) {
      const int xs = u % segs, y = (u / segs) % p.H, n = u / (segs * p.H);
      mbar_wait(&tfull_bar[acc], acc_phase);
      acc_phase ^= 1;
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + acc * C + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int kc = 0; kc < KC; ++kc, ++step) {
        uint8_t* my_row = stg_g + b * STG_PAD + row * RB;
        // both 32-column TMEM loads of this k-chunk are in flight before the first wait
        uint32_t r[CH];
        if (CH == 64) { tmem_ld32(taddr + kc * CH, r); tmem_ld32(taddr + kc * CH + 32, r + 32); }
        else tmem_ld16(taddr, r);
        if (p.aux_mode) {
          mbar_wait(&aux_g[b], (aux_phase >> b) & 1u);
          aux_phase ^= 1u << b;
        }
        tmem_ld_wait();
        if (kc == KC - 1) {                      // all TMEM reads of this unit are done: release the accumulator early
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += EW) {
          float v[EW];
#pragma unroll
          for (int j = 0; j < EW; ++j) v[j] = __uint_as_float(r[c0 + j]) + s_bias[kc * CH + c0 + j];
          const int j0 = c0 / 8;                 // first logical 16-byte chunk of this column chunk
          uint4 aux[EW / 8];
          if (p.aux_mode) {
#pragma unroll
            for (int q = 0; q < EW / 8; ++q) aux[q] = *reinterpret_cast<const uint4*>(my_row + (((j0 + q) ^ sw_key) * 16));
          }
          if (p.aux_mode == 1) {
#pragma unroll
            for (int q = 0; q < EW / 8; ++q) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&aux[q]);
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
          if (p.aux_mode == 2) {
#pragma unroll
            for (int q = 0; q < EW / 8; ++q) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&aux[q]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (!(__low2float(h[e]) > 0.f)) v[q * 8 + 2 * e] = 0.f;
                if (!(__high2float(h[e]) > 0.f)) v[q * 8 + 2 * e + 1] = 0.f;
              }
            }
          }
#pragma unroll
          for (int q = 0; q < EW / 8; ++q) {
            uint4 t;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
            *reinterpret_cast<uint4*>(my_row + (((j0 + q) ^ sw_key) * 16)) = t;
          }
        }
        fence_proxy_async_smem();              // staging writes -> visible to the TMA store
        if (leader) tma_store_wait_read0();    // the store of step-1 (other buffer) has finished reading its sub-tile
        epi_bar_sync(grp);
        if (leader) {
          tma_store_4d(&tmap_out, stg_g + b * STG_PAD, kc * CH, xs * 128, y, n);
          tma_store_commit();
          if (want_red) mbar_arrive(&sfull_bar[grp * NB + b]);   // statistics are taken over the values as STORED (bf16)
        }
        if (want_red && step >= 1) {           // the statistics MMAs of the previous step have read the other buffer
          mbar_wait(&sdone_bar[grp * NB + bprev], (sdone_phase >> bprev) & 1u);
          sdone_phase ^= 1u << bprev;
        }
        // the operand of step + NB - 1 goes into the buffer of step - 1: its store has been read (wait above) and, with statistics, so have they
        if (leader && p.aux_mode) prefetch_aux(step + NB - 1);
        bprev = b;
        b = b + 1 == NB ? 0 : b + 1;
      }
    }
    if (leader) tma_store_wait_read0();   // the CTA may exit once the stores have READ shared memory; the writes complete with the grid
    if (leader) dbg_stamp(5 + grp);
  }
  // ---- teardown ----------------------------------------------------------------------------------------------------
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    dbg_stamp(7);
    dbg_stamp(9);
  }
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

// NHWC bf16 tensor [N][H][W][C] viewed as (C, W, H, N); box = {min(C,64), box_w, 1, 1}
static int make_nhwc_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int box_w) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -1;
  const int CH = C < 64 ? C : 64;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)CH, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CH == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace erfk

using namespace erfk;

// tools only: buf = [launches][300][10] u64 (nullptr switches it off); launch k of the calls that follow writes row block k % launches
extern "C" int erf_conv1d_tc_set_debug(unsigned long long* buf, int launches) {
  h_conv1d_dbg = buf != nullptr;
  h_conv1d_dbg_launches = launches > 0 ? launches : 1;
  h_conv1d_dbg_next = 0;
  return cudaMemcpyToSymbol(g_conv1d_dbg, &buf, sizeof(buf)) == cudaSuccess ? ERF_OK : ERF_ERR_CUDA;
}

static int g_cap_main = 0, g_cap_side = 0;
int erfk::launch_cap_main() { return g_cap_main; }
int erfk::launch_cap_side() { return g_cap_side; }
// SMs a 3-tap conv / data-gradient launch (main) and a weight-gradient launch (side) may take; 0 = all.  Host-side state read at launch (or
// graph-capture) time: the trainer sets it around a backward pass whose weight gradients run on a second stream.
extern "C" int erf_set_launch_cap(int main_sms, int side_sms) {
  g_cap_main = main_sms > 0 ? main_sms : 0;
  g_cap_side = side_sms > 0 ? side_sms : 0;
  return ERF_OK;
}

extern "C" int erf_conv1d_tc_supported(int C, int W, int dil) {
  // any width from 16 pixels: rows are cut into 128-pixel units, the last one may be ragged (round 1 required W % 128 == 0, which left the
  // reference fork's own 640 x 480 frames -- widths 320 / 160 / 80 -- and the 224 x 224 ImageNet net on the CUDA-core kernels)
  return (C == 16 || C == 64 || C == 128) && W >= 16 && dil >= 1 && dil <= 32;
}

extern "C" int erf_conv1d_tc(const void* in, const void* w_img, const float* bias, void* out, const void* res, const void* mask, int N, int H, int W,
                             int C, int horizontal, int dil, int relu, double* stats, float* colsum, erf_stream_t stream) {
  ERF_CHECK_ARG(in && w_img && out, "erf_conv1d_tc: null pointer");
  ERF_CHECK_ARG(N > 0 && H > 0, "erf_conv1d_tc: empty tensor");
  ERF_CHECK_ARG(!(res && mask), "erf_conv1d_tc: res and mask are mutually exclusive");
  if (!erf_conv1d_tc_supported(C, W, dil)) {
    set_error("erf_conv1d_tc: unsupported shape C=%d W=%d dil=%d (need C in {16,64,128}, W >= 16, dil <= 32)", C, W, dil);
    return ERF_ERR_UNSUPPORTED;
  }
  const int CH = C < 64 ? C : 64, RB = CH * 2, KC = C / CH;
  const int box_w = horizontal ? 128 + 2 * dil : 128;
  const void* aux = res ? res : mask;
  CUtensorMap tin, tout, taux;
  if (make_nhwc_map(&tin, in, N, H, W, C, box_w) != 0 || make_nhwc_map(&tout, out, N, H, W, C, 128) != 0 ||
      make_nhwc_map(&taux, aux ? aux : out, N, H, W, C, 128) != 0) {
    set_error("erf_conv1d_tc: cuTensorMapEncodeTiled failed");
    return ERF_ERR_CUDA;
  }
  Conv1dTcParams p;
  p.w_img = (const __nv_bfloat16*)w_img;
  p.bias = bias;
  p.N = N; p.H = H; p.W = W;
  p.horizontal = horizontal; p.dil = dil; p.relu = relu & 1; p.w_static = (relu >> 1) & 1;   // relu: bit 0 = ReLU, bit 1 = static weights
  p.aux_mode = res ? 1 : (mask ? 2 : 0);
  p.stats = stats;
  p.colsum = colsum;
  p.dbg_launch = h_conv1d_dbg ? (h_conv1d_dbg_next++ % h_conv1d_dbg_launches) : 0;
  p.units = N * H * ((W + 127) / 128);
  if (horizontal) {
    p.sub_bytes = ((((box_w + 7) / 8) * 8 * RB + 1023) / 1024) * 1024;
  } else {
    p.sub_bytes = ((128 * RB + 1023) / 1024) * 1024;
  }
  p.slot_bytes = p.sub_bytes;   // one (tap, k-chunk) tile [vertical] or one k-chunk of the haloed row tile [horizontal]
  const int w_pad = ((3 * KC * C * RB + 1023) / 1024) * 1024;
  p.nb = (p.aux_mode && C <= 64) ? 3 : 2;
  {
    const char* nb_env = getenv("ERF_CONV1D_NB");   // experiments: 2 = the two-buffer scheme everywhere
    if (nb_env && atoi(nb_env) == 2) p.nb = 2;
  }
  const int stg_pad = 2 * p.nb * (((128 * RB + 1023) / 1024) * 1024);
  const int budget = 222 * 1024 - w_pad - stg_pad;   // + 1 KB alignment slack + ~3 KB of static shared memory (barriers, bias, statistics) <= 227 KB
  p.slots = budget / p.slot_bytes;
  if (p.slots > kMaxSlots) p.slots = kMaxSlots;
  if (p.slots < 2) {
    set_error("erf_conv1d_tc: tile does not fit in shared memory");
    return ERF_ERR_UNSUPPORTED;
  }
  size_t smem = (size_t)w_pad + stg_pad + (size_t)p.slots * p.slot_bytes + 1024;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // C = 16 tiles are 4 KB: the per-unit TMA -> MMA -> epilogue -> store chain is latency-bound, and the kernel's footprint (~60 KB of
  // shared memory, 72 registers, 32 TMEM columns) lets two CTAs share an SM
  const int per_sm = C == 16 ? 2 : 1;   // measured: 35 -> 23 us (vertical), 21 -> 15 us (horizontal) at batch 6; 3 per SM is slower
  if (launch_cap_main() > 0 && launch_cap_main() < sms) sms = launch_cap_main();
  int grid = p.units < sms * per_sm ? p.units : sms * per_sm;
  cudaError_t e;
  cudaStream_t st = (cudaStream_t)stream;
  const bool red = stats != nullptr || colsum != nullptr;
#define CONV1D_LAUNCH3(C_, RED_, NB_)                                                                                                 \
  e = cudaFuncSetAttribute(conv1d_tc_kernel<C_, RED_, NB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
  if (e == cudaSuccess && C_ == 16) e = prefer_max_smem(conv1d_tc_kernel<C_, RED_, NB_>);                                             \
  if (e == cudaSuccess) launch_pdl(conv1d_tc_kernel<C_, RED_, NB_>, dim3(grid), dim3(320 + 32 * StatWarps<C_, RED_>::value), smem, (cudaStream_t)(st), tin, tout, taux, p);
#define CONV1D_LAUNCH(C_, RED_) if (p.nb == 3) { CONV1D_LAUNCH3(C_, RED_, 3) } else { CONV1D_LAUNCH3(C_, RED_, 2) }
  if (C == 16) {
    if (red) { CONV1D_LAUNCH(16, true) } else { CONV1D_LAUNCH(16, false) }
  } else if (C == 64) {
    if (red) { CONV1D_LAUNCH(64, true) } else { CONV1D_LAUNCH(64, false) }
  } else {
    if (red) { CONV1D_LAUNCH3(128, true, 2) } else { CONV1D_LAUNCH3(128, false, 2) }
  }
#undef CONV1D_LAUNCH3
#undef CONV1D_LAUNCH
  if (e != cudaSuccess) {
    set_error("erf_conv1d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
  ERF_CHECK_LAUNCH("erf_conv1d_tc");
  return ERF_OK;
}
