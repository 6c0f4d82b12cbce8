// Train-mode BatchNorm in ONE launch per direction (nn.BatchNorm2d(C, eps=1e-3) inside eval/erfnet.py:12-61,99-108).
//
//   forward : the batch statistics arrive as fp64 sums from the producing conv's epilogue (Gram / ones MMAs); every CTA derives the
//             per-channel scale / shift itself (C <= 128 values), CTA 0 also writes mean / invstd for the backward pass and updates the
//             running statistics -- no separate finalisation launch;  out = relu?( (x * sc + sh) * drop[n,c] (+ res) ).
//   backward: dy = gamma * is * (gb - mean(gb) - xhat * mean(gb * xhat)),  gb = g * (mask > 0) * drop[n,c].  The two per-channel means need the
//             whole tensor, so the kernel runs in two phases separated by a GRID BARRIER (all CTAs are co-resident: one per SM).  What a
//             thread read in phase 1 (gm = g * (mask > 0), exact in bf16, and y) stays in its registers and in shared memory, so phase 2
//             reads nothing again for the tensors of the bench workload (the round-1 pair of kernels read the three tensors twice).
//             Items beyond the cache (larger batches) are simply re-read.
#include "bn_stream.cuh"

namespace erfk {

// ---------------------------------------------------------------------------------------------------------------------------------
template <typename T, bool DROP>
__global__ void __launch_bounds__(BS_THREADS, 2) bn_fwd_apply_kernel(const T* __restrict__ x, int x_ld, const double* __restrict__ stats, double count,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* rm, float* rv,
                                                                     float momentum, float eps, const float* __restrict__ drop, const T* __restrict__ res,
                                                                     int res_ld, T* __restrict__ out, int out_ld, long long HW, int C, long long n_px,
                                                                     int px_per_block, int relu, float* mean_out, float* invstd_out) {
  pdl_enter();
  extern __shared__ __align__(16) float s_dyn[];   // [2C]: scale, shift
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s = stats[c], sq = stats[C + c];
    const double mu = s / count;
    double var = sq / count - mu * mu;
    if (var < 0) var = 0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * is;
    s_dyn[c] = sc;
    s_dyn[C + c] = beta[c] - (float)mu * sc;
    if (blockIdx.x == 0) {
      mean_out[c] = (float)mu;
      invstd_out[c] = is;
      if (rm) {   // momentum update with the UNBIASED variance (torch semantics)
        const double unb = count > 1 ? var * count / (count - 1) : var;
        rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mu;
        rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unb;
      }
    }
  }
  __syncthreads();
  const PxRange r = px_range(C, n_px, px_per_block);
  float scn[8], shn[8];     // scale / shift of the current sample (Dropout2d multiplier folded in)
  lds8(s_dyn + r.c0, scn);
  lds8(s_dyn + C + r.c0, shn);
  long long cur_n = -1;
  for (long long p = r.p0 + r.lane_px; p < r.p1; p += UNR * r.lanes) {
    Raw8<T> xv[UNR], rv8[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        xv[u].load(x + (size_t)q * x_ld + r.c0);
        if (res) rv8[u].load(res + (size_t)q * res_ld + r.c0);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long q = p + u * r.lanes;
      if (q < r.p1) {
        if (DROP) {
          const long long n = img_of(q, HW);
          if (n != cur_n) {   // Dropout2d multiplier of this (sample, channel), folded into scale and shift like bn_finalize does
            cur_n = n;
            lds8(s_dyn + r.c0, scn);
            lds8(s_dyn + C + r.c0, shn);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float d = drop[n * C + r.c0 + k];
              scn[k] *= d;
              shn[k] *= d;
            }
          }
        }
        float f[8], t[8];
        xv[u].unpack(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f[k] * scn[k] + shn[k];
        if (res) {
          rv8[u].unpack(t);
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

// ---------------------------------------------------------------------------------------------------------------------------------
struct BnBwdFusedParams {
  const void *g, *y, *mask;
  int g_ld, y_ld, mask_ld;
  const float *drop, *gamma, *beta, *mean, *invstd;
  long long HW, n_px;
  int C, px_per_block, smem_items;
  int ring_gy, ring_m;   // bf16 kernel: first slot of the (g, y) ring of items beyond the cache, first slot of the mask ring
  double* sums;        // [sum_copies][2C], zeroed by the caller
  int sum_copies;      // lean kernel: CTA b adds into copy b % sum_copies (148 same-address fp64 atomics serialise in the L2: ~3 us at the end of phase 1)
  unsigned* counter;   // zeroed by the caller
  void* dy;
  int dy_ld;
  float *dgamma, *dbeta;
  void* gmasked;
  int gmasked_ld;
};

template <typename T> struct Pack8 {   // how 8 values are kept between the phases
  static __device__ __forceinline__ void pack(Raw8<T>& dst, const float (&f)[8]);
};
template <> struct Pack8<__nv_bfloat16> {
  static __device__ __forceinline__ void pack(Raw8<__nv_bfloat16>& dst, const float (&f)[8]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&dst.q);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  }
};
template <> struct Pack8<float> {
  static __device__ __forceinline__ void pack(Raw8<float>& dst, const float (&f)[8]) {
    dst.a = make_float4(f[0], f[1], f[2], f[3]);
    dst.b = make_float4(f[4], f[5], f[6], f[7]);
  }
};

template <typename T, int REGI, int SUNR>
__global__ void __launch_bounds__(BS_THREADS, 1) bn_bwd_fused_kernel(const BnBwdFusedParams p) {
  pdl_enter();
  extern __shared__ __align__(16) float s_dyn[];
  const int C = p.C;
  float* s_coef = s_dyn;                                   // phase 1: [2C] mean, invstd; phase 2: [4C] k1, k2, k3, mean
  float* s_part = s_dyn + 4 * C;                           // [warps][2C] reduction rows
  Raw8<T>* s_cache = reinterpret_cast<Raw8<T>*>(s_part + (BS_THREADS / 32) * 2 * C);   // [smem_items][2][BS_THREADS]
  const T* __restrict__ g = (const T*)p.g;
  const T* __restrict__ y = (const T*)p.y;
  const T* __restrict__ mask = (const T*)p.mask;
  T* __restrict__ dy = (T*)p.dy;
  T* __restrict__ gmasked = (T*)p.gmasked;
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_coef[i] = p.mean[i]; s_coef[C + i] = p.invstd[i]; }
  __syncthreads();
  const PxRange r = px_range(C, p.n_px, p.px_per_block);
  const int items = p.px_per_block / r.lanes;              // px_per_block is a multiple of lanes
  const long long q0 = r.p0 + r.lane_px;
  float acc[2][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = 0.f;
  Raw8<T> cg[REGI], cy[REGI];

  // gm = g * (mask > 0) (kept), gb = gm * drop; acc += (gb, gb * xhat)
  auto phase1 = [&](long long q, Raw8<T>& gv, const Raw8<T>& yv, const Raw8<T>& mv) {
    float gg[8], t[8];
    gv.unpack(gg);
    if (mask) {
      mv.unpack(t);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (!(t[k] > 0.f)) gg[k] = 0.f;
      Pack8<T>::pack(gv, gg);
    }
    if (p.drop) {
      const float* dp = p.drop + img_of(q, p.HW) * C + r.c0;
#pragma unroll
      for (int k = 0; k < 8; ++k) gg[k] *= dp[k];
    }
    float mu[8], is[8];
    yv.unpack(t);
    lds8(s_coef + r.c0, mu);
    lds8(s_coef + C + r.c0, is);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[0][k] += gg[k];
      acc[1][k] += gg[k] * (t[k] - mu[k]) * is[k];
    }
  };

  // ---- phase 1: register-cached items (all loads of these items are issued before the first use) ----
  {
    Raw8<T> mv[REGI];
#pragma unroll
    for (int k = 0; k < REGI; ++k) {
      const long long q = q0 + (long long)k * r.lanes;
      if (k < items && q < r.p1) {
        cg[k].load(g + (size_t)q * p.g_ld + r.c0);
        cy[k].load(y + (size_t)q * p.y_ld + r.c0);
        if (mask) mv[k].load(mask + (size_t)q * p.mask_ld + r.c0);
      }
    }
#pragma unroll
    for (int k = 0; k < REGI; ++k) {
      const long long q = q0 + (long long)k * r.lanes;
      if (k < items && q < r.p1) phase1(q, cg[k], cy[k], mv[k]);
    }
  }
  // ---- phase 1: remaining items, SUNR in flight; the first smem_items of them are parked in shared memory ----
  for (int k = REGI; k < items; k += SUNR) {
    Raw8<T> gv[SUNR], yv[SUNR], mv[SUNR];
#pragma unroll
    for (int u = 0; u < SUNR; ++u) {
      const long long q = q0 + (long long)(k + u) * r.lanes;
      if (k + u < items && q < r.p1) {
        gv[u].load(g + (size_t)q * p.g_ld + r.c0);
        yv[u].load(y + (size_t)q * p.y_ld + r.c0);
        if (mask) mv[u].load(mask + (size_t)q * p.mask_ld + r.c0);
      }
    }
#pragma unroll
    for (int u = 0; u < SUNR; ++u) {
      const long long q = q0 + (long long)(k + u) * r.lanes;
      if (k + u < items && q < r.p1) {
        phase1(q, gv[u], yv[u], mv[u]);
        const int si = k + u - REGI;
        if (si < p.smem_items) {
          s_cache[(size_t)(2 * si) * BS_THREADS + threadIdx.x] = gv[u];
          s_cache[(size_t)(2 * si + 1) * BS_THREADS + threadIdx.x] = yv[u];
        }
      }
    }
  }
  block_channel_reduce<2>(acc, C, r.c0, s_part, [&](int i, float t) { atomicAdd(&p.sums[i], (double)t); });

  grid_barrier(p.counter, gridDim.x);

  // ---- phase 2: dy = k1 * gb + k2 * (y - mu) + k3 ----
  {
    const float inv_m = 1.f / (float)p.n_px;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      const float is = p.invstd[i], ga = p.gamma[i], mu = p.mean[i];
      const float s1 = (float)__ldcg(&p.sums[i]), s2 = (float)__ldcg(&p.sums[C + i]);
      s_coef[i] = ga * is;
      s_coef[C + i] = -ga * is * is * s2 * inv_m;
      s_coef[2 * C + i] = -ga * is * s1 * inv_m;
      s_coef[3 * C + i] = mu;
      if (blockIdx.x == 0) {
        if (p.dbeta) p.dbeta[i] = s1;
        if (p.dgamma) p.dgamma[i] = s2;
      }
    }
  }
  __syncthreads();
  auto phase2 = [&](long long q, const Raw8<T>& gmv, const Raw8<T>& yv) {
    float gg[8], t[8], k[8];
    gmv.unpack(gg);
    if (gmasked) Raw8<T>::store(gmasked + (size_t)q * p.gmasked_ld + r.c0, gg);
    if (p.drop) {
      const float* dp = p.drop + img_of(q, p.HW) * C + r.c0;
#pragma unroll
      for (int e = 0; e < 8; ++e) gg[e] *= dp[e];
    }
    lds8(s_coef + r.c0, k);
#pragma unroll
    for (int e = 0; e < 8; ++e) gg[e] *= k[e];
    yv.unpack(t);
    lds8(s_coef + 3 * C + r.c0, k);
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] -= k[e];
    lds8(s_coef + C + r.c0, k);
#pragma unroll
    for (int e = 0; e < 8; ++e) gg[e] += t[e] * k[e];
    lds8(s_coef + 2 * C + r.c0, k);
#pragma unroll
    for (int e = 0; e < 8; ++e) gg[e] += k[e];
    Raw8<T>::store(dy + (size_t)q * p.dy_ld + r.c0, gg);
  };
#pragma unroll
  for (int k = 0; k < REGI; ++k) {
    const long long q = q0 + (long long)k * r.lanes;
    if (k < items && q < r.p1) phase2(q, cg[k], cy[k]);
  }
  for (int k = REGI; k < items; ++k) {
    const long long q = q0 + (long long)k * r.lanes;
    if (q >= r.p1) break;
    const int si = k - REGI;
    Raw8<T> gv, yv;
    if (si < p.smem_items) {
      gv = s_cache[(size_t)(2 * si) * BS_THREADS + threadIdx.x];
      yv = s_cache[(size_t)(2 * si + 1) * BS_THREADS + threadIdx.x];
    } else {   // beyond the cache: read again and redo the mask
      gv.load(g + (size_t)q * p.g_ld + r.c0);
      yv.load(y + (size_t)q * p.y_ld + r.c0);
      if (mask) {
        Raw8<T> mv;
        mv.load(mask + (size_t)q * p.mask_ld + r.c0);
        float gg[8], t[8];
        gv.unpack(gg);
        mv.unpack(t);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (!(t[e] > 0.f)) gg[e] = 0.f;
        Pack8<T>::pack(gv, gg);
      }
    }
    phase2(q, gv, yv);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// bf16 fast path.  Same two-phase structure, built for memory-level parallelism and a short instruction stream:
//   * one 512-thread CTA per SM; a thread owns one 8-channel group and `items` pixels.  Its first REGI items live in registers (loaded
//     up front, 3 x REGI 16-byte loads in flight), the others arrive by cp.async straight into the thread's OWN shared-memory slots
//     (depth-D prefetch, no registers, no block barrier: a thread only ever touches its own slots); slots of cached items simply stay.
//   * ReLU mask either from the stored activation (packed bf16x2 compare + AND) or, for a plain BN -> ReLU (no dropout, no residual),
//     recomputed from y itself exactly as the forward pass did (fma(y, sc, sh) > 0): one tensor read less.
//   * sums: S1 = sum gb, S2c = sum gb * (y - mean) in fp32 per thread, fp64 across CTAs; dy = k1 * gb + k2 * y + k3.
enum { BN_MASK_NONE = 0, BN_MASK_TENSOR = 1, BN_MASK_AFFINE = 2 };
constexpr int BF_D = 4;   // cp.async prefetch depth (items)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void unpack_bf8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Timeline instrumentation (tools/timeline_bn.py): when a buffer is registered every CTA records globaltimer stamps (ns) of its phases.
__device__ unsigned long long* g_bn_dbg = nullptr;
__device__ __forceinline__ void bn_stamp(int k) {
#ifndef ERF_NO_TIMELINE
  if (g_bn_dbg && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_bn_dbg[(size_t)blockIdx.x * 8 + k] = t;
  }
#endif
}

template <int MODE, int REGI>
__global__ void __launch_bounds__(BS_THREADS, 1) bn_bwd_fused_bf16_kernel(const BnBwdFusedParams p) {
  bn_stamp(0);
  pdl_enter();
  bn_stamp(1);
  extern __shared__ __align__(16) float s_dyn[];
  const int C = p.C, tid = threadIdx.x;
  float* s_coef = s_dyn;                                   // [6C]: sc, sh (mask from y), mean | after the barrier also k1, k2, k3
  float* s_part = s_dyn + 6 * C;                           // [warps][2C]
  uint4* s_slot = reinterpret_cast<uint4*>(s_part + (BS_THREADS / 32) * 2 * C);   // [slots][BS_THREADS]
  const __nv_bfloat16* __restrict__ g = (const __nv_bfloat16*)p.g;
  const __nv_bfloat16* __restrict__ y = (const __nv_bfloat16*)p.y;
  const __nv_bfloat16* __restrict__ mask = (const __nv_bfloat16*)p.mask;
  __nv_bfloat16* __restrict__ dy = (__nv_bfloat16*)p.dy;
  __nv_bfloat16* __restrict__ gmasked = (__nv_bfloat16*)p.gmasked;
  for (int i = tid; i < C; i += blockDim.x) {
    const float is = p.invstd[i], mu = p.mean[i];
    if (MODE == BN_MASK_AFFINE) {   // the forward pass's scale / shift (bn_fwd_apply_kernel / bn_finalize_kernel), same expressions
      const float sc = p.gamma[i] * is;
      s_coef[i] = sc;
      s_coef[C + i] = p.beta[i] - mu * sc;
    }
    s_coef[2 * C + i] = mu;
  }
  __syncthreads();
  const PxRange r = px_range(C, p.n_px, p.px_per_block);
  const int items = p.px_per_block / r.lanes;
  const int ns = items > REGI ? items - REGI : 0;          // streamed items
  const long long q0 = r.p0 + r.lane_px;
  float acc[2][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = 0.f;
  uint4 cg[REGI], cy[REGI];

  // Dropout2d multipliers of the current sample for this thread's 8 channels: a thread's pixels ascend, so the row changes at most
  // N - 1 times (two 16-byte loads each) instead of eight scalar loads and a division per item
  float dropv[8];
  long long drop_n = -1, drop_end = 0;      // pixels [.., drop_end) belong to sample drop_n
  auto load_drop = [&](long long q) {
    if (q >= drop_end || drop_n < 0) {
      drop_n = img_of(q, p.HW);
      drop_end = (drop_n + 1) * p.HW;
      const float4* dp = reinterpret_cast<const float4*>(p.drop + drop_n * C + r.c0);
      const float4 a = dp[0], b = dp[1];
      dropv[0] = a.x; dropv[1] = a.y; dropv[2] = a.z; dropv[3] = a.w;
      dropv[4] = b.x; dropv[5] = b.y; dropv[6] = b.z; dropv[7] = b.w;
    }
  };
  // gq <- g * (mask > 0), in place
  auto apply_mask = [&](uint4& gq, const uint4& yq, const uint4& mq) {
    uint32_t* gw = reinterpret_cast<uint32_t*>(&gq);
    if (MODE == BN_MASK_TENSOR) {
      const uint32_t* mw = reinterpret_cast<const uint32_t*>(&mq);
      const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) gw[i] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&mw[i]), zero);
    } else if (MODE == BN_MASK_AFFINE) {
      float yf[8], sc[8], sh[8];
      unpack_bf8(yq, yf);
      lds8(s_coef + r.c0, sc);
      lds8(s_coef + C + r.c0, sh);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t keep = 0u;
        if (fmaf(yf[2 * i], sc[2 * i], sh[2 * i]) > 0.f) keep |= 0x0000FFFFu;
        if (fmaf(yf[2 * i + 1], sc[2 * i + 1], sh[2 * i + 1]) > 0.f) keep |= 0xFFFF0000u;
        gw[i] &= keep;
      }
    }
  };
  auto proc1 = [&](long long q, uint4& gq, const uint4& yq, const uint4& mq) {
    apply_mask(gq, yq, mq);
    float gf[8], yf[8], mu[8];
    unpack_bf8(gq, gf);
    if (p.drop) {
      load_drop(q);
#pragma unroll
      for (int k = 0; k < 8; ++k) gf[k] *= dropv[k];
    }
    unpack_bf8(yq, yf);
    lds8(s_coef + 2 * C + r.c0, mu);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[0][k] += gf[k];
      acc[1][k] = fmaf(gf[k], yf[k] - mu[k], acc[1][k]);
    }
  };
  auto slot_of = [&](int j, int& gs, int& ys, int& ms) {       // shared-memory slots of streamed item j
    if (j < p.smem_items) { gs = 2 * j; ys = 2 * j + 1; }
    else { gs = p.ring_gy + 2 * (j % BF_D); ys = gs + 1; }
    ms = p.ring_m + (j % BF_D);
  };
  auto issue = [&](int j, bool with_mask) {                     // start the copies of streamed item j (no-op past the end)
    const long long q = q0 + (long long)(REGI + j) * r.lanes;
    if (j < ns && q < r.p1) {
      int gs, ys, ms;
      slot_of(j, gs, ys, ms);
      cp_async16(&s_slot[(size_t)gs * BS_THREADS + tid], g + (size_t)q * p.g_ld + r.c0);
      cp_async16(&s_slot[(size_t)ys * BS_THREADS + tid], y + (size_t)q * p.y_ld + r.c0);
      if (with_mask) cp_async16(&s_slot[(size_t)ms * BS_THREADS + tid], mask + (size_t)q * p.mask_ld + r.c0);
    }
    cp_async_commit();
  };

  // ---- phase 1 ----
#pragma unroll
  for (int j = 0; j < BF_D - 1; ++j) issue(j, MODE == BN_MASK_TENSOR);
  {
    uint4 mv[REGI];
#pragma unroll
    for (int k = 0; k < REGI; ++k) {
      const long long q = q0 + (long long)k * r.lanes;
      if (k < items && q < r.p1) {
        cg[k] = *reinterpret_cast<const uint4*>(g + (size_t)q * p.g_ld + r.c0);
        cy[k] = *reinterpret_cast<const uint4*>(y + (size_t)q * p.y_ld + r.c0);
        if (MODE == BN_MASK_TENSOR) mv[k] = *reinterpret_cast<const uint4*>(mask + (size_t)q * p.mask_ld + r.c0);
      }
    }
#pragma unroll
    for (int k = 0; k < REGI; ++k) {
      const long long q = q0 + (long long)k * r.lanes;
      if (k < items && q < r.p1) proc1(q, cg[k], cy[k], mv[k]);
    }
  }
  for (int j = 0; j < ns; ++j) {
    issue(j + BF_D - 1, MODE == BN_MASK_TENSOR);
    cp_async_wait<BF_D - 1>();
    const long long q = q0 + (long long)(REGI + j) * r.lanes;
    if (q >= r.p1) continue;
    int gs, ys, ms;
    slot_of(j, gs, ys, ms);
    uint4 gq = s_slot[(size_t)gs * BS_THREADS + tid];
    const uint4 yq = s_slot[(size_t)ys * BS_THREADS + tid];
    uint4 mq = make_uint4(0u, 0u, 0u, 0u);
    if (MODE == BN_MASK_TENSOR) mq = s_slot[(size_t)ms * BS_THREADS + tid];
    proc1(q, gq, yq, mq);
    if (MODE != BN_MASK_NONE && j < p.smem_items) s_slot[(size_t)gs * BS_THREADS + tid] = gq;   // keep the masked gradient
  }
  cp_async_wait<0>();
  bn_stamp(2);
  block_channel_reduce<2>(acc, C, r.c0, s_part, [&](int i, float t) { atomicAdd(&p.sums[i], (double)t); });
  bn_stamp(3);

  grid_barrier(p.counter, gridDim.x);
  bn_stamp(4);

  // ---- phase 2 ----
  const int first_uncached = p.smem_items;                    // streamed items [first_uncached, ns) are read again
  for (int j = first_uncached; j < first_uncached + BF_D - 1; ++j) issue(j, MODE == BN_MASK_TENSOR);   // in flight while the cached items are processed
  {
    const double inv_m = 1.0 / (double)p.n_px;
    for (int i = tid; i < C; i += blockDim.x) {
      const double is = p.invstd[i], ga = p.gamma[i], mu = p.mean[i];
      const double s1 = __ldcg(&p.sums[i]);
      const double s2 = is * __ldcg(&p.sums[C + i]);            // sum gb * xhat
      const double k2 = -ga * is * is * s2 * inv_m;
      s_coef[3 * C + i] = (float)(ga * is);
      s_coef[4 * C + i] = (float)k2;
      s_coef[5 * C + i] = (float)(-ga * is * s1 * inv_m - k2 * mu);
      if (blockIdx.x == 0) {
        if (p.dbeta) p.dbeta[i] = (float)s1;
        if (p.dgamma) p.dgamma[i] = (float)s2;
      }
    }
  }
  __syncthreads();
  bn_stamp(5);
  drop_n = -1;                               // phase 2 walks the pixels again from the start
  auto proc2 = [&](long long q, const uint4& gmq, const uint4& yq) {
    if (gmasked) *reinterpret_cast<uint4*>(gmasked + (size_t)q * p.gmasked_ld + r.c0) = gmq;
    float gf[8], yf[8], k[8];
    unpack_bf8(gmq, gf);
    if (p.drop) {
      load_drop(q);
#pragma unroll
      for (int e = 0; e < 8; ++e) gf[e] *= dropv[e];
    }
    lds8(s_coef + 3 * C + r.c0, k);
#pragma unroll
    for (int e = 0; e < 8; ++e) gf[e] *= k[e];
    unpack_bf8(yq, yf);
    lds8(s_coef + 4 * C + r.c0, k);
#pragma unroll
    for (int e = 0; e < 8; ++e) gf[e] = fmaf(yf[e], k[e], gf[e]);
    lds8(s_coef + 5 * C + r.c0, k);
    uint4 o;
    o.x = pack_bf2(gf[0] + k[0], gf[1] + k[1]);
    o.y = pack_bf2(gf[2] + k[2], gf[3] + k[3]);
    o.z = pack_bf2(gf[4] + k[4], gf[5] + k[5]);
    o.w = pack_bf2(gf[6] + k[6], gf[7] + k[7]);
    *reinterpret_cast<uint4*>(dy + (size_t)q * p.dy_ld + r.c0) = o;
  };
#pragma unroll
  for (int k = 0; k < REGI; ++k) {
    const long long q = q0 + (long long)k * r.lanes;
    if (k < items && q < r.p1) proc2(q, cg[k], cy[k]);
  }
  {
    const int nc = ns < p.smem_items ? ns : p.smem_items;
    for (int j = 0; j < nc; ++j) {
      const long long q = q0 + (long long)(REGI + j) * r.lanes;
      if (q >= r.p1) break;
      proc2(q, s_slot[(size_t)(2 * j) * BS_THREADS + tid], s_slot[(size_t)(2 * j + 1) * BS_THREADS + tid]);
    }
  }
  for (int j = first_uncached; j < ns; ++j) {
    issue(j + BF_D - 1, MODE == BN_MASK_TENSOR);
    cp_async_wait<BF_D - 1>();
    const long long q = q0 + (long long)(REGI + j) * r.lanes;
    if (q >= r.p1) continue;
    int gs, ys, ms;
    slot_of(j, gs, ys, ms);
    uint4 gq = s_slot[(size_t)gs * BS_THREADS + tid];
    const uint4 yq = s_slot[(size_t)ys * BS_THREADS + tid];
    uint4 mq = make_uint4(0u, 0u, 0u, 0u);
    if (MODE == BN_MASK_TENSOR) mq = s_slot[(size_t)ms * BS_THREADS + tid];
    apply_mask(gq, yq, mq);
    proc2(q, gq, yq);
  }
  cp_async_wait<0>();
  bn_stamp(6);
}


// ---------------------------------------------------------------------------------------------------------------------------------
// bf16 fast path, third generation.  The per-CTA timeline of the kernel above (tools/timeline_bn.py) showed both phases bound by
// INSTRUCTION ISSUE, not by memory: ~200 SASS instructions per 16-byte item in phase 1 and ~110 in phase 2 (64-bit index arithmetic and
// a bounds check per item, coefficient vectors re-fetched from shared memory, scalar fp32 math) -- 15 + 12 us at C = 64 for 92 MB.
// Same two-phase structure and the same on-chip cache, but a short instruction stream:
//   * addresses are running pointers with one per-thread stride; a thread knows its item count up front (no per-item bounds checks);
//   * per-channel coefficients stay in registers, the math is packed fp32x2 (FFMA2 / FADD2 / FMUL2: bit-identical to the scalar forms);
//   * Dropout2d: a CTA's pixel range crosses at most one sample boundary (checked by the launcher), so the multipliers are reloaded
//     once at a precomputed item index instead of dividing per item;
//   * with a mask TENSOR the masked gradient gm is written in phase 1 (it is an output anyway), so phase 2 re-reads only (gm, y) for
//     the items beyond the cache and never touches the mask again.
// Modes served: BN_MASK_AFFINE (plain BN -> ReLU, mask recomputed from y; no dropout, no gm output) and BN_MASK_TENSOR with the gm
// output (BN2 of a block), with or without Dropout2d.  Everything else stays on the kernel above.
constexpr int BL_D = 4;   // cp.async prefetch depth (items)

__device__ __forceinline__ void unpack_bf8_2(const uint4& q, float2 (&f)[4]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
}
__device__ __forceinline__ void ld8_as_f2(const float* __restrict__ src, float2 (&f)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  f[0] = make_float2(a.x, a.y); f[1] = make_float2(a.z, a.w); f[2] = make_float2(b.x, b.y); f[3] = make_float2(b.z, b.w);
}

template <int MODE, bool DROP, int REGI>
__global__ void __launch_bounds__(BS_THREADS, 1) bn_bwd_lean_kernel(const BnBwdFusedParams p) {
  bn_stamp(0);
  pdl_enter();
  bn_stamp(1);
  static_assert(MODE == BN_MASK_AFFINE || MODE == BN_MASK_TENSOR, "modes of the lean kernel");
  extern __shared__ __align__(16) float s_dyn[];
  const int C = p.C, tid = threadIdx.x;
  float* s_coef = s_dyn;                                   // [5C]: sc, sh (mask from y) | k1, k2, k3 (after the barrier)
  float* s_part = s_dyn + 5 * C;                           // [warps][2C]
  char* s_mine = reinterpret_cast<char*>(s_part + (BS_THREADS / 32) * 2 * C) + tid * 16;   // this thread's 16 bytes of every slot
  constexpr int SLOT = BS_THREADS * 16;                    // bytes between slots
  const int c8 = C >> 3, c0 = (tid % c8) * 8, lanes = BS_THREADS / c8;
  const long long p0 = (long long)blockIdx.x * p.px_per_block;
  const long long p1 = p0 + p.px_per_block < p.n_px ? p0 + p.px_per_block : p.n_px;
  const long long q0 = p0 + tid / c8;
  const int cnt = q0 < p1 ? (int)((p1 - q0 + lanes - 1) / lanes) : 0;   // items of this thread (pixels q0 + k * lanes)
  const int ns = cnt > REGI ? cnt - REGI : 0;                            // streamed items
  const int nc = p.smem_items;                                           // streamed items [0, nc) have their own (g, y) slots

  if (MODE == BN_MASK_AFFINE) {
    for (int i = tid; i < C; i += blockDim.x) {   // the forward pass's scale / shift (bn_fwd_apply_kernel / bn_finalize_kernel), same expressions
      const float sc = p.gamma[i] * p.invstd[i];
      s_coef[i] = sc;
      s_coef[C + i] = p.beta[i] - p.mean[i] * sc;
    }
    __syncthreads();
  }
  float2 nmu[4];
  {
    ld8_as_f2(p.mean + c0, nmu);
#pragma unroll
    for (int i = 0; i < 4; ++i) nmu[i] = make_float2(-nmu[i].x, -nmu[i].y);
  }
  // running pointers (bytes) and per-item strides
  const char* gp = reinterpret_cast<const char*>(p.g) + ((size_t)q0 * p.g_ld + c0) * 2;
  const char* yp = reinterpret_cast<const char*>(p.y) + ((size_t)q0 * p.y_ld + c0) * 2;
  const char* mp = reinterpret_cast<const char*>(p.mask) + ((size_t)q0 * p.mask_ld + c0) * 2;
  char* dyp = reinterpret_cast<char*>(p.dy) + ((size_t)q0 * p.dy_ld + c0) * 2;
  char* gmp = reinterpret_cast<char*>(p.gmasked) + ((size_t)q0 * p.gmasked_ld + c0) * 2;
  const size_t gs = (size_t)lanes * p.g_ld * 2, ys = (size_t)lanes * p.y_ld * 2, ms = (size_t)lanes * p.mask_ld * 2;
  const size_t dys = (size_t)lanes * p.dy_ld * 2, gms = (size_t)lanes * p.gmasked_ld * 2;
  // Dropout2d: items [0, kb) belong to sample n0, items [kb, cnt) to sample n0 + 1 (px_per_block <= HW)
  float2 dv[4];
  int kb = 0x7fffffff;
  const float* drop_next = nullptr;
  if (DROP) {
    const long long n0 = img_of(q0 < p.n_px ? q0 : 0, p.HW);
    const long long edge = (n0 + 1) * p.HW;
    kb = (int)((edge - q0 + lanes - 1) / lanes);
    ld8_as_f2(p.drop + n0 * C + c0, dv);
    drop_next = p.drop + (n0 + 1) * C + c0;
  }
  float2 acc0[4], acc1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
  uint4 cg[REGI], cy[REGI];

  // gq <- g * (mask > 0) in place (exact in bf16); sums += (gb, gb * (y - mean)), gb = gq * drop
  auto proc1 = [&](uint4& gq, const uint4& yq, const uint4& mq) {
    float2 yf[4], gf[4];
    unpack_bf8_2(yq, yf);
    uint32_t* gw = reinterpret_cast<uint32_t*>(&gq);
    if (MODE == BN_MASK_TENSOR) {
      const uint32_t* mw = reinterpret_cast<const uint32_t*>(&mq);
      const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) gw[i] &= __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&mw[i]), zero);
    } else {
      float2 sc[4], sh[4];
      ld8_as_f2(s_coef + c0, sc);
      ld8_as_f2(s_coef + C + c0, sh);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 z = __ffma2_rn(yf[i], sc[i], sh[i]);
        const uint32_t lo = z.x > 0.f ? 0x0000FFFFu : 0u, hi = z.y > 0.f ? 0xFFFF0000u : 0u;
        gw[i] &= lo | hi;
      }
    }
    unpack_bf8_2(gq, gf);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (DROP) gf[i] = __fmul2_rn(gf[i], dv[i]);
      acc0[i] = __fadd2_rn(acc0[i], gf[i]);
      acc1[i] = __ffma2_rn(gf[i], __fadd2_rn(yf[i], nmu[i]), acc1[i]);
    }
  };
  auto next_sample = [&]() { if (DROP) ld8_as_f2(drop_next, dv); };

  // streamed item j: slots of its g / y (own slots below nc, a ring of BL_D pairs beyond) and, with a mask tensor, a ring of BL_D mask slots
  int iss = 0, iss_r = 0, iss_m = 0;                       // issue side: next item, ring positions
  const char* ig = gp + (size_t)REGI * gs;
  const char* iy = yp + (size_t)REGI * ys;
  const char* im = mp + (size_t)REGI * ms;
  auto issue1 = [&]() {
    if (iss < ns) {
      const int gsl = iss < nc ? 2 * iss : p.ring_gy + 2 * iss_r;
      cp_async16(s_mine + (size_t)gsl * SLOT, ig);
      cp_async16(s_mine + (size_t)(gsl + 1) * SLOT, iy);
      if (MODE == BN_MASK_TENSOR) cp_async16(s_mine + (size_t)(p.ring_m + iss_m) * SLOT, im);
      if (iss >= nc) iss_r = iss_r + 1 == BL_D ? 0 : iss_r + 1;
      iss_m = iss_m + 1 == BL_D ? 0 : iss_m + 1;
      ig += gs; iy += ys; im += ms;
    }
    ++iss;
    cp_async_commit();
  };

  // ---- phase 1 ----
#pragma unroll
  for (int j = 0; j < BL_D - 1; ++j) issue1();
  {
    uint4 mv[REGI];
#pragma unroll
    for (int k = 0; k < REGI; ++k) {
      if (k < cnt) {
        cg[k] = *reinterpret_cast<const uint4*>(gp + (size_t)k * gs);
        cy[k] = *reinterpret_cast<const uint4*>(yp + (size_t)k * ys);
        if (MODE == BN_MASK_TENSOR) mv[k] = *reinterpret_cast<const uint4*>(mp + (size_t)k * ms);
      }
    }
#pragma unroll
    for (int k = 0; k < REGI; ++k) {
      if (k < cnt) {
        if (DROP && k == kb) next_sample();
        proc1(cg[k], cy[k], mv[k]);
        if (MODE == BN_MASK_TENSOR) *reinterpret_cast<uint4*>(gmp + (size_t)k * gms) = cg[k];
      }
    }
  }
  {
    int con_r = 0, con_m = 0;                              // consume side ring positions
    char* og = gmp + (size_t)REGI * gms;
    for (int j = 0; j < ns; ++j) {
      issue1();
      cp_async_wait<BL_D - 1>();
      const int gsl = j < nc ? 2 * j : p.ring_gy + 2 * con_r;
      uint4 gq = *reinterpret_cast<const uint4*>(s_mine + (size_t)gsl * SLOT);
      const uint4 yq = *reinterpret_cast<const uint4*>(s_mine + (size_t)(gsl + 1) * SLOT);
      uint4 mq = make_uint4(0u, 0u, 0u, 0u);
      if (MODE == BN_MASK_TENSOR) mq = *reinterpret_cast<const uint4*>(s_mine + (size_t)(p.ring_m + con_m) * SLOT);
      if (DROP && REGI + j == kb) next_sample();
      proc1(gq, yq, mq);
      if (j < nc) *reinterpret_cast<uint4*>(s_mine + (size_t)gsl * SLOT) = gq;   // keep the masked gradient
      else con_r = con_r + 1 == BL_D ? 0 : con_r + 1;
      con_m = con_m + 1 == BL_D ? 0 : con_m + 1;
      if (MODE == BN_MASK_TENSOR) { *reinterpret_cast<uint4*>(og) = gq; og += gms; }
    }
  }
  cp_async_wait<0>();
  bn_stamp(2);
  {
    float acc[2][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[0][2 * i] = acc0[i].x; acc[0][2 * i + 1] = acc0[i].y;
      acc[1][2 * i] = acc1[i].x; acc[1][2 * i + 1] = acc1[i].y;
    }
    double* my_sums = p.sums + (size_t)(blockIdx.x % (unsigned)p.sum_copies) * 2 * C;
    block_channel_reduce<2>(acc, C, c0, s_part, [&](int i, float t) { atomicAdd(&my_sums[i], (double)t); });
  }
  bn_stamp(3);

  grid_barrier(p.counter, gridDim.x);
  bn_stamp(4);

  // ---- phase 2: dy = k1 * gb + k2 * y + k3 ----
  // items beyond the cache are read again: (gm, y) when gm was written in phase 1, (g, y) + the recomputed mask otherwise
  const char* rg = MODE == BN_MASK_TENSOR ? gmp : gp;
  const size_t rgs = MODE == BN_MASK_TENSOR ? gms : gs;
  int iss2 = nc, iss2_r = 0;
  const char* ig2 = rg + (size_t)(REGI + nc) * rgs;
  const char* iy2 = yp + (size_t)(REGI + nc) * ys;
  auto issue2 = [&]() {
    if (iss2 < ns) {
      const int gsl = p.ring_gy + 2 * iss2_r;
      cp_async16(s_mine + (size_t)gsl * SLOT, ig2);
      cp_async16(s_mine + (size_t)(gsl + 1) * SLOT, iy2);
      iss2_r = iss2_r + 1 == BL_D ? 0 : iss2_r + 1;
      ig2 += rgs; iy2 += ys;
    }
    ++iss2;
    cp_async_commit();
  };
#pragma unroll
  for (int j = 0; j < BL_D - 1; ++j) issue2();   // in flight while the cached items are processed
  {
    const double inv_m = 1.0 / (double)p.n_px;
    for (int i = tid; i < C; i += blockDim.x) {
      const double is = p.invstd[i], ga = p.gamma[i], mu = p.mean[i];
      double s1 = 0.0, s2 = 0.0;
      for (int k = 0; k < p.sum_copies; ++k) {
        s1 += __ldcg(&p.sums[(size_t)k * 2 * C + i]);
        s2 += __ldcg(&p.sums[(size_t)k * 2 * C + C + i]);
      }
      s2 *= is;                                                 // sum gb * xhat
      const double k2 = -ga * is * is * s2 * inv_m;
      s_coef[2 * C + i] = (float)(ga * is);
      s_coef[3 * C + i] = (float)k2;
      s_coef[4 * C + i] = (float)(-ga * is * s1 * inv_m - k2 * mu);
      if (blockIdx.x == 0) {
        if (p.dbeta) p.dbeta[i] = (float)s1;
        if (p.dgamma) p.dgamma[i] = (float)s2;
      }
    }
  }
  __syncthreads();
  bn_stamp(5);
  float2 k1[4], k2[4], k3[4], k1d[4];
  ld8_as_f2(s_coef + 2 * C + c0, k1);
  ld8_as_f2(s_coef + 3 * C + c0, k2);
  ld8_as_f2(s_coef + 4 * C + c0, k3);
  if (DROP) {
    ld8_as_f2(drop_next - C, dv);              // back to the first sample
#pragma unroll
    for (int i = 0; i < 4; ++i) k1d[i] = __fmul2_rn(k1[i], dv[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) k1d[i] = k1[i];
  }
  auto next_sample2 = [&]() {
    if (DROP) {
      ld8_as_f2(drop_next, dv);
#pragma unroll
      for (int i = 0; i < 4; ++i) k1d[i] = __fmul2_rn(k1[i], dv[i]);
    }
  };
  auto proc2 = [&](char* dst, const uint4& gmq, const uint4& yq) {
    float2 gf[4], yf[4];
    unpack_bf8_2(gmq, gf);
    unpack_bf8_2(yq, yf);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 r2 = __ffma2_rn(gf[i], k1d[i], __ffma2_rn(yf[i], k2[i], k3[i]));
      ow[i] = pack_bf2(r2.x, r2.y);
    }
    *reinterpret_cast<uint4*>(dst) = o;
  };
#pragma unroll
  for (int k = 0; k < REGI; ++k) {
    if (k < cnt) {
      if (DROP && k == kb) next_sample2();
      proc2(dyp + (size_t)k * dys, cg[k], cy[k]);
    }
  }
  char* od = dyp + (size_t)REGI * dys;
  {
    const int ncc = ns < nc ? ns : nc;
    for (int j = 0; j < ncc; ++j) {
      if (DROP && REGI + j == kb) next_sample2();
      proc2(od, *reinterpret_cast<const uint4*>(s_mine + (size_t)(2 * j) * SLOT), *reinterpret_cast<const uint4*>(s_mine + (size_t)(2 * j + 1) * SLOT));
      od += dys;
    }
  }
  {
    int con_r = 0;
    for (int j = nc; j < ns; ++j) {
      issue2();
      cp_async_wait<BL_D - 1>();
      const int gsl = p.ring_gy + 2 * con_r;
      uint4 gq = *reinterpret_cast<const uint4*>(s_mine + (size_t)gsl * SLOT);
      const uint4 yq = *reinterpret_cast<const uint4*>(s_mine + (size_t)(gsl + 1) * SLOT);
      con_r = con_r + 1 == BL_D ? 0 : con_r + 1;
      if (MODE == BN_MASK_AFFINE) {            // redo the mask
        float2 yf[4], sc[4], sh[4];
        unpack_bf8_2(yq, yf);
        ld8_as_f2(s_coef + c0, sc);
        ld8_as_f2(s_coef + C + c0, sh);
        uint32_t* gw = reinterpret_cast<uint32_t*>(&gq);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 z = __ffma2_rn(yf[i], sc[i], sh[i]);
          const uint32_t lo = z.x > 0.f ? 0x0000FFFFu : 0u, hi = z.y > 0.f ? 0xFFFF0000u : 0u;
          gw[i] &= lo | hi;
        }
      }
      if (DROP && REGI + j == kb) next_sample2();
      proc2(od, gq, yq);
      od += dys;
    }
  }
  cp_async_wait<0>();
  bn_stamp(6);
}

template <int MODE, bool DROP>
static int launch_bn_bwd_lean(BnBwdFusedParams p, cudaStream_t stream) {
  constexpr int REGI = 6;
  const BsLaunch l = bs_launch(p.C, p.n_px, 1);
  if (DROP && l.px_per_block > p.HW) return 1;      // a thread's items must cross at most one sample boundary
  const int lanes = BS_THREADS / (p.C / 8);
  const int items = l.px_per_block / lanes;
  const int ns = items > REGI ? items - REGI : 0;
  const size_t fixed = (size_t)(5 * p.C + (BS_THREADS / 32) * 2 * p.C) * sizeof(float);
  const size_t slot_bytes = (size_t)BS_THREADS * 16;
  const size_t budget = 220 * 1024;
  const int slots = (int)((budget - fixed) / slot_bytes);
  const int mask_slots = MODE == BN_MASK_TENSOR ? BL_D : 0;
  int cache_items = ns;
  if (2 * cache_items + mask_slots > slots) {       // not everything fits: the rest goes through a (g, y) ring and is read again in phase 2
    cache_items = (slots - mask_slots - 2 * BL_D) / 2;
    if (cache_items < 0) return 1;
  }
  const bool ring = cache_items < ns;
  p.smem_items = cache_items;
  p.ring_gy = 2 * cache_items;
  p.ring_m = 2 * cache_items + (ring ? 2 * BL_D : 0);
  const int used = p.ring_m + mask_slots;
  const size_t smem = fixed + (size_t)used * slot_bytes;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  auto kern = bn_bwd_lean_kernel<MODE, DROP, REGI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) return -1;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BS_THREADS, smem) != cudaSuccess) return -1;
  if (occ < 1 || (long long)l.grid > (long long)occ * sms) return 1;
  p.px_per_block = l.px_per_block;
  launch_pdl(bn_bwd_lean_kernel<MODE, DROP, REGI>, dim3(l.grid), dim3(BS_THREADS), smem, stream, p);
  return 0;
}

constexpr int BF_REGI = 6;

template <int MODE>
static int launch_bn_bwd_fused_bf16(BnBwdFusedParams p, cudaStream_t stream) {
  const BsLaunch l = bs_launch(p.C, p.n_px, 1);
  const int lanes = BS_THREADS / (p.C / 8);
  const int items = l.px_per_block / lanes;
  const int ns = items > BF_REGI ? items - BF_REGI : 0;
  const size_t fixed = (size_t)(6 * p.C + (BS_THREADS / 32) * 2 * p.C) * sizeof(float);
  const size_t slot_bytes = (size_t)BS_THREADS * 16;
  const size_t budget = 220 * 1024;
  int slots = (int)((budget - fixed) / slot_bytes);
  const int mask_slots = MODE == BN_MASK_TENSOR ? BF_D : 0;
  int cache_items = ns;
  if (2 * cache_items + mask_slots > slots) {       // not everything fits: the rest goes through a (g, y) ring and is read again in phase 2
    cache_items = (slots - mask_slots - 2 * BF_D) / 2;
    if (cache_items < 0) return 1;
  }
  const bool ring = cache_items < ns;
  p.smem_items = cache_items;
  p.ring_gy = 2 * cache_items;
  p.ring_m = 2 * cache_items + (ring ? 2 * BF_D : 0);
  const int used = p.ring_m + mask_slots;
  const size_t smem = fixed + (size_t)used * slot_bytes;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  if (cudaFuncSetAttribute(bn_bwd_fused_bf16_kernel<MODE, BF_REGI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) return -1;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_bf16_kernel<MODE, BF_REGI>, BS_THREADS, smem) != cudaSuccess) return -1;
  if (occ < 1 || (long long)l.grid > (long long)occ * sms) return 1;
  p.px_per_block = l.px_per_block;
  launch_pdl(bn_bwd_fused_bf16_kernel<MODE, BF_REGI>, dim3(l.grid), dim3(BS_THREADS), smem, stream, p);
  return 0;
}

template <typename T> struct RegItems;
template <> struct RegItems<__nv_bfloat16> { static constexpr int value = 6; };
template <> struct RegItems<float> { static constexpr int value = 3; };

template <typename T, int SUNR>
static int launch_bn_bwd_fused(BnBwdFusedParams p, cudaStream_t stream) {
  constexpr int REGI = RegItems<T>::value;
  const BsLaunch l = bs_launch(p.C, p.n_px, 1);    // ONE 512-thread CTA per SM with 128 registers per thread: more items stay on chip
  const int lanes = BS_THREADS / (p.C / 8);
  const int items = l.px_per_block / lanes;
  const size_t fixed = (size_t)(4 * p.C + (BS_THREADS / 32) * 2 * p.C) * sizeof(float);
  const size_t per_item = (size_t)2 * BS_THREADS * sizeof(Raw8<T>);
  const size_t budget = 220 * 1024;   // one CTA per SM
  int smem_items = items > REGI ? items - REGI : 0;
  const int fit = fixed < budget ? (int)((budget - fixed) / per_item) : 0;
  if (smem_items > fit) smem_items = fit;
  const size_t smem = fixed + (size_t)smem_items * per_item;
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  auto kern = bn_bwd_fused_kernel<T, REGI, SUNR>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess) return -1;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BS_THREADS, smem) != cudaSuccess) return -1;
  if (occ < 1 || (long long)l.grid > (long long)occ * sms) return 1;   // the grid would not be co-resident: caller falls back to two launches
  p.px_per_block = l.px_per_block;
  p.smem_items = smem_items;
  launch_pdl(bn_bwd_fused_kernel<T, REGI, SUNR>, dim3(l.grid), dim3(BS_THREADS), smem, stream, p);
  return 0;
}

}  // namespace erfk

using namespace erfk;

extern "C" int erf_bn_set_debug(unsigned long long* buf) {   // tools only: [grid][8] globaltimer stamps per launch (nullptr switches it off)
  return cudaMemcpyToSymbol(g_bn_dbg, &buf, sizeof(buf)) == cudaSuccess ? ERF_OK : ERF_ERR_CUDA;
}

static bool fused_disabled() {
  static const bool off = getenv("ERF_BN_FUSED") && atoi(getenv("ERF_BN_FUSED")) == 0;
  return off;
}

extern "C" int erf_bn_fused_supported(int C) { return !fused_disabled() && bs_fast_ok(C) && C <= 256; }

extern "C" int erf_bn_fwd_apply(const void* x, int x_ld, const double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, float momentum, float eps, const float* drop, const void* res, int res_ld, void* out, int out_ld,
                                int N, int HW, int C, int relu, float* mean, float* invstd, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_bn_fwd_apply: bad dtype %d", dtype);
  ERF_CHECK_ARG(x && stats && gamma && beta && out && mean && invstd && N > 0 && HW > 0 && count > 0, "erf_bn_fwd_apply: bad arguments");
  ERF_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "erf_bn_fwd_apply: running_mean/var must be given together");
  if (!(bs_fast_ok(C) && C <= 256 && (x_ld % 8) == 0 && (out_ld % 8) == 0 && (!res || (res_ld % 8) == 0) && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
        ((uintptr_t)res % 16) == 0)) {
    set_error("erf_bn_fwd_apply: unsupported layout (C=%d must be 8..256 with C/8 dividing 32, 16-byte aligned rows)", C);
    return ERF_ERR_UNSUPPORTED;
  }
  const long long n_px = (long long)N * HW;
  const BsLaunch l = bs_launch(C, n_px, 4);
  const size_t smem = 2 * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define ERF_BN_FWD(TT, DR)                                                                                                                                    \
  launch_pdl(bn_fwd_apply_kernel<TT, DR>, dim3(l.grid), dim3(BS_THREADS), smem, st, (const TT*)x, x_ld, stats, count, gamma, beta, running_mean, running_var, \
             momentum, eps, drop, (const TT*)res, res_ld, (TT*)out, out_ld, (long long)HW, C, n_px, l.px_per_block, relu, mean, invstd)
  if (dtype == ERF_F32) {
    if (drop) ERF_BN_FWD(float, true); else ERF_BN_FWD(float, false);
  } else {
    if (drop) ERF_BN_FWD(__nv_bfloat16, true); else ERF_BN_FWD(__nv_bfloat16, false);
  }
#undef ERF_BN_FWD
  ERF_CHECK_LAUNCH("erf_bn_fwd_apply");
  return ERF_OK;
}

extern "C" int erf_bn_bwd_fused(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld, const float* beta_for_mask,
                                const float* drop, const float* gamma, const float* mean, const float* invstd, double* sums, unsigned* counter, int N,
                                int HW, int C, void* dy, int dy_ld, float* dgamma, float* dbeta, void* gmasked, int gmasked_ld, int dtype,
                                erf_stream_t stream) {
  return erf_bn_bwd_fused_spread(g, g_ld, y, y_ld, mask, mask_ld, beta_for_mask, drop, gamma, mean, invstd, sums, 1, counter, N, HW, C, dy, dy_ld, dgamma, dbeta,
                                 gmasked, gmasked_ld, dtype, stream);
}

extern "C" int erf_bn_bwd_fused_spread(const void* g, int g_ld, const void* y, int y_ld, const void* mask, int mask_ld, const float* beta_for_mask,
                                       const float* drop, const float* gamma, const float* mean, const float* invstd, double* sums, int sum_copies,
                                       unsigned* counter, int N, int HW, int C, void* dy, int dy_ld, float* dgamma, float* dbeta, void* gmasked,
                                       int gmasked_ld, int dtype, erf_stream_t stream) {
  ERF_CHECK_ARG(sum_copies >= 1 && sum_copies <= 64, "erf_bn_bwd_fused: sum_copies=%d must be in [1, 64]", sum_copies);
  ERF_CHECK_ARG(dtype == ERF_F32 || dtype == ERF_BF16, "erf_bn_bwd_fused: bad dtype %d", dtype);
  ERF_CHECK_ARG(g && y && gamma && mean && invstd && sums && counter && dy && N > 0 && HW > 0, "erf_bn_bwd_fused: bad arguments");
  ERF_CHECK_ARG(!beta_for_mask || (mask && !drop), "erf_bn_bwd_fused: the mask can only be recomputed from y for a plain BN -> ReLU (mask tensor still required as the fallback operand, no dropout)");
  const bool ok = !fused_disabled() && bs_fast_ok(C) && C <= 256 && (g_ld % 8) == 0 && (y_ld % 8) == 0 && (dy_ld % 8) == 0 && (!mask || (mask_ld % 8) == 0) &&
                  (!gmasked || (gmasked_ld % 8) == 0) && ((uintptr_t)g % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)mask % 16) == 0 &&
                  ((uintptr_t)dy % 16) == 0 && ((uintptr_t)gmasked % 16) == 0 && ((uintptr_t)drop % 16) == 0;
  if (ok) {
    BnBwdFusedParams p;
    p.g = g; p.y = y; p.mask = mask;
    p.g_ld = g_ld; p.y_ld = y_ld; p.mask_ld = mask_ld;
    p.drop = drop; p.gamma = gamma; p.beta = beta_for_mask; p.mean = mean; p.invstd = invstd;
    p.ring_gy = p.ring_m = 0;
    p.HW = HW; p.n_px = (long long)N * HW;
    p.C = C; p.px_per_block = 0; p.smem_items = 0;
    p.sums = sums; p.sum_copies = sum_copies; p.counter = counter;   // only the lean kernel spreads; the others use copy 0
    p.dy = dy; p.dy_ld = dy_ld; p.dgamma = dgamma; p.dbeta = dbeta; p.gmasked = gmasked; p.gmasked_ld = gmasked_ld;
    static const int v1 = getenv("ERF_BN_V1") ? atoi(getenv("ERF_BN_V1")) : 0;   // 1: the register / LDG kernel also for bf16 (experiments)
    const char* lean_env = getenv("ERF_BN_LEAN");   // read per call: tests and tools switch kernels inside one process
    const bool lean_ok = !(lean_env && atoi(lean_env) == 0) && ((uintptr_t)mean % 16) == 0 && ((uintptr_t)drop % 16) == 0 && (C % 8) == 0;
    int rc;
    if (dtype == ERF_F32)
      rc = launch_bn_bwd_fused<float, 2>(p, (cudaStream_t)stream);
    else if (v1)
      rc = launch_bn_bwd_fused<__nv_bfloat16, 2>(p, (cudaStream_t)stream);
    else {
      rc = 1;   // 1 = this kernel does not take the case, try the next one
      if (lean_ok && mask && beta_for_mask && !gmasked && !drop)
        rc = launch_bn_bwd_lean<BN_MASK_AFFINE, false>(p, (cudaStream_t)stream);
      else if (lean_ok && mask && !beta_for_mask && gmasked)
        rc = drop ? launch_bn_bwd_lean<BN_MASK_TENSOR, true>(p, (cudaStream_t)stream) : launch_bn_bwd_lean<BN_MASK_TENSOR, false>(p, (cudaStream_t)stream);
      if (rc == 1) {
        if (!mask)
          rc = launch_bn_bwd_fused_bf16<BN_MASK_NONE>(p, (cudaStream_t)stream);
        else if (beta_for_mask)
          rc = launch_bn_bwd_fused_bf16<BN_MASK_AFFINE>(p, (cudaStream_t)stream);
        else
          rc = launch_bn_bwd_fused_bf16<BN_MASK_TENSOR>(p, (cudaStream_t)stream);
      }
    }
    if (rc == 0) {
      ERF_CHECK_LAUNCH("erf_bn_bwd_fused");
      return ERF_OK;
    }
    if (rc < 0) {
      set_error("erf_bn_bwd_fused: kernel attribute / occupancy query failed: %s", cudaGetErrorString(cudaGetLastError()));
      return ERF_ERR_CUDA;
    }
  }
  // not co-resident / unsupported layout: the two-launch form (same arithmetic)
  int rc = erf_bn_bwd_reduce(g, g_ld, y, y_ld, mask, mask_ld, drop, mean, invstd, N, HW, C, sums, dtype, stream);
  if (rc != ERF_OK) return rc;
  return erf_bn_bwd_apply(g, g_ld, y, y_ld, mask, mask_ld, drop, gamma, mean, invstd, sums, N, HW, C, dy, dy_ld, dgamma, dbeta, gmasked, gmasked_ld, dtype, stream);
}
