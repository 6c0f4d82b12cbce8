// tcgen05 probe: one CTA computes D[128 x N] = A[row_shift .. row_shift+127, :K] * B[N, :K]^T with the A operand
// descriptor's start address shifted by `row_shift` rows.  The fused conv kernels implement a convolution tap as
// exactly such a row-shifted view of one shared-memory activation tile, so tests/test_gpu_umma.py uses this to pin
// the descriptor semantics (no-swizzle and 128B-swizzle K-major layouts) against a CPU matmul, and to time the
// MMA issue rate of both layouts.
#include "common.cuh"
#include "tc.cuh"

namespace erfk {
using namespace tc;

__global__ void __launch_bounds__(128) umma_probe_kernel(const __nv_bfloat16* __restrict__ a, int a_rows, const __nv_bfloat16* __restrict__ b,
                                                         int n, int ksteps, int row_shift, int layout, float* __restrict__ d, int iters,
                                                         long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int K = ksteps * 16;
  const int kchunks = K / 8;  // 16-byte chunks per row
  // operand footprints
  const uint32_t a_bytes = layout == LAYOUT_SW128 ? a_rows * 128 : a_rows * 16 * kchunks;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 1023) & ~1023u);
  const int tid = threadIdx.x;
  // ---- stage operands with plain stores in the chosen canonical layout
  for (int i = tid; i < a_rows * kchunks; i += blockDim.x) {
    int r = i / kchunks, c = i % kchunks;
    uint4 v = *reinterpret_cast<const uint4*>(a + (size_t)r * K + c * 8);
    uint32_t off = layout == LAYOUT_SW128 ? r * 128 + ((c ^ (r & 7)) * 16) : c * (a_rows * 16) + r * 16;
    *reinterpret_cast<uint4*>(sa + off) = v;
  }
  for (int i = tid; i < n * kchunks; i += blockDim.x) {
    int r = i / kchunks, c = i % kchunks;
    uint4 v = *reinterpret_cast<const uint4*>(b + (size_t)r * K + c * 8);
    uint32_t off = layout == LAYOUT_SW128 ? r * 128 + ((c ^ (r & 7)) * 16) : c * (n * 16) + r * 16;
    *reinterpret_cast<uint4*>(sb + off) = v;
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  uint32_t cols = 32;
  while ((int)cols < n) cols <<= 1;
  __syncthreads();
  if (tid < 32) tmem_alloc(&tmem_base_s, cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_bf16(128, n);
  long long t0 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < ksteps; ++k) {
        uint64_t da, db;
        if (layout == LAYOUT_SW128) {
          da = make_smem_desc(smem_u32(sa) + row_shift * 128 + k * 32, 16, 1024, LAYOUT_SW128);
          db = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024, LAYOUT_SW128);
        } else {
          da = make_smem_desc(smem_u32(sa) + row_shift * 16 + k * 2 * (a_rows * 16), a_rows * 16, 128, LAYOUT_NONE);
          db = make_smem_desc(smem_u32(sb) + k * 2 * (n * 16), n * 16, 128, LAYOUT_NONE);
        }
        umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tcgen05_fence_after();
  if (tid == 0 && cycles) *cycles = clock64() - t0;
  const int warp = tid >> 5, lane = tid & 31;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) d[(size_t)row * n + c0 + j] = __uint_as_float(r[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, cols);
}

}  // namespace erfk

extern "C" int erf_umma_probe(const void* a, int a_rows, const void* b, int n, int ksteps, int row_shift, int layout, float* d, int iters,
                              long long* cycles, erf_stream_t stream) {
  using namespace erfk;
  ERF_CHECK_ARG(a && b && d, "erf_umma_probe: null pointer");
  ERF_CHECK_ARG(layout == tc::LAYOUT_NONE || layout == tc::LAYOUT_SW128, "erf_umma_probe: layout must be 0 or 2");
  ERF_CHECK_ARG(n >= 16 && n <= 256 && n % 16 == 0, "erf_umma_probe: N=%d must be a multiple of 16 in [16,256]", n);
  ERF_CHECK_ARG(ksteps >= 1 && (layout != tc::LAYOUT_SW128 || ksteps <= 4), "erf_umma_probe: ksteps out of range");
  ERF_CHECK_ARG(row_shift >= 0 && a_rows >= 128 + row_shift && iters >= 1, "erf_umma_probe: need a_rows >= 128 + row_shift");
  const int K = ksteps * 16;
  size_t a_bytes = layout == tc::LAYOUT_SW128 ? (size_t)a_rows * 128 : (size_t)a_rows * 2 * K;
  size_t b_bytes = layout == tc::LAYOUT_SW128 ? (size_t)n * 128 : (size_t)n * 2 * K;
  size_t smem = ((a_bytes + 1023) & ~(size_t)1023) + b_bytes + 2048;
  ERF_CHECK_ARG(smem <= 200 * 1024, "erf_umma_probe: operands too large for shared memory");
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("erf_umma_probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return ERF_ERR_CUDA;
  }
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, a_rows, (const __nv_bfloat16*)b, n, ksteps, row_shift, layout, d,
                                                            iters, cycles);
  ERF_CHECK_LAUNCH("erf_umma_probe");
  return ERF_OK;
}
