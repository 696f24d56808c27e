// Bring-up self test of the f16f8 operand scheme (umma.cuh): D[128 x N] = A[128 x K] * B[N x K]^T computed as
// one fp16 kind::f16 term + two kind::f8f6f4 correction terms into one fp32 TMEM accumulator that holds 4096 * D.
// Uses the same operand layouts, descriptors and split as the decode kernels: per K = 32 step
//   A: [a16: 4 K groups | r8: 2 | a8: 2] of 128 rows x 16 B     B: [w16: 4 K groups | w8: 2 | s8: 2] of N rows x 16 B
#include "common.cuh"
#include "umma.cuh"

namespace ddmi {
namespace ummak {
using namespace umma;

constexpr int SF_KG = 128 * 16;                         // one K group of A
constexpr int SF_OFF_A = 0, SF_A_BYTES = 8 * 8 * SF_KG; // K <= 256: 8 steps x 8 K groups
constexpr int SF_OFF_B = SF_A_BYTES, SF_B_BYTES = 256 * 128;
constexpr int SF_OFF_BAR = SF_OFF_B + SF_B_BYTES, SF_SMEM = SF_OFF_BAR + 64;

__device__ __forceinline__ uint4 pack_f16x8(const float* v) {
  uint32_t w[4];
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 pack_e4m3x16(const float* v) {
  uint32_t w[4];
  for (int i = 0; i < 4; ++i) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * i], v[4 * i + 1]), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * i + 2], v[4 * i + 3]), __NV_SATFINITE, __NV_E4M3);
    w[i] = lo | (hi << 16);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(160, 1)
selftest_f16f8_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t abuf = sbase + SF_OFF_A, bst = sbase + SF_OFF_B, bar = sbase + SF_OFF_BAR;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(bar + 8, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SF_OFF_BAR + 8);
  if (tid < 128) {
    for (int s = 0; s < K / 32; ++s) {
      float y[32];
      for (int i = 0; i < 32; ++i) y[i] = a[(size_t)tid * K + s * 32 + i];
      uint4 a16[4], r8[2], a8[2];
      split32_f16f8(y, a16, r8, a8);
      const uint32_t base = abuf + s * 8 * SF_KG + tid * 16;
      for (int g = 0; g < 4; ++g) st_shared_v4(base + g * SF_KG, a16[g]);
      for (int g = 0; g < 2; ++g) {
        st_shared_v4(base + (4 + g) * SF_KG, r8[g]);
        st_shared_v4(base + (6 + g) * SF_KG, a8[g]);
      }
    }
  }
  const uint32_t id16 = idesc_f16_f32(128, N), id8 = idesc_f8_f32(128, N);
  uint32_t ph = 0;
  for (int s = 0; s < K / 32; ++s) {
    if (tid < 128) {
      for (int r = tid; r < N; r += 128) {
        float w[32], ws[32], sr[32];
        for (int i = 0; i < 32; ++i) {
          w[i] = b[(size_t)r * K + s * 32 + i];
          ws[i] = w[i] * kF8Scale;
          const __half h = __float2half_rn(ws[i]);
          sr[i] = ws[i] - __half2float(h);
        }
        uint4 w16[4], w8[2], s8[2];
        for (int g = 0; g < 4; ++g) w16[g] = pack_f16x8(ws + 8 * g);     // fp16(S w)
        for (int g = 0; g < 2; ++g) {
          w8[g] = pack_e4m3x16(w + 16 * g);                              // e4m3(w)
          s8[g] = pack_e4m3x16(sr + 16 * g);                             // e4m3(S w - fp16(S w))
        }
        for (int g = 0; g < 4; ++g) st_shared_v4(bst + (g * N + r) * 16, w16[g]);
        for (int g = 0; g < 2; ++g) {
          st_shared_v4(bst + ((4 + g) * N + r) * 16, w8[g]);
          st_shared_v4(bst + ((6 + g) * N + r) * 16, s8[g]);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 128) {
      tc_fence_after();
      const uint32_t ab = abuf + s * 8 * SF_KG;
      mma_bf16(tmem, smem_desc(ab, SF_KG, 128), smem_desc(bst, N * 16, 128), id16, s > 0 ? 1u : 0u);
      mma_bf16(tmem, smem_desc(ab + 2 * SF_KG, SF_KG, 128), smem_desc(bst + 2 * N * 16, N * 16, 128), id16, 1u);
      mma_f8(tmem, smem_desc(ab + 4 * SF_KG, SF_KG, 128), smem_desc(bst + 4 * N * 16, N * 16, 128), id8, 1u);
      mma_f8(tmem, smem_desc(ab + 6 * SF_KG, SF_KG, 128), smem_desc(bst + 6 * N * 16, N * 16, 128), id8, 1u);
      mma_commit(bar);
    }
    mbar_wait(bar, ph);
    ph ^= 1;
    tc_fence_after();
  }
  if (tid < 128) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32 && c0 + i < N; ++i) d[(size_t)tid * N + c0 + i] = v[i] * kF8InvScale;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace ummak

int launch_selftest_f16f8(const float* a, const float* b, float* d, int N, int K, cudaStream_t st) {
  using namespace ummak;
  DDMI_CUDA(cudaFuncSetAttribute(selftest_f16f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM));
  selftest_f16f8_kernel<<<1, 160, SF_SMEM, st>>>(a, b, d, N, K);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
