// Bring-up self test of the CTA-pair (cta_group::2) tensor-core path:
// D[256 x N] = A[256 x K] * B[N x K]^T with the bf16x3 split; CTA r of the pair owns rows
// 128r..128r+127 of A and D and rows (N/2)r.. of B, exactly as the paired decode kernels do.
#include "common.cuh"
#include "umma.cuh"

namespace ddmi {
namespace ummak2 {
using namespace umma;

constexpr int KG_BYTES = 128 * 16;
constexpr int ST2_A = 32 * KG_BYTES;                    // A hi (K <= 256), then A lo
constexpr int ST2_OFF_B = 2 * ST2_A;                    // one K step of the local B half: [hi | lo] <= 8 KB
constexpr int ST2_OFF_BAR = ST2_OFF_B + 8192;
constexpr int ST2_SMEM = ST2_OFF_BAR + 64;
// barriers: +0 ready (leader: both CTAs' operands written, count 2), +8 done (multicast commit), +16 tmem slot

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1)
selftest2_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_hi = sbase, a_lo = sbase + ST2_A, bst = sbase + ST2_OFF_B, bar = sbase + ST2_OFF_BAR;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int NH = N / 2;
  if (tid == 0) {
    mbar_init(bar, 2);
    mbar_init(bar + 8, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc2(bar + 16, 256);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + ST2_OFF_BAR + 16);
  const uint32_t ready_leader = mapa_rank(bar, 0);
  if (tid < 128) {   // this CTA's 128 rows of A
    const size_t grow = (size_t)rank * 128 + tid;
    for (int g = 0; g < K / 8; ++g) {
      float y[8];
      for (int i = 0; i < 8; ++i) y[i] = a[grow * K + g * 8 + i];
      uint4 hi, lo;
      split8(y, hi, lo);
      st_shared_v4(a_hi + g * KG_BYTES + tid * 16, hi);
      st_shared_v4(a_lo + g * KG_BYTES + tid * 16, lo);
    }
  }
  const uint32_t idesc = idesc2_bf16_f32(N);
  uint32_t ph = 0;
  for (int j = 0; j < K / 16; ++j) {
    if (tid < 128) {   // this CTA's N/2 rows of B for K step j
      for (int r = tid; r < NH; r += 128) {
        for (int g = 0; g < 2; ++g) {
          float y[8];
          for (int i = 0; i < 8; ++i) y[i] = b[((size_t)rank * NH + r) * K + j * 16 + g * 8 + i];
          uint4 hi, lo;
          split8(y, hi, lo);
          st_shared_v4(bst + (g * NH + r) * 16, hi);
          st_shared_v4(bst + NH * 32 + (g * NH + r) * 16, lo);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) mbar_arrive_cluster(ready_leader);          // one arrival per CTA on the leader's barrier
    if (rank == 0 && warp == 4) {
      mbar_wait_cluster(bar, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bhi = smem_desc(bst, NH * 16, 128), blo = smem_desc(bst + NH * 32, NH * 16, 128);
        const uint64_t ahi = smem_desc(a_hi + j * 2 * KG_BYTES, KG_BYTES, 128);
        const uint64_t alo = smem_desc(a_lo + j * 2 * KG_BYTES, KG_BYTES, 128);
        mma2_bf16(tmem, ahi, bhi, idesc, j > 0 ? 1u : 0u);
        mma2_bf16(tmem, alo, bhi, idesc, 1u);
        mma2_bf16(tmem, ahi, blo, idesc, 1u);
        mma2_commit_mc(bar + 8, 3);
      }
      __syncwarp();
    }
    mbar_wait(bar + 8, ph);   // both CTAs: the MMAs reading this K step's operands are done
    ph ^= 1;
    tc_fence_after();
  }
  if (tid < 128) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      float2 v[8];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) {
        d[((size_t)rank * 128 + tid) * N + c0 + 2 * i] = v[i].x;
        d[((size_t)rank * 128 + tid) * N + c0 + 2 * i + 1] = v[i].y;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 4) tmem_dealloc2(tmem, 256);
}
}  // namespace ummak2

int launch_selftest_umma2(const float* a, const float* b, float* d, int N, int K, cudaStream_t st) {
  using namespace ummak2;
  DDMI_CUDA(cudaFuncSetAttribute(selftest2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST2_SMEM));
  selftest2_kernel<<<2, 160, ST2_SMEM, st>>>(a, b, d, N, K);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}
}  // namespace ddmi
