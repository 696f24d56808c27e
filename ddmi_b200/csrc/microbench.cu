// Diagnostics: micro-benchmarks of the epilogue-side building blocks of the tcgen05 engine (ddmi_debug_microbench).
// One CTA, 8 "E" warps laid out as in the decode kernels (warp w owns TMEM lanes 32 * (w % 4)..+31, sub = w / 4 picks 32 of
// each 64-column quarter) + one warp that allocates tensor memory.  Each mode repeats one stage-sized unit of work (the
// thread's 128 accumulator values) `iters` times and reports the average cycles per repetition of warp 0 and the CTA-wide
// span, so the cost model of an epilogue stage is measured, not guessed.
//   mode 0: drain      4 x tcgen05.ld.32x32b.x32 + wait::ld, 8 warps
//   mode 1: drain      same, 4 warps only (one per lane quadrant)
//   mode 2: convert    f16f8 split of 128 values per thread (registers only)
//   mode 3: publish    32 x st.shared.v4 per thread in the A-operand layout (+ fence.proxy.async)
//   mode 4: stage      drain + bias/lrelu + split + publish (one full image epilogue stage without the hand-shakes)
//   mode 5: tmem store 4 x tcgen05.st.32x32b.x16 x 2 (128 values) + wait::st
//   mode 6: convert    bf16 hi/lo split of 128 values per thread
//   mode 7: publish    as mode 3 without the fence.proxy.async after each quarter
//   mode 8: drain      one tcgen05.ld.32x32b.x32 + wait::ld at a time (4 dependent round trips)
#include <string.h>
#include "common.cuh"
#include "umma.cuh"
#include "tma.cuh"

namespace ddmi {
namespace mbench {
using namespace umma;

constexpr int H_KG = 32, KG_BYTES = 128 * 16;            // as in umma_engine.cuh
template <int NP>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float2 (&b)[NP]) {
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
    b[2 * i] = make_float2(v.x, v.y);
    b[2 * i + 1] = make_float2(v.z, v.w);
  }
}

constexpr int MB_OFF_BAR = 2 * H_KG * KG_BYTES;          // [H hi | H lo] then the TMEM slot
constexpr int MB_SMEM = MB_OFF_BAR + 64;

__global__ void __launch_bounds__(288, 1)
microbench_kernel(int mode, int iters, const float* __restrict__ seed, unsigned long long* __restrict__ out, float* __restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase, h_lo = sbase + H_KG * KG_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5;
  __shared__ long long t_first, t_last;
  if (tid == 0) { t_first = 0x7fffffffffffffffLL; t_last = 0; }
  if (warp == 8) tmem_alloc(sbase + MB_OFF_BAR, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + MB_OFF_BAR);
  if (warp < 8) {
    const int row = tid & 127, sub = warp >> 2;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    float2 v[4][16];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 16; ++i) v[q][i] = make_float2(__ldg(seed + ((tid * 7 + q * 16 + i) & 1023)), __ldg(seed + ((tid * 5 + q + i * 3) & 1023)));
    // give the accumulator defined contents
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 t[8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = v[q][c * 8 + i];
        tmem_st16(tmem_lane + q * 64 + sub * 32 + c * 16, t);
      }
    }
    tmem_st_wait();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const bool active = !(mode == 1 && warp >= 4);
    const long long t0 = clock64();
    float acc = 0.f;
    if (active) {
      for (int it = 0; it < iters; ++it) {
        if (mode == 0 || mode == 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_ld32(tmem_lane + q * 64 + sub * 32, v[q]);
          tmem_ld_wait();
          acc += v[0][0].x;
        } else if (mode == 2 || mode == 6) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (mode == 2) {
                uint4 a16[2], r8, a8;
                split16_f16f8(&v[q][c * 8], a16, r8, a8);
                v[q][c * 8].x += __uint_as_float((a16[0].x ^ a16[1].y ^ r8.z ^ a8.w) & 0x007fffffu) * 1e-30f;
              } else {
                uint4 hi[2], lo[2];
                split8(&v[q][c * 8], hi[0], lo[0]);
                split8(&v[q][c * 8 + 4], hi[1], lo[1]);
                v[q][c * 8].x += __uint_as_float((hi[0].x ^ hi[1].y ^ lo[0].z ^ lo[1].w) & 0x007fffffu) * 1e-30f;
              }
            }
          }
        } else if (mode == 8) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            tmem_ld32(tmem_lane + q * 64 + sub * 32, v[q]);
            tmem_ld_wait();
          }
          acc += v[0][0].x;
        } else if (mode == 3 || mode == 7) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col0 = q * 64 + sub * 32;
            const uint32_t a = h_hi + (col0 / 8) * KG_BYTES + row * 16, b = h_lo + (col0 / 8) * KG_BYTES + row * 16;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              st_shared_v4(a + g * KG_BYTES, make_uint4(it, g, q, tid));
              st_shared_v4(b + g * KG_BYTES, make_uint4(tid, q, g, it));
            }
            if (mode == 3) fence_proxy_async();
          }
        } else if (mode == 4) {
#pragma unroll
          for (int q = 0; q < 4; ++q) tmem_ld32(tmem_lane + q * 64 + sub * 32, v[q]);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col0 = q * 64 + sub * 32;
            float2 b[16];
            load_vec<16>(seed + col0, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float2 t = __ffma2_rn(v[q][i], make_float2(kF8InvScale, kF8InvScale), b[i]);
              const float2 u = __fmul2_rn(t, make_float2(0.2f, 0.2f));
              v[q][i] = make_float2(fmaxf(t.x, u.x), fmaxf(t.y, u.y));
            }
            store_step_f16f8(h_hi + (col0 / 8) * KG_BYTES + row * 16, h_lo + (col0 / 8) * KG_BYTES + row * 16, KG_BYTES, v[q]);
            fence_proxy_async();
          }
          acc += v[1][3].y;
        } else if (mode == 5) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float2 t[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) t[i] = v[q][c * 8 + i];
              tmem_st16(tmem_lane + q * 64 + sub * 32 + c * 16, t);
            }
          }
          tmem_st_wait();
        }
      }
    }
    const long long t1 = clock64();
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[q][i].x + v[q][i].y;
    sink[tid] = acc;
    if (active && (tid & 31) == 0) {
      atomicMin((long long*)&t_first, t0);
      atomicMax((long long*)&t_last, t1);
    }
    if (tid == 0) out[0] = (unsigned long long)(t1 - t0);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (tid == 0) out[1] = (unsigned long long)(t_last - t_first);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

// Weight-stream micro-benchmark: every CTA (one per SM) streams `iters` slots of `slot_bytes` from a `span_bytes` window of
// global memory (L2-resident when small) through a ring of `nslots` shared-memory slots with 1-D bulk copies, re-issuing a
// slot as soon as it has landed (no consumer): bytes in flight per SM = nslots * slot_bytes, so the sustained rate shows the
// L2 -> SM bulk-copy latency (Little's law) and where the chip-wide L2 bandwidth caps it.  out[0] = cycles of CTA 0.
__global__ void __launch_bounds__(32, 1)
ringbench_kernel(const uint8_t* __restrict__ src, unsigned long long span_bytes, int slot_bytes, int nslots, int iters,
                 unsigned long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + (uint32_t)nslots * (uint32_t)slot_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) mbar_init(bar + 8 * s, 1);
    fence_mbar_init();
  }
  __syncwarp();
  const unsigned long long nchunk = span_bytes / (unsigned long long)slot_bytes;
  unsigned long long c = (blockIdx.x * 37ull) % nchunk;
  const long long t0 = clock64();
  for (int i = 0; i < nslots && i < iters; ++i) {
    if (elect_one()) {
      mbar_expect_tx(bar + 8 * i, slot_bytes);
      bulk_g2s(sbase + i * slot_bytes, src + c * slot_bytes, slot_bytes, bar + 8 * i);
    }
    c = c + 1 == nchunk ? 0 : c + 1;
  }
  uint32_t slot = 0, ph = 0;
  for (int i = 0; i < iters; ++i) {
    mbar_wait(bar + 8 * slot, ph);
    if (i + nslots < iters && elect_one()) {
      mbar_expect_tx(bar + 8 * slot, slot_bytes);
      bulk_g2s(sbase + slot * slot_bytes, src + c * slot_bytes, slot_bytes, bar + 8 * slot);
    }
    c = c + 1 == nchunk ? 0 : c + 1;
    if (++slot == (uint32_t)nslots) { slot = 0; ph ^= 1; }
  }
  const long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
}

// Scattered-gather micro-benchmark: one CTA per SM with `smem_kb` of dynamic shared memory (the decode kernels leave ~4 KB of
// L1), 256 threads; 8 threads fetch one 256-byte channels-last texel (2 x float4 each) at pseudo-random positions of a
// `ntexel`-texel table, `U` texels in flight per thread, through one of several load forms:
//   0 ld.global.nc (__ldg)   1 ld.global.cg   2 ld.global.nc.L1::no_allocate   3 ld.global.cv   4 ld.global.L1::evict_first
// out[0] = cycles of CTA 0 for `iters` rounds of U texel-loads per 8-thread group.
template <int VAR>
__device__ __forceinline__ float4 ld_var(const float4* p) {
  float4 v;
  if (VAR == 0) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (VAR == 1) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (VAR == 2) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (VAR == 3) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  if (VAR == 4) asm volatile("ld.global.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
template <int VAR, int U>
__global__ void __launch_bounds__(256, 1)
gatherbench_kernel(const float* __restrict__ table, unsigned ntexel, int iters, unsigned long long* __restrict__ out, float* __restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, sub = tid & 7;
  unsigned state = (blockIdx.x * 256u + (tid >> 3)) * 2654435761u + 12345u;
  float acc = 0.f;
  if (tid == 0) smem[0] = 1;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float4 v[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      state = state * 1664525u + 1013904223u;
      const unsigned tx = (state >> 8) % ntexel;
      const float4* p = reinterpret_cast<const float4*>(table + (size_t)tx * 64) + sub * 2;
      v[u][0] = ld_var<VAR>(p);
      v[u][1] = ld_var<VAR>(p + 1);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u][0].x + v[u][0].w + v[u][1].y + v[u][1].z;
  }
  const long long t1 = clock64();
  sink[blockIdx.x * 256 + tid] = acc + smem[0];
  if (blockIdx.x == 0 && tid == 0) out[0] = (unsigned long long)(t1 - t0);
}

}  // namespace mbench

template <int VAR, int U>
static int launch_gb(const float* table, unsigned ntexel, int iters, int smem_kb, int ctas, unsigned long long* out, float* sink,
                     cudaStream_t st) {
  using namespace mbench;
  DDMI_CUDA(cudaFuncSetAttribute(gatherbench_kernel<VAR, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024));
  gatherbench_kernel<VAR, U><<<ctas, 256, smem_kb * 1024, st>>>(table, ntexel, iters, out, sink);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}
int launch_gatherbench(int var, int u, const float* table, unsigned ntexel, int iters, int smem_kb, int ctas,
                       unsigned long long* out, float* sink, cudaStream_t st) {
#define GB(V, UU) if (var == V && u == UU) return launch_gb<V, UU>(table, ntexel, iters, smem_kb, ctas, out, sink, st);
  GB(0, 4) GB(0, 12) GB(1, 4) GB(1, 12) GB(2, 4) GB(2, 12) GB(3, 4) GB(3, 12) GB(4, 4) GB(4, 12)
#undef GB
  set_error("gatherbench: variant %d / unroll %d not built", var, u);
  return DDMI_ERR_UNSUPPORTED;
}

// Bring-up self test of the plane-window TMA path: one 64 x 2 x 64 box of an NCHW fp32 plane -> shared memory -> out.
// variant 0: tensor map as a __grid_constant__ kernel parameter; 1: tensor map read from global memory.
namespace mbench {
__global__ void __launch_bounds__(128, 1)
tma_selftest_kernel(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int x, int y, int c, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + 32768;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 32768);
    tma::load_3d(sbase, gmap ? gmap : &pmap, x, y, c, bar);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 8192; i += 128) out[i] = reinterpret_cast<const float*>(smem)[i];
}
}  // namespace mbench

int launch_tma_selftest(const float* plane, int batch, int C, int H, int W, int x, int y, int c, int variant, void* map_dev,
                        float* out, cudaStream_t st) {
  using namespace mbench;
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  if (!tma::make_plane_map(&m, plane, batch, C, H, W, 64, 2, 64)) {
    set_error("cuTensorMapEncodeTiled failed (or is unavailable) for a %dx%d plane", H, W);
    return DDMI_ERR_UNSUPPORTED;
  }
  if (variant == 1) DDMI_CUDA(cudaMemcpyAsync(map_dev, &m, sizeof(m), cudaMemcpyHostToDevice, st));
  DDMI_CUDA(cudaFuncSetAttribute(tma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64));
  tma_selftest_kernel<<<1, 128, 32768 + 64, st>>>(m, variant == 1 ? (const CUtensorMap*)map_dev : nullptr, x, y, c, out);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_ringbench(const void* src, unsigned long long span_bytes, int slot_bytes, int nslots, int iters, int ctas,
                     unsigned long long* out, cudaStream_t st) {
  using namespace mbench;
  const int smem = 200 * 1024;     // one CTA per SM
  DDMI_CUDA(cudaFuncSetAttribute(ringbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ringbench_kernel<<<ctas, 32, smem, st>>>((const uint8_t*)src, span_bytes, slot_bytes, nslots, iters, out);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}


// tcgen05.mma issue / execution rate in the decode kernels' configuration (CTA pair, M = 256, operands zero-filled):
// one elected thread of the leader issues `iters` rounds of 8 MMAs and one commit per round, then waits for the last commit.
//   variant bit 0: N = 128 instead of 256      bit 1: A operand in tensor memory (TS) instead of shared memory (SS)
//   bit 2: FP8 (kind::f8f6f4, K = 32) instead of fp16 (kind::f16, K = 16)    bit 3: the f16f8 mix (f16 f16 f8 f8) x 2
//   bit 4: 8 other warps store to shared memory (st.shared.v4, A-operand-like pattern) for the duration
// out[0] = cycles from the first issue to the completion of the last commit, out[1] = MMAs issued.
namespace mbench {
constexpr int MM_A = 0, MM_B = 65536, MM_ST = 65536 + 16384, MM_BAR = MM_ST + 65536, MM_SMEM = MM_BAR + 64;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
mmabench_kernel(int variant, int iters, unsigned long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + MM_BAR;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < MM_BAR / 16; i += 320) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  volatile int* flag = reinterpret_cast<volatile int*>(smem + MM_BAR + 32);
  if (tid == 0) {
    mbar_init(bar, 1);
    *flag = 0;
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc2(bar + 16, 512);
  fence_proxy_async();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + MM_BAR + 16);
  const int n = (variant & 1) ? 128 : 256, nloc = n / 2;
  const bool ts = variant & 2, f8 = variant & 4, mix = variant & 8, stores = variant & 16;
  if (warp == 9 && rank == 0) {
    constexpr uint64_t kDescHi = ((uint64_t)(128 >> 4) | (1ull << 14)) << 32;
    const uint32_t idesc16 = idesc_f16_f32(256, 0) | ((uint32_t)n << 14), idesc8 = idesc_f8_f32(256, 0) | ((uint32_t)n << 14);
    const uint32_t a_lo = ((sbase + MM_A) >> 4) | ((2048 >> 4) << 16);
    const uint32_t b_lo = ((sbase + MM_B) >> 4) | ((uint32_t)nloc << 16);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool use8 = mix ? ((j & 3) >= 2) : f8;
          const uint64_t bd = kDescHi | (b_lo + (uint32_t)(j & 3) * (uint32_t)(nloc * 2));
          const uint32_t acc = tmem + ((variant & 1) ? (uint32_t)(j & 1) * 128u : 0u);
          if (ts) {
            if (use8) mma2_f8_ts(acc, tmem + 256 + j * 8, bd, idesc8, 1u);
            else mma2_bf16_ts(acc, tmem + 256 + j * 8, bd, idesc16, 1u);
          } else {
            const uint64_t ad = kDescHi | (a_lo + (uint32_t)j * (4096 >> 4));
            if (use8) mma2_f8(acc, ad, bd, idesc8, 1u);
            else mma2_bf16(acc, ad, bd, idesc16, 1u);
          }
        }
        if (it == iters - 1) mma2_commit_mc(bar, 3);
      }
      __syncwarp();
    }
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    *flag = 1;
    if ((tid & 31) == 0 && blockIdx.x == 0) {
      out[0] = (unsigned long long)(t1 - t0);
      out[1] = (unsigned long long)iters * 8ull;
    }
  } else if (warp == 9) {
    mbar_wait(bar, 0);
    *flag = 1;
  } else if (warp < 8 && stores) {
    const int row = tid & 127;
    uint32_t k = 0;
    while (*flag == 0) {
#pragma unroll
      for (int g = 0; g < 8; ++g) st_shared_v4(sbase + MM_ST + ((k + g) & 31) * 2048 + row * 16, make_uint4(k, g, row, tid));
      k += 8;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 9) tmem_dealloc2(tmem, 512);
}
}  // namespace mbench

int launch_microbench(int mode, int iters, const float* seed, unsigned long long* out, float* sink, cudaStream_t st) {
  using namespace mbench;
  if (mode >= 100) {     // tcgen05.mma rate: 100 + variant on one CTA pair, 200 + variant on every SM
    const int full = mode >= 200, variant = mode - (full ? 200 : 100);
    DDMI_CUDA(cudaFuncSetAttribute(mmabench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MM_SMEM));
    int sms = 0;
    DDMI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    mmabench_kernel<<<full ? (sms / 2) * 2 : 2, 320, MM_SMEM, st>>>(variant, iters, out);
    DDMI_CUDA(cudaGetLastError());
    return DDMI_OK;
  }
  DDMI_CUDA(cudaFuncSetAttribute(microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  microbench_kernel<<<1, 288, MB_SMEM, st>>>(mode, iters, seed, out, sink);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
