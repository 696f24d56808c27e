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
#include "common.cuh"
#include "umma.cuh"

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

}  // namespace mbench

int launch_microbench(int mode, int iters, const float* seed, unsigned long long* out, float* sink, cudaStream_t st) {
  using namespace mbench;
  DDMI_CUDA(cudaFuncSetAttribute(microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MB_SMEM));
  microbench_kernel<<<1, 288, MB_SMEM, st>>>(mode, iters, seed, out, sink);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
