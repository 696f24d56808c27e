// tcgen05 (5th-gen tensor core) decode kernels: DDMI_PREC_BF16X3.
//
// image_umma_kernel -- MLP.forward (models/d2c_vae/mlp.py:34-66) fused end to end:
// one persistent CTA per SM walks 128-coordinate tiles; per tile
//   gather   : 3 bilinear plane lookups (align_corners=false, border) per coordinate,
//              written straight into the MMA A-operand layout as bf16 hi/lo pairs
//   13 GEMM groups on the tensor core: D[128 x 256] (+)= A[128 x K] * W^T, fp32
//              accumulators in TMEM, every product as 3 bf16 MMAs
//              (Ahi*Bhi + Alo*Bhi + Ahi*Blo) so the result carries ~16 mantissa bits
//   epilogues: TMEM -> registers (tcgen05.ld), bias + leaky-ReLU (the reference's
//              fused_bias_act op, op/fused_bias_act_kernel.cu:28-47) + residual,
//              re-split to bf16 hi/lo and written back as the next layer's A operand
//   ToRGB    : N = 16 MMA, 3 columns stored.
// Weights stream from L2 through a 4-stage shared-memory ring filled by 1-D bulk
// async copies (cp.async.bulk + mbarrier complete_tx); the host packs them in the
// exact consumption order and shared-memory image (ddmi_b200/packing.py).
//
// Warp roles (10 warps): 0-7 gather + epilogue (warp w owns TMEM lanes 32*(w%4).. and
// column half w/4), 8 weight producer (one lane), 9 MMA issuer (one lane) + TMEM owner.
#include "common.cuh"
#include "umma.cuh"

namespace ddmi {
namespace ummak {

using namespace umma;

constexpr int TILE = 128;             // coordinates per tile == MMA M
constexpr int NEPI = 256;             // epilogue / gather threads
constexpr int NTHREADS = 384;          // 3 warpgroups: 2 x gather/epilogue (208 regs), 1 x {producer, MMA, 2 idle warps} (88 regs); 208*256 + 88*128 == 168*384: setmaxnreg can only hand out what was released
constexpr int KG_BYTES = TILE * 16;   // one 8-wide K group of an A operand: 128 rows x 16 B
constexpr int H_BYTES = 32 * KG_BYTES;   // 256-wide activation, one of hi / lo
constexpr int X_BYTES = 8 * KG_BYTES;    // 64-wide PE features, one of hi / lo
constexpr int STAGE_BYTES = 16384;    // one K step (16) of a 256-wide layer: hi 8 KB | lo 8 KB
constexpr int NSTAGE = 4;
constexpr int CHUNKS_PER_TILE = 233;  // 232 K steps of N = 256, + 1 chunk holding ToRGB's 16 K steps of N = 16

constexpr int OFF_HHI = 0;
constexpr int OFF_HLO = OFF_HHI + H_BYTES;
constexpr int OFF_XHI = OFF_HLO + H_BYTES;
constexpr int OFF_XLO = OFF_XHI + X_BYTES;
constexpr int OFF_W = OFF_XLO + X_BYTES;
constexpr int OFF_BAR = OFF_W + NSTAGE * STAGE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
// barriers (8 B each) at OFF_BAR: w_full[4], w_empty[4], mma_done, a_ready; then the TMEM base address
constexpr int BAR_WFULL = 0, BAR_WEMPTY = 32, BAR_MMADONE = 64, BAR_AREADY = 72, TMEM_SLOT = 80;

constexpr uint32_t IDESC_N256 = idesc_bf16_f32(256);
constexpr uint32_t IDESC_N16 = idesc_bf16_f32(16);

// Diagnostics: cycle counters of CTA 0 (see ddmi_debug_profile in the header).
// [0] epilogue thread 0: cycles parked waiting for MMA groups   [1] cycles in epilogue stages
// [2] cycles in gathers   [3] MMA thread: cycles waiting for operands (a_ready)
// [4] MMA thread: cycles waiting for weight chunks   [5] MMA thread: total   [6] tiles   [7] spare
__device__ unsigned long long g_prof[8];

constexpr float kSqrt2 = 1.41421356237309504880f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;

// ---------------------------------------------------------------------------
// epilogue helpers (one thread = one tile row, 128 of the 256 output columns)
// ---------------------------------------------------------------------------
// y[0..31] (16 pairs) -> bf16 hi/lo, written as 4 K groups of the 256-wide A operand
__device__ __forceinline__ void store_act32(uint32_t h_hi, uint32_t h_lo, int row, int col0, const float2 (&y)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 hi, lo;
    split8(&y[g * 4], hi, lo);
    const uint32_t off = (uint32_t)((col0 / 8 + g) * KG_BYTES + row * 16);
    st_shared_v4(h_hi + off, hi);
    st_shared_v4(h_lo + off, lo);
  }
}

template <int NP>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float2 (&b)[NP]) {
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
    b[2 * i] = make_float2(v.x, v.y);
    b[2 * i + 1] = make_float2(v.z, v.w);
  }
}
// 16 columns (8 pairs) -> two K groups of the A operand
__device__ __forceinline__ void store_act16(uint32_t h_hi, uint32_t h_lo, int row, int col0, const float2 (&y)[8]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    uint4 hi, lo;
    split8(&y[g * 4], hi, lo);
    const uint32_t off = (uint32_t)((col0 / 8 + g) * KG_BYTES + row * 16);
    st_shared_v4(h_hi + off, hi);
    st_shared_v4(h_lo + off, lo);
  }
}

// One epilogue stage for this thread's row and 128 of the 256 columns.  sqrt2 gains are folded into
// the weights / biases on the host (leaky ReLU is positively homogeneous).  The TMEM load of the
// next chunk is in flight while the current one is converted.
// MODE 0: H = lrelu(acc1 + b)                              (conv1 / conv2; 32-column chunks)
// MODE 1: H = lrelu(acc1 + b) + acc2 + cs                  (conv3 + skip; res1, res2; 16-column chunks)
// MODE 2: as 1, and acc2 <- H / sqrt2                      (res3: stash res4's identity skip)
// MODE 3: H = lrelu(acc1 + b) + acc2                       (res4: acc2 holds h3 / sqrt2)
template <int MODE>
__device__ __forceinline__ void epilogue(uint32_t tmem, uint32_t h_hi, uint32_t h_lo, int row, int lane_base,
                                         int col_half, const float* __restrict__ bias,
                                         const float* __restrict__ cs) {
  const uint32_t t0 = tmem + ((uint32_t)lane_base << 16) + (uint32_t)(col_half * 128);
  if (MODE == 0) {
    float2 v[2][16];
    tmem_ld32(t0, v[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int cur = c & 1, col0 = col_half * 128 + c * 32;
      float2 b[16];
      load_vec<16>(bias + col0, b);
      tmem_ld_wait();
      if (c < 3) tmem_ld32(t0 + (c + 1) * 32, v[cur ^ 1]);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[cur][i] = bias_lrelu_pair(v[cur][i], b[i], 0.2f);
      store_act32(h_hi, h_lo, row, col0, v[cur]);
    }
  } else {
    float2 v[2][8], s[2][8];
    tmem_ld16(t0, v[0]);
    tmem_ld16(t0 + 256, s[0]);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int cur = c & 1, col0 = col_half * 128 + c * 16;
      float2 b[8];
      load_vec<8>(bias + col0, b);
      tmem_ld_wait();
      if (c < 7) {
        tmem_ld16(t0 + (c + 1) * 16, v[cur ^ 1]);
        tmem_ld16(t0 + 256 + (c + 1) * 16, s[cur ^ 1]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) v[cur][i] = __fadd2_rn(bias_lrelu_pair(v[cur][i], b[i], 0.2f), s[cur][i]);
      if (MODE == 1 || MODE == 2) {
        load_vec<8>(cs + col0, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[cur][i] = __fadd2_rn(v[cur][i], b[i]);
      }
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[cur][i] = __fmul2_rn(v[cur][i], make_float2(kInvSqrt2, kInvSqrt2));
        tmem_st16(t0 + 256 + c * 16, s[cur]);
      }
      store_act16(h_hi, h_lo, row, col0, v[cur]);
    }
    if (MODE == 2) tmem_st_wait();
  }
}

// ---------------------------------------------------------------------------
// the fused image kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
image_umma_kernel(PlaneSet ps, const float* __restrict__ cx, const float* __restrict__ cy, long long n,
                  int tiles_per_item, long long total_tiles, const uint8_t* __restrict__ wstream,
                  const float* __restrict__ vec, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase + OFF_HHI, h_lo = sbase + OFF_HLO;
  const uint32_t x_hi = sbase + OFF_XHI, x_lo = sbase + OFF_XLO;
  const uint32_t wst = sbase + OFF_W;
  const uint32_t bar = sbase + OFF_BAR;
  const uint32_t b_mma = bar + BAR_MMADONE, b_ardy = bar + BAR_AREADY;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int C = 64;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar + BAR_WFULL + 8 * s, 1);
      mbar_init(bar + BAR_WEMPTY + 8 * s, 1);
    }
    mbar_init(b_mma, 1);
    mbar_init(b_ardy, NEPI);
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc(bar + TMEM_SLOT, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + TMEM_SLOT);

  const long long first = blockIdx.x, stride = gridDim.x;

  if (warp < 8) {
    // =================== gather + epilogue threads ===================
    reg_inc<208>();
    const int row = tid & 127;             // tile row == TMEM lane
    const int lane_base = (warp & 3) * 32; // this warp's TMEM lane quadrant
    const int col_half = warp >> 2;        // output columns [128*col_half, +128)
    const int ghalf = tid >> 7;            // gather: channels [32*ghalf, +32)
    uint32_t ph_mma = 0;
    const bool prof = (blockIdx.x == 0 && tid == 0);
    long long p_wait = 0, p_epi = 0, p_gather = 0, p_t = 0;

    auto gather = [&](long long tile, int s) {
      const long long g0 = clock64();
      const int b = (int)(tile / tiles_per_item);
      long long gi = (tile % tiles_per_item) * TILE + row;
      if (gi > n - 1) gi = n - 1;
      const Tap t = make_tap<false>(__ldg(cx + gi), __ldg(cy + gi), ps.h[s], ps.w[s]);
      const size_t hw = (size_t)ps.h[s] * ps.w[s];
      const float* base = ps.data[s] + ((size_t)b * C + ghalf * 32) * hw;
#pragma unroll 1
      for (int g = 0; g < 4; ++g) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = tap_sample(base + (size_t)(g * 8 + i) * hw, t);
        uint4 hi, lo;
        split8(y, hi, lo);
        const uint32_t off = (uint32_t)((ghalf * 4 + g) * KG_BYTES + row * 16);
        st_shared_v4(x_hi + off, hi);
        st_shared_v4(x_lo + off, lo);
      }
      p_gather += clock64() - g0;
    };
    auto publish = [&]() {   // make this thread's smem / TMEM writes visible to the MMA warp, then signal
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(b_ardy);
      p_epi += clock64() - p_t;
    };
    auto wait_mma = [&]() {
      const long long w0 = clock64();
      mbar_wait(b_mma, ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      p_t = clock64();
      p_wait += p_t - w0;
    };

    gather(first, 0);
    publish();
    for (long long tile = first; tile < total_tiles; tile += stride) {
      const float* bv = vec;
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk, bv += 1024) {
        // conv1 (+ skip into acc2 for blk < 3)
        wait_mma();
        epilogue<0>(tmem, h_hi, h_lo, row, lane_base, col_half, bv, nullptr);
        publish();
        // the PE buffer is free now: prefetch the next scale (or the next tile's coarse scale)
        if (blk < 2) gather(tile, blk + 1);
        else if (blk == 2 && tile + stride < total_tiles) gather(tile + stride, 0);
        // conv2
        wait_mma();
        epilogue<0>(tmem, h_hi, h_lo, row, lane_base, col_half, bv + 256, nullptr);
        publish();
        // conv3 + skip
        wait_mma();
        if (blk < 2) epilogue<1>(tmem, h_hi, h_lo, row, lane_base, col_half, bv + 512, bv + 768);
        else if (blk == 2) epilogue<2>(tmem, h_hi, h_lo, row, lane_base, col_half, bv + 512, bv + 768);
        else epilogue<3>(tmem, h_hi, h_lo, row, lane_base, col_half, bv + 512, nullptr);
        publish();
      }
      // ToRGB: acc1[:, 0:16]
      wait_mma();
      if (col_half == 0) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)lane_base << 16), v);   // only columns 0..2 are meaningful
        tmem_ld_wait();
        const int b = (int)(tile / tiles_per_item);
        const long long gi = (tile % tiles_per_item) * TILE + row;
        if (gi < n) {
          const float* brgb = vec + 4096 + 768;
#pragma unroll
          for (int c = 0; c < 3; ++c) out[((size_t)b * 3 + c) * n + gi] = v[c] + __ldg(brgb + c);
        }
      }
      publish();
    }
    if (prof) {
      atomicAdd(&g_prof[0], (unsigned long long)p_wait);
      atomicAdd(&g_prof[1], (unsigned long long)(p_epi - p_gather));
      atomicAdd(&g_prof[2], (unsigned long long)p_gather);
    }
  } else {
   reg_dec<88>();   // the whole third warpgroup (warps 8-11) executes this one instruction
   if (warp == 8) {
    // =================== weight producer ===================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (long long tile = first; tile < total_tiles; tile += stride) {
        const uint8_t* src = wstream;
        for (int c = 0; c < CHUNKS_PER_TILE; ++c, src += STAGE_BYTES) {
          mbar_wait(bar + BAR_WEMPTY + 8 * s, ph ^ 1);
          mbar_expect_tx(bar + BAR_WFULL + 8 * s, STAGE_BYTES);
          bulk_g2s(wst + s * STAGE_BYTES, src, STAGE_BYTES, bar + BAR_WFULL + 8 * s);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      }
    }
   } else if (warp == 9) {
    // =================== MMA issuer ===================
    if (lane == 0) {
      uint32_t s = 0, ph = 0, ph_a = 0;
      const uint32_t acc1 = tmem, acc2 = tmem + 256;
      long long q_a = 0, q_w = 0, q_tiles = 0;
      const long long q_start = clock64();
      // one GEMM segment: acc (+)= A[:, nk*16] * W^T, W K-steps taken from the ring
      auto seg = [&](uint32_t a_hi, uint32_t a_lo, int nk, uint32_t acc, uint32_t first_acc) {
        for (int j = 0; j < nk; ++j) {
          if (!mbar_try_wait(bar + BAR_WFULL + 8 * s, ph)) {
            const long long w0 = clock64();
            mbar_wait(bar + BAR_WFULL + 8 * s, ph);
            q_w += clock64() - w0;
          }
          tc_fence_after();
          const uint32_t wb = wst + s * STAGE_BYTES;
          const uint64_t bhi = smem_desc(wb, 256 * 16, 128), blo = smem_desc(wb + 8192, 256 * 16, 128);
          const uint64_t ahi = smem_desc(a_hi + j * 2 * KG_BYTES, KG_BYTES, 128);
          const uint64_t alo = smem_desc(a_lo + j * 2 * KG_BYTES, KG_BYTES, 128);
          mma_bf16(acc, ahi, bhi, IDESC_N256, (j > 0) ? 1u : first_acc);
          mma_bf16(acc, alo, bhi, IDESC_N256, 1u);
          mma_bf16(acc, ahi, blo, IDESC_N256, 1u);
          mma_commit(bar + BAR_WEMPTY + 8 * s);
          if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
      };
      auto wait_a = [&]() {
        const long long w0 = clock64();
        mbar_wait(b_ardy, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        q_a += clock64() - w0;
      };
      for (long long tile = first; tile < total_tiles; tile += stride) {
        for (int blk = 0; blk < 4; ++blk) {
          wait_a();
          if (blk == 0) {
            seg(x_hi, x_lo, 4, acc2, 0u);
            seg(x_hi, x_lo, 4, acc1, 0u);
          } else if (blk < 3) {
            seg(h_hi, h_lo, 16, acc2, 0u);
            seg(x_hi, x_lo, 4, acc2, 1u);
            seg(h_hi, h_lo, 16, acc1, 0u);
            seg(x_hi, x_lo, 4, acc1, 1u);
          } else {
            seg(h_hi, h_lo, 16, acc1, 0u);
          }
          mma_commit(b_mma);
          wait_a();
          seg(h_hi, h_lo, 16, acc1, 0u);
          mma_commit(b_mma);
          wait_a();
          seg(h_hi, h_lo, 16, acc1, 0u);
          mma_commit(b_mma);
        }
        // ToRGB: one ring chunk = 16 K steps of [hi 512 B | lo 512 B] (N = 16)
        wait_a();
        mbar_wait(bar + BAR_WFULL + 8 * s, ph);
        tc_fence_after();
        const uint32_t wb = wst + s * STAGE_BYTES;
        for (int j = 0; j < 16; ++j) {
          const uint64_t bhi = smem_desc(wb + j * 1024, 16 * 16, 128), blo = smem_desc(wb + j * 1024 + 512, 16 * 16, 128);
          const uint64_t ahi = smem_desc(h_hi + j * 2 * KG_BYTES, KG_BYTES, 128);
          const uint64_t alo = smem_desc(h_lo + j * 2 * KG_BYTES, KG_BYTES, 128);
          mma_bf16(acc1, ahi, bhi, IDESC_N16, j > 0 ? 1u : 0u);
          mma_bf16(acc1, alo, bhi, IDESC_N16, 1u);
          mma_bf16(acc1, ahi, blo, IDESC_N16, 1u);
        }
        mma_commit(bar + BAR_WEMPTY + 8 * s);
        if (++s == NSTAGE) { s = 0; ph ^= 1; }
        mma_commit(b_mma);
        ++q_tiles;
      }
      if (blockIdx.x == 0) {
        atomicAdd(&g_prof[3], (unsigned long long)q_a);
        atomicAdd(&g_prof[4], (unsigned long long)q_w);
        atomicAdd(&g_prof[5], (unsigned long long)(clock64() - q_start));
        atomicAdd(&g_prof[6], (unsigned long long)q_tiles);
      }
    }
    __syncwarp();
   }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------
// bring-up self test: D[128 x N] = A[128 x K] * B[N x K]^T through the same
// descriptors, split, MMA and TMEM load paths as the decode kernel.
// ---------------------------------------------------------------------------
constexpr int ST_OFF_AHI = 0, ST_OFF_ALO = H_BYTES, ST_OFF_B = 2 * H_BYTES, ST_OFF_BAR = ST_OFF_B + STAGE_BYTES;
constexpr int ST_SMEM = ST_OFF_BAR + 64;

__global__ void __launch_bounds__(160, 1)
selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int N, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_hi = sbase + ST_OFF_AHI, a_lo = sbase + ST_OFF_ALO, bst = sbase + ST_OFF_B;
  const uint32_t bar = sbase + ST_OFF_BAR;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(bar + 8, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + ST_OFF_BAR + 8);
  // A: row = tid (128 rows), all K groups
  if (tid < 128) {
    for (int g = 0; g < K / 8; ++g) {
      float y[8];
      for (int i = 0; i < 8; ++i) y[i] = a[(size_t)tid * K + g * 8 + i];
      uint4 hi, lo;
      split8(y, hi, lo);
      st_shared_v4(a_hi + g * KG_BYTES + tid * 16, hi);
      st_shared_v4(a_lo + g * KG_BYTES + tid * 16, lo);
    }
  }
  const uint32_t idesc = idesc_bf16_f32(N);
  uint32_t ph = 0;
  for (int j = 0; j < K / 16; ++j) {
    // B K-step block: [hi: 2 kgroups x N rows x 16 B | lo: same]
    if (tid < 128) {
      for (int r = tid; r < N; r += 128) {
        for (int g = 0; g < 2; ++g) {
          float y[8];
          for (int i = 0; i < 8; ++i) y[i] = b[(size_t)r * K + j * 16 + g * 8 + i];
          uint4 hi, lo;
          split8(y, hi, lo);
          st_shared_v4(bst + (g * N + r) * 16, hi);
          st_shared_v4(bst + N * 32 + (g * N + r) * 16, lo);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 128) {
      tc_fence_after();
      const uint64_t bhi = smem_desc(bst, N * 16, 128), blo = smem_desc(bst + N * 32, N * 16, 128);
      const uint64_t ahi = smem_desc(a_hi + j * 2 * KG_BYTES, KG_BYTES, 128);
      const uint64_t alo = smem_desc(a_lo + j * 2 * KG_BYTES, KG_BYTES, 128);
      mma_bf16(tmem, ahi, bhi, idesc, j > 0 ? 1u : 0u);
      mma_bf16(tmem, alo, bhi, idesc, 1u);
      mma_bf16(tmem, ahi, blo, idesc, 1u);
      mma_commit(bar);
    }
    mbar_wait(bar, ph);   // everyone: the B block may be overwritten once the MMAs are done
    ph ^= 1;
    tc_fence_after();
  }
  if (tid < 128) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 32 && c0 + i < N; ++i) d[(size_t)tid * N + c0 + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace ummak

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
int launch_image_umma(const PlaneSet& ps, int batch, int C, const float* cx, const float* cy, long long n,
                      const void* gemm, size_t gemm_bytes, const float* vec, size_t vec_floats, float* out,
                      cudaStream_t st) {
  using namespace ummak;
  if (C != 64) {
    set_error("tcgen05 image kernel is built for 64-channel planes");
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(gemm_bytes == (size_t)CHUNKS_PER_TILE * STAGE_BYTES, "packed bf16x3 stream is %zu bytes, expected %zu",
               gemm_bytes, (size_t)CHUNKS_PER_TILE * STAGE_BYTES);
  DDMI_REQUIRE(vec_floats == 4096 + 768 + 3, "packed vec blob is %zu floats, expected 4867", vec_floats);
  int dev = 0, sms = 0;
  DDMI_CUDA(cudaGetDevice(&dev));
  DDMI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tpi = (n + TILE - 1) / TILE;
  const long long total = tpi * batch;
  if (tpi > 2147483647LL) {
    set_error("n_coords %lld too large for one launch", n);
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_CUDA(cudaFuncSetAttribute(image_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  image_umma_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(ps, cx, cy, n, (int)tpi, total, (const uint8_t*)gemm, vec, out);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int debug_profile(unsigned long long* out, int reset) {
  DDMI_CUDA(cudaMemcpyFromSymbol(out, ummak::g_prof, sizeof(unsigned long long) * 8));
  if (reset) {
    unsigned long long z[8] = {};
    DDMI_CUDA(cudaMemcpyToSymbol(ummak::g_prof, z, sizeof(z)));
  }
  return DDMI_OK;
}

int launch_selftest_umma(const float* a, const float* b, float* d, int N, int K, cudaStream_t st) {
  using namespace ummak;
  DDMI_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
  selftest_kernel<<<1, 160, ST_SMEM, st>>>(a, b, d, N, K);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
