// tcgen05 (5th-gen tensor core) decode kernels: DDMI_PREC_BF16X3 and DDMI_PREC_F16F8 (template SCHEME; umma.cuh describes the
// f16f8 operand scheme: an fp16 main MMA + one e4m3 correction MMA per 16-wide K step, issued a 32-wide step pair at a time).
//
// Engine (shared by every decoder family): one persistent CTA per SM walks 128-row tiles.
//   * warps 0-7  : gather + epilogue ("E" threads; warp w owns TMEM lanes 32*(w%4)..+31)
//   * warp  8    : weight producer -- one lane streams weight units (8 KB per CTA of a pair) from L2 into a
//                  shared-memory ring (4 or 8 slots) with 1-D bulk async copies (cp.async.bulk + mbarrier)
//   * warp  9    : MMA issuer -- one lane interprets the host-built PROGRAM (packing.py): a list
//                  of UNIT ops (one 16-wide K step of a 128 x N block = 3 tcgen05.mma: Ahi*Bhi +
//                  Alo*Bhi + Ahi*Blo, fp32 accumulate in TMEM), WAIT ops (operands ready) and
//                  COMMIT ops (tcgen05.commit -> the E threads may drain the accumulator).
//   The program and the weight stream are generated together, so the ring is consumed in exactly
//   the order it is produced and neither warp knows anything about the network.
//
// K-split software pipeline: every epilogue thread first drains ALL of its accumulator values
// into registers (the TMEM accumulator is free again), then converts + stores the 256 output columns
// in four quarters (= 4 K steps each of the next layer's A operand), signalling operand barrier A_q
// after each.  The program places the next layer's K steps of quarter q (full N = 256) after WAIT q,
// so the tensor core is already working while the later quarters are still being converted.
//
// CTA pairs (PAIR = 1, the default): two CTAs of a cluster run tcgen05.mma.cta_group::2 (M = 256): each CTA
// owns its own 128-row tile (A operand, accumulators, epilogue) but only HALF of every weight K step
// (N/2 rows of B), so the per-CTA shared-memory operand traffic and the L2 -> SM weight stream are cut
// (the single-CTA kernel is shared-memory-bandwidth bound, profiles/r01_*).  Only the leader CTA's warp 9
// issues MMAs; the peer's warp 9 forwards "my half of slot s has landed" to the leader; epilogue warps
// of both CTAs arrive on the leader's operand barriers; tcgen05.commit multicasts to both CTAs.
//
// image_umma_kernel -- MLP.forward (models/d2c_vae/mlp.py:34-66): gather of the 3 PE planes
// (align_corners=false, border) straight into the A-operand layout (bf16 hi/lo), 13 GEMM groups,
// epilogues = the reference's fused_bias_act (op/fused_bias_act_kernel.cu:28-47) + residuals,
// ToRGB as an N=16 block.
#include <string.h>
#include "umma_engine.cuh"
#include "tma.cuh"
#include "image_ts_issuer.cuh"
#include "decode_umma_occ.cuh"
#include "decode_umma_nerf.cuh"
#include "decode_umma_video.cuh"

namespace ddmi {
namespace ummak {

// One image epilogue stage for this thread's row.  The 256 output columns are produced in four
// quarters of 64 (= 4 K steps of the next layer); within a quarter the two warps that share a
// TMEM lane quadrant take 32 columns each (sub = 0 / 1).  All 128 accumulator values of the
// thread are drained into registers first (the accumulator is free for the next layer at once),
// then converted quarter by quarter with `signal(q)` after each.
// Bias / skip-constant vectors come from global memory and there is no L1 left beside 224 KB of
// shared memory, so they are PREFETCHED: quarter 0's before the thread parks on the MMA barrier
// (StagePrefetch), quarter q+1's while quarter q is converted.
// sqrt2 gains are folded into weights / biases on the host.
// MODE 0: H = lrelu(acc1 + b)                     (conv1 / conv2)
// MODE 1: H = lrelu(acc1 + b) + acc2 + cs         (conv3 + skip; res1, res2)
// MODE 2: as 1, and acc2 <- H / sqrt2             (res3: stash res4's identity skip)
// MODE 3: H = lrelu(acc1 + b) + acc2              (res4: acc2 holds h3 / sqrt2)
struct StagePrefetch {
  float2 b[16];    // bias of quarter 0 (this thread's 32 columns)
  float2 c[16];    // skip constant of quarter 0 (MODE 1 / 2)
};
template <bool WITH_CS>
__device__ __forceinline__ void stage_prefetch(StagePrefetch& pf, const float* __restrict__ bias,
                                               const float* __restrict__ cs, int sub) {
  load_vec<16>(bias + sub * 32, pf.b);
  if (WITH_CS) load_vec<16>(cs + sub * 32, pf.c);
}

// SCHEME 1 (f16f8): the accumulators hold 4096 x the GEMM result; 1/4096 rides in the bias / skip FMAs.
template <int SCHEME>
__device__ __forceinline__ float2 image_act(float2 t, float2 b) {
  if (SCHEME) {
    t = __ffma2_rn(t, make_float2(kF8InvScale, kF8InvScale), b);
    const float2 u = __fmul2_rn(t, make_float2(0.2f, 0.2f));
    return make_float2(fmaxf(t.x, u.x), fmaxf(t.y, u.y));
  }
  return bias_lrelu_pair(t, b, 0.2f);
}
// NOISE: nz = noise.weight (x activation gain) * noise[b, row] of this StyledConv, added to every column before the bias.
// N split: a layer's GEMM group may finish its output columns 0..127 (COMMIT 0) a K-half before columns 128..255
// (COMMIT 1) -- see packing.pack_image.  The stage is entered after COMMIT 0: it drains and publishes the first half
// (quarters 0, 1: the operand columns 0..127 they overwrite are dead by then, the program reads them in the first K-half
// only), waits for COMMIT 1 after quarter 0 (`wait_b`), drains the second half, reports the accumulator free (`drained`:
// operand barrier 4) and carries on with quarters 1..3.  With an unsplit program both commits arrive together.
template <int MODE, int SCHEME, int NOISE, class Signal, class WaitB, class Drained>
__device__ __forceinline__ void image_stage(uint32_t tmem_lane, uint32_t h_hi, uint32_t h_lo, int row, int sub,
                                            const float* __restrict__ bias, const float* __restrict__ cs,
                                            StagePrefetch& pf, Signal signal, WaitB wait_b, Drained drained, float nz, bool tr,
                                            uint32_t& trn) {
  constexpr bool WITH_CS = (MODE == 1 || MODE == 2);
  if (dbg(1)) {
    wait_b();
    drained();
#pragma unroll
    for (int q = 0; q < 4; ++q) signal(q);
    return;
  }
  const float2 nz2 = make_float2(nz, nz);
  if (NOISE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pf.b[i] = __fadd2_rn(pf.b[i], nz2);
  }
  float2 v[4][16];
#pragma unroll
  for (int q = 0; q < 2; ++q) tmem_ld32(tmem_lane + q * 64 + sub * 32, v[q]);
  tmem_ld_wait();
  trace(tr, 0x03, trn, 0);                              // first accumulator half drained
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col0 = q * 64 + sub * 32;
    if (q == 1) {
      wait_b();
#pragma unroll
      for (int q2 = 2; q2 < 4; ++q2) tmem_ld32(tmem_lane + q2 * 64 + sub * 32, v[q2]);
      tmem_ld_wait();
      drained();
    }
    float2 bn[16], cn[16];
    if (q < 3) {                                        // next quarter's vectors: in flight during this conversion
      load_vec<16>(bias + col0 + 64, bn);
      if (WITH_CS) load_vec<16>(cs + col0 + 64, cn);
    }
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[q][i] = image_act<SCHEME>(v[q][i], pf.b[i]);
    } else {
#pragma unroll
      for (int c = 0; c < 2; ++c) {                    // acc2 in 16-column pieces (TMEM reads are ~50 cycles)
        float2 s[8];
        tmem_ld16(tmem_lane + 256 + col0 + c * 16, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float2 y = image_act<SCHEME>(v[q][c * 8 + i], pf.b[c * 8 + i]);
          y = SCHEME ? __ffma2_rn(s[i], make_float2(kF8InvScale, kF8InvScale), y) : __fadd2_rn(y, s[i]);
          if (WITH_CS) y = __fadd2_rn(y, pf.c[c * 8 + i]);
          v[q][c * 8 + i] = y;
          constexpr float kStash = SCHEME ? kInvSqrt2 * kF8Scale : kInvSqrt2;
          if (MODE == 2) s[i] = __fmul2_rn(y, make_float2(kStash, kStash));
        }
        if (MODE == 2) tmem_st16(tmem_lane + 256 + col0 + c * 16, s);
      }
    }
    if (SCHEME) store_step_f16f8(h_hi + (col0 / 8) * KG_BYTES + row * 16, h_lo + (col0 / 8) * KG_BYTES + row * 16, KG_BYTES, v[q]);
    else store_act<16>(h_hi, h_lo, row, col0, v[q]);
    if (MODE == 2 && q == 3) tmem_st_wait();
    signal(q);
    if (q < 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        pf.b[i] = NOISE ? __fadd2_rn(bn[i], nz2) : bn[i];
        if (WITH_CS) pf.c[i] = cn[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// TS = 1 (f16f8, CTA pairs; packing._pack_image_ts): the 256-wide running activation H lives in TENSOR memory.
//   TMEM columns [0, 256) the ONE accumulator | [256, 384) H fp16 (two per column) | [384, 512) H FP8, per 32 K columns
//   [r8: 8 columns | a8: 8 columns].  tcgen05.mma with the A operand in tensor memory runs at the tensor pipe's own rate
//   (136 cycles per 256 x 256 x 16 step against 162-173 with A in shared memory: profiles/r02_mmabench.txt), and an epilogue
//   stage publishes with tcgen05.st (265 cycles per 128 values) instead of 32 st.shared.v4 (1556 cycles: the shared-memory
//   store bandwidth bounded the epilogue, profiles/r02_microbench_epilogue.txt).
//   With one accumulator the skip GEMM of a block is its own group in front of conv1 and its result is PARKED: the 128 KB of
//   shared memory that held H (K groups 0..63) keep, per epilogue thread, 32 private 16-byte chunks (chunk c at c * 4096 +
//   tid * 16: conflict-free) with the thread's 128 fp32 skip values until conv3's epilogue adds them.
//   Every group is N-split (COMMIT 0: output columns 0..127, COMMIT 1: 128..255; see image_stage) and every stage reports
//   "columns 0..127 / 128..255 of the accumulator are in registers" on operand barriers 5 / 4.
//   ToRGB (256 -> 3) is evaluated by the epilogue threads in fp32 from res4's output (weights: kernel parameter).
// MODE 0: H = lrelu(acc + b)                          (conv1 / conv2)
// MODE 1: H = lrelu(acc + b) + parked + cs            (conv3 + skip; res1, res2)
// MODE 2: as 1, and parked <- H / sqrt2               (res3: res4's identity skip)
// MODE 3: y = lrelu(acc + b) + parked -> rgb += Wrgb y (res4 + ToRGB; nothing is published)
// MODE 4: parked <- acc                               (skip GEMM)
// ---------------------------------------------------------------------------
struct RgbParam {
  float w[3 * 256];
};

// this thread's 32 columns [col0, col0 + 32) of H -> tensor memory
__device__ __forceinline__ void publish_ts(uint32_t tmem_lane, int col0, const float2* y) {
  uint4 a16[4], r8[2], a8[2];
  split32_f16f8(y, a16, r8, a8);
  uint32_t w16[16], wr[8], wa[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    w16[4 * g] = a16[g].x; w16[4 * g + 1] = a16[g].y; w16[4 * g + 2] = a16[g].z; w16[4 * g + 3] = a16[g].w;
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    wr[4 * g] = r8[g].x; wr[4 * g + 1] = r8[g].y; wr[4 * g + 2] = r8[g].z; wr[4 * g + 3] = r8[g].w;
    wa[4 * g] = a8[g].x; wa[4 * g + 1] = a8[g].y; wa[4 * g + 2] = a8[g].z; wa[4 * g + 3] = a8[g].w;
  }
  tmem_st16w(tmem_lane + 256 + col0 / 2, w16);
  tmem_st8w(tmem_lane + 384 + (col0 / 32) * 16, wr);
  tmem_st8w(tmem_lane + 384 + (col0 / 32) * 16 + 8, wa);
}

constexpr uint32_t PARK_STRIDE = NEPI * 16;   // chunk c of a thread: park + c * PARK_STRIDE

// CODE SIZE is a first-order concern here: with every stage, quarter and call site unrolled the epilogue threads walked ~160 KB
// of straight-line code per tile and 35 % of their stall samples were instruction-cache misses (profiles/r02_image_ts_ncu.md).
// A stage is therefore a two-trip loop over accumulator halves (two quarters unrolled in the body), conv1 / conv2 share one
// call site, and the three conv3 flavours are one body with run-time switches.
// KIND 0: H = lrelu(acc + b)                                         (conv1 / conv2)
// KIND 1: y = lrelu(acc + b) + parked;  blk 2: parked <- y / sqrt2 (res4's identity skip);
//         blk < 3: H = y, else rgb += Wrgb y (ToRGB; nothing is published)
// KIND 2: parked <- acc + cs                                         (skip GEMM; cs = the folded scale-injection columns)
// One quarter of a stage (this thread's 32 columns starting at col0 = q * 64 + sub * 32); pb = the quarter's bias / skip constant.
template <int KIND, int NOISE, class Sig>
__device__ __forceinline__ void image_ts_quarter(uint32_t tmem_lane, uint32_t park, int q, int col0, float2 (&v)[16],
                                                 float2 (&pb)[16], Sig sig, float2 nz2, int blk, const RgbParam& rgbw,
                                                 float (&rgb)[3]) {
  if (KIND == 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = __ffma2_rn(pb[2 * i], make_float2(kF8Scale, kF8Scale), v[2 * i]);
      const float2 c = __ffma2_rn(pb[2 * i + 1], make_float2(kF8Scale, kF8Scale), v[2 * i + 1]);
      st_shared_v4(park + (q * 8 + i) * PARK_STRIDE,
                   make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(c.x), __float_as_uint(c.y)));
    }
    return;
  }
  if (NOISE) {
#pragma unroll
    for (int i = 0; i < 16; ++i) pb[i] = __fadd2_rn(pb[i], nz2);
  }
  if (KIND == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = image_act<1>(v[i], pb[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = ld_shared_v4(park + (q * 8 + i) * PARK_STRIDE);
      v[2 * i] = __ffma2_rn(make_float2(__uint_as_float(u.x), __uint_as_float(u.y)), make_float2(kF8InvScale, kF8InvScale),
                            image_act<1>(v[2 * i], pb[2 * i]));
      v[2 * i + 1] = __ffma2_rn(make_float2(__uint_as_float(u.z), __uint_as_float(u.w)), make_float2(kF8InvScale, kF8InvScale),
                                image_act<1>(v[2 * i + 1], pb[2 * i + 1]));
    }
    if (blk == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 a = __fmul2_rn(v[2 * i], make_float2(kInvSqrt2 * kF8Scale, kInvSqrt2 * kF8Scale));
        const float2 c = __fmul2_rn(v[2 * i + 1], make_float2(kInvSqrt2 * kF8Scale, kInvSqrt2 * kF8Scale));
        st_shared_v4(park + (q * 8 + i) * PARK_STRIDE,
                     make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(c.x), __float_as_uint(c.y)));
      }
    }
  }
  if (KIND == 1 && blk == 3) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        a0 = fmaf(v[i].x, rgbw.w[c * 256 + col0 + 2 * i], a0);
        a1 = fmaf(v[i].y, rgbw.w[c * 256 + col0 + 2 * i + 1], a1);
      }
      rgb[c] += a0 + a1;
    }
  } else {
    publish_ts(tmem_lane, col0, v);
    tmem_st_wait();
    sig(q);
  }
}

// CODE SIZE is a first-order concern here: with every stage, quarter and call site unrolled the epilogue threads walked ~160 KB
// of straight-line code per tile and 35 % of their stall samples were instruction-cache misses
// (profiles/r02_image_ts_ncu_before_code_diet.md).  A stage is therefore a two-trip loop over accumulator halves (two quarters
// unrolled in the body), conv1 / conv2 share one call site, and the three conv3 flavours are one body with run-time switches.
// KIND 0: H = lrelu(acc + b)                                         (conv1 / conv2)
// KIND 1: y = lrelu(acc + b) + parked;  blk 2: parked <- y / sqrt2 (res4's identity skip);
//         blk < 3: H = y, else rgb += Wrgb y (ToRGB; nothing is published)
// KIND 2: parked <- acc + cs                                         (skip GEMM; cs = the folded scale-injection columns)
template <int KIND, int NOISE, class Sig, class WaitA, class WaitB>
__device__ __forceinline__ void image_ts_stage(uint32_t tmem_lane, uint32_t park, int sub, const float* __restrict__ vecs,
                                               Sig sig, WaitA wait_a, WaitB wait_b, float nz, int blk, bool more,
                                               const RgbParam& rgbw, float (&rgb)[3]) {
  // vecs: KIND 0 / 1 the bias, KIND 2 the skip constant (256 floats); this thread's 32 columns of quarter q start at q * 64 + sub * 32
  const float* vp = vecs + sub * 32;
  float2 pa[16], pb[16];                                  // vectors of the even / odd quarter: no copies between them
  load_vec<16>(vp, pa);                                   // quarter 0's: in flight while the thread parks on the MMA barrier
  const float2 nz2 = make_float2(nz, nz);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    if (h == 0) wait_a(); else wait_b();                  // COMMIT 0 / 1: output columns 0..127 / 128..255 are complete
    float2 v[2][16];
#pragma unroll
    for (int j = 0; j < 2; ++j) tmem_ld32(tmem_lane + (2 * h + j) * 64 + sub * 32, v[j]);
    tmem_ld_wait();
    sig(5 - h);                                           // this accumulator half is in registers
    if (KIND == 1 && h == 1 && blk == 3 && more) {        // the next tile's first groups read X only: let them start
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) sig(q2);
    }
    load_vec<16>(vp + (2 * h + 1) * 64, pb);              // in flight during the even quarter's conversion
    image_ts_quarter<KIND, NOISE>(tmem_lane, park, 2 * h, (2 * h) * 64 + sub * 32, v[0], pa, sig, nz2, blk, rgbw, rgb);
    if (h == 0) load_vec<16>(vp + 128, pa);               // quarter 2's, in flight during the odd quarter's conversion
    image_ts_quarter<KIND, NOISE>(tmem_lane, park, 2 * h + 1, (2 * h + 1) * 64 + sub * 32, v[1], pb, sig, nz2, blk, rgbw, rgb);
  }
}

// ---------------------------------------------------------------------------
// the fused image kernel
// ---------------------------------------------------------------------------
#ifdef DDMI_EXP_BIGRING
// experiment (timing only, results are garbage): the PE-feature region becomes ring space -> 12 slots = 96 KB of weights in flight
using ImgL = Layout<0, 98304>;
#else
using ImgL = Layout<8>;
#endif
// Plane patches for coherent query tiles (regular grids): when the bilinear taps of a tile's 128 coordinates fall inside a
// PATCH_W x PATCH_H texel window of the plane, ONE 3-D TMA load (box PATCH_W x PATCH_H x 64 channels = 32 KB, straight from the
// NCHW plane the VAE decoder emits) stages the window in the X region and every thread blends its row from shared memory --
// instead of 4 x 32 scalar L2 loads per thread (1024^2 / 2048^2 grids on 64^2 .. 256^2 planes: windows of 9 .. 34 x 2 texels).
// The coordinates are still taken literally from the caller's tensor: the window is the bounding box of the actual taps, and a
// tile whose taps do not fit (scattered queries, tiles that straddle image rows, native-resolution grids) gathers as before.
constexpr int PATCH_W = 64, PATCH_H = 2, PATCH_BYTES = PATCH_W * PATCH_H * 64 * 4;
static_assert(PATCH_BYTES <= 2 * 8 * KG_BYTES * 2, "the patch is staged in the X region");
constexpr int IMG_OFF_PBAR = TMEM_SLOT + 8;                        // inside the barrier block
constexpr int IMG_OFF_SCRATCH = BAR_BYTES;                         // 8 warps x {xmin, xmax, ymin, ymax}
constexpr int IMG_OFF_RGBX = IMG_OFF_SCRATCH + 128;                  // TS: [3][128] partial ToRGB sums of the sub = 1 threads
constexpr int IMG_SMEM = ImgL::SMEM_BYTES + 128 + 3 * 128 * 4;

template <int PAIR, int SCHEME, int NOISE, int TS = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
image_umma_kernel(PlaneSet ps, const float* __restrict__ cx, const float* __restrict__ cy, long long n,
                  int tiles_per_item, long long total_tiles, const uint8_t* __restrict__ wstream,
                  const __grid_constant__ ProgramParam prog, const float* __restrict__ vec, void* __restrict__ out, int store,
                  NoiseArgs na, const __grid_constant__ CUtensorMap pmap0, const __grid_constant__ CUtensorMap pmap1,
                  const __grid_constant__ CUtensorMap pmap2, int patch_mask, const __grid_constant__ RgbParam rgbw) {
  static_assert(!TS || (PAIR && SCHEME), "the TMEM-resident-activation kernel is f16f8 on CTA pairs");
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase + ImgL::KG_HHI * KG_BYTES, h_lo = sbase + ImgL::KG_HLO * KG_BYTES;
  const uint32_t x_hi = sbase + ImgL::KG_XHI * KG_BYTES, x_lo = sbase + ImgL::KG_XLO * KG_BYTES;
  const uint32_t ring = sbase + ImgL::OFF_RING;
  const uint32_t bar = sbase + ImgL::OFF_BAR;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int C = 64;
  const uint32_t pbar = bar + IMG_OFF_PBAR;               // plane-patch (TMA) completion barrier
  int* pscratch = reinterpret_cast<int*>(smem + ImgL::OFF_BAR + IMG_OFF_SCRATCH);
  if (tid == 32) mbar_init(pbar, 1);                      // published by engine_begin's barrier init fence + sync
  const uint32_t tmem = engine_begin<PAIR, SCHEME>(smem, ImgL::OFF_BAR);

  // work split: CTA (or CTA pair) w of W takes iterations w, w + W, ...; a pair iteration = tiles 2u and 2u + 1
  const long long nwork = PAIR ? (total_tiles + 1) / 2 : total_tiles;
  const long long wfirst = PAIR ? blockIdx.x / 2 : blockIdx.x, wstride = PAIR ? gridDim.x / 2 : gridDim.x;
  const long long ntiles = wfirst < nwork ? (nwork - wfirst + wstride - 1) / wstride : 0;
  auto tile_of = [&](long long i) { const long long u = wfirst + i * wstride; return PAIR ? 2 * u + rank : u; };

  if (warp < 8) {
    // =================== gather + epilogue threads ===================
    reg_inc<216>();
    const int row = tid & 127;                          // tile row == TMEM lane
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int sub = warp >> 2;                          // which 32 columns of each 64-column quarter
    const int ghalf = tid >> 7;                         // gather: channels [32*ghalf, +32)
    const uint32_t a_bar = PAIR ? mapa_rank(bar + BAR_A0, 0) : bar + BAR_A0;   // operand barriers live in the leader
    uint32_t ph_mma = 0;
    const bool prof = DDMI_PROFILE && (blockIdx.x == 0 && tid == 0);
    bool tr = false;                                    // this thread traces the current tile iteration (profiling build)
    uint32_t trn = 0;
    long long p_wait = 0, p_epi = 0, p_gather = 0, p_t = prof_clock();

    // `part` 0 / 1 = first / second 16 of this thread's 32 channels (-1: both): the two halves are issued in two different
    // idle windows of the epilogue threads (after the conv1 and after the conv2 epilogue, while conv2 / conv3 run on the
    // tensor core), so the L2-latency-bound gather delays neither epilogue by much
    // ---- staged path state: decided in the first gather window, used in the second
    bool pf_fits = false;
    int pf_x = 0, pf_y = 0;
    Tap pf_tap;
    uint32_t ph_patch = 0;
    // this thread's tap + the tile's tap window (bounding box over the 128 rows); true if a patch load covers it
    // this row's query position of a tile: the three scales of a tile share it, so it is fetched once per tile (an exposed
    // L2 round trip per scale otherwise), and `coord_prefetch` starts the next tile's fetch a stage before it is needed
    long long c_tile = -1;
    float c_x = 0.f, c_y = 0.f;
    auto coord_fetch = [&](long long tile) {
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      if (tile == c_tile) return;
      long long gi = (tile % tiles_per_item) * TILE + row;
      if (gi > n - 1) gi = n - 1;
      c_x = __ldg(cx + gi);
      c_y = __ldg(cy + gi);
      c_tile = tile;
    };
    auto patch_plan = [&](long long tile, int s) {
      coord_fetch(tile);
      const int W = ps.w[s];
      pf_tap = make_tap<false>(c_x, c_y, ps.h[s], W);
      const int x0 = pf_tap.o00 % W, y0 = pf_tap.o00 / W, x1 = pf_tap.o11 % W, y1 = pf_tap.o11 / W;
      const int xmn = __reduce_min_sync(0xffffffffu, x0), xmx = __reduce_max_sync(0xffffffffu, x1);
      const int ymn = __reduce_min_sync(0xffffffffu, y0), ymx = __reduce_max_sync(0xffffffffu, y1);
      if (lane == 0) {
        pscratch[warp * 4 + 0] = xmn; pscratch[warp * 4 + 1] = xmx; pscratch[warp * 4 + 2] = ymn; pscratch[warp * 4 + 3] = ymx;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int a = pscratch[0], bq = pscratch[1], c2 = pscratch[2], d = pscratch[3];
#pragma unroll
      for (int w8 = 1; w8 < 8; ++w8) {
        a = min(a, pscratch[w8 * 4]); bq = max(bq, pscratch[w8 * 4 + 1]);
        c2 = min(c2, pscratch[w8 * 4 + 2]); d = max(d, pscratch[w8 * 4 + 3]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // the scratch may be rewritten
      a &= ~3;                                            // TMA: the window's first texel must sit on a 16-byte boundary
      pf_x = a; pf_y = c2;
      pf_fits = ((patch_mask >> s) & 1) && (bq - a < PATCH_W) && (d - c2 < PATCH_H);
    };
    // X <- features of scale s blended from a TMA-staged plane window (all 32 channels of this thread)
    // the window load is issued in the first gather window (X is free from there on), the blend runs in the second
    auto patch_issue = [&](long long tile, int s) {
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      const int b = (int)(tile / tiles_per_item);
      if (tid == 0) {
        mbar_expect_tx(pbar, PATCH_BYTES);
        tma::load_3d(x_hi, s == 0 ? &pmap0 : (s == 1 ? &pmap1 : &pmap2), pf_x, pf_y, b * C, pbar);
      }
    };
    auto gather_staged = [&](long long tile, int s) {
      const long long g0 = prof_clock();
      trace(tr, 0x20, trn, 0);
      mbar_wait(pbar, ph_patch);
      ph_patch ^= 1;
      const int W = ps.w[s];
      const int rx0 = pf_tap.o00 % W - pf_x, ry0 = pf_tap.o00 / W - pf_y, rx1 = pf_tap.o11 % W - pf_x, ry1 = pf_tap.o11 / W - pf_y;
      const float* P = reinterpret_cast<const float*>(smem + ImgL::KG_XHI * KG_BYTES) + (size_t)(ghalf * 32) * (PATCH_H * PATCH_W);
      const int i00 = ry0 * PATCH_W + rx0, i01 = ry0 * PATCH_W + rx1, i10 = ry1 * PATCH_W + rx0, i11 = ry1 * PATCH_W + rx1;
      float y[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float* pc = P + c * (PATCH_H * PATCH_W);
        float acc = pc[i00] * pf_tap.w00;                   // the order of tap_sample(): results are bit-identical to the direct path
        acc = fmaf(pc[i01], pf_tap.w01, acc);
        acc = fmaf(pc[i10], pf_tap.w10, acc);
        acc = fmaf(pc[i11], pf_tap.w11, acc);
        y[c] = acc;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");      // every thread has read the window: X may be overwritten
      if (SCHEME) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float2 y2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y2[i] = make_float2(y[g * 16 + 2 * i], y[g * 16 + 2 * i + 1]);
          uint4 a16[2], r8, a8;
          split16_f16f8(y2, a16, r8, a8);
          const uint32_t off = (uint32_t)(ghalf * 4 * KG_BYTES + row * 16);
          st_shared_v4(x_hi + off + (2 * g) * KG_BYTES, a16[0]);
          st_shared_v4(x_hi + off + (2 * g + 1) * KG_BYTES, a16[1]);
          st_shared_v4(x_lo + off + g * KG_BYTES, r8);
          st_shared_v4(x_lo + off + (2 + g) * KG_BYTES, a8);
        }
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y8[i] = y[g * 8 + i];
          uint4 hi, lo;
          split8(y8, hi, lo);
          const uint32_t off = (uint32_t)((ghalf * 4 + g) * KG_BYTES + row * 16);
          st_shared_v4(x_hi + off, hi);
          st_shared_v4(x_lo + off, lo);
        }
      }
      p_gather += prof_clock() - g0;
      trace(tr, 0x21, trn, 0);
    };
    auto gather_direct = [&](long long tile, int s, int part) {
      const long long g0 = prof_clock();
      if (dbg(4)) return;
      trace(tr, 0x20, trn, 0);
      if (tile > total_tiles - 1) tile = total_tiles - 1;   // odd tail of a pair: decode a duplicate, store nothing
      const int b = (int)(tile / tiles_per_item);
      long long gi = (tile % tiles_per_item) * TILE + row;
      if (gi > n - 1) gi = n - 1;
      const Tap t = make_tap<false>(__ldg(cx + gi), __ldg(cy + gi), ps.h[s], ps.w[s]);
      const size_t hw = (size_t)ps.h[s] * ps.w[s];
      const float* base = ps.data[s] + ((size_t)b * C + ghalf * 32) * hw;
      if (SCHEME) {
        // f16f8: this thread's 32 channels = K step `ghalf` of X: fp16 groups 4*ghalf.., FP8 groups [r8 r8 a8 a8] at x_lo
#pragma unroll 1
        for (int g = (part == 1 ? 1 : 0); g < (part == 0 ? 1 : 2); ++g) {
          float2 y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            y[i].x = tap_sample(base + (size_t)(g * 16 + 2 * i) * hw, t);
            y[i].y = tap_sample(base + (size_t)(g * 16 + 2 * i + 1) * hw, t);
          }
          uint4 a16[2], r8, a8;
          split16_f16f8(y, a16, r8, a8);
          const uint32_t off = (uint32_t)(ghalf * 4 * KG_BYTES + row * 16);
          st_shared_v4(x_hi + off + (2 * g) * KG_BYTES, a16[0]);
          st_shared_v4(x_hi + off + (2 * g + 1) * KG_BYTES, a16[1]);
          st_shared_v4(x_lo + off + g * KG_BYTES, r8);
          st_shared_v4(x_lo + off + (2 + g) * KG_BYTES, a8);
        }
      } else {
#pragma unroll 1
        for (int g = (part == 1 ? 2 : 0); g < (part == 0 ? 2 : 4); ++g) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = tap_sample(base + (size_t)(g * 8 + i) * hw, t);
          uint4 hi, lo;
          split8(y, hi, lo);
          const uint32_t off = (uint32_t)((ghalf * 4 + g) * KG_BYTES + row * 16);
          st_shared_v4(x_hi + off, hi);
          st_shared_v4(x_lo + off, lo);
        }
      }
      p_gather += prof_clock() - g0;
      trace(tr, 0x21, trn, 0);
    };
    // first window (part 0): decide; direct gathers take their first half now, staged ones wait for the second window
    // (the X region is the staging buffer, and one TMA round trip + 32 channels from shared memory fit there easily)
    auto gather = [&](long long tile, int s, int part) {
      if (part != 1) {
        const long long g0 = prof_clock();
        patch_plan(tile, s);
        if (pf_fits) patch_issue(tile, s);
        p_gather += prof_clock() - g0;
      }
      if (pf_fits) {
        if (part != 0) gather_staged(tile, s);
      } else {
        gather_direct(tile, s, part);
      }
      if (TS) fence_proxy_async();                        // TS signals carry no proxy fence of their own (H is in TMEM)
    };
    // make this warp's smem / TMEM writes visible to the MMA warp (of the leader CTA), then signal one quarter
    auto signal = [&](int q) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(a_bar + 8 * q);
      trace(tr, 0x10 + q, trn, 0);
      if (q == 3) p_epi += prof_clock() - p_t;
    };
    // the second accumulator half is in registers: operand barrier 4 (the program waits it before it writes columns 128..255)
    auto drained = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(a_bar + 8 * 4);
      trace(tr, 0x14, trn, 0);
    };
    auto signal_all = [&]() {
      drained();
#pragma unroll
      for (int q = 0; q < 4; ++q) signal(q);
    };
    auto wait_mma = [&]() {                               // COMMIT 0: output columns 0..127 (or the whole group) are complete
      const long long w0 = prof_clock();
      trace(tr, 0x01, trn, 0);
      mbar_wait(bar + BAR_MMADONE, ph_mma);
      tc_fence_after();
      trace(tr, 0x02, trn, 0);
      p_t = prof_clock();
      p_wait += p_t - w0;
    };
    auto wait_b = [&]() {                                 // COMMIT 1: columns 128..255 too
      const long long w0 = prof_clock();
      mbar_wait(bar + BAR_MMADONE + 8, ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      trace(tr, 0x04, trn, 0);
      const long long w1 = prof_clock();
      p_wait += w1 - w0;
      p_t += w1 - w0;
    };

    if constexpr (TS) {
      // ---- TMEM-resident activation: 15 GEMM groups per tile (skip, conv1, conv2, conv3 per block; res4 has no skip)
      const uint32_t park = sbase + (uint32_t)tid * 16;
      float* rgbx = reinterpret_cast<float*>(smem + ImgL::OFF_BAR + IMG_OFF_RGBX);
      auto sig = [&](int i) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(a_bar + 8 * i);
        trace(tr, 0x10 + i, trn, 0);
      };
      if (ntiles > 0) {
        gather(tile_of(0), 0, -1);
        sig(5);
        sig(4);
#pragma unroll
        for (int q = 0; q < 4; ++q) sig(q);
      }
      for (long long it = 0; it < ntiles; ++it) {
        const long long tile = tile_of(it);
        const float* bv = vec;
        const bool more = it + 1 < ntiles;
        tr = prof && it == kTraceIter;
        float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk, bv += 1024) {
          float nz[3] = {0.f, 0.f, 0.f};
          if (NOISE) {
            const long long tt = tile < total_tiles ? tile : total_tiles - 1;
            long long gi = (tt % tiles_per_item) * TILE + row;
            if (gi > n - 1) gi = n - 1;
            noise_block3(na, blk, (size_t)(tt / tiles_per_item), n, gi, nz);
#pragma unroll
            for (int j = 0; j < 3; ++j) nz[j] *= __ldg(vec + 4096 + 768 + 3 + 3 * blk + j);
          }
          if (blk == 2 && more) coord_fetch(tile_of(it + 1));     // in flight during the skip stage; used after conv1
          // ---- skip GEMM of the block -> parked (+ the block's skip constant)
          if (blk < 3) image_ts_stage<2, NOISE>(tmem_lane, park, sub, bv + 768, sig, wait_mma, wait_b, 0.f, blk, more, rgbw, rgb);
          // ---- conv1, conv2; after each, one half of the next PE scale (or of the next tile's coarse scale) is gathered:
          // X is free once conv1's group is complete
#pragma unroll 1
          for (int cv = 0; cv < 2; ++cv) {
            image_ts_stage<0, NOISE>(tmem_lane, park, sub, bv + 256 * cv, sig, wait_mma, wait_b, cv ? nz[1] : nz[0], blk, more, rgbw, rgb);
            if (blk < 2 || (blk == 2 && more)) gather(blk < 2 ? tile : tile_of(it + 1), blk < 2 ? blk + 1 : 0, cv);
          }
          // ---- conv3 + parked skip (res4: + ToRGB)
          image_ts_stage<1, NOISE>(tmem_lane, park, sub, bv + 512, sig, wait_mma, wait_b, nz[2], blk, more, rgbw, rgb);
        }
        // ---- ToRGB: the two threads of a row add their partial sums
        if (sub == 1) {
#pragma unroll
          for (int c = 0; c < 3; ++c) rgbx[c * 128 + row] = rgb[c];
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (sub == 0 && tile < total_tiles) {
          const int b = (int)(tile / tiles_per_item);
          const long long gi = (tile % tiles_per_item) * TILE + row;
          if (gi < n) {
            const float* brgb = vec + 4096 + 768;
#pragma unroll
            for (int c = 0; c < 3; ++c) store_rgb(out, store, b, n, gi, c, rgb[c] + rgbx[c * 128 + row] + __ldg(brgb + c));
          }
        }
      }
    } else {
    if (ntiles > 0) gather(tile_of(0), 0, -1);
    if (ntiles > 0) signal_all();
    for (long long it = 0; it < ntiles; ++it) {
      const long long tile = tile_of(it);
      const float* bv = vec;
      tr = prof && it == kTraceIter;
#pragma unroll 1
      for (int blk = 0; blk < 4; ++blk, bv += 1024) {
        // noise of this block's three StyledConvs at this thread's row, scaled by noise.weight (x the folded gain)
        float nz[3] = {0.f, 0.f, 0.f};
        if (NOISE) {
          const long long tt = tile < total_tiles ? tile : total_tiles - 1;
          long long gi = (tt % tiles_per_item) * TILE + row;
          if (gi > n - 1) gi = n - 1;
          noise_block3(na, blk, (size_t)(tt / tiles_per_item), n, gi, nz);
#pragma unroll
          for (int j = 0; j < 3; ++j) nz[j] *= __ldg(vec + 4096 + 768 + 3 + 3 * blk + j);
        }
        // ---- conv1 (+ skip into acc2 for blk < 3)
        StagePrefetch pf;
        stage_prefetch<false>(pf, bv, nullptr, sub);
        wait_mma();
        image_stage<0, SCHEME, NOISE>(tmem_lane, h_hi, h_lo, row, sub, bv, nullptr, pf, signal, wait_b, drained, nz[0], tr, trn);
        // the PE buffer is free now: prefetch the next scale (or the next tile's coarse scale), first half of the channels
        if (blk < 2) gather(tile, blk + 1, 0);
        else if (blk == 2 && it + 1 < ntiles) gather(tile_of(it + 1), 0, 0);
        // ---- conv2
        stage_prefetch<false>(pf, bv + 256, nullptr, sub);
        wait_mma();
        image_stage<0, SCHEME, NOISE>(tmem_lane, h_hi, h_lo, row, sub, bv + 256, nullptr, pf, signal, wait_b, drained, nz[1], tr, trn);
        if (blk < 2) gather(tile, blk + 1, 1);                                   // second half, under conv3's GEMM
        else if (blk == 2 && it + 1 < ntiles) gather(tile_of(it + 1), 0, 1);
        // ---- conv3 + skip
        if (blk < 3) stage_prefetch<true>(pf, bv + 512, bv + 768, sub);
        else stage_prefetch<false>(pf, bv + 512, nullptr, sub);
        wait_mma();
        if (blk < 2) image_stage<1, SCHEME, NOISE>(tmem_lane, h_hi, h_lo, row, sub, bv + 512, bv + 768, pf, signal, wait_b, drained, nz[2], tr, trn);
        else if (blk == 2) image_stage<2, SCHEME, NOISE>(tmem_lane, h_hi, h_lo, row, sub, bv + 512, bv + 768, pf, signal, wait_b, drained, nz[2], tr, trn);
        else image_stage<3, SCHEME, NOISE>(tmem_lane, h_hi, h_lo, row, sub, bv + 512, nullptr, pf, signal, wait_b, drained, nz[2], tr, trn);
      }
      // ---- ToRGB: acc1[:, 0:16]
      wait_mma();
      wait_b();
      if (sub == 0 && tile < total_tiles) {
        float2 v[8];
        tmem_ld16(tmem_lane, v);   // only columns 0..2 are meaningful
        tmem_ld_wait();
        const int b = (int)(tile / tiles_per_item);
        const long long gi = (tile % tiles_per_item) * TILE + row;
        if (gi < n) {
          const float* brgb = vec + 4096 + 768;
          constexpr float kOut = SCHEME ? kF8InvScale : 1.0f;
          store_rgb(out, store, b, n, gi, 0, v[0].x * kOut + __ldg(brgb + 0));
          store_rgb(out, store, b, n, gi, 1, v[0].y * kOut + __ldg(brgb + 1));
          store_rgb(out, store, b, n, gi, 2, v[1].x * kOut + __ldg(brgb + 2));
        }
      }
      if (it + 1 < ntiles) signal_all();
    }
    }
    if (prof) {
      prof_add(0, p_wait);
      prof_add(1, p_epi);
      prof_add(2, p_gather);
    }
  } else {
    if constexpr (TS) {
      // producer and forwarder interpret the op table (byte counts only); the issuer is image_ts_issuer.cuh's straight-line code
      reg_dec<72>();
      if (warp == 8) producer_loop<PAIR, ImgL::RING_BYTES, SCHEME>(prog.op, wstream, ring, bar, ntiles, rank);
      else if (warp == 9 && rank == 0) image_ts_issue_loop(sbase, ring, bar, tmem, ImgL::KG_XHI, ImgL::KG_XLO, ntiles);
      else if (warp == 9) forward_loop<ImgL::RING_BYTES, SCHEME>(prog.op, bar, ntiles);
    } else {
      engine_service_warps<PAIR, ImgL::RING_BYTES, SCHEME, 0>(prog.op, wstream, sbase, ring, bar, tmem, ntiles, rank);
    }
  }
  engine_end<PAIR>(tmem);
}

// ---------------------------------------------------------------------------
// bring-up self test: D[128 x N] = A[128 x K] * B[N x K]^T through the same
// descriptors, split, MMA and TMEM load paths as the decode kernels.
// ---------------------------------------------------------------------------
constexpr int ST_H_BYTES = H_KG * KG_BYTES;
constexpr int ST_OFF_AHI = 0, ST_OFF_ALO = ST_H_BYTES, ST_OFF_B = 2 * ST_H_BYTES, ST_OFF_BAR = ST_OFF_B + 16384;
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
                      const void* gemm, size_t gemm_bytes, const uint32_t* program_host, size_t program_words,
                      const uint32_t* program_dev, const float* vec, size_t vec_floats, void* out, int store, int pair,
                      int f16f8, const NoiseArgs& na, int no_patch, int ts, const float* vec_host, cudaStream_t st) {
  using namespace ummak;
  if (C != 64) {
    set_error("tcgen05 image kernel is built for 64-channel planes");
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(program_host && program_dev && program_words >= 2, "bf16x3 weights carry no MMA program");
  const long long need = ummak::program_stream_bytes(program_host, program_words);
  DDMI_REQUIRE(need > 0 && (size_t)need == gemm_bytes, "MMA program consumes %lld weight bytes but the stream has %zu",
               need, gemm_bytes);
  ProgramParam pp;
  DDMI_REQUIRE(make_program_param(program_host, program_words, &pp), "MMA program has %zu words, at most %d fit the kernel parameter",
               program_words, PROG_MAX);
  DDMI_REQUIRE(!f16f8 || pair, "the f16f8 image kernel runs as CTA pairs only");
  DDMI_REQUIRE(!ts || (f16f8 && vec_host), "a TMEM-resident-activation program needs DDMI_PREC_F16F8 and weights->vec_host");
  RgbParam rgbv = {};     // ToRGB weights as a launch parameter (constant bank); unused (zero) by the other programs
  if (ts) {
    // the issuer of the TS kernel is the op table written out as code (image_ts_issuer.cuh): refuse any other table
    uint32_t h = 0x811C9DC5u;
    size_t nops = 0;
    while (nops < program_words && (program_host[nops] & 3) != OP_END) ++nops;
    for (size_t i = 0; i < nops; ++i)
      for (int b = 0; b < 4; ++b) h = (h ^ ((program_host[i] >> (8 * b)) & 0xFF)) * 0x01000193u;
    DDMI_REQUIRE(nops == kImageTsProgramOps && h == kImageTsProgramHash,
                 "the MMA program (%zu ops, FNV-1a %08x) is not the one image_ts_issuer.cuh implements (%d ops, %08x): "
                 "packing._pack_image_ts and the issuer must change together", nops, h, kImageTsProgramOps, kImageTsProgramHash);
    memcpy(rgbv.w, vec_host + 4096, sizeof(rgbv.w));
  }
  DDMI_REQUIRE(vec_floats == 4096 + 768 + 3 + 12, "packed vec blob is %zu floats, expected 4879", vec_floats);
  int dev = 0, sms = 0;
  DDMI_CUDA(cudaGetDevice(&dev));
  DDMI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tpi = (n + TILE - 1) / TILE;
  const long long total = tpi * batch;
  if (tpi > 2147483647LL) {
    set_error("n_coords %lld too large for one launch", n);
    return DDMI_ERR_UNSUPPORTED;
  }
  // CTA pairs: a 2-CTA cluster per pair of tiles (weights for the pair are packed [half 0 | half 1] per K step)
  const uint8_t* ws = (const uint8_t*)gemm;
  const int tpi_i = (int)tpi;
  const long long work = (total + 1) / 2, npairs = work < sms / 2 ? work : sms / 2;
  const unsigned ctas = pair ? (unsigned)(2 * npairs) : (unsigned)(total < sms ? total : sms);
  // plane windows through TMA where the plane layout allows a tensor map (16-byte aligned base, width a multiple of 4)
  CUtensorMap pm[3];
  int patch_mask = 0;
  memset(pm, 0, sizeof(pm));
  if (!no_patch) {
    for (int s = 0; s < 3; ++s)
      if (tma::make_plane_map(&pm[s], ps.data[s], batch, C, ps.h[s], ps.w[s], PATCH_W, PATCH_H, 64)) patch_mask |= 1 << s;
  }
#define DDMI_IMG_LAUNCH(P, S, Z)                                                                                          \
  DDMI_CUDA(launch_engine(image_umma_kernel<P, S, Z>, P, ctas, IMG_SMEM, st, ps, cx, cy, n, tpi_i, total, ws, pp, vec, out, \
                          store, na, pm[0], pm[1], pm[2], patch_mask, rgbv))
#define DDMI_IMG_LAUNCH_TS(Z)                                                                                              \
  DDMI_CUDA(launch_engine(image_umma_kernel<1, 1, Z, 1>, 1, ctas, IMG_SMEM, st, ps, cx, cy, n, tpi_i, total, ws, pp, vec, out, \
                          store, na, pm[0], pm[1], pm[2], patch_mask, rgbv))
  const int nz = na.mode != 0;
  if (ts && nz) { DDMI_IMG_LAUNCH_TS(1); }
  else if (ts) { DDMI_IMG_LAUNCH_TS(0); }
  else if (pair && f16f8 && nz) { DDMI_IMG_LAUNCH(1, 1, 1); }
  else if (pair && f16f8) { DDMI_IMG_LAUNCH(1, 1, 0); }
  else if (pair && nz) { DDMI_IMG_LAUNCH(1, 0, 1); }
  else if (pair) { DDMI_IMG_LAUNCH(1, 0, 0); }
  else if (nz) { DDMI_IMG_LAUNCH(0, 0, 1); }
  else { DDMI_IMG_LAUNCH(0, 0, 0); }
#undef DDMI_IMG_LAUNCH
#undef DDMI_IMG_LAUNCH_TS
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_occupancy_umma_entry(const PlaneSet& ps, int batch, int C, const float* pts, long long n, long long batch_stride,
                                float divisor, float upper, const void* gemm, size_t gemm_bytes,
                                const uint32_t* program_host, size_t program_words, const uint32_t* program_dev,
                                const float* vec, size_t vec_floats, float* logits, int pair, int nhwc, int f16f8,
                                cudaStream_t st) {
  return launch_occupancy_umma(ps, batch, C, pts, n, batch_stride, divisor, upper, gemm, gemm_bytes, program_host,
                               program_words, program_dev, vec, vec_floats, logits, pair, nhwc, f16f8, st);
}

int launch_nerf_umma_entry(const PlaneSet& ps, int batch, int C, const float* rays, long long n_rays, int ray_stride,
                           const float* t_vals, int z_stride, int n_samples, float plane_extent, float slope, int white_bkgd,
                           const void* gemm, size_t gemm_bytes, const uint32_t* program_host, size_t program_words,
                           const uint32_t* program_dev, const float* vec, size_t vec_floats, float* rgb_map, float* raw,
                           int fuse, int f16f8, cudaStream_t st) {
  return launch_nerf_umma(ps, batch, C, rays, n_rays, ray_stride, t_vals, z_stride, n_samples, plane_extent, slope, white_bkgd, gemm,
                          gemm_bytes, program_host, program_words, program_dev, vec, vec_floats, rgb_map, raw, fuse, f16f8, st);
}

int launch_video_umma_entry(const PlaneSet& ps, int batch, int C, const float* cxy, const float* cyt, const float* cxt, int T,
                            int H, int W, const void* gemm, size_t gemm_bytes, const uint32_t* program_host,
                            size_t program_words, const uint32_t* program_dev, const float* vec, size_t vec_floats,
                            void* out, int store, int pair, int f16f8, void* workspace, size_t workspace_bytes,
                            cudaStream_t st) {
  return launch_video_umma(ps, batch, C, cxy, cyt, cxt, T, H, W, gemm, gemm_bytes, program_host, program_words, program_dev,
                           vec, vec_floats, out, store, pair, f16f8, workspace, workspace_bytes, st);
}
size_t video_workspace_bytes(int batch, int T, int H, int W) { return video_table_bytes(batch, T, H, W); }

size_t occupancy_lattice_workspace_bytes(int batch, int nx, int ny, int nz) { return occupancy_lattice_bytes(batch, nx, ny, nz); }
int launch_occupancy_lattice_umma_entry(const PlaneSet& ps, int batch, int C, const float* axes, int nx, int ny, int nz,
                                        float divisor, float upper, const void* gemm, size_t gemm_bytes,
                                        const uint32_t* program_host, size_t program_words, const uint32_t* program_dev,
                                        const float* vec, size_t vec_floats, float* logits, int pair, int nhwc, int f16f8,
                                        void* workspace, size_t workspace_bytes, cudaStream_t st) {
  ummak::OccLattice lat = {nullptr, axes, nx, ny, nz};
  return launch_occupancy_umma(ps, batch, C, nullptr, (long long)nx * ny * nz, 0, divisor, upper, gemm, gemm_bytes, program_host,
                               program_words, program_dev, vec, vec_floats, logits, pair, nhwc, f16f8, st, &lat, workspace,
                               workspace_bytes);
}

int debug_trace(unsigned long long* out, int cap, int* n, int reset) {
  // the whole buffer (unwritten slots are 0: the caller drops them)
  int cnt = ummak::kTraceCap < cap ? ummak::kTraceCap : cap;
  DDMI_CUDA(cudaMemcpyFromSymbol(out, ummak::g_trace, sizeof(unsigned long long) * cnt));
  *n = cnt;
  if (reset) {
    static unsigned long long zeros[ummak::kTraceCap] = {};
    DDMI_CUDA(cudaMemcpyToSymbol(ummak::g_trace, zeros, sizeof(zeros)));
  }
  return DDMI_OK;
}

int debug_set(int flags) {
  DDMI_CUDA(cudaMemcpyToSymbol(ummak::g_dbg, &flags, sizeof(flags)));
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
