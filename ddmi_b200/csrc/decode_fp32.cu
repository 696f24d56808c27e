// fp32 CUDA-core decode kernels: the exact-arithmetic path (DDMI_PREC_FP32).
//
// One CTA = one tile of 64 query rows; the whole decoder (plane gather ->
// every MLP layer -> output) runs inside the CTA with activations resident in
// shared memory, so HBM sees only planes (L2-resident), coordinates and
// outputs.  Layer weights ([K][N] fp32, host-folded, packed by
// ddmi_b200/packing.py) stream through a double-buffered cp.async ring.
//
// Reference math restated here (paths relative to the reference checkout):
//   image      models/d2c_vae/mlp.py:34-66, blocks.py:187-283,312-356,604-638
//   occupancy  models/d2c_vae/mlp.py:82-111, blocks.py:673-716
//   video      models/d2c_vae/mlp.py:128-157, utils/general_utils.py:134-145
//   nerf       models/d2c_vae/mlp.py:241-281, utils/nerf_helpers.py:296-530
#include "common.cuh"

namespace ddmi {
namespace fp32 {

constexpr int TM = 64;     // rows per tile
constexpr int NT = 256;    // threads per CTA
constexpr int HS = 260;    // row stride (floats) of the 256-wide activation buffers
constexpr int KC = 16;     // K rows per streamed weight chunk
constexpr int WBUF = KC * 256;  // floats per ring slot

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[4][4*NJ] += A[64 x K] * Wt[K x 64*NJ].  Thread (ty, tx) owns rows
// ty*4..+3 and columns j*64 + tx*4..+3.  A lives in shared memory (row stride
// lda, multiple of 4); Wt is global, row-major [K][N], K a multiple of 16.
// Ends with a __syncthreads(): on return every thread is done reading A/wbuf.
template <int NJ>
__device__ __forceinline__ void gemm_seg(float (&acc)[4][4 * NJ], const float* __restrict__ As,
                                         int lda, int K, bool relu_in,
                                         const float* __restrict__ Wg, float* wbuf) {
  constexpr int N = 64 * NJ;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nchunk = K / KC;
  auto issue = [&](int c) {
    const float* src = Wg + (size_t)c * KC * N;
    float* dst = wbuf + (c & 1) * WBUF;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      int p = tid + NT * i;  // 16-byte piece index
      cp_async16(dst + p * 4, src + p * 4);
    }
    cp_async_commit();
  };
  issue(0);
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) {
      issue(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * WBUF;
    const float* a0 = As + (ty * 4) * lda + c * KC;
#pragma unroll
    for (int q = 0; q < KC / 4; ++q) {
      float a[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = *reinterpret_cast<const float4*>(a0 + i * lda + q * 4);
        if (relu_in) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        a[i][0] = v.x; a[i][1] = v.y; a[i][2] = v.z; a[i][3] = v.w;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float* wrow = wb + (q * 4 + kk) * N + tx * 4;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float4 w = *reinterpret_cast<const float4*>(wrow + j * 64);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][j * 4 + 0] = fmaf(a[i][kk], w.x, acc[i][j * 4 + 0]);
            acc[i][j * 4 + 1] = fmaf(a[i][kk], w.y, acc[i][j * 4 + 1]);
            acc[i][j * 4 + 2] = fmaf(a[i][kk], w.z, acc[i][j * 4 + 2]);
            acc[i][j * 4 + 3] = fmaf(a[i][kk], w.w, acc[i][j * 4 + 3]);
          }
        }
      }
    }
    __syncthreads();
  }
}

template <int NJ>
__device__ __forceinline__ void zero_acc(float (&acc)[4][4 * NJ]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4 * NJ; ++j) acc[i][j] = 0.f;
}

// dst[row][col] = f(row, col, acc) for the thread's 4 x 4NJ patch.
template <int NJ, class F>
__device__ __forceinline__ void store_acc(const float (&acc)[4][4 * NJ], float* dst, int ldd, F f) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = ty * 4 + i;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int col = j * 64 + tx * 4;
      float4 o;
      o.x = f(row, col + 0, acc[i][j * 4 + 0]);
      o.y = f(row, col + 1, acc[i][j * 4 + 1]);
      o.z = f(row, col + 2, acc[i][j * 4 + 2]);
      o.w = f(row, col + 3, acc[i][j * 4 + 3]);
      *reinterpret_cast<float4*>(dst + row * ldd + col) = o;
    }
  }
}

// ===========================================================================
// image: MLP.forward
// ===========================================================================
constexpr int IMG_XS = 68;
constexpr size_t IMG_SMEM = (size_t)(2 * TM * HS + TM * IMG_XS + 2 * WBUF) * sizeof(float);

__global__ void __launch_bounds__(NT, 1)
image_kernel(PlaneSet ps, int C, const float* __restrict__ cx, const float* __restrict__ cy,
             long long n, int tiles_per_item, const float* __restrict__ Wg,
             const float* __restrict__ vec, void* __restrict__ out, int store, NoiseArgs na) {
  extern __shared__ float4 smem4[];
  __shared__ float NZ[3 * TM];        // noise.weight * noise of the block's three StyledConvs, per tile row
  float* H = reinterpret_cast<float*>(smem4);
  float* Hn = H + TM * HS;
  float* X = Hn + TM * HS;
  float* wbuf = X + TM * IMG_XS;
  const int tid = threadIdx.x;
  const int b = blockIdx.x / tiles_per_item;
  const long long n0 = (long long)(blockIdx.x % tiles_per_item) * TM;
  const float kSqrt2 = 1.41421356237309504880f, kInvSqrt2 = 0.70710678118654752440f;

  // query position of this thread's gather row (rows past the end replay the last one)
  const int gr = tid & 63, gg = tid >> 6;
  long long gi = n0 + gr;
  if (gi > n - 1) gi = n - 1;
  const float gx = __ldg(cx + gi), gy = __ldg(cy + gi);
  const int cpg = C / 4;
  auto gather = [&](int s) {
    Tap t = make_tap<false>(gx, gy, ps.h[s], ps.w[s]);
    size_t hw = (size_t)ps.h[s] * ps.w[s];
    const float* base = ps.data[s] + ((size_t)b * C + gg * cpg) * hw;
    for (int c = 0; c < cpg; ++c) X[gr * IMG_XS + gg * cpg + c] = tap_sample(base + c * hw, t);
  };

  const float* wp = Wg;
  float acc[4][16];
  for (int blk = 0; blk < 4; ++blk) {
    const bool hasH = blk > 0, hasX = blk < 3;
    const float* bv = vec + blk * 1024;
    if (na.mode != 0 && tid < TM) {
      long long g = n0 + tid;
      if (g > n - 1) g = n - 1;
      float z[3];
      noise_block3(na, blk, (size_t)b, n, g, z);
      for (int j = 0; j < 3; ++j) NZ[j * TM + tid] = z[j] * __ldg(vec + 4096 + 768 + 3 + 3 * blk + j);
    } else if (tid < TM) {
      for (int j = 0; j < 3; ++j) NZ[j * TM + tid] = 0.f;
    }
    if (hasX) gather(blk);
    __syncthreads();
    // conv1
    zero_acc<4>(acc);
    if (hasH) { gemm_seg<4>(acc, H, HS, 256, false, wp, wbuf); wp += 256 * 256; }
    if (hasX) { gemm_seg<4>(acc, X, IMG_XS, C, false, wp, wbuf); wp += C * 256; }
    store_acc<4>(acc, Hn, HS, [&](int row, int col, float v) { return kSqrt2 * lrelu(v + NZ[row] + __ldg(bv + col), 0.2f); });
    __syncthreads();
    // conv2
    zero_acc<4>(acc);
    gemm_seg<4>(acc, Hn, HS, 256, false, wp, wbuf); wp += 256 * 256;
    store_acc<4>(acc, Hn, HS, [&](int row, int col, float v) { return kSqrt2 * lrelu(v + NZ[TM + row] + __ldg(bv + 256 + col), 0.2f); });
    __syncthreads();
    // conv3 (its sqrt2 cancels the block's 1/sqrt2)
    zero_acc<4>(acc);
    gemm_seg<4>(acc, Hn, HS, 256, false, wp, wbuf); wp += 256 * 256;
    store_acc<4>(acc, Hn, HS, [&](int row, int col, float v) { return lrelu(v + NZ[2 * TM + row] + __ldg(bv + 512 + col), 0.2f); });
    // skip (weights pre-scaled by 1/sqrt2 on the host); own-element reads of Hn only
    if (blk < 3) {
      zero_acc<4>(acc);
      if (hasH) { gemm_seg<4>(acc, H, HS, 256, false, wp, wbuf); wp += 256 * 256; }
      gemm_seg<4>(acc, X, IMG_XS, C, false, wp, wbuf); wp += C * 256;
      store_acc<4>(acc, H, HS, [&](int row, int col, float v) { return v + __ldg(bv + 768 + col) + Hn[row * HS + col]; });
    } else {
      zero_acc<4>(acc);
      store_acc<4>(acc, H, HS, [&](int row, int col, float) { return H[row * HS + col] * kInvSqrt2 + Hn[row * HS + col]; });
    }
    __syncthreads();
  }
  // ToRGB: out[b, c, n] = Wrgb[c] . H[row] + brgb[c]
  const float* wrgb = vec + 4096;
  const int r = tid >> 2, c = tid & 3;
  if (c < 3 && n0 + r < n) {
    float s = 0.f;
    for (int k = 0; k < 256; ++k) s = fmaf(H[r * HS + k], __ldg(wrgb + c * 256 + k), s);
    store_rgb(out, store, b, n, n0 + r, c, s + __ldg(wrgb + 768 + c));
  }
}

// ===========================================================================
// ResnetBlockFC chain shared by occupancy (MLP3D) and video (MLPVideo)
// ===========================================================================
// Packed gemm order: R1 fc0[KX x NH1], sc[KX x 256], fc1[NH1 x 256];
// R2/R3: fc0[256 x 256][KX x 256], sc[256 x 256][KX x 256], fc1[256 x 256];
// R4: fc0[256 x 256], fc1[256 x 256].   NH1 = 64*NJ1 = min(KX, 256).
// vec (bias part): R1 b0[NH1], b1[256]; R2 b0,b1; R3 b0,b1; R4 b0,b1 (256 each).
template <int KX, int NJ1, class Gather, class R1Extra>
__device__ __forceinline__ const float* resnet_chain(float* H, float* Hn, float* X, float* wbuf,
                                                     const float* __restrict__ Wg,
                                                     const float* __restrict__ vec, Gather gather,
                                                     R1Extra r1_extra) {
  constexpr int XS = KX + 4;
  constexpr int NH1 = 64 * NJ1;
  const float* wp = Wg;
  const float* bv = vec;
  float acc[4][16];
  // ---- R1: x = X0
  gather(0);
  __syncthreads();
  {
    float a1[4][4 * NJ1];
    zero_acc<NJ1>(a1);
    gemm_seg<NJ1>(a1, X, XS, KX, true, wp, wbuf); wp += KX * NH1;
    store_acc<NJ1>(a1, Hn, HS, [&](int, int col, float v) { return fmaxf(v + __ldg(bv + col), 0.f); });
    __syncthreads();
  }
  zero_acc<4>(acc);
  gemm_seg<4>(acc, X, XS, KX, false, wp, wbuf); wp += KX * 256;
  gemm_seg<4>(acc, Hn, HS, NH1, false, wp, wbuf); wp += NH1 * 256;
  store_acc<4>(acc, H, HS, [&](int row, int col, float v) { return v + __ldg(bv + NH1 + col) + r1_extra(row, col); });
  bv += NH1 + 256;
  __syncthreads();
  // ---- R2, R3: x = [H, X_s]
  for (int s = 1; s < 3; ++s) {
    gather(s);
    __syncthreads();
    zero_acc<4>(acc);
    gemm_seg<4>(acc, H, HS, 256, true, wp, wbuf); wp += 256 * 256;
    gemm_seg<4>(acc, X, XS, KX, true, wp, wbuf); wp += KX * 256;
    store_acc<4>(acc, Hn, HS, [&](int, int col, float v) { return fmaxf(v + __ldg(bv + col), 0.f); });
    __syncthreads();
    zero_acc<4>(acc);
    gemm_seg<4>(acc, H, HS, 256, false, wp, wbuf); wp += 256 * 256;
    gemm_seg<4>(acc, X, XS, KX, false, wp, wbuf); wp += KX * 256;
    gemm_seg<4>(acc, Hn, HS, 256, false, wp, wbuf); wp += 256 * 256;
    store_acc<4>(acc, H, HS, [&](int, int col, float v) { return v + __ldg(bv + 256 + col); });
    bv += 512;
    __syncthreads();
  }
  // ---- R4: identity shortcut
  zero_acc<4>(acc);
  gemm_seg<4>(acc, H, HS, 256, true, wp, wbuf); wp += 256 * 256;
  store_acc<4>(acc, Hn, HS, [&](int, int col, float v) { return fmaxf(v + __ldg(bv + col), 0.f); });
  __syncthreads();
  zero_acc<4>(acc);
  gemm_seg<4>(acc, Hn, HS, 256, false, wp, wbuf); wp += 256 * 256;
  store_acc<4>(acc, H, HS, [&](int row, int col, float v) { return H[row * HS + col] + v + __ldg(bv + 256 + col); });
  bv += 512;
  __syncthreads();
  return bv;  // start of the decoder-specific tail of vec
}

// ---- occupancy ------------------------------------------------------------
constexpr int OCC_KX = 64;
constexpr size_t OCC_SMEM = (size_t)(2 * TM * HS + TM * (OCC_KX + 4) + 2 * WBUF + TM * 4) * sizeof(float);

__global__ void __launch_bounds__(NT, 1)
occupancy_kernel(PlaneSet ps, int C, const float* __restrict__ pts, long long n,
                 long long batch_stride, int tiles_per_item, float divisor, float upper,
                 const float* __restrict__ Wg, const float* __restrict__ vec,
                 float* __restrict__ logits) {
  extern __shared__ float4 smem4[];
  float* H = reinterpret_cast<float*>(smem4);
  float* Hn = H + TM * HS;
  float* X = Hn + TM * HS;
  float* wbuf = X + TM * (OCC_KX + 4);
  float* P = wbuf + 2 * WBUF;  // [64][4] raw points
  const int tid = threadIdx.x;
  const int b = blockIdx.x / tiles_per_item;
  const long long n0 = (long long)(blockIdx.x % tiles_per_item) * TM;
  const int gr = tid & 63, gg = tid >> 6;
  long long gi = n0 + gr;
  if (gi > n - 1) gi = n - 1;
  const float* pp = pts + (size_t)b * batch_stride + gi * 3;
  const float p0 = __ldg(pp), p1 = __ldg(pp + 1), p2 = __ldg(pp + 2);
  if (gg == 0) { P[gr * 4 + 0] = p0; P[gr * 4 + 1] = p1; P[gr * 4 + 2] = p2; P[gr * 4 + 3] = 0.f; }
  const float g0 = occ_normalize(p0, divisor, upper);
  const float g1 = occ_normalize(p1, divisor, upper);
  const float g2 = occ_normalize(p2, divisor, upper);
  const int cpg = C / 4;
  auto gather = [&](int s) {
    // axis 0 'xy' -> (p0,p1), 1 'yz' -> (p1,p2), 2 'xz' -> (p0,p2); first -> column index
    Tap txy = make_tap<true>(g0, g1, ps.h[0 * 3 + s], ps.w[0 * 3 + s]);
    Tap tyz = make_tap<true>(g1, g2, ps.h[1 * 3 + s], ps.w[1 * 3 + s]);
    Tap txz = make_tap<true>(g0, g2, ps.h[2 * 3 + s], ps.w[2 * 3 + s]);
    size_t hw0 = (size_t)ps.h[s] * ps.w[s], hw1 = (size_t)ps.h[3 + s] * ps.w[3 + s],
           hw2 = (size_t)ps.h[6 + s] * ps.w[6 + s];
    const float* b0 = ps.data[s] + ((size_t)b * C + gg * cpg) * hw0;
    const float* b1 = ps.data[3 + s] + ((size_t)b * C + gg * cpg) * hw1;
    const float* b2 = ps.data[6 + s] + ((size_t)b * C + gg * cpg) * hw2;
    for (int c = 0; c < cpg; ++c) {
      float v = tap_sample(b0 + c * hw0, txy);
      v = __fadd_rn(v, tap_sample(b1 + c * hw1, tyz));
      v = __fadd_rn(v, tap_sample(b2 + c * hw2, txz));
      X[gr * (OCC_KX + 4) + gg * cpg + c] = v;
    }
  };
  // vec tail: netp_w[3][256], out_w[256], out_b[1]   (netp bias is folded into R1.b1)
  const float* tail = vec + (64 + 256) + 3 * 512;
  auto r1_extra = [&](int row, int col) {
    return fmaf(P[row * 4 + 2], __ldg(tail + 512 + col),
                fmaf(P[row * 4 + 1], __ldg(tail + 256 + col), P[row * 4 + 0] * __ldg(tail + col)));
  };
  resnet_chain<OCC_KX, 1>(H, Hn, X, wbuf, Wg, vec, gather, r1_extra);
  // logits = out_w . H + out_b   (4 threads per row)
  const float* ow = tail + 768;
  const int r = tid >> 2, q = tid & 3;
  float s = 0.f;
  for (int k = q * 64; k < q * 64 + 64; ++k) s = fmaf(H[r * HS + k], __ldg(ow + k), s);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (q == 0 && n0 + r < n) logits[(size_t)b * n + n0 + r] = s + __ldg(ow + 256);
}

// ---- video ------------------------------------------------------------------
constexpr int VID_KX = 192;
constexpr size_t VID_SMEM = (size_t)(2 * TM * HS + TM * (VID_KX + 4) + 2 * WBUF) * sizeof(float);

__global__ void __launch_bounds__(NT, 1)
video_kernel(PlaneSet ps, int C, const float* __restrict__ cxy, const float* __restrict__ cyt,
             const float* __restrict__ cxt, int T, int Hh, int Ww, int tiles_per_item,
             const float* __restrict__ Wg, const float* __restrict__ vec, void* __restrict__ out, int store) {
  extern __shared__ float4 smem4[];
  float* H = reinterpret_cast<float*>(smem4);
  float* Hn = H + TM * HS;
  float* X = Hn + TM * HS;
  float* wbuf = X + TM * (VID_KX + 4);
  const int tid = threadIdx.x;
  const int b = blockIdx.x / tiles_per_item;
  const long long n = (long long)T * Hh * Ww;
  const long long n0 = (long long)(blockIdx.x % tiles_per_item) * TM;
  const int gr = tid & 63, gg = tid >> 6;
  long long gi = n0 + gr;
  if (gi > n - 1) gi = n - 1;
  const int w = (int)(gi % Ww), h = (int)((gi / Ww) % Hh), t = (int)(gi / ((long long)Ww * Hh));
  // grids taken literally: channel 0 -> last plane axis, channel 1 -> second-to-last
  const float xy0 = __ldg(cxy + (size_t)h * Ww + w), xy1 = __ldg(cxy + (size_t)Hh * Ww + (size_t)h * Ww + w);
  const float yt0 = __ldg(cyt + (size_t)t * Hh + h), yt1 = __ldg(cyt + (size_t)T * Hh + (size_t)t * Hh + h);
  const float xt0 = __ldg(cxt + (size_t)t * Ww + w), xt1 = __ldg(cxt + (size_t)T * Ww + (size_t)t * Ww + w);
  const int cpg = C / 4;
  auto gather = [&](int s) {
    const float g0[3] = {xy0, yt0, xt0}, g1[3] = {xy1, yt1, xt1};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int ph = ps.h[a * 3 + s], pw = ps.w[a * 3 + s];
      Tap tp = make_tap<true>(g0[a], g1[a], ph, pw);
      size_t hw = (size_t)ph * pw;
      const float* base = ps.data[a * 3 + s] + ((size_t)b * C + gg * cpg) * hw;
      for (int c = 0; c < cpg; ++c)
        X[gr * (VID_KX + 4) + a * C + gg * cpg + c] = tap_sample(base + c * hw, tp);
    }
  };
  const float* tail = resnet_chain<VID_KX, 3>(H, Hn, X, wbuf, Wg, vec, gather,
                                              [](int, int) { return 0.f; });
  // out[b, c, voxel] = out_w[c] . lrelu(H, 0.2) + out_b[c]
  const int r = tid >> 2, c = tid & 3;
  if (c < 3 && n0 + r < n) {
    float s = 0.f;
    for (int k = 0; k < 256; ++k) s = fmaf(lrelu(H[r * HS + k], 0.2f), __ldg(tail + c * 256 + k), s);
    store_rgb(out, store, b, n, n0 + r, c, s + __ldg(tail + 768 + c));
  }
}

// ===========================================================================
// NeRF: MLPNeRF.forward (+ fused sample generation / gather / embedding)
// ===========================================================================
// gemm order: L1[160x256] L2[256x256] L3[160x256][256x256] L4 L5[160..][256..] L6
//             Lfinal[256x256] Ldir[256x128][32x128]
// vec: b1..b6 (256 each), bfinal(256), bdir(128), sigma_w(256), sigma_b(1),
//      rgb_w[3][128], rgb_b[3]
constexpr int NRF_XS = 164, NRF_DS = 36;
constexpr size_t NRF_SMEM = (size_t)(2 * TM * HS + TM * NRF_XS + TM * NRF_DS + 2 * WBUF + TM) * sizeof(float);

template <bool kFused>
__global__ void __launch_bounds__(NT, 1)
nerf_kernel(PlaneSet ps, int C, const float* __restrict__ xin, int x_stride, int sigma_only,
            const float* __restrict__ rays, int ray_stride, const float* __restrict__ t_vals, int z_stride,
            int n_samples, float plane_extent, long long n /* rows per object */,
            int tiles_per_item, float slope, const float* __restrict__ Wg,
            const float* __restrict__ vec, float* __restrict__ out) {
  extern __shared__ float4 smem4[];
  float* H = reinterpret_cast<float*>(smem4);
  float* Hn = H + TM * HS;
  float* X = Hn + TM * HS;
  float* D = X + TM * NRF_XS;
  float* wbuf = D + TM * NRF_DS;
  float* S = wbuf + 2 * WBUF;  // sigma per row
  const int tid = threadIdx.x;
  const int b = blockIdx.x / tiles_per_item;
  const long long n0 = (long long)(blockIdx.x % tiles_per_item) * TM;
  const int gr = tid & 63, gg = tid >> 6;
  long long gi = n0 + gr;
  if (gi > n - 1) gi = n - 1;

  if (kFused) {
    // utils/nerf_helpers.py:356-396 (perturb = 0, lindisp = False)
    const long long ray = gi / n_samples;
    const int smp = (int)(gi % n_samples);
    const float* rr = rays + (size_t)ray * ray_stride;
    const float z = nerf_z(t_vals, z_stride, ray, smp, __ldg(rr + 6), __ldg(rr + 7));
    float p[3], g[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      p[i] = __fadd_rn(__ldg(rr + i), __fmul_rn(__ldg(rr + 3 + i), z));
      g[i] = __fdiv_rn(p[i], plane_extent);
    }
    // triplane gather: xy=(x,y) yz=(y,z) xz=(x,z); concat [xy,yz,xz]
    const int cpg = C / 4;
    const float ga[3] = {g[0], g[1], g[0]}, gb[3] = {g[1], g[2], g[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      Tap tp = make_tap<true>(ga[a], gb[a], ps.h[a], ps.w[a]);
      size_t hw = (size_t)ps.h[a] * ps.w[a];
      const float* base = ps.data[a] + ((size_t)b * C + gg * cpg) * hw;
      for (int c = 0; c < cpg; ++c) X[gr * NRF_XS + a * C + gg * cpg + c] = tap_sample(base + c * hw, tp);
    }
    // positional embeddings (Embedder.embed, nerf_helpers.py:82-112)
    if (gg == 0) {
      float* e = X + gr * NRF_XS + 3 * C;
      e[0] = p[0]; e[1] = p[1]; e[2] = p[2];
      float f = 1.f;
      for (int l = 0; l < 10; ++l) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float a = __fmul_rn(p[i], f);
          e[3 + l * 6 + i] = sinf(a);
          e[3 + l * 6 + 3 + i] = cosf(a);
        }
        f *= 2.f;
      }
      for (int k = 3 * C + 63; k < NRF_XS; ++k) X[gr * NRF_XS + k] = 0.f;
    } else if (gg == 1) {
      float* e = D + gr * NRF_DS;
      float v[3] = {__ldg(rr + 8), __ldg(rr + 9), __ldg(rr + 10)};
      e[0] = v[0]; e[1] = v[1]; e[2] = v[2];
      float f = 1.f;
      for (int l = 0; l < 4; ++l) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float a = __fmul_rn(v[i], f);
          e[3 + l * 6 + i] = sinf(a);
          e[3 + l * 6 + 3 + i] = cosf(a);
        }
        f *= 2.f;
      }
      for (int k = 27; k < NRF_DS; ++k) e[k] = 0.f;
    }
  } else {
    const float* xr = xin + (size_t)gi * x_stride;
    for (int k = gg; k < NRF_XS; k += 4) X[gr * NRF_XS + k] = k < 159 ? __ldg(xr + k) : 0.f;
    for (int k = gg; k < NRF_DS; k += 4) D[gr * NRF_DS + k] = (k < 27 && !sigma_only) ? __ldg(xr + 159 + k) : 0.f;
  }
  __syncthreads();

  const float* wp = Wg;
  float acc[4][16];
  for (int l = 0; l < 6; ++l) {
    zero_acc<4>(acc);
    if (l == 0 || l == 2 || l == 4) { gemm_seg<4>(acc, X, NRF_XS, 160, false, wp, wbuf); wp += 160 * 256; }
    if (l > 0) { gemm_seg<4>(acc, H, HS, 256, false, wp, wbuf); wp += 256 * 256; }
    const float* bl = vec + l * 256;
    store_acc<4>(acc, H, HS, [&](int, int col, float v) { return lrelu(v + __ldg(bl + col), slope); });
    __syncthreads();
  }
  const float* sw = vec + 7 * 256 + 128;
  {
    const int r = tid >> 2, q = tid & 3;
    float s = 0.f;
    for (int k = q * 64; k < q * 64 + 64; ++k) s = fmaf(H[r * HS + k], __ldg(sw + k), s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __ldg(sw + 256);
    if (sigma_only) {
      if (q == 0 && n0 + r < n) out[(size_t)b * n + n0 + r] = s;
    } else if (q == 0) {
      S[r] = s;
    }
  }
  if (sigma_only) return;
  // xyz_encoding_final (no activation), then dir_encoding on [final, dir]
  zero_acc<4>(acc);
  gemm_seg<4>(acc, H, HS, 256, false, wp, wbuf); wp += 256 * 256;
  {
    const float* bf = vec + 6 * 256;
    store_acc<4>(acc, Hn, HS, [&](int, int col, float v) { return v + __ldg(bf + col); });
  }
  __syncthreads();
  {
    float a2[4][8];
    zero_acc<2>(a2);
    gemm_seg<2>(a2, Hn, HS, 256, false, wp, wbuf); wp += 256 * 128;
    gemm_seg<2>(a2, D, NRF_DS, 32, false, wp, wbuf); wp += 32 * 128;
    const float* bd = vec + 7 * 256;
    store_acc<2>(a2, H, HS, [&](int, int col, float v) { return lrelu(v + __ldg(bd + col), slope); });
  }
  __syncthreads();
  const float* rw = sw + 257;
  const int r = tid >> 2, c = tid & 3;
  if (n0 + r < n) {
    float v;
    if (c < 3) {
      float s = 0.f;
      for (int k = 0; k < 128; ++k) s = fmaf(H[r * HS + k], __ldg(rw + c * 128 + k), s);
      s += __ldg(rw + 384 + c);
      v = 1.f / (1.f + expf(-s));
    } else {
      v = S[r];
    }
    out[((size_t)b * n + n0 + r) * 4 + c] = v;
  }
}

// raw2outputs (utils/nerf_helpers.py:487-530), raw_noise_std = 0.  One thread per ray.
__global__ void nerf_composite_kernel(const float* __restrict__ raw, const float* __restrict__ rays,
                                      int ray_stride, const float* __restrict__ t_vals, int z_stride, int n_samples,
                                      long long n_rays, int batch, int white_bkgd,
                                      float* __restrict__ rgb_map) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * batch) return;
  const long long ray = i % n_rays;
  const float* rr = rays + (size_t)ray * ray_stride;
  const float near = __ldg(rr + 6), far = __ldg(rr + 7);
  const float dx = __ldg(rr + 3), dy = __ldg(rr + 4), dz = __ldg(rr + 5);
  const float dn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  const float4* rw = reinterpret_cast<const float4*>(raw) + (size_t)i * n_samples;
  float T = 1.f, acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
  float z = nerf_z(t_vals, z_stride, ray, 0, near, far);
  for (int s = 0; s < n_samples; ++s) {
    float dist;
    if (s + 1 < n_samples) {
      float zn = nerf_z(t_vals, z_stride, ray, s + 1, near, far);
      dist = __fsub_rn(zn, z);
      z = zn;
    } else {
      dist = 1e10f;
    }
    dist = __fmul_rn(dist, dn);
    float4 v = __ldg(rw + s);
    float alpha = __fsub_rn(1.f, expf(-__fmul_rn(softplus20(v.w), dist)));
    float wgt = __fmul_rn(alpha, T);
    cr = fmaf(wgt, v.x, cr); cg = fmaf(wgt, v.y, cg); cb = fmaf(wgt, v.z, cb);
    acc = __fadd_rn(acc, wgt);
    T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
  }
  if (white_bkgd) {
    float bg = __fsub_rn(1.f, acc);
    cr += bg; cg += bg; cb += bg;
  }
  rgb_map[i * 3 + 0] = cr; rgb_map[i * 3 + 1] = cg; rgb_map[i * 3 + 2] = cb;
}

// (batch, C, HW) -> (batch, HW, C): channels-last planes for vectorised scattered gathers.
// One CTA = 32 pixels x all channels of one item, through a padded shared-memory tile.
__global__ void planes_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][65];
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads: 32 x 8
  for (int c0 = 0; c0 < C; c0 += 64) {
    for (int c = ty; c < 64 && c0 + c < C; c += 8)
      tile[tx][c] = (p0 + tx < HW) ? src[((size_t)b * C + c0 + c) * HW + p0 + tx] : 0.f;
    __syncthreads();
    for (int p = ty; p < 32; p += 8)
      for (int c = tx; c < 64 && c0 + c < C; c += 32)
        if (p0 + p < HW) dst[((size_t)b * HW + p0 + p) * C + c0 + c] = tile[p][c];
    __syncthreads();
  }
}

}  // namespace fp32

int launch_planes_to_nhwc(const float* src, float* dst, int batch, int C, int HW, cudaStream_t st) {
  dim3 grid((HW + 31) / 32, batch);
  fp32::planes_to_nhwc_kernel<<<grid, 256, 0, st>>>(src, dst, C, HW);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
template <class K>
static int set_smem(K kernel, size_t bytes) {
  DDMI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return DDMI_OK;
}

static int check_grid(long long tiles) {
  if (tiles <= 0 || tiles > 2147483647LL) {
    set_error("tile count %lld out of range for one launch", tiles);
    return DDMI_ERR_UNSUPPORTED;
  }
  return DDMI_OK;
}

// Hierarchical sampling, utils/nerf_helpers.py:166-209 (sample_pdf): per ray, pdf = (w + 1e-5) / sum, cdf = [0, cumsum(pdf)],
// then for every u: inds = searchsorted(cdf, u, right=True), below = max(inds - 1, 0), above = min(inds, n_bins - 1),
// t = (u - cdf[below]) / (denom < 1e-5 ? 1 : denom), sample = bins[below] + t * (bins[above] - bins[below]).
// One warp per ray; the cdf is accumulated sequentially (the order of torch.cumsum) in shared memory.
namespace fp32 {
constexpr int PDF_MAX_BINS = 1024;
__global__ void __launch_bounds__(128)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights, const float* __restrict__ u,
                  long long n_rays, int n_bins, int n_samples, float* __restrict__ out) {
  __shared__ float cdf_s[4][PDF_MAX_BINS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * 4 + warp;
  if (ray >= n_rays) return;
  float* cdf = cdf_s[warp];
  const float* w = weights + ray * (n_bins - 1);
  const float* b = bins + ray * n_bins;
  float part = 0.f;
  for (int i = lane; i < n_bins - 1; i += 32) part += __fadd_rn(__ldg(w + i), 1e-5f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) {
    float c = 0.f;
    cdf[0] = 0.f;
    for (int i = 0; i < n_bins - 1; ++i) {
      c = __fadd_rn(c, __fdiv_rn(__fadd_rn(__ldg(w + i), 1e-5f), part));
      cdf[i + 1] = c;
    }
  }
  __syncwarp();
  for (int j = lane; j < n_samples; j += 32) {
    const float uj = __ldg(u + ray * n_samples + j);
    int lo = 0, hi = n_bins;                       // first index with cdf[idx] > uj  (searchsorted right=True)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= uj) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, n_bins - 1);
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fsub_rn(uj, cdf[below]), denom);
    const float b0 = __ldg(b + below), b1 = __ldg(b + above);
    out[ray * n_samples + j] = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
  }
}
}  // namespace fp32

int launch_sample_pdf(const float* bins, const float* weights, const float* u, long long n_rays, int n_bins, int n_samples,
                      float* out, cudaStream_t st) {
  fp32::sample_pdf_kernel<<<(unsigned)((n_rays + 3) / 4), 128, 0, st>>>(bins, weights, u, n_rays, n_bins, n_samples, out);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_image_fp32(const PlaneSet& ps, int batch, int C, const float* cx, const float* cy,
                      long long n, const float* Wg, const float* vec, void* out, int store, const NoiseArgs& na,
                      cudaStream_t st) {
  long long tpi = (n + fp32::TM - 1) / fp32::TM;
  int rc = check_grid(tpi * batch);
  if (rc) return rc;
  rc = set_smem(fp32::image_kernel, fp32::IMG_SMEM);
  if (rc) return rc;
  fp32::image_kernel<<<(unsigned)(tpi * batch), fp32::NT, fp32::IMG_SMEM, st>>>(ps, C, cx, cy, n, (int)tpi, Wg, vec, out, store, na);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_occupancy_fp32(const PlaneSet& ps, int batch, int C, const float* pts, long long n,
                          long long batch_stride, float divisor, float upper, const float* Wg,
                          const float* vec, float* logits, cudaStream_t st) {
  long long tpi = (n + fp32::TM - 1) / fp32::TM;
  int rc = check_grid(tpi * batch);
  if (rc) return rc;
  rc = set_smem(fp32::occupancy_kernel, fp32::OCC_SMEM);
  if (rc) return rc;
  fp32::occupancy_kernel<<<(unsigned)(tpi * batch), fp32::NT, fp32::OCC_SMEM, st>>>(
      ps, C, pts, n, batch_stride, (int)tpi, divisor, upper, Wg, vec, logits);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_video_fp32(const PlaneSet& ps, int batch, int C, const float* cxy, const float* cyt,
                      const float* cxt, int T, int H, int W, const float* Wg, const float* vec,
                      void* out, int store, cudaStream_t st) {
  long long n = (long long)T * H * W;
  long long tpi = (n + fp32::TM - 1) / fp32::TM;
  int rc = check_grid(tpi * batch);
  if (rc) return rc;
  rc = set_smem(fp32::video_kernel, fp32::VID_SMEM);
  if (rc) return rc;
  fp32::video_kernel<<<(unsigned)(tpi * batch), fp32::NT, fp32::VID_SMEM, st>>>(
      ps, C, cxy, cyt, cxt, T, H, W, (int)tpi, Wg, vec, out, store);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_nerf_mlp_fp32(const float* x, long long n, int x_stride, int sigma_only, float slope,
                         const float* Wg, const float* vec, float* out, cudaStream_t st) {
  long long tpi = (n + fp32::TM - 1) / fp32::TM;
  int rc = check_grid(tpi);
  if (rc) return rc;
  rc = set_smem(fp32::nerf_kernel<false>, fp32::NRF_SMEM);
  if (rc) return rc;
  PlaneSet ps = {};
  fp32::nerf_kernel<false><<<(unsigned)tpi, fp32::NT, fp32::NRF_SMEM, st>>>(
      ps, 32, x, x_stride, sigma_only, nullptr, 0, nullptr, 0, 1, 1.f, n, (int)tpi, slope, Wg, vec, out);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_nerf_composite(const float* raw, const float* rays, int ray_stride, const float* t_vals, int z_stride, int n_samples,
                          long long n_rays, int batch, int white_bkgd, float* rgb_map, cudaStream_t st) {
  long long tot = n_rays * batch;
  fp32::nerf_composite_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(raw, rays, ray_stride, t_vals, z_stride, n_samples, n_rays,
                                                                             batch, white_bkgd, rgb_map);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_nerf_render_fp32(const PlaneSet& ps, int batch, int C, const float* rays, long long n_rays,
                            int ray_stride, const float* t_vals, int z_stride, int n_samples, float plane_extent,
                            float slope, int white_bkgd, const float* Wg, const float* vec,
                            float* rgb_map, float* raw, cudaStream_t st) {
  long long n = n_rays * n_samples;
  long long tpi = (n + fp32::TM - 1) / fp32::TM;
  int rc = check_grid(tpi * batch);
  if (rc) return rc;
  rc = set_smem(fp32::nerf_kernel<true>, fp32::NRF_SMEM);
  if (rc) return rc;
  fp32::nerf_kernel<true><<<(unsigned)(tpi * batch), fp32::NT, fp32::NRF_SMEM, st>>>(
      ps, C, nullptr, 0, 0, rays, ray_stride, t_vals, z_stride, n_samples, plane_extent, n, (int)tpi, slope,
      Wg, vec, raw);
  DDMI_CUDA(cudaGetLastError());
  long long tot = n_rays * batch;
  fp32::nerf_composite_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(
      raw, rays, ray_stride, t_vals, z_stride, n_samples, n_rays, batch, white_bkgd, rgb_map);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
