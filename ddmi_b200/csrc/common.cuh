// Shared device/host helpers for the ddmi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ddmi_b200.h"

namespace ddmi {

// ---------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DDMI_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      ::ddmi::set_error(__VA_ARGS__);                \
      return DDMI_ERR_BAD_ARG;                       \
    }                                                \
  } while (0)

#define DDMI_CUDA(call)                                              \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return ::ddmi::cuda_fail(e__, #call);    \
  } while (0)

// Nine planes at most (3 axes x 3 scales); passed to kernels by value.
struct PlaneSet {
  const float* data[9];
  int h[9];
  int w[9];
};

// ---------------------------------------------------------------------------
// Bilinear tap of F.grid_sample(mode='bilinear', padding_mode='border').
// Index maps as in ATen's grid_sampler (the arithmetic behind
// utils/general_utils.py:122-137 and utils/nerf_helpers.py:391-393):
//   align_corners = true : ix = ((g + 1) / 2) * (W - 1)
//   align_corners = false: ix = ((g + 1) * W - 1) / 2
//   border: ix = min(W - 1, max(ix, 0)); 4 taps around floor(ix).
// The +1 neighbour is clamped into range; its weight is exactly 0 whenever the
// clamp is active, so the result equals ATen's "skip out-of-bounds tap".
// ---------------------------------------------------------------------------
struct Tap {
  int o00, o01, o10, o11;      // element offsets inside one channel image
  float w00, w01, w10, w11;    // nw, ne, sw, se
};

template <bool kAlignCorners>
__device__ __forceinline__ float unnormalize(float g, int size) {
  if (kAlignCorners) {
    return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1));
  } else {
    return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 2.f);
  }
}

template <bool kAlignCorners>
__device__ __forceinline__ Tap make_tap(float gx, float gy, int H, int W) {
  float ix = unnormalize<kAlignCorners>(gx, W);
  float iy = unnormalize<kAlignCorners>(gy, H);
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  float fx = floorf(ix), fy = floorf(iy);
  int x0 = (int)fx, y0 = (int)fy;
  float tx1 = __fsub_rn(__fadd_rn(fx, 1.f), ix);  // ix_se - ix
  float tx0 = __fsub_rn(ix, fx);                  // ix - ix_nw
  float ty1 = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  float ty0 = __fsub_rn(iy, fy);
  int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  Tap t;
  t.o00 = y0 * W + x0; t.o01 = y0 * W + x1;
  t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
  t.w00 = __fmul_rn(tx1, ty1); t.w01 = __fmul_rn(tx0, ty1);
  t.w10 = __fmul_rn(tx1, ty0); t.w11 = __fmul_rn(tx0, ty0);
  return t;
}

__device__ __forceinline__ float tap_sample(const float* __restrict__ ch, const Tap& t) {
  float v00 = __ldg(ch + t.o00), v01 = __ldg(ch + t.o01);
  float v10 = __ldg(ch + t.o10), v11 = __ldg(ch + t.o11);
  float acc = v00 * t.w00;
  acc = fmaf(v01, t.w01, acc);
  acc = fmaf(v10, t.w10, acc);
  acc = fmaf(v11, t.w11, acc);
  return acc;
}

// NCH consecutive channels of an NCHW plane at one tap: all 4 * NCH loads are issued before any is consumed, so a thread
// pays one L2 round trip per call instead of one per few channels (the kernels keep no L1 beside their shared memory).
template <int NCH>
__device__ __forceinline__ void tap_sample_n(const float* __restrict__ ch0, size_t hw, const Tap& t, float (&out)[NCH]) {
  float v[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const float* ch = ch0 + (size_t)c * hw;
    v[c][0] = __ldg(ch + t.o00); v[c][1] = __ldg(ch + t.o01);
    v[c][2] = __ldg(ch + t.o10); v[c][3] = __ldg(ch + t.o11);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c)   // same order as tap_sample
    out[c] = fmaf(v[c][3], t.w11, fmaf(v[c][2], t.w10, fmaf(v[c][1], t.w01, v[c][0] * t.w00)));
}

// Channels-last variant: 8 consecutive channels of one pixel are two float4 loads.
// img = base of one (H, W, C) item; c = first channel (multiple of 4).
__device__ __forceinline__ void tap_sample8_nhwc(const float* __restrict__ img, const Tap& t, int C, int c, float (&out)[8]) {
  const float4* p00 = reinterpret_cast<const float4*>(img + (size_t)t.o00 * C + c);
  const float4* p01 = reinterpret_cast<const float4*>(img + (size_t)t.o01 * C + c);
  const float4* p10 = reinterpret_cast<const float4*>(img + (size_t)t.o10 * C + c);
  const float4* p11 = reinterpret_cast<const float4*>(img + (size_t)t.o11 * C + c);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = __ldg(p00 + h), b = __ldg(p01 + h), cc = __ldg(p10 + h), d = __ldg(p11 + h);
    out[4 * h + 0] = fmaf(d.x, t.w11, fmaf(cc.x, t.w10, fmaf(b.x, t.w01, a.x * t.w00)));
    out[4 * h + 1] = fmaf(d.y, t.w11, fmaf(cc.y, t.w10, fmaf(b.y, t.w01, a.y * t.w00)));
    out[4 * h + 2] = fmaf(d.z, t.w11, fmaf(cc.z, t.w10, fmaf(b.z, t.w01, a.z * t.w00)));
    out[4 * h + 3] = fmaf(d.w, t.w11, fmaf(cc.w, t.w10, fmaf(b.w, t.w01, a.w * t.w00)));
  }
}

// normalize_coordinate + sample_plane_feature (utils/general_utils.py:71-94,
// 115-119): u = p / (1 + padding + 10e-6) + 0.5, clamped to [0, 1 - 10e-6],
// g = 2u - 1.  `divisor` and `upper` are computed on the host exactly the way
// the Python scalars are (double arithmetic, then cast to fp32).
__device__ __forceinline__ float occ_normalize(float p, float divisor, float upper) {
  float u = __fadd_rn(__fdiv_rn(p, divisor), 0.5f);
  if (u >= 1.f) u = upper;
  if (u < 0.f) u = 0.f;
  return __fsub_rn(__fmul_rn(2.f, u), 1.f);
}

// Depth of sample `smp` on ray `ray`: z_stride == 0 -> near * (1 - t) + far * t from the shared t table (nerf_helpers.py:356-358,
// perturb = 0, lindisp = False); z_stride = n_samples -> the caller's per-ray table (stratified `perturb`, `lindisp`: :359-380).
__device__ __forceinline__ float nerf_z(const float* __restrict__ zt, int z_stride, long long ray, int smp, float near, float far) {
  if (z_stride) return __ldg(zt + (size_t)ray * z_stride + smp);
  const float tv = __ldg(zt + smp);
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, tv)), __fmul_rn(far, tv));
}

// Output store modes of the image / video decoders (DDMI_STORE_* in ddmi_b200.h): channel c (of 3) of coordinate gi of
// item b, n coordinates per item.  0: (b, 3, n) fp32 as the reference returns it; 1: the same, clamp(v, -1, 1)
// (evals/eval.py:162,226); 2: uint8((clamp(v,-1,1) + 1) * 127.5) channels-last (b, n, 3) -- the reference's
// `rearrange((fake.clamp(-1,1) + 1) * 127.5, 'b c t h w -> b t h w c').type(torch.uint8)` (evals/eval.py:289,336-337),
// fp32 ops in that order, truncating cast.
__device__ __forceinline__ void store_rgb(void* out, int store, size_t b, long long n, long long gi, int c, float v) {
  if (store != 0) v = fminf(fmaxf(v, -1.f), 1.f);
  if (store == 2)
    reinterpret_cast<unsigned char*>(out)[((size_t)b * n + gi) * 3 + c] =
        (unsigned char)__float2uint_rz(__fmul_rn(__fadd_rn(v, 1.f), 127.5f));
  else
    reinterpret_cast<float*>(out)[((size_t)b * 3 + c) * n + gi] = v;
}

// ---------------------------------------------------------------------------
// Noise injection of the image decoder's 12 StyledConv layers (models/d2c_vae/blocks.py:286-297, 349-356):
// out = conv(x) + noise.weight * noise[b, 0, y, x] before the bias / leaky ReLU.  The reference draws N(0,1) inside forward
// (not reproducible); here the noise is either 12 EXPLICIT tensors (mode 1: layer l reads p[l], (batch, n) fp32) or a
// DOCUMENTED counter-based stream (mode 2), identical for every kernel, precision and tiling:
//   x = Philox4x32-10(key = (seed lo, seed hi), counter = (g lo, g hi, b, blk)),  g = coordinate index, b = item, blk = l / 3
//   u_i = ((x_i >> 9) + 0.5) * 2^-23;  z0, z1 = sqrt(-2 ln u0) * (cos, sin)(2 pi u1);  z2 = sqrt(-2 ln u2) * cos(2 pi u3)
//   noise of conv1, conv2, conv3 of block blk = z0, z1, z2   (oracle/ddmi_oracle.py::philox_noise is the same in numpy)
// ---------------------------------------------------------------------------
struct NoiseArgs {
  const float* p[12];
  unsigned long long seed;
  int mode;                 // DDMI_NOISE_NONE / _TENSORS / _PHILOX
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    c[0] = h1 ^ c[1] ^ k0; c[1] = l1; c[2] = h0 ^ c[3] ^ k1; c[3] = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void philox_normal3(unsigned long long seed, unsigned long long g, uint32_t b, uint32_t blk, float (&z)[3]) {
  uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), b, blk};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  float u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = __fmul_rn(__fadd_rn((float)(c[i] >> 9), 0.5f), 1.1920928955078125e-7f);   // 2^-23
  const float r0 = sqrtf(__fmul_rn(-2.f, logf(u[0]))), r1 = sqrtf(__fmul_rn(-2.f, logf(u[2])));
  float s0, c0;
  sincosf(__fmul_rn(6.2831853071795864769f, u[1]), &s0, &c0);
  z[0] = __fmul_rn(r0, c0);
  z[1] = __fmul_rn(r0, s0);
  z[2] = __fmul_rn(r1, cosf(__fmul_rn(6.2831853071795864769f, u[3])));
}
// noise of the three StyledConv layers of block `blk` at (item b, coordinate gi)
__device__ __forceinline__ void noise_block3(const NoiseArgs& na, int blk, size_t b, long long n, long long gi, float (&z)[3]) {
  if (na.mode == DDMI_NOISE_TENSORS) {
#pragma unroll
    for (int j = 0; j < 3; ++j) z[j] = __ldg(na.p[3 * blk + j] + b * (size_t)n + gi);
  } else {
    philox_normal3(na.seed, (unsigned long long)gi, (uint32_t)b, (uint32_t)blk, z);
  }
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// torch.nn.functional.softplus (beta = 1, threshold = 20)
__device__ __forceinline__ float softplus20(float v) { return v > 20.f ? v : log1pf(expf(v)); }

}  // namespace ddmi
