// Plane-producer tail: the last layers of the D2C-VAE decoders, which EMIT the PE planes the decode
// kernels sample (models/d2c_vae/autoencoder_unet.py:770-771,812-814 -- the per-level `hdbf` 1x1 convolutions -- and :822-827 --
// norm_out (GroupNorm 32, eps 1e-6) -> swish -> conv_out (3x3, pad 1) [-> tanh]; same tail in the video / triplane decoders,
// :1111-1142, :1531-1562).  Fused here so that a plane leaves the producer once, in the layout its consumer gathers from:
// NCHW for the image / video kernels (TMA windows, scalar taps) or channels-last for the scattered-query kernels (occupancy,
// NeRF) -- no transposition pass, no normalised / activated copy of the feature map in HBM.
//
// fp32 CUDA-core kernels (the exact-arithmetic class of decode_fp32.cu): a 3x3 convolution of 128 -> 64 channels over B x 256^2
// pixels is 0.6 TFLOP per batch of 64 against 128 TFLOP of decode, so the tail is ~2 % of a generation step; a tcgen05
// implicit-GEMM version is the next step if it ever shows up in a profile.
//   gn_stats_kernel   : mean / rstd per (item, group), two passes over the group (biased variance, as torch.nn.GroupNorm)
//   plane_conv_kernel : one block = 256 output pixels (8 rows x 32 columns) x COUT channels, input channels streamed through
//                       shared memory in chunks of 8 (tile with halo, normalised + swish'd on the way in) next to the chunk's
//                       weights; one thread = 4 consecutive pixels x 16 output channels (inputs of a row are shared by the
//                       three x-taps: 18 shared-memory loads per 192 FMAs)
#include "common.cuh"

namespace ddmi {
namespace ptail {

__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, int C, int groups, long long hw, float eps, float* __restrict__ stats) {
  // block (b, g): channels [g * cpg, (g + 1) * cpg) are contiguous in NCHW
  const int cpg = C / groups;
  const long long n = (long long)cpg * hw;
  const float* p = x + ((long long)blockIdx.x * cpg) * hw;     // blockIdx.x = b * groups + g
  __shared__ float sh[8];
  __shared__ float s_mean;
  auto block_sum = [&](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w];
    __syncthreads();
    return t;
  };
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) s += __ldg(p + i);
  const float mean = block_sum(s) / (float)n;
  if (threadIdx.x == 0) s_mean = mean;
  float q = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float d = __ldg(p + i) - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum(q) / (float)n;
  if (threadIdx.x == 0) {
    stats[2 * blockIdx.x] = s_mean;
    stats[2 * blockIdx.x + 1] = rsqrtf(var + eps);
  }
}

constexpr int TY = 8, TX = 32;              // output tile
constexpr int PX = 4, CG = 16;              // register tile of one thread: 4 consecutive pixels x 16 output channels

template <int KS, int COUT, bool NORM>
__global__ void __launch_bounds__(TY * TX, 2)
plane_conv_kernel(const float* __restrict__ x, int C, int H, int W, const float* __restrict__ stats, int groups,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ wt,
                  const float* __restrict__ bias, int tanh_out, int nhwc, float* __restrict__ out) {
  // 256 threads = (COUT / 16) channel groups x pixel groups; with COUT = 32 a block covers two 8 x 32 tiles stacked in y
  constexpr int R = KS / 2, NCG = COUT / CG, ROWS = TY * (4 / NCG), IY = ROWS + 2 * R, IXP = TX + 4 /* halo, padded to 16 B */,
                TAPS = KS * KS, NIN = PX + 2 * R;
  constexpr int CK = KS == 1 ? (COUT == 64 ? 32 : 16) : 8;   // input channels per shared-memory chunk (<= 48 KB static in all)
  __shared__ __align__(16) float s_in[CK][IY][IXP];
  __shared__ __align__(16) float s_w[CK][TAPS][COUT];
  const int tiles_x = (W + TX - 1) / TX;
  const int tx0 = (blockIdx.x % tiles_x) * TX, ty0 = (blockIdx.x / tiles_x) * ROWS, b = blockIdx.y;
  const int cg = threadIdx.x / (TY * TX / NCG) , pg = threadIdx.x % (TY * TX / NCG);   // a warp shares one channel group
  const int ly = pg / (TX / PX), lx = (pg % (TX / PX)) * PX;
  const int ox = tx0 + lx, oy = ty0 + ly;
  const size_t hw = (size_t)H * W;
  const float* xb = x + (size_t)b * C * hw;
  const int cpg = NORM ? C / groups : 1;
  float acc[PX][CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    const float bj = __ldg(bias + cg * CG + j);
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[p][j] = bj;
  }
  for (int c0 = 0; c0 < C; c0 += CK) {
    // ---- stage the chunk: input tile with halo (zero padded AFTER the activation, like conv2d's padding of its input);
    // column q of the tile = image column tx0 + q - R
    // (unrolled by 4: the loads of four of a thread's ~11 elements are in flight together; two blocks per SM cover the rest)
    constexpr int NSTAGE = CK * IY * (TX + 2 * R);
#pragma unroll 4
    for (int k = 0; k < (NSTAGE + TY * TX - 1) / (TY * TX); ++k) {
      const int i = threadIdx.x + k * TY * TX;
      if (i < NSTAGE) {
        const int c = i / (IY * (TX + 2 * R)), r = (i / (TX + 2 * R)) % IY, q = i % (TX + 2 * R);
        const int gy = ty0 + r - R, gx = tx0 + q - R, ch = c0 + c;
        float v = 0.f;
        if (ch < C && gy >= 0 && gy < H && gx >= 0 && gx < W) {
          v = __ldg(xb + (size_t)ch * hw + (size_t)gy * W + gx);
          if (NORM) {
            const int g = ch / cpg;
            const float mean = __ldg(stats + 2 * (b * groups + g)), rstd = __ldg(stats + 2 * (b * groups + g) + 1);
            v = fmaf((v - mean) * rstd, __ldg(gamma + ch), __ldg(beta + ch));     // GroupNorm affine
            v = v / (1.f + __expf(-v));                                          // swish: x * sigmoid(x)
          }
        }
        s_in[c][r][q] = v;
      }
    }
    // weights of the chunk: wt arrives as (C, KS * KS, COUT) (the host transposes nn.Conv2d's (COUT, C, KS, KS) once), so a
    // chunk is one contiguous run of CK * TAPS * COUT floats
    constexpr int NW4 = CK * TAPS * COUT / 4;
#pragma unroll
    for (int k = 0; k < (NW4 + TY * TX - 1) / (TY * TX); ++k) {
      const int i = threadIdx.x + k * TY * TX;
      if (i < NW4) {
        const int ch = c0 + (4 * i) / (TAPS * COUT);
        reinterpret_cast<float4*>(&s_w[0][0][0])[i] =
            ch < C ? __ldg(reinterpret_cast<const float4*>(wt + (size_t)c0 * TAPS * COUT) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CK; ++c) {
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        float in[NIN + 2];                                        // the PX + 2R inputs of this row, two aligned vector loads
        const float* row = &s_in[c][ly + dy][lx];
        *reinterpret_cast<float4*>(&in[0]) = *reinterpret_cast<const float4*>(row);
        if (KS > 1) *reinterpret_cast<float2*>(&in[4]) = *reinterpret_cast<const float2*>(row + 4);
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float4* w4 = reinterpret_cast<const float4*>(&s_w[c][dy * KS + dx][cg * CG]);
#pragma unroll
          for (int j = 0; j < CG / 4; ++j) {
            const float4 w = w4[j];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              const float v = in[p + dx];
              acc[p][4 * j] = fmaf(v, w.x, acc[p][4 * j]);
              acc[p][4 * j + 1] = fmaf(v, w.y, acc[p][4 * j + 1]);
              acc[p][4 * j + 2] = fmaf(v, w.z, acc[p][4 * j + 2]);
              acc[p][4 * j + 3] = fmaf(v, w.w, acc[p][4 * j + 3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (oy < H) {
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (ox + p >= W) continue;
      if (tanh_out) {
#pragma unroll
        for (int j = 0; j < CG; ++j) acc[p][j] = tanhf(acc[p][j]);
      }
      if (nhwc) {
        float4* o = reinterpret_cast<float4*>(out + (((size_t)b * H + oy) * W + ox + p) * COUT + cg * CG);
#pragma unroll
        for (int j = 0; j < CG / 4; ++j) o[j] = make_float4(acc[p][4 * j], acc[p][4 * j + 1], acc[p][4 * j + 2], acc[p][4 * j + 3]);
      } else {
        float* o = out + ((size_t)b * COUT + cg * CG) * hw + (size_t)oy * W + ox + p;
#pragma unroll
        for (int j = 0; j < CG; ++j) o[(size_t)j * hw] = acc[p][j];
      }
    }
  }
}

}  // namespace ptail

// KS = 1: hdbf head (no norm); KS = 3: norm_out -> swish -> conv_out.  stats: B * groups * 2 floats of scratch (KS = 3).
int launch_plane_conv(const float* x, int B, int C, int H, int W, int ks, const float* gamma, const float* beta, int groups,
                      float eps, const float* wt, const float* bias, int cout, int tanh_out, int nhwc, float* stats, float* out,
                      cudaStream_t st) {
  using namespace ptail;
  const int rows = cout == 64 ? TY : 2 * TY;    // output rows per block (see plane_conv_kernel)
  const dim3 grid((unsigned)(((W + TX - 1) / TX) * ((H + rows - 1) / rows)), (unsigned)B);
  if (ks == 3) {
    gn_stats_kernel<<<(unsigned)(B * groups), 256, 0, st>>>(x, C, groups, (long long)H * W, eps, stats);
    if (cout == 64) plane_conv_kernel<3, 64, true><<<grid, TY * TX, 0, st>>>(x, C, H, W, stats, groups, gamma, beta, wt, bias, tanh_out, nhwc, out);
    else plane_conv_kernel<3, 32, true><<<grid, TY * TX, 0, st>>>(x, C, H, W, stats, groups, gamma, beta, wt, bias, tanh_out, nhwc, out);
  } else {
    if (cout == 64) plane_conv_kernel<1, 64, false><<<grid, TY * TX, 0, st>>>(x, C, H, W, nullptr, 1, nullptr, nullptr, wt, bias, 0, nhwc, out);
    else plane_conv_kernel<1, 32, false><<<grid, TY * TX, 0, st>>>(x, C, H, W, nullptr, 1, nullptr, nullptr, wt, bias, 0, nhwc, out);
  }
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
