/*
 * ddmi_b200.h -- C ABI of the B200-native D2C-VAE continuous-decoding path.
 *
 * One entry point per decoder family of the reference (mlvlab/DDMI).  Only PODs
 * cross this boundary: device pointers, sizes, scalars and a CUDA stream handle.
 * No torch / C++ types, no exceptions: every function returns an int status
 * (DDMI_OK == 0) and ddmi_last_error() gives the message for the calling thread.
 *
 * Ownership: the caller owns every buffer (planes, coordinates, packed weights,
 * outputs); the library keeps no state between calls besides per-device kernel
 * attributes.  All work is enqueued on `stream` of the CURRENT device and is
 * stream-ordered; calls are re-entrant.  There is no CPU path: a machine without
 * an sm_100 device gets DDMI_ERR_CUDA / DDMI_ERR_UNSUPPORTED, never a fallback.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   ddmi_decode_image      models/d2c_vae/mlp.py:34-66   MLP.forward
 *                          (+ utils/general_utils.py:122-123, blocks.py:187-356,
 *                             604-638, op/fused_bias_act_kernel.cu:18-49)
 *   ddmi_decode_occupancy  models/d2c_vae/mlp.py:82-111  MLP3D.forward
 *                          (+ general_utils.py:71-94,115-119,126-131, blocks.py:673-716)
 *   ddmi_decode_video      models/d2c_vae/mlp.py:128-157 MLPVideo.forward
 *                          (+ general_utils.py:134-145)
 *   ddmi_nerf_mlp          models/d2c_vae/mlp.py:241-281 MLPNeRF.forward
 *   ddmi_nerf_render       utils/nerf_helpers.py:296-452 render_rays
 *                          (+ :455-475 run_network, :82-112 Embedder.embed,
 *                             :487-530 raw2outputs)
 *   ddmi_sample_pdf        utils/nerf_helpers.py:166-209 sample_pdf
 *   ddmi_plane_head / _tail models/d2c_vae/autoencoder_unet.py:770-771,812-814 (hdbf 1x1 convs) / :822-827 (norm_out -> swish ->
 *                          conv_out [-> tanh]); same tail at :1111-1142, :1531-1562
 *   ddmi_mcubes_*          convocc/src/utils/libmcubes/marchingcubes.h:23-193 mc::marching_cubes (libmcubes.marching_cubes,
 *                          pywrapper.cpp:90-107) + the vertex post-processing of convocc/src/conv_onet/generation.py:152-186
 * The reference binds its native ops with pybind11 inside a JIT torch extension
 * (models/d2c_vae/op/fused_bias_act.cpp:18-20); INTEGRATION.md shows the ctypes
 * stub a maintainer adds instead.
 *
 * Packed weights: produced by ddmi_b200/packing.py (layouts in DESIGN.md §4);
 * `precision` selects both the layout and the kernel family:
 *   DDMI_PREC_FP32    fp32 CUDA-core kernels (exact-arithmetic path)
 *   DDMI_PREC_BF16X3  tcgen05 tensor-core kernels, bf16 hi/lo split operands,
 *                     3 MMAs per product, fp32 accumulation in TMEM
 *   DDMI_PREC_F16F8   tcgen05 kernels, fp16 main term + two FP8 correction terms (e5m2 activation side x
 *                     e4m3 weight side) at twice the MMA rate: 2/3 of the tensor-pipe time of BF16X3 at
 *                     ~2e-4 max-abs (6e-5 of |out|max) error (all four decoders; CTA-pair kernels only,
 *                     weights must be pair-packed).  OPERAND RANGE: |weight| < 16 (checked by the packer);
 *                     the relative accuracy holds for |plane value|, |hidden activation| <= 2.8e4 (the e5m2
 *                     residual operand 4096 (a - fp16(a)) <= 2 |a| saturates at 57344); beyond that the
 *                     correction terms clamp (accuracy degrades towards single-pass fp16, ~1e-3 relative)
 *                     and at 65504 the fp16 main term clamps.  Nothing is detected at run time: callers with
 *                     such signals use DDMI_PREC_BF16X3 (bf16 range).
 */
#ifndef DDMI_B200_H
#define DDMI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DDMI_API __attribute__((visibility("default")))
#else
#define DDMI_API
#endif

#define DDMI_ABI_VERSION 10

enum {
  DDMI_OK = 0,
  DDMI_ERR_BAD_ARG = 1,      /* null pointer, non-positive size, misaligned buffer */
  DDMI_ERR_UNSUPPORTED = 2,  /* shape / width / precision this build has no kernel for */
  DDMI_ERR_CUDA = 3          /* CUDA runtime error (launch, attribute, no device)    */
};

enum { DDMI_PREC_FP32 = 0, DDMI_PREC_BF16X3 = 1, DDMI_PREC_F16F8 = 2 };
/* Output store modes of the image / video decoders (the epilogues the reference's callers apply to the decoded signal):
 *   DDMI_STORE_F32               (batch, 3, n) fp32, the value the reference's forward returns
 *   DDMI_STORE_F32_CLAMP         same layout, clamp(x, -1, 1)                  evals/eval.py:162,226, tools/ldm/image.py:246
 *   DDMI_STORE_U8_CHANNELS_LAST  (batch, n, 3) uint8 = trunc((clamp(x,-1,1) + 1) * 127.5), i.e. image (b,h,w,c) / video
 *                                (b,t,h,w,c): rearrange(...).type(torch.uint8) of evals/eval.py:289,336-337          */
enum { DDMI_STORE_F32 = 0, DDMI_STORE_F32_CLAMP = 1, DDMI_STORE_U8_CHANNELS_LAST = 2 };

/* plane memory layout: as the reference's VAE decoder emits them, or channels-last */
enum { DDMI_LAYOUT_NCHW = 0, DDMI_LAYOUT_NHWC = 1 };

/* One positional-embedding plane batch: fp32, contiguous (batch, channels, height, width). */
typedef struct {
  const float* data;
  int32_t height;
  int32_t width;
} ddmi_plane_t;

/* Host-folded, packed MLP weights (device memory). */
typedef struct {
  int32_t precision;   /* DDMI_PREC_* */
  int32_t reserved;    /* bit 0: tcgen05 stream is packed for CTA pairs ([half 0 | half 1] per K step);
                          bit 1: image decode: never stage plane windows with TMA (diagnostic: every tile gathers directly)
                          bit 2: image decode, DDMI_PREC_F16F8: `program` drives the TMEM-resident-activation kernel
                                 (packing._pack_image_ts; needs vec_host) */
  const void* gemm;    /* GEMM operands, layout per precision (device)            */
  uint64_t gemm_bytes;
  const float* vec;    /* fp32 vectors: biases, folded constants, small heads (device) */
  uint64_t vec_floats;
  /* DDMI_PREC_BF16X3 / DDMI_PREC_F16F8: the MMA program that consumes `gemm`, one
     copy in device memory for the kernel and one in host memory for validation.     */
  const uint32_t* program;
  const uint32_t* program_host;
  uint64_t program_words;
  /* host copy of `vec` (may be NULL unless reserved bit 2 is set): small heads that a kernel takes as launch parameters
     (constant bank) instead of reading them from device memory -- the image decoder's ToRGB weights */
  const float* vec_host;
} ddmi_weights_t;

DDMI_API int ddmi_abi_version(void);
DDMI_API const char* ddmi_last_error(void);
DDMI_API const char* ddmi_status_string(int status);

/* Device properties the host side shards / sizes by.  Returns status. */
DDMI_API int ddmi_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/*
 * Image decode.  planes[s] = hdbf[s], s = 0 (coarse) .. 2 (fine), each
 * (batch, channels=64, S_s, S_s).  coord_x / coord_y: n_coords query positions in
 * [-1,1] (channel 0 / 1 of the reference's (1,2,h,w) coords tensor, flattened);
 * shared by all batch items.  Sampling: bilinear, border padding,
 * align_corners = false.  out: (batch, 3, n_coords) fp32.
 */
DDMI_API int ddmi_decode_image(const ddmi_plane_t planes[3], int32_t batch, int32_t channels,
                      const float* coord_x, const float* coord_y, int64_t n_coords,
                      const ddmi_weights_t* weights, float* out, void* stream);

/* ddmi_decode_image with an output store mode (DDMI_STORE_*); `out` is float* or uint8_t* accordingly. */
DDMI_API int ddmi_decode_image_store(const ddmi_plane_t planes[3], int32_t batch, int32_t channels,
                            const float* coord_x, const float* coord_y, int64_t n_coords,
                            const ddmi_weights_t* weights, int32_t store, void* out, void* stream);

/*
 * Image decode of a checkpoint whose NoiseInjection weights are non-zero (models/d2c_vae/blocks.py:286-297: every one of
 * the 12 StyledConv layers adds noise.weight * N(0,1)[b,1,h,w] before its bias / leaky ReLU; the reference draws the noise
 * inside forward).  noise_mode:
 *   DDMI_NOISE_NONE     no noise term (what ddmi_decode_image[_store] do; right when every noise.weight is 0)
 *   DDMI_NOISE_TENSORS  noise[l], l = 0..11 (net_res1.conv1, conv2, conv3, net_res2.conv1, ...): device pointers to
 *                       (batch, n_coords) fp32 -- the explicit `noise=` tensors of NoiseInjection.forward
 *   DDMI_NOISE_PHILOX   a counter-based N(0,1) stream keyed by `seed` (Philox4x32-10 + Box-Muller, defined in
 *                       csrc/common.cuh and restated in oracle/ddmi_oracle.py::philox_noise): the same values for every
 *                       precision, tiling and GPU count
 * The per-layer weights (noise.weight times the folded activation gain) are the last 12 floats of the packed vec blob.
 */
enum { DDMI_NOISE_NONE = 0, DDMI_NOISE_TENSORS = 1, DDMI_NOISE_PHILOX = 2 };
DDMI_API int ddmi_decode_image_noise(const ddmi_plane_t planes[3], int32_t batch, int32_t channels,
                            const float* coord_x, const float* coord_y, int64_t n_coords,
                            const ddmi_weights_t* weights, int32_t store, int32_t noise_mode,
                            const float* const* noise, uint64_t seed, void* out, void* stream);

/*
 * (batch, C, H, W) -> (batch, H, W, C).  Scattered queries (3-D points, ray samples) gather all
 * channels of a texel with float4 loads from the channels-last copy; planes are a few MB per item,
 * so the one-off re-layout is noise next to the decode.  src / dst: device, distinct.
 */
DDMI_API int ddmi_planes_to_channels_last(const float* src, float* dst, int32_t batch, int32_t channels,
                                          int32_t height, int32_t width, void* stream);

/*
 * Occupancy decode.  planes[a*3+s]: axis a = 0 'xy', 1 'yz', 2 'xz'; scale s =
 * 0..2; each (batch, 64, R_s, R_s).  points: (batch, n_points, 3) fp32 with
 * batch stride `point_batch_stride` floats (0 = the same points for every item).
 * Coordinates are normalised as normalize_coordinate(padding) does, then sampled
 * bilinear / border / align_corners = true and summed over the three axes.
 * logits: (batch, n_points) fp32.  plane_layout: DDMI_LAYOUT_NCHW (each (batch, 64, R, R)) or
 * DDMI_LAYOUT_NHWC (each (batch, R, R, 64), tcgen05 kernel only).
 */
DDMI_API int ddmi_decode_occupancy(const ddmi_plane_t planes[9], int32_t batch, int32_t channels, int32_t plane_layout,
                          const float* points, int64_t n_points, int64_t point_batch_stride,
                          float padding, const ddmi_weights_t* weights, float* logits,
                          void* stream);

/*
 * Video decode.  planes[a*3+s]: a = 0 'xy' (batch,64,Hs,Ws), 1 'yt' (batch,64,Ts,Hs),
 * 2 'xt' (batch,64,Ts,Ws).  coords_xy (2,H,W), coords_yt (2,T,H), coords_xt (2,T,W):
 * the reference's grids taken literally -- channel 0 indexes the plane's LAST
 * axis, channel 1 its second-to-last (so yt/xt are read with transposed axes,
 * SURVEY.md F6).  Features are concatenated [xy,yt,xt] per scale.
 * out: (batch, 3, T, H, W) fp32.
 */
DDMI_API int ddmi_decode_video(const ddmi_plane_t planes[9], int32_t batch, int32_t channels,
                      const float* coords_xy, const float* coords_yt, const float* coords_xt,
                      int32_t T, int32_t H, int32_t W,
                      const ddmi_weights_t* weights, float* out, void* stream);

/* ddmi_decode_video with an output store mode (DDMI_STORE_*); `out` is float* or uint8_t* accordingly. */
DDMI_API int ddmi_decode_video_store(const ddmi_plane_t planes[9], int32_t batch, int32_t channels,
                            const float* coords_xy, const float* coords_yt, const float* coords_xt,
                            int32_t T, int32_t H, int32_t W, const ddmi_weights_t* weights, int32_t store,
                            void* out, void* stream);

/*
 * Occupancy logits on a query LATTICE {xs[i]} x {ys[j]} x {zs[k]} (ABI 10): the dense grid of the reference's mesh generator,
 * box_size * make_3d_grid(...) (convocc/src/conv_onet/generation.py:90-97, convocc/src/common.py:145-164).  axes = device array
 * [xs (nx) | ys (ny) | zs (nz)]; logits: (batch, nx * ny * nz), point index (i * ny + j) * nz + k -- bit-identical to
 * ddmi_decode_occupancy on the expanded point list.  On a lattice the 'xy' sample depends on (i, j) only, 'yz' on (j, k), 'xz'
 * on (i, k): the nx ny + ny nz + nx nz distinct vectors per scale are sampled once into `workspace`
 * (ddmi_occupancy_lattice_workspace_bytes() bytes, 16-byte aligned, scratch for this call) and every point reads three
 * records instead of twelve texels.  tcgen05 precisions, pair-packed weights.
 */
DDMI_API int64_t ddmi_occupancy_lattice_workspace_bytes(int32_t batch, int32_t nx, int32_t ny, int32_t nz);
DDMI_API int ddmi_decode_occupancy_lattice(const ddmi_plane_t planes[9], int32_t batch, int32_t channels, int32_t plane_layout,
                                  const float* axes, int32_t nx, int32_t ny, int32_t nz, float padding,
                                  const ddmi_weights_t* weights, float* logits, void* workspace,
                                  uint64_t workspace_bytes, void* stream);

/*
 * ddmi_decode_video_store with a caller-provided device workspace (ABI 10).  The three query grids of a video are separable
 * by construction (xy by (h,w), yt by (t,h), xt by (t,w): utils/general_utils.py:38-52), so only H W + T H + T W distinct
 * feature vectors exist per scale; with a workspace of ddmi_video_workspace_bytes() bytes (16-byte aligned; 0 = this
 * precision has no table path) the DDMI_PREC_F16F8 kernel samples each of them once into operand-format tables and the decode
 * reads those instead of gathering 36 taps x 64 channels per voxel.  Results are bit-identical to workspace = NULL (direct gathers).
 * The workspace is scratch: it may be reused or freed once the call's work on `stream` has completed.
 */
DDMI_API int64_t ddmi_video_workspace_bytes(int32_t batch, int32_t T, int32_t H, int32_t W, int32_t precision);
DDMI_API int ddmi_decode_video_ws(const ddmi_plane_t planes[9], int32_t batch, int32_t channels,
                         const float* coords_xy, const float* coords_yt, const float* coords_xt,
                         int32_t T, int32_t H, int32_t W, const ddmi_weights_t* weights, int32_t store,
                         void* out, void* workspace, uint64_t workspace_bytes, void* stream);

/*
 * NeRF MLP on pre-embedded rows.  x: (n, 186) = [latent 96 | embed(pts) 63 |
 * embed(dir) 27] (or (n,159) when sigma_only).  out: (n,4) = [rgb, sigma]
 * (or (n,1)).  `negative_slope` is the LeakyReLU slope (the reference's
 * nn.LeakyReLU(True) => 1.0).
 */
DDMI_API int ddmi_nerf_mlp(const float* x, int64_t n, int32_t x_stride, int32_t sigma_only,
                  float negative_slope, const ddmi_weights_t* weights, float* out,
                  void* stream);

/*
 * NeRF ray render: sample generation, triplane gather, positional embedding,
 * MLP and volume compositing for `batch` objects that share one ray set.
 * planes[a] = 'xy','yz','xz', each (batch, 32, R, R).  rays: (n_rays, ray_stride)
 * rows [o(3) d(3) near far viewdir(3)].  t_vals: (n_samples) the caller's
 * linspace(0,1,n_samples).  pts are divided by `plane_extent` (3.5) before
 * sampling (bilinear / border / align_corners = true).
 * plane_layout: DDMI_LAYOUT_NCHW (fp32 kernel) or DDMI_LAYOUT_NHWC (tcgen05 kernel, each (batch, R, R, 32)).
 * rgb_map: (batch, n_rays, 3).  raw: optional (batch, n_rays, n_samples, 4)
 * model outputs [rgb, sigma] (pass NULL to skip the store when the kernel
 * composites in place; some kernels need it as workspace and then it is required
 * -- the function says so with DDMI_ERR_BAD_ARG).
 */
DDMI_API int ddmi_nerf_render(const ddmi_plane_t planes[3], int32_t batch, int32_t channels, int32_t plane_layout,
                     const float* rays, int64_t n_rays, int32_t ray_stride,
                     const float* t_vals, int32_t n_samples, float plane_extent,
                     float negative_slope, int32_t white_bkgd,
                     const ddmi_weights_t* weights, float* rgb_map, float* raw,
                     void* stream);

/*
 * Same with per-ray sample depths: z_vals (n_rays, n_samples) fp32 (device), shared by all batch items, replaces
 * near * (1 - t) + far * t.  This is how the host side passes the reference's stratified `perturb` samples and `lindisp`
 * spacing (utils/nerf_helpers.py:359-380); compositing uses the same table for its distances (:487-495).
 */
DDMI_API int ddmi_nerf_render_z(const ddmi_plane_t planes[3], int32_t batch, int32_t channels, int32_t plane_layout,
                       const float* rays, int64_t n_rays, int32_t ray_stride,
                       const float* z_vals, int32_t n_samples, float plane_extent,
                       float negative_slope, int32_t white_bkgd, const ddmi_weights_t* weights,
                       float* rgb_map, float* raw, void* stream);

/*
 * Hierarchical (inverse-CDF) sampling of utils/nerf_helpers.py:166-209.  bins (n_rays, n_bins), weights (n_rays, n_bins - 1),
 * u (n_rays, n_samples) uniform numbers in [0,1) drawn by the caller (torch.rand / linspace, as the reference does);
 * out (n_rays, n_samples) fp32.  All device memory.
 */
DDMI_API int ddmi_sample_pdf(const float* bins, const float* weights, const float* u, int64_t n_rays, int32_t n_bins,
                             int32_t n_samples, float* out, void* stream);

/*
 * Plane-producer tail: the layers of the D2C-VAE decoders that EMIT the PE planes, fused, writing each
 * plane once in the layout its consumer gathers from (out_layout: DDMI_LAYOUT_NCHW (batch, C_out, H, W) for the image / video
 * decoders, DDMI_LAYOUT_NHWC (batch, H, W, C_out) for the scattered-query decoders -- hand the latter to the occupancy / NeRF
 * entry points with plane_layout = DDMI_LAYOUT_NHWC and no transposition runs).  fp32, all device memory.
 *   ddmi_plane_head: out = conv1x1(h; weight, bias)                                     -- `up[i].hdbf[0]`, nn.Conv2d(block_in, out_ch, 1)
 *   ddmi_plane_tail: out = [tanh] conv3x3(swish(GroupNorm(h; groups, eps, gn_weight, gn_bias)); weight, bias), zero
 *                    padding 1 -- `norm_out` (GroupNorm(32, eps 1e-6)), x * sigmoid(x), `conv_out`, `tanh_out`.
 *                    stats: scratch of batch * groups * 2 floats (mean / rstd per item and group).
 * `weight` is nn.Conv2d's (C_out, C_in, k, k) tensor TRANSPOSED ONCE by the caller to (C_in, k * k, C_out), contiguous (the kernels
 * stream input-channel chunks of it with vector loads).
 * C_out must be 64 or 32.
 */
DDMI_API int ddmi_plane_head(const float* h, int32_t batch, int32_t in_channels, int32_t height, int32_t width, const float* weight,
                             const float* bias, int32_t out_channels, int32_t out_layout, float* out, void* stream);
DDMI_API int ddmi_plane_tail(const float* h, int32_t batch, int32_t in_channels, int32_t height, int32_t width, const float* gn_weight,
                             const float* gn_bias, int32_t groups, float eps, const float* weight, const float* bias,
                             int32_t out_channels, int32_t tanh_out, int32_t out_layout, float* stats, float* out, void* stream);

/*
 * Occupancy post-step: marching cubes on a decoded logit grid, on the GPU (the reference copies the grid to the host and runs
 * libmcubes there: generation.py:130-144,166-168).  grid: (nx, ny, nz) fp32 device memory, z fastest -- the order of
 * make_3d_grid / value_grid (generation.py:90-97).  pad = 1 surrounds it with one layer of pad_value (the reference pads with
 * -1e6 so the mesh closes: generation.py:164-165), pad = 0 takes it as is.  The mesh is the reference's bit for bit: vertices
 * (float64, + 0.5 grid units as libmcubes leaves them) and triangles in the order its sequential sweep appends them, the
 * duplicated vertices on the low faces of the volume included.
 *   1. ddmi_mcubes_workspace_bytes -> size of the scratch buffer (16 B per cell of the padded volume)
 *   2. ddmi_mcubes_count: classify + prefix sums; totals_dev[0] = vertices, totals_dev[1] = triangle corners (3 per triangle),
 *      two uint64 in device memory -- read them back to size the outputs
 *   3. ddmi_mcubes_emit: vertices (totals[0] x 3 float64) and triangles (totals[1] int64 vertex indices), device memory.
 *      affine = NULL, or 7 host doubles {sub0, sub1, div_x, div_y, div_z, sub2, mul}: every vertex component becomes
 *      mul * ((((v - sub0) - sub1) / div_axis) - sub2), float64, in that order -- Generator3D.extract_mesh's
 *      "vertices -= 0.5; vertices -= 1; vertices /= (n - 1); vertices = box_size * (vertices - 0.5)".
 * At most 2^28 - 1 cells per call.
 */
DDMI_API int ddmi_mcubes_workspace_bytes(int32_t nx, int32_t ny, int32_t nz, int32_t pad, uint64_t* bytes);
DDMI_API int ddmi_mcubes_count(const float* grid, int32_t nx, int32_t ny, int32_t nz, int32_t pad, double pad_value, double isovalue,
                               void* workspace, uint64_t workspace_bytes, uint64_t* totals_dev, void* stream);
DDMI_API int ddmi_mcubes_emit(const float* grid, int32_t nx, int32_t ny, int32_t nz, int32_t pad, double pad_value, double isovalue,
                              const void* workspace, const double* affine, double* vertices, int64_t* triangles, void* stream);

/*
 * Bring-up self test of the plane-window TMA path (cp.async.bulk.tensor.3d over an NCHW fp32 plane batch): the 64 (x) x 2 (y)
 * x 64 (channel) box at element coordinates (x, y, c) -> out[64][2][64] fp32 (channel-major), out-of-range elements 0.
 * variant 0 hands the tensor map to the kernel as a parameter, 1 through `map_dev` (128 B of device memory, 64-byte aligned).
 */
DDMI_API int ddmi_selftest_tma(const float* plane, int32_t batch, int32_t channels, int32_t height, int32_t width, int32_t x,
                               int32_t y, int32_t c, int32_t variant, void* map_dev, float* out, void* stream);

/*
 * Bring-up self test of the tcgen05 path: one 128 x N x K bf16 GEMM through the
 * same descriptors / TMEM epilogue the decode kernels use.  a: (128,K) fp32,
 * b: (N,K) fp32 (device); d: (128,N) fp32 = a * b^T computed with the bf16x3
 * split.  N in {16..256, multiple of 16}, K multiple of 16, K <= 256.
 */
DDMI_API int ddmi_selftest_umma(const float* a, const float* b, float* d, int32_t N, int32_t K,
                       void* stream);

/*
 * Same for the CTA-pair path (tcgen05.mma.cta_group::2, M = 256 over a 2-CTA cluster):
 * a: (256,K), b: (N,K), d: (256,N); CTA r owns rows 128r.. of a / d and rows (N/2)r.. of b.
 */
DDMI_API int ddmi_selftest_umma2(const float* a, const float* b, float* d, int32_t N, int32_t K,
                                 void* stream);

/*
 * Same shapes as ddmi_selftest_umma through the DDMI_PREC_F16F8 operand scheme (one fp16 kind::f16 term + two
 * kind::f8f6f4 correction terms, e5m2 x e4m3, into one accumulator); K a multiple of 32.
 */
DDMI_API int ddmi_selftest_f16f8(const float* a, const float* b, float* d, int32_t N, int32_t K,
                                 void* stream);

/*
 * Diagnostics.  The shipping library is built WITHOUT in-kernel instrumentation: these return zeros / an empty trace
 * unless the library was built with -DDDMI_PROFILE=1 (`make -C ddmi_b200/csrc prof` -> libddmi_b200_prof.so, loaded by
 * the dev tools through DDMI_B200_LIB).
 * ddmi_debug_profile: cycle counters accumulated by CTA 0 of the tcgen05 image kernel since the last
 * reset (synchronises the device).  out[0] epilogue thread: cycles parked waiting for MMA groups,
 * [1] cycles in epilogue stages, [2] cycles in plane gathers, [3] MMA thread: cycles waiting for
 * operands, [4] cycles waiting for weight chunks, [5] MMA thread total, [6] tiles, [7] spare.
 * ddmi_debug_trace: (event id << 48 | SM clock) records of one tile iteration of CTA 0 (epilogue thread 0 and the MMA
 * lane); *count = records copied to `out` (host memory, `capacity` entries).
 * ddmi_debug_microbench: epilogue building blocks in isolation (csrc/microbench.cu); out_dev[0] = cycles warp 0 spent,
 * out_dev[1] = span over the 8 warps, for `iters` repetitions of one stage-sized unit.  All buffers are device memory.
 * Modes 100 + v (one CTA pair) / 200 + v (every SM): tcgen05.mma rate, `iters` rounds of 8 MMAs; v bit 0: N = 128 (else 256),
 * bit 1: A operand in tensor memory, bit 2: FP8, bit 3: the f16f8 mix, bit 4: concurrent shared-memory stores; out_dev[0] =
 * cycles, out_dev[1] = MMAs issued.
 */
DDMI_API int ddmi_debug_profile(uint64_t out[8], int32_t reset);
DDMI_API int ddmi_debug_trace(uint64_t* out, int32_t capacity, int32_t* count, int32_t reset);
/* profiling build only: what-if switches of the image kernel (bit 0: epilogue stages only signal, bit 1: the issuer skips
 * the MMA instructions, bit 2: no plane gathers) -- outputs are garbage, the timings bound each side of the pipeline */
DDMI_API int ddmi_debug_set(int32_t flags);
/* L2 -> shared-memory bulk-copy stream in isolation (csrc/microbench.cu::ringbench_kernel): `ctas` CTAs (one per SM) each
 * stream `iters` slots of `slot_bytes` from a `span_bytes` window at `src` through a ring of `nslots` slots with no consumer;
 * out_dev[0] = cycles CTA 0 took. */
/* scattered 256-byte texel gathers in isolation (csrc/microbench.cu::gatherbench_kernel): load form `variant` (0 ld.global.nc,
 * 1 .cg, 2 .nc.L1::no_allocate, 3 .cv, 4 L1::evict_first), `unroll` (4 or 12) texels in flight per thread, `smem_kb` of
 * dynamic shared memory held by each of `ctas` CTAs (what is left of the 228 KB is L1); sink_dev: ctas * 256 floats. */
DDMI_API int ddmi_debug_gatherbench(int32_t variant, int32_t unroll, const float* table, uint32_t ntexel, int32_t iters,
                                    int32_t smem_kb, int32_t ctas, uint64_t* out_dev, float* sink_dev, void* stream);
DDMI_API int ddmi_debug_ringbench(const void* src, uint64_t span_bytes, int32_t slot_bytes, int32_t nslots, int32_t iters,
                                  int32_t ctas, uint64_t* out_dev, void* stream);
DDMI_API int ddmi_debug_microbench(int32_t mode, int32_t iters, const float* seed, uint64_t* out_dev, float* sink_dev,
                                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DDMI_B200_H */
