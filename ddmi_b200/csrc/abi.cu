// extern "C" boundary of libddmi_b200.so: argument validation + dispatch.
// Declarations and the reference interfaces they replace: include/ddmi_b200.h.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace ddmi {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return DDMI_ERR_CUDA;
}

// fp32 CUDA-core path (decode_fp32.cu)
int launch_image_fp32(const PlaneSet&, int, int, const float*, const float*, long long, const float*, const float*, void*, int, const NoiseArgs&, cudaStream_t);
int launch_occupancy_fp32(const PlaneSet&, int, int, const float*, long long, long long, float, float, const float*, const float*, float*, cudaStream_t);
int launch_video_fp32(const PlaneSet&, int, int, const float*, const float*, const float*, int, int, int, const float*, const float*, void*, int, cudaStream_t);
int launch_nerf_mlp_fp32(const float*, long long, int, int, float, const float*, const float*, float*, cudaStream_t);
int launch_nerf_render_fp32(const PlaneSet&, int, int, const float*, long long, int, const float*, int, int, float, float, int, const float*, const float*, float*, float*, cudaStream_t);
// tcgen05 path (decode_umma.cu)
int launch_image_umma(const PlaneSet&, int, int, const float*, const float*, long long, const void*, size_t, const uint32_t*, size_t, const uint32_t*, const float*, size_t, void*, int, int, int, const NoiseArgs&, int, int, const float*, cudaStream_t);
int launch_selftest_umma(const float*, const float*, float*, int, int, cudaStream_t);
int launch_occupancy_umma_entry(const PlaneSet&, int, int, const float*, long long, long long, float, float, const void*, size_t, const uint32_t*, size_t, const uint32_t*, const float*, size_t, float*, int, int, int, cudaStream_t);
int launch_planes_to_nhwc(const float*, float*, int, int, int, cudaStream_t);
int launch_video_umma_entry(const PlaneSet&, int, int, const float*, const float*, const float*, int, int, int, const void*, size_t, const uint32_t*, size_t, const uint32_t*, const float*, size_t, void*, int, int, int, void*, size_t, cudaStream_t);
size_t video_workspace_bytes(int, int, int, int);
size_t occupancy_lattice_workspace_bytes(int, int, int, int);
int launch_occupancy_lattice_umma_entry(const PlaneSet&, int, int, const float*, int, int, int, float, float, const void*, size_t, const uint32_t*, size_t, const uint32_t*, const float*, size_t, float*, int, int, int, void*, size_t, cudaStream_t);
int launch_nerf_composite(const float*, const float*, int, const float*, int, int, long long, int, int, float*, cudaStream_t);
int launch_nerf_umma_entry(const PlaneSet&, int, int, const float*, long long, int, const float*, int, int, float, float, int, const void*, size_t, const uint32_t*, size_t, const uint32_t*, const float*, size_t, float*, float*, int, int, cudaStream_t);
int debug_profile(unsigned long long*, int);
int debug_trace(unsigned long long*, int, int*, int);
int debug_set(int);
int launch_tma_selftest(const float*, int, int, int, int, int, int, int, int, void*, float*, cudaStream_t);
int launch_sample_pdf(const float*, const float*, const float*, long long, int, int, float*, cudaStream_t);
int launch_gatherbench(int, int, const float*, unsigned, int, int, int, unsigned long long*, float*, cudaStream_t);
int launch_ringbench(const void*, unsigned long long, int, int, int, int, unsigned long long*, cudaStream_t);
int launch_microbench(int, int, const float*, unsigned long long*, float*, cudaStream_t);
int launch_selftest_umma2(const float*, const float*, float*, int, int, cudaStream_t);
int launch_selftest_f16f8(const float*, const float*, float*, int, int, cudaStream_t);

static int check_planes(const ddmi_plane_t* planes, int count, PlaneSet* ps) {
  DDMI_REQUIRE(planes != nullptr, "planes is NULL");
  for (int i = 0; i < count; ++i) {
    DDMI_REQUIRE(planes[i].data != nullptr, "planes[%d].data is NULL", i);
    DDMI_REQUIRE(planes[i].height >= 1 && planes[i].width >= 1, "planes[%d] has empty extent %dx%d", i,
                 planes[i].height, planes[i].width);
    ps->data[i] = planes[i].data;
    ps->h[i] = planes[i].height;
    ps->w[i] = planes[i].width;
  }
  return DDMI_OK;
}

static int check_weights(const ddmi_weights_t* w, uint64_t need_gemm_bytes, uint64_t need_vec) {
  DDMI_REQUIRE(w != nullptr, "weights is NULL");
  DDMI_REQUIRE(w->gemm != nullptr && w->vec != nullptr, "weights->gemm / weights->vec is NULL");
  DDMI_REQUIRE(((uintptr_t)w->gemm & 127) == 0, "weights->gemm must be 128-byte aligned");
  DDMI_REQUIRE(((uintptr_t)w->vec & 15) == 0, "weights->vec must be 16-byte aligned");
  DDMI_REQUIRE(w->gemm_bytes == need_gemm_bytes, "packed gemm blob is %llu bytes, this decoder expects %llu",
               (unsigned long long)w->gemm_bytes, (unsigned long long)need_gemm_bytes);
  DDMI_REQUIRE(w->vec_floats == need_vec, "packed vec blob is %llu floats, this decoder expects %llu",
               (unsigned long long)w->vec_floats, (unsigned long long)need_vec);
  return DDMI_OK;
}

int launch_plane_conv(const float*, int, int, int, int, int, const float*, const float*, int, float, const float*, const float*, int, int, int, float*, float*, cudaStream_t);
int mcubes_workspace_bytes(int, int, int, int, unsigned long long*);
int launch_mcubes_count(const float*, int, int, int, int, double, double, void*, unsigned long long, unsigned long long*, cudaStream_t);
int launch_mcubes_emit(const float*, int, int, int, int, double, double, const void*, const double*, double*, long long*, cudaStream_t);
}  // namespace ddmi

using namespace ddmi;

extern "C" {

DDMI_API int ddmi_abi_version(void) { return DDMI_ABI_VERSION; }

DDMI_API const char* ddmi_last_error(void) { return g_err; }

DDMI_API const char* ddmi_status_string(int status) {
  switch (status) {
    case DDMI_OK: return "ok";
    case DDMI_ERR_BAD_ARG: return "bad argument";
    case DDMI_ERR_UNSUPPORTED: return "unsupported configuration";
    case DDMI_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

DDMI_API int ddmi_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  DDMI_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DDMI_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return DDMI_OK;
}

DDMI_API int ddmi_decode_image(const ddmi_plane_t planes[3], int32_t batch, int32_t channels,
                      const float* coord_x, const float* coord_y, int64_t n_coords,
                      const ddmi_weights_t* weights, float* out, void* stream) {
  return ddmi_decode_image_store(planes, batch, channels, coord_x, coord_y, n_coords, weights, DDMI_STORE_F32, out, stream);
}

DDMI_API int ddmi_decode_image_store(const ddmi_plane_t planes[3], int32_t batch, int32_t channels,
                                     const float* coord_x, const float* coord_y, int64_t n_coords,
                                     const ddmi_weights_t* weights, int32_t store, void* out, void* stream) {
  return ddmi_decode_image_noise(planes, batch, channels, coord_x, coord_y, n_coords, weights, store, DDMI_NOISE_NONE, nullptr,
                                 0, out, stream);
}

DDMI_API int ddmi_decode_image_noise(const ddmi_plane_t planes[3], int32_t batch, int32_t channels,
                                     const float* coord_x, const float* coord_y, int64_t n_coords,
                                     const ddmi_weights_t* weights, int32_t store, int32_t noise_mode,
                                     const float* const* noise, uint64_t seed, void* out, void* stream) {
  DDMI_REQUIRE(store >= DDMI_STORE_F32 && store <= DDMI_STORE_U8_CHANNELS_LAST, "unknown store mode %d", store);
  DDMI_REQUIRE(noise_mode >= DDMI_NOISE_NONE && noise_mode <= DDMI_NOISE_PHILOX, "unknown noise mode %d", noise_mode);
  NoiseArgs na = {};
  na.mode = noise_mode;
  na.seed = seed;
  if (noise_mode == DDMI_NOISE_TENSORS) {
    DDMI_REQUIRE(noise != nullptr, "noise_mode = DDMI_NOISE_TENSORS needs 12 noise pointers");
    for (int l = 0; l < 12; ++l) {
      DDMI_REQUIRE(noise[l] != nullptr, "noise[%d] is NULL", l);
      na.p[l] = noise[l];
    }
  }
  PlaneSet ps = {};
  int rc = check_planes(planes, 3, &ps);
  if (rc) return rc;
  DDMI_REQUIRE(batch >= 1, "batch must be >= 1 (got %d)", batch);
  DDMI_REQUIRE(n_coords >= 1, "n_coords must be >= 1 (got %lld)", (long long)n_coords);
  DDMI_REQUIRE(coord_x && coord_y && out, "coord_x / coord_y / out is NULL");
  if (channels != 64) {
    set_error("image decode is built for latent_dim = 64 planes (got %d channels)", channels);
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(weights != nullptr, "weights is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (weights->precision == DDMI_PREC_FP32) {
    // res1: c1[64] c2 c3 skip[64]; res2/3: c1[320] c2 c3 skip[320]; res4: c1 c2 c3
    const uint64_t kfl = (64 + 256 + 256 + 64) + 2 * (320 + 256 + 256 + 320) + 3 * 256;
    rc = check_weights(weights, kfl * 256 * sizeof(float), 4096 + 768 + 3 + 12);
    if (rc) return rc;
    return launch_image_fp32(ps, batch, channels, coord_x, coord_y, n_coords, (const float*)weights->gemm,
                             weights->vec, out, store, na, st);
  } else if (weights->precision == DDMI_PREC_BF16X3 || weights->precision == DDMI_PREC_F16F8) {
    DDMI_REQUIRE(weights->gemm && weights->vec, "weights->gemm / weights->vec is NULL");
    DDMI_REQUIRE(((uintptr_t)weights->gemm & 127) == 0, "weights->gemm must be 128-byte aligned");
    const int f16f8 = weights->precision == DDMI_PREC_F16F8;
    DDMI_REQUIRE(!f16f8 || (weights->reserved & 1), "DDMI_PREC_F16F8 weights must be packed for CTA pairs");
    return launch_image_umma(ps, batch, channels, coord_x, coord_y, n_coords, weights->gemm, weights->gemm_bytes,
                             weights->program_host, weights->program_words, weights->program, weights->vec,
                             weights->vec_floats, out, store, weights->reserved & 1, f16f8, na, (weights->reserved >> 1) & 1,
                             (weights->reserved >> 2) & 1, weights->vec_host, st);
  }
  set_error("unknown precision %d", weights->precision);
  return DDMI_ERR_UNSUPPORTED;
}

DDMI_API int ddmi_planes_to_channels_last(const float* src, float* dst, int32_t batch, int32_t channels, int32_t height,
                                          int32_t width, void* stream) {
  DDMI_REQUIRE(src && dst && src != dst, "src / dst is NULL or aliased");
  DDMI_REQUIRE(batch >= 1 && channels >= 1 && height >= 1 && width >= 1, "empty plane");
  DDMI_REQUIRE(batch <= 65535, "batch %d too large for one launch", batch);
  return launch_planes_to_nhwc(src, dst, batch, channels, height * width, (cudaStream_t)stream);
}

DDMI_API int ddmi_decode_occupancy(const ddmi_plane_t planes[9], int32_t batch, int32_t channels, int32_t plane_layout,
                          const float* points, int64_t n_points, int64_t point_batch_stride,
                          float padding, const ddmi_weights_t* weights, float* logits, void* stream) {
  PlaneSet ps = {};
  int rc = check_planes(planes, 9, &ps);
  if (rc) return rc;
  DDMI_REQUIRE(batch >= 1, "batch must be >= 1 (got %d)", batch);
  DDMI_REQUIRE(n_points >= 1, "n_points must be >= 1 (got %lld)", (long long)n_points);
  DDMI_REQUIRE(points && logits, "points / logits is NULL");
  DDMI_REQUIRE(plane_layout == DDMI_LAYOUT_NCHW || plane_layout == DDMI_LAYOUT_NHWC, "unknown plane_layout %d", plane_layout);
  DDMI_REQUIRE(point_batch_stride == 0 || point_batch_stride >= 3 * n_points,
               "point_batch_stride %lld overlaps items", (long long)point_batch_stride);
  if (channels != 64) {
    set_error("occupancy decode is built for latent_dim = 64 planes (got %d channels)", channels);
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(weights != nullptr, "weights is NULL");
  // the Python scalars of normalize_coordinate: double arithmetic, then fp32
  const float divisor = (float)(1.0 + (double)padding + 10e-6);
  const float upper = (float)(1.0 - 10e-6);
  if (weights->precision == DDMI_PREC_BF16X3 || weights->precision == DDMI_PREC_F16F8) {
    DDMI_REQUIRE(weights->gemm && weights->vec, "weights->gemm / weights->vec is NULL");
    DDMI_REQUIRE(((uintptr_t)weights->gemm & 127) == 0, "weights->gemm must be 128-byte aligned");
    return launch_occupancy_umma_entry(ps, batch, channels, points, n_points, point_batch_stride, divisor, upper,
                                       weights->gemm, weights->gemm_bytes, weights->program_host, weights->program_words,
                                       weights->program, weights->vec, weights->vec_floats, logits,
                                       weights->reserved & 1, plane_layout, weights->precision == DDMI_PREC_F16F8,
                                       (cudaStream_t)stream);
  }
  if (plane_layout != DDMI_LAYOUT_NCHW) {
    set_error("the fp32 occupancy kernel reads NCHW planes only");
    return DDMI_ERR_UNSUPPORTED;
  }
  if (weights->precision != DDMI_PREC_FP32) {
    set_error("occupancy decode: unknown precision %d", weights->precision);
    return DDMI_ERR_UNSUPPORTED;
  }
  const uint64_t gfl = (64 * 64 + 64 * 256 + 64 * 256) + 2 * ((320 + 320 + 256) * 256) + 2 * 256 * 256;
  rc = check_weights(weights, gfl * sizeof(float), 1856 + 768 + 256 + 1);
  if (rc) return rc;
  return launch_occupancy_fp32(ps, batch, channels, points, n_points, point_batch_stride, divisor, upper,
                               (const float*)weights->gemm, weights->vec, logits, (cudaStream_t)stream);
}

DDMI_API int64_t ddmi_occupancy_lattice_workspace_bytes(int32_t batch, int32_t nx, int32_t ny, int32_t nz) {
  if (batch < 1 || nx < 1 || ny < 1 || nz < 1 || (int64_t)nx * ny * nz > 2147483647LL) return 0;
  return (int64_t)occupancy_lattice_workspace_bytes(batch, nx, ny, nz);
}

DDMI_API int ddmi_decode_occupancy_lattice(const ddmi_plane_t planes[9], int32_t batch, int32_t channels, int32_t plane_layout,
                                           const float* axes, int32_t nx, int32_t ny, int32_t nz, float padding,
                                           const ddmi_weights_t* weights, float* logits, void* workspace,
                                           uint64_t workspace_bytes, void* stream) {
  PlaneSet ps = {};
  int rc = check_planes(planes, 9, &ps);
  if (rc) return rc;
  DDMI_REQUIRE(batch >= 1, "batch must be >= 1 (got %d)", batch);
  DDMI_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1 && (int64_t)nx * ny * nz <= 2147483647LL, "lattice %d x %d x %d", nx, ny, nz);
  DDMI_REQUIRE(axes && logits && workspace, "axes / logits / workspace is NULL");
  DDMI_REQUIRE(plane_layout == DDMI_LAYOUT_NCHW || plane_layout == DDMI_LAYOUT_NHWC, "unknown plane_layout %d", plane_layout);
  if (channels != 64) {
    set_error("occupancy decode is built for latent_dim = 64 planes (got %d channels)", channels);
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(weights != nullptr, "weights is NULL");
  if (weights->precision != DDMI_PREC_BF16X3 && weights->precision != DDMI_PREC_F16F8) {
    set_error("lattice queries exist for the tcgen05 precisions only (expand the lattice and call ddmi_decode_occupancy)");
    return DDMI_ERR_UNSUPPORTED;
  }
  if (!(weights->reserved & 1)) {
    set_error("lattice queries need weights packed for CTA pairs");
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(weights->gemm && weights->vec, "weights->gemm / weights->vec is NULL");
  DDMI_REQUIRE(((uintptr_t)weights->gemm & 127) == 0, "weights->gemm must be 128-byte aligned");
  const float divisor = (float)(1.0 + (double)padding + 10e-6);
  const float upper = (float)(1.0 - 10e-6);
  return launch_occupancy_lattice_umma_entry(ps, batch, channels, axes, nx, ny, nz, divisor, upper, weights->gemm,
                                             weights->gemm_bytes, weights->program_host, weights->program_words,
                                             weights->program, weights->vec, weights->vec_floats, logits, 1, plane_layout,
                                             weights->precision == DDMI_PREC_F16F8, workspace, (size_t)workspace_bytes,
                                             (cudaStream_t)stream);
}

DDMI_API int ddmi_decode_video(const ddmi_plane_t planes[9], int32_t batch, int32_t channels,
                      const float* coords_xy, const float* coords_yt, const float* coords_xt,
                      int32_t T, int32_t H, int32_t W, const ddmi_weights_t* weights, float* out,
                      void* stream) {
  return ddmi_decode_video_store(planes, batch, channels, coords_xy, coords_yt, coords_xt, T, H, W, weights,
                                 DDMI_STORE_F32, out, stream);
}

DDMI_API int ddmi_decode_video_store(const ddmi_plane_t planes[9], int32_t batch, int32_t channels,
                                     const float* coords_xy, const float* coords_yt, const float* coords_xt,
                                     int32_t T, int32_t H, int32_t W, const ddmi_weights_t* weights, int32_t store,
                                     void* out, void* stream) {
  return ddmi_decode_video_ws(planes, batch, channels, coords_xy, coords_yt, coords_xt, T, H, W, weights, store, out, nullptr, 0,
                              stream);
}

DDMI_API int64_t ddmi_video_workspace_bytes(int32_t batch, int32_t T, int32_t H, int32_t W, int32_t precision) {
  if (precision != DDMI_PREC_F16F8 || batch < 1 || T < 1 || H < 1 || W < 1) return 0;
  return (int64_t)video_workspace_bytes(batch, T, H, W);
}

DDMI_API int ddmi_decode_video_ws(const ddmi_plane_t planes[9], int32_t batch, int32_t channels,
                                  const float* coords_xy, const float* coords_yt, const float* coords_xt,
                                  int32_t T, int32_t H, int32_t W, const ddmi_weights_t* weights, int32_t store,
                                  void* out, void* workspace, uint64_t workspace_bytes, void* stream) {
  DDMI_REQUIRE(store >= DDMI_STORE_F32 && store <= DDMI_STORE_U8_CHANNELS_LAST, "unknown store mode %d", store);
  PlaneSet ps = {};
  int rc = check_planes(planes, 9, &ps);
  if (rc) return rc;
  DDMI_REQUIRE(batch >= 1, "batch must be >= 1 (got %d)", batch);
  DDMI_REQUIRE(T >= 1 && H >= 1 && W >= 1, "empty query volume %dx%dx%d", T, H, W);
  DDMI_REQUIRE(coords_xy && coords_yt && coords_xt && out, "coords / out is NULL");
  if (channels != 64) {
    set_error("video decode is built for latent_dim = 64 planes (got %d channels)", channels);
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(weights != nullptr, "weights is NULL");
  if (weights->precision == DDMI_PREC_BF16X3 || weights->precision == DDMI_PREC_F16F8) {
    DDMI_REQUIRE(weights->gemm && weights->vec, "weights->gemm / weights->vec is NULL");
    DDMI_REQUIRE(((uintptr_t)weights->gemm & 127) == 0, "weights->gemm must be 128-byte aligned");
    return launch_video_umma_entry(ps, batch, channels, coords_xy, coords_yt, coords_xt, T, H, W, weights->gemm,
                                   weights->gemm_bytes, weights->program_host, weights->program_words, weights->program,
                                   weights->vec, weights->vec_floats, out, store, weights->reserved & 1,
                                   weights->precision == DDMI_PREC_F16F8, workspace, (size_t)workspace_bytes,
                                   (cudaStream_t)stream);
  }
  if (weights->precision != DDMI_PREC_FP32) {
    set_error("video decode: unknown precision %d", weights->precision);
    return DDMI_ERR_UNSUPPORTED;
  }
  const uint64_t gfl = (192 * 192 + 192 * 256 + 192 * 256) + 2 * ((448 + 448 + 256) * 256) + 2 * 256 * 256;
  rc = check_weights(weights, gfl * sizeof(float), 448 + 3 * 512 + 768 + 3);
  if (rc) return rc;
  return launch_video_fp32(ps, batch, channels, coords_xy, coords_yt, coords_xt, T, H, W,
                           (const float*)weights->gemm, weights->vec, out, store, (cudaStream_t)stream);
}

static const uint64_t kNerfGemmFloats =
    (uint64_t)(160 + 256 + (160 + 256) + 256 + (160 + 256) + 256 + 256) * 256 + (256 + 32) * 128;
static const uint64_t kNerfVecFloats = 7 * 256 + 128 + 256 + 1 + 384 + 3;

DDMI_API int ddmi_nerf_mlp(const float* x, int64_t n, int32_t x_stride, int32_t sigma_only, float negative_slope,
                  const ddmi_weights_t* weights, float* out, void* stream) {
  DDMI_REQUIRE(x && out, "x / out is NULL");
  DDMI_REQUIRE(n >= 1, "n must be >= 1 (got %lld)", (long long)n);
  DDMI_REQUIRE(x_stride >= (sigma_only ? 159 : 186), "x_stride %d too small", x_stride);
  DDMI_REQUIRE(weights != nullptr, "weights is NULL");
  if (weights->precision != DDMI_PREC_FP32) {
    set_error("nerf mlp: precision %d has no kernel in this build (fp32 only)", weights->precision);
    return DDMI_ERR_UNSUPPORTED;
  }
  int rc = check_weights(weights, kNerfGemmFloats * sizeof(float), kNerfVecFloats);
  if (rc) return rc;
  return launch_nerf_mlp_fp32(x, n, x_stride, sigma_only ? 1 : 0, negative_slope, (const float*)weights->gemm,
                              weights->vec, out, (cudaStream_t)stream);
}

static int nerf_render_impl(const ddmi_plane_t planes[3], int32_t batch, int32_t channels, int32_t plane_layout,
                            const float* rays, int64_t n_rays, int32_t ray_stride, const float* t_vals, int z_stride,
                            int32_t n_samples, float plane_extent, float negative_slope, int32_t white_bkgd,
                            const ddmi_weights_t* weights, float* rgb_map, float* raw, void* stream);

DDMI_API int ddmi_nerf_render(const ddmi_plane_t planes[3], int32_t batch, int32_t channels, int32_t plane_layout,
                     const float* rays, int64_t n_rays, int32_t ray_stride, const float* t_vals, int32_t n_samples,
                     float plane_extent, float negative_slope, int32_t white_bkgd,
                     const ddmi_weights_t* weights, float* rgb_map, float* raw, void* stream) {
  return nerf_render_impl(planes, batch, channels, plane_layout, rays, n_rays, ray_stride, t_vals, 0, n_samples, plane_extent,
                          negative_slope, white_bkgd, weights, rgb_map, raw, stream);
}

DDMI_API int ddmi_nerf_render_z(const ddmi_plane_t planes[3], int32_t batch, int32_t channels, int32_t plane_layout,
                       const float* rays, int64_t n_rays, int32_t ray_stride, const float* z_vals, int32_t n_samples,
                       float plane_extent, float negative_slope, int32_t white_bkgd,
                       const ddmi_weights_t* weights, float* rgb_map, float* raw, void* stream) {
  return nerf_render_impl(planes, batch, channels, plane_layout, rays, n_rays, ray_stride, z_vals, n_samples, n_samples,
                          plane_extent, negative_slope, white_bkgd, weights, rgb_map, raw, stream);
}

static int nerf_render_impl(const ddmi_plane_t planes[3], int32_t batch, int32_t channels, int32_t plane_layout,
                            const float* rays, int64_t n_rays, int32_t ray_stride, const float* t_vals, int z_stride,
                            int32_t n_samples, float plane_extent, float negative_slope, int32_t white_bkgd,
                            const ddmi_weights_t* weights, float* rgb_map, float* raw, void* stream) {
  PlaneSet ps = {};
  int rc = check_planes(planes, 3, &ps);
  if (rc) return rc;
  DDMI_REQUIRE(batch >= 1, "batch must be >= 1 (got %d)", batch);
  DDMI_REQUIRE(n_rays >= 1 && n_samples >= 1, "empty ray set (%lld rays x %d samples)", (long long)n_rays, n_samples);
  DDMI_REQUIRE(rays && t_vals && rgb_map, "rays / t_vals (z_vals) / rgb_map is NULL");
  DDMI_REQUIRE(ray_stride >= 11, "ray rows need [o d near far viewdir] = 11 floats (stride %d)", ray_stride);
  DDMI_REQUIRE(plane_extent > 0.f, "plane_extent must be positive");
  if (channels != 32) {
    set_error("nerf render is built for 32-channel triplanes (got %d)", channels);
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(weights != nullptr, "weights is NULL");
  DDMI_REQUIRE(raw == nullptr || ((uintptr_t)raw & 15) == 0, "raw must be 16-byte aligned");
  if (weights->precision == DDMI_PREC_BF16X3 || weights->precision == DDMI_PREC_F16F8) {
    // tcgen05 kernel: channels-last planes, CTA pairs; compositing is fused when one tile is one ray
    if (plane_layout != DDMI_LAYOUT_NHWC || !(weights->reserved & 1)) {
      set_error("the tcgen05 NeRF kernel needs channels-last planes and pair-packed weights");
      return DDMI_ERR_UNSUPPORTED;
    }
    DDMI_REQUIRE(weights->gemm && weights->vec, "weights->gemm / weights->vec is NULL");
    DDMI_REQUIRE(((uintptr_t)weights->gemm & 127) == 0, "weights->gemm must be 128-byte aligned");
    const int fuse = n_samples == 128;
    DDMI_REQUIRE(fuse || raw != nullptr, "n_samples != 128: compositing runs as a second kernel over `raw`; pass the buffer");
    rc = launch_nerf_umma_entry(ps, batch, channels, rays, n_rays, ray_stride, t_vals, z_stride, n_samples, plane_extent,
                                negative_slope, white_bkgd, weights->gemm, weights->gemm_bytes, weights->program_host,
                                weights->program_words, weights->program, weights->vec, weights->vec_floats, rgb_map, raw,
                                fuse, weights->precision == DDMI_PREC_F16F8, (cudaStream_t)stream);
    if (rc || fuse) return rc;
    return launch_nerf_composite(raw, rays, ray_stride, t_vals, z_stride, n_samples, n_rays, batch, white_bkgd, rgb_map,
                                 (cudaStream_t)stream);
  }
  if (weights->precision != DDMI_PREC_FP32) {
    set_error("nerf render: unknown precision %d", weights->precision);
    return DDMI_ERR_UNSUPPORTED;
  }
  if (plane_layout != DDMI_LAYOUT_NCHW) {
    set_error("the fp32 NeRF kernel reads NCHW planes only");
    return DDMI_ERR_UNSUPPORTED;
  }
  rc = check_weights(weights, kNerfGemmFloats * sizeof(float), kNerfVecFloats);
  if (rc) return rc;
  DDMI_REQUIRE(raw != nullptr, "the fp32 render kernel composites from `raw`; pass a (batch,n_rays,n_samples,4) buffer");
  return launch_nerf_render_fp32(ps, batch, channels, rays, n_rays, ray_stride, t_vals, z_stride, n_samples, plane_extent,
                                 negative_slope, white_bkgd, (const float*)weights->gemm, weights->vec, rgb_map,
                                 raw, (cudaStream_t)stream);
}

DDMI_API int ddmi_sample_pdf(const float* bins, const float* weights, const float* u, int64_t n_rays, int32_t n_bins,
                             int32_t n_samples, float* out, void* stream) {
  DDMI_REQUIRE(bins && weights && u && out, "bins / weights / u / out is NULL");
  DDMI_REQUIRE(n_rays >= 1 && n_samples >= 1, "empty ray set (%lld rays x %d samples)", (long long)n_rays, n_samples);
  DDMI_REQUIRE(n_bins >= 2 && n_bins <= 1024, "n_bins must be in [2, 1024] (got %d)", n_bins);
  DDMI_REQUIRE(n_rays <= 4LL * 2147483647LL, "too many rays for one launch");
  return launch_sample_pdf(bins, weights, u, n_rays, n_bins, n_samples, out, (cudaStream_t)stream);
}

DDMI_API int ddmi_plane_head(const float* h, int32_t batch, int32_t in_channels, int32_t height, int32_t width, const float* weight,
                             const float* bias, int32_t out_channels, int32_t out_layout, float* out, void* stream) {
  DDMI_REQUIRE(h && weight && bias && out, "h / weight / bias / out is NULL");
  DDMI_REQUIRE(batch >= 1 && in_channels >= 1 && height >= 1 && width >= 1, "empty feature map");
  DDMI_REQUIRE(out_channels == 64 || out_channels == 32, "plane heads emit 64 (or 32: srn-cars) channels, got %d", out_channels);
  DDMI_REQUIRE(out_layout == DDMI_LAYOUT_NCHW || out_layout == DDMI_LAYOUT_NHWC, "out_layout must be DDMI_LAYOUT_NCHW / _NHWC");
  DDMI_REQUIRE(batch <= 65535, "batch %d too large for one launch", batch);
  return launch_plane_conv(h, batch, in_channels, height, width, 1, nullptr, nullptr, 1, 0.f, weight, bias, out_channels, 0,
                           out_layout == DDMI_LAYOUT_NHWC, nullptr, out, (cudaStream_t)stream);
}

DDMI_API int ddmi_plane_tail(const float* h, int32_t batch, int32_t in_channels, int32_t height, int32_t width, const float* gn_weight,
                             const float* gn_bias, int32_t groups, float eps, const float* weight, const float* bias,
                             int32_t out_channels, int32_t tanh_out, int32_t out_layout, float* stats, float* out, void* stream) {
  DDMI_REQUIRE(h && gn_weight && gn_bias && weight && bias && stats && out, "a tensor pointer is NULL");
  DDMI_REQUIRE(batch >= 1 && in_channels >= 1 && height >= 1 && width >= 1, "empty feature map");
  DDMI_REQUIRE(groups >= 1 && in_channels % groups == 0, "in_channels %d is not a multiple of groups %d", in_channels, groups);
  DDMI_REQUIRE(out_channels == 64 || out_channels == 32, "the tail emits 64 (or 32: srn-cars) channels, got %d", out_channels);
  DDMI_REQUIRE(out_layout == DDMI_LAYOUT_NCHW || out_layout == DDMI_LAYOUT_NHWC, "out_layout must be DDMI_LAYOUT_NCHW / _NHWC");
  DDMI_REQUIRE(batch <= 65535, "batch %d too large for one launch", batch);
  return launch_plane_conv(h, batch, in_channels, height, width, 3, gn_weight, gn_bias, groups, eps, weight, bias, out_channels,
                           tanh_out, out_layout == DDMI_LAYOUT_NHWC, stats, out, (cudaStream_t)stream);
}

DDMI_API int ddmi_mcubes_workspace_bytes(int32_t nx, int32_t ny, int32_t nz, int32_t pad, uint64_t* bytes) {
  DDMI_REQUIRE(bytes, "bytes is NULL");
  unsigned long long b = 0;
  const int rc = mcubes_workspace_bytes(nx, ny, nz, pad, &b);
  *bytes = b;
  return rc;
}

DDMI_API int ddmi_mcubes_count(const float* grid, int32_t nx, int32_t ny, int32_t nz, int32_t pad, double pad_value, double isovalue,
                               void* workspace, uint64_t workspace_bytes, uint64_t* totals_dev, void* stream) {
  DDMI_REQUIRE(grid && workspace && totals_dev, "grid / workspace / totals_dev is NULL");
  return launch_mcubes_count(grid, nx, ny, nz, pad, pad_value, isovalue, workspace, workspace_bytes,
                             (unsigned long long*)totals_dev, (cudaStream_t)stream);
}

DDMI_API int ddmi_mcubes_emit(const float* grid, int32_t nx, int32_t ny, int32_t nz, int32_t pad, double pad_value, double isovalue,
                              const void* workspace, const double* affine, double* vertices, int64_t* triangles, void* stream) {
  DDMI_REQUIRE(grid && workspace, "grid / workspace is NULL");
  DDMI_REQUIRE(vertices && triangles, "vertices / triangles is NULL (an empty mesh needs no emit call)");
  return launch_mcubes_emit(grid, nx, ny, nz, pad, pad_value, isovalue, workspace, affine, vertices, (long long*)triangles,
                            (cudaStream_t)stream);
}

DDMI_API int ddmi_selftest_tma(const float* plane, int32_t batch, int32_t channels, int32_t height, int32_t width, int32_t x,
                               int32_t y, int32_t c, int32_t variant, void* map_dev, float* out, void* stream) {
  DDMI_REQUIRE(plane && out && (variant == 0 || map_dev), "plane / out / map_dev is NULL");
  return launch_tma_selftest(plane, batch, channels, height, width, x, y, c, variant, map_dev, out, (cudaStream_t)stream);
}

DDMI_API int ddmi_selftest_umma(const float* a, const float* b, float* d, int32_t N, int32_t K, void* stream) {
  DDMI_REQUIRE(a && b && d, "a / b / d is NULL");
  DDMI_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "N must be a multiple of 16 in [16,256] (got %d)", N);
  DDMI_REQUIRE(K >= 16 && K <= 256 && K % 16 == 0, "K must be a multiple of 16 in [16,256] (got %d)", K);
  return launch_selftest_umma(a, b, d, N, K, (cudaStream_t)stream);
}

DDMI_API int ddmi_selftest_umma2(const float* a, const float* b, float* d, int32_t N, int32_t K, void* stream) {
  DDMI_REQUIRE(a && b && d, "a / b / d is NULL");
  DDMI_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "N must be a multiple of 16 in [16,256] (got %d)", N);
  DDMI_REQUIRE(K >= 16 && K <= 256 && K % 16 == 0, "K must be a multiple of 16 in [16,256] (got %d)", K);
  return launch_selftest_umma2(a, b, d, N, K, (cudaStream_t)stream);
}

DDMI_API int ddmi_selftest_f16f8(const float* a, const float* b, float* d, int32_t N, int32_t K, void* stream) {
  DDMI_REQUIRE(a && b && d, "a / b / d is NULL");
  DDMI_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "N must be a multiple of 16 in [16,256] (got %d)", N);
  DDMI_REQUIRE(K >= 32 && K <= 256 && K % 32 == 0, "K must be a multiple of 32 in [32,256] (got %d)", K);
  return launch_selftest_f16f8(a, b, d, N, K, (cudaStream_t)stream);
}

DDMI_API int ddmi_debug_profile(uint64_t out[8], int32_t reset) {
  DDMI_REQUIRE(out != nullptr, "out is NULL");
  return debug_profile((unsigned long long*)out, reset);
}

DDMI_API int ddmi_debug_trace(uint64_t* out, int32_t capacity, int32_t* count, int32_t reset) {
  DDMI_REQUIRE(out != nullptr && count != nullptr && capacity >= 1, "out / count is NULL or capacity < 1");
  return debug_trace((unsigned long long*)out, capacity, count, reset);
}

DDMI_API int ddmi_debug_set(int32_t flags) { return debug_set(flags); }

DDMI_API int ddmi_debug_gatherbench(int32_t variant, int32_t unroll, const float* table, uint32_t ntexel, int32_t iters,
                                    int32_t smem_kb, int32_t ctas, uint64_t* out_dev, float* sink_dev, void* stream) {
  DDMI_REQUIRE(table && out_dev && sink_dev && ntexel >= 1 && iters >= 1 && ctas >= 1, "bad gatherbench arguments");
  DDMI_REQUIRE(smem_kb >= 1 && smem_kb <= 227, "smem_kb must be 1..227");
  return launch_gatherbench(variant, unroll, table, ntexel, iters, smem_kb, ctas, (unsigned long long*)out_dev, sink_dev,
                            (cudaStream_t)stream);
}

DDMI_API int ddmi_debug_ringbench(const void* src, uint64_t span_bytes, int32_t slot_bytes, int32_t nslots, int32_t iters,
                                  int32_t ctas, uint64_t* out_dev, void* stream) {
  DDMI_REQUIRE(src && out_dev && ((uintptr_t)src & 127) == 0, "src / out_dev is NULL or src not 128-byte aligned");
  DDMI_REQUIRE(slot_bytes >= 1024 && slot_bytes % 1024 == 0 && nslots >= 1 && nslots <= 64 &&
                   (long long)slot_bytes * nslots <= 196 * 1024,
               "ring of %d x %d bytes does not fit", nslots, slot_bytes);
  DDMI_REQUIRE(span_bytes >= (uint64_t)slot_bytes && iters >= 1 && ctas >= 1, "empty span / iters / ctas");
  return launch_ringbench(src, span_bytes, slot_bytes, nslots, iters, ctas, (unsigned long long*)out_dev, (cudaStream_t)stream);
}

DDMI_API int ddmi_debug_microbench(int32_t mode, int32_t iters, const float* seed, uint64_t* out_dev, float* sink_dev, void* stream) {
  DDMI_REQUIRE(((mode >= 0 && mode <= 8) || (mode >= 100 && mode < 132) || (mode >= 200 && mode < 232)) && iters >= 1,
               "mode must be 0..8 (epilogue blocks) or 100.. / 200.. + variant (tcgen05.mma rate) and iters >= 1");
  DDMI_REQUIRE(seed && out_dev && sink_dev, "seed (1024 floats) / out_dev (2 x u64) / sink_dev (256 floats) is NULL");
  return launch_microbench(mode, iters, seed, (unsigned long long*)out_dev, sink_dev, (cudaStream_t)stream);
}

}  // extern "C"
