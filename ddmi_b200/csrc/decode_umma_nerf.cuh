// NeRF ray decode (utils/nerf_helpers.py:296-530 render_rays + run_network + raw2outputs, with
// MLPNeRF.forward models/d2c_vae/mlp.py:241-281) on the tcgen05 engine.
//
// One tile = 128 consecutive (ray, sample) rows.  Per tile: sample generation, triplane gather
// (channels-last planes, align_corners = true), positional embeddings of the un-normalised point
// (L = 10) written straight into the A-operand layout, the 8-layer MLP (skips re-read the 160-wide
// [latent | embedding] operand X, which therefore stays resident), the view-direction layer
// (N = 128 block, its 27-wide embedding reuses X's space once the last skip layer is done), the
// sigma / rgb heads in the epilogue, and -- when n_samples == 128, i.e. one tile == one ray --
// the volume compositing with a warp-shuffle product scan, so only 3 floats per ray leave the SM.
//
// X lives in TENSOR memory (columns 256..415 behind the single 256-column accumulator; the MMAs that consume it
// use the A-from-TMEM form), written with tcgen05.st by the thread that owns the row.  With X in shared memory
// (80 KB next to H's 128 KB) only a 2-slot weight ring fitted and the MMA phases ran ~1.5x slower than the image
// kernel's, bound by ring refill latency; now the ring has 8 slots.
//
// vec layout (floats): b1..b6 [6][256] | b_final[256] | b_dir[128] | w_sigma[256] | b_sigma[1] pad[3]
//                      | w_rgb[3][128] | b_rgb[3]
#pragma once
#include "decode_umma_occ.cuh"

namespace ddmi {
namespace ummak {

using NrfL = Layout<0, 65536>;    // shared memory: H + 8 x 8 KB ring slots (CTA pairs only); X is in tensor memory
constexpr int NRF_KG_XH = 64, NRF_KG_XL = 84;   // X K groups (4 TMEM columns each): hi 64..83 -> columns 256..335, lo 84..103 -> 336..415
constexpr int NRF_OFF_SCRATCH = NrfL::OFF_BAR + BAR_BYTES;
constexpr int NRF_OFF_WRGB = NRF_OFF_SCRATCH + 5120;   // [w_rgb 3 x 128 | b_rgb 3 | pad] staged once per CTA: the rgb head runs
constexpr int NRF_SMEM = NRF_OFF_WRGB + 1552;          // between tiles, where six dependent L2 round trips were exposed
constexpr int NV_B = 0, NV_BF = 1536, NV_BD = 1792, NV_WS = 1920, NV_BS = 2176, NV_WRGB = 2180, NV_BRGB = 2564, NV_TOTAL = 2567;

__device__ __forceinline__ float2 lrelu_pair(float2 t, float slope) {
  return make_float2(fmaxf(t.x, 0.f) + slope * fminf(t.x, 0.f), fmaxf(t.y, 0.f) + slope * fminf(t.y, 0.f));
}

// K group j (8 columns of X) of this thread's row -> tensor memory.  SCHEME 1: fp16 group at XH + j, and 8 bytes each of
// the pair's r8 / a8 groups ([r8 r8 a8 a8] per 32 columns behind XL; 8 bytes = 2 TMEM columns)
template <int SCHEME>
__device__ __forceinline__ void x_store8(uint32_t tmem_lane, int j, const float (&y)[8]) {
  if (SCHEME) {
    uint4 a16;
    uint2 r8, a8;
    split8_f16f8(y, a16, r8, a8);
    tmem_st4(tmem_lane + (NRF_KG_XH + j) * 4, a16);
    const uint32_t c8 = (uint32_t)((NRF_KG_XL + (j >> 2) * 4) * 4 + (j & 3) * 2);
    tmem_st2(tmem_lane + c8, r8);
    tmem_st2(tmem_lane + c8 + 8, a8);
  } else {
    uint4 hi, lo;
    split8(y, hi, lo);
    tmem_st4(tmem_lane + (NRF_KG_XH + j) * 4, hi);
    tmem_st4(tmem_lane + (NRF_KG_XL + j) * 4, lo);
  }
}

// gamma(p), L = 10 (Embedder.embed, nerf_helpers.py:82-112): [p, sin(2^0 p), cos(2^0 p), ..., sin(2^9 p), cos(2^9 p)] = 63 values
// (+ one zero pad).  Elements [32 HALF, +32) with compile-time indexing; sin and cos of one argument share a sincosf.
// LEVELS = 10, N = 32: one half of gamma(pts); LEVELS = 4, N = 16: one half of gamma(viewdir) (27 values + 5 zero pads).
template <int HALF, int LEVELS, int N>
__device__ __forceinline__ void embed_half(const float (&p)[3], float (&y)[N]) {
  constexpr int E0 = N * HALF;
#pragma unroll
  for (int i = 0; i < N; ++i) y[i] = 0.f;                  // covers the pad elements
  if (HALF == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) y[a] = p[a];
  }
#pragma unroll
  for (int l = 0; l < LEVELS; ++l) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int es = 3 + 6 * l + a, ec = es + 3;            // element indices of sin / cos (compile-time after unrolling)
      const bool ns = es >= E0 && es < E0 + N, nc = ec >= E0 && ec < E0 + N;
      if (ns || nc) {
        float sv, cv;
        sincosf(__fmul_rn(p[a], (float)(1 << l)), &sv, &cv);
        if (ns) y[es - E0] = sv;
        if (nc) y[ec - E0] = cv;
      }
    }
  }
}

template <int NK>
struct XTaps {   // the tap loads of NK K groups of one plane + the bilinear weights
  float4 q[NK][8];
  float w[4];
};

template <int SCHEME>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
nerf_umma_kernel(PlaneSet ps, int C, const float* __restrict__ rays, int ray_stride, const float* __restrict__ t_vals,
                 int z_stride, int n_samples, float plane_extent, long long n /* rows per object */, int tiles_per_item,
                 long long total_tiles, float slope, int white_bkgd, int fuse,
                 const uint8_t* __restrict__ wstream, const __grid_constant__ ProgramParam prog,
                 const float* __restrict__ vec, float* __restrict__ rgb_map, float* __restrict__ raw) {
  constexpr int PAIR = 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase, h_lo = sbase + H_KG * KG_BYTES;
  const uint32_t ring = sbase + NrfL::OFF_RING, bar = sbase + NrfL::OFF_BAR;
  float* scratch = reinterpret_cast<float*>(smem + NRF_OFF_SCRATCH);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t tmem = engine_begin<PAIR, SCHEME>(smem, NrfL::OFF_BAR);
  float* wrgb = reinterpret_cast<float*>(smem + NRF_OFF_WRGB);
  for (int i = threadIdx.x; i < 387; i += NTHREADS) wrgb[i] = __ldg(vec + NV_WRGB + i);   // NV_BRGB = NV_WRGB + 384
  __syncthreads();

  const long long nwork = (total_tiles + 1) / 2;
  const long long wfirst = blockIdx.x / 2, wstride = gridDim.x / 2;
  const long long ntiles = wfirst < nwork ? (nwork - wfirst + wstride - 1) / wstride : 0;
  auto tile_of = [&](long long i) { return 2 * (wfirst + i * wstride) + rank; };

  if (warp < 8) {
    reg_inc<216>();
    const int row = tid & 127;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int sub = warp >> 2;
    const int ghalf = tid >> 7;
    const uint32_t a_bar = mapa_rank(bar + BAR_A0, 0);
    uint32_t ph_mma = 0;
    bool tr = false;        // profiling build: E thread 0 of CTA 0 traces tile iteration kTraceIter
    uint32_t trn = 0;

    auto signal = [&](int q) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(a_bar + 8 * q);
      trace(tr, 0x10 + q, trn, 0);
    };
    auto signal_all = [&]() {
#pragma unroll
      for (int q = 0; q < 4; ++q) signal(q);
    };
    auto wait_mma = [&]() {
      trace(tr, 0x01, trn, 0);
      mbar_wait(bar + BAR_MMADONE, ph_mma);
      ph_mma ^= 1;
      tc_fence_after();
      trace(tr, 0x02, trn, 0);
    };
    // row -> (object, ray, sample); rows past the end replay the last one
    struct RowInfo { int b; long long ray; int smp; long long gi; };
    auto row_of = [&](long long tile) {
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      RowInfo r;
      r.b = (int)(tile / tiles_per_item);
      r.gi = (tile % tiles_per_item) * TILE + row;
      if (r.gi > n - 1) r.gi = n - 1;
      r.ray = r.gi / n_samples;
      r.smp = (int)(r.gi % n_samples);
      return r;
    };
    auto zval = [&](const float* rr, long long ray, int smp) {   // nerf_helpers.py:356-380 (common.cuh::nerf_z)
      return nerf_z(t_vals, z_stride, ray, smp, __ldg(rr + 6), __ldg(rr + 7));
    };
    // X = [latent xy|yz|xz (3 x 32) | gamma(pts) 63 | 0] = 12 gathered + 8 embedding K groups per row, split evenly over the
    // row's two threads: thread 0 gathers 'xy' K groups 0, 1 and 'yz' 4..7, thread 1 'xy' 2, 3 and 'xz' 8..11 (one plane = one
    // tap set per batch); embedding groups [12 + 4 ghalf, +4).  The gathers are L2-latency bound (there is no L1 beside ~200 KB
    // of shared memory), so all tap loads of a batch are issued before any is consumed.
    // The NEXT tile's X is built in three windows where these threads would otherwise park: X is dead once the second skip
    // layer has committed and only its first 4 K groups are reused (view-direction embedding of the last GEMM), so
    //   x_point + x_embed: the sample point and gamma(pts) (16 sincosf per thread, ~5 K cycles) under layer 6's GEMM,
    //   x_planes<4>:       the 'yz' | 'xz' latents after the final layer's epilogue, under the last two GEMMs, together with
    //   x_taps<2>:         the LOADS of the 'xy' taps, which are blended + stored (x_finish<2>) once the last GEMM has
    //                      committed; the tile is handed over right after
    // (the whole build between tiles kept the tensor core idle for ~10 K cycles per tile: profiles/r02b_nerf_timeline_before.txt)
    auto x_point = [&](long long tile, float (&p)[3]) {   // -> object index
      const RowInfo ri = row_of(tile);
      const float* rr = rays + (size_t)ri.ray * ray_stride;
      const float z = zval(rr, ri.ray, ri.smp);
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(__ldg(rr + i), __fmul_rn(__ldg(rr + 3 + i), z));
      return ri.b;
    };
    auto x_embed = [&](const float (&p)[3]) {
      float e[32];
      if (ghalf == 0) embed_half<0, 10, 32>(p, e); else embed_half<1, 10, 32>(p, e);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = e[jj * 8 + i];
        x_store8<SCHEME>(tmem_lane, 12 + ghalf * 4 + jj, y);
      }
    };
    // tap loads of K groups j0 .. j0 + NK - 1 (all of plane a = j0 / 4: xy = (x, y), yz = (y, z), xz = (x, z); first coordinate
    // -> column) of object b at point p
    auto x_taps = [&](int b, const float (&p)[3], int j0, auto& t) {
      constexpr int NK = sizeof(t.q) / sizeof(t.q[0]);
      float g[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) g[i] = __fdiv_rn(p[i], plane_extent);
      const int a = j0 >> 2;
      const float ga = a == 1 ? g[1] : g[0], gb = a == 0 ? g[1] : g[2];
      const Tap tp = make_tap<true>(ga, gb, ps.h[a], ps.w[a]);
      const float* img = ps.data[a] + (size_t)b * ps.h[a] * ps.w[a] * C + (j0 & 3) * 8;
      const int o[4] = {tp.o00, tp.o01, tp.o10, tp.o11};
      t.w[0] = tp.w00; t.w[1] = tp.w01; t.w[2] = tp.w10; t.w[3] = tp.w11;
#pragma unroll
      for (int u = 0; u < NK; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4* pp = reinterpret_cast<const float4*>(img + (size_t)o[k] * C + u * 8);
          t.q[u][2 * k] = __ldg(pp);
          t.q[u][2 * k + 1] = __ldg(pp + 1);
        }
      }
    };
    auto x_finish = [&](int j0, const auto& t) {   // blend -> X K groups j0.. (the caller waits for the tensor-memory stores)
      constexpr int NK = sizeof(t.q) / sizeof(t.q[0]);
#pragma unroll
      for (int u = 0; u < NK; ++u) {
        float y[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // same FMA order as common.cuh::tap_sample8_nhwc
          const float4 a0 = t.q[u][h], b0 = t.q[u][2 + h], c0 = t.q[u][4 + h], d0 = t.q[u][6 + h];
          y[4 * h + 0] = fmaf(d0.x, t.w[3], fmaf(c0.x, t.w[2], fmaf(b0.x, t.w[1], a0.x * t.w[0])));
          y[4 * h + 1] = fmaf(d0.y, t.w[3], fmaf(c0.y, t.w[2], fmaf(b0.y, t.w[1], a0.y * t.w[0])));
          y[4 * h + 2] = fmaf(d0.z, t.w[3], fmaf(c0.z, t.w[2], fmaf(b0.z, t.w[1], a0.z * t.w[0])));
          y[4 * h + 3] = fmaf(d0.w, t.w[3], fmaf(c0.w, t.w[2], fmaf(b0.w, t.w[1], a0.w * t.w[0])));
        }
        x_store8<SCHEME>(tmem_lane, j0 + u, y);
      }
    };
    auto x_planes4 = [&](int b, const float (&p)[3]) {   // 'yz' (thread 0) / 'xz' (thread 1): 32 float4 loads in flight
      XTaps<4> t;
      x_taps(b, p, 4 + 4 * ghalf, t);
      x_finish(4 + 4 * ghalf, t);
    };
    // h = lrelu(acc1 + b, slope) -> H, quarter by quarter
    auto stage_act = [&](const float* __restrict__ b, bool act, float2 (&v)[4][16]) {   // waits for the GEMM first
      const float sl = act ? slope : 1.f;
      // nn.LeakyReLU(True) has negative_slope = 1.0: the reference's activation is the identity -- do not spend
      // 4 instructions per pair evaluating it
      if (sl == 1.f)
        biased_stage<SCHEME>(tmem_lane, 0, sub, row, h_hi, h_lo, b, 4, v, wait_mma, [](float2 t) { return t; }, signal, 0);
      else
        biased_stage<SCHEME>(tmem_lane, 0, sub, row, h_hi, h_lo, b, 4, v, wait_mma, [sl](float2 t) { return lrelu_pair(t, sl); }, signal, 0);
    };

    if (ntiles > 0) {
      float p0[3];
      const int b0 = x_point(tile_of(0), p0);
      XTaps<2> t0;
      x_taps(b0, p0, 2 * ghalf, t0);
      x_finish(2 * ghalf, t0);
      x_planes4(b0, p0);
      x_embed(p0);
      tmem_st_wait();
      signal_all();
    }
    float pn[3] = {0.f, 0.f, 0.f};   // the next tile's sample point / object (set under layer 6's GEMM)
    int bn = 0;
    for (long long it = 0; it < ntiles; ++it) {
      const long long tile = tile_of(it);
      tr = DDMI_PROFILE && blockIdx.x == 0 && tid == 0 && it == kTraceIter;
      const RowInfo ri = row_of(tile);
      const float* rr = rays + (size_t)ri.ray * ray_stride;
      float sigma = 0.f;
      // ---- xyz_encoding_1..6
#pragma unroll 1
      for (int l = 0; l < 6; ++l) {
        float2 v[4][16];
        stage_act(vec + NV_B + l * 256, true, v);
        // X is dead from here on (the second skip layer, xyz_encoding_5, has committed): the next tile's gamma(pts) -- 16
        // sincosf per thread, ~5 K cycles -- goes under layer 6's GEMM
        if (l == 4 && it + 1 < ntiles) {
          trace(tr, 0x20, trn, 0);
          bn = x_point(tile_of(it + 1), pn);
          x_embed(pn);
          tmem_st_wait();
          trace(tr, 0x21, trn, 0);
        }
        if (l == 5) {
          // X is dead (the last skip layer committed): view-direction embedding -> X's first 4 K groups,
          // sigma = w_sigma . h6 + b_sigma (this thread's 128 columns, summed across the two sub-warps below)
          {   // 2 of the 4 K groups per thread of the row
            const float vd[3] = {__ldg(rr + 8), __ldg(rr + 9), __ldg(rr + 10)};
            float e[16];
            if (ghalf == 0) embed_half<0, 4, 16>(vd, e); else embed_half<1, 4, 16>(vd, e);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              float y[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = e[jj * 8 + i];
              x_store8<SCHEME>(tmem_lane, 2 * ghalf + jj, y);
            }
            tmem_st_wait();
          }
          float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 w[16];
            load_vec<16>(vec + NV_WS + q * 64 + sub * 32, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) s2 = __ffma2_rn(v[q][i], w[i], s2);
          }
          scratch[sub * 128 + row] = s2.x + s2.y;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          sigma = scratch[row] + scratch[128 + row] + __ldg(vec + NV_BS);
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      // ---- xyz_encoding_final (no activation); the direction embedding written above rides on these signals
      {
        float2 v[4][16];
        stage_act(vec + NV_BF, false, v);
      }
      const bool more = it + 1 < ntiles;
      XTaps<2> xy;
      if (more) {
        trace(tr, 0x20, trn, 0);
        x_planes4(bn, pn);
        x_taps(bn, pn, 2 * ghalf, xy);
        tmem_st_wait();
        trace(tr, 0x21, trn, 0);
      }
      // ---- dir_encoding (N = 128) + rgb head
      float rgb[3];
      {
        float2 v[2][16], bd[2][16];
        load_vec<16>(vec + NV_BD + sub * 32, bd[0]);          // before parking on the barrier (L2 round trip under the GEMM)
        load_vec<16>(vec + NV_BD + 64 + sub * 32, bd[1]);
        wait_mma();
        tmem_ld32(tmem_lane + sub * 32, v[0]);
        tmem_ld32(tmem_lane + 64 + sub * 32, v[1]);
        tmem_ld_wait();
        // Every MMA of this tile has committed and its last accumulator is in registers: finish the next tile's X (the 'xy'
        // latent, where the direction embedding sat) and hand over NOW, so that the rgb head and the compositing below run
        // under the next tile's first GEMM.
        if (more) {
          x_finish(2 * ghalf, xy);
          tmem_st_wait();
          signal_all();
        }
        float2 acc3[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[q][i] = acc_plus<SCHEME>(v[q][i], bd[q][i]);
            v[q][i] = slope == 1.f ? v[q][i] : lrelu_pair(v[q][i], slope);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float2* w = reinterpret_cast<const float2*>(wrgb + c * 128 + q * 64 + sub * 32);   // warp-uniform: broadcast
#pragma unroll
            for (int i = 0; i < 16; ++i) acc3[c] = __ffma2_rn(v[q][i], w[i], acc3[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) scratch[(c * 2 + sub) * 128 + row] = acc3[c].x + acc3[c].y;
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float s = scratch[(c * 2) * 128 + row] + scratch[(c * 2 + 1) * 128 + row] + wrgb[384 + c];
          rgb[c] = 1.f / (1.f + expf(-s));
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const bool valid = tile < total_tiles && (tile % tiles_per_item) * TILE + row < n;
      if (sub == 0) {
        if (raw != nullptr && valid) {
          reinterpret_cast<float4*>(raw)[(size_t)ri.b * n + ri.gi] = make_float4(rgb[0], rgb[1], rgb[2], sigma);
        }
        if (fuse) {
          // raw2outputs (nerf_helpers.py:487-530): one tile == one ray, row == sample index
          const float z0 = zval(rr, ri.ray, ri.smp);
          float dist = ri.smp + 1 < n_samples ? __fsub_rn(zval(rr, ri.ray, ri.smp + 1), z0) : 1e10f;
          const float dx = __ldg(rr + 3), dy = __ldg(rr + 4), dz = __ldg(rr + 5);
          dist = __fmul_rn(dist, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))));
          const float alpha = __fsub_rn(1.f, expf(-__fmul_rn(softplus20(sigma), dist)));
          float t = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
          // inclusive product scan over the warp's 32 samples, then across the 4 warps
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const float u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t *= u;
          }
          float T = __shfl_up_sync(0xffffffffu, t, 1);
          if (lane == 0) T = 1.f;
          if (lane == 31) scratch[1024 + warp] = t;
          asm volatile("bar.sync 2, 128;" ::: "memory");
          for (int w = 0; w < warp; ++w) T *= scratch[1024 + w];
          const float wgt = alpha * T;
          float s0 = wgt * rgb[0], s1 = wgt * rgb[1], s2 = wgt * rgb[2], sa = wgt;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
          }
          if (lane == 0) {
            scratch[1040 + warp * 4 + 0] = s0; scratch[1040 + warp * 4 + 1] = s1;
            scratch[1040 + warp * 4 + 2] = s2; scratch[1040 + warp * 4 + 3] = sa;
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (row == 0 && tile < total_tiles) {
            float o3[4] = {0.f, 0.f, 0.f, 0.f};
            for (int w = 0; w < 4; ++w)
              for (int c = 0; c < 4; ++c) o3[c] += scratch[1040 + w * 4 + c];
            const float bg = white_bkgd ? 1.f - o3[3] : 0.f;
            float* dst = rgb_map + ((size_t)ri.b * (n / n_samples) + ri.ray) * 3;
            dst[0] = o3[0] + bg; dst[1] = o3[1] + bg; dst[2] = o3[2] + bg;
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
      }
      trace(tr, 0x05, trn, 0);   // tile done on the epilogue side
    }
  } else {
    engine_service_warps<PAIR, NrfL::RING_BYTES, SCHEME>(prog.op, wstream, sbase, ring, bar, tmem, ntiles, rank);
  }
  engine_end<PAIR>(tmem);
}

}  // namespace ummak

inline int launch_nerf_umma(const PlaneSet& ps, int batch, int C, const float* rays, long long n_rays, int ray_stride,
                            const float* t_vals, int z_stride, int n_samples, float plane_extent, float slope, int white_bkgd,
                            const void* gemm, size_t gemm_bytes, const uint32_t* program_host, size_t program_words,
                            const uint32_t* program_dev, const float* vec, size_t vec_floats, float* rgb_map, float* raw,
                            int fuse, int f16f8, cudaStream_t st) {
  using namespace ummak;
  DDMI_REQUIRE(program_host && program_dev && program_words >= 2, "bf16x3 weights carry no MMA program");
  const long long need = program_stream_bytes(program_host, program_words);
  DDMI_REQUIRE(need > 0 && (size_t)need == gemm_bytes, "MMA program consumes %lld weight bytes but the stream has %zu",
               need, gemm_bytes);
  ProgramParam pp;
  DDMI_REQUIRE(make_program_param(program_host, program_words, &pp), "MMA program has %zu words, at most %d fit the kernel parameter",
               program_words, PROG_MAX);
  DDMI_REQUIRE(vec_floats == (size_t)NV_TOTAL, "packed vec blob is %zu floats, expected %d", vec_floats, NV_TOTAL);
  int dev = 0, sms = 0;
  DDMI_CUDA(cudaGetDevice(&dev));
  DDMI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long n = n_rays * n_samples;
  const long long tpi = (n + TILE - 1) / TILE;
  const long long total = tpi * batch;
  if (tpi > 2147483647LL) {
    set_error("ray set too large for one launch");
    return DDMI_ERR_UNSUPPORTED;
  }
  const long long work = (total + 1) / 2, npairs = work < sms / 2 ? work : sms / 2;
  if (f16f8) {
    DDMI_CUDA(cudaFuncSetAttribute(nerf_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NRF_SMEM));
    nerf_umma_kernel<1><<<(unsigned)(2 * npairs), NTHREADS, NRF_SMEM, st>>>(
        ps, C, rays, ray_stride, t_vals, z_stride, n_samples, plane_extent, n, (int)tpi, total, slope, white_bkgd, fuse,
        (const uint8_t*)gemm, pp, vec, rgb_map, raw);
  } else {
    DDMI_CUDA(cudaFuncSetAttribute(nerf_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, NRF_SMEM));
    nerf_umma_kernel<0><<<(unsigned)(2 * npairs), NTHREADS, NRF_SMEM, st>>>(
        ps, C, rays, ray_stride, t_vals, z_stride, n_samples, plane_extent, n, (int)tpi, total, slope, white_bkgd, fuse,
        (const uint8_t*)gemm, pp, vec, rgb_map, raw);
  }
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
