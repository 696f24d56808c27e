// Video decoder (MLPVideo.forward, models/d2c_vae/mlp.py:128-157; triplane 'concat' sampling,
// utils/general_utils.py:134-145) on the tcgen05 engine.
//
// Per 128-voxel tile (voxel order (t, h, w), w fastest): the per-scale feature is the CONCAT
// [xy(h,w) | yt(t,h) | xt(t,w)] (192 wide) -- never materialised.  It cannot sit in shared memory
// in raw + relu form next to the 256-wide activation, so it streams through the 64-wide Xa (raw) /
// Xb (relu) buffers as three PIECES; a concat is a K split, so each piece just adds its K = 64 runs
// to the shortcut (acc2) and fc_0 (acc1) accumulators.  Per ResnetBlockFC with input [h | X]:
//   E: publish RAW h (A0..A3)            | MMA: shortcut over h            -> COMMIT D0
//   E: gather piece 0 meanwhile
//   E: wait D0, publish relu(h) (A4..A7) | MMA: fc_0 over h quarter 0, piece 0 (both accumulators) -> COMMIT D1
//                                        |      fc_0 over h quarters 1..3
//   E: wait D1, gather piece 1, A0       | MMA: piece 1 -> COMMIT D1
//   E: wait D1, gather piece 2, A1       | MMA: piece 2 -> COMMIT D0
//   E: wait D0, net = relu(acc1 + b0) (A4..A7) | MMA: fc_1 ONTO acc2 -> COMMIT D0
// The grids are taken literally (the 'yt' / 'xt' planes are read with transposed axes).
//
// vec layout (floats): b0_1[192] b1_1[256] b0_2[256] b1_2[256] b0_3[256] b1_3[256] b0_4[256]
//                      (b1_3 + b1_4)[256] w_out[3][256] b_out[3]
#pragma once
#include "decode_umma_occ.cuh"

namespace ddmi {
namespace ummak {

using VidL = OccL;   // same carve-up: H | [Xa hi, Xb hi] | [Xa lo, Xb lo] | 4 x 8 KB ring | barriers
constexpr int VID_SMEM = VidL::OFF_BAR + BAR_BYTES;
// [3][2][128] fp32 partial outputs live at the start of H: at the output stage the last GEMM that reads H has
// committed and nothing writes H again before the named barrier that ends the stage.
constexpr int VV_B01 = 0, VV_B11 = 192, VV_B02 = 448, VV_B12 = 704, VV_B03 = 960, VV_B13 = 1216, VV_B04 = 1472,
              VV_B14 = 1728, VV_WOUT = 1984, VV_BOUT = 2752, VV_TOTAL = 2755;

template <int PAIR, int SCHEME>
__global__ void __launch_bounds__(NTHREADS, 1)
video_umma_kernel(PlaneSet ps, const float* __restrict__ cxy, const float* __restrict__ cyt,
                  const float* __restrict__ cxt, int T, int Hh, int Ww, int tiles_per_item, long long total_tiles,
                  const uint8_t* __restrict__ wstream, const __grid_constant__ ProgramParam prog,
                  const float* __restrict__ vec, void* __restrict__ out, int store) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase, h_lo = sbase + H_KG * KG_BYTES;
  const uint32_t xa_hi = sbase + OCC_KG_XAH * KG_BYTES, xa_lo = sbase + OCC_KG_XAL * KG_BYTES;
  const uint32_t xb_hi = sbase + OCC_KG_XBH * KG_BYTES, xb_lo = sbase + OCC_KG_XBL * KG_BYTES;
  const uint32_t ring = sbase + VidL::OFF_RING, bar = sbase + VidL::OFF_BAR;
  float* part = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int C = 64;
  const long long n = (long long)T * Hh * Ww;
  const uint32_t tmem = engine_begin<PAIR, SCHEME>(smem, VidL::OFF_BAR);

  const long long nwork = PAIR ? (total_tiles + 1) / 2 : total_tiles;
  const long long wfirst = PAIR ? blockIdx.x / 2 : blockIdx.x, wstride = PAIR ? gridDim.x / 2 : gridDim.x;
  const long long ntiles = wfirst < nwork ? (nwork - wfirst + wstride - 1) / wstride : 0;
  auto tile_of = [&](long long i) { const long long u = wfirst + i * wstride; return PAIR ? 2 * u + rank : u; };

  if (warp < 8) {
    reg_inc<216>();
    const int row = tid & 127;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int sub = warp >> 2;
    const int ghalf = tid >> 7;
    const uint32_t a_bar = PAIR ? mapa_rank(bar + BAR_A0, 0) : bar + BAR_A0;
    uint32_t ph_done = 0;   // bit i = parity of completion barrier i

    auto signal = [&](int i) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(a_bar + 8 * i);
    };
    auto wait_done = [&](int i) {
      mbar_wait(bar + BAR_MMADONE + 8 * i, (ph_done >> i) & 1);
      ph_done ^= 1u << i;
      tc_fence_after();
    };
    // piece j (0 'xy', 1 'yt', 2 'xt') of scale s: raw -> Xa, relu -> Xb; 32 of the 64 channels per thread
    auto gather = [&](long long tile, int s, int j) {
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      const int b = (int)(tile / tiles_per_item);
      long long gi = (tile % tiles_per_item) * TILE + row;
      if (gi > n - 1) gi = n - 1;
      const int w = (int)(gi % Ww), h = (int)((gi / Ww) % Hh), t = (int)(gi / ((long long)Ww * Hh));
      float g0, g1;   // channel 0 -> last plane axis, channel 1 -> second-to-last (grid_sample convention)
      if (j == 0) { g0 = __ldg(cxy + (size_t)h * Ww + w); g1 = __ldg(cxy + (size_t)Hh * Ww + (size_t)h * Ww + w); }
      else if (j == 1) { g0 = __ldg(cyt + (size_t)t * Hh + h); g1 = __ldg(cyt + (size_t)T * Hh + (size_t)t * Hh + h); }
      else { g0 = __ldg(cxt + (size_t)t * Ww + w); g1 = __ldg(cxt + (size_t)T * Ww + (size_t)t * Ww + w); }
      const int pi = j * 3 + s;
      const Tap tp = make_tap<true>(g0, g1, ps.h[pi], ps.w[pi]);
      const size_t hw = (size_t)ps.h[pi] * ps.w[pi];
      const float* base = ps.data[pi] + ((size_t)b * C + ghalf * 32) * hw;
#pragma unroll 1
      for (int g2 = 0; g2 < 2; ++g2) {          // 16 channels per round trip (64 loads in flight per thread)
        float y16[16];
        tap_sample_n<16>(base + (size_t)(g2 * 16) * hw, hw, tp, y16);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float y[8], yr[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            y[i] = y16[g * 8 + i];
            yr[i] = fmaxf(y[i], 0.f);
          }
          store8<SCHEME>(xa_hi, xa_lo, row, ghalf * 4 + g2 * 2 + g, y);
          store8<SCHEME>(xb_hi, xb_lo, row, ghalf * 4 + g2 * 2 + g, yr);
        }
      }
    };
    // net = relu(acc1 + b0) -> H quarters, published on A4..A7 (NQ = number of 64-column quarters: 3 for R1)
    auto stage_net = [&](const float* __restrict__ b0, int nq) {   // waits completion barrier D0 first
      float2 v[4][16];
      biased_stage<SCHEME>(tmem_lane, 0, sub, row, h_hi, h_lo, b0, nq, v, [&]() { wait_done(0); },
                   [](float2 t) { return relu_pair(t); }, signal, 4);
    };

    if (ntiles > 0) gather(tile_of(0), 0, 0);
    for (long long it = 0; it < ntiles; ++it) {
      const long long tile = tile_of(it);
      // ================= R1: x = X0 (three pieces), hidden 192 =================
      signal(0);                       // piece 0 (prefetched)
      wait_done(1);
      gather(tile, 0, 1); signal(1);
      wait_done(1);
      gather(tile, 0, 2); signal(2);
      stage_net(vec + VV_B01, 3);      // (waits D0) fc_1 (K = 192) accumulates onto the shortcut in acc2
      // ================= R2, R3: x = [h | X_s] =================
#pragma unroll 1
      for (int blk = 1; blk < 3; ++blk) {
        float2 v[4][16];
        output_stage<SCHEME, false>(tmem_lane, 256, sub, row, h_hi, h_lo, vec + (blk == 1 ? VV_B11 : VV_B12), v,
                                    [&]() { wait_done(0); }, [](int, float2 (&)[16]) {}, signal, 0);               // raw h
        gather(tile, blk, 0);                                                                              // overlaps the shortcut GEMM
        wait_done(0);
#pragma unroll
        for (int q = 0; q < 4; ++q) { put_quarter<true, SCHEME>(h_hi, h_lo, row, q, sub, v[q]); signal(4 + q); }   // relu(h) (+ piece 0)
        wait_done(1);
        gather(tile, blk, 1); signal(0);
        wait_done(1);
        gather(tile, blk, 2); signal(1);
        stage_net(vec + (blk == 1 ? VV_B02 : VV_B03), 4);
        // Xa / Xb are free (piece 2 committed): prefetch the next tile's first piece behind R3
        if (blk == 2 && it + 1 < ntiles) gather(tile_of(it + 1), 0, 0);
      }
      // ================= R4: identity shortcut, acc2 keeps accumulating =================
      {
        float2 v[4][16];
        output_stage<SCHEME, true>(tmem_lane, 256, sub, row, h_hi, h_lo, vec + VV_B13, v, [&]() { wait_done(0); },
                                   [](int, float2 (&)[16]) {}, signal, 0);
      }
      stage_net(vec + VV_B04, 4);
      // ================= out = w_out . lrelu(acc2 + b1_3 + b1_4, 0.2) + b_out =================
      wait_done(0);
      {
        float2 v[4][16];
        drain128(tmem_lane, 256, sub, v);
        float2 a3[3] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          add_vec<SCHEME, 16>(v[q], vec + VV_B14 + q * 64 + sub * 32);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 u = __fmul2_rn(v[q][i], make_float2(0.2f, 0.2f));
            v[q][i] = make_float2(fmaxf(v[q][i].x, u.x), fmaxf(v[q][i].y, u.y));
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float2 w[16];
            load_vec<16>(vec + VV_WOUT + c * 256 + q * 64 + sub * 32, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) a3[c] = __ffma2_rn(v[q][i], w[i], a3[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) part[(c * 2 + sub) * 128 + row] = a3[c].x + a3[c].y;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (sub == 0 && tile < total_tiles) {
          const int b = (int)(tile / tiles_per_item);
          const long long gi = (tile % tiles_per_item) * TILE + row;
          if (gi < n) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              store_rgb(out, store, b, n, gi, c, part[(c * 2) * 128 + row] + part[(c * 2 + 1) * 128 + row] + __ldg(vec + VV_BOUT + c));
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  } else {
    engine_service_warps<PAIR, VidL::RING_BYTES, SCHEME, 0>(prog.op, wstream, sbase, ring, bar, tmem, ntiles, rank);
  }
  engine_end<PAIR>(tmem);
}

}  // namespace ummak

inline int launch_video_umma(const PlaneSet& ps, int batch, int C, const float* cxy, const float* cyt, const float* cxt,
                             int T, int H, int W, const void* gemm, size_t gemm_bytes, const uint32_t* program_host,
                             size_t program_words, const uint32_t* program_dev, const float* vec, size_t vec_floats,
                             void* out, int store, int pair, int f16f8, cudaStream_t st) {
  using namespace ummak;
  DDMI_REQUIRE(!f16f8 || pair, "the f16f8 video kernel runs as CTA pairs only");
  if (C != 64) {
    set_error("tcgen05 video kernel is built for 64-channel planes");
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(program_host && program_dev && program_words >= 2, "bf16x3 weights carry no MMA program");
  const long long need = program_stream_bytes(program_host, program_words);
  DDMI_REQUIRE(need > 0 && (size_t)need == gemm_bytes, "MMA program consumes %lld weight bytes but the stream has %zu",
               need, gemm_bytes);
  ProgramParam pp;
  DDMI_REQUIRE(make_program_param(program_host, program_words, &pp), "MMA program has %zu words, at most %d fit the kernel parameter",
               program_words, PROG_MAX);
  DDMI_REQUIRE(vec_floats == (size_t)VV_TOTAL, "packed vec blob is %zu floats, expected %d", vec_floats, VV_TOTAL);
  int dev = 0, sms = 0;
  DDMI_CUDA(cudaGetDevice(&dev));
  DDMI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long n = (long long)T * H * W;
  const long long tpi = (n + TILE - 1) / TILE;
  const long long total = tpi * batch;
  if (tpi > 2147483647LL) {
    set_error("query volume too large for one launch");
    return DDMI_ERR_UNSUPPORTED;
  }
  const uint8_t* ws = (const uint8_t*)gemm;
  const int tpi_i = (int)tpi;
  const long long work = (total + 1) / 2, npairs = work < sms / 2 ? work : sms / 2;
  if (f16f8) {
    DDMI_CUDA(launch_engine(video_umma_kernel<1, 1>, 1, (unsigned)(2 * npairs), VID_SMEM, st, ps, cxy, cyt, cxt, T, H, W, tpi_i,
                            total, ws, pp, vec, out, store));
  } else if (pair) {
    DDMI_CUDA(launch_engine(video_umma_kernel<1, 0>, 1, (unsigned)(2 * npairs), VID_SMEM, st, ps, cxy, cyt, cxt, T, H, W, tpi_i,
                            total, ws, pp, vec, out, store));
  } else {
    DDMI_CUDA(launch_engine(video_umma_kernel<0, 0>, 0, (unsigned)(total < sms ? total : sms), VID_SMEM, st, ps, cxy, cyt, cxt, T,
                            H, W, tpi_i, total, ws, pp, vec, out, store));
  }
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
