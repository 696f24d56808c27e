// Video decoder (MLPVideo.forward, models/d2c_vae/mlp.py:128-157; triplane 'concat' sampling,
// utils/general_utils.py:134-145) on the tcgen05 engine.
//
// Per 128-voxel tile (voxel order (t, h, w), w fastest): the per-scale feature is the CONCAT
// [xy(h,w) | yt(t,h) | xt(t,w)] (192 wide) -- never materialised.  It cannot sit in shared memory
// in raw + relu form next to the 256-wide activation, so it streams through the 64-wide Xa (raw) /
// Xb (relu) buffers as three PIECES; a concat is a K split, so each piece just adds its K = 64 runs
// to the shortcut (acc2) and fc_0 (acc1) accumulators.  R1 (input = X0 only) is the exception: the activation buffer H is
// idle until R1's hidden layer is published, so its pieces 1 and 2 are parked in H (raw at K groups 16 (j - 1).., relu 8 further)
// beside piece 0 in Xa / Xb and the three pieces run back to back; they are stored -- and signalled -- inside the PREVIOUS
// tile's output stage, right after the accumulator drain, so the tensor core works on R1 while the output head is evaluated.
// Per ResnetBlockFC with input [h | X] (R2, R3):
//   E: store piece 0 (Xa / Xb are free), then publish RAW h (A0..A3) | MMA: shortcut over h -> COMMIT D0
//   E: wait D0, publish relu(h) (A4..A7)        | MMA: piece 0 (fc_0 starts acc1 here, shortcut adds to acc2) -> COMMIT D1
//                                               |      fc_0 over relu(h) quarters 0, 1
//   E: wait D1, store piece 1, A0               | MMA: piece 1 -> COMMIT D1;  fc_0 over quarters 2, 3
//   E: wait D1, store piece 2, A1               | MMA: piece 2 -> COMMIT D0
//   E: wait D0, net = relu(acc1 + b0) (A4..A7)  | MMA: fc_1 ONTO acc2 -> COMMIT D0
// (every hand-shake round trip of the epilogue threads has MMAs of another operand to hide behind)
// The grids are taken literally (the 'yt' / 'xt' planes are read with transposed axes).
//
//
// FEATURE TABLES (TAB = 1, f16f8; round 2).  A video's query grids are separable by construction -- the 'xy' grid is indexed
// by (h, w), 'yt' by (t, h), 'xt' by (t, w) (utils/general_utils.py:38-52) -- so a T x H x W volume holds only
// H W + T H + T W distinct feature vectors per scale, not T H W (1/14 for 16 x 256 x 256).  `video_table_kernel` samples each
// of them once (the direct gather's arithmetic) and stores it in OPERAND format, one 256-byte record per vector:
//   [a16: 64 fp16 | r8: 64 e5m2 residuals | a8: 64 e5m2]   (exactly the bytes store8<1> writes for one row)
// The decoder's gather then is 8 x 16-byte loads + 16 shared-memory stores per thread and piece: no taps, no conversions (the
// relu copy is the raw record with negative channels masked out), and the loads are issued BEFORE the thread parks on the
// completion barrier that frees Xa / Xb.  Bit-identical to the direct gather (tests/test_parity_gpu.py).
//
// vec layout (floats): b0_1[192] b1_1[256] b0_2[256] b1_2[256] b0_3[256] b1_3[256] b0_4[256]
//                      (b1_3 + b1_4)[256] w_out[3][256] b_out[3]
#pragma once
#include "decode_umma_occ.cuh"

namespace ddmi {
namespace ummak {

using VidL = OccL;   // same carve-up: H | [Xa hi, Xb hi] | [Xa lo, Xb lo] | 4 x 8 KB ring | barriers
// [3][128] fp32 partial outputs of the output head (the row's second thread hands its half to the first) behind the barriers
constexpr int VID_OFF_PART = VidL::OFF_BAR + BAR_BYTES;
constexpr int VID_SMEM = VID_OFF_PART + 3 * 128 * 4;
static_assert(VID_SMEM <= 232448, "video kernel shared memory");
constexpr int VV_B01 = 0, VV_B11 = 192, VV_B02 = 448, VV_B12 = 704, VV_B03 = 960, VV_B13 = 1216, VV_B04 = 1472,
              VV_B14 = 1728, VV_WOUT = 1984, VV_BOUT = 2752, VV_TOTAL = 2755;

constexpr int VID_REC_BYTES = 256;   // one feature-table record
// one thread = 8 channels (K group kg) of one record; consecutive threads walk consecutive grid entries (coalesced NCHW taps)
__global__ void __launch_bounds__(256)
video_table_kernel(PlaneSet ps, const float* __restrict__ cxy, const float* __restrict__ cyt, const float* __restrict__ cxt,
                   int T, int Hh, int Ww, int batch, uint8_t* __restrict__ table) {
  constexpr int C = 64;
  const long long nxy = (long long)Hh * Ww, nyt = (long long)T * Hh, nxt = (long long)T * Ww, ntot = nxy + nyt + nxt;
  const long long total = ntot * 8 * 3 * batch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % ntot;
    const int kg = (int)((i / ntot) & 7);
    const int bs = (int)(i / (ntot * 8));          // b * 3 + s
    const int s = bs % 3, b = bs / 3;
    int j;
    float g0, g1;   // the grids are taken literally, like the direct gather
    if (e < nxy) { j = 0; g0 = __ldg(cxy + e); g1 = __ldg(cxy + nxy + e); }
    else if (e < nxy + nyt) { j = 1; g0 = __ldg(cyt + (e - nxy)); g1 = __ldg(cyt + nyt + (e - nxy)); }
    else { j = 2; g0 = __ldg(cxt + (e - nxy - nyt)); g1 = __ldg(cxt + nxt + (e - nxy - nyt)); }
    const int pi = j * 3 + s;
    const Tap tp = make_tap<true>(g0, g1, ps.h[pi], ps.w[pi]);
    const size_t hw = (size_t)ps.h[pi] * ps.w[pi];
    float y[8];
    tap_sample_n<8>(ps.data[pi] + ((size_t)b * C + kg * 8) * hw, hw, tp, y);
    uint4 a16;
    uint2 r8, a8;
    split8_f16f8(y, a16, r8, a8);
    uint8_t* rec = table + ((size_t)bs * ntot + e) * VID_REC_BYTES;
    *reinterpret_cast<uint4*>(rec + kg * 16) = a16;
    *reinterpret_cast<uint2*>(rec + 128 + kg * 8) = r8;
    *reinterpret_cast<uint2*>(rec + 192 + kg * 8) = a8;
  }
}
struct VidRec { uint4 a[4], r[2], e[2]; };   // this thread's 32 channels of one record
struct VidDst { uint32_t a_hi, a_lo, b_hi, b_lo; };   // where a piece goes: raw (a) and relu (b) copies, hi / lo K groups

template <int PAIR, int SCHEME, int TAB>
__global__ void __launch_bounds__(NTHREADS, 1)
video_umma_kernel(PlaneSet ps, const float* __restrict__ cxy, const float* __restrict__ cyt,
                  const float* __restrict__ cxt, int T, int Hh, int Ww, int tiles_per_item, long long total_tiles,
                  const uint8_t* __restrict__ wstream, const __grid_constant__ ProgramParam prog,
                  const float* __restrict__ vec, void* __restrict__ out, int store, const uint8_t* __restrict__ table) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase, h_lo = sbase + H_KG * KG_BYTES;
  const uint32_t xa_hi = sbase + OCC_KG_XAH * KG_BYTES, xa_lo = sbase + OCC_KG_XAL * KG_BYTES;
  const uint32_t xb_hi = sbase + OCC_KG_XBH * KG_BYTES, xb_lo = sbase + OCC_KG_XBL * KG_BYTES;
  const uint32_t ring = sbase + VidL::OFF_RING, bar = sbase + VidL::OFF_BAR;
  float* part = reinterpret_cast<float*>(smem + VID_OFF_PART);
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
    bool tr = false;        // profiling build: E thread 0 of CTA 0 traces tile iteration kTraceIter
    uint32_t trn = 0;

    auto signal = [&](int i) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(a_bar + 8 * i);
      trace(tr, 0x10 + i, trn, 0);
    };
    auto wait_done = [&](int i) {
      trace(tr, 0x01, trn, 0);
      mbar_wait(bar + BAR_MMADONE + 8 * i, (ph_done >> i) & 1);
      ph_done ^= 1u << i;
      tc_fence_after();
      trace(tr, 0x02, trn, 0);
    };
    // piece j (0 'xy', 1 'yt', 2 'xt') of scale s: raw -> Xa, relu -> Xb; 32 of the 64 channels per thread
    auto gather = [&](long long tile, int s, int j, const VidDst& d) {
      trace(tr, 0x20, trn, 0);
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
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = y16[g * 8 + i];
          store8_raw_relu<SCHEME>(d.a_hi, d.a_lo, d.b_hi, d.b_lo, row, ghalf * 4 + g2 * 2 + g, y);
        }
      }
      trace(tr, 0x21, trn, 0);
    };
    // table path: loads (any time) and stores (once Xa / Xb are free) of piece j of scale s
    auto tab_load = [&](long long tile, int s, int j, VidRec& x) {
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      const int b = (int)(tile / tiles_per_item);
      long long gi = (tile % tiles_per_item) * TILE + row;
      if (gi > n - 1) gi = n - 1;
      const long long nxy = (long long)Hh * Ww, nyt = (long long)T * Hh, nxt = (long long)T * Ww;
      const long long e = j == 0 ? gi % nxy : (j == 1 ? nxy + gi / Ww : nxy + nyt + (gi / nxy) * Ww + gi % Ww);
      const uint4* p = reinterpret_cast<const uint4*>(table + ((size_t)(b * 3 + s) * (nxy + nyt + nxt) + e) * VID_REC_BYTES);
#pragma unroll
      for (int i = 0; i < 4; ++i) x.a[i] = __ldg(p + 4 * ghalf + i);
#pragma unroll
      for (int i = 0; i < 2; ++i) { x.r[i] = __ldg(p + 8 + 2 * ghalf + i); x.e[i] = __ldg(p + 12 + 2 * ghalf + i); }
    };
    auto tab_store = [&](const VidRec& x, const VidDst& d) {
      trace(tr, 0x20, trn, 0);
      const uint32_t o16 = (uint32_t)(ghalf * 4 * KG_BYTES + row * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        st_shared_v4(d.a_hi + o16 + i * KG_BYTES, x.a[i]);
        st_shared_v4(d.b_hi + o16 + i * KG_BYTES,
                     make_uint4(relu_h2(x.a[i].x), relu_h2(x.a[i].y), relu_h2(x.a[i].z), relu_h2(x.a[i].w)));
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint4 m = make_uint4(~neg_mask8(x.e[i].x), ~neg_mask8(x.e[i].y), ~neg_mask8(x.e[i].z), ~neg_mask8(x.e[i].w));
        st_shared_v4(d.a_lo + o16 + i * KG_BYTES, x.r[i]);
        st_shared_v4(d.a_lo + o16 + (2 + i) * KG_BYTES, x.e[i]);
        st_shared_v4(d.b_lo + o16 + i * KG_BYTES, make_uint4(x.r[i].x & m.x, x.r[i].y & m.y, x.r[i].z & m.z, x.r[i].w & m.w));
        st_shared_v4(d.b_lo + o16 + (2 + i) * KG_BYTES, make_uint4(x.e[i].x & m.x, x.e[i].y & m.y, x.e[i].z & m.z, x.e[i].w & m.w));
      }
      trace(tr, 0x21, trn, 0);
    };
    // net = relu(acc1 + b0) -> H quarters, published on A4..A7 (NQ = number of 64-column quarters: 3 for R1)
    auto stage_net = [&](const float* __restrict__ b0, int nq) {   // waits completion barrier D0 first
      float2 v[4][16];
      biased_stage<SCHEME>(tmem_lane, 0, sub, row, h_hi, h_lo, b0, nq, v, [&]() { wait_done(0); },
                   [](float2 t) { return relu_pair(t); }, signal, 4);
    };

    // piece (s, j) of `tile` into `d`, which the completion barrier `done` (-1: none) frees: the table path issues its
    // loads before parking on the barrier
    const VidDst dX = {xa_hi, xa_lo, xb_hi, xb_lo};
    auto piece = [&](long long tile, int s, int j, int done, const VidDst& d) {
      if (TAB) {
        VidRec x;
        tab_load(tile, s, j, x);
        if (done >= 0) wait_done(done);
        tab_store(x, d);
      } else {
        if (done >= 0) wait_done(done);
        gather(tile, s, j, d);
      }
    };
    // R1's pieces 1 and 2 of `tile` -> H (idle between R4.fc_1's commit and R1's hidden layer), published on A1 / A2
    auto r1_dst = [&](int j) {
      const uint32_t o = (uint32_t)(16 * (j - 1) * KG_BYTES);
      return VidDst{h_hi + o, h_lo + o, h_hi + o + 8 * KG_BYTES, h_lo + o + 8 * KG_BYTES};
    };
    auto r1_pieces = [&](long long tile) {
#pragma unroll
      for (int j = 1; j < 3; ++j) {
        piece(tile, 0, j, -1, r1_dst(j));
        signal(j);
      }
    };
    if (ntiles > 0) {
      piece(tile_of(0), 0, 0, -1, dX);
      signal(0);
      r1_pieces(tile_of(0));
    }
    for (long long it = 0; it < ntiles; ++it) {
      const long long tile = tile_of(it);
      tr = DDMI_PROFILE && blockIdx.x == 0 && tid == 0 && it == kTraceIter;
      // ================= R1: x = X0 (three pieces, published by the previous tile's output stage), hidden 192 =================
      stage_net(vec + VV_B01, 3);      // (waits D0) fc_1 (K = 192) accumulates onto the shortcut in acc2
      // ================= R2, R3: x = [h | X_s] =================
#pragma unroll 1
      for (int blk = 1; blk < 3; ++blk) {
        float2 v[4][16];
        piece(tile, blk, 0, -1, dX);            // Xa / Xb are free (the previous stage_net waited the commit behind piece 2)
        output_stage<SCHEME, false>(tmem_lane, 256, sub, row, h_hi, h_lo, vec + (blk == 1 ? VV_B11 : VV_B12), v,
                                    [&]() { wait_done(0); }, [](int, float2 (&)[16]) {}, signal, 0);               // raw h (+ piece 0)
        wait_done(0);
#pragma unroll
        for (int q = 0; q < 4; ++q) { put_quarter<true, SCHEME>(h_hi, h_lo, row, q, sub, v[q]); signal(4 + q); }   // relu(h)
        piece(tile, blk, 1, 1, dX); signal(0);
        piece(tile, blk, 2, 1, dX); signal(1);
        stage_net(vec + (blk == 1 ? VV_B02 : VV_B03), 4);
        // Xa / Xb are free (piece 2 committed): prefetch the next tile's first piece behind R3
        if (blk == 2 && it + 1 < ntiles) piece(tile_of(it + 1), 0, 0, -1, dX);
      }
      // ================= R4: identity shortcut, acc2 keeps accumulating =================
      {
        float2 v[4][16];
        output_stage<SCHEME, true>(tmem_lane, 256, sub, row, h_hi, h_lo, vec + VV_B13, v, [&]() { wait_done(0); },
                                   [](int, float2 (&)[16]) {}, signal, 0);
      }
      stage_net(vec + VV_B04, 4);
      // ================= out = w_out . lrelu(acc2 + b1_3 + b1_4, 0.2) + b_out =================
      const bool more = it + 1 < ntiles;
      if (TAB && more) {
        // the next tile's R1 pieces 1 and 2: records loaded while this thread would park anyway, stored as soon as H is free
        // (R4.fc_1 committed) and published at once; piece 0 (already in Xa / Xb) is published after the drain below --
        // R1's first MMAs overwrite acc2
        VidRec x1, x2;
        tab_load(tile_of(it + 1), 0, 1, x1);
        tab_load(tile_of(it + 1), 0, 2, x2);
        wait_done(0);
        tab_store(x1, r1_dst(1)); signal(1);
        tab_store(x2, r1_dst(2)); signal(2);
      } else {
        wait_done(0);
      }
      {
        float2 v[4][16];
        drain128(tmem_lane, 256, sub, v);
        if (more) signal(0);   // acc2 is in registers: the tensor core works on the next tile's R1 while the head is evaluated
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
        if (sub == 1) {
#pragma unroll
          for (int c = 0; c < 3; ++c) part[c * 128 + row] = a3[c].x + a3[c].y;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (sub == 0 && tile < total_tiles) {
          const int b = (int)(tile / tiles_per_item);
          const long long gi = (tile % tiles_per_item) * TILE + row;
          if (gi < n) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
              store_rgb(out, store, b, n, gi, c, (a3[c].x + a3[c].y) + part[c * 128 + row] + __ldg(vec + VV_BOUT + c));
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (more && !TAB) r1_pieces(tile_of(it + 1));
      }
    }
  } else {
    engine_service_warps<PAIR, VidL::RING_BYTES, SCHEME, 0>(prog.op, wstream, sbase, ring, bar, tmem, ntiles, rank);
  }
  engine_end<PAIR>(tmem);
}

}  // namespace ummak

// bytes of the feature tables of one launch: batch x 3 scales x (H W + T H + T W) records
inline size_t video_table_bytes(int batch, int T, int H, int W) {
  return (size_t)batch * 3 * ((size_t)H * W + (size_t)T * H + (size_t)T * W) * ummak::VID_REC_BYTES;
}

inline int launch_video_umma(const PlaneSet& ps, int batch, int C, const float* cxy, const float* cyt, const float* cxt,
                             int T, int H, int W, const void* gemm, size_t gemm_bytes, const uint32_t* program_host,
                             size_t program_words, const uint32_t* program_dev, const float* vec, size_t vec_floats,
                             void* out, int store, int pair, int f16f8, void* workspace, size_t workspace_bytes,
                             cudaStream_t st) {
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
  const uint8_t* table = nullptr;
  if (f16f8 && workspace) {   // feature tables (see the top of this file); without a workspace every tile gathers directly
    const size_t need_ws = video_table_bytes(batch, T, H, W);
    DDMI_REQUIRE(workspace_bytes >= need_ws, "video workspace is %zu bytes, ddmi_video_workspace_bytes says %zu", workspace_bytes,
                 need_ws);
    DDMI_REQUIRE(((uintptr_t)workspace & 15) == 0, "video workspace must be 16-byte aligned");
    const long long work_items = (long long)(need_ws / VID_REC_BYTES) * 8;
    const long long blocks = (work_items + 255) / 256;
    video_table_kernel<<<(unsigned)(blocks < (long long)sms * 16 ? blocks : (long long)sms * 16), 256, 0, st>>>(
        ps, cxy, cyt, cxt, T, H, W, batch, (uint8_t*)workspace);
    DDMI_CUDA(cudaGetLastError());
    table = (const uint8_t*)workspace;
  }
  if (f16f8 && table) {
    DDMI_CUDA(launch_engine(video_umma_kernel<1, 1, 1>, 1, (unsigned)(2 * npairs), VID_SMEM, st, ps, cxy, cyt, cxt, T, H, W,
                            tpi_i, total, ws, pp, vec, out, store, table));
  } else if (f16f8) {
    DDMI_CUDA(launch_engine(video_umma_kernel<1, 1, 0>, 1, (unsigned)(2 * npairs), VID_SMEM, st, ps, cxy, cyt, cxt, T, H, W,
                            tpi_i, total, ws, pp, vec, out, store, table));
  } else if (pair) {
    DDMI_CUDA(launch_engine(video_umma_kernel<1, 0, 0>, 1, (unsigned)(2 * npairs), VID_SMEM, st, ps, cxy, cyt, cxt, T, H, W,
                            tpi_i, total, ws, pp, vec, out, store, table));
  } else {
    DDMI_CUDA(launch_engine(video_umma_kernel<0, 0, 0>, 0, (unsigned)(total < sms ? total : sms), VID_SMEM, st, ps, cxy, cyt,
                            cxt, T, H, W, tpi_i, total, ws, pp, vec, out, store, table));
  }
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
