// Occupancy decoder (MLP3D.forward, models/d2c_vae/mlp.py:82-111) on the tcgen05 engine.
//
// Per 128-point tile: triplane gather (normalize_coordinate + 3 bilinear lookups summed,
// utils/general_utils.py:71-94,115-131) -> ResnetBlockFC x4 (blocks.py:673-716) -> logit.
// A ResnetBlockFC needs its input twice: raw for the shortcut, relu'd for fc_0.  The running
// activation lives once in shared memory, so each block with a shortcut runs in two operand
// phases: the epilogue keeps h (fp32) in registers, publishes RAW h (shortcut GEMM -> acc2),
// and after that GEMM committed rewrites the same buffer with relu(h) (fc_0 GEMM -> acc1);
// fc_1 then accumulates ONTO the shortcut in acc2, so x_s + dx never leaves TMEM.
// The 64-wide plane features are kept in both forms (Xa raw, Xb relu); in R2 / R3 they are published on their own barrier
// (A4) and their MMAs (shortcut part + the start of fc_0's accumulator) run right behind the shortcut's commit, i.e. while the
// epilogue threads rewrite H -- the tensor core is not idle across that hand-over.
//
// What bounds this kernel (round 2, profiles/r02_occupancy_timeline_grid.txt, r02_gatherbench_scattered_texels.txt): the
// GATHERS.  A point needs 9 planes x 4 taps x 256 B of channels-last texels (9 KB); a B200 SM sustains ~32 B/clk on scattered
// 256-byte texels through LDG whatever the load form, unroll depth or L1 size, so one scale costs the 256 epilogue threads
// ~12-13 K cycles, three times per tile, against 39 K cycles of MMA work -- and the epilogue threads that gather are the
// ones the next stage waits for.  Tried and measured slower: one X buffer + a 64 KB weight ring (the ring is not the limit
// here), dedicated gather warps staging texels with cp.async (scattered LDGSTS tops out near 13 B/clk/SM; two warps cannot
// hold enough LDG results in registers either).  Second half of round 2: the PE operands' MMAs commit on their own completion
// barrier (D1) right behind the shortcut, so the next gather runs under fc_0's MMAs over h instead of in front of the next
// stage; lattice queries (below) move a quarter of the bytes.  With that the kernel sits at ~70 % of its shared-memory-operand
// tensor bound (profiles/r02b_occupancy_timeline_lattice.txt).
//
// LATTICE QUERIES (LAT = 1; round 2).  On a query lattice {xs[i]} x {ys[j]} x {zs[k]} (the mesh generator's and BASELINE
// configs[3]'s dense grid) the 'xy' sample depends on (i, j) only, 'yz' on (j, k), 'xz' on (i, k): `occ_table_kernel` samples
// the nx ny + ny nz + nx nz distinct vectors per scale once (fp32, the direct gather's arithmetic) and a point's feature is
// three 256-byte table reads + the same two additions instead of twelve scattered texels: a quarter of the bytes, contiguous
// along z, and 1/43 of the tap arithmetic at 128^3.  Bit-identical to querying the expanded point list.
//
// net_p (the Linear(3, 256) on the query point) rides on R1.fc_1's accumulator: the point is published as columns 64..95 of H.
// vec layout (floats): b0_1[64] b1_1'[256] Wp[3][256] (unused by this kernel) b0_2[256] b1_2[256] b0_3[256] b1_3[256]
//                      b0_4[256] (b1_3+b1_4)[256] w_out[256] b_out[1]          (b1_1' = b1_1 + net_p.bias)
#pragma once
#include "umma_engine.cuh"

namespace ddmi {
namespace ummak {

using OccL = Layout<16, 32768>;   // X region: [Xa hi 8 | Xb hi 8] [Xa lo 8 | Xb lo 8] K groups; 4 x 8 KB ring slots
constexpr int OCC_KG_XAH = 64, OCC_KG_XBH = 72, OCC_KG_XAL = 80, OCC_KG_XBL = 88;
constexpr int OCC_OFF_PART = OccL::OFF_BAR + BAR_BYTES;          // [2][128] fp32 partial logits
constexpr int OCC_SMEM = OCC_OFF_PART + 1024;
constexpr int OV_B01 = 0, OV_B11 = 64, OV_WP = 320, OV_B02 = 1088, OV_B12 = 1344, OV_B03 = 1600, OV_B13 = 1856,
              OV_B04 = 2112, OV_B14 = 2368, OV_WOUT = 2624, OV_BOUT = 2880, OV_TOTAL = 2881;

__device__ __forceinline__ float2 relu_pair(float2 t) { return make_float2(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f)); }

// this thread's 128 values of a 256-wide accumulator: quarter q -> columns [64q + 32 sub, +32)
__device__ __forceinline__ void drain128(uint32_t tmem_lane, int acc_col, int sub, float2 (&v)[4][16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) tmem_ld32(tmem_lane + acc_col + q * 64 + sub * 32, v[q]);
  tmem_ld_wait();
}
// y <- accumulator values (de-scaled) + vector
template <int SCHEME, int NP>
__device__ __forceinline__ void add_vec(float2 (&y)[NP], const float* __restrict__ p) {
  float2 b[NP];
  load_vec<NP>(p, b);
#pragma unroll
  for (int i = 0; i < NP; ++i) y[i] = acc_plus<SCHEME>(y[i], b[i]);
}
// Generic epilogue stage: prefetch quarter 0's bias, park on the MMA barrier (`wait`), drain this thread's values of the
// accumulator at `acc_col`, then per quarter q < nq: v = f(v + bias) -> H, signal(sig0 + q); the bias of quarter q + 1 is
// loaded while quarter q is converted (there is no L1 beside ~224 KB of shared memory: an unprefetched bias costs an
// exposed L2 round trip per stage).  v is left holding the stored values.
template <int SCHEME, class Wait, class Act, class Signal>
__device__ __forceinline__ void biased_stage(uint32_t tmem_lane, int acc_col, int sub, int row, uint32_t h_hi, uint32_t h_lo,
                                             const float* __restrict__ bias, int nq, float2 (&v)[4][16], Wait wait, Act act,
                                             Signal signal, int sig0) {
  float2 b[16];
  load_vec<16>(bias + sub * 32, b);
  wait();
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (q < nq) tmem_ld32(tmem_lane + acc_col + q * 64 + sub * 32, v[q]);
  tmem_ld_wait();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < nq) {
      float2 bn[16];
      if (q + 1 < nq) load_vec<16>(bias + (q + 1) * 64 + sub * 32, bn);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[q][i] = act(acc_plus<SCHEME>(v[q][i], b[i]));
      {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float2 y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = v[q][c * 8 + i];
          store16<SCHEME>(h_hi, h_lo, row, q * 64 + sub * 32 + c * 16, y);
        }
      }
      signal(sig0 + q);
      if (q + 1 < nq) {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = bn[i];
      }
    }
  }
}

// write one quarter (32 columns of this thread) of the 256-wide operand, optionally relu'd
template <bool RELU, int SCHEME>
__device__ __forceinline__ void put_quarter(uint32_t h_hi, uint32_t h_lo, int row, int q, int sub, const float2 (&vq)[16]) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float2 y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = RELU ? relu_pair(vq[c * 8 + i]) : vq[c * 8 + i];
    store16<SCHEME>(h_hi, h_lo, row, q * 64 + sub * 32 + c * 16, y);
  }
}

// Block-output stage: h = acc (de-scaled) + bias [+ extra], published quarter by quarter (raw or relu'd) on barriers
// sig0..sig0+3; v keeps h (fp32) for a later phase.  Like biased_stage, quarter 0's bias is loaded before the thread parks on
// the MMA barrier and quarter q + 1's while quarter q is converted (an unprefetched vector is an exposed L2 round trip).
template <int SCHEME, bool RELU, class Wait, class Extra, class Signal>
__device__ __forceinline__ void output_stage(uint32_t tmem_lane, int acc_col, int sub, int row, uint32_t h_hi, uint32_t h_lo,
                                             const float* __restrict__ bias, float2 (&v)[4][16], Wait wait, Extra extra,
                                             Signal signal, int sig0) {
  float2 b[16];
  load_vec<16>(bias + sub * 32, b);
  wait();
  drain128(tmem_lane, acc_col, sub, v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float2 bn[16];
    if (q < 3) load_vec<16>(bias + (q + 1) * 64 + sub * 32, bn);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[q][i] = acc_plus<SCHEME>(v[q][i], b[i]);
    extra(q, v[q]);
    put_quarter<RELU, SCHEME>(h_hi, h_lo, row, q, sub, v[q]);
    signal(sig0 + q);
    if (q < 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) b[i] = bn[i];
    }
  }
}

// lattice queries: axes = [xs (nx) | ys (ny) | zs (nz)] (device), point index = (i * ny + j) * nz + k; table = per (item, scale)
// [xy: nx ny | yz: ny nz | xz: nx nz] records of 64 fp32
struct OccLattice {
  const float* table;
  const float* axes;
  int nx, ny, nz;
};
// one thread = 8 channels of one record
template <int NHWC>
__global__ void __launch_bounds__(256)
occ_table_kernel(PlaneSet ps, OccLattice lat, float divisor, float upper, int batch, float* __restrict__ table) {
  constexpr int C = 64;
  const long long nxy = (long long)lat.nx * lat.ny, nyz = (long long)lat.ny * lat.nz, nxz = (long long)lat.nx * lat.nz;
  const long long ntot = nxy + nyz + nxz, total = ntot * 8 * 3 * batch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // channels-last: the 8 threads of a record read one texel's 256 contiguous bytes; NCHW: consecutive threads walk entries
    const int kg = NHWC ? (int)(i & 7) : (int)((i / ntot) & 7);
    const long long e = NHWC ? (i >> 3) % ntot : i % ntot;
    const int bs = NHWC ? (int)((i >> 3) / ntot) : (int)(i / (ntot * 8));
    const int s = bs % 3, b = bs / 3;
    int pi;
    float ga, gb;   // (gx, gy) of make_tap: the direct gather's txy(g0, g1), tyz(g1, g2), txz(g0, g2)
    if (e < nxy) {
      pi = s;
      ga = occ_normalize(__ldg(lat.axes + e / lat.ny), divisor, upper);
      gb = occ_normalize(__ldg(lat.axes + lat.nx + e % lat.ny), divisor, upper);
    } else if (e < nxy + nyz) {
      pi = 3 + s;
      ga = occ_normalize(__ldg(lat.axes + lat.nx + (e - nxy) / lat.nz), divisor, upper);
      gb = occ_normalize(__ldg(lat.axes + lat.nx + lat.ny + (e - nxy) % lat.nz), divisor, upper);
    } else {
      pi = 6 + s;
      ga = occ_normalize(__ldg(lat.axes + (e - nxy - nyz) / lat.nz), divisor, upper);
      gb = occ_normalize(__ldg(lat.axes + lat.nx + lat.ny + (e - nxy - nyz) % lat.nz), divisor, upper);
    }
    const Tap tp = make_tap<true>(ga, gb, ps.h[pi], ps.w[pi]);
    const size_t hw = (size_t)ps.h[pi] * ps.w[pi];
    float y[8];
    if (NHWC) tap_sample8_nhwc(ps.data[pi] + (size_t)b * hw * C, tp, C, kg * 8, y);
    else tap_sample_n<8>(ps.data[pi] + ((size_t)b * C + kg * 8) * hw, hw, tp, y);
    float4* rec = reinterpret_cast<float4*>(table + ((size_t)bs * ntot + e) * C + kg * 8);
    rec[0] = make_float4(y[0], y[1], y[2], y[3]);
    rec[1] = make_float4(y[4], y[5], y[6], y[7]);
  }
}

template <int PAIR, int NHWC, int SCHEME, int LAT = 0>   // NHWC: planes are channels-last (batch, H, W, C) -- vectorised scattered gathers
__global__ void __launch_bounds__(NTHREADS, 1)
occupancy_umma_kernel(PlaneSet ps, const float* __restrict__ pts, long long n, long long batch_stride,
                      int tiles_per_item, long long total_tiles, float divisor, float upper,
                      const uint8_t* __restrict__ wstream, const __grid_constant__ ProgramParam prog,
                      const float* __restrict__ vec, float* __restrict__ logits, OccLattice lat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t h_hi = sbase, h_lo = sbase + H_KG * KG_BYTES;
  const uint32_t xa_hi = sbase + OCC_KG_XAH * KG_BYTES, xa_lo = sbase + OCC_KG_XAL * KG_BYTES;
  const uint32_t xb_hi = sbase + OCC_KG_XBH * KG_BYTES, xb_lo = sbase + OCC_KG_XBL * KG_BYTES;
  const uint32_t ring = sbase + OccL::OFF_RING, bar = sbase + OccL::OFF_BAR;
  float* part = reinterpret_cast<float*>(smem + OCC_OFF_PART);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int C = 64;
  const uint32_t tmem = engine_begin<PAIR, SCHEME>(smem, OccL::OFF_BAR);

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
    uint32_t ph_d1 = 0;
    auto wait_d1 = [&]() {   // completion barrier D1: the PE operands of R2 / R3 are consumed (Xa / Xb free)
      trace(tr, 0x01, trn, 0);
      mbar_wait(bar + BAR_MMADONE + 8, ph_d1);
      ph_d1 ^= 1;
      tc_fence_after();
      trace(tr, 0x04, trn, 0);
    };
    // this thread's query point of a tile (rows past the end replay the last point)
    auto point_of = [&](long long tile, float (&p)[3]) {
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      const int b = (int)(tile / tiles_per_item);
      long long gi = (tile % tiles_per_item) * TILE + row;
      if (gi > n - 1) gi = n - 1;
      if (LAT) {
        const unsigned g = (unsigned)gi, k = g % (unsigned)lat.nz, ij = g / (unsigned)lat.nz;
        p[0] = __ldg(lat.axes + ij / (unsigned)lat.ny);
        p[1] = __ldg(lat.axes + lat.nx + ij % (unsigned)lat.ny);
        p[2] = __ldg(lat.axes + lat.nx + lat.ny + k);
        return b;
      }
      const float* pp = pts + (size_t)b * batch_stride + gi * 3;
      p[0] = __ldg(pp); p[1] = __ldg(pp + 1); p[2] = __ldg(pp + 2);
      return b;
    };
    // triplane 'add' gather of scale s: raw -> Xa, relu -> Xb.
    // NHWC: 8 threads cooperate on one point -- thread (tid & 7) owns K group kg = 8 consecutive channels
    //   (two float4 per tap), so one warp instruction touches 4 points x 2 cache lines; 4 passes of 32 points.
    // NCHW: one thread per (point, 32-channel half), scalar loads (layout the VAE decoder emits).
    // part: -1 = the whole tile, 0 / 1 = its first / second half (rows in the vectorised forms, channels in the NCHW one)
    auto gather = [&](long long tile, int s, int part = -1) {
      const int p0 = part == 1 ? 2 : 0, p1 = part == 0 ? 2 : 4;
      trace(tr, 0x20, trn, 0);
      if (tile > total_tiles - 1) tile = total_tiles - 1;
      const int b = (int)(tile / tiles_per_item);
      const long long r0 = (tile % tiles_per_item) * TILE;
      const size_t hw0 = (size_t)ps.h[s] * ps.w[s], hw1 = (size_t)ps.h[3 + s] * ps.w[3 + s],
                   hw2 = (size_t)ps.h[6 + s] * ps.w[6 + s];
      if (LAT) {
        // 8 threads per point (thread kg: 8 channels = two float4 of each of the point's three records), 4 passes of 32
        // points; every load of the tile is issued before the first is consumed
        const int kg = tid & 7;
        const unsigned nxy = (unsigned)lat.nx * lat.ny, nyz = (unsigned)lat.ny * lat.nz, nxz = (unsigned)lat.nx * lat.nz;
        const float4* tb = reinterpret_cast<const float4*>(lat.table + (size_t)(b * 3 + s) * ((size_t)nxy + nyz + nxz) * C) + kg * 2;
        float4 u[4][6];
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          if (pass < p0 || pass >= p1) continue;
          long long gi = r0 + pass * 32 + (tid >> 3);
          if (gi > n - 1) gi = n - 1;
          const unsigned g = (unsigned)gi, k = g % (unsigned)lat.nz, ij = g / (unsigned)lat.nz;
          const unsigned i = ij / (unsigned)lat.ny, j = ij % (unsigned)lat.ny;
          const float4* pxy = tb + (size_t)ij * 16;
          const float4* pyz = tb + ((size_t)nxy + j * (unsigned)lat.nz + k) * 16;
          const float4* pxz = tb + ((size_t)nxy + nyz + i * (unsigned)lat.nz + k) * 16;
          u[pass][0] = __ldg(pxy); u[pass][1] = __ldg(pxy + 1);
          u[pass][2] = __ldg(pyz); u[pass][3] = __ldg(pyz + 1);
          u[pass][4] = __ldg(pxz); u[pass][5] = __ldg(pxz + 1);
        }
#pragma unroll
        for (int pass = 0; pass < 4; ++pass) {
          if (pass < p0 || pass >= p1) continue;
          const float a[8] = {u[pass][0].x, u[pass][0].y, u[pass][0].z, u[pass][0].w, u[pass][1].x, u[pass][1].y, u[pass][1].z, u[pass][1].w};
          const float c[8] = {u[pass][2].x, u[pass][2].y, u[pass][2].z, u[pass][2].w, u[pass][3].x, u[pass][3].y, u[pass][3].z, u[pass][3].w};
          const float d[8] = {u[pass][4].x, u[pass][4].y, u[pass][4].z, u[pass][4].w, u[pass][5].x, u[pass][5].y, u[pass][5].z, u[pass][5].w};
          float y[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) y[q] = __fadd_rn(__fadd_rn(a[q], c[q]), d[q]);
          store8_raw_relu<SCHEME>(xa_hi, xa_lo, xb_hi, xb_lo, pass * 32 + (tid >> 3), kg, y);
        }
      } else if (NHWC) {
        const int kg = tid & 7;
        const float* b0 = ps.data[s] + (size_t)b * hw0 * C;
        const float* b1 = ps.data[3 + s] + (size_t)b * hw1 * C;
        const float* b2 = ps.data[6 + s] + (size_t)b * hw2 * C;
#pragma unroll 1
        for (int pass = p0; pass < p1; ++pass) {
          const int prow = pass * 32 + (tid >> 3);
          long long gi = r0 + prow;
          if (gi > n - 1) gi = n - 1;
          const float* pp = pts + (size_t)b * batch_stride + gi * 3;
          const float g0 = occ_normalize(__ldg(pp), divisor, upper), g1 = occ_normalize(__ldg(pp + 1), divisor, upper),
                      g2 = occ_normalize(__ldg(pp + 2), divisor, upper);
          const Tap txy = make_tap<true>(g0, g1, ps.h[s], ps.w[s]);
          const Tap tyz = make_tap<true>(g1, g2, ps.h[3 + s], ps.w[3 + s]);
          const Tap txz = make_tap<true>(g0, g2, ps.h[6 + s], ps.w[6 + s]);
          float y[8], u[8], w[8];
          tap_sample8_nhwc(b0, txy, C, kg * 8, y);
          tap_sample8_nhwc(b1, tyz, C, kg * 8, u);
          tap_sample8_nhwc(b2, txz, C, kg * 8, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __fadd_rn(__fadd_rn(y[i], u[i]), w[i]);
          store8_raw_relu<SCHEME>(xa_hi, xa_lo, xb_hi, xb_lo, prow, kg, y);
        }
      } else {
        float p[3];
        point_of(tile, p);
        const float g0 = occ_normalize(p[0], divisor, upper), g1 = occ_normalize(p[1], divisor, upper),
                    g2 = occ_normalize(p[2], divisor, upper);
        const Tap txy = make_tap<true>(g0, g1, ps.h[s], ps.w[s]);
        const Tap tyz = make_tap<true>(g1, g2, ps.h[3 + s], ps.w[3 + s]);
        const Tap txz = make_tap<true>(g0, g2, ps.h[6 + s], ps.w[6 + s]);
        const float* b0 = ps.data[s] + ((size_t)b * C + ghalf * 32) * hw0;
        const float* b1 = ps.data[3 + s] + ((size_t)b * C + ghalf * 32) * hw1;
        const float* b2 = ps.data[6 + s] + ((size_t)b * C + ghalf * 32) * hw2;
#pragma unroll 1
        for (int g = p0; g < p1; ++g) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = g * 8 + i;
            float v = tap_sample(b0 + c * hw0, txy);
            v = __fadd_rn(v, tap_sample(b1 + c * hw1, tyz));
            v = __fadd_rn(v, tap_sample(b2 + c * hw2, txz));
            y[i] = v;
          }
          store8_raw_relu<SCHEME>(xa_hi, xa_lo, xb_hi, xb_lo, row, ghalf * 4 + g, y);
        }
      }
      trace(tr, 0x21, trn, 0);
    };
    // wait for the fc_0 GEMM, then relu(acc1 + b) -> H, quarter by quarter
    auto stage_net = [&](const float* __restrict__ b0) {
      float2 v[4][16];
      biased_stage<SCHEME>(tmem_lane, 0, sub, row, h_hi, h_lo, b0, 4, v, wait_mma, [](float2 t) { return relu_pair(t); }, signal, 0);
    };

    if (ntiles > 0) {
      gather(tile_of(0), 0);
      signal_all();
    }
    for (long long it = 0; it < ntiles; ++it) {
      const long long tile = tile_of(it);
      tr = DDMI_PROFILE && blockIdx.x == 0 && tid == 0 && it == kTraceIter;
      // ---- R1.fc_0 (N = 64): net = relu(acc1[:, 0:64] + b0) -> H[:, 0:64]
      {
        float2 v[16], b[16];
        load_vec<16>(vec + OV_B01 + sub * 32, b);      // before parking on the barrier: the L2 round trip overlaps the GEMM
        wait_mma();
        tmem_ld32(tmem_lane + sub * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = relu_pair(acc_plus<SCHEME>(v[i], b[i]));
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float2 y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = v[c * 8 + i];
          store16<SCHEME>(h_hi, h_lo, row, sub * 32 + c * 16, y);
        }
        {   // the query point as columns 64..95 of H: net_p runs on the tensor core behind R1.fc_1
          float p[3];
          point_of(tile, p);
          float2 y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = make_float2(0.f, 0.f);
          if (sub == 0) { y[0] = make_float2(p[0], p[1]); y[1] = make_float2(p[2], 0.f); }
          store16<SCHEME>(h_hi, h_lo, row, 64 + sub * 16, y);
        }
        signal_all();
      }
      // R1's fc_0 AND shortcut ran in that group, so both feature buffers are free; R1.fc_1 is a single K = 64 run, too short
      // to hide a gather: half of scale 1 is gathered here (these threads would park on fc_1's commit for as long), the other
      // half behind R1's output stage, under R2's shortcut GEMM over h
      gather(tile, 1, 0);
      // ---- R1 output h1 = acc2 + b1' + net_p(p); R2, R3: two operand phases (raw, then relu)
#pragma unroll 1
      for (int blk = 1; blk < 3; ++blk) {
        float2 v[4][16];
        const float* b1 = vec + (blk == 1 ? OV_B11 : OV_B12);
        // phase 1: raw h (shortcut operand)
        output_stage<SCHEME, false>(tmem_lane, 256, sub, row, h_hi, h_lo, b1, v, wait_mma, [](int, float2 (&)[16]) {}, signal, 0);
        // the PE operands have their own barrier (A4): the program runs them between the two h phases
        if (blk == 1) { gather(tile, 1, 1); signal(4); }
        // phase 2: relu(h) (fc_0 operand) once the shortcut GEMM has consumed the raw copy
        wait_mma();
#pragma unroll
        for (int q = 0; q < 4; ++q) { put_quarter<true, SCHEME>(h_hi, h_lo, row, q, sub, v[q]); signal(q); }
        // the PE operands ran right behind the shortcut's commit (D1): Xa / Xb are free, what comes next is gathered under
        // fc_0's MMAs over h
        // -- in two halves, one under fc_0's MMAs over h and one under fc_1's: a scattered gather (12 K cycles at the SM's L2
        // read rate) does not fit one of those windows
        wait_d1();
        if (blk == 1) gather(tile, 2, 0);
        else if (it + 1 < ntiles) gather(tile_of(it + 1), 0, 0);
        stage_net(vec + (blk == 1 ? OV_B02 : OV_B03));
        if (blk == 1) { gather(tile, 2, 1); signal(4); }
        else if (it + 1 < ntiles) gather(tile_of(it + 1), 0, 1);
      }
      // ---- R3 output h3 = acc2 + b1_3; R4 has an identity shortcut: only relu(h3) is needed, acc2 keeps accumulating
      {
        float2 v[4][16];
        output_stage<SCHEME, true>(tmem_lane, 256, sub, row, h_hi, h_lo, vec + OV_B13, v, wait_mma,
                                   [](int, float2 (&)[16]) {}, signal, 0);
      }
      // ---- R4.fc_0 epilogue
      stage_net(vec + OV_B04);
      // ---- logits = w_out . (acc2 + b1_3 + b1_4) + b_out   (vectors prefetched like in the other stages)
      {
        float2 v[4][16], b[16], w[16];
        load_vec<16>(vec + OV_B14 + sub * 32, b);
        load_vec<16>(vec + OV_WOUT + sub * 32, w);
        wait_mma();
        drain128(tmem_lane, 256, sub, v);
        // acc2 is in registers and the next tile's PE operands are in place: the tensor core starts the next tile's R1 while
        // the head is evaluated
        if (it + 1 < ntiles) signal_all();
        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 bn[16], wn[16];
          if (q < 3) {
            load_vec<16>(vec + OV_B14 + (q + 1) * 64 + sub * 32, bn);
            load_vec<16>(vec + OV_WOUT + (q + 1) * 64 + sub * 32, wn);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) s2 = __ffma2_rn(acc_plus<SCHEME>(v[q][i], b[i]), w[i], s2);
          if (q < 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { b[i] = bn[i]; w[i] = wn[i]; }
          }
        }
        part[sub * 128 + row] = s2.x + s2.y;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 256 E threads only
        if (sub == 0 && tile < total_tiles) {
          const int item = (int)(tile / tiles_per_item);
          const long long gi = (tile % tiles_per_item) * TILE + row;
          if (gi < n) logits[(size_t)item * n + gi] = part[row] + part[128 + row] + __ldg(vec + OV_BOUT);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // part[] may be rewritten by the next tile
      }
    }
  } else {
    engine_service_warps<PAIR, OccL::RING_BYTES, SCHEME, 0>(prog.op, wstream, sbase, ring, bar, tmem, ntiles, rank);
  }
  engine_end<PAIR>(tmem);
}

}  // namespace ummak

// bytes of the lattice tables of one launch: batch x 3 scales x (nx ny + ny nz + nx nz) records of 64 fp32
inline size_t occupancy_lattice_bytes(int batch, int nx, int ny, int nz) {
  return (size_t)batch * 3 * ((size_t)nx * ny + (size_t)ny * nz + (size_t)nx * nz) * 64 * sizeof(float);
}

// lattice != nullptr: the query is the lattice axes[0..nx) x axes[nx..nx+ny) x axes[nx+ny..) (pts / batch_stride unused, n = nx ny nz)
// and `workspace` receives its feature tables
inline int launch_occupancy_umma(const PlaneSet& ps, int batch, int C, const float* pts, long long n, long long batch_stride,
                                 float divisor, float upper, const void* gemm, size_t gemm_bytes,
                                 const uint32_t* program_host, size_t program_words, const uint32_t* program_dev,
                                 const float* vec, size_t vec_floats, float* logits, int pair, int nhwc, int f16f8,
                                 cudaStream_t st, const ummak::OccLattice* lattice = nullptr, void* workspace = nullptr,
                                 size_t workspace_bytes = 0) {
  using namespace ummak;
  DDMI_REQUIRE(!f16f8 || pair, "the f16f8 occupancy kernel runs as CTA pairs only");
  if (C != 64) {
    set_error("tcgen05 occupancy kernel is built for 64-channel planes");
    return DDMI_ERR_UNSUPPORTED;
  }
  DDMI_REQUIRE(program_host && program_dev && program_words >= 2, "bf16x3 weights carry no MMA program");
  const long long need = program_stream_bytes(program_host, program_words);
  DDMI_REQUIRE(need > 0 && (size_t)need == gemm_bytes, "MMA program consumes %lld weight bytes but the stream has %zu",
               need, gemm_bytes);
  ProgramParam pp;
  DDMI_REQUIRE(make_program_param(program_host, program_words, &pp), "MMA program has %zu words, at most %d fit the kernel parameter",
               program_words, PROG_MAX);
  DDMI_REQUIRE(vec_floats == (size_t)OV_TOTAL, "packed vec blob is %zu floats, expected %d", vec_floats, OV_TOTAL);
  int dev = 0, sms = 0;
  DDMI_CUDA(cudaGetDevice(&dev));
  DDMI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long tpi = (n + TILE - 1) / TILE;
  const long long total = tpi * batch;
  if (tpi > 2147483647LL) {
    set_error("n_points %lld too large for one launch", n);
    return DDMI_ERR_UNSUPPORTED;
  }
  const uint8_t* ws = (const uint8_t*)gemm;
  const int tpi_i = (int)tpi;
  const long long work = (total + 1) / 2, npairs = work < sms / 2 ? work : sms / 2;
  const unsigned ctas = pair ? (unsigned)(2 * npairs) : (unsigned)(total < sms ? total : sms);
  OccLattice lat = {};
  if (lattice) {
    DDMI_REQUIRE(pair, "lattice queries run on the CTA-pair kernels");
    DDMI_REQUIRE(n == (long long)lattice->nx * lattice->ny * lattice->nz && n <= 2147483647LL, "lattice of %d x %d x %d points",
                 lattice->nx, lattice->ny, lattice->nz);
    const size_t need_ws = occupancy_lattice_bytes(batch, lattice->nx, lattice->ny, lattice->nz);
    DDMI_REQUIRE(workspace && workspace_bytes >= need_ws, "lattice workspace is %zu bytes, ddmi_occupancy_lattice_workspace_bytes says %zu",
                 workspace_bytes, need_ws);
    DDMI_REQUIRE(((uintptr_t)workspace & 15) == 0, "lattice workspace must be 16-byte aligned");
    lat = *lattice;
    lat.table = (const float*)workspace;
    const long long items = (long long)(need_ws / 32), blocks = (items + 255) / 256;
    const unsigned grid = (unsigned)(blocks < (long long)sms * 16 ? blocks : (long long)sms * 16);
    if (nhwc) occ_table_kernel<1><<<grid, 256, 0, st>>>(ps, lat, divisor, upper, batch, (float*)workspace);
    else occ_table_kernel<0><<<grid, 256, 0, st>>>(ps, lat, divisor, upper, batch, (float*)workspace);
    DDMI_CUDA(cudaGetLastError());
  }
#define DDMI_OCC_LAUNCH(P, L, S, T)                                                                                        \
  DDMI_CUDA(launch_engine(occupancy_umma_kernel<P, L, S, T>, P, ctas, OCC_SMEM, st, ps, pts, n, batch_stride, tpi_i, total, \
                          divisor, upper, ws, pp, vec, logits, lat))
  if (lattice && f16f8) { DDMI_OCC_LAUNCH(1, 0, 1, 1); }
  else if (lattice) { DDMI_OCC_LAUNCH(1, 0, 0, 1); }
  else if (f16f8 && nhwc) { DDMI_OCC_LAUNCH(1, 1, 1, 0); }
  else if (f16f8) { DDMI_OCC_LAUNCH(1, 0, 1, 0); }
  else if (pair && nhwc) { DDMI_OCC_LAUNCH(1, 1, 0, 0); }
  else if (pair) { DDMI_OCC_LAUNCH(1, 0, 0, 0); }
  else if (nhwc) { DDMI_OCC_LAUNCH(0, 1, 0, 0); }
  else { DDMI_OCC_LAUNCH(0, 0, 0, 0); }
#undef DDMI_OCC_LAUNCH
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
