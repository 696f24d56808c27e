// Shared tcgen05 engine of the bf16x3 decode kernels (see decode_umma.cu for the design notes):
// program interpreters (weight producer, MMA issuer, pair forwarder), shared-memory layout,
// barrier map and the epilogue building blocks.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace ddmi {
namespace ummak {

using namespace umma;

constexpr int TILE = 128;              // rows per tile == MMA M
constexpr int NEPI = 256;              // gather / epilogue threads
constexpr int NTHREADS = 384;          // 3 warpgroups: 2 x E (216 regs), 1 x {producer, MMA, 2 idle warps} (72 regs);
                                       // 216*256 + 72*128 == 168*384: setmaxnreg only hands out what was released
constexpr int KG_BYTES = TILE * 16;    // one 8-wide K group of an A operand: 128 rows x 16 B
constexpr int H_KG = 32;               // 256-wide running activation
// weight ring: RING bytes cut into slots of one K step of a 256-wide block: 16 KB, or 8 KB per CTA of a
// pair (this CTA's 128 of the 256 rows)

// program op encoding (ddmi_b200/packing.py::UmmaProgram)
constexpr uint32_t OP_UNIT = 0, OP_WAIT = 1, OP_COMMIT = 2, OP_END = 3;
// The op table travels as a KERNEL PARAMETER (constant bank): the interpreters' op fetches are then uniform constant loads
// (ULDC) straight into uniform registers, so the whole decode -> descriptor -> tcgen05.mma chain of the issuer stays on the
// uniform datapath.  Fetched with __ldg from global memory the ops landed in vector registers and every descriptor paid an
// R2UR move: ~350 cycles per 4-MMA iteration, i.e. the issuer -- not the tensor pipe -- paced the kernel.
constexpr int PROG_MAX = 256;
struct ProgramParam {
  uint32_t op[PROG_MAX];
};

// barriers, 8 B each, relative to the barrier block
constexpr int MAX_SLOTS = 16;        // ring slots a kernel may use (barrier block capacity)
constexpr int BAR_WFULL = 0, BAR_WEMPTY = 8 * MAX_SLOTS, BAR_PFULL = 16 * MAX_SLOTS /* leader: the peer's half of slot s landed */,
              BAR_MMADONE = 24 * MAX_SLOTS /* D0..D3: COMMIT targets */, BAR_A0 = BAR_MMADONE + 32 /* A0..A7: operand barriers (WAIT ops) */,
              TMEM_SLOT = BAR_A0 + 64;
constexpr int BAR_BYTES = TMEM_SLOT + 32;
// Protocol rule for every mbarrier here: it must never complete two phases before its waiter has consumed the
// first (a parity wait cannot tell 0 from 2 completed phases).  E threads therefore re-signal operand barrier i
// only after waiting a COMMIT that follows the previous WAIT i in program order, and the program re-commits
// done-barrier j only after a WAIT whose signal the E threads issue after consuming the previous COMMIT j.

template <int XKG, int RING = 65536>   // XKG: K groups of the feature operand region behind H
struct Layout {
  static constexpr int A_BYTES = (2 * H_KG + 2 * XKG) * KG_BYTES;   // [H hi | H lo | X hi | X lo]
  static constexpr int OFF_RING = A_BYTES;
  static constexpr int RING_BYTES = RING;
  static constexpr int OFF_BAR = OFF_RING + RING;
  static constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES;
  static constexpr int KG_HHI = 0, KG_HLO = H_KG, KG_XHI = 2 * H_KG, KG_XLO = 2 * H_KG + XKG;
};

constexpr float kInvSqrt2 = 0.70710678118654752440f;

// Diagnostics (ddmi_debug_profile / ddmi_debug_trace): compiled in only with -DDDMI_PROFILE=1, i.e. in the separate
// libddmi_b200_prof.so the dev tools load (`make prof`); the shipping library carries no clock reads or counters.
// Counters of CTA 0:
// [0] E thread 0: cycles parked waiting for MMA groups   [1] cycles in epilogue stages (excl. gathers)
// [2] cycles in gathers   [3] MMA thread: cycles waiting for operands   [4] cycles waiting for weights
// [5] MMA thread total   [6] tiles   [7] spare
// Trace: (event id << 48 | clock) records of CTA 0 during tile iteration kTraceIter (E thread 0 and the MMA lane).
#ifndef DDMI_PROFILE
#define DDMI_PROFILE 0
#endif
__device__ unsigned long long g_prof[8];
// what-if switches of the profiling build (ddmi_debug_set; results are garbage, timings are the point):
// bit 0: epilogue stages only signal (no drain / convert / publish)   bit 1: the issuer skips the tcgen05.mma instructions
// (hand-shakes and commits stay)   bit 2: no plane gathers
__constant__ int g_dbg;   // constant bank: a what-if test costs the issuer no global-memory round trip
__device__ __forceinline__ bool dbg(int bit) { return DDMI_PROFILE && (g_dbg & bit); }
constexpr int kTraceCap = 4096, kTraceIter = 5, kTraceRegion = 1024;   // regions: E thread 0, MMA lane, 2 x kernel-specific
__device__ unsigned long long g_trace[kTraceCap];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ long long prof_clock() { return DDMI_PROFILE ? clock64() : 0ll; }
__device__ __forceinline__ void prof_add(int i, long long v) {
  if (DDMI_PROFILE) atomicAdd(&g_prof[i], (unsigned long long)v);
}
// one record; `on` = this thread is a designated tracer and the CTA is in its traced tile iteration.  Fire-and-forget stores
// (no atomics: a returning atomic costs the tracer an L2 round trip per record): E thread 0 fills the lower half of the
// buffer, the MMA lane the upper half, each with its own running index `n`; unwritten slots stay 0.
__device__ __forceinline__ void trace(bool on, uint32_t id, uint32_t& n, uint32_t base) {
  if (DDMI_PROFILE && on) {
    if (n < (uint32_t)kTraceRegion) g_trace[base + n] = ((unsigned long long)id << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull);
    ++n;
  }
}

// ---------------------------------------------------------------------------
// engine: producer + MMA issuer (program interpreters)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t op_n(uint32_t op) {
  const uint32_t c = (op >> 2) & 3;
  return c == 0 ? 128u : (c == 2 ? 16u : (c == 3 ? 64u : 256u));
}

// SCHEME 1 (f16f8): the ring is managed in PAIRS of slots (one 32-wide K step pair = 4 MMAs): one empty barrier (the odd
// slot's), one full barrier (the even slot's), one expect_tx and two bulk copies per pair -- at 256 tensor cycles per slot
// a per-slot handshake (each a ~90-cycle barrier round trip) barely keeps the ring full.
template <int PAIR, int RING_BYTES, int SCHEME = 0>
__device__ __forceinline__ void producer_loop(const uint32_t* __restrict__ program, const uint8_t* __restrict__ wstream,
                                              uint32_t ring, uint32_t bar, long long ntiles, uint32_t rank) {
  constexpr uint32_t SLOT_BYTES = PAIR ? 8192 : 16384, NSLOT = RING_BYTES / SLOT_BYTES;
  uint32_t slot = 0, ph = 0;
  for (long long t = 0; t < ntiles; ++t) {
    const uint8_t* src = wstream;
    for (int pc = 0;; ++pc) {
      const uint32_t op = program[pc];
      const uint32_t kind = op & 3;
      if (kind == OP_END) break;
      if (kind != OP_UNIT) continue;
      // this CTA's share of one K step; half-width units (op bit 30, f16f8): of one 32-wide step PAIR of the 128-row block
      const bool half = SCHEME && ((op >> 30) & 1);
      const uint32_t bytes = half ? 8192u : op_n(op) * (PAIR ? 32 : 64);
      const int cnt = (int)((op >> 24) & 31) + 1;
      if (SCHEME) {
        for (int j = 0; j < cnt; j += half ? 4 : 2) {
          const uint32_t full = bar + BAR_WFULL + 8 * slot;
          mbar_wait(bar + BAR_WEMPTY + 8 * slot + 8, ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full, 2 * bytes);
            bulk_g2s(ring + slot * SLOT_BYTES, src + rank * bytes, bytes, full);
            bulk_g2s(ring + (slot + 1) * SLOT_BYTES, src + (2 + rank) * bytes, bytes, full);
          }
          src += bytes * 4;
          slot += 2;
          if (slot == NSLOT) { slot = 0; ph ^= 1; }
        }
      } else {
        for (int j = 0; j < cnt; ++j) {
          mbar_wait(bar + BAR_WEMPTY + 8 * slot, ph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bar + BAR_WFULL + 8 * slot, bytes);
            bulk_g2s(ring + slot * SLOT_BYTES, src + rank * bytes, bytes, bar + BAR_WFULL + 8 * slot);
          }
          src += bytes * (PAIR ? 2 : 1);
          if (++slot == NSLOT) { slot = 0; ph ^= 1; }
        }
      }
    }
  }
}

// Peer CTA of a pair: tell the leader when this CTA's half of each ring slot has landed.
template <int RING_BYTES, int SCHEME = 0>
__device__ __forceinline__ void forward_loop(const uint32_t* __restrict__ program, uint32_t bar, long long ntiles) {
  constexpr uint32_t NSLOT = RING_BYTES / 8192;
  uint32_t slot = 0, ph = 0;
  const uint32_t leader_pfull = mapa_rank(bar + BAR_PFULL, 0);
  for (long long t = 0; t < ntiles; ++t) {
    for (int pc = 0;; ++pc) {
      const uint32_t op = program[pc];
      const uint32_t kind = op & 3;
      if (kind == OP_END) break;
      if (kind != OP_UNIT) continue;
      const int cnt = (int)((op >> 24) & 31) + 1;
      const int per_unit = SCHEME ? (((op >> 30) & 1) ? 4 : 2) : 1;
      for (int j = 0; j < cnt; j += per_unit) {
        mbar_wait(bar + BAR_WFULL + 8 * slot, ph);
        if (elect_one()) mbar_arrive_remote(leader_pfull + 8 * slot);
        slot += SCHEME ? 2 : 1;
        if (slot == NSLOT) { slot = 0; ph ^= 1; }
      }
    }
  }
}

// UNIT op = a run of `cnt` consecutive K steps of one (128 x PAIR?2:1) x N block.  The whole warp runs the loop
// (warp-uniform), one elected lane issues.
// SCHEME 0 (bf16x3): a step = one ring slot = 3 MMAs (Ahi*Bhi + Alo*Bhi + Ahi*Blo).
// SCHEME 1 (f16f8, pairs only): same 16-wide steps, slots and A-operand stepping, but a step is TWO MMAs: the fp16 main
// term (a16 x w16, K = 16) and one K = 32 FP8 (e5m2 x e4m3) correction MMA -- r8 x w8 on even steps, a8 x s8 on odd steps of a
// 32-wide pair (the A region keeps [r8 r8 a8 a8] K groups per pair, the slot [w16: 2 K groups | w8 or s8: 2 K groups]); the
// ring hand-shake unit is the slot PAIR.
// Steps are issued in CHUNKS of up to 4 slots (= one 64-wide K quarter: 8 f16f8 / 12 bf16x3 MMAs, 1024 / 1536 tensor cycles):
// all of the chunk's ring barriers are probed back to back (their ~90-cycle latencies overlap), then every MMA and one
// tcgen05.commit per hand-shake unit are issued in one elected region.  Per-iteration fixed costs (probe round trips, elect,
// descriptor set-up) are paid once per chunk, so the issuer runs ahead of the tensor pipe instead of pacing it.
// TS: the kernel uses A-from-TMEM units (op bit 29); compiled out otherwise to keep the loop small.
template <int PAIR, int RING_BYTES, int SCHEME = 0, int TS = 1>
__device__ __forceinline__ void mma_loop(const uint32_t* __restrict__ program, uint32_t a_base, uint32_t ring,
                                         uint32_t bar, uint32_t tmem, long long ntiles) {
  static_assert(SCHEME == 0 || PAIR == 1, "the f16f8 scheme is built for CTA pairs");
  constexpr uint32_t SLOT_BYTES = PAIR ? 8192 : 16384, NSLOT = RING_BYTES / SLOT_BYTES;
  constexpr uint32_t USLOTS = SCHEME ? 2 : 1;          // ring slots per hand-shake unit
  // hand-shake units per issue chunk: a 64-wide K quarter, but never more than half of the ring (the other half refills)
  constexpr int CH = (4 / USLOTS) < (NSLOT / USLOTS / 2) ? (4 / USLOTS) : (NSLOT / USLOTS / 2);
  static_assert(NSLOT % USLOTS == 0 && CH >= 1, "ring too small for chunked issue");
  uint32_t slot = 0, ph = 0, ph_a = 0;   // ph_a: bit i = parity of operand barrier i
  long long q_a = 0, q_w = 0;
  const long long q_start = prof_clock();
  // descriptors: hi word is constant (SBO = 128 B, version 1); lo word = addr >> 4 | LBO >> 4 << 16
  constexpr uint64_t kDescHi = ((uint64_t)(128 >> 4) | (1ull << 14)) << 32;
  const uint32_t a_lo32 = (a_base >> 4) | ((KG_BYTES >> 4) << 16);
  const uint32_t ring_lo32 = ring >> 4;
  for (long long t = 0; t < ntiles; ++t) {
    const bool tr = DDMI_PROFILE && blockIdx.x == 0 && t == kTraceIter && (threadIdx.x & 31) == 0;
    uint32_t trn = 0;
    for (int pc = 0;; ++pc) {
      const uint32_t op = program[pc];
      const uint32_t kind = op & 3;
      if (kind == OP_UNIT) {
        const uint32_t n = op_n(op);
        const uint32_t nloc = PAIR ? n / 2 : n;                                   // B rows held by one CTA
        const uint32_t idesc = (SCHEME ? idesc_f16_f32(256, 0) : PAIR ? idesc2_bf16_f32(0) : idesc_bf16_f32(0)) | (n << 14);   // N >> 3 at bit 17
        const uint32_t idesc8 = idesc_f8_f32(256, 0) | (n << 14);
        const uint32_t acc = tmem + ((op >> 5) & 7) * 64;
        // A operand: K groups of the shared-memory A region, or (bit 29, CTA pairs only) of tensor memory, where one
        // K group = 4 columns counted from the TMEM base (so K group 64 sits right behind a 256-column accumulator)
        const bool a_in_tmem = PAIR && TS && ((op >> 29) & 1);
        const uint32_t a_step = a_in_tmem ? 8u : 2 * (KG_BYTES >> 4);
        uint32_t ahi32 = a_in_tmem ? tmem + ((op >> 8) & 0xFF) * 4 : a_lo32 + ((op >> 8) & 0xFF) * (KG_BYTES >> 4);
        uint32_t alo32 = a_in_tmem ? tmem + ((op >> 16) & 0xFF) * 4 : a_lo32 + ((op >> 16) & 0xFF) * (KG_BYTES >> 4);
        uint32_t accum = (op >> 4) & 1;
        // half-width units (op bit 30; f16f8, N = 128): a hand-shake unit is still two 8 KB slots, but each slot holds one
        // 32-wide step PAIR of the 128-row block ([w16 | FP8] of its two steps, 2 KB each per CTA), i.e. the unit covers 64
        // K columns in 8 MMAs of half the length -- the N split that lets half of a layer's output columns finish early
        const bool half = SCHEME && ((op >> 30) & 1);
        const uint32_t ustep = half ? 2 * USLOTS : USLOTS;                          // 16-wide A steps per hand-shake unit
        const int units = ((int)((op >> 24) & 31) + 1) / (int)ustep;              // hand-shake units in this run
        trace(tr, 0x100 + pc, trn, kTraceRegion);                                // UNIT starts
        for (int j = 0; j < units; j += CH) {
          const int nu = units - j < CH ? units - j : CH;
          // ---- probe every ring barrier of the chunk back to back
          bool ready = true;
          {
            uint32_t s = slot, p = ph;
#pragma unroll
            for (int u = 0; u < CH; ++u) {
              if (u < nu) {
                const bool r = mbar_try_wait(bar + BAR_WFULL + 8 * s, p);
                const bool r2 = PAIR ? mbar_try_wait(bar + BAR_PFULL + 8 * s, p) : true;
                ready = ready && r && r2;
                s += USLOTS;
                if (s == NSLOT) { s = 0; p ^= 1; }
              }
            }
          }
          if (!ready) {
            const long long w0 = prof_clock();
            uint32_t s = slot, p = ph;
            for (int u = 0; u < nu; ++u) {
              mbar_wait(bar + BAR_WFULL + 8 * s, p);
              if (PAIR) mbar_wait(bar + BAR_PFULL + 8 * s, p);
              s += USLOTS;
              if (s == NSLOT) { s = 0; p ^= 1; }
            }
            q_w += prof_clock() - w0;
          }
          tc_fence_after();
          // ---- issue the chunk
          const bool skip_mma = dbg(2);
          if (elect_one()) {
            uint32_t s = slot, ah = ahi32, al = alo32, ac = accum;
#pragma unroll
            for (int u = 0; u < CH; ++u) {
              if (u < nu) {
                const uint32_t w0lo = (ring_lo32 + s * (SLOT_BYTES >> 4)) | (nloc << 16);   // LBO = nloc * 16 bytes
                if (SCHEME) {
                  const uint32_t w1lo = w0lo + (SLOT_BYTES >> 4);
                  // slot: [w16: 2 K groups | FP8: 2 K groups] x nloc rows x 16 B
                  const uint64_t b16_0 = kDescHi | w0lo, b8_0 = kDescHi | (w0lo + nloc * 2);
                  const uint64_t b16_1 = kDescHi | w1lo, b8_1 = kDescHi | (w1lo + nloc * 2);
                  if (half) {
                    // slot s: steps 0, 1; slot s + 1: steps 2, 3 -- each [w16 step 0 | FP8 even | w16 step 1 | FP8 odd] x 64 rows
                    if (!skip_mma) {
#pragma unroll
                      for (uint32_t p2 = 0; p2 < 2; ++p2) {
                        const uint32_t base = w0lo + p2 * (SLOT_BYTES >> 4);
                        const uint64_t c16_0 = kDescHi | base, c8_0 = kDescHi | (base + 128);
                        const uint64_t c16_1 = kDescHi | (base + 256), c8_1 = kDescHi | (base + 384);
                        if (a_in_tmem) {
                          mma2_bf16_ts(acc, ah + (2 * p2) * a_step, c16_0, idesc, p2 ? 1u : ac);
                          mma2_bf16_ts(acc, ah + (2 * p2 + 1) * a_step, c16_1, idesc, 1u);
                          mma2_f8_ts(acc, al + (2 * p2) * a_step, c8_0, idesc8, 1u);       // r8 x w8
                          mma2_f8_ts(acc, al + (2 * p2 + 1) * a_step, c8_1, idesc8, 1u);   // a8 x s8
                        } else {
                          mma2_bf16(acc, kDescHi | (ah + (2 * p2) * a_step), c16_0, idesc, p2 ? 1u : ac);
                          mma2_bf16(acc, kDescHi | (ah + (2 * p2 + 1) * a_step), c16_1, idesc, 1u);
                          mma2_f8(acc, kDescHi | (al + (2 * p2) * a_step), c8_0, idesc8, 1u);       // r8 x w8
                          mma2_f8(acc, kDescHi | (al + (2 * p2 + 1) * a_step), c8_1, idesc8, 1u);   // a8 x s8
                        }
                      }
                    }
                  } else if (a_in_tmem) {
                    mma2_bf16_ts(acc, ah, b16_0, idesc, ac);          // kind::f16 with fp16 formats (idesc)
                    mma2_bf16_ts(acc, ah + a_step, b16_1, idesc, 1u);
                    mma2_f8_ts(acc, al, b8_0, idesc8, 1u);               // r8 x w8
                    mma2_f8_ts(acc, al + a_step, b8_1, idesc8, 1u);      // a8 x s8
                  } else if (!skip_mma) {
                    mma2_bf16(acc, kDescHi | ah, b16_0, idesc, ac);
                    mma2_bf16(acc, kDescHi | (ah + a_step), b16_1, idesc, 1u);
                    mma2_f8(acc, kDescHi | al, b8_0, idesc8, 1u);
                    mma2_f8(acc, kDescHi | (al + a_step), b8_1, idesc8, 1u);
                  }
                  mma2_commit_mc(bar + BAR_WEMPTY + 8 * s + 8, 3);          // releases the slot pair
                } else {
                  const uint64_t bhi = kDescHi | w0lo, blo = kDescHi | (w0lo + nloc * 2);   // lo block at + nloc * 32 bytes
                  if (PAIR && a_in_tmem) {
                    mma2_bf16_ts(acc, ah, bhi, idesc, ac);
                    mma2_bf16_ts(acc, al, bhi, idesc, 1u);
                    mma2_bf16_ts(acc, ah, blo, idesc, 1u);
                    mma2_commit_mc(bar + BAR_WEMPTY + 8 * s, 3);
                  } else if (PAIR) {
                    mma2_bf16(acc, kDescHi | ah, bhi, idesc, ac);
                    mma2_bf16(acc, kDescHi | al, bhi, idesc, 1u);
                    mma2_bf16(acc, kDescHi | ah, blo, idesc, 1u);
                    mma2_commit_mc(bar + BAR_WEMPTY + 8 * s, 3);
                  } else {
                    mma_bf16(acc, kDescHi | ah, bhi, idesc, ac);
                    mma_bf16(acc, kDescHi | al, bhi, idesc, 1u);
                    mma_bf16(acc, kDescHi | ah, blo, idesc, 1u);
                    mma_commit(bar + BAR_WEMPTY + 8 * s);
                  }
                }
                ac = 1u;
                ah += ustep * a_step;
                al += ustep * a_step;
                s += USLOTS;
                if (s == NSLOT) s = 0;
              }
            }
          }
          // every lane advances the (warp-uniform) run state
          accum = 1u;
          ahi32 += (uint32_t)nu * ustep * a_step;
          alo32 += (uint32_t)nu * ustep * a_step;
          for (int u = 0; u < nu; ++u) {
            slot += USLOTS;
            if (slot == NSLOT) { slot = 0; ph ^= 1; }
          }
        }
      } else if (kind == OP_WAIT) {
        const uint32_t i = (op >> 2) & 7;
        const long long w0 = prof_clock();
        trace(tr, 0x200 + pc, trn, kTraceRegion);                    // WAIT begins
        mbar_wait(bar + BAR_A0 + 8 * i, (ph_a >> i) & 1);
        ph_a ^= 1u << i;
        tc_fence_after();
        trace(tr, 0x300 + pc, trn, kTraceRegion);                    // WAIT satisfied
        q_a += prof_clock() - w0;
      } else if (kind == OP_COMMIT) {
        const uint32_t db = bar + BAR_MMADONE + 8 * ((op >> 2) & 3);
        trace(tr, 0x400 + pc, trn, kTraceRegion);                    // COMMIT issued
        if (elect_one()) {
          if (PAIR) mma2_commit_mc(db, 3);
          else      mma_commit(db);
        }
      } else {
        break;
      }
    }
  }
  if (DDMI_PROFILE && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    prof_add(3, q_a);
    prof_add(4, q_w);
    prof_add(5, prof_clock() - q_start);
    prof_add(6, ntiles);
  }
}

// ---------------------------------------------------------------------------
// epilogue helpers: one thread = one tile row (TMEM lane), 64 columns of one accumulator half
// ---------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float2 (&b)[NP]) {
#pragma unroll
  for (int i = 0; i < NP / 2; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
    b[2 * i] = make_float2(v.x, v.y);
    b[2 * i + 1] = make_float2(v.z, v.w);
  }
}
// NP pairs (2*NP consecutive columns starting at col0) -> bf16 hi/lo K groups of the 256-wide operand
template <int NP>
__device__ __forceinline__ void store_act(uint32_t h_hi, uint32_t h_lo, int row, int col0, const float2 (&y)[NP]) {
#pragma unroll
  for (int g = 0; g < NP / 4; ++g) {
    uint4 hi, lo;
    split8(&y[g * 4], hi, lo);
    const uint32_t off = (uint32_t)((col0 / 8 + g) * KG_BYTES + row * 16);
    st_shared_v4(h_hi + off, hi);
    st_shared_v4(h_lo + off, lo);
  }
}

// SCHEME-generic epilogue pieces (SCHEME 0: bf16x3, h_hi / h_lo = bases of the bf16 hi / lo K groups of an operand
// region; SCHEME 1: f16f8, h_hi = base of its fp16 K groups, h_lo = base of its FP8 K groups [r8 r8 a8 a8] per 32 columns,
// and accumulators hold 4096 x the GEMM result).
template <int SCHEME>
__device__ __forceinline__ float2 acc_plus(float2 acc, float2 b) {   // accumulator value (de-scaled) + b
  return SCHEME ? __ffma2_rn(acc, make_float2(kF8InvScale, kF8InvScale), b) : __fadd2_rn(acc, b);
}
// 16 consecutive columns starting at col0 (a multiple of 16)
template <int SCHEME>
__device__ __forceinline__ void store16(uint32_t h_hi, uint32_t h_lo, int row, int col0, const float2 (&y)[8]) {
  if (SCHEME) {
    uint4 a16[2], r8, a8;
    split16_f16f8(y, a16, r8, a8);
    const uint32_t o16 = (uint32_t)((col0 / 8) * KG_BYTES + row * 16);
    const uint32_t o8 = (uint32_t)(((col0 / 32) * 4 + ((col0 / 16) & 1)) * KG_BYTES + row * 16);
    st_shared_v4(h_hi + o16, a16[0]);
    st_shared_v4(h_hi + o16 + KG_BYTES, a16[1]);
    st_shared_v4(h_lo + o8, r8);
    st_shared_v4(h_lo + o8 + 2 * KG_BYTES, a8);
  } else {
    store_act<8>(h_hi, h_lo, row, col0, y);
  }
}
// 8 consecutive columns = K group kg of the region (gathers: one thread owns 8 channels at a time)
template <int SCHEME>
__device__ __forceinline__ void store8(uint32_t h_hi, uint32_t h_lo, int row, int kg, const float (&y)[8]) {
  if (SCHEME) {
    uint4 a16;
    uint2 r8, a8;
    split8_f16f8(y, a16, r8, a8);
    st_shared_v4(h_hi + (uint32_t)(kg * KG_BYTES + row * 16), a16);
    const uint32_t o8 = (uint32_t)(((kg >> 2) * 4 + ((kg >> 1) & 1)) * KG_BYTES + row * 16 + (kg & 1) * 8);
    st_shared_v2(h_lo + o8, r8);
    st_shared_v2(h_lo + o8 + 2 * KG_BYTES, a8);
  } else {
    uint4 hi, lo;
    split8(y, hi, lo);
    const uint32_t off = (uint32_t)(kg * KG_BYTES + row * 16);
    st_shared_v4(h_hi + off, hi);
    st_shared_v4(h_lo + off, lo);
  }
}

// relu of an f16f8 operand word = the word with its negative channels cleared (fp16 pairs / e5m2 quads; the sign of every
// format is the value's own: cvt keeps it even when the magnitude rounds to zero), i.e. split(relu(v)) bit for bit
__device__ __forceinline__ uint32_t relu_h2(uint32_t h) { return h & ~(((h >> 15) & 0x00010001u) * 0xFFFFu); }
__device__ __forceinline__ uint32_t neg_mask8(uint32_t a8) { return ((a8 >> 7) & 0x01010101u) * 0xFFu; }
// K group kg of a row, raw -> (a_hi, a_lo) and relu'd -> (b_hi, b_lo): f16f8 converts once and masks
template <int SCHEME>
__device__ __forceinline__ void store8_raw_relu(uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int row, int kg,
                                                const float (&y)[8]) {
  if (SCHEME) {
    uint4 a16;
    uint2 r8, a8;
    split8_f16f8(y, a16, r8, a8);
    const uint32_t o16 = (uint32_t)(kg * KG_BYTES + row * 16);
    const uint32_t o8 = (uint32_t)(((kg >> 2) * 4 + ((kg >> 1) & 1)) * KG_BYTES + row * 16 + (kg & 1) * 8);
    st_shared_v4(a_hi + o16, a16);
    st_shared_v2(a_lo + o8, r8);
    st_shared_v2(a_lo + o8 + 2 * KG_BYTES, a8);
    const uint2 m = make_uint2(~neg_mask8(a8.x), ~neg_mask8(a8.y));
    st_shared_v4(b_hi + o16, make_uint4(relu_h2(a16.x), relu_h2(a16.y), relu_h2(a16.z), relu_h2(a16.w)));
    st_shared_v2(b_lo + o8, make_uint2(r8.x & m.x, r8.y & m.y));
    st_shared_v2(b_lo + o8 + 2 * KG_BYTES, make_uint2(a8.x & m.x, a8.y & m.y));
  } else {
    float yr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) yr[i] = fmaxf(y[i], 0.f);
    store8<SCHEME>(a_hi, a_lo, row, kg, y);
    store8<SCHEME>(b_hi, b_lo, row, kg, yr);
  }
}

// ---------------------------------------------------------------------------
// kernel prologue / epilogue shared by every decoder
// ---------------------------------------------------------------------------
// Initialise the barrier block, allocate 512 TMEM columns (warp 9), sync; returns the TMEM base.
// a_hi_count: arrivals that complete operand barriers A4..A7 (default: one per E warp of the CTA or CTA pair, like A0..A3;
// kernels whose A4.. are signalled by other warps -- the occupancy kernel's gather warps -- pass their own count)
template <int PAIR, int SCHEME = 0>
__device__ __forceinline__ uint32_t engine_begin(uint8_t* smem, int off_bar, int a_hi_count = 8 * (1 + PAIR)) {
  const uint32_t bar = smem_u32(smem) + off_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < MAX_SLOTS; ++s) {
      mbar_init(bar + BAR_WFULL + 8 * s, 1);
      mbar_init(bar + BAR_WEMPTY + 8 * s, 1);
      mbar_init(bar + BAR_PFULL + 8 * s, 1);
    }
    for (int q = 0; q < 4; ++q) mbar_init(bar + BAR_MMADONE + 8 * q, 1);
    for (int q = 0; q < 8; ++q) mbar_init(bar + BAR_A0 + 8 * q, q < 4 ? 8 * (1 + PAIR) : a_hi_count);   // A0..A3: one arrival per E warp (of both CTAs)
    fence_mbar_init();
  }
  if (warp == 9) {
    if (PAIR) tmem_alloc2(bar + TMEM_SLOT, 512);
    else tmem_alloc(bar + TMEM_SLOT, 512);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem + off_bar + TMEM_SLOT);
}
template <int PAIR>
__device__ __forceinline__ void engine_end(uint32_t tmem) {
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if ((threadIdx.x >> 5) == 9) {
    if (PAIR) tmem_dealloc2(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}
// Warps 8..11: weight producer, MMA issuer (leader) / forwarder (peer).
template <int PAIR, int RING_BYTES, int SCHEME = 0, int TS = 1>
__device__ __forceinline__ void engine_service_warps(const uint32_t* __restrict__ program, const uint8_t* __restrict__ wstream,
                                                     uint32_t sbase, uint32_t ring, uint32_t bar, uint32_t tmem,
                                                     long long ntiles, uint32_t rank) {
  reg_dec<72>();   // the whole third warpgroup (warps 8-11) executes this one instruction
  const int warp = threadIdx.x >> 5;
  if (warp == 8) {
    producer_loop<PAIR, RING_BYTES, SCHEME>(program, wstream, ring, bar, ntiles, rank);   // whole warp, one elected lane issues
  } else if (warp == 9) {
    if (rank == 0) mma_loop<PAIR, RING_BYTES, SCHEME, TS>(program, sbase, ring, bar, tmem, ntiles);
    else forward_loop<RING_BYTES, SCHEME>(program, bar, ntiles);
  }
}

// Host: launch a decode kernel as 2-CTA clusters (pair) or plain CTAs.
template <class Kernel, class... Args>
static inline cudaError_t launch_engine(Kernel kernel, int pair, unsigned ctas, size_t smem, cudaStream_t st, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kernel);
    fprintf(stderr, "ddmi_b200: cudaFuncSetAttribute(smem=%zu) failed: static smem %zu, regs %d, max dyn %d\n", smem,
            fa.sharedSizeBytes, fa.numRegs, fa.maxDynamicSharedSizeBytes);
    return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// Host: the op table as a kernel parameter (END-padded); false if it does not fit.
static inline bool make_program_param(const uint32_t* prog, size_t words, ProgramParam* out) {
  if (words > (size_t)PROG_MAX) return false;
  for (int i = 0; i < PROG_MAX; ++i) out->op[i] = (size_t)i < words ? prog[i] : OP_END;
  return true;
}

// Walk a program on the host: number of weight bytes it consumes; -1 if malformed.
static inline long long program_stream_bytes(const uint32_t* prog, size_t words) {
  long long bytes = 0;
  for (size_t i = 0; i < words; ++i) {
    const uint32_t kind = prog[i] & 3;
    if (kind == OP_END) return (i + 1 < words) ? bytes : -1;   // needs >= 1 END of padding after the first
    if (kind == OP_UNIT) {
      const uint32_t c = (prog[i] >> 2) & 3;
      bytes += (long long)(c == 0 ? 128 : (c == 1 ? 256 : (c == 2 ? 16 : 64))) * 64 * (((prog[i] >> 24) & 31) + 1);
    }
  }
  return -1;
}

}  // namespace ummak
}  // namespace ddmi
