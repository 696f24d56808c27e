// MMA issuer of image_umma_kernel<.., TS = 1>: the op table of packing._pack_image_ts ('full' N split) written out as
// straight-line code.
//
// Why not the interpreter (umma_engine.cuh::mma_loop): a TMEM-operand half-width unit is only 8 x 74 = 592 tensor cycles, and
// the interpreter needs ~1500 issuer cycles per such unit (op fetch + decode ~360, a satisfied WAIT ~160, probes + descriptor
// moves to uniform registers + 8 MMAs + commit ~980: profiles/r02_image_ts_timeline_interpreter.txt) -- the ISSUER, a single
// warp executing dependent scalar code, paced the kernel at 2.4e8 coords/s while the tensor pipe idled.  Here every K-group
// number, accumulator column, ring slot and barrier index is a compile-time constant: an MMA costs one or two uniform adds.
// The weight producer and the peer CTA's forwarder still interpret the op table (they only need byte counts), so program,
// stream and this file must describe the same sequence; tests/test_umma_program_cpu.py pins the program, the GPU parity
// tests pin this file against it.
//
// Ring arithmetic: a hand-shake unit is two 8 KB slots; a tile consumes 116 units = 29 trips round the 8-slot ring, so every
// tile starts at slot 0 with the ring phase flipped.
#pragma once
#include "umma_engine.cuh"

namespace ddmi {
namespace ummak {

// The op table this file writes out (packing._pack_image_ts): number of ops and FNV-1a over their bytes.  The launcher checks the
// program it is handed against these, tests/test_umma_program_cpu.py pins them on the packing side.
constexpr int kImageTsProgramOps = 190;
constexpr uint32_t kImageTsProgramHash = 0x97232c81u;

struct TsIssuer {
  uint32_t bar, tmem, ring_lo32, x16_lo32, x8_lo32;   // x*: descriptor low words of X's fp16 / FP8 K groups in shared memory
  uint32_t ph;                                        // ring phase of the slots about to be consumed
  uint32_t ph_a;                                      // bit i: parity of operand barrier i
  bool tr;
  uint32_t trn;
};

constexpr uint64_t kTsDescHi = ((uint64_t)(128 >> 4) | (1ull << 14)) << 32;   // SBO = 128 B, descriptor version 1
constexpr uint32_t kTsSlot16 = 8192 >> 4;                                      // one ring slot in descriptor units

// compact spin (mbar_wait's clock-based watchdog is ~12 instructions per site): the issuer has ~130 wait sites per tile
__device__ __forceinline__ void ts_spin(uint32_t bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity))
    if (++n > (1u << 24)) __trap();   // watchdog by iteration count (0.3 s of spinning at least): a protocol bug traps instead of hanging the GPU
}

template <int I>
__device__ __forceinline__ void ts_wait(TsIssuer& c) {
  const uint32_t par = (c.ph_a >> I) & 1;
  if (!mbar_try_wait(c.bar + BAR_A0 + 8 * I, par)) ts_spin(c.bar + BAR_A0 + 8 * I, par);
  c.ph_a ^= 1u << I;
  tc_fence_after();
}
template <int J>
__device__ __forceinline__ void ts_commit(TsIssuer& c) {
  if (elect_one()) mma2_commit_mc(c.bar + BAR_MMADONE + 8 * J, 3);
  trace(c.tr, 0x400 + J, c.trn, kTraceRegion);
}
// the slot pair SLOT, SLOT + 1 has landed in both CTAs
template <int SLOT>
__device__ __forceinline__ void ts_acquire(TsIssuer& c) {
  const bool r = mbar_try_wait(c.bar + BAR_WFULL + 8 * SLOT, c.ph);
  const bool r2 = mbar_try_wait(c.bar + BAR_PFULL + 8 * SLOT, c.ph);
  if (!(r && r2)) {
    ts_spin(c.bar + BAR_WFULL + 8 * SLOT, c.ph);
    ts_spin(c.bar + BAR_PFULL + 8 * SLOT, c.ph);
  }
  tc_fence_after();
}
template <int SLOT>
__device__ __forceinline__ void ts_release(TsIssuer& c) {   // inside the elected region
  mma2_commit_mc(c.bar + BAR_WEMPTY + 8 * SLOT + 8, 3);
}
template <int SLOT>
__device__ __forceinline__ void ts_advance(TsIssuer& c) {
  if (SLOT == 6) c.ph ^= 1;
}

// One tcgen05.mma whose operand addresses are a (warp-uniform) base register plus an immediate, formed INSIDE the asm block:
// written in C++ the compiler hoists all ~900 descriptors of a tile out of the tile loop, spills them and feeds every MMA
// through local-memory loads and R2UR moves (~90 instructions per 8-MMA unit); like this an MMA is two uniform adds.
// A_TMEM: the A operand is a tensor-memory address (a_base + A_OFF columns), else a shared-memory descriptor low word.
template <bool F8, bool A_TMEM, uint32_t A_OFF, uint32_t B_OFF, uint32_t IDESC, uint32_t ACCUM>
__device__ __forceinline__ void ts_mma(uint32_t acc, uint32_t a_base, uint32_t b_base) {
  if (A_TMEM) {
    if (F8)
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b32 a, lo, ac;\n\t.reg .b64 d;\n\t"
          "mov.b32 ac, %5;\n\tsetp.ne.b32 p, ac, 0;\n\t"
          "add.u32 a, %1, %3;\n\tadd.u32 lo, %2, %4;\n\tmov.b64 d, {lo, 0x4008};\n\t"
          "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [a], d, %6, p;\n\t}" ::"r"(acc), "r"(a_base), "r"(b_base), "n"(A_OFF),
          "n"(B_OFF), "n"(ACCUM), "n"(IDESC)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b32 a, lo, ac;\n\t.reg .b64 d;\n\t"
          "mov.b32 ac, %5;\n\tsetp.ne.b32 p, ac, 0;\n\t"
          "add.u32 a, %1, %3;\n\tadd.u32 lo, %2, %4;\n\tmov.b64 d, {lo, 0x4008};\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], [a], d, %6, p;\n\t}" ::"r"(acc), "r"(a_base), "r"(b_base), "n"(A_OFF),
          "n"(B_OFF), "n"(ACCUM), "n"(IDESC)
          : "memory");
  } else {
    if (F8)
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b32 a, lo, ac;\n\t.reg .b64 d, e;\n\t"
          "mov.b32 ac, %5;\n\tsetp.ne.b32 p, ac, 0;\n\t"
          "add.u32 a, %1, %3;\n\tadd.u32 lo, %2, %4;\n\tmov.b64 d, {lo, 0x4008};\n\tmov.b64 e, {a, 0x4008};\n\t"
          "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], e, d, %6, p;\n\t}" ::"r"(acc), "r"(a_base), "r"(b_base), "n"(A_OFF),
          "n"(B_OFF), "n"(ACCUM), "n"(IDESC)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b32 a, lo, ac;\n\t.reg .b64 d, e;\n\t"
          "mov.b32 ac, %5;\n\tsetp.ne.b32 p, ac, 0;\n\t"
          "add.u32 a, %1, %3;\n\tadd.u32 lo, %2, %4;\n\tmov.b64 d, {lo, 0x4008};\n\tmov.b64 e, {a, 0x4008};\n\t"
          "tcgen05.mma.cta_group::2.kind::f16 [%0], e, d, %6, p;\n\t}" ::"r"(acc), "r"(a_base), "r"(b_base), "n"(A_OFF),
          "n"(B_OFF), "n"(ACCUM), "n"(IDESC)
          : "memory");
  }
}
static_assert((kTsDescHi >> 32) == 0x4008, "descriptor high word is spelled out in ts_mma");
constexpr uint32_t kTsI16h = idesc_f16_f32(256, 0) | (128u << 14), kTsI8h = idesc_f8_f32(256, 0) | (128u << 14);   // N = 128
constexpr uint32_t kTsI16f = idesc_f16_f32(256, 0) | (256u << 14), kTsI8f = idesc_f8_f32(256, 0) | (256u << 14);   // N = 256
constexpr uint32_t kTsLboH = 64u << 16, kTsLboF = 128u << 16;     // B operand: N/2 rows x 16 B between K groups
constexpr uint32_t kTsKstep = 2 * (KG_BYTES >> 4);                // one 16-wide step of a shared-memory A operand

// Half-width (N = 128) unit over H quarter Q (64 K columns) from tensor memory into accumulator columns ACC..ACC+127.
// Slot s + p: the 32-wide step pair p, [w16 step 0 | FP8 even | w16 step 1 | FP8 odd] x 64 rows x 32 B.
template <int SLOT, int ACC, int Q, bool FIRST>
__device__ __forceinline__ void ts_unit_h(TsIssuer& c) {
  trace(c.tr, 0x500, c.trn, kTraceRegion);
  ts_acquire<SLOT>(c);
  trace(c.tr, 0x100 + Q + ACC / 32, c.trn, kTraceRegion);
  if (elect_one()) {
    const uint32_t acc = c.tmem + ACC;
    constexpr uint32_t A16 = 256 + 32 * Q, A8 = 384 + 32 * Q;
    constexpr uint32_t B0 = SLOT * kTsSlot16 + kTsLboH, B1 = B0 + kTsSlot16;
    ts_mma<false, true, A16, B0, kTsI16h, FIRST ? 0u : 1u>(acc, c.tmem, c.ring_lo32);
    ts_mma<false, true, A16 + 8, B0 + 256, kTsI16h, 1u>(acc, c.tmem, c.ring_lo32);
    ts_mma<true, true, A8, B0 + 128, kTsI8h, 1u>(acc, c.tmem, c.ring_lo32);          // r8 x w8
    ts_mma<true, true, A8 + 8, B0 + 384, kTsI8h, 1u>(acc, c.tmem, c.ring_lo32);      // a8 x s8
    ts_mma<false, true, A16 + 16, B1, kTsI16h, 1u>(acc, c.tmem, c.ring_lo32);
    ts_mma<false, true, A16 + 24, B1 + 256, kTsI16h, 1u>(acc, c.tmem, c.ring_lo32);
    ts_mma<true, true, A8 + 16, B1 + 128, kTsI8h, 1u>(acc, c.tmem, c.ring_lo32);
    ts_mma<true, true, A8 + 24, B1 + 384, kTsI8h, 1u>(acc, c.tmem, c.ring_lo32);
    trace(c.tr, 0x501, c.trn, kTraceRegion);
    ts_release<SLOT>(c);
    trace(c.tr, 0x502, c.trn, kTraceRegion);
  }
  ts_advance<SLOT>(c);
}
// Half-width unit over X (K = 64, shared memory) -- block 0, whose input is the PE features alone.
template <int SLOT, int ACC>
__device__ __forceinline__ void ts_unit_x_half(TsIssuer& c) {
  ts_acquire<SLOT>(c);
  trace(c.tr, 0x180 + ACC / 32, c.trn, kTraceRegion);
  if (elect_one()) {
    const uint32_t acc = c.tmem + ACC;
    constexpr uint32_t B0 = SLOT * kTsSlot16 + kTsLboH, B1 = B0 + kTsSlot16;
    ts_mma<false, false, 0, B0, kTsI16h, 0u>(acc, c.x16_lo32, c.ring_lo32);
    ts_mma<false, false, kTsKstep, B0 + 256, kTsI16h, 1u>(acc, c.x16_lo32, c.ring_lo32);
    ts_mma<true, false, 0, B0 + 128, kTsI8h, 1u>(acc, c.x8_lo32, c.ring_lo32);
    ts_mma<true, false, kTsKstep, B0 + 384, kTsI8h, 1u>(acc, c.x8_lo32, c.ring_lo32);
    ts_mma<false, false, 2 * kTsKstep, B1, kTsI16h, 1u>(acc, c.x16_lo32, c.ring_lo32);
    ts_mma<false, false, 3 * kTsKstep, B1 + 256, kTsI16h, 1u>(acc, c.x16_lo32, c.ring_lo32);
    ts_mma<true, false, 2 * kTsKstep, B1 + 128, kTsI8h, 1u>(acc, c.x8_lo32, c.ring_lo32);
    ts_mma<true, false, 3 * kTsKstep, B1 + 384, kTsI8h, 1u>(acc, c.x8_lo32, c.ring_lo32);
    ts_release<SLOT>(c);
  }
  ts_advance<SLOT>(c);
}
// Full-width (N = 256) unit over X's 32-wide step pair P, accumulating onto columns 0..255.
// Slot s = step 0, slot s + 1 = step 1, each [w16: 2 K groups | FP8: 2 K groups] x 128 rows x 16 B.
template <int SLOT, int P>
__device__ __forceinline__ void ts_unit_x_full(TsIssuer& c) {
  ts_acquire<SLOT>(c);
  trace(c.tr, 0x1a0 + P, c.trn, kTraceRegion);
  if (elect_one()) {
    constexpr uint32_t B0 = SLOT * kTsSlot16 + kTsLboF, B1 = B0 + kTsSlot16;
    ts_mma<false, false, (2 * P) * kTsKstep, B0, kTsI16f, 1u>(c.tmem, c.x16_lo32, c.ring_lo32);
    ts_mma<false, false, (2 * P + 1) * kTsKstep, B1, kTsI16f, 1u>(c.tmem, c.x16_lo32, c.ring_lo32);
    ts_mma<true, false, (2 * P) * kTsKstep, B0 + 256, kTsI8f, 1u>(c.tmem, c.x8_lo32, c.ring_lo32);
    ts_mma<true, false, (2 * P + 1) * kTsKstep, B1 + 256, kTsI8f, 1u>(c.tmem, c.x8_lo32, c.ring_lo32);
    ts_release<SLOT>(c);
  }
  ts_advance<SLOT>(c);
}

// One GEMM group over [H (256) | X (64, optional)]: packing._pack_image_ts.group(W, 256, k_x, published), 'full' split.
// SLOT0: ring slot at entry (0 or 4); 8 (+ 2 with X) hand-shake units.
template <int SLOT0, bool HAS_X, bool PUBLISHED>
__device__ __forceinline__ void ts_group_h(TsIssuer& c) {
  constexpr int S = SLOT0;
  ts_wait<5>(c);
  if (PUBLISHED) ts_wait<0>(c);
  ts_unit_h<(S + 0) % 8, 0, 0, true>(c);
  if (PUBLISHED) ts_wait<1>(c);
  ts_unit_h<(S + 2) % 8, 0, 1, false>(c);
  ts_wait<4>(c);
  ts_unit_h<(S + 4) % 8, 128, 0, true>(c);
  ts_unit_h<(S + 6) % 8, 128, 1, false>(c);
  constexpr int T = HAS_X ? S + 4 : S;      // (S + 8 + 4) % 8 after the two X units
  if (HAS_X) {
    ts_unit_x_full<(S + 0) % 8, 0>(c);
    ts_unit_x_full<(S + 2) % 8, 1>(c);
  }
  if (PUBLISHED) ts_wait<2>(c);
  ts_unit_h<(T + 0) % 8, 0, 2, false>(c);
  if (PUBLISHED) ts_wait<3>(c);
  ts_unit_h<(T + 2) % 8, 0, 3, false>(c);
  ts_commit<0>(c);
  ts_unit_h<(T + 4) % 8, 128, 2, false>(c);
  ts_unit_h<(T + 6) % 8, 128, 3, false>(c);
  ts_commit<1>(c);
}
// Block 0's groups: X only (two half-width units).
template <int SLOT0, bool PUBLISHED>
__device__ __forceinline__ void ts_group_x(TsIssuer& c) {
  if (PUBLISHED) {
    ts_wait<0>(c);
    ts_wait<1>(c);
    ts_wait<2>(c);
    ts_wait<3>(c);
  }
  ts_wait<5>(c);
  ts_unit_x_half<SLOT0, 0>(c);
  ts_commit<0>(c);
  ts_wait<4>(c);
  ts_unit_x_half<(SLOT0 + 2) % 8, 128>(c);
  ts_commit<1>(c);
}

__device__ __forceinline__ void image_ts_issue_loop(uint32_t sbase, uint32_t ring, uint32_t bar, uint32_t tmem, int x16_kg,
                                                    int x8_kg, long long ntiles) {
  TsIssuer c;
  // __shfl_sync(.., 0) tells the compiler these are warp-uniform (the TMEM base comes out of shared memory): descriptors are
  // then built with uniform adds of immediates instead of being precomputed per lane, spilled and moved over with R2UR
  c.bar = __shfl_sync(0xffffffffu, bar, 0);
  c.tmem = __shfl_sync(0xffffffffu, tmem, 0);
  c.ring_lo32 = __shfl_sync(0xffffffffu, ring >> 4, 0);
  const uint32_t a_lo32 = (__shfl_sync(0xffffffffu, sbase, 0) >> 4) | ((KG_BYTES >> 4) << 16);
  c.x16_lo32 = a_lo32 + x16_kg * (KG_BYTES >> 4);
  c.x8_lo32 = a_lo32 + x8_kg * (KG_BYTES >> 4);
  c.ph = 0;
  c.ph_a = 0;
  c.trn = 0;
  const long long q_start = prof_clock();
  for (long long t = 0; t < ntiles; ++t) {
    c.tr = DDMI_PROFILE && blockIdx.x == 0 && t == kTraceIter && (threadIdx.x & 31) == 0;
    c.trn = 0;
    // block 0 (res1): skip, conv1 over X; conv2, conv3 over H
    ts_group_x<0, true>(c);
    ts_group_x<4, false>(c);
#pragma unroll 1
    for (int g = 0; g < 2; ++g) ts_group_h<0, false, true>(c);
#pragma unroll 1
    for (int blk = 1; blk < 3; ++blk) {   // res2, res3: skip and conv1 over [H | X] (10 units each: 4 -> 0), conv2, conv3
      ts_group_h<0, true, true>(c);
      ts_group_h<4, true, false>(c);
#pragma unroll 1
      for (int g = 0; g < 2; ++g) ts_group_h<0, false, true>(c);
    }
#pragma unroll 1
    for (int g = 0; g < 3; ++g) ts_group_h<0, false, true>(c);   // res4: conv1, conv2, conv3
  }
  if (DDMI_PROFILE && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    prof_add(5, prof_clock() - q_start);
    prof_add(6, ntiles);
  }
}

}  // namespace ummak
}  // namespace ddmi
