// 5G NR LDPC layered normalized min-sum decoder - throughput kernel.
//
// Two codeblocks of the same (base graph, lifting size) are decoded together: the two 16-bit halves of every 32-bit
// word belong to codeblock A and codeblock B, so one instruction works on both and - because both use the same lifted
// graph - no lane rotation is ever needed. Thread j owns lifted check node j of every base-graph row.
//
// Messages and soft bits are small integers held exactly in IEEE half precision (|x| <= 2048 is exact), which buys:
//   * +-infinity ("fixed bit", the reference's +-127) with hardware-native sticky semantics: INF - c = INF,
//   * packed add/fma/relu on the FMA pipe (HADD2/HFMA2), leaving the ALU pipe (LOP3/PRMT/HMNMX2, 64 lanes/clk/SM on
//     B200 - the bound of this kernel) for sign logic and min/second-min only; both pipes issue in parallel.
// Check-to-variable messages are kept compressed per (row, check) as {scaled min1, scaled min2, one "negative" bit per
// edge, the index of the edge that held the minimum}, from which every message is exactly reconstructible; they live in a
// per-CTA scratch in global memory (mostly L2: see DESIGN.md 4.1 for what reaches DRAM) and are prefetched one row ahead.
// For the hot shape (BG1, Z = 384) the layered schedule is compiled in (row_programs.inc, spec_row); every other shape runs
// the same arithmetic (row_math) from the edge tables.
//
// Bit-exact target: ldpc_decoder_impl::decode (lib/phy/upper/channel_coding/ldpc/ldpc_decoder_impl.cpp:60-318) with the
// AVX2/AVX512 or generic node kernels (ldpc_decoder_avx512.cpp:81-290, ldpc_decoder_generic.cpp:30-128); see SURVEY 8a
// R7-R12 for the integer semantics reproduced here.
#pragma once

#include "pdc_device.cuh"
#include <cuda_fp16.h>
#include <utility>

namespace pdc {
namespace h2 {

typedef uint32_t hh; // two halves: low = codeblock A, high = codeblock B

__device__ __forceinline__ __half2 H(hh x)
{
  return *reinterpret_cast<__half2*>(&x);
}
__device__ __forceinline__ hh U(__half2 x)
{
  return *reinterpret_cast<hh*>(&x);
}
__device__ __forceinline__ hh lop_and_or(hh a, hh b, hh c) // (a & b) | c
{
  hh r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// min(|a|, |b|) carrying the product of the two signs (HMNMX2.XORSIGN): a clamp of a to +-|b| when b is the bound, and
// a running minimum of magnitudes that accumulates the sign parity in its sign bit.
__device__ __forceinline__ __half2 min_abs_xorsign(__half2 a, __half2 b)
{
  hh r;
  asm("min.xorsign.abs.f16x2 %0, %1, %2;" : "=r"(r) : "r"(*reinterpret_cast<hh*>(&a)), "r"(*reinterpret_cast<hh*>(&b)));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ hh lop_xor_and(hh a, hh b, hh c) // a ^ (b & c)
{
  hh r;
  asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// Half-precision constants, both halves.
constexpr hh H_ZERO  = 0x00000000u;
constexpr hh H_ONE   = 0x3C003C00u;
constexpr hh H_NEG1  = 0xBC00BC00u;
constexpr hh H_TWO   = 0x40004000u;
constexpr hh H_HALF  = 0x38003800u;
constexpr hh H_120   = 0x57805780u;
constexpr hh H_N120  = 0xD780D780u;
constexpr hh H_N230  = 0xDB30DB30u;
constexpr hh H_N128  = 0xD800D800u;
constexpr hh H_1024  = 0x64006400u;
constexpr hh H_N1024 = 0xE400E400u;
constexpr hh H_BIG   = 0x7BFF7BFFu; // 65504: one excess unit times BIG overflows to infinity
constexpr hh H_NBIG  = 0xFBFFFBFFu;
constexpr hh H_SIGN  = 0x80008000u;
constexpr hh H_544   = 0x60406040u; // promotion slope: 121 * 544 overflows, 120 * 544 = 65280 does not
constexpr hh H_N65280 = 0xFBF8FBF8u;
constexpr hh H_0P8   = 0x3A663A66u; // 0.7998046875
constexpr hh H_N0P6  = 0xB8CDB8CDu; // -0.60009765625
constexpr hh H_1023P5 = 0x63FF63FFu; // 1023.5
constexpr hh H_2M13  = 0x08000800u; // 2^-13

// Compressed check-to-variable messages of one (row, check), both codeblocks: one 128-bit word.
//   rows of degree <= 16: {scaled min1 (half2), scaled min2 (half2), "negative" flags, index of the minimum (half2)}
//   rows of degree 19   : {scaled minima as bytes {m1 A, m1 B, m2 A, m2 B}, flags of edges 0-15, flags of edges 16-18,
//                          index of the minimum}
// A flag word holds, per codeblock half, one bit per edge: edges 0-7 of the word in bits 7..0 (first edge in bit 7),
// edges 8-15 in bits 15..8; the bits above a single group are don't-care. The index of the edge whose message is min2
// is kept as a small integer in half precision (compared with HSET2, selected with one LOP3).
typedef uint4 RowState;

__device__ __forceinline__ hh& st_word(RowState& s, int k)
{
  return (k == 0) ? s.x : (k == 1) ? s.y : (k == 2) ? s.z : s.w;
}
__device__ __forceinline__ hh st_word(const RowState& s, int k)
{
  return (k == 0) ? s.x : (k == 1) ? s.y : (k == 2) ? s.z : s.w;
}

// The row state is re-read every iteration and must stay in L2 while LLR batches stream through it: its accesses carry
// an evict-last policy, the soft-bit input an evict-first one.
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_state(const uint4* p, uint64_t pol)
{
  uint4 v;
  asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
// The same from base + off (a compile-time byte offset) if `far`, else from base: two predicated loads instead of a select
// and a 64-bit add in front of one.
// (Always fetching slot M + 1, with row 0 leaving a copy of its messages in the slot behind the last row in use, saves
// the second load - and measured slower in the step: 3.06 vs 3.03 ms.)
template <uint32_t OFF>
__device__ __forceinline__ uint4 ld_state_sel(const uint4* base, bool far, uint64_t pol)
{
  uint4 v;
  asm volatile("{\n\t"
               ".reg .pred p;\n\t"
               "setp.ne.u32 p, %6, 0;\n\t"
               "@p ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4+%7], %5;\n\t"
               "@!p ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;\n\t"
               "}"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(base), "l"(pol), "r"((uint32_t)far), "n"(OFF));
  return v;
}
__device__ __forceinline__ void st_state(uint4* p, const uint4& v, uint64_t pol)
{
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p, uint64_t pol)
{
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// Shared memory through explicit 32-bit shared-window addresses (a generic pointer makes the compiler rebuild the window
// base from a special register in every row).
#ifdef PDC_DEBUG_BOUNDS
// [a, a + n) inside the kernel's dynamic shared memory?
__device__ __forceinline__ bool smem_range_ok(uint32_t a, uint32_t n)
{
  extern __shared__ __align__(16) unsigned char smem_dbg_base[];
  uint32_t size;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(size));
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_dbg_base);
  return a >= base && a + n <= base + size && (a % n) == 0;
}
#endif

__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
  PDC_ASSERT(smem_range_ok(a, 4));
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v)
{
  PDC_ASSERT(smem_range_ok(a, 4));
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a)
{
  PDC_ASSERT(smem_range_ok(a, 16));
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

// The lifted graph of this (base graph, Z) in shared memory.
struct GraphSmem {
  // Per edge {circulant shift in bytes, shared-window address of the first soft word of its variable node}; every row
  // starts on a 16-byte boundary so that one 128-bit load brings two edges.
  uint2    einfo[MAX_EDGES + MAX_ROWS];
  uint32_t row_info[MAX_ROWS]; // first einfo entry | degree << 16 | "no barrier needed" << 31
};


// Global-memory images of what a CTA needs per (base graph, lifting size), built once on the host (upload_h2_images):
// the per-edge table in the layout of GraphSmem (column bases relative to the soft-bit array), the row table and the CRC
// word weights. A CTA copies them with coalesced loads instead of deriving them from the constant-memory base graph
// with one (serialised) constant access per thread, and only when the shape differs from that of its previous pair.
#ifndef H2_NO_REUSE
#define H2_NO_REUSE 0
#endif
constexpr int H2_Z_SLOTS = 51;
__device__ uint2    g_h2_einfo[2][H2_Z_SLOTS][MAX_EDGES + MAX_ROWS];
__device__ uint32_t g_h2_row_info[2][MAX_ROWS];
__device__ uint32_t g_h2_xpow32[3][XPOW_ENTRIES];
__constant__ uint8_t c_h2_z_slot[MAX_Z + 1]; // uniform index: one constant-cache access

// CTA barrier that threads of one warp may reach from different places (partial last warp of a lifting size that is not
// a multiple of 32): the non-aligned form.
__device__ __forceinline__ void row_barrier()
{
  asm volatile("barrier.sync 0;" ::: "memory");
}

// Bit pattern of the half-precision number e (0 <= e < 1024), in both halves.
__host__ __device__ constexpr hh half2_of_small_int(int e)
{
  if (e == 0) {
    return 0u;
  }
  int ex = 0;
  while ((2 << ex) <= e) {
    ++ex;
  }
  const uint32_t h = ((uint32_t)(15 + ex) << 10) | (((uint32_t)e - (1u << ex)) << (10 - ex));
  return h * 0x00010001u;
}
static_assert(half2_of_small_int(1) == 0x3C003C00u && half2_of_small_int(3) == 0x42004200u &&
                  half2_of_small_int(18) == 0x4C804C80u,
              "half-precision encoding of small integers");

// The arithmetic of one base-graph row (layer) of degree DEG for one check of both codeblocks, given the shared-window
// addresses of the row's soft words (formed by the caller BEFORE the barrier that orders the row after the previous one:
// they do not depend on it, so the wait overlaps useful work).
//   st    : in: compressed messages of this row from the previous iteration; out: those of the NEXT row (fetched from
//           sp_next as soon as this row's are consumed). This row's new messages are stored to sp.
//   SCALE : PDC_SCALE_X86 / PDC_SCALE_GENERIC compiled in, or -1: scale_mode decides at run time.
// What pass 1 of a row hands to pass 2.
template <int DEG>
struct RowPass {
  __half2 vc[DEG]; // soft - c2v_old clamped to +-120, infinite if the soft bit was infinite
  __half2 min1, min2;
  hh      par;
};

// Pass 1: variable-to-check messages of the row's edges (st: the row's compressed messages of the previous iteration),
// their two smallest magnitudes and the sign parity.
template <int DEG>
__device__ __forceinline__ void row_pass1(const uint32_t (&addr)[DEG], const RowState& st, RowPass<DEG>& rp)
{
  constexpr bool PACKED_MIN = DEG > 16;
  constexpr int  F0         = PACKED_MIN ? 1 : 2; // index of the first flag word

  __half2 (&vc)[DEG] = rp.vc;

  const __half2 h120 = H(H_120);

  // Old scaled minima: the message of an edge has magnitude min1, that of the edge which held the minimum min2.
  __half2 m1, m2;
  if (PACKED_MIN) {
    m1 = __hadd2(H(__byte_perm(st.x, 0x64646464u, 0x4140)), H(H_N1024));
    m2 = __hadd2(H(__byte_perm(st.x, 0x64646464u, 0x4342)), H(H_N1024));
  } else {
    m1 = H(st.x);
    m2 = H(st.y);
  }
  const __half2 d12 = __hsub2(m1, m2);
  const __half2 idx_old = H(st.w);

  // min1 carries the sign parity of the edges seen so far in its sign bits (min_abs_xorsign); its magnitude is the minimum.
  __half2 min1 = h120, min2 = h120, c_prev = h120;
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    const hh      f    = st_word(st, F0 + (e >> 4));
    const hh      ps   = f << ((e & 8) ? (e & 7) : 8 + (e & 7)); // "negative" flag of this edge in the sign position
    const __half2 s    = H(lds_u32(addr[e]));
    const __half2 eq   = __heq2(idx_old, H(half2_of_small_int(e)));
    const __half2 sg   = H(lop_and_or(ps, H_SIGN, H_ONE));
    const __half2 nmag = __hfma2(eq, d12, __hneg2(m1)); // -min2 for the edge that held the minimum, -min1 otherwise
    const __half2 v    = __hfma2(sg, nmag, s);          // s - c2v_old; |v| <= 222, or infinite with s
    // Clamp to +-120 but keep infinity: v * 2^-13 is less than half a unit in the last place of the clamped value for
    // every finite v (|v| <= 222), so the sum rounds back to it, and infinite for an infinite v.
    const __half2 cl = min_abs_xorsign(v, h120);
    const __half2 c  = __hfma2(v, H(H_2M13), cl);
    vc[e]            = c;
    // Two smallest magnitudes of the row, edges taken in pairs (shorter dependency chain, three-input minimum), on the
    // clamped values: an infinite magnitude counts as 120, so the minima never exceed 120.
    if ((e & 1) == 0 && e != DEG - 1) {
      c_prev = cl;
    } else if (e == 1) {
      min1 = min_abs_xorsign(c_prev, cl);
      min2 = __hmax2(__habs2(c_prev), __habs2(cl));
    } else if (e & 1) {
      const __half2 lo = min_abs_xorsign(c_prev, cl), hi = __hmax2(__habs2(c_prev), __habs2(cl));
      min2             = __hmin2(__hmin2(min2, hi), __hmax2(__habs2(min1), __habs2(lo)));
      min1             = min_abs_xorsign(min1, lo);
    } else {
      min2 = __hmin2(min2, __hmax2(__habs2(min1), __habs2(cl)));
      min1 = min_abs_xorsign(min1, cl);
    }
  }

  rp.min1 = __habs2(min1), rp.min2 = min2, rp.par = U(min1);
}

// Pass 2: scaled minima, new check-to-variable messages, soft-bit update, and the row's new compressed messages to sp.
template <int DEG, int SCALE>
__device__ __forceinline__ void row_pass2(const uint32_t (&addr)[DEG], const RowPass<DEG>& rp, uint4* sp, uint64_t pol,
                                          int scale_mode, bool keep = true)
{
  constexpr bool PACKED_MIN = DEG > 16;
  const __half2 (&vc)[DEG]  = rp.vc;
  const __half2 one = H(H_ONE), min1 = rp.min1, min2 = rp.min2;
  const hh      par = rp.par;

  // Scaled minima. x86: (x * 52428) >> 16 == ceil(0.8 x) - 1 for x >= 1, 0 for x = 0;
  // generic: round(0.8 x). Both are computed exactly through round-to-nearest in the [1024, 2048) binade.
  __half2 s1, s2;
  if ((SCALE < 0) ? (scale_mode == PDC_SCALE_X86) : (SCALE == PDC_SCALE_X86)) {
    // 0.7998 x + 1023.5 rounds (once, in the binade of 1024) to 1024 + ceil(0.8 x) - 1 for every x in 1..120 and stays
    // 1023.5 for x = 0, which the relu of the subtraction turns into 0.
    s1 = __hfma2_relu(__hfma2(min1, H(H_0P8), H(H_1023P5)), one, H(H_N1024));
    s2 = __hfma2_relu(__hfma2(min2, H(H_0P8), H(H_1023P5)), one, H(H_N1024));
  } else {
    s1 = __hadd2(__hfma2(min1, H(H_0P8), H(H_1024)), H(H_N1024));
    s2 = __hadd2(__hfma2(min2, H(H_0P8), H(H_1024)), H(H_N1024));
  }
  const hh par_s = lop_and_or(par, H_SIGN, H_ONE);

  // Pass 2 is balanced between the two math pipes: the magnitude select and the index of the minimum are bit operations
  // on the comparison mask (ALU pipe), the message sum, the promotion and the sign flags are packed half-precision
  // arithmetic (FMA pipe).
  const hh s_diff = U(s1) ^ U(s2);
  __half2  acc_s  = H(H_ZERO);
  hh       idx    = 0;
  hh       grp[3] = {0, 0, 0};
#pragma unroll
  for (int e = 0; e != DEG; ++e) {
    const __half2 c    = vc[e];
    const hh      ism  = __heq2_mask(__habs2(c), min1);       // 0xffff per half if this edge holds the minimum
    const __half2 mag  = H(lop_xor_and(U(s1), ism, s_diff));  // min2 for the minimum edge, min1 otherwise (scaled)
    const __half2 sgn  = H(lop_xor_and(par_s, U(c), H_SIGN)); // +-1: sign parity of the row without this edge
    const __half2 x    = __hfma2(sgn, mag, c);
    // Promotion (LLR.cpp:74-87): |x| > 120 -> +-infinity. pe = relu(544 |x| - 65280) is 0 up to |x| = 120 and at
    // least 544 from 121 on, where x * pe + x exceeds the half-precision range; an infinite x stays infinite.
    const __half2 pe = __hfma2_relu(__habs2(x), H(H_544), H(H_N65280));
    const __half2 r  = __hfma2(x, pe, x);
    sts_u32(addr[e], U(r));
    // sum of +-2^k over the edges of a group of eight; turned into "negative" bits below
    acc_s = ((e & 7) == 0) ? sgn : __hfma2(acc_s, H(H_TWO), sgn);
    // index of an edge that holds the minimum (with several, min2 == min1 and any of them serves): (ism & e) | (~ism & idx)
    // (edge 0 leaves the initial 0)
    if (e != 0) {
      asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(idx) : "r"(ism), "r"(half2_of_small_int(e)), "r"(idx));
    }
    if ((e & 7) == 7 || e == DEG - 1) {
      // acc_s = sum over the n edges of the group of sgn * 2^(n-1-k): negative-edge bits = ((2^n - 1) - acc_s) / 2,
      // left-aligned for a partial group and converted to an integer byte through the 1024 binade (bits 7..0 of each
      // half; the bits above are the exponent of 1024: don't-care for a reader that shifts its flag to the sign bit).
      const int     n_in_group = (e & 7) + 1;
      const int     fill       = 8 - n_in_group;
      const float   full       = (float)((1 << n_in_group) - 1) * 0.5f * (float)(1 << fill);
      const __half2 sc         = H(0x38003800u + (uint32_t)fill * 0x04000400u); // 0.5 * 2^fill
      if (n_in_group < 8) {
        // full + 1024 is an integer below 2048: one fused operation
        grp[e >> 3] = U(__hfma2(acc_s, __hneg2(sc), __float2half2_rn(full + 1024.0f)));
      } else {
        grp[e >> 3] = U(__hadd2(__hfma2(acc_s, __hneg2(sc), __float2half2_rn(full)), H(H_1024)));
      }
      acc_s                    = H(H_ZERO);
    }
  }
  RowState st_out;
  if (PACKED_MIN) {
    st_out.x = __byte_perm(U(__hadd2(s1, H(H_1024))), U(__hadd2(s2, H(H_1024))), 0x6420);
    st_out.y = __byte_perm(grp[0], grp[1], 0x6240);
    st_out.z = grp[2];
  } else {
    st_out.x = U(s1);
    st_out.y = U(s2);
    st_out.z = (DEG > 8) ? __byte_perm(grp[0], grp[1], 0x6240) : grp[0];
  }
  st_out.w = idx;
  if (keep) { // (nobody reads the messages of the last iteration)
    st_state(sp, st_out, pol);
  }
}

// One row: pass 1, fetch of the next row's compressed messages into the registers this row's were in, pass 2.
template <int DEG, int SCALE>
__device__ __forceinline__ void row_math(const uint32_t (&addr)[DEG], RowState& st, uint4* sp, const uint4* sp_next,
                                         uint64_t pol, int scale_mode, bool keep = true)
{
  RowPass<DEG> rp;
  row_pass1<DEG>(addr, st, rp);
  st = ld_state(sp_next, pol);
  row_pass2<DEG, SCALE>(addr, rp, sp, pol, scale_mode, keep);
}

// Table-driven row: the addresses of the row's soft words come from the edge table in shared memory.
//   e_info : shared-window address of the row's edge table
template <int DEG>
__device__ __forceinline__ void process_row(uint32_t e_info, uint32_t j4, uint32_t neg_Z4, RowState& st, uint4* sp,
                                            const uint4* sp_next, uint64_t pol, int scale_mode, bool need_barrier,
                                            bool keep = true)
{
  uint32_t addr[DEG];
  {
    uint4 ei = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int e = 0; e != DEG; ++e) {
      if ((e & 1) == 0) {
        ei = lds_u128(e_info + 8 * e);
      }
      const uint32_t t = j4 + ((e & 1) ? ei.z : ei.x);
      // wrap: t - Z underflows to a huge value when t < Z (one fused add + unsigned minimum)
      addr[e] = ((e & 1) ? ei.w : ei.y) + __viaddmin_u32(t, neg_Z4, t);
    }
  }
  if (need_barrier) {
    row_barrier();
  }
  row_math<DEG, -1>(addr, st, sp, sp_next, pol, scale_mode, keep);
}

__device__ __forceinline__ void dispatch_row(int deg, uint32_t e_info, uint32_t j4, uint32_t neg_Z4, RowState& st,
                                             uint4* sp, const uint4* sp_next, uint64_t pol, int scale_mode,
                                             bool need_barrier)
{
  // Most frequent degrees first (BG1: 18 rows of degree 5, 8 of degree 6, ...).
  if (deg == 5) {
    process_row<5>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 6) {
    process_row<6>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 4) {
    process_row<4>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 7) {
    process_row<7>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 19) {
    process_row<19>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 3) {
    process_row<3>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 8) {
    process_row<8>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else if (deg == 9) {
    process_row<9>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  } else {
    process_row<10>(e_info, j4, neg_Z4, st, sp, sp_next, pol, scale_mode, need_barrier);
  }
}


// ---- compile-time row program of the hot shape ----------------------------------------------------------------------
//
// For the shapes listed in tools/gen_row_program.py the layered schedule of the leading rows is compiled in: every
// edge's variable node and circulant shift is an immediate operand, the rows are unrolled, and a third of the edge
// addresses need no instruction at all (see the generator: per-row thread offset tau_m, extension nodes stored rotated
// by it). The arithmetic (row_math) is shared with the table-driven rows.
//
// How many rows: unrolled code is 16 bytes x ~25 instructions per edge, and the row loop has a size above which the
// instruction fetch cannot keep up. With ~29 instructions per edge (first versions of this kernel; 8192 codeblocks BG1
// Z = 384, 46 rows, 6 iterations; table-driven loop 4.14 ms) rows 0-3 compiled in measured 4.13 ms, 0-7 4.03, 0-11 3.92,
// 0-15 3.84, 0-23 3.82, and all 46 (147 KB of loop) 6.24 ms. With the arithmetic of today (131 KB for all 46 rows) the
// cliff is not reached: 28 rows 3.39 ms per step, 30 3.33, 32 3.32, 34 3.28, 40 3.28, 43 3.28, all 46 3.20 ms - a
// table-driven row costs ~200 instructions (27 of them loop control and the jump to the body of its degree) where a
// compiled-in row of degree 5 costs ~160. Anything that grows the code per edge has to be re-measured against this.
// (Two consecutive rows that share no variable node as ONE block of code - both first passes, then both second passes, so
// that the compiler has two independent instruction streams to interleave - measured 3 % SLOWER than row after row. A
// variant with per-row address stubs jumping into one shared body per degree (41 KB) measured 4.26 ms: the indirect branch
// per row costs more than the instructions it saves.)
#include "row_programs.inc"

#ifndef H2_SPEC_ROWS
#define H2_SPEC_ROWS 46
#endif
// The first H2_SPEC_FROM rows (0 or 4: the four degree-19 rows of BG1, 21 % of the unrolled code) run from ONE
// table-driven copy of their code instead: 48 instructions more per row for the addresses, 20 KB less code to fetch -
// measured 2.96 vs 3.05 ms per step (instruction-cache hit rate of the all-unrolled loop: 89 %). The same for the longest
// runs of equal degree further down (rows 28-36, degree 5; rows 16-21, degree 6) measured SLOWER: 3.01 and 3.03 vs 2.98 ms.
#ifndef H2_SPEC_FROM
#define H2_SPEC_FROM 4
#endif

template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f)
{
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f)
{
  static_for_impl(std::make_integer_sequence<int, N>{}, f);
}

// Shared-window address of the soft word edge E of row M reads and writes for thread t.
//   jb  = soft_s + 4 t          (edges at shift 0 relative to the thread: address = jb + immediate)
//   jn4 = 4 (t - Z), "negative" (adding 4 d either stays negative - no wrap, 4 Z is added back - or not: the unsigned
//                                minimum of the two candidates is the wrapped offset)
template <class P, int M, int E>
__device__ __forceinline__ uint32_t spec_edge_addr(uint32_t jb, uint32_t jn4, uint32_t soft_s)
{
  constexpr int      d  = P::DLT[M][E];
  constexpr uint32_t cb = (uint32_t)P::COL[M][E] * (uint32_t)P::Z * 4u;
  if constexpr (d == 0) {
    return jb + cb;
  } else {
    const uint32_t u = jn4 + 4u * (uint32_t)d;
    return __viaddmin_u32(u, 4u * (uint32_t)P::Z, u) + (soft_s + cb);
  }
}

template <class P, int M, int SCALE>
__device__ __forceinline__ void spec_row(uint32_t jb, uint32_t jn4, uint32_t soft_s, RowState& st, uint4* st_thread,
                                         int layers, uint64_t pol, int scale_mode, bool keep)
{
  constexpr int DEG    = P::DEG[M];
  constexpr int STRIDE = (P::Z + 31) & ~31;
  uint32_t      addr[DEG];
  static_for<DEG>([&](auto ec) {
    constexpr int e = decltype(ec)::value;
    addr[e]         = spec_edge_addr<P, M, e>(jb, jn4, soft_s);
  });
  if (!P::NOBAR[M]) {
    // every thread of the CTA is active in a compiled-in shape (Z a multiple of 32) and arrives from the same place: the
    // aligned barrier, without the divergence check the non-aligned form is compiled with
    static_assert(P::Z % 32 == 0, "compiled-in shapes have whole warps");
    asm volatile("barrier.sync.aligned 0;" ::: "memory");
  }
  // Pass 1, fetch of the old messages of the next row in use into the registers this row's were in, pass 2.
  RowPass<DEG> rp;
  row_pass1<DEG>(addr, st, rp);
  if (M + 1 < 4) {
    st = ld_state(st_thread + (M + 1) * STRIDE, pol);
  } else if (M + 1 == P::ROWS) {
    st = ld_state(st_thread, pol);
  } else {
    st = ld_state_sel<(uint32_t)((M + 1) * STRIDE * sizeof(uint4))>(st_thread, M + 1 < layers, pol);
  }
  row_pass2<DEG, SCALE>(addr, rp, st_thread + M * STRIDE, pol, scale_mode, keep);
}

// Rows M .. END-1 of one iteration; stops after the last row in use (at least four rows are always in use).
template <class P, int M, int END, int SCALE>
__device__ __forceinline__ void spec_rows_from(uint32_t jb, uint32_t jn4, uint32_t soft_s, RowState& st,
                                               uint4* st_thread, int layers, uint64_t pol, int scale_mode, bool keep)
{
  if constexpr (M < END) {
    if (M >= 4 && M >= layers) {
      return;
    }
    spec_row<P, M, SCALE>(jb, jn4, soft_s, st, st_thread, layers, pol, scale_mode, keep);
    spec_rows_from<P, M + 1, END, SCALE>(jb, jn4, soft_s, st, st_thread, layers, pol, scale_mode, keep);
  }
}

// ---- hard decisions, "a message soft bit is zero" flags and CRC of both codeblocks ---------------------------------------
//
// acc[0..3] = unreduced CRC remainders {A low, A high, B low, B high}, acc[4] = zero flags (one 16-bit half per
// codeblock). out[2h], out[2h + 1]: where the packed hard decisions of codeblock h go (batch output, HARQ entry), or null.

// 128 consecutive positions per warp step, four per lane (one 128-bit shared load); only for steps whose positions are
// all inside the checked bits of every codeblock present. Eight lanes make one 32-bit word of hard decisions.
template <bool CRC, bool PACK>
__device__ __forceinline__ void sweep128(const hh* soft, const uint2* wgt, int n_steps, int warp, int n_warps, int lane_id,
                                         uint32_t (&acc)[5], uint8_t* const (&out)[4])
{
  const int      l8  = lane_id & 7, grp = lane_id >> 3;
  const uint32_t sh3 = 28u - 4u * (uint32_t)l8;                         // shift of the lane's LAST bit: 31 - (4 l8 + 3)
  // CRC: a lane keeps its four bit positions inside a word for the whole sweep, so the shift by 31 - b is applied once
  // at the end: per bit and step only "accumulator ^= weight & mask", on both codeblocks at once (the weights are stored
  // as {low halves of A and B, high halves of A and B}, the masks are one 16-bit half per codeblock).
  uint32_t lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
  for (int s = warp; s < n_steps; s += n_warps) {
    const uint4 v4 = *reinterpret_cast<const uint4*>(soft + 128 * s + 4 * lane_id);
    const int   t  = 4 * s + grp;
    const hh    sv[4] = {v4.x, v4.y, v4.z, v4.w};
    uint32_t    hm[4];
#pragma unroll
    for (int k = 0; k != 4; ++k) {
      hm[k] = __hle2_mask(H(sv[k]), H(H_ZERO)); // 0xffff per half: hard bit 1 (soft <= 0)
    }
    if (CRC) {
      const uint2 wg = wgt[t];
#pragma unroll
      for (int k = 0; k != 4; ++k) {
        acc[4] |= __heq2_mask(H(sv[k]), H(H_ZERO));
        lo[k] = lop_xor_and(lo[k], wg.x, hm[k]);
        hi[k] = lop_xor_and(hi[k], wg.y, hm[k]);
      }
    }
    if (PACK) {
      // nibbles {A: bits 0-3, B: bits 16-19}, first position in the most significant bit
      const uint32_t one = 0x00010001u;
      const uint32_t nib = ((hm[0] & one) << 3) | ((hm[1] & one) << 2) | ((hm[2] & one) << 1) | (hm[3] & one);
      // Eight lanes -> one word per codeblock with three full-warp exchanges (nibbles -> bytes -> halves -> words), both
      // codeblocks travelling in one register (a group-masked redux.sync makes the four groups of a warp take turns).
      const uint32_t n1  = __shfl_xor_sync(0xffffffffu, nib, 1);
      const uint32_t by  = (l8 & 1) ? ((n1 << 4) | nib) : ((nib << 4) | n1);   // A: bits 0-7, B: bits 16-23
      const uint32_t b2  = __shfl_xor_sync(0xffffffffu, by, 2);
      const uint32_t hf  = (l8 & 2) ? (b2 | (by << 8)) : (by | (b2 << 8));     // A: bits 0-15, B: bits 16-31
      const uint32_t h4  = __shfl_xor_sync(0xffffffffu, hf, 4);
      const uint32_t wlo = (l8 & 4) ? h4 : hf, whi = (l8 & 4) ? hf : h4;       // bytes 0-1 and 2-3 of the word
      const uint32_t wa  = __byte_perm(wlo, whi, 0x5410);
      const uint32_t wb  = __byte_perm(wlo, whi, 0x7632);
      if (l8 == 0) {
        if (out[0] != nullptr) {
          reinterpret_cast<uint32_t*>(out[0])[t] = wa;
          reinterpret_cast<uint32_t*>(out[1])[t] = wa;
        }
        if (out[2] != nullptr) {
          reinterpret_cast<uint32_t*>(out[2])[t] = wb;
          reinterpret_cast<uint32_t*>(out[3])[t] = wb;
        }
      }
    }
  }
  if (CRC) {
#pragma unroll
    for (int k = 0; k != 4; ++k) {
      const uint32_t a  = __byte_perm(lo[k], hi[k], 0x5410), b = __byte_perm(lo[k], hi[k], 0x7632);
      const uint32_t sh = sh3 + 3u - (uint32_t)k;
      acc[0] ^= a << sh;
      acc[1] ^= __funnelshift_l(a, 0u, sh);
      acc[2] ^= b << sh;
      acc[3] ^= __funnelshift_l(b, 0u, sh);
    }
  }
}

// One 32-bit word (lane = bit) with bound tests: the words around the end of the checked bits and of the message.
template <bool CRC>
__device__ __forceinline__ void sweep_word(const hh* soft, const uint2* wgt, int w, int K, int nb0, int nb1, int lane_id,
                                           uint32_t (&acc)[5], uint8_t* const (&out)[4])
{
  const int      i  = 32 * w + lane_id;
  const __half2  sw = H((i < K) ? soft[i] : H_ONE);
  const uint32_t hm = __hle2_mask(sw, H(H_ZERO));
  if (CRC) {
    const uint32_t shl = 31u - (uint32_t)lane_id;
    const uint32_t m0  = (i < nb0) ? __byte_perm(hm, 0, 0x1010) : 0u;
    const uint32_t m1  = (i < nb1) ? __byte_perm(hm, 0, 0x3232) : 0u;
    const uint2    wp  = wgt[w]; // {low halves of A and B, high halves}
    const uint2    wg  = make_uint2(__byte_perm(wp.x, wp.y, 0x5410), __byte_perm(wp.x, wp.y, 0x7632));
    acc[4] |= __heq2_mask(sw, H(H_ZERO));
    acc[0] = lop_xor_and(acc[0], wg.x << shl, m0);
    acc[1] = lop_xor_and(acc[1], __funnelshift_l(wg.x, 0u, shl), m0);
    acc[2] = lop_xor_and(acc[2], wg.y << shl, m1);
    acc[3] = lop_xor_and(acc[3], __funnelshift_l(wg.y, 0u, shl), m1);
  }
  if (out[0] != nullptr || out[2] != nullptr) {
    // Lane k holds bit k of the ballot = bit 31-k of the MSB-first word: reversing the bits of each byte gives the byte
    // string in memory order.
    const uint32_t b0 = __ballot_sync(0xffffffffu, (hm & 0xffffu) != 0);
    const uint32_t b1 = __ballot_sync(0xffffffffu, (hm >> 16) != 0);
    const int      nbytes = (K + 7) / 8;
    uint8_t* const o0 = lane_id ? out[2] : out[0];
    uint8_t* const o1 = lane_id ? out[3] : out[1];
    if (lane_id < 2 && o0 != nullptr) {
      const uint32_t le = __byte_perm(__brev(lane_id ? b1 : b0), 0, 0x0123);
      if (4 * w + 4 <= nbytes) {
        reinterpret_cast<uint32_t*>(o0)[w] = le;
        reinterpret_cast<uint32_t*>(o1)[w] = le;
      } else {
        for (int k = 0; 4 * w + k < nbytes; ++k) {
          o0[4 * w + k] = (uint8_t)(le >> (8 * k));
          o1[4 * w + k] = (uint8_t)(le >> (8 * k));
        }
      }
    }
  }
}

// -DH2_PHASE_TIMING (tools/phase_probe.py): CTA 0 records the global timer at the phase boundaries of its first pair.
#ifdef H2_PHASE_TIMING
__device__ unsigned long long g_h2_phase[16];
#define H2_PHASE(k)                                                                                                    \
  do {                                                                                                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0) {                                                                         \
      unsigned long long t_;                                                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                                           \
      g_h2_phase[k] = t_;                                                                                              \
    }                                                                                                                  \
  } while (0)
#else
#define H2_PHASE(k) ((void)0)
#endif

// Per-codeblock bookkeeping of the pair.
struct LaneInfo {
  int valid; // a codeblock is decoded in this half during the current pass
  int cb;    // index in the batch
  int layers;
  int F, crc_kind, early, max_iter;
  int done;  // 1: early stop reached, 2: all-zero input
};

// Static shared memory of the kernel (one object, see the kernel).
template <int MAX_THREADS>
struct PairShared {
  __align__(16) uint32_t sh_desc[2][8];              // the two descriptors of the pair
  unsigned long long     sh_part[2][MAX_THREADS / 32]; // per-warp unreduced CRC remainders
  uint32_t               sh_red[2][32];              // x^(order + k) mod P of each codeblock's CRC
  LaneInfo               lane[2];
  int                    sh_last[2];
  int                    sh_zpart[MAX_THREADS / 32]; // per-warp "a message soft bit is zero" flags
  int                    sh_publish[2];
  pdc_cb_result          sh_result[2];
  int                    sh_defer_b; // codeblock B could not be decoded together with A: it gets its own pass
  uint32_t               sh_next_pair;
};
// Shared-window address of the dynamic shared memory of a CTA that is not part of a cluster: the driver reserves the
// first KB, the static object follows, the dynamic part starts at the next multiple of 16. With the address a compile-
// time constant the compiled-in rows reach every soft word as [register + immediate]; the kernel compares it with the
// real address and takes the table-driven rows if a future toolkit lays shared memory out differently.
template <int MAX_THREADS>
__host__ __device__ constexpr uint32_t spec_soft_base()
{
  return 0x400u + (((uint32_t)sizeof(PairShared<MAX_THREADS>) + 15u) & ~15u);
}

__host__ __device__ inline size_t h2_smem_bytes(int bg, int Z)
{
  int    n_full = (bg == 1) ? 68 : 52;
  int    kb     = (bg == 1) ? 22 : 10;
  size_t soft   = (size_t)n_full * Z * 4;
  size_t bits   = (size_t)((kb * Z + 31) / 32) * 8 + 16; // CRC word weights of both codeblocks
  return soft + bits + sizeof(GraphSmem) + 64;
}

// SPEC_Z: lifting size whose BG1 row program is compiled in (0: none); pairs of any other shape take the table-driven
// row loop of the same kernel.
// HIGH_RATE: the instantiation for batches of first transmissions whose soft bits end inside the first 24 variable nodes
// (four rows in use - a 273-PRB 256QAM slot): it also holds a compiled-in copy of rows 0-3 for them. (In the general
// instantiation that copy, although never executed by a 46-row codeblock, cost the 8192-codeblock step 3 % through the
// code layout alone: 2.98 vs 2.90 ms; four-row slots lose 7 % without it.)
template <int MAX_THREADS, int MIN_BLOCKS, int SPEC_Z, bool HIGH_RATE = false>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
    ldpc_decode_h2_kernel(BatchParams prm, hh* state_scratch, uint32_t scratch_stride_words, uint32_t* work_counter,
                          uint32_t counter_base)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // All statically allocated shared memory of the kernel is ONE object, so that its size - and with it the shared-window
  // address at which the dynamic part (the soft bits) starts - is known at compile time (SPEC_SOFT_BASE below).
  __shared__ PairShared<MAX_THREADS> ps;
  LaneInfo (&lane)[2]                             = ps.lane;
  int (&sh_last)[2]                               = ps.sh_last;
  unsigned long long (&sh_part)[2][MAX_THREADS / 32] = ps.sh_part;
  int (&sh_zpart)[MAX_THREADS / 32]               = ps.sh_zpart;
  uint32_t (&sh_red)[2][32]                       = ps.sh_red;
  int (&sh_publish)[2]                            = ps.sh_publish;
  pdc_cb_result (&sh_result)[2]                   = ps.sh_result;
  int&      sh_defer_b                            = ps.sh_defer_b;
  uint32_t& sh_next_pair                          = ps.sh_next_pair;
  uint32_t (&sh_desc)[2][8]                       = ps.sh_desc;

  const int tid  = threadIdx.x;
  const int nthr = blockDim.x;
  H2_PHASE(0);

  // Persistent CTAs: pair p = codeblocks 2p and 2p+1 of the batch. They are decoded together when they share the lifted
  // graph, the iteration count and the number of rows in use; otherwise one after the other.
  // The first pair of a CTA is its index; further pairs come from a counter, so that CTAs whose codeblocks stopped early
  // take over work from those still iterating.
  const uint32_t n_pairs = (prm.n_cb + 1) / 2;
  // Shape of the tables currently in shared memory (a CTA usually decodes pairs of one shape back to back).
  int      cur_bg = 0, cur_Z = 0;
  uint32_t cur_wkey0 = 0xffffffffu, cur_wkey1 = 0xffffffffu;
  for (uint32_t pair = blockIdx.x; pair < n_pairs;) {
    for (int pass = 0; pass != 2; ++pass) {
      const uint32_t cb0 = 2 * pair;
      __syncthreads();
      // Both descriptors (7 words each) in one round trip.
      static_assert(sizeof(pdc_cb_desc) == 28, "descriptor layout");
      if (tid < 14) {
        const uint32_t h = tid / 7u, wd = tid % 7u;
        if (cb0 + h < prm.n_cb) {
          sh_desc[h][wd] = __ldg(reinterpret_cast<const uint32_t*>(prm.cbs + cb0 + h) + wd);
        }
      }
      __syncthreads();
      if (tid == 0) {
        const pdc_cb_desc* d[2];
        bool               ok[2];
        for (int h = 0; h != 2; ++h) {
          uint32_t cb = cb0 + h;
          ok[h]       = false;
          d[h]        = nullptr;
          if (cb < prm.n_cb) {
            d[h]  = reinterpret_cast<const pdc_cb_desc*>(sh_desc[h]);
            ok[h] = (d[h]->flags & PDC_CB_DECODE) != 0;
            if (ok[h]) {
              const int bg = d[h]->base_graph, Z = d[h]->lifting_size;
              if ((bg != 1 && bg != 2) || Z < 2 || Z > MAX_Z || c_tab.set_index[Z] == 0xff || d[h]->max_iter == 0 ||
                  d[h]->harq_id >= prm.harq_entries || d[h]->crc_kind > PDC_CRC24B ||
                  (int)d[h]->nof_filler >= ((bg == 1) ? 22 : 10) * Z) {
                if (pass == 0) {
                  pdc_cb_result r;
                  r.crc_ok = 0, r.iters = d[h]->max_iter, r.status = 2, r.nlayers = 0;
                  prm.results[cb] = r;
                }
                ok[h] = false;
              }
            }
          }
        }
        bool compatible = ok[0] && ok[1] && d[0]->base_graph == d[1]->base_graph &&
                          d[0]->lifting_size == d[1]->lifting_size && d[0]->max_iter == d[1]->max_iter;
        if (pass == 0) {
          sh_defer_b = (ok[1] && !compatible) ? 1 : 0;
        }
        for (int h = 0; h != 2; ++h) {
          bool take = (pass == 0) ? (h == 0 ? ok[0] : (ok[1] && compatible)) : (h == 1 && ok[1] && sh_defer_b);
          lane[h].valid  = take;
          lane[h].cb     = cb0 + h;
          lane[h].layers = 0;
          lane[h].done   = 0;
          if (take) {
            lane[h].F        = d[h]->nof_filler;
            lane[h].crc_kind = d[h]->crc_kind;
            lane[h].early    = ((d[h]->flags & PDC_CB_EARLY_STOP) && d[h]->crc_kind != PDC_CRC_NONE) ? 1 : 0;
            lane[h].max_iter = d[h]->max_iter;
          }
          sh_last[h] = 0;
        }
      }
      __syncthreads();
      if (!lane[0].valid && !lane[1].valid) {
        continue;
      }
      const int          lead    = lane[0].valid ? 0 : 1;
      const pdc_cb_desc& dl      = *reinterpret_cast<const pdc_cb_desc*>(sh_desc[lead]);
      const pdc_cb_desc* const dsc[2] = {reinterpret_cast<const pdc_cb_desc*>(sh_desc[0]),
                                         reinterpret_cast<const pdc_cb_desc*>(sh_desc[1])};
      const int          bg      = dl.base_graph;
      const int          Z       = dl.lifting_size;
      const int          b       = bg - 1;
      const int          kb      = (bg == 1) ? 22 : 10;
      const int          n_full  = (bg == 1) ? 68 : 52;
      const int          rows    = (bg == 1) ? 46 : 42;
      const int          K       = kb * Z;
      const int          N       = (n_full - 2) * Z;
      const int          n_words = (K + 31) / 32;
      // the compiled-in row program applies (and the soft bits are where it expects them)
      const bool spec = (SPEC_Z != 0) && bg == 1 && Z == SPEC_Z &&
                        (uint32_t)__cvta_generic_to_shared(smem_raw) == spec_soft_base<MAX_THREADS>();
      PDC_ASSERT((uint32_t)__cvta_generic_to_shared(smem_raw) == spec_soft_base<MAX_THREADS>());

      // Carve shared memory.
      size_t off  = 0;
      hh*    soft = reinterpret_cast<hh*>(smem_raw);
      off += (size_t)n_full * Z * 4;
      uint2* wgt = reinterpret_cast<uint2*>(smem_raw + off); // CRC word weights {low halves of A and B, high halves}
      off += (size_t)n_words * 8;
      off          = (off + 15) & ~(size_t)15;
      GraphSmem& g = *reinterpret_cast<GraphSmem*>(smem_raw + off);

      // Compressed messages: one uint4 per (row, check), private to thread j.
      uint4*         st_base   = reinterpret_cast<uint4*>(state_scratch + (size_t)blockIdx.x * scratch_stride_words);
      const uint32_t st_stride = (uint32_t)((Z + 31) & ~31);
      const uint32_t soft_s    = (uint32_t)__cvta_generic_to_shared(smem_raw);

      // Programmatic dependent launch: this kernel may have started while the rate dematcher of the same batch was still
      // draining. Everything above reads only what was complete before the dematcher started (descriptors); what follows
      // reads what it wrote (soft bits, last non-zero positions).
      H2_PHASE(1);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      H2_PHASE(2);
      // Last non-zero input of each codeblock (ldpc_decoder_impl.cpp:86-99), recorded by the rate dematcher: the loads
      // are issued here and consumed after the tables are in place.
      const int8_t* in[2];
      int4          slots[2];
      for (int h = 0; h != 2; ++h) {
        in[h]    = lane[h].valid ? prm.harq + (size_t)dsc[h]->harq_id * PDC_MAX_CB_SOFT : nullptr;
        slots[h] = lane[h].valid ? __ldg(reinterpret_cast<const int4*>(prm.harq_last) + dsc[h]->harq_id)
                                 : make_int4(0, 0, 0, 0);
      }
      if (H2_NO_REUSE || bg != cur_bg || Z != cur_Z) {
        const uint2* img = g_h2_einfo[b][c_h2_z_slot[Z]];
        for (int i = tid; i < MAX_EDGES + MAX_ROWS; i += nthr) {
          uint2 e = img[i];
          e.y += soft_s; // the image holds column bases relative to the soft-bit array
          g.einfo[i] = e;
        }
        for (int m = tid; m < rows; m += nthr) {
          g.row_info[m] = g_h2_row_info[b][m];
        }
        cur_bg    = bg;
        cur_Z     = Z;
        cur_wkey0 = 0xffffffffu; // the weights live behind the soft bits: a new shape moves them
      }
      // CRC word weights: x^(32 (T-1-t)) mod P for the T words of the K - F checked bits, zero beyond; and
      // x^(order + k) mod P for the final reduction.
      {
        int kind[2], T[2];
        for (int h = 0; h != 2; ++h) {
          const bool with_crc = lane[h].valid && lane[h].crc_kind != PDC_CRC_NONE;
          kind[h]             = with_crc ? lane[h].crc_kind - 1 : 0;
          T[h]                = with_crc ? (K - lane[h].F + 31) / 32 : 0;
        }
        const uint32_t wkey0 = (uint32_t)kind[0] | ((uint32_t)T[0] << 8) | ((uint32_t)n_words << 20);
        const uint32_t wkey1 = (uint32_t)kind[1] | ((uint32_t)T[1] << 8) | ((uint32_t)n_words << 20);
        if (H2_NO_REUSE || wkey0 != cur_wkey0 || wkey1 != cur_wkey1) {
          for (int t = tid; t < n_words; t += nthr) {
            const uint32_t wa = (t < T[0]) ? g_h2_xpow32[kind[0]][T[0] - 1 - t] : 0u;
            const uint32_t wb = (t < T[1]) ? g_h2_xpow32[kind[1]][T[1] - 1 - t] : 0u;
            wgt[t]            = make_uint2(__byte_perm(wa, wb, 0x5410), __byte_perm(wa, wb, 0x7632));
          }
          for (int idx = tid; idx < 64; idx += nthr) {
            const int h = idx >> 5, k = idx & 31;
            uint32_t  r = 0;
            if (T[h] != 0) {
              const uint32_t poly = crc_poly(kind[h] + 1), top = 1u << crc_order(kind[h] + 1);
              r = poly ^ top; // x^order mod P
              for (int i = 0; i != k; ++i) {
                r <<= 1;
                if (r & top) {
                  r ^= poly;
                }
              }
            }
            sh_red[h][k] = r;
          }
          cur_wkey0 = wkey0;
          cur_wkey1 = wkey1;
        }
      }
      H2_PHASE(3);
      // Last non-zero input of each codeblock (ldpc_decoder_impl.cpp:86-99): recorded by the rate dematcher for the
      // entries it wrote; entries of unknown content are scanned.
      const uint64_t pol_stream = l2_policy_evict_first();
#pragma unroll
      for (int h = 0; h != 2; ++h) {
        if (in[h] == nullptr) {
          continue;
        }
        // (a record that describes fewer than N positions, or whose last non-zero soft bit lies beyond this codeblock -
        // the entry was last written for another shape - says nothing exact about [0, N): scan)
        const int known = harq_last_known(slots[h], N);
        if (known >= 0 && known <= N) {
          if (tid == 0) {
            sh_last[h] = known;
          }
          continue;
        }
        int last = 0;
        for (int q = tid; q < ((N + 3) >> 2); q += nthr) {
          const uint32_t w4 = ld_stream_u32(reinterpret_cast<const uint32_t*>(in[h]) + q, pol_stream);
#pragma unroll
          for (int k = 0; k != 4; ++k) {
            if (((w4 >> (8 * k)) & 0xffu) != 0 && 4 * q + k < N) {
              last = max(last, 4 * q + k + 1);
            }
          }
        }
        for (int o = 16; o > 0; o >>= 1) {
          last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        }
        if ((tid & 31) == 0 && last > 0) {
          atomicMax(&sh_last[h], last);
        }
      }
      __syncthreads();
      if (tid == 0) {
        for (int h = 0; h != 2; ++h) {
          if (!lane[h].valid) {
            continue;
          }
          if (sh_last[h] == 0) {
            // All-zero input: not decodable (ldpc_decoder_impl.cpp:88-94).
            pdc_cb_result r;
            r.crc_ok = 0, r.iters = (uint8_t)lane[h].max_iter, r.status = 1, r.nlayers = 0;
            prm.results[lane[h].cb] = r;
            lane[h].valid           = 0;
            lane[h].done            = 2;
            continue;
          }
          int cb_len     = max(sh_last[h] + 2 * Z, K + 4 * Z);
          cb_len         = ((cb_len + Z - 1) / Z) * Z;
          lane[h].layers = cb_len / Z - kb;
        }
        // Different numbers of rows in use: B is decoded alone in the second pass.
        if (lane[0].valid && lane[1].valid && lane[0].layers != lane[1].layers) {
          lane[1].valid = 0;
          sh_defer_b    = 2;
        }
      }
      __syncthreads();
      H2_PHASE(4);
      // load_soft_bits (ldpc_decoder_impl.cpp:149-184): two punctured nodes at zero, whole nodes clamped to +-64; only
      // the variable nodes of the rows in use are needed. int8 -> half through the 1024 binade: 0x6400 | (v + 128) is
      // the half 1024 + v + 128.
      {
        const int n_rows = max(lane[0].valid ? lane[0].layers : 0, lane[1].valid ? lane[1].layers : 0);
        const int n_load = min(N, (kb + n_rows - 2) * Z);
        for (int i = tid; i < 2 * Z; i += nthr) {
          soft[i] = 0;
        }
        const uint32_t* src[2];
        for (int h = 0; h != 2; ++h) {
          src[h] = lane[h].valid ? reinterpret_cast<const uint32_t*>(in[h]) : nullptr;
        }
        const __half2 h64 = H(0x54005400u), hn64 = H(0xD400D400u), hn1152 = H(0xE480E480u);
        // Six words of each codeblock in flight per thread (the loop is latency bound: one CTA, streaming loads; six
        // cover the four-row codeblocks of a high-rate slot in one round trip).
        constexpr int IN_FLIGHT = 6;
        const int     n_q       = (n_load + 3) >> 2;
        for (int q0 = tid; q0 < n_q; q0 += IN_FLIGHT * nthr) {
          uint32_t wa[IN_FLIGHT], wb[IN_FLIGHT];
#pragma unroll
          for (int u = 0; u != IN_FLIGHT; ++u) {
            const int q = q0 + u * nthr;
            wa[u]       = (src[0] && q < n_q) ? ld_stream_u32(src[0] + q, pol_stream) : 0u;
            wb[u]       = (src[1] && q < n_q) ? ld_stream_u32(src[1] + q, pol_stream) : 0u;
          }
#pragma unroll
          for (int u = 0; u != IN_FLIGHT; ++u) {
            const int q = q0 + u * nthr;
            if (q < n_q) {
#pragma unroll
              for (int k = 0; k != 4; ++k) {
                // bytes {A_k, A_k, B_k, B_k} -> halves {0x64 : A_k ^ 0x80, 0x64 : B_k ^ 0x80}
                uint32_t t = __byte_perm(wa[u], wb[u], (uint32_t)(k | (k << 4) | ((4 + k) << 8) | ((4 + k) << 12)));
                uint32_t x;
                asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(x) : "r"(t), "r"(0x00ff00ffu), "r"(0x64806480u));
                __half2 v = __hadd2(H(x), hn1152);
                v         = __hmax2(__hmin2(v, h64), hn64);
                if (4 * q + k < n_load) {
                  int dst = 2 * Z + 4 * q + k;
                  if constexpr (SPEC_Z != 0) {
                    // Extension nodes of the compiled-in rows are stored rotated by the row's thread offset
                    // (tools/gen_row_program.py).
                    const int c = (4 * q) / SPEC_Z + 2;
                    if (spec && c >= kb + 4 && c - kb < H2_SPEC_ROWS) {
                      int pos = 4 * q + k - (c - 2) * SPEC_Z - (int)ROWPROG_1_384_TAU[c - kb];
                      pos += (pos < 0) ? SPEC_Z : 0;
                      dst = c * SPEC_Z + pos;
                    }
                  }
                  PDC_ASSERT(dst >= 2 * Z && dst < n_full * Z && 4 * q + k < PDC_MAX_CB_SOFT);
                  soft[dst] = U(v);
                }
              }
            }
          }
        }
      }
      __syncthreads();
      H2_PHASE(5);
      // All-zero codeblocks without a CRC calculator output all ones.
      for (int h = 0; h != 2; ++h) {
        if (lane[h].done == 2 && lane[h].crc_kind == PDC_CRC_NONE) {
          uint8_t* out = prm.cb_bits + (size_t)lane[h].cb * PDC_MAX_CB_BYTES;
          for (int i = tid; i < (K + 7) / 8; i += nthr) {
            int rem = K - 8 * i;
            out[i]  = (rem >= 8) ? 0xff : (uint8_t)(0xff << (8 - rem));
          }
        }
      }
      if (!lane[0].valid && !lane[1].valid) {
        continue;
      }
      const int  layers     = lane[lane[0].valid ? 0 : 1].layers;
      const int  max_iter   = lane[lead].max_iter;
      const int  scale_mode = prm.scale_mode;
      const int  j          = tid;
      const bool active     = j < Z;

      // Compressed messages: one uint4 per (row, check), only ever touched by thread j. They start at zero ("no message
      // yet") and the next row is fetched while the current one is processed.
      uint4* const   st_thread = st_base + j;
      // every (row, check) entry this thread will touch lies inside the CTA's slice of the scratch
      PDC_ASSERT((size_t)layers * st_stride * 4 <= scratch_stride_words && (!active || (uint32_t)j < st_stride));
      PDC_ASSERT(layers >= 4 && layers <= rows);
      const uint64_t pol_keep  = l2_policy_evict_last();
      if (active) {
        for (int m = 0; m < layers; ++m) {
          st_state(st_thread + (uint32_t)m * st_stride, make_uint4(0, 0, 0, 0), pol_keep);
        }
      }
      const uint32_t j4 = 4u * (uint32_t)j, neg_Z4 = 0u - 4u * (uint32_t)Z;
      const uint32_t einfo_s    = (uint32_t)__cvta_generic_to_shared(g.einfo);
      const uint32_t row_info_s = (uint32_t)__cvta_generic_to_shared(g.row_info);
      RowState       st         = make_uint4(0, 0, 0, 0);
      const uint32_t jb = soft_s + j4, jn4 = j4 - 4u * (uint32_t)Z;
      H2_PHASE(6);
      for (int it = 0; it < max_iter; ++it) {
        if (it == 1) {
          H2_PHASE(7);
        }
        uint4* sp = st_thread;
        int    m0 = 0;
        bool   table_rows = true;
        // The messages a row stores are read by the next iteration: the last one need not store them (a sixth of the
        // state writes of a six-iteration decode, and dirty lines nobody will read).
        const bool keep_state = (it + 1 != max_iter);
        if constexpr (SPEC_Z != 0) {
          if (spec) {
            // The leading rows of the hot shape run from the compiled-in program, the rest from the tables.
            // (the kernel with a compiled-in program is only launched for the x86 scale rule)
            constexpr uint32_t soft_c = spec_soft_base<MAX_THREADS>(); // == soft_s (checked above)
#if H2_SPEC_FROM > 0
            if (!HIGH_RATE || layers != H2_SPEC_FROM) {
              // The first rows (all of one degree in BG1, always in use) from ONE copy of the code, table-driven addresses.
#pragma unroll 1
              for (int m = 0; m != H2_SPEC_FROM; ++m) {
                // (a codeblock that uses only these rows wraps around to row 0 behind the last of them)
                const uint4* spn = (m + 1 < layers) ? st_thread + (uint32_t)(m + 1) * st_stride : st_thread;
                process_row<19>(einfo_s + 8u * 20u * (uint32_t)m, j4, neg_Z4, st, st_thread + (uint32_t)m * st_stride, spn,
                                pol_keep, PDC_SCALE_X86, true, keep_state);
              }
              spec_rows_from<RowProgram<1, (SPEC_Z != 0 ? SPEC_Z : 384)>, H2_SPEC_FROM, H2_SPEC_ROWS, PDC_SCALE_X86>(
                  soft_c + j4, jn4, soft_c, st, st_thread, layers, pol_keep, scale_mode, keep_state);
            } else {
              // A high-rate codeblock uses these rows only: their compiled-in copy (a loop that small fits the
              // instruction cache; the table-driven copy above costs it 48 instructions per row).
              spec_rows_from<RowProgram<1, (SPEC_Z != 0 ? SPEC_Z : 384)>, 0, H2_SPEC_FROM, PDC_SCALE_X86>(
                  soft_c + j4, jn4, soft_c, st, st_thread, layers, pol_keep, scale_mode, keep_state);
            }
#else
            spec_rows_from<RowProgram<1, (SPEC_Z != 0 ? SPEC_Z : 384)>, H2_SPEC_FROM, H2_SPEC_ROWS, PDC_SCALE_X86>(
                soft_c + j4, jn4, soft_c, st, st_thread, layers, pol_keep, scale_mode, keep_state);
#endif
            if constexpr (H2_SPEC_ROWS >= 46) {
              table_rows = false; // the whole schedule is compiled in: nothing left for the table-driven loop
            } else {
              m0 = min(layers, H2_SPEC_ROWS);
              sp = (m0 < layers) ? st_thread + (uint32_t)m0 * st_stride : st_thread;
            }
          }
        }
        for (int m = m0; table_rows && m < layers; ++m) {
          if (active) {
            const uint32_t info = lds_u32(row_info_s + 4u * (uint32_t)m);
            const int      e0   = info & 0xffffu;
            const int      deg  = (info >> 16) & 0x7fff;
            uint4* const   spn  = (m + 1 < layers) ? sp + st_stride : st_thread;
            dispatch_row(deg, einfo_s + 8u * (uint32_t)e0, j4, neg_Z4, st, sp, spn, pol_keep, scale_mode,
                         (info >> 31) == 0);
            sp = spn;
          } else if ((g.row_info[m] >> 31) == 0) {
            row_barrier();
          }
        }

        if (it + 1 == max_iter) {
          H2_PHASE(8);
        }
        const bool last_it   = (it + 1 == max_iter);
        const bool any_early = (lane[0].valid && lane[0].early && !lane[0].done) ||
                               (lane[1].valid && lane[1].early && !lane[1].done);
        if (any_early || last_it) {
          row_barrier(); // the last row is complete
          // get_hard_bits (:126-134) of both codeblocks in one sweep: bit = soft <= 0; a zero among the K message soft
          // bits blocks the early stop. The CRC is computed in the same sweep as M(x) mod P == 0 with
          // M(x) = sum_t W_t(x) x^(32 (T-1-t)): the lane holding bit b of word t adds the UNREDUCED product
          // wgt[t] x^(31-b); the sum is reduced modulo P once, one bit per lane, by the first warp.
          // On the last iteration every codeblock still active is published whatever its CRC says, so the same sweep
          // writes its packed hard decisions; a codeblock that stops earlier gets a pack-only sweep below.
          const int lane_id = tid & 31, warp = tid >> 5, n_warps = nthr >> 5;
          {
            const int nb0 = lane[0].valid ? K - lane[0].F : 0, nb1 = lane[1].valid ? K - lane[1].F : 0;
            // Positions inside the checked bits of every codeblock present need no bound test: whole 128-position steps
            // (four words) of them take the vector path, the rest goes word by word.
            const int n_steps = min(lane[0].valid ? nb0 : K, lane[1].valid ? nb1 : K) >> 7;
            uint32_t  acc[5]  = {0, 0, 0, 0, 0}; // a0l, a0h, a1l, a1h, zero flags
            uint8_t*  out[4]  = {nullptr, nullptr, nullptr, nullptr};
            if (last_it) {
              for (int h = 0; h != 2; ++h) {
                if (lane[h].valid && !lane[h].done) {
                  out[2 * h]     = prm.cb_bits + (size_t)lane[h].cb * PDC_MAX_CB_BYTES;
                  out[2 * h + 1] = prm.harq_data + (size_t)dsc[h]->harq_id * PDC_MAX_CB_BYTES;
                }
              }
              sweep128<true, true>(soft, wgt, n_steps, warp, n_warps, lane_id, acc, out);
            } else {
              sweep128<true, false>(soft, wgt, n_steps, warp, n_warps, lane_id, acc, out);
            }
            for (int w = 4 * n_steps + warp; w < n_words; w += n_warps) {
              sweep_word<true>(soft, wgt, w, K, nb0, nb1, lane_id, acc, out);
            }
            // Warp-wide XOR / OR reductions in one instruction each (redux.sync).
            const int za = __any_sync(0xffffffffu, (acc[4] & 0xffffu) != 0);
            const int zb = __any_sync(0xffffffffu, (acc[4] >> 16) != 0);
#pragma unroll
            for (int k = 0; k != 4; ++k) {
              acc[k] = __reduce_xor_sync(0xffffffffu, acc[k]);
            }
            if (lane_id == 0) {
              sh_part[0][warp] = ((uint64_t)acc[1] << 32) | acc[0];
              sh_part[1][warp] = ((uint64_t)acc[3] << 32) | acc[2];
              sh_zpart[warp]   = (za ? 1 : 0) | (zb ? 2 : 0);
            }
            __syncthreads();
            if (warp == 0) {
              const uint64_t p0 = (lane_id < n_warps) ? sh_part[0][lane_id] : 0ull;
              const uint64_t p1 = (lane_id < n_warps) ? sh_part[1][lane_id] : 0ull;
              const uint64_t v0 = ((uint64_t)__reduce_xor_sync(0xffffffffu, (uint32_t)(p0 >> 32)) << 32) |
                                  __reduce_xor_sync(0xffffffffu, (uint32_t)p0);
              const uint64_t v1 = ((uint64_t)__reduce_xor_sync(0xffffffffu, (uint32_t)(p1 >> 32)) << 32) |
                                  __reduce_xor_sync(0xffffffffu, (uint32_t)p1);
              const int zz = (int)__reduce_or_sync(0xffffffffu, (uint32_t)((lane_id < n_warps) ? sh_zpart[lane_id] : 0));
              for (int h = 0; h != 2; ++h) {
                const uint64_t v     = h ? v1 : v0;
                const int      kind  = lane[h].crc_kind;
                const int      order = crc_order(kind);
                const uint32_t part  =
                    __reduce_xor_sync(0xffffffffu, ((v >> (order + lane_id)) & 1ull) ? sh_red[h][lane_id] : 0u);
                const uint32_t crc = part ^ ((uint32_t)v & ((1u << order) - 1u));
                if (lane_id == 0) {
                  const bool active   = lane[h].valid && !lane[h].done && (lane[h].early || last_it);
                  const bool pass_crc = (kind != PDC_CRC_NONE) && (crc == 0);
                  const bool stop     = lane[h].early && pass_crc && !((zz >> h) & 1);
                  sh_publish[h]       = (active && (stop || last_it)) ? 1 : 0;
                  pdc_cb_result r;
                  r.crc_ok     = lane[h].early ? (stop ? 1 : 0) : (pass_crc ? 1 : 0);
                  r.iters      = (uint8_t)(stop ? it + 1 : lane[h].max_iter);
                  r.status     = 0;
                  r.nlayers    = (uint8_t)lane[h].layers;
                  sh_result[h] = r;
                  if (active && stop) {
                    lane[h].done = 1;
                  }
                }
              }
            }
            __syncthreads();
          }
          // Publish finished codeblocks. On the last iteration their hard decisions are already out; a codeblock that
          // stopped before gets a pack-only sweep now.
          if (sh_publish[0] || sh_publish[1]) {
            if (!last_it) {
              uint8_t* out[4] = {nullptr, nullptr, nullptr, nullptr};
              for (int h = 0; h != 2; ++h) {
                if (sh_publish[h]) {
                  out[2 * h]     = prm.cb_bits + (size_t)lane[h].cb * PDC_MAX_CB_BYTES;
                  out[2 * h + 1] = prm.harq_data + (size_t)dsc[h]->harq_id * PDC_MAX_CB_BYTES;
                }
              }
              uint32_t  acc[5] = {0, 0, 0, 0, 0};
              const int n_steps = K >> 7;
              sweep128<false, true>(soft, wgt, n_steps, warp, n_warps, lane_id, acc, out);
              for (int w = 4 * n_steps + warp; w < n_words; w += n_warps) {
                sweep_word<false>(soft, wgt, w, K, 0, 0, lane_id, acc, out);
              }
            }
            if (tid < 2 && sh_publish[tid]) {
              prm.results[lane[tid].cb] = sh_result[tid];
            }
          }
          const bool all_done = (!lane[0].valid || lane[0].done) && (!lane[1].valid || lane[1].done);
          if (all_done) {
            break;
          }
        }
      }
    }
    __syncthreads();
    H2_PHASE(9);
    if (tid == 0) {
      // The counter is never reset: every decoded pair takes exactly one ticket, so a launch advances it by its number
      // of pairs and the host passes the value it had before the launch.
      sh_next_pair = gridDim.x + (atomicAdd(work_counter, 1u) - counter_base);
    }
    __syncthreads();
    pair = sh_next_pair;
  }
  // A CTA that never reached the wait above (only invalid codeblocks) must not let the grid complete ahead of the
  // kernel it was serialized behind: the next kernel's wait relies on this one for the ordering.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  H2_PHASE(10);
}

} // namespace h2

// Launch plan of the throughput kernel for a batch whose largest lifting size is max_Z.
struct H2Plan {
  int    threads;
  int    grid;
  size_t smem;
  size_t scratch_words_per_cta;
  bool   big;  // the 384-thread instantiation (two CTAs per SM)
  bool   spec; // ... with the BG1 Z = 384 row program compiled in
  bool   high_rate; // ... and the instantiation for four-row codeblocks (see the kernel)
};

// PDC_NO_SPEC=1: table-driven row loop for every shape (A/B measurements, tests of the general path).
inline bool h2_no_spec()
{
  static const bool v = [] {
    const char* e = getenv("PDC_NO_SPEC");
    return e != nullptr && e[0] == '1';
  }();
  return v;
}

// Builds the per-(base graph, lifting size) images the kernel copies (after upload_tables, once per context).
inline cudaError_t upload_h2_images()
{
  const BgTables& h = host_tables();
  static uint2    einfo[2][h2::H2_Z_SLOTS][MAX_EDGES + MAX_ROWS];
  static uint32_t row_info[2][MAX_ROWS];
  static uint8_t  z_slot[MAX_Z + 1];
  static_assert(NR_LDPC_NOF_LIFTING_SIZES == h2::H2_Z_SLOTS, "lifting size table");
  memset(einfo, 0, sizeof(einfo));
  memset(row_info, 0, sizeof(row_info));
  memset(z_slot, 0, sizeof(z_slot));
  for (int bg = 0; bg != 2; ++bg) {
    const int rows = bg ? 42 : 46, n_edges = bg ? BG2_NOF_EDGES : BG1_NOF_EDGES;
    for (int zi = 0; zi != h2::H2_Z_SLOTS; ++zi) {
      const int Z = NR_LDPC_LIFTING_SIZES[zi], set = NR_LDPC_SET_INDEX[zi];
      z_slot[Z]   = (uint8_t)zi;
      for (int i = 0; i != n_edges; ++i) {
        const int m  = h.row[bg][i];
        const int pi = h.row_pstart[bg][m] + (i - h.row_start[bg][m]);
        // {circulant shift in bytes (ldpc_luts_impl.cpp:4536-4541), column base relative to the soft-bit array}
        einfo[bg][zi][pi] = make_uint2((uint32_t)(4 * (h.v[bg][set][i] % Z)), (uint32_t)(h.col[bg][i] * Z * 4));
      }
    }
    for (int m = 0; m != rows; ++m) {
      row_info[bg][m] = ((uint32_t)h.row_free[bg][m] << 31) | (uint32_t)h.row_pstart[bg][m] |
                        ((uint32_t)(h.row_start[bg][m + 1] - h.row_start[bg][m]) << 16);
    }
  }
  cudaError_t e = cudaMemcpyToSymbol(h2::g_h2_einfo, einfo, sizeof(einfo));
  if (e == cudaSuccess) {
    e = cudaMemcpyToSymbol(h2::g_h2_row_info, row_info, sizeof(row_info));
  }
  if (e == cudaSuccess) {
    e = cudaMemcpyToSymbol(h2::c_h2_z_slot, z_slot, sizeof(z_slot));
  }
  if (e == cudaSuccess) {
    e = cudaMemcpyToSymbol(h2::g_h2_xpow32, h.xpow32, sizeof(h.xpow32));
  }
  return e;
}

typedef void (*h2_kernel_t)(BatchParams, uint32_t*, uint32_t, uint32_t*, uint32_t);

// The dynamic shared-memory limit is an attribute of a kernel PER DEVICE: set once per context (pdc_create, with the
// context's device current) to the largest size any batch can ask for.
inline cudaError_t h2_configure_device()
{
  const int   big = (int)h2::h2_smem_bytes(1, 384), small = (int)h2::h2_smem_bytes(1, 192);
  cudaError_t e   = cudaFuncSetAttribute((h2_kernel_t)h2::ldpc_decode_h2_kernel<384, 2, 384>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  if (e == cudaSuccess) {
    e = cudaFuncSetAttribute((h2_kernel_t)h2::ldpc_decode_h2_kernel<384, 2, 384, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  }
  if (e == cudaSuccess) {
    e = cudaFuncSetAttribute((h2_kernel_t)h2::ldpc_decode_h2_kernel<384, 2, 0>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  }
  if (e == cudaSuccess) {
    e = cudaFuncSetAttribute((h2_kernel_t)h2::ldpc_decode_h2_kernel<192, 4, 0>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, small);
  }
  return e;
}

inline cudaError_t h2_plan(int max_Z, bool any_bg1, uint32_t n_cb, int sm_count, H2Plan& plan,
                           int scale_mode = PDC_SCALE_X86, bool high_rate = false)
{
  const int bg               = any_bg1 ? 1 : 2;
  plan.threads               = ((max_Z + 31) / 32) * 32;
  plan.smem                  = h2::h2_smem_bytes(bg, max_Z);
  const int rows             = (bg == 1) ? 46 : 42;
  plan.scratch_words_per_cta = (size_t)rows * 4 * ((max_Z + 31) & ~31);
  plan.big                   = plan.threads > 192;
  plan.spec                  = plan.big && any_bg1 && max_Z == 384 && scale_mode == PDC_SCALE_X86 && !h2_no_spec();
  plan.high_rate             = plan.spec && high_rate;
  h2_kernel_t   k            = plan.high_rate ? (h2_kernel_t)h2::ldpc_decode_h2_kernel<384, 2, 384, true>
                               : plan.spec    ? (h2_kernel_t)h2::ldpc_decode_h2_kernel<384, 2, 384>
                               : plan.big ? (h2_kernel_t)h2::ldpc_decode_h2_kernel<384, 2, 0>
                                          : (h2_kernel_t)h2::ldpc_decode_h2_kernel<192, 4, 0>;
  int         per_sm = 0;
  cudaError_t e      = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, plan.threads, plan.smem);
  if (e != cudaSuccess) {
    return e;
  }
  if (per_sm < 1) {
    per_sm = 1;
  }
  const uint32_t n_pairs = (n_cb + 1) / 2;
  plan.grid              = (int)std::min<uint32_t>(n_pairs, (uint32_t)(sm_count * per_sm));
  return cudaSuccess;
}

// work_counter: a uint32 that only ever grows; counter_base = its value before this launch (the launch adds one per
// pair). The kernel is launched with programmatic stream serialization: it may begin while the previous kernel of the
// stream (the rate dematcher) drains, and waits (griddepcontrol.wait) before touching what that kernel wrote.
inline cudaError_t launch_ldpc_decode_h2(const BatchParams& p, const H2Plan& plan, uint32_t* scratch,
                                         uint32_t* work_counter, uint32_t counter_base, cudaStream_t s)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = dim3((unsigned)plan.grid);
  cfg.blockDim           = dim3((unsigned)plan.threads);
  cfg.dynamicSmemBytes   = plan.smem;
  cfg.stream             = s;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                          = attr;
  cfg.numAttrs                                       = 1;
  const uint32_t stride = (uint32_t)plan.scratch_words_per_cta;
  if (plan.high_rate) {
    return cudaLaunchKernelEx(&cfg, h2::ldpc_decode_h2_kernel<384, 2, 384, true>, p, scratch, stride, work_counter,
                              counter_base);
  }
  if (plan.spec) {
    return cudaLaunchKernelEx(&cfg, h2::ldpc_decode_h2_kernel<384, 2, 384>, p, scratch, stride, work_counter,
                              counter_base);
  }
  if (plan.big) {
    return cudaLaunchKernelEx(&cfg, h2::ldpc_decode_h2_kernel<384, 2, 0>, p, scratch, stride, work_counter,
                              counter_base);
  }
  return cudaLaunchKernelEx(&cfg, h2::ldpc_decode_h2_kernel<192, 4, 0>, p, scratch, stride, work_counter, counter_base);
}

} // namespace pdc
