// Rate dematching + HARQ soft combining (HBM-bound gather kernel).
//
// Reference behaviour: ldpc_rate_dematcher_impl::rate_dematch / allot_llrs / deinterleave_llrs
// (lib/phy/upper/channel_coding/ldpc/ldpc_rate_dematcher_impl.cpp:46-114, :128-201, :203-257) and the SIMD
// combine_softbits of ldpc_rate_dematcher_avx512_impl.cpp:29-64.
//
// The reference walks the circular buffer sequentially. Here every output position computes its own final value:
//   value(p) = base(p) (+) in[i0] (+) in[i0 + D] (+) in[i0 + 2D] ...      D = Ncb - F data positions per lap
// with the first term copied instead of combined on the first lap of a new transmission. All the side effects of the
// sequential walk (zeroed head, filler fill, final zeroing relative to the END of the N-long buffer, regions left stale)
// are reproduced through base(p). Saturating combination is not associative, so the laps are applied in order.
#pragma once

#include "pdc_device.cuh"

namespace pdc {

// Kernel geometry. The rate-matched input of a codeblock is first staged in shared memory in DEINTERLEAVED order
// (coalesced 64/128-bit global loads, bit planes separated with byte permutes), so the gather that follows reads
// consecutive bytes. Inputs longer than the staging buffer are gathered from global memory instead.
constexpr int DM_THREADS     = 256;
constexpr int DM_STAGE_BYTES = 25600; // one lap of the largest codeblock; longer inputs take the unstaged path (eight CTAs per SM matter more)
constexpr int DM_MAX_PARTS   = 4; // CTAs per codeblock (small batches are latency bound: split the N positions)

struct DematchGeom {
  int N, Ncb, Z, K_sys, F, info, k0, d0, Dn, E, qm, Kq;
  float inv_Kq;
  int new_data;
  int first_pass_len; // data positions from the start of the walk to the end of the circular buffer
  int wrapped;        // the walk went past the end of the circular buffer at least once
  int zlo;            // new data: [0, zlo) is zeroed
  int zf_lo;          // new data, not wrapped: [zf_lo, N) is zeroed at the end; N if not applicable
  int simd_width;
  int staged;         // the deinterleaved input is staged in shared memory
  const uint32_t* seq; // deferred descrambling: scrambling sequence of the codeword (nullptr: the input is descrambled)
  uint32_t seq_base;  // sequence element of the codeblock's first rate-matched soft bit
};

__device__ __forceinline__ bool dm_geometry(const pdc_cb_desc& d, int simd_width, DematchGeom& g)
{
  int bg = d.base_graph;
  int Z  = d.lifting_size;
  if ((bg != 1 && bg != 2) || Z < 2 || Z > MAX_Z || c_tab.set_index[Z] == 0xff || d.rv > 3 || d.qm == 0 ||
      d.rm_length == 0 || (d.rm_length % d.qm) != 0) {
    return false;
  }
  int n_short = (bg == 1) ? 66 : 50;
  int kb      = (bg == 1) ? 22 : 10;
  g.Z         = Z;
  g.N         = n_short * Z;
  g.Ncb       = (d.nref > 0) ? min((int)d.nref, g.N) : g.N;
  g.K_sys     = (kb - 2) * Z;
  g.F         = d.nof_filler;
  g.info      = g.K_sys - g.F;
  if (g.F >= g.K_sys || g.Ncb <= g.K_sys) {
    return false;
  }
  const int sf1[4] = {0, 17, 33, 56};
  const int sf2[4] = {0, 13, 25, 43};
  int       sf     = (bg == 1) ? sf1[d.rv] : sf2[d.rv];
  g.k0             = ((sf * g.Ncb) / g.N) * Z; // ldpc_rate_dematcher_impl.cpp:104-105
  g.Dn             = g.Ncb - g.F;
  g.d0             = (g.k0 < g.info) ? g.k0 : ((g.k0 < g.K_sys) ? g.info : g.k0 - g.F);
  g.E              = d.rm_length;
  g.qm             = d.qm;
  g.Kq             = g.E / g.qm;
  g.inv_Kq         = 1.0f / (float)g.Kq;
  g.new_data       = (d.flags & PDC_CB_NEW_DATA) ? 1 : 0;
  g.first_pass_len = g.Dn - g.d0;
  g.wrapped        = g.E > g.first_pass_len;
  g.zlo            = (g.k0 < g.info) ? g.k0 : g.info;
  g.zf_lo          = g.N;
  g.simd_width     = simd_width;
  g.staged = (g.qm == 1 || g.qm == 2 || g.qm == 4 || g.qm == 6 || g.qm == 8) && (g.E <= DM_STAGE_BYTES);
  if (g.new_data && !g.wrapped) {
    int de      = g.d0 + g.E;
    int idx_end = (g.K_sys + max(0, de - g.info)) % g.Ncb;
    if (idx_end != 0) {
      g.zf_lo = g.N - (g.Ncb - idx_end);
    }
  }
  return true;
}

// Is the combine at position p, fed by deinterleaved input index i, inside the SIMD part of its chunk?
// (Only evaluated for non-finite operands, where the SIMD and scalar rules of the reference differ.)
__device__ __noinline__ bool dm_in_simd_block(const DematchGeom& g, int p, int i)
{
  if (g.simd_width <= 0) {
    return false;
  }
  int w = (g.d0 + i) / g.Dn; // lap
  int start_pos, i_start;
  if (p < g.info) {
    start_pos = (w == 0) ? g.k0 : 0;
    i_start   = (w == 0) ? 0 : w * g.Dn - g.d0;
    int len   = min(g.info - start_pos, g.E - i_start);
    return (p - start_pos) < (len / g.simd_width) * g.simd_width;
  }
  start_pos = (w == 0 && g.k0 > g.K_sys) ? g.k0 : g.K_sys;
  i_start   = (w == 0) ? ((g.k0 < g.info) ? g.info - g.k0 : 0) : w * g.Dn - g.d0 + g.info;
  int len   = min(g.Ncb - start_pos, g.E - i_start);
  return (p - start_pos) < (len / g.simd_width) * g.simd_width;
}

// out (+) in with the reference's semantics (log_likelihood_ratio::operator+, LLR.cpp:40-72; "a + b" is "b += a").
__device__ __forceinline__ int dm_combine(const DematchGeom& g, int p, int i, int a, int b)
{
  bool a_fin = (a >= -LLR_MAX) && (a <= LLR_MAX);
  bool b_fin = (b >= -LLR_MAX) && (b <= LLR_MAX);
  if (a_fin && b_fin) {
    return max(-LLR_MAX, min(LLR_MAX, a + b));
  }
  if (dm_in_simd_block(g, p, i)) {
    int s = max(-128, min(127, a + b));
    return max(-LLR_MAX, min(LLR_MAX, s));
  }
  if (b == (int)(int8_t)(-a)) {
    return 0;
  }
  if (!b_fin) {
    return b;
  }
  return a; // a is the non-finite one
}

// Deinterleaved input element i (ldpc_rate_dematcher_impl.cpp:203-257). STAGED: the deinterleaved sequence is in sh[].
template <bool STAGED>
__device__ __forceinline__ int dm_fetch(const DematchGeom& g, const int8_t* __restrict__ llr, const uint8_t* sh, int i)
{
  PDC_ASSERT(i >= 0 && i < g.E);
  if (STAGED) {
    PDC_ASSERT(i < DM_STAGE_BYTES);
    return (int)(int8_t)sh[i];
  }
  int k = i;
  if (g.qm != 1) {
    int j   = i / g.Kq; // bit plane
    int sym = i - j * g.Kq;
    k       = sym * g.qm + j;
  }
  PDC_ASSERT(k >= 0 && k < g.E);
  int v = (int)__ldg(llr + k);
  if (g.seq != nullptr && seq_bit(g.seq, g.seq_base + (uint32_t)k)) {
    v = (int)(int8_t)(uint8_t)(0u - (uint32_t)v);
  }
  return v;
}

template <bool STAGED>
__device__ __forceinline__ int dm_position(const DematchGeom& g, const int8_t* __restrict__ llr, const uint8_t* sh,
                                           int p, int old)
{
  if (p >= g.N) {
    return old;
  }
  int val = old;
  if (g.new_data) {
    if (p >= g.info && p < g.K_sys) {
      return LLR_INF;
    }
    if (p < g.zlo || p >= g.zf_lo) {
      val = 0;
    }
  } else if (p >= g.info && p < g.K_sys) {
    return old;
  }
  if (p >= g.Ncb) {
    return val;
  }
  int d  = (p < g.info) ? p : p - g.F;
  int i  = d - g.d0;
  bool first_lap = (i >= 0);
  if (i < 0) {
    i += g.Dn;
  }
  if (i >= g.E) {
    return val;
  }
  if (g.new_data && first_lap) {
    val = dm_fetch<STAGED>(g, llr, sh, i);
    i += g.Dn;
  }
  for (; i < g.E; i += g.Dn) {
    val = dm_combine(g, p, i, val, dm_fetch<STAGED>(g, llr, sh, i));
  }
  return val;
}

// Position of the last non-zero soft bit seen by a thread: the word with the highest index and its value.
struct DmLast {
  int      w   = -1;
  uint32_t val = 0;
  __device__ __forceinline__ void note(int word, uint32_t v)
  {
    if (v != 0 && word > w) {
      w   = word;
      val = v;
    }
  }
  __device__ __forceinline__ int position() const // 1 + index of the last non-zero byte, 0 if none
  {
    return (w < 0) ? 0 : 4 * w + 4 - (__clz((int)val) >> 3);
  }
};

// General path: the four positions of word w one by one.
template <bool STAGED>
__device__ __noinline__ uint32_t dm_word_general(const DematchGeom& g, const int8_t* __restrict__ llr, const uint8_t* sh,
                                                 uint32_t* out, int w)
{
  PDC_ASSERT(w >= 0 && 4 * w < g.N && 4 * w + 4 <= PDC_MAX_CB_SOFT);
  const uint32_t old = out[w];
  uint32_t       res = 0;
#pragma unroll
  for (int k = 0; k != 4; ++k) {
    int o = (int)(int8_t)(old >> (8 * k));
    int v = dm_position<STAGED>(g, llr, sh, 4 * w + k, o);
    res |= (uint32_t)(uint8_t)(int8_t)v << (8 * k);
  }
  if (res != old) {
    out[w] = res;
  }
  // Bytes beyond N keep their old value and do not count as soft bits.
  if (4 * w + 3 >= g.N) {
    res &= 0xffffffffu >> (8 * (4 * w + 4 - g.N));
  }
  return res;
}

// True if any of the four packed int8 is outside [-120, 120] (|u - 0x80| < 8 for the unsigned byte u).
__device__ __forceinline__ bool dm_any_nonfinite(uint32_t x)
{
  uint32_t v = __vabsdiffu4(x, 0x80808080u);
  return ((v - 0x08080808u) & ~v & 0x80808080u) != 0;
}

// Four saturating LLR sums at once for finite operands: clamp(a + b, +-120) per int8, computed in 16-bit lanes.
__device__ __forceinline__ uint32_t dm_combine4_finite(uint32_t a, uint32_t b)
{
  // Sign-extend even / odd bytes to 16-bit lanes (PRMT selector bit 3 replicates the sign of the selected byte; the
  // __byte_perm intrinsic only defines selector values 0-7, hence the explicit PTX).
  uint32_t ae, ao, be, bo;
  asm("prmt.b32 %0, %1, 0, 0xA280;" : "=r"(ae) : "r"(a));
  asm("prmt.b32 %0, %1, 0, 0xB391;" : "=r"(ao) : "r"(a));
  asm("prmt.b32 %0, %1, 0, 0xA280;" : "=r"(be) : "r"(b));
  asm("prmt.b32 %0, %1, 0, 0xB391;" : "=r"(bo) : "r"(b));
  uint32_t se = __vmaxs2(__vmins2(__vadd2(ae, be), 0x00780078u), 0xff88ff88u);
  uint32_t so = __vmaxs2(__vmins2(__vadd2(ao, bo), 0x00780078u), 0xff88ff88u);
  return __byte_perm(se, so, 0x6240);
}

// ---- single-lap transmissions: the N positions fall into a few segments of uniform behaviour ------------------------
//
// Between two consecutive breakpoints (region limits of the circular buffer, start and end of the walk) every position
// is treated alike and is fed by the deinterleaved input at a constant distance, i = p + ioff.
enum DmAction : int {
  DM_KEEP = 0,  // stale: the walk does not touch the position
  DM_ZERO,      // zeroed by a new transmission
  DM_FILL,      // filler bits of a new transmission: +infinity
  DM_COPY,      // first lap of a new transmission
  DM_COMB_ZERO, // combined with a zeroed position
  DM_COMB       // combined with the stored soft bit
};
struct DmSeg {
  int p0, p1, action, ioff;
};
constexpr int DM_MAX_BP = 12;

__device__ __forceinline__ int dm_classify(const DematchGeom& g, int p, int& ioff)
{
  ioff = 0;
  if (p >= g.info && p < g.K_sys) {
    return g.new_data ? DM_FILL : DM_KEEP;
  }
  const bool base_zero = g.new_data && (p < g.zlo || p >= g.zf_lo);
  const int  idle      = base_zero ? DM_ZERO : DM_KEEP;
  if (p >= g.Ncb) {
    return idle;
  }
  int        d         = (p < g.info) ? p : p - g.F;
  int        i         = d - g.d0;
  const bool first_lap = i >= 0;
  if (!first_lap) {
    i += g.Dn;
  }
  if (i >= g.E) {
    return idle;
  }
  ioff = i - p;
  if (g.new_data && first_lap) {
    return DM_COPY;
  }
  return base_zero ? DM_COMB_ZERO : DM_COMB;
}

__device__ __forceinline__ int dm_breakpoint(const DematchGeom& g, int k)
{
  auto pos_of = [&](int d) { return (d < g.info) ? d : d + g.F; };
  int  v;
  switch (k) {
    case 0:
      v = 0;
      break;
    case 1:
      v = g.zlo;
      break;
    case 2:
      v = g.zf_lo;
      break;
    case 3:
      v = g.info;
      break;
    case 4:
      v = g.K_sys;
      break;
    case 5:
      v = g.Ncb;
      break;
    case 6:
      v = pos_of(g.d0);
      break;
    case 7:
      v = g.wrapped ? pos_of(g.E - g.first_pass_len) : pos_of(g.d0 + g.E);
      break;
    default:
      v = g.N;
      break;
  }
  return max(0, min(v, g.N));
}

__device__ __forceinline__ uint32_t dm_lds_u32_unaligned(const uint8_t* sh, int a)
{
  const uint32_t* w = reinterpret_cast<const uint32_t*>(sh + (a & ~3));
  return __funnelshift_r(w[0], w[1], 8 * (a & 3));
}

// Words [wa, wb) of one segment, all four positions of each inside the segment.
__device__ __forceinline__ void dm_segment_words(const DematchGeom& g, const int8_t* __restrict__ llr, const uint8_t* sh,
                                                 uint32_t* out, int action, int ioff, int wa, int wb, DmLast& last)
{
  PDC_ASSERT(wa >= 0 && wa <= wb && 4 * wb <= g.N + 3 && 4 * wb <= PDC_MAX_CB_SOFT);
  const int tid = threadIdx.x, T = blockDim.x;
  // The write-only actions go out in 16-byte stores: words [wa, va) and [vb, wb) one by one, [va, vb) four at a time
  // (the HARQ entry is 16-byte aligned). These loops are what the kernel spends its issue slots on.
  const int va = min(wb, (wa + 3) & ~3), vb = max(va, wb & ~3);
  uint4*    out4 = reinterpret_cast<uint4*>(out);
  switch (action) {
    case DM_ZERO:
      for (int w = wa + tid; w < va; w += T) {
        out[w] = 0u;
      }
      for (int q = (va >> 2) + tid; q < (vb >> 2); q += T) {
        out4[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      for (int w = vb + tid; w < wb; w += T) {
        out[w] = 0u;
      }
      break;
    case DM_FILL:
      for (int w = wa + tid; w < va; w += T) {
        out[w] = 0x7f7f7f7fu;
        last.note(w, 0x7f7f7f7fu);
      }
      for (int q = (va >> 2) + tid; q < (vb >> 2); q += T) {
        out4[q] = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
        last.note(4 * q + 3, 0x7f7f7f7fu);
      }
      for (int w = vb + tid; w < wb; w += T) {
        out[w] = 0x7f7f7f7fu;
        last.note(w, 0x7f7f7f7fu);
      }
      break;
    case DM_COPY:
      for (int w = wa + tid; w < va; w += T) {
        const uint32_t r = dm_lds_u32_unaligned(sh, 4 * w + ioff);
        out[w]           = r;
        last.note(w, r);
      }
      for (int q = (va >> 2) + tid; q < (vb >> 2); q += T) {
        // sixteen staged bytes from an arbitrary byte offset: five aligned words, four funnel shifts
        const int       a   = 16 * q + ioff;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(sh + (a & ~3));
        const uint32_t  sh8 = 8u * (uint32_t)(a & 3);
        const uint32_t  s0 = src[0], s1 = src[1], s2 = src[2], s3 = src[3], s4 = src[4];
        uint4           r;
        r.x = __funnelshift_r(s0, s1, sh8);
        r.y = __funnelshift_r(s1, s2, sh8);
        r.z = __funnelshift_r(s2, s3, sh8);
        r.w = __funnelshift_r(s3, s4, sh8);
        out4[q] = r;
        last.note(4 * q, r.x);
        last.note(4 * q + 1, r.y);
        last.note(4 * q + 2, r.z);
        last.note(4 * q + 3, r.w);
      }
      for (int w = vb + tid; w < wb; w += T) {
        const uint32_t r = dm_lds_u32_unaligned(sh, 4 * w + ioff);
        out[w]           = r;
        last.note(w, r);
      }
      break;
    case DM_COMB_ZERO:
      for (int w = wa + tid; w < wb; w += T) {
        const uint32_t in = dm_lds_u32_unaligned(sh, 4 * w + ioff);
        uint32_t       r  = in; // clamp(0 + in) for finite soft bits
        if (dm_any_nonfinite(in)) {
          r = dm_word_general<true>(g, llr, sh, out, w);
        } else {
          out[w] = r;
        }
        last.note(w, r);
      }
      break;
    case DM_COMB:
      for (int w = wa + tid; w < wb; w += 2 * T) {
        const int      w2   = w + T;
        const bool     has2 = w2 < wb;
        const uint32_t o1   = out[w];
        const uint32_t o2   = has2 ? out[w2] : 0u;
        const uint32_t i1   = dm_lds_u32_unaligned(sh, 4 * w + ioff);
        const uint32_t i2   = has2 ? dm_lds_u32_unaligned(sh, 4 * w2 + ioff) : 0u;
        uint32_t       r1, r2 = 0;
        if (dm_any_nonfinite(o1) || dm_any_nonfinite(i1)) {
          r1 = dm_word_general<true>(g, llr, sh, out, w);
        } else {
          r1 = dm_combine4_finite(o1, i1);
          if (r1 != o1) {
            out[w] = r1;
          }
        }
        last.note(w, r1);
        if (has2) {
          if (dm_any_nonfinite(o2) || dm_any_nonfinite(i2)) {
            r2 = dm_word_general<true>(g, llr, sh, out, w2);
          } else {
            r2 = dm_combine4_finite(o2, i2);
            if (r2 != o2) {
              out[w2] = r2;
            }
          }
          last.note(w2, r2);
        }
      }
      break;
    default: // DM_KEEP: only the position of the last non-zero soft bit is needed
      for (int w = wa + tid; w < wb; w += 2 * T) {
        const int      w2 = w + T;
        const uint32_t o1 = out[w];
        const uint32_t o2 = (w2 < wb) ? out[w2] : 0u;
        last.note(w, o1);
        last.note(w2, o2);
      }
      break;
  }
}

// Bit-plane separation of a group of four symbols (w[]: 4 * QM bytes, symbol after symbol): o[j] = the four soft bits of
// plane j. Byte permutes: one per plane for QPSK, three for 16QAM / 256QAM (pairs of symbols, then the pair of pairs);
// 64QAM (symbols of six bytes straddle words) goes byte by byte.
template <int QM>
__device__ __forceinline__ void dm_planes(const uint32_t (&w)[QM], uint32_t (&o)[QM])
{
  if (QM == 2) {
    o[0] = __byte_perm(w[0], w[1], 0x6420);
    o[1] = __byte_perm(w[0], w[1], 0x7531);
  } else if (QM == 4) {
#pragma unroll
    for (int j = 0; j != QM; ++j) {
      const uint32_t sel = (uint32_t)(((4 + j) << 4) | j);
      o[j] = __byte_perm(__byte_perm(w[0], w[1], sel), __byte_perm(w[2], w[3], sel), 0x5410);
    }
  } else if (QM == 8) {
#pragma unroll
    for (int j = 0; j != QM; ++j) {
      const uint32_t sel = (uint32_t)(((4 + (j & 3)) << 4) | (j & 3));
      const int      h   = j >> 2; // planes 4-7 sit in the second word of a symbol
      o[j] = __byte_perm(__byte_perm(w[h], w[2 + h], sel), __byte_perm(w[4 + h], w[6 + h], sel), 0x5410);
    }
  } else {
#pragma unroll
    for (int j = 0; j != QM; ++j) {
      uint32_t v = 0;
#pragma unroll
      for (int s4 = 0; s4 != 4; ++s4) {
        const int off = s4 * QM + j;
        v |= ((w[off >> 2] >> (8 * (off & 3))) & 0xffu) << (8 * s4);
      }
      o[j] = v;
    }
  }
}

// Stage the rate-matched input in shared memory in deinterleaved order: sh[j * Kq + sym] = llr[sym * QM + j].
// Four symbols per thread step: coalesced vector loads, bit planes separated with byte permutes.
template <int QM>
__device__ __forceinline__ void dm_stage_planes(const DematchGeom& g, const int8_t* __restrict__ llr, uint8_t* sh,
                                                int t0, int nt)
{
  const uint8_t*  src      = reinterpret_cast<const uint8_t*>(llr);
  const int       n_groups = g.Kq >> 2;
  constexpr int   VEC      = (QM % 4 == 0) ? 16 : 8; // bytes per vector load (a group is 4 * QM bytes)
  const bool      aligned  = (reinterpret_cast<uintptr_t>(src) % VEC) == 0;
  const bool      word_st  = (g.Kq & 3) == 0;        // plane starts are word aligned
  for (int grp = t0; grp < n_groups; grp += nt) {
    uint32_t       w[QM];
    const uint8_t* p = src + (size_t)grp * 4 * QM;
    if (aligned) {
      if (VEC == 16) {
#pragma unroll
        for (int r = 0; r != QM / 4; ++r) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + r);
          w[4 * r] = v.x, w[4 * r + 1] = v.y, w[4 * r + 2] = v.z, w[4 * r + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int r = 0; r != QM / 2; ++r) {
          const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + r);
          w[2 * r] = v.x, w[2 * r + 1] = v.y;
        }
      }
    } else {
      const uint32_t  sh8 = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
      const uint32_t* a   = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
      uint32_t        cur = __ldg(a);
#pragma unroll
      for (int r = 0; r != QM; ++r) {
        // The next aligned word holds at least one byte of this group whenever the pointer is misaligned.
        const uint32_t nxt = (sh8 != 0 || r + 1 != QM) ? __ldg(a + r + 1) : 0u;
        w[r]               = __funnelshift_r(cur, nxt, sh8);
        cur                = nxt;
      }
    }
    if (g.seq != nullptr) {
      // Deferred descrambling: the 4 * QM soft bits of the group against their sequence elements.
      const uint32_t bits = seq_bits32(g.seq, g.seq_base + (uint32_t)grp * 4u * QM);
#pragma unroll
      for (int r = 0; r != QM; ++r) {
        w[r] = negate4(w[r], (bits >> (4 * r)) & 0xfu);
      }
    }
    uint32_t ow[QM];
    dm_planes<QM>(w, ow);
#pragma unroll
    for (int j = 0; j != QM; ++j) {
      const uint32_t o   = ow[j];
      uint8_t*       dst = sh + j * g.Kq + 4 * grp;
      if (word_st) {
        *reinterpret_cast<uint32_t*>(dst) = o;
      } else {
        dst[0] = (uint8_t)o, dst[1] = (uint8_t)(o >> 8), dst[2] = (uint8_t)(o >> 16), dst[3] = (uint8_t)(o >> 24);
      }
    }
  }
  // Symbols after the last complete group.
  const int sym0 = 4 * n_groups;
  for (int idx = t0; idx < (g.Kq - sym0) * QM; idx += nt) {
    const int s4 = idx / QM, j = idx - s4 * QM;
    const int k  = (sym0 + s4) * QM + j;
    uint8_t   v  = src[(size_t)k];
    if (g.seq != nullptr && seq_bit(g.seq, g.seq_base + (uint32_t)k)) {
      v = (uint8_t)(0u - v);
    }
    sh[j * g.Kq + sym0 + s4] = v;
  }
}

// Thread t0 of the nt staging threads.
__device__ __forceinline__ void dm_stage(const DematchGeom& g, const int8_t* __restrict__ llr, uint8_t* sh, int t0, int nt)
{
  switch (g.qm) {
    case 2:
      dm_stage_planes<2>(g, llr, sh, t0, nt);
      break;
    case 4:
      dm_stage_planes<4>(g, llr, sh, t0, nt);
      break;
    case 6:
      dm_stage_planes<6>(g, llr, sh, t0, nt);
      break;
    case 8:
      dm_stage_planes<8>(g, llr, sh, t0, nt);
      break;
    default:
      // One bit per symbol: the stream is already in order. (Other modulation orders are never staged.)
      for (int i = t0; i < g.E; i += nt) {
        uint8_t v = (uint8_t)__ldg(llr + i);
        if (g.seq != nullptr && seq_bit(g.seq, g.seq_base + (uint32_t)i)) {
          v = (uint8_t)(0u - v);
        }
        sh[i] = v;
      }
      break;
  }
}

// gridDim = (codeblocks, parts). Each CTA owns a contiguous range of the 32-bit words of the HARQ entry.
// ---- express path ------------------------------------------------------------------------------------------------------
// The common new transmission - redundancy version 0, one lap of the circular buffer, everything a multiple of four
// soft bits - needs neither the staging buffer nor the segment machinery: every output word is four consecutive symbols
// of one bit plane (straight from a coalesced vector load of 4 symbols x QM planes, separated with byte permutes), a
// filler word, a stale word (only inspected for the last non-zero soft bit) or zero:
//   [0, info)            <- deinterleaved input [0, info)
//   [info, K_sys)        <- +infinity (filler bits)
//   [K_sys, E + F)       <- deinterleaved input [info, E)
//   [E + F, zf_lo)       stale (limited buffer only: the reference zeroes relative to the END of the N-long buffer)
//   [zf_lo, N)           <- 0
// Same bytes as the general path (the parity tests compare whole HARQ entries); about a quarter of its instructions.
// Decided by every thread from the descriptor alone (before anything is staged); implies everything dm_express needs
// from the geometry: no offset (rv 0), one lap at most (E <= Ncb - F), input beyond the systematic part, word alignment
// of every region boundary but the start of the final zeroing (that word is cut, see dm_express).
__device__ __forceinline__ bool dm_express_pre(const pdc_cb_desc& d, const int8_t* llr)
{
  const int qm = d.qm, bg = d.base_graph, Z = d.lifting_size;
  if (!(d.flags & PDC_CB_NEW_DATA) || d.rv != 0 || !(qm == 2 || qm == 4 || qm == 6 || qm == 8) || (bg != 1 && bg != 2) ||
      Z < 2 || Z > MAX_Z) {
    return false;
  }
  const int      N = ((bg == 1) ? 66 : 50) * Z, K_sys = ((bg == 1) ? 20 : 8) * Z, F = d.nof_filler, E = (int)d.rm_length;
  const int      Ncb = (d.nref > 0) ? min((int)d.nref, N) : N;
  const uint32_t vec = (qm % 4 == 0) ? 16u : 8u;
  return (E % (4 * qm)) == 0 && (F & 3) == 0 && (N & 3) == 0 && F < K_sys && Ncb > K_sys && E > K_sys - F &&
         E + F <= Ncb && (reinterpret_cast<uintptr_t>(llr) % vec) == 0;
}
__device__ __forceinline__ bool dm_express_ok(const DematchGeom& g)
{
  return g.new_data && !g.wrapped && g.k0 == 0 && g.E > g.info && g.E + g.F <= g.Ncb && (g.N & 3) == 0 &&
         (g.info & 3) == 0 && g.zf_lo >= g.E + g.F && g.zf_lo <= g.N;
}

template <int QM>
__device__ __forceinline__ void dm_express_copy(const DematchGeom& g, const int8_t* __restrict__ llr, uint32_t* out,
                                                int t0, int nt, DmLast& last)
{
  const uint8_t* src      = reinterpret_cast<const uint8_t*>(llr);
  const int      n_groups = g.Kq >> 2;
  constexpr int  VEC      = (QM % 4 == 0) ? 16 : 8;
  constexpr int  U        = (QM <= 2) ? 4 : (QM <= 4 ? 2 : 1); // groups in flight per thread (the loop is latency bound)
  for (int grp0 = t0; grp0 < n_groups; grp0 += U * nt) {
    uint32_t w[U][QM];
#pragma unroll
    for (int u = 0; u != U; ++u) {
      const int      grp = grp0 + u * nt;
      const uint8_t* p   = src + (size_t)grp * 4 * QM;
      if (grp < n_groups) {
        if (VEC == 16) {
#pragma unroll
          for (int r = 0; r != QM / 4; ++r) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + r);
            w[u][4 * r] = v.x, w[u][4 * r + 1] = v.y, w[u][4 * r + 2] = v.z, w[u][4 * r + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int r = 0; r != QM / 2; ++r) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + r);
            w[u][2 * r] = v.x, w[u][2 * r + 1] = v.y;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u != U; ++u) {
      const int grp = grp0 + u * nt;
      if (grp >= n_groups) {
        break;
      }
      if (g.seq != nullptr) {
        const uint32_t bits = seq_bits32(g.seq, g.seq_base + (uint32_t)grp * 4u * QM);
#pragma unroll
        for (int r = 0; r != QM; ++r) {
          w[u][r] = negate4(w[u][r], (bits >> (4 * r)) & 0xfu);
        }
      }
      uint32_t ow[QM];
      dm_planes<QM>(w[u], ow);
#pragma unroll
      for (int j = 0; j != QM; ++j) {
        const uint32_t o = ow[j];
        const int      i = j * g.Kq + 4 * grp; // deinterleaved index of the word's first soft bit
        const int      pw = ((i < g.info) ? i : i + g.F) >> 2;
        PDC_ASSERT(4 * pw + 4 <= g.E + g.F && 4 * pw + 4 <= PDC_MAX_CB_SOFT);
        out[pw] = o;
        last.note(pw, o);
      }
    }
  }
}

// old_last: the entry's previous record of its last non-zero soft bit (1 + index; -1 = unknown), see dm_express_combine.
__device__ __forceinline__ void dm_express(const DematchGeom& g, const int8_t* __restrict__ llr, uint32_t* out, int t0,
                                           int nt, DmLast& last, int old_last)
{
  switch (g.qm) {
    case 2:
      dm_express_copy<2>(g, llr, out, t0, nt, last);
      break;
    case 4:
      dm_express_copy<4>(g, llr, out, t0, nt, last);
      break;
    case 6:
      dm_express_copy<6>(g, llr, out, t0, nt, last);
      break;
    default:
      dm_express_copy<8>(g, llr, out, t0, nt, last);
      break;
  }
  for (int w = (g.info >> 2) + t0; w < (g.K_sys >> 2); w += nt) {
    out[w] = 0x7f7f7f7fu;
    last.note(w, 0x7f7f7f7fu);
  }
  // Stale soft bits (limited buffer) still count for the decoder's trimming. They are unchanged, so the previous record
  // usually tells without reading them: nothing non-zero up there, or the very soft bit it points at.
  const int s_lo = g.E + g.F, s_hi = g.zf_lo & ~3;
  if (old_last >= 0 && old_last <= s_lo) {
    // nothing
  } else if (old_last >= 0 && old_last - 1 < s_hi) {
    if (t0 == 0) {
      const int w = (old_last - 1) >> 2;
      last.note(w, out[w]);
    }
  } else {
    for (int w = (s_lo >> 2) + t0; w < (s_hi >> 2); w += nt) {
      last.note(w, out[w]);
    }
  }
  if ((g.zf_lo & 3) != 0 && t0 == 0) {
    // the word the final zeroing starts in: its first bytes stay, the rest is zeroed
    const int      w = g.zf_lo >> 2;
    const uint32_t r = out[w] & (0xffffffffu >> (8 * (4 - (g.zf_lo & 3))));
    out[w]           = r;
    last.note(w, r);
  }
  for (int w = ((g.zf_lo + 3) >> 2) + t0; w < (g.N >> 2); w += nt) {
    out[w] = 0u;
  }
}

// ---- retransmissions: express combine -----------------------------------------------------------------------------------
//
// A retransmission that walks the circular buffer at most once (E <= Ncb - F), with everything a multiple of four soft
// bits, needs neither the staging buffer nor the segments either: four consecutive symbols of one bit plane (one word
// out of the coalesced vector load, as in dm_express_copy) meet ONE aligned word of the HARQ entry,
//   deinterleaved index i -> data position q = (d0 + i) mod (Ncb - F) -> buffer position p = q (q < info) or q + F,
// and are combined four at a time (VIADDMNMX.S16x2; a word with a non-finite operand on either side goes position by
// position through the general path, which knows the reference's two combine rules). The positions the walk does not
// touch are only inspected for the last non-zero soft bit. Same bytes as the staged path; the HARQ-buffer parity tests
// compare whole entries.
__device__ __forceinline__ bool dm_xcomb_pre(const pdc_cb_desc& d, const int8_t* llr)
{
  const int qm = d.qm, bg = d.base_graph, Z = d.lifting_size;
  if ((d.flags & PDC_CB_NEW_DATA) || d.rv > 3 || !(qm == 2 || qm == 4 || qm == 6 || qm == 8) || (bg != 1 && bg != 2) ||
      Z < 4 || Z > MAX_Z || (Z & 3) != 0) {
    return false;
  }
  const int N = ((bg == 1) ? 66 : 50) * Z, K_sys = ((bg == 1) ? 20 : 8) * Z, F = d.nof_filler, E = (int)d.rm_length;
  const int Ncb = (d.nref > 0) ? min((int)d.nref, N) : N;
  if ((F & 3) != 0 || F >= K_sys || Ncb <= K_sys || (E % (4 * qm)) != 0) {
    return false;
  }
  const int sf1[4] = {0, 17, 33, 56};
  const int sf2[4] = {0, 13, 25, 43};
  const int k0 = ((((bg == 1) ? sf1[d.rv] : sf2[d.rv]) * Ncb) / N) * Z, info = K_sys - F, Dn = Ncb - F;
  const int d0 = (k0 < info) ? k0 : ((k0 < K_sys) ? info : k0 - F);
  const uint32_t vec = (qm % 4 == 0) ? 16u : 8u;
  return E <= Dn && (d0 + E <= Dn || (Dn & 3) == 0) && (reinterpret_cast<uintptr_t>(llr) % vec) == 0;
}
__device__ __forceinline__ bool dm_xcomb_ok(const DematchGeom& g)
{
  return !g.new_data && g.E <= g.Dn && (g.d0 & 3) == 0 && (g.info & 3) == 0 && (g.F & 3) == 0 && (g.N & 3) == 0 &&
         (!g.wrapped || (g.Dn & 3) == 0);
}

// Buffer word the four symbols of bit plane j, group grp meet.
__device__ __forceinline__ int dm_xcomb_word(const DematchGeom& g, int j, int grp)
{
  int q = g.d0 + j * g.Kq + 4 * grp;
  q -= (q >= g.Dn) ? g.Dn : 0;
  return ((q < g.info) ? q : q + g.F) >> 2;
}

// `slow`: list (in the unused staging buffer) of the words with a non-finite operand, `n_slow` its length; they are left
// untouched here and combined position by position once the sweep is over (dm_express_combine) - a call to the general
// path inside this loop would cost every iteration its spills.
template <int QM>
__device__ __forceinline__ void dm_xcomb_words(const DematchGeom& g, const int8_t* __restrict__ llr, uint32_t* out, int t0,
                                               int nt, DmLast& last, uint16_t* slow, int* n_slow)
{
  const uint8_t* src      = reinterpret_cast<const uint8_t*>(llr);
  const int      n_groups = g.Kq >> 2;
  constexpr int  VEC      = (QM % 4 == 0) ? 16 : 8;
  constexpr int  U        = (QM <= 2) ? 2 : 1; // groups in flight per thread
  for (int grp0 = t0; grp0 < n_groups; grp0 += U * nt) {
    uint32_t w[U][QM], old[U][QM];
    int      pw[U][QM];
#pragma unroll
    for (int u = 0; u != U; ++u) {
      const int      grp = grp0 + u * nt;
      const uint8_t* p   = src + (size_t)grp * 4 * QM;
      if (grp < n_groups) {
        if (VEC == 16) {
#pragma unroll
          for (int r = 0; r != QM / 4; ++r) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + r);
            w[u][4 * r] = v.x, w[u][4 * r + 1] = v.y, w[u][4 * r + 2] = v.z, w[u][4 * r + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int r = 0; r != QM / 2; ++r) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p) + r);
            w[u][2 * r] = v.x, w[u][2 * r + 1] = v.y;
          }
        }
        // the HARQ words these symbols meet, requested together with them
        int q = g.d0 + 4 * grp;
#pragma unroll
        for (int j = 0; j != QM; ++j) {
          const int qq = q - ((q >= g.Dn) ? g.Dn : 0);
          pw[u][j]     = ((qq < g.info) ? qq : qq + g.F) >> 2;
          PDC_ASSERT(pw[u][j] >= 0 && 4 * pw[u][j] + 4 <= g.Ncb && 4 * pw[u][j] + 4 <= PDC_MAX_CB_SOFT);
          old[u][j] = out[pw[u][j]];
          q += g.Kq;
        }
      }
    }
#pragma unroll
    for (int u = 0; u != U; ++u) {
      const int grp = grp0 + u * nt;
      if (grp >= n_groups) {
        break;
      }
      if (g.seq != nullptr) {
        const uint32_t bits = seq_bits32(g.seq, g.seq_base + (uint32_t)grp * 4u * QM);
#pragma unroll
        for (int r = 0; r != QM; ++r) {
          w[u][r] = negate4(w[u][r], (bits >> (4 * r)) & 0xfu);
        }
      }
      uint32_t ow[QM];
      dm_planes<QM>(w[u], ow);
#pragma unroll
      for (int j = 0; j != QM; ++j) {
        if (dm_any_nonfinite(old[u][j]) || dm_any_nonfinite(ow[j])) {
          slow[atomicAdd(n_slow, 1)] = (uint16_t)pw[u][j];
        } else {
          const uint32_t res = dm_combine4_finite(old[u][j], ow[j]);
          out[pw[u][j]]      = res;
          last.note(pw[u][j], res);
        }
      }
    }
  }
}

// old_last: what the entry's previous dematcher pass recorded (maximum of its slots, -1 if any is unknown); cta_last: the
// CTA's running maximum (shared memory, zero on entry).
__device__ __forceinline__ void dm_express_combine(const DematchGeom& g, const int8_t* __restrict__ llr, uint32_t* out,
                                                   int t0, int nt, DmLast& last, uint16_t* slow, int* n_slow, int old_last,
                                                   int* cta_last)
{
  switch (g.qm) {
    case 2:
      dm_xcomb_words<2>(g, llr, out, t0, nt, last, slow, n_slow);
      break;
    case 4:
      dm_xcomb_words<4>(g, llr, out, t0, nt, last, slow, n_slow);
      break;
    case 6:
      dm_xcomb_words<6>(g, llr, out, t0, nt, last, slow, n_slow);
      break;
    default:
      dm_xcomb_words<8>(g, llr, out, t0, nt, last, slow, n_slow);
      break;
  }
  // The words set aside: position by position (the reference's two rules for non-finite operands).
  __syncthreads();
  {
    const int n = *n_slow;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const int w = slow[k];
      last.note(w, dm_word_general<false>(g, llr, nullptr, out, w));
    }
  }
  // What the walk does not touch still counts for the decoder's trimming: the filler bits, what lies beyond the circular
  // buffer, and the data positions outside [d0, d0 + E) (one interval if the walk wrapped, two if not), each of which
  // maps to at most two stretches of the buffer (before and behind the filler bits). A word is touched as a whole or not
  // at all and belongs to the stretch its first position is in.
  int lo[6], hi[6];
  lo[0] = g.info, hi[0] = g.K_sys;
  lo[1] = g.Ncb, hi[1] = g.N;
  {
    int qa[2], qb[2];
    if (g.wrapped) {
      qa[0] = g.d0 + g.E - g.Dn, qb[0] = g.d0;
      qa[1] = 0, qb[1] = 0;
    } else {
      qa[0] = 0, qb[0] = g.d0;
      qa[1] = g.d0 + g.E, qb[1] = g.Dn;
    }
#pragma unroll
    for (int k = 0; k != 2; ++k) {
      lo[2 + 2 * k] = qa[k], hi[2 + 2 * k] = min(qb[k], g.info);
      lo[3 + 2 * k] = max(qa[k], g.info) + g.F, hi[3 + 2 * k] = qb[k] + g.F;
    }
  }
  // They are unchanged, so the entry's previous "last non-zero" (old_last = 1 + its index, -1 = unknown) usually settles
  // it without reading them: if that soft bit is itself untouched it still is the last one among them; if not, the last
  // untouched non-zero lies below `bound` (the end of the highest untouched stretch under old_last), and a part whose own
  // touched words reach at least that far need not look (the decoder takes the maximum over the parts).
  bool scan = true;
  if (old_last >= 0) {
    int  bound     = 0;
    bool untouched = false;
#pragma unroll
    for (int k = 0; k != 6; ++k) {
      const int l4 = (lo[k] + 3) & ~3; // first position of the first word of the stretch
      if (l4 < hi[k] && l4 < old_last) {
        bound     = max(bound, min(old_last, (hi[k] + 3) & ~3));
        untouched = untouched || (old_last - 1 >= l4 && ((old_last - 1) & ~3) < hi[k]);
      }
    }
    int pos = last.position();
    for (int o = 16; o > 0; o >>= 1) {
      pos = max(pos, __shfl_xor_sync(0xffffffffu, pos, o));
    }
    if ((threadIdx.x & 31) == 0 && pos > 0) {
      atomicMax(cta_last, pos);
    }
    __syncthreads();
    if (untouched) {
      scan = false;
      if (t0 == 0) {
        atomicMax(cta_last, old_last);
      }
    } else if (*cta_last >= bound) {
      scan = false;
    }
  }
  if (scan) {
#pragma unroll
    for (int k = 0; k != 6; ++k) {
      // words whose first position p = 4 w lies in [lo, hi)
      for (int w = ((lo[k] + 3) >> 2) + t0; 4 * w < hi[k]; w += nt) {
        PDC_ASSERT(w >= 0 && 4 * w < g.N);
        last.note(w, out[w]);
      }
    }
  }
}

__global__ void __launch_bounds__(DM_THREADS, 6) rate_dematch_kernel(BatchParams prm)
{
  __shared__ __align__(16) uint8_t sh_in[DM_STAGE_BYTES + 16];
  __shared__ DematchGeom           g_sh;
  __shared__ int                   ok;
  __shared__ int                   sh_last;
  __shared__ int                   sh_n_slow;
  __shared__ int                   sh_bp[DM_MAX_BP];
  __shared__ DmSeg                 sh_seg[DM_MAX_BP];
  const int                        tid = threadIdx.x;
  uint32_t                         cb  = blockIdx.x;
  const pdc_cb_desc                d   = prm.cbs[cb];
  // Launched with programmatic serialization: the descriptor was uploaded before the previous kernel began; everything
  // else (soft bits of the front end, the codeblock-to-codeword map, HARQ entries a decoder may still be reading) waits.
  pdl_wait();
  if (tid == 0 && blockIdx.y == 0 && prm.results != nullptr) {
    // "Not run" until the decoder says otherwise: saves the batch a separate clear of the result array.
    pdc_cb_result none;
    none.crc_ok = 0, none.iters = 0, none.status = 0, none.nlayers = 0;
    prm.results[cb] = none;
  }
  if (!(d.flags & PDC_CB_DEMATCH)) {
    return;
  }
  // The first warp derives the geometry and the segments while the others stage the input, for which the modulation
  // order and the length are enough.
  const int8_t*   llr      = prm.llrs + d.llr_offset;
  const uint32_t* seq      = nullptr;
  uint32_t        seq_base = 0;
  if (prm.cb_scr != nullptr) {
    const uint4 e = prm.cb_scr[cb];
    if (e.x != CB_NOT_SCRAMBLED) {
      // The codeblock's soft bits were left scrambled in the raw codeword: descramble while staging.
      llr      = prm.raw + e.z;
      seq      = prm.seq + e.x;
      seq_base = e.y;
    }
  }
  uint32_t* out = reinterpret_cast<uint32_t*>(prm.harq + (size_t)min(d.harq_id, prm.harq_entries - 1) * PDC_MAX_CB_SOFT);
  const int     qm = d.qm, E = d.rm_length;
  // The entry's record of its last non-zero soft bit, requested now (the express paths use it long after the round trip).
  const int4 slots = (gridDim.y == 1)
                         ? __ldg(reinterpret_cast<const int4*>(prm.harq_last) + min(d.harq_id, prm.harq_entries - 1))
                         : make_int4(-1, -1, -1, -1);
  // Express path candidates skip the staging; the geometry (first warp) has the last word.
  const bool    express_pre = dm_express_pre(d, llr);
  const bool    xcomb_pre   = !express_pre && dm_xcomb_pre(d, llr);
  const bool    staged = !express_pre && !xcomb_pre && (qm == 1 || qm == 2 || qm == 4 || qm == 6 || qm == 8) && E > 0 &&
                      (E % qm) == 0 && E <= DM_STAGE_BYTES;
  if (tid < 32) {
    if (tid == 0) {
      ok            = dm_geometry(d, prm.simd_width, g_sh) && (d.harq_id < prm.harq_entries);
      g_sh.seq      = seq;
      g_sh.seq_base = seq_base;
      sh_last       = 0;
      sh_n_slow     = 0;
    }
    __syncwarp();
    if (ok && (express_pre || xcomb_pre)) {
      if (tid == 0) {
        g_sh.staged = 0; // nothing was staged: whatever the express path does not take goes position by position
      }
    } else if (ok && g_sh.staged && g_sh.E <= g_sh.Dn) {
      // Sort the breakpoints (rank by counting) and classify the segments between them.
      const int mine = dm_breakpoint(g_sh, min(tid, DM_MAX_BP - 1));
      int       rank = 0;
      for (int k = 0; k != DM_MAX_BP; ++k) {
        const int c = __shfl_sync(0xffffffffu, mine, k);
        rank += (c < mine || (c == mine && k < tid)) ? 1 : 0;
      }
      if (tid < DM_MAX_BP) {
        sh_bp[rank] = mine;
      }
      __syncwarp();
      if (tid < DM_MAX_BP - 1) {
        DmSeg sg;
        sg.p0     = sh_bp[tid];
        sg.p1     = sh_bp[tid + 1];
        sg.ioff   = 0;
        sg.action = DM_KEEP;
        if (sg.p0 < sg.p1) {
          sg.action = dm_classify(g_sh, sg.p0, sg.ioff);
        }
        sh_seg[tid] = sg;
      }
    }
  }
  if (staged && tid >= 32) {
    DematchGeom gs; // the fields staging needs
    gs.qm = qm, gs.E = E, gs.Kq = E / qm, gs.seq = seq, gs.seq_base = seq_base;
    dm_stage(gs, llr, sh_in, tid - 32, (int)blockDim.x - 32);
  }
  __syncthreads();
  if (!ok) {
    return; // the decode kernel reports the invalid descriptor
  }
  const int  N    = g_sh.N;
  const int  nw   = (N + 3) >> 2;
  const bool fast = staged && g_sh.E <= g_sh.Dn; // single lap
  const int  per  = (nw + (int)gridDim.y - 1) / (int)gridDim.y;
  const int  w_lo = (int)blockIdx.y * per;
  const int  w_hi = min(nw, w_lo + per);
  DmLast last;
  PDC_ASSERT(!express_pre || dm_express_ok(g_sh));
  PDC_ASSERT(!xcomb_pre || dm_xcomb_ok(g_sh));
  // The entry's previous record of its last non-zero soft bit (with several CTAs per codeblock a fast one may have
  // replaced its slot before a slow one reads them: unknown then).
  // The record must describe all N positions of THIS codeblock: an entry last used by a shorter codeblock may still hold
  // what a longer one left behind it.
  int old_last = -1;
  if (gridDim.y == 1 && (express_pre || xcomb_pre)) {
    old_last = harq_last_known(slots, N);
  }
  if (express_pre && dm_express_ok(g_sh)) {
    dm_express(g_sh, llr, out, (int)blockIdx.y * (int)blockDim.x + tid, (int)gridDim.y * (int)blockDim.x, last, old_last);
  } else if (xcomb_pre && dm_xcomb_ok(g_sh)) {
    // (a CTA meets at most N / 4 <= 6336 words: their 16-bit indices fit the staging buffer this path does not use)
    static_assert(DM_STAGE_BYTES >= 2 * (PDC_MAX_CB_SOFT / 4), "slow-word list");
    dm_express_combine(g_sh, llr, out, (int)blockIdx.y * (int)blockDim.x + tid, (int)gridDim.y * (int)blockDim.x, last,
                       reinterpret_cast<uint16_t*>(sh_in), &sh_n_slow, old_last, &sh_last);
  } else if (fast) {
    // Words cut by a breakpoint (and the incomplete last word): one thread per soft bit, the four of a word side by
    // side; the HARQ word is requested before the segment loops and used after them.
    const bool cut_thread = tid < 4 * DM_MAX_BP;
    bool       cut_mine   = false;
    int        cut_w      = 0;
    uint32_t   cut_old    = 0;
    if (cut_thread) {
      const int b = sh_bp[tid >> 2];
      cut_w       = b >> 2;
      cut_mine    = (b & 3) != 0 && cut_w >= w_lo && cut_w < w_hi;
      if (cut_mine && tid >= 4) {
        const int prev = sh_bp[(tid >> 2) - 1];
        cut_mine       = !((prev & 3) != 0 && (prev >> 2) == cut_w);
      }
      if (cut_mine) {
        cut_old = out[cut_w];
      }
    }
    for (int s = 0; s != DM_MAX_BP - 1; ++s) {
      const DmSeg sg = sh_seg[s];
      const int   wa = max((sg.p0 + 3) >> 2, w_lo);
      const int   wb = min(sg.p1 >> 2, w_hi);
      if (wa < wb) {
        dm_segment_words(g_sh, llr, sh_in, out, sg.action, sg.ioff, wa, wb, last);
      }
    }
    if (cut_thread) {
      const int k   = tid & 3;
      uint32_t  res = 0;
      if (cut_mine) {
        const int v = dm_position<true>(g_sh, llr, sh_in, 4 * cut_w + k, (int)(int8_t)(cut_old >> (8 * k)));
        res         = (uint32_t)(uint8_t)(int8_t)v << (8 * k);
      }
      const uint32_t lanes = (tid < 32) ? 0xffffffffu : ((1u << (4 * DM_MAX_BP - 32)) - 1u);
      res |= __shfl_xor_sync(lanes, res, 1);
      res |= __shfl_xor_sync(lanes, res, 2);
      if (cut_mine && k == 0) {
        if (res != cut_old) {
          out[cut_w] = res;
        }
        if (4 * cut_w + 3 >= g_sh.N) { // bytes beyond N keep their old value and do not count as soft bits
          res &= 0xffffffffu >> (8 * (4 * cut_w + 4 - g_sh.N));
        }
        last.note(cut_w, res);
      }
    }
  } else if (staged) {
    for (int w = w_lo + tid; w < w_hi; w += blockDim.x) {
      last.note(w, dm_word_general<true>(g_sh, llr, sh_in, out, w));
    }
  } else {
    for (int w = w_lo + tid; w < w_hi; w += blockDim.x) {
      last.note(w, dm_word_general<false>(g_sh, llr, sh_in, out, w));
    }
  }
  // The decoder trims trailing zeros (ldpc_decoder_impl.cpp:86-99): hand it the position of the last non-zero soft
  // bit, one slot per part (the decoder takes the maximum).
  int pos = last.position();
  for (int o = 16; o > 0; o >>= 1) {
    pos = max(pos, __shfl_xor_sync(0xffffffffu, pos, o));
  }
  if ((tid & 31) == 0 && pos > 0) {
    atomicMax(&sh_last, pos);
  }
  __syncthreads();
  if (tid == 0) {
    // (the parts together have seen or written every one of the N positions: the record is exact for [0, N))
    int32_t* slot    = prm.harq_last + (size_t)d.harq_id * DM_MAX_PARTS;
    slot[blockIdx.y] = harq_last_pack(sh_last, N);
    if (blockIdx.y == 0) {
      for (int k = (int)gridDim.y; k < DM_MAX_PARTS; ++k) {
        slot[k] = harq_last_pack(0, N);
      }
    }
  }
}

inline cudaError_t launch_rate_dematch(const BatchParams& p, int sm_count, cudaStream_t s)
{
  if (p.n_cb == 0) {
    return cudaSuccess;
  }
  // Small batches are latency bound: up to DM_MAX_PARTS CTAs per codeblock until the GPU is filled a few times over.
  int parts = 1;
  while (parts < DM_MAX_PARTS && (size_t)p.n_cb * parts < (size_t)8 * sm_count) {
    parts *= 2;
  }
  return launch_pdl(rate_dematch_kernel, dim3(p.n_cb, parts), dim3(DM_THREADS), 0, s, p);
}

} // namespace pdc
