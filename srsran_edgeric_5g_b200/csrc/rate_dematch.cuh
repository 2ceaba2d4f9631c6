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

struct DematchGeom {
  int N, Ncb, Z, K_sys, F, info, k0, d0, Dn, E, qm, Kq;
  float inv_Kq;
  int new_data;
  int first_pass_len; // data positions from the start of the walk to the end of the circular buffer
  int wrapped;        // the walk went past the end of the circular buffer at least once
  int zlo;            // new data: [0, zlo) is zeroed
  int zf_lo;          // new data, not wrapped: [zf_lo, N) is zeroed at the end; N if not applicable
  int simd_width;
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

__device__ __forceinline__ int dm_fetch(const DematchGeom& g, const int8_t* __restrict__ llr, int i)
{
  if (g.qm == 1) {
    return __ldg(llr + i);
  }
  int j   = i / g.Kq; // bit plane
  int sym = i - j * g.Kq;
  return __ldg(llr + sym * g.qm + j);
}

__device__ __forceinline__ int dm_position(const DematchGeom& g, const int8_t* __restrict__ llr, int p, int old)
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
    val = dm_fetch(g, llr, i);
    i += g.Dn;
  }
  for (; i < g.E; i += g.Dn) {
    val = dm_combine(g, p, i, val, dm_fetch(g, llr, i));
  }
  return val;
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

// Word fast path: the four positions 4w..4w+3 lie in one region of the circular buffer, are fed by four consecutive
// deinterleaved inputs of one bit plane (or by none) and the transmission makes a single lap. Returns false if the
// word needs the general per-position path.
__device__ __forceinline__ bool dm_word_fast(const DematchGeom& g, const int8_t* __restrict__ llr, int p0, uint32_t old,
                                             uint32_t& res)
{
  if (p0 + 3 >= g.Ncb) {
    return false;
  }
  const bool in_info = p0 + 3 < g.info;
  if (!in_info && p0 < g.K_sys) {
    return false;
  }
  int  i         = (in_info ? p0 : p0 - g.F) - g.d0;
  bool first_lap = i >= 0;
  if (!first_lap) {
    if (i + 3 >= 0) {
      return false;
    }
    i += g.Dn;
  }
  // Base value of the four positions (what the sequential walk leaves there before combining).
  uint32_t base = old;
  if (g.new_data) {
    const bool z0 = (p0 < g.zlo) || (p0 >= g.zf_lo);
    const bool z3 = (p0 + 3 < g.zlo) || (p0 + 3 >= g.zf_lo);
    if (z0 != z3) {
      return false;
    }
    if (z0) {
      base = 0;
    }
  }
  if (i >= g.E) {
    res = base;
    return true;
  }
  if (i + 3 >= g.E) {
    return false;
  }
  // Four consecutive inputs of one bit plane.
  uint32_t in;
  if (g.qm == 1) {
    in = (uint32_t)(uint8_t)__ldg(llr + i) | ((uint32_t)(uint8_t)__ldg(llr + i + 1) << 8) |
         ((uint32_t)(uint8_t)__ldg(llr + i + 2) << 16) | ((uint32_t)(uint8_t)__ldg(llr + i + 3) << 24);
  } else {
    int j = __float2int_rz(__fmul_rz((float)i, g.inv_Kq));
    int sym = i - j * g.Kq;
    if (sym < 0) {
      --j;
      sym += g.Kq;
    } else if (sym >= g.Kq) {
      ++j;
      sym -= g.Kq;
    }
    if (sym + 3 >= g.Kq) {
      return false;
    }
    const int8_t* src = llr + sym * g.qm + j;
    in = (uint32_t)(uint8_t)__ldg(src) | ((uint32_t)(uint8_t)__ldg(src + g.qm) << 8) |
         ((uint32_t)(uint8_t)__ldg(src + 2 * g.qm) << 16) | ((uint32_t)(uint8_t)__ldg(src + 3 * g.qm) << 24);
  }
  if (g.new_data && first_lap) {
    res = in; // copy
    return true;
  }
  if (dm_any_nonfinite(base) || dm_any_nonfinite(in)) {
    return false;
  }
  res = dm_combine4_finite(base, in);
  return true;
}

// One CTA per codeblock; each thread owns 4 consecutive soft bits (one 32-bit read-modify-write of the HARQ entry).
__global__ void __launch_bounds__(256) rate_dematch_kernel(BatchParams prm)
{
  __shared__ DematchGeom g;
  __shared__ int         ok;
  __shared__ int         sh_last;
  uint32_t               cb = blockIdx.x;
  const pdc_cb_desc&     d  = prm.cbs[cb];
  if (!(d.flags & PDC_CB_DEMATCH)) {
    return;
  }
  if (threadIdx.x == 0) {
    ok      = dm_geometry(d, prm.simd_width, g) && (d.harq_id < prm.harq_entries);
    sh_last = 0;
  }
  __syncthreads();
  if (!ok) {
    return; // the decode kernel reports the invalid descriptor
  }
  const int8_t* llr        = prm.llrs + d.llr_offset;
  uint32_t*     out        = reinterpret_cast<uint32_t*>(prm.harq + (size_t)d.harq_id * PDC_MAX_CB_SOFT);
  const int     nw         = (g.N + 3) >> 2;
  const bool    single_lap = g.E <= g.Dn;
  int           last       = 0;
  for (int w = threadIdx.x; w < nw; w += blockDim.x) {
    const uint32_t old = out[w];
    uint32_t       res = 0;
    if (!(single_lap && dm_word_fast(g, llr, 4 * w, old, res))) {
      res = 0;
#pragma unroll
      for (int k = 0; k != 4; ++k) {
        int o = (int)(int8_t)(old >> (8 * k));
        int v = dm_position(g, llr, 4 * w + k, o);
        res |= (uint32_t)(uint8_t)(int8_t)v << (8 * k);
      }
    }
    if (res != old) {
      out[w] = res;
    }
    if (res != 0) {
      // Highest non-zero position of this word (bytes beyond N keep their old value and do not count).
#pragma unroll
      for (int k = 0; k != 4; ++k) {
        if (((res >> (8 * k)) & 0xffu) != 0 && 4 * w + k < g.N) {
          last = max(last, 4 * w + k + 1);
        }
      }
    }
  }
  // The decoder trims trailing zeros (ldpc_decoder_impl.cpp:86-99): hand it the position of the last non-zero soft bit.
  for (int o = 16; o > 0; o >>= 1) {
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  if ((threadIdx.x & 31) == 0 && last > 0) {
    atomicMax(&sh_last, last);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    prm.harq_last[d.harq_id] = sh_last;
  }
}

inline cudaError_t launch_rate_dematch(const BatchParams& p, cudaStream_t s)
{
  if (p.n_cb == 0) {
    return cudaSuccess;
  }
  rate_dematch_kernel<<<p.n_cb, 256, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace pdc
