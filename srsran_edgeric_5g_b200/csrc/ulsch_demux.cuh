// Codeword front end on the device: TS 38.211 5.2.1 scrambling sequence, descrambling and UL-SCH demultiplexing.
//
// Reference behaviour: pseudo_random_generator_impl (lib/phy/upper/sequence_generators/pseudo_random_generator_impl.cpp),
// revert_scrambling + the per-block sequence generation of pusch_demodulator_impl::demodulate
// (lib/phy/upper/channel_processors/pusch/pusch_demodulator_impl.cpp:38-128, :254-259) and ulsch_demultiplex_impl
// (lib/phy/upper/channel_processors/pusch/ulsch_demultiplex_impl.cpp:111-198 placeholders, :501-589 demultiplexing).
//
// The reference walks the codeword OFDM symbol by OFDM symbol and pushes resource elements into four decoder buffers.
// Here every OUTPUT soft bit finds its input: the host plans, per symbol, where each stream starts and - for the few
// symbols that carry UCI - which resource elements belong to it (ulsch_plan.h); the kernels are gathers.
//   prg_kernel       : the scrambling sequence of every codeword, 128 bits per thread, started anywhere in the sequence
//                      by a polynomial jump (x^n mod g from a three-level table) instead of a serial advance.
//   ulsch_sch_kernel : UL-SCH stream, one aligned 32-bit output word per thread step (HBM-bound: 1 B in, 1 B out per
//                      soft bit + 1/8 B of sequence).
//   ulsch_uci_kernel : HARQ-ACK / CSI streams with the placeholder rules, one soft bit per thread.
#pragma once

#include "pdc_device.cuh"
#include "ulsch_plan.h"

namespace pdc {

constexpr uint32_t PRG_NC    = 1600;
constexpr uint32_t PRG_TAPS1 = 0x9u; // x1: x^31 + x^3 + 1
constexpr uint32_t PRG_TAPS2 = 0xfu; // x2: x^31 + x^3 + x^2 + x + 1
constexpr int      PRG_WORDS_PER_THREAD = 16; // large batches; small ones use PRG_WORDS_PER_THREAD_SMALL (latency bound)
constexpr int      PRG_WORDS_PER_THREAD_SMALL = 4;

// x^(c * 128^level) mod g of both generators, c < 128, level < 3 (sequence positions below 2^21).
__device__ uint32_t g_prg_jump[2][3][128];

// (a * b) mod g over GF(2), g = x^31 + taps.
__host__ __device__ inline uint32_t prg_mulmod(uint32_t a, uint32_t b, uint32_t taps)
{
  uint32_t r = 0;
  for (int i = 30; i >= 0; --i) {
    r <<= 1;
    if (r & 0x80000000u) {
      r = (r & 0x7fffffffu) ^ taps;
    }
    if ((b >> i) & 1u) {
      r ^= a;
    }
  }
  return r;
}

// The first 62 elements of the generator's sequence from its 31-bit initial state (bit j = element j).
__host__ __device__ inline uint64_t prg_prefix(uint32_t state, uint32_t taps)
{
  uint64_t s = state & 0x7fffffffu;
  for (int n = 0; n != 31; ++n) {
    // element n + 31 = XOR of the elements n + t over the taps t
    uint64_t f = 0;
    for (int t = 0; t != 4; ++t) {
      if ((taps >> t) & 1u) {
        f ^= (s >> (n + t)) & 1ull;
      }
    }
    s |= f << (n + 31);
  }
  return s;
}

// State (elements n .. n+30) of a generator: with x^n = sum p_i x^i (mod g), element n + j = sum p_i element(i + j).
__device__ __forceinline__ uint32_t prg_state_at(uint32_t p, uint64_t prefix)
{
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i != 31; ++i) {
    const uint32_t window = (uint32_t)(prefix >> i) & 0x7fffffffu;
    s ^= ((p >> i) & 1u) ? window : 0u;
  }
  return s;
}

__device__ __forceinline__ uint32_t prg_xpow(int gen, uint32_t n, uint32_t taps)
{
  const uint32_t a = g_prg_jump[gen][0][n & 127u], b = g_prg_jump[gen][1][(n >> 7) & 127u],
                 c = g_prg_jump[gen][2][(n >> 14) & 127u];
  return prg_mulmod(prg_mulmod(a, b, taps), c, taps);
}

// Sixteen sequence elements at once: the state holds elements n .. n+30, the recurrences reach back at most 31.
__device__ __forceinline__ uint32_t prg_step16(uint32_t& x1, uint32_t& x2)
{
  const uint32_t out = (x1 ^ x2) & 0xffffu;
  const uint32_t f1  = ((x1 >> 3) ^ x1) & 0xffffu;
  const uint32_t f2  = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 0xffffu;
  x1                 = (x1 >> 16) | (f1 << 15);
  x2                 = (x2 >> 16) | (f2 << 15);
  return out;
}

// seq[cw.seq_word_off + j] = elements cw.prg_offset + 32 j .. + 31 of c(n) for cw.c_init, element k in bit k.
template <int WPT>
__global__ void __launch_bounds__(128) prg_kernel(const UlschCodeword* __restrict__ cws, uint32_t* __restrict__ seq)
{
  const UlschCodeword& cw      = cws[blockIdx.y];
  const uint32_t       n_words = (cw.n_in + 31u) / 32u;
  const uint32_t       w0      = (blockIdx.x * blockDim.x + threadIdx.x) * WPT;
  if (w0 >= n_words) {
    return;
  }
  const uint32_t n  = PRG_NC + cw.prg_offset + 32u * w0;
  uint32_t       x1 = prg_state_at(prg_xpow(0, n, PRG_TAPS1), prg_prefix(1u, PRG_TAPS1));
  uint32_t       x2 = prg_state_at(prg_xpow(1, n, PRG_TAPS2), prg_prefix(cw.c_init, PRG_TAPS2));
  uint32_t*      out = seq + cw.seq_word_off + w0;
#pragma unroll
  for (int k = 0; k != WPT; ++k) {
    const uint32_t lo = prg_step16(x1, x2);
    const uint32_t hi = prg_step16(x1, x2);
    out[k]            = lo | (hi << 16);
  }
}

struct UlschArgs {
  const UlschCodeword* cws;
  const UlschSymbol*   syms;
  const uint16_t*      lists;
  const uint32_t*      seq;
  const int8_t*        in;
  int8_t*              sch;
  int8_t*              uci;
};

constexpr int ULSCH_MAX_SYMBOLS = 14;

// Symbol of stream k that holds soft bit o of the stream (symbols without elements of the stream are skipped).
__device__ __forceinline__ int ulsch_find_symbol(const UlschSymbol* S, int n_sym, int k, uint32_t o, uint32_t bpre)
{
  int s = 0;
  for (int i = 0; i != n_sym; ++i) {
    if (S[i].n_out_re[k] != 0 && S[i].out_off[k] <= o) {
      s = i;
    }
  }
  (void)bpre;
  return s;
}

// One soft bit of the UL-SCH stream, the general way.
__device__ __forceinline__ int ulsch_sch_byte(const UlschArgs& a, const UlschCodeword& cw, const UlschSymbol* S, uint32_t o)
{
  const int          s   = ulsch_find_symbol(S, (int)cw.n_sym, 0, o, cw.bpre);
  const UlschSymbol& sym = S[s];
  const uint32_t     rel = o - sym.out_off[0];
  const uint32_t     r = rel / cw.bpre, q = rel - r * cw.bpre;
  uint32_t           re = r;
  if (sym.list_off[0] != ULSCH_IDENTITY) {
    const uint32_t e = a.lists[sym.list_off[0] + r];
    if (e & ULSCH_PUNCTURED) {
      return 0;
    }
    re = e;
  }
  const uint32_t i = sym.in_off + re * cw.bpre + q;
  int            v = a.in[cw.in_off + i];
  if ((cw.flags & PDC_CW_SCRAMBLED) && seq_bit(a.seq + cw.seq_word_off, i)) {
    v = (int)(int8_t)(uint8_t)(0u - (uint32_t)v);
  }
  return v;
}

// Sixteen scrambling bits starting at element i of the codeword.
__device__ __forceinline__ uint32_t seq_bits16(const uint32_t* __restrict__ seq, uint32_t i)
{
  return seq_bits32(seq, i) & 0xffffu;
}

// Sixteen input bytes from any address: one 128-bit load when aligned, else five aligned words and funnel shifts.
__device__ __forceinline__ uint4 ldg_u128_unaligned(const int8_t* p)
{
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if ((a & 15u) == 0) {
    return __ldg(reinterpret_cast<const uint4*>(p));
  }
  const uint32_t* w  = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t  sh = (uint32_t)(a & 3u) * 8u;
  const uint32_t  w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
  const uint32_t  w4 = sh ? __ldg(w + 4) : 0u;
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                    __funnelshift_r(w3, w4, sh));
}

// UL-SCH stream: every thread step produces sixteen consecutive output soft bits (one 128-bit store) when they come
// from sixteen consecutive inputs of one symbol; runs that touch a symbol with UCI or a symbol boundary go soft bit by
// soft bit.
__global__ void __launch_bounds__(256) ulsch_sch_kernel(UlschArgs a)
{
  __shared__ UlschCodeword cw;
  __shared__ UlschSymbol   S[ULSCH_MAX_SYMBOLS];
  pdl_wait(); // programmatic serialization behind prg_kernel / the soft demapper
  if (threadIdx.x == 0) {
    cw = a.cws[blockIdx.y];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < cw.n_sym * (sizeof(UlschSymbol) / 4); i += blockDim.x) {
    reinterpret_cast<uint32_t*>(S)[i] = reinterpret_cast<const uint32_t*>(a.syms + cw.sym_first)[i];
  }
  __syncthreads();
  if (cw.flags & ULSCH_CW_DEFERRED) {
    return; // descrambled by the rate dematcher while it stages the codeblocks
  }
  const uint32_t  n_sch     = cw.n_out[0];
  const uint32_t  n_chunks  = (n_sch + 15u) / 16u;
  const bool      scrambled = (cw.flags & PDC_CW_SCRAMBLED) != 0;
  const uint32_t* seq       = a.seq + cw.seq_word_off;
  int8_t*         out       = a.sch + cw.sch_off;
  const bool      out_al16  = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += gridDim.x * blockDim.x) {
    const uint32_t     o   = 16u * c;
    const int          s   = ulsch_find_symbol(S, (int)cw.n_sym, 0, o, cw.bpre);
    const UlschSymbol& sym = S[s];
    const uint32_t     rel = o - sym.out_off[0];
    if (out_al16 && sym.list_off[0] == ULSCH_IDENTITY && rel + 16u <= sym.n_out_re[0] * cw.bpre && o + 16u <= n_sch) {
      const uint32_t i = sym.in_off + rel;
      uint4          v = ldg_u128_unaligned(a.in + cw.in_off + i);
      if (scrambled) {
        const uint32_t bits = seq_bits16(seq, i);
        v.x                 = negate4(v.x, bits & 0xfu);
        v.y                 = negate4(v.y, (bits >> 4) & 0xfu);
        v.z                 = negate4(v.z, (bits >> 8) & 0xfu);
        v.w                 = negate4(v.w, bits >> 12);
      }
      *reinterpret_cast<uint4*>(out + o) = v;
    } else {
      for (uint32_t b = 0; b != 16u && o + b < n_sch; ++b) {
        out[o + b] = (int8_t)ulsch_sch_byte(a, cw, S, o + b);
      }
    }
  }
}

// HARQ-ACK, CSI Part 1 and CSI Part 2 streams, one soft bit per thread, with the placeholder rules of
// on_uci_placeholder_1bit / _2bit: with one or two message bits the "x" placeholders (bits 2.. of every modulation
// symbol) get their scrambling back, and with one message bit the "y" placeholder (bit 1) takes the scrambling of bit 0.
__global__ void __launch_bounds__(256) ulsch_uci_kernel(UlschArgs a)
{
  __shared__ UlschCodeword cw;
  __shared__ UlschSymbol   S[ULSCH_MAX_SYMBOLS];
  pdl_wait();
  if (threadIdx.x == 0) {
    cw = a.cws[blockIdx.y];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < cw.n_sym * (sizeof(UlschSymbol) / 4); i += blockDim.x) {
    reinterpret_cast<uint32_t*>(S)[i] = reinterpret_cast<const uint32_t*>(a.syms + cw.sym_first)[i];
  }
  __syncthreads();
  const uint32_t  total     = cw.n_out[1] + cw.n_out[2] + cw.n_out[3];
  const bool      scrambled = (cw.flags & PDC_CW_SCRAMBLED) != 0;
  const uint32_t* seq       = a.seq + cw.seq_word_off;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int          k   = (t < cw.n_out[1]) ? 1 : (t < cw.n_out[1] + cw.n_out[2]) ? 2 : 3;
    const uint32_t     o   = t - cw.uci_base[k];
    const int          s   = ulsch_find_symbol(S, (int)cw.n_sym, k, o, cw.bpre);
    const UlschSymbol& sym = S[s];
    const uint32_t     rel = o - sym.out_off[k];
    const uint32_t     r = rel / cw.bpre, q = rel - r * cw.bpre;
    const uint32_t     e  = a.lists[sym.list_off[k] + r];
    const uint32_t     re = e & 0x7fffu;
    const uint32_t     i  = sym.in_off + re * cw.bpre + q;
    int                v  = (e & ULSCH_PUNCTURED) ? 0 : a.in[cw.in_off + i];
    if (scrambled && seq_bit(seq, i)) {
      v = (int)(int8_t)(uint8_t)(0u - (uint32_t)v);
    }
    const uint32_t nb = sym.uci_bits[k];
    if ((nb == 1u || nb == 2u) && cw.qm > 1u) {
      const uint32_t bq = q % cw.qm; // position inside the modulation symbol
      bool           flip = false;
      if (bq == 1u) {
        flip = (nb == 1u) && (seq_bit(seq, i - 1u) != seq_bit(seq, i));
      } else if (bq >= 2u) {
        flip = seq_bit(seq, i) != 0u;
      }
      if (flip) {
        v = (int)(int8_t)(uint8_t)(0u - (uint32_t)v);
      }
    }
    a.uci[cw.uci_off + t] = (int8_t)v;
  }
}

// For every codeblock of the batch: the deferred codeword (if any) whose UL-SCH space holds its rate-matched soft bits.
__global__ void cb_descramble_map_kernel(const pdc_cb_desc* __restrict__ cbs, uint32_t n_cb,
                                         const UlschCodeword* __restrict__ cws, uint32_t n_cw, uint4* __restrict__ out)
{
  pdl_wait();
  const uint32_t cb = blockIdx.x * blockDim.x + threadIdx.x;
  if (cb >= n_cb) {
    return;
  }
  const uint32_t off = cbs[cb].llr_offset;
  uint4          e   = make_uint4(CB_NOT_SCRAMBLED, 0, 0, 0);
  for (uint32_t c = 0; c != n_cw; ++c) {
    const UlschCodeword& cw = cws[c];
    if ((cw.flags & ULSCH_CW_DEFERRED) && off >= cw.sch_off && off - cw.sch_off < cw.n_out[0]) {
      const uint32_t s = off - cw.sch_off;
      e                = make_uint4(cw.seq_word_off, s, cw.in_off + s, 0);
      break;
    }
  }
  out[cb] = e;
}

inline cudaError_t upload_prg_tables()
{
  static uint32_t h[2][3][128];
  const uint32_t  taps[2] = {PRG_TAPS1, PRG_TAPS2};
  for (int g = 0; g != 2; ++g) {
    uint32_t step = 2u; // x^1
    for (int level = 0; level != 3; ++level) {
      h[g][level][0] = 1u;
      for (int c = 1; c != 128; ++c) {
        h[g][level][c] = prg_mulmod(h[g][level][c - 1], step, taps[g]);
      }
      step = prg_mulmod(h[g][level][127], step, taps[g]); // x^(128^(level+1))
    }
  }
  return cudaMemcpyToSymbol(g_prg_jump, h, sizeof(h));
}

} // namespace pdc
