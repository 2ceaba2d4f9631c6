// Downlink twin of the decode path: 5G NR LDPC encoding (TS 38.212 5.3.2) fused with rate matching
// (bit selection + interleaving, TS 38.212 5.4.2), one CTA per codeblock.
//
// Reference: ldpc_encoder_impl::encode (lib/phy/upper/channel_coding/ldpc/ldpc_encoder_impl.cpp:30-80, the systematic /
// high-rate / extended regions of ldpc_encoder_generic.cpp / _avx2.cpp) and ldpc_rate_matcher_impl::rate_match
// (ldpc_rate_matcher_impl.cpp:30-150: select_bits, interleave_bits). Pure GF(2) work: bit-exact by construction; the
// oracle's orc_ldpc_encode / orc_rate_match (pinned against the compiled reference) are the checkers.
//
// Thread j owns lifted position j of every variable node; the codeblock lives in shared memory one bit per byte
// (68 x 384 bytes at most). Rows 0-3 give the four core parity nodes (the sum of the four rows cancels everything but one
// rotation of the first, the others follow by back-substitution); every extension row is then independent of the others
// and only the rows the rate matcher will read are computed. The rate matcher is a gather: output bit o of the
// interleaved sequence finds its position in the circular buffer in closed form (the filler bits are a gap in it).
// The base graph is walked in constant memory with uniform indices (one broadcast per edge); taking the edges from the
// decoder's per-shape images instead (shifts already reduced modulo Z), from global or staged in shared memory, measured
// slower: the kernel then needs 39-54 registers instead of 32 and loses a resident CTA per SM (16 transport blocks:
// 271 -> 343-382 us).
#pragma once

#include "pdc_device.cuh"
#include "tables.cuh"

namespace pdc {

struct EncodeParams {
  const pdc_enc_desc* cbs;
  uint32_t            n_cb;
  const uint8_t*      msgs; // packed message bits, MSB first
  uint8_t*            out;  // one bit per byte
  uint32_t            out_capacity;
};

constexpr int ENC_MAX_THREADS = 384;

__device__ __forceinline__ int enc_wrap(int t, int Z)
{
  return (t >= Z) ? t - Z : t;
}

// mode 0: encode + rate match (E bits at out_offset, one per byte, or packed for PDC_ENC_PACKED); mode 1: the whole codeword, N = (n_full - 2) Z bits at out_offset
// (ldpc_encoder::encode alone).
__global__ void __launch_bounds__(ENC_MAX_THREADS) ldpc_encode_rm_kernel(EncodeParams prm, int mode)
{
  extern __shared__ __align__(16) unsigned char enc_smem[];
  __shared__ int sh_ok;
  const int           tid = threadIdx.x, nthr = blockDim.x;
  const pdc_enc_desc  d   = prm.cbs[blockIdx.x];
  const int           bg = d.base_graph, Z = d.lifting_size, b = bg - 1;
  if (tid == 0) {
    sh_ok = ((bg == 1 || bg == 2) && Z >= 2 && Z <= MAX_Z && c_tab.set_index[Z] != 0xff) ? 1 : 0;
  }
  __syncthreads();
  if (!sh_ok) {
    return; // the host validated the descriptors; nothing is written for an invalid one
  }
  const int kb = (bg == 1) ? 22 : 10, n_full = (bg == 1) ? 68 : 52, rows = (bg == 1) ? 46 : 42;
  const int K = kb * Z, N = (n_full - 2) * Z, set = c_tab.set_index[Z];
  uint8_t*  c   = enc_smem;                 // n_full * Z bits
  uint8_t*  lam = enc_smem + n_full * Z;    // 4 * Z: information part of the four core rows

  // Rate-matching geometry (ldpc_rate_matcher_impl.cpp:52-90).
  const int F     = d.nof_filler;
  const int Ncb   = (d.nref > 0) ? min((int)d.nref, N) : N;
  const int K_sys = (kb - 2) * Z;
  const int qm    = max(1, (int)d.qm);
  const int E     = (mode == 0) ? (int)d.rm_length : N;
  int       k0    = 0;
  {
    const int sf = (bg == 1) ? ((d.rv == 1) ? 17 : (d.rv == 2) ? 33 : (d.rv == 3) ? 56 : 0)
                             : ((d.rv == 1) ? 13 : (d.rv == 2) ? 25 : (d.rv == 3) ? 43 : 0);
    k0 = (int)floor((double)sf * (double)Ncb / (double)N) * Z;
  }
  // Rows whose parity the output can reach: everything up to Ncb if the walk wraps, else up to its end.
  int rows_needed = rows;
  if (mode == 0 && k0 + E + F < Ncb) {
    rows_needed = max(4, min(rows, (k0 + E + F + 2 * Z + Z - 1) / Z - kb));
  } else if (mode == 0) {
    rows_needed = max(4, min(rows, (Ncb + 2 * Z + Z - 1) / Z - kb));
  }

  // Message bits (fillers are zeros in the input), parity nodes cleared.
  const uint8_t* msg = prm.msgs + d.msg_offset;
  for (int i = tid; i < K; i += nthr) {
    c[i] = (uint8_t)((msg[i >> 3] >> (7 - (i & 7))) & 1u);
  }
  for (int i = K + tid; i < (kb + rows_needed) * Z; i += nthr) {
    c[i] = 0;
  }
  __syncthreads();

  const int  j      = tid;
  const bool active = j < Z;
  // Core rows, information columns only.
  if (active) {
    for (int m = 0; m != 4; ++m) {
      uint8_t acc = 0;
      for (int i = c_tab.row_start[b][m]; i != c_tab.row_start[b][m + 1]; ++i) {
        const int col = c_tab.col[b][i];
        if (col < kb) {
          acc ^= c[col * Z + enc_wrap(j + c_tab.v[b][set][i] % Z, Z)];
        }
      }
      lam[m * Z + j] = acc;
    }
  }
  __syncthreads();
  // First core parity node: column kb appears in three core rows, two of them with the same shift; the sum of the four
  // rows is one rotation of it.
  {
    int sh[3] = {0, 0, 0}, n_sh = 0;
    for (int i = c_tab.row_start[b][0]; i != c_tab.row_start[b][4]; ++i) {
      if (c_tab.col[b][i] == kb && n_sh < 3) {
        sh[n_sh++] = c_tab.v[b][set][i] % Z;
      }
    }
    const int dsh = (sh[0] == sh[1]) ? sh[2] : ((sh[0] == sh[2]) ? sh[1] : sh[0]);
    if (active) {
      c[kb * Z + enc_wrap(j + dsh, Z)] = lam[j] ^ lam[Z + j] ^ lam[2 * Z + j] ^ lam[3 * Z + j];
    }
  }
  __syncthreads();
  // Back-substitution: three rounds, each taking the core rows with exactly one unknown core parity node.
  {
    int known = 1; // bit k: core parity node kb + k is known
    for (int round = 0; round != 3; ++round) {
      int solved = 0;
      for (int m = 0; m != 4; ++m) {
        int unknown = -1, n_unknown = 0, unknown_shift = 0;
        for (int i = c_tab.row_start[b][m]; i != c_tab.row_start[b][m + 1]; ++i) {
          const int col = c_tab.col[b][i];
          if (col >= kb && col < kb + 4 && !((known >> (col - kb)) & 1)) {
            unknown       = col;
            unknown_shift = c_tab.v[b][set][i] % Z;
            ++n_unknown;
          }
        }
        if (n_unknown != 1) {
          continue;
        }
        if (active) {
          uint8_t acc = lam[m * Z + j];
          for (int i = c_tab.row_start[b][m]; i != c_tab.row_start[b][m + 1]; ++i) {
            const int col = c_tab.col[b][i];
            if (col >= kb && col < kb + 4 && ((known >> (col - kb)) & 1)) {
              acc ^= c[col * Z + enc_wrap(j + c_tab.v[b][set][i] % Z, Z)];
            }
          }
          c[unknown * Z + enc_wrap(j + unknown_shift, Z)] = acc;
        }
        solved |= 1 << (unknown - kb);
        // Rows solved in the same round do not depend on each other's result only if they target different nodes and
        // read nodes known before the round: make every solved node visible before it is used.
        __syncthreads();
        known |= solved;
      }
    }
  }
  __syncthreads();
  // Extension rows: a single identity column kb + m each.
  if (active) {
    for (int m = 4; m < rows_needed; ++m) {
      uint8_t acc = 0;
      for (int i = c_tab.row_start[b][m]; i != c_tab.row_start[b][m + 1]; ++i) {
        const int col = c_tab.col[b][i];
        if (col != kb + m) {
          acc ^= c[col * Z + enc_wrap(j + c_tab.v[b][set][i] % Z, Z)];
        }
      }
      c[(kb + m) * Z + j] = acc;
    }
  }
  __syncthreads();

  uint8_t* out = prm.out + d.out_offset;
  if (mode == 1) {
    for (int i = tid; i < N; i += nthr) {
      out[i] = c[2 * Z + i];
    }
    return;
  }
  // Bit selection + interleaving as a gather. The circular buffer [0, Ncb) without the filler positions
  // [K_sys - F, K_sys) has L positions; the walk starts at k0 (or right after the fillers if k0 falls on one).
  const int f_lo = min(K_sys - F, Ncb), f_hi = min(K_sys, Ncb), Fp = f_hi - f_lo;
  const int L    = Ncb - Fp;
  int       k0e  = k0 % Ncb;
  if (k0e >= f_lo && k0e < f_hi) {
    k0e = f_hi % Ncb;
  }
  const int c0  = (k0e < f_lo) ? k0e : k0e - Fp; // compressed coordinate of the start
  const int per = E / qm;
  if (d.flags & PDC_ENC_PACKED) {
    // Eight output bits per thread step, first bit in the most significant bit; the last byte is zero padded.
    const int nbytes = (E + 7) >> 3;
    for (int bi = tid; bi < nbytes; bi += nthr) {
      const int o0 = 8 * bi;
      int       i = o0 / qm, jj = o0 - i * qm;
      uint32_t  v = 0;
#pragma unroll 1
      for (int bit = 0; bit != 8; ++bit) {
        if (o0 + bit < E) {
          const int k   = jj * per + i;
          const int ci  = (c0 + k) % L;
          const int pos = (ci < f_lo) ? ci : ci + Fp;
          v |= (uint32_t)(c[2 * Z + pos] & 1u) << (7 - bit);
        }
        if (++jj == qm) {
          jj = 0;
          ++i;
        }
      }
      out[bi] = (uint8_t)v;
    }
    return;
  }
  for (int o = tid; o < E; o += nthr) {
    const int i = o / qm, jj = o - i * qm;
    const int k = jj * per + i;               // index in the selected (not yet interleaved) sequence
    int       ci = (c0 + k) % L;
    const int pos = (ci < f_lo) ? ci : ci + Fp;
    out[o] = c[2 * Z + pos];
  }
}

inline size_t enc_smem_bytes(int bg, int Z)
{
  return (size_t)((bg == 1 ? 68 : 52) + 4) * Z + 16;
}

} // namespace pdc
