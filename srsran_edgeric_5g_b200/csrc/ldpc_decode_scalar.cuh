// 5G NR LDPC layered normalized min-sum decoder - general kernel (any base graph, any lifting size).
//
// One CTA per codeblock, one thread per lifted check node j in [0, Z): the Z checks of a base-graph row are independent,
// the rows (layers) are processed in order with one CTA barrier per layer. Soft bits live in shared memory as int8;
// check-to-variable messages are kept compressed per (row, j) as {scaled min1, scaled min2, arg-min, sign bits} - from
// which every message is exactly reconstructible - also in shared memory.
//
// Reference behaviour (bit-exact target): ldpc_decoder_impl::decode and helpers
// (lib/phy/upper/channel_coding/ldpc/ldpc_decoder_impl.cpp:60-318) with the AVX2/AVX512 node kernels
// (ldpc_decoder_avx512.cpp:81-290) or the generic ones (ldpc_decoder_generic.cpp:30-128), selected by scale_mode.
#pragma once

#include "pdc_device.cuh"

namespace pdc {

struct DecodeShape {
  int bg, Z, kb, n_full, rows, n_edges, K;
};

// Shared-memory footprint for lifting size Z (bytes). Layout:
//   soft[n_full * Z] int8 | vbuf[MAX_DEG * Z] int8 | state0[rows * Z] u32 | state1[4 * Z] u32 | bits[ceil(K/32)] u32 |
//   shift[n_edges] u16 | col[n_edges] u8 | misc
__host__ __device__ inline size_t decode_smem_bytes(int bg, int Z)
{
  int    n_full  = (bg == 1) ? 68 : 52;
  int    rows    = (bg == 1) ? 46 : 42;
  int    kb      = (bg == 1) ? 22 : 10;
  size_t soft    = ((size_t)n_full * Z + 15) & ~(size_t)15;
  size_t vbuf    = ((size_t)MAX_DEG * Z + 15) & ~(size_t)15;
  size_t state   = (size_t)(rows + 4) * Z * 4;
  size_t bits    = (size_t)((kb * Z + 31) / 32) * 4 + 16;
  size_t tables  = (size_t)MAX_EDGES * 2 + MAX_EDGES + 64;
  return soft + vbuf + state + bits + ((tables + 15) & ~(size_t)15) + 64;
}

__global__ void ldpc_decode_scalar_kernel(BatchParams prm, const int8_t* direct_in, uint32_t direct_n)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sh_last;     // 1 + index of the last non-zero input LLR
  __shared__ uint32_t sh_crc; // XOR-reduced CRC remainder

  const uint32_t     cb   = blockIdx.x;
  const pdc_cb_desc& d    = prm.cbs[cb];
  const int          tid  = threadIdx.x;
  const int          nthr = blockDim.x;
  if (!(d.flags & PDC_CB_DECODE)) {
    return;
  }
  pdc_cb_result res;
  res.crc_ok  = 0;
  res.iters   = d.max_iter;
  res.status  = 0;
  res.nlayers = 0;

  const int bg = d.base_graph;
  const int Z  = d.lifting_size;
  if ((bg != 1 && bg != 2) || Z < 2 || Z > MAX_Z || c_tab.set_index[Z] == 0xff || d.max_iter == 0 ||
      d.harq_id >= prm.harq_entries || d.crc_kind > PDC_CRC24B || (int)d.nof_filler >= ((bg == 1) ? 22 : 10) * Z) {
    if (tid == 0) {
      res.status      = 2;
      prm.results[cb] = res;
    }
    return;
  }
  const int b       = bg - 1;
  const int kb      = (bg == 1) ? 22 : 10;
  const int n_full  = (bg == 1) ? 68 : 52;
  const int n_short = n_full - 2;
  const int rows    = (bg == 1) ? 46 : 42;
  const int n_edges = (bg == 1) ? BG1_EDGES_N : BG2_EDGES_N;
  const int K       = kb * Z;
  const int N       = n_short * Z;
  const int F       = d.nof_filler;
  const int set     = c_tab.set_index[Z];

  // Carve shared memory.
  size_t   off    = 0;
  int8_t*  soft   = reinterpret_cast<int8_t*>(smem_raw);
  off += ((size_t)n_full * Z + 15) & ~(size_t)15;
  int8_t* vbuf = reinterpret_cast<int8_t*>(smem_raw + off);
  off += ((size_t)MAX_DEG * Z + 15) & ~(size_t)15;
  uint32_t* state0 = reinterpret_cast<uint32_t*>(smem_raw + off);
  off += (size_t)rows * Z * 4;
  uint32_t* state1 = reinterpret_cast<uint32_t*>(smem_raw + off);
  off += (size_t)4 * Z * 4;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem_raw + off);
  off += (size_t)((K + 31) / 32) * 4 + 16;
  uint16_t* e_shift = reinterpret_cast<uint16_t*>(smem_raw + off);
  off += (size_t)MAX_EDGES * 2;
  uint8_t* e_col = smem_raw + off;

  // Input: the HARQ entry after dematching (N soft bits), or a caller-provided LLR vector (single-codeblock API).
  const int8_t* in   = direct_in ? direct_in : prm.harq + (size_t)d.harq_id * PDC_MAX_CB_SOFT;
  const int     n_in = direct_in ? (int)direct_n : N;

  if (tid == 0) {
    sh_last = 0;
  }
  // Lifted shifts of this (base graph, Z): V mod Z (ldpc_luts_impl.cpp:4536-4541).
  for (int i = tid; i < n_edges; i += nthr) {
    e_shift[i] = (uint16_t)(c_tab.v[b][set][i] % Z);
    e_col[i]   = c_tab.col[b][i];
  }
  for (int i = tid; i < (rows + 4) * Z; i += nthr) {
    state0[i] = 0; // "check-to-variable messages not initialised" = all-zero messages
  }
  __syncthreads();

  // load_soft_bits (ldpc_decoder_impl.cpp:149-184) + search of the last non-zero LLR (:86-99).
  {
    int full  = (n_in / Z) * Z;
    int last  = 0;
    int total = n_full * Z;
    for (int i = tid; i < total; i += nthr) {
      int k = i - 2 * Z;
      int v = 0;
      if (k >= 0 && k < n_in) {
        v = in[k];
        if (v != 0) {
          last = k + 1;
        }
        if (k < full) {
          v = max(-CLAMP_IN, min(CLAMP_IN, v));
        }
      }
      soft[i] = (int8_t)v;
    }
    // Block-wide maximum.
    for (int o = 16; o > 0; o >>= 1) {
      last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    }
    if ((tid & 31) == 0 && last > 0) {
      atomicMax(&sh_last, last);
    }
  }
  __syncthreads();
  const int trimmed = sh_last;
  uint8_t*  out     = prm.cb_bits + (size_t)cb * PDC_MAX_CB_BYTES;
  uint8_t*  out_h   = prm.harq_data + (size_t)d.harq_id * PDC_MAX_CB_BYTES;
  const int crc_kind   = d.crc_kind;
  const bool early     = (d.flags & PDC_CB_EARLY_STOP) && (crc_kind != PDC_CRC_NONE);
  const int  n_words   = (K + 31) / 32;
  if (trimmed == 0) {
    // All-zero input: not decodable; with no CRC calculator the output is set to all ones (:88-94).
    if (crc_kind == PDC_CRC_NONE) {
      for (int i = tid; i < (K + 7) / 8; i += nthr) {
        int rem = K - 8 * i;
        out[i]  = (rem >= 8) ? 0xff : (uint8_t)(0xff << (8 - rem));
      }
    }
    if (tid == 0) {
      res.status      = 1;
      prm.results[cb] = res;
    }
    return;
  }
  int cb_len = max(trimmed + 2 * Z, K + 4 * Z);
  cb_len     = ((cb_len + Z - 1) / Z) * Z;
  const int layers = cb_len / Z - kb;
  res.nlayers      = (uint8_t)layers;

  const int  j      = tid;
  const bool active = j < Z;
  const int  scale_mode = prm.scale_mode;
  const uint16_t* row_start = c_tab.row_start[b];

  int  iters_done = d.max_iter;
  bool crc_ok     = false;
  for (int it = 0; it < d.max_iter; ++it) {
    for (int m = 0; m < layers; ++m) {
      if (active) {
        const int e0  = row_start[m];
        const int deg = row_start[m + 1] - e0;
        uint32_t  st  = state0[m * Z + j];
        uint32_t  sg  = st >> 19;
        if (m < 4) {
          sg |= state1[m * Z + j] << 13;
        }
        const int om1 = st & 0x7f, om2 = (st >> 7) & 0x7f, oarg = (st >> 14) & 0x1f;
        int       min1 = LLR_MAX, min2 = LLR_MAX, arg = 0;
        uint32_t  par = 0, vneg = 0;
        for (int e = 0; e < deg; ++e) {
          int pos = j + e_shift[e0 + e];
          pos     = (pos >= Z) ? pos - Z : pos;
          int s   = soft[e_col[e0 + e] * Z + pos];
          int mag = (e == oarg) ? om2 : om1;
          int c   = ((sg >> e) & 1u) ? -mag : mag;
          int v   = max(-LLR_MAX, min(LLR_MAX, s - c));
          v       = (s >= LLR_INF) ? LLR_INF : v;
          v       = (s <= -LLR_INF) ? -LLR_INF : v;
          vbuf[e * Z + j] = (int8_t)v;
          int a   = abs(v);
          if (a < min2) {
            min2 = (a < min1) ? min1 : a;
          }
          if (a < min1) {
            min1 = a;
            arg  = e;
          }
          uint32_t ng = (v < 0) ? 1u : 0u;
          par ^= ng;
          vneg |= ng << e;
        }
        const int s1 = scale_c2v(min1, scale_mode);
        const int s2 = scale_c2v(min2, scale_mode);
        // Sign of each new message: parity of the row XOR the sign of its own v2c.
        uint32_t nsg = vneg ^ (par ? ((1u << deg) - 1u) : 0u);
        for (int e = 0; e < deg; ++e) {
          int pos = j + e_shift[e0 + e];
          pos     = (pos >= Z) ? pos - Z : pos;
          int v   = vbuf[e * Z + j];
          int mag = (e == arg) ? s2 : s1;
          int c   = ((nsg >> e) & 1u) ? -mag : mag;
          int r   = v + c;
          r       = (r > LLR_MAX) ? LLR_INF : r;
          r       = (r < -LLR_MAX) ? -LLR_INF : r;
          r       = (v > LLR_MAX) ? LLR_INF : r;
          r       = (v < -LLR_MAX) ? -LLR_INF : r;
          soft[e_col[e0 + e] * Z + pos] = (int8_t)r;
        }
        state0[m * Z + j] = (uint32_t)s1 | ((uint32_t)s2 << 7) | ((uint32_t)arg << 14) | (nsg << 19);
        if (m < 4) {
          state1[m * Z + j] = nsg >> 13;
        }
      }
      __syncthreads();
    }

    const bool last_it = (it + 1 == d.max_iter);
    if (early || last_it) {
      // get_hard_bits (:126-134): bit = soft <= 0, MSB first; any zero among the K message soft bits blocks early stop.
      int any_zero = 0;
      for (int w = tid >> 5; w < n_words; w += nthr >> 5) {
        int      i  = 32 * w + (tid & 31);
        int      s  = (i < K) ? soft[i] : 1;
        uint32_t bw = __brev(__ballot_sync(0xffffffffu, s <= 0));
        any_zero |= (s == 0);
        if ((tid & 31) == 0) {
          bits[w] = bw;
        }
      }
      if (tid == 0) {
        sh_crc = 0;
      }
      any_zero = __syncthreads_or(any_zero);
      bool pass = false;
      if (crc_kind != PDC_CRC_NONE) {
        // M(x) mod P == 0 over the first K - F bits: word t contributes W_t(x) * x^(32 (T-1-t)) (the message is
        // left-aligned, multiplying by a power of x does not change whether the remainder is zero).
        const int      nb    = K - F;
        const int      T     = (nb + 31) / 32;
        const uint32_t poly  = crc_poly(crc_kind);
        const int      order = crc_order(crc_kind);
        uint32_t       acc   = 0;
        for (int t = tid; t < T; t += nthr) {
          uint32_t w = bits[t];
          if (t == T - 1 && (nb & 31)) {
            w &= 0xffffffffu << (32 - (nb & 31));
          }
          acc ^= gf2_mulmod(w, c_tab.xpow32[crc_kind - 1][T - 1 - t], poly, order);
        }
        for (int o = 16; o > 0; o >>= 1) {
          acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if ((tid & 31) == 0 && acc) {
          atomicXor(&sh_crc, acc);
        }
        __syncthreads();
        pass = (sh_crc == 0);
        __syncthreads(); // sh_crc is reset on the next iteration
      }
      if (early) {
        if (pass && !any_zero) {
          crc_ok     = true;
          iters_done = it + 1;
          break;
        }
      } else if (last_it) {
        crc_ok = pass;
      }
    }
  }

  // Decoded bits, packed MSB first.
  for (int w = tid; w < n_words; w += nthr) {
    uint32_t v = bits[w];
    int      base = 4 * w;
    int      nbytes = (K + 7) / 8;
#pragma unroll
    for (int k = 0; k != 4; ++k) {
      if (base + k < nbytes) {
        out[base + k]   = (uint8_t)(v >> (24 - 8 * k));
        out_h[base + k] = (uint8_t)(v >> (24 - 8 * k));
      }
    }
  }
  if (tid == 0) {
    res.crc_ok      = crc_ok ? 1 : 0;
    res.iters       = (uint8_t)iters_done;
    prm.results[cb] = res;
  }
}

// Per device, once per context (see h2_configure_device).
inline cudaError_t scalar_configure_device()
{
  return cudaFuncSetAttribute(ldpc_decode_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)decode_smem_bytes(1, MAX_Z));
}

inline cudaError_t launch_ldpc_decode(const BatchParams& p, int max_Z, int max_bg_rows68, const int8_t* direct_in,
                                      uint32_t direct_n, cudaStream_t s)
{
  if (p.n_cb == 0) {
    return cudaSuccess;
  }
  int    bg      = max_bg_rows68 ? 1 : 2;
  size_t smem    = decode_smem_bytes(bg, max_Z);
  int    threads = ((max_Z + 31) / 32) * 32;
  ldpc_decode_scalar_kernel<<<p.n_cb, threads, smem, s>>>(p, direct_in, direct_n);
  return cudaGetLastError();
}

} // namespace pdc
