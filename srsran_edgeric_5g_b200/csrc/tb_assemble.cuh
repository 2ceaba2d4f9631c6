// Transport-block assembly: codeblock concatenation + TB CRC (CRC24A) on the device.
//
// Reference behaviour: pusch_decoder_impl::join_and_notify / concatenate_codeblocks
// (lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.cpp:384-450, :452-497): with one codeblock the TB CRC is the
// codeblock CRC; with several, and only if every codeblock CRC passed, the payloads (without CB CRC, filler and zero
// padding) are concatenated and CRC24A over the TB must equal the 24 bits that follow the payload of the last codeblock.
//
// TB_SPLIT CTAs per transport block. The concatenated stream (payload + TB checksum) is produced 32 bits per thread
// with a funnel shift from at most two codeblocks; each thread then runs a table-driven CRC24A over its run of
// consecutive words and weights it by x^(32 * words after the run) mod P; the XOR of all weighted remainders (across the
// CTAs of the TB through a global accumulator, last CTA publishes) is zero iff the CRC matches.
#pragma once

#include "pdc_device.cuh"

namespace pdc {

constexpr int TB_THREADS = 512;
constexpr int TB_SPLIT   = 16;

// Two-level table of x^(32 e) mod CRC24A, e = 256 hi + lo (transport blocks of up to 65536 words), and the byte table
// T[b] = (b * x^24) mod P; filled at start-up.
__device__ uint32_t   g_xpow_crc24a_hi[256]; // x^(32 * 256 * i)
__device__ uint32_t   g_xpow_crc24a_lo[256]; // x^(32 * i)
__constant__ uint32_t c_crc24a_table[256];

// 32 bits of a codeblock's decoded message starting at bit `off` (MSB first), zero beyond the buffer.
__device__ __forceinline__ uint32_t cb_bits32(const uint8_t* __restrict__ src, uint32_t off)
{
  const uint32_t* w  = reinterpret_cast<const uint32_t*>(src);
  uint32_t        i  = off >> 5, sh = off & 31u;
  uint32_t        hi = __byte_perm(w[i], 0, 0x0123); // bytes in memory are MSB first
  if (sh == 0) {
    return hi;
  }
  uint32_t lo = (i + 1 < PDC_MAX_CB_BYTES / 4) ? __byte_perm(w[i + 1], 0, 0x0123) : 0u;
  return __funnelshift_l(lo, hi, sh);
}

// sync: per TB {XOR accumulator, arrival counter}, zero between launches (the last CTA of a TB resets them).
__global__ void __launch_bounds__(TB_THREADS) tb_assemble_kernel(TbParams prm, const uint8_t* harq_data, uint32_t* sync)
{
  __shared__ uint32_t sh_crc;
  __shared__ uint32_t sh_table[256];
  const pdc_tb_desc&  tb  = prm.tbs[blockIdx.x];
  const int           tid = threadIdx.x;
  int                 ok  = 1;
  if (tid < 256) {
    sh_table[tid] = c_crc24a_table[tid];
  }
  if (tid == 0) {
    sh_crc = 0;
  }
  pdl_wait(); // launched behind the decoder with programmatic serialization: its results are read from here on
  for (uint32_t i = tid; i < tb.nof_cb; i += blockDim.x) {
    const pdc_cb_desc& d = prm.cbs[tb.first_cb + i];
    if ((d.flags & PDC_CB_DECODE) && !prm.cb_results[tb.first_cb + i].crc_ok) {
      ok = 0;
    }
  }
  ok = __syncthreads_and(ok);
  pdc_tb_result r;
  r.tb_crc_ok  = 0;
  r.all_cb_ok  = (uint8_t)ok;
  r.reserved   = 0;
  uint8_t* out = prm.tb_bytes + tb.out_offset;
  const uint32_t part = blockIdx.y, n_parts = gridDim.y;
  if (ok && tb.nof_cb == 1) {
    if (part == 0) {
      const uint8_t* src = harq_data + (size_t)prm.cbs[tb.first_cb].harq_id * PDC_MAX_CB_BYTES;
      for (uint32_t i = tid; i < tb.tbs_bits / 8; i += blockDim.x) {
        out[i] = src[i];
      }
    }
    r.tb_crc_ok = 1;
  } else if (ok) {
    const pdc_cb_desc& d0     = prm.cbs[tb.first_cb];
    const uint32_t     K      = ((d0.base_graph == 1) ? 22u : 10u) * d0.lifting_size;
    const uint32_t     n_data = K - 24u - d0.nof_filler;
    const uint32_t     total  = tb.tbs_bits + 24u; // payload followed by the TB checksum
    const uint32_t     T      = (total + 31u) / 32u;
    uint32_t*          out_w  = reinterpret_cast<uint32_t*>(out);
    const uint32_t     poly   = crc_poly(PDC_CRC24A);
    // Thread t of part p owns words [(p * blockDim + t) * per, ...).
    const uint32_t n_thr = blockDim.x * n_parts;
    const uint32_t per   = (T + n_thr - 1) / n_thr;
    const uint32_t first = (part * blockDim.x + tid) * per;
    const uint32_t last  = min(T, first + per);
    uint32_t       crc   = 0;
    if (first < T) {
      uint32_t q   = 32u * first;
      uint32_t cb  = q / n_data;
      uint32_t off = q - cb * n_data;
      for (uint32_t t = first; t < last; ++t) {
        PDC_ASSERT(cb < tb.nof_cb && off < n_data && n_data <= 8448u);
        const uint8_t* src = harq_data + (size_t)prm.cbs[tb.first_cb + cb].harq_id * PDC_MAX_CB_BYTES;
        uint32_t       w   = cb_bits32(src, off);
        uint32_t       rem = n_data - off; // bits left in this codeblock
        if (rem < 32u) {
          // The word continues in the next codeblock (the last codeblock never runs out before `total`).
          w &= 0xffffffffu << (32u - rem);
          if (cb + 1 < tb.nof_cb) {
            const uint8_t* nxt = harq_data + (size_t)prm.cbs[tb.first_cb + cb + 1].harq_id * PDC_MAX_CB_BYTES;
            w |= cb_bits32(nxt, 0) >> rem;
          }
          cb += 1;
          off = 32u - rem;
        } else {
          off += 32u;
          if (off == n_data) {
            cb += 1;
            off = 0;
          }
        }
        if (t == T - 1 && (total & 31u)) {
          w &= 0xffffffffu << (32u - (total & 31u));
        }
        out_w[t] = __byte_perm(w, 0, 0x0123); // big-endian bit order -> byte order in memory
#pragma unroll
        for (int k = 3; k >= 0; --k) {
          uint32_t byte = (w >> (8 * k)) & 0xffu;
          crc           = ((crc << 8) ^ sh_table[((crc >> 16) ^ byte) & 0xffu]) & 0xffffffu;
        }
      }
      // Weight by x^(32 * words after this run) mod P.
      const uint32_t e = T - last;
      crc              = gf2_mulmod(crc, __ldg(&g_xpow_crc24a_hi[(e >> 8) & 255u]), poly, 24);
      crc              = gf2_mulmod(crc, __ldg(&g_xpow_crc24a_lo[e & 255u]), poly, 24);
    }
    for (int o = 16; o > 0; o >>= 1) {
      crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
    }
    if ((tid & 31) == 0 && crc) {
      atomicXor(&sh_crc, crc);
    }
    __syncthreads();
  }
  // Combine the parts: XOR the CTA's remainder into the TB accumulator; the last CTA to arrive publishes and resets.
  if (tid == 0) {
    uint32_t* acc = sync + 2 * blockIdx.x;
    if (sh_crc) {
      atomicXor(acc, sh_crc);
    }
    __threadfence();
    if (atomicAdd(acc + 1, 1u) == n_parts - 1) {
      __threadfence();
      const uint32_t total_crc = atomicExch(acc, 0u);
      acc[1]                   = 0;
      if (ok && tb.nof_cb > 1) {
        r.tb_crc_ok = (total_crc == 0) ? 1 : 0;
      }
      prm.tb_results[blockIdx.x] = r;
    }
  }
}

inline cudaError_t upload_tb_tables()
{
  static uint32_t lo[256], hi[256];
  uint32_t        poly = crc_poly(PDC_CRC24A);
  uint32_t        x32  = 1; // x^32 mod P
  for (int b = 0; b != 32; ++b) {
    x32 <<= 1;
    if (x32 & (1u << 24)) {
      x32 ^= poly;
    }
  }
  lo[0] = 1;
  for (int i = 1; i != 256; ++i) {
    lo[i] = gf2_mulmod(lo[i - 1], x32, poly, 24);
  }
  const uint32_t x8192 = gf2_mulmod(lo[255], x32, poly, 24); // x^(32 * 256)
  hi[0]                = 1;
  for (int i = 1; i != 256; ++i) {
    hi[i] = gf2_mulmod(hi[i - 1], x8192, poly, 24);
  }
  cudaError_t e = cudaMemcpyToSymbol(g_xpow_crc24a_lo, lo, sizeof(lo));
  if (e != cudaSuccess) {
    return e;
  }
  e = cudaMemcpyToSymbol(g_xpow_crc24a_hi, hi, sizeof(hi));
  if (e != cudaSuccess) {
    return e;
  }
  uint32_t table[256];
  for (uint32_t b = 0; b != 256; ++b) {
    uint32_t v = b << 16;
    for (int k = 0; k != 8; ++k) {
      v <<= 1;
      if (v & (1u << 24)) {
        v ^= poly;
      }
    }
    table[b] = v & 0xffffffu;
  }
  return cudaMemcpyToSymbol(c_crc24a_table, table, sizeof(table));
}

inline cudaError_t launch_tb_assemble(const TbParams& p, const uint8_t* harq_data, uint32_t* sync, cudaStream_t s)
{
  if (p.n_tb == 0) {
    return cudaSuccess;
  }
  return launch_pdl(tb_assemble_kernel, dim3(p.n_tb, TB_SPLIT), dim3(TB_THREADS), 0, s, p, harq_data, sync);
}

} // namespace pdc
