// Transport-block assembly: codeblock concatenation + TB CRC (CRC24A) on the device.
//
// Reference behaviour: pusch_decoder_impl::join_and_notify / concatenate_codeblocks
// (lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.cpp:384-450, :452-497): with one codeblock the TB CRC is the
// codeblock CRC; with several, and only if every codeblock CRC passed, the payloads (without CB CRC, filler and zero
// padding) are concatenated and CRC24A over the TB must equal the 24 bits that follow the payload of the last codeblock.
#pragma once

#include "pdc_device.cuh"

namespace pdc {

// x^(32 * 2^i) mod CRC24A, i = 0..19, filled at start-up.
__constant__ uint32_t c_xpow_crc24a_pow2[20];

__device__ __forceinline__ uint32_t tb_stream_byte(const uint8_t* __restrict__ data, const pdc_cb_desc* cbs,
                                                   uint32_t first_cb, uint32_t n_data, uint32_t q)
{
  // 8 bits of the concatenated stream starting at bit q.
  uint32_t v = 0;
#pragma unroll
  for (int k = 0; k != 8; ++k) {
    uint32_t        qq  = q + k;
    uint32_t        cb  = qq / n_data;
    uint32_t        off = qq - cb * n_data;
    const uint8_t*  src = data + (size_t)cbs[first_cb + cb].harq_id * PDC_MAX_CB_BYTES;
    v                   = (v << 1) | ((src[off >> 3] >> (7 - (off & 7))) & 1u);
  }
  return v;
}

__global__ void __launch_bounds__(256) tb_assemble_kernel(TbParams prm, const uint8_t* harq_data)
{
  __shared__ uint32_t sh_crc;
  const pdc_tb_desc&  tb  = prm.tbs[blockIdx.x];
  const int           tid = threadIdx.x;
  int                 ok  = 1;
  for (uint32_t i = tid; i < tb.nof_cb; i += blockDim.x) {
    const pdc_cb_desc& d = prm.cbs[tb.first_cb + i];
    if ((d.flags & PDC_CB_DECODE) && !prm.cb_results[tb.first_cb + i].crc_ok) {
      ok = 0;
    }
  }
  if (tid == 0) {
    sh_crc = 0;
  }
  ok = __syncthreads_and(ok);
  pdc_tb_result r;
  r.tb_crc_ok = 0;
  r.all_cb_ok = (uint8_t)ok;
  r.reserved  = 0;
  uint8_t* out = prm.tb_bytes + tb.out_offset;
  if (ok && tb.nof_cb == 1) {
    const uint8_t* src = harq_data + (size_t)prm.cbs[tb.first_cb].harq_id * PDC_MAX_CB_BYTES;
    for (uint32_t i = tid; i < tb.tbs_bits / 8; i += blockDim.x) {
      out[i] = src[i];
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
    uint32_t           acc    = 0;
    for (uint32_t t = tid; t < T; t += blockDim.x) {
      uint32_t w = 0;
#pragma unroll
      for (int k = 0; k != 4; ++k) {
        uint32_t q = 32u * t + 8u * k;
        uint32_t b = (q < total) ? tb_stream_byte(harq_data, prm.cbs, tb.first_cb, n_data, q) : 0u;
        w          = (w << 8) | b;
      }
      if (t == T - 1 && (total & 31u)) {
        w &= 0xffffffffu << (32u - (total & 31u));
      }
      out_w[t] = __byte_perm(w, 0, 0x0123); // big-endian bit order -> byte order in memory
      // x^(32 (T-1-t)) mod P by square-and-multiply over the precomputed x^(32 2^i).
      uint32_t e  = T - 1 - t;
      uint32_t xp = 1;
      for (int i = 0; e != 0; ++i, e >>= 1) {
        if (e & 1u) {
          xp = gf2_mulmod(xp, c_xpow_crc24a_pow2[i], poly, 24);
        }
      }
      acc ^= gf2_mulmod(w, xp, poly, 24);
    }
    for (int o = 16; o > 0; o >>= 1) {
      acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if ((tid & 31) == 0 && acc) {
      atomicXor(&sh_crc, acc);
    }
    __syncthreads();
    r.tb_crc_ok = (sh_crc == 0) ? 1 : 0;
  }
  if (tid == 0) {
    prm.tb_results[blockIdx.x] = r;
  }
}

inline cudaError_t upload_tb_tables()
{
  uint32_t h[20];
  uint32_t poly = crc_poly(PDC_CRC24A);
  uint32_t x    = 1;
  for (int b = 0; b != 32; ++b) {
    x <<= 1;
    if (x & (1u << 24)) {
      x ^= poly;
    }
  }
  for (int i = 0; i != 20; ++i) {
    h[i] = x;
    x    = gf2_mulmod(x, x, poly, 24);
  }
  return cudaMemcpyToSymbol(c_xpow_crc24a_pow2, h, sizeof(h));
}

inline cudaError_t launch_tb_assemble(const TbParams& p, const uint8_t* harq_data, cudaStream_t s)
{
  if (p.n_tb == 0) {
    return cudaSuccess;
  }
  tb_assemble_kernel<<<p.n_tb, 256, 0, s>>>(p, harq_data);
  return cudaGetLastError();
}

} // namespace pdc
