// Shared device-side definitions of the PUSCH decode path (sm_100a).
#pragma once

#include "../../include/pusch_dec_cuda.h"
#include <cuda_runtime.h>
#include <stdint.h>

// -DPDC_DEBUG_BOUNDS: every shared-memory / global index the kernels form is checked on the device (a violation prints
// the site and traps: the launch fails, the test sees it), and the host puts canary words behind every device
// allocation (pdc_debug_canaries_ok). compute-sanitizer is not available on the pool the GPU tests run on, so the GPU
// suite is run once per round against this build instead (tools/run_debug_bounds.sh).
#ifdef PDC_DEBUG_BOUNDS
#include <cstdio>
#define PDC_ASSERT(cond)                                                                                               \
  do {                                                                                                                 \
    if (!(cond)) {                                                                                                     \
      printf("PDC_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x,     \
             (int)threadIdx.x);                                                                                        \
      __trap();                                                                                                        \
    }                                                                                                                  \
  } while (0)
#else
#define PDC_ASSERT(cond) ((void)0)
#endif

namespace pdc {

constexpr int MAX_Z        = 384;
constexpr int BG1_EDGES_N  = 316;
constexpr int BG2_EDGES_N  = 197;
constexpr int MAX_EDGES    = 316;
constexpr int MAX_ROWS     = 46;
constexpr int MAX_DEG      = 19;
constexpr int LLR_MAX      = 120;
constexpr int LLR_INF      = 127;
constexpr int CLAMP_IN     = 64; // ldpc_decoder_impl.h:193-195
constexpr int XPOW_ENTRIES = 272;

// Base-graph description in constant memory, the standard's sparse form (see bg_tables.inc).
struct BgTables {
  uint8_t  row[2][MAX_EDGES];
  uint8_t  col[2][MAX_EDGES];
  uint16_t v[2][8][MAX_EDGES];
  uint16_t row_start[2][MAX_ROWS + 2];
  uint16_t row_pstart[2][MAX_ROWS + 2]; // row starts when every row is padded to an even number of edges
  uint8_t  row_free[2][MAX_ROWS + 2];   // 1: the row shares no variable node with the rows since the last barrier
  uint8_t  set_index[MAX_Z + 1]; // 0xff = not a lifting size
  // x^(32 k) mod P for the three CRC polynomials (index PDC_CRC16-1 .. PDC_CRC24B-1).
  uint32_t xpow32[3][XPOW_ENTRIES];
};

__host__ __device__ inline uint32_t crc_poly(int kind)
{
  return (kind == PDC_CRC16) ? 0x11021u : (kind == PDC_CRC24A) ? 0x1864CFBu : 0x1800063u;
}
__host__ __device__ inline int crc_order(int kind)
{
  return (kind == PDC_CRC16) ? 16 : 24;
}

// All six NR generator polynomials (crc_calculator.h:35-48) for the stand-alone CRC call; codeblocks and transport blocks
// only use the first three.
constexpr int CRC_KINDS = 6;
__host__ __device__ inline uint32_t crc_poly_any(int kind)
{
  return (kind == PDC_CRC24C) ? 0x1B2B117u : (kind == PDC_CRC11) ? 0xE21u : (kind == PDC_CRC6) ? 0x61u : crc_poly(kind);
}
__host__ __device__ inline int crc_order_any(int kind)
{
  return (kind == PDC_CRC24C) ? 24 : (kind == PDC_CRC11) ? 11 : (kind == PDC_CRC6) ? 6 : crc_order(kind);
}

// (W(x) * A(x)) mod P, W of degree < 32, A of degree < order.
__host__ __device__ inline uint32_t gf2_mulmod(uint32_t w, uint32_t a, uint32_t poly, int order)
{
  uint32_t top = 1u << order;
  uint32_t r   = 0;
#pragma unroll 4
  for (int i = 31; i >= 0; --i) {
    r <<= 1;
    if (r & top) {
      r ^= poly;
    }
    if ((w >> i) & 1u) {
      r ^= a;
    }
  }
  return r;
}

// Check-to-variable scaling: the only arithmetic difference between the reference variants.
__device__ __forceinline__ int scale_c2v(int x, int mode)
{
  if (mode == PDC_SCALE_X86) {
    return (x * 52428) >> 16;
  }
  if (mode == PDC_SCALE_GENERIC) {
    return (int)((float)x * 0.8f + 0.5f);
  }
  return (x * 204) >> 8;
}

// ---- scrambling-sequence helpers shared by the front end and the rate dematcher ---------------------------------------
// Sequences are stored one element per bit, element k of a codeword in bit (k & 31) of word (k >> 5).
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t seq_bit(const uint32_t* __restrict__ seq, uint32_t i)
{
  return (__ldg(seq + (i >> 5)) >> (i & 31u)) & 1u;
}
// Up to 32 elements starting at element i (element i + k in bit k).
__device__ __forceinline__ uint32_t seq_bits32(const uint32_t* __restrict__ seq, uint32_t i)
{
  const uint32_t* w = seq + (i >> 5);
  return __funnelshift_r(__ldg(w), __ldg(w + 1), i & 31u);
}
// Per byte: -v where the bit of `bits` is set (two's complement, -128 stays -128), v otherwise.
__device__ __forceinline__ uint32_t negate4(uint32_t v, uint32_t bits)
{
  const uint32_t one = (bits * 0x00204081u) & 0x01010101u; // bit k -> byte k
  const uint32_t m   = one * 0xffu;
  const uint32_t t   = v ^ m;
  return ((t & 0x7f7f7f7fu) + one) ^ (t & 0x80808080u);
}
#endif

constexpr uint32_t CB_NOT_SCRAMBLED = 0xffffffffu;

// Launch parameters shared by the kernels.
// Record of the last non-zero soft bit of a HARQ entry (BatchParams::harq_last). An entry is a PDC_MAX_CB_SOFT-long slot
// that codeblocks of different sizes use over time (the reference's buffer pool hands codeblock buffers out again without
// clearing them, rx_buffer_pool_impl.cpp:44): a record is exact for the first `extent` positions of the entry - the N of
// the codeblock whose rate dematcher wrote it - and says nothing about what older, longer codeblocks left behind them.
__host__ __device__ constexpr int32_t harq_last_pack(int last, int extent)
{
  return (int32_t)(((uint32_t)extent << 16) | (uint32_t)last);
}
static_assert(PDC_MAX_CB_SOFT < 32768, "extent and position share a non-negative 32-bit word");
// Position recorded in four slots written by one rate-dematcher launch, or -1 if any slot is unknown or the record
// describes fewer than n leading positions.
__device__ __forceinline__ int harq_last_known(const int4& slots, int n)
{
  const int mn = min(min(slots.x, slots.y), min(slots.z, slots.w));
  if (mn < 0 || (mn >> 16) < n) {
    return -1;
  }
  return max(max(slots.x, slots.y), max(slots.z, slots.w)) & 0xffff;
}

struct BatchParams {
  const pdc_cb_desc* cbs;
  uint32_t           n_cb;
  const int8_t*      llrs;
  int8_t*            harq;       // arena: entries x PDC_MAX_CB_SOFT
  uint32_t           harq_entries;
  pdc_cb_result*     results;
  uint8_t*           cb_bits;    // n_cb x PDC_MAX_CB_BYTES (device staging, always present)
  uint8_t*           harq_data;  // entries x PDC_MAX_CB_BYTES: decoded message bits kept with the HARQ entry
  int32_t*           harq_last;  // per entry, 4 slots (one per dematcher part; take the maximum), see harq_last_pack:
                                 // low half = 1 + index of the last non-zero soft bit, high half = how many leading
                                 // positions of the entry the record describes; -1 = unknown (scan the entry)
  int                scale_mode;
  int                simd_width;
  // Deferred descrambling (codewords without UCI whose UL-SCH soft bits were not materialised): per codeblock
  // {first word of the codeword's sequence, sequence element of the codeblock's first soft bit, offset of that soft bit
  // in `raw`, -}; x = CB_NOT_SCRAMBLED: read `llrs` as usual. nullptr: no codeblock is affected.
  const uint4*       cb_scr;
  const uint32_t*    seq;
  const int8_t*      raw;
};

struct TbParams {
  const pdc_tb_desc* tbs;
  uint32_t           n_tb;
  const pdc_cb_desc* cbs;
  const pdc_cb_result* cb_results;
  const uint8_t*     cb_bits;
  pdc_tb_result*     tb_results;
  uint8_t*           tb_bytes;
};

// Launch with programmatic stream serialization: the kernel may begin while the previous kernel of the stream drains
// (its launch latency and prologue overlap that tail) and must execute pdl_wait() before it touches anything the
// previous kernel wrote - or anything the previous kernel may still read, if it writes it.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = grid;
  cfg.blockDim           = block;
  cfg.dynamicSmemBytes   = smem;
  cfg.stream             = s;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs                                          = attr;
  cfg.numAttrs                                       = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

} // namespace pdc
