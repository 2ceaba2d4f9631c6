// Soft demapper: equalised symbols + post-equalisation noise variances -> int8 soft bits.
//
// Bit-exact target: demodulation_mapper_impl::demodulate_soft (lib/phy/upper/channel_modulation/
// demodulation_mapper_impl.cpp:78-106) as an x86 build of the reference computes it. That build picks its kernels at
// compile time: whole blocks of 16 / 8 / 16 / 4 symbols (QPSK / 16QAM / 64QAM / 256QAM) of every demodulate_soft CALL
// go through the AVX2 kernels (demodulation_mapper_qpsk.cpp:39-78, _qam16.cpp:41-116, _qam64.cpp:185-283,
// _qam256.cpp:228-273 with avx2_helpers.h:121-271), the remainder through a scalar loop whose arithmetic differs
// (division by the noise variance or the interval width instead of multiplication by a reciprocal, round-half-away
// instead of round-half-even, one |z|^2 near-zero test instead of one per component). Both are reproduced, selected by
// the position of the symbol inside its call, so the call boundaries (one call per OFDM symbol in
// pusch_demodulator_impl.cpp:231-247) are part of the input. GCC fuses "slope * value + intercept" (and 16QAM's
// "0.8 - gain * |x|" in the scalar loop) into one FMA; every other operation is a separate IEEE single-precision
// operation, written with the _rn intrinsics so that nvcc cannot contract anything else.
//
// HBM-bound: 12 bytes read and Qm bytes written per symbol. One CTA per tile of DEMOD_TILE symbols of one call; soft bits
// are staged in shared memory and leave in 16-byte words.
#pragma once

#include "pdc_device.cuh"

namespace pdc {

constexpr int      DEMOD_THREADS = 256;
constexpr int      DEMOD_TILE    = 1024;
constexpr uint32_t DEMOD_SCALAR_ONLY = 1u; // flags: the portable build of the reference (no SIMD blocks)

struct DemodCall {
  uint32_t sym_off; // first symbol of the call in the symbol / noise-variance arrays
  uint32_t n_sym;
  uint32_t llr_off; // first soft bit of the call in the output
  uint32_t mod;     // PDC_MOD_*
};

struct DemodTile {
  uint32_t call;
  uint32_t first; // first symbol of the tile inside its call
};

struct DemodArgs {
  const DemodCall* calls;
  const DemodTile* tiles;
  const float2*    symbols;
  const float*     noise_vars;
  int8_t*          llrs;
  uint32_t         flags;
};

namespace demod {

// 1 / sqrt(10), 1 / sqrt(42), 1 / sqrt(170) in single precision ("1.0F / std::sqrt(10.0F)", demodulation_mapper_qam16.cpp:39).
constexpr float U10  = 0.3162277638912201f;  // 0x3ea1e89b
constexpr float U42  = 0.15430335700511932f; // 0x3e1e01b3
constexpr float U170 = 0.0766965001821518f;  // 0x3d9d130e
constexpr float GAIN_QPSK = 2.8284270763397217f; // 2 * float(sqrt(2)), 0x403504f3

// Piecewise-linear LLR tables, 7 of them: 64QAM bits 01 / 23 / 45 (demodulation_mapper_qam64.cpp:43-80), 256QAM bits
// 01 / 23 / 45 / 67 (demodulation_mapper_qam256.cpp:43-165). slope = k * unit, intercept = num / den.
struct TableDef {
  int8_t n, width_units;
  int8_t slope_k[16];
  int8_t icpt_num[16];
};
__constant__ TableDef c_demod_tab[7] = {
    {8, 2, {16, 12, 8, 4, 4, 8, 12, 16}, {24, 12, 4, 0, 0, -4, -12, -24}},
    {8, 2, {8, 4, 4, 8, -8, -4, -4, -8}, {20, 8, 8, 12, 12, 8, 8, 20}},
    {4, 4, {4, -4, 4, -4}, {12, -4, -4, 12}},
    {16, 2, {32, 28, 24, 20, 16, 12, 8, 4, 4, 8, 12, 16, 20, 24, 28, 32},
     {112, 84, 60, 40, 24, 12, 4, 0, 0, -4, -12, -24, -40, -60, -84, -112}},
    {16, 2, {16, 12, 8, 4, 4, 8, 12, 16, -16, -12, -8, -4, -4, -8, -12, -16},
     {88, 60, 36, 16, 16, 28, 36, 40, 40, 36, 28, 16, 16, 36, 60, 88}},
    {16, 2, {8, 4, 4, 8, -8, -4, -4, -8, 8, 4, 4, 8, -8, -4, -4, -8},
     {52, 24, 24, 44, -20, -8, -8, -12, -12, -8, -8, -20, 44, 24, 24, 52}},
    {8, 4, {4, -4, 4, -4, 4, -4, 4, -4}, {28, -20, 12, -4, -4, 12, -20, 28}},
};

// The tables in shared memory: per table 16 slopes, 16 intercepts, then {width, 1 / width}.
struct SmemTables {
  float slope[7][16];
  float icpt[7][16];
  float width[7];
  float inv_width[7];
};

__device__ __forceinline__ void build_tables(SmemTables& t, int tid)
{
  if (tid < 7 * 16) {
    const int   g = tid >> 4, i = tid & 15;
    const float unit = (g < 3) ? U42 : U170, den = (g < 3) ? 21.0f : 85.0f;
    t.slope[g][i] = __fmul_rn((float)c_demod_tab[g].slope_k[i], unit);
    t.icpt[g][i]  = __fdiv_rn((float)c_demod_tab[g].icpt_num[i], den);
    if (i == 0) {
      const float w  = __fmul_rn((float)c_demod_tab[g].width_units, unit);
      t.width[g]     = w;
      t.inv_width[g] = __fdiv_rn(1.0f, w);
    }
  }
}

// Conversion of an integer-valued float the way cvttss2si / cvtps2dq do it: NaN and out-of-range give 0x80000000.
__device__ __forceinline__ uint32_t x86_f2i(float v)
{
  return (v >= -2147483648.0f && v < 2147483648.0f) ? (uint32_t)(int32_t)v : 0x80000000u;
}

// log_likelihood_ratio::quantize (lib/phy/upper/log_likelihood_ratio.cpp:89-98): the scalar loop's quantiser.
__device__ __forceinline__ uint32_t quantize_scalar(float value, float range_limit)
{
  float clipped = value;
  if (fabsf(value) > range_limit) {
    clipped = copysignf(range_limit, value);
  }
  float q = __fdiv_rn(clipped, range_limit);
  q       = __fmul_rn(q, 120.0f);
  return x86_f2i(roundf(q)) & 0xffu;
}

// mm256::quantize_ps (avx2_helpers.h:121-166), one element; scale = 120 / range_limit (6 or 5, exact).
__device__ __forceinline__ uint32_t quantize_simd(float value, float scale)
{
  float v = __fmul_rn(value, scale);
  if (v > 120.0f) {
    v = 120.0f;
  }
  if (v < -120.0f) {
    v = -120.0f;
  }
  v = rintf(v);
  const int32_t i = (int32_t)x86_f2i(v);
  return (i > 120 || i < -120) ? 0u : ((uint32_t)i & 0xffu);
}

__device__ __forceinline__ int interval_index(float q, int n)
{
  const int32_t idx = (int32_t)(x86_f2i(floorf(q)) + (uint32_t)(n >> 1));
  return min(max(idx, 0), n - 1);
}

// interval_function (demodulation_mapper_intervals.h:31-63).
__device__ __forceinline__ float interval_scalar(const SmemTables& t, int g, int n, float value, float rcp_noise)
{
  const int k = interval_index(__fdiv_rn(value, t.width[g]), n);
  return __fmul_rn(__fmaf_rn(t.slope[g][k], value, t.icpt[g][k]), rcp_noise);
}

// mm256::interval_function (avx2_helpers.h:234-254).
__device__ __forceinline__ float interval_simd(const SmemTables& t, int g, int n, float value, float rcp_noise)
{
  const int k = interval_index(__fmul_rn(value, t.inv_width[g]), n);
  float     l = __fmul_rn(__fmaf_rn(t.slope[g][k], value, t.icpt[g][k]), rcp_noise);
  if (fabsf(value) <= 1e-9f) {
    l = 0.0f;
  }
  return l;
}

// Soft bits of one symbol, byte k of the result = soft bit k. i_call = index of the symbol inside its call.
__device__ __forceinline__ uint64_t demod_symbol(const SmemTables& t, float re, float im, float nv, uint32_t mod,
                                                 bool simd, uint32_t i_call)
{
  uint64_t out = 0;
  if (mod <= PDC_MOD_BPSK) {
    // demod_BPSK_symbol (demodulation_mapper_impl.cpp:34-42); odd symbols of pi/2-BPSK are rotated by -90 degrees (:58-75).
    if (nv > 0.0f) {
      const bool  rot = (mod == PDC_MOD_PI_2_BPSK) && (i_call & 1u);
      const float sum = rot ? __fsub_rn(im, re) : __fadd_rn(re, im);
      out             = quantize_scalar(__fdiv_rn(__fmul_rn(sum, GAIN_QPSK), nv), 24.0f);
    }
    return out;
  }
  const float c[2] = {re, im};
  if (simd) {
    const float rcp = (nv > 0.0f) ? __fdiv_rn(1.0f, nv) : 0.0f; // safe_div (avx2_helpers.h:259-271)
    if (mod == PDC_MOD_QPSK) {
#pragma unroll
      for (int k = 0; k != 2; ++k) {
        out |= (uint64_t)quantize_simd(__fmul_rn(__fmul_rn(GAIN_QPSK, c[k]), rcp), 5.0f) << (8 * k);
      }
    } else if (mod == PDC_MOD_QAM16) {
      const float gain = __fmul_rn(4.0f, U10), thr = __fmul_rn(2.0f, U10);
#pragma unroll
      for (int k = 0; k != 2; ++k) {
        const float first  = __fmul_rn(gain, c[k]);
        const float second = __fsub_rn(__fmul_rn(2.0f, first), copysignf(0.8f, c[k]));
        float       l01    = (fabsf(c[k]) > thr) ? second : first;
        float       l23    = __fsub_rn(0.8f, fabsf(first));
        l01                = __fmul_rn(l01, rcp);
        l23                = __fmul_rn(l23, rcp);
        if (fabsf(c[k]) <= 1e-9f) {
          l01 = 0.0f;
          l23 = 0.0f;
        }
        out |= (uint64_t)quantize_simd(l01, 6.0f) << (8 * k);
        out |= (uint64_t)quantize_simd(l23, 6.0f) << (8 * (2 + k));
      }
    } else {
      const int g0 = (mod == PDC_MOD_QAM64) ? 0 : 3, ng = (int)mod >> 1;
      for (int g = 0; g != ng; ++g) {
        const int n = c_demod_tab[g0 + g].n;
#pragma unroll
        for (int k = 0; k != 2; ++k) {
          out |= (uint64_t)quantize_simd(interval_simd(t, g0 + g, n, c[k], rcp), 6.0f) << (8 * (2 * g + k));
        }
      }
    }
    return out;
  }
  // Scalar remainder loops.
  if (mod == PDC_MOD_QPSK) {
    // demod_QPSK_symbol (demodulation_mapper_qpsk.cpp:121-129).
    if (nv > 0.0f) {
#pragma unroll
      for (int k = 0; k != 2; ++k) {
        out |= (uint64_t)quantize_scalar(__fdiv_rn(__fmul_rn(GAIN_QPSK, c[k]), nv), 24.0f) << (8 * k);
      }
    }
    return out;
  }
  // is_near_zero(cf_t) (include/srsran/support/math_utils.h:91-94): |z|^2 as fma(re, re, im * im).
  if (1e-9f > __fmaf_rn(re, re, __fmul_rn(im, im))) {
    return 0;
  }
  if (mod == PDC_MOD_QAM16) {
    // demod_16QAM_symbol_01 / _23 (demodulation_mapper_qam16.cpp:192-222).
    if (nv > 0.0f) {
      const float gain = __fmul_rn(4.0f, U10), thr = __fmul_rn(2.0f, U10);
#pragma unroll
      for (int k = 0; k != 2; ++k) {
        float l01 = __fmul_rn(gain, c[k]);
        if (fabsf(c[k]) > thr) {
          l01 = __fsub_rn(__fmul_rn(2.0f, l01), copysignf(0.8f, c[k]));
        }
        const float l23 = __fmaf_rn(-gain, fabsf(c[k]), 0.8f);
        out |= (uint64_t)quantize_scalar(__fdiv_rn(l01, nv), 20.0f) << (8 * k);
        out |= (uint64_t)quantize_scalar(__fdiv_rn(l23, nv), 20.0f) << (8 * (2 + k));
      }
    }
    return out;
  }
  const float rcp = (nv > 0.0f) ? __fdiv_rn(1.0f, nv) : 0.0f;
  const int   g0 = (mod == PDC_MOD_QAM64) ? 0 : 3, ng = (int)mod >> 1;
  for (int g = 0; g != ng; ++g) {
    const int n = c_demod_tab[g0 + g].n;
#pragma unroll
    for (int k = 0; k != 2; ++k) {
      out |= (uint64_t)quantize_scalar(interval_scalar(t, g0 + g, n, c[k], rcp), 20.0f) << (8 * (2 * g + k));
    }
  }
  return out;
}

} // namespace demod

__global__ void __launch_bounds__(DEMOD_THREADS) demod_kernel(DemodArgs a)
{
  __shared__ demod::SmemTables tab;
  __shared__ __align__(16) unsigned char stage[DEMOD_THREADS * 8];
  const int tid = threadIdx.x;
  demod::build_tables(tab, tid);
  const DemodTile tile = a.tiles[blockIdx.x];
  const DemodCall call = a.calls[tile.call];
  const uint32_t  qm   = (call.mod == PDC_MOD_PI_2_BPSK) ? 1u : call.mod;
  const uint32_t  block =
      (call.mod == PDC_MOD_QPSK || call.mod == PDC_MOD_QAM64) ? 16u : (call.mod == PDC_MOD_QAM16) ? 8u : 4u;
  const uint32_t n_simd =
      ((a.flags & DEMOD_SCALAR_ONLY) || call.mod <= PDC_MOD_BPSK) ? 0u : (call.n_sym / block) * block;
  const uint32_t end = min(call.n_sym, tile.first + (uint32_t)DEMOD_TILE);
  __syncthreads();
  for (uint32_t base = tile.first; base < end; base += DEMOD_THREADS) {
    const uint32_t i = base + tid;
    if (i < end) {
      const float2   z  = __ldg(a.symbols + call.sym_off + i);
      const float    nv = __ldg(a.noise_vars + call.sym_off + i);
      const uint64_t v  = demod::demod_symbol(tab, z.x, z.y, nv, call.mod, i < n_simd, i);
      unsigned char* p  = stage + tid * qm;
      if (qm == 8) {
        *reinterpret_cast<uint2*>(p) = make_uint2((uint32_t)v, (uint32_t)(v >> 32));
      } else if (qm == 6) {
        reinterpret_cast<uint16_t*>(p)[0] = (uint16_t)v;
        reinterpret_cast<uint16_t*>(p)[1] = (uint16_t)(v >> 16);
        reinterpret_cast<uint16_t*>(p)[2] = (uint16_t)(v >> 32);
      } else if (qm == 4) {
        *reinterpret_cast<uint32_t*>(p) = (uint32_t)v;
      } else if (qm == 2) {
        *reinterpret_cast<uint16_t*>(p) = (uint16_t)v;
      } else {
        *p = (unsigned char)v;
      }
    }
    __syncthreads();
    // Staged soft bits of this step leave in 16-byte words where the destination allows it.
    const uint32_t nbytes = min((uint32_t)DEMOD_THREADS, end - base) * qm;
    int8_t*        dst    = a.llrs + (size_t)call.llr_off + (size_t)base * qm;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
      const uint32_t n16 = nbytes >> 4;
      for (uint32_t w = tid; w < n16; w += DEMOD_THREADS) {
        reinterpret_cast<uint4*>(dst)[w] = reinterpret_cast<const uint4*>(stage)[w];
      }
      for (uint32_t b = (n16 << 4) + tid; b < nbytes; b += DEMOD_THREADS) {
        dst[b] = (int8_t)stage[b];
      }
    } else {
      for (uint32_t b = tid; b < nbytes; b += DEMOD_THREADS) {
        dst[b] = (int8_t)stage[b];
      }
    }
    __syncthreads();
  }
}

} // namespace pdc
