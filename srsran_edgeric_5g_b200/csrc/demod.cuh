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
// 12 bytes read and Qm bytes written per symbol. One CTA per tile of DEMOD_TILE symbols of one call, four symbols per
// thread (all loads of a tile first; persistent CTAs striding over smaller tiles, with or without a one-tile look-ahead, measured slower: 26-28 us against 19.5 us per 16-cell slot); soft bits are staged in shared memory and leave in 16-byte words. The call table travels
// in the kernel parameters, the piecewise-linear tables are computed once per context. At ~115 instructions per 256QAM
// symbol the kernel sits between the issue and the HBM rooflines.
#pragma once

#include "pdc_device.cuh"

namespace pdc {

constexpr int      DEMOD_THREADS = 256;
constexpr int      DEMOD_TILE    = 1024; // symbols per tile (one tile per CTA), four per thread
constexpr int      DEMOD_PER_THREAD = DEMOD_TILE / DEMOD_THREADS;
constexpr int      DEMOD_MAX_CALLS = 224; // calls per launch: the call table travels in the kernel parameters
constexpr uint32_t DEMOD_SCALAR_ONLY = 1u; // flags: the portable build of the reference (no SIMD blocks)

struct DemodCall {
  uint32_t sym_off; // first symbol of the call in the symbol / noise-variance arrays
  uint32_t n_sym;
  uint32_t llr_off; // first soft bit of the call in the output
  uint32_t mod;     // PDC_MOD_*
};

// Kernel parameters: the calls of this launch and the first tile (CTA) of each; no table in device memory, hence no
// upload in front of the kernel.
struct DemodArgs {
  DemodCall     calls[DEMOD_MAX_CALLS];
  uint32_t      tile_start[DEMOD_MAX_CALLS + 1];
  const float*  tables; // SmemTables image computed once per context (demod_tables_kernel)
  uint32_t      n_calls;
  uint32_t      flags;
  const float2* symbols;
  const float*  noise_vars;
  int8_t*       llrs;
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

// The tables in shared memory: per table 16 {slope, intercept} pairs, then {width, 1 / width}.
struct SmemTables {
  float2 si[7][16];
  float  width[7];
  float  inv_width[7];
};

constexpr int TABLE_FLOATS = sizeof(SmemTables) / sizeof(float);

// The tables, computed on the device with the reference's single-precision operations; run once per context.
__global__ void demod_tables_kernel(float* image)
{
  SmemTables& t   = *reinterpret_cast<SmemTables*>(image);
  const int   tid = threadIdx.x;
  if (tid < 7 * 16) {
    const int   g = tid >> 4, i = tid & 15;
    const float unit = (g < 3) ? U42 : U170, den = (g < 3) ? 21.0f : 85.0f;
    t.si[g][i] = make_float2(__fmul_rn((float)c_demod_tab[g].slope_k[i], unit),
                             __fdiv_rn((float)c_demod_tab[g].icpt_num[i], den));
    if (i == 0) {
      const float w  = __fmul_rn((float)c_demod_tab[g].width_units, unit);
      t.width[g]     = w;
      t.inv_width[g] = __fdiv_rn(1.0f, w);
    }
  }
}

__device__ __forceinline__ void load_tables(SmemTables& t, const float* image, int tid)
{
  if (tid < TABLE_FLOATS) {
    reinterpret_cast<float*>(&t)[tid] = __ldg(image + tid);
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

__device__ __forceinline__ int interval_index(float q, int n)
{
  const int32_t idx = (int32_t)(x86_f2i(floorf(q)) + (uint32_t)(n >> 1));
  return min(max(idx, 0), n - 1);
}

// interval_function (demodulation_mapper_intervals.h:31-63).
__device__ __forceinline__ float interval_scalar(const SmemTables& t, int g, int n, float value, float rcp_noise)
{
  const int    k  = interval_index(__fdiv_rn(value, t.width[g]), n);
  const float2 si = t.si[g][k];
  return __fmul_rn(__fmaf_rn(si.x, value, si.y), rcp_noise);
}

// ---- SIMD-block semantics, written for instruction count (this path takes all but the last few symbols of a call) ----

__device__ __forceinline__ float max_nan(float a, float b)
{
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float min_nan(float a, float b)
{
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// mm256::quantize_ps, one element: scale, clip to +-120 with a NaN passing through (the NaN-propagating min / max),
// round to nearest even and convert - cvt.rni turns the NaN into the 0 that check_bounds_epi32 makes of it.
__device__ __forceinline__ uint32_t quantize_fast(float value, float scale)
{
  return (uint32_t)__float2int_rn(min_nan(max_nan(__fmul_rn(value, scale), -120.0f), 120.0f));
}

// compute_interval_idx (avx2_helpers.h:174-193): clamp(int(floor(value / width)) + n / 2, 0, n - 1), where the x86
// conversion of a floor at or beyond +-2^31 gives 0x80000000 and so lands in interval 0. cvt.rmi saturates instead
// (+2^31 -> INT_MAX), but INT_MAX + n / 2 wraps negative as well and is clamped to the same interval 0; a NaN converts
// to 0 here and to 0x80000000 there, which is immaterial because slope * NaN + intercept is NaN for every interval.
__device__ __forceinline__ int interval_index_fast(float q, int half_n, int n_minus_1)
{
  return min(__viaddmax_s32(__float2int_rd(q), half_n, 0), n_minus_1);
}

// Four soft bits {a, b, c, d} -> one little-endian word.
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// One symbol through the AVX2-kernel arithmetic (demodulation_mapper_qpsk.cpp:39-78, _qam16.cpp:41-116,
// _qam64.cpp:185-283, _qam256.cpp:228-273). The "force zero where |value| <= 1e-9" blend of interval_function is applied
// as a zero reciprocal of that component (a finite value times zero quantises to the same 0).
template <uint32_t MOD>
__device__ __forceinline__ uint64_t demod_symbol_simd(const SmemTables& t, float re, float im, float nv)
{
  const float rcp  = (nv > 0.0f) ? __frcp_rn(nv) : 0.0f; // safe_div(1, noise) (avx2_helpers.h:259-271); IEEE 1 / x
  const float c[2] = {re, im};
  if (MOD == PDC_MOD_QPSK) {
    const uint32_t a = quantize_fast(__fmul_rn(__fmul_rn(GAIN_QPSK, re), rcp), 5.0f);
    const uint32_t b = quantize_fast(__fmul_rn(__fmul_rn(GAIN_QPSK, im), rcp), 5.0f);
    return __byte_perm(a, b, 0x0040) & 0xffffu;
  }
  if (MOD == PDC_MOD_QAM16) {
    const float gain = __fmul_rn(4.0f, U10), thr = __fmul_rn(2.0f, U10);
    uint32_t    q[4];
#pragma unroll
    for (int k = 0; k != 2; ++k) {
      const float first  = __fmul_rn(gain, c[k]);
      const float second = __fsub_rn(__fmul_rn(2.0f, first), copysignf(0.8f, c[k]));
      const float rc     = (fabsf(c[k]) <= 1e-9f) ? 0.0f : rcp;
      const float l01    = (fabsf(c[k]) > thr) ? second : first;
      const float l23    = __fsub_rn(0.8f, fabsf(first));
      // |value| <= 1e-9: both values are finite, so the zero reciprocal gives the blended zero
      q[k]     = quantize_fast(__fmul_rn(l01, rc), 6.0f);
      q[2 + k] = quantize_fast(__fmul_rn(l23, rc), 6.0f);
    }
    return pack4(q[0], q[1], q[2], q[3]);
  }
  // 64QAM / 256QAM: the groups of width 2 units share their interval index, the last group has width 4 units.
  constexpr bool Q256 = (MOD == PDC_MOD_QAM256);
  constexpr int  G0 = Q256 ? 3 : 0, NG = Q256 ? 4 : 3, HALF_A = Q256 ? 8 : 4, HALF_B = Q256 ? 4 : 2;
  const float    inv_a = t.inv_width[G0], inv_b = t.inv_width[G0 + NG - 1];
  uint32_t       q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k != 2; ++k) {
    const float v  = c[k];
    const float rc = (fabsf(v) <= 1e-9f) ? 0.0f : rcp;
    const int   ka = interval_index_fast(__fmul_rn(v, inv_a), HALF_A, 2 * HALF_A - 1);
    const int   kb = interval_index_fast(__fmul_rn(v, inv_b), HALF_B, 2 * HALF_B - 1);
#pragma unroll
    for (int g = 0; g != NG; ++g) {
      const float2 si = t.si[G0 + g][(g == NG - 1) ? kb : ka];
      q[2 * g + k]    = quantize_fast(__fmul_rn(__fmaf_rn(si.x, v, si.y), rc), 6.0f);
    }
  }
  return (uint64_t)pack4(q[0], q[1], q[2], q[3]) | ((uint64_t)pack4(q[4], q[5], q[6], q[7]) << 32);
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
  if (simd) {
    return (mod == PDC_MOD_QAM256)  ? demod_symbol_simd<PDC_MOD_QAM256>(t, re, im, nv)
           : (mod == PDC_MOD_QAM64) ? demod_symbol_simd<PDC_MOD_QAM64>(t, re, im, nv)
           : (mod == PDC_MOD_QAM16) ? demod_symbol_simd<PDC_MOD_QAM16>(t, re, im, nv)
                                    : demod_symbol_simd<PDC_MOD_QPSK>(t, re, im, nv);
  }
  const float c[2] = {re, im};
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

// Stores the qm soft bits of one symbol (byte k of v = soft bit k) into the staging area.
__device__ __forceinline__ void stage_symbol(unsigned char* p, uint64_t v, uint32_t qm)
{
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

// A tile whose symbols all take the SIMD-block arithmetic: all loads before any arithmetic.
template <uint32_t MOD>
__device__ __forceinline__ void tile_simd(const SmemTables& tab, const DemodArgs& a, const DemodCall& call, uint32_t first,
                                          uint32_t end, unsigned char* stage, int tid)
{
  float2 z[DEMOD_PER_THREAD];
  float  nv[DEMOD_PER_THREAD];
#pragma unroll
  for (int r = 0; r != DEMOD_PER_THREAD; ++r) {
    const uint32_t i = first + r * DEMOD_THREADS + tid;
    if (i < end) {
      z[r]  = __ldg(a.symbols + call.sym_off + i);
      nv[r] = __ldg(a.noise_vars + call.sym_off + i);
    }
  }
#pragma unroll
  for (int r = 0; r != DEMOD_PER_THREAD; ++r) {
    const uint32_t i = first + r * DEMOD_THREADS + tid;
    if (i < end) {
      stage_symbol(stage + (size_t)(i - first) * MOD, demod_symbol_simd<MOD>(tab, z[r].x, z[r].y, nv[r]), MOD);
    }
  }
}

} // namespace demod

__global__ void __launch_bounds__(DEMOD_THREADS, 8) demod_kernel(const __grid_constant__ DemodArgs a)
{
  __shared__ demod::SmemTables tab;
  __shared__ __align__(16) unsigned char stage[DEMOD_TILE * 8];
  const int tid = threadIdx.x;
  demod::load_tables(tab, a.tables, tid);
  __syncthreads();
  {
    const uint32_t tile = blockIdx.x;
    // The call of this tile: last call whose first tile is not beyond it.
    uint32_t lo = 0, hi = a.n_calls;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (a.tile_start[mid] <= tile) {
        lo = mid;
      } else {
        hi = mid;
      }
    }
    const DemodCall call  = a.calls[lo];
    const uint32_t  first = (tile - a.tile_start[lo]) * DEMOD_TILE;
    const uint32_t  qm    = (call.mod == PDC_MOD_PI_2_BPSK) ? 1u : call.mod;
    const uint32_t  block =
        (call.mod == PDC_MOD_QPSK || call.mod == PDC_MOD_QAM64) ? 16u : (call.mod == PDC_MOD_QAM16) ? 8u : 4u;
    const uint32_t n_simd =
        ((a.flags & DEMOD_SCALAR_ONLY) || call.mod <= PDC_MOD_BPSK) ? 0u : (call.n_sym / block) * block;
    const uint32_t end = min(call.n_sym, first + (uint32_t)DEMOD_TILE);
    if (end <= n_simd) {
      // Every symbol of the tile lies in the SIMD blocks of its call (all tiles but the last of a call, and the last too
      // unless the call has a remainder): straight-line code for the modulation.
      if (call.mod == PDC_MOD_QAM256) {
        demod::tile_simd<PDC_MOD_QAM256>(tab, a, call, first, end, stage, tid);
      } else if (call.mod == PDC_MOD_QAM64) {
        demod::tile_simd<PDC_MOD_QAM64>(tab, a, call, first, end, stage, tid);
      } else if (call.mod == PDC_MOD_QAM16) {
        demod::tile_simd<PDC_MOD_QAM16>(tab, a, call, first, end, stage, tid);
      } else {
        demod::tile_simd<PDC_MOD_QPSK>(tab, a, call, first, end, stage, tid);
      }
    } else {
      // Tiles with symbols of the scalar remainder (and BPSK, which has no SIMD kernel): symbol by symbol.
#pragma unroll 1
      for (uint32_t i = first + tid; i < end; i += DEMOD_THREADS) {
        const float2 z  = __ldg(a.symbols + call.sym_off + i);
        const float  nv = __ldg(a.noise_vars + call.sym_off + i);
        demod::stage_symbol(stage + (size_t)(i - first) * qm,
                            demod::demod_symbol(tab, z.x, z.y, nv, call.mod, i < n_simd, i), qm);
      }
    }
    __syncthreads();
    // The soft bits of the tile leave in 16-byte words where the destination allows it.
    const uint32_t nbytes = (end - first) * qm;
    int8_t*        dst    = a.llrs + (size_t)call.llr_off + (size_t)first * qm;
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
  }
}

} // namespace pdc
