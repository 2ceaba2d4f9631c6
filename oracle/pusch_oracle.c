/* TEST INFRASTRUCTURE - NOT PRODUCT CODE. See pusch_oracle.h for scope, parity status and usage rules.
 *
 * Plain-C restatement of the reference algorithm, written from the behaviour of the files cited at each function
 * (paths relative to /root/reference/srsRAN-5G-ER/). Scalar, single-threaded, no SIMD: the point is clarity.
 */
#define _POSIX_C_SOURCE 199309L
#include "pusch_oracle.h"
#include "bg_tables.inc"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_Z 384
#define MAX_NFULL 68
#define MAX_M 46
#define MAX_DEG 19
#define LLR_MAX 120
#define LLR_INF 127
#define MAX_CB_SIZE (66 * MAX_Z)
#define MAX_CB_BYTES (22 * MAX_Z / 8)

/* ------------------------------------------------------------------------------------------------------------------ */
/* LLR algebra - include/srsran/phy/upper/log_likelihood_ratio.h:46-244, lib/phy/upper/log_likelihood_ratio.cpp:40-87. */
/* ------------------------------------------------------------------------------------------------------------------ */

/* isinf: anything outside [-120, 120] (log_likelihood_ratio.h:159-165). */
static int is_inf(int v)
{
  return (v > LLR_MAX) || (v < -LLR_MAX);
}

/* log_likelihood_ratio::operator+ : opposite values cancel, infinities are sticky, otherwise saturate at +-120.
 * "a + b" is evaluated by the reference as "b += a" (log_likelihood_ratio.h:103-107), so an infinite b wins over an
 * infinite a; the negation wraps in int8 like the reference's unary minus. */
static int8_t llr_add(int a, int b)
{
  if (b == (int8_t)(-a)) {
    return 0;
  }
  if (is_inf(b)) {
    return (int8_t)b;
  }
  if (is_inf(a)) {
    return (int8_t)a;
  }
  int s = a + b;
  if (s > LLR_MAX) {
    s = LLR_MAX;
  }
  if (s < -LLR_MAX) {
    s = -LLR_MAX;
  }
  return (int8_t)s;
}

/* What the SIMD dematchers do on whole blocks: saturating int8 add, then clamp to +-120, no infinity rules
 * (ldpc/ldpc_rate_dematcher_avx512_impl.cpp:45-59, ldpc/ldpc_rate_dematcher_avx2_impl.cpp:45-59). */
static int8_t simd_add(int a, int b)
{
  int s = a + b;
  if (s > 127) {
    s = 127;
  }
  if (s < -128) {
    s = -128;
  }
  if (s > LLR_MAX) {
    s = LLR_MAX;
  }
  if (s < -LLR_MAX) {
    s = -LLR_MAX;
  }
  return (int8_t)s;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* CRC - crc_calculator_generic_impl.cpp:29-57 (polynomials), :111-133 (bitwise remainder).                            */
/* ------------------------------------------------------------------------------------------------------------------ */

static void crc_params(int kind, uint32_t* poly, int* order)
{
  switch (kind) {
    case ORC_CRC16:
      *poly  = 0x11021;
      *order = 16;
      break;
    case ORC_CRC24A:
      *poly  = 0x1864CFB;
      *order = 24;
      break;
    case ORC_CRC24C:
      *poly  = 0x1B2B117;
      *order = 24;
      break;
    case ORC_CRC11:
      *poly  = 0xE21;
      *order = 11;
      break;
    case ORC_CRC6:
      *poly  = 0x61;
      *order = 6;
      break;
    default:
      *poly  = 0x1800063;
      *order = 24;
      break;
  }
}

static int get_bit(const uint8_t* packed, int i)
{
  return (packed[i >> 3] >> (7 - (i & 7))) & 1;
}

static void put_bit(uint8_t* packed, int i, int b)
{
  uint8_t m = (uint8_t)(0x80 >> (i & 7));
  if (b) {
    packed[i >> 3] |= m;
  } else {
    packed[i >> 3] &= (uint8_t)~m;
  }
}

uint32_t orc_crc(int crc_kind, const uint8_t* packed, int nbits)
{
  uint32_t poly;
  int      order;
  crc_params(crc_kind, &poly, &order);
  uint32_t top = 1u << order;
  uint32_t reg = 0;
  for (int i = 0; i != nbits + order; ++i) {
    int b = (i < nbits) ? get_bit(packed, i) : 0;
    reg   = (reg << 1) | (uint32_t)b;
    if (reg & top) {
      reg ^= poly;
    }
  }
  return reg & (top - 1);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Rate dematcher.                                                                                                     */
/* ------------------------------------------------------------------------------------------------------------------ */

/* combine_softbits over one chunk: generic ldpc_rate_dematcher_impl.cpp:116-126; SIMD variants process
 * floor(len/width) blocks with simd_add and the remaining tail with the generic rule. */
static void combine_chunk(int8_t* out, const int8_t* in, int len, int simd_width)
{
  int n_simd = (simd_width > 0) ? (len / simd_width) * simd_width : 0;
  for (int i = 0; i != n_simd; ++i) {
    out[i] = simd_add(out[i], in[i]);
  }
  for (int i = n_simd; i != len; ++i) {
    out[i] = llr_add(out[i], in[i]);
  }
}

void orc_rate_dematch(int8_t*       out,
                      int           N,
                      const int8_t* in,
                      int           E,
                      int           new_data,
                      int           rv,
                      int           qm,
                      int           nref,
                      int           nof_filler,
                      int           simd_width)
{
  static const double sf_bg1[4] = {0, 17, 33, 56};
  static const double sf_bg2[4] = {0, 13, 25, 43};

  /* ldpc_rate_dematcher_impl.cpp:61-105: circular buffer length, base graph from N, k0. */
  int           Ncb = (nref > 0) ? ((nref < N) ? nref : N) : N;
  const double* sf;
  int           n_short, k_b;
  if (N % 66 == 0) {
    sf      = sf_bg1;
    n_short = 66;
    k_b     = 22;
  } else {
    sf      = sf_bg2;
    n_short = 50;
    k_b     = 10;
  }
  int Z     = N / n_short;
  int K_sys = (k_b - 2) * Z;
  int info  = K_sys - nof_filler;
  int k0    = (int)((uint16_t)floor((sf[rv] * Ncb) / N)) * Z;

  /* Deinterleaver :203-257: out[(E/Qm) j + i] = in[i Qm + j]. */
  int8_t* deint = NULL;
  if (qm > 1) {
    deint = (int8_t*)malloc((size_t)E);
    int K = E / qm;
    for (int i = 0; i != K; ++i) {
      for (int j = 0; j != qm; ++j) {
        deint[K * j + i] = in[i * qm + j];
      }
    }
    in = deint;
  }

  /* allot_llrs :128-201. The walk below keeps the reference's order of side effects, including the ones that look
   * accidental: in copy mode out[0, idx) is zeroed every time an information chunk is written, out[0, info) is zeroed
   * when the walk starts beyond the information bits, the final zeroing addresses the LAST (Ncb - idx) entries of the
   * N-long buffer, and positions in [Ncb, N) or skipped by k0 are otherwise left as they were. */
  int copy_mode = new_data;
  int idx       = k0;
  int left      = E;
  while (left > 0) {
    if (idx < info) {
      int n = info - idx;
      if (n > left) {
        n = left;
      }
      if (copy_mode) {
        memset(out, 0, (size_t)idx);
        memcpy(out + idx, in, (size_t)n);
      } else {
        combine_chunk(out + idx, in, n, simd_width);
      }
      idx += n;
      in += n;
      left -= n;
    } else if (copy_mode) {
      memset(out, 0, (size_t)info);
    }
    if (copy_mode) {
      memset(out + info, LLR_INF, (size_t)nof_filler);
    }
    if (idx < K_sys) {
      idx = K_sys;
    }
    int n = Ncb - idx;
    if (n > left) {
      n = left;
    }
    if (copy_mode) {
      memcpy(out + idx, in, (size_t)n);
    } else {
      combine_chunk(out + idx, in, n, simd_width);
    }
    idx = (idx + n) % Ncb;
    in += n;
    left -= n;
    if (left > 0) {
      copy_mode = 0;
    }
  }
  if (copy_mode && (idx != 0)) {
    memset(out + N - (Ncb - idx), 0, (size_t)(Ncb - idx));
  }
  free(deint);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Base graph access.                                                                                                  */
/* ------------------------------------------------------------------------------------------------------------------ */

typedef struct {
  int n_full, n_short, k_b, m;
  int deg[MAX_M];
  int col[MAX_M][MAX_DEG];
  int shift[MAX_M][MAX_DEG];
} graph_t;

static int set_index_of(int Z)
{
  for (int i = 0; i != NR_LDPC_NOF_LIFTING_SIZES; ++i) {
    if (NR_LDPC_LIFTING_SIZES[i] == Z) {
      return NR_LDPC_SET_INDEX[i];
    }
  }
  return -1;
}

/* ldpc_graph_impl.h:64-78 + ldpc_luts_impl.cpp:4521-4544: shift = V mod Z, edges of a row in ascending column order. */
static void build_graph(graph_t* g, int bg, int Z)
{
  int ls               = set_index_of(Z);
  int n_edges          = (bg == 1) ? BG1_NOF_EDGES : BG2_NOF_EDGES;
  const unsigned short(*e)[10] = (bg == 1) ? BG1_EDGES : BG2_EDGES;
  g->n_full            = (bg == 1) ? 68 : 52;
  g->n_short           = (bg == 1) ? 66 : 50;
  g->k_b               = (bg == 1) ? 22 : 10;
  g->m                 = (bg == 1) ? 46 : 42;
  memset(g->deg, 0, sizeof(g->deg));
  for (int i = 0; i != n_edges; ++i) {
    int m              = e[i][0];
    int d              = g->deg[m]++;
    g->col[m][d]       = e[i][1];
    g->shift[m][d]     = e[i][2 + ls] % Z;
  }
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* LDPC decoder.                                                                                                       */
/* ------------------------------------------------------------------------------------------------------------------ */

static int scale_c2v(int x, int mode)
{
  switch (mode) {
    case ORC_SCALE_GENERIC:
      return (int)roundf((float)x * 0.8f);
    case ORC_SCALE_NEON:
      return (x * 204) >> 8;
    default:
      return (x * 52428) >> 16;
  }
}

int orc_ldpc_decode(int           bg,
                    int           Z,
                    const int8_t* in,
                    int           n_in,
                    int           nof_filler,
                    int           crc_kind,
                    int           max_iter,
                    int           scale_mode,
                    uint8_t*      out,
                    int8_t*       soft_out)
{
  graph_t* g = (graph_t*)malloc(sizeof(graph_t));
  build_graph(g, bg, Z);
  int K = g->k_b * Z;

  /* ldpc_decoder_impl.cpp:85-94: trim trailing zeros; an all-zero input cannot be decoded. */
  int trimmed = n_in;
  while ((trimmed > 0) && (in[trimmed - 1] == 0)) {
    --trimmed;
  }
  if (trimmed == 0) {
    if (crc_kind == ORC_CRC_NONE) {
      memset(out, 0xff, (size_t)((K + 7) / 8));
      if (K % 8) {
        out[K / 8] = (uint8_t)(0xff << (8 - K % 8));
      }
    }
    free(g);
    return 0;
  }

  /* load_soft_bits :149-184: two punctured nodes at zero, full nodes clamped to +-64, a partial last node copied
   * unclamped (the reference leaves the rest of that node as it was; a fresh decoder has zeros there). */
  int8_t* soft = (int8_t*)calloc((size_t)(MAX_NFULL * MAX_Z), 1);
  int     full = (n_in / Z) * Z;
  for (int i = 0; i != full; ++i) {
    int v = in[i];
    if (v > 64) {
      v = 64;
    }
    if (v < -64) {
      v = -64;
    }
    soft[2 * Z + i] = (int8_t)v;
  }
  for (int i = full; i != n_in; ++i) {
    soft[2 * Z + i] = in[i];
  }

  /* :103-114: number of layers actually processed. */
  int cb_len = trimmed + 2 * Z;
  if (cb_len < K + 4 * Z) {
    cb_len = K + 4 * Z;
  }
  if (cb_len % Z != 0) {
    cb_len = (cb_len / Z + 1) * Z;
  }
  int layers = cb_len / Z - g->k_b;

  /* Check-to-variable messages, all zero = "not initialised" (:206-210: the first visit copies the soft bits). */
  int8_t(*c2v)[MAX_DEG][MAX_Z] = calloc((size_t)MAX_M, sizeof(*c2v));
  int8_t(*v2c)[MAX_Z]          = calloc((size_t)MAX_DEG, sizeof(*v2c));

  int result = 0;
  for (int it = 0; it != max_iter; ++it) {
    for (int m = 0; m != layers; ++m) {
      int deg = g->deg[m];
      /* update_variable_to_check_messages :186-229 with compute_var_to_check_msgs (avx512 :81-121). */
      for (int e = 0; e != deg; ++e) {
        const int8_t* s = soft + g->col[m][e] * Z;
        for (int j = 0; j != Z; ++j) {
          int d = s[j] - c2v[m][e][j];
          if (d > LLR_MAX) {
            d = LLR_MAX;
          }
          if (d < -LLR_MAX) {
            d = -LLR_MAX;
          }
          if (s[j] >= LLR_INF) {
            d = LLR_INF;
          }
          if (s[j] <= -LLR_INF) {
            d = -LLR_INF;
          }
          v2c[e][j] = (int8_t)d;
        }
      }
      /* update_check_to_variable_messages :246-318: per lifted check j, scan the row in adjacency order. */
      for (int j = 0; j != Z; ++j) {
        int min1 = LLR_MAX, min2 = LLR_MAX, arg = 0, par = 0;
        for (int e = 0; e != deg; ++e) {
          int x = v2c[e][(j + g->shift[m][e]) % Z];
          int a = abs(x);
          if (a < min2) {
            min2 = (a < min1) ? min1 : a;
          }
          if (a < min1) {
            min1 = a;
            arg  = e;
          }
          par ^= (x < 0);
        }
        int s1 = scale_c2v(min1, scale_mode);
        int s2 = scale_c2v(min2, scale_mode);
        for (int e = 0; e != deg; ++e) {
          int pos = (j + g->shift[m][e]) % Z;
          int x   = v2c[e][pos];
          int mag = (arg == e) ? s2 : s1;
          c2v[m][e][pos] = (int8_t)((par ^ (x < 0)) ? -mag : mag);
        }
      }
      /* update_soft_bits :231-244 with compute_soft_bits (avx512 :218-259) = promotion_sum (LLR.cpp:74-87). */
      for (int e = 0; e != deg; ++e) {
        int8_t* s = soft + g->col[m][e] * Z;
        for (int j = 0; j != Z; ++j) {
          int v = v2c[e][j];
          int c = c2v[m][e][j];
          int r;
          if (is_inf(v)) {
            r = (v > 0) ? LLR_INF : -LLR_INF;
          } else {
            r = v + c;
            if (r > LLR_MAX) {
              r = LLR_INF;
            }
            if (r < -LLR_MAX) {
              r = -LLR_INF;
            }
          }
          s[j] = (int8_t)r;
        }
      }
    }
    /* :125-134: hard decision (bit = soft <= 0, LLR.cpp:313-339), gate on "no zero soft bit", CRC over K - F bits. */
    if (crc_kind != ORC_CRC_NONE) {
      int no_zero = 1;
      for (int i = 0; i != K; ++i) {
        put_bit(out, i, soft[i] <= 0);
        no_zero &= (soft[i] != 0);
      }
      if (no_zero && (orc_crc(crc_kind, out, K - nof_filler) == 0)) {
        result = it + 1;
        break;
      }
    }
  }
  if (crc_kind == ORC_CRC_NONE) {
    for (int i = 0; i != K; ++i) {
      put_bit(out, i, soft[i] <= 0);
    }
  }
  if (soft_out) {
    memcpy(soft_out, soft, (size_t)(g->n_full * Z));
  }
  free(v2c);
  free(c2v);
  free(soft);
  free(g);
  return result;
}

/* pusch_codeblock_decoder.cpp:35-71. */
int orc_cb_decode(int8_t*       rm_buffer,
                  int           N,
                  const int8_t* in,
                  int           E,
                  int           new_data,
                  int           rv,
                  int           qm,
                  int           nref,
                  int           nof_filler,
                  int           crc_kind,
                  int           use_early_stop,
                  int           max_iter,
                  int           scale_mode,
                  int           simd_width,
                  uint8_t*      out)
{
  int bg = (N % 66 == 0) ? 1 : 2;
  int Z  = N / ((bg == 1) ? 66 : 50);
  int K  = ((bg == 1) ? 22 : 10) * Z;
  orc_rate_dematch(rm_buffer, N, in, E, new_data, rv, qm, nref, nof_filler, simd_width);
  if (use_early_stop) {
    return orc_ldpc_decode(bg, Z, rm_buffer, N, nof_filler, crc_kind, max_iter, scale_mode, out, NULL);
  }
  orc_ldpc_decode(bg, Z, rm_buffer, N, nof_filler, ORC_CRC_NONE, max_iter, scale_mode, out, NULL);
  return (orc_crc(crc_kind, out, K - nof_filler) == 0) ? max_iter : 0;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Segmentation - ldpc.h:128-228, ldpc_segmenter_impl.cpp:58-68, :254-331.                                             */
/* ------------------------------------------------------------------------------------------------------------------ */

static int tb_crc_bits(int tbs)
{
  return (tbs <= 3824) ? 16 : 24;
}

static int nof_codeblocks(int tbs, int bg)
{
  int b       = tbs + tb_crc_bits(tbs);
  int max_seg = (bg == 1) ? 8448 : 3840;
  return (b <= max_seg) ? 1 : (b + (max_seg - 24) - 1) / (max_seg - 24);
}

static int lifting_size_for(int tbs, int bg, int C)
{
  int b   = tbs + tb_crc_bits(tbs);
  int ref = 22;
  if (bg == 2) {
    ref = (b > 640) ? 10 : (b > 560) ? 9 : (b > 192) ? 8 : 6;
  }
  int b_out = b + ((C > 1) ? 24 * C : 0);
  for (int i = 0; i != NR_LDPC_NOF_LIFTING_SIZES; ++i) {
    if (NR_LDPC_LIFTING_SIZES[i] * ref * C >= b_out) {
      return NR_LDPC_LIFTING_SIZES[i];
    }
  }
  return 0;
}

int orc_segment_rx(int tbs_bits, int bg, int qm, int nof_layers, int n_llr, orc_cb_meta* meta)
{
  int C        = nof_codeblocks(tbs_bits, bg);
  int b_in     = tbs_bits + tb_crc_bits(tbs_bits);
  int b_out    = b_in + ((C > 1) ? 24 * C : 0);
  int Z        = lifting_size_for(tbs_bits, bg, C);
  int K        = ((bg == 1) ? 22 : 10) * Z;
  int crc_bits = (C > 1) ? 24 : 0;
  int max_info = (b_out + C - 1) / C - crc_bits;
  int sym_pl   = (n_llr / qm) / nof_layers;
  int n_short  = C - (sym_pl % C);
  int offset   = 0;
  for (int i = 0; i != C; ++i) {
    int per_cb           = (i < n_short) ? (sym_pl / C) : ((sym_pl + C - 1) / C);
    meta[i].Z            = Z;
    meta[i].full_length  = ((bg == 1) ? 3 : 5) * K;
    meta[i].rm_length    = per_cb * nof_layers * qm;
    meta[i].nof_filler   = K - (max_info + crc_bits);
    meta[i].cw_offset    = offset;
    meta[i].nof_crc_bits = (C == 1) ? tb_crc_bits(tbs_bits) : 24;
    offset += meta[i].rm_length;
  }
  return C;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Transport-block decoder - pusch_decoder_impl.cpp.                                                                   */
/* ------------------------------------------------------------------------------------------------------------------ */

static void copy_bits(uint8_t* dst, int dst_off, const uint8_t* src, int src_off, int n)
{
  for (int i = 0; i != n; ++i) {
    put_bit(dst, dst_off + i, get_bit(src, src_off + i));
  }
}

void orc_pusch_decode(orc_harq*            harq,
                      const int8_t*        llrs,
                      int                  n_llr,
                      int                  tb_bytes,
                      const orc_pusch_cfg* cfg,
                      int                  scale_mode,
                      int                  simd_width,
                      uint8_t*             tb_out,
                      int*                 stats)
{
  orc_cb_meta meta[256];
  int         tbs = tb_bytes * 8;
  int         C   = orc_segment_rx(tbs, cfg->bg, cfg->qm, cfg->nof_layers, n_llr, meta);

  /* select_crc :35-46. */
  int crc_kind = (C > 1) ? ORC_CRC24B : ((tbs > 3824) ? ORC_CRC24A : ORC_CRC16);

  /* new_data :131-135. */
  if (cfg->new_data) {
    memset(harq->crc_ok, 0, (size_t)C);
  }

  int n_obs = 0, it_min = 0, it_max = 0, it_sum = 0;
  for (int cb = 0; cb != C; ++cb) {
    int8_t*  rm  = harq->soft + (size_t)cb * MAX_CB_SIZE;
    uint8_t* msg = harq->data + (size_t)cb * MAX_CB_BYTES;
    int      N   = meta[cb].full_length;
    /* fork_codeblock_task :335-345: codeblocks already decoded keep combining but are not decoded again. */
    if (harq->crc_ok[cb]) {
      orc_rate_dematch(rm, N, llrs + meta[cb].cw_offset, meta[cb].rm_length, cfg->new_data, cfg->rv, cfg->qm,
                       cfg->nref, meta[cb].nof_filler, simd_width);
      continue;
    }
    int it = orc_cb_decode(rm, N, llrs + meta[cb].cw_offset, meta[cb].rm_length, cfg->new_data, cfg->rv, cfg->qm,
                           cfg->nref, meta[cb].nof_filler, crc_kind, cfg->use_early_stop, cfg->max_iter, scale_mode,
                           simd_width, msg);
    /* :357-363. */
    int obs = it;
    if (it > 0) {
      harq->crc_ok[cb] = 1;
    } else {
      obs = cfg->max_iter;
    }
    it_min = (n_obs == 0 || obs < it_min) ? obs : it_min;
    it_max = (n_obs == 0 || obs > it_max) ? obs : it_max;
    it_sum += obs;
    ++n_obs;
  }

  /* join_and_notify :384-450. */
  int tb_ok = 0;
  if (C == 1) {
    tb_ok = harq->crc_ok[0];
    if (tb_ok) {
      memcpy(tb_out, harq->data, (size_t)tb_bytes);
    }
  } else {
    int all = 1;
    for (int cb = 0; cb != C; ++cb) {
      all &= harq->crc_ok[cb];
    }
    if (all) {
      /* concatenate_codeblocks :452-497. */
      int      tb_off   = 0;
      uint32_t checksum = 0;
      for (int cb = 0; cb != C; ++cb) {
        int            K        = meta[cb].full_length / ((cfg->bg == 1) ? 3 : 5);
        int            n_data   = K - meta[cb].nof_crc_bits - meta[cb].nof_filler;
        int            free_bits = tbs - tb_off;
        int            n_new    = (free_bits < n_data) ? free_bits : n_data;
        const uint8_t* msg      = harq->data + (size_t)cb * MAX_CB_BYTES;
        copy_bits(tb_out, tb_off, msg, 0, n_new);
        if (cb == C - 1) {
          for (int i = 0; i != 24; ++i) {
            checksum = (checksum << 1) | (uint32_t)get_bit(msg, n_new + i);
          }
        }
        tb_off += n_new;
      }
      if (orc_crc(ORC_CRC24A, tb_out, tbs) == checksum) {
        tb_ok = 1;
      } else {
        memset(harq->crc_ok, 0, (size_t)C);
      }
    }
  }
  stats[0] = tb_ok;
  stats[1] = C;
  stats[2] = n_obs;
  stats[3] = it_min;
  stats[4] = it_max;
  stats[5] = it_sum;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* TX side (input synthesis only): TS 38.212 5.2.2 segmentation, 5.3.2 encoding, 5.4.2 rate matching.                  */
/* ------------------------------------------------------------------------------------------------------------------ */

/* acc[j] ^= v[(j + s) mod Z] */
static void xor_rot(uint8_t* acc, const uint8_t* v, int s, int Z)
{
  for (int j = 0; j != Z; ++j) {
    acc[j] ^= v[(j + s) % Z];
  }
}

void orc_ldpc_encode(int bg, int Z, const uint8_t* msg, uint8_t* cw)
{
  graph_t* g = (graph_t*)malloc(sizeof(graph_t));
  build_graph(g, bg, Z);
  int      kb = g->k_b;
  uint8_t* c  = (uint8_t*)calloc((size_t)(g->n_full * Z), 1);
  memcpy(c, msg, (size_t)(kb * Z));

  /* Core parity: rows 0..3, columns kb..kb+3. lam[i] = sum over information columns of row i. */
  uint8_t lam[4][MAX_Z];
  memset(lam, 0, sizeof(lam));
  for (int i = 0; i != 4; ++i) {
    for (int e = 0; e != g->deg[i]; ++e) {
      if (g->col[i][e] < kb) {
        xor_rot(lam[i], c + g->col[i][e] * Z, g->shift[i][e], Z);
      }
    }
  }
  /* Column kb appears in three core rows, two of them with the same shift: summing the four rows cancels everything
   * but one rotated copy of p0. */
  int sh[4], n_sh = 0;
  for (int i = 0; i != 4; ++i) {
    for (int e = 0; e != g->deg[i]; ++e) {
      if (g->col[i][e] == kb) {
        sh[n_sh++] = g->shift[i][e];
      }
    }
  }
  int d = (sh[0] == sh[1]) ? sh[2] : ((sh[0] == sh[2]) ? sh[1] : sh[0]);
  uint8_t sum[MAX_Z];
  for (int j = 0; j != Z; ++j) {
    sum[j] = lam[0][j] ^ lam[1][j] ^ lam[2][j] ^ lam[3][j];
  }
  uint8_t* p0 = c + kb * Z;
  for (int j = 0; j != Z; ++j) {
    p0[(j + d) % Z] = sum[j];
  }
  /* Back-substitution: repeatedly take a core row with exactly one unknown core parity column. */
  int known[4] = {1, 0, 0, 0};
  for (int round = 0; round != 3; ++round) {
    for (int i = 0; i != 4; ++i) {
      int unknown = -1, n_unknown = 0, unknown_shift = 0;
      for (int e = 0; e != g->deg[i]; ++e) {
        int col = g->col[i][e];
        if ((col >= kb) && (col < kb + 4) && !known[col - kb]) {
          unknown       = col;
          unknown_shift = g->shift[i][e];
          ++n_unknown;
        }
      }
      if (n_unknown != 1) {
        continue;
      }
      uint8_t acc[MAX_Z];
      memcpy(acc, lam[i], (size_t)Z);
      for (int e = 0; e != g->deg[i]; ++e) {
        int col = g->col[i][e];
        if ((col >= kb) && (col < kb + 4) && known[col - kb]) {
          xor_rot(acc, c + col * Z, g->shift[i][e], Z);
        }
      }
      uint8_t* p = c + unknown * Z;
      for (int j = 0; j != Z; ++j) {
        p[(j + unknown_shift) % Z] = acc[j];
      }
      known[unknown - kb] = 1;
    }
  }
  /* Extension parity: row m >= 4 has a single identity column kb + m. */
  for (int m = 4; m != g->m; ++m) {
    uint8_t* p = c + (kb + m) * Z;
    for (int e = 0; e != g->deg[m]; ++e) {
      if (g->col[m][e] != kb + m) {
        xor_rot(p, c + g->col[m][e] * Z, g->shift[m][e], Z);
      }
    }
  }
  memcpy(cw, c + 2 * Z, (size_t)((g->n_full - 2) * Z));
  free(c);
  free(g);
}

void orc_rate_match(const uint8_t* cw, int N, int E, int rv, int qm, int nref, int nof_filler, uint8_t* out)
{
  static const double sf_bg1[4] = {0, 17, 33, 56};
  static const double sf_bg2[4] = {0, 13, 25, 43};
  int           Ncb  = (nref > 0) ? ((nref < N) ? nref : N) : N;
  int           bg1  = (N % 66 == 0);
  const double* sf   = bg1 ? sf_bg1 : sf_bg2;
  int           Z    = N / (bg1 ? 66 : 50);
  int           K_sys = ((bg1 ? 22 : 10) - 2) * Z;
  int           k0   = (int)((uint16_t)floor((sf[rv] * Ncb) / N)) * Z;
  uint8_t*      sel  = (uint8_t*)malloc((size_t)E);
  int           idx  = k0 % Ncb;
  for (int k = 0; k != E;) {
    if ((idx >= K_sys - nof_filler) && (idx < K_sys)) {
      idx = K_sys % Ncb;
      continue;
    }
    sel[k++] = cw[idx];
    idx      = (idx + 1) % Ncb;
  }
  int per = E / qm;
  for (int i = 0; i != per; ++i) {
    for (int j = 0; j != qm; ++j) {
      out[i * qm + j] = sel[j * per + i];
    }
  }
  free(sel);
}

int orc_tb_encode(const uint8_t* tb, int tb_bytes, int bg, int rv, int qm, int nref, int nof_layers, int n_llr,
                  uint8_t* cw_bits)
{
  orc_cb_meta meta[256];
  int         tbs      = tb_bytes * 8;
  int         C        = orc_segment_rx(tbs, bg, qm, nof_layers, n_llr, meta);
  int         tcrc     = tb_crc_bits(tbs);
  int         b_in     = tbs + tcrc;
  uint8_t*    tb_crc   = (uint8_t*)calloc((size_t)(b_in / 8 + 8), 1);
  memcpy(tb_crc, tb, (size_t)tb_bytes);
  uint32_t chk = orc_crc((tcrc == 16) ? ORC_CRC16 : ORC_CRC24A, tb, tbs);
  for (int i = 0; i != tcrc; ++i) {
    put_bit(tb_crc, tbs + i, (chk >> (tcrc - 1 - i)) & 1);
  }
  int      Z       = meta[0].Z;
  int      K       = ((bg == 1) ? 22 : 10) * Z;
  int      cb_crc  = (C > 1) ? 24 : 0;
  int      info    = K - meta[0].nof_filler - cb_crc; /* bits taken per CB, incl. TB CRC and zero padding in the last */
  uint8_t* seg     = (uint8_t*)malloc((size_t)(K / 8 + 8));
  uint8_t* msg     = (uint8_t*)malloc((size_t)K);
  uint8_t* cw      = (uint8_t*)malloc((size_t)(66 * MAX_Z));
  int      tb_off  = 0;
  for (int cb = 0; cb != C; ++cb) {
    memset(seg, 0, (size_t)(K / 8 + 8));
    int take = info;
    if (tb_off + take > b_in) {
      take = b_in - tb_off; /* the rest is zero padding */
    }
    copy_bits(seg, 0, tb_crc, tb_off, take);
    tb_off += take;
    if (cb_crc) {
      uint32_t c = orc_crc(ORC_CRC24B, seg, info);
      for (int i = 0; i != 24; ++i) {
        put_bit(seg, info + i, (c >> (23 - i)) & 1);
      }
    }
    for (int i = 0; i != K; ++i) {
      msg[i] = (uint8_t)get_bit(seg, i);
    }
    orc_ldpc_encode(bg, Z, msg, cw);
    orc_rate_match(cw, meta[cb].full_length, meta[cb].rm_length, rv, qm, nref, meta[cb].nof_filler,
                   cw_bits + meta[cb].cw_offset);
  }
  free(cw);
  free(msg);
  free(seg);
  free(tb_crc);
  return C;
}

double orc_bench_cb_batch(int           n_cb,
                          const int8_t* llrs,
                          int           E,
                          int           N,
                          int           rv,
                          int           qm,
                          int           nref,
                          int           nof_filler,
                          int           crc_kind,
                          int           use_early_stop,
                          int           max_iter,
                          int*          iters_out)
{
  int8_t*         rm = (int8_t*)malloc((size_t)N);
  uint8_t         bits[MAX_CB_BYTES + 8];
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int cb = 0; cb != n_cb; ++cb) {
    int it = orc_cb_decode(rm, N, llrs + (size_t)cb * E, E, 1, rv, qm, nref, nof_filler, crc_kind, use_early_stop,
                           max_iter, ORC_SCALE_X86, 64, bits);
    if (iters_out) {
      iters_out[cb] = it;
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(rm);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- codeword front end ------------------------------------------------------------------------------------------------ */

/* TS 38.211 5.2.1: x1(n+31) = x1(n+3) + x1(n), x2(n+31) = x2(n+3) + x2(n+2) + x2(n+1) + x2(n), x1(0) = 1,
 * x2 = c_init, c(n) = x1(n + 1600) + x2(n + 1600). Bit-serial; state bit j = x(n + j). */
void orc_prg_bits(uint32_t c_init, uint32_t offset, uint32_t n, uint8_t* bits)
{
  uint32_t x1 = 1u, x2 = c_init & 0x7fffffffu;
  uint32_t skip = 1600u + offset;
  for (uint32_t i = 0; i != skip + n; ++i) {
    if (i >= skip) {
      bits[i - skip] = (uint8_t)((x1 ^ x2) & 1u);
    }
    uint32_t f1 = ((x1 >> 3) ^ x1) & 1u;
    uint32_t f2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
    x1          = (x1 >> 1) | (f1 << 30);
    x2          = (x2 >> 1) | (f2 << 30);
  }
}

void orc_revert_scrambling(int8_t* out, const int8_t* in, const uint8_t* seq_bits, uint32_t n)
{
  for (uint32_t i = 0; i != n; ++i) {
    out[i] = seq_bits[i] ? (int8_t)(uint8_t)(0u - (uint8_t)in[i]) : in[i];
  }
}

#define ORC_MAX_RE (275 * 12)

typedef struct {
  uint8_t v[ORC_MAX_RE];
  int     size;
} re_set;

static void re_fill(re_set* s, int size, int value)
{
  s->size = size;
  memset(s->v, value ? 1 : 0, (size_t)size);
}
static int re_count(const re_set* s)
{
  int c = 0;
  for (int i = 0; i != s->size; ++i) {
    c += s->v[i];
  }
  return c;
}
/* First m_re_count elements of src taking one out of d (re_set_select, ulsch_demultiplex_impl.cpp:78-102). */
static void re_select(re_set* dst, const re_set* src, int d, int m_re_count)
{
  re_fill(dst, src->size, 0);
  int count = 0, d_count = 0;
  for (int i = 0; i != src->size && count != m_re_count; ++i) {
    if (!src->v[i]) {
      continue;
    }
    if (d_count % d == 0) {
      dst->v[i] = 1;
      ++count;
    }
    ++d_count;
  }
}

static int ulsch_l1(int mask)
{
  int first = -1;
  for (int l = 0; l != 14; ++l) {
    if ((mask >> l) & 1) {
      first = l;
      break;
    }
  }
  if (first < 0) {
    return -1;
  }
  for (int l = first; l != 14; ++l) {
    if (!((mask >> l) & 1)) {
      return l;
    }
  }
  return -1;
}
static int ulsch_l1_csi(int mask)
{
  for (int l = 0; l != 14; ++l) {
    if (!((mask >> l) & 1)) {
      return l;
    }
  }
  return -1;
}
static int ulsch_re_dmrs_symbol(const orc_ulsch_cfg* c)
{
  int per_prb = c->nof_cdm_groups_without_data * (c->dmrs_type == 1 ? 6 : 4);
  return (12 - per_prb) * c->nof_prb;
}

uint32_t orc_ulsch_codeword_length(const orc_ulsch_cfg* c)
{
  uint32_t total = 0;
  for (int l = c->start_symbol_index; l != c->start_symbol_index + c->nof_symbols; ++l) {
    int m = ((c->dmrs_symbol_mask >> l) & 1) ? ulsch_re_dmrs_symbol(c) : c->nof_prb * 12;
    total += (uint32_t)(m * c->qm * c->nof_layers);
  }
  return total;
}

/* on_uci_placeholder_1bit / _2bit (ulsch_demultiplex_impl.cpp:111-198): x placeholders get their scrambling undone, the
 * y placeholder of the 1-bit case takes the scrambling of the first bit. */
static void uci_placeholder(int8_t* out, const int8_t* data, const uint8_t* seq, int qm, int nbits, int nof_uci_bits)
{
  if (qm == 1) {
    memcpy(out, data, (size_t)nbits);
    return;
  }
  for (int s = 0; s != nbits / qm; ++s) {
    const int8_t*  d = data + s * qm;
    const uint8_t* m = seq + s * qm;
    int8_t*        o = out + s * qm;
    o[0]             = d[0];
    if (nof_uci_bits == 1) {
      o[1] = ((m[0] ^ m[1]) == 1) ? (int8_t)-d[1] : d[1];
    } else {
      o[1] = d[1];
    }
    for (int b = 2; b != qm; ++b) {
      o[b] = (m[b] == 1) ? (int8_t)-d[b] : d[b];
    }
  }
}

int orc_ulsch_demux(const orc_ulsch_cfg* c,
                    const int8_t*        in,
                    const uint8_t*       seq_bits,
                    uint32_t             n_in,
                    int8_t*              sch,
                    int8_t*              harq_ack,
                    int8_t*              csi_part1,
                    int8_t*              csi_part2,
                    uint32_t*            n_out)
{
  static re_set ulsch, uci, rvd, hack, csi1, csi2, tmp;
  static int8_t sym[ORC_MAX_RE * 8 * 4];
  const int     bpre   = c->qm * c->nof_layers;
  const int     l1     = ulsch_l1(c->dmrs_symbol_mask);
  const int     l1_csi = ulsch_l1_csi(c->dmrs_symbol_mask);
  if (l1 < 0 || l1_csi < 0 || bpre <= 0 || n_in != orc_ulsch_codeword_length(c)) {
    return -1;
  }
  int m_rvd = 0, m_hack = 0, m_csi1 = 0, m_csi2 = 0;
  int hack_open = c->nof_harq_ack_bits != 0, csi1_open = c->nof_csi_part1_bits != 0, csi2_open = 0;
  int csi2_bits = 0, csi2_enc = 0; /* configured when CSI Part 1 ends (set_csi_part2) */
  uint32_t pos = 0;
  n_out[0] = n_out[1] = n_out[2] = n_out[3] = 0;

  for (int l = c->start_symbol_index; l != c->start_symbol_index + c->nof_symbols; ++l) {
    /* configure_current_ofdm_symbol (:371-473) */
    const int dmrs    = (c->dmrs_symbol_mask >> l) & 1;
    const int M_ulsch = dmrs ? ulsch_re_dmrs_symbol(c) : c->nof_prb * 12;
    const int nsoft   = M_ulsch * bpre;
    if (nsoft == 0) {
      continue;
    }
    re_fill(&ulsch, M_ulsch, 1);
    re_fill(&uci, M_ulsch, !dmrs);
    re_fill(&rvd, M_ulsch, 0);
    re_fill(&hack, M_ulsch, 0);
    re_fill(&csi1, M_ulsch, 0);
    re_fill(&csi2, M_ulsch, 0);
    int M_uci    = re_count(&uci);
    int rem_rvd  = (int)((unsigned)(c->nof_harq_ack_rvd - m_rvd) / (unsigned)bpre);
    if (l >= l1 && M_uci > 0 && rem_rvd > 0) {
      int d = 1, m = M_uci;
      if (rem_rvd < M_uci) {
        d = M_uci / rem_rvd;
        m = rem_rvd;
      }
      re_select(&rvd, &ulsch, d, m);
      m_rvd += m * bpre;
    }
    const int rem_hack = (int)((unsigned)(c->nof_enc_harq_ack_bits - m_hack) / (unsigned)bpre);
    if (l >= l1 && M_uci > 0 && c->nof_harq_ack_bits > 2 && rem_hack > 0) {
      int d = 1, m = M_uci;
      if (rem_hack < M_uci) {
        d = M_uci / rem_hack;
        m = rem_hack;
      }
      re_select(&hack, &uci, d, m);
      for (int i = 0; i != M_ulsch; ++i) {
        if (hack.v[i]) {
          ulsch.v[i] = 0;
          uci.v[i]   = 0;
        }
      }
      M_uci = re_count(&uci);
      m_hack += m * bpre;
    }
    const int rem_csi1 = (int)((unsigned)(c->nof_enc_csi_part1_bits - m_csi1) / (unsigned)bpre);
    const int M_rvd    = re_count(&rvd);
    if (l >= l1_csi && (M_uci - M_rvd) > 0 && rem_csi1 > 0) {
      int d = 1, m = M_uci - M_rvd;
      if (rem_csi1 < M_uci - M_rvd) {
        d = (M_uci - M_rvd) / rem_csi1;
        m = rem_csi1;
      }
      re_fill(&tmp, M_ulsch, 0);
      for (int i = 0; i != M_ulsch; ++i) {
        tmp.v[i] = (uint8_t)(!rvd.v[i] && uci.v[i]);
      }
      re_select(&csi1, &tmp, d, m);
      for (int i = 0; i != M_ulsch; ++i) {
        if (csi1.v[i]) {
          ulsch.v[i] = 0;
          uci.v[i]   = 0;
        }
      }
      m_csi1 += m * bpre;
    }
    /* configure_csi_part2_current_ofdm_symbol (:475-499); also run when CSI Part 2 is set in the middle of the symbol. */
#define CONFIGURE_CSI2()                                                                                               \
  do {                                                                                                                 \
    int M_uci2   = re_count(&uci);                                                                                     \
    int rem_csi2 = (int)((unsigned)(csi2_enc - m_csi2) / (unsigned)bpre);                                              \
    if (l >= l1_csi && M_uci2 > 0 && rem_csi2 > 0) {                                                                   \
      int d2 = 1, m2 = M_uci2;                                                                                         \
      if (rem_csi2 < M_uci2) {                                                                                         \
        d2 = M_uci2 / rem_csi2;                                                                                        \
        m2 = rem_csi2;                                                                                                 \
      }                                                                                                                \
      re_select(&csi2, &uci, d2, m2);                                                                                  \
      for (int i = 0; i != M_ulsch; ++i) {                                                                             \
        if (csi2.v[i]) {                                                                                               \
          ulsch.v[i] = 0;                                                                                              \
          uci.v[i]   = 0;                                                                                              \
        }                                                                                                              \
      }                                                                                                                \
      m_csi2 += m2 * bpre;                                                                                             \
    }                                                                                                                  \
  } while (0)
    CONFIGURE_CSI2();
    if (M_rvd > 0 && c->nof_harq_ack_bits <= 2 && rem_hack > 0) {
      int d = 1, m = M_rvd;
      if (rem_hack < M_rvd) {
        d = M_rvd / rem_hack;
        m = rem_hack;
      }
      re_select(&hack, &rvd, d, m);
      m_hack += m * bpre;
    }

    /* demux_current_ofdm_symbol (:501-589) on a copy of the symbol (HARQ-ACK placeholders are zeroed in place). */
    memcpy(sym, in + pos, (size_t)nsoft);
    const uint8_t* seq = seq_bits + pos;
    if (re_count(&hack) != 0) {
      if (!hack_open) {
        return -1;
      }
      for (int i = 0; i != M_ulsch; ++i) {
        if (!hack.v[i]) {
          continue;
        }
        int8_t* re = sym + i * bpre;
        if (c->nof_harq_ack_bits == 1 || c->nof_harq_ack_bits == 2) {
          uci_placeholder(harq_ack + n_out[1], re, seq + i * bpre, c->qm, bpre, c->nof_harq_ack_bits);
          memset(re, 0, (size_t)bpre);
        } else {
          memcpy(harq_ack + n_out[1], re, (size_t)bpre);
        }
        n_out[1] += (uint32_t)bpre;
      }
      if (m_hack == c->nof_enc_harq_ack_bits) {
        hack_open = 0;
      }
    }
    if (re_count(&csi1) != 0) {
      if (!csi1_open) {
        return -1;
      }
      for (int i = 0; i != M_ulsch; ++i) {
        if (!csi1.v[i]) {
          continue;
        }
        const int8_t* re = sym + i * bpre;
        if (c->nof_csi_part1_bits == 1 || c->nof_csi_part1_bits == 2) {
          uci_placeholder(csi_part1 + n_out[2], re, seq + i * bpre, c->qm, bpre, c->nof_csi_part1_bits);
        } else {
          memcpy(csi_part1 + n_out[2], re, (size_t)bpre);
        }
        n_out[2] += (uint32_t)bpre;
      }
      if (m_csi1 == c->nof_enc_csi_part1_bits) {
        csi1_open = 0;
        /* CSI Part 1 decoded: the processor configures CSI Part 2 now, for the symbol being demultiplexed. */
        if (c->nof_enc_csi_part2_bits != 0) {
          csi2_open = 1;
          csi2_bits = c->nof_csi_part2_bits;
          csi2_enc  = c->nof_enc_csi_part2_bits;
          CONFIGURE_CSI2();
        }
      }
    }
    if (re_count(&csi2) != 0) {
      if (!csi2_open) {
        return -1;
      }
      for (int i = 0; i != M_ulsch; ++i) {
        if (!csi2.v[i]) {
          continue;
        }
        const int8_t* re = sym + i * bpre;
        if (csi2_bits == 1 || csi2_bits == 2) {
          uci_placeholder(csi_part2 + n_out[3], re, seq + i * bpre, c->qm, bpre, csi2_bits);
        } else {
          memcpy(csi_part2 + n_out[3], re, (size_t)bpre);
        }
        n_out[3] += (uint32_t)bpre;
      }
      if (m_csi2 == csi2_enc) {
        csi2_open = 0;
      }
    }
    for (int i = 0; i != M_ulsch; ++i) {
      if (ulsch.v[i]) {
        memcpy(sch + n_out[0], sym + i * bpre, (size_t)bpre);
        n_out[0] += (uint32_t)bpre;
      }
    }
    pos += (uint32_t)nsoft;
  }
#undef CONFIGURE_CSI2
  return (hack_open || csi1_open || csi2_open) ? -1 : 0;
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* Soft demapper - lib/phy/upper/channel_modulation/demodulation_mapper_*.cpp.                      */
/*                                                                                                                    */
/* The reference chooses its kernels at COMPILE time: on an x86 build (-mavx2 -mfma, what oracle/Makefile and the       */
/* reference's AUTO_DETECT_ISA build use) whole blocks of 16/8/16/4 symbols (QPSK/16QAM/64QAM/256QAM) of every          */
/* demodulate_soft call take the AVX2 kernel and the remainder a scalar loop whose arithmetic differs (division instead */
/* of reciprocal multiplication, round-half-away instead of round-half-even, near-zero test on |z|^2 instead of per     */
/* component). An AVX512 build adds wider blocks with the same per-element arithmetic, hence the same output.           */
/* GCC contracts "slope * value + intercept" into one fused multiply-add in both paths (checked in the object code of   */
/* oracle/_ref); fmaf() below is that contraction, every other operation is a separate single-precision operation       */
/* (this file is compiled with -ffp-contract=off).                                                                     */
/* ------------------------------------------------------------------------------------------------------------------ */

/* cvttss2si / cvtps2dq of an integer-valued float: out-of-range and NaN give the "integer indefinite" 0x80000000. */
static int32_t x86_f2i(float v)
{
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) {
    return INT32_MIN;
  }
  return (int32_t)v;
}

/* log_likelihood_ratio::quantize (lib/phy/upper/log_likelihood_ratio.cpp:89-98); the cast of the rounded value goes
 * through cvttss2si and keeps the low byte, so a NaN becomes 0. */
static int8_t quantize_scalar(float value, float range_limit)
{
  float clipped = value;
  if (fabsf(value) > range_limit) {
    clipped = copysignf(range_limit, value);
  }
  float q = clipped / range_limit;
  q       = q * (float)LLR_MAX;
  return (int8_t)(uint8_t)(uint32_t)x86_f2i(roundf(q));
}

/* mm256::quantize_ps (lib/phy/upper/channel_modulation/avx2_helpers.h:121-166), one element. */
static int8_t quantize_simd(float value, float range_limit)
{
  float scale = (float)LLR_MAX / range_limit;
  float v     = value * scale;
  /* clip_ps: ordered compares, a NaN passes through. */
  if (v > (float)LLR_MAX) {
    v = (float)LLR_MAX;
  }
  if (v < (float)-LLR_MAX) {
    v = (float)-LLR_MAX;
  }
  v         = nearbyintf(v); /* _MM_FROUND_NINT: to nearest, ties to even (default rounding mode) */
  int32_t i = x86_f2i(v);
  if (i > LLR_MAX || i < -LLR_MAX) { /* check_bounds_epi32: NaN (integer indefinite) -> 0 */
    i = 0;
  }
  return (int8_t)i;
}

typedef struct {
  float width;
  int   n;
  float slope[16];
  float icpt[16];
} itab_t;

/* Tables of demodulation_mapper_qam64.cpp:43-80 and demodulation_mapper_qam256.cpp:43-165: slopes are small integers
 * times 1/sqrt(42) or 1/sqrt(170) (single-precision product), intercepts integers over 21 or 85. */
static void make_itab(itab_t* t, float unit, int width_units, int n, const int* slope_k, const int* icpt_num, float den)
{
  t->width = (float)width_units * unit;
  t->n     = n;
  for (int i = 0; i != n; ++i) {
    t->slope[i] = (float)slope_k[i] * unit;
    t->icpt[i]  = (float)icpt_num[i] / den;
  }
}

static itab_t g_q64[3], g_q256[4];
static int    g_itab_ready;

static void init_itabs(void)
{
  if (g_itab_ready) {
    return;
  }
  const float u42 = 1.0f / sqrtf(42.0f), u170 = 1.0f / sqrtf(170.0f);
  static const int s64_01[8] = {16, 12, 8, 4, 4, 8, 12, 16}, i64_01[8] = {24, 12, 4, 0, 0, -4, -12, -24};
  static const int s64_23[8] = {8, 4, 4, 8, -8, -4, -4, -8}, i64_23[8] = {20, 8, 8, 12, 12, 8, 8, 20};
  static const int s64_45[4] = {4, -4, 4, -4}, i64_45[4] = {12, -4, -4, 12};
  make_itab(&g_q64[0], u42, 2, 8, s64_01, i64_01, 21.0f);
  make_itab(&g_q64[1], u42, 2, 8, s64_23, i64_23, 21.0f);
  make_itab(&g_q64[2], u42, 4, 4, s64_45, i64_45, 21.0f);
  static const int s256_01[16] = {32, 28, 24, 20, 16, 12, 8, 4, 4, 8, 12, 16, 20, 24, 28, 32};
  static const int i256_01[16] = {112, 84, 60, 40, 24, 12, 4, 0, 0, -4, -12, -24, -40, -60, -84, -112};
  static const int s256_23[16] = {16, 12, 8, 4, 4, 8, 12, 16, -16, -12, -8, -4, -4, -8, -12, -16};
  static const int i256_23[16] = {88, 60, 36, 16, 16, 28, 36, 40, 40, 36, 28, 16, 16, 36, 60, 88};
  static const int s256_45[16] = {8, 4, 4, 8, -8, -4, -4, -8, 8, 4, 4, 8, -8, -4, -4, -8};
  static const int i256_45[16] = {52, 24, 24, 44, -20, -8, -8, -12, -12, -8, -8, -20, 44, 24, 24, 52};
  static const int s256_67[8]  = {4, -4, 4, -4, 4, -4, 4, -4};
  static const int i256_67[8]  = {28, -20, 12, -4, -4, 12, -20, 28};
  make_itab(&g_q256[0], u170, 2, 16, s256_01, i256_01, 85.0f);
  make_itab(&g_q256[1], u170, 2, 16, s256_23, i256_23, 85.0f);
  make_itab(&g_q256[2], u170, 2, 16, s256_45, i256_45, 85.0f);
  make_itab(&g_q256[3], u170, 4, 8, s256_67, i256_67, 85.0f);
  g_itab_ready = 1;
}

static int clamp_idx(int32_t idx, int n)
{
  return (idx < 0) ? 0 : (idx > n - 1) ? n - 1 : idx;
}

/* interval_function (demodulation_mapper_intervals.h:31-63): index from a DIVISION by the interval width. */
static float interval_scalar(float value, float rcp_noise, const itab_t* t)
{
  float   q   = value / t->width;
  int32_t idx = (int32_t)((uint32_t)x86_f2i(floorf(q)) + (uint32_t)(t->n / 2));
  int     k   = clamp_idx(idx, t->n);
  float   l   = fmaf(t->slope[k], value, t->icpt[k]);
  return l * rcp_noise;
}

/* mm256::interval_function (avx2_helpers.h:234-254): index from a MULTIPLICATION by 1 / width, result forced to zero for
 * |value| <= 1e-9 (ordered compare). */
static float interval_simd(float value, float rcp_noise, const itab_t* t)
{
  float   inv = 1.0f / t->width;
  float   q   = value * inv;
  int32_t idx = (int32_t)((uint32_t)x86_f2i(floorf(q)) + (uint32_t)(t->n / 2));
  int     k   = clamp_idx(idx, t->n);
  float   l   = fmaf(t->slope[k], value, t->icpt[k]);
  l           = l * rcp_noise;
  if (fabsf(value) <= 1e-9f) {
    l = 0.0f;
  }
  return l;
}

/* safe_div(1, noise) (avx2_helpers.h:259-271) and the scalar "if (noise > 0) rcp = 1 / noise". */
static float rcp_noise_of(float nv)
{
  return (nv > 0.0f) ? 1.0f / nv : 0.0f;
}

/* is_near_zero(cf_t) (include/srsran/support/math_utils.h:91-94): |z|^2 < 1e-9, |z|^2 evaluated as fma(re, re, im * im). */
static int near_zero_cf(float re, float im)
{
  float t = im * im;
  t       = fmaf(re, re, t);
  return 1e-9f > t;
}

static void demod_symbol_simd(int8_t* out, float re, float im, float nv, int mod)
{
  const float rcp = rcp_noise_of(nv);
  const float c[2] = {re, im};
  if (mod == 2) {
    /* demod_QPSK_avx2 (demodulation_mapper_qpsk.cpp:39-78). */
    const float gain = 2.0f * 1.41421356237309504880f;
    for (int k = 0; k != 2; ++k) {
      float l = gain * c[k];
      out[k]  = quantize_simd(l * rcp, 24.0f);
    }
  } else if (mod == 4) {
    /* demod_QAM16_avx2 (demodulation_mapper_qam16.cpp:41-116). */
    const float u10 = 1.0f / sqrtf(10.0f), gain = 4.0f * u10, thr = 2.0f * u10;
    for (int k = 0; k != 2; ++k) {
      float first  = gain * c[k];
      float second = 2.0f * first - copysignf(0.8f, c[k]);
      float l01    = (fabsf(c[k]) > thr) ? second : first;
      float l23    = 0.8f - fabsf(first);
      l01 *= rcp;
      l23 *= rcp;
      if (fabsf(c[k]) <= 1e-9f) {
        l01 = 0.0f;
        l23 = 0.0f;
      }
      out[k]     = quantize_simd(l01, 20.0f);
      out[2 + k] = quantize_simd(l23, 20.0f);
    }
  } else if (mod == 6) {
    for (int g = 0; g != 3; ++g) {
      for (int k = 0; k != 2; ++k) {
        out[2 * g + k] = quantize_simd(interval_simd(c[k], rcp, &g_q64[g]), 20.0f);
      }
    }
  } else {
    for (int g = 0; g != 4; ++g) {
      for (int k = 0; k != 2; ++k) {
        out[2 * g + k] = quantize_simd(interval_simd(c[k], rcp, &g_q256[g]), 20.0f);
      }
    }
  }
}

static void demod_symbol_scalar(int8_t* out, float re, float im, float nv, int mod)
{
  const float c[2] = {re, im};
  if (mod == 2) {
    /* demod_QPSK_symbol (demodulation_mapper_qpsk.cpp:121-129). */
    const float gain = 2.0f * 1.41421356237309504880f;
    for (int k = 0; k != 2; ++k) {
      if (!(nv > 0.0f)) {
        out[k] = 0;
      } else {
        float l = gain * c[k];
        out[k]  = quantize_scalar(l / nv, 24.0f);
      }
    }
    return;
  }
  if (near_zero_cf(re, im)) {
    memset(out, 0, (size_t)mod);
    return;
  }
  if (mod == 4) {
    /* demod_16QAM_symbol_01/_23 (demodulation_mapper_qam16.cpp:192-222): "0.8 - gain * |x|" is one fused operation. */
    const float u10 = 1.0f / sqrtf(10.0f), gain = 4.0f * u10, thr = 2.0f * u10;
    for (int k = 0; k != 2; ++k) {
      if (!(nv > 0.0f)) {
        out[k] = out[2 + k] = 0;
        continue;
      }
      float l01 = gain * c[k];
      if (fabsf(c[k]) > thr) {
        l01 = 2.0f * l01 - copysignf(0.8f, c[k]);
      }
      float l23  = fmaf(-gain, fabsf(c[k]), 0.8f);
      out[k]     = quantize_scalar(l01 / nv, 20.0f);
      out[2 + k] = quantize_scalar(l23 / nv, 20.0f);
    }
    return;
  }
  const float   rcp = rcp_noise_of(nv);
  const itab_t* t   = (mod == 6) ? g_q64 : g_q256;
  for (int g = 0; g != mod / 2; ++g) {
    for (int k = 0; k != 2; ++k) {
      out[2 * g + k] = quantize_scalar(interval_scalar(c[k], rcp, &t[g]), 20.0f);
    }
  }
}

/* demod_BPSK_symbol (demodulation_mapper_impl.cpp:34-42). */
static int8_t demod_bpsk(float re, float im, float nv)
{
  if (!(nv > 0.0f)) {
    return 0;
  }
  float l = re + im;
  l       = l * (2.0f * 1.41421356237309504880f);
  return quantize_scalar(l / nv, 24.0f);
}

void orc_demodulate_soft(int8_t* llr, const float* symbols, const float* noise_vars, uint32_t n, int mod, int simd)
{
  init_itabs();
  if (mod == ORC_MOD_BPSK || mod == ORC_MOD_PI_2_BPSK) {
    /* demodulate_soft_BPSK / _PI_2_BPSK (demodulation_mapper_impl.cpp:44-76): odd symbols of pi/2-BPSK are rotated by
     * -90 degrees, i.e. re + im becomes im - re. */
    for (uint32_t i = 0; i != n; ++i) {
      float re = symbols[2 * i], im = symbols[2 * i + 1];
      if (mod == ORC_MOD_PI_2_BPSK && (i & 1)) {
        float t = re;
        re      = im;
        im      = -t;
      }
      llr[i] = demod_bpsk(re, im, noise_vars[i]);
    }
    return;
  }
  const uint32_t block  = (mod == 2) ? 16 : (mod == 4) ? 8 : (mod == 6) ? 16 : 4;
  const uint32_t n_simd = simd ? (n / block) * block : 0;
  for (uint32_t i = 0; i != n; ++i) {
    if (i < n_simd) {
      demod_symbol_simd(llr + (size_t)i * mod, symbols[2 * i], symbols[2 * i + 1], noise_vars[i], mod);
    } else {
      demod_symbol_scalar(llr + (size_t)i * mod, symbols[2 * i], symbols[2 * i + 1], noise_vars[i], mod);
    }
  }
}
