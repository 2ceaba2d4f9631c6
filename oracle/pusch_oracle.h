/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's uplink PUSCH decode path (rate dematching + HARQ combining, layered
 * normalized-min-sum LDPC decoding with CRC early stop, CB/TB CRC, TB assembly). It exists only to CHECK the CUDA
 * path: it may be called from tests/, from __graft_entry__.smoke() and from the cpu_baseline / --impl reference legs of
 * bench.py, never from the product.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py compares every function below with the reference itself
 * (oracle/_ref/libsrsref.so, built from /root/reference by oracle/Makefile) and tests/golden/ holds vectors produced by
 * that library, so the restatement is also checked where /root/reference does not exist.
 *
 * File references are relative to /root/reference/srsRAN-5G-ER/.
 */
#ifndef PUSCH_ORACLE_H
#define PUSCH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Check-to-variable scaling rule (the one arithmetic difference between the reference's decoder variants). */
#define ORC_SCALE_X86 0     /* (x * 52428) >> 16   - ldpc/avx2_support.h:65-106, ldpc/avx512_support.h:65-107 */
#define ORC_SCALE_GENERIC 1 /* round(x * 0.8f)     - ldpc/ldpc_decoder_generic.cpp:70-79                       */
#define ORC_SCALE_NEON 2    /* (x * 204) >> 8      - ldpc/neon_support.h:64-100                                */

/* CRC kinds. */
#define ORC_CRC_NONE 0
#define ORC_CRC16 1
#define ORC_CRC24A 2
#define ORC_CRC24B 3
#define ORC_CRC24C 4 /* the remaining polynomials of crc_calculator.h:35-48, for the stand-alone CRC only */
#define ORC_CRC11 5
#define ORC_CRC6 6

/* crc_calculator::calculate - lib/phy/upper/channel_coding/crc_calculator_generic_impl.cpp:111-133.
 * Remainder of the first nbits (MSB first) of packed[] followed by `order` zero bits. */
uint32_t orc_crc(int crc_kind, const uint8_t* packed, int nbits);

/* ldpc_rate_dematcher_impl::rate_dematch - ldpc/ldpc_rate_dematcher_impl.cpp:46-114 with allot_llrs :128-201 and the
 * deinterleaver :203-257. out[N] is the in/out HARQ buffer. simd_width = 64 (avx512), 32 (avx2) or 0 (generic): the
 * SIMD variants combine whole blocks without the infinity rules of log_likelihood_ratio::operator+ and handle only
 * the tail of each chunk with them (ldpc_rate_dematcher_avx512_impl.cpp:29-64); this matters only for non-finite stale
 * buffer contents. */
void orc_rate_dematch(int8_t*       out,
                      int           N,
                      const int8_t* in,
                      int           E,
                      int           new_data,
                      int           rv,
                      int           qm,
                      int           nref,
                      int           nof_filler,
                      int           simd_width);

/* ldpc_decoder_impl::decode - ldpc/ldpc_decoder_impl.cpp:60-147 (+ :149-318 and the per-variant kernels).
 * Returns the iteration count when the CRC passed, 0 otherwise (the reference's empty optional). out holds ceil(K/8)
 * bytes, bits MSB first. soft_out, when not NULL, receives the final N_full*Z soft bits (debug aid). */
int orc_ldpc_decode(int           bg,
                    int           Z,
                    const int8_t* in,
                    int           n_in,
                    int           nof_filler,
                    int           crc_kind,
                    int           max_iter,
                    int           scale_mode,
                    uint8_t*      out,
                    int8_t*       soft_out);

/* pusch_codeblock_decoder::decode - lib/phy/upper/channel_processors/pusch/pusch_codeblock_decoder.cpp:35-71. */
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
                  uint8_t*      out);

/* One codeblock of the rx segmentation. */
typedef struct {
  int Z;
  int full_length;  /* N = 66Z | 50Z */
  int rm_length;    /* E             */
  int nof_filler;   /* F             */
  int cw_offset;
  int nof_crc_bits; /* 16 | 24       */
} orc_cb_meta;

/* ldpc_segmenter_impl::segment (rx) - ldpc/ldpc_segmenter_impl.cpp:254-331 with the helpers of
 * include/srsran/phy/upper/channel_coding/ldpc/ldpc.h:128-228. Returns the number of codeblocks (<= 162?  the caller
 * provides room for 256). */
int orc_segment_rx(int tbs_bits, int bg, int qm, int nof_layers, int n_llr, orc_cb_meta* meta);

/* HARQ buffer of one (rnti, harq_id): what rx_buffer holds (include/srsran/phy/upper/rx_buffer.h:42-81). */
typedef struct {
  int      nof_cb;
  int8_t*  soft;   /* nof_cb x 25344 */
  uint8_t* data;   /* nof_cb x 1056  */
  uint8_t* crc_ok; /* nof_cb         */
} orc_harq;

typedef struct {
  int bg, rv, qm, nref, nof_layers, max_iter, use_early_stop, new_data;
} orc_pusch_cfg;

/* pusch_decoder_impl: new_data :89-138, fork_codeblock_task :309-382, join_and_notify :384-450,
 * concatenate_codeblocks :452-497 (lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.cpp).
 * stats = {tb_crc_ok, nof_codeblocks_total, nof_observations, min_iter, max_iter, sum_iter}. */
void orc_pusch_decode(orc_harq*            harq,
                      const int8_t*        llrs,
                      int                  n_llr,
                      int                  tb_bytes,
                      const orc_pusch_cfg* cfg,
                      int                  scale_mode,
                      int                  simd_width,
                      uint8_t*             tb_out,
                      int*                 stats);

/* Systematic 5G NR LDPC encoder (TS 38.212 5.3.2), used only to synthesise valid codewords for tests and benches.
 * msg[K] one bit per byte (fillers as 0), cw[N_full*Z - 2Z] one bit per byte (the 2Z punctured bits removed). */
void orc_ldpc_encode(int bg, int Z, const uint8_t* msg, uint8_t* cw);

/* TX rate matching (bit selection + interleaving, TS 38.212 5.4.2) of one codeblock, for synthesising inputs.
 * cw[N] one bit per byte with filler positions marked by the caller through nof_filler; out[E] one bit per byte. */
void orc_rate_match(const uint8_t* cw, int N, int E, int rv, int qm, int nref, int nof_filler, uint8_t* out);

/* Complete TX chain for a transport block: CRC attachment, segmentation, encoding, rate matching, concatenation.
 * Returns the number of codeblocks; cw_bits[n_llr] one bit per byte. */
int orc_tb_encode(const uint8_t* tb, int tb_bytes, int bg, int rv, int qm, int nref, int nof_layers, int n_llr,
                  uint8_t* cw_bits);

/* Timing helper for bench.py's cpu_baseline ("port"): decodes n_cb codeblocks of one shape, single thread.
 * Returns seconds. */
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
                          int*          iters_out);


/* ---- codeword front end: scrambling sequence, descrambling, UL-SCH demultiplexing ---------------- */

/* TS 38.211 5.2.1 pseudo-random sequence c(offset .. offset+n-1) for c_init, one bit per byte
 * (lib/phy/upper/sequence_generators/pseudo_random_generator_impl.cpp:47-68 init/advance, generate). */
void orc_prg_bits(uint32_t c_init, uint32_t offset, uint32_t n, uint8_t* bits);

/* out[i] = seq[i] ? -in[i] : in[i] with two's-complement wrap of -128
 * (revert_scrambling, lib/phy/upper/channel_processors/pusch/pusch_demodulator_impl.cpp:38-128). */
void orc_revert_scrambling(int8_t* out, const int8_t* in, const uint8_t* seq_bits, uint32_t n);

/* ulsch_demultiplex::configuration (include/srsran/phy/upper/channel_processors/pusch/ulsch_demultiplex.h:46-76) plus
 * the CSI Part 2 sizes the PUSCH processor hands over when CSI Part 1 has been decoded (pusch_processor_impl.cpp:61-82;
 * 0 = no CSI Part 2). */
typedef struct {
  int qm;                          /* bits per symbol of the modulation                          */
  int nof_layers;
  int nof_prb;
  int start_symbol_index;
  int nof_symbols;
  int nof_harq_ack_rvd;
  int dmrs_type;                   /* 1 or 2                                                     */
  int dmrs_symbol_mask;            /* bit l set: OFDM symbol l of the slot carries DM-RS         */
  int nof_cdm_groups_without_data;
  int nof_harq_ack_bits;
  int nof_enc_harq_ack_bits;
  int nof_csi_part1_bits;
  int nof_enc_csi_part1_bits;
  int nof_csi_part2_bits;
  int nof_enc_csi_part2_bits;
} orc_ulsch_cfg;

/* Number of soft bits of the codeword (all OFDM symbols of the allocation). */
uint32_t orc_ulsch_codeword_length(const orc_ulsch_cfg* cfg);

/* ulsch_demultiplex_impl (lib/phy/upper/channel_processors/pusch/ulsch_demultiplex_impl.cpp:200-589): in[] are the
 * descrambled soft bits of the codeword in resource-element order, seq_bits[] the scrambling sequence applied to them
 * (needed to undo the scrambling of the UCI placeholders). Outputs are the soft-bit streams handed to the four decoder
 * buffers; n_out[0..3] = lengths of sch, harq_ack, csi_part1, csi_part2. Returns 0, or -1 on an inconsistent
 * configuration (where the reference asserts). */
int orc_ulsch_demux(const orc_ulsch_cfg* cfg,
                    const int8_t*        in,
                    const uint8_t*       seq_bits,
                    uint32_t             n_in,
                    int8_t*              sch,
                    int8_t*              harq_ack,
                    int8_t*              csi_part1,
                    int8_t*              csi_part2,
                    uint32_t*            n_out);

/* ---- soft demapper ------------------------------------------------------------------------------ */

#define ORC_MOD_PI_2_BPSK 0
#define ORC_MOD_BPSK 1 /* 2, 4, 6, 8: QPSK, 16QAM, 64QAM, 256QAM (bits per symbol) */

/* One demodulation_mapper::demodulate_soft call (lib/phy/upper/channel_modulation/demodulation_mapper_impl.cpp:78-106):
 * symbols = n interleaved {re, im} pairs (equaliser output), noise_vars = n post-equalisation noise variances,
 * llr = n * bits-per-symbol soft bits. simd != 0: the x86 build of the reference (AVX2 kernels on whole blocks of the
 * call, scalar remainder); simd == 0: the portable build (scalar loop only). */
void orc_demodulate_soft(int8_t* llr, const float* symbols, const float* noise_vars, uint32_t n, int mod, int simd);

#ifdef __cplusplus
}
#endif
#endif
