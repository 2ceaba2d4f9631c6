/*
 * pusch_dec_cuda.h - C ABI of the B200 (sm_100a) uplink PUSCH decode path.
 *
 * Rate dematching + HARQ soft combining -> 5G NR LDPC layered normalized min-sum decoding with CRC early stop ->
 * codeblock / transport-block CRC, batched over every codeblock of a slot. Plain pointers and sizes only; no C++,
 * CUDA runtime or torch types cross this boundary. Every function returns PDC_OK (0) or a negative error code and never
 * throws; there is no CPU fallback - if no CUDA device is usable, pdc_create fails.
 *
 * Each entry point names the reference interface it stands behind (paths relative to /root/reference/srsRAN-5G-ER/):
 *
 *   ldpc_decoder::decode                  include/srsran/phy/upper/channel_coding/ldpc/ldpc_decoder.h:73-74
 *   ldpc_rate_dematcher::rate_dematch     include/srsran/phy/upper/channel_coding/ldpc/ldpc_rate_dematcher.h:52-55
 *   crc_calculator::calculate             include/srsran/phy/upper/channel_coding/crc_calculator.h:62-84
 *   hal::hw_accelerator_pusch_dec         include/srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec.h:36-115
 *   hal::hw_accelerator<int8_t,uint8_t>   include/srsran/hal/hw_accelerator.h:35-57
 *   pusch_decoder (batched)               include/srsran/phy/upper/channel_processors/pusch/pusch_decoder.h:54-99
 *   rx_buffer (HARQ state)                include/srsran/phy/upper/rx_buffer.h:42-81
 *   ulsch_demultiplex / pusch_codeword_buffer (codeword front end)
 *                                         include/srsran/phy/upper/channel_processors/pusch/ulsch_demultiplex.h:41-103,
 *                                         include/srsran/phy/upper/channel_processors/pusch/pusch_codeword_buffer.h
 *   pseudo_random_generator::generate     include/srsran/phy/upper/sequence_generators/pseudo_random_generator.h
 *   ldpc_encoder::encode                  include/srsran/phy/upper/channel_coding/ldpc/ldpc_encoder.h:37-38 (downlink twin)
 *   ldpc_rate_matcher::rate_match         include/srsran/phy/upper/channel_coding/ldpc/ldpc_rate_matcher.h:37
 *   demodulation_mapper::demodulate_soft  include/srsran/phy/upper/channel_modulation/demodulation_mapper.h:62-65
 *                                         (factory: channel_modulation_factories.h:32-41)
 *
 * Threading: each batch queue ("stream") is driven by one thread at a time (pdc_submit .. pdc_wait of a queue may come
 * from different threads one after the other; a thread that submits to a queue another thread is submitting to gets
 * PDC_ERR_CAPACITY). Different queues are independent. The synchronous single-object calls (pdc_ldpc_decode,
 * pdc_rate_dematch, pdc_crc, pdc_ulsch_demux, pdc_scrambling_sequence, pdc_demodulate_soft, pdc_encode,
 * pdc_ldpc_encode) may be called from any number of threads - the reference runs pools of decoder / dematcher objects
 * concurrently (concurrent_thread_local_object_pool.h:41-110): they share one private queue and one set of scratch
 * buffers inside the context and serialise on a mutex there; they never use the batch queues, so they cannot collide
 * with a batch in flight. pdc_launch_codewords_device / pdc_launch_demod_device (caller's stream, plan buffers shared
 * with the synchronous calls) remain one call at a time per context.
 *
 * INTEGRATION.md shows the C++ adapter classes that bind these calls behind create_ldpc_decoder_factory_sw("cuda"),
 * create_ldpc_rate_dematcher_factory_sw("cuda") and the hal::hw_accelerator_pusch_dec factory.
 */
#ifndef PUSCH_DEC_CUDA_H
#define PUSCH_DEC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDC_OK 0
#define PDC_ERR_INVALID (-1)   /* bad argument or descriptor */
#define PDC_ERR_CUDA (-2)      /* CUDA runtime error; pdc_last_error() has the text */
#define PDC_ERR_CAPACITY (-3)  /* batch larger than the context was created for ("queue full": retry after a wait) */
#define PDC_ERR_NO_DEVICE (-4) /* no usable sm_100 device */

/* Largest lifted codeblock: 66 x 384 soft bits, 22 x 384 message bits. */
#define PDC_MAX_CB_SOFT 25344
#define PDC_MAX_CB_BYTES 1056

/* CRC attached to a codeblock (hal::hw_dec_cb_crc_type, hw_accelerator_pusch_dec.h:36). */
#define PDC_CRC_NONE 0
#define PDC_CRC16 1
#define PDC_CRC24A 2
#define PDC_CRC24B 3
/* The remaining NR polynomials, accepted by pdc_crc only (crc_calculator.h:35-48). */
#define PDC_CRC24C 4
#define PDC_CRC11 5
#define PDC_CRC6 6

/* Check-to-variable scaling rule = which reference decoder variant is reproduced bit for bit. */
#define PDC_SCALE_X86 0     /* ldpc_decoder_avx2 / ldpc_decoder_avx512 ("auto" on any x86 host) - default */
#define PDC_SCALE_GENERIC 1 /* ldpc_decoder_generic */
#define PDC_SCALE_NEON 2    /* ldpc_decoder_neon */

/* Per-codeblock flags. */
#define PDC_CB_NEW_DATA 0x01   /* first transmission: copy instead of combine (rate_dematch new_data = true)          */
#define PDC_CB_EARLY_STOP 0x02 /* check the CRC after every iteration (decode with a crc_calculator)                 */
#define PDC_CB_DECODE 0x04     /* run the LDPC decoder; clear = dematch/combine only (codeblock already CRC-ok,
                                  pusch_decoder_impl.cpp:335-345)                                                    */
#define PDC_CB_DEMATCH 0x08    /* run the rate dematcher; clear = decode the HARQ entry as it is                     */

/*
 * One codeblock operation. Carries the fields of hal::hw_pusch_decoder_configuration
 * (hw_accelerator_pusch_dec.h:39-72) / codeblock_metadata (include/srsran/phy/upper/codeblock_metadata.h:42-80).
 */
typedef struct {
  uint32_t llr_offset;   /* first rate-matched LLR of this codeblock inside the batch LLR buffer                      */
  uint32_t rm_length;    /* E: number of rate-matched LLRs (multiple of qm)                                           */
  uint32_t harq_id;      /* absolute codeblock id = entry of the device HARQ arena (rx_buffer.h:58-65)                */
  uint32_t nref;         /* limited-buffer length N_ref, 0 = unlimited                                                */
  uint16_t lifting_size; /* Z                                                                                         */
  uint16_t nof_filler;   /* F                                                                                         */
  uint8_t  base_graph;   /* 1 or 2                                                                                    */
  uint8_t  qm;           /* bits per symbol: 1, 2, 4, 6, 8                                                            */
  uint8_t  rv;           /* redundancy version 0..3                                                                   */
  uint8_t  crc_kind;     /* PDC_CRC16 / PDC_CRC24A / PDC_CRC24B                                                       */
  uint8_t  max_iter;     /* LDPC iterations (ldpc_decoder::configuration::algorithm_details::max_iterations)          */
  uint8_t  flags;        /* PDC_CB_*                                                                                  */
  uint16_t tb_index;     /* transport block this codeblock belongs to (index into the pdc_tb_desc array), or 0xffff   */
} pdc_cb_desc;

/* Result of one codeblock (hal::hw_pusch_decoder_outputs, hw_accelerator_pusch_dec.h:75-80). */
typedef struct {
  uint8_t crc_ok;  /* CRC over the first K - F decoded bits is zero                                                   */
  uint8_t iters;   /* iterations run: the early-stop iteration, else max_iter                                         */
  uint8_t status;  /* 0 ok, 1 = all-zero input (not decodable, ldpc_decoder_impl.cpp:88-94), 2 = invalid descriptor   */
  uint8_t nlayers; /* base-graph rows actually processed (ldpc_decoder_impl.cpp:114)                                  */
} pdc_cb_result;

/*
 * One transport block of the batch: codeblocks first_cb .. first_cb + nof_cb - 1 of the descriptor array, in order.
 * The device concatenates the decoded codeblocks and checks the TB CRC exactly as pusch_decoder_impl::join_and_notify
 * does (pusch_decoder_impl.cpp:384-450, concatenate_codeblocks :452-497).
 */
typedef struct {
  uint32_t first_cb;
  uint32_t nof_cb;
  uint32_t tbs_bits;      /* transport block size without CRC (at most 2 097 088 bits; the largest NR TBS is 1 277 992) */
  uint32_t out_offset;    /* byte offset of this TB inside the batch TB output buffer                                 */
  uint32_t prev_ok_mask_offset; /* reserved, 0                                                                         */
} pdc_tb_desc;

typedef struct {
  uint8_t  tb_crc_ok;
  uint8_t  all_cb_ok;
  uint16_t reserved;
} pdc_tb_result;

/*
 * One PUSCH codeword of the front end: the soft bits of one UE in one slot as the soft demapper
 * emits them, in resource-element order over the OFDM symbols of the allocation. Carries the fields of
 * ulsch_demultiplex::configuration (include/srsran/phy/upper/channel_processors/pusch/ulsch_demultiplex.h:46-76), the
 * scrambling seed of pusch_demodulator_impl::demodulate (c_init = rnti * 2^15 + n_id, pusch_demodulator_impl.cpp:139-140)
 * and the CSI Part 2 sizes the PUSCH processor passes to set_csi_part2 once CSI Part 1 is decoded
 * (pusch_processor_impl.cpp:61-82; both 0 = no CSI Part 2).
 */
#define PDC_CW_SCRAMBLED 1u /* the input still carries the scrambling sequence: descramble on the device             */
#define PDC_CW_DEFER_DESCRAMBLING 2u /* with PDC_CW_SCRAMBLED, for a codeword without UCI: do not materialise its UL-SCH
                                        soft bits; the rate dematcher of the NEXT batch that reads its LLRs from that
                                        UL-SCH space (pdc_submit after pdc_submit_codewords; pdc_launch_device with
                                        d_llrs = the d_sch of pdc_launch_codewords_device) descrambles while it stages
                                        the codeblocks. The UL-SCH space itself is then left unwritten.              */

typedef struct {
  uint32_t in_offset;   /* first soft bit of the codeword in the input buffer                                         */
  uint32_t sch_offset;  /* where its UL-SCH soft bits start in the UL-SCH space (multiple of 4); pdc_cb_desc::llr_offset
                           of the transport block's codeblocks refers to that space                                   */
  uint32_t uci_offset;  /* where its UCI soft bits go in the UCI output: HARQ-ACK, CSI Part 1, CSI Part 2 back to back  */
  uint32_t c_init;
  uint32_t flags;       /* PDC_CW_*                                                                                    */
  uint8_t  qm;          /* bits per symbol of the modulation: 1, 2, 4, 6, 8                                            */
  uint8_t  nof_layers;  /* 1..4                                                                                        */
  uint8_t  start_symbol_index;
  uint8_t  nof_symbols;
  uint8_t  dmrs_type;   /* 1 or 2                                                                                      */
  uint8_t  nof_cdm_groups_without_data;
  uint16_t nof_prb;
  uint16_t dmrs_symbol_mask; /* bit l: OFDM symbol l of the slot carries DM-RS                                        */
  uint16_t reserved;
  uint32_t nof_harq_ack_rvd;
  uint32_t nof_harq_ack_bits;
  uint32_t nof_enc_harq_ack_bits;
  uint32_t nof_csi_part1_bits;
  uint32_t nof_enc_csi_part1_bits;
  uint32_t nof_csi_part2_bits;
  uint32_t nof_enc_csi_part2_bits;
} pdc_cw_desc;

/* Soft bits delivered to each of the four decoder buffers of ulsch_demultiplex::demultiplex / set_csi_part2. */
typedef struct {
  uint32_t n_sch;
  uint32_t n_harq_ack;
  uint32_t n_csi_part1;
  uint32_t n_csi_part2;
} pdc_cw_result;

/*
 * Soft demapper. Modulation codes = bits per symbol, 0 for pi/2-BPSK
 * (modulation_scheme, include/srsran/ran/sch/modulation_scheme.h).
 */
#define PDC_MOD_PI_2_BPSK 0
#define PDC_MOD_BPSK 1
#define PDC_MOD_QPSK 2
#define PDC_MOD_QAM16 4
#define PDC_MOD_QAM64 6
#define PDC_MOD_QAM256 8

/* Which build of the reference's demapper is reproduced bit for bit (its SIMD path is chosen at compile time). */
#define PDC_DEMOD_X86 0    /* x86 build (AVX2 / AVX512 kernels on whole blocks of a call, scalar remainder) - default */
#define PDC_DEMOD_SCALAR 1 /* portable build: the scalar loop for every symbol                                        */

#define PDC_CW_PLAIN_BPSK 16u /* pdc_submit_symbols, qm = 1: BPSK instead of the pi/2-BPSK that PUSCH uses            */

/* One demodulation_mapper::demodulate_soft call: n_sym symbols -> n_sym * bits-per-symbol soft bits. */
typedef struct {
  uint32_t sym_offset; /* first symbol of the call in the symbol / noise-variance arrays                               */
  uint32_t n_sym;
  uint32_t llr_offset; /* first soft bit of the call in the output                                                     */
  uint32_t modulation; /* PDC_MOD_*                                                                                    */
} pdc_demod_call;

/*
 * Downlink twin: one codeblock to encode and rate match. Carries the fields of codeblock_metadata
 * (include/srsran/phy/upper/codeblock_metadata.h:42-80) the encoder and the rate matcher read.
 */
/* pdc_enc_desc.flags: the codeblock's output is packed, eight bits per byte, first bit in the most significant bit:    */
/* (rm_length + 7) / 8 bytes at out_offset, the unused bits of the last byte zero (what a bit_buffer / the packed span  */
/* of hal::hw_accelerator_pdsch_enc::dequeue_operation holds): an eighth of the bytes on the way back to the host.       */
#define PDC_ENC_PACKED 1u

typedef struct {
  uint32_t msg_offset;   /* byte offset of the codeblock's K message bits (packed MSB first, filler bits as zeros)      */
  uint32_t out_offset;   /* where its rate-matched bits start in the output: byte offset (one bit per byte, or packed)  */
  uint32_t rm_length;    /* E: number of rate-matched bits (multiple of qm)                                             */
  uint32_t nref;         /* limited-buffer length N_ref, 0 = unlimited                                                  */
  uint16_t lifting_size; /* Z                                                                                           */
  uint16_t nof_filler;   /* F                                                                                           */
  uint8_t  base_graph;   /* 1 or 2                                                                                      */
  uint8_t  qm;           /* bits per symbol of the bit interleaver: 1, 2, 4, 6, 8                                       */
  uint8_t  rv;           /* redundancy version 0..3                                                                     */
  uint8_t  flags;        /* PDC_ENC_*                                                                                   */
} pdc_enc_desc;

typedef struct {
  int32_t  device;          /* CUDA device ordinal                                                                     */
  uint32_t max_cbs;         /* largest batch, in codeblocks                                                            */
  uint32_t max_llrs;        /* largest batch, in rate-matched LLRs                                                     */
  uint32_t harq_entries;    /* device HARQ arena size in codeblocks (25344 soft bits + 1056 message bytes each)        */
  uint32_t max_tbs;         /* largest number of transport blocks per batch (0 = codeblock interface only)             */
  uint32_t max_tb_bytes;    /* bytes of TB output per batch                                                            */
  int32_t  scale_mode;      /* PDC_SCALE_*                                                                             */
  int32_t  combine_simd_width; /* 64 / 32 / 0: which reference dematcher's treatment of non-finite stale soft bits is
                                  reproduced (avx512 / avx2 / generic); irrelevant for finite values                  */
  uint32_t nof_streams;     /* independent in-flight batches ("hardware queues", reserve_queue/free_queue)             */
  int32_t  demod_mode;      /* PDC_DEMOD_*                                                                             */
} pdc_config;

typedef struct pdc_ctx pdc_ctx;

/* ------------------------------------------------------------------------------------------------------------------ */
/* Context                                                                                                             */
/* ------------------------------------------------------------------------------------------------------------------ */

void pdc_default_config(pdc_config* cfg);
int  pdc_create(const pdc_config* cfg, pdc_ctx** out);
void pdc_destroy(pdc_ctx* ctx);
/* Text of the last error on this thread. */
const char* pdc_last_error(void);
/* Library and device facts: sm count, sm major/minor, number of kernel launches issued so far by this context. */
int pdc_device_info(pdc_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor);
uint64_t pdc_launch_count(pdc_ctx* ctx);

/*
 * Measures the 32-bit integer instruction throughput of the device in lane-operations per second (the decoder's
 * roofline denominator): mode 0 = LOP3/IADD3 only (ALU pipe), mode 1 = LOP3/IADD3 interleaved with IMAD (ALU + FMA pipes).
 */
int pdc_measure_int_peak(pdc_ctx* ctx, int mode, double* lane_ops_per_s);

/*
 * Debug builds of the library (-DPDC_DEBUG_BOUNDS: device-side asserts on every shared-memory / global index, canary words
 * behind every device allocation): 1 if every canary of the context's device is intact, 0 if one was overwritten.
 * Release builds answer -1 (nothing to check).
 */
int pdc_debug_canaries_ok(pdc_ctx* ctx);

/* Pinned host memory for zero-copy-staging of LLR batches (the caller may also pass pageable memory, at a price). */
void* pdc_host_alloc(size_t bytes);
void  pdc_host_free(void* p);
/* The same for INPUT staging only (soft bits the CPU writes and never reads back): write-combined, portable memory. */
void* pdc_host_alloc_input(size_t bytes);

/* ------------------------------------------------------------------------------------------------------------------ */
/* Batched codeblock interface = hal::hw_accelerator_pusch_dec (enqueue_operation / dequeue_operation) and the batched */
/* pusch_decoder. One batch per stream ("queue") at a time.                                                            */
/* ------------------------------------------------------------------------------------------------------------------ */

/*
 * Enqueues one batch: H2D copy of the descriptors and of n_llrs LLRs, rate dematching into the HARQ arena, LDPC decoding,
 * CB CRC, optional TB assembly + TB CRC, D2H copy of results, hard bits and TB bytes. Returns as soon as the work is
 * queued on the stream. Host buffers must stay valid until pdc_wait returns.
 *   cb_results[n_cb], cb_bits[n_cb * PDC_MAX_CB_BYTES] (may be NULL), tb_results[n_tb], tb_bytes (may be NULL).
 * llrs, cb_bits and tb_bytes in page-locked memory (pdc_host_alloc) are copied from / to directly; pageable buffers go
 * through the context's pinned staging at the price of one host memcpy.
 */
int pdc_submit(pdc_ctx*           ctx,
               uint32_t           stream,
               const pdc_cb_desc* cbs,
               uint32_t           n_cb,
               const int8_t*      llrs,
               size_t             n_llrs,
               const pdc_tb_desc* tbs,
               uint32_t           n_tb,
               pdc_cb_result*     cb_results,
               uint8_t*           cb_bits,
               pdc_tb_result*     tb_results,
               uint8_t*           tb_bytes);
/* Blocks until the batch on this stream is complete and its outputs are in the host buffers given to pdc_submit. */
int pdc_wait(pdc_ctx* ctx, uint32_t stream);

/*
 * Codeword front end of a batch = pusch_demodulator_impl's descrambling (pusch_demodulator_impl.cpp:254-259) +
 * ulsch_demultiplex_impl (lib/phy/upper/channel_processors/pusch/ulsch_demultiplex_impl.cpp:200-589), on the device:
 * H2D copy of the n_raw soft bits of all codewords, scrambling sequences, descrambling, demultiplexing into the UL-SCH
 * soft bits of the batch (which stay on the device, where the rate dematcher reads them) and into the UCI soft bits
 * (copied to uci_out). Call it before pdc_submit on the same stream and pass llrs = NULL, n_llrs = 0 to pdc_submit:
 * the codeblock descriptors then address the demultiplexed UL-SCH space. results[n_cw] and uci_out are valid after
 * pdc_wait (results also right away: the lengths are known from the plan).
 */
int pdc_submit_codewords(pdc_ctx*           ctx,
                         uint32_t           stream,
                         const pdc_cw_desc* cws,
                         uint32_t           n_cw,
                         const int8_t*      raw_llrs,
                         size_t             n_raw,
                         int8_t*            uci_out,
                         size_t             uci_capacity,
                         pdc_cw_result*     results);
/*
 * The same front end fed one step earlier, by the channel equaliser's output instead of soft bits
 * (pusch_demodulator_impl.cpp:231-259: demodulate_soft, then descrambling, per OFDM symbol): H2D copy of n_sym equalised
 * symbols ({re, im} float pairs) and their noise variances, soft demapping on the device with one demodulate_soft call
 * per OFDM symbol of each codeword (the blocks pusch_demodulator_impl hands to the demapper), then exactly what
 * pdc_submit_codewords does with the resulting soft bits - which never exist in host memory. cws[i].in_offset places
 * the codeword's soft bits in a device-internal space; sym_offsets[i] is its first symbol in symbols[] / noise_vars[]
 * (symbol k of the codeword yields soft bits k * qm .. k * qm + qm - 1 of it). Every codeword must carry
 * PDC_CW_SCRAMBLED: the demapper's output is scrambled.
 */
int pdc_submit_symbols(pdc_ctx*           ctx,
                       uint32_t           stream,
                       const pdc_cw_desc* cws,
                       uint32_t           n_cw,
                       const uint32_t*    sym_offsets,
                       const float*       symbols,
                       const float*       noise_vars,
                       size_t             n_sym,
                       int8_t*            uci_out,
                       size_t             uci_capacity,
                       pdc_cw_result*     results);
/* Non-blocking: *done = 1 when pdc_wait would not block. */
int pdc_poll(pdc_ctx* ctx, uint32_t stream, int* done);

/*
 * Device-resident variant (measurement and pipelines that already hold LLRs in HBM): same work as pdc_submit without
 * host copies. d_* are device pointers; cuda_stream is a cudaStream_t (0 = default stream); d_tb_* may be 0 when
 * n_tb == 0. The host does not read the descriptors, so the caller states what the batch contains:
 * max_lifting_size = largest Z, flags_union = OR of all pdc_cb_desc::flags, any_bg1 = some codeblock uses base graph 1.
 * flags_union may also carry PDC_LAUNCH_HIGH_RATE, a hint (results never depend on it): every codeblock is a first
 * transmission with rv 0 whose rm_length + nof_filler does not exceed 24 Z (base graph 1), i.e. only four base-graph rows are in use
 * (a 273-PRB 256QAM slot); pdc_submit works this out from the descriptors.
 */
#define PDC_LAUNCH_HIGH_RATE 0x200u
int pdc_launch_device(pdc_ctx*    ctx,
                      const void* d_cbs,
                      uint32_t    n_cb,
                      const void* d_llrs,
                      const void* d_tbs,
                      uint32_t    n_tb,
                      void*       d_cb_results,
                      void*       d_cb_bits,
                      void*       d_tb_results,
                      void*       d_tb_bytes,
                      uint32_t    max_lifting_size,
                      uint32_t    flags_union,
                      int         any_bg1,
                      void*       cuda_stream);

/* ------------------------------------------------------------------------------------------------------------------ */
/* HARQ arena = device-resident rx_buffer soft bits ("external soft bits", rx_buffer_codeblock_pool.h:63-72).          */
/* ------------------------------------------------------------------------------------------------------------------ */

int pdc_harq_read(pdc_ctx* ctx, uint32_t harq_id, int8_t* soft, uint32_t n);
int pdc_harq_write(pdc_ctx* ctx, uint32_t harq_id, const int8_t* soft, uint32_t n);
/* free_harq_context_entry: the entry may be reused; contents are left as they are (the reference pool does not clear
 * soft bits either, include/srsran/phy/upper/rx_buffer_pool.h:62-63). An entry is a PDC_MAX_CB_SOFT-long slot: codeblocks
 * of any size may follow each other on it, and what a longer codeblock left behind a shorter one is still there when a
 * third one arrives (it shows in the stale stretch of a limited-buffer transmission, exactly as in the reference). The
 * library's own bookkeeping about an entry (position of its last non-zero soft bit, for the decoder's trimming,
 * ldpc_decoder_impl.cpp:86-99) is tied to the codeblock length it was taken for; pdc_harq_write voids it. */
int pdc_harq_free(pdc_ctx* ctx, uint32_t harq_id);
/* Device address of the arena (harq_entries x PDC_MAX_CB_SOFT int8), for device-resident pipelines. */
void* pdc_harq_device_ptr(pdc_ctx* ctx);

/* ------------------------------------------------------------------------------------------------------------------ */
/* Synchronous single-codeblock calls = the "cuda" variants of ldpc_decoder / ldpc_rate_dematcher / crc_calculator.   */
/* Latency path (one codeblock cannot fill a GPU); used for configs 1-2 parity and by the factory adapters.            */
/* ------------------------------------------------------------------------------------------------------------------ */

/*
 * ldpc_decoder::decode. llrs[n_llrs] with K + 2Z <= n_llrs <= N; out receives ceil(K/8) bytes (bits MSB first).
 * crc_kind = PDC_CRC_NONE reproduces decode(..., crc = nullptr, ...). *iters = iteration count when the CRC passed,
 * 0 otherwise (the reference's empty optional).
 */
int pdc_ldpc_decode(pdc_ctx*      ctx,
                    int           base_graph,
                    int           lifting_size,
                    const int8_t* llrs,
                    uint32_t      n_llrs,
                    uint32_t      nof_filler,
                    int           crc_kind,
                    int           max_iter,
                    uint8_t*      out,
                    int*          iters);

/* ldpc_rate_dematcher::rate_dematch. buffer[N] is the in/out soft buffer in host memory, N = 66Z or 50Z. */
int pdc_rate_dematch(pdc_ctx*      ctx,
                     int8_t*       buffer,
                     uint32_t      N,
                     const int8_t* llrs,
                     uint32_t      E,
                     int           new_data,
                     int           rv,
                     int           qm,
                     uint32_t      nref,
                     uint32_t      nof_filler);

/*
 * The front end on buffers that are already on the device (descriptors on the host: the plan is made there), queued on
 * the caller's CUDA stream: d_raw_llrs -> d_sch (UL-SCH space, feed it to pdc_launch_device as d_llrs) and d_uci
 * (may be NULL when no codeword carries UCI). One call at a time per context: the plan buffers are shared with the
 * synchronous calls below.
 */
int pdc_launch_codewords_device(pdc_ctx*           ctx,
                                const pdc_cw_desc* cws,
                                uint32_t           n_cw,
                                const void*        d_raw_llrs,
                                size_t             n_raw,
                                void*              d_sch,
                                size_t             sch_capacity,
                                void*              d_uci,
                                size_t             uci_capacity,
                                pdc_cw_result*     results,
                                void*              cuda_stream);

/*
 * ulsch_demultiplex::demultiplex for n_cw codewords, synchronous, host buffers. seq_bits: the scrambling sequence of
 * the input, packed MSB first and indexed like llrs (bit in_offset + i belongs to soft bit i of the codeword), as
 * pusch_codeword_buffer::on_new_block receives it; NULL = generated on the device from c_init. sch_out receives the
 * UL-SCH soft bits at sch_offset, uci_out the UCI soft bits at uci_offset.
 */
int pdc_ulsch_demux(pdc_ctx*           ctx,
                    const pdc_cw_desc* cws,
                    uint32_t           n_cw,
                    const int8_t*      llrs,
                    size_t             n_llrs,
                    const uint8_t*     seq_bits,
                    int8_t*            sch_out,
                    size_t             sch_capacity,
                    int8_t*            uci_out,
                    size_t             uci_capacity,
                    pdc_cw_result*     results);
/*
 * demodulation_mapper::demodulate_soft, synchronous, host buffers: symbols = n {re, im} pairs, noise_vars[n],
 * llrs[n * bits per symbol]. One call = one call of the reference (block / remainder split of its SIMD build).
 */
int pdc_demodulate_soft(pdc_ctx*     ctx,
                        int8_t*      llrs,
                        const float* symbols,
                        const float* noise_vars,
                        uint32_t     n,
                        int          modulation);
/*
 * Batched, device-resident: n_calls demodulate_soft calls over d_symbols / d_noise_vars into d_llrs, queued on the
 * caller's CUDA stream (call table on the host). One call at a time per context.
 */
int pdc_launch_demod_device(pdc_ctx*              ctx,
                            const pdc_demod_call* calls,
                            uint32_t              n_calls,
                            const void*           d_symbols,
                            const void*           d_noise_vars,
                            size_t                n_sym,
                            void*                 d_llrs,
                            size_t                llr_capacity,
                            void*                 cuda_stream);

/* TS 38.211 5.2.1 pseudo-random sequence c(offset .. offset + n - 1), packed MSB first (pseudo_random_generator::generate). */
int pdc_scrambling_sequence(pdc_ctx* ctx, uint32_t c_init, uint32_t offset, uint32_t n, uint8_t* packed);

/* crc_calculator::calculate (crc_calculator.h:62-84) over the first nbits (MSB first) of packed[]; crc_kind PDC_CRC16 ..
 * PDC_CRC6: remainder of the message followed by `order` zero bits, no reflection, no final xor. */
int pdc_crc(pdc_ctx* ctx, int crc_kind, const uint8_t* packed, uint32_t nbits, uint32_t* checksum);

/* ------------------------------------------------------------------------------------------------------------------ */
/* Downlink twin: LDPC encoding + rate matching (ldpc_encoder::encode followed by ldpc_rate_matcher::rate_match, as     */
/* pdsch_encoder_impl.cpp:52-70 chains them for every codeblock), one kernel for a batch of codeblocks.                 */
/* ------------------------------------------------------------------------------------------------------------------ */

/*
 * Synchronous, host buffers: msgs = the codeblocks' message bits (packed), out = the rate-matched bits of all codeblocks,
 * one bit per byte, each codeblock at its out_offset (concatenating them gives the codeword the modulator consumes).
 */
int pdc_encode(pdc_ctx*            ctx,
               const pdc_enc_desc* cbs,
               uint32_t            n_cb,
               const uint8_t*      msgs,
               size_t              msg_bytes,
               uint8_t*            out,
               size_t              out_capacity);
/* Device-resident variant on the caller's CUDA stream (descriptors on the device too; the caller vouches for them). */
int pdc_launch_encode_device(pdc_ctx*    ctx,
                             const void* d_cbs,
                             uint32_t    n_cb,
                             const void* d_msgs,
                             void*       d_out,
                             size_t      out_capacity,
                             uint32_t    max_lifting_size,
                             int         any_bg1,
                             void*       cuda_stream);
/* ldpc_encoder::encode alone: K message bits (packed) -> the N = 66 Z or 50 Z bits of the codeblock, one bit per byte. */
int pdc_ldpc_encode(pdc_ctx* ctx, int base_graph, int lifting_size, const uint8_t* msg_packed, uint8_t* codeblock_bits);

#ifdef __cplusplus
}
#endif
#endif /* PUSCH_DEC_CUDA_H */
