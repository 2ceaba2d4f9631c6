/// \file
/// \brief "cuda" variants of the reference's plug-in interfaces, implemented on top of the C ABI
///        (include/pusch_dec_cuda.h). This translation unit sees the reference's headers and the C header only; no CUDA.
///
///   ldpc_decoder_cuda             : srsran::ldpc_decoder            (phy/upper/channel_coding/ldpc/ldpc_decoder.h:37-75)
///   ldpc_rate_dematcher_cuda      : srsran::ldpc_rate_dematcher     (.../ldpc/ldpc_rate_dematcher.h:35-56)
///   crc_calculator_cuda           : srsran::crc_calculator          (.../crc_calculator.h:62-84)
///   hw_accelerator_pusch_dec_cuda : srsran::hal::hw_accelerator_pusch_dec
///                                   (hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec.h:83-115),
///                                   driven unchanged by pusch_decoder_hw_impl (pusch_decoder_hw_impl.cpp:132-342).
///   ulsch_demultiplex_cuda        : srsran::ulsch_demultiplex + pusch_codeword_buffer
///                                   (phy/upper/channel_processors/pusch/ulsch_demultiplex.h:41-103,
///                                   pusch_codeword_buffer.h), the step that feeds pusch_decoder_buffer::on_new_softbits.
///   demodulation_mapper_cuda      : srsran::demodulation_mapper     (phy/upper/channel_modulation/demodulation_mapper.h:41-69),
///                                   created by create_channel_modulation_cuda_factory() in the place of
///                                   create_channel_modulation_sw_factory() (channel_modulation_factories.h:32-41).
///   pusch_decoder_cuda / pusch_decoder_batch_cuda : srsran::pusch_decoder (phy/upper/channel_processors/pusch/
///                                   pusch_decoder.h:54-99), the BATCHED decoder: every decoder created from one batch
///                                   object queues its transport block at on_end_softbits; flush() decodes all of them
///                                   (every UE and cell of the slot) with ONE pdc_submit, TB concatenation and TB CRC on
///                                   the device, and then fires the notifiers.
///   ldpc_encoder_cuda             : srsran::ldpc_encoder (phy/upper/channel_coding/ldpc/ldpc_encoder.h:37-38), the downlink
///                                   twin; pdc_encode additionally fuses the rate matcher for whole batches.
/// and the factories a "cuda" branch of create_ldpc_decoder_factory_sw / create_ldpc_rate_dematcher_factory_sw /
/// create_crc_calculator_factory_sw / create_hw_accelerator_pusch_dec_factory returns.
#pragma once

#include "pusch_dec_cuda.h"
#include "srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec.h"
#include "srsran/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc.h"
#include "srsran/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc_factory.h"
#include "srsran/hal/phy/upper/channel_processors/pusch/hw_accelerator_pusch_dec_factory.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/phy/upper/channel_modulation/channel_modulation_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_codeword_buffer.h"
#include "srsran/phy/upper/channel_coding/ldpc/ldpc_encoder.h"
#include "srsran/phy/upper/channel_coding/ldpc/ldpc_segmenter_rx.h"
#include "srsran/phy/upper/channel_processors/pusch/factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/ulsch_demultiplex.h"
#include <atomic>
#include <memory>
#include <mutex>
#include <optional>
#include <vector>

namespace srsran {
namespace cuda {

/// Shared, reference-counted GPU context (one per device; owns the HARQ arena, streams and pinned staging).
class context
{
public:
  /// Configuration of the device context.
  struct config {
    int      device       = 0;
    unsigned max_cbs      = 4096;
    unsigned harq_entries = 4096;
    /// Which reference decoder is reproduced bit for bit: PDC_SCALE_X86 (avx2/avx512, "auto" on x86) or generic.
    int scale_mode = PDC_SCALE_X86;
    /// Batch queues (CUDA streams with their staging buffers): one per hal::hw_accelerator_pusch_dec /
    /// pusch_decoder_batch_cuda instance, i.e. per concurrent PUSCH processor.
    unsigned nof_queues = 4;
  };

  /// Returns nullptr if no usable GPU is present (there is no software fallback).
  static std::shared_ptr<context> create(const config& cfg);
  ~context();
  pdc_ctx* get() const { return ctx; }
  unsigned nof_queues() const { return queues; }
  /// Hands out an unused batch queue, or no_queue when all are taken: queues are never shared between objects, and an
  /// object gives its queue back when it is destroyed.
  static constexpr unsigned no_queue = ~0U;
  unsigned                  claim_queue()
  {
    std::lock_guard<std::mutex> lock(queue_mutex);
    for (unsigned q = 0; q != queues; ++q) {
      if (!(in_use & (1ULL << q))) {
        in_use |= 1ULL << q;
        return q;
      }
    }
    return no_queue;
  }
  /// Marks a specific queue as taken (objects constructed with an explicit queue number); false if it already is.
  bool claim_queue(unsigned q)
  {
    std::lock_guard<std::mutex> lock(queue_mutex);
    if (q >= queues || (in_use & (1ULL << q))) {
      return false;
    }
    in_use |= 1ULL << q;
    return true;
  }
  void release_queue(unsigned q)
  {
    std::lock_guard<std::mutex> lock(queue_mutex);
    if (q < queues) {
      in_use &= ~(1ULL << q);
    }
  }

private:
  context(pdc_ctx* c, unsigned nq) : ctx(c), queues(nq > 64 ? 64 : nq) {}
  pdc_ctx*   ctx;
  unsigned   queues;
  std::mutex queue_mutex;
  uint64_t   in_use = 0;
};

/// LDPC decoder, single-codeblock synchronous call (latency path).
class ldpc_decoder_cuda : public ldpc_decoder
{
public:
  explicit ldpc_decoder_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::optional<unsigned>
  decode(bit_buffer& output, span<const log_likelihood_ratio> input, crc_calculator* crc, const configuration& cfg) override;

private:
  std::shared_ptr<context> ctx;
};

/// LDPC rate dematcher, single-codeblock synchronous call.
class ldpc_rate_dematcher_cuda : public ldpc_rate_dematcher
{
public:
  explicit ldpc_rate_dematcher_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  void rate_dematch(span<log_likelihood_ratio>       output,
                    span<const log_likelihood_ratio> input,
                    bool                             new_data,
                    const codeblock_metadata&        cfg) override;

private:
  std::shared_ptr<context> ctx;
};

/// CRC calculator (CRC16 / CRC24A / CRC24B).
class crc_calculator_cuda : public crc_calculator
{
public:
  crc_calculator_cuda(std::shared_ptr<context> c, crc_generator_poly p) : ctx(std::move(c)), poly(p) {}
  crc_calculator_checksum_t calculate_byte(span<const uint8_t> data) override;
  crc_calculator_checksum_t calculate_bit(span<const uint8_t> data) override;
  crc_calculator_checksum_t calculate(const bit_buffer& data) override;
  crc_generator_poly        get_generator_poly() const override { return poly; }

private:
  std::shared_ptr<context> ctx;
  crc_generator_poly       poly;
};

/// \brief PUSCH decoder accelerator in the hal::hw_accelerator_pusch_dec slot.
///
/// External HARQ: soft bits live in the device arena, indexed by the absolute codeblock id of the rx_buffer pool
/// (rx_buffer.h:58-65), so pusch_decoder_hw_impl enqueues every codeblock of the transport block before the first
/// dequeue (pusch_decoder_hw_impl.cpp:246-262). The first dequeue submits all of them as ONE GPU batch.
class hw_accelerator_pusch_dec_cuda : public hal::hw_accelerator_pusch_dec
{
public:
  hw_accelerator_pusch_dec_cuda(std::shared_ptr<context> c, unsigned queue);
  ~hw_accelerator_pusch_dec_cuda() override;

  void reserve_queue() override;
  void free_queue() override;
  bool enqueue_operation(span<const int8_t> data, span<const int8_t> soft_data = {}, unsigned cb_index = 0) override;
  bool dequeue_operation(span<uint8_t> data, span<int8_t> soft_data = {}, unsigned segment_index = 0) override;
  void configure_operation(const hal::hw_pusch_decoder_configuration& config, unsigned cb_index = 0) override;
  void read_operation_outputs(hal::hw_pusch_decoder_outputs& out,
                              unsigned                       cb_index       = 0,
                              unsigned                       absolute_cb_id = 0) override;
  void free_harq_context_entry(unsigned absolute_cb_id) override;
  bool is_external_harq_supported() const override { return true; }

private:
  void flush();

  std::shared_ptr<context>   ctx;
  unsigned                   queue;
  std::vector<pdc_cb_desc>   pending_cfg;   // by codeblock index within the TB
  std::vector<int>           slot_of_cb;    // position in the submitted batch, -1 if not enqueued
  std::vector<pdc_cb_desc>   batch;
  std::vector<pdc_cb_result> results;
  std::vector<uint8_t>       bits;
  int8_t*                    llr_staging = nullptr; // pinned
  size_t                     llr_capacity = 0, llr_used = 0;
  bool                       submitted = false;
};

/// \brief UL-SCH demultiplexer in the ulsch_demultiplex slot (created per PUSCH processor, like ulsch_demultiplex_impl).
///
/// The demodulator writes the descrambled soft bits of the whole codeword into the staging buffer this object hands out
/// (get_next_block_view never limits a block to the current OFDM symbol) and passes the scrambling sequence of every
/// block; on_end_codeword runs the demultiplexing of the whole codeword on the GPU (pdc_ulsch_demux) and feeds the four
/// decoder buffers. CSI Part 2: the CSI Part 1 buffer is completed first; if its decoder answers with set_csi_part2 the
/// codeword is demultiplexed a second time with the CSI Part 2 sizes (HARQ-ACK and CSI Part 1 do not depend on them).
class ulsch_demultiplex_cuda : public ulsch_demultiplex, private pusch_codeword_buffer
{
public:
  explicit ulsch_demultiplex_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  void set_csi_part2(pusch_decoder_buffer& csi_part2, unsigned nof_csi_part2_bits, unsigned nof_csi_part2_enc_bits) override;
  pusch_codeword_buffer& demultiplex(pusch_decoder_buffer& sch_data,
                                     pusch_decoder_buffer& harq_ack,
                                     pusch_decoder_buffer& csi_part1,
                                     const configuration&  config) override;

private:
  span<log_likelihood_ratio> get_next_block_view(unsigned block_size) override;
  void on_new_block(span<const log_likelihood_ratio> data, const bit_buffer& scrambling_seq) override;
  void on_end_codeword() override;
  bool run(pdc_cw_result& res);

  std::shared_ptr<context>          ctx;
  configuration                     cfg;
  pusch_decoder_buffer*             sch = nullptr;
  pusch_decoder_buffer*             ack = nullptr;
  pusch_decoder_buffer*             csi1 = nullptr;
  pusch_decoder_buffer*             csi2 = nullptr;
  unsigned                          csi2_bits = 0, csi2_enc = 0;
  std::vector<log_likelihood_ratio> codeword; // descrambled soft bits, resource-element order
  std::vector<uint8_t>              seq;      // scrambling sequence, packed MSB first
  std::vector<int8_t>               out_sch, out_uci;
  size_t                            count = 0;
};

/// \brief Soft demapper in the demodulation_mapper slot: one demodulate_soft call = one pdc_demodulate_soft call (bit-exact
/// with the x86 build of demodulation_mapper_impl, SIMD blocks and scalar remainder of the call included).
class demodulation_mapper_cuda : public demodulation_mapper
{
public:
  explicit demodulation_mapper_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  void demodulate_soft(span<log_likelihood_ratio> llrs,
                       span<const cf_t>           symbols,
                       span<const float>          noise_vars,
                       modulation_scheme          mod) override;

private:
  std::shared_ptr<context> ctx;
};

/// Channel-modulation factory whose demodulation mapper runs on the GPU; the modulation mapper and the EVM calculator are
/// the software ones (they are not on the uplink decode path).
std::shared_ptr<channel_modulation_factory> create_channel_modulation_cuda_factory(std::shared_ptr<context> ctx);

class pusch_decoder_cuda;

/// \brief Slot batch of the batched PUSCH decoder (the throughput path of INTEGRATION.md 3).
///
/// Owns one accelerator queue. pusch_decoder_cuda instances created from it behave like pusch_decoder_impl up to
/// on_end_softbits, which queues the transport block instead of decoding it; flush() - called once the PUSCH processors
/// of the slot have delivered their soft bits - builds the codeblock and transport-block descriptors with the
/// reference's own ldpc_segmenter_rx, submits them as one batch and reports every transport block through its
/// notifier exactly as pusch_decoder_impl::join_and_notify does (pusch_decoder_impl.cpp:384-450): codeblock CRC flags
/// in the rx buffer, iteration statistics, TB bytes, TB CRC, buffer release / unlock. Deferring on_sch_data is what the
/// interface allows (the software decoder also completes on another thread, pusch_decoder_impl.cpp:365-367).
class pusch_decoder_batch_cuda
{
public:
  /// \param queue Queue ("stream") of the context this batch submits on.
  pusch_decoder_batch_cuda(std::shared_ptr<context> ctx, unsigned queue, std::unique_ptr<ldpc_segmenter_rx> segmenter);
  ~pusch_decoder_batch_cuda();
  /// pusch_decoder_factory::create(): a decoder bound to this batch (one per PUSCH processor).
  std::unique_ptr<pusch_decoder> create();
  /// Number of transport blocks waiting for flush().
  unsigned pending() const { return static_cast<unsigned>(queued.size()); }
  /// Decodes everything queued with one GPU submission and notifies. Returns false if the submission failed (every
  /// queued transport block is then reported with tb_crc_ok = false).
  bool flush();

private:
  friend class pusch_decoder_cuda;
  struct queued_tb {
    pusch_decoder_cuda*               decoder;
    span<uint8_t>                     transport_block;
    unique_rx_buffer                  rm_buffer;
    pusch_decoder_notifier*           notifier;
    pusch_decoder::configuration      cfg;
    std::vector<log_likelihood_ratio> llrs;
  };
  void queue(queued_tb&& tb) { queued.push_back(std::move(tb)); }

  std::shared_ptr<context>           ctx;
  unsigned                           queue_id;
  std::unique_ptr<ldpc_segmenter_rx> segmenter;
  std::vector<queued_tb>             queued;
  std::vector<pdc_cb_desc>           cbs;
  std::vector<pdc_tb_desc>           tbs;
  std::vector<pdc_cb_result>         cb_results;
  std::vector<pdc_tb_result>         tb_results;
  std::vector<uint8_t>               decode_mask;
  int8_t*                            llr_staging  = nullptr; // page-locked
  size_t                             llr_capacity = 0;
  uint8_t*                           tb_staging   = nullptr; // page-locked
  size_t                             tb_capacity  = 0;
};

/// One PUSCH decoder bound to a pusch_decoder_batch_cuda; the pusch_decoder_buffer it hands out is itself.
class pusch_decoder_cuda : public pusch_decoder, private pusch_decoder_buffer
{
public:
  explicit pusch_decoder_cuda(pusch_decoder_batch_cuda& b) : batch(b) {}
  pusch_decoder_buffer& new_data(span<uint8_t>           transport_block,
                                 unique_rx_buffer        rm_buffer,
                                 pusch_decoder_notifier& notifier,
                                 const configuration&    cfg) override;
  void                  set_nof_softbits(units::bits nof_softbits) override;

private:
  friend class pusch_decoder_batch_cuda;
  span<log_likelihood_ratio> get_next_block_view(unsigned block_size) override;
  void                       on_new_softbits(span<const log_likelihood_ratio> softbits) override;
  void                       on_end_softbits() override;

  enum class state { idle, collecting, decoding };
  pusch_decoder_batch_cuda&          batch;
  state                              st = state::idle;
  pusch_decoder_batch_cuda::queued_tb current;
  std::vector<log_likelihood_ratio>  view;
  std::optional<unsigned>            expected_softbits;
};

/// What a "cuda" choice in place of create_pusch_decoder_factory_sw / _hw (phy/upper/channel_processors/pusch/
/// factories.h:56-99) returns: every create() yields a decoder bound to the given slot batch.
std::shared_ptr<pusch_decoder_factory> create_pusch_decoder_factory_cuda(std::shared_ptr<pusch_decoder_batch_cuda> batch);

/// LDPC encoder, single-codeblock synchronous call (the throughput interface is pdc_encode: encoder + rate matcher for a
/// batch of codeblocks in one kernel).
class ldpc_encoder_cuda : public ldpc_encoder
{
public:
  explicit ldpc_encoder_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  void encode(bit_buffer& output, const bit_buffer& input, const codeblock_metadata::tb_common_metadata& cfg) override;

private:
  std::shared_ptr<context> ctx;
  std::vector<uint8_t>     bits;
};
/// \brief PDSCH encoder accelerator in the hal::hw_accelerator_pdsch_enc slot (hal/phy/upper/channel_processors/
/// hw_accelerator_pdsch_enc.h:77-105), codeblock mode: pdsch_encoder_hw_impl (lib/phy/upper/channel_processors/
/// pdsch_encoder_hw_impl.cpp:35-200), unchanged, segments the transport block and attaches the CRCs, enqueues every
/// segment and dequeues the rate-matched codeblocks; the first dequeue encodes and rate-matches all enqueued segments as
/// ONE pdc_encode batch.
class hw_accelerator_pdsch_enc_cuda : public hal::hw_accelerator_pdsch_enc
{
public:
  explicit hw_accelerator_pdsch_enc_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  void reserve_queue() override;
  void free_queue() override {}
  bool enqueue_operation(span<const uint8_t> data, span<const uint8_t> aux_data = {}, unsigned cb_index = 0) override;
  bool dequeue_operation(span<uint8_t> data, span<uint8_t> aux_data = {}, unsigned segment_index = 0) override;
  void configure_operation(const hal::hw_pdsch_encoder_configuration& config, unsigned cb_index = 0) override;
  bool get_cb_mode() const override { return true; }
  unsigned get_max_tb_size() const override { return 0; }

private:
  std::shared_ptr<context>  ctx;
  std::vector<pdc_enc_desc> descs;    // by codeblock index
  std::vector<uint8_t>      enqueued; // by codeblock index
  std::vector<uint8_t>      msgs, out;
  bool                      encoded = false;
};
std::shared_ptr<hal::hw_accelerator_pdsch_enc_factory> create_hw_accelerator_pdsch_enc_factory_cuda(std::shared_ptr<context> ctx);

/// What a "cuda" branch of create_ldpc_encoder_factory_sw (channel_coding_factories.h:68) returns.
std::shared_ptr<ldpc_encoder_factory> create_ldpc_encoder_factory_cuda(std::shared_ptr<context> ctx);

std::shared_ptr<ldpc_decoder_factory>        create_ldpc_decoder_factory_cuda(std::shared_ptr<context> ctx);
std::shared_ptr<ldpc_rate_dematcher_factory> create_ldpc_rate_dematcher_factory_cuda(std::shared_ptr<context> ctx);
std::shared_ptr<crc_calculator_factory>      create_crc_calculator_factory_cuda(std::shared_ptr<context> ctx);
std::shared_ptr<hal::hw_accelerator_pusch_dec_factory>
create_hw_accelerator_pusch_dec_factory_cuda(std::shared_ptr<context> ctx);
/// What a "cuda" branch of create_ulsch_demultiplex_factory_sw (phy/upper/channel_processors/pusch/factories.h:161)
/// creates for every PUSCH processor.
std::unique_ptr<ulsch_demultiplex> create_ulsch_demultiplex_cuda(std::shared_ptr<context> ctx);

} // namespace cuda
} // namespace srsran
