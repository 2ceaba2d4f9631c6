#include "pusch_dec_cuda_adapters.h"
#include "srsran/ran/dmrs.h"
#include "srsran/ran/pusch/pusch_constants.h"
#include "srsran/support/srsran_assert.h"
#include <algorithm>
#include "srsran/ran/sch/modulation_scheme.h"
#include "srsran/support/error_handling.h"
#include "srsran/support/srsran_assert.h"
#include <atomic>
#include <cstring>

using namespace srsran;
using namespace srsran::cuda;

static int to_crc_kind(crc_generator_poly poly)
{
  switch (poly) {
    case crc_generator_poly::CRC16:
      return PDC_CRC16;
    case crc_generator_poly::CRC24A:
      return PDC_CRC24A;
    case crc_generator_poly::CRC24B:
      return PDC_CRC24B;
    case crc_generator_poly::CRC24C:
      return PDC_CRC24C;
    case crc_generator_poly::CRC11:
      return PDC_CRC11;
    case crc_generator_poly::CRC6:
      return PDC_CRC6;
    default:
      return -1;
  }
}

namespace {
/// Errors of the synchronous device calls are logged (rate limited), never fatal.
void log_device_error(const char* what)
{
  static std::atomic<unsigned> count{0};
  unsigned                     n = count.fetch_add(1, std::memory_order_relaxed);
  if (n < 16 || (n & (n - 1)) == 0) {
    fmt::print(stderr, "[pusch_dec_cuda] {} failed ({} so far): {}\n", what, n + 1, pdc_last_error());
  }
}
} // namespace

std::shared_ptr<context> context::create(const config& cfg)
{
  pdc_config c;
  pdc_default_config(&c);
  c.device       = cfg.device;
  c.max_cbs      = cfg.max_cbs;
  c.max_llrs     = cfg.max_cbs * 12288u;
  c.harq_entries = cfg.harq_entries;
  c.scale_mode   = cfg.scale_mode;
  c.nof_streams  = std::max(1U, cfg.nof_queues);
  pdc_ctx* h     = nullptr;
  if (pdc_create(&c, &h) != PDC_OK) {
    return nullptr;
  }
  return std::shared_ptr<context>(new context(h, c.nof_streams));
}

context::~context()
{
  pdc_destroy(ctx);
}

std::optional<unsigned> ldpc_decoder_cuda::decode(bit_buffer&                      output,
                                                  span<const log_likelihood_ratio> input,
                                                  crc_calculator*                  crc,
                                                  const configuration&             cfg)
{
  const auto& tb = cfg.block_conf.tb_common;
  int         bg = (tb.base_graph == ldpc_base_graph_type::BG1) ? 1 : 2;
  unsigned    Z  = static_cast<unsigned>(tb.lifting_size);
  unsigned    K  = ((bg == 1) ? 22 : 10) * Z;
  // Same contract as ldpc_decoder_impl::decode (ldpc_decoder_impl.cpp:69-83).
  srsran_assert(output.size() == K, "The output size {} is not equal to the message length {}.", output.size(), K);
  int kind = PDC_CRC_NONE;
  if (crc != nullptr) {
    kind = to_crc_kind(crc->get_generator_poly());
    srsran_assert(kind > 0, "Invalid CRC calculator.");
  }
  int iters = 0;
  int rc    = pdc_ldpc_decode(ctx->get(),
                           bg,
                           static_cast<int>(Z),
                           reinterpret_cast<const int8_t*>(input.data()),
                           input.size(),
                           cfg.block_conf.cb_specific.nof_filler_bits,
                           kind,
                           static_cast<int>(cfg.algorithm_conf.max_iterations),
                           output.get_buffer().data(),
                           &iters);
  if (rc != PDC_OK) {
    // A failed device call is reported like a dropped accelerator operation - CRC failure, i.e. no iteration count
    // (hw_accelerator_pusch_dec_acc100_impl.cpp:176-187, 233-247) - and logged; the gNB keeps running and HARQ recovers.
    log_device_error("pdc_ldpc_decode");
    return std::nullopt;
  }
  if (iters > 0) {
    return static_cast<unsigned>(iters);
  }
  return std::nullopt;
}

void ldpc_rate_dematcher_cuda::rate_dematch(span<log_likelihood_ratio>       output,
                                            span<const log_likelihood_ratio> input,
                                            bool                             new_data,
                                            const codeblock_metadata&        cfg)
{
  int rc = pdc_rate_dematch(ctx->get(),
                            reinterpret_cast<int8_t*>(output.data()),
                            output.size(),
                            reinterpret_cast<const int8_t*>(input.data()),
                            input.size(),
                            new_data,
                            static_cast<int>(cfg.tb_common.rv),
                            static_cast<int>(get_bits_per_symbol(cfg.tb_common.mod)),
                            cfg.tb_common.Nref,
                            cfg.cb_specific.nof_filler_bits);
  if (rc != PDC_OK) {
    // The soft buffer keeps what it held: the codeblock fails its CRC and is retransmitted.
    log_device_error("pdc_rate_dematch");
  }
}

crc_calculator_checksum_t crc_calculator_cuda::calculate_byte(span<const uint8_t> data)
{
  uint32_t c  = 0;
  int      rc = pdc_crc(ctx->get(), to_crc_kind(poly), data.data(), data.size() * 8, &c);
  if (rc != PDC_OK) {
    log_device_error("pdc_crc");
    return ~0U; // never the remainder of a message that carries its checksum: the check fails
  }
  return c;
}

crc_calculator_checksum_t crc_calculator_cuda::calculate_bit(span<const uint8_t> data)
{
  std::vector<uint8_t> packed((data.size() + 7) / 8, 0);
  for (size_t i = 0; i != data.size(); ++i) {
    packed[i / 8] |= static_cast<uint8_t>((data[i] & 1U) << (7 - (i % 8)));
  }
  uint32_t c  = 0;
  int      rc = pdc_crc(ctx->get(), to_crc_kind(poly), packed.data(), data.size(), &c);
  if (rc != PDC_OK) {
    log_device_error("pdc_crc");
    return ~0U; // never the remainder of a message that carries its checksum: the check fails
  }
  return c;
}

crc_calculator_checksum_t crc_calculator_cuda::calculate(const bit_buffer& data)
{
  uint32_t c  = 0;
  int      rc = pdc_crc(ctx->get(), to_crc_kind(poly), data.get_buffer().data(), data.size(), &c);
  if (rc != PDC_OK) {
    log_device_error("pdc_crc");
    return ~0U; // never the remainder of a message that carries its checksum: the check fails
  }
  return c;
}

hw_accelerator_pusch_dec_cuda::hw_accelerator_pusch_dec_cuda(std::shared_ptr<context> c, unsigned q) :
  ctx(std::move(c)), queue(q), pending_cfg(MAX_NOF_SEGMENTS), slot_of_cb(MAX_NOF_SEGMENTS, -1)
{
  llr_capacity = static_cast<size_t>(MAX_NOF_SEGMENTS) * 12288u;
  llr_staging  = static_cast<int8_t*>(pdc_host_alloc(llr_capacity));
  report_fatal_error_if_not(llr_staging != nullptr, "pdc_host_alloc failed");
  results.resize(MAX_NOF_SEGMENTS);
  bits.resize(static_cast<size_t>(MAX_NOF_SEGMENTS) * PDC_MAX_CB_BYTES);
}

hw_accelerator_pusch_dec_cuda::~hw_accelerator_pusch_dec_cuda()
{
  ctx->release_queue(queue);
  pdc_host_free(llr_staging);
}

void hw_accelerator_pusch_dec_cuda::reserve_queue()
{
  batch.clear();
  std::fill(slot_of_cb.begin(), slot_of_cb.end(), -1);
  llr_used  = 0;
  submitted = false;
}

void hw_accelerator_pusch_dec_cuda::free_queue()
{
  batch.clear();
  submitted = false;
}

void hw_accelerator_pusch_dec_cuda::configure_operation(const hal::hw_pusch_decoder_configuration& cfg,
                                                        unsigned                                   cb_index)
{
  pdc_cb_desc d;
  std::memset(&d, 0, sizeof(d));
  d.rm_length    = cfg.cw_length;
  d.harq_id      = cfg.absolute_cb_id;
  d.nref         = cfg.Nref;
  d.lifting_size = static_cast<uint16_t>(cfg.lifting_size);
  d.nof_filler   = static_cast<uint16_t>(cfg.nof_filler_bits);
  d.base_graph   = (cfg.base_graph_index == ldpc_base_graph_type::BG1) ? 1 : 2;
  d.qm           = static_cast<uint8_t>(get_bits_per_symbol(cfg.modulation));
  d.rv           = static_cast<uint8_t>(cfg.rv);
  d.crc_kind     = (cfg.cb_crc_type == hal::hw_dec_cb_crc_type::CRC16)    ? PDC_CRC16
                   : (cfg.cb_crc_type == hal::hw_dec_cb_crc_type::CRC24A) ? PDC_CRC24A
                                                                          : PDC_CRC24B;
  d.max_iter     = static_cast<uint8_t>(cfg.max_nof_ldpc_iterations);
  d.flags        = PDC_CB_DEMATCH | PDC_CB_DECODE | (cfg.new_data ? PDC_CB_NEW_DATA : 0) |
            (cfg.use_early_stop ? PDC_CB_EARLY_STOP : 0);
  d.tb_index            = 0xffff;
  pending_cfg[cb_index] = d;
}

bool hw_accelerator_pusch_dec_cuda::enqueue_operation(span<const int8_t> data,
                                                      span<const int8_t> /*soft_data*/,
                                                      unsigned cb_index)
{
  // "Queue full": the caller retries after dequeuing (pusch_decoder_hw_impl.cpp:237-241).
  if (submitted || llr_used + data.size() > llr_capacity || batch.size() >= MAX_NOF_SEGMENTS) {
    return false;
  }
  pdc_cb_desc d = pending_cfg[cb_index];
  d.llr_offset  = static_cast<uint32_t>(llr_used);
  std::memcpy(llr_staging + llr_used, data.data(), data.size());
  llr_used += (data.size() + 15) & ~static_cast<size_t>(15);
  slot_of_cb[cb_index] = static_cast<int>(batch.size());
  batch.push_back(d);
  return true;
}

void hw_accelerator_pusch_dec_cuda::flush()
{
  int rc = pdc_submit(ctx->get(), queue, batch.data(), batch.size(), llr_staging, llr_used, nullptr, 0, results.data(),
                      bits.data(), nullptr, nullptr);
  if (rc == PDC_OK) {
    rc = pdc_wait(ctx->get(), queue);
  }
  if (rc != PDC_OK) {
    // A failed batch reports CRC failure with the maximum number of iterations, like a dropped accelerator operation
    // (hw_accelerator_pusch_dec_acc100_impl.cpp:233-247).
    for (size_t i = 0; i != batch.size(); ++i) {
      results[i].crc_ok = 0;
      results[i].iters  = batch[i].max_iter;
    }
  }
  submitted = true;
}

bool hw_accelerator_pusch_dec_cuda::dequeue_operation(span<uint8_t> data,
                                                      span<int8_t> /*soft_data*/,
                                                      unsigned segment_index)
{
  int slot = slot_of_cb[segment_index];
  if (slot < 0) {
    return false;
  }
  if (!submitted) {
    flush();
  }
  std::memcpy(data.data(), bits.data() + static_cast<size_t>(slot) * PDC_MAX_CB_BYTES, data.size());
  return true;
}

void hw_accelerator_pusch_dec_cuda::read_operation_outputs(hal::hw_pusch_decoder_outputs& out,
                                                           unsigned                       cb_index,
                                                           unsigned /*absolute_cb_id*/)
{
  int slot = slot_of_cb[cb_index];
  srsran_assert(slot >= 0 && submitted, "No completed operation for this codeblock.");
  out.CRC_pass            = results[slot].crc_ok != 0;
  out.nof_ldpc_iterations = results[slot].iters;
}

void hw_accelerator_pusch_dec_cuda::free_harq_context_entry(unsigned absolute_cb_id)
{
  pdc_harq_free(ctx->get(), absolute_cb_id);
}

namespace {

class ldpc_decoder_factory_cuda : public ldpc_decoder_factory
{
public:
  explicit ldpc_decoder_factory_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::unique_ptr<ldpc_decoder> create() override { return std::make_unique<ldpc_decoder_cuda>(ctx); }

private:
  std::shared_ptr<context> ctx;
};

class ldpc_rate_dematcher_factory_cuda : public ldpc_rate_dematcher_factory
{
public:
  explicit ldpc_rate_dematcher_factory_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::unique_ptr<ldpc_rate_dematcher> create() override { return std::make_unique<ldpc_rate_dematcher_cuda>(ctx); }

private:
  std::shared_ptr<context> ctx;
};

class crc_calculator_factory_cuda : public crc_calculator_factory
{
public:
  explicit crc_calculator_factory_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::unique_ptr<crc_calculator> create(crc_generator_poly poly) override
  {
    if (to_crc_kind(poly) < 0) {
      return nullptr;
    }
    return std::make_unique<crc_calculator_cuda>(ctx, poly);
  }

private:
  std::shared_ptr<context> ctx;
};

class hw_accelerator_pusch_dec_factory_cuda : public hal::hw_accelerator_pusch_dec_factory
{
public:
  explicit hw_accelerator_pusch_dec_factory_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::unique_ptr<hal::hw_accelerator_pusch_dec> create() override
  {
    // One batch queue ("hardware queue") per accelerator instance, never shared: two PUSCH processors driving one
    // queue from two threads would find it busy. Beyond the context's queues creation fails, like a hardware
    // accelerator that ran out of queues (the caller sizes context::config::nof_queues to its PUSCH processors).
    unsigned q = ctx->claim_queue();
    if (q == context::no_queue) {
      return nullptr;
    }
    return std::make_unique<hw_accelerator_pusch_dec_cuda>(ctx, q);
  }

private:
  std::shared_ptr<context> ctx;
};

} // namespace

std::shared_ptr<ldpc_decoder_factory> srsran::cuda::create_ldpc_decoder_factory_cuda(std::shared_ptr<context> ctx)
{
  return ctx ? std::make_shared<ldpc_decoder_factory_cuda>(std::move(ctx)) : nullptr;
}

std::shared_ptr<ldpc_rate_dematcher_factory>
srsran::cuda::create_ldpc_rate_dematcher_factory_cuda(std::shared_ptr<context> ctx)
{
  return ctx ? std::make_shared<ldpc_rate_dematcher_factory_cuda>(std::move(ctx)) : nullptr;
}

std::shared_ptr<crc_calculator_factory> srsran::cuda::create_crc_calculator_factory_cuda(std::shared_ptr<context> ctx)
{
  return ctx ? std::make_shared<crc_calculator_factory_cuda>(std::move(ctx)) : nullptr;
}

std::shared_ptr<hal::hw_accelerator_pusch_dec_factory>
srsran::cuda::create_hw_accelerator_pusch_dec_factory_cuda(std::shared_ptr<context> ctx)
{
  return ctx ? std::make_shared<hw_accelerator_pusch_dec_factory_cuda>(std::move(ctx)) : nullptr;
}

// ---- ulsch_demultiplex_cuda ---------------------------------------------------------------------------------------------

pusch_codeword_buffer& ulsch_demultiplex_cuda::demultiplex(pusch_decoder_buffer& sch_data,
                                                           pusch_decoder_buffer& harq_ack,
                                                           pusch_decoder_buffer& csi_part1,
                                                           const configuration&  config)
{
  cfg       = config;
  sch       = &sch_data;
  ack       = (config.nof_harq_ack_bits != 0) ? &harq_ack : nullptr;
  csi1      = (config.nof_csi_part1_bits != 0) ? &csi_part1 : nullptr;
  csi2      = nullptr;
  csi2_bits = csi2_enc = 0;
  count     = 0;
  // Size of the codeword: every OFDM symbol of the allocation (ulsch_demultiplex_impl.cpp:371-382).
  const unsigned bpre    = get_bits_per_symbol(cfg.modulation) * cfg.nof_layers;
  const unsigned re_dmrs = (NRE - cfg.nof_cdm_groups_without_data * (cfg.dmrs == dmrs_type::TYPE1 ? 6 : 4)) * cfg.nof_prb;
  size_t         total   = 0;
  for (unsigned l = cfg.start_symbol_index; l != cfg.start_symbol_index + cfg.nof_symbols; ++l) {
    total += (cfg.dmrs_symbol_mask.test(l) ? re_dmrs : NRE * cfg.nof_prb) * bpre;
  }
  codeword.resize(total);
  seq.assign((total + 7) / 8, 0);
  return *this;
}

void ulsch_demultiplex_cuda::set_csi_part2(pusch_decoder_buffer& csi_part2, unsigned nof_bits, unsigned nof_enc_bits)
{
  csi2      = &csi_part2;
  csi2_bits = nof_bits;
  csi2_enc  = nof_enc_bits;
}

span<log_likelihood_ratio> ulsch_demultiplex_cuda::get_next_block_view(unsigned block_size)
{
  block_size = static_cast<unsigned>(std::min<size_t>(block_size, codeword.size() - count));
  return span<log_likelihood_ratio>(codeword).subspan(count, block_size);
}

void ulsch_demultiplex_cuda::on_new_block(span<const log_likelihood_ratio> data, const bit_buffer& scrambling_seq)
{
  srsran_assert(count + data.size() <= codeword.size(), "More soft bits than the allocation holds.");
  if (data.data() != codeword.data() + count) {
    std::copy(data.begin(), data.end(), codeword.begin() + count);
  }
  // Append the scrambling sequence at bit position `count` (MSB first).
  for (size_t i = 0, n = data.size(); i != n;) {
    const size_t   pos  = count + i;
    const unsigned room = 8 - static_cast<unsigned>(pos & 7);
    const unsigned take = static_cast<unsigned>(std::min<size_t>(room, n - i));
    const unsigned bits = scrambling_seq.extract(i, take); // first bit in the most significant position
    seq[pos >> 3] |= static_cast<uint8_t>(bits << (room - take));
    i += take;
  }
  count += data.size();
}

bool ulsch_demultiplex_cuda::run(pdc_cw_result& res)
{
  pdc_cw_desc d                 = {};
  d.qm                          = get_bits_per_symbol(cfg.modulation);
  d.nof_layers                  = cfg.nof_layers;
  d.start_symbol_index          = cfg.start_symbol_index;
  d.nof_symbols                 = cfg.nof_symbols;
  d.dmrs_type                   = (cfg.dmrs == dmrs_type::TYPE1) ? 1 : 2;
  d.nof_cdm_groups_without_data = cfg.nof_cdm_groups_without_data;
  d.nof_prb                     = cfg.nof_prb;
  for (unsigned l = 0; l != 14; ++l) {
    if (l < cfg.dmrs_symbol_mask.size() && cfg.dmrs_symbol_mask.test(l)) {
      d.dmrs_symbol_mask |= static_cast<uint16_t>(1u << l);
    }
  }
  d.nof_harq_ack_rvd       = cfg.nof_harq_ack_rvd;
  d.nof_harq_ack_bits      = cfg.nof_harq_ack_bits;
  d.nof_enc_harq_ack_bits  = cfg.nof_enc_harq_ack_bits;
  d.nof_csi_part1_bits     = cfg.nof_csi_part1_bits;
  d.nof_enc_csi_part1_bits = cfg.nof_enc_csi_part1_bits;
  d.nof_csi_part2_bits     = csi2_bits;
  d.nof_enc_csi_part2_bits = csi2_enc;
  out_sch.resize(codeword.size() + 16);
  out_uci.resize(codeword.size() + 16);
  return pdc_ulsch_demux(ctx->get(), &d, 1, reinterpret_cast<const int8_t*>(codeword.data()), codeword.size(),
                         seq.data(), out_sch.data(), out_sch.size(), out_uci.data(), out_uci.size(), &res) == PDC_OK;
}

void ulsch_demultiplex_cuda::on_end_codeword()
{
  auto deliver = [](pusch_decoder_buffer* b, const int8_t* p, size_t n) {
    if (b != nullptr && n != 0) {
      b->on_new_softbits(span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(p), n));
    }
  };
  pdc_cw_result res = {};
  bool          ok  = run(res);
  srsran_assert(ok, "UL-SCH demultiplexing failed: {}", pdc_last_error());
  if (ack != nullptr) {
    deliver(ack, out_uci.data(), res.n_harq_ack);
    ack->on_end_softbits();
  }
  if (csi1 != nullptr) {
    deliver(csi1, out_uci.data() + res.n_harq_ack, res.n_csi_part1);
    csi1->on_end_softbits(); // the CSI Part 1 decoder may answer with set_csi_part2
  }
  if (csi2 != nullptr && csi2_enc != 0) {
    ok = run(res); // same HARQ-ACK and CSI Part 1; CSI Part 2 taken out of the UL-SCH elements
    srsran_assert(ok, "UL-SCH demultiplexing failed: {}", pdc_last_error());
    deliver(csi2, out_uci.data() + res.n_harq_ack + res.n_csi_part1, res.n_csi_part2);
    csi2->on_end_softbits();
  }
  deliver(sch, out_sch.data(), res.n_sch);
  sch->on_end_softbits();
  sch = ack = csi1 = csi2 = nullptr;
}

std::unique_ptr<ulsch_demultiplex> srsran::cuda::create_ulsch_demultiplex_cuda(std::shared_ptr<context> ctx)
{
  if (!ctx) {
    return nullptr;
  }
  return std::make_unique<ulsch_demultiplex_cuda>(std::move(ctx));
}

// ---- soft demapper ---------------------------------------------------------------------------------------------------

void demodulation_mapper_cuda::demodulate_soft(span<log_likelihood_ratio> llrs,
                                               span<const cf_t>           symbols,
                                               span<const float>          noise_vars,
                                               modulation_scheme          mod)
{
  srsran_assert(symbols.size() == noise_vars.size(), "Inputs symbols and noise_vars must have the same length.");
  srsran_assert(symbols.size() * get_bits_per_symbol(mod) == llrs.size(), "Input and output lengths are incompatible.");
  int m = PDC_MOD_QAM256;
  switch (mod) {
    case modulation_scheme::PI_2_BPSK:
      m = PDC_MOD_PI_2_BPSK;
      break;
    case modulation_scheme::BPSK:
      m = PDC_MOD_BPSK;
      break;
    case modulation_scheme::QPSK:
      m = PDC_MOD_QPSK;
      break;
    case modulation_scheme::QAM16:
      m = PDC_MOD_QAM16;
      break;
    case modulation_scheme::QAM64:
      m = PDC_MOD_QAM64;
      break;
    default:
      break;
  }
  // log_likelihood_ratio is a one-byte wrapper of int8_t and cf_t is two floats: the spans are passed as they are.
  static_assert(sizeof(log_likelihood_ratio) == 1 && sizeof(cf_t) == 8, "unexpected layout");
  int rc = pdc_demodulate_soft(ctx->get(),
                               reinterpret_cast<int8_t*>(llrs.data()),
                               reinterpret_cast<const float*>(symbols.data()),
                               noise_vars.data(),
                               static_cast<uint32_t>(symbols.size()),
                               m);
  srsran_assert(rc == PDC_OK, "pdc_demodulate_soft failed: {}", pdc_last_error());
  (void)rc;
}

namespace {
class channel_modulation_factory_cuda : public channel_modulation_factory
{
public:
  explicit channel_modulation_factory_cuda(std::shared_ptr<context> c) :
    ctx(std::move(c)), sw(create_channel_modulation_sw_factory())
  {
  }
  std::unique_ptr<modulation_mapper>   create_modulation_mapper() override { return sw->create_modulation_mapper(); }
  std::unique_ptr<demodulation_mapper> create_demodulation_mapper() override
  {
    return std::make_unique<demodulation_mapper_cuda>(ctx);
  }
  std::unique_ptr<evm_calculator> create_evm_calculator() override { return sw->create_evm_calculator(); }

private:
  std::shared_ptr<context>                    ctx;
  std::shared_ptr<channel_modulation_factory> sw;
};
} // namespace

std::shared_ptr<channel_modulation_factory> srsran::cuda::create_channel_modulation_cuda_factory(std::shared_ptr<context> ctx)
{
  return std::make_shared<channel_modulation_factory_cuda>(std::move(ctx));
}

// ---- batched PUSCH decoder -------------------------------------------------------------------------------------------

pusch_decoder_batch_cuda::pusch_decoder_batch_cuda(std::shared_ptr<context>           c,
                                                   unsigned                           queue,
                                                   std::unique_ptr<ldpc_segmenter_rx> seg) :
  ctx(std::move(c)), queue_id(queue), segmenter(std::move(seg))
{
  // queue = context::no_queue: take any free one. A queue is never shared with another batch or accelerator object.
  if (queue_id == context::no_queue) {
    queue_id = ctx->claim_queue();
  } else if (!ctx->claim_queue(queue_id)) {
    queue_id = context::no_queue;
  }
  report_fatal_error_if_not(queue_id != context::no_queue,
                            "pusch_decoder_batch_cuda: no free batch queue (context::config::nof_queues)");
}

pusch_decoder_batch_cuda::~pusch_decoder_batch_cuda()
{
  pdc_wait(ctx->get(), queue_id);
  ctx->release_queue(queue_id);
  pdc_host_free(llr_staging);
  pdc_host_free(tb_staging);
}

std::unique_ptr<pusch_decoder> pusch_decoder_batch_cuda::create()
{
  return std::make_unique<pusch_decoder_cuda>(*this);
}

pusch_decoder_buffer& pusch_decoder_cuda::new_data(span<uint8_t>           transport_block,
                                                   unique_rx_buffer        rm_buffer,
                                                   pusch_decoder_notifier& notifier,
                                                   const configuration&    cfg)
{
  srsran_assert(st == state::idle, "Invalid state: the decoder is busy.");
  const unsigned nof_cb = ldpc::compute_nof_codeblocks(units::bytes(transport_block.size()).to_bits(), cfg.base_graph);
  srsran_assert(nof_cb == rm_buffer.get().get_nof_codeblocks(),
                "Wrong number of codeblocks {} (expected {}).",
                rm_buffer.get().get_nof_codeblocks(),
                nof_cb);
  current.decoder         = this;
  current.transport_block = transport_block;
  current.rm_buffer       = std::move(rm_buffer);
  current.notifier        = &notifier;
  current.cfg             = cfg;
  current.llrs.clear();
  expected_softbits.reset();
  if (cfg.new_data) {
    current.rm_buffer.get().reset_codeblocks_crc(); // pusch_decoder_impl.cpp:125-127
  }
  st = state::collecting;
  return *this;
}

void pusch_decoder_cuda::set_nof_softbits(units::bits nof_softbits)
{
  // The batch starts at flush(); like pusch_decoder_hw_impl::set_nof_softbits this only records the expectation.
  expected_softbits = nof_softbits.value();
}

span<log_likelihood_ratio> pusch_decoder_cuda::get_next_block_view(unsigned block_size)
{
  srsran_assert(st == state::collecting, "Invalid state.");
  view.resize(block_size);
  return view;
}

void pusch_decoder_cuda::on_new_softbits(span<const log_likelihood_ratio> softbits)
{
  srsran_assert(st == state::collecting, "Invalid state.");
  current.llrs.insert(current.llrs.end(), softbits.begin(), softbits.end());
}

void pusch_decoder_cuda::on_end_softbits()
{
  srsran_assert(st == state::collecting, "Invalid state.");
  srsran_assert(!expected_softbits.has_value() || *expected_softbits == current.llrs.size(),
                "The number of UL-SCH softbits does not match the expected value.");
  srsran_assert(current.llrs.size() % get_bits_per_symbol(current.cfg.mod) == 0,
                "The number of soft bits must be multiple of the modulation order.");
  st = state::decoding;
  batch.queue(std::move(current));
}

bool pusch_decoder_batch_cuda::flush()
{
  if (queued.empty()) {
    return true;
  }
  std::vector<queued_tb> work;
  work.swap(queued);
  cbs.clear();
  tbs.clear();
  decode_mask.clear();
  size_t n_llr = 0, tb_bytes = 0;
  for (const queued_tb& q : work) {
    n_llr += q.llrs.size();
    tb_bytes += (q.transport_block.size() * 8 + 24 + 31) / 32 * 4;
  }
  if (n_llr > llr_capacity) {
    pdc_host_free(llr_staging);
    llr_capacity = n_llr + n_llr / 4 + 64;
    llr_staging  = static_cast<int8_t*>(pdc_host_alloc(llr_capacity));
  }
  if (tb_bytes > tb_capacity) {
    pdc_host_free(tb_staging);
    tb_capacity = tb_bytes + tb_bytes / 4 + 64;
    tb_staging  = static_cast<uint8_t*>(pdc_host_alloc(tb_capacity));
  }
  srsran_assert(llr_staging != nullptr && tb_staging != nullptr, "pdc_host_alloc failed: {}", pdc_last_error());
  size_t llr_off = 0, tb_off = 0;
  for (size_t i_tb = 0; i_tb != work.size(); ++i_tb) {
    queued_tb&     q        = work[i_tb];
    const unsigned tbs_bits = static_cast<unsigned>(q.transport_block.size() * 8);
    // Segmentation by the reference's own segmenter (ldpc_segmenter_rx_impl), as in pusch_decoder_impl::on_end_softbits.
    segmenter_config seg_cfg;
    seg_cfg.base_graph     = q.cfg.base_graph;
    seg_cfg.rv             = q.cfg.rv;
    seg_cfg.mod            = q.cfg.mod;
    seg_cfg.Nref           = q.cfg.Nref;
    seg_cfg.nof_layers     = q.cfg.nof_layers;
    seg_cfg.nof_ch_symbols = static_cast<unsigned>(q.llrs.size()) / get_bits_per_symbol(q.cfg.mod);
    static_vector<described_rx_codeblock, MAX_NOF_SEGMENTS> blocks;
    segmenter->segment(blocks, q.llrs, tbs_bits, seg_cfg);
    const unsigned C = static_cast<unsigned>(blocks.size());
    // A lone codeblock carries the transport-block CRC (CRC16 up to 3824 bits, CRC24A above), the others CRC24B.
    const int  crc_kind = (C > 1) ? PDC_CRC24B : (tbs_bits > 3824 ? PDC_CRC24A : PDC_CRC16);
    span<bool> crcs     = q.rm_buffer.get().get_codeblocks_crc();
    memcpy(llr_staging + llr_off, q.llrs.data(), q.llrs.size());
    pdc_tb_desc t   = {};
    t.first_cb      = static_cast<uint32_t>(cbs.size());
    t.nof_cb        = C;
    t.tbs_bits      = tbs_bits;
    t.out_offset    = static_cast<uint32_t>(tb_off);
    tbs.push_back(t);
    tb_off += (tbs_bits + 24 + 31) / 32 * 4;
    for (unsigned k = 0; k != C; ++k) {
      const codeblock_metadata& m = blocks[k].second;
      pdc_cb_desc               d = {};
      d.llr_offset                = static_cast<uint32_t>(llr_off + m.cb_specific.cw_offset);
      d.rm_length                 = m.cb_specific.rm_length;
      d.harq_id                   = q.rm_buffer.get().get_absolute_codeblock_id(k);
      d.nref                      = m.tb_common.Nref;
      d.lifting_size              = static_cast<uint16_t>(m.tb_common.lifting_size);
      d.nof_filler                = static_cast<uint16_t>(m.cb_specific.nof_filler_bits);
      d.base_graph                = (m.tb_common.base_graph == ldpc_base_graph_type::BG1) ? 1 : 2;
      d.qm                        = static_cast<uint8_t>(get_bits_per_symbol(m.tb_common.mod));
      d.rv                        = static_cast<uint8_t>(m.tb_common.rv);
      d.crc_kind                  = static_cast<uint8_t>(crc_kind);
      d.max_iter                  = static_cast<uint8_t>(q.cfg.nof_ldpc_iterations);
      // Codeblocks whose CRC already passed are only combined, not decoded again (pusch_decoder_impl.cpp:335-345).
      const bool decode = !crcs[k];
      d.flags           = static_cast<uint8_t>(PDC_CB_DEMATCH | (q.cfg.new_data ? PDC_CB_NEW_DATA : 0) |
                                     (q.cfg.use_early_stop ? PDC_CB_EARLY_STOP : 0) | (decode ? PDC_CB_DECODE : 0));
      d.tb_index        = static_cast<uint16_t>(i_tb);
      cbs.push_back(d);
      decode_mask.push_back(decode ? 1 : 0);
    }
    llr_off += q.llrs.size();
  }
  cb_results.assign(cbs.size(), pdc_cb_result{});
  tb_results.assign(tbs.size(), pdc_tb_result{});
  int rc = pdc_submit(ctx->get(), queue_id, cbs.data(), static_cast<uint32_t>(cbs.size()), llr_staging, n_llr, tbs.data(),
                      static_cast<uint32_t>(tbs.size()), cb_results.data(), nullptr, tb_results.data(), tb_staging);
  if (rc == PDC_OK) {
    rc = pdc_wait(ctx->get(), queue_id);
  }
  const bool ok = (rc == PDC_OK);
  for (size_t i_tb = 0; i_tb != work.size(); ++i_tb) {
    queued_tb&           q = work[i_tb];
    const pdc_tb_desc&   t = tbs[i_tb];
    span<bool>           crcs = q.rm_buffer.get().get_codeblocks_crc();
    pusch_decoder_result result;
    result.tb_crc_ok            = false;
    result.nof_codeblocks_total = t.nof_cb;
    if (ok) {
      for (unsigned k = 0; k != t.nof_cb; ++k) {
        if (!decode_mask[t.first_cb + k]) {
          continue;
        }
        const pdc_cb_result& r = cb_results[t.first_cb + k];
        if (r.crc_ok) {
          crcs[k] = true;
          result.ldpc_decoder_stats.update(r.iters);
        } else {
          result.ldpc_decoder_stats.update(q.cfg.nof_ldpc_iterations);
        }
      }
      const bool all_ok = std::all_of(crcs.begin(), crcs.end(), [](bool b) { return b; });
      if (t.nof_cb == 1) {
        // The codeblock CRC is the transport-block CRC (pusch_decoder_impl.cpp:404-412).
        result.tb_crc_ok = crcs[0];
        if (result.tb_crc_ok) {
          memcpy(q.transport_block.data(), tb_staging + t.out_offset, q.transport_block.size());
        }
      } else if (all_ok) {
        memcpy(q.transport_block.data(), tb_staging + t.out_offset, q.transport_block.size());
        if (tb_results[i_tb].tb_crc_ok) {
          result.tb_crc_ok = true;
        } else {
          // All codeblocks pass but the transport block does not: start over (pusch_decoder_impl.cpp:428-433).
          q.rm_buffer.get().reset_codeblocks_crc();
        }
      }
    }
    if (result.tb_crc_ok) {
      q.rm_buffer.release();
    } else {
      q.rm_buffer.unlock();
    }
    q.decoder->st = pusch_decoder_cuda::state::idle;
    q.notifier->on_sch_data(result);
  }
  return ok;
}

namespace {
class pusch_decoder_factory_cuda : public pusch_decoder_factory
{
public:
  explicit pusch_decoder_factory_cuda(std::shared_ptr<pusch_decoder_batch_cuda> b) : batch(std::move(b)) {}
  std::unique_ptr<pusch_decoder> create() override { return batch->create(); }

private:
  std::shared_ptr<pusch_decoder_batch_cuda> batch;
};
} // namespace

std::shared_ptr<pusch_decoder_factory>
srsran::cuda::create_pusch_decoder_factory_cuda(std::shared_ptr<pusch_decoder_batch_cuda> batch)
{
  return std::make_shared<pusch_decoder_factory_cuda>(std::move(batch));
}

// ---- downlink twin: LDPC encoder ---------------------------------------------------------------------------------------

void ldpc_encoder_cuda::encode(bit_buffer& output, const bit_buffer& input, const codeblock_metadata::tb_common_metadata& cfg)
{
  const unsigned bg = (cfg.base_graph == ldpc_base_graph_type::BG1) ? 1 : 2;
  const unsigned Z  = static_cast<unsigned>(cfg.lifting_size);
  const unsigned K = (bg == 1 ? 22 : 10) * Z, N = (bg == 1 ? 66 : 50) * Z;
  srsran_assert(input.size() == K, "Input size {} does not match the codeblock size {}.", input.size(), K);
  srsran_assert(output.size() <= N, "Output size {} exceeds the full codeblock size {}.", output.size(), N);
  // bit_buffer stores bits packed MSB first: the message goes as it is (filler bits read as zeros, like in the
  // reference's encoders, ldpc_encoder_impl.cpp:52-60).
  std::vector<uint8_t> msg((K + 7) / 8, 0);
  for (unsigned i = 0; i != K; ++i) {
    const uint8_t b = input.extract(i, 1);
    msg[i >> 3] |= static_cast<uint8_t>((b & 1u) << (7 - (i & 7)));
  }
  bits.resize(N);
  int rc = pdc_ldpc_encode(ctx->get(), static_cast<int>(bg), static_cast<int>(Z), msg.data(), bits.data());
  srsran_assert(rc == PDC_OK, "pdc_ldpc_encode failed: {}", pdc_last_error());
  (void)rc;
  // The caller may ask for a prefix of the codeblock only (ldpc_encoder.h:47-50).
  for (unsigned i = 0, n = output.size(); i != n; ++i) {
    output.insert(bits[i], i, 1);
  }
}

namespace {
class ldpc_encoder_factory_cuda : public ldpc_encoder_factory
{
public:
  explicit ldpc_encoder_factory_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::unique_ptr<ldpc_encoder> create() override { return std::make_unique<ldpc_encoder_cuda>(ctx); }

private:
  std::shared_ptr<context> ctx;
};
} // namespace

std::shared_ptr<ldpc_encoder_factory> srsran::cuda::create_ldpc_encoder_factory_cuda(std::shared_ptr<context> ctx)
{
  return std::make_shared<ldpc_encoder_factory_cuda>(std::move(ctx));
}

// ---- downlink twin: PDSCH encoder accelerator (codeblock mode) --------------------------------------------------------

void hw_accelerator_pdsch_enc_cuda::reserve_queue()
{
  descs.clear();
  enqueued.clear();
  msgs.clear();
  out.clear();
  encoded = false;
}

void hw_accelerator_pdsch_enc_cuda::configure_operation(const hal::hw_pdsch_encoder_configuration& cfg, unsigned cb_index)
{
  if (descs.size() <= cb_index) {
    descs.resize(cb_index + 1);
    enqueued.resize(cb_index + 1, 0);
  }
  pdc_enc_desc& d = descs[cb_index];
  d               = {};
  d.rm_length     = cfg.rm_length;
  d.nref          = cfg.Nref;
  d.lifting_size  = static_cast<uint16_t>(cfg.lifting_size);
  d.nof_filler    = static_cast<uint16_t>(cfg.nof_filler_bits);
  d.base_graph    = (cfg.base_graph_index == ldpc_base_graph_type::BG1) ? 1 : 2;
  d.qm            = static_cast<uint8_t>(get_bits_per_symbol(cfg.modulation));
  d.rv            = static_cast<uint8_t>(cfg.rv);
}

bool hw_accelerator_pdsch_enc_cuda::enqueue_operation(span<const uint8_t> data, span<const uint8_t>, unsigned cb_index)
{
  srsran_assert(cb_index < descs.size(), "Codeblock {} was not configured.", cb_index);
  pdc_enc_desc&  d = descs[cb_index];
  const unsigned K = ((d.base_graph == 1) ? 22u : 10u) * d.lifting_size;
  // The segment arrives without its filler bits (pdsch_encoder_hw_impl.cpp:92-95): they are zeros behind it.
  d.msg_offset = static_cast<uint32_t>(msgs.size());
  msgs.resize(msgs.size() + (K + 7) / 8, 0);
  const unsigned n_bits = K - d.nof_filler;
  memcpy(msgs.data() + d.msg_offset, data.data(), std::min<size_t>(data.size(), (n_bits + 7) / 8));
  if (n_bits % 8) {
    msgs[d.msg_offset + n_bits / 8] &= static_cast<uint8_t>(0xff00u >> (n_bits % 8));
  }
  enqueued[cb_index] = 1;
  encoded            = false;
  return true;
}

bool hw_accelerator_pdsch_enc_cuda::dequeue_operation(span<uint8_t> data, span<uint8_t> aux_data, unsigned segment_index)
{
  if (segment_index >= descs.size() || !enqueued[segment_index]) {
    return false;
  }
  if (!encoded) {
    // Everything enqueued so far in one batch.
    std::vector<pdc_enc_desc> batch;
    uint32_t                  off = 0;
    for (size_t i = 0; i != descs.size(); ++i) {
      if (enqueued[i]) {
        descs[i].out_offset = off;
        off += descs[i].rm_length;
        batch.push_back(descs[i]);
      }
    }
    out.assign(off + 8, 0);
    int rc = pdc_encode(ctx->get(), batch.data(), static_cast<uint32_t>(batch.size()), msgs.data(), msgs.size(), out.data(),
                        out.size());
    srsran_assert(rc == PDC_OK, "pdc_encode failed: {}", pdc_last_error());
    if (rc != PDC_OK) {
      return false;
    }
    encoded = true;
  }
  const pdc_enc_desc& d = descs[segment_index];
  srsran_assert(data.size() == d.rm_length, "Wrong codeblock length.");
  memcpy(data.data(), out.data() + d.out_offset, d.rm_length);
  for (size_t i = 0, n = std::min<size_t>(aux_data.size(), (d.rm_length + 7) / 8); i != n; ++i) {
    uint8_t v = 0;
    for (unsigned b = 0; b != 8 && 8 * i + b < d.rm_length; ++b) {
      v |= static_cast<uint8_t>((out[d.out_offset + 8 * i + b] & 1u) << (7 - b));
    }
    aux_data[i] = v;
  }
  return true;
}

namespace {
class hw_accelerator_pdsch_enc_factory_cuda : public hal::hw_accelerator_pdsch_enc_factory
{
public:
  explicit hw_accelerator_pdsch_enc_factory_cuda(std::shared_ptr<context> c) : ctx(std::move(c)) {}
  std::unique_ptr<hal::hw_accelerator_pdsch_enc> create() override
  {
    return std::make_unique<hw_accelerator_pdsch_enc_cuda>(ctx);
  }

private:
  std::shared_ptr<context> ctx;
};
} // namespace

std::shared_ptr<hal::hw_accelerator_pdsch_enc_factory>
srsran::cuda::create_hw_accelerator_pdsch_enc_factory_cuda(std::shared_ptr<context> ctx)
{
  return std::make_shared<hw_accelerator_pdsch_enc_factory_cuda>(std::move(ctx));
}
