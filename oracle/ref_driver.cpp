// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// C-ABI driver over the *unmodified* reference implementation (srsRAN-5G-ER inside /root/reference). It is compiled
// together with the reference's own sources (where they lie, see oracle/Makefile) into oracle/_ref/libsrsref.so and is
// used (a) to pin oracle/pusch_oracle.c against the real reference, (b) to generate the fixtures in tests/golden/ and
// (c) as the "reference" CPU arm of bench.py. Nothing under srsran_edgeric_5g_b200/ may load this library.
//
// Every entry point is a thin call into a reference class through its public factory:
//   create_ldpc_decoder_factory_sw / create_ldpc_rate_dematcher_factory_sw / create_crc_calculator_factory_sw
//     (include/srsran/phy/upper/channel_coding/channel_coding_factories.h:50-77)
//   pusch_decoder_impl (lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.h:43)
//   ldpc_segmenter_tx + ldpc_encoder + ldpc_rate_matcher, used as in pdsch_encoder_impl.cpp:30-75.

#include "lib/phy/upper/channel_processors/pusch/pusch_codeblock_decoder.h"
#include "lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.h"
#include "lib/phy/upper/channel_processors/pusch/ulsch_demultiplex_impl.h"
#include "lib/phy/upper/sequence_generators/pseudo_random_generator_impl.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_codeword_buffer.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_buffer.h"
#include "srsran/phy/upper/channel_coding/channel_coding_factories.h"
#include "srsran/phy/upper/channel_modulation/channel_modulation_factories.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include "srsran/ran/pusch/pusch_mcs.h"
#include "srsran/ran/sch/ldpc_base_graph.h"
#include "srsran/ran/sch/tbs_calculator.h"
#include "srsran/srsvec/bit.h"
#include "srsran/support/cpu_features.h"
#include <atomic>
#include <functional>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

using namespace srsran;

namespace {

crc_generator_poly to_poly(int crc_kind)
{
  // 1 = CRC16, 2 = CRC24A, 3 = CRC24B, 4 = CRC24C, 5 = CRC11, 6 = CRC6.
  switch (crc_kind) {
    case 1:
      return crc_generator_poly::CRC16;
    case 2:
      return crc_generator_poly::CRC24A;
    case 4:
      return crc_generator_poly::CRC24C;
    case 5:
      return crc_generator_poly::CRC11;
    case 6:
      return crc_generator_poly::CRC6;
    default:
      return crc_generator_poly::CRC24B;
  }
}

modulation_scheme to_mod(int qm)
{
  switch (qm) {
    case 1:
      return modulation_scheme::BPSK;
    case 2:
      return modulation_scheme::QPSK;
    case 4:
      return modulation_scheme::QAM16;
    case 6:
      return modulation_scheme::QAM64;
    default:
      return modulation_scheme::QAM256;
  }
}

// A bit_buffer view over caller memory (the span constructor of bit_buffer is protected).
bit_buffer view_bits(uint8_t* ptr, unsigned nbits)
{
  return bit_buffer::from_bytes(span<uint8_t>(ptr, (nbits + 7) / 8)).first(nbits);
}

struct cb_tools {
  std::unique_ptr<ldpc_decoder>        dec;
  std::unique_ptr<ldpc_rate_dematcher> dem;
  std::unique_ptr<crc_calculator>      crc[7];
};

std::unique_ptr<cb_tools> make_tools(const char* type)
{
  auto t     = std::make_unique<cb_tools>();
  auto dec_f = create_ldpc_decoder_factory_sw(type);
  auto dem_f = create_ldpc_rate_dematcher_factory_sw(type);
  auto crc_f = create_crc_calculator_factory_sw("auto");
  if (!dec_f || !dem_f || !crc_f) {
    return nullptr;
  }
  t->dec = dec_f->create();
  t->dem = dem_f->create();
  for (int k = 1; k <= 6; ++k) {
    t->crc[k] = crc_f->create(to_poly(k));
  }
  return t;
}

codeblock_metadata make_meta(int bg, int Z, int rv, int qm, int nref, int nof_filler, int crc_bits, int rm_length)
{
  codeblock_metadata m;
  m.tb_common.base_graph        = (bg == 2) ? ldpc_base_graph_type::BG2 : ldpc_base_graph_type::BG1;
  m.tb_common.lifting_size      = static_cast<ldpc::lifting_size_t>(Z);
  m.tb_common.rv                = rv;
  m.tb_common.mod               = to_mod(qm);
  m.tb_common.Nref              = nref;
  m.cb_specific.full_length     = ((bg == 2) ? 50 : 66) * Z;
  m.cb_specific.rm_length       = rm_length;
  m.cb_specific.nof_filler_bits = nof_filler;
  m.cb_specific.nof_crc_bits    = crc_bits;
  return m;
}

// A HARQ buffer the driver owns, so that soft bits, data bits and CRC flags can be read back after a decode.
class driver_rx_buffer : public unique_rx_buffer::callback
{
public:
  driver_rx_buffer(unsigned nof_cb, unsigned max_cb_size) :
    soft(nof_cb, std::vector<log_likelihood_ratio>(max_cb_size)),
    data(nof_cb, std::vector<uint8_t>(max_cb_size / 8 + 8)),
    crcs(new bool[nof_cb]()),
    n(nof_cb)
  {
  }
  ~driver_rx_buffer() override { delete[] crcs; }
  unsigned   get_nof_codeblocks() const override { return n; }
  void       reset_codeblocks_crc() override { std::fill(crcs, crcs + n, false); }
  span<bool> get_codeblocks_crc() override { return span<bool>(crcs, n); }
  unsigned   get_absolute_codeblock_id(unsigned cb) const override { return cb; }
  span<log_likelihood_ratio> get_codeblock_soft_bits(unsigned cb, unsigned size) override
  {
    return span<log_likelihood_ratio>(soft[cb]).first(size);
  }
  bit_buffer get_codeblock_data_bits(unsigned cb, unsigned size) override
  {
    return view_bits(data[cb].data(), size);
  }
  void lock() override {}
  void unlock() override {}
  void release() override { released = true; }

  std::vector<std::vector<log_likelihood_ratio>> soft;
  std::vector<std::vector<uint8_t>>              data;
  bool*                                          crcs;
  unsigned                                       n;
  bool                                           released = false;
};

class driver_notifier : public pusch_decoder_notifier
{
public:
  void                 on_sch_data(const pusch_decoder_result& r) override { result = r; }
  pusch_decoder_result result;
};

struct pusch_handle {
  std::unique_ptr<pusch_decoder_impl> decoder;
  std::unique_ptr<driver_rx_buffer>   buffer;
};

} // namespace

extern "C" {

/// Which variant "auto" resolves to on this host: 3 = avx512, 2 = avx2, 1 = generic
/// (dispatch rule of channel_coding_factories.cpp:100-124).
int ref_auto_variant()
{
#ifdef __x86_64__
  if (cpu_supports_feature(cpu_feature::avx512f) && cpu_supports_feature(cpu_feature::avx512bw)) {
    return 3;
  }
  if (cpu_supports_feature(cpu_feature::avx2)) {
    return 2;
  }
#endif
  return 1;
}

void* ref_tools_create(const char* type)
{
  return make_tools(type).release();
}

void ref_tools_destroy(void* h)
{
  delete static_cast<cb_tools*>(h);
}

/// ldpc_decoder::decode. \c out holds ceil(K/8) bytes. crc_kind 0 = nullptr (no early stop).
/// Returns the iteration count (>=1) when the optional has a value, 0 when it is empty.
int ref_ldpc_decode(void*         h,
                    int           bg,
                    int           Z,
                    const int8_t* llr,
                    int           n_llr,
                    int           nof_filler,
                    int           crc_kind,
                    int           max_iter,
                    uint8_t*      out)
{
  auto*    t = static_cast<cb_tools*>(h);
  unsigned K = ((bg == 2) ? 10 : 22) * Z;

  ldpc_decoder::configuration cfg;
  cfg.block_conf                    = make_meta(bg, Z, 0, 2, 0, nof_filler, (crc_kind == 1) ? 16 : 24, n_llr);
  cfg.algorithm_conf.max_iterations = max_iter;

  bit_buffer                       bits = view_bits(out, K);
  span<const log_likelihood_ratio> in(reinterpret_cast<const log_likelihood_ratio*>(llr), n_llr);
  std::optional<unsigned>          r = t->dec->decode(bits, in, (crc_kind == 0) ? nullptr : t->crc[crc_kind].get(), cfg);
  return r.has_value() ? static_cast<int>(r.value()) : 0;
}

/// ldpc_rate_dematcher::rate_dematch. \c out is the in/out HARQ buffer of length N (66Z or 50Z).
void ref_rate_dematch(void*         h,
                      int8_t*       out,
                      int           N,
                      const int8_t* in,
                      int           E,
                      int           new_data,
                      int           rv,
                      int           qm,
                      int           nref,
                      int           nof_filler)
{
  auto* t  = static_cast<cb_tools*>(h);
  int   bg = (N % 66 == 0) ? 1 : 2;
  int   Z  = N / ((bg == 1) ? 66 : 50);
  codeblock_metadata m = make_meta(bg, Z, rv, qm, nref, nof_filler, 24, E);
  t->dem->rate_dematch(span<log_likelihood_ratio>(reinterpret_cast<log_likelihood_ratio*>(out), N),
                       span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(in), E),
                       new_data != 0,
                       m);
}

/// crc_calculator::calculate over the first nbits of a packed (MSB first) buffer.
unsigned ref_crc(void* h, int crc_kind, const uint8_t* packed, int nbits)
{
  auto*      t = static_cast<cb_tools*>(h);
  bit_buffer bits = view_bits(const_cast<uint8_t*>(packed), nbits);
  return t->crc[crc_kind]->calculate(bits);
}

/// pusch_codeblock_decoder::decode (dematch + decode + CRC bookkeeping) for one codeblock.
int ref_cb_decode(void*         h,
                  int8_t*       rm_buffer,
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
                  uint8_t*      out)
{
  auto*              t  = static_cast<cb_tools*>(h);
  int                bg = (N % 66 == 0) ? 1 : 2;
  int                Z  = N / ((bg == 1) ? 66 : 50);
  unsigned           K  = ((bg == 2) ? 10 : 22) * Z;
  codeblock_metadata m  = make_meta(bg, Z, rv, qm, nref, nof_filler, (crc_kind == 1) ? 16 : 24, E);

  span<log_likelihood_ratio>       rm(reinterpret_cast<log_likelihood_ratio*>(rm_buffer), N);
  span<const log_likelihood_ratio> llr(reinterpret_cast<const log_likelihood_ratio*>(in), E);
  t->dem->rate_dematch(rm, llr, new_data != 0, m);

  ldpc_decoder::configuration cfg;
  cfg.block_conf                    = m;
  cfg.algorithm_conf.max_iterations = max_iter;
  bit_buffer      bits = view_bits(out, K);
  crc_calculator* crc = t->crc[crc_kind].get();
  if (use_early_stop) {
    std::optional<unsigned> r = t->dec->decode(bits, rm, crc, cfg);
    return r.has_value() ? static_cast<int>(r.value()) : 0;
  }
  t->dec->decode(bits, rm, nullptr, cfg);
  return (crc->calculate(bits.first(K - nof_filler)) == 0) ? max_iter : 0;
}

/// TX chain of the reference (segmenter_tx + encoder + rate matcher): TB bytes -> codeword bits (one bit per byte).
/// Returns the number of codeblocks.
int ref_tb_encode(const uint8_t* tb,
                  int            tb_bytes,
                  int            bg,
                  int            rv,
                  int            qm,
                  int            nref,
                  int            nof_layers,
                  int            nof_ch_symbols,
                  uint8_t*       cw_bits)
{
  auto crc_f = create_crc_calculator_factory_sw("auto");
  auto seg   = create_ldpc_segmenter_tx_factory_sw(crc_f)->create();
  auto enc   = create_ldpc_encoder_factory_sw("auto")->create();
  auto rm    = create_ldpc_rate_matcher_factory_sw()->create();

  segmenter_config cfg;
  cfg.base_graph     = (bg == 2) ? ldpc_base_graph_type::BG2 : ldpc_base_graph_type::BG1;
  cfg.rv             = rv;
  cfg.mod            = to_mod(qm);
  cfg.Nref           = nref;
  cfg.nof_layers     = nof_layers;
  cfg.nof_ch_symbols = nof_ch_symbols;

  static_vector<described_segment, MAX_NOF_SEGMENTS> segments;
  seg->segment(segments, span<const uint8_t>(tb, tb_bytes), cfg);

  dynamic_bit_buffer full;
  dynamic_bit_buffer packed;
  unsigned           offset = 0;
  for (const described_segment& s : segments) {
    const codeblock_metadata& m = s.get_metadata();
    full.resize(m.cb_specific.full_length);
    enc->encode(full, s.get_data(), m.tb_common);
    packed.resize(m.cb_specific.rm_length);
    rm->rate_match(packed, full, m);
    srsvec::bit_unpack(span<uint8_t>(cw_bits + offset, m.cb_specific.rm_length), packed);
    offset += m.cb_specific.rm_length;
  }
  return static_cast<int>(segments.size());
}

/// Encodes one codeblock: msg holds K bits, one per byte (fillers as 0); out receives n_out (<= 66Z/50Z) bits.
void ref_ldpc_encode(int bg, int Z, const uint8_t* msg, uint8_t* out, int n_out)
{
  static thread_local auto enc = create_ldpc_encoder_factory_sw("auto")->create();
  unsigned                 K   = ((bg == 2) ? 10 : 22) * Z;
  dynamic_bit_buffer       in(K);
  srsvec::bit_pack(in, span<const uint8_t>(msg, K));
  dynamic_bit_buffer                     cw(n_out);
  codeblock_metadata::tb_common_metadata c;
  c.base_graph   = (bg == 2) ? ldpc_base_graph_type::BG2 : ldpc_base_graph_type::BG1;
  c.lifting_size = static_cast<ldpc::lifting_size_t>(Z);
  enc->encode(cw, in, c);
  srsvec::bit_unpack(span<uint8_t>(out, n_out), cw);
}

/// Rx segmentation metadata, 6 ints per codeblock: {Z, full_length, rm_length, nof_filler, cw_offset, nof_crc_bits}.
int ref_segment_rx(int tbs_bits, int bg, int rv, int qm, int nref, int nof_layers, int n_llr, int* meta)
{
  auto                              seg = create_ldpc_segmenter_rx_factory_sw()->create();
  std::vector<log_likelihood_ratio> dummy(n_llr);
  segmenter_config                  cfg;
  cfg.base_graph     = (bg == 2) ? ldpc_base_graph_type::BG2 : ldpc_base_graph_type::BG1;
  cfg.rv             = rv;
  cfg.mod            = to_mod(qm);
  cfg.Nref           = nref;
  cfg.nof_layers     = nof_layers;
  cfg.nof_ch_symbols = n_llr / qm;
  static_vector<described_rx_codeblock, MAX_NOF_SEGMENTS> cbs;
  seg->segment(cbs, dummy, tbs_bits, cfg);
  int i = 0;
  for (const auto& cb : cbs) {
    const codeblock_metadata& m = cb.second;
    meta[i++]                   = static_cast<int>(m.tb_common.lifting_size);
    meta[i++]                   = m.cb_specific.full_length;
    meta[i++]                   = m.cb_specific.rm_length;
    meta[i++]                   = m.cb_specific.nof_filler_bits;
    meta[i++]                   = m.cb_specific.cw_offset;
    meta[i++]                   = m.cb_specific.nof_crc_bits;
  }
  return static_cast<int>(cbs.size());
}

/// Creates a pusch_decoder_impl (synchronous: no executor) with a driver-owned HARQ buffer of nof_cb codeblocks.
void* ref_pusch_create(const char* type, int nof_cb)
{
  auto crc_f = create_crc_calculator_factory_sw("auto");
  auto dec_f = create_ldpc_decoder_factory_sw(type);
  auto dem_f = create_ldpc_rate_dematcher_factory_sw(type);
  auto seg_f = create_ldpc_segmenter_rx_factory_sw();
  if (!crc_f || !dec_f || !dem_f || !seg_f) {
    return nullptr;
  }
  std::vector<std::unique_ptr<pusch_codeblock_decoder>> cbd(1);
  pusch_codeblock_decoder::sch_crc                      c1;
  c1.crc16  = crc_f->create(crc_generator_poly::CRC16);
  c1.crc24A = crc_f->create(crc_generator_poly::CRC24A);
  c1.crc24B = crc_f->create(crc_generator_poly::CRC24B);
  cbd[0]    = std::make_unique<pusch_codeblock_decoder>(dem_f->create(), dec_f->create(), c1);
  auto pool = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cbd));

  pusch_decoder_impl::sch_crc c2;
  c2.crc16  = crc_f->create(crc_generator_poly::CRC16);
  c2.crc24A = crc_f->create(crc_generator_poly::CRC24A);
  c2.crc24B = crc_f->create(crc_generator_poly::CRC24B);

  auto* h    = new pusch_handle;
  h->decoder = std::make_unique<pusch_decoder_impl>(seg_f->create(), pool, std::move(c2), nullptr, MAX_RB, 4);
  h->buffer  = std::make_unique<driver_rx_buffer>(nof_cb, ldpc::MAX_CODEBLOCK_SIZE);
  return h;
}

void ref_pusch_destroy(void* h)
{
  delete static_cast<pusch_handle*>(h);
}

/// Fills every soft buffer with a sentinel so that "stale" regions are reproducible.
void ref_pusch_fill_soft(void* hv, int value)
{
  auto* h = static_cast<pusch_handle*>(hv);
  for (auto& v : h->buffer->soft) {
    std::fill(v.begin(), v.end(), log_likelihood_ratio(value));
  }
}

/// One (re)transmission through pusch_decoder_impl::new_data / on_new_softbits / on_end_softbits.
/// stats = {tb_crc_ok, nof_codeblocks_total, nof_observations, min_iter, max_iter, mean_iter*1000}.
void ref_pusch_decode(void*         hv,
                      const int8_t* llrs,
                      int           n_llr,
                      int           tb_bytes,
                      int           bg,
                      int           rv,
                      int           qm,
                      int           nref,
                      int           nof_layers,
                      int           max_iter,
                      int           use_early_stop,
                      int           new_data,
                      int           reset_crcs,
                      uint8_t*      tb_out,
                      int*          stats)
{
  auto* h = static_cast<pusch_handle*>(hv);
  if (reset_crcs) {
    h->buffer->reset_codeblocks_crc();
  }
  pusch_decoder::configuration cfg;
  cfg.base_graph          = (bg == 2) ? ldpc_base_graph_type::BG2 : ldpc_base_graph_type::BG1;
  cfg.rv                  = rv;
  cfg.mod                 = to_mod(qm);
  cfg.Nref                = nref;
  cfg.nof_layers          = nof_layers;
  cfg.nof_ldpc_iterations = max_iter;
  cfg.use_early_stop      = use_early_stop != 0;
  cfg.new_data            = new_data != 0;

  driver_notifier       notifier;
  unique_rx_buffer      ub(*h->buffer);
  pusch_decoder_buffer& buf = h->decoder->new_data(span<uint8_t>(tb_out, tb_bytes), std::move(ub), notifier, cfg);
  buf.on_new_softbits(span<const log_likelihood_ratio>(reinterpret_cast<const log_likelihood_ratio*>(llrs), n_llr));
  buf.on_end_softbits();

  const pusch_decoder_result& r = notifier.result;
  stats[0]                      = r.tb_crc_ok;
  stats[1]                      = r.nof_codeblocks_total;
  stats[2]                      = r.ldpc_decoder_stats.get_nof_observations();
  stats[3]                      = r.ldpc_decoder_stats.get_nof_observations() ? r.ldpc_decoder_stats.get_min() : 0;
  stats[4]                      = r.ldpc_decoder_stats.get_nof_observations() ? r.ldpc_decoder_stats.get_max() : 0;
  stats[5] = r.ldpc_decoder_stats.get_nof_observations() ? static_cast<int>(r.ldpc_decoder_stats.get_mean() * 1000) : 0;
}

/// Copies out the HARQ state of one codeblock: soft bits (n bytes), CRC flag.
int ref_pusch_get_cb(void* hv, int cb, int8_t* soft, int n)
{
  auto* h = static_cast<pusch_handle*>(hv);
  std::memcpy(soft, h->buffer->soft[cb].data(), n);
  return h->buffer->crcs[cb];
}

/// CPU baseline: nof_threads threads, one pusch_codeblock_decoder-equivalent per thread, codeblocks statically
/// partitioned (the reference's threading model, pusch/factories.cpp:84-96). Every codeblock: rate_dematch(new_data)
/// into a per-thread HARQ buffer, then decode with or without early stop. Returns seconds of wall clock for
/// \c repeats passes over the batch of n_cb codeblocks (all of the same shape). iters_out[n_cb] receives the result of
/// the last pass.
double ref_bench_cb_batch(const char*   type,
                          int           nof_threads,
                          int           repeats,
                          int           n_cb,
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
                          int*          iters_out,
                          uint8_t*      bits_out)
{
  int      bg      = (N % 66 == 0) ? 1 : 2;
  int      Z       = N / ((bg == 1) ? 66 : 50);
  unsigned K       = ((bg == 2) ? 10 : 22) * Z;
  unsigned K_bytes = (K + 7) / 8;

  std::vector<std::unique_ptr<cb_tools>> tools;
  for (int i = 0; i != nof_threads; ++i) {
    tools.push_back(make_tools(type));
    if (!tools.back()) {
      return -1.0;
    }
  }
  std::atomic<int> ready{0};
  std::atomic<int> go{0};
  auto             worker = [&](int tid) {
    std::vector<int8_t>  rm(N);
    std::vector<uint8_t> local_bits(K_bytes + 8);
    int                  lo = static_cast<int>(static_cast<long>(n_cb) * tid / nof_threads);
    int                  hi = static_cast<int>(static_cast<long>(n_cb) * (tid + 1) / nof_threads);
    ready.fetch_add(1);
    while (go.load() == 0) {
    }
    for (int r = 0; r != repeats; ++r) {
      for (int cb = lo; cb != hi; ++cb) {
        uint8_t* out = bits_out ? bits_out + static_cast<size_t>(cb) * K_bytes : local_bits.data();
        int      it  = ref_cb_decode(tools[tid].get(),
                               rm.data(),
                               N,
                               llrs + static_cast<size_t>(cb) * E,
                               E,
                               1,
                               rv,
                               qm,
                               nref,
                               nof_filler,
                               crc_kind,
                               use_early_stop,
                               max_iter,
                               out);
        if (iters_out) {
          iters_out[cb] = it;
        }
      }
    }
  };
  std::vector<std::thread> threads;
  for (int i = 0; i != nof_threads; ++i) {
    threads.emplace_back(worker, i);
  }
  while (ready.load() != nof_threads) {
  }
  auto t0 = std::chrono::steady_clock::now();
  go.store(1);
  for (auto& t : threads) {
    t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}


// ---- codeword front end: pseudo-random sequence and UL-SCH demultiplexer of the reference ----------------------------

/// c(offset .. offset + n - 1) of pseudo_random_generator_impl, one bit per byte.
void ref_prg_bits(unsigned c_init, unsigned offset, unsigned n, uint8_t* bits)
{
  pseudo_random_generator_impl prg;
  prg.init(c_init);
  prg.advance(offset);
  dynamic_bit_buffer seq(n);
  prg.generate(seq);
  for (unsigned i = 0; i != n; ++i) {
    bits[i] = seq.extract(i, 1);
  }
}

/// tbs_calculator_calculate (lib/ran/sch/tbs_calculator.cpp:166-188) for a PUSCH MCS (pusch_mcs_get_config, table 0 = qam64,
/// 1 = qam256, no transform precoding) and get_ldpc_base_graph. out = {tbs bits, base graph (1 / 2), bits per symbol,
/// target code rate x 1024 x 2}.
void ref_pusch_tbs(int table, int mcs, int nof_symb_sh, int nof_dmrs_prb, int nof_oh_prb, int nof_layers, int n_prb, int* out)
{
  sch_mcs_description d =
      pusch_mcs_get_config(table == 0 ? pusch_mcs_table::qam64 : pusch_mcs_table::qam256, sch_mcs_index(mcs), false);
  tbs_calculator_configuration c = {};
  c.nof_symb_sh                  = nof_symb_sh;
  c.nof_dmrs_prb                 = nof_dmrs_prb;
  c.nof_oh_prb                   = nof_oh_prb;
  c.mcs_descr                    = d;
  c.nof_layers                   = nof_layers;
  c.tb_scaling_field             = 0;
  c.n_prb                        = n_prb;
  unsigned tbs                   = tbs_calculator_calculate(c);
  out[0]                         = static_cast<int>(tbs);
  out[1] = (get_ldpc_base_graph(d.get_normalised_target_code_rate(), units::bits(tbs)) == ldpc_base_graph_type::BG1) ? 1 : 2;
  out[2] = static_cast<int>(get_bits_per_symbol(d.modulation));
  out[3] = static_cast<int>(d.target_code_rate * 2.0F);
}

/// One demodulation_mapper::demodulate_soft call (include/srsran/phy/upper/channel_modulation/demodulation_mapper.h:62)
/// on the mapper of create_channel_modulation_sw_factory(). symbols: n interleaved {re, im} pairs.
/// mod: 0 = pi/2-BPSK, 1 = BPSK, 2 = QPSK, 4 = 16QAM, 6 = 64QAM, 8 = 256QAM.
void ref_demodulate_soft(int8_t* llr, const float* symbols, const float* noise_vars, unsigned n, int mod)
{
  static std::unique_ptr<demodulation_mapper> demapper =
      create_channel_modulation_sw_factory()->create_demodulation_mapper();
  modulation_scheme scheme = (mod == 0) ? modulation_scheme::PI_2_BPSK : to_mod(mod);
  unsigned          qm     = (mod == 0) ? 1 : mod;
  std::vector<log_likelihood_ratio> out(static_cast<size_t>(n) * qm);
  demapper->demodulate_soft(
      out, span<const cf_t>(reinterpret_cast<const cf_t*>(symbols), n), span<const float>(noise_vars, n), scheme);
  for (size_t i = 0; i != out.size(); ++i) {
    llr[i] = out[i].to_value_type();
  }
}

namespace {
/// Decoder buffer that records the soft bits it is given.
class spy_decoder_buffer : public pusch_decoder_buffer
{
public:
  std::vector<log_likelihood_ratio> data;
  std::vector<log_likelihood_ratio> scratch;
  bool                              ended = false;
  std::function<void()>             on_end;

  span<log_likelihood_ratio> get_next_block_view(unsigned block_size) override
  {
    scratch.resize(block_size);
    return scratch;
  }
  void on_new_softbits(span<const log_likelihood_ratio> softbits) override
  {
    data.insert(data.end(), softbits.begin(), softbits.end());
  }
  void on_end_softbits() override
  {
    ended = true;
    if (on_end) {
      on_end();
    }
  }
};
} // namespace

/// Drives ulsch_demultiplex_impl the way pusch_demodulator_impl does (pusch_demodulator_impl.cpp:160-283): per OFDM
/// symbol, blocks as large as get_next_block_view grants. cfg[15]: the fields of orc_ulsch_cfg in order.
/// max_block (> 0) additionally caps the block size in resource elements. Returns 0.
int ref_ulsch_demux(const int*     cfg,
                    const int8_t*  in,
                    const uint8_t* seq_bits,
                    unsigned       n_in,
                    unsigned       max_block_re,
                    int8_t*        sch,
                    int8_t*        harq_ack,
                    int8_t*        csi_part1,
                    int8_t*        csi_part2,
                    unsigned*      n_out)
{
  ulsch_demultiplex::configuration c;
  const int                        qm = cfg[0];
  c.modulation  = (qm == 1)   ? modulation_scheme::PI_2_BPSK
                  : (qm == 2) ? modulation_scheme::QPSK
                  : (qm == 4) ? modulation_scheme::QAM16
                  : (qm == 6) ? modulation_scheme::QAM64
                              : modulation_scheme::QAM256;
  c.nof_layers                  = cfg[1];
  c.nof_prb                     = cfg[2];
  c.start_symbol_index          = cfg[3];
  c.nof_symbols                 = cfg[4];
  c.nof_harq_ack_rvd            = cfg[5];
  c.dmrs                        = (cfg[6] == 1) ? dmrs_type::TYPE1 : dmrs_type::TYPE2;
  c.dmrs_symbol_mask            = symbol_slot_mask(14);
  for (unsigned l = 0; l != 14; ++l) {
    if ((cfg[7] >> l) & 1) {
      c.dmrs_symbol_mask.set(l);
    }
  }
  c.nof_cdm_groups_without_data = cfg[8];
  c.nof_harq_ack_bits           = cfg[9];
  c.nof_enc_harq_ack_bits       = cfg[10];
  c.nof_csi_part1_bits          = cfg[11];
  c.nof_enc_csi_part1_bits      = cfg[12];
  const unsigned csi2_bits = cfg[13], csi2_enc = cfg[14];

  auto               demux = std::make_unique<ulsch_demultiplex_impl>();
  spy_decoder_buffer b_sch, b_ack, b_csi1, b_csi2;
  b_csi1.on_end = [&]() {
    if (csi2_enc != 0) {
      demux->set_csi_part2(b_csi2, csi2_bits, csi2_enc);
    }
  };
  pusch_codeword_buffer& cw = demux->demultiplex(b_sch, b_ack, b_csi1, c);

  const unsigned bpre        = qm * c.nof_layers;
  const unsigned per_prb_dm  = c.nof_cdm_groups_without_data * ((cfg[6] == 1) ? 6 : 4);
  unsigned       pos         = 0;
  for (unsigned l = c.start_symbol_index; l != c.start_symbol_index + c.nof_symbols; ++l) {
    unsigned nof_re = c.dmrs_symbol_mask.test(l) ? (12 - per_prb_dm) * c.nof_prb : 12 * c.nof_prb;
    unsigned count  = 0;
    while (count != nof_re) {
      unsigned want = nof_re - count;
      if (max_block_re != 0) {
        want = std::min(want, max_block_re);
      }
      span<log_likelihood_ratio> view = cw.get_next_block_view(want * bpre);
      unsigned                   n    = view.size();
      if (n == 0) {
        // The reference's demodulator would spin here (an allocation whose FIRST symbol carries no data together with
        // UCI): not a case the reference supports.
        return -3;
      }
      for (unsigned i = 0; i != n; ++i) {
        view[i] = log_likelihood_ratio(in[pos + i]);
      }
      dynamic_bit_buffer seq(n);
      for (unsigned i = 0; i != n; ++i) {
        seq.insert(seq_bits[pos + i], i, 1);
      }
      cw.on_new_block(view, seq);
      pos += n;
      count += n / bpre;
    }
  }
  if (pos != n_in) {
    return -2;
  }
  cw.on_end_codeword();
  spy_decoder_buffer* bufs[4] = {&b_sch, &b_ack, &b_csi1, &b_csi2};
  int8_t*             outs[4] = {sch, harq_ack, csi_part1, csi_part2};
  for (int k = 0; k != 4; ++k) {
    n_out[k] = bufs[k]->data.size();
    for (unsigned i = 0; i != n_out[k]; ++i) {
      outs[k][i] = bufs[k]->data[i].to_value_type();
    }
  }
  return 0;
}

} // extern "C"
