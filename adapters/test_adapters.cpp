// Drop-in check of the "cuda" adapters against the reference's own software objects, driven through the reference's
// own classes: the UNMODIFIED pusch_decoder_hw_impl (lib/phy/upper/channel_processors/pusch/pusch_decoder_hw_impl.cpp)
// runs on top of hw_accelerator_pusch_dec_cuda and is compared with the reference's software pusch_decoder_impl on the
// same LLRs, over HARQ retransmissions. Test infrastructure: links oracle/_ref/libsrsref.so (the compiled reference).
#include "lib/phy/upper/channel_processors/pdsch_encoder_hw_impl.h"
#include "lib/phy/upper/channel_processors/pdsch_encoder_impl.h"
#include "lib/phy/upper/channel_processors/pusch/pusch_codeblock_decoder.h"
#include "lib/phy/upper/channel_processors/pusch/pusch_decoder_hw_impl.h"
#include "lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.h"
#include "lib/phy/upper/channel_processors/pusch/ulsch_demultiplex_impl.h"
#include "pusch_dec_cuda_adapters.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include <cmath>
#include <functional>
#include <cstdio>
#include <random>
#include <thread>
#include <atomic>

using namespace srsran;

extern "C" int ref_tb_encode(const uint8_t* tb,
                             int            tb_bytes,
                             int            bg,
                             int            rv,
                             int            qm,
                             int            nref,
                             int            nof_layers,
                             int            nof_ch_symbols,
                             uint8_t*       cw_bits);

namespace {

class test_rx_buffer : public unique_rx_buffer::callback
{
public:
  test_rx_buffer(unsigned nof_cb, unsigned first_id) :
    soft(nof_cb, std::vector<log_likelihood_ratio>(ldpc::MAX_CODEBLOCK_SIZE)),
    data(nof_cb, std::vector<uint8_t>(ldpc::MAX_CODEBLOCK_SIZE / 8 + 8)),
    crcs(new bool[nof_cb]()),
    n(nof_cb),
    first(first_id)
  {
  }
  ~test_rx_buffer() override { delete[] crcs; }
  unsigned   get_nof_codeblocks() const override { return n; }
  void       reset_codeblocks_crc() override { std::fill(crcs, crcs + n, false); }
  span<bool> get_codeblocks_crc() override { return span<bool>(crcs, n); }
  unsigned   get_absolute_codeblock_id(unsigned cb) const override { return first + cb; }
  span<log_likelihood_ratio> get_codeblock_soft_bits(unsigned cb, unsigned size) override
  {
    return span<log_likelihood_ratio>(soft[cb]).first(size);
  }
  bit_buffer get_codeblock_data_bits(unsigned cb, unsigned size) override
  {
    return bit_buffer::from_bytes(span<uint8_t>(data[cb].data(), (size + 7) / 8)).first(size);
  }
  void lock() override {}
  void unlock() override {}
  void release() override {}

  std::vector<std::vector<log_likelihood_ratio>> soft;
  std::vector<std::vector<uint8_t>>              data;
  bool*                                          crcs;
  unsigned                                       n, first;
};

class spy : public pusch_decoder_notifier
{
public:
  void                 on_sch_data(const pusch_decoder_result& r) override { result = r; }
  pusch_decoder_result result;
};

modulation_scheme to_mod(int qm)
{
  return qm == 2   ? modulation_scheme::QPSK
         : qm == 4 ? modulation_scheme::QAM16
         : qm == 6 ? modulation_scheme::QAM64
                   : modulation_scheme::QAM256;
}


/// Decoder buffer that records what it is given; the CSI Part 1 one answers with set_csi_part2 like the PUSCH processor.
class recording_buffer : public pusch_decoder_buffer
{
public:
  std::vector<log_likelihood_ratio> data, scratch;
  bool                              ended = false;
  std::function<void()>             on_end;
  span<log_likelihood_ratio>        get_next_block_view(unsigned n) override
  {
    scratch.resize(n);
    return scratch;
  }
  void on_new_softbits(span<const log_likelihood_ratio> s) override { data.insert(data.end(), s.begin(), s.end()); }
  void on_end_softbits() override
  {
    ended = true;
    if (on_end) {
      on_end();
    }
  }
};

/// Feeds one codeword into a demultiplexer the way pusch_demodulator_impl does (:160-283) and records the four streams.
struct demux_run {
  recording_buffer sch, ack, csi1, csi2;
  void             run(ulsch_demultiplex&                      demux,
                       const ulsch_demultiplex::configuration& cfg,
                       unsigned                                csi2_bits,
                       unsigned                                csi2_enc,
                       const std::vector<int8_t>&              llr,
                       const std::vector<uint8_t>&             seq_bits)
  {
    csi1.on_end = [&]() {
      if (csi2_enc != 0) {
        demux.set_csi_part2(csi2, csi2_bits, csi2_enc);
      }
    };
    pusch_codeword_buffer& cw   = demux.demultiplex(sch, ack, csi1, cfg);
    const unsigned         bpre = get_bits_per_symbol(cfg.modulation) * cfg.nof_layers;
    const unsigned         re_dmrs =
        (12 - cfg.nof_cdm_groups_without_data * (cfg.dmrs == dmrs_type::TYPE1 ? 6 : 4)) * cfg.nof_prb;
    size_t pos = 0;
    for (unsigned l = cfg.start_symbol_index; l != cfg.start_symbol_index + cfg.nof_symbols; ++l) {
      unsigned nof_re = cfg.dmrs_symbol_mask.test(l) ? re_dmrs : 12 * cfg.nof_prb;
      unsigned done   = 0;
      while (done != nof_re) {
        span<log_likelihood_ratio> view = cw.get_next_block_view((nof_re - done) * bpre);
        unsigned                   n    = view.size();
        dynamic_bit_buffer         seq(n);
        for (unsigned i = 0; i != n; ++i) {
          view[i] = log_likelihood_ratio(llr[pos + i]);
          seq.insert(seq_bits[pos + i], i, 1);
        }
        cw.on_new_block(view, seq);
        pos += n;
        done += n / bpre;
      }
    }
    cw.on_end_codeword();
  }
};

int failures = 0;
#define CHECK(cond, ...)                                                                                               \
  do {                                                                                                                 \
    if (!(cond)) {                                                                                                     \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__);                                                                 \
      std::printf(__VA_ARGS__);                                                                                        \
      std::printf("\n");                                                                                               \
      ++failures;                                                                                                      \
    }                                                                                                                  \
  } while (0)

} // namespace

int main()
{
  auto ctx = cuda::context::create({});
  if (!ctx) {
    std::printf("no GPU context: %s\n", pdc_last_error());
    return 2;
  }
  std::mt19937 rng(12345);

  auto crc_f = create_crc_calculator_factory_sw("auto");
  // ---- single-codeblock adapters vs the reference software objects ------------------------------------------------------
  {
    auto dec_sw  = create_ldpc_decoder_factory_sw("auto")->create();
    auto dec_gpu = cuda::create_ldpc_decoder_factory_cuda(ctx)->create();
    auto dem_sw  = create_ldpc_rate_dematcher_factory_sw("auto")->create();
    auto dem_gpu = cuda::create_ldpc_rate_dematcher_factory_cuda(ctx)->create();
    auto crc_sw  = crc_f->create(crc_generator_poly::CRC24B);
    auto crc_gpu = cuda::create_crc_calculator_factory_cuda(ctx)->create(crc_generator_poly::CRC24B);
    for (int trial = 0; trial != 20; ++trial) {
      unsigned           Zs[] = {384, 96, 52, 13, 7, 256};
      unsigned           Z    = Zs[trial % 6];
      bool               bg1  = trial % 2;
      unsigned           N = (bg1 ? 66 : 50) * Z, K = (bg1 ? 22 : 10) * Z;
      unsigned           qm = 2 + 2 * (trial % 4);
      unsigned           E  = ((N / 2 + rng() % N) / qm) * qm;
      codeblock_metadata m;
      m.tb_common.base_graph        = bg1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
      m.tb_common.lifting_size      = static_cast<ldpc::lifting_size_t>(Z);
      m.tb_common.rv                = trial % 4;
      m.tb_common.mod               = to_mod(qm);
      m.tb_common.Nref              = 0;
      m.cb_specific.full_length     = N;
      m.cb_specific.rm_length       = E;
      m.cb_specific.nof_filler_bits = rng() % Z;
      m.cb_specific.nof_crc_bits    = 24;
      std::vector<log_likelihood_ratio> llr(E), a(N), b(N);
      for (auto& v : llr) {
        v = static_cast<int>(rng() % 61) - 30;
      }
      for (unsigned i = 0; i != N; ++i) {
        a[i] = b[i] = static_cast<int>(rng() % 201) - 100;
      }
      dem_sw->rate_dematch(a, llr, trial % 3 != 0, m);
      dem_gpu->rate_dematch(b, llr, trial % 3 != 0, m);
      CHECK(std::equal(a.begin(), a.end(), b.begin()), "rate_dematch differs (trial %d)", trial);

      ldpc_decoder::configuration cfg;
      cfg.block_conf                    = m;
      cfg.algorithm_conf.max_iterations = 1 + trial % 6;
      dynamic_bit_buffer o1(K), o2(K);
      auto               r1 = dec_sw->decode(o1, a, trial % 2 ? crc_sw.get() : nullptr, cfg);
      auto               r2 = dec_gpu->decode(o2, a, trial % 2 ? crc_gpu.get() : nullptr, cfg);
      CHECK(r1 == r2, "decode iterations differ (trial %d)", trial);
      CHECK(o1 == o2, "decoded bits differ (trial %d)", trial);
      CHECK(crc_sw->calculate(o1) == crc_gpu->calculate(o1), "crc differs (trial %d)", trial);
    }
  }

  // ---- the same objects from several threads at once -------------------------------------------------------------------
  // pusch_decoder_impl runs a pool of decoder / dematcher objects on a pool of worker threads (pusch_decoder_impl.cpp:
  // 309-382, concurrent_thread_local_object_pool.h:41-110): eight threads, each with its own adapter objects on the ONE
  // shared context, each comparing every result with its own reference software objects.
  {
    constexpr int      n_threads = 8;
    std::atomic<int>   bad{0}, done{0};
    std::vector<std::thread> pool;
    for (int t = 0; t != n_threads; ++t) {
      pool.emplace_back([&, t]() {
        std::mt19937 trng(777 + t);
        auto dec_sw  = create_ldpc_decoder_factory_sw("auto")->create();
        auto dec_gpu = cuda::create_ldpc_decoder_factory_cuda(ctx)->create();
        auto dem_sw  = create_ldpc_rate_dematcher_factory_sw("auto")->create();
        auto dem_gpu = cuda::create_ldpc_rate_dematcher_factory_cuda(ctx)->create();
        auto crc_sw  = crc_f->create(crc_generator_poly::CRC24B);
        auto crc_gpu = cuda::create_crc_calculator_factory_cuda(ctx)->create(crc_generator_poly::CRC24B);
        for (int trial = 0; trial != 12; ++trial) {
          unsigned           Zs[] = {384, 96, 52, 13, 7, 256};
          unsigned           Z    = Zs[(trial + t) % 6];
          bool               bg1  = (trial + t) % 2;
          unsigned           N = (bg1 ? 66 : 50) * Z, K = (bg1 ? 22 : 10) * Z;
          unsigned           qm = 2 + 2 * (trial % 4);
          unsigned           E  = ((N / 2 + trng() % N) / qm) * qm;
          codeblock_metadata m;
          m.tb_common.base_graph        = bg1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
          m.tb_common.lifting_size      = static_cast<ldpc::lifting_size_t>(Z);
          m.tb_common.rv                = trial % 4;
          m.tb_common.mod               = to_mod(qm);
          m.tb_common.Nref              = 0;
          m.cb_specific.full_length     = N;
          m.cb_specific.rm_length       = E;
          m.cb_specific.nof_filler_bits = trng() % Z;
          m.cb_specific.nof_crc_bits    = 24;
          std::vector<log_likelihood_ratio> llr(E), a(N), b(N);
          for (auto& v : llr) {
            v = static_cast<int>(trng() % 61) - 30;
          }
          for (unsigned i = 0; i != N; ++i) {
            a[i] = b[i] = static_cast<int>(trng() % 201) - 100;
          }
          dem_sw->rate_dematch(a, llr, trial % 3 != 0, m);
          dem_gpu->rate_dematch(b, llr, trial % 3 != 0, m);
          ldpc_decoder::configuration cfg;
          cfg.block_conf                    = m;
          cfg.algorithm_conf.max_iterations = 1 + trial % 6;
          dynamic_bit_buffer o1(K), o2(K);
          auto               r1 = dec_sw->decode(o1, a, trial % 2 ? crc_sw.get() : nullptr, cfg);
          auto               r2 = dec_gpu->decode(o2, a, trial % 2 ? crc_gpu.get() : nullptr, cfg);
          if (!std::equal(a.begin(), a.end(), b.begin()) || r1 != r2 || !(o1 == o2) ||
              crc_sw->calculate(o1) != crc_gpu->calculate(o1)) {
            ++bad;
          }
          ++done;
        }
      });
    }
    for (auto& th : pool) {
      th.join();
    }
    CHECK(bad == 0 && done == n_threads * 12, "concurrent single-codeblock adapters: %d of %d calls differ", bad.load(),
          done.load());
  }

  // ---- the reference's pusch_decoder_hw_impl on the CUDA accelerator vs the reference's software pusch_decoder_impl ----
  auto hw_factory = cuda::create_hw_accelerator_pusch_dec_factory_cuda(ctx);
  {
    // One batch queue per accelerator object, never shared; beyond the context's queues creation fails, and a destroyed
    // object hands its queue back.
    std::vector<std::unique_ptr<hal::hw_accelerator_pusch_dec>> accs;
    for (unsigned i = 0; i != ctx->nof_queues(); ++i) {
      accs.push_back(hw_factory->create());
      CHECK(accs.back() != nullptr, "accelerator %u of %u not created", i, ctx->nof_queues());
    }
    CHECK(hw_factory->create() == nullptr, "an accelerator beyond the context's queues was created");
    accs.pop_back();
    CHECK(hw_factory->create() != nullptr, "a released queue was not handed out again");
  }
  for (int trial = 0; trial != 12; ++trial) {
    int      bg       = (trial % 3) ? 1 : 2;
    int      tb_bytes = 40 + static_cast<int>(rng() % (bg == 1 ? 6000 : 700));
    int      qm       = 2 + 2 * (trial % 4);
    int      nl       = 1 + trial % 2;
    double   rate     = (bg == 1) ? 0.6 + 0.03 * (trial % 10) : 0.25 + 0.03 * (trial % 10);
    int      nsym     = static_cast<int>(std::ceil(tb_bytes * 8 / rate / qm / nl)) * nl;
    int      n_llr    = nsym * qm;
    unsigned C        = ldpc::compute_nof_codeblocks(units::bits(tb_bytes * 8),
                                              bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2);
    unsigned nref     = (trial % 2) ? ldpc::compute_N_ref(units::bytes(tb_bytes + 40), C).value() : 0;
    bool     early    = trial % 2;
    unsigned max_it   = 2 + trial % 5;
    double   snr_db   = ((bg == 1) ? 8.0 : 3.0) * rate / 0.8 - 1.0 - 0.25 * (trial % 8);
    double   sigma2   = std::pow(10.0, -snr_db / 10.0);

    std::vector<uint8_t> tb(tb_bytes);
    for (auto& v : tb) {
      v = rng() & 0xff;
    }

    // Software reference.
    std::vector<std::unique_ptr<pusch_codeblock_decoder>> cbd(1);
    pusch_codeblock_decoder::sch_crc                      c1{crc_f->create(crc_generator_poly::CRC16),
                                        crc_f->create(crc_generator_poly::CRC24A),
                                        crc_f->create(crc_generator_poly::CRC24B)};
    cbd[0]    = std::make_unique<pusch_codeblock_decoder>(create_ldpc_rate_dematcher_factory_sw("auto")->create(),
                                                       create_ldpc_decoder_factory_sw("auto")->create(),
                                                       c1);
    auto pool = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cbd));
    pusch_decoder_impl::sch_crc c2{crc_f->create(crc_generator_poly::CRC16),
                                   crc_f->create(crc_generator_poly::CRC24A),
                                   crc_f->create(crc_generator_poly::CRC24B)};
    pusch_decoder_impl sw(create_ldpc_segmenter_rx_factory_sw()->create(), pool, std::move(c2), nullptr, MAX_RB, 4);
    // Reference hardware-decoder front end on the CUDA accelerator.
    pusch_decoder_hw_impl::sch_crc c3{crc_f->create(crc_generator_poly::CRC16),
                                      crc_f->create(crc_generator_poly::CRC24A),
                                      crc_f->create(crc_generator_poly::CRC24B)};
    pusch_decoder_hw_impl hw(create_ldpc_segmenter_rx_factory_sw()->create(), c3, hw_factory->create());

    test_rx_buffer buf_sw(C, 0), buf_hw(C, 100 + 16 * trial);
    bool           done   = false;
    const int      rvs[4] = {0, 2, 3, 1};
    for (int t = 0; t != 4 && !done; ++t) {
      std::vector<uint8_t> cw(n_llr);
      ref_tb_encode(tb.data(), tb_bytes, bg, rvs[t], qm, nref, nl, nsym, cw.data());
      std::vector<log_likelihood_ratio> llrs(n_llr);
      std::normal_distribution<double>  noise(0.0, std::sqrt(sigma2));
      for (int i = 0; i != n_llr; ++i) {
        double y = (1.0 - 2.0 * cw[i]) + noise(rng);
        double l = std::round(12.0 * y / sigma2);
        llrs[i]  = static_cast<int>(std::max(-120.0, std::min(120.0, l)));
      }
      pusch_decoder::configuration cfg;
      cfg.base_graph          = bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
      cfg.rv                  = rvs[t];
      cfg.mod                 = to_mod(qm);
      cfg.Nref                = nref;
      cfg.nof_layers          = nl;
      cfg.nof_ldpc_iterations = max_it;
      cfg.use_early_stop      = early;
      cfg.new_data            = (t == 0);

      std::vector<uint8_t> out_sw(tb_bytes), out_hw(tb_bytes);
      spy                  n_sw, n_hw;
      {
        pusch_decoder_buffer& b = sw.new_data(out_sw, unique_rx_buffer(buf_sw), n_sw, cfg);
        b.on_new_softbits(llrs);
        b.on_end_softbits();
      }
      {
        pusch_decoder_buffer& b = hw.new_data(out_hw, unique_rx_buffer(buf_hw), n_hw, cfg);
        b.on_new_softbits(llrs);
        b.on_end_softbits();
      }
      const pusch_decoder_result &a = n_sw.result, &b = n_hw.result;
      CHECK(a.tb_crc_ok == b.tb_crc_ok, "tb_crc_ok differs (trial %d tx %d)", trial, t);
      CHECK(a.nof_codeblocks_total == b.nof_codeblocks_total, "nof_codeblocks differs (trial %d)", trial);
      CHECK(a.ldpc_decoder_stats.get_nof_observations() == b.ldpc_decoder_stats.get_nof_observations(),
            "observations differ (trial %d tx %d): %zu vs %zu",
            trial,
            t,
            a.ldpc_decoder_stats.get_nof_observations(),
            b.ldpc_decoder_stats.get_nof_observations());
      if (a.ldpc_decoder_stats.get_nof_observations() && b.ldpc_decoder_stats.get_nof_observations()) {
        CHECK(a.ldpc_decoder_stats.get_min() == b.ldpc_decoder_stats.get_min() &&
                  a.ldpc_decoder_stats.get_max() == b.ldpc_decoder_stats.get_max(),
              "iteration statistics differ (trial %d tx %d)",
              trial,
              t);
      }
      for (unsigned cb = 0; cb != C; ++cb) {
        CHECK(buf_sw.crcs[cb] == buf_hw.crcs[cb], "CB CRC flag differs (trial %d tx %d cb %u)", trial, t, cb);
      }
      if (a.tb_crc_ok) {
        CHECK(out_sw == out_hw && out_sw == tb, "TB bytes differ (trial %d tx %d)", trial, t);
        done = true;
      }
      std::printf("trial %2d tx %d: bg%d tbs=%d C=%u qm=%d es=%d it=%u -> crc sw=%d hw=%d obs=%zu\n",
                  trial,
                  t,
                  bg,
                  tb_bytes * 8,
                  C,
                  qm,
                  early,
                  max_it,
                  a.tb_crc_ok,
                  b.tb_crc_ok,
                  b.ldpc_decoder_stats.get_nof_observations());
    }
  }

  // ---- batched PUSCH decoder: several UEs per slot in ONE submission vs the software decoder, over HARQ retransmissions ---
  {
    struct ue_t {
      int                  bg, qm, nl, tb_bytes, nsym, n_llr;
      unsigned             C, nref, max_it;
      bool                 early, done = false;
      double               sigma2;
      std::vector<uint8_t> tb;
      std::unique_ptr<test_rx_buffer> buf_sw, buf_gpu;
      std::unique_ptr<pusch_decoder_impl> sw;
      std::unique_ptr<pusch_decoder>      gpu;
    };
    auto batch_ptr = std::make_shared<cuda::pusch_decoder_batch_cuda>(ctx, cuda::context::no_queue,
                                                                      create_ldpc_segmenter_rx_factory_sw()->create());
    cuda::pusch_decoder_batch_cuda& batch = *batch_ptr;
    auto dec_factory = cuda::create_pusch_decoder_factory_cuda(batch_ptr);
    std::vector<ue_t>              ues(7);
    unsigned                       next_id = 1200;
    for (size_t u = 0; u != ues.size(); ++u) {
      ue_t& e    = ues[u];
      e.bg       = (u % 3 == 2) ? 2 : 1;
      e.qm       = 2 + 2 * static_cast<int>(u % 4);
      e.nl       = 1 + static_cast<int>(u % 2);
      e.tb_bytes = (e.bg == 1) ? 600 + 2100 * static_cast<int>(u) : 40 + 90 * static_cast<int>(u);
      double rate = (e.bg == 1) ? 0.62 + 0.04 * (u % 5) : 0.3 + 0.04 * (u % 5);
      e.nsym      = static_cast<int>(std::ceil(e.tb_bytes * 8 / rate / e.qm / e.nl)) * e.nl;
      e.n_llr     = e.nsym * e.qm;
      e.C         = ldpc::compute_nof_codeblocks(units::bits(e.tb_bytes * 8),
                                         e.bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2);
      e.nref      = (u % 2) ? ldpc::compute_N_ref(units::bytes(e.tb_bytes + 40), e.C).value() : 0;
      e.early     = u % 2;
      e.max_it    = 3 + u % 4;
      double snr  = ((e.bg == 1) ? 8.0 : 3.0) * rate / 0.8 - 1.6 - 0.6 * (u % 4);
      e.sigma2    = std::pow(10.0, -snr / 10.0);
      e.tb.resize(e.tb_bytes);
      for (auto& v : e.tb) {
        v = rng() & 0xff;
      }
      e.buf_sw  = std::make_unique<test_rx_buffer>(e.C, 0);
      e.buf_gpu = std::make_unique<test_rx_buffer>(e.C, next_id);
      next_id += e.C;
      std::vector<std::unique_ptr<pusch_codeblock_decoder>> cbd(1);
      pusch_codeblock_decoder::sch_crc c1{crc_f->create(crc_generator_poly::CRC16),
                                          crc_f->create(crc_generator_poly::CRC24A),
                                          crc_f->create(crc_generator_poly::CRC24B)};
      cbd[0]    = std::make_unique<pusch_codeblock_decoder>(create_ldpc_rate_dematcher_factory_sw("auto")->create(),
                                                         create_ldpc_decoder_factory_sw("auto")->create(),
                                                         c1);
      auto pool = std::make_shared<pusch_decoder_impl::codeblock_decoder_pool>(std::move(cbd));
      pusch_decoder_impl::sch_crc c2{crc_f->create(crc_generator_poly::CRC16),
                                     crc_f->create(crc_generator_poly::CRC24A),
                                     crc_f->create(crc_generator_poly::CRC24B)};
      e.sw  = std::make_unique<pusch_decoder_impl>(create_ldpc_segmenter_rx_factory_sw()->create(), pool, std::move(c2),
                                                  nullptr, MAX_RB, 4);
      e.gpu = dec_factory->create();
    }
    const int rvs[4]     = {0, 2, 3, 1};
    unsigned  slots_done = 0, tbs_compared = 0;
    for (int t = 0; t != 4; ++t) {
      std::vector<std::vector<uint8_t>> out_sw(ues.size()), out_gpu(ues.size());
      std::vector<spy>                  n_sw(ues.size()), n_gpu(ues.size());
      std::vector<size_t>               active;
      for (size_t u = 0; u != ues.size(); ++u) {
        ue_t& e = ues[u];
        if (e.done) {
          continue;
        }
        active.push_back(u);
        std::vector<uint8_t> cw(e.n_llr);
        ref_tb_encode(e.tb.data(), e.tb_bytes, e.bg, rvs[t], e.qm, e.nref, e.nl, e.nsym, cw.data());
        std::vector<log_likelihood_ratio> llrs(e.n_llr);
        std::normal_distribution<double>  noise(0.0, std::sqrt(e.sigma2));
        for (int i = 0; i != e.n_llr; ++i) {
          double y = (1.0 - 2.0 * cw[i]) + noise(rng);
          double l = std::round(12.0 * y / e.sigma2);
          llrs[i]  = static_cast<int>(std::max(-120.0, std::min(120.0, l)));
        }
        pusch_decoder::configuration cfg;
        cfg.base_graph          = e.bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
        cfg.rv                  = rvs[t];
        cfg.mod                 = to_mod(e.qm);
        cfg.Nref                = e.nref;
        cfg.nof_layers          = e.nl;
        cfg.nof_ldpc_iterations = e.max_it;
        cfg.use_early_stop      = e.early;
        cfg.new_data            = (t == 0);
        out_sw[u].assign(e.tb_bytes, 0);
        out_gpu[u].assign(e.tb_bytes, 0);
        {
          pusch_decoder_buffer& b = e.sw->new_data(out_sw[u], unique_rx_buffer(*e.buf_sw), n_sw[u], cfg);
          b.on_new_softbits(llrs);
          b.on_end_softbits();
        }
        {
          // Soft bits arrive in two blocks through the view the decoder hands out, like from the demodulator.
          pusch_decoder_buffer& b = e.gpu->new_data(out_gpu[u], unique_rx_buffer(*e.buf_gpu), n_gpu[u], cfg);
          e.gpu->set_nof_softbits(units::bits(e.n_llr));
          const unsigned             first = (e.n_llr / 2 / e.qm) * e.qm;
          span<log_likelihood_ratio> v     = b.get_next_block_view(first);
          std::copy(llrs.begin(), llrs.begin() + first, v.begin());
          b.on_new_softbits(v);
          b.on_new_softbits(span<const log_likelihood_ratio>(llrs).last(e.n_llr - first));
          b.on_end_softbits();
        }
      }
      if (active.empty()) {
        break;
      }
      CHECK(batch.pending() == active.size(), "batched decoder: %u transport blocks queued, %zu expected", batch.pending(),
            active.size());
      CHECK(batch.flush(), "batched decoder: flush failed: %s", pdc_last_error());
      ++slots_done;
      for (size_t u : active) {
        ue_t&                       e = ues[u];
        const pusch_decoder_result &a = n_sw[u].result, &b = n_gpu[u].result;
        CHECK(a.tb_crc_ok == b.tb_crc_ok, "batched decoder: tb_crc_ok differs (ue %zu tx %d)", u, t);
        CHECK(a.nof_codeblocks_total == b.nof_codeblocks_total, "batched decoder: nof_codeblocks differs (ue %zu)", u);
        CHECK(a.ldpc_decoder_stats.get_nof_observations() == b.ldpc_decoder_stats.get_nof_observations(),
              "batched decoder: observations differ (ue %zu tx %d)", u, t);
        if (a.ldpc_decoder_stats.get_nof_observations() && b.ldpc_decoder_stats.get_nof_observations()) {
          CHECK(a.ldpc_decoder_stats.get_min() == b.ldpc_decoder_stats.get_min() &&
                    a.ldpc_decoder_stats.get_max() == b.ldpc_decoder_stats.get_max() &&
                    a.ldpc_decoder_stats.get_mean() == b.ldpc_decoder_stats.get_mean(),
                "batched decoder: iteration statistics differ (ue %zu tx %d)", u, t);
        }
        for (unsigned cb = 0; cb != e.C; ++cb) {
          CHECK(e.buf_sw->crcs[cb] == e.buf_gpu->crcs[cb], "batched decoder: CB CRC flag differs (ue %zu tx %d cb %u)", u,
                t, cb);
        }
        if (a.tb_crc_ok) {
          CHECK(out_sw[u] == out_gpu[u] && out_sw[u] == e.tb, "batched decoder: TB bytes differ (ue %zu tx %d)", u, t);
          e.done = true;
        }
        ++tbs_compared;
      }
    }
    std::printf("pusch_decoder_batch_cuda: %u slots, %u transport blocks compared with pusch_decoder_impl\n", slots_done,
                tbs_compared);
    CHECK(tbs_compared >= 7 && slots_done >= 2, "batched decoder: too little was compared");
  }

  // ---- ulsch_demultiplex_cuda vs the reference's ulsch_demultiplex_impl ---------------------------------------------------
  {
    auto     demux_gpu = cuda::create_ulsch_demultiplex_cuda(ctx);
    unsigned checked   = 0;
    for (int trial = 0; trial != 400 && checked != 60; ++trial) {
      const unsigned qms[] = {2, 4, 6, 8};
      const unsigned qm = qms[rng() % 4], nl = 1 + rng() % 4, nprb = 1 + rng() % 30;
      ulsch_demultiplex::configuration cfg;
      cfg.modulation                  = to_mod(qm);
      cfg.nof_layers                  = nl;
      cfg.nof_prb                     = nprb;
      cfg.start_symbol_index          = rng() % 3;
      cfg.nof_symbols                 = 6 + rng() % (9 - cfg.start_symbol_index);
      cfg.dmrs                        = (rng() % 2) ? dmrs_type::TYPE1 : dmrs_type::TYPE2;
      cfg.nof_cdm_groups_without_data = 1 + rng() % 2;
      cfg.dmrs_symbol_mask            = symbol_slot_mask(14);
      cfg.dmrs_symbol_mask.set(cfg.start_symbol_index + 1 + rng() % 3); // never the first symbol of the allocation
      const unsigned bpre   = qm * nl;
      const unsigned nre    = cfg.nof_prb * 12 * (cfg.nof_symbols - 1);
      const unsigned acks[] = {0, 1, 2, 5, 20};
      cfg.nof_harq_ack_bits = acks[rng() % 5];
      cfg.nof_enc_harq_ack_bits = cfg.nof_harq_ack_bits ? (1 + rng() % (nre / 8 + 1)) * bpre : 0;
      cfg.nof_harq_ack_rvd      = (cfg.nof_harq_ack_bits <= 2) ? cfg.nof_enc_harq_ack_bits + (rng() % 4) * bpre : 0;
      cfg.nof_csi_part1_bits    = (rng() % 2) ? 1 + rng() % 20 : 0;
      cfg.nof_enc_csi_part1_bits = cfg.nof_csi_part1_bits ? (1 + rng() % (nre / 8 + 1)) * bpre : 0;
      unsigned csi2_bits = 0, csi2_enc = 0;
      if (cfg.nof_csi_part1_bits != 0 && rng() % 2) {
        csi2_bits = 1 + rng() % 9;
        csi2_enc  = (1 + rng() % (nre / 8 + 1)) * bpre;
      }
      // Codeword length.
      const unsigned re_dmrs =
          (12 - cfg.nof_cdm_groups_without_data * (cfg.dmrs == dmrs_type::TYPE1 ? 6 : 4)) * cfg.nof_prb;
      size_t total = 0;
      for (unsigned l = cfg.start_symbol_index; l != cfg.start_symbol_index + cfg.nof_symbols; ++l) {
        total += (cfg.dmrs_symbol_mask.test(l) ? re_dmrs : 12 * cfg.nof_prb) * bpre;
      }
      std::vector<int8_t>  llr(total);
      std::vector<uint8_t> seq(total);
      for (size_t i = 0; i != total; ++i) {
        llr[i] = static_cast<int8_t>(static_cast<int>(rng() % 241) - 120);
        seq[i] = rng() & 1;
      }
      // Only descriptions in which all UCI fits (the reference asserts otherwise; the GPU path reports an error).
      {
        pdc_cw_desc d                 = {};
        d.qm                          = qm;
        d.nof_layers                  = nl;
        d.start_symbol_index          = cfg.start_symbol_index;
        d.nof_symbols                 = cfg.nof_symbols;
        d.dmrs_type                   = (cfg.dmrs == dmrs_type::TYPE1) ? 1 : 2;
        d.nof_cdm_groups_without_data = cfg.nof_cdm_groups_without_data;
        d.nof_prb                     = cfg.nof_prb;
        for (unsigned l = 0; l != 14; ++l) {
          d.dmrs_symbol_mask |= cfg.dmrs_symbol_mask.test(l) ? (1u << l) : 0u;
        }
        d.nof_harq_ack_rvd       = cfg.nof_harq_ack_rvd;
        d.nof_harq_ack_bits      = cfg.nof_harq_ack_bits;
        d.nof_enc_harq_ack_bits  = cfg.nof_enc_harq_ack_bits;
        d.nof_csi_part1_bits     = cfg.nof_csi_part1_bits;
        d.nof_enc_csi_part1_bits = cfg.nof_enc_csi_part1_bits;
        d.nof_csi_part2_bits     = csi2_bits;
        d.nof_enc_csi_part2_bits = csi2_enc;
        std::vector<int8_t> o1(total + 16), o2(total + 16);
        pdc_cw_result       r;
        if (pdc_ulsch_demux(ctx->get(), &d, 1, llr.data(), total, nullptr, o1.data(), o1.size(), o2.data(), o2.size(),
                            &r) != PDC_OK) {
          continue;
        }
      }
      // Value-initialised like the reference's factory does (std::make_unique): the class leaves softbit_count
      // without an initialiser.
      auto      demux_sw = std::make_unique<ulsch_demultiplex_impl>();
      demux_run a, b;
      a.run(*demux_sw, cfg, csi2_bits, csi2_enc, llr, seq);
      b.run(*demux_gpu, cfg, csi2_bits, csi2_enc, llr, seq);
      auto same = [](const recording_buffer& x, const recording_buffer& y) {
        return x.data.size() == y.data.size() && x.ended == y.ended &&
               std::equal(x.data.begin(), x.data.end(), y.data.begin());
      };
      CHECK(same(a.sch, b.sch), "ulsch demux: SCH stream differs (trial %d, %zu vs %zu)", trial, a.sch.data.size(),
            b.sch.data.size());
      CHECK(same(a.ack, b.ack), "ulsch demux: HARQ-ACK stream differs (trial %d)", trial);
      CHECK(same(a.csi1, b.csi1), "ulsch demux: CSI Part 1 stream differs (trial %d)", trial);
      CHECK(same(a.csi2, b.csi2), "ulsch demux: CSI Part 2 stream differs (trial %d)", trial);
      ++checked;
    }
    std::printf("ulsch_demultiplex_cuda: %u configurations compared with ulsch_demultiplex_impl\n", checked);
    CHECK(checked >= 30, "too few UL-SCH demultiplexing configurations were checked");
  }
  // ---- soft demapper: demodulation_mapper_cuda vs demodulation_mapper_impl ------------------------------------------------
  {
    auto dm_sw  = create_channel_modulation_sw_factory()->create_demodulation_mapper();
    auto dm_gpu = cuda::create_channel_modulation_cuda_factory(ctx)->create_demodulation_mapper();
    const modulation_scheme mods[] = {modulation_scheme::PI_2_BPSK, modulation_scheme::BPSK, modulation_scheme::QPSK,
                                      modulation_scheme::QAM16, modulation_scheme::QAM64, modulation_scheme::QAM256};
    std::normal_distribution<float>       gauss(0.0F, 0.8F);
    std::uniform_real_distribution<float> uni(0.0F, 1.0F);
    unsigned                              checked = 0;
    for (int trial = 0; trial != 120; ++trial) {
      modulation_scheme mod = mods[trial % 6];
      unsigned          n   = 1 + rng() % 4000;
      unsigned          qm  = get_bits_per_symbol(mod);
      std::vector<cf_t>  sym(n);
      std::vector<float> nv(n);
      for (unsigned i = 0; i != n; ++i) {
        sym[i] = cf_t(gauss(rng), gauss(rng));
        nv[i]  = 0.001F + 0.1F * uni(rng);
        if (trial % 5 == 4 && rng() % 16 == 0) {
          // ill-formed inputs: zero / negative / NaN / infinite noise variances, tiny and huge symbols
          const float bad[] = {0.0F, -1.0F, NAN, INFINITY, 1e-30F};
          nv[i]             = bad[rng() % 5];
          if (rng() % 2) {
            sym[i] = cf_t(1e-10F * gauss(rng), (rng() % 2) ? 1e20F : 0.0F);
          }
        }
      }
      std::vector<log_likelihood_ratio> a(n * qm), b(n * qm);
      dm_sw->demodulate_soft(a, sym, nv, mod);
      dm_gpu->demodulate_soft(b, sym, nv, mod);
      CHECK(std::equal(a.begin(), a.end(), b.begin()), "demodulation mapper: soft bits differ (trial %d, %s, %u symbols)",
            trial, to_string(mod).c_str(), n);
      ++checked;
    }
    std::printf("demodulation_mapper_cuda: %u demodulate_soft calls compared with demodulation_mapper_impl\n", checked);
  }
  // ---- downlink twin: ldpc_encoder_cuda vs the reference's encoder, and pdc_encode vs encoder + rate matcher -----------------
  {
    auto enc_sw  = create_ldpc_encoder_factory_sw("auto")->create();
    auto enc_gpu = cuda::create_ldpc_encoder_factory_cuda(ctx)->create();
    auto rm_sw   = create_ldpc_rate_matcher_factory_sw()->create();
    unsigned checked = 0;
    for (int trial = 0; trial != 40; ++trial) {
      const unsigned Zs[] = {384, 352, 208, 96, 52, 30, 13, 7, 2, 256};
      const unsigned Z = Zs[trial % 10], bg = 1 + (trial % 2);
      const unsigned K = (bg == 1 ? 22 : 10) * Z, N = (bg == 1 ? 66 : 50) * Z;
      const unsigned F = (trial % 3 == 0) ? rng() % Z : 0;
      const unsigned qm = 2 * (1 + trial % 4);
      codeblock_metadata m;
      m.tb_common.base_graph        = bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
      m.tb_common.lifting_size      = static_cast<ldpc::lifting_size_t>(Z);
      m.tb_common.rv                = trial % 4;
      m.tb_common.mod               = to_mod(qm);
      m.tb_common.Nref              = (trial % 5 == 0) ? (N - rng() % (N / 4)) : 0;
      m.cb_specific.full_length     = N;
      m.cb_specific.rm_length       = ((N / 3 + rng() % (2 * N)) / qm) * qm;
      m.cb_specific.nof_filler_bits = F;
      dynamic_bit_buffer msg(K), cw_sw(N), cw_gpu(N);
      std::vector<uint8_t> packed((K + 7) / 8, 0);
      for (unsigned i = 0; i != K; ++i) {
        const uint8_t b = (i < K - F) ? (rng() & 1) : 0;
        msg.insert(b, i, 1);
        packed[i >> 3] |= static_cast<uint8_t>(b << (7 - (i & 7)));
      }
      enc_sw->encode(cw_sw, msg, m.tb_common);
      enc_gpu->encode(cw_gpu, msg, m.tb_common);
      bool same = true;
      for (unsigned i = 0; i != N; ++i) {
        // filler positions of the systematic part are "filler bit" markers in the reference; compare what is transmitted
        if (i >= K - 2 * Z - F && i < K - 2 * Z) {
          continue;
        }
        same = same && (cw_sw.extract(i, 1) == cw_gpu.extract(i, 1));
      }
      CHECK(same, "ldpc encoder: codeblock differs (trial %d bg%u Z=%u)", trial, bg, Z);
      // encoder + rate matcher of the reference vs the fused device kernel
      const unsigned     E = m.cb_specific.rm_length;
      dynamic_bit_buffer rm_out(E);
      rm_sw->rate_match(rm_out, cw_sw, m);
      pdc_enc_desc d = {};
      d.rm_length = E, d.nref = m.tb_common.Nref, d.lifting_size = static_cast<uint16_t>(Z);
      d.nof_filler = static_cast<uint16_t>(F), d.base_graph = static_cast<uint8_t>(bg), d.qm = static_cast<uint8_t>(qm);
      d.rv = static_cast<uint8_t>(m.tb_common.rv);
      std::vector<uint8_t> out(E);
      CHECK(pdc_encode(ctx->get(), &d, 1, packed.data(), packed.size(), out.data(), out.size()) == PDC_OK,
            "pdc_encode failed: %s", pdc_last_error());
      bool same_rm = true;
      for (unsigned i = 0; i != E; ++i) {
        same_rm = same_rm && (rm_out.extract(i, 1) == out[i]);
      }
      CHECK(same_rm, "encoder + rate matcher: bits differ (trial %d bg%u Z=%u E=%u rv=%u qm=%u F=%u)", trial, bg, Z, E,
            m.tb_common.rv, qm, F);
      ++checked;
    }
    std::printf("ldpc_encoder_cuda / pdc_encode: %u codeblocks compared with the reference's encoder and rate matcher\n",
                checked);
  }
  // ---- downlink twin: the reference's pdsch_encoder_hw_impl, unchanged, on the CUDA accelerator vs pdsch_encoder_impl --------
  {
    pdsch_encoder_impl sw(create_ldpc_segmenter_tx_factory_sw(crc_f)->create(),
                          create_ldpc_encoder_factory_sw("auto")->create(),
                          create_ldpc_rate_matcher_factory_sw()->create());
    pdsch_encoder_hw_impl::sch_crc c4{crc_f->create(crc_generator_poly::CRC16), crc_f->create(crc_generator_poly::CRC24A),
                                      crc_f->create(crc_generator_poly::CRC24B)};
    pdsch_encoder_hw_impl hw(c4, create_ldpc_segmenter_tx_factory_sw(crc_f)->create(),
                             cuda::create_hw_accelerator_pdsch_enc_factory_cuda(ctx)->create());
    unsigned checked = 0;
    for (int trial = 0; trial != 24; ++trial) {
      const int      bg       = (trial % 3 == 2) ? 2 : 1;
      const unsigned qm       = 2 * (1 + trial % 4), nl = 1 + trial % 4;
      const unsigned tb_bytes = (bg == 1) ? 500 + 3700 * (trial % 7) : 30 + 60 * (trial % 7);
      const double   rate     = (bg == 1) ? 0.5 + 0.04 * (trial % 10) : 0.2 + 0.04 * (trial % 10);
      const unsigned nsym     = static_cast<unsigned>(std::ceil(tb_bytes * 8 / rate / qm / nl)) * nl;
      const unsigned C        = ldpc::compute_nof_codeblocks(units::bits(tb_bytes * 8),
                                                      bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2);
      pdsch_encoder::configuration cfg;
      cfg.base_graph     = bg == 1 ? ldpc_base_graph_type::BG1 : ldpc_base_graph_type::BG2;
      cfg.rv             = trial % 4;
      cfg.mod            = to_mod(qm);
      cfg.Nref           = (trial % 2) ? ldpc::compute_N_ref(units::bytes(tb_bytes + 50), C).value() : 0;
      cfg.nof_layers     = nl;
      cfg.nof_ch_symbols = nsym;
      std::vector<uint8_t> tb(tb_bytes), cw_sw(nsym * qm), cw_hw(nsym * qm);
      for (auto& v : tb) {
        v = rng() & 0xff;
      }
      sw.encode(cw_sw, tb, cfg);
      hw.encode(cw_hw, tb, cfg);
      CHECK(cw_sw == cw_hw, "pdsch encoder: codeword differs (trial %d bg%d tbs=%u C=%u qm=%u nl=%u rv=%u)", trial, bg,
            tb_bytes * 8, C, qm, nl, cfg.rv);
      ++checked;
    }
    std::printf("hw_accelerator_pdsch_enc_cuda: %u transport blocks through pdsch_encoder_hw_impl compared with "
                "pdsch_encoder_impl\n", checked);
  }
  std::printf(failures ? "FAILED: %d checks\n" : "PASS (%d failures)\n", failures);
  return failures ? 1 : 0;
}
