// Drop-in check of the "cuda" adapters against the reference's own software objects, driven through the reference's
// own classes: the UNMODIFIED pusch_decoder_hw_impl (lib/phy/upper/channel_processors/pusch/pusch_decoder_hw_impl.cpp)
// runs on top of hw_accelerator_pusch_dec_cuda and is compared with the reference's software pusch_decoder_impl on the
// same LLRs, over HARQ retransmissions. Test infrastructure: links oracle/_ref/libsrsref.so (the compiled reference).
#include "lib/phy/upper/channel_processors/pusch/pusch_codeblock_decoder.h"
#include "lib/phy/upper/channel_processors/pusch/pusch_decoder_hw_impl.h"
#include "lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.h"
#include "pusch_dec_cuda_adapters.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_notifier.h"
#include "srsran/phy/upper/channel_processors/pusch/pusch_decoder_result.h"
#include "srsran/phy/upper/unique_rx_buffer.h"
#include <cmath>
#include <cstdio>
#include <random>

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

  // ---- the reference's pusch_decoder_hw_impl on the CUDA accelerator vs the reference's software pusch_decoder_impl ----
  auto hw_factory = cuda::create_hw_accelerator_pusch_dec_factory_cuda(ctx);
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
  std::printf(failures ? "FAILED: %d checks\n" : "PASS (%d failures)\n", failures);
  return failures ? 1 : 0;
}
