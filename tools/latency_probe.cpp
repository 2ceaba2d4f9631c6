// Single-slot latency through the C ABI from C++ (what a gNB thread sees: no Python between the calls): pdc_submit +
// pdc_wait of one batch at a time, host clock from "soft bits in page-locked memory" to "transport-block bytes and CRC
// flags back in host memory" (SURVEY 8d). The batch (descriptors and soft bits) comes from files bench.py writes, so the
// transport blocks are valid codewords and every stage of the chain does its full work.
//
//   latency_probe <cbs.bin> <tbs.bin> <llrs.bin> <tb_bytes_per_batch> <slots> [device] [want_cb_bits]
//
// With transport blocks in the batch the codeblock hard bits are not brought back unless want_cb_bits is 1 (the transport
// block is assembled on the device; a gNB only needs its bytes and the flags).
//
// Prints one JSON object. Only include/pusch_dec_cuda.h is needed to build it.
#include "pusch_dec_cuda.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static std::vector<unsigned char> read_file(const char* path)
{
  std::vector<unsigned char> data;
  FILE*                      f = fopen(path, "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path);
    exit(2);
  }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  data.resize((size_t)n);
  if (n && fread(data.data(), 1, (size_t)n, f) != (size_t)n) {
    fprintf(stderr, "short read on %s\n", path);
    exit(2);
  }
  fclose(f);
  return data;
}

int main(int argc, char** argv)
{
  if (argc < 6) {
    fprintf(stderr, "usage: latency_probe cbs.bin tbs.bin llrs.bin tb_bytes slots [device]\n");
    return 2;
  }
  const std::vector<unsigned char> cb_raw = read_file(argv[1]), tb_raw = read_file(argv[2]), llr_raw = read_file(argv[3]);
  const size_t                     tb_bytes = (size_t)atol(argv[4]);
  const int                        slots    = atoi(argv[5]);
  const uint32_t n_cb = (uint32_t)(cb_raw.size() / sizeof(pdc_cb_desc)), n_tb = (uint32_t)(tb_raw.size() / sizeof(pdc_tb_desc));
  const size_t   n_llr = llr_raw.size();

  pdc_config cfg;
  pdc_default_config(&cfg);
  cfg.device       = (argc > 6) ? atoi(argv[6]) : 0;
  cfg.max_cbs      = n_cb;
  cfg.max_llrs     = (uint32_t)n_llr + 64;
  cfg.harq_entries = n_cb;
  cfg.max_tbs      = n_tb;
  cfg.max_tb_bytes = (uint32_t)tb_bytes + 64;
  cfg.nof_streams  = 1;
  pdc_ctx* ctx = nullptr;
  if (pdc_create(&cfg, &ctx) != PDC_OK) {
    fprintf(stderr, "pdc_create: %s\n", pdc_last_error());
    return 1;
  }
  int8_t*  llrs = static_cast<int8_t*>(pdc_host_alloc(n_llr + 64));
  uint8_t* bits = static_cast<uint8_t*>(pdc_host_alloc((size_t)n_cb * PDC_MAX_CB_BYTES));
  uint8_t* tbo  = static_cast<uint8_t*>(pdc_host_alloc(tb_bytes + 64));
  if (!llrs || !bits || !tbo) {
    fprintf(stderr, "pdc_host_alloc failed\n");
    return 1;
  }
  memcpy(llrs, llr_raw.data(), n_llr);
  std::vector<pdc_cb_result> cb_res(n_cb);
  std::vector<pdc_tb_result> tb_res(n_tb);
  const pdc_cb_desc*         cbs = reinterpret_cast<const pdc_cb_desc*>(cb_raw.data());
  const pdc_tb_desc*         tbs = reinterpret_cast<const pdc_tb_desc*>(tb_raw.data());

  const bool want_bits = (n_tb == 0) || (argc > 7 && atoi(argv[7]) != 0);
  std::vector<double> lat;
  lat.reserve((size_t)slots);
  bool all_ok = true;
  for (int i = 0; i != slots + 20; ++i) {
    const auto t0 = std::chrono::steady_clock::now();
    int rc = pdc_submit(ctx, 0, cbs, n_cb, llrs, n_llr, tbs, n_tb, cb_res.data(), want_bits ? bits : nullptr, tb_res.data(), tbo);
    if (rc == PDC_OK) {
      rc = pdc_wait(ctx, 0);
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (rc != PDC_OK) {
      fprintf(stderr, "pdc_submit / pdc_wait: %s\n", pdc_last_error());
      return 1;
    }
    if (i >= 20) {
      lat.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
    }
    for (uint32_t t = 0; t != n_tb; ++t) {
      all_ok = all_ok && tb_res[t].tb_crc_ok;
    }
  }
  std::sort(lat.begin(), lat.end());
  double mean = 0;
  for (double v : lat) {
    mean += v;
  }
  mean /= (double)lat.size();
  printf("{\"p50\": %.2f, \"p99\": %.2f, \"min\": %.2f, \"mean\": %.2f, \"slots\": %d, \"codeblocks\": %u, "
         "\"transport_blocks\": %u, \"cb_bits_copied_back\": %s, \"tb_crc_ok\": %s, \"path\": \"C++: pdc_submit + pdc_wait, page-locked host buffers\"}\n",
         lat[lat.size() / 2], lat[std::min(lat.size() - 1, (size_t)((double)lat.size() * 0.99))], lat.front(), mean, slots, n_cb,
         n_tb, want_bits ? "true" : "false", all_ok ? "true" : "false");
  pdc_host_free(llrs);
  pdc_host_free(bits);
  pdc_host_free(tbo);
  pdc_destroy(ctx);
  return 0;
}
