// CPU check of the host-side UL-SCH demultiplexing plan (srsran_edgeric_5g_b200/csrc/ulsch_plan.h): executes the plan
// with the same index arithmetic as the device gathers, so that tests/test_host_logic.py can compare the streams with
// the oracle without a GPU. Test infrastructure; built by the test with g++.
#include "../../srsran_edgeric_5g_b200/csrc/ulsch_plan.h"
#include <cstring>

using namespace pdc;

static int find_symbol(const UlschSymbol* S, int n_sym, int k, uint32_t o)
{
  int s = 0;
  for (int i = 0; i != n_sym; ++i) {
    if (S[i].n_out_re[k] != 0 && S[i].out_off[k] <= o) {
      s = i;
    }
  }
  return s;
}

extern "C" int plan_demux_cpu(const pdc_cw_desc* d, const int8_t* in, const uint8_t* seq /* one bit per byte */,
                              int8_t* sch, int8_t* uci, uint32_t* n_out)
{
  UlschPlan plan;
  if (!ulsch_plan_codeword(*d, plan)) {
    return -1;
  }
  const UlschCodeword& cw = plan.cws[0];
  const UlschSymbol*   S  = plan.syms.data() + cw.sym_first;
  for (int k = 0; k != 4; ++k) {
    n_out[k] = cw.n_out[k];
  }
  for (uint32_t o = 0; o != cw.n_out[0]; ++o) {
    const UlschSymbol& sym = S[find_symbol(S, (int)cw.n_sym, 0, o)];
    uint32_t           rel = o - sym.out_off[0], r = rel / cw.bpre, q = rel % cw.bpre, re = r;
    if (sym.list_off[0] != ULSCH_IDENTITY) {
      uint16_t e = plan.lists[sym.list_off[0] + r];
      if (e & ULSCH_PUNCTURED) {
        sch[o] = 0;
        continue;
      }
      re = e;
    }
    sch[o] = in[sym.in_off + re * cw.bpre + q];
  }
  const uint32_t total = cw.n_out[1] + cw.n_out[2] + cw.n_out[3];
  for (uint32_t t = 0; t != total; ++t) {
    const int          k   = (t < cw.n_out[1]) ? 1 : (t < cw.n_out[1] + cw.n_out[2]) ? 2 : 3;
    const uint32_t     o   = t - cw.uci_base[k];
    const UlschSymbol& sym = S[find_symbol(S, (int)cw.n_sym, k, o)];
    const uint32_t     rel = o - sym.out_off[k], r = rel / cw.bpre, q = rel % cw.bpre;
    const uint32_t     e   = plan.lists[sym.list_off[k] + r];
    const uint32_t     re  = e & 0x7fffu;
    const uint32_t     i   = sym.in_off + re * cw.bpre + q;
    int                v   = (e & ULSCH_PUNCTURED) ? 0 : in[i];
    const uint32_t     nb  = sym.uci_bits[k];
    if ((nb == 1 || nb == 2) && cw.qm > 1) {
      const uint32_t bq   = q % cw.qm;
      bool           flip = false;
      if (bq == 1) {
        flip = nb == 1 && seq[i - 1] != seq[i];
      } else if (bq >= 2) {
        flip = seq[i] != 0;
      }
      if (flip) {
        v = (int8_t)(uint8_t)(0u - (unsigned)v);
      }
    }
    uci[t] = (int8_t)v;
  }
  return 0;
}
