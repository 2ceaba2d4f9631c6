// Host-side plan of the UL-SCH demultiplexing of one PUSCH codeword (TS 38.212 6.2.7).
//
// Mirrors what ulsch_demultiplex_impl derives symbol by symbol (configure_current_ofdm_symbol /
// configure_csi_part2_current_ofdm_symbol, lib/phy/upper/channel_processors/pusch/ulsch_demultiplex_impl.cpp:371-499):
// for every OFDM symbol of the allocation, which resource elements go to UL-SCH, HARQ-ACK, CSI Part 1 and CSI Part 2.
// The result is a table the device kernels index: per symbol the input offset, the offset of its first soft bit in each
// output stream, and - only for symbols that carry UCI - the ordered list of resource elements of each stream.
// CSI Part 2 is planned from the sizes the caller supplies; it takes effect in the symbol in which CSI Part 1 ends, as it
// does when the PUSCH processor calls set_csi_part2 from the CSI Part 1 decoder (pusch_processor_impl.cpp:61-82).
#pragma once

#include "../../include/pusch_dec_cuda.h"
#include <cstdint>
#include <vector>

namespace pdc {

constexpr uint32_t ULSCH_IDENTITY  = 0xffffffffu; // list offset: every resource element of the symbol, in order
constexpr uint16_t ULSCH_PUNCTURED = 0x8000u;     // list entry: the soft bits of this element read as zero (taken by HARQ-ACK)
constexpr int      ULSCH_STREAMS   = 4;           // UL-SCH, HARQ-ACK, CSI Part 1, CSI Part 2
constexpr uint32_t ULSCH_CW_DEFERRED = 0x100u;    // internal codeword flag: UL-SCH not materialised, the dematcher descrambles

// One OFDM symbol that carries soft bits. Offsets are in soft bits relative to the codeword / to each output stream.
struct UlschSymbol {
  uint32_t in_off;
  uint32_t n_re;
  uint32_t out_off[ULSCH_STREAMS];
  uint32_t n_out_re[ULSCH_STREAMS];
  uint32_t list_off[ULSCH_STREAMS]; // into the uint16 list pool, or ULSCH_IDENTITY (UL-SCH only)
  uint32_t uci_bits[ULSCH_STREAMS]; // number of UCI message bits of the stream (1 or 2: placeholder rules apply)
};

// One codeword as the kernels see it.
struct UlschCodeword {
  uint32_t in_off, sch_off, uci_off; // absolute offsets in the raw input, the UL-SCH space and the UCI output
  uint32_t c_init, flags;
  uint32_t qm, bpre;
  uint32_t sym_first, n_sym;
  uint32_t n_in;
  uint32_t n_out[ULSCH_STREAMS];  // soft bits of each stream
  uint32_t uci_base[ULSCH_STREAMS]; // offset of each UCI stream inside the codeword's UCI area ([0] unused)
  uint32_t seq_word_off;          // first 32-bit word of this codeword's scrambling sequence
  uint32_t prg_offset;            // sequence element of the codeword's first soft bit (0 for a PUSCH codeword)
};

struct UlschPlan {
  std::vector<UlschCodeword> cws;
  std::vector<UlschSymbol>   syms;
  std::vector<uint16_t>      lists;
  uint32_t                   seq_words = 0;
  void                       clear()
  {
    cws.clear();
    syms.clear();
    lists.clear();
    seq_words = 0;
  }
};

namespace ulsch_detail {

typedef std::vector<uint8_t> ReSet;

inline int count(const ReSet& s)
{
  int c = 0;
  for (uint8_t v : s) {
    c += v;
  }
  return c;
}

// The first m elements of src, taking one out of every d.
inline void select(ReSet& dst, const ReSet& src, int d, int m)
{
  dst.assign(src.size(), 0);
  int taken = 0, seen = 0;
  for (size_t i = 0; i != src.size() && taken != m; ++i) {
    if (!src[i]) {
      continue;
    }
    if (seen % d == 0) {
      dst[i] = 1;
      ++taken;
    }
    ++seen;
  }
}

inline void remove(ReSet& from, const ReSet& what)
{
  for (size_t i = 0; i != from.size(); ++i) {
    if (what[i]) {
      from[i] = 0;
    }
  }
}

// d and m of "spread `remaining` elements over `available` ones".
inline void spread(int remaining, int available, int& d, int& m)
{
  d = 1;
  m = available;
  if (remaining < available) {
    d = available / remaining;
    m = remaining;
  }
}

} // namespace ulsch_detail

// Appends the plan of one codeword. Returns false if the description is inconsistent (no DM-RS symbol, UCI that does not
// fit the allocation, sizes that are not multiples of the bits per resource element, ...).
inline bool ulsch_plan_codeword(const pdc_cw_desc& d, UlschPlan& plan)
{
  using namespace ulsch_detail;
  const int qm = d.qm, bpre = d.qm * d.nof_layers;
  if (!(qm == 1 || qm == 2 || qm == 4 || qm == 6 || qm == 8) || d.nof_layers < 1 || d.nof_layers > 4 || d.nof_prb == 0 ||
      d.nof_prb > 275 || d.nof_symbols == 0 || d.start_symbol_index + d.nof_symbols > 14 ||
      (d.dmrs_type != 1 && d.dmrs_type != 2) || d.nof_cdm_groups_without_data < 1 ||
      d.nof_cdm_groups_without_data > (d.dmrs_type == 1 ? 2 : 3)) {
    return false;
  }
  const int mask = d.dmrs_symbol_mask & 0x3fff;
  // l1: first symbol without DM-RS after the first DM-RS symbol; l1_csi: first symbol without DM-RS.
  int first_dmrs = -1, l1 = -1, l1_csi = -1;
  for (int l = 0; l != 14; ++l) {
    if ((mask >> l) & 1) {
      if (first_dmrs < 0) {
        first_dmrs = l;
      }
    } else {
      if (l1_csi < 0) {
        l1_csi = l;
      }
      if (first_dmrs >= 0 && l1 < 0) {
        l1 = l;
      }
    }
  }
  if (first_dmrs < 0 || l1 < 0 || l1_csi < 0) {
    return false;
  }
  const int re_dmrs_symbol = (12 - d.nof_cdm_groups_without_data * (d.dmrs_type == 1 ? 6 : 4)) * d.nof_prb;

  UlschCodeword cw = {};
  cw.in_off        = d.in_offset;
  cw.sch_off       = d.sch_offset;
  cw.uci_off       = d.uci_offset;
  cw.c_init        = d.c_init;
  cw.flags         = d.flags;
  cw.qm            = (uint32_t)qm;
  cw.bpre          = (uint32_t)bpre;
  cw.sym_first     = (uint32_t)plan.syms.size();

  unsigned m_rvd = 0, m_ack = 0, m_csi1 = 0, m_csi2 = 0;
  unsigned csi2_bits = 0, csi2_enc = 0;
  bool     ack_open = d.nof_harq_ack_bits != 0, csi1_open = d.nof_csi_part1_bits != 0, csi2_open = false;
  uint32_t in_pos = 0;
  uint32_t out_pos[ULSCH_STREAMS] = {0, 0, 0, 0};
  ReSet    ulsch, uci, rvd, ack, csi1, csi2, tmp;

  for (int l = d.start_symbol_index; l != d.start_symbol_index + d.nof_symbols; ++l) {
    const bool dmrs = (mask >> l) & 1;
    const int  n_re = dmrs ? re_dmrs_symbol : d.nof_prb * 12;
    if (n_re == 0) {
      continue;
    }
    // Once every UCI stream has been placed, the remaining symbols are plain UL-SCH (the reserved elements alone do
    // not change any stream): no element sets needed.
    const bool ack_left  = m_ack < d.nof_enc_harq_ack_bits;
    const bool csi1_left = m_csi1 < d.nof_enc_csi_part1_bits;
    const bool csi2_left = csi2_open ? (m_csi2 < csi2_enc) : (csi1_left && d.nof_enc_csi_part2_bits != 0);
    if (!ack_left && !csi1_left && !csi2_left) {
      UlschSymbol s = {};
      s.in_off      = in_pos;
      s.n_re        = (uint32_t)n_re;
      for (int k = 0; k != ULSCH_STREAMS; ++k) {
        s.out_off[k]  = out_pos[k];
        s.list_off[k] = ULSCH_IDENTITY;
      }
      s.n_out_re[0] = (uint32_t)n_re;
      out_pos[0] += (uint32_t)(n_re * bpre);
      plan.syms.push_back(s);
      in_pos += (uint32_t)(n_re * bpre);
      continue;
    }
    ulsch.assign((size_t)n_re, 1);
    uci.assign((size_t)n_re, dmrs ? 0 : 1);
    rvd.assign((size_t)n_re, 0);
    ack.assign((size_t)n_re, 0);
    csi1.assign((size_t)n_re, 0);
    csi2.assign((size_t)n_re, 0);
    bool any_uci = false;

    int       M_uci   = count(uci);
    const int rem_rvd = (int)((d.nof_harq_ack_rvd - m_rvd) / (unsigned)bpre);
    if (l >= l1 && M_uci > 0 && rem_rvd > 0) {
      int dd, m;
      spread(rem_rvd, M_uci, dd, m);
      select(rvd, ulsch, dd, m);
      m_rvd += (unsigned)(m * bpre);
    }
    const int rem_ack = (int)((d.nof_enc_harq_ack_bits - m_ack) / (unsigned)bpre);
    if (l >= l1 && M_uci > 0 && d.nof_harq_ack_bits > 2 && rem_ack > 0) {
      int dd, m;
      spread(rem_ack, M_uci, dd, m);
      select(ack, uci, dd, m);
      remove(ulsch, ack);
      remove(uci, ack);
      M_uci = count(uci);
      m_ack += (unsigned)(m * bpre);
      any_uci = true;
    }
    const int rem_csi1 = (int)((d.nof_enc_csi_part1_bits - m_csi1) / (unsigned)bpre);
    const int M_rvd    = count(rvd);
    if (l >= l1_csi && (M_uci - M_rvd) > 0 && rem_csi1 > 0) {
      int dd, m;
      spread(rem_csi1, M_uci - M_rvd, dd, m);
      tmp.assign((size_t)n_re, 0);
      for (int i = 0; i != n_re; ++i) {
        tmp[(size_t)i] = (uint8_t)(!rvd[(size_t)i] && uci[(size_t)i]);
      }
      select(csi1, tmp, dd, m);
      remove(ulsch, csi1);
      remove(uci, csi1);
      m_csi1 += (unsigned)(m * bpre);
      any_uci = true;
    }
    auto place_csi2 = [&]() {
      const int M   = count(uci);
      const int rem = (int)((csi2_enc - m_csi2) / (unsigned)bpre);
      if (l >= l1_csi && M > 0 && rem > 0) {
        int dd, m;
        spread(rem, M, dd, m);
        select(csi2, uci, dd, m);
        remove(ulsch, csi2);
        remove(uci, csi2);
        m_csi2 += (unsigned)(m * bpre);
        any_uci = true;
      }
    };
    place_csi2();
    bool punctured = false;
    if (M_rvd > 0 && d.nof_harq_ack_bits <= 2 && rem_ack > 0) {
      int dd, m;
      spread(rem_ack, M_rvd, dd, m);
      select(ack, rvd, dd, m);
      m_ack += (unsigned)(m * bpre);
      // One- and two-bit HARQ-ACK punctures the UL-SCH: the elements stay in the UL-SCH stream, zeroed.
      punctured = d.nof_harq_ack_bits == 1 || d.nof_harq_ack_bits == 2;
      any_uci   = true;
    }
    // Streams that end in this symbol (demux_current_ofdm_symbol :501-589).
    if (count(ack) != 0) {
      if (!ack_open) {
        return false;
      }
      if (m_ack == d.nof_enc_harq_ack_bits) {
        ack_open = false;
      }
    }
    if (count(csi1) != 0) {
      if (!csi1_open) {
        return false;
      }
      if (m_csi1 == d.nof_enc_csi_part1_bits) {
        csi1_open = false;
        if (d.nof_enc_csi_part2_bits != 0) {
          csi2_open = true;
          csi2_bits = d.nof_csi_part2_bits;
          csi2_enc  = d.nof_enc_csi_part2_bits;
          place_csi2();
        }
      }
    }
    if (count(csi2) != 0) {
      if (!csi2_open) {
        return false;
      }
      if (m_csi2 == csi2_enc) {
        csi2_open = false;
      }
    }

    UlschSymbol s = {};
    s.in_off      = in_pos;
    s.n_re        = (uint32_t)n_re;
    const ReSet*   sets[ULSCH_STREAMS] = {&ulsch, &ack, &csi1, &csi2};
    const uint32_t bits[ULSCH_STREAMS] = {0, d.nof_harq_ack_bits, d.nof_csi_part1_bits, csi2_bits};
    for (int k = 0; k != ULSCH_STREAMS; ++k) {
      s.out_off[k]  = out_pos[k];
      s.n_out_re[k] = (uint32_t)count(*sets[k]);
      s.uci_bits[k] = bits[k];
      s.list_off[k] = ULSCH_IDENTITY;
      if (k != 0 || any_uci) {
        s.list_off[k] = (uint32_t)plan.lists.size();
        for (int i = 0; i != n_re; ++i) {
          if ((*sets[k])[(size_t)i]) {
            uint16_t e = (uint16_t)i;
            // The punctured elements are zeroed before CSI Part 2 (which may sit on reserved elements) and UL-SCH
            // are read (demux_current_ofdm_symbol handles HARQ-ACK first, :509-531).
            if ((k == 0 || k == 3) && punctured && ack[(size_t)i]) {
              e |= ULSCH_PUNCTURED;
            }
            plan.lists.push_back(e);
          }
        }
      }
      out_pos[k] += s.n_out_re[k] * (uint32_t)bpre;
    }
    plan.syms.push_back(s);
    in_pos += (uint32_t)(n_re * bpre);
  }
  if (ack_open || csi1_open || csi2_open || in_pos == 0) {
    return false;
  }
  cw.n_sym = (uint32_t)plan.syms.size() - cw.sym_first;
  cw.n_in  = in_pos;
  uint32_t base = 0;
  for (int k = 0; k != ULSCH_STREAMS; ++k) {
    cw.n_out[k] = out_pos[k];
    if (k != 0) {
      cw.uci_base[k] = base;
      base += out_pos[k];
    }
  }
  cw.seq_word_off = plan.seq_words;
  plan.seq_words += (in_pos + 31) / 32 + 16; // generated sixteen words at a time
  plan.seq_words = (plan.seq_words + 3) & ~3u;
  plan.cws.push_back(cw);
  return true;
}

} // namespace pdc
