// Constant-memory tables: TS 38.212 base graphs (sparse form) and CRC shift polynomials.
#pragma once

#include "pdc_device.cuh"
#include <string.h>

namespace pdc {

#include "bg_tables.inc"

__constant__ BgTables c_tab;

// Host copy of the tables (filled by upload_tables; the decoder's per-shape images are derived from it).
inline BgTables& host_tables()
{
  static BgTables h;
  return h;
}

inline cudaError_t upload_tables()
{
  BgTables& h = host_tables();
  memset(&h, 0, sizeof(h));
  for (int bg = 0; bg != 2; ++bg) {
    int n                        = bg ? BG2_NOF_EDGES : BG1_NOF_EDGES;
    const unsigned short(*e)[10] = bg ? BG2_EDGES : BG1_EDGES;
    int rows                     = bg ? 42 : 46;
    int r                        = 0;
    h.row_start[bg][0]           = 0;
    for (int i = 0; i != n; ++i) {
      h.row[bg][i] = (uint8_t)e[i][0];
      h.col[bg][i] = (uint8_t)e[i][1];
      for (int s = 0; s != 8; ++s) {
        h.v[bg][s][i] = e[i][2 + s];
      }
      while (r < e[i][0]) {
        h.row_start[bg][++r] = (uint16_t)i;
      }
    }
    while (r < rows) {
      h.row_start[bg][++r] = (uint16_t)n;
    }
    // Rows that touch none of the variable nodes of the rows processed since the last barrier need no barrier of their
    // own (the extension rows of both base graphs are pairwise column-disjoint in consecutive pairs).
    {
      bool used[MAX_EDGES] = {false}; // indexed by column (< 68)
      for (int m = 0; m != rows; ++m) {
        bool clash = (m == 0);
        for (int i = h.row_start[bg][m]; i != h.row_start[bg][m + 1]; ++i) {
          clash = clash || used[h.col[bg][i]];
        }
        if (clash) {
          for (bool& u : used) {
            u = false;
          }
        }
        for (int i = h.row_start[bg][m]; i != h.row_start[bg][m + 1]; ++i) {
          used[h.col[bg][i]] = true;
        }
        h.row_free[bg][m] = clash ? 0 : 1;
      }
    }
    h.row_pstart[bg][0] = 0;
    for (int m = 0; m != rows; ++m) {
      int deg                 = h.row_start[bg][m + 1] - h.row_start[bg][m];
      h.row_pstart[bg][m + 1] = (uint16_t)(h.row_pstart[bg][m] + ((deg + 1) & ~1));
    }
  }
  memset(h.set_index, 0xff, sizeof(h.set_index));
  for (int i = 0; i != NR_LDPC_NOF_LIFTING_SIZES; ++i) {
    h.set_index[NR_LDPC_LIFTING_SIZES[i]] = NR_LDPC_SET_INDEX[i];
  }
  for (int k = 0; k != 3; ++k) {
    int      kind  = k + 1;
    uint32_t poly  = crc_poly(kind);
    int      order = crc_order(kind);
    uint32_t top   = 1u << order;
    uint32_t x     = 1;
    for (int j = 0; j != XPOW_ENTRIES; ++j) {
      h.xpow32[k][j] = x;
      for (int b = 0; b != 32; ++b) {
        x <<= 1;
        if (x & top) {
          x ^= poly;
        }
      }
    }
  }
  return cudaMemcpyToSymbol(c_tab, &h, sizeof(h));
}

} // namespace pdc
