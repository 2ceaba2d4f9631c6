"""Shared-channel link adaptation arithmetic the PUSCH path is configured from: MCS tables, transport block size and
LDPC base graph selection. Host-side mirror of the reference's helpers, written from the standard's procedures:

    pusch_mcs_get_config           include/srsran/ran/pusch/pusch_mcs.h:55      (TS 38.214 Tables 5.1.3.1-1/-2, 6.1.4.1)
    tbs_calculator_calculate       lib/ran/sch/tbs_calculator.cpp:166-188       (TS 38.214 5.1.3.2 steps 1-4)
    get_ldpc_base_graph            include/srsran/ran/sch/ldpc_base_graph.h:38-48 (TS 38.212 6.2.2 / 7.2.2)

The reference evaluates the intermediate number of information bits in single precision; so does this module
(numpy.float32), and tests/test_host_logic.py compares the two for every MCS and allocation size.
"""
import numpy as np

BG1, BG2 = 1, 2

# (bits per symbol, target code rate x 1024): TS 38.214 Table 5.1.3.1-1 (64QAM table) and 5.1.3.1-2 (256QAM table).
MCS_TABLE_QAM64 = [(2, 120), (2, 157), (2, 193), (2, 251), (2, 308), (2, 379), (2, 449), (2, 526), (2, 602), (2, 679),
                   (4, 340), (4, 378), (4, 434), (4, 490), (4, 553), (4, 616), (4, 658), (6, 438), (6, 466), (6, 517),
                   (6, 567), (6, 616), (6, 666), (6, 719), (6, 772), (6, 822), (6, 873), (6, 910), (6, 948)]
MCS_TABLE_QAM256 = [(2, 120), (2, 193), (2, 308), (2, 449), (2, 602), (4, 378), (4, 434), (4, 490), (4, 553), (4, 616),
                    (4, 658), (6, 466), (6, 517), (6, 567), (6, 616), (6, 666), (6, 719), (6, 772), (6, 822), (6, 873),
                    (8, 682.5), (8, 711), (8, 754), (8, 797), (8, 841), (8, 885), (8, 916.5), (8, 948)]

# TS 38.214 Table 5.1.3.2-1: the transport block sizes up to 3824 bits.
TBS_TABLE = [24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 136, 144, 152, 160, 168, 176, 184, 192, 208, 224,
             240, 256, 272, 288, 304, 320, 336, 352, 368, 384, 408, 432, 456, 480, 504, 528, 552, 576, 608, 640, 672,
             704, 736, 768, 808, 848, 888, 928, 984, 1032, 1064, 1128, 1160, 1192, 1224, 1256, 1288, 1320, 1352, 1416,
             1480, 1544, 1608, 1672, 1736, 1800, 1864, 1928, 2024, 2088, 2152, 2216, 2280, 2408, 2472, 2536, 2600, 2664,
             2728, 2792, 2856, 2976, 3104, 3240, 3368, 3496, 3624, 3752, 3824]


def pusch_mcs_get_config(table, index):
    """(bits per symbol, target code rate x 1024) of an MCS index; table: "qam64" or "qam256" (no transform precoding)."""
    t = MCS_TABLE_QAM64 if table == "qam64" else MCS_TABLE_QAM256
    return t[index]


def tbs_calculate(nof_symb_sh, nof_dmrs_prb, nof_oh_prb, qm, tcr_x1024, nof_layers, n_prb, tb_scaling_field=0):
    """TS 38.214 5.1.3.2. Returns the transport block size in bits."""
    f = np.float32
    nof_re = min(12 * nof_symb_sh - nof_dmrs_prb - nof_oh_prb, 156) * n_prb          # step 1
    tcr = f(tcr_x1024) * f(1.0 / 1024)
    scaling = f(1.0) / f(1 << tb_scaling_field)
    nof_info = scaling * f(nof_re) * tcr * f(qm) * f(nof_layers)                     # step 2, left to right
    if nof_info <= f(3824):                                                          # step 3
        n = 3
        if nof_info > f(512):
            n = int(np.floor(np.log2(nof_info))) - 6
        nof_info_prime = max(24, (1 << n) * int(np.floor(nof_info / f(1 << n))))
        return next(t for t in TBS_TABLE if t >= nof_info_prime)
    n = int(np.floor(np.log2(nof_info - f(24))) - f(5.0))                            # step 4
    q = (nof_info - f(24)) / f(1 << n)
    r = int(np.floor(q + f(0.5))) if q >= 0 else -int(np.floor(-q + f(0.5)))         # std::round: half away from zero
    nof_info_prime = max(3840, (1 << n) * r)
    C = 1
    if tcr <= f(0.25):
        C = -(-(nof_info_prime + 24) // 3816)
    elif nof_info_prime > 8424:
        C = -(-(nof_info_prime + 24) // 8424)
    return 8 * C * (-(-(nof_info_prime + 24) // (8 * C))) - 24


def get_ldpc_base_graph(rate, tbs_bits):
    """TS 38.212 6.2.2: base graph 2 for small or low-rate transport blocks."""
    r = np.float32(rate)
    if tbs_bits <= 292 or r <= np.float32(0.25) or (tbs_bits <= 3824 and r <= np.float32(0.67)):
        return BG2
    return BG1


def pusch_allocation(table, mcs_index, n_prb, nof_symbols=14, nof_dmrs_symbols=1, nof_cdm_groups_without_data=2,
                     dmrs_type=1, nof_layers=1):
    """Everything the decoder needs to know about one PUSCH allocation without UCI: dict with tbs_bits, base_graph, qm,
    nof_layers, n_llr (soft bits of the codeword) and the target code rate."""
    qm, r1024 = pusch_mcs_get_config(table, mcs_index)
    dmrs_re_per_prb_symbol = nof_cdm_groups_without_data * (6 if dmrs_type == 1 else 4)
    nof_dmrs_prb = dmrs_re_per_prb_symbol * nof_dmrs_symbols
    tbs = tbs_calculate(nof_symbols, nof_dmrs_prb, 0, qm, r1024, nof_layers, n_prb)
    n_re = (12 * nof_symbols - nof_dmrs_prb) * n_prb
    return dict(tbs_bits=tbs, base_graph=get_ldpc_base_graph(r1024 / 1024.0, tbs), qm=qm, nof_layers=nof_layers,
                n_llr=n_re * qm * nof_layers, rate=r1024 / 1024.0, n_prb=n_prb, mcs=mcs_index)
