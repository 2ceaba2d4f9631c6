import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from srsran_edgeric_5g_b200 import capi, ldpc
from oracle.pyoracle import Oracle
from tests.vectors import make_tb_llrs
orc = Oracle(); rng = np.random.default_rng(3)
tbs_bits, n_llr, qm, nl = 1277992, 1362816, 8, 4
C = ldpc.compute_nof_codeblocks(tbs_bits, 1); nref = ldpc.compute_N_ref(tbs_bits // 8, C)
tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
llrs, _ = make_tb_llrs(orc, tb, 1, 0, qm, nref, nl, n_llr, 8.4, rng)
metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n_llr)
cells = 16; n_cb = C * cells; tb_stride = (tbs_bits + 24 + 31) // 32 * 4
cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE); tbd = np.zeros(cells, capi.TB_DESC_DTYPE)
flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA
cws = np.zeros(cells, capi.CW_DESC_DTYPE); raws = []
for c in range(cells):
    tbd[c] = (c * C, C, tbs_bits, c * tb_stride, 0)
    for k, m in enumerate(metas):
        cbs[c * C + k] = (c * n_llr + m.cw_offset, m.rm_length, c * C + k, nref, m.lifting_size, m.nof_filler_bits, 1, qm, 0, capi.CRC24B, 6, flags, c)
    c_init = (0x4601 + c) * 32768 + 17 * c
    cws[c]["in_offset"], cws[c]["sch_offset"], cws[c]["c_init"] = c * n_llr, c * n_llr, c_init
    cws[c]["flags"] = capi.CW_SCRAMBLED | capi.CW_DEFER_DESCRAMBLING
    for k, v in (("qm", qm), ("nof_layers", nl), ("nof_prb", 273), ("nof_symbols", 14), ("dmrs_type", 1), ("dmrs_symbol_mask", 1 << 2), ("nof_cdm_groups_without_data", 2)):
        cws[c][k] = v
    raws.append(orc.revert_scrambling(llrs, orc.prg_bits(c_init, 0, n_llr)))
ctx = capi.Context(device=0, max_cbs=n_cb, max_llrs=cells * n_llr + 64, harq_entries=2 * n_cb, max_tbs=cells, max_tb_bytes=cells * tb_stride + 64, nof_streams=2)
raw_pin = [capi.PinnedBuffer(cells * n_llr) for _ in range(2)]
bits_pin = [capi.PinnedBuffer(n_cb * capi.PDC_MAX_CB_BYTES, np.uint8) for _ in range(2)]
tb_pin = [capi.PinnedBuffer(cells * tb_stride + 64, np.uint8) for _ in range(2)]
for b in raw_pin: b.array[:] = np.concatenate(raws)
sch_pin = capi.PinnedBuffer(cells * n_llr); sch_pin.array[:] = np.tile(llrs, cells)
T = time.perf_counter
def t(f):
    a = T(); r = f(); return (T() - a) * 1e6, r
for rep in range(3):
    a, _ = t(lambda: ctx.submit_codewords(cws, raw_pin[0].array, stream=0))
    b, _ = t(lambda: ctx.submit(cbs, None, tbd, stream=0, out_bits=bits_pin[0].array, out_tb=tb_pin[0].array))
    c, _ = t(lambda: ctx.wait(0))
    print("front+decode: submit_codewords %.0f us, submit %.0f us, wait %.0f us" % (a, b, c))
for rep in range(3):
    b, _ = t(lambda: ctx.submit(cbs, sch_pin.array, tbd, stream=0, out_bits=bits_pin[0].array, out_tb=tb_pin[0].array))
    c, _ = t(lambda: ctx.wait(0))
    print("plain: submit %.0f us, wait %.0f us" % (b, c))
for rep in range(3):
    a, _ = t(lambda: ctx.submit_codewords(cws, raw_pin[0].array, stream=0))
    c, _ = t(lambda: ctx.wait(0))
    print("front only: submit_codewords %.0f us, wait %.0f us" % (a, c))
# pipelined plain
def run_plain(n):
    ctx.submit(cbs, sch_pin.array, tbd, stream=0, out_bits=bits_pin[0].array, out_tb=tb_pin[0].array)
    for i in range(1, n):
        q = i & 1
        c2 = cbs
        ctx.submit(c2, sch_pin.array, tbd, stream=q, out_bits=bits_pin[q].array, out_tb=tb_pin[q].array)
        ctx.wait((i - 1) & 1)
    ctx.wait((n - 1) & 1)
run_plain(4)
a, _ = t(lambda: run_plain(20)); print("pipelined plain: %.0f us/slot" % (a / 20))
def run_fe(n, cw):
    def slot(i):
        q = i & 1
        ctx.submit_codewords(cw, raw_pin[q].array, stream=q)
        ctx.submit(cbs, None, tbd, stream=q, out_bits=bits_pin[q].array, out_tb=tb_pin[q].array)
    slot(0)
    for i in range(1, n):
        slot(i)
        ctx.wait((i - 1) & 1)
    ctx.wait((n - 1) & 1)
cws_nd = cws.copy(); cws_nd["flags"] = capi.CW_SCRAMBLED
for name, cw in (("deferred", cws), ("materialised", cws_nd)):
    run_fe(10, cw)
    a, _ = t(lambda: run_fe(20, cw)); print("pipelined front end (%s) + decode: %.0f us/slot" % (name, a / 20))
def run_fe_only(n):
    ctx.submit_codewords(cws_nd, raw_pin[0].array, stream=0)
    for i in range(1, n):
        q = i & 1
        ctx.submit_codewords(cws_nd, raw_pin[q].array, stream=q)
        ctx.wait((i - 1) & 1)
    ctx.wait((n - 1) & 1)
run_fe_only(10)
a, _ = t(lambda: run_fe_only(20)); print("pipelined front end only: %.0f us/slot" % (a / 20))
