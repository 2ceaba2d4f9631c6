"""Synthetic test vectors: valid 5G NR codewords through an AWGN channel, quantised like the reference's demodulator.

LLR model: y = +-1 + N(0, sigma^2), LLR = clamp(round(6 * 2y / sigma^2), +-120) - 6 LSB per natural-log unit,
the scale of log_likelihood_ratio::quantize with range 20 (lib/phy/upper/log_likelihood_ratio.cpp:89-98).
Uses the oracle's TX chain (checked against the reference's own in tests/test_oracle_vs_reference.py).
"""
import numpy as np

from oracle import pyoracle as po

LIFTING_SIZES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52,
                 56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352,
                 384]


def awgn_llr(bits, snr_db, rng):
    sigma2 = 10.0 ** (-snr_db / 10.0)
    y = (1.0 - 2.0 * bits.astype(np.float64)) + rng.normal(0.0, np.sqrt(sigma2), bits.size)
    return np.clip(np.round(6.0 * 2.0 * y / sigma2), -120, 120).astype(np.int8)


def random_message(orc, bg, Z, nof_filler, crc_kind, rng):
    """K message bits (one per byte): payload + CRC + zero fillers."""
    K = (22 if bg == 1 else 10) * Z
    crc_bits = {po.CRC_NONE: 0, po.CRC16: 16, po.CRC24A: 24, po.CRC24B: 24}[crc_kind]
    msg = rng.integers(0, 2, K).astype(np.uint8)
    msg[K - nof_filler:] = 0
    nb = K - nof_filler - crc_bits
    assert nb > 0
    if crc_bits:
        c = orc.crc(crc_kind, np.packbits(msg[:nb]), nb)
        msg[nb:nb + crc_bits] = [(c >> (crc_bits - 1 - i)) & 1 for i in range(crc_bits)]
    return msg


class CbBatch:
    """n_cb codeblocks of one shape: rate-matched LLRs plus everything needed to run them on the GPU or the oracle."""

    def __init__(self, bg, Z, E, qm, rv, nref, nof_filler, crc_kind, llrs, msgs):
        self.bg, self.Z, self.E, self.qm, self.rv, self.nref = bg, Z, E, qm, rv, nref
        self.nof_filler, self.crc_kind = nof_filler, crc_kind
        self.llrs = llrs  # (n_cb, E) int8
        self.msgs = msgs  # (n_cb, K) bits
        self.n_cb = llrs.shape[0]
        self.N = (66 if bg == 1 else 50) * Z
        self.K = (22 if bg == 1 else 10) * Z

    def descriptors(self, capi, max_iter, early_stop, new_data=True, harq_base=0, decode=True):
        cbs = np.zeros(self.n_cb, capi.CB_DESC_DTYPE)
        flags = capi.CB_DEMATCH | (capi.CB_DECODE if decode else 0) | (capi.CB_NEW_DATA if new_data else 0) | (
            capi.CB_EARLY_STOP if early_stop else 0)
        for i in range(self.n_cb):
            cbs[i] = (i * self.E, self.E, harq_base + i, self.nref, self.Z, self.nof_filler, self.bg, self.qm, self.rv,
                      self.crc_kind, max_iter, flags, 0xffff)
        return cbs

    def run_gpu(self, ctx, max_iter=6, early_stop=True, new_data=True, harq_base=0, read_harq=True, harq_init=0):
        """harq_init: value the HARQ entries are filled with before a new transmission (None = leave them as they
        are); the reference does not clear soft buffers either, so stale regions are part of the comparison."""
        from srsran_edgeric_5g_b200 import capi
        cbs = self.descriptors(capi, max_iter, early_stop, new_data, harq_base)
        if new_data and harq_init is not None:
            for i in range(self.n_cb):
                ctx.harq_write(harq_base + i, np.full(self.N, harq_init, np.int8))
        ctx.submit(cbs, np.ascontiguousarray(self.llrs.reshape(-1)), None, stream=0, want_bits=True)
        out = ctx.wait(0)
        kb = (self.K + 7) // 8
        res = {
            "crc_ok": out["cb_results"]["crc_ok"].astype(bool),
            "iters": out["cb_results"]["iters"].astype(np.int32),
            "status": out["cb_results"]["status"].astype(np.int32),
            "nlayers": out["cb_results"]["nlayers"].astype(np.int32),
            "bits": out["cb_bits"][:, :kb].copy(),
        }
        if read_harq:
            res["harq"] = np.stack([ctx.harq_read(harq_base + i, self.N) for i in range(self.n_cb)])
        return res

    def run_oracle(self, orc, max_iter=6, early_stop=True, new_data=True, harq=None, scale=po.SCALE_X86,
                   simd_width=64):
        kb = (self.K + 7) // 8
        harq = np.zeros((self.n_cb, self.N), np.int8) if harq is None else harq
        crc_ok = np.zeros(self.n_cb, bool)
        iters = np.zeros(self.n_cb, np.int32)
        bits = np.zeros((self.n_cb, kb), np.uint8)
        for i in range(self.n_cb):
            it, b = orc.cb_decode(harq[i], self.llrs[i], new_data, self.rv, self.qm, self.nref, self.nof_filler,
                                  self.crc_kind, early_stop, max_iter, scale, simd_width)
            crc_ok[i] = it > 0
            iters[i] = it if it > 0 else max_iter
            bits[i] = b
        return {"crc_ok": crc_ok, "iters": iters, "bits": bits, "harq": harq}


def make_cb_batch(orc, bg, Z, n_cb, E, qm, rv, snr_db, seed, crc_kind=po.CRC24B, nof_filler=0, nref=0):
    rng = np.random.default_rng(seed)
    N = (66 if bg == 1 else 50) * Z
    K = (22 if bg == 1 else 10) * Z
    llrs = np.zeros((n_cb, E), np.int8)
    msgs = np.zeros((n_cb, K), np.uint8)
    for i in range(n_cb):
        msg = random_message(orc, bg, Z, nof_filler, crc_kind, rng)
        cw = orc.ldpc_encode(bg, Z, msg)
        tx = orc.rate_match(cw, E, rv, qm, nref, nof_filler)
        llrs[i] = awgn_llr(tx, snr_db, rng)
        msgs[i] = msg
    assert N == cw.size
    return CbBatch(bg, Z, E, qm, rv, nref, nof_filler, crc_kind, llrs, msgs)


def make_tb_llrs(orc, tb, bg, rv, qm, nref, nof_layers, n_llr, snr_db, rng):
    cw, n_cb = orc.tb_encode(tb, bg, rv, qm, nref, nof_layers, n_llr)
    return awgn_llr(cw, snr_db, rng), n_cb


# ---- codeword front end (UL-SCH demultiplexing) -------------------------------------------------------------------------

def random_ulsch_cfg(rng, max_prb=40):
    """A random, self-consistent ulsch_demultiplex configuration (dict with the fields of orc_ulsch_cfg): every
    modulation, 1-4 layers, both DM-RS types, HARQ-ACK with 0 / 1 / 2 / more bits, CSI Part 1 and Part 2. The first OFDM
    symbol of the allocation always carries data (the reference's demodulator cannot start on an empty symbol)."""
    qm = int(rng.choice([1, 2, 4, 6, 8]))
    nl = int(rng.integers(1, 5)) if qm > 1 else 1
    nprb = int(rng.integers(1, max_prb))
    start = int(rng.integers(0, 4))
    nsym = int(rng.integers(4, 15 - start))
    dt = int(rng.integers(1, 3))
    cdm = int(rng.integers(1, 3 if dt == 1 else 4))
    full_dmrs = cdm * (6 if dt == 1 else 4) == 12
    lo = start + 1 if full_dmrs else start
    mask = 0
    for p in set(int(x) for x in rng.integers(lo, start + nsym - 1, size=int(rng.integers(1, 4)))):
        mask |= 1 << p
    cfg = dict(qm=qm, nof_layers=nl, nof_prb=nprb, start_symbol_index=start, nof_symbols=nsym, dmrs_type=dt,
               dmrs_symbol_mask=mask, nof_cdm_groups_without_data=cdm)
    bpre = qm * nl
    per_dmrs = (12 - cdm * (6 if dt == 1 else 4)) * nprb
    nre = sum(per_dmrs if (mask >> l) & 1 else 12 * nprb for l in range(start, start + nsym))
    ack_bits = int(rng.choice([0, 0, 1, 2, 3, 7, 20]))
    cfg["nof_harq_ack_bits"] = ack_bits
    if ack_bits > 0:
        cfg["nof_enc_harq_ack_bits"] = int(rng.integers(1, max(2, nre // 6))) * bpre
    if ack_bits <= 2:
        cfg["nof_harq_ack_rvd"] = max(cfg.get("nof_enc_harq_ack_bits", 0), int(rng.integers(0, max(1, nre // 5))) * bpre)
    if rng.random() < 0.5:
        cfg["nof_csi_part1_bits"] = int(rng.choice([1, 2, 5, 30]))
        cfg["nof_enc_csi_part1_bits"] = int(rng.integers(1, max(2, nre // 6))) * bpre
        if rng.random() < 0.5:
            cfg["nof_csi_part2_bits"] = int(rng.choice([1, 2, 9]))
            cfg["nof_enc_csi_part2_bits"] = int(rng.integers(1, max(2, nre // 6))) * bpre
    return cfg


def ulsch_case(orc, rng, max_prb=40):
    """(cfg, descrambled LLRs, scrambling bits) of a configuration the oracle accepts (all UCI fits)."""
    while True:
        cfg = random_ulsch_cfg(rng, max_prb)
        n = orc.ulsch_codeword_length(cfg)
        llr = rng.integers(-120, 121, n).astype(np.int8)
        seq = rng.integers(0, 2, n).astype(np.uint8)
        rc, outs = orc.ulsch_demux(cfg, llr, seq)
        if rc == 0:
            return cfg, llr, seq, outs


# ---- soft demapper ------------------------------------------------------------------------------------------------------

DEMOD_MODS = (0, 1, 2, 4, 6, 8)  # pi/2-BPSK, BPSK, QPSK, 16QAM, 64QAM, 256QAM
_QAM_UNIT = {0: 1.0, 1: 1.0, 2: 1.0, 4: 1 / np.sqrt(10), 6: 1 / np.sqrt(42), 8: 1 / np.sqrt(170)}


def demod_inputs(rng, n, mod, kind):
    """Symbols and noise variances of one demodulate_soft call. kind 0: plausible equaliser output; 1: magnitudes over
    nine decades; 2: values on (and one or two ulps around) the interval boundaries of the piecewise-linear LLR
    functions with noise variances that make rounding ties; 3: zeros, denormals, near-zero thresholds, infinities, NaNs,
    negative / zero / infinite / NaN noise variances."""
    if kind == 0:
        s = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.8
        nv = np.abs(rng.standard_normal(n)) * 0.05 + 1e-3
    elif kind == 1:
        s = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 10.0 ** rng.uniform(-6, 3, n)
        nv = 10.0 ** rng.uniform(-4, 2, n)
    elif kind == 2:
        s = ((rng.integers(-20, 21, n) + 1j * rng.integers(-20, 21, n)) * np.float32(_QAM_UNIT[mod])).astype(np.complex64)
        re = s.real.copy().view(np.int32)
        re += rng.integers(-2, 3, n).astype(np.int32)
        im = s.imag.copy()
        s = np.empty(n, np.complex64)
        s.real, s.imag = re.view(np.float32), im
        nv = rng.choice(np.array([0.5, 1, 0.25, 0.1, 1 / 3., 0.7, 2.0], np.float32), n)
    else:
        pool = np.array([0, -0.0, 1e-10, -1e-10, 1e-9, 2e-5, 3.2e-5, 3.1e-5, np.inf, -np.inf, np.nan, 1e30, -1e30, 3e9,
                         -3e9, 2147483520.0, 0.3, -0.7, 1e-38, 1e-45], np.float32)
        s = np.empty(n, np.complex64)
        s.real, s.imag = rng.choice(pool, n), rng.choice(pool, n)
        nv = rng.choice(np.array([0, -1, np.nan, np.inf, 1e-30, 1e30, 0.1, 1, 1e-45, -0.0], np.float32), n)
    with np.errstate(all="ignore"):
        return np.ascontiguousarray(s, np.complex64), np.ascontiguousarray(nv, np.float32)


def ofdm_symbol_sizes(cfg):
    """Modulation symbols of every OFDM symbol of a codeword that carries data = sizes of the demodulate_soft calls of
    pusch_demodulator_impl.cpp:160-247 (resource elements x layers)."""
    per_dmrs = 12 - cfg["nof_cdm_groups_without_data"] * (6 if cfg["dmrs_type"] == 1 else 4)
    sizes = []
    for l in range(cfg["start_symbol_index"], cfg["start_symbol_index"] + cfg["nof_symbols"]):
        n_re = (per_dmrs if (cfg["dmrs_symbol_mask"] >> l) & 1 else 12) * cfg["nof_prb"]
        if n_re:
            sizes.append(n_re * cfg["nof_layers"])
    return sizes


def demod_codeword(orc, cfg, symbols, noise_vars, mod=None, simd=True):
    """The soft bits pusch_demodulator_impl obtains for a codeword: one oracle demodulate_soft call per OFDM symbol."""
    mod = cfg["qm"] if mod is None else mod
    out, pos = [], 0
    for n in ofdm_symbol_sizes(cfg):
        out.append(orc.demodulate_soft(symbols[pos:pos + n], noise_vars[pos:pos + n], mod, simd))
        pos += n
    assert pos == symbols.size
    return np.concatenate(out)


def modulate(bits, qm, rng=None):
    """TS 38.211 5.1 modulation mapper (QPSK .. 256QAM), written from the standard's formulas."""
    b = bits.reshape(-1, qm).astype(np.float64)
    s = 1 - 2 * b
    if qm == 2:
        re, im, norm = s[:, 0], s[:, 1], np.sqrt(2)
    elif qm == 4:
        re, im, norm = s[:, 0] * (2 - s[:, 2]), s[:, 1] * (2 - s[:, 3]), np.sqrt(10)
    elif qm == 6:
        re, im = s[:, 0] * (4 - s[:, 2] * (2 - s[:, 4])), s[:, 1] * (4 - s[:, 3] * (2 - s[:, 5]))
        norm = np.sqrt(42)
    else:
        re = s[:, 0] * (8 - s[:, 2] * (4 - s[:, 4] * (2 - s[:, 6])))
        im = s[:, 1] * (8 - s[:, 3] * (4 - s[:, 5] * (2 - s[:, 7])))
        norm = np.sqrt(170)
    return ((re + 1j * im) / norm).astype(np.complex64)


# ---- randomised HARQ sequences through the queued batch path ------------------------------------------------------------

def harq_sequence_rounds(ctx, orc, rng, n_ent, rounds, harq_base=0, stream=0, max_Z=384, debug=None):
    """A population of HARQ entries of random shapes lives through `rounds` slots of ONE pdc_submit each: per entry and
    slot a new transmission (new message) or a retransmission of the current one, any redundancy version, aligned and
    unaligned lengths from a fraction of a lap to three laps, limited buffers, filler bits, AWGN soft bits of the valid
    codeword - some thinned out or zeroed (the position of the last non-zero soft bit moves, up and down), some carrying
    the non-finite values -128 / +-127 -, per-codeblock iteration limits and early-stop flags. After every slot: whole
    HARQ entries (stale regions included), CRC flags, iteration counts and decoded bits against the oracle's
    pusch_codeblock_decoder (rate dematcher + decoder on the entry it keeps: pusch_codeblock_decoder.cpp:35-71).
    Returns the number of codeblocks compared; raises AssertionError with the case on a mismatch."""
    from srsran_edgeric_5g_b200 import capi
    sizes = [z for z in LIFTING_SIZES if z <= max_Z]
    ents = []
    for i in range(n_ent):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice(sizes[8:] if rng.random() < 0.7 else sizes))
        kb = 22 if bg == 1 else 10
        N, K, Ksys = (66 if bg == 1 else 50) * Z, kb * Z, (kb - 2) * Z
        qm = int(rng.choice([1, 2, 4, 6, 8]))
        crc_kind = po.CRC24B if K > 64 else po.CRC16
        f_max = min(Ksys - 1, 2 * Z, K - 26)
        F = int(rng.integers(0, max(1, f_max))) if rng.random() < 0.5 else 0
        if rng.random() < 0.7:
            F -= F % 4
        nref = 0
        if rng.random() < 0.4:
            nref = int(rng.integers(Ksys + 2 * Z, N + 1))
            if rng.random() < 0.7:
                nref &= ~3
        ents.append(dict(bg=bg, Z=Z, kb=kb, N=N, K=K, qm=qm, F=F, nref=nref, crc=crc_kind, cw=None,
                         buf=ctx.harq_read(harq_base + i, N)))
    n_cmp = 0
    for rnd in range(rounds):
        cbs = np.zeros(n_ent, capi.CB_DESC_DTYPE)
        llrs, off, want, before = [], 0, [], []
        mi_all, es_all = int(rng.integers(1, 9)), bool(rng.integers(0, 2))
        for i, e in enumerate(ents):
            new = e["cw"] is None or rng.random() < 0.35
            if new:
                e["cw"] = orc.ldpc_encode(e["bg"], e["Z"], random_message(orc, e["bg"], e["Z"], e["F"], e["crc"], rng))
            ncb = min(e["nref"], e["N"]) if e["nref"] else e["N"]
            qm, lap = e["qm"], ncb - e["F"]
            mode = rng.random()
            if mode < 0.6:     # at most one lap, a multiple of four symbols (the express paths when the rest is aligned)
                E = int(rng.integers(1, max(2, lap // (4 * qm)) + 1)) * 4 * qm
            elif mode < 0.85:  # any length up to one lap and a bit
                E = int(rng.integers(1, max(2, (lap + 2 * e["Z"]) // qm))) * qm
            else:              # several laps
                E = int(rng.integers(max(1, lap // qm), max(2, 3 * lap // qm))) * qm
            rv = 0 if (new and rng.random() < 0.7) else int(rng.integers(0, 4))
            l = awgn_llr(orc.rate_match(e["cw"], E, rv, qm, e["nref"], e["F"]), float(rng.uniform(-4.0, 9.0)), rng)
            thin = rng.random()
            if thin < 0.15:
                l = (l * (rng.random(E) < rng.uniform(0.02, 0.5))).astype(np.int8)
            elif thin < 0.20:
                l[:] = 0
            elif thin < 0.30:
                k = rng.integers(0, E, max(1, E // 50))
                l[k] = rng.choice(np.array([-128, -127, 127], np.int8), k.size)
            mi = mi_all if rng.random() < 0.8 else int(rng.integers(1, 9))
            es = es_all if rng.random() < 0.8 else bool(rng.integers(0, 2))
            flags = capi.CB_DEMATCH | capi.CB_DECODE | (capi.CB_NEW_DATA if new else 0) | (capi.CB_EARLY_STOP if es else 0)
            cbs[i] = (off, E, harq_base + i, e["nref"], e["Z"], e["F"], e["bg"], qm, rv, e["crc"], mi, flags, 0xffff)
            llrs.append(l)
            off += E
            if debug is not None:
                before.append(e["buf"].copy())
            it, bits = orc.cb_decode(e["buf"], l, new, rv, qm, e["nref"], e["F"], e["crc"], es, mi)
            want.append((it > 0, it if it > 0 else mi, bits, (new, rv, E, mi, es)))
        if debug is not None:
            debug.setdefault("history", []).append(cbs.copy())
        ctx.submit(cbs, np.concatenate(llrs), None, stream=stream, want_bits=True)
        out = ctx.wait(stream)
        res = out["cb_results"]
        for i, e in enumerate(ents):
            case = (rnd, i, {k: e[k] for k in ("bg", "Z", "qm", "F", "nref")}, want[i][3])
            got = ctx.harq_read(harq_base + i, e["N"])
            assert (got == e["buf"]).all(), ("HARQ entry", case, np.nonzero(got != e["buf"])[0][:8])
            if not e["buf"].any():
                assert res["status"][i] == 1 and res["crc_ok"][i] == 0, ("all-zero entry", case, res[i])
                continue
            assert res["status"][i] == 0, ("status", case, res[i])
            assert bool(res["crc_ok"][i]) == want[i][0] and int(res["iters"][i]) == want[i][1], ("CRC / iterations", case, res[i], want[i][:2])
            kbytes = (e["K"] + 7) // 8
            if debug is not None and not (out["cb_bits"][i, :kbytes] == want[i][2]).all():
                debug.update(ent=e, llr=llrs[i], desc=cbs[i].copy(), res=res[i].copy(), got=out["cb_bits"][i, :kbytes].copy(),
                             want=want[i][2], before=before[i], cbs=cbs.copy(), index=i)
            assert (out["cb_bits"][i, :kbytes] == want[i][2]).all(), ("decoded bits", case)
            n_cmp += 1
    return n_cmp
