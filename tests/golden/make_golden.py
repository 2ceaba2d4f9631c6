#!/usr/bin/env python3
"""Generates the golden fixtures of tests/golden/ (run in the build container, where /root/reference exists).

  ldpc_examples.npz     - message/codeword pairs of the reference tree's own on-disk vectors
                          (srs-4G-UE/lib/src/phy/fec/ldpc/test/examplesBG{1,2}.dat, format of ldpc_dec_c_test.c:93-135):
                          first NOF_EXAMPLES of the 10 pairs of every (base graph, lifting size), bit-packed.
  ref_decoder.npz       - LLR inputs and the outputs (bits, iterations) of the compiled reference's ldpc_decoder
                          ("auto" = avx512/avx2, and "generic"), produced through oracle/_ref/libsrsref.so.
  ref_dematcher.npz     - HARQ buffer before/after the compiled reference's ldpc_rate_dematcher.
  ref_frontend.npz      - codeword front end: pseudo-random sequences of the compiled reference's
                          pseudo_random_generator_impl and inputs/outputs of its ulsch_demultiplex_impl (fed block by
                          block like pusch_demodulator_impl does) for 40 random configurations.
  ref_pusch.npz         - transport-block level: LLRs of 4 HARQ transmissions, expected TB bytes/CRC/statistics and
                          CRC32 of every codeblock soft buffer after each transmission (reference pusch_decoder_impl).
"""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as po  # noqa: E402
from tests.vectors import LIFTING_SIZES, awgn_llr, random_message  # noqa: E402

OUT = Path(__file__).resolve().parent
EXAMPLES = Path("/root/reference/srs-4G-UE/lib/src/phy/fec/ldpc/test")
NOF_EXAMPLES = 2


def parse_examples(path, bg):
    text = path.read_text().split("\n")
    out = {}
    i = 0
    while i < len(text):
        line = text[i].strip()
        if line.startswith("ls") and line.endswith("msgs"):
            Z = int(line[2:-4])
            msgs = text[i + 1:i + 11]
            assert text[i + 11].strip() == f"ls{Z}cwds"
            cwds = text[i + 12:i + 22]
            out[Z] = (msgs, cwds)
            i += 22
        else:
            i += 1
    return out


def gen_examples():
    data = {}
    for bg, name in ((1, "examplesBG1.dat"), (2, "examplesBG2.dat")):
        ex = parse_examples(EXAMPLES / name, bg)
        assert sorted(ex) == LIFTING_SIZES
        for Z, (msgs, cwds) in ex.items():
            K = (22 if bg == 1 else 10) * Z
            N = (66 if bg == 1 else 50) * Z
            m = np.zeros((NOF_EXAMPLES, K), np.uint8)
            f = np.zeros((NOF_EXAMPLES, K), np.uint8)
            c = np.zeros((NOF_EXAMPLES, N), np.uint8)
            for k in range(NOF_EXAMPLES):
                assert len(msgs[k]) == K and len(cwds[k]) == N
                m[k] = [0 if ch == "-" else int(ch) for ch in msgs[k]]
                f[k] = [1 if ch == "-" else 0 for ch in msgs[k]]
                c[k] = [0 if ch == "-" else int(ch) for ch in cwds[k]]
            data[f"bg{bg}_z{Z}_msg"] = np.packbits(m, axis=1)
            data[f"bg{bg}_z{Z}_filler"] = f.sum(axis=1).astype(np.int32)
            data[f"bg{bg}_z{Z}_cw"] = np.packbits(c, axis=1)
    np.savez_compressed(OUT / "ldpc_examples.npz", **data)


def gen_decoder(orc):
    rng = np.random.default_rng(2024)
    ref = {"auto": po.Reference("auto"), "generic": po.Reference("generic")}
    cases = []
    llrs = []
    outs = {"auto": [], "generic": []}
    for trial in range(120):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice(LIFTING_SIZES)) if trial % 3 else int(rng.choice([384, 352, 320, 256, 64, 2, 3]))
        kb = 22 if bg == 1 else 10
        K = kb * Z
        crc_kind = int(rng.integers(0, 4))
        crc_bits = 16 if crc_kind == 1 else 24
        F = int(rng.integers(0, max(1, min(Z, K - crc_bits - 1)))) if rng.random() < 0.5 else 0
        if K - F - crc_bits <= 0:
            continue
        msg = random_message(orc, bg, Z, F, crc_kind, rng)
        cw = orc.ldpc_encode(bg, Z, msg)
        nodes = int(rng.integers(kb + 2, (66 if bg == 1 else 50) + 1))
        rate = K / (nodes * Z)
        snr = {True: 10 * np.log10(2 ** (2 * rate) - 1) + 1.5, False: 0}[True] + rng.uniform(-1.5, 1.5)
        llr = awgn_llr(cw, snr, rng)
        llr[K - 2 * Z - F:K - 2 * Z] = 127
        n_in = nodes * Z
        if rng.random() < 0.3:
            llr_in = llr[:n_in].copy()
        else:
            llr_in = llr.copy()
            llr_in[n_in:] = 0
        mi = int(rng.integers(1, 9))
        cases.append((bg, Z, F, crc_kind, mi, llr_in.size))
        llrs.append(llr_in)
        for name, r in ref.items():
            it, bits = r.ldpc_decode(bg, Z, llr_in, F, crc_kind, mi)
            outs[name].append((it, bits))
    data = {"cases": np.array(cases, np.int32), "llrs": np.concatenate(llrs)}
    for name in outs:
        data[f"iters_{name}"] = np.array([o[0] for o in outs[name]], np.int32)
        data[f"bits_{name}"] = np.concatenate([o[1] for o in outs[name]])
    np.savez_compressed(OUT / "ref_decoder.npz", **data)


def gen_dematcher():
    rng = np.random.default_rng(77)
    ref = po.Reference("auto")
    width = 64 if ref.auto_variant() == "avx512" else 32
    cases, bufs0, llrs, bufs1 = [], [], [], []
    for trial in range(150):
        bg = int(rng.integers(1, 3))
        Z = int(rng.choice(LIFTING_SIZES))
        kb = 22 if bg == 1 else 10
        N = (66 if bg == 1 else 50) * Z
        Ksys = (kb - 2) * Z
        qm = int(rng.choice([1, 2, 4, 6, 8]))
        F = int(rng.integers(0, min(Ksys - 1, 2 * Z))) if rng.random() < 0.6 else 0
        nref = int(rng.integers(Ksys + 2 * Z, N + 50)) if rng.random() < 0.4 else 0
        E = int(rng.integers(1, max(2, 3 * N // qm))) * qm
        rv = int(rng.integers(0, 4))
        mode = rng.random()
        buf0 = (rng.integers(-120, 121, N) if mode < 0.6 else rng.integers(-128, 128, N) if mode < 0.8 else
                np.zeros(N)).astype(np.int8)
        llr = (rng.integers(-120, 121, E) if rng.random() < 0.8 else rng.integers(-128, 128, E)).astype(np.int8)
        new_data = int(rng.integers(0, 2))
        out = buf0.copy()
        ref.rate_dematch(out, llr, new_data, rv, qm, nref, F)
        cases.append((N, E, new_data, rv, qm, nref, F))
        bufs0.append(buf0)
        llrs.append(llr)
        bufs1.append(out)
    np.savez_compressed(OUT / "ref_dematcher.npz", cases=np.array(cases, np.int32), simd_width=np.int32(width),
                        buf0=np.concatenate(bufs0), llrs=np.concatenate(llrs), buf1=np.concatenate(bufs1))


def gen_pusch(orc):
    rng = np.random.default_rng(99)
    cases, tbs, llrs, stats, tb_out, crcs = [], [], [], [], [], []
    for trial in range(10):
        bg = 1 if trial % 3 else 2
        tb_bytes = int(rng.integers(20, 700 if bg == 2 else 4000))
        qm = int(rng.choice([2, 4, 6, 8]))
        nl = int(rng.integers(1, 3))
        rate = rng.uniform(0.55, 0.9) if bg == 1 else rng.uniform(0.25, 0.6)
        nsym = int(np.ceil(tb_bytes * 8 / rate / qm / nl)) * nl
        n_llr = nsym * qm
        C = len(orc.segment_rx(tb_bytes * 8, bg, qm, nl, n_llr))
        nref = 0 if trial % 2 else min(25344, (tb_bytes + 40) * 8 * 3 // (2 * C))
        es, mi = bool(trial % 2), int(rng.integers(2, 7))
        snr = (8 if bg == 1 else 3) * rate / 0.8 + rng.uniform(-4, -1)
        tb = rng.integers(0, 256, tb_bytes).astype(np.uint8)
        fill = int(rng.integers(-50, 50))
        rp = po.ReferencePusch(C, "auto", fill)
        cases.append((bg, tb_bytes, qm, nl, n_llr, nref, int(es), mi, fill, C))
        tbs.append(tb)
        for t, rv in enumerate([0, 2, 3, 1]):
            cw, _ = orc.tb_encode(tb, bg, rv, qm, nref, nl, n_llr)
            llr = awgn_llr(cw, snr, rng)
            out, st = rp.decode(llr, tb_bytes, bg, rv, qm, nref, nl, mi, es, t == 0)
            llrs.append(llr)
            stats.append(st[:5])
            tb_out.append(out)
            N = orc.segment_rx(tb_bytes * 8, bg, qm, nl, n_llr)[0].full_length
            crcs.append([zlib.crc32(rp.get_cb(cb, N)[0].tobytes()) for cb in range(C)] + [0] * (16 - C))
    np.savez_compressed(OUT / "ref_pusch.npz", cases=np.array(cases, np.int32), tbs=np.concatenate(tbs),
                        llrs=np.concatenate(llrs), stats=np.array(stats, np.int32), tb_out=np.concatenate(tb_out),
                        soft_crc32=np.array(crcs, np.uint32))


def gen_frontend(orc):
    from tests.vectors import ulsch_case
    ref = po.Reference("auto")
    rng = np.random.default_rng(2024)
    prg_cases = [(0, 0, 256), (1, 0, 300), (12345, 0, 2000), (0x7FFFFFFF, 777, 1000), (1 << 30, 100000, 500),
                 (4660 * 32768 + 321, 1362816 - 640, 640)]
    prg = [np.packbits(ref.prg_bits(c, o, n)) for c, o, n in prg_cases]
    cfgs, llrs, seqs, outs, lens = [], [], [], [], []
    while len(cfgs) < 40:
        cfg, llr, seq, _ = ulsch_case(orc, rng, max_prb=12)
        rc, ro = ref.ulsch_demux(cfg, llr, seq, int(rng.choice([0, 0, 7, 50])))
        if rc != 0:
            continue
        cfgs.append(po.ulsch_cfg_array(cfg))
        llrs.append(llr)
        seqs.append(np.packbits(seq))
        outs.extend(ro)
        lens.append([llr.size] + [o.size for o in ro])
    np.savez_compressed(OUT / "ref_frontend.npz", prg_cases=np.array(prg_cases, np.int64), prg_bits=np.concatenate(prg),
                        cfgs=np.array(cfgs, np.int32), llrs=np.concatenate(llrs), seq_bits=np.concatenate(seqs),
                        outs=np.concatenate(outs), lens=np.array(lens, np.int64))


def gen_demod():
    """demodulation_mapper_impl::demodulate_soft of the compiled reference (x86 build: AVX2 blocks + scalar remainder) on
    every modulation, call sizes with and without a remainder, and the four input families of tests.vectors.demod_inputs."""
    from tests.vectors import DEMOD_MODS, demod_inputs
    ref = po.Reference()
    rng = np.random.default_rng(2024)
    cases, syms, nvs, outs = [], [], [], []
    for mod in DEMOD_MODS:
        for kind in range(4):
            for n in (1, 3, 4, 7, 16, 24, 36, 45, 132, 300):
                s, nv = demod_inputs(rng, n, mod, kind)
                o = ref.demodulate_soft(s, nv, mod)
                cases.append((mod, kind, n))
                syms.append(s)
                nvs.append(nv)
                outs.append(o)
    np.savez_compressed(OUT / "ref_demod.npz", cases=np.array(cases, np.int32), symbols=np.concatenate(syms),
                        noise_vars=np.concatenate(nvs), llrs=np.concatenate(outs))


if __name__ == "__main__":
    orc = po.Oracle()
    if "--demod-only" in sys.argv:
        gen_demod()
        sys.exit(0)
    if "--frontend-only" in sys.argv:
        gen_frontend(orc)
        sys.exit(0)
    gen_examples()
    gen_frontend(orc)
    gen_demod()
    gen_decoder(orc)
    gen_dematcher()
    gen_pusch(orc)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
