"""TEST INFRASTRUCTURE - ctypes bindings for the two checkers under oracle/_ref/.

* ``Oracle``    - oracle/pusch_oracle.c, the plain-C restatement (always available once ``make -C oracle oracle`` ran).
* ``Reference`` - the unmodified reference compiled from /root/reference (``make -C oracle ref``); the prebuilt
  library travels to the GPU box, the sources do not.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
ORACLE_SO = REF_DIR / "liboracle.so"
REFERENCE_SO = REF_DIR / "libsrsref.so"

CRC_NONE, CRC16, CRC24A, CRC24B = 0, 1, 2, 3
CRC24C, CRC11, CRC6 = 4, 5, 6
SCALE_X86, SCALE_GENERIC, SCALE_NEON = 0, 1, 2
MAX_CB_SIZE = 66 * 384
MAX_CB_BYTES = 22 * 384 // 8

_c_int = ctypes.c_int
_vp = ctypes.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def build_oracle():
    """Compile the C restatement (gcc only, a second). One process at a time: several ranks may load the oracle at once."""
    import fcntl
    (HERE / "_ref").mkdir(exist_ok=True)
    with open(HERE / "_ref" / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        subprocess.check_call(["make", "-s", "-C", str(HERE), "oracle"])


def build_reference():
    """Compile the reference where /root/reference exists (about 20 s on 8 cores)."""
    subprocess.check_call(["make", "-s", "-j", str(os.cpu_count() or 4), "-C", str(HERE), "ref"])


class CbMeta(ctypes.Structure):
    _fields_ = [(n, _c_int) for n in ("Z", "full_length", "rm_length", "nof_filler", "cw_offset", "nof_crc_bits")]


class _Harq(ctypes.Structure):
    _fields_ = [("nof_cb", _c_int), ("soft", _vp), ("data", _vp), ("crc_ok", _vp)]


class _PuschCfg(ctypes.Structure):
    _fields_ = [(n, _c_int) for n in ("bg", "rv", "qm", "nref", "nof_layers", "max_iter", "use_early_stop", "new_data")]


class Harq:
    """HARQ buffer of one (rnti, harq_id) for Oracle.pusch_decode."""

    def __init__(self, nof_cb, fill=0):
        self.nof_cb = nof_cb
        self.soft = np.full((nof_cb, MAX_CB_SIZE), fill, np.int8)
        self.data = np.zeros((nof_cb, MAX_CB_BYTES), np.uint8)
        self.crc_ok = np.zeros(nof_cb, np.uint8)

    def _c(self):
        return _Harq(self.nof_cb, _ptr(self.soft), _ptr(self.data), _ptr(self.crc_ok))


ULSCH_CFG_FIELDS = ("qm", "nof_layers", "nof_prb", "start_symbol_index", "nof_symbols", "nof_harq_ack_rvd", "dmrs_type",
                    "dmrs_symbol_mask", "nof_cdm_groups_without_data", "nof_harq_ack_bits", "nof_enc_harq_ack_bits",
                    "nof_csi_part1_bits", "nof_enc_csi_part1_bits", "nof_csi_part2_bits", "nof_enc_csi_part2_bits")


def ulsch_cfg_array(cfg):
    """orc_ulsch_cfg as an int32 array (the struct is 15 ints)."""
    return np.array([int(cfg.get(k, 0)) for k in ULSCH_CFG_FIELDS], np.int32)


def _oracle_stale():
    """The shared object is missing or older than the C source it is compiled from."""
    if not ORACLE_SO.exists():
        return True
    t = ORACLE_SO.stat().st_mtime
    return any((HERE / f).stat().st_mtime > t for f in ("pusch_oracle.c", "pusch_oracle.h", "bg_tables.inc"))


class Oracle:
    def __init__(self):
        if _oracle_stale():
            build_oracle()
        L = ctypes.CDLL(str(ORACLE_SO))
        L.orc_crc.restype = ctypes.c_uint32
        L.orc_crc.argtypes = [_c_int, _vp, _c_int]
        L.orc_rate_dematch.restype = None
        L.orc_rate_dematch.argtypes = [_vp, _c_int, _vp] + [_c_int] * 7
        L.orc_ldpc_decode.restype = _c_int
        L.orc_ldpc_decode.argtypes = [_c_int, _c_int, _vp] + [_c_int] * 5 + [_vp, _vp]
        L.orc_cb_decode.restype = _c_int
        L.orc_cb_decode.argtypes = [_vp, _c_int, _vp] + [_c_int] * 11 + [_vp]
        L.orc_segment_rx.restype = _c_int
        L.orc_segment_rx.argtypes = [_c_int] * 5 + [_vp]
        L.orc_pusch_decode.restype = None
        L.orc_pusch_decode.argtypes = [_vp, _vp, _c_int, _c_int, _vp, _c_int, _c_int, _vp, _vp]
        L.orc_ldpc_encode.restype = None
        L.orc_ldpc_encode.argtypes = [_c_int, _c_int, _vp, _vp]
        L.orc_rate_match.restype = None
        L.orc_rate_match.argtypes = [_vp] + [_c_int] * 6 + [_vp]
        L.orc_tb_encode.restype = _c_int
        L.orc_tb_encode.argtypes = [_vp] + [_c_int] * 7 + [_vp]
        L.orc_bench_cb_batch.restype = ctypes.c_double
        L.orc_bench_cb_batch.argtypes = [_c_int, _vp] + [_c_int] * 9 + [_vp]
        L.orc_prg_bits.restype = None
        L.orc_prg_bits.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, _vp]
        L.orc_revert_scrambling.restype = None
        L.orc_revert_scrambling.argtypes = [_vp, _vp, _vp, ctypes.c_uint32]
        L.orc_ulsch_codeword_length.restype = ctypes.c_uint32
        L.orc_ulsch_codeword_length.argtypes = [_vp]
        L.orc_ulsch_demux.restype = _c_int
        L.orc_ulsch_demux.argtypes = [_vp, _vp, _vp, ctypes.c_uint32, _vp, _vp, _vp, _vp, _vp]
        L.orc_demodulate_soft.restype = None
        L.orc_demodulate_soft.argtypes = [_vp, _vp, _vp, ctypes.c_uint32, _c_int, _c_int]
        self.L = L

    # -- soft demapper --------------------------------------------------------------------------------------------------
    def demodulate_soft(self, symbols, noise_vars, mod, simd=True):
        """One demodulate_soft call. symbols: complex64[n]; noise_vars: float32[n]; mod: 0 (pi/2-BPSK), 1, 2, 4, 6, 8."""
        symbols = np.ascontiguousarray(symbols, np.complex64)
        noise_vars = np.ascontiguousarray(noise_vars, np.float32)
        out = np.zeros(symbols.size * max(mod, 1), np.int8)
        self.L.orc_demodulate_soft(_ptr(out), _ptr(symbols), _ptr(noise_vars), symbols.size, mod, 1 if simd else 0)
        return out

    # -- codeword front end ---------------------------------------------------------------------------------------------
    def prg_bits(self, c_init, offset, n):
        bits = np.zeros(n, np.uint8)
        self.L.orc_prg_bits(c_init, offset, n, _ptr(bits))
        return bits

    def revert_scrambling(self, llrs, seq_bits):
        llrs = np.ascontiguousarray(llrs, np.int8)
        seq_bits = np.ascontiguousarray(seq_bits, np.uint8)
        out = np.zeros_like(llrs)
        self.L.orc_revert_scrambling(_ptr(out), _ptr(llrs), _ptr(seq_bits), llrs.size)
        return out

    def ulsch_codeword_length(self, cfg):
        c = ulsch_cfg_array(cfg)
        return int(self.L.orc_ulsch_codeword_length(_ptr(c)))

    def ulsch_demux(self, cfg, llrs, seq_bits):
        """cfg: dict with the fields of orc_ulsch_cfg. Returns (status, [sch, harq_ack, csi_part1, csi_part2])."""
        c = ulsch_cfg_array(cfg)
        llrs = np.ascontiguousarray(llrs, np.int8)
        seq_bits = np.ascontiguousarray(seq_bits, np.uint8)
        outs = [np.zeros(llrs.size, np.int8) for _ in range(4)]
        n_out = np.zeros(4, np.uint32)
        rc = self.L.orc_ulsch_demux(_ptr(c), _ptr(llrs), _ptr(seq_bits), llrs.size, _ptr(outs[0]), _ptr(outs[1]),
                                    _ptr(outs[2]), _ptr(outs[3]), _ptr(n_out))
        return rc, [o[:n] for o, n in zip(outs, n_out)]

    def crc(self, kind, packed, nbits):
        packed = np.ascontiguousarray(packed, np.uint8)
        return int(self.L.orc_crc(kind, _ptr(packed), nbits))

    def rate_dematch(self, out, llr, new_data, rv, qm, nref=0, nof_filler=0, simd_width=64):
        assert out.dtype == np.int8 and out.flags.c_contiguous
        llr = np.ascontiguousarray(llr, np.int8)
        self.L.orc_rate_dematch(_ptr(out), out.size, _ptr(llr), llr.size, int(new_data), rv, qm, nref, nof_filler,
                                simd_width)
        return out

    def ldpc_decode(self, bg, Z, llr, nof_filler=0, crc_kind=CRC_NONE, max_iter=6, scale=SCALE_X86, want_soft=False):
        llr = np.ascontiguousarray(llr, np.int8)
        K = (22 if bg == 1 else 10) * Z
        out = np.zeros((K + 7) // 8, np.uint8)
        soft = np.zeros((68 if bg == 1 else 52) * Z, np.int8) if want_soft else None
        it = self.L.orc_ldpc_decode(bg, Z, _ptr(llr), llr.size, nof_filler, crc_kind, max_iter, scale, _ptr(out),
                                    _ptr(soft))
        return (it, out, soft) if want_soft else (it, out)

    def cb_decode(self, rm_buffer, llr, new_data, rv, qm, nref, nof_filler, crc_kind, early_stop, max_iter,
                  scale=SCALE_X86, simd_width=64):
        llr = np.ascontiguousarray(llr, np.int8)
        N = rm_buffer.size
        K = N // 3 if N % 66 == 0 else N // 5
        out = np.zeros((K + 7) // 8, np.uint8)
        it = self.L.orc_cb_decode(_ptr(rm_buffer), N, _ptr(llr), llr.size, int(new_data), rv, qm, nref, nof_filler,
                                  crc_kind, int(early_stop), max_iter, scale, simd_width, _ptr(out))
        return it, out

    def segment_rx(self, tbs_bits, bg, qm, nof_layers, n_llr):
        meta = (CbMeta * 256)()
        n = self.L.orc_segment_rx(tbs_bits, bg, qm, nof_layers, n_llr, ctypes.byref(meta))
        return [meta[i] for i in range(n)]

    def pusch_decode(self, harq, llrs, tb_bytes, bg, rv, qm, nref, nof_layers, max_iter, early_stop, new_data,
                     scale=SCALE_X86, simd_width=64):
        llrs = np.ascontiguousarray(llrs, np.int8)
        cfg = _PuschCfg(bg, rv, qm, nref, nof_layers, max_iter, int(early_stop), int(new_data))
        tb = np.zeros(tb_bytes, np.uint8)
        stats = np.zeros(6, np.int32)
        h = harq._c()
        self.L.orc_pusch_decode(ctypes.byref(h), _ptr(llrs), llrs.size, tb_bytes, ctypes.byref(cfg), scale, simd_width,
                                _ptr(tb), _ptr(stats))
        return tb, stats

    def ldpc_encode(self, bg, Z, msg_bits):
        msg_bits = np.ascontiguousarray(msg_bits, np.uint8)
        cw = np.zeros((66 if bg == 1 else 50) * Z, np.uint8)
        self.L.orc_ldpc_encode(bg, Z, _ptr(msg_bits), _ptr(cw))
        return cw

    def rate_match(self, cw, E, rv, qm, nref=0, nof_filler=0):
        cw = np.ascontiguousarray(cw, np.uint8)
        out = np.zeros(E, np.uint8)
        self.L.orc_rate_match(_ptr(cw), cw.size, E, rv, qm, nref, nof_filler, _ptr(out))
        return out

    def tb_encode(self, tb, bg, rv, qm, nref, nof_layers, n_llr):
        tb = np.ascontiguousarray(tb, np.uint8)
        cw = np.zeros(n_llr, np.uint8)
        n = self.L.orc_tb_encode(_ptr(tb), tb.size, bg, rv, qm, nref, nof_layers, n_llr, _ptr(cw))
        return cw, n

    def bench_cb_batch(self, llrs, E, N, rv, qm, nref, nof_filler, crc_kind, early_stop, max_iter):
        llrs = np.ascontiguousarray(llrs, np.int8)
        n_cb = llrs.size // E
        iters = np.zeros(n_cb, np.int32)
        sec = self.L.orc_bench_cb_batch(n_cb, _ptr(llrs), E, N, rv, qm, nref, nof_filler, crc_kind, int(early_stop),
                                        max_iter, _ptr(iters))
        return sec, iters


class Reference:
    """The compiled reference (oracle/_ref/libsrsref.so)."""

    @staticmethod
    def available():
        return REFERENCE_SO.exists()

    def __init__(self, variant="auto"):
        L = ctypes.CDLL(str(REFERENCE_SO))
        L.ref_tools_create.restype = _vp
        L.ref_tools_create.argtypes = [ctypes.c_char_p]
        L.ref_tools_destroy.argtypes = [_vp]
        L.ref_ldpc_decode.restype = _c_int
        L.ref_ldpc_decode.argtypes = [_vp, _c_int, _c_int, _vp] + [_c_int] * 4 + [_vp]
        L.ref_rate_dematch.restype = None
        L.ref_rate_dematch.argtypes = [_vp, _vp, _c_int, _vp] + [_c_int] * 6
        L.ref_crc.restype = ctypes.c_uint
        L.ref_crc.argtypes = [_vp, _c_int, _vp, _c_int]
        L.ref_cb_decode.restype = _c_int
        L.ref_cb_decode.argtypes = [_vp, _vp, _c_int, _vp] + [_c_int] * 9 + [_vp]
        L.ref_tb_encode.restype = _c_int
        L.ref_tb_encode.argtypes = [_vp] + [_c_int] * 7 + [_vp]
        L.ref_ldpc_encode.restype = None
        L.ref_ldpc_encode.argtypes = [_c_int, _c_int, _vp, _vp, _c_int]
        L.ref_segment_rx.restype = _c_int
        L.ref_segment_rx.argtypes = [_c_int] * 7 + [_vp]
        L.ref_prg_bits.restype = None
        L.ref_prg_bits.argtypes = [ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, _vp]
        L.ref_ulsch_demux.restype = _c_int
        L.ref_ulsch_demux.argtypes = [_vp, _vp, _vp, ctypes.c_uint, ctypes.c_uint, _vp, _vp, _vp, _vp, _vp]
        L.ref_pusch_tbs.restype = None
        L.ref_pusch_tbs.argtypes = [_c_int] * 7 + [_vp]
        L.ref_demodulate_soft.restype = None
        L.ref_demodulate_soft.argtypes = [_vp, _vp, _vp, ctypes.c_uint, _c_int]
        L.ref_pusch_create.restype = _vp
        L.ref_pusch_create.argtypes = [ctypes.c_char_p, _c_int]
        L.ref_pusch_destroy.argtypes = [_vp]
        L.ref_pusch_fill_soft.argtypes = [_vp, _c_int]
        L.ref_pusch_decode.restype = None
        L.ref_pusch_decode.argtypes = [_vp, _vp] + [_c_int] * 11 + [_vp, _vp]
        L.ref_pusch_get_cb.restype = _c_int
        L.ref_pusch_get_cb.argtypes = [_vp, _c_int, _vp, _c_int]
        L.ref_bench_cb_batch.restype = ctypes.c_double
        L.ref_bench_cb_batch.argtypes = [ctypes.c_char_p] + [_c_int] * 3 + [_vp] + [_c_int] * 9 + [_vp, _vp]
        self.L = L
        self.variant = variant
        self.h = L.ref_tools_create(variant.encode())
        if not self.h:
            raise RuntimeError(f"reference variant {variant!r} not available on this CPU")

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_tools_destroy(self.h)
            self.h = None

    def auto_variant(self):
        return {3: "avx512", 2: "avx2", 1: "generic"}[self.L.ref_auto_variant()]

    def crc(self, kind, packed, nbits):
        packed = np.ascontiguousarray(packed, np.uint8)
        return int(self.L.ref_crc(self.h, kind, _ptr(packed), nbits))

    def rate_dematch(self, out, llr, new_data, rv, qm, nref=0, nof_filler=0):
        llr = np.ascontiguousarray(llr, np.int8)
        self.L.ref_rate_dematch(self.h, _ptr(out), out.size, _ptr(llr), llr.size, int(new_data), rv, qm, nref,
                                nof_filler)
        return out

    def ldpc_decode(self, bg, Z, llr, nof_filler=0, crc_kind=CRC_NONE, max_iter=6):
        llr = np.ascontiguousarray(llr, np.int8)
        K = (22 if bg == 1 else 10) * Z
        out = np.zeros((K + 7) // 8, np.uint8)
        it = self.L.ref_ldpc_decode(self.h, bg, Z, _ptr(llr), llr.size, nof_filler, crc_kind, max_iter, _ptr(out))
        return it, out

    def cb_decode(self, rm_buffer, llr, new_data, rv, qm, nref, nof_filler, crc_kind, early_stop, max_iter):
        llr = np.ascontiguousarray(llr, np.int8)
        N = rm_buffer.size
        K = N // 3 if N % 66 == 0 else N // 5
        out = np.zeros((K + 7) // 8, np.uint8)
        it = self.L.ref_cb_decode(self.h, _ptr(rm_buffer), N, _ptr(llr), llr.size, int(new_data), rv, qm, nref,
                                  nof_filler, crc_kind, int(early_stop), max_iter, _ptr(out))
        return it, out

    def ldpc_encode(self, bg, Z, msg_bits, n_out=None):
        msg_bits = np.ascontiguousarray(msg_bits, np.uint8)
        n_out = n_out or (66 if bg == 1 else 50) * Z
        cw = np.zeros(n_out, np.uint8)
        self.L.ref_ldpc_encode(bg, Z, _ptr(msg_bits), _ptr(cw), n_out)
        return cw

    def tb_encode(self, tb, bg, rv, qm, nref, nof_layers, n_llr):
        tb = np.ascontiguousarray(tb, np.uint8)
        cw = np.zeros(n_llr, np.uint8)
        n = self.L.ref_tb_encode(_ptr(tb), tb.size, bg, rv, qm, nref, nof_layers, n_llr // qm, _ptr(cw))
        return cw, n

    def segment_rx(self, tbs_bits, bg, rv, qm, nref, nof_layers, n_llr):
        meta = np.zeros(256 * 6, np.int32)
        n = self.L.ref_segment_rx(tbs_bits, bg, rv, qm, nref, nof_layers, n_llr, _ptr(meta))
        return meta[:n * 6].reshape(n, 6)

    def prg_bits(self, c_init, offset, n):
        bits = np.zeros(n, np.uint8)
        self.L.ref_prg_bits(c_init, offset, n, _ptr(bits))
        return bits

    def pusch_tbs(self, table, mcs, nof_symb_sh, nof_dmrs_prb, nof_oh_prb, nof_layers, n_prb):
        """(tbs bits, base graph, bits per symbol, target code rate x 1024) from the reference's pusch_mcs_get_config,
        tbs_calculator_calculate and get_ldpc_base_graph. table: "qam64" / "qam256"."""
        out = np.zeros(4, np.int32)
        self.L.ref_pusch_tbs(0 if table == "qam64" else 1, mcs, nof_symb_sh, nof_dmrs_prb, nof_oh_prb, nof_layers, n_prb,
                             _ptr(out))
        return int(out[0]), int(out[1]), int(out[2]), out[3] / 2.0

    def demodulate_soft(self, symbols, noise_vars, mod):
        """demodulation_mapper_impl::demodulate_soft of the compiled reference (one call)."""
        symbols = np.ascontiguousarray(symbols, np.complex64)
        noise_vars = np.ascontiguousarray(noise_vars, np.float32)
        out = np.zeros(symbols.size * max(mod, 1), np.int8)
        self.L.ref_demodulate_soft(_ptr(out), _ptr(symbols), _ptr(noise_vars), symbols.size, mod)
        return out

    def ulsch_demux(self, cfg, llrs, seq_bits, max_block_re=0):
        """ulsch_demultiplex_impl fed like pusch_demodulator_impl does. Returns (status, [sch, ack, csi1, csi2])."""
        c = ulsch_cfg_array(cfg)
        llrs = np.ascontiguousarray(llrs, np.int8)
        seq_bits = np.ascontiguousarray(seq_bits, np.uint8)
        outs = [np.zeros(llrs.size, np.int8) for _ in range(4)]
        n_out = np.zeros(4, np.uint32)
        rc = self.L.ref_ulsch_demux(_ptr(c), _ptr(llrs), _ptr(seq_bits), llrs.size, max_block_re, _ptr(outs[0]),
                                    _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), _ptr(n_out))
        return rc, [o[:n] for o, n in zip(outs, n_out)]

    def bench_cb_batch(self, llrs, E, N, rv, qm, nref, nof_filler, crc_kind, early_stop, max_iter, threads=1,
                       repeats=1, want_bits=False):
        llrs = np.ascontiguousarray(llrs, np.int8)
        n_cb = llrs.size // E
        iters = np.zeros(n_cb, np.int32)
        K = N // 3 if N % 66 == 0 else N // 5
        bits = np.zeros((n_cb, (K + 7) // 8), np.uint8) if want_bits else None
        sec = self.L.ref_bench_cb_batch(self.variant.encode(), threads, repeats, n_cb, _ptr(llrs), E, N, rv, qm, nref,
                                        nof_filler, crc_kind, int(early_stop), max_iter, _ptr(iters), _ptr(bits))
        return (sec, iters, bits) if want_bits else (sec, iters)


class ReferencePusch:
    """pusch_decoder_impl of the reference with a driver-owned HARQ buffer."""

    def __init__(self, nof_cb, variant="auto", fill=None):
        self.ref = Reference(variant)
        self.L = self.ref.L
        self.h = self.L.ref_pusch_create(variant.encode(), nof_cb)
        self.nof_cb = nof_cb
        if fill is not None:
            self.L.ref_pusch_fill_soft(self.h, fill)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_pusch_destroy(self.h)
            self.h = None

    def decode(self, llrs, tb_bytes, bg, rv, qm, nref, nof_layers, max_iter, early_stop, new_data, reset_crcs=False):
        llrs = np.ascontiguousarray(llrs, np.int8)
        tb = np.zeros(tb_bytes, np.uint8)
        stats = np.zeros(6, np.int32)
        self.L.ref_pusch_decode(self.h, _ptr(llrs), llrs.size, tb_bytes, bg, rv, qm, nref, nof_layers, max_iter,
                                int(early_stop), int(new_data), int(reset_crcs), _ptr(tb), _ptr(stats))
        return tb, stats

    def get_cb(self, cb, n):
        soft = np.zeros(n, np.int8)
        ok = self.L.ref_pusch_get_cb(self.h, cb, _ptr(soft), n)
        return soft, bool(ok)
