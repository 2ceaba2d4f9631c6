"""ctypes binding of include/pusch_dec_cuda.h (the C-ABI of the CUDA library).

The library is loaded from the package directory (in-tree build). There is no fallback: a missing library or a missing
GPU raises.
"""
import ctypes
import os
from pathlib import Path

import numpy as np

from . import build as _build

PDC_OK = 0
PDC_ERR_INVALID, PDC_ERR_CUDA, PDC_ERR_CAPACITY, PDC_ERR_NO_DEVICE = -1, -2, -3, -4
PDC_MAX_CB_SOFT = 25344
PDC_MAX_CB_BYTES = 1056
CRC_NONE, CRC16, CRC24A, CRC24B = 0, 1, 2, 3
CRC24C, CRC11, CRC6 = 4, 5, 6  # pdc_crc only
SCALE_X86, SCALE_GENERIC, SCALE_NEON = 0, 1, 2
CB_NEW_DATA, CB_EARLY_STOP, CB_DECODE, CB_DEMATCH = 1, 2, 4, 8
LAUNCH_HIGH_RATE = 0x200  # pdc_launch_device flags_union hint: see PDC_LAUNCH_HIGH_RATE

_u8, _u16, _u32, _i32 = ctypes.c_uint8, ctypes.c_uint16, ctypes.c_uint32, ctypes.c_int32
_vp = ctypes.c_void_p


class CbDesc(ctypes.Structure):
    _fields_ = [("llr_offset", _u32), ("rm_length", _u32), ("harq_id", _u32), ("nref", _u32), ("lifting_size", _u16),
                ("nof_filler", _u16), ("base_graph", _u8), ("qm", _u8), ("rv", _u8), ("crc_kind", _u8),
                ("max_iter", _u8), ("flags", _u8), ("tb_index", _u16)]


class CbResult(ctypes.Structure):
    _fields_ = [("crc_ok", _u8), ("iters", _u8), ("status", _u8), ("nlayers", _u8)]


class TbDesc(ctypes.Structure):
    _fields_ = [("first_cb", _u32), ("nof_cb", _u32), ("tbs_bits", _u32), ("out_offset", _u32), ("reserved", _u32)]


class TbResult(ctypes.Structure):
    _fields_ = [("tb_crc_ok", _u8), ("all_cb_ok", _u8), ("reserved", _u16)]


class Config(ctypes.Structure):
    _fields_ = [("device", _i32), ("max_cbs", _u32), ("max_llrs", _u32), ("harq_entries", _u32), ("max_tbs", _u32),
                ("max_tb_bytes", _u32), ("scale_mode", _i32), ("combine_simd_width", _i32), ("nof_streams", _u32),
                ("demod_mode", _i32)]


CB_DESC_DTYPE = np.dtype([("llr_offset", "<u4"), ("rm_length", "<u4"), ("harq_id", "<u4"), ("nref", "<u4"),
                          ("lifting_size", "<u2"), ("nof_filler", "<u2"), ("base_graph", "u1"), ("qm", "u1"),
                          ("rv", "u1"), ("crc_kind", "u1"), ("max_iter", "u1"), ("flags", "u1"), ("tb_index", "<u2")])
CB_RESULT_DTYPE = np.dtype([("crc_ok", "u1"), ("iters", "u1"), ("status", "u1"), ("nlayers", "u1")])
TB_DESC_DTYPE = np.dtype([("first_cb", "<u4"), ("nof_cb", "<u4"), ("tbs_bits", "<u4"), ("out_offset", "<u4"),
                          ("reserved", "<u4")])
TB_RESULT_DTYPE = np.dtype([("tb_crc_ok", "u1"), ("all_cb_ok", "u1"), ("reserved", "<u2")])
CW_SCRAMBLED = 1
CW_DEFER_DESCRAMBLING = 2
CW_PLAIN_BPSK = 16
# soft demapper: modulation codes (bits per symbol, 0 = pi/2-BPSK), reference build reproduced, pdc_demod_call
MOD_PI_2_BPSK, MOD_BPSK, MOD_QPSK, MOD_QAM16, MOD_QAM64, MOD_QAM256 = 0, 1, 2, 4, 6, 8
DEMOD_X86, DEMOD_SCALAR = 0, 1
# pdc_enc_desc (downlink twin: LDPC encoding + rate matching)
ENC_DESC_DTYPE = np.dtype([("msg_offset", "<u4"), ("out_offset", "<u4"), ("rm_length", "<u4"), ("nref", "<u4"),
                           ("lifting_size", "<u2"), ("nof_filler", "<u2"), ("base_graph", "u1"), ("qm", "u1"),
                           ("rv", "u1"), ("flags", "u1")])
ENC_PACKED = 1
DEMOD_CALL_DTYPE = np.dtype([("sym_offset", "<u4"), ("n_sym", "<u4"), ("llr_offset", "<u4"), ("modulation", "<u4")])
# pdc_cw_desc / pdc_cw_result (codeword front end)
CW_DESC_DTYPE = np.dtype([("in_offset", "<u4"), ("sch_offset", "<u4"), ("uci_offset", "<u4"), ("c_init", "<u4"),
                          ("flags", "<u4"), ("qm", "u1"), ("nof_layers", "u1"), ("start_symbol_index", "u1"),
                          ("nof_symbols", "u1"), ("dmrs_type", "u1"), ("nof_cdm_groups_without_data", "u1"),
                          ("nof_prb", "<u2"), ("dmrs_symbol_mask", "<u2"), ("reserved", "<u2"),
                          ("nof_harq_ack_rvd", "<u4"), ("nof_harq_ack_bits", "<u4"), ("nof_enc_harq_ack_bits", "<u4"),
                          ("nof_csi_part1_bits", "<u4"), ("nof_enc_csi_part1_bits", "<u4"),
                          ("nof_csi_part2_bits", "<u4"), ("nof_enc_csi_part2_bits", "<u4")])
CW_RESULT_DTYPE = np.dtype([("n_sch", "<u4"), ("n_harq_ack", "<u4"), ("n_csi_part1", "<u4"), ("n_csi_part2", "<u4")])
assert CW_DESC_DTYPE.itemsize == 60 and CW_RESULT_DTYPE.itemsize == 16 and ENC_DESC_DTYPE.itemsize == 24
assert CB_DESC_DTYPE.itemsize == ctypes.sizeof(CbDesc) == 28
assert TB_DESC_DTYPE.itemsize == ctypes.sizeof(TbDesc) == 20

EXPORTS = ["pdc_default_config", "pdc_create", "pdc_destroy", "pdc_last_error", "pdc_device_info", "pdc_launch_count",
           "pdc_measure_int_peak", "pdc_debug_canaries_ok", "pdc_host_alloc", "pdc_host_alloc_input", "pdc_host_free", "pdc_submit", "pdc_wait", "pdc_poll", "pdc_launch_device",
           "pdc_harq_read", "pdc_harq_write", "pdc_harq_free", "pdc_harq_device_ptr", "pdc_ldpc_decode",
           "pdc_rate_dematch", "pdc_crc", "pdc_submit_codewords", "pdc_ulsch_demux", "pdc_scrambling_sequence",
           "pdc_launch_codewords_device", "pdc_submit_symbols", "pdc_demodulate_soft", "pdc_launch_demod_device",
           "pdc_encode", "pdc_launch_encode_device", "pdc_ldpc_encode"]

_lib = None


class PdcError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"pusch_dec_cuda error {code}: {text}")
        self.code = code


def library_path() -> Path:
    return _build.LIB


def load():
    """Loads (building first if the sources are newer) the in-tree CUDA library. Raises if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    # PDC_LIBRARY: load a prebuilt library from elsewhere (deployment, A/B measurements) instead of the in-tree build.
    path = os.environ.get("PDC_LIBRARY") or _build.build()
    L = ctypes.CDLL(str(path))
    L.pdc_default_config.argtypes = [ctypes.POINTER(Config)]
    L.pdc_default_config.restype = None
    L.pdc_create.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(_vp)]
    L.pdc_destroy.argtypes = [_vp]
    L.pdc_destroy.restype = None
    L.pdc_last_error.restype = ctypes.c_char_p
    L.pdc_device_info.argtypes = [_vp] + [ctypes.POINTER(ctypes.c_int)] * 3
    L.pdc_launch_count.argtypes = [_vp]
    L.pdc_launch_count.restype = ctypes.c_uint64
    L.pdc_measure_int_peak.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.pdc_debug_canaries_ok.argtypes = [_vp]
    L.pdc_debug_canaries_ok.restype = ctypes.c_int
    L.pdc_host_alloc.argtypes = [ctypes.c_size_t]
    L.pdc_host_alloc.restype = _vp
    L.pdc_host_alloc_input.argtypes = [ctypes.c_size_t]
    L.pdc_host_alloc_input.restype = _vp
    L.pdc_host_free.argtypes = [_vp]
    L.pdc_host_free.restype = None
    L.pdc_submit.argtypes = [_vp, _u32, _vp, _u32, _vp, ctypes.c_size_t, _vp, _u32, _vp, _vp, _vp, _vp]
    L.pdc_wait.argtypes = [_vp, _u32]
    L.pdc_poll.argtypes = [_vp, _u32, ctypes.POINTER(ctypes.c_int)]
    L.pdc_launch_device.argtypes = [_vp, _vp, _u32, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _u32, _u32, ctypes.c_int, _vp]
    L.pdc_harq_read.argtypes = [_vp, _u32, _vp, _u32]
    L.pdc_harq_write.argtypes = [_vp, _u32, _vp, _u32]
    L.pdc_harq_free.argtypes = [_vp, _u32]
    L.pdc_harq_device_ptr.argtypes = [_vp]
    L.pdc_harq_device_ptr.restype = _vp
    L.pdc_ldpc_decode.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp, _u32, _u32, ctypes.c_int, ctypes.c_int, _vp,
                                  ctypes.POINTER(ctypes.c_int)]
    L.pdc_rate_dematch.argtypes = [_vp, _vp, _u32, _vp, _u32, ctypes.c_int, ctypes.c_int, ctypes.c_int, _u32, _u32]
    L.pdc_crc.argtypes = [_vp, ctypes.c_int, _vp, _u32, ctypes.POINTER(_u32)]
    L.pdc_submit_codewords.argtypes = [_vp, _u32, _vp, _u32, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp]
    L.pdc_ulsch_demux.argtypes = [_vp, _vp, _u32, _vp, ctypes.c_size_t, _vp, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t,
                                  _vp]
    L.pdc_scrambling_sequence.argtypes = [_vp, _u32, _u32, _u32, _vp]
    L.pdc_launch_codewords_device.argtypes = [_vp, _vp, _u32, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp,
                                              ctypes.c_size_t, _vp, _vp]
    L.pdc_submit_symbols.argtypes = [_vp, _u32, _vp, _u32, _vp, _vp, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp]
    L.pdc_demodulate_soft.argtypes = [_vp, _vp, _vp, _vp, _u32, ctypes.c_int]
    L.pdc_launch_demod_device.argtypes = [_vp, _vp, _u32, _vp, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t, _vp]
    L.pdc_encode.argtypes = [_vp, _vp, _u32, _vp, ctypes.c_size_t, _vp, ctypes.c_size_t]
    L.pdc_launch_encode_device.argtypes = [_vp, _vp, _u32, _vp, _vp, ctypes.c_size_t, _u32, ctypes.c_int, _vp]
    L.pdc_ldpc_encode.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp]
    _lib = L
    return L


def check(rc):
    if rc != PDC_OK:
        raise PdcError(rc, load().pdc_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


class PinnedBuffer:
    """Pinned host memory from pdc_host_alloc exposed as a numpy array."""

    def __init__(self, nbytes, dtype=np.int8, input_only=False):
        """input_only: write-combined memory (pdc_host_alloc_input) - fill it once, never read it on the CPU."""
        self._L = load()
        self.ptr = (self._L.pdc_host_alloc_input if input_only else self._L.pdc_host_alloc)(nbytes)
        if not self.ptr:
            raise PdcError(PDC_ERR_CUDA, "pdc_host_alloc failed")
        n = nbytes // np.dtype(dtype).itemsize
        self.array = np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(nbytes,))
        self.array = self.array.view(dtype)[:n]

    def __del__(self):
        if getattr(self, "ptr", None):
            self._L.pdc_host_free(self.ptr)
            self.ptr = None


class Context:
    """pdc_ctx: one per GPU. Owns the device HARQ arena, the streams ("queues") and the pinned staging."""

    def __init__(self, device=0, max_cbs=4096, max_llrs=None, harq_entries=4096, max_tbs=256, max_tb_bytes=4 << 20,
                 scale_mode=SCALE_X86, combine_simd_width=64, nof_streams=2, demod_mode=DEMOD_X86):
        L = load()
        cfg = Config()
        L.pdc_default_config(ctypes.byref(cfg))
        cfg.device = device
        cfg.max_cbs = max_cbs
        cfg.max_llrs = max_llrs if max_llrs is not None else max_cbs * PDC_MAX_CB_SOFT
        cfg.harq_entries = harq_entries
        cfg.max_tbs = max_tbs
        cfg.max_tb_bytes = max_tb_bytes
        cfg.scale_mode = scale_mode
        cfg.combine_simd_width = combine_simd_width
        cfg.nof_streams = nof_streams
        cfg.demod_mode = demod_mode
        self.cfg = cfg
        self._L = L
        self.h = _vp()
        check(L.pdc_create(ctypes.byref(cfg), ctypes.byref(self.h)))
        self._pending = {}
        self._pending_fe = {}

    def debug_canaries_ok(self):
        """1 / 0 on a -DPDC_DEBUG_BOUNDS build of the library (canaries intact / overwritten), -1 on a release build."""
        return int(self._L.pdc_debug_canaries_ok(self.h))

    def close(self):
        if getattr(self, "h", None):
            if os.environ.get("PDC_CHECK_CANARIES") == "1":
                # debug-bounds run of the GPU suite (tools/run_debug_bounds.sh): every context leaves with its canaries intact
                ok = self.debug_canaries_ok()
                self._L.pdc_destroy(self.h)
                self.h = None
                if ok != 1:
                    raise PdcError(PDC_ERR_CUDA, "device canaries overwritten (or not a PDC_DEBUG_BOUNDS build): %d" % ok)
                return
            self._L.pdc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def device_info(self):
        a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(self._L.pdc_device_info(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"sm_count": a.value, "cc": (b.value, c.value)}

    def launch_count(self):
        return int(self._L.pdc_launch_count(self.h))

    def measure_int_peak(self, mode=0):
        """32-bit integer lane-operations per second (mode 0: ALU pipe only, 1: ALU + FMA pipes)."""
        v = ctypes.c_double()
        check(self._L.pdc_measure_int_peak(self.h, mode, ctypes.byref(v)))
        return v.value

    # -- batched interface --------------------------------------------------------------------------------------------
    def submit(self, cbs, llrs, tbs=None, stream=0, want_bits=True, want_tb=True, out_bits=None, out_tb=None):
        """cbs: numpy array of CB_DESC_DTYPE; llrs: int8 array (ideally a PinnedBuffer view); tbs: TB_DESC_DTYPE.
        out_bits / out_tb: caller-owned uint8 output arrays (PinnedBuffer views are written by the GPU directly)."""
        cbs = np.ascontiguousarray(cbs, CB_DESC_DTYPE)
        # llrs = None: the LLRs are the UL-SCH soft bits submit_codewords() left on the device.
        assert llrs is None or (llrs.dtype == np.int8 and llrs.flags.c_contiguous)
        n_cb = cbs.size
        n_tb = 0 if tbs is None else tbs.size
        out = {
            "cbs": cbs,
            "llrs": llrs,
            "cb_results": np.zeros(n_cb, CB_RESULT_DTYPE),
            "cb_bits": (out_bits[:n_cb * PDC_MAX_CB_BYTES].reshape(n_cb, PDC_MAX_CB_BYTES) if out_bits is not None else
                        np.zeros((n_cb, PDC_MAX_CB_BYTES), np.uint8)) if want_bits else None,
            "tbs": None,
            "tb_results": None,
            "tb_bytes": None,
        }
        if n_tb:
            tbs = np.ascontiguousarray(tbs, TB_DESC_DTYPE)
            out["tbs"] = tbs
            out["tb_results"] = np.zeros(n_tb, TB_RESULT_DTYPE)
            if want_tb:
                need = int(max(t["out_offset"] + (int(t["tbs_bits"]) + 24 + 31) // 32 * 4 for t in tbs))
                out["tb_bytes"] = out_tb[:need] if out_tb is not None else np.zeros(need, np.uint8)
        check(self._L.pdc_submit(self.h, stream, _ptr(cbs), n_cb, _ptr(llrs), 0 if llrs is None else llrs.size,
                                 _ptr(out["tbs"]), n_tb,
                                 _ptr(out["cb_results"]), _ptr(out["cb_bits"]), _ptr(out["tb_results"]),
                                 _ptr(out["tb_bytes"])))
        self._pending[stream] = out
        return out

    def wait(self, stream=0):
        check(self._L.pdc_wait(self.h, stream))
        out = self._pending.pop(stream, None)
        fe = self._pending_fe.pop(stream, None)
        if fe is not None:
            if out is None:
                return fe
            out["codewords"] = fe
        return out

    # -- codeword front end ---------------------------------------------------------------------------------------------
    def submit_codewords(self, cws, raw_llrs, stream=0, out_uci=None):
        """pdc_submit_codewords: descrambling + UL-SCH demultiplexing of the batch's codewords on the device; follow it
        with submit(cbs, None, ...) on the same stream. cws: CW_DESC_DTYPE array; raw_llrs: int8 (ideally pinned)."""
        cws = np.ascontiguousarray(cws, CW_DESC_DTYPE)
        assert raw_llrs.dtype == np.int8 and raw_llrs.flags.c_contiguous
        res = np.zeros(cws.size, CW_RESULT_DTYPE)
        if out_uci is None:
            # UCI area: the codewords' UCI soft bits back to back from their uci_offset.
            need = int((cws["uci_offset"].astype(np.int64) + cws["nof_enc_harq_ack_bits"] + cws["nof_enc_csi_part1_bits"] +
                        cws["nof_enc_csi_part2_bits"]).max()) if cws.size else 0
            out_uci = np.zeros(max(1, need), np.int8)
        uci = out_uci
        check(self._L.pdc_submit_codewords(self.h, stream, _ptr(cws), cws.size, _ptr(raw_llrs), raw_llrs.size, _ptr(uci),
                                           uci.size, _ptr(res)))
        fe = {"cws": cws, "raw_llrs": raw_llrs, "cw_results": res, "uci": uci}
        self._pending_fe[stream] = fe
        return fe

    def submit_symbols(self, cws, sym_offsets, symbols, noise_vars, stream=0, out_uci=None):
        """pdc_submit_symbols: the front end fed by the equaliser's output - soft demapping (one demodulate_soft call per
        OFDM symbol of each codeword), descrambling and UL-SCH demultiplexing on the device; follow it with
        submit(cbs, None, ...) on the same stream. symbols: complex64 (ideally pinned); noise_vars: float32."""
        cws = np.ascontiguousarray(cws, CW_DESC_DTYPE)
        sym_offsets = np.ascontiguousarray(sym_offsets, np.uint32)
        assert symbols.dtype == np.complex64 and symbols.flags.c_contiguous
        assert noise_vars.dtype == np.float32 and noise_vars.flags.c_contiguous and noise_vars.size == symbols.size
        res = np.zeros(cws.size, CW_RESULT_DTYPE)
        if out_uci is None:
            need = int((cws["uci_offset"].astype(np.int64) + cws["nof_enc_harq_ack_bits"] + cws["nof_enc_csi_part1_bits"] +
                        cws["nof_enc_csi_part2_bits"]).max()) if cws.size else 0
            out_uci = np.zeros(max(1, need), np.int8)
        check(self._L.pdc_submit_symbols(self.h, stream, _ptr(cws), cws.size, _ptr(sym_offsets), _ptr(symbols),
                                         _ptr(noise_vars), symbols.size, _ptr(out_uci), out_uci.size, _ptr(res)))
        fe = {"cws": cws, "sym_offsets": sym_offsets, "symbols": symbols, "noise_vars": noise_vars, "cw_results": res,
              "uci": out_uci}
        self._pending_fe[stream] = fe
        return fe

    def demodulate_soft(self, symbols, noise_vars, modulation):
        """pdc_demodulate_soft: one demodulation_mapper::demodulate_soft call (synchronous, host buffers)."""
        symbols = np.ascontiguousarray(symbols, np.complex64)
        noise_vars = np.ascontiguousarray(noise_vars, np.float32)
        assert symbols.size == noise_vars.size
        out = np.zeros(symbols.size * max(int(modulation), 1), np.int8)
        check(self._L.pdc_demodulate_soft(self.h, _ptr(out), _ptr(symbols), _ptr(noise_vars), symbols.size,
                                          int(modulation)))
        return out

    def launch_demod_device(self, calls, d_symbols, d_noise_vars, n_sym, d_llrs, llr_capacity, cuda_stream=0):
        """pdc_launch_demod_device: a batch of demodulate_soft calls on device buffers, queued on the caller's stream."""
        calls = np.ascontiguousarray(calls, DEMOD_CALL_DTYPE)
        check(self._L.pdc_launch_demod_device(self.h, _ptr(calls), calls.size, d_symbols, d_noise_vars, n_sym, d_llrs,
                                              llr_capacity, cuda_stream or None))

    def launch_codewords_device(self, cws, d_raw, n_raw, d_sch, sch_capacity, d_uci=0, uci_capacity=0, cuda_stream=0):
        """pdc_launch_codewords_device: the front end on device buffers, queued on the caller's stream."""
        cws = np.ascontiguousarray(cws, CW_DESC_DTYPE)
        res = np.zeros(cws.size, CW_RESULT_DTYPE)
        check(self._L.pdc_launch_codewords_device(self.h, _ptr(cws), cws.size, d_raw, n_raw, d_sch, sch_capacity,
                                                  d_uci or None, uci_capacity, _ptr(res), cuda_stream or None))
        return res

    def ulsch_demux(self, cws, llrs, seq_bits_packed=None, sch_capacity=None, uci_capacity=None):
        """pdc_ulsch_demux (synchronous). Returns (cw_results, sch, uci)."""
        cws = np.ascontiguousarray(cws, CW_DESC_DTYPE)
        llrs = np.ascontiguousarray(llrs, np.int8)
        sch = np.zeros(sch_capacity if sch_capacity is not None else llrs.size + 4 * cws.size, np.int8)
        uci = np.zeros(uci_capacity if uci_capacity is not None else max(1, llrs.size), np.int8)
        res = np.zeros(cws.size, CW_RESULT_DTYPE)
        if seq_bits_packed is not None:
            seq_bits_packed = np.ascontiguousarray(seq_bits_packed, np.uint8)
        check(self._L.pdc_ulsch_demux(self.h, _ptr(cws), cws.size, _ptr(llrs), llrs.size, _ptr(seq_bits_packed), _ptr(sch),
                                      sch.size, _ptr(uci), uci.size, _ptr(res)))
        return res, sch, uci

    def scrambling_sequence(self, c_init, offset, n):
        """TS 38.211 5.2.1 sequence c(offset .. offset + n - 1), one bit per element."""
        packed = np.zeros((n + 7) // 8, np.uint8)
        check(self._L.pdc_scrambling_sequence(self.h, c_init, offset, n, _ptr(packed)))
        return np.unpackbits(packed)[:n]

    def poll(self, stream=0):
        d = ctypes.c_int()
        check(self._L.pdc_poll(self.h, stream, ctypes.byref(d)))
        return bool(d.value)

    def launch_device(self, d_cbs, n_cb, d_llrs, d_cb_results, d_cb_bits, max_lifting_size, flags_union, any_bg1,
                      cuda_stream=0, d_tbs=0, n_tb=0, d_tb_results=0, d_tb_bytes=0):
        check(self._L.pdc_launch_device(self.h, d_cbs, n_cb, d_llrs, d_tbs or None, n_tb, d_cb_results, d_cb_bits,
                                        d_tb_results or None, d_tb_bytes or None, max_lifting_size, flags_union,
                                        int(any_bg1), cuda_stream or None))

    # -- HARQ arena ---------------------------------------------------------------------------------------------------
    def harq_read(self, harq_id, n=PDC_MAX_CB_SOFT):
        soft = np.zeros(n, np.int8)
        check(self._L.pdc_harq_read(self.h, harq_id, _ptr(soft), n))
        return soft

    def harq_write(self, harq_id, soft):
        soft = np.ascontiguousarray(soft, np.int8)
        check(self._L.pdc_harq_write(self.h, harq_id, _ptr(soft), soft.size))

    def harq_free(self, harq_id):
        check(self._L.pdc_harq_free(self.h, harq_id))

    def harq_device_ptr(self):
        return self._L.pdc_harq_device_ptr(self.h)

    # -- single-codeblock synchronous calls -----------------------------------------------------------------------------
    def ldpc_decode(self, bg, Z, llrs, nof_filler=0, crc_kind=CRC_NONE, max_iter=6, out=None):
        llrs = np.ascontiguousarray(llrs, np.int8)
        K = (22 if bg == 1 else 10) * Z
        if out is None:
            out = np.zeros((K + 7) // 8, np.uint8)
        it = ctypes.c_int()
        check(self._L.pdc_ldpc_decode(self.h, bg, Z, _ptr(llrs), llrs.size, nof_filler, crc_kind, max_iter, _ptr(out),
                                      ctypes.byref(it)))
        return it.value, out

    def rate_dematch(self, buffer, llrs, new_data, rv, qm, nref=0, nof_filler=0):
        assert buffer.dtype == np.int8 and buffer.flags.c_contiguous
        llrs = np.ascontiguousarray(llrs, np.int8)
        check(self._L.pdc_rate_dematch(self.h, _ptr(buffer), buffer.size, _ptr(llrs), llrs.size, int(new_data), rv, qm,
                                       nref, nof_filler))
        return buffer

    # -- downlink twin ------------------------------------------------------------------------------------------------
    def encode(self, cbs, msgs_packed, out_capacity=None):
        """pdc_encode: LDPC encoding + rate matching of a batch of codeblocks. cbs: ENC_DESC_DTYPE array; msgs_packed: the
        message bits (packed, each codeblock at its msg_offset). Returns the rate-matched bits, one per byte (eight per
        byte, MSB first, for descriptors flagged ENC_PACKED)."""
        cbs = np.ascontiguousarray(cbs, ENC_DESC_DTYPE)
        msgs_packed = np.ascontiguousarray(msgs_packed, np.uint8)
        if out_capacity is None:
            n_out = np.where(cbs["flags"] & ENC_PACKED, (cbs["rm_length"].astype(np.int64) + 7) // 8, cbs["rm_length"])
            out_capacity = int((cbs["out_offset"].astype(np.int64) + n_out).max())
        out = np.zeros(out_capacity, np.uint8)
        check(self._L.pdc_encode(self.h, _ptr(cbs), cbs.size, _ptr(msgs_packed), msgs_packed.size, _ptr(out), out.size))
        return out

    def launch_encode_device(self, d_cbs, n_cb, d_msgs, d_out, out_capacity, max_lifting_size, any_bg1, cuda_stream=0):
        check(self._L.pdc_launch_encode_device(self.h, d_cbs, n_cb, d_msgs, d_out, out_capacity, max_lifting_size,
                                               int(any_bg1), cuda_stream or None))

    def ldpc_encode(self, bg, Z, msg_bits):
        """pdc_ldpc_encode: ldpc_encoder::encode of one codeblock. msg_bits: K bits (one per element). Returns the
        N = 66 Z / 50 Z codeblock bits, one per byte."""
        packed = np.packbits(np.ascontiguousarray(msg_bits, np.uint8))
        out = np.zeros((66 if bg == 1 else 50) * Z, np.uint8)
        check(self._L.pdc_ldpc_encode(self.h, bg, Z, _ptr(packed), _ptr(out)))
        return out

    def crc(self, crc_kind, packed, nbits):
        packed = np.ascontiguousarray(packed, np.uint8)
        c = _u32()
        check(self._L.pdc_crc(self.h, crc_kind, _ptr(packed), nbits, ctypes.byref(c)))
        return c.value
