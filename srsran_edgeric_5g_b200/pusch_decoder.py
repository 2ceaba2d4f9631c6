"""Batched PUSCH decoder - host-side mirror of the reference's pusch_decoder for the "cuda" variant.

Interface mirrored (same names and call order):
  pusch_decoder::new_data / set_nof_softbits     include/srsran/phy/upper/channel_processors/pusch/pusch_decoder.h:54-99
  pusch_decoder_buffer::on_new_softbits / on_end_softbits          .../pusch/pusch_decoder_buffer.h:34-58
  pusch_decoder_notifier::on_sch_data(pusch_decoder_result)        .../pusch/pusch_decoder_notifier.h:30-39
  rx_buffer / unique_rx_buffer / rx_buffer_pool                    include/srsran/phy/upper/rx_buffer.h:42-81
Behaviour mirrored: pusch_decoder_impl (lib/phy/upper/channel_processors/pusch/pusch_decoder_impl.cpp), in particular
  * CB CRC selection (select_crc :35-46), CRC flags reset on new data (:131-135),
  * codeblocks whose CRC flag is already set are combined but not decoded and push no statistic (:335-345),
  * statistics: iterations on success, nof_ldpc_iterations on failure (:357-363),
  * join: single CB -> TB CRC = CB CRC; several -> concatenate + CRC24A, reset all CB CRCs on mismatch (:384-450),
  * release() the buffer on success, unlock() otherwise (:432-436).

What differs by design: on_end_softbits() only queues the transport block; every TB queued by any decoder
of the same PuschDecoderBatch - all UEs and cells of a slot - is decoded by ONE batched GPU submission in flush(), and
the notifiers are called from flush(). The HARQ soft bits live in the device arena ("external soft bits").
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi
from .ldpc import BG1, compute_nof_codeblocks, segment_rx

MAX_BITS_CRC16 = 3824


@dataclass
class pusch_decoder_configuration:
    """pusch_decoder::configuration (pusch_decoder.h:58-76)."""
    base_graph: int = BG1
    rv: int = 0
    mod: int = 1  # bits per symbol
    Nref: int = 0
    nof_layers: int = 1
    nof_ldpc_iterations: int = 6
    use_early_stop: bool = True
    new_data: bool = True


class sample_statistics:
    """The part of sample_statistics<unsigned> (include/srsran/support/stats.h) the decoder result exposes."""

    def __init__(self):
        self._v: List[int] = []

    def update(self, v):
        self._v.append(int(v))

    def reset(self):
        self._v = []

    def get_nof_observations(self):
        return len(self._v)

    def get_min(self):
        return min(self._v)

    def get_max(self):
        return max(self._v)

    def get_mean(self):
        return sum(self._v) / len(self._v)


@dataclass
class pusch_decoder_result:
    """pusch_decoder_result.h:30-41."""
    tb_crc_ok: bool = False
    nof_codeblocks_total: int = 0
    ldpc_decoder_stats: sample_statistics = field(default_factory=sample_statistics)


class rx_buffer:
    """One (rnti, harq_id) HARQ buffer: CRC flags on the host, soft/message bits in the device arena."""

    def __init__(self, pool, key, absolute_ids):
        self._pool, self.key = pool, key
        self._ids = list(absolute_ids)
        self._crc = np.zeros(len(self._ids), bool)
        self.locked = False

    def get_nof_codeblocks(self):
        return len(self._ids)

    def reset_codeblocks_crc(self):
        self._crc[:] = False

    def get_codeblocks_crc(self):
        return self._crc

    def get_absolute_codeblock_id(self, codeblock_id):
        return self._ids[codeblock_id]

    # unique_rx_buffer semantics (unique_rx_buffer.h:33-139)
    def lock(self):
        self.locked = True

    def unlock(self):
        self.locked = False

    def release(self):
        self.locked = False
        self._pool._free(self)

    def is_valid(self):
        return True


class rx_buffer_pool:
    """rx_buffer_pool with external soft bits: hands out entries of the device HARQ arena (rx_buffer_pool.h:44-92).

    reserve() fails (returns None) when the pool is exhausted, when the buffer is locked, or when a retransmission asks
    for a different number of codeblocks - the reservation failures listed in rx_buffer_pool.h:62-75.
    """

    def __init__(self, ctx: capi.Context, first_entry=0, nof_entries=None):
        nof_entries = ctx.cfg.harq_entries - first_entry if nof_entries is None else nof_entries
        self._free_ids = list(range(first_entry + nof_entries - 1, first_entry - 1, -1))
        self._buffers = {}
        self._ctx = ctx

    def reserve(self, slot, key, nof_codeblocks, new_data) -> Optional[rx_buffer]:
        buf = self._buffers.get(key)
        if buf is not None:
            if buf.locked:
                return None
            if buf.get_nof_codeblocks() != nof_codeblocks:
                if not new_data:
                    return None
                self._free(buf)
                buf = None
        if buf is None:
            if not new_data or len(self._free_ids) < nof_codeblocks:
                return None
            buf = rx_buffer(self, key, [self._free_ids.pop() for _ in range(nof_codeblocks)])
            self._buffers[key] = buf
        buf.lock()
        return buf

    def _free(self, buf):
        if self._buffers.get(buf.key) is buf:
            del self._buffers[buf.key]
            for i in buf._ids:
                self._ctx.harq_free(i)
            self._free_ids.extend(reversed(buf._ids))


class _QueuedTb:
    __slots__ = ("decoder", "transport_block", "rm_buffer", "notifier", "cfg", "llrs", "metas", "decode_mask")


class pusch_decoder:
    """One pusch_decoder instance (one per PUSCH processor in the reference); bound to a PuschDecoderBatch."""

    def __init__(self, batch):
        self._batch = batch
        self._state = "idle"

    def new_data(self, transport_block: np.ndarray, rm_buffer: rx_buffer, notifier, cfg: pusch_decoder_configuration):
        if self._state != "idle":
            raise RuntimeError(f"Invalid state. It was expected to be idle but it was {self._state}.")
        tbs = transport_block.size * 8
        nof_cb = compute_nof_codeblocks(tbs, cfg.base_graph)
        if nof_cb != rm_buffer.get_nof_codeblocks():
            raise ValueError(f"Wrong number of codeblocks {rm_buffer.get_nof_codeblocks()} (expected {nof_cb}).")
        self._tb, self._buf, self._notifier, self._cfg = transport_block, rm_buffer, notifier, cfg
        self._chunks, self._count, self._expected = [], 0, None
        if cfg.new_data:
            rm_buffer.reset_codeblocks_crc()
        self._state = "collecting"
        return self

    def set_nof_softbits(self, nof_softbits: int):
        # The batched decoder starts at the end of the slot batch; like pusch_decoder_hw_impl::set_nof_softbits
        # (pusch_decoder_hw_impl.h:98-101) this only records the expectation.
        if nof_softbits % self._cfg.mod != 0:
            raise ValueError("The number of soft bits must be multiple of the modulation order.")
        self._expected = nof_softbits

    # pusch_decoder_buffer
    def on_new_softbits(self, softbits: np.ndarray):
        if self._state != "collecting":
            raise RuntimeError(f"Invalid state. It was expected to be collecting but it was {self._state}.")
        self._chunks.append(np.ascontiguousarray(softbits, np.int8))
        self._count += softbits.size

    def on_end_softbits(self):
        if self._state != "collecting":
            raise RuntimeError(f"Invalid state. It was expected to be collecting but it was {self._state}.")
        if self._expected is not None and self._expected != self._count:
            raise ValueError(f"The number of UL-SCH softbits, i.e., {self._count}, does not match the expected value.")
        if self._count % self._cfg.mod != 0:
            raise ValueError("The number of soft bits must be multiple of the modulation order.")
        q = _QueuedTb()
        q.decoder, q.transport_block, q.rm_buffer, q.notifier, q.cfg = self, self._tb, self._buf, self._notifier, self._cfg
        q.llrs = self._chunks[0] if len(self._chunks) == 1 else np.concatenate(self._chunks)
        self._state = "decoding"
        self._batch._queue(q)


class PuschDecoderBatch:
    """Slot batch: collects the transport blocks of every pusch_decoder created from it and decodes them together."""

    def __init__(self, ctx: capi.Context, stream=0):
        self._ctx, self._stream = ctx, stream
        self._queued: List[_QueuedTb] = []

    def create(self) -> pusch_decoder:
        """pusch_decoder_factory::create()."""
        return pusch_decoder(self)

    def _queue(self, q):
        self._queued.append(q)

    def pending(self):
        return len(self._queued)

    def flush(self):
        """Decodes everything queued with one GPU submission and notifies each TB's notifier."""
        if not self._queued:
            return
        tbs_q, self._queued = self._queued, []
        n_cb = sum(q.rm_buffer.get_nof_codeblocks() for q in tbs_q)
        n_llr = sum(q.llrs.size for q in tbs_q)
        cbs = np.zeros(n_cb, capi.CB_DESC_DTYPE)
        tbs = np.zeros(len(tbs_q), capi.TB_DESC_DTYPE)
        llrs = np.empty(n_llr, np.int8)
        i_cb = llr_off = tb_off = 0
        for i_tb, q in enumerate(tbs_q):
            cfg = q.cfg
            tbs_bits = q.transport_block.size * 8
            q.metas = segment_rx(tbs_bits, cfg.base_graph, cfg.rv, cfg.mod, cfg.Nref, cfg.nof_layers, q.llrs.size)
            C = len(q.metas)
            crc_kind = capi.CRC24B if C > 1 else (capi.CRC24A if tbs_bits > MAX_BITS_CRC16 else capi.CRC16)
            crcs = q.rm_buffer.get_codeblocks_crc()
            q.decode_mask = ~crcs.copy()
            llrs[llr_off:llr_off + q.llrs.size] = q.llrs
            tbs[i_tb] = (i_cb, C, tbs_bits, tb_off, 0)
            tb_off += (tbs_bits + 24 + 31) // 32 * 4
            for k, m in enumerate(q.metas):
                flags = capi.CB_DEMATCH
                flags |= capi.CB_NEW_DATA if cfg.new_data else 0
                flags |= capi.CB_EARLY_STOP if cfg.use_early_stop else 0
                flags |= capi.CB_DECODE if q.decode_mask[k] else 0
                cbs[i_cb] = (llr_off + m.cw_offset, m.rm_length, q.rm_buffer.get_absolute_codeblock_id(k), cfg.Nref,
                             m.lifting_size, m.nof_filler_bits, cfg.base_graph, cfg.mod, cfg.rv, crc_kind,
                             cfg.nof_ldpc_iterations, flags, i_tb)
                i_cb += 1
            llr_off += q.llrs.size
        self._ctx.submit(cbs, llrs, tbs, stream=self._stream, want_bits=False, want_tb=True)
        out = self._ctx.wait(self._stream)
        self.last_output = out
        for i_tb, q in enumerate(tbs_q):
            first, C = int(tbs[i_tb]["first_cb"]), int(tbs[i_tb]["nof_cb"])
            res = out["cb_results"][first:first + C]
            crcs = q.rm_buffer.get_codeblocks_crc()
            result = pusch_decoder_result(False, C)
            for k in range(C):
                if not q.decode_mask[k]:
                    continue
                if res[k]["crc_ok"]:
                    crcs[k] = True
                    result.ldpc_decoder_stats.update(res[k]["iters"])
                else:
                    result.ldpc_decoder_stats.update(q.cfg.nof_ldpc_iterations)
            tb_res = out["tb_results"][i_tb]
            nbytes = q.transport_block.size
            o = int(tbs[i_tb]["out_offset"])
            if C == 1:
                result.tb_crc_ok = bool(crcs[0])
                if result.tb_crc_ok:
                    q.transport_block[:] = out["tb_bytes"][o:o + nbytes]
            elif crcs.all():
                q.transport_block[:] = out["tb_bytes"][o:o + nbytes]
                if tb_res["tb_crc_ok"]:
                    result.tb_crc_ok = True
                else:
                    q.rm_buffer.reset_codeblocks_crc()
            if result.tb_crc_ok:
                q.rm_buffer.release()
            else:
                q.rm_buffer.unlock()
            q.decoder._state = "idle"
            q.notifier.on_sch_data(result)


class pusch_decoder_notifier_spy:
    """Test double with the interface of the reference's pusch_decoder_notifier_spy
    (tests/unittests/phy/upper/channel_processors/pusch/pusch_decoder_notifier_spy.h:31-43)."""

    def __init__(self):
        self._entries: List[pusch_decoder_result] = []

    def on_sch_data(self, result):
        self._entries.append(result)

    def get_entries(self):
        return self._entries
