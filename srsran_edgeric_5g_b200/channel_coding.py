"""Host-side mirror of the reference's channel-coding plug-in surface for the "cuda" variant.

Same names, argument meaning and error behaviour as
  create_ldpc_decoder_factory_sw / create_ldpc_rate_dematcher_factory_sw / create_crc_calculator_factory_sw
  (include/srsran/phy/upper/channel_coding/channel_coding_factories.h:50-77,
   lib/phy/upper/channel_coding/channel_coding_factories.cpp:61-125, :158-193):
a factory for an unknown type is None (the reference returns nullptr), contract violations raise (the reference
asserts). Everything computes on the GPU through the C ABI; there is no software path here.
"""
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import capi
from .ldpc import BG1, BG2, CodeblockMetadata

_shared_ctx = None


def shared_context():
    """Process-wide context for the single-codeblock (latency path) objects."""
    global _shared_ctx
    if _shared_ctx is None:
        _shared_ctx = capi.Context(max_cbs=256, harq_entries=256, max_tbs=16, max_tb_bytes=1 << 20)
    return _shared_ctx


class crc_generator_poly:
    """crc_calculator.h:31-39."""
    CRC24A, CRC24B, CRC24C, CRC16, CRC11, CRC6 = range(6)


_POLY_TO_KIND = {crc_generator_poly.CRC16: capi.CRC16, crc_generator_poly.CRC24A: capi.CRC24A,
                 crc_generator_poly.CRC24B: capi.CRC24B}


class crc_calculator:
    """crc_calculator (crc_calculator.h:62-84) - CRC16 / CRC24A / CRC24B on the GPU."""

    def __init__(self, ctx, poly):
        if poly not in _POLY_TO_KIND:
            raise ValueError("the cuda crc_calculator supports CRC16, CRC24A and CRC24B")
        self._ctx, self._poly = ctx, poly

    def get_generator_poly(self):
        return self._poly

    def calculate(self, packed_bits, nof_bits=None):
        """calculate(const bit_buffer&): packed_bits is uint8 (MSB first), nof_bits defaults to all of it."""
        packed_bits = np.ascontiguousarray(packed_bits, np.uint8)
        nof_bits = packed_bits.size * 8 if nof_bits is None else nof_bits
        return self._ctx.crc(_POLY_TO_KIND[self._poly], packed_bits, nof_bits)

    def calculate_byte(self, data):
        return self.calculate(data)

    def calculate_bit(self, bits):
        bits = np.ascontiguousarray(bits, np.uint8)
        return self.calculate(np.packbits(bits), bits.size)


class crc_calculator_factory:
    def __init__(self, ctx):
        self._ctx = ctx

    def create(self, poly):
        return crc_calculator(self._ctx, poly)


def create_crc_calculator_factory_sw(crc_type: str):
    if crc_type != "cuda":
        return None
    return crc_calculator_factory(shared_context())


@dataclass
class ldpc_decoder_configuration:
    """ldpc_decoder::configuration (ldpc_decoder.h:44-57)."""

    @dataclass
    class algorithm_details:
        max_iterations: int = 6
        scaling_factor: float = 0.8

    block_conf: CodeblockMetadata = None
    algorithm_conf: "ldpc_decoder_configuration.algorithm_details" = field(default_factory=algorithm_details)


class ldpc_decoder:
    """ldpc_decoder (ldpc_decoder.h:37-75): decode(output, input, crc, cfg) -> optional iteration count."""

    def __init__(self, ctx):
        self._ctx = ctx

    def decode(self, output: np.ndarray, input_llrs: np.ndarray, crc: Optional[crc_calculator],
               cfg: ldpc_decoder_configuration) -> Optional[int]:
        m = cfg.block_conf
        Z, bg = m.lifting_size, m.base_graph
        K = (22 if bg == BG1 else 10) * Z
        N = (66 if bg == BG1 else 50) * Z
        # Same contract as ldpc_decoder_impl::decode (ldpc_decoder_impl.cpp:69-83); the reference asserts.
        if output.size != (K + 7) // 8:
            raise ValueError(f"The output size {output.size * 8} is not equal to the message length {K}.")
        if input_llrs.size > N:
            raise ValueError(f"The input size {input_llrs.size} exceeds the maximum message length {N}.")
        if input_llrs.size < K + 2 * Z:
            raise ValueError(f"The input length {input_llrs.size} does not reach minimum {K + 2 * Z}")
        if abs(cfg.algorithm_conf.scaling_factor - 0.8) > 1e-6:
            raise ValueError("the cuda decoder implements the reference's default scaling factor 0.8")
        if m.nof_crc_bits not in (16, 24):
            raise ValueError("Invalid number of CRC bits.")
        kind = capi.CRC_NONE if crc is None else _POLY_TO_KIND[crc.get_generator_poly()]
        iters, _ = self._ctx.ldpc_decode(bg, Z, input_llrs, m.nof_filler_bits, kind, cfg.algorithm_conf.max_iterations,
                                         out=output)
        return iters if iters > 0 else None


class ldpc_decoder_factory:
    def __init__(self, ctx):
        self._ctx = ctx

    def create(self):
        return ldpc_decoder(self._ctx)


def create_ldpc_decoder_factory_sw(dec_type: str):
    if dec_type != "cuda":
        return None
    return ldpc_decoder_factory(shared_context())


class ldpc_rate_dematcher:
    """ldpc_rate_dematcher (ldpc_rate_dematcher.h:35-56): rate_dematch(output, input, new_data, cfg)."""

    def __init__(self, ctx):
        self._ctx = ctx

    def rate_dematch(self, output: np.ndarray, input_llrs: np.ndarray, new_data: bool, cfg: CodeblockMetadata):
        N = output.size
        if N % 66 != 0 and N % 50 != 0:
            raise ValueError("LDPC rate dematching: invalid input length.")
        if input_llrs.size % cfg.mod != 0:
            raise ValueError("The input length should be a multiple of the modulation order.")
        if not 0 <= cfg.rv <= 3:
            raise ValueError("RV should an integer between 0 and 3.")
        self._ctx.rate_dematch(output, input_llrs, new_data, cfg.rv, cfg.mod, cfg.Nref, cfg.nof_filler_bits)


class ldpc_rate_dematcher_factory:
    def __init__(self, ctx):
        self._ctx = ctx

    def create(self):
        return ldpc_rate_dematcher(self._ctx)


def create_ldpc_rate_dematcher_factory_sw(dematcher_type: str):
    if dematcher_type != "cuda":
        return None
    return ldpc_rate_dematcher_factory(shared_context())
