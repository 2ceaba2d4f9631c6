"""B200-native (sm_100a) uplink PUSCH decode path for the EdgeRIC-instrumented srsRAN 5G gNB.

Rate dematching + HARQ combining -> LDPC layered min-sum decoding (CRC early stop) -> CB/TB CRC, as hand-written CUDA
kernels behind a C ABI (include/pusch_dec_cuda.h). This package is the Python host side used by tests and benches; it
mirrors the reference's plug-in interfaces (see channel_coding.py and pusch_decoder.py). The C++ adapters that plug the
same library into the gNB are under adapters/.

Importing this package does not touch the GPU; creating a capi.Context does, and fails loudly without one.
"""
from . import build, capi, ldpc  # noqa: F401
from .channel_coding import (  # noqa: F401
    create_crc_calculator_factory_sw, create_ldpc_decoder_factory_sw, create_ldpc_rate_dematcher_factory_sw,
    crc_generator_poly, ldpc_decoder_configuration)
from .pusch_decoder import (  # noqa: F401
    PuschDecoderBatch, pusch_decoder_configuration, pusch_decoder_notifier_spy, pusch_decoder_result, rx_buffer_pool)

__all__ = ["build", "capi", "ldpc"]
