"""GPU test (-m gpu): the C++ adapters behind the reference's own classes.

adapters/_build/test_adapters is built in the build container (it needs the reference headers) and shipped with the
snapshot. It runs the UNMODIFIED pusch_decoder_hw_impl of the reference on top of hw_accelerator_pusch_dec_cuda and
compares it with the reference's software pusch_decoder_impl over HARQ retransmissions, plus the single-codeblock
ldpc_decoder / ldpc_rate_dematcher / crc_calculator adapters against their "auto" software counterparts."""
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
BIN = Path(__file__).resolve().parent.parent / "adapters" / "_build" / "test_adapters"


def test_reference_hw_decoder_front_end_on_cuda_accelerator():
    if not BIN.exists():
        pytest.skip("adapters/_build/test_adapters not built (needs /root/reference at build time)")
    out = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "PASS" in out.stdout
