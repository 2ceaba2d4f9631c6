import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def orc():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_available():
    from oracle.pyoracle import Reference
    return Reference.available()


@pytest.fixture(scope="session")
def ctx():
    """One GPU context for the whole session. Fails loudly when the CUDA library or the GPU is missing."""
    from srsran_edgeric_5g_b200 import capi
    c = capi.Context(device=0, max_cbs=2048, max_llrs=2048 * 12288, harq_entries=2048, max_tbs=64,
                     max_tb_bytes=2 << 20)
    yield c
    c.close()
