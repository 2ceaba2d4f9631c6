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


# ---- the reference's OWN harnesses on the "cuda" variants (integration/Makefile) ------------------------------------------
HARNESS = Path(__file__).resolve().parent.parent / "integration" / "_build"


def test_reference_crc_calculator_test_with_cuda_factory():
    """tests/unittests/phy/upper/channel_coding/crc_calculator_test.cpp of the reference, unmodified, `-F cuda`: all six
    generator polynomials, byte / bit / bit-buffer interfaces, against its own bit-serial model."""
    exe = HARNESS / "crc_calculator_test"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "-F", "cuda"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_reference_ldpc_decoder_benchmark_with_cuda_factory():
    """tests/benchmarks/phy/upper/channel_coding/ldpc/ldpc_decoder_benchmark.cpp of the reference, unmodified,
    `-T cuda`: create_ldpc_decoder_factory_sw("cuda") hands out the GPU decoder and the benchmark runs to completion
    (encoded codeblocks with CRC16, both base graphs)."""
    exe = HARNESS / "ldpc_decoder_benchmark"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "-T", "cuda", "-L", "96", "-I", "6", "-C", "-R", "20"], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "LDPC decoder cuda" in out.stdout and "BG=2 LS=96" in out.stdout


def test_reference_ldpc_encoder_benchmark_with_cuda_factory():
    """tests/benchmarks/phy/upper/channel_coding/ldpc/ldpc_encoder_benchmark.cpp of the reference, unmodified,
    `-T cuda`: create_ldpc_encoder_factory_sw("cuda") hands out the GPU encoder (downlink twin) and the benchmark runs to
    completion over both base graphs and all lifting sizes."""
    exe = HARNESS / "ldpc_encoder_benchmark"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "-T", "cuda", "-R", "5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "LDPC encoder cuda" in out.stdout


def test_reference_pusch_processor_benchmark_with_cuda_decoder_and_dematcher():
    """tests/benchmarks/phy/upper/channel_processors/pusch/pusch_processor_benchmark.cpp of the reference, unmodified:
    the whole PUSCH chain (channel estimator, equaliser, demodulator, UL-SCH demultiplexer, pusch_decoder_impl) with
    `-D cuda -M cuda`, i.e. create_ldpc_decoder_factory_sw("cuda") / create_ldpc_rate_dematcher_factory_sw("cuda") inside
    its own pusch_decoder_impl, on several worker threads at once. The benchmark asserts that every transport block
    passes its CRC (pusch_processor_benchmark.cpp: TESTASSERT on the notifier's result)."""
    exe = HARNESS / "pusch_processor_benchmark"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    for profile, threads in (("scs15_5MHz_qpsk_rv0_1port_1layer", 4), ("scs30_100MHz_256qam_rv0_4port_nlayer", 2)):
        out = subprocess.run([str(exe), "-m", "throughput_total", "-R", "2", "-B", "2", "-T", str(threads), "-t", "0", "-D",
                              "cuda", "-M", "cuda", "-P", profile], capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "PUSCH RB=" in out.stdout, out.stdout[-2000:]


def test_reference_pusch_decoder_hwacc_benchmark_with_cuda_accelerator():
    """tests/benchmarks/phy/upper/channel_processors/pusch/pusch_decoder_hwacc_benchmark.cpp of the reference with the
    accelerator name "cuda" added next to "acc100" (integration/apply_cuda_branch.py): generic pusch_decoder_impl against
    the reference's pusch_decoder_hw_impl on hal::create_hw_accelerator_pusch_dec_factory({acc_type = "cuda"}) - the hal
    registry with the "cuda" branch - for PRB {25, 52, 106, 270} x {QPSK, 16QAM, 64QAM, 256QAM}."""
    exe = HARNESS / "pusch_decoder_hwacc_benchmark"
    if not exe.exists():
        pytest.skip("integration/_build not built (needs /root/reference at build time)")
    out = subprocess.run([str(exe), "-T", "cuda", "-i", "6"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("PUSCH RB=") == 16 and "cuda" in out.stdout, out.stdout[-2000:]
    print(out.stdout)


def test_cpp_client_of_the_c_abi(orc, tmp_path):
    """tools/latency_probe.cpp (pdc_create / pdc_host_alloc / pdc_submit / pdc_wait from C++, nothing but the header): a
    two-codeblock transport block encoded by the oracle, ideal soft bits, decoded and assembled through the C ABI."""
    import json
    import subprocess
    from pathlib import Path

    import numpy as np

    from srsran_edgeric_5g_b200 import capi, ldpc
    exe = Path(__file__).resolve().parent.parent / "tools" / "_build" / "latency_probe"
    if not exe.exists():
        pytest.skip("tools/_build/latency_probe not built (__graft_entry__.build())")
    rng = np.random.default_rng(5)
    bg, qm, nl, tb_bytes = 1, 6, 2, 1800
    tbs_bits = tb_bytes * 8
    n_llr = int(np.ceil(tbs_bits / 0.6 / qm / nl)) * nl * qm
    C = ldpc.compute_nof_codeblocks(tbs_bits, bg)
    assert C == 2
    nref = ldpc.compute_N_ref(tb_bytes, C)
    tb = rng.integers(0, 256, tb_bytes).astype(np.uint8)
    cw, _ = orc.tb_encode(tb, bg, 0, qm, nref, nl, n_llr)
    llrs = np.where(cw == 0, 40, -40).astype(np.int8)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
    cbs = np.zeros(C, capi.CB_DESC_DTYPE)
    for k, m in enumerate(ldpc.segment_rx(tbs_bits, bg, 0, qm, nref, nl, n_llr)):
        cbs[k] = (m.cw_offset, m.rm_length, k, nref, m.lifting_size, m.nof_filler_bits, bg, qm, 0, capi.CRC24B, 6, flags, 0)
    tbd = np.zeros(1, capi.TB_DESC_DTYPE)
    tbd[0] = (0, C, tbs_bits, 0, 0)
    paths = []
    for name, arr in (("cbs.bin", cbs), ("tbs.bin", tbd), ("llrs.bin", llrs)):
        paths.append(str(tmp_path / name))
        arr.tofile(paths[-1])
    run = subprocess.run([str(exe)] + paths + [str(tb_bytes + 8), "20"], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stderr
    res = json.loads(run.stdout.strip().splitlines()[-1])
    assert res["tb_crc_ok"] is True and res["codeblocks"] == 2 and res["transport_blocks"] == 1 and res["p50"] > 0
