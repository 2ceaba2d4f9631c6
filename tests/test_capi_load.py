"""CPU tests: the C-ABI library builds, loads and exports every symbol include/pusch_dec_cuda.h declares.
No compute call is made (there may be no GPU); creating a context without a GPU must fail loudly, not fall back."""
import ctypes
import re
from pathlib import Path

import pytest

from srsran_edgeric_5g_b200 import build, capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "pusch_dec_cuda.h"


def declared_functions():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdc_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_the_abi():
    lib = build.build()
    assert lib.exists()
    L = ctypes.CDLL(str(lib))
    decl = declared_functions()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(capi.EXPORTS) == decl, "capi.EXPORTS and the header disagree"


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(capi.CbDesc) == 28 and capi.CB_DESC_DTYPE.itemsize == 28
    assert ctypes.sizeof(capi.CbResult) == 4 and capi.CB_RESULT_DTYPE.itemsize == 4
    assert ctypes.sizeof(capi.TbDesc) == 20 and capi.TB_DESC_DTYPE.itemsize == 20
    assert ctypes.sizeof(capi.TbResult) == 4 and capi.TB_RESULT_DTYPE.itemsize == 4
    assert ctypes.sizeof(capi.Config) == 40
    assert capi.DEMOD_CALL_DTYPE.itemsize == 16
    for f, _ in capi.CbDesc._fields_:
        assert capi.CB_DESC_DTYPE.fields[f][1] == getattr(capi.CbDesc, f).offset


def test_constants_match_the_header():
    """Every PDC_* value the header defines and the Python mirror also names (with or without the prefix) is the same."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    defines = {}
    for name, value in re.findall(r"^#define\s+(PDC_[A-Z0-9_]+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+))u?\)?\s*$", text, flags=re.M):
        defines[name] = int(value, 0)
    assert len(defines) >= 30
    checked = 0
    for name, value in defines.items():
        for mirror in (name, name[len("PDC_"):]):
            if hasattr(capi, mirror):
                assert getattr(capi, mirror) == value, f"{name}: header {value}, capi.{mirror} {getattr(capi, mirror)}"
                checked += 1
    assert checked >= 20


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.PdcError) as e:
        capi.Context(max_cbs=4, harq_entries=4)
    assert e.value.code == capi.PDC_ERR_NO_DEVICE


def test_sm100a_code_only():
    # The library carries sm_100a SASS and nothing else (no multi-arch fatbin, no PTX JIT fallback).
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", str(build.LIB)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_cubin_is_sm100a_and_keeps_the_occupancy_the_design_counts_on():
    """What DESIGN.md section 4 assumes about the compiled kernels, read from the library itself (cuobjdump; no GPU): only
    sm_100a code, the throughput decoder within 80 registers (two 384-thread CTAs per SM), the rate dematcher within 40
    (six 256-thread CTAs per SM), no local-memory arrays, and every kernel of the path present."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(tool).exists():
        pytest.skip("cuobjdump not available")
    lib = str(build.build())
    elfs = subprocess.run([tool, "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_\w+)\.cubin", elfs))
    assert archs == {"sm_100a"}, archs
    usage = subprocess.run([tool, "-res-usage", lib], capture_output=True, text=True).stdout
    regs = {}
    for name, reg, local in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) .*?LOCAL:(\d+)", usage):
        regs[name] = int(reg)
        assert int(local) == 0, (name, "local-memory array")
    decoders = {k: v for k, v in regs.items() if "ldpc_decode_h2_kernelILi384ELi2E" in k}
    assert len(decoders) >= 3 and max(decoders.values()) <= 80, decoders
    dematcher = [v for k, v in regs.items() if "rate_dematch_kernel" in k]
    assert dematcher and max(dematcher) <= 40, dematcher
    for kernel in ("demod_kernel", "tb_assemble_kernel", "ulsch_sch_kernel", "ldpc_encode_rm_kernel", "prg_kernel"):
        assert any(kernel in k for k in regs), kernel
