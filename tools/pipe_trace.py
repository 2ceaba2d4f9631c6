#!/usr/bin/env python3
"""Where the time of a pipelined pdc_submit goes: the 16-cell slot of BASELINE config 4 through tools/latency_probe with
PDC_PIPE_TRACE=1 (device timestamps of every group's copy and kernels, printed by pdc_wait every 50th slot)."""
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    from oracle.pyoracle import Oracle
    from srsran_edgeric_5g_b200 import capi, ldpc
    from tests.vectors import make_tb_llrs
    cells = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    orc = Oracle()
    rng = np.random.default_rng(3)
    tbs_bits, n_llr, qm, nl = 1277992, 1362816, 8, 4
    C = ldpc.compute_nof_codeblocks(tbs_bits, 1)
    nref = ldpc.compute_N_ref(tbs_bits // 8, C)
    tb = rng.integers(0, 256, tbs_bits // 8).astype(np.uint8)
    llrs, _ = make_tb_llrs(orc, tb, 1, 0, qm, nref, nl, n_llr, 8.4, rng)
    metas = ldpc.segment_rx(tbs_bits, 1, 0, qm, nref, nl, n_llr)
    cbs = np.zeros(C * cells, capi.CB_DESC_DTYPE)
    tbd = np.zeros(cells, capi.TB_DESC_DTYPE)
    flags = capi.CB_DEMATCH | capi.CB_DECODE | capi.CB_NEW_DATA | capi.CB_EARLY_STOP
    tb_stride = (tbs_bits + 24 + 31) // 32 * 4
    for c in range(cells):
        tbd[c] = (c * C, C, tbs_bits, c * tb_stride, 0)
        for k, m in enumerate(metas):
            cbs[c * C + k] = (c * n_llr + m.cw_offset, m.rm_length, c * C + k, nref, m.lifting_size, m.nof_filler_bits, 1, qm,
                              0, capi.CRC24B, 6, flags, c)
    exe = ROOT / "tools" / "_build" / "latency_probe"
    with tempfile.TemporaryDirectory() as tmp:
        paths = [os.path.join(tmp, n) for n in ("cbs.bin", "tbs.bin", "llrs.bin")]
        for path, arr in zip(paths, (cbs, tbd, np.tile(llrs, cells))):
            np.ascontiguousarray(arr).tofile(path)
        env = dict(os.environ, PDC_PIPE_TRACE="1")
        run = subprocess.run([str(exe)] + paths + [str(cells * tb_stride), "300"], capture_output=True, text=True, env=env)
        print(run.stdout.strip())
        print("\n".join(run.stderr.strip().splitlines()[-4:]))


if __name__ == "__main__":
    main()
