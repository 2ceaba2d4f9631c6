"""In-tree build of the CUDA library (nvcc cross-compiles for sm_100a without a GPU)."""
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libpusch_dec_cuda.so"
SOURCES = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) +
                 [PKG.parent / "include" / "pusch_dec_cuda.h"])

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(s.stat().st_mtime > t for s in SOURCES)


def build(force=False, verbose=False):
    """Compile csrc/pusch_dec_cuda.cu -> libpusch_dec_cuda.so next to this file. Safe to call from several processes at
    once (one rank per GPU under torchrun): one compiles - into a temporary file that is renamed when complete - the
    others wait for the lock and find the library up to date."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return LIB
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
        tmp = PKG / ("libpusch_dec_cuda.tmp%d.so" % os.getpid())
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(tmp), str(CSRC / "pusch_dec_cuda.cu")]
        # The image exports CC/CXX pointing at a relocated gcc; nvcc must use the system host compiler.
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        out = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if out.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError("nvcc failed:\n" + out.stdout + out.stderr)
        os.replace(tmp, LIB)
        if verbose:
            print(out.stderr)
        return LIB
