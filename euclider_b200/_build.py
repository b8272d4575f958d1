"""In-tree build of libeuclider_b200.so (host front end + CUDA kernels for sm_100a).

nvcc cross-compiles without a GPU, so this runs in the CPU-only container; the built .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
BUILD = ROOT / ("build" + os.environ.get("EUCL_LIB_SUFFIX", ""))
SUFFIX = os.environ.get("EUCL_LIB_SUFFIX", "")  # experiment builds: EUCL_LIB_SUFFIX=_x EUCL_NVCC_EXTRA="-D..."
LIB = PKG / f"libeuclider_b200{SUFFIX}.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the reference (Rust/LLVM) never contracts a*b+c; bit-matching its hit decisions
# needs separate DMUL/DADD.  No fast-math anywhere.
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
              "-Xcompiler", "-fPIC,-Wall,-ffp-contract=off", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-Wextra", "-ffp-contract=off"]
NVCC_FLAGS += os.environ.get("EUCL_NVCC_EXTRA", "").split()


def _sources():
    host = sorted((CSRC / "host").glob("*.cc"))
    cuda = sorted((CSRC / "cuda").glob("*.cu"))
    return host, cuda


def _stamp(paths) -> str:
    h = hashlib.sha256()
    deps = list(paths) + sorted(CSRC.rglob("*.h")) + sorted(CSRC.rglob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    for p in deps:
        h.update(str(p).encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS + GENCODE).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    host, cuda = _sources()
    BUILD.mkdir(exist_ok=True)
    stamp_file = BUILD / "lib.stamp"
    stamp = _stamp(host + cuda)
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    inc = ["-I", str(ROOT / "include"), "-I", str(CSRC / "host"), "-I", str(CSRC / "cuda")]
    objs = []
    log = []
    for src in host:
        obj = BUILD / (src.stem + ".o")
        cmd = [CXX, *CXX_FLAGS, *inc, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"host compile failed: {src}\n{r.stderr}")
        objs.append(obj)
    # kernels.cu is compiled twice: real = double (namespace eucl) and real = float (namespace eucl_f32, the
    # reference's `low_precision` feature); api_device.cu dispatches on the scene's precision
    jobs = [(src, BUILD / (src.stem + ".cu.o"), []) for src in cuda]
    jobs += [(src, BUILD / (src.stem + "_f32.cu.o"), ["-DEUCL_REAL=float", "-DEUCL_NS=eucl_f32"]) for src in cuda if src.name == "kernels.cu"]
    procs = [(src, obj, subprocess.Popen([NVCC, *GENCODE, *NVCC_FLAGS, *extra, *inc, "-c", str(src), "-o", str(obj)],
                                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)) for src, obj, extra in jobs]
    for src, obj, proc in procs:
        _, err = proc.communicate()
        log.append(err)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed: {src} -> {obj.name}\n{err}")
        objs.append(obj)
    cmd = [NVCC, *GENCODE, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed\n{r.stderr}")
    (BUILD / "ptxas.log").write_text("\n".join(log))
    stamp_file.write_text(stamp)
    if verbose:
        print("\n".join(log), file=sys.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
