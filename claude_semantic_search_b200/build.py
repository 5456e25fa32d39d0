"""Build libcss_b200.so (sm_100a) and the oracle's C restatement.

    python -m claude_semantic_search_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the CPU-only container and
the resulting .so travels to the GPU box with the repo snapshot.
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
LIB = PKG / "libcss_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--use_fast_math" if False else "-DCSS_NO_FAST_MATH",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _deps():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc"))
                  + [ROOT / "include" / "css_b200.h", Path(__file__)])


def _digest() -> str:
    h = hashlib.sha256()
    for p in _deps():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_native(force: bool = False, verbose: bool = False) -> Path:
    stamp = PKG / "csrc" / ".build_stamp"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == dig:
        return LIB
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        cmd = [NVCC, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src.name}:\n{out}\n")
    (objdir / "nvcc.log").write_text("\n".join(log))
    if failed:
        raise RuntimeError("nvcc compilation failed (see output above)")
    if verbose:
        print("\n".join(log))
    # cudart linked statically: the .so loads (and exports its symbols) on a box
    # without a driver; the driver API (TMA descriptor encode) is resolved at run
    # time through cudaGetDriverEntryPoint.
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    subprocess.run(cmd, check=True)
    stamp.write_text(dig)
    return LIB


def build_oracle(force: bool = False) -> Path:
    odir = ROOT / "oracle"
    out = odir / "_build" / "liboracle_flat.so"
    src = odir / "flat_ip.c"
    if not src.exists():
        return out
    if not force and out.exists() and out.stat().st_mtime >= src.stat().st_mtime:
        return out
    out.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", str(out), str(src), "-lm"],
                   check=True)
    return out


if __name__ == "__main__":
    force = "--force" in sys.argv
    lib = build_native(force=force, verbose="--verbose" in sys.argv)
    print("built", lib)
    print("built", build_oracle(force=force))
