"""In-tree build of the C-ABI library (nvcc -> realtime-codec-agent_b200/libmagicodec_b200.so).

sm_100a only: `-gencode arch=compute_100a,code=sm_100a`.  The built .so is git-ignored but travels
to the GPU box with the repo snapshot.  `python -m rca_b200_loader build` or
`__graft_entry__.build()`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libmagicodec_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
_STAMP = os.path.join(PKG_DIR, ".build_stamp")


def _sources_digest() -> str:
    hsh = hashlib.sha256()
    roots = [CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")]
    for root in roots:
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(root, name), "rb") as f:
                    hsh.update(name.encode())
                    hsh.update(f.read())
    return hsh.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; the B200 engine cannot be built")


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True adds -DMC_TRACE (kernels log their dependency-wait times; tools/trace_stream.py) — a diagnostic build,
    written next to the shipped library as libmagicodec_b200_trace.so and loaded only when $MAGICODEC_B200_LIB names it."""
    digest = _sources_digest() + ("+trace" if trace else "")
    lib_path = LIB_PATH.replace(".so", "_trace.so") if trace else LIB_PATH
    stamp = _STAMP + ("_trace" if trace else "")
    if not force and os.path.isfile(lib_path) and os.path.isfile(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return lib_path
    cmd = [
        nvcc_path(), "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", *(["-DMC_TRACE"] if trace else []),
        "-Xptxas", "-v" if verbose else "-O3",
        "-o", lib_path, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "audio_decode.cpp"),
        "-lcudart",
    ]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libmagicodec_b200.so:\n" + proc.stderr[-4000:])
    with open(stamp, "w") as f:
        f.write(digest)
    return lib_path


def main(argv=None) -> None:
    argv = sys.argv[1:] if argv is None else argv
    print(build(force="--force" in argv, verbose="-v" in argv, trace="--trace" in argv))


if __name__ == "__main__":
    main()
