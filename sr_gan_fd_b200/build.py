"""In-tree build of libb200sr.so (nvcc cross-compiles sm_100a without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb200sr.so")
SOURCES = ["b200sr.cu"]
HEADERS = ["ptx.cuh", "conv_kernel.cuh", "wgrad_kernel.cuh", "aux_kernels.cuh", "optim_kernels.cuh", "iqa_kernels.cuh", "vgg_kernels.cuh",
           "disc_kernels.cuh", os.path.join("..", "..", "include", "b200sr.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "550",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


STAMP_PATH = LIB_PATH + ".srchash"


def source_hash() -> str:
    """SHA-256 over the sources, headers and compiler flags the library is built from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        h.update(f.encode())
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_stale() -> bool:
    """The library is current iff it exists and was built from exactly these sources (hash stamp next to it; file times are
    meaningless after a checkout or a copy to the GPU box)."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as fh:
        return fh.read().strip() != source_hash()


LAST_BUILD = {"ran_nvcc": False, "reason": "not called"}


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> sr_gan_fd_b200/libb200sr.so.  Returns the library path; LAST_BUILD says whether nvcc ran."""
    if not force and not is_stale():
        LAST_BUILD.update(ran_nvcc=False, reason="library matches the source hash " + source_hash()[:12])
        return LIB_PATH
    try:
        nvcc = _nvcc()
    except RuntimeError:
        if os.path.exists(LIB_PATH) and not force:  # GPU box without a toolchain: the shipped library is what there is
            LAST_BUILD.update(ran_nvcc=False, reason="nvcc not found; using the shipped library (source hash NOT verified)")
            return LIB_PATH
        raise
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    with open(STAMP_PATH, "w") as fh:
        fh.write(source_hash())
    LAST_BUILD.update(ran_nvcc=True, reason="forced" if force else "sources changed (hash stamp missing or different)")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
