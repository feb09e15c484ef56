"""In-tree build of the native extensions (called by ``__graft_entry__.build()``).

CUDA sources are compiled for sm_100a only (``-gencode arch=compute_100a,code=sm_100a -lineinfo``); nvcc cross-compiles
on a machine without a GPU.  Outputs land in ``ps_pytorch_b200/ops/_build/*.so`` and are re-used until a source changes
(content hash), so the GPU box imports the shipped ``.so`` without compiling."""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")

CUDA_SOURCES = ["bindings.cpp", "arena.cpp", "optim_kernels.cu", "comm_kernels.cu", "bn_kernels.cu", "elementwise.cu",
                "gemm_tcgen05.cu", "conv2_tcgen05.cu", "wgrad_halo_tcgen05.cu", "wgrad2_tcgen05.cu", "stem_tcgen05.cu", "gemm_bindings.cpp", "stem_wgrad.cu"]
HOST_SOURCES = ["host_ext.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--use_fast_math", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _digest(files):
    h = hashlib.sha256()
    for f in sorted(files):
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def _build_one(name, sources, cuda, verbose):
    from torch.utils import cpp_extension
    srcs = [os.path.join(CSRC, s) for s in sources if os.path.isfile(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    dig = _digest(srcs + hdrs)
    so = os.path.join(BUILD, name + ".so")
    stamp = os.path.join(BUILD, name + ".sha256")
    if os.path.isfile(so) and os.path.isfile(stamp) and open(stamp).read().strip() == dig:
        return so, False
    work = os.path.join(BUILD, "obj_" + name)
    os.makedirs(work, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")   # our -gencode is what matters; this only silences the auto-detect
    os.environ.setdefault("MAX_JOBS", str(min(8, os.cpu_count() or 4)))
    cpp_extension.load(name=name, sources=srcs, extra_cflags=["-O3", "-std=c++17"],
                       extra_cuda_cflags=NVCC_FLAGS if cuda else None, extra_include_paths=[CSRC],
                       build_directory=work, verbose=verbose, is_python_module=False, with_cuda=cuda)
    shutil.copyfile(os.path.join(work, name + ".so"), so)
    with open(stamp, "w") as f:
        f.write(dig)
    return so, True


def build_all(verbose: bool = False):
    os.makedirs(BUILD, exist_ok=True)
    out = []
    out.append(_build_one("ps_b200_C", CUDA_SOURCES, True, verbose))
    out.append(_build_one("ps_b200_host", HOST_SOURCES, False, verbose))
    return out


if __name__ == "__main__":
    for so, rebuilt in build_all(verbose="-v" in sys.argv):
        print(("built   " if rebuilt else "cached  ") + so)
