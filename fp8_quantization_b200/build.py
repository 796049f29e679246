"""Builds libfp8fq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so travels
to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "fp8fq_kernels.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "fp8fq_core.h"), os.path.join(HERE, "..", "include", "fp8fq.h")]
OUT = os.path.join(HERE, "libfp8fq.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # no --use_fast_math: log2f/powf/division must be the IEEE / libdevice versions ATen uses
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, defines=(), out=None):
    """``defines`` / ``out``: an alternative build of the same library (e.g. ``-DFP8FQ_FOLD_ACT=1``) next to the default
    one; ``FP8FQ_LIB=<out>`` makes the package load it."""
    target = out or OUT
    if not force and not defines and out is None and not needs_build():
        return OUT
    cmd = ([find_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])
           + ["-o", target, SRC])
    env = dict(os.environ)
    env.pop("CC", None)  # the image's CC points at a gcc nvcc does not need
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libfp8fq.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return target


TORCH_SRC = os.path.join(HERE, "csrc", "fp8fq_torch.cpp")
TORCH_OUT = os.path.join(HERE, "libfp8fq_torch.so")


def build_torch_ext(force=False):
    """Builds libfp8fq_torch.so: the TORCH_LIBRARY(fp8fq, ...) operator layer over the C ABI (csrc/fp8fq_torch.cpp), a plain
    g++ compile against torch's headers, linked to libfp8fq.so next to it ($ORIGIN).  In-tree, like libfp8fq.so."""
    deps = [TORCH_SRC, DEPS[2], OUT]
    if not force and os.path.exists(TORCH_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(TORCH_OUT) for d in deps):
        return TORCH_OUT
    import torch

    tdir = os.path.dirname(torch.__file__)
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(find_nvcc())), "include")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include", f"-I{cuda_inc}", TORCH_SRC, "-o", TORCH_OUT,
           f"-L{tdir}/lib", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", f"-L{HERE}", "-l:libfp8fq.so",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tdir}/lib"]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building libfp8fq_torch.so")
    return TORCH_OUT


if __name__ == "__main__":
    # python -m fp8_quantization_b200.build [--force] [-v] [-DNAME=VALUE ...] [--out PATH]
    argv = sys.argv[1:]
    print(build(force="--force" in argv, verbose="-v" in argv, defines=[a[2:] for a in argv if a.startswith("-D")],
                out=argv[argv.index("--out") + 1] if "--out" in argv else None))
    if "--out" not in argv and not any(a.startswith("-D") for a in argv):
        print(build_torch_ext(force="--force" in argv))
