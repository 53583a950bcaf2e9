"""Build libmsq_b200.so in-tree with nvcc for sm_100a (B200).

    python -m maxsquareloss_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the built library sits in
``maxsquareloss_b200/lib/`` (git-ignored, but it travels with the working tree).
``libmsq_torch.so`` next to it is the PyTorch host binding (``csrc/torch_binding.cpp``, g++): C++ autograd nodes that
call the C ABI of ``libmsq_b200.so``.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libmsq_b200.so")
SOURCES = ["api.cu", "confusion.cu", "eval_flip.cu", "prob_loss.cu", "softce.cu", "fused_loss.cu", "guidance.cu", "source_ce.cu", "host_pipe.cu", "comm.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "fused_common.cuh"), os.path.join(os.path.dirname(PKG), "include", "msq_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC",
              # share libcudart.so.12 with PyTorch: one runtime instance => one notion of the current
              # device, so torch.cuda.set_device(local_rank) also governs this library's launches
              "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmsq_b200.so cannot be built (there is no CPU fallback)")


TORCH_LIB = os.path.join(LIBDIR, "libmsq_torch.so")
TORCH_SRC = os.path.join(CSRC, "torch_binding.cpp")


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def torch_binding_is_stale():
    if not os.path.exists(TORCH_LIB):
        return True
    t = os.path.getmtime(TORCH_LIB)
    return any(os.path.getmtime(d) > t for d in (TORCH_SRC, HEADERS[-1]))


def build_torch_binding(force=False):
    """g++ the PyTorch host binding against this interpreter's torch headers and link it to libmsq_b200.so
    (``$ORIGIN`` run path: the two libraries travel together).  Returns its path."""
    if not force and not torch_binding_is_stale():
        return TORCH_LIB
    import torch
    tdir = os.path.dirname(os.path.abspath(torch.__file__))
    # the SYSTEM g++ (the one whose libstdc++.so.6 PyTorch runs on).  $CXX is deliberately ignored: this image exports a
    # CXX whose libstdc++ is linked statically, and a binding with a private copy of the iostreams crashes inside
    # TORCH_CHECK's message formatting (seen on the B200 box: SIGSEGV in std::ostream::_M_insert<long>)
    cxx = os.environ.get("MSQ_CXX") or next((c for c in ("/usr/bin/g++", shutil.which("g++")) if c and os.path.exists(c)), None)
    if not cxx:
        raise RuntimeError("g++ not found: libmsq_torch.so cannot be built")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-I" + os.path.join(tdir, "include"), "-I" + os.path.join(tdir, "include", "torch", "csrc", "api", "include"),
           "-I" + cuda_inc, TORCH_SRC, "-o", TORCH_LIB + ".tmp", "-L" + os.path.join(tdir, "lib"),
           "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-L" + LIBDIR, "-lmsq_b200", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr[-4000:])
    needed = subprocess.run(["readelf", "-d", TORCH_LIB + ".tmp"], capture_output=True, text=True).stdout
    if "libstdc++.so" not in needed:
        raise RuntimeError(f"{cxx} linked libstdc++ statically into libmsq_torch.so; set MSQ_CXX to the system g++")
    os.replace(TORCH_LIB + ".tmp", TORCH_LIB)
    return TORCH_LIB


def build(force=False, verbose=False, defines=(), out=None):
    """Compile every CUDA source (one nvcc per file, in parallel) and link one shared library.
    Returns its path.  ``defines``/``out`` build an experimental variant next to the default library."""
    from concurrent.futures import ThreadPoolExecutor
    out = out or LIB
    if not force and not defines and not is_stale():
        build_torch_binding()
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" + ("_" + "_".join(defines).replace("=", "-") if defines else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    cflags = [f for f in NVCC_FLAGS if f not in ("--shared",)]
    cflags = cflags[:cflags.index("-cudart")]            # link-only flags stay out of the compile step

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        deps = [path] + HEADERS
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
            return obj, None
        cmd = [nvcc] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-D" + d for d in defines] + \
              ["-c", "-o", obj, path]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, err in results:
            if err:
                sys.stderr.write(err)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "shared",
            "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-ldl", "-o", out + ".tmp"] + [o for o, _ in results]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    os.replace(out + ".tmp", out)
    if out == LIB:
        build_torch_binding(force=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
