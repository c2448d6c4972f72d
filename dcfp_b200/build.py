"""In-tree build of the native pieces (no JIT cache: the built .so files travel with the repo).

  dcfp_b200/lib/libdcfp_b200.so   hand-written sm_100a kernels + the C ABI (nvcc, static cudart)
  dcfp_b200/lib/dcfp_torch_ops.so TORCH_LIBRARY binding torch.ops.dcfp.* over that C ABI (g++)

`python -m dcfp_b200.build` rebuilds what is stale; nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
INCLUDE = os.path.join(ROOT, "include")
ABI_LIB = os.path.join(LIBDIR, "libdcfp_b200.so")
OPS_LIB = os.path.join(LIBDIR, "dcfp_torch_ops.so")

CU_SOURCES = ["class_stats.cu", "bn_fused.cu", "eic_select.cu", "gather.cu", "balance.cu", "abi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-extended-lambda",
              "-Xcompiler", "-fPIC", "-cudart", "static"]
if os.environ.get("DCFP_K1_TRACE") == "1":  # development build: per-phase timestamps inside the fused K1 kernel (scripts/k1_trace.py)
    NVCC_FLAGS.append("-DDCFP_K1_TRACE")


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = os.path.join(cuda_home, "bin", "nvcc")
    return exe if os.path.exists(exe) else (shutil.which("nvcc") or "nvcc")


def _digest(sources, extra=""):
    """Content hash of the inputs of a build step (file names + bytes + flags)."""
    import hashlib
    h = hashlib.sha256(extra.encode())
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, sources, extra=""):
    """Rebuild when the target is missing or was built from different inputs.  Content hashes, not mtimes: the repo
    travels to the GPU box as a snapshot whose timestamps say nothing about what the shipped .so was built from."""
    stamp = target + ".stamp"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    with open(stamp) as f:
        return f.read().strip() != _digest(sources, extra)


def _write_stamp(target, sources, extra=""):
    with open(target + ".stamp", "w") as f:
        f.write(_digest(sources, extra))


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("build command failed (%d):\n%s\n%s" % (proc.returncode, " ".join(cmd), proc.stdout))
    if verbose and proc.stdout.strip():
        print(proc.stdout)


def build_abi(force=False, verbose=False, ptxas_verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(INCLUDE, "dcfp_b200.h")]
    flags = " ".join(NVCC_FLAGS)
    if not force and not _stale(ABI_LIB, deps, flags):
        return ABI_LIB
    objs = []
    procs = []
    for s in srcs:  # one nvcc per translation unit, in parallel
        o = os.path.join(LIBDIR, os.path.basename(s)[:-3] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_verbose else []) + ["-I", INCLUDE, "-I", CSRC, "-c", s, "-o", o]
        if verbose:
            print("+", " ".join(cmd), flush=True)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), out))
        if (verbose or ptxas_verbose) and out.strip():
            print(out)
    _run([_nvcc(), "-shared", "-cudart", "static", "-o", ABI_LIB] + objs, verbose)
    for o in objs:
        os.remove(o)
    _write_stamp(ABI_LIB, deps, flags)
    return ABI_LIB


def build_torch_ops(force=False, verbose=False):
    import torch
    from torch.utils import cpp_extension as ce

    src = os.path.join(CSRC, "torch_binding.cpp")
    deps = [src, os.path.join(INCLUDE, "dcfp_b200.h")]
    extra = "torch " + torch.__version__ + " abi " + (open(ABI_LIB + ".stamp").read() if os.path.exists(ABI_LIB + ".stamp") else "")
    if not force and not _stale(OPS_LIB, deps, extra):
        return OPS_LIB
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = []
    for p in ce.include_paths() + [os.path.join(cuda_home, "include"), INCLUDE]:
        inc += ["-I", p]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
            "-DTORCH_API_INCLUDE_EXTENSION_H", "-w"] + inc + [src, "-o", OPS_LIB, "-L", LIBDIR, "-ldcfp_b200", "-L", torch_lib,
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib])
    _run(cmd, verbose)
    _write_stamp(OPS_LIB, deps, extra)
    return OPS_LIB


def build_all(force=False, verbose=False):
    build_abi(force=force, verbose=verbose)
    build_torch_ops(force=force, verbose=verbose)
    return ABI_LIB, OPS_LIB


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built:", ABI_LIB, OPS_LIB)
