"""Build libmaz_b200.so in-tree with nvcc for sm_100a (B200).  `python -m mazero_b200.build [--force]`.

The tree kernels are compiled with -fmad=false: the reference is bit-exact x86-64 code without FMA, and
every fp operation that matters is additionally written with explicit *_rn intrinsics.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmaz_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "maz_tree.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


# per-file flags: the tree engine must not contract a*b+c (bit-exact with x86-64 reference code);
# the inference kernels are ordinary floating point and want FMAs.
FILE_FLAGS = {"maz_tree.cu": ["-fmad=false"]}


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, log = [], ""
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", *FILE_FLAGS.get(os.path.basename(src), []),
               "-Xcompiler", "-fPIC,-O2", "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        log += res.stdout + res.stderr
        objs.append(obj)
    cmd = [_nvcc(), *ARCH, "-shared", "-o", OUT, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
