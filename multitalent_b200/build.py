"""In-tree build of libmtb200.so with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot).

`python -m multitalent_b200.build` or `__graft_entry__.build()`.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libmtb200.so")
OBJ_DIR = os.path.join(PKG_DIR, "build")

SOURCES = ["api.cu", "conv_ffma.cu", "conv_umma.cu", "conv_halo.cu", "conv_line.cu", "conv_pw.cu", "head_bwd.cu", "conv_gm.cu", "conv_c1.cu", "wgrad_line.cu", "wgrad_line_s2.cu", "wgrad_rows.cu", "norm.cu", "norm_tma.cu", "loss.cu", "loss_softmax.cu", "sliding.cu", "optim_pack.cu", "dataio.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-I", INCLUDE, "-I", CSRC,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libmtb200.so for sm_100a)")


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every CUDA translation unit to an object (in parallel) and link the shared library."""
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "mtb200.h"))
    procs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src + ".o")
        stamp = obj + ".sha"
        dig = _digest([sp] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", sp, "-o", obj]
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, stamp, dig, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out.decode(errors="replace")))
        else:
            if verbose or out.strip():
                sys.stderr.write(out.decode(errors="replace"))
            with open(stamp, "w") as f:
                f.write(dig)
    if failed:
        raise RuntimeError("libmtb200 build failed")
    if procs or not os.path.exists(LIB_PATH) or force:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout.decode(errors="replace"))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
