"""Build libcidnet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python hvi-cidnet_b200/build.py [--force] [--bf16]

One translation unit per .cu under csrc/, compiled in parallel, linked with
`nvcc -shared`.  No torch, no cuDNN/cuBLAS: only the CUDA runtime (static) --
the driver API entry point for TMA descriptors is fetched at run time through
cudaGetDriverEntryPoint, so libcuda is not needed at link time.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libcidnet_b200.so")
LIB_BF16 = os.path.join(HERE, "libcidnet_b200_bf16.so")     # the -DCIDNET_ACT_BF16 build (select with CIDNET_LIB=<path>)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_digest(extra):
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()


def compile_one(src, extra, hdig, force, verbose, objdir=OBJ):
    obj = os.path.join(objdir, src[:-3] + ".o")
    stamp = obj + ".stamp"
    sdig = hashlib.sha1(open(os.path.join(CSRC, src), "rb").read()).hexdigest() + hdig
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == sdig:
        return src, "", False
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(sdig)
    return src, r.stderr, True


def build(force=False, bf16=False, verbose=False, groups3=False):
    """groups3: experimental build with a third epilogue warpgroup in the 1x1 conv GEMMs (libcidnet_b200_g3.so; see
    conv_gemm.cu launch_conv_gemm) -- used once per round to re-check that experiment, never shipped as the default."""
    objdir = OBJ + ("_bf16" if bf16 else "") + ("_g3" if groups3 else "")
    lib = LIB_BF16 if bf16 else LIB
    if groups3:
        lib = lib[:-3] + "_g3.so"
    os.makedirs(objdir, exist_ok=True)
    extra = (["-DCIDNET_ACT_BF16"] if bf16 else []) + (["-DCIDNET_GEMM_GROUPS3"] if groups3 else [])
    hdig = headers_digest(extra)
    srcs = sources()
    rebuilt = False
    logs = {}
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, log, did in ex.map(lambda s: compile_one(s, extra, hdig, force, verbose, objdir), srcs):
            rebuilt |= did
            logs[src] = log
    if rebuilt or not os.path.exists(lib):
        objs = [os.path.join(objdir, s[:-3] + ".o") for s in srcs]
        cmd = [nvcc_path(), "-shared", "-o", lib] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for s, log in logs.items():
            if log:
                print(f"==== {s}\n{log}")
    return lib


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--groups3", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.bf16, a.verbose, a.groups3))
