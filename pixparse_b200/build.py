"""In-tree build of the sm_100a C-ABI library (`pixparse_b200/csrc/libpixparse_b200.so`).

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with the
gpurun snapshot. Only files whose source (or a header) changed are recompiled.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
LIB_PATH = os.path.join(CSRC, "libpixparse_b200.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src, hdr_digest, verbose):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = obj + ".stamp"
    with open(os.path.join(CSRC, src), "rb") as fh:
        digest = hashlib.sha256(fh.read() + hdr_digest.encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, False, ""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(res.stderr)
    return obj, True, res.stderr if verbose else ""


def build(verbose=False, force=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            if f.endswith(".stamp"):
                os.remove(os.path.join(OBJ_DIR, f))
    hdr = _headers_digest()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, hdr, verbose), srcs))
    objs = [r[0] for r in results]
    rebuilt = any(r[1] for r in results)
    if verbose:
        for r in results:
            if r[2]:
                print(r[2])
    if rebuilt or not os.path.exists(LIB_PATH):
        cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


def build_variant(name, extra_flags, only=None):
    """A/B builds: csrc/libab_<name>.so compiled with extra nvcc flags (e.g. -DB200_GEMM_NAUX=1); loaded through
    PIXPARSE_B200_LIB (scripts/gpu_ab.sh, scripts/gpu_gemm_ab.sh). Objects of files not in `only` are reused."""
    out_dir = os.path.join(CSRC, "build", "ab_" + name)
    os.makedirs(out_dir, exist_ok=True)
    build()
    objs = []
    for src in _sources():
        if only is not None and src not in only:
            objs.append(os.path.join(OBJ_DIR, src[:-3] + ".o"))
            continue
        obj = os.path.join(out_dir, src[:-3] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
        objs.append(obj)
    lib = os.path.join(CSRC, f"libab_{name}.so")
    res = subprocess.run([_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)
