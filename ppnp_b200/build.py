"""Build libppnp_b200.so (all csrc/*.cu, sm_100a only) in-tree with nvcc."""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libppnp_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return cand


def source_hash(paths):
    """Content hash of the build inputs: mtimes do not survive the copy to the GPU box."""
    import hashlib
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps(sources):
    return sources + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(_HERE, "..", "include", "*.h"))


def _stale(sources):
    if not os.path.exists(OUT) or not os.path.exists(OUT + ".srchash"):
        return True
    with open(OUT + ".srchash") as f:
        return f.read().strip() != source_hash(_deps(sources))


def build_library(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (e.g. defines=['PPNP_SPMM_U4=8'], out='libvariant.so')
    that ppnp_b200._lib loads when PPNP_B200_LIB points at it."""
    if out is not None:
        return _build(sorted(glob.glob(os.path.join(CSRC, '*.cu'))), list(defines), out, verbose, tag=os.path.basename(out))
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not sources:
        raise RuntimeError("no CUDA sources found")
    if not force and not _stale(sources):
        return OUT
    _build(sources, [], OUT, verbose, tag="")
    with open(OUT + ".srchash", "w") as f:
        f.write(source_hash(_deps(sources)))
    return OUT


def _build(sources, defines, out_path, verbose, tag):
    objdir = os.path.join(_HERE, "csrc", "_obj" + ("_" + tag if tag else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs, procs = [], []
    for s in sources:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = [nvcc, *NVCC_FLAGS, *["-D" + d for d in defines], "-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(s)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
    cmd = [nvcc, "-shared", "-o", out_path, *objs, "-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return out_path


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
