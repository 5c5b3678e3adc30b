"""In-tree build of the CUDA library (sm_100a only).  `python -m dynamics_aware_diffusion_b200.build`."""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libdad_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
SOURCES = ["dad_api.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "conv_tc.cuh", "kernels_f32.cuh", "step_kernel.cuh"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put it on PATH)")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(REPO_ROOT, "include", "dad_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into libdad_b200.so next to this file.  nvcc cross-compiles without a GPU.
    `defines` / `out` build a tuning variant under another file name (selected at run time with DAD_LIB_PATH)."""
    target = out or LIB_PATH
    if not force and not out and not _stale():
        return LIB_PATH
    cmd = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(REPO_ROOT, "include"), "-I", CSRC,
           "-o", target] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stdout.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-q" not in sys.argv, defines=defs, out=outs[0] if outs else None))
