"""In-tree build of the CUDA library (sm_100a only).  `python -m dynamics_aware_diffusion_b200.build`."""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_NAME = "libdad_b200.so"
LIB_PATH = os.path.join(PKG_DIR, LIB_NAME)
# translation units: (source, extra defines, object name).  conv_chain_kernel is instantiated once per GroupNorm
# width in its own object so that the widths compile in parallel.
UNITS = [("dad_api.cu", [], "dad_api")] + [("chain_inst.cu", ["CHAIN_GW=%d" % gw], "chain_inst_%d" % gw)
                                            for gw in (16, 32, 64, 128, 256)]
OBJ_DIR = os.path.join(PKG_DIR, "_build")


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put it on PATH)")


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    # every file of csrc/ is a dependency (dad_api.cu includes all the .cuh kernels): no hand-kept list to go stale
    deps = glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(REPO_ROOT, "include", "*.h")) + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into libdad_b200.so next to this file.  nvcc cross-compiles without a GPU.
    `defines` / `out` build a tuning variant under another file name (used with DAD_TUNING=1 DAD_LIB_PATH=...)."""
    from concurrent.futures import ThreadPoolExecutor
    target = out or LIB_PATH
    if not force and not out and not _stale():
        return LIB_PATH
    tag = "" if not out else "_" + os.path.splitext(os.path.basename(out))[0]
    os.makedirs(OBJ_DIR, exist_ok=True)
    common = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-I", os.path.join(REPO_ROOT, "include"), "-I", CSRC] + ["-D" + d for d in defines]
    if verbose:
        common += ["-Xptxas", "-v"]

    def compile_one(unit):
        src, defs, name = unit
        obj = os.path.join(OBJ_DIR, name + tag + ".o")
        cmd = common + ["-D" + d for d in defs] + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, cmd, res

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, UNITS))
    for obj, cmd, res in results:
        if verbose:
            print(" ".join(cmd))
            sys.stdout.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", target] + \
        [obj for obj, _, _ in results] + ["-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-q" not in sys.argv, defines=defs, out=outs[0] if outs else None))
