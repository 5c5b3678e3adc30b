"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the sampling path, staged so they travel to the GPU box.

Test / measurement infrastructure only.  The reference is pure Python (there is nothing to compile), and
/root/reference does not exist on the GPU box; `gpurun` ships the repo directory including git-ignored files, so this
script copies the handful of reference files the path needs, byte for byte, from /root/reference into oracle/_ref/
(listed in .gitignore: reference sources never enter the history).  `__graft_entry__.build()` runs it whenever
/root/reference is present.  Consumers: `bench.py --impl reference` and bench.py's cpu_baseline leg (the reference's
own classes timed on the host cores, `kind: "reference"`), through oracle/ref_shim.py; without oracle/_ref they fall
back to the pinned restatement oracle/torch_port.py (`kind: "port"`).

    python oracle/make_ref.py            # copy (idempotent); prints the staged files and their sha256
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("DAD_REFERENCE_SRC", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
# the files of SURVEY.md 8(a): U-Net, diffusion, policies, projector builder, dynamics fit (+ the package markers the
# relative imports need)
FILES = [
    "m_diffuser/models/__init__.py",
    "m_diffuser/models/temporal_unet.py",
    "m_diffuser/models/diffusion.py",
    "m_diffuser/guides/__init__.py",
    "m_diffuser/guides/policies.py",
    "m_diffuser/dynamics/projection.py",
    "m_diffuser/dynamics/data_driven.py",
]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC_ROOT, "m_diffuser", "models")):
        if verbose:
            print("reference tree not present at %s: nothing staged" % SRC_ROOT)
        return False
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(SRC_ROOT, rel), os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
    with open(os.path.join(DST_ROOT, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("\n".join(manifest))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
