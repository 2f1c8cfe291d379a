"""TEST INFRASTRUCTURE — recipe that makes the reference itself travel to the GPU box.

The reference's hot path is pure Python (no build system, nothing to compile), so "building" it for the CPU arm of
bench.py means placing the UNMODIFIED module files where the harness can import them on a box that has no
/root/reference:

    python oracle/build_ref.py        # run in the build container (also called by __graft_entry__.build())

copies  mpu/*.py, model/{__init__,gpt2_modeling,distributed}.py, vqvae/*.py (generation/sampling.py imports
pretrain_gpt2 / data_utils -> lmdb, tensorboardX, which this image lacks; its loop semantics are exercised through
GPT2Model.forward with hidden-state mems instead) from /root/reference into oracle/_ref/ (git-ignored: reference sources never enter this repository's history; NOT
gpurun-ignored, so the directory ships with the snapshot like a built .so) and writes MANIFEST.json with the
sha256 of every file.  oracle/ref_harness.py falls back to oracle/_ref when /root/reference is absent, with the
same four shims; bench.py --impl reference then times the reference's own modules (cpu_baseline.kind = "reference").
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("COGVIEW_REFERENCE", "/root/reference")
FILES = {
    "mpu": None,                         # every .py of the package
    "vqvae": None,
    "model": ["__init__.py", "gpt2_modeling.py", "distributed.py"],
}


def build(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "mpu")):
        if verbose:
            print("oracle/build_ref: no reference tree at %s (GPU box?) — keeping the prebuilt oracle/_ref" % SRC)
        return os.path.isdir(os.path.join(DEST, "mpu"))
    manifest = {}
    for pkg, names in FILES.items():
        sdir, ddir = os.path.join(SRC, pkg), os.path.join(DEST, pkg)
        os.makedirs(ddir, exist_ok=True)
        if names is None:
            names = sorted(f for f in os.listdir(sdir) if f.endswith(".py"))
        for n in names:
            s = os.path.join(sdir, n)
            if not os.path.exists(s):
                continue
            shutil.copyfile(s, os.path.join(ddir, n))
            manifest["%s/%s" % (pkg, n)] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    if verbose:
        print("oracle/build_ref: %d reference files -> %s" % (len(manifest), DEST))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
