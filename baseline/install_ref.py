"""Installs the UNMODIFIED reference (YiwuZhong/Sub-GC) for the reference arm of bench.py into baseline/_ref/ (git-ignored; it travels
to the GPU box with the repository snapshot like the built .so).

    python baseline/install_ref.py [--src /root/reference]

The reference is plain Python with no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to
build ("neither 'setup.py' nor 'pyproject.toml' found"); the files of its model path are copied as they are instead:
models/ (AttModel, CaptionModel, loss_wrapper, lib/*), misc/__init__.py + misc/utils.py, and the two class-name tables the model's
constructor reads.  Nothing else of the reference is needed to run `models.setup(opt)` and `model(..., mode='sample')` on CPU.
"""
import argparse
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/__init__.py", "models/AttModel.py", "models/CaptionModel.py", "models/loss_wrapper.py", "models/lib/__init__.py",
         "models/lib/gcn_backbone.py", "models/lib/gpn.py", "models/lib/graph_conv.py", "models/lib/graph_conv_unit.py", "misc/__init__.py",
         "misc/utils.py", "data/object_names_1600-0-20.npy", "data/predicate_names_1600-0-20.npy", "LICENSE"]


def install(src="/root/reference", quiet=False):
    if not os.path.isdir(src):
        return False
    h = hashlib.sha256()
    for rel in FILES:
        s, t = os.path.join(src, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            raise FileNotFoundError(s)
        os.makedirs(os.path.dirname(t), exist_ok=True)
        shutil.copyfile(s, t)
        h.update(rel.encode())
        h.update(open(s, "rb").read())
    with open(os.path.join(DST, "INSTALLED"), "w") as fh:
        fh.write(f"source {src}\nsha256 {h.hexdigest()}\nfiles {len(FILES)} (copied unmodified)\n")
    if not quiet:
        print(f"reference installed into {DST} ({len(FILES)} files, sha256 {h.hexdigest()[:16]})")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    sys.exit(0 if install(ap.parse_args().src) else 1)
