"""Make the reference's own hot-path files available where /root/reference is not mounted (the GPU box).

TEST INFRASTRUCTURE.  ``vendor()`` copies ``openset_imagenet/{losses,metrics,util,model}.py`` VERBATIM (byte for
byte, sha256 recorded in ``MANIFEST.json``) from the read-only reference checkout into ``oracle/_ref/openset_imagenet/``.
That directory is git-ignored -- no reference source enters this repository's history -- but it is not gpurun-ignored,
so it travels to the GPU box together with the built ``libosb.so``.  There ``oracle/ref_loader.py`` loads the files
exactly as it loads them from /root/reference here (same ``vast.tools`` / ``matplotlib`` stubs, which carry no
arithmetic), and ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's OWN code on the box's host
cores (``kind: "reference"``).  ``__graft_entry__.build()`` calls this whenever the reference checkout is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("OSB_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ("losses.py", "metrics.py", "util.py", "model.py")


def vendor(quiet: bool = False) -> bool:
    """Copy the files if the reference checkout is present.  Returns True when oracle/_ref is complete afterwards."""
    src_pkg = os.path.join(SRC_ROOT, "openset_imagenet")
    dst_pkg = os.path.join(DST_ROOT, "openset_imagenet")
    if os.path.isfile(os.path.join(src_pkg, "losses.py")):
        os.makedirs(dst_pkg, exist_ok=True)
        manifest = {"source": SRC_ROOT, "files": {}}
        for name in FILES:
            shutil.copyfile(os.path.join(src_pkg, name), os.path.join(dst_pkg, name))
            manifest["files"][name] = hashlib.sha256(open(os.path.join(dst_pkg, name), "rb").read()).hexdigest()
        with open(os.path.join(DST_ROOT, "MANIFEST.json"), "w") as f:
            json.dump(manifest, f, indent=1)
        if not quiet:
            print(f"vendored {len(FILES)} reference files into {dst_pkg}")
    return all(os.path.isfile(os.path.join(dst_pkg, n)) for n in FILES)


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
