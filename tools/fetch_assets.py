#!/usr/bin/env python
"""Copies the reference's INPUT DATA (scenes/*.json, resources/*) into assets/_ref/.

The reference checkout (/root/reference) exists only in the build container; the GPU box gets a
snapshot of this repo.  assets/_ref/ is git-ignored (reference data stays out of this repo's
history) but is not gpurun-ignored, so it travels with the snapshot like the built .so files.
No reference source code is copied -- only scene descriptions and texture images.
"""
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DST = ROOT / "assets" / "_ref"


def main() -> int:
    if not REF.exists():
        print("reference checkout not present; keeping existing assets/_ref", file=sys.stderr)
        return 0 if DST.exists() else 1
    for sub, pattern in (("scenes", "*.json"), ("resources", "*")):
        (DST / sub).mkdir(parents=True, exist_ok=True)
        for src in sorted((REF / sub).glob(pattern)):
            dst = DST / sub / src.name
            if not dst.exists() or dst.stat().st_size != src.stat().st_size:
                shutil.copyfile(src, dst)
    print(f"assets in {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
