"""Build-time recipe for `oracle/_ref/` (TEST / BASELINE INFRASTRUCTURE, never part of the product path).

The reference is a set of Python scripts with no build system, so "building" it for the reference arm of bench.py means
placing its two hot-path modules where the GPU box can import them: `/root/reference` does not exist there, while
`oracle/_ref/` travels with the snapshot like the built `.so` files (it is listed in .gitignore, so the reference's
sources never enter this repository's history).  `__graft_entry__.build()` calls `make_ref()` whenever `/root/reference`
is mounted.  With it in place `bench.py --impl reference` times the UNMODIFIED reference `LowLightEnhance`
(`compute_loss` -> `backward` -> `optimizer.step`, model.py:313-316) and reports `kind: "reference"`; without it the arm
falls back to the pinned restatement `oracle/sshslie_oracle.py` (`kind: "port"`).
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("model.py", "utils.py", "metrics.py")    # model.py:12-13 imports metrics and utils; nothing else is on the path


def make_ref():
    """Copy the reference's hot-path modules into oracle/_ref/ (returns the directory, or None if not mounted)."""
    if not all(os.path.exists(os.path.join(REF_SRC, f)) for f in FILES):
        return None
    os.makedirs(REF_DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DST, f))
    return REF_DST


def ref_dir():
    """oracle/_ref if the build has populated it, else None."""
    return REF_DST if all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES) else None


if __name__ == "__main__":
    print(make_ref())
