"""TEST INFRASTRUCTURE - recipe that places the UNMODIFIED reference package next to the oracle.

The reference (SeonghwanSeo/PharmacoNet) is pure Python + numba: there is nothing to compile, but `/root/reference`
does not exist on the GPU box. This recipe copies `src/pmnet` of the reference, byte for byte, into the git-ignored
`oracle/_ref/pmnet` (it travels with `gpurun` like a built `.so`; it never enters the history), so that

  * `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own numba path
    (`GraphMatcher.run` under `multiprocessing.Pool`, screening.py:46-68) on the bench box's host cores, and
  * `tools/cnn_bench.py` can time the reference's own `nn.Module`s through torch / cuDNN on the same GPU.

`__graft_entry__.build()` runs it whenever `/root/reference` is present. Nothing under `pharmaconet_b200/` reads it.
"""

from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PMNET_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def make_ref(verbose: bool = True) -> str | None:
    src = os.path.join(SRC, "src", "pmnet")
    if not os.path.isdir(src):
        if verbose:
            print(f"make_ref: {src} not present (GPU box): keeping {DST} as shipped")
        return DST if os.path.isdir(os.path.join(DST, "pmnet")) else None
    dst = os.path.join(DST, "pmnet")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    cmp = filecmp.dircmp(src, dst, ignore=["__pycache__"])
    assert not cmp.diff_files and not cmp.left_only, "oracle/_ref/pmnet differs from the reference"
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of /root/reference/src/pmnet made by oracle/make_ref.py (git-ignored; ships via gpurun).\n")
    if verbose:
        print(f"make_ref: copied {src} -> {dst}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
