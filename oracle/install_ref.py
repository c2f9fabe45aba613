"""TEST / BENCH INFRASTRUCTURE ONLY — place the UNMODIFIED reference sources of the hot path under git-ignored baseline/_ref/.

    python oracle/install_ref.py          (build container only: needs /root/reference; run by __graft_entry__.build())

The reference (SidRama/Longitudinal-VAE) is a directory of plain Python modules with no setup.py / pyproject, so
`pip install --target baseline/_ref /root/reference` has nothing to install (recorded in DESIGN.md); the equivalent is a
byte-for-byte copy of the modules the path runs through: elbo_functions.py (the bound), kernel_gen.py / kernel_spec.py
(kernel construction) and GP_model.py.  `baseline/_ref/` is listed in .gitignore (reference sources never enter the
history) and NOT in .gpurunignore, so the copy travels to the GPU box, where /root/reference does not exist.

`bench.py --impl reference` imports these files from baseline/_ref (never from this repo's package) with
`oracle/gpytorch_standin` on sys.path for the absent third-party GPyTorch and times
elbo_functions.minibatch_KLD_upper_bound[_iter] + backward + the natural-gradient update on the box's host cores.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LVAE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ("elbo_functions.py", "kernel_gen.py", "kernel_spec.py", "GP_model.py")


def install(verbose=True):
    """Copy FILES from the reference checkout; returns the destination, or None when the checkout is absent."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"install_ref: {REF} not present (GPU box?) — using the copy already under {DST}, if any")
        return DST if all(os.path.exists(os.path.join(DST, f)) for f in FILES) else None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
        fh.write(f"byte-for-byte copies of {', '.join(FILES)} from {REF} (oracle/install_ref.py); not tracked by git\n")
    if verbose:
        print(f"install_ref: {len(FILES)} reference modules under {DST}")
    return DST


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
