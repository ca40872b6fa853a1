"""Put a travelling copy of the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that the
reference-literal CPU baseline — the reference's own per-environment loop
    u = controller.get_control_efforts(x); x = dynamics.simulate(x, u)      (scripts/test_vhjb_policy.py:146-151)
— can be timed on the GPU box's host cores, where /root/reference does not exist.

The reference is a plain Python package without dependencies of its own in setup.py (install_requires=[]); its imports of
jax / gin / matplotlib are satisfied by the inert stand-ins of oracle/_stubs (every isinstance(x, jnp.ndarray) is False, so
the reference runs its NumPy branch, unmodified).  First choice is the contract's offline pip install into baseline/_ref;
the reference's setup.py names no packages, so pip installs nothing importable and the four package directories are
copied instead (only .py and .gin files).  Nothing from the reference enters git history.

    python tools/install_reference_baseline.py [--src /root/reference]
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("dynamics", "controller", "configs", "utils")


def install(src: str = "/root/reference", verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(src, "dynamics")):
        raise SystemExit(f"reference tree not found at {src}")
    os.makedirs(DEST, exist_ok=True)
    note = []
    with tempfile.TemporaryDirectory() as tmp:          # the source tree is read-only: pip builds in a copy
        work = os.path.join(tmp, "ref")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "*.ipynb", "*.mat"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DEST, work]
        r = subprocess.run(cmd, capture_output=True, text=True)
        note.append(f"pip install rc={r.returncode}")
    importable = all(os.path.isfile(os.path.join(DEST, p, "__init__.py")) or os.path.isdir(os.path.join(DEST, p)) for p in PACKAGES)
    if not importable:
        for pkg in PACKAGES:
            dst = os.path.join(DEST, pkg)
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(os.path.join(src, pkg), dst,
                            ignore=lambda d, names: [n for n in names if not (n.endswith((".py", ".gin")) or
                                                                              os.path.isdir(os.path.join(d, n)))])
        note.append("package directories copied (.py, .gin)")
    with open(os.path.join(DEST, "SOURCE.txt"), "w") as fh:
        fh.write(f"unmodified copy of {src} for the CPU baseline; {'; '.join(note)}\n")
    if verbose:
        print(f"{DEST}: {'; '.join(note)}")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    install(ap.parse_args().src)
