"""Stage the UNMODIFIED reference for the GPU box (test infrastructure; nothing under h-net-mamba-asr_b200/ imports this).

`/root/reference` exists in the build container only.  `stage()` -- called by `__graft_entry__.build()` when the checkout
is present -- installs the reference package as it lies into `baseline/_ref/` (`pip install --no-index --no-deps --target`,
from a scratch copy because the checkout is read-only) and copies the reference's own test files of the hot path next to
it (`baseline/_ref/ref_tests/`), together with a conftest that this script writes: it puts the reference package and this
repo's package on sys.path, provides empty stand-ins for the two third-party imports that are not installed here
(`editdistance`, `omegaconf`: SURVEY.md §8c) and calls `dcasr_b200.install()` BEFORE any reference module is imported by a
test.  `baseline/_ref/` is git-ignored (reference sources never enter the history) but travels to the GPU box with the
snapshot, where `tests/test_gpu_reference_suite.py` runs those files against the drop-in and trains the reference's own
`ASRModel` under the reference's own `Trainer`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")
TESTS_DST = os.path.join(DST, "ref_tests")
# the reference's tests of the path and of the callers either side of it
HOT_PATH_TESTS = ["test_hnet_chunk.py", "test_mamba_block.py", "test_encoder.py", "test_fixed_pool.py", "test_asr_task.py",
                  "test_ctc.py", "test_loss.py"]

CONFTEST = '''"""Written by oracle/stage_reference.py: run the reference's own tests against the dcasr_b200 drop-in."""
import os, sys, types
HERE = os.path.dirname(os.path.abspath(__file__))
REF_PKG = os.path.dirname(HERE)                                   # baseline/_ref (holds the installed `dcasr`)
REPO = os.path.dirname(os.path.dirname(REF_PKG))
for p in (REF_PKG, os.path.join(REPO, "h-net-mamba-asr_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
for name in ("editdistance", "omegaconf"):                        # not installed in this image; unused by these tests
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.eval = lambda a, b: sum(x != y for x, y in zip(a, b)) + abs(len(a) - len(b))
            m.OmegaConf = type("OmegaConf", (), {})
            sys.modules[name] = m
if os.environ.get("HNB_REF_SUITE_NO_INSTALL") != "1":
    import dcasr_b200
    dcasr_b200.install()

import pytest, torch


@pytest.fixture(autouse=True)
def _hnb_cuda_default_device(request):
    """tests/test_fixed_pool.py and tests/test_ctc.py of the reference create their tensors on the default device (they are
    CPU tests of pure-torch modules); the drop-in is CUDA-only, so those files run with the GPU as default device -- the
    other files set it themselves."""
    if request.module.__name__.split(".")[-1] in ("test_fixed_pool", "test_ctc") and torch.cuda.is_available():
        torch.set_default_device("cuda")
        yield
        torch.set_default_device("cpu")
    else:
        yield
'''


def stage(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF, "src", "dcasr")):
        return False
    if not os.path.isdir(os.path.join(DST, "dcasr")):
        os.makedirs(DST, exist_ok=True)
        with tempfile.TemporaryDirectory() as tmp:
            src = os.path.join(tmp, "reference")
            shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
            cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
                   "--find-links", "/opt/wheelhouse", "--target", DST, src]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("pip install of the reference failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    os.makedirs(TESTS_DST, exist_ok=True)
    n = 0
    for name in HOT_PATH_TESTS:
        s = os.path.join(REF, "tests", name)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(TESTS_DST, name))
            n += 1
    with open(os.path.join(TESTS_DST, "conftest.py"), "w") as f:
        f.write(CONFTEST)
    if verbose:
        print(f"staged the reference package and {n} of its test files under {DST}")
    return True


if __name__ == "__main__":
    stage()
