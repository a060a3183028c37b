#!/usr/bin/env python
"""Run the REFERENCE's own pytest suite against wlsqm_b200 (`import wlsqm` aliased to the drop-in).

The reference's tests are not part of this repository.  In the build container,
    python tools/reference_tests_against_b200.py --stage
copies /root/reference/tests into oracle/_ref/_reftests (git-ignored, travels to the GPU box with the compiled
reference); on the GPU box,
    python tools/reference_tests_against_b200.py --run
runs the in-scope files (simple / expert / interp / edge cases / stencil / noise robustness / parallel / package /
lapackdrivers) and prints pytest's summary; --unstage removes the copy again.  Out of scope by design (SURVEY.md 8b,
DESIGN.md 8): test_cimport.py (the .pxd-level API).  test_package.py (import graph, version, ScalingAlgo,
number_of_dofs) needs no device and passes too.
"""
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
STAGE = ROOT / "oracle" / "_ref" / "_reftests"
IN_SCOPE = ["test_simple.py", "test_expert.py", "test_interp.py", "test_edge_cases.py", "test_stencil.py",
            "test_noise_robustness.py", "test_parallel.py", "test_package.py", "test_lapackdrivers.py"]


def main():
    if "--stage" in sys.argv:
        src = Path("/root/reference/tests")
        STAGE.mkdir(parents=True, exist_ok=True)
        for f in src.glob("*.py"):
            shutil.copy(f, STAGE / f.name)
        print("staged", len(list(STAGE.glob("*.py"))), "files in", STAGE)
    if "--run" in sys.argv:
        files = [str(STAGE / f) for f in IN_SCOPE if (STAGE / f).exists()]
        if not files:
            sys.exit("no staged reference tests (run --stage in the build container first)")
        env_path = str(ROOT / "tools" / "refshim")
        extra = [a for a in sys.argv[1:] if a not in ("--stage", "--run", "--unstage")]
        cmd = [sys.executable, "-m", "pytest", "-q", *extra, "-p", "wlsqm_alias_plugin", "--rootdir", str(STAGE), "-o",
               "cache_dir=/tmp/refshim_cache", *files]
        import os
        env = dict(os.environ)
        env["PYTHONPATH"] = env_path + os.pathsep + env.get("PYTHONPATH", "")
        sys.exit(subprocess.call(cmd, env=env, cwd=str(STAGE)))
    if "--unstage" in sys.argv:
        shutil.rmtree(STAGE, ignore_errors=True)
        print("removed", STAGE)


if __name__ == "__main__":
    main()
