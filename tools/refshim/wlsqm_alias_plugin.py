"""pytest plugin: `import wlsqm` resolves to wlsqm_b200 (the drop-in), so that the REFERENCE's own test-suite runs
against the B200 implementation unchanged.  Used by tools/reference_tests_against_b200.py only."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT / "python-wlsqm_b200")]

import wlsqm_b200                                   # noqa: E402
import wlsqm_b200.fitter.defs                       # noqa: E402,F401
import wlsqm_b200.fitter.simple                     # noqa: E402,F401
import wlsqm_b200.fitter.expert                     # noqa: E402,F401
import wlsqm_b200.fitter.interp                     # noqa: E402,F401
import wlsqm_b200.utils.lapackdrivers               # noqa: E402,F401

sys.modules["wlsqm"] = wlsqm_b200
for name, mod in list(sys.modules.items()):
    if name.startswith("wlsqm_b200."):
        sys.modules["wlsqm." + name[len("wlsqm_b200."):]] = mod
