"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/wlsqm_b200.h declares; the ctypes table covers them; the product never touches the oracle;
without a CUDA device the compute entry points fail loudly instead of falling back."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    hdr = (ROOT / "include" / "wlsqm_b200.h").read_text()
    return sorted(set(re.findall(r"WLSQM_API\s+[\w\s\*]+?\b(wlsqm_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from wlsqm_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/wlsqm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signature table and header are out of sync"
    assert _lib.lib().wlsqm_b200_abi_version() == 1


def test_number_of_dofs_tables():
    # reference tests/test_package.py:47-53
    import wlsqm_b200 as w
    from wlsqm_b200 import _lib
    for dim, exp in ((1, [1, 2, 3, 4, 5]), (2, [1, 3, 6, 10, 15]), (3, [1, 4, 10, 20, 35])):
        assert [w.number_of_dofs(dim, o) for o in range(5)] == exp
        assert [_lib.lib().wlsqm_number_of_dofs(dim, o) for o in range(5)] == exp
    assert w.number_of_dofs(0, 1) == -1 and w.number_of_dofs(2, 7) == -2
    assert _lib.lib().wlsqm_number_of_dofs(5, 1) == -1 and _lib.lib().wlsqm_number_of_dofs(2, -1) == -2


def test_package_surface_matches_reference():
    # reference tests/test_package.py:24-32: the flat namespace re-exports
    import wlsqm_b200 as w
    for name in ["ALGO_BASIC", "ALGO_ITERATIVE", "WEIGHT_UNIFORM", "WEIGHT_CENTER", "ExpertSolver", "number_of_dofs",
                 "interpolate_fit", "lambdify_fit", "b2_F", "i3_XYZ2", "SIZE3", "i3_0th_end"] + \
                [f"fit_{d}D{it}{m}" for d in (1, 2, 3) for it in ("", "_iterative") for m in ("", "_many", "_many_parallel")]:
        assert hasattr(w, name), name
    assert not hasattr(w, "i1_0th_end") and not hasattr(w, "i2_0th_end")     # defs.pyx:316,346
    assert w.b3_XYZ2 == 1 << w.i3_XYZ2 == 1 << 34
    import wlsqm_b200.utils.lapackdrivers as ld
    for name in ("mgeneral", "mgeneralp", "mgeneralfactor", "mgeneralfactorp", "mgeneralfactored", "mgeneralfactoredp"):
        assert hasattr(ld, name)


def test_argument_validation_needs_no_device():
    import wlsqm_b200 as w
    nk = np.array([5, 5], np.int32)
    od = np.array([1, 1], np.int32)
    kn = np.zeros(2, np.int64)
    wm = np.ones(2, np.int32)
    with pytest.raises(ValueError, match="same length"):
        w.ExpertSolver(2, nk, od[:1], kn, wm)
    with pytest.raises(ValueError, match="Dimension"):
        w.ExpertSolver(5, nk, od, kn, wm)
    with pytest.raises(TypeError, match="integer"):      # typed `int algorithm` (expert.pyx:92-93): what the compiled reference raises
        w.ExpertSolver(2, nk, od, kn, wm, algorithm=None)
    with pytest.raises(ValueError, match="Unknown algorithm"):
        w.ExpertSolver(2, nk, od, kn, wm, algorithm=3)
    with pytest.raises(ValueError, match="ntasks"):
        w.ExpertSolver(2, nk, od, kn, wm, ntasks=0)
    with pytest.raises(ValueError, match="dtype"):
        w.ExpertSolver(2, nk.astype(np.int64), od, kn, wm)
    with pytest.raises(ValueError):
        w.interpolate_fit(np.zeros(2), np.zeros(6), 4, 2, np.zeros((3, 2)))
    with pytest.raises(ValueError):
        w.lambdify_fit(np.zeros(2), np.zeros(6), 2, 9)


def test_no_cpu_fallback_without_device():
    from wlsqm_b200 import _lib
    import wlsqm_b200 as w
    if _lib.lib().wlsqm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    nk = np.array([5], np.int32)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        w.ExpertSolver(2, nk, np.array([1], np.int32), np.zeros(1, np.int64), np.ones(1, np.int32))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        w.fit_2D(np.zeros((5, 2)), np.zeros(5), np.zeros(2), np.zeros(3), None, order=1, knowns=0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        w.interpolate_fit(np.zeros(2), np.zeros(3), 2, 1, np.zeros((4, 2)))


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use oracle/"""
    bad = []
    for p in (ROOT / "python-wlsqm_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".pyx") and p.is_file():
            txt = p.read_text(errors="replace")
            if re.search(r"^\s*(import|from)\s+oracle\b|oracle/|wlsqm_oracle|libwlsqm_oracle|_ref/", txt, re.M):
                bad.append(str(p))
    hdr = (ROOT / "include" / "wlsqm_b200.h").read_text()
    assert "oracle" not in hdr
    assert not bad, bad


def test_as_arr_mirrors_memoryview_checks():
    from wlsqm_b200 import _lib
    a = np.zeros((4, 6))
    assert _lib.as_arr(a, np.float64, 2, "a").strides == (6, 1)
    assert _lib.as_arr(a[:, ::2], np.float64, 2, "a", last_contig=False, allow_copy=True).strides == (3, 1)
    with pytest.raises(ValueError, match="contiguous"):
        _lib.as_arr(a[:, ::2], np.float64, 2, "a")
    with pytest.raises(ValueError, match="dtype"):
        _lib.as_arr(a.astype(np.float32), np.float64, 2, "a")
    with pytest.raises(ValueError, match="dimensions"):
        _lib.as_arr(a, np.float64, 3, "a")
    ro = a.copy()
    ro.setflags(write=False)
    with pytest.raises(ValueError, match="read-only"):
        _lib.as_arr(ro, np.float64, 2, "a", writable=True)
    assert _lib.as_arr(a[::2], np.float64, 2, "a").strides == (12, 1)      # pitched rows are fine


def test_meta_summary_needs_no_device():
    """wlsqm_meta_summary (host-only): max nk, order range and uniformity of the metadata arrays in one pass"""
    from wlsqm_b200 import _lib
    rng = np.random.default_rng(1)
    n = 100_003
    nk = rng.integers(5, 60, n).astype(np.int32)
    od = rng.integers(0, 5, n).astype(np.int32)
    kn = rng.integers(0, 4, n).astype(np.int64)
    wm = rng.integers(1, 3, n).astype(np.int32)
    assert _lib.meta_summary(nk, od, kn, wm) == (int(nk.max()), int(od.min()), int(od.max()), False)
    u = (np.full(n, 30, np.int32), np.full(n, 4, np.int32), np.zeros(n, np.int64), np.full(n, 1, np.int32))
    assert _lib.meta_summary(*u) == (30, 4, 4, True)
    u[2][-1] = 1                      # one case differs in its knowns only
    assert _lib.meta_summary(*u) == (30, 4, 4, False)
    e = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int32))
    assert _lib.meta_summary(*e) == (0, 0, 0, True)
    # batches of 262144 cases and more are scanned in pieces on the library's host threads
    for m in (1, 2, 262143, 262144, 600_001):
        a, b = rng.integers(5, 31, m).astype(np.int32), rng.integers(0, 5, m).astype(np.int32)
        same = bool((a == a[0]).all() and (b == b[0]).all())
        assert _lib.meta_summary(a, b, np.zeros(m, np.int64), np.ones(m, np.int32)) == (int(a.max()), int(b.min()), int(b.max()), same)
        w = np.ones(m, np.int32)
        w[m // 2] = 1 if m == 1 else 2          # a single case in the middle differs in its weighting only
        assert _lib.meta_summary(np.full(m, 7, np.int32), np.full(m, 2, np.int32), np.zeros(m, np.int64), w) == (7, 2, 2, m == 1)


def test_batched_driver_argument_checks_need_no_device():
    """shape / dtype / layout validation of the batched drivers mirrors the memoryview signatures of
    lapackdrivers.pyx (double[::1,:,:] etc.) and happens before any device work"""
    from wlsqm_b200.utils import lapackdrivers as ld
    A = np.zeros((4, 4, 3), order="F")
    b = np.zeros((4, 3), order="F")
    ip = np.zeros((4, 3), dtype=np.int32, order="F")
    for fn, args in ((ld.mgeneralfactor, (np.zeros((4, 4, 3)), ip)),            # C order
                     (ld.msymmetricfactor, (np.zeros((4, 4, 3)), ip)),
                     (ld.msymmetricfactored, (A, ip.astype(np.int64), b)),     # wrong ipiv dtype
                     (ld.mgeneral, (A.astype(np.float32), b)),                   # wrong data dtype
                     (ld.msymmetrize, (np.zeros((4, 4)),))):                     # wrong rank
        with pytest.raises(ValueError):
            fn(*args)
    for fn, args in ((ld.msymmetricfactor, (A, np.zeros((5, 3), dtype=np.int32, order="F"))),
                     (ld.msymmetric, (A, np.zeros((4, 2), order="F"))),
                     (ld.mgeneralfactored, (A, ip, np.zeros((3, 3), order="F"))),
                     (ld.msymmetrize, (np.zeros((4, 5, 3), order="F"),))):
        with pytest.raises(ValueError, match="shape mismatch"):
            fn(*args)
    assert set(ld.__all__) >= {"mgeneral", "mgeneralp", "msymmetric", "msymmetricp", "msymmetricfactor",
                               "msymmetricfactored", "msymmetrize", "msymmetrizep"}


def test_constructor_errors_match_the_live_reference():
    """the same bad arguments to the unmodified reference (oracle/_ref, where it is built) and to the drop-in:
    same exception type for every check ExpertSolver makes before it touches its native state (expert.pyx:131-159)"""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as orc
    ref = orc.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref is not built here")
    import wlsqm_b200 as w
    nk, od, kn, wm = np.full(2, 6, np.int32), np.full(2, 1, np.int32), np.zeros(2, np.int64), np.ones(2, np.int32)
    table = [
        ("length mismatch", (2, nk, od[:1], kn, wm), {}),
        ("bad dimension", (5, nk, od, kn, wm), {}),
        ("algorithm None", (2, nk, od, kn, wm), {"algorithm": None}),
        ("unknown algorithm", (2, nk, od, kn, wm), {"algorithm": 3}),
        ("ntasks 0", (2, nk, od, kn, wm), {"ntasks": 0}),
        ("max_iter None", (2, nk, od, kn, wm), {"max_iter": None}),
        ("do_sens None", (2, nk, od, kn, wm), {"do_sens": None}),
        ("nk int64", (2, nk.astype(np.int64), od, kn, wm), {}),
        ("knowns int32", (2, nk, od, kn.astype(np.int32), wm), {}),
        ("empty batch", (2, nk[:0], od[:0], kn[:0], wm[:0]), {}),
        ("nk 2-D", (2, nk.reshape(1, 2), od, kn, wm), {}),
        ("dimension None", (None, nk, od, kn, wm), {}),
    ]
    import warnings
    for what, args, kw in table:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")      # (the reference's __del__ complains about its half-built object)
            with pytest.raises(Exception) as e_ref:
                ref.ExpertSolver(*args, **kw)
        with pytest.raises(Exception) as e_new:
            w.ExpertSolver(*args, **kw)
        assert issubclass(e_new.type, e_ref.type), (what, e_ref.type, e_new.type, str(e_ref.value), str(e_new.value))
    # None for a typed int: the compiled signature of the reference refuses it with TypeError before its own
    # `cannot be None` checks can run (expert.pyx:92-93 vs :140-149) -- the Cython shim here does exactly the same
    with pytest.raises(TypeError, match="an integer is required"):
        w.ExpertSolver(2, nk, od, kn, wm, max_iter=None)
    with pytest.raises(TypeError, match="an integer is required"):
        w.fit_2D_many(np.zeros((2, 6, 2)), np.zeros((2, 6)), nk, np.zeros((2, 2)), np.zeros((2, 3)), None, None, od, kn, wm)


def test_cython_shim_is_the_binding_and_ctypes_is_the_fallback():
    """the shipped binding is the Cython C-ABI shim (wlsqm_b200/_shim.pyx, built by __graft_entry__.build()); with
    WLSQM_BINDING=ctypes the ctypes twins serve the same names and the same validation errors"""
    import subprocess
    import sys
    import wlsqm_b200 as w
    from wlsqm_b200.fitter import expert, simple
    assert expert.BINDING == "cython" and simple.BINDING == "cython", "build the shim: python python-wlsqm_b200/build_shim.py"
    assert w.ExpertSolver.__module__ == "wlsqm_b200._shim" and w.fit_3D_iterative_many_parallel.__module__ == "wlsqm_b200._shim"
    assert w.fit_2D_many_parallel.__name__ == "fit_2D_many_parallel"
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):          # Cython's own memoryview coercion error
        w.fit_2D_many_parallel(np.zeros((3, 4, 2), np.float32), np.zeros((3, 4)), np.full(3, 4, np.int32), np.zeros((3, 2)),
                               np.zeros((3, 6)), None, 0, np.full(3, 2, np.int32), np.zeros(3, np.int64), np.ones(3, np.int32))
    with pytest.raises(ValueError, match="wrong number of dimensions"):
        w.fit_2D_many_parallel(np.zeros((3, 4)), np.zeros((3, 4)), np.full(3, 4, np.int32), np.zeros((3, 2)),
                               np.zeros((3, 6)), None, 0, np.full(3, 2, np.int32), np.zeros(3, np.int64), np.ones(3, np.int32))
    with pytest.raises(TypeError):
        w.fit_2D_many_parallel(np.zeros((3, 4, 2)), np.zeros((3, 4)), np.full(3, 4, np.int32), np.zeros((3, 2)),
                               np.zeros((3, 6)), None, None, np.full(3, 2, np.int32), np.zeros(3, np.int64), np.ones(3, np.int32))
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, wlsqm_b200 as w\n"
            "from wlsqm_b200.fitter import expert, simple\n"
            "assert expert.BINDING == 'ctypes' and simple.BINDING == 'ctypes' and w.ExpertSolver.__module__.endswith('expert')\n"
            "try:\n    w.ExpertSolver(2, np.zeros(2, np.int32), np.zeros(2, np.int32), np.zeros(2, np.int64), np.ones(2, np.int32), algorithm=None)\n"
            "except TypeError: print('ok')\n" % str(ROOT / "python-wlsqm_b200"))
    import os
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, WLSQM_BINDING="ctypes"))
    assert r.stdout.strip() == "ok", r.stderr[-600:]
