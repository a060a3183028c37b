"""the randomized parity tests of tests/test_gpu_parity.py over more seeds than the test suite runs, and under the library's
A/B switches (shared-memory prepare kernel, per-order launches from 500 cases on, 1024-case staging chunks);
usage: python tools/fuzz_sweep.py SEED_LO SEED_HI"""
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/oracle", "/root/repo/python-wlsqm_b200"]
import test_gpu_parity as t

seeds = range(int(sys.argv[1]), int(sys.argv[2]))
MODES = {"default": {}, "smem-prepare": {"WLSQM_PREP_KERNEL": "smem"}, "bucketed-prepare": {"WLSQM_PREP_BUCKET_MIN": "500"},
         "small-chunks": {"WLSQM_SOLVE_CHUNK": "1024"}}
bad = 0
for mode, env in MODES.items():
    for k in ("WLSQM_PREP_KERNEL", "WLSQM_PREP_BUCKET_MIN", "WLSQM_SOLVE_CHUNK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for seed in (seeds if mode == "default" else list(seeds)[:2]):
        for dim in (1, 2, 3):
            for algo in (1, 2):
                try:
                    t.test_random_knowns_masks_orders_and_sizes(dim, algo, seed)
                    print("ok   [%s] masks %dD algo %d seed %d" % (mode, dim, algo, seed), flush=True)
                except Exception as exc:      # noqa: BLE001
                    bad += 1
                    print("FAIL [%s] masks %dD algo %d seed %d: %s" % (mode, dim, algo, seed, str(exc)[:2500]), flush=True)
        for (dim, order, k, nkn, algo) in ((1, 4, 9, 2, 1), (2, 3, 24, 3, 1), (2, 4, 30, 5, 2), (3, 4, 60, 34, 1), (3, 3, 40, 7, 1),
                                          (2, 1, 8, 1, 1), (3, 1, 12, 2, 2), (1, 2, 6, 1, 2)):
            try:
                t.test_random_knowns_patterns_of_equal_count(dim, order, k, nkn, algo, seed)
                print("ok   [%s] equal-count %dD o%d nkn %d algo %d seed %d" % (mode, dim, order, nkn, algo, seed), flush=True)
            except Exception as exc:      # noqa: BLE001
                bad += 1
                print("FAIL [%s] equal-count %dD o%d nkn %d algo %d seed %d: %s" % (mode, dim, order, nkn, algo, seed, str(exc)[:2500]), flush=True)
print("failures:", bad)
