"""tests/test_gpu_parity.py::test_random_knowns_masks_orders_and_sizes over more seeds than the test suite runs"""
import sys, traceback
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/oracle", "/root/repo/python-wlsqm_b200"]
import test_gpu_parity as t
seeds = range(int(sys.argv[1]), int(sys.argv[2]))
bad = 0
for dim in (1, 2, 3):
    for algo in (1, 2):
        for seed in seeds:
            try:
                t.test_random_knowns_masks_orders_and_sizes.__wrapped__(dim, algo, seed) if hasattr(
                    t.test_random_knowns_masks_orders_and_sizes, "__wrapped__") else t.test_random_knowns_masks_orders_and_sizes(dim, algo, seed)
                print("ok   %dD algo %d seed %d" % (dim, algo, seed), flush=True)
            except Exception as exc:      # noqa: BLE001
                bad += 1
                print("FAIL %dD algo %d seed %d: %s" % (dim, algo, seed, str(exc)[:3000]), flush=True)
print("failures:", bad)
