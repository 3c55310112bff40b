import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import abi_util as au, oracle_bind as oracle
rng = np.random.RandomState(0)
for (m, n, k, ta, tb) in [(3136, 128, 8, True, False), (3136, 128, 8, False, False), (3136, 128, 32, True, False),
                          (3072, 128, 8, True, False), (512, 128, 8, True, False), (3136, 256, 8, True, False),
                          (8, 128, 12, False, True), (8, 3136, 128, False, True)]:
    a = rng.uniform(-1, 1, (k, m) if ta else (m, k)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, k) if tb else (k, n)).astype(np.float32)
    got = au.matmul(a, b, ta, tb); ref = oracle.matmul(a, b, ta, tb)
    bad = np.argwhere(np.abs(got - ref) > 1e-2 * np.abs(ref).max())
    print((m, n, k, ta, tb), "rel", au.rel_err(got, ref), "nan", np.isnan(got).sum(), "bad rows", sorted(set(bad[:, 0]))[:8], len(bad))
