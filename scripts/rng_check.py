"""Host query draw (mimrl_legacy_permutation_head) against np.random.permutation: values, generator state, time."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from mimrl_b200.model import legacy_permutation_head as draw

for N, m, seed in [(1, 1, 0), (2, 1, 1), (5, 5, 2), (63, 3, 8), (64, 64, 9), (65, 1, 10), (1284, 64, 3), (65536, 100, 4),
                   (65537, 1000, 5), (1 << 20, 4096, 6), (1000003, 7, 7)]:
    np.random.seed(seed); np.random.rand(3)
    a = np.random.permutation(N)[:m]; ra = np.random.randint(0, 1 << 30, 5)
    np.random.seed(seed); np.random.rand(3)
    b = draw(N, m, 0); rb = np.random.randint(0, 1 << 30, 5)
    assert np.array_equal(a, b) and np.array_equal(ra, rb), (N, m)
print("ok")
N = 1 << 20
for _ in range(3):
    t = time.perf_counter(); np.random.permutation(N)[:4096]; t1 = time.perf_counter() - t
    t = time.perf_counter(); draw(N, 4096); t2 = time.perf_counter() - t
    print("numpy %.2f ms  ours %.2f ms" % (t1 * 1e3, t2 * 1e3))
