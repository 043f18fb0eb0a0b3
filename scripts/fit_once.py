import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA
d = int(sys.argv[1]) if len(sys.argv) > 1 else 200
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
per = int(sys.argv[3]) if len(sys.argv) > 3 else 100
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rng = np.random.RandomState(1)
q, _ = np.linalg.qr(rng.randn(d, d))
a_b = q * np.sqrt(2.0 * np.exp(-np.arange(d) / (0.15 * d)))[None, :]
z = rng.randn(k, d)
labels = np.repeat(np.arange(k), per).astype(np.uint64)
x = 0.5 + (z @ a_b.T)[labels.astype(np.int64)] + rng.randn(k * per, d)
p = PLDA()
p.fit(x, labels, iters)
p.fit(x, labels, iters)
print(p.fit_timings())
