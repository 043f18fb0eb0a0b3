import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
z = torch.randn(k, d, device=dev, generator=g)
x = (0.5 + z).repeat_interleave(n // k, dim=0) + torch.randn(n, d, device=dev, generator=g)
labels = np.repeat(np.arange(k), n // k).astype(np.uint64)
p = PLDA()
p.fit(x, labels, 2); p.fit(x, labels, 2)
print(p.fit_timings())
