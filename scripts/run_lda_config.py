"""BASELINE configs[4]: LDA 1M x 200, 5k classes, fit + predict_log_proba over 1M test vectors on one B200."""
import os, sys, time, json
import numpy as np
import torch
import ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import LDA, _ffi

n, d, k, nt = 1_000_000, 200, 5_000, 1_000_000
rng = np.random.RandomState(3)
centers = (rng.randn(k, d) * 0.7).astype(np.float32)
y = (np.arange(n) % k).astype(np.int64)
x = centers[y] + rng.randn(n, d).astype(np.float32)
m = LDA()
t0 = time.perf_counter(); m.fit(x, y); fit_host_s = time.perf_counter() - t0
t0 = time.perf_counter(); m.fit(x, y); fit_host_s = time.perf_counter() - t0
dev = torch.device("cuda", 0)
xt = torch.from_numpy(centers[rng.randint(0, k, nt)] + rng.randn(nt, d).astype(np.float32)).to(dev)
out = m.predict_log_proba(xt)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = m.predict_log_proba(xt)
torch.cuda.synchronize(); pred_s = time.perf_counter() - t0
torch.cuda.synchronize(); t0 = time.perf_counter()
dec = m.decision_function(xt)
torch.cuda.synchronize(); dec_s = time.perf_counter() - t0
lp = out[:2000].double().cpu().numpy()
print(json.dumps({"config": "c5", "n": n, "d": d, "classes": k, "nt": nt, "fit_host_rows_s": fit_host_s,
                  "predict_log_proba_s": pred_s, "rows_per_s": nt / pred_s,
                  "algorithmic_tflops": 2.0 * d * k * nt / pred_s / 1e12,
                  "decision_function_s": dec_s, "out_gb": nt * k * 4 / 1e9,
                  "sum_exp_rows_close_to_1": bool(np.allclose(np.exp(lp).sum(1), 1.0, atol=2e-3)),
                  "mem_gb": torch.cuda.max_memory_allocated() / 1e9}))
