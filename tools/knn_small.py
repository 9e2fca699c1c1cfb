"""Tiny kNN cases for compute-sanitizer / quick triage: every (algo, selection class) once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gadm_b200  # noqa
from gadm_b200 import ops
from oracle import knn_oracle as ko
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for algo in ("brute", "grid"):
    for k in (1, 16, 20):
        s = rng.random((2, 300, 3), dtype=np.float32); q = rng.random((2, 130, 3), dtype=np.float32)
        jobs = ops.make_jobs([(0, 0, 0, 300, 130, 130 * k, 300, 130, k, 2)])
        idx = ops.knn3d_jobs(torch.from_numpy(s).to(dev).view(-1, 3), torch.from_numpy(q).to(dev).view(-1, 3), jobs, 2 * 130 * k, algo)
        torch.cuda.synchronize()
        ok = np.array_equal(idx.view(2, 130, k).cpu().numpy().astype(np.int64), ko.knn_port(s, q, k))
        print(algo, k, "OK" if ok else "MISMATCH", flush=True)
