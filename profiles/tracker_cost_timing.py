import sys, time
sys.path.insert(0, '.')
import numpy as np
from swiftwatcher_b200 import segment_tracking as st
rng = np.random.default_rng(3)
for n in (50, 500, 540):
    ws = st.CostWorkspace(0, max_segments=2048)
    p = rng.uniform(0, 1080, (n, 2)); c = p + rng.normal(0, 5, (n, 2)); first = p - rng.normal(0, 20, (n, 2))
    has = (rng.random(n) < 0.7).astype(np.uint8)
    for _ in range(5): ws.costs(p, first, has, c)
    t0 = time.perf_counter()
    for _ in range(50): m = ws.costs(p, first, has, c)
    dt = (time.perf_counter() - t0) / 50
    class S:
        def __init__(self, cen, hist): self.centroid, self.segment_history = tuple(cen), hist
    prev = [S(p[i], [S(first[i], [])] if has[i] else []) for i in range(n)]
    curr = [S(c[j], []) for j in range(n)]
    t0 = time.perf_counter(); st.match_costs(prev, curr); dn = time.perf_counter() - t0
    t0 = time.perf_counter(); a = st.apply_hungarian_algorithm(m); dh = time.perf_counter() - t0
    print("tracker cost matrix %dx%d: CUDA %.3f ms (matrix in pinned memory), numpy match block %.2f ms, Hungarian %.2f ms" % (n, n, dt*1e3, dn*1e3, dh*1e3))
    ws.close()
