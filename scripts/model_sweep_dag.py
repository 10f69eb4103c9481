"""Design aid (CPU only): list-scheduling model of the GCFM sweep's dependency DAG.

The sweep visits agents in a random permutation; agent i needs the NEW state of every earlier agent within reach.  W warps draw
tickets in permutation order; an agent takes t_pre before it needs its dependencies and t_post after the last one has published.
Prints mean dependencies per agent, DAG depth and the modelled sweep time for the candidate-search variants of round 2
(reach = cutoff + displacement margin; field-of-view culling with the 0.7 pi half-angle of pedestrians.py:259-262).

    python scripts/model_sweep_dag.py          # 12.5 k agents on 819 x 102 m, 100 k on 819 x 819 m, 1000 on 20 x 20 m
"""
import heapq

import numpy as np
from scipy.spatial import cKDTree


def run(N, Lx, Ly, reach, W, t_pre, t_post, fov=False, seed=0, t_other=0.0):
    rng = np.random.RandomState(seed)
    x, y, th = rng.uniform(0, Lx, N), rng.uniform(0, Ly, N), rng.uniform(0, 2 * np.pi, N)
    perm = rng.permutation(N)
    rank = np.empty(N, int)
    rank[perm] = np.arange(N)
    pairs = cKDTree(np.c_[x, y]).query_pairs(reach, output_type='ndarray')
    a, b = pairs[:, 0], pairs[:, 1]
    sw = rank[a] > rank[b]
    late, early = np.where(sw, a, b), np.where(sw, b, a)          # `late` depends on `early`
    if fov:
        ex, ey = x[early] - x[late], y[early] - y[late]
        c = (np.cos(th[late]) * ex + np.sin(th[late]) * ey) / np.hypot(ex, ey)
        keep = c > np.cos(0.7 * np.pi) - 0.05
        late, early = late[keep], early[keep]
    deps = [[] for _ in range(N)]
    for l, e in zip(late, early):
        deps[l].append(e)
    depth = np.zeros(N, int)
    for r in range(N):
        i = perm[r]
        if deps[i]:
            depth[i] = 1 + max(depth[j] for j in deps[i])
    fin = np.zeros(N)
    workers = [0.0] * W
    heapq.heapify(workers)
    for r in range(N):
        i = perm[r]
        ready = heapq.heappop(workers) + t_pre
        if deps[i]:
            f = max(ready, max(fin[j] for j in deps[i])) + t_post
        else:
            f = ready + t_other
        fin[i] = f
        heapq.heappush(workers, f)
    return np.mean([len(d) for d in deps]), depth.max(), fin.max()


if __name__ == "__main__":
    for (N, Lx, Ly) in [(12500, 819.2, 102.4), (100000, 819.2, 819.2), (1000, 20, 20)]:
        for reach in (5.0, 4.25):
            for fov in (False, True):
                for (tp, tq) in ((12, 10), (3, 8), (3, 5)):
                    nd, d, T = run(N, Lx, Ly, reach, 2368, tp, tq, fov, t_other=min(tq, 2))
                    print(f"N={N} reach={reach} fov={fov} t_pre={tp}us t_post={tq}us: {nd:.2f} deps/agent, depth {d}, sweep {T:.0f} us")
