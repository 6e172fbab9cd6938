import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oraclelib as ol
from cmc_cosmic_b200.cluster import Cluster
d = np.load(os.path.join(ROOT, "tmp_debug", "sigma_bad_state.npz"))
sr, sm, vr, vt, N = d["r"], d["m"], d["vr"], d["vt"], int(d["N"])
for n in (len(vr), len(vr) - 1, len(vr) - 2048, 500000, 98305, 98304 + 2048, 98304, 49153 + 2048 * 47):
    c = Cluster(N, n_star_dim=N + 64)
    c.load(np.arange(1, N + 1) * 1.0, np.zeros(N), np.zeros(N), np.ones(N))
    c.star_r[:n + 2] = sr[:n + 2]; c.star_m[:n + 2] = sm[:n + 2]
    c.star_r[n + 1] = 1e10
    c.field("vr")[1:n + 1] = vr[:n]; c.field("vt")[1:n + 1] = vt[:n]
    c.clus.N_MAX = c.clus.N_MAX_NEW = n; c.Start[0], c.End[0] = 1, n
    c.upload()
    out = []
    for rep in range(3):
        a, sg = c.dev_calc_sigma_r(20)
        out.append(dict(c.sigma_stats))
    want = ol.calc_sigma_r(20, c.star_r[:n + 2].copy(), c.star_m[:n + 2].copy(), n, N, vr[:n], vt[:n])[1]
    print("n", n, "tiles", out[0]["tiles"], "rerun", [o["rerun"] for o in out], "hard_tiles", [o["hard_tiles"] for o in out], "ok", np.array_equal(sg[1:].view(np.uint64), want[1:].view(np.uint64)), flush=True)
    c.close()
