"""Development probe: AE size / accepted-vector statistics per level of a hierarchy.
usage: level_stats.py n levels [epa]   (writes gpurun_out/level_stats_<n>.json)"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import saamge_b200 as sab

n = int(sys.argv[1]); levels = int(sys.argv[2]); epa = int(sys.argv[3]) if len(sys.argv) > 3 else 64
p = sab.default_params(num_levels=levels, first_elems_per_agg=52, elems_per_agg=epa, partition_kind=2, block=(32, 32, 32))
pr = sab.Problem(3, n, coef_kind=1); na = pr.partition(p)
t = time.time(); H = sab.ml_build(pr, p); print("ml_build %.2fs" % (time.time() - t), flush=True)
out = {"times": H.times()}
for l in range(levels - 1):
    I = H.get("AE_to_dof.I", l); sz = np.diff(I)
    h_m = None
    try:
        sab.ml_download(H) if l == 0 else None
        h_m = H.get("ae_m", l)
    except Exception as ex:
        print("no ae_m", ex)
    q = [0, 5, 25, 50, 75, 95, 100]
    d = {"ND": H.scalar("ND", l), "nparts": int(len(sz)), "n_pct": dict(zip(q, np.percentile(sz, q).tolist())),
         "sum_n3": float((sz.astype(float) ** 3).sum()), "sum_n2": float((sz.astype(float) ** 2).sum()),
         "sizes_sorted_desc_head": np.sort(sz)[::-1][:40].tolist()}
    if h_m is not None and len(h_m) == len(sz):
        d["m_pct"] = dict(zip(q, np.percentile(h_m, q).tolist())); d["sum_m"] = int(h_m.sum())
        d["sum_n2m"] = float((sz.astype(float) ** 2 * h_m).sum())
    out["level%d" % l] = d
    print(l, json.dumps(d), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "level_stats_%d.json" % n), "w"), indent=1)
