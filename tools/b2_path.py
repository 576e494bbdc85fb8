"""The reference's own call pattern: the tf-idf matrix B2 is handed to the engine (normFlag False).  Times the tree from B2 with
and without factor recovery and checks it against the tree from raw counts + in-engine tf-idf.  python tools/b2_path.py c2"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
df = torch.bincount(col.long(), minlength=g).double()
idf = torch.where(df > 0, torch.log(float(n) / df.clamp(min=1)), torch.zeros_like(df))
b2 = val * idf[col.long()]
ref = T.hierarchical_spectral_cluster(T.DeviceB.adopt_device(n, g, rp, col, val.clone(), normalize=True, device=0))
for mode in ("recover", "f64"):
    if mode == "f64": os.environ["TMC_NO_RECOVER"] = "1"
    for it in range(3):
        w = b2.clone(); torch.cuda.synchronize(); t0 = time.perf_counter()
        b = T.DeviceB.adopt_device(n, g, rp, col, w, normalize=False, device=0); t1 = time.perf_counter()
        tr = T.hierarchical_spectral_cluster(b); t2 = time.perf_counter()
        info = b.info(); b.free()
    same = tr.leaf_sets() == ref.leaf_sets()
    print(f"{name} {mode}: storage {info['value_type']} panel {info['panel_cols']} cols x {info['panel_bits']} bits; adopt {1e3*(t1-t0):.1f} ms tree {1e3*(t2-t1):.1f} ms; "
          f"nodes {tr.n_nodes} leaf sets equal to the counts+tf-idf tree: {same}", flush=True)
