import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1]
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
h_ptr = rp.cpu().numpy(); nnz = val.numel()
np_col = col.cpu().numpy().astype(np.uint16); np_val = val.cpu().numpy().astype(np.uint16)
rt = torch.cuda.cudart()
for pin in (0, 1):
    if pin:
        print("register ->", rt.cudaHostRegister(np_col.ctypes.data, np_col.nbytes, 0), rt.cudaHostRegister(np_val.ctypes.data, np_val.nbytes, 0), flush=True)
    host = (h_ptr, np_col, np_val, (n, g))
    for it in range(4):
        t = time.perf_counter()
        tr = T.hierarchical_spectral_cluster(host, normalize=True, device=0)
        print(f"pinned={pin} it {it}: {1e3*(time.perf_counter()-t):.1f} ms engine {tr.stats['engine_ms']:.1f}", flush=True)
# torch-pinned tensors for comparison
pc = torch.from_numpy(np_col.astype(np.int32)).pin_memory(); pv = torch.from_numpy(np_val.astype(np.float64)).pin_memory()
host = (h_ptr, pc.numpy(), pv.numpy(), (n, g))
for it in range(3):
    t = time.perf_counter(); tr = T.hierarchical_spectral_cluster(host, normalize=True, device=0)
    print(f"torch-pinned int32/f64 it {it}: {1e3*(time.perf_counter()-t):.1f} ms", flush=True)
