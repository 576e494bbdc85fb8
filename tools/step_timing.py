import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
work = torch.empty_like(val)
for it in range(int(os.environ.get("TMC_ITERS", "5"))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    work.copy_(val); torch.cuda.synchronize(); t1 = time.perf_counter()
    b = T.DeviceB.adopt_device(n, g, rp, col, work, normalize=True, device=0); t2 = time.perf_counter()
    tr = T.hierarchical_spectral_cluster(b, profile=(it % 2 == 1)); t3 = time.perf_counter()
    b.free(); t4 = time.perf_counter()
    print(f"it {it}: copy {1e3*(t1-t0):.1f} adopt {1e3*(t2-t1):.1f} tree {1e3*(t3-t2):.1f} (engine {tr.stats['engine_ms']:.1f}, spmv {tr.stats['spmv_ms']:.1f}) free {1e3*(t4-t3):.1f} ms", flush=True)
b = T.DeviceB.adopt_device(n, g, rp, col, work.copy_(val), normalize=True, device=0)
print("pair (sorted root):", T.bench_spmv_pair(b, 10))
