import ctypes as C, os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
import torch
from too_many_cells_b200 import synth
lib = C.CDLL(os.path.join(HERE, "libexp.so"))
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
nnz = val.numel()
df = torch.bincount(col.long(), minlength=g)
rank = torch.empty(g, dtype=torch.int32, device=dev)
rank[torch.argsort(df, descending=True)] = torch.arange(g, dtype=torch.int32, device=dev)
colr = rank[col.long()].contiguous()                      # hot columns get small ids
v16 = val.to(torch.int32).clamp(max=65535).to(torch.int16).contiguous()    # bit pattern of uint16
perm = torch.arange(n, dtype=torch.int32, device=dev)
y = torch.rand(g, dtype=torch.float64, device=dev); w = torch.zeros(n, dtype=torch.float64, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for rpc in (8, 16):
    ms = timeit(lambda: lib.exp_gather_u16(P(rp), P(colr), P(v16), P(perm), P(y), P(w), n, rpc))
    print(f"u16 baseline rpc={rpc}: {ms:.4f} ms  {6.0*nnz/ms/1e6:.0f} GB/s  chk {float(w.sum()):.8e}", flush=True)
for nt in (256, 512, 1024):
    for H in (4096, 8192, 12288, 16384, 24576):
        if H * 8 > 200 * 1024: continue
        for rpc in (nt // 32 * 4, nt // 32 * 16):
            ms = timeit(lambda: lib.exp_gather_smem(nt, P(rp), P(colr), P(v16), P(perm), P(y), P(w), n, rpc, H))
            print(f"smem nt={nt} H={H} rpc={rpc}: {ms:.4f} ms  {6.0*nnz/ms/1e6:.0f} GB/s  chk {float(w.sum()):.8e}", flush=True)
