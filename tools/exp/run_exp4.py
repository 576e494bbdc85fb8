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
rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32), (rp[1:] - rp[:-1]))
v16 = val.to(torch.int32).clamp(max=65535).to(torch.int16).contiguous()
x = torch.rand(n, dtype=torch.float64, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
yref = None
for pbs in (4096, 8192, 16384):
    pb = (rows // pbs).to(torch.int64)
    key = (pb * g + col.to(torch.int64)) * n + rows.to(torch.int64)
    order = torch.sort(key).indices
    tcol = col[order].contiguous(); tpos = rows[order].contiguous(); tv = v16[order].contiguous()
    nb = (n + pbs - 1) // pbs
    cnt = torch.bincount(pb, minlength=nb)
    blk = torch.zeros(nb + 1, dtype=torch.int64, device=dev); blk[1:] = torch.cumsum(cnt, 0)
    del order, key
    for shares in (64, 256, 1024):
        y = torch.zeros(g, dtype=torch.float64, device=dev)
        lib.exp_tposblk(pbs, P(tcol), P(tpos), P(tv), P(blk), nb, P(x), P(y), n, shares); torch.cuda.synchronize()
        if yref is None: yref = y.clone()
        err = float((y - yref).abs().max() / yref.abs().max())
        ms = timeit(lambda: lib.exp_tposblk(pbs, P(tcol), P(tpos), P(tv), P(blk), nb, P(x), P(y), n, shares))
        print(f"posblk={pbs} shares={shares} ctas={nb*shares}: {ms:.4f} ms  {10.0*nnz/ms/1e6:.0f} GB/s (10 B/nnz)  relerr {err:.1e}", flush=True)
    del tcol, tpos, tv
