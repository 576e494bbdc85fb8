import ctypes as C, os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
lib = C.CDLL(os.path.join(HERE, "libexp.so"))
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)
nnz = val.numel()
x = torch.rand(n, dtype=torch.float64, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32), (rp[1:] - rp[:-1]))
def make(order_rows):
    key = col.to(torch.int64) * (2 * n) + order_rows.to(torch.int64)
    order = torch.sort(key).indices
    return col[order].contiguous(), rows[order].contiguous(), val[order].contiguous()
variants = {"sorted": rows}
W = 9472
win = (rows // W) * W + torch.randint(0, W, (nnz,), device=dev, dtype=torch.int32) % W     # random order inside windows of W rows
variants["window-shuffled(9472)"] = win
variants["fully shuffled"] = torch.randint(0, n, (nnz,), device=dev, dtype=torch.int32)
for nm, o in variants.items():
    ccol, cpos, cval = make(o)
    pad = lambda t: torch.cat([t, torch.zeros(16, dtype=t.dtype, device=dev)])
    ccol, cpos, cval = pad(ccol), pad(cpos), pad(cval)
    for mode, mn in ((0, "merge+atomics"), (1, "no atomics"), (2, "no x gather")):
        yt = torch.zeros(g, dtype=torch.float64, device=dev)
        ms = timeit(lambda: lib.exp_tcoo_v8(mode, P(ccol), P(cpos), P(cval), P(x), P(yt), C.c_longlong(nnz)))
        print(f"{nm:24s} {mn:14s}: {ms:.4f} ms  {16.0*nnz/ms/1e6:.0f} GB/s", flush=True)
    del ccol, cpos, cval
