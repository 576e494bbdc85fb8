import ctypes as C, os, sys, subprocess
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
import torch
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
lib = C.CDLL(os.path.join(HERE, "libexp.so"))
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n, g, k, d, s, binary = synth.CONFIGS[name]
dev = torch.device("cuda", 0)
rp, col, val = synth.synth_counts_torch(n, g, k, d, s, dev, binary)
b = T.DeviceB.adopt_device(n, g, rp, col, val, normalize=True, device=0)   # val normalised in place
nnz = val.numel(); bytes_half = 12.0 * nnz
perm = torch.arange(n, dtype=torch.int32, device=dev)
permr = torch.randperm(n, device=dev).to(torch.int32)
x = torch.rand(n, dtype=torch.float64, device=dev); y = torch.rand(g, dtype=torch.float64, device=dev)
w = torch.zeros(n, dtype=torch.float64, device=dev)
P = lambda t: C.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
ref = None
print("product bench_spmv_pair (scatter, gather ms):", T.bench_spmv_pair(b, 10), flush=True)
yz = torch.zeros(g, dtype=torch.float64, device=dev)
def alt():
    yz.zero_()
    lib.exp_scatter(0, P(rp), P(col), P(val), P(perm), P(x), P(yz), n, 48, g, 1, C.c_int64(g))
    lib.exp_gather(0, P(rp), P(col), P(val), P(perm), P(yz), P(w), n, 48)
print("alternating zero+scatter+gather v0:", timeit(alt), "ms", flush=True)
for pm, pname in ((permr, "randperm"),):
    for rpc in (8, 16, 48):
        for v in (0, 3, 4, 6):
            ms = timeit(lambda: lib.exp_gather(v, P(rp), P(col), P(val), P(pm), P(y), P(w), n, rpc))
            chk = float(w.sum()) if v != 2 else 0.0
            print(f"gather v{v} perm={pname} rpc={rpc}: {ms:.4f} ms  {bytes_half/ms/1e6:.0f} GB/s  chk {chk:.10e}", flush=True)
for rep in (1, 4, 16):
    yy = torch.zeros(rep * g, dtype=torch.float64, device=dev)
    for v in (0, 1):
        ms = timeit(lambda: lib.exp_scatter(v, P(rp), P(col), P(val), P(perm), P(x), P(yy), n, 48, g, rep, C.c_int64(g)))
        print(f"scatter v{v} replicas={rep}: {ms:.4f} ms  {bytes_half/ms/1e6:.0f} GB/s", flush=True)
# column-sorted COO
rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int32), (rp[1:] - rp[:-1]))
order = torch.sort(col.to(torch.int64) * n + rows.to(torch.int64)).indices
ccol = col[order].contiguous(); cpos = rows[order].contiguous(); cval = val[order].contiguous()
del order
yref = torch.zeros(g, dtype=torch.float64, device=dev)
lib.exp_scatter(0, P(rp), P(col), P(val), P(perm), P(x), P(yref), n, 48, g, 1, C.c_int64(g)); torch.cuda.synchronize()
for E in (4, 8, 16):
    yt = torch.zeros(g, dtype=torch.float64, device=dev)
    lib.exp_tcoo(E, P(ccol), P(cpos), P(cval), P(x), P(yt), C.c_int64(nnz)); torch.cuda.synchronize()
    err = float((yt - yref).abs().max() / yref.abs().max())
    ms = timeit(lambda: lib.exp_tcoo(E, P(ccol), P(cpos), P(cval), P(x), P(yt), C.c_int64(nnz)))
    print(f"tcoo E={E}: {ms:.4f} ms  {16.0*nnz/ms/1e6:.0f} GB/s (16 B/nnz)  relerr {err:.2e}", flush=True)
