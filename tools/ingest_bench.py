"""Throughput of the GPU MatrixMarket ingest against scipy.io.mmread (CPU), c1-sized text."""
import io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.io, scipy.sparse as sp
import too_many_cells_b200 as T
from too_many_cells_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
ip, ix, dv, _ = synth.synth_config(name)
n, g = synth.CONFIGS[name][:2]
a = sp.csr_matrix((dv, ix, ip), shape=(n, g))
coo = sp.coo_matrix(a.T)
t0 = time.perf_counter()
lines = np.char.add(np.char.add(np.char.add((coo.row + 1).astype(str), " "), np.char.add((coo.col + 1).astype(str), " ")), coo.data.astype(np.int64).astype(str))
text = ("%%MatrixMarket matrix coordinate integer general\n" + f"{g} {n} {coo.nnz}\n" + "\n".join(lines.tolist()) + "\n").encode()
print(f"text: {len(text)/1e6:.1f} MB, {coo.nnz} entries (built in {time.perf_counter()-t0:.1f} s)", flush=True)
T.parse_matrix_market(text, transpose=True)   # warm
t0 = time.perf_counter(); m = T.parse_matrix_market(text, transpose=True); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); r = scipy.io.mmread(io.BytesIO(text)).T.tocsr(); t_cpu = time.perf_counter() - t0
r.sort_indices()
ok = np.array_equal(m.indptr, r.indptr) and np.array_equal(m.indices, r.indices) and np.array_equal(m.data, r.data)
print(f"GPU ingest (H2D + parse + sort + D2H): {t_gpu*1e3:.0f} ms = {len(text)/t_gpu/1e6:.0f} MB/s; scipy.io.mmread + transpose: {t_cpu*1e3:.0f} ms = {len(text)/t_cpu/1e6:.0f} MB/s; identical: {ok}")
