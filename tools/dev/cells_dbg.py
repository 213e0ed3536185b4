import sys, time, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from oarfish_b200 import synth, DeviceStore
n_cells, reads, M = 296, 50000, 200000
s, crp = synth.make_cells([reads] * n_cells, M, 6.0, seed=5)
# compact (cell, txp) ids with numpy
lens = np.diff(s.row_ptr.astype(np.int64))
cell_of_row = np.repeat(np.arange(n_cells), np.diff(crp.astype(np.int64)))
cell_of_aln = np.repeat(cell_of_row, lens)
key = cell_of_aln.astype(np.int64) * M + s.txp_id.astype(np.int64)
uniq, inv = np.unique(key, return_inverse=True)
gtxp = inv.astype(np.uint32); T = len(uniq)
print("T", T, "nnz", s.nnz)
ds = DeviceStore(s.row_ptr, gtxp, s.prob, T)
print(ds.layout_info())
prev = torch.full((T,), 0.25, dtype=torch.float64, device="cuda"); curr = torch.zeros(T, dtype=torch.float64, device="cuda")
ds.sweep_timed(prev, curr, 3)
print("sweep us", ds.sweep_timed(prev, curr, 20) / 20 * 1e3)
prev.zero_()
print("sweep us (prev=0)", ds.sweep_timed(prev, curr, 20) / 20 * 1e3)
