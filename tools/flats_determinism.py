import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from richdem_b200 import _lib
L = _lib.lib(); _lib.init(0); _lib.use_torch_stream()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(d.data_ptr(), N, N))
outs = []
for rep in range(3):
    w = d.clone()
    _lib.check(L.rdb200_dev_resolve_flats_epsilon_f32(w.data_ptr(), N, N, -9999.0))
    st = _lib.stats()
    print(rep, st["flat_cells_raised"], st["flat_bfs_levels"], flush=True)
    outs.append(w)
for k in (1, 2):
    diff = (outs[0].view(torch.int32) != outs[k].view(torch.int32))
    nd = int(diff.sum())
    print("rep", k, "differs from rep 0 in", nd, "cells")
    if nd:
        idx = diff.nonzero()[:10]
        for y, x in idx.tolist():
            a = outs[0][y, x].item(); b = outs[k][y, x].item(); z = d[y, x].item()
            ua = outs[0].view(torch.int32)[y, x].item() - d.view(torch.int32)[y, x].item()
            ub = outs[k].view(torch.int32)[y, x].item() - d.view(torch.int32)[y, x].item()
            print(f"   ({x},{y}) filled={z!r} ulps0={ua} ulps{k}={ub}")
