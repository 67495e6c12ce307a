import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from richdem_b200 import _lib
import richdem_b200 as rd
L = _lib.lib(); _lib.init(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(d.data_ptr(), N, N))
dem = rd.rdarray(d.cpu().numpy(), no_data=-9999.0)
del d
res = []
for rep in range(3):
    m, l = rd.FlatMask(dem)
    res.append((m, l))
    print(rep, "mask sum", int(m.astype(np.int64).sum()), "labelled", int((l != 0).sum()), flush=True)
for k in (1, 2):
    dm = res[0][0] != res[k][0]; dl = res[0][1] != res[k][1]
    print("rep", k, "mask diff", int(dm.sum()), "label diff", int(dl.sum()))
    if dm.any():
        ys, xs = np.nonzero(dm)
        print("  first mask diffs:", [(int(x), int(y), int(res[0][0][y, x]), int(res[k][0][y, x]), int(res[0][1][y, x]), int(res[k][1][y, x])) for y, x in list(zip(ys, xs))[:6]])
        # bounding box of the first differing flat
        lab = res[0][1][ys[0], xs[0]]
        yy, xx = np.nonzero(res[0][1] == lab)
        print("  flat label", int(lab), "size", len(yy), "bbox x", int(xx.min()), int(xx.max()), "y", int(yy.min()), int(yy.max()))
