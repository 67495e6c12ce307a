import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from richdem_b200 import _lib
L = _lib.lib(); _lib.init(0); _lib.use_torch_stream()
# k=v arguments are rdb200_set_param switches (e.g. flats_uf_tiled=1); RDB200_PROFILE=1 prints the phase laps
for kv in [a for a in sys.argv[1:] if "=" in a]:
    _lib.set_param(kv.split("=")[0], int(kv.split("=")[1]))
for N in [int(a) for a in sys.argv[1:] if "=" not in a] or [4096, 16384]:
    d = torch.empty((N, N), dtype=torch.float32, device="cuda")
    _lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
    _lib.check(L.rdb200_dev_fill_depressions_d8_f32(d.data_ptr(), N, N)); s0 = _lib.stats()
    for rep in range(2):
        w = d.clone()
        t = time.time()
        _lib.check(L.rdb200_dev_resolve_flats_epsilon_f32(w.data_ptr(), N, N, -9999.0))
        dt = time.time() - t
        st = _lib.stats()
    acc = torch.empty((N, N), dtype=torch.float64, device="cuda")
    _lib.check(L.rdb200_dev_fa_d8_f32_f64(w.data_ptr(), acc.data_ptr(), N, N, -9999.0, 1)); s2 = _lib.stats()
    _lib.check(L.rdb200_dev_fa_tarboton_f32_f64(w.data_ptr(), acc.data_ptr(), N, N, -9999.0, 1)); s3 = _lib.stats()
    print(f"N={N} fill={s0['ms_total']:.1f}ms flats={st['ms_total']:.1f}ms (levels={st['flat_bfs_levels']} raised={st['flat_cells_raised']} launches={st['kernel_launches']}) fa_d8={s2['ms_total']:.1f}ms fa_dinf={s3['ms_total']:.1f}ms (rounds={s3['accum_rounds']}) acc_max={float(acc.max()):.0f}", flush=True)
    del d, w, acc
