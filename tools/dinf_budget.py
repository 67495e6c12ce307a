import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from richdem_b200 import _lib
L = _lib.lib(); _lib.init(0); _lib.use_torch_stream()
N = int(sys.argv[1])
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(d.data_ptr(), N, N))
filled = d.clone()
_lib.check(L.rdb200_dev_resolve_flats_epsilon_f32(d.data_ptr(), N, N, -9999.0))
acc = torch.empty((N, N), dtype=torch.float64, device="cuda")
ref = None
for name, dem in (("resolved", d), ("filled-only", filled)):
    for b in [int(x) for x in sys.argv[2:]]:
        _lib.set_param("accum_budget", b)
        _lib.check(L.rdb200_dev_fa_tarboton_f32_f64(dem.data_ptr(), acc.data_ptr(), N, N, -9999.0, 1))
        st = _lib.stats()
        print(f"N={N} {name} budget={b} fa_dinf={st['ms_total']:.1f}ms levels={st['accum_rounds']} max={float(acc.max()):.3f}", flush=True)
