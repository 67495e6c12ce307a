"""Unit-weight D-infinity accumulation engines side by side on one raster (filled, and filled + flat-resolved):
    python tools/dinf_profile.py 32768 "accum_dinf_packed=0" "accum_dinf_packed=1 accum_dinf_stats=1" ...
Each quoted argument is one configuration of rdb200_set_param switches; results are compared with the first one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from richdem_b200 import _lib  # noqa: E402

L = _lib.lib()
_lib.init(0)
_lib.use_torch_stream()
N = int(sys.argv[1])
D8 = len(sys.argv) > 2 and sys.argv[2] == "d8"  # time FA_D8 instead (python tools/dinf_profile.py 32768 d8 "accum_walk_scan=0" ...)
configs = [c.strip() for a in sys.argv[(3 if D8 else 2):] for c in a.split(";")] or [""]  # (";" separates configurations too)
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(d.data_ptr(), N, N))
r = d.clone()
_lib.check(L.rdb200_dev_resolve_flats_epsilon_f32(r.data_ptr(), N, N, -9999.0))
acc = torch.empty((N, N), dtype=torch.float64, device="cuda")
for name, dem in (("filled", d), ("resolved", r)):
    ref = None
    for cfg in configs:
        _lib.reset_params()
        for kv in cfg.split():
            _lib.set_param(kv.split("=")[0], int(kv.split("=")[1]))
        times = []
        for rep in range(2):
            fa = L.rdb200_dev_fa_d8_f32_f64 if D8 else L.rdb200_dev_fa_tarboton_f32_f64
            _lib.check(fa(dem.data_ptr(), acc.data_ptr(), N, N, -9999.0, 1))
            times.append(_lib.stats()["ms_total"])
        if ref is None:
            ref = acc.clone()
            err = 0.0
        else:
            err = float(((acc - ref).abs() / ref.abs().clamp(min=1.0)).max())
        print(f"N={N} {name:8s} [{cfg or 'defaults'}] {'fa_d8' if D8 else 'fa_dinf'} {min(times):8.1f} ms (runs {', '.join(f'{t:.1f}' for t in times)}) "
              f"rounds={_lib.stats()['accum_rounds']} max rel diff vs first {err:.2e}", flush=True)
_lib.reset_params()
