import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from richdem_b200 import _lib
L = _lib.lib(); _lib.init(0); _lib.use_torch_stream()
_lib.set_param("fill_ordered", 0)
_lib.set_param("fill_rounds_per_sync", 1)
for N in (64, 128, 256, 512, 1024):
    d = torch.empty((N, N), dtype=torch.float32, device="cuda")
    _lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
    for rep in range(3):
        w = d.clone()
        _lib.check(L.rdb200_dev_fill_depressions_d8_f32(w.data_ptr(), N, N))
        st = _lib.stats()
    print(f"N={N} sweep_ms={st['ms_main_kernel']:.3f} rounds={st['fill_rounds']} visits={st['fill_tile_visits']} passes={st['fill_tile_iters']} -> us/round={1e3*st['ms_main_kernel']/max(st['fill_rounds'],1):.1f} passes/visit={st['fill_tile_iters']/max(st['fill_tile_visits'],1):.1f}", flush=True)
