import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from richdem_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = _lib.lib(); _lib.init(0); _lib.use_torch_stream()
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
ref = d.clone()
_lib.set_param("fill_ordered", 0)
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(ref.data_ptr(), N, N))
ref.copy_(d)
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(ref.data_ptr(), N, N))
st = _lib.stats()
print(f"N={N} baseline ms_total={st['ms_total']:.2f} sweep_ms={st['ms_main_kernel']:.2f} rounds={st['fill_rounds']} visits={st['fill_tile_visits']} iters={st['fill_tile_iters']}", flush=True)
# e.g.  python tools/fill_profile.py 32768 "" fill_async=1 fill_async=1,fill_ordered=0   (wrap in `timeout`)
configs = sys.argv[2:] or [""]
for cfg in configs:
    kv = [a.split("=") for a in cfg.split(",") if a]
    for k, v in kv:
        _lib.set_param(k, int(v))
    for prof in ((1, 0) if "--prof" in os.environ.get("FP_FLAGS", "") else (0,)):
        _lib.set_param("fill_profile", prof)
        w = d.clone()
        _lib.check(L.rdb200_dev_fill_depressions_d8_f32(w.data_ptr(), N, N))
        st = _lib.stats()
        ok = bool(torch.equal(w, ref))
        print(f"N={N} [{cfg}] profile={prof} same={ok} ms_total={st['ms_total']:.2f} sweep_ms={st['ms_main_kernel']:.2f} rounds={st['fill_rounds']} visits={st['fill_tile_visits']} iters={st['fill_tile_iters']}", flush=True)
