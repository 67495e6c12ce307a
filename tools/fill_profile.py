"""Fill timing / work counters for a list of rdb200_set_param configurations:

    python tools/fill_profile.py [N] [cfg ...]        cfg = "k=v,k=v" ("" = defaults)
    e.g.  timeout 300 python tools/fill_profile.py 32768 "" fill_ordered=0 fill_multigrid=0 fill_vcycle=0

Every configuration starts from the defaults, runs twice on the same N x N fBm DEM (seed 42) and reports the
faster run; `same` compares the result with the first configuration's.  FP_FLAGS=--prof adds the in-tile counters.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from richdem_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = _lib.lib()
_lib.init(0)
_lib.use_torch_stream()
d = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
ref = None
for cfg in sys.argv[2:] or [""]:
    _lib.reset_params()
    for k, v in [a.split("=") for a in cfg.split(",") if a]:
        _lib.set_param(k, int(v))
    if "--prof" in os.environ.get("FP_FLAGS", ""):
        _lib.set_param("fill_profile", 1)
    best = None
    for _ in range(2):
        w = d.clone()
        _lib.check(L.rdb200_dev_fill_depressions_d8_f32(w.data_ptr(), N, N))
        st = _lib.stats()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    if ref is None:
        ref = w
    nt = ((N + 63) // 64) ** 2
    print(f"N={N} [{cfg or 'defaults'}] same={bool(torch.equal(w, ref))} ms_total={best['ms_total']:.2f} "
          f"sweep_ms={best['ms_main_kernel']:.2f} rounds={best['fill_rounds']} visits={best['fill_tile_visits']} "
          f"({best['fill_tile_visits'] / nt:.2f} raster-eq) passes/visit={best['fill_tile_iters'] / max(1, best['fill_tile_visits']):.2f}",
          flush=True)
_lib.reset_params()
