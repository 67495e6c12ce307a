"""A/B timing of the accumulation switches that round 1 left off by default:

    python tools/accum_switches.py [N] [--dinf]

On an N x N fBm DEM (filled on the device; with --dinf also flat-resolved and run through FA_Tarboton), every
configuration is run 3 times; prints the best ms_total, the main-kernel time, the level count, and whether the
result equals the baseline configuration's (bit-exact for unit-weight D8, 1e-9 relative for D-infinity).
Wrap in `timeout`: the switches have only been checked on the CPU emulation of the kernels (tests/emu).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from richdem_b200 import _lib  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 16384
DINF = "--dinf" in sys.argv
ND = -9999.0
L = _lib.lib()
_lib.init(0)
_lib.use_torch_stream()

dem = torch.empty((N, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(dem.data_ptr(), N, N, 0, 42, 12, 0.0))
_lib.check(L.rdb200_dev_fill_depressions_d8_f32(dem.data_ptr(), N, N))
if DINF:
    _lib.check(L.rdb200_dev_resolve_flats_epsilon_f32(dem.data_ptr(), N, N, ND))
acc = torch.empty((N, N), dtype=torch.float64, device="cuda")

ALL = ("accum_fused_prep", "accum_walk_lanes", "accum_agg", "accum_tail", "accum_tail_budget", "accum_async")
if DINF:
    configs = [{}, {"accum_agg": 1}, {"accum_tail": 1024}, {"accum_tail": 4096}, {"accum_tail": 4096, "accum_tail_budget": 128},
               {"accum_agg": 1, "accum_tail": 4096}, {"accum_agg": 1, "accum_tail": 16384, "accum_tail_budget": 64},
               {"accum_async": 1}]
else:
    configs = [{}, {"accum_fused_prep": 1}, {"accum_walk_lanes": 1}, {"accum_fused_prep": 1, "accum_walk_lanes": 1}]


def run(cfg):
    for k in ALL:
        _lib.set_param(k, 0)
    for k, v in cfg.items():
        _lib.set_param(k, v)
    best = None
    for _ in range(3):
        fn = L.rdb200_dev_fa_tarboton_f32_f64 if DINF else L.rdb200_dev_fa_d8_f32_f64
        _lib.check(fn(dem.data_ptr(), acc.data_ptr(), N, N, ND, 1))
        st = _lib.stats()
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    return best


base = None
for cfg in configs:
    st = run(cfg)
    if base is None:
        base = acc.clone()
        same = True
    elif DINF:
        same = bool(((acc - base).abs() <= 1e-9 * base.abs().clamp(min=1.0)).all())
    else:
        same = bool(torch.equal(acc, base))
    print(f"N={N} {'FA_Dinf' if DINF else 'FA_D8'} {cfg or 'baseline'}: ms_total={st['ms_total']:.2f} main_kernel={st['ms_main_kernel']:.2f} "
          f"levels={st['accum_rounds']} same={same}", flush=True)
for k in ALL:
    _lib.set_param(k, 0)
if not DINF:  # d8_flow_directions: default kernel vs the rolling-window one
    dirs = torch.empty((N, N), dtype=torch.uint8, device="cuda")
    base_dirs = None
    for roll in (0, 1):
        _lib.set_param("flowdirs_rolling", roll)
        best = 1e30
        for _ in range(3):
            _lib.check(L.rdb200_dev_d8_flow_directions_f32(dem.data_ptr(), dirs.data_ptr(), N, N, ND))
            best = min(best, _lib.stats()["ms_total"])
        if base_dirs is None:
            base_dirs = dirs.clone()
        print(f"N={N} d8_flow_directions flowdirs_rolling={roll}: ms_total={best:.2f} same={bool(torch.equal(dirs, base_dirs))}", flush=True)
    _lib.set_param("flowdirs_rolling", 0)
