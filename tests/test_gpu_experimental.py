"""GPU parity of the switches that round 1 prepared but could not time or run on a B200:
fill_async, accum_fused_prep, accum_walk_lanes, accum_agg, accum_tail, flats_uf_tiled.  They are off by default and
have only been checked on the CPU model of the kernels (tests/test_emulated_kernels.py), so these tests are opt-in:

    RDB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q

Once a switch has passed here and has been timed (tools/accum_switches.py, tools/fill_profile.py,
tools/flats_profile.py), move its case into tests/test_gpu_parity.py::test_algorithm_variants_agree.
"""
import os

import numpy as np
import pytest

import oracle
import richdem_b200 as rd
from richdem_b200 import _lib

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("RDB_TEST_EXPERIMENTAL"), reason="opt-in: RDB_TEST_EXPERIMENTAL=1")]
ND = -9999.0
SWITCHES = ("fill_multigrid", "fill_multigrid_min", "fill_vcycle", "fill_async", "flowdirs_rolling", "accum_async", "accum_fused_prep", "accum_walk_lanes", "accum_agg", "accum_tail", "accum_tail_budget",
            "flats_uf_tiled")


def R(a, nd=ND):
    return rd.rdarray(np.ascontiguousarray(a), no_data=nd)


@pytest.fixture()
def switches():
    yield
    for k in SWITCHES:
        _lib.set_param(k, 0)


CONFIGS = [{"accum_fused_prep": 1}, {"accum_walk_lanes": 1}, {"accum_fused_prep": 1, "accum_walk_lanes": 1},
           {"accum_agg": 1}, {"accum_tail": 2048}, {"accum_agg": 1, "accum_tail": 64, "accum_tail_budget": 3},
           {"flats_uf_tiled": 1}, {"fill_async": 1}, {"accum_async": 1}, {"flowdirs_rolling": 1},
           {"fill_multigrid": 8, "fill_multigrid_min": 256}, {"fill_multigrid": 4, "fill_multigrid_min": 256, "fill_async": 1},
           {"fill_multigrid": 8, "fill_multigrid_min": 256, "fill_vcycle": 4}]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()))
@pytest.mark.parametrize("shape,q", [((1500, 2040), 0.5), ((2048, 2048), None), ((777, 1028), 2.0)])
def test_switch_matches_reference(checker, switches, cfg, shape, q):
    for k, v in cfg.items():
        _lib.set_param(k, v)
    dem = oracle.fbm_terrain(*shape, seed=shape[0], quantum=q)
    dem[shape[0] // 4: shape[0] // 4 + 40, shape[1] // 3: shape[1] // 3 + 60] = ND
    filled = np.asarray(rd.FillDepressions(R(dem)))
    f_ref = checker.fill_depressions(dem)
    assert np.array_equal(filled, f_ref)
    res = np.asarray(rd.ResolveFlats(R(f_ref)))
    r_ref = checker.resolve_flats(f_ref, ND)
    assert np.array_equal(res.view(np.uint32), r_ref.view(np.uint32))
    assert np.array_equal(np.asarray(rd.FlowDirectionsD8(R(r_ref))), checker.d8_flow_directions(r_ref, ND))
    assert np.array_equal(np.asarray(rd.FlowAccumulation(R(r_ref), "D8")), checker.fa_d8(r_ref, ND))
    assert np.array_equal(np.asarray(rd.FlowAccumulation(R(f_ref), "D8")), checker.fa_d8(f_ref, ND))
    np.testing.assert_allclose(np.asarray(rd.FlowAccumulation(R(r_ref), "Dinf")), checker.fa_dinf(r_ref, ND), rtol=1e-9, atol=0)


@pytest.mark.parametrize("cfg", [{"fill_async": 1}, {"accum_fused_prep": 1, "accum_walk_lanes": 1}, {"fill_multigrid": 8},
                                 {"fill_multigrid": 4, "fill_async": 1}, {"fill_multigrid": 8, "fill_vcycle": 4}],
                         ids=lambda c: ",".join(c))
def test_switch_at_8192_matches_default(switches, cfg):
    import torch
    N = 8192
    L = _lib.lib()
    _lib.use_torch_stream()
    d = torch.empty((N, N), dtype=torch.float32, device="cuda")
    _lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 77, 12, 0.0))
    base = d.clone()
    _lib.check(L.rdb200_dev_fill_depressions_d8_f32(base.data_ptr(), N, N))
    acc0 = torch.empty((N, N), dtype=torch.float64, device="cuda")
    _lib.check(L.rdb200_dev_fa_d8_f32_f64(base.data_ptr(), acc0.data_ptr(), N, N, ND, 1))
    for k, v in cfg.items():
        _lib.set_param(k, v)
    w = d.clone()
    _lib.check(L.rdb200_dev_fill_depressions_d8_f32(w.data_ptr(), N, N))
    assert torch.equal(w, base)
    acc = torch.empty((N, N), dtype=torch.float64, device="cuda")
    _lib.check(L.rdb200_dev_fa_d8_f32_f64(w.data_ptr(), acc.data_ptr(), N, N, ND, 1))
    assert torch.equal(acc, acc0)


@pytest.mark.parametrize("cfg", [{}, {"accum_agg": 1, "accum_tail": 2048}, {"accum_async": 1}], ids=lambda c: ",".join(c) or "default")
def test_eight_receiver_proportions(checker, switches, cfg):
    """FlowAccumulation(props) with up to 8 receivers per cell (equal shares to every lower neighbour)."""
    for k, v in cfg.items():
        _lib.set_param(k, v)
    dem = checker.resolve_flats(checker.fill_depressions(oracle.fbm_terrain(900, 1100, seed=5)), ND)
    h, w = dem.shape
    p = np.zeros((h, w, 9), np.float32)
    d8x = [0, -1, -1, 0, 1, 1, 1, 0, -1]
    d8y = [0, 0, -1, -1, -1, 0, 1, 1, 1]
    for n in range(1, 9):
        sh = np.full_like(dem, np.inf)
        ys = slice(max(0, -d8y[n]), h - max(0, d8y[n]))
        xs = slice(max(0, -d8x[n]), w - max(0, d8x[n]))
        sh[ys, xs] = dem[ys.start + d8y[n]:ys.stop + d8y[n], xs.start + d8x[n]:xs.stop + d8x[n]]
        p[:, :, n] = sh < dem
    p[0, :, :] = p[-1, :, :] = 0
    p[:, 0, :] = p[:, -1, :] = 0
    s_ = p[:, :, 1:].sum(axis=2, keepdims=True)
    p[:, :, 1:] = np.where(s_ > 0, p[:, :, 1:] / np.maximum(s_, 1), 0)
    got = np.asarray(rd.FlowAccumFromProps(rd.rd3array(p, no_data=-2)))
    np.testing.assert_allclose(got, checker.flow_accumulation(p), rtol=1e-9, atol=0)
