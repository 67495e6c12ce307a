"""Kernel-logic parity WITHOUT a GPU: the shipped CUDA sources (richdem_b200/csrc/*.cu, unmodified) are
rewritten against a single-threaded fiber model of CUDA (tests/emu/) and compiled with g++ into
tests/_bin/librdb200_emu_test.so; the GPU parity tests' own checks are then run against that library.

This is test infrastructure, not a backend: the product loader (richdem_b200/_lib.py) refuses to load
the emulation build, and only this module -- by swapping the already-loaded library handle inside a
fixture -- ever points the Python layer at it.  It checks the kernels' logic (worklists, dirty-block
lists, union-find, dependency counters, level schedules, band protocols); it cannot say anything about
TMA staging, memory ordering, races or speed -- the `-m gpu` suite on the B200 remains the parity gate.
"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np
import pytest

import oracle
from richdem_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def emu_lib():
    if sys.platform != "linux" or os.uname().machine != "x86_64":
        pytest.skip("the fiber switch of tests/emu is x86-64 SysV only")
    build_emu = _load_module("build_emu", os.path.join(HERE, "emu", "build_emu.py"))
    path = build_emu.build()
    L = C.CDLL(str(path))
    assert L.rdb200_emulated() == 1
    for name, argtypes in _lib.SIGNATURES.items():
        f = getattr(L, name)
        f.argtypes = argtypes
        f.restype = C.c_int
    L.rdb200_last_error.restype = C.c_char_p
    L.rdb200_last_error.argtypes = []
    L.rdb200_version.restype = C.c_int
    L.rdb200_shutdown.restype = None
    return L


@pytest.fixture()
def emulated(emu_lib, monkeypatch):
    """Point the Python layer at the emulation for ONE test, then restore the real (absent) handle."""
    monkeypatch.setattr(_lib, "_lib", emu_lib)
    _lib.init(0)
    _lib.set_param("fill_use_tma", 0)  # TMA / mbarrier PTX is not emulated
    yield emu_lib
    _lib.reset_params()


@pytest.fixture(scope="module")
def gp():
    return _load_module("gpu_parity_checks", os.path.join(HERE, "test_gpu_parity.py"))


def test_loader_refuses_the_emulation_build(emu_lib, monkeypatch):
    build_emu = sys.modules.get("build_emu") or _load_module("build_emu", os.path.join(HERE, "emu", "build_emu.py"))
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(build_emu.LIB))
    with pytest.raises(_lib.RichdemB200Error, match="no CPU fallback"):
        _lib.lib()


def test_golden_fixtures(emulated, gp, golden):
    gp.test_fill_known_answer(golden)
    gp.test_d8_flow_accum_known_answers(golden)
    gp.test_data_dems(golden)
    gp.test_synthetic_golden(golden)


def test_beauford_crop(emulated, gp, golden):
    gp.test_beauford_crop_golden(golden)


@pytest.mark.parametrize("shape,seed,q,patch", [
    ((150, 220), 1, None, False), ((257, 131), 2, 0.5, False), ((400, 500), 3, 2.0, True),
    ((65, 129), 4, 1.0, False), ((64, 64), 5, 10.0, False), ((63, 200), 6, None, True),
])
def test_pipeline_vs_oracle(emulated, gp, checker, shape, seed, q, patch):
    gp.test_pipeline_vs_oracle(checker, shape, seed, q, patch)


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (2, 2), (3, 3), (3, 64), (5, 1), (4, 130)])
def test_degenerate_shapes(emulated, gp, checker, shape):
    gp.test_degenerate_shapes(checker, shape)


def test_special_rasters(emulated, gp, checker):
    gp.test_special_rasters(checker)


def test_in_place_and_copy_semantics(emulated, gp):
    gp.test_in_place_and_copy_semantics()


@pytest.mark.parametrize("param,value", [
    ("fill_ordered", 0), ("fill_max_iters", 1), ("fill_max_iters", 2), ("fill_rounds_per_sync", 1), ("fill_order_rounds", 40),
    ("flats_tiled", 0), ("accum_packed", 0), ("accum_budget", 1), ("accum_budget", 64),
    ("accum_walk_lanes", 0), ("accum_fused_prep", 0), ("flats_uf_tiled", 0), ("flats_fused_classify", 0), ("flats_pair", 0), ("flowdirs_rolling", 0), ("accum_dinf_packed", 1), ("accum_dinf_packed", 0),
    ("fill_multigrid", 4), ("fill_multigrid", 0), ("fill_vcycle", 0), ("fill_vcycle", 2),
])
def test_algorithm_variants_agree(emulated, gp, checker, param, value):
    """Every tunable is a schedule / layout choice; none may change a result."""
    _lib.set_param("fill_multigrid_min", 32)  # so that a 300 x 420 raster gets two coarse levels
    _lib.set_param(param, value)
    dem = oracle.fbm_terrain(300, 420, seed=17, quantum=0.5)
    dem[40:70, 100:180] = gp.ND
    gp.check_pipeline(dem, gp.ND, checker, dinf_rtol=gp.ACC_RTOL if (param, value) == ("accum_dinf_packed", 0) else None)


def test_level_schedule_engages_and_matches(emulated, gp, checker):
    """A raster wide enough (>= 10 tiles across) for the level-ordered admission to be active."""
    import richdem_b200 as rd
    dem = oracle.fbm_terrain(200, 900, seed=23)
    got = np.asarray(rd.FillDepressions(gp.R(dem)))
    assert np.array_equal(got, checker.fill_depressions(dem))
    s = _lib.stats()
    assert s["fill_tile_visits"] > 0 and s["fill_rounds"] > 0


# ---- row-band (multi-GPU) protocols: the G-bands-on-one-device drivers of test_gpu_sharded.py, on host memory ----
@pytest.fixture()
def band_drivers(emulated, monkeypatch):
    import torch
    from richdem_b200 import sharded

    def host_view(ptr, shape, typestr, device):
        dt = np.dtype(typestr)
        n = int(np.prod(shape))
        buf = (C.c_char * (n * dt.itemsize)).from_address(int(ptr))
        return torch.from_numpy(np.frombuffer(buf, dtype=dt, count=n).reshape(shape))

    monkeypatch.setattr(sharded, "_on_device", lambda t: True)       # "device" memory is host memory here
    monkeypatch.setattr(sharded, "_view", host_view)
    monkeypatch.setattr(_lib, "use_torch_stream", lambda: None)
    gs = _load_module("gpu_sharded_drivers", os.path.join(HERE, "test_gpu_sharded.py"))
    gs.DEV = "cpu"
    return gs


@pytest.mark.parametrize("G", [1, 2, 3, 5])
def test_band_fill(band_drivers, checker, G):
    dem = oracle.fbm_terrain(330, 300, seed=31, quantum=0.5)
    got, rounds = band_drivers.emulate_bands(dem, G)
    assert np.array_equal(got, checker.fill_depressions(dem)), f"G={G} after {rounds} exchanges"


def test_band_fill_ghost_row_on_tile_boundary(band_drivers, checker):
    band_drivers.test_band_fill_ghost_row_on_tile_boundary(checker)


@pytest.mark.parametrize("G", [2, 3, 5])
@pytest.mark.parametrize("dinf", [False, True, "nolanes"])
def test_band_accumulation(band_drivers, checker, G, dinf):
    nd = -9999.0
    if dinf == "nolanes":  # unit-weight D8 with one thread per source instead of the persistent-lane walk
        _lib.set_param("accum_walk_lanes", 0)
        dinf = False
    dem = oracle.fbm_terrain(260, 212, seed=41, quantum=0.25)  # W % 4 == 0: the packed / fused band path
    dem[100:140, 60:120] = nd
    resolved = checker.resolve_flats(checker.fill_depressions(dem), nd)
    got, _ = band_drivers.emulate_fa_bands(resolved, G, nd, dinf)
    if dinf:  # unit weights: the packed fixed-point walk in band mode; then the level kernel's double atomics
        np.testing.assert_allclose(got, checker.fa_dinf(resolved, nd), rtol=5e-7, atol=0)
        _lib.set_param("accum_dinf_packed", 0)
        got0, _ = band_drivers.emulate_fa_bands(resolved, G, nd, dinf)
        np.testing.assert_allclose(got0, checker.fa_dinf(resolved, nd), rtol=1e-9, atol=0)
    else:
        assert np.array_equal(got, checker.fa_d8(resolved, nd))


def test_band_accumulation_with_weights(band_drivers, checker):
    band_drivers.test_band_accumulation_with_weights(checker)


@pytest.mark.parametrize("G", [2, 3, 5, "uf_global"])
def test_band_flat_resolution(band_drivers, checker, G):
    nd = -9999.0
    if G == "uf_global":  # one-level union-find instead of the tiled one
        _lib.set_param("flats_uf_tiled", 0)
        G = 3
    dem = oracle.fbm_terrain(300, 260, seed=51, quantum=0.5)
    dem[150:170, 60:120] = nd
    filled = checker.fill_depressions(dem)
    expected = checker.resolve_flats(filled, nd)
    got, it1, it2 = band_drivers.emulate_flats_bands(filled, G, nd)
    assert np.array_equal(got.view(np.uint32), expected.view(np.uint32)), (G, it1, it2)


def test_band_flat_resolution_snaking_flat(band_drivers, checker):
    band_drivers.test_band_flat_resolution_snaking_flat(checker)


@pytest.mark.parametrize("shape", [(70, 2040), (130, 1016), (67, 1020), (3, 1024), (200, 4), (129, 8), (66, 1012)])
def test_fused_d8_preparation_block_seams(emulated, gp, checker, shape):
    """accum_fused_prep: blocks own 1016 output columns / 64 rows and overlap by a halo; widths and heights around
    those seams, with and without the persistent-lane walk."""
    import richdem_b200 as rd
    dem = oracle.fbm_terrain(*shape, seed=sum(shape), quantum=0.5)
    dem[shape[0] // 3: shape[0] // 3 + 5, shape[1] // 2: shape[1] // 2 + 9] = gp.ND
    expected = checker.fa_d8(dem, gp.ND)
    for lanes in (0, 1):
        _lib.set_param("accum_walk_lanes", lanes)
        got = np.asarray(rd.FlowAccumulation(gp.R(dem), "D8"))
        assert np.array_equal(got, expected), (shape, lanes)


@pytest.mark.parametrize("k", [2, 3, 4, 8])
@pytest.mark.parametrize("engine", ["rounds", "vcycle1", "vcycle3"])
def test_multigrid_seeded_fill(emulated, gp, checker, k, engine):
    """fill_multigrid: the flood starts from the lifted fill of the k x k max-pooled raster (recursively) instead of
    +inf; ragged block edges, NoData, plateaus.  Any upper bound must relax to the exact surface."""
    import richdem_b200 as rd
    _lib.set_param("fill_multigrid", k)
    _lib.set_param("fill_multigrid_min", 32)
    _lib.set_param("fill_vcycle", 0)
    if engine.startswith("vcycle"):  # coarse-grid corrections (restrict / coarse relax / prolong) every 1 or 3 fine rounds
        _lib.set_param("fill_vcycle", int(engine[-1]))
        _lib.set_param("fill_rounds_per_sync", 2)
    for shape, q in (((301, 423), 0.5), ((130, 70), None), ((65, 129), 5.0), ((33, 35), None), ((500, 640), None)):
        dem = oracle.fbm_terrain(*shape, seed=k + shape[0], quantum=q)
        dem[shape[0] // 3: shape[0] // 3 + 7, shape[1] // 2: shape[1] // 2 + 9] = gp.ND
        got = np.asarray(rd.FillDepressions(gp.R(dem)))
        assert np.array_equal(got, checker.fill_depressions(dem)), (k, engine, shape)


def test_vcycle_prolongation_wakes_neighbouring_tiles(emulated, gp, checker):
    """ADVICE round 1 regression (walled lake next to a tile seam) on the CPU model of the kernels."""
    import richdem_b200 as rd
    dem = gp.walled_lake_case()
    expected = checker.fill_depressions(dem)
    for cfg in ({"fill_multigrid": 8, "fill_multigrid_min": 32, "fill_vcycle": 1},
                {"fill_multigrid": 4, "fill_multigrid_min": 32, "fill_vcycle": 2},
                {"fill_multigrid": 8, "fill_multigrid_min": 32, "fill_vcycle": 0}):
        for k, v in cfg.items():
            _lib.set_param(k, v)
        assert np.array_equal(np.asarray(rd.FlowDirectionsD8(gp.R(dem))).shape, dem.shape)
        assert np.array_equal(np.asarray(rd.FillDepressions(gp.R(dem))), expected), cfg


@pytest.mark.parametrize("shape,seed,q", [((150, 220), 1, None), ((400, 500), 3, 2.0), ((64, 64), 5, 10.0), ((3, 3), 9, None)])
def test_d4_fill(emulated, gp, checker, shape, seed, q):
    gp.test_d4_fill_vs_oracle(checker, shape, seed, q)
    _lib.set_param("fill_multigrid", 4)
    _lib.set_param("fill_multigrid_min", 32)
    _lib.set_param("fill_vcycle", 2)
    gp.test_d4_fill_vs_oracle(checker, shape, seed, q)


@pytest.mark.parametrize("share", [0, 3, -1])
def test_packed_dinf_work_sharing(emulated, gp, checker, share):
    """accum_dinf_packed: fixed-point D-infinity walk in phases; share = ring entries above which a warp asks for a
    rebalancing phase once a quarter of the warps wait at the barrier (0: any queued cell does).  Re-run with concurrent blocks by test_cooperative_kernels_with_several_blocks."""
    import richdem_b200 as rd
    _lib.set_param("accum_dinf_packed", 1)
    _lib.set_param("accum_dinf_share", share)
    dem = checker.resolve_flats(checker.fill_depressions(oracle.fbm_terrain(420, 520, seed=71, quantum=0.5)), gp.ND)
    dem[200:230, 100:160] = gp.ND
    got = np.asarray(rd.FlowAccumulation(gp.R(dem), "Dinf"))
    np.testing.assert_allclose(got, checker.fa_dinf(dem, gp.ND), rtol=gp.DINF_UNIT_RTOL, atol=0)


def test_direction_grid_flat_resolution(emulated, gp, checker, golden):
    gp.test_flow_directions_with_resolved_flats_golden(golden)
    gp.test_flow_directions_with_resolved_flats_vs_oracle(checker, (300, 420), 2, 0.5)
    gp.test_flow_directions_with_resolved_flats_vs_oracle(checker, (64, 70), 5, 10.0)


def _spread_to_all_lower_neighbours(dem):
    """An 8-receiver proportions grid (equal shares to every lower neighbour); edge cells carry no flow, as in every
    FM_* output (the reference's accumulation never bounds-checks receivers, flow_accumulation_generic.hpp:84-87)."""
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
    return p


def test_eight_receiver_proportions(emulated, gp, checker):
    """FlowAccumulation(props) on graphs with up to 8 receivers per cell (a cone with huge fan-in, and fBm)."""
    import richdem_b200 as rd
    yy, xx = np.mgrid[0:201, 0:231]
    cone = np.hypot(yy - 100, xx - 115).astype(np.float32)
    for dem in (cone, checker.resolve_flats(checker.fill_depressions(oracle.fbm_terrain(260, 300, seed=5)), gp.ND)):
        props = _spread_to_all_lower_neighbours(dem)
        got = np.asarray(rd.FlowAccumFromProps(rd.rd3array(props, no_data=-2)))
        np.testing.assert_allclose(got, checker.flow_accumulation(props), rtol=1e-9, atol=0)


def test_cooperative_kernels_with_several_blocks():
    """The cooperative kernels (multi-receiver level kernel, the persistent BFS) size their grid from the SM count; re-run their cases with 3 emulated SMs so that they execute as
    3 blocks side by side, each with its own shared memory and a real grid barrier."""
    import subprocess
    if os.environ.get("RDB_EMU_SMS"):
        pytest.skip("already inside the multi-block run")
    env = dict(os.environ, RDB_EMU_SMS="3", RDB_EMU_CHAOS="7")  # CHAOS: atomics yield at random -> other interleavings
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", os.path.abspath(__file__), "-k",
                        "variants or band_accumulation or special_rasters or degenerate or eight_receiver or packed_dinf"],
                       env=env, cwd=os.path.dirname(HERE), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("shape,q", [((40, 200), 2.0), ((16, 64), 5.0), ((17, 65), 5.0), ((100, 130), 20.0), ((33, 129), 1000.0)])
def test_tiled_union_find_seams(emulated, gp, checker, shape, q):
    """flats_uf_tiled: flats that cross the 64x16 union-find tiles in every direction (coarse quantisation makes
    large plateaus; the last case is one flat covering the raster)."""
    import richdem_b200 as rd
    dem = checker.fill_depressions(oracle.fbm_terrain(*shape, seed=shape[1], quantum=q))
    dem[shape[0] // 2, shape[1] // 3: shape[1] // 3 + 4] = gp.ND
    m, l = rd.FlatMask(gp.R(dem))
    m_ref, l_ref = checker.flat_mask(dem, gp.ND)
    assert np.array_equal(m, m_ref)
    pairs = np.unique(np.stack([l[l != 0], l_ref[l_ref != 0]]), axis=1)
    assert len(np.unique(pairs[0])) == pairs.shape[1] == len(np.unique(pairs[1]))
    got = np.asarray(rd.ResolveFlats(gp.R(dem)))
    assert np.array_equal(got.view(np.uint32), checker.resolve_flats(dem, gp.ND).view(np.uint32))


@pytest.mark.parametrize("shape", [(70, 2052), (130, 1024), (65, 1028), (3, 8), (200, 4), (64, 64), (1, 12)])
def test_rolling_flow_directions(emulated, gp, checker, shape):
    """flowdirs_rolling: 4 columns per thread, 64-row chunks, halo columns by shuffle; widths around the block seams."""
    import richdem_b200 as rd
    dem = checker.resolve_flats(checker.fill_depressions(oracle.fbm_terrain(*shape, seed=shape[1], quantum=0.5)), gp.ND)
    dem[shape[0] // 2:, shape[1] // 2: shape[1] // 2 + 3] = gp.ND
    assert np.array_equal(np.asarray(rd.FlowDirectionsD8(gp.R(dem))), checker.d8_flow_directions(dem, gp.ND))


def test_terrain_attributes(emulated, gp, checker, golden):
    """the TA_* stencil kernel -- window seams, NoData and raster-edge neighbours, all eight attributes."""
    gp.test_terrain_attributes_golden(golden)
    for shape in [(1, 9), (7, 1), (2, 2), (130, 129), (17, 257)]:
        gp.test_terrain_attributes_vs_oracle(checker, shape)
    gp.test_terrain_attribute_rejects_unknown_names()
