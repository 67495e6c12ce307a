"""Row-band fill on the GPU: the CUDA band solver (rdb200_dev_fill_begin/run/read_row/update_row/
finish) driven through the same protocol as richdem_b200/sharded.py, with all bands living on one
device (sequential emulation of G ranks).  The result must equal the single-band fill exactly for
every G -- the pattern of the reference's programs/parallel_priority_flood/test.py."""
import numpy as np
import pytest

import oracle
from richdem_b200 import _lib, sharded

pytestmark = pytest.mark.gpu
DEV = "cuda"  # tests/test_emulated_kernels.py re-runs these drivers on host memory against the kernel emulation


def emulate_bands(dem: np.ndarray, G: int):
    import torch
    h, w = dem.shape
    solvers, metas = [], []
    for g in range(G):
        r0, r1, gt, gb = sharded.local_rows(h, G, g)
        local = torch.from_numpy(np.ascontiguousarray(dem[r0 - gt:r1 + gb])).to(DEV, copy=True)
        if gt:
            local[0].fill_(float("inf"))
        if gb:
            local[-1].fill_(float("inf"))
        solvers.append(sharded.CudaBandSolver(local.contiguous()))
        metas.append((r0, r1, gt, gb, local.shape[0]))
    rounds = 0
    while True:
        changed = [s.run() for s in solvers]
        rounds += 1
        any_change = False
        for g, (r0, r1, gt, gb, lh) in enumerate(metas):
            if gt and ((changed[g] & 1) or rounds == 1):
                any_change = True
            if gb and ((changed[g] & 2) or rounds == 1):
                any_change = True
        if not any_change or G == 1:
            break
        ups = {g: solvers[g].read_row(1) for g, m in enumerate(metas) if m[2]}
        dns = {g: solvers[g].read_row(m[4] - 2) for g, m in enumerate(metas) if m[3]}
        for g, (r0, r1, gt, gb, lh) in enumerate(metas):
            if gt:
                solvers[g].update_row(0, dns[g - 1])
            if gb:
                solvers[g].update_row(lh - 1, ups[g + 1])
        assert rounds < 1000
    out = np.empty_like(dem)
    for g, (r0, r1, gt, gb, lh) in enumerate(metas):
        full = solvers[g].finish().cpu().numpy()
        out[r0:r1] = full[gt:gt + (r1 - r0)]
    return out, rounds


def test_band_fill_ghost_row_on_tile_boundary(checker):
    """Band heights chosen so a ghost row is the first row of a 64-row tile (its only real row):
    replacing it must wake the tile row above, whose apron it is."""
    dem = oracle.fbm_terrain(256, 320, seed=33, quantum=0.5)   # 2 bands of 128 rows -> ghost at local row 128
    expected = checker.fill_depressions(dem)
    for G in (2, 4):
        got, _ = emulate_bands(dem, G)
        assert np.array_equal(got, expected), f"G={G}"


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_band_fill_equals_single_fill(checker, G):
    dem = oracle.fbm_terrain(700, 900, seed=31, quantum=0.5)
    expected = checker.fill_depressions(dem)
    got, rounds = emulate_bands(dem, G)
    assert np.array_equal(got, expected), f"G={G}: {(got != expected).sum()} cells differ after {rounds} rounds"


def emulate_fa_bands(dem: np.ndarray, G: int, nodata: float, dinf: bool, weights=None):
    """Drive G CudaBandAccumulators on one device through the fa_band protocol."""
    import torch
    h, w = dem.shape
    accs, metas, outs = [], [], []
    for g in range(G):
        r0, r1, gt, gb = sharded.local_rows(h, G, g)
        local = torch.from_numpy(np.ascontiguousarray(dem[r0 - gt:r1 + gb])).to(DEV, copy=True).contiguous()
        if weights is None:
            acc = torch.empty(local.shape, dtype=torch.float64, device=DEV)
        else:
            acc = torch.from_numpy(np.ascontiguousarray(weights[r0 - gt:r1 + gb])).to(DEV, copy=True).contiguous()
        A = sharded.CudaBandAccumulator(local, acc, nodata, gt, gb, dinf, weights is None)
        accs.append(A)
        outs.append(acc)
        metas.append((r0, r1, gt, gb))
    for g, (r0, r1, gt, gb) in enumerate(metas):  # exchange edge codes
        if gt:
            c = accs[g - 1].edge_codes(1)
            accs[g].set_ghost_codes(0, c[0], c[1])
        if gb:
            c = accs[g + 1].edge_codes(0)
            accs[g].set_ghost_codes(1, c[0], c[1])
    rounds = 0
    while True:
        sent = [A.run() for A in accs]
        rounds += 1
        if not any(a + b for a, b in sent):
            break
        ups = {g: accs[g].take_outflow(0) for g, m in enumerate(metas) if m[2]}
        dns = {g: accs[g].take_outflow(1) for g, m in enumerate(metas) if m[3]}
        for g, (r0, r1, gt, gb) in enumerate(metas):
            if gt:
                accs[g].apply_inflow(0, *dns[g - 1])
            if gb:
                accs[g].apply_inflow(1, *ups[g + 1])
        assert rounds < 10000
    out = np.empty((h, w), np.float64)
    for g, (r0, r1, gt, gb) in enumerate(metas):
        accs[g].finish()
        out[r0:r1] = outs[g][gt:gt + (r1 - r0)].cpu().numpy()
    return out, rounds


@pytest.mark.parametrize("G", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("dinf", [False, True])
def test_band_accumulation_equals_single(checker, G, dinf):
    nd = -9999.0
    dem = oracle.fbm_terrain(640, 500, seed=41, quantum=0.25)
    dem[100:140, 200:260] = nd
    filled = checker.fill_depressions(dem)
    resolved = checker.resolve_flats(filled, nd)
    expected = checker.fa_dinf(resolved, nd) if dinf else checker.fa_d8(resolved, nd)
    got, rounds = emulate_fa_bands(resolved, G, nd, dinf)
    if dinf:  # unit weights: the packed fixed-point walk (relative error < 2^-23 by construction)
        np.testing.assert_allclose(got, expected, rtol=5e-7, atol=0)
        _lib.set_param("accum_dinf_packed", 0)  # the level kernel's double atomics
        try:
            got0, _ = emulate_fa_bands(resolved, G, nd, dinf)
        finally:
            _lib.reset_params()
        np.testing.assert_allclose(got0, expected, rtol=1e-9, atol=0)
    else:
        assert np.array_equal(got, expected), f"G={G}: {(got != expected).sum()} cells differ ({rounds} rounds)"


def test_band_accumulation_with_weights(checker):
    nd = -9999.0
    dem = checker.resolve_flats(checker.fill_depressions(oracle.fbm_terrain(300, 280, seed=43)), nd)
    wts = np.random.default_rng(1).random(dem.shape)
    got, _ = emulate_fa_bands(dem, 4, nd, False, weights=wts)
    np.testing.assert_allclose(got, checker.fa_d8(dem, nd, wts), rtol=1e-9, atol=0)


# ---- row-band flat resolution: G bands on one device, same seam protocol as sharded.resolve_flats_band ----
def _relax_emulated(solvers, metas):
    rounds = 0
    while True:
        changed = [s.run() for s in solvers]
        rounds += 1
        any_edge = any((m[2] and ((c & 1) or rounds == 1)) or (m[3] and ((c & 2) or rounds == 1))
                       for c, m in zip(changed, metas))
        any_active = any(c & 4 for c in changed)
        if not any_edge and not any_active:
            return rounds
        if not any_edge:
            continue
        ups = {g: solvers[g].read_row(1) for g, m in enumerate(metas) if m[2]}
        dns = {g: solvers[g].read_row(m[4] - 2) for g, m in enumerate(metas) if m[3]}
        for g, m in enumerate(metas):
            if m[2]:
                solvers[g].update_row(0, dns[g - 1])
            if m[3]:
                solvers[g].update_row(m[4] - 1, ups[g + 1])
        assert rounds < 10000


def emulate_flats_bands(dem: np.ndarray, G: int, nodata: float):
    import torch
    h, w = dem.shape
    F, metas, locals_ = [], [], []
    for g in range(G):
        r0, r1, gt, gb = sharded.local_rows(h, G, g)
        local = torch.from_numpy(np.ascontiguousarray(dem[r0 - gt:r1 + gb])).to(DEV, copy=True).contiguous()
        locals_.append(local)
        F.append(sharded.CudaFlatsBand(local, nodata, gt, gb))
        metas.append((r0, r1, gt, gb, local.shape[0]))

    def exchange_ft():
        rows = [(f.ft[f.rows(0)[0]].clone(), f.ft[f.rows(1)[0]].clone()) for f in F]
        for g, m in enumerate(metas):
            if m[2]:
                F[g].ft[0].copy_(rows[g - 1][1])
            if m[3]:
                F[g].ft[m[4] - 1].copy_(rows[g + 1][0])

    def merge(payload, apply):
        iters = 0
        while G > 1:
            iters += 1
            pay = [(getattr(f, payload)(0) if m[2] else None, getattr(f, payload)(1) if m[3] else None)
                   for f, m in zip(F, metas)]
            ch = False
            for g, m in enumerate(metas):
                if m[2]:
                    ch |= getattr(F[g], apply)(0, pay[g - 1][1])
                if m[3]:
                    ch |= getattr(F[g], apply)(1, pay[g + 1][0])
            if not ch:
                break
            assert iters < 100
        return iters

    exchange_ft()
    for f in F:
        f.step("edges")
    exchange_ft()
    for f in F:
        f.step("components")
    it1 = merge("flag_payload", "merge_flags")
    for f in F:
        f.step("labels")
    it2 = 0
    for away in (True, False):
        solvers = [f.gradient_begin(away) for f in F]
        _relax_emulated(solvers, metas)
        for f, s in zip(F, solvers):
            f.gradient_end(away, s)
        if away:
            it2 = merge("height_payload", "merge_heights")
    out = np.empty_like(dem)
    for g, (r0, r1, gt, gb, lh) in enumerate(metas):
        F[g].step("apply")
        F[g].finish()
        out[r0:r1] = locals_[g][gt:gt + (r1 - r0)].cpu().numpy()
    return out, it1, it2


@pytest.mark.parametrize("G", [1, 2, 3, 5, 8])
def test_band_flat_resolution_equals_single(checker, G):
    nd = -9999.0
    dem = oracle.fbm_terrain(640, 520, seed=51, quantum=0.5)
    dem[300:330, 100:180] = nd
    filled = checker.fill_depressions(dem)
    expected = checker.resolve_flats(filled, nd)
    assert (expected != filled).mean() > 0.05
    got, it1, it2 = emulate_flats_bands(filled, G, nd)
    bad = (got.view(np.uint32) != expected.view(np.uint32)).sum()
    assert bad == 0, f"G={G}: {bad} cells differ (flag iterations {it1}, height iterations {it2})"


def test_band_flat_resolution_snaking_flat(checker):
    """One flat that crosses every seam several times (a comb), so outlet flags and flat heights have
    to travel through several seam iterations."""
    nd = -9999.0
    dem = np.full((96, 64), 10.0, np.float32)
    dem[:, ::4] = 5.0                      # vertical channels at elevation 5 ...
    dem[2, :] = 5.0                        # ... joined at the top
    dem[93, 1::8] = 5.0
    dem[0, :] = dem[-1, :] = 20.0
    dem[:, 0] = dem[:, -1] = 20.0
    dem[94, 4] = 1.0                       # a single outlet near the bottom
    dem[95, 4] = 0.0
    filled = checker.fill_depressions(dem)
    expected = checker.resolve_flats(filled, nd)
    for G in (2, 4, 6):
        got, it1, it2 = emulate_flats_bands(filled, G, nd)
        assert np.array_equal(got.view(np.uint32), expected.view(np.uint32)), (G, it1, it2)
