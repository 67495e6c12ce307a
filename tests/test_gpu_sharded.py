"""Row-band fill on the GPU: the CUDA band solver (rdb200_dev_fill_begin/run/read_row/update_row/
finish) driven through the same protocol as richdem_b200/sharded.py, with all bands living on one
device (sequential emulation of G ranks).  The result must equal the single-band fill exactly for
every G -- the pattern of the reference's programs/parallel_priority_flood/test.py."""
import numpy as np
import pytest

import oracle
from richdem_b200 import sharded

pytestmark = pytest.mark.gpu


def emulate_bands(dem: np.ndarray, G: int):
    import torch
    h, w = dem.shape
    solvers, metas = [], []
    for g in range(G):
        r0, r1, gt, gb = sharded.local_rows(h, G, g)
        local = torch.from_numpy(np.ascontiguousarray(dem[r0 - gt:r1 + gb])).cuda()
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
        local = torch.from_numpy(np.ascontiguousarray(dem[r0 - gt:r1 + gb])).cuda().contiguous()
        if weights is None:
            acc = torch.empty(local.shape, dtype=torch.float64, device="cuda")
        else:
            acc = torch.from_numpy(np.ascontiguousarray(weights[r0 - gt:r1 + gb])).cuda().contiguous()
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
    if dinf:
        np.testing.assert_allclose(got, expected, rtol=1e-9, atol=0)
    else:
        assert np.array_equal(got, expected), f"G={G}: {(got != expected).sum()} cells differ ({rounds} rounds)"


def test_band_accumulation_with_weights(checker):
    nd = -9999.0
    dem = checker.resolve_flats(checker.fill_depressions(oracle.fbm_terrain(300, 280, seed=43)), nd)
    wts = np.random.default_rng(1).random(dem.shape)
    got, _ = emulate_fa_bands(dem, 4, nd, False, weights=wts)
    np.testing.assert_allclose(got, checker.fa_d8(dem, nd, wts), rtol=1e-9, atol=0)
