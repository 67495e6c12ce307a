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


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
def test_band_fill_equals_single_fill(checker, G):
    dem = oracle.fbm_terrain(700, 900, seed=31, quantum=0.5)
    expected = checker.fill_depressions(dem)
    got, rounds = emulate_bands(dem, G)
    assert np.array_equal(got, expected), f"G={G}: {(got != expected).sum()} cells differ after {rounds} rounds"
