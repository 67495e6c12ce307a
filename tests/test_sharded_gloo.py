"""world_size-2/3 gloo tests (CPU) of the row-band protocol in richdem_b200/sharded.py.

The band solver is injected: here a CPU stand-in built on the oracle that solves
"fill with fixed first/last rows" exactly like the CUDA band solver's contract
(include/richdem_b200.h, rdb200_dev_fill_*).  What is under test is the partitioning, the halo
exchange, the ghost-row bookkeeping and the termination rule -- the sharded result must equal the
single-band answer bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from richdem_b200 import sharded  # noqa: E402


def test_band_bounds():
    assert sharded.band_bounds(10, 1) == [(0, 10)]
    assert sharded.band_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    b = sharded.band_bounds(32768, 8)
    assert b[0] == (0, 4096) and b[-1] == (28672, 32768)
    assert sharded.local_rows(100, 4, 0) == (0, 25, 0, 1)
    assert sharded.local_rows(100, 4, 2) == (50, 75, 1, 1)
    assert sharded.local_rows(100, 4, 3) == (75, 100, 1, 0)
    with pytest.raises(ValueError):
        sharded.band_bounds(3, 4)


class OracleBandSolver:
    """CPU stand-in with the CudaBandSolver contract: first/last local rows are fixed boundary
    values, left/right columns are raster border.  A fill of the local raster in which the fixed
    rows are border rows is exactly that (the priority flood pins all four sides)."""

    def __init__(self, local_dem: torch.Tensor):
        import oracle
        self.O = oracle.port()
        self.Z = local_dem.numpy().copy()
        self.h, self.w = self.Z.shape
        self.W = None
        self.prev = None

    def _solve(self):
        z = self.Z.copy()
        big = np.float32(3.0e38)
        z[~np.isfinite(z)] = big  # +inf ghost rows: "no information yet"
        w = self.O.fill_depressions(z)
        w[w >= big] = np.inf
        return w

    def run(self) -> int:
        new = self._solve()
        ch = 0
        if self.W is None or not np.array_equal(new[1], self.W[1]):
            ch |= 1
        if self.W is None or not np.array_equal(new[self.h - 2], self.W[self.h - 2]):
            ch |= 2
        self.W = new
        return ch

    def read_row(self, y):
        return torch.from_numpy(self.W[y].copy())

    def update_row(self, y, row):
        assert y in (0, self.h - 1)
        new = row.numpy()
        assert (new <= self.Z[y]).all()
        self.Z[y] = new

    def finish(self):
        return torch.from_numpy(self.W.copy())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dem, expected, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h, w = dem.shape
        local, (r0, r1, gt, gb) = sharded.scatter_rows(dem if rank == 0 else None, h, w, torch.float32, "cpu")
        filled, rounds = sharded.fill_band(local, gt, gb, solver_cls=OracleBandSolver)
        own = filled[gt: gt + (r1 - r0)].numpy()
        ok = np.array_equal(own, expected[r0:r1])
        out_q.put((rank, bool(ok), int(rounds)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,seed", [(2, (90, 70), 3), (3, (120, 64), 4)])
def test_sharded_fill_protocol_gloo(world, shape, seed):
    import oracle
    dem = oracle.fbm_terrain(*shape, seed=seed, quantum=2.0)
    expected = oracle.port().fill_depressions(dem)
    assert (expected != dem).mean() > 0.02
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dem, expected, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in results), results
    assert max(r for _, _, r in results) >= 2  # at least one real exchange happened
