"""The multi-process row-band pipeline (richdem_b200/sharded.py: fill_band -> resolve_flats_band -> exchange_rows ->
fa_band, D8 and D-infinity) over torch.distributed with the gloo backend, one process per band, with the band solvers
running the SHIPPED kernels on the CPU model of tests/emu (host memory stands in for HBM).  This is the code path
`torchrun` takes on N GPUs -- CudaBandSolver / CudaFlatsBand / CudaBandAccumulator, halo exchange, seam merges,
termination votes -- which the gloo test with the oracle solver (test_sharded_gloo.py) only covers for the fill.
"""
import ctypes as C
import importlib.util
import multiprocessing as mp
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ND = -9999.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _load_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, lib_path, dem, expected, params, out_q):
    import torch
    import torch.distributed as dist
    from richdem_b200 import _lib, sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    try:
        # point this process's Python layer at the kernel emulation (tests only; the loader itself refuses it)
        L = C.CDLL(lib_path)
        for name, argtypes in _lib.SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = argtypes
            f.restype = C.c_int
        L.rdb200_last_error.restype = C.c_char_p
        L.rdb200_last_error.argtypes = []
        _lib._lib = L
        _lib.use_torch_stream = lambda: None
        sharded._on_device = lambda t: True

        def host_view(ptr, shape, typestr, device):
            dt = np.dtype(typestr)
            n = int(np.prod(shape))
            buf = (C.c_char * (n * dt.itemsize)).from_address(int(ptr))
            return torch.from_numpy(np.frombuffer(buf, dtype=dt, count=n).reshape(shape))

        sharded._view = host_view
        _lib.init(0)
        _lib.set_param("fill_use_tma", 0)
        # the C++ band driver (rdb200_mgpu_*, over the callback communicator here) takes its multigrid settings from the
        # library switches; the test raster is small, so the coarse levels are allowed down to 16 cells
        band_multigrid = params.pop("_band_multigrid", 0)
        band_vcycle = params.pop("_band_vcycle", 0)
        _lib.set_param("fill_multigrid", band_multigrid)
        _lib.set_param("fill_multigrid_min", 16)
        if band_vcycle:
            _lib.set_param("fill_vcycle", band_vcycle)
        for k, v in params.items():
            _lib.set_param(k, v)

        dist.init_process_group("gloo", rank=rank, world_size=world)
        h, w = dem.shape
        local, (r0, r1, gt, gb) = sharded.scatter_rows(dem if rank == 0 else None, h, w, torch.float32, "cpu")
        own = slice(gt, gt + (r1 - r0))
        res = {}
        filled, _ = sharded.fill_band(local, gt, gb, multigrid=band_multigrid, row0=r0 - gt, height=h, vcycle=band_vcycle)
        res["fill"] = np.array_equal(filled[own].numpy(), expected["fill"][r0:r1])
        filled = filled.contiguous()
        sharded.resolve_flats_band(filled, gt, gb, ND)
        res["flats"] = np.array_equal(filled[own].numpy().view(np.uint32), expected["flats"][r0:r1].view(np.uint32))
        sharded.exchange_rows(filled, gt, gb)
        acc, _ = sharded.fa_band(filled, gt, gb, ND, dinf=False)
        res["fa_d8"] = np.array_equal(acc[own].numpy(), expected["fa_d8"][r0:r1])
        acc, _ = sharded.fa_band(filled, gt, gb, ND, dinf=True)
        a, e = acc[own].numpy(), expected["fa_dinf"][r0:r1]
        res["fa_dinf"] = bool(np.all(np.abs(a - e) <= 5e-7 * np.maximum(1.0, np.abs(e))))  # packed fixed-point sums (< 2^-23)
        out_q.put((rank, res, None))
    except Exception as exc:  # surface the failure in the parent instead of a silent non-zero exit
        import traceback
        out_q.put((rank, {}, traceback.format_exc() + repr(exc)))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("world,params", [(2, {}), (3, {}), (3, {"accum_walk_lanes": 0, "flats_uf_tiled": 0, "accum_fused_prep": 0}),
                                          (3, {"_band_multigrid": 4}), (2, {"_band_multigrid": 8}),
                                          (4, {"_band_multigrid": 3, "fill_multigrid": 2, "fill_multigrid_min": 16, "fill_band_multigrid": 3}),
                                          (3, {"_band_multigrid": 4, "_band_vcycle": 1, "fill_band_rounds": 2, "fill_rounds_per_sync": 2}),
                                          (2, {"_band_multigrid": 8, "_band_vcycle": 2, "fill_band_rounds": 1, "fill_rounds_per_sync": 1})],
                         ids=["2-ranks", "3-ranks", "3-ranks-round1-kernels", "3-ranks-multigrid", "2-ranks-multigrid8",
                              "4-ranks-multigrid-recursive", "3-ranks-vcycle", "2-ranks-vcycle"])
def test_sharded_pipeline_on_emulated_kernels(world, params):
    if sys.platform != "linux" or os.uname().machine != "x86_64":
        pytest.skip("the fiber switch of tests/emu is x86-64 SysV only")
    import oracle
    lib_path = str(_load_module("build_emu", os.path.join(HERE, "emu", "build_emu.py")).build())
    O = oracle.best()
    dem = oracle.fbm_terrain(300, 260, seed=61, quantum=0.5)
    dem[140:165, 60:130] = ND
    expected = {"fill": O.fill_depressions(dem)}
    expected["flats"] = O.resolve_flats(expected["fill"], ND)
    expected["fa_d8"] = O.fa_d8(expected["flats"], ND)
    expected["fa_dinf"] = O.fa_dinf(expected["flats"], ND)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lib_path, dem, expected, dict(params), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, res, err in results:
        assert err is None, f"rank {rank}: {err}"
        assert res == {"fill": True, "flats": True, "fa_d8": True, "fa_dinf": True}, (rank, res)
    assert all(p.exitcode == 0 for p in procs)
