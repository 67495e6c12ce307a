"""Timeline of the C++ band drivers (rdb200_mgpu_*; csrc/fill.cu mgpu_fill_band, csrc/accum.cu mgpu_fa_band).

    python tools/band_profile.py [N]                          one band = the whole raster, no communicator (1 GPU)
    torchrun --nproc-per-node G tools/band_profile.py [N]     G row bands over NCCL

Prints the per-phase trace of the last of three fills (fill_trace=1) and the wall times of fill / FA_D8 / FA_Dinf."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from richdem_b200 import _lib, sharded  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_lib.init(lr)
_lib.use_torch_stream()
L = _lib.lib()
r0, r1, gt, gb = sharded.local_rows(N, world, rank)
hloc = r1 - r0 + gt + gb
dem = torch.empty((hloc, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(dem.data_ptr(), N, hloc, r0 - gt, 42, 12, 0.0))


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


# further arguments: configurations of rdb200_set_param switches ("fill_vcycle=16 fill_band_multigrid=32"), each run in turn
variants = [a for a in sys.argv[2:] if a != "dinf"] or [""]
for variant in variants:
    _lib.reset_params()
    for kv in variant.split():
        _lib.set_param(kv.split("=")[0], int(kv.split("=")[1]))
    if rank == 0:
        print(f"== variant [{variant or 'defaults'}]", flush=True)
    for rep in range(3):
        if rep == 2:
            _lib.set_param("fill_trace", 1)
        w = dem.clone()
        sync()
        t = time.perf_counter()
        filled, xr, st = sharded.fill_band(w, gt, gb, row0=r0 - gt, height=N, return_stats=True)
        sync()
        tf = time.perf_counter() - t
        _lib.set_param("fill_trace", 0)
        t = time.perf_counter()
        acc, ar, st2 = sharded.fa_band(filled, gt, gb, -9999.0, dinf=False, return_stats=True)
        sync()
        ta = time.perf_counter() - t
        if rank == 0:
            print(f"rep {rep}: world={world} fill {tf * 1e3:.2f} ms ({xr} cycles, sweep {st['ms_main_kernel']:.2f} ms, "
                  f"{st['fill_rounds']} live rounds, {st['fill_tile_visits']} visits)  fa_d8 {ta * 1e3:.2f} ms ({ar} rounds)", flush=True)
_lib.reset_params()
if "dinf" in sys.argv[2:]:
    t = time.perf_counter()
    acc, ar = sharded.fa_band(filled, gt, gb, -9999.0, dinf=True)
    sync()
    if rank == 0:
        print(f"fa_dinf {(time.perf_counter() - t) * 1e3:.2f} ms ({ar} rounds)", flush=True)
if world > 1:
    dist.destroy_process_group()
