"""Under torchrun: stage timings of the full sharded pipeline (fill -> flats -> FA_D8 -> FA_Dinf) on an
N x N fBm DEM.  Writes gpurun_out/mgpu_pipeline_<world>.json (rank 0)."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from richdem_b200 import _lib, sharded

MG = int(os.environ.get("RDB_BAND_MULTIGRID", "0"))  # k >= 2: multigrid start of the band fill (fill_band(..., multigrid=k))
VC = int(os.environ.get("RDB_BAND_VCYCLE", "0"))     # n > 0: coarse-grid correction after every n halo exchanges

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
_lib.init(lr); _lib.use_torch_stream()
L = _lib.lib()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ND = -9999.0
r0, r1, gt, gb = sharded.local_rows(N, world, rank)
hloc = r1 - r0 + gt + gb
dem0 = torch.empty((hloc, N), dtype=torch.float32, device="cuda")
_lib.check(L.rdb200_dev_generate_fbm_f32(dem0.data_ptr(), N, hloc, r0 - gt, 42, 12, 0.0))

def timed(fn):
    torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(); dist.barrier()
    return out, (time.perf_counter() - t) * 1e3

res = {}
for rep in range(2):   # second repetition is the reported one (workspace warm)
    work = dem0.clone()
    (filled, fr), res["fill_ms"] = timed(lambda: sharded.fill_band(work, gt, gb, multigrid=MG, row0=r0 - gt, height=N, vcycle=VC))
    it, res["resolve_flats_ms"] = timed(lambda: sharded.resolve_flats_band(filled, gt, gb, ND))
    _, res["halo_refresh_ms"] = timed(lambda: sharded.exchange_rows(filled, gt, gb))
    (a8, r8), res["fa_d8_ms"] = timed(lambda: sharded.fa_band(filled, gt, gb, ND, dinf=False))
    (ai, ri), res["fa_dinf_ms"] = timed(lambda: sharded.fa_band(filled, gt, gb, ND, dinf=True))
    res.update({"fill_exchange_rounds": fr, "flats_seam_iterations": it, "fa_d8_rounds": r8, "fa_dinf_rounds": ri})
mx = torch.tensor([float(a8[gt:gt + r1 - r0].max()), float(ai[gt:gt + r1 - r0].max())], dtype=torch.float64, device="cuda")
dist.all_reduce(mx, op=dist.ReduceOp.MAX)
if rank == 0:
    res.update({"N": N, "world": world, "fa_d8_max": float(mx[0]), "fa_dinf_max": float(mx[1]),
                "pipeline_fill_flats_d8_ms": res["fill_ms"] + res["resolve_flats_ms"] + res["halo_refresh_ms"] + res["fa_d8_ms"]})
    res["pipeline_Mcells_s"] = N * N / (res["pipeline_fill_flats_d8_ms"] * 1e-3) / 1e6
    print(json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/mgpu_pipeline_{world}.json", "w"), indent=1)
dist.destroy_process_group()
