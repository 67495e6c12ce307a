"""Developer GPU check: every stage vs the CPU oracle at a few sizes + fill timings.
Run on the GPU box:  python tools/gpu_check.py [--big]   (writes gpurun_out/gpu_check.json)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import richdem_b200 as rd
from richdem_b200 import _lib

O = oracle.best()
print("oracle kind:", O.kind, flush=True)
res = {"oracle": O.kind, "cases": []}
ND = -9999.0

def cmp(name, a, b, case, tol=None):
    if tol is None:
        ok = bool(np.array_equal(a, b))
        bad = int((a != b).sum()) if not ok else 0
    else:
        denom = np.maximum(np.abs(b), 1e-30)
        rel = np.abs(a - b) / denom
        ok = bool((rel <= tol).all())
        bad = int((rel > tol).sum())
    case[name] = {"ok": ok, "bad": bad}
    print(f"   {name:22s} {'OK' if ok else 'MISMATCH'} bad={bad}", flush=True)
    return ok

def run_case(h, w, seed, q=None, nodata_patch=False, use_tma=1):
    _lib.set_param("fill_use_tma", use_tma)
    dem = oracle.fbm_terrain(h, w, seed=seed, quantum=q)
    if nodata_patch:
        dem[h // 4: h // 4 + h // 8, w // 3: w // 3 + w // 6] = ND
        dem[0: h // 10, 0: w // 10] = ND
    case = {"h": h, "w": w, "seed": seed, "q": q, "tma": use_tma}
    print(f"case {h}x{w} seed={seed} q={q} nd={nodata_patch} tma={use_tma}", flush=True)
    rdem = rd.rdarray(dem, no_data=ND)
    t = time.time(); f_gpu = rd.FillDepressions(rdem); tg = time.time() - t
    st = rd.stats(); case["fill_stats"] = st
    t = time.time(); f_ref = O.fill_depressions(dem); tc = time.time() - t
    case["fill_gpu_s"], case["fill_cpu_s"] = tg, tc
    cmp("fill", np.asarray(f_gpu), f_ref, case)
    print(f"     rounds={st['fill_rounds']} visits={st['fill_tile_visits']} iters={st['fill_tile_iters']} ms_kernel={st['ms_main_kernel']:.2f} ms_total={st['ms_total']:.2f} cpu={tc:.2f}s", flush=True)
    fr = rd.rdarray(f_ref, no_data=ND)
    m_g, l_g = rd.FlatMask(fr); m_r, l_r = O.flat_mask(f_ref, ND)
    cmp("flat_mask", m_g, m_r, case); cmp("labels!=0", l_g != 0, l_r != 0, case)
    r_gpu = rd.ResolveFlats(fr); st = rd.stats()
    r_ref = O.resolve_flats(f_ref, ND)
    cmp("resolve_flats", np.asarray(r_gpu), r_ref, case)
    print(f"     flats: levels={st['flat_bfs_levels']} raised={st['flat_cells_raised']} ms={st['ms_total']:.2f}", flush=True)
    rr = rd.rdarray(r_ref, no_data=ND)
    cmp("d8_dirs", np.asarray(rd.FlowDirectionsD8(rr)), O.d8_flow_directions(r_ref, ND), case)
    dirs = O.d8_flow_directions(r_ref, ND)
    cmp("d8_flow_accum", np.asarray(rd.D8FlowAccum(dirs)), O.d8_flow_accum(dirs), case)
    cmp("fm_d8", np.asarray(rd.FlowProportions(rr, "D8")), O.fm_d8(r_ref, ND), case)
    pg, pr = np.asarray(rd.FlowProportions(rr, "Dinf")), O.fm_dinf(r_ref, ND)
    case["fm_dinf_exact"] = bool(np.array_equal(pg, pr))
    ulp = np.abs(pg.view(np.int32).astype(np.int64) - pr.view(np.int32).astype(np.int64))
    case["fm_dinf_max_ulp"] = int(ulp.max()); case["fm_dinf_n_diff"] = int((ulp > 0).sum())
    print(f"   fm_dinf exact={case['fm_dinf_exact']} max_ulp={case['fm_dinf_max_ulp']} ndiff={case['fm_dinf_n_diff']}", flush=True)
    a_g = rd.FlowAccumulation(rr, "D8"); st = rd.stats()
    cmp("fa_d8", np.asarray(a_g), O.fa_d8(r_ref, ND), case)
    print(f"     fa_d8 ms_total={st['ms_total']:.2f} ms_kernel={st['ms_main_kernel']:.2f}", flush=True)
    cmp("fa_dinf(1e-9)", np.asarray(rd.FlowAccumulation(rr, "Dinf")), O.fa_dinf(r_ref, ND), case, tol=1e-9)
    wts = np.random.default_rng(seed).random((h, w))
    cmp("fa_d8 weights", np.asarray(rd.FlowAccumulation(rr, "D8", weights=rd.rdarray(wts, no_data=-1))),
        O.fa_d8(r_ref, ND, wts), case, tol=1e-9)
    cmp("facc_props(dinf)", np.asarray(rd.FlowAccumFromProps(rd.rd3array(pr, no_data=-2))),
        O.flow_accumulation(pr), case, tol=1e-9)
    cmp("facc_props(d8)", np.asarray(rd.FlowAccumFromProps(rd.rd3array(O.fm_d8(r_ref, ND), no_data=-2))),
        O.flow_accumulation(O.fm_d8(r_ref, ND)), case)
    res["cases"].append(case)

try:
    run_case(150, 220, 1, use_tma=0)
    run_case(150, 220, 1)
    run_case(257, 131, 2, q=0.5)
    run_case(400, 500, 3, q=2.0, nodata_patch=True)
    run_case(1024, 1024, 5)
    run_case(2048, 3000, 6, q=1.0)
    if "--big" in sys.argv:
        run_case(4096, 4096, 7)
except Exception as e:
    import traceback; traceback.print_exc()
    res["error"] = repr(e)

# fill timings on device-generated terrain (no oracle): scaling of rounds / time with size
import torch
for N in ([2048, 4096, 8192, 16384] + ([32768] if "--huge" in sys.argv else [])):
    try:
        d = torch.empty((N, N), dtype=torch.float32, device="cuda")
        L = _lib.lib()
        _lib.check(L.rdb200_dev_generate_fbm_f32(d.data_ptr(), N, N, 0, 42, 12, 0.0))
        torch.cuda.synchronize()
        t = time.time()
        _lib.check(L.rdb200_dev_fill_depressions_d8_f32(d.data_ptr(), N, N))
        dt = time.time() - t
        st = rd.stats()
        acc = torch.empty((N, N), dtype=torch.float64, device="cuda")
        t = time.time()
        _lib.check(L.rdb200_dev_fa_d8_f32_f64(d.data_ptr(), acc.data_ptr(), N, N, ND, 1))
        dta = time.time() - t
        sta = rd.stats()
        row = {"N": N, "fill_s": dt, "fill_Mcells_s": N * N / dt / 1e6, "fill_stats": st, "fa_d8_s": dta,
               "fa_d8_Mcells_s": N * N / dta / 1e6, "fa_stats": sta, "acc_max": float(acc.max())}
        print(json.dumps(row), flush=True)
        res.setdefault("timings", []).append(row)
        del d, acc
    except Exception as e:
        import traceback; traceback.print_exc()
        res.setdefault("timing_errors", []).append(repr(e))
        break

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gpu_check.json", "w"), indent=1, default=str)
print("DONE", flush=True)
