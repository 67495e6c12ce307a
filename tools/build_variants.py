"""Builds tuning variants of the library (different -D knobs) into gpurun_variants/*.so"""
import os, subprocess, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from richdem_b200 import build as B
variants = dict(a.split("=", 1) for a in sys.argv[1:])
out = os.path.join(ROOT, "variants"); os.makedirs(out, exist_ok=True)
for name, defs in variants.items():
    os.environ["RDB_DEFS"] = defs
    B.NVCC_FLAGS[:] = [f for f in B.NVCC_FLAGS if not f.startswith("-DRDB_")] + defs.split()
    B.build(force=True)
    shutil.copy(B.LIB, os.path.join(out, f"lib_{name}.so"))
    print("built", name, defs)
os.environ["RDB_DEFS"] = ""
B.NVCC_FLAGS[:] = [f for f in B.NVCC_FLAGS if not f.startswith("-DRDB_")]
B.build(force=True)
