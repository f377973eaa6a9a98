"""Flake hunt: runs tests/test_gpu_mf.py::test_other_embedding_widths_vs_oracle in-process N times per case and
prints every assertion that fires. Usage: python scripts/repeat_mf_width_test.py [N]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_mf as T

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fails = 0
for d in (32, 256, 512, 1024):
    for name, wd in (("adam", 0.0), ("sgd", 1e-3)):
        for i in range(N):
            try:
                T.test_other_embedding_widths_vs_oracle(d, name, wd)
            except AssertionError as e:
                fails += 1
                tb = traceback.extract_tb(e.__traceback__)[-1]
                print(f"FAIL d={d} {name} wd={wd} rep={i}: line {tb.lineno} {tb.line} :: {str(e)[:200]}", flush=True)
        print(f"done d={d} {name} wd={wd}", flush=True)
print("total failures", fails)
