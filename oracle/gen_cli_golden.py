"""Generate tests/golden/cli/: run the REFERENCE's own CLI mains (construct_boundary_alpha.py, boundary_curvature_localpoly.py;
matplotlib replaced by a no-op stub) on small inputs taken from reference_vectors.npz and keep their CSV / TXT outputs.
Build container only (reads /root/reference); nothing of the reference's code is copied, only its outputs.

    python oracle/gen_cli_golden.py
"""
import sys, types, runpy, os, shutil
import numpy as np

class _Noop(types.ModuleType):
    def __getattr__(self, name):
        def f(*a, **k):
            return _Noop("x")
        return f
    def __call__(self, *a, **k):
        return _Noop("x")
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = _Noop(name)

sys.modules["matplotlib"].__dict__["pyplot"] = sys.modules["matplotlib.pyplot"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(ROOT + "/tests/golden/reference_vectors.npz")
out = ROOT + "/tests/golden/cli"
os.makedirs(out, exist_ok=True)
work = "/tmp/lm_cli_golden_work"; shutil.rmtree(work, ignore_errors=True); os.makedirs(work + "/outputs")
np.savetxt(work + "/construct_points.csv", g["alpha_points"], delimiter=",")                       # headerless, like construct_stage1_clean.py:178
sys.argv = ["construct_boundary_alpha.py", "--input_csv", work + "/construct_points.csv", "--alpha", "6.0", "--output_prefix", work + "/outputs/construct"]
runpy.run_path("/root/reference/construct_boundary_alpha.py", run_name="__main__")
np.savetxt(work + "/loop_boundary.csv", g["curv_closed_P"], delimiter=",", header="x,y", comments="")  # the format <prefix>_boundary.csv has
sys.argv = ["boundary_curvature_localpoly.py", "--input_csv", work + "/loop_boundary.csv", "--output_prefix", work + "/outputs/loop", "--neighbors", "7"]
runpy.run_path("/root/reference/boundary_curvature_localpoly.py", run_name="__main__")
for f in ("construct_points.csv", "loop_boundary.csv"):
    shutil.copy(work + "/" + f, out + "/" + f)
for f in sorted(os.listdir(work + "/outputs")):
    if f.endswith((".csv", ".txt")):
        shutil.copy(work + "/outputs/" + f, out + "/" + f)
# README step 2 as documented: construct_boundary_alpha_spyder_v2.py is a parameter-block script; run it with its three
# parameters pointed at the fixture (alpha 12 gives one outer loop and eight holes on this cloud)
import re
src = open("/root/reference/construct_boundary_alpha_spyder_v2.py").read()
src = re.sub(r'^input_csv\s*=.*$', f'input_csv = "{out}/construct_points.csv"', src, flags=re.M)
src = re.sub(r'^alpha\s*=.*$', 'alpha = 12.0', src, flags=re.M)
src = re.sub(r'^output_prefix\s*=.*$', f'output_prefix = "{work}/outputs/construct_v2"', src, flags=re.M)
ns = {"__name__": "v2"}
exec(compile(src, "construct_boundary_alpha_spyder_v2.py", "exec"), ns)
for f in ("construct_v2_boundary.csv", "construct_v2_edges.csv", "construct_v2_meta.txt"):
    shutil.copy(work + "/outputs/" + f, out + "/" + f)
np.save(out + "/construct_v2_ordered_idx.npy", np.asarray(ns["ordered_idx"], dtype=np.int32))
print(sorted(os.listdir(out)))
