''' ad-hoc: evals/s of the device path for one library build (RACELINE_B200_LIB) '''
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import build_product
sys.path.insert(0, ROOT)
import bench
from aircraft_trajectory_optimization_b200.functions import NlpFunctions, load_library
from aircraft_trajectory_optimization_b200.models import vehicle_params
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
case = sys.argv[2] if len(sys.argv) > 2 else bench.CASE
prod = build_product(case)
st = prod.structure
F = NlpFunctions(st, prod.vehicle_config, device=0)
X, L, VP = bench.make_inputs(st, vehicle_params(prod.vehicle_config), B, 0)
dev = torch.device('cuda', 0)
f64 = dict(dtype=torch.float64, device=dev)
x_d, l_d, vp_d = (torch.from_numpy(a).to(dev) for a in (X, L, VP))
sig = torch.ones(B, **f64); f_d = torch.empty(B, **f64); gf = torch.empty(B, st.nw, **f64); g = torch.empty(B, st.ng, **f64)
j = torch.empty(B, st.nnz_jac, **f64); h = torch.empty(B, st.nnz_hess, **f64)
step = lambda: F.eval_device(x_d, l_d, sig, vp_d, None, f_d, gf, g, j, h)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
ab = bench.algorithmic_bytes(st)
print(os.environ.get('RACELINE_B200_LIB', 'default'), case, f'B={B} {ms:.2f} ms/step {B / ms * 1e3:.0f} evals/s; {ab} B/eval -> {ab * B / ms / 1e6:.0f} GB/s algorithmic; checksum {float(h.sum()):.6e}')
