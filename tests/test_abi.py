''' the C-ABI library loads and exports every symbol include/raceline_b200.h declares (no compute) '''
import os
import re


def test_header_symbols_exported(built_library):
    from aircraft_trajectory_optimization_b200.functions import load_library, EXPORTS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, 'include', 'raceline_b200.h')).read()
    declared = set(re.findall(r'\b(rb_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    lib = load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(EXPORTS) <= declared


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'aircraft_trajectory_optimization_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), fn


def test_no_gpu_fails_loudly(built_library):
    ''' on a box without CUDA the function object refuses to exist (no CPU fallback) '''
    import ctypes
    import pytest
    from aircraft_trajectory_optimization_b200.functions import load_library
    lib = load_library()
    cnt = ctypes.c_int(0)
    rc = lib.rb_device_count(ctypes.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip('a GPU is present')
    from cases import build_product
    prod = build_product('race_param_rk4_point', small=True)
    with pytest.raises(RuntimeError):
        prod.functions


def test_solver_fails_loudly_without_gpu(built_library):
    ''' the nlpsol-shaped solver has no CPU path either '''
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from cases import build_product
    prod = build_product('race_param_rk4_point', small=True)
    with pytest.raises(RuntimeError):
        prod.solve()


def test_solve_util_matches_reference_signature():
    ''' drone3d/utils/solve_util.py:11-24: argument names and defaults '''
    import inspect
    from aircraft_trajectory_optimization_b200.solve_util import solve_util
    sig = inspect.signature(solve_util)
    assert list(sig.parameters) == ['line', 'global_frame', 'drone', 'use_quaternion', 'global_r', 'use_ws', 'solve',
                                    'fix_gate_center', 'verbose', 'use_rk4', 'N', 'v0']
    d = {k: v.default for k, v in sig.parameters.items() if v.default is not inspect.Parameter.empty}
    assert d == dict(use_quaternion=False, global_r=True, use_ws=False, solve=True, fix_gate_center=False,
                     verbose=True, use_rk4=False, N=50, v0=1.0)
    from cases import make_line
    solver, rl = solve_util(make_line('fig8'), global_frame=False, drone=False, solve=False, verbose=False, N=8)
    assert type(solver).__name__ == 'ParametricPointRaceline' and len(rl.states) == 8 * 8
