''' the C-ABI library loads and exports every symbol include/raceline_b200.h declares (no compute) '''
import os
import re


def test_header_symbols_exported(built_library):
    from aircraft_trajectory_optimization_b200.functions import load_library, EXPORTS
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, 'include', 'raceline_b200.h')).read()
    declared = set(re.findall(r'\b(rb_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    # the CasADi-shaped symbols are declared through RB_CASADI_DECLARE(NAME): expand the macro
    macro = hdr[hdr.index('#define RB_CASADI_DECLARE(NAME)'):hdr.index('RB_CASADI_DECLARE(nlp_f)')]
    suffixes = [''] + re.findall(r'NAME##(_[a-z_]+)\(', macro)
    names = re.findall(r'^RB_CASADI_DECLARE\((\w+)\)', hdr, re.M)
    assert len(names) == 5 and len(suffixes) == 16
    declared |= {n + sfx for n in names for sfx in suffixes}
    lib = load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert set(EXPORTS) <= declared


def test_casadi_shaped_metadata(built_library):
    ''' the generated-code style companions answer without a GPU: arity, names, work sizes '''
    import ctypes
    from aircraft_trajectory_optimization_b200.functions import load_library
    lib = load_library()
    want = {'nlp_f': (['x', 'p'], ['f']), 'nlp_g': (['x', 'p'], ['g']),
            'nlp_grad_f': (['x', 'p'], ['f', 'grad_f_x']), 'nlp_jac_g': (['x', 'p'], ['g', 'jac_g_x']),
            'nlp_hess_l': (['x', 'p', 'lam_f', 'lam_g'], ['triu_hess_gamma_x_x'])}
    for name, (ins, outs) in want.items():
        n_in, n_out = getattr(lib, name + '_n_in'), getattr(lib, name + '_n_out')
        n_in.restype = n_out.restype = ctypes.c_longlong
        assert n_in() == len(ins) and n_out() == len(outs)
        for kind, lst in (('_name_in', ins), ('_name_out', outs)):
            fn = getattr(lib, name + kind)
            fn.restype, fn.argtypes = ctypes.c_char_p, [ctypes.c_longlong]
            assert [fn(i).decode() for i in range(len(lst))] == lst
            assert fn(len(lst)) is None
        sz = [ctypes.c_longlong(-1) for _ in range(4)]
        assert getattr(lib, name + '_work')(*[ctypes.byref(z) for z in sz]) == 0
        assert [z.value for z in sz] == [len(ins), len(outs), 0, 0]
        # nothing bound: the call fails with a message instead of touching a device
        f = getattr(lib, name)
        f.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
        lib.rb_casadi_bind.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        assert lib.rb_casadi_bind(None, None) == 0
        assert f(None, None, None, None, 0) != 0
        assert b'rb_casadi_bind' in lib.rb_last_error()
        sp = getattr(lib, name + '_sparsity_out')
        sp.restype, sp.argtypes = ctypes.c_void_p, [ctypes.c_longlong]
        assert sp(0) is None


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
