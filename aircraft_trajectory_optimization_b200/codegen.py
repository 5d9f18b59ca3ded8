'''
Build-time code generator: vehicle ODE right-hand sides -> straight-line fp64 CUDA.

For every model variant (models.Variant) it differentiates the scalar expression graph of
f(x, fc, vp) (x = [z; u]) and emits one header `csrc/generated/pf_<variant>.cuh` with

  f    (x)                      -> f                       primal
  jvp  (x, dx)                  -> f, J dx                 forward tangent      (RK4 kernel)
  hvp  (x, dx, kb, dkb)         -> J'kb, d/de[J(x+e dx)'(kb+e dkb)]   second-order adjoint (RK4 kernel)
  fJW  (x, mu)                  -> f, nnz(J), nnz(triu sum_i mu_i d2f_i)   (collocation kernel)

plus `csrc/generated/pf_<variant>.json` with the structural patterns the host needs to lay out
jac_g / hess_l: the J and W patterns of f, and for the RK4 shooting cell
out = cont(rk4(z, u, h)) (reference base_raceline.py:363-391, :1052-1112; drone_raceline.py:42-45)
the dependency table of every output and of every output's Hessian on v = (z, u, h).

Common sub-expressions are shared automatically (hash-consed graph), constants are printed with
repr() so they round-trip exactly.
'''
import json
import os

from . import symbolic as sx
from .models import Variant, zdot, NFC

HERE = os.path.dirname(os.path.abspath(__file__))
GEN_DIR = os.path.join(HERE, 'csrc', 'generated')

ALL_VARIANTS = [Variant(v, o, f, d)
                for v, o in (('drone', 'quat'), ('drone', 'ypr'), ('point', 'quat'))
                for f in ('global', 'param_gr', 'param_lr')
                for d in (False, True)]

# variants compiled into the shared library (the five BASELINE configs use the first four)
COMPILED_VARIANTS = [
    Variant('drone', 'quat', 'global', False),
    Variant('drone', 'quat', 'param_gr', False),
    Variant('point', 'quat', 'global', False),
    Variant('point', 'quat', 'param_gr', False),
    Variant('drone', 'ypr', 'param_gr', False),
    Variant('drone', 'ypr', 'global', False),
    # frame-relative orientation (`global_r=False`, drone3d/dynamics/drone_models.py:249-292, point_model.py:149-213)
    Variant('drone', 'quat', 'param_lr', False),
    Variant('drone', 'ypr', 'param_lr', False),
    Variant('point', 'quat', 'param_lr', False),
    # linear drag b != 0 changes the sparsity pattern: the two BASELINE drone variants with drag
    Variant('drone', 'quat', 'global', True),
    Variant('drone', 'quat', 'param_gr', True),
]


def _syms(g, name, n):
    return [sx.SX(g.input(f'{name}{k}')) for k in range(n)]


class PointFunctionGraphs:
    ''' the symbolic graphs of one variant '''

    def __init__(self, variant: Variant):
        self.variant = variant
        nz, nu = variant.nz, variant.nu
        nx = nz + nu
        nvp = len(variant.vp_names)
        g = sx.new_graph()
        self.g = g
        self.x = _syms(g, 'x', nx)
        self.fc = _syms(g, 'fc', NFC)
        self.vp = _syms(g, 'vp', nvp)
        self.dx = _syms(g, 'dx', nx)
        self.kb = _syms(g, 'kb', nz)
        self.dkb = _syms(g, 'dkb', nz)
        f = zdot(variant, self.x[:nz], self.x[nz:], self.fc, self.vp)
        self.f = [sx._id(e) for e in f]
        xi = [e.i for e in self.x]

        # materialised Jacobian (column-major sorted nnz) ------------------------------------
        wrt = {n: c for c, n in enumerate(xi)}
        rows = g.forward_sparse(self.f, wrt)
        self.J = sorted(((r, c) for r, d in enumerate(rows) for c in d), key=lambda rc: (rc[1], rc[0]))
        self.J_nodes = [rows[r][c] for r, c in self.J]

        # W = sum_i kb_i d2 f_i / dx2, upper triangle ------------------------------------------
        adj = g.reverse(self.f, [e.i for e in self.kb])
        self.xb = [adj.get(n, g.zero) for n in xi]
        hrows = g.forward_sparse(self.xb, wrt)
        W = {}
        for i, d in enumerate(hrows):
            for j, v in d.items():
                if i <= j:
                    W[(i, j)] = v
                else:
                    W.setdefault((j, i), v)
        self.W = sorted(W, key=lambda rc: (rc[1], rc[0]))
        self.W_nodes = [W[k] for k in self.W]

        # directional derivatives --------------------------------------------------------------
        one_col = {n: 0 for n in xi}
        seeds = {n: d.i for n, d in zip(xi, self.dx)}
        self.df = [d.get(0, g.zero) for d in g.forward_sparse(self.f, one_col, seeds=seeds)]
        for n, d in zip(self.kb, self.dkb):
            one_col[n.i] = 0
            seeds[n.i] = d.i
        self.dxb = [d.get(0, g.zero) for d in g.forward_sparse(self.xb, one_col, seeds=seeds)]

    def input_exprs(self):
        ''' C expression for every INPUT slot of the graph, in creation order '''
        v = self.variant
        nx = v.nz + v.nu
        names = [f'x[{k}]' for k in range(nx)] + [f'fc[{k}]' for k in range(NFC)] \
            + [f'vp[{k}]' for k in range(len(v.vp_names))] + [f'dx[{k}]' for k in range(nx)] \
            + [f'kb[{k}]' for k in range(v.nz)] + [f'dkb[{k}]' for k in range(v.nz)]
        assert len(names) == len(self.g.input_names)
        return names


def rk4_cell_tables(variant: Variant):
    '''
    structural tables of the shooting cell out = cont(rk4(z, u, h)) over v = (z, u, h):
      J_dep[c][j]    out_c depends on v_j
      H_dep[c]       list of (i, j), i <= j, with d2 out_c / dv_i dv_j structurally non-zero
    '''
    nz, nu = variant.nz, variant.nu
    g = sx.new_graph()
    z = _syms(g, 'z', nz)
    u = _syms(g, 'u', nu)
    h = sx.SX(g.input('h'))
    fc = _syms(g, 'fc', NFC)
    vp = _syms(g, 'vp', len(variant.vp_names))

    def f(zz):
        return zdot(variant, zz, u, fc, vp)

    k1 = f(z)
    k2 = f([z[i] + h / 2 * k1[i] for i in range(nz)])
    k3 = f([z[i] + h / 2 * k2[i] for i in range(nz)])
    k4 = f([z[i] + h * k3[i] for i in range(nz)])
    zn = [z[i] + h / 6 * (k1[i] + k2[i] * 2 + k3[i] * 2 + k4[i]) for i in range(nz)]
    if variant.vehicle == 'drone' and variant.orient == 'quat':
        nrm = sx.norm_2(zn[3:7])
        zn[3:7] = [e / nrm for e in zn[3:7]]
    out = [sx._id(e) for e in zn]
    v = [e.i for e in (*z, *u, h)]
    nv = len(v)
    wrt = {n: c for c, n in enumerate(v)}
    dep = g.dependencies(out, wrt)
    J_dep = [[(d >> j) & 1 for j in range(nv)] for d in dep]
    H_dep = []
    for c in range(nz):
        adj = g.reverse([out[c]], [g.one])
        grads = [adj.get(n, g.zero) for n in v]
        hd = g.dependencies(grads, wrt)
        ent = set()
        for i, d in enumerate(hd):
            for j in range(nv):
                if (d >> j) & 1:
                    ent.add((min(i, j), max(i, j)))
        H_dep.append(sorted(ent))
    return J_dep, H_dep


def _emit_fn(pf: PointFunctionGraphs, name, args, outputs, out_exprs, prologue=''):
    code, n_ops = pf.g.emit_c(outputs, pf.input_exprs(), out_exprs, indent='    ')
    sig = ', '.join(args)
    return (f'  // {n_ops} arithmetic instructions\n'
            f'  __device__ __forceinline__ static void {name}({sig}) {{\n{prologue}{code}\n  }}\n'), n_ops


def generate_variant(variant: Variant, out_dir=GEN_DIR):
    ''' write pf_<variant>.cuh and pf_<variant>.json; returns the metadata dict '''
    os.makedirs(out_dir, exist_ok=True)
    pf = PointFunctionGraphs(variant)
    nz, nu = variant.nz, variant.nu
    nx = nz + nu
    name = variant.name
    cx, cfc, cvp = 'const double* __restrict__ x', 'const double* __restrict__ fc', \
        'const double* __restrict__ vp'
    parts = []
    ops = {}
    code, ops['f'] = _emit_fn(pf, 'f', [cx, cfc, cvp, 'double* __restrict__ f'], pf.f,
                              [f'f[{i}]' for i in range(nz)])
    parts.append(code)
    code, ops['jvp'] = _emit_fn(pf, 'jvp', [cx, 'const double* __restrict__ dx', cfc, cvp,
                                           'double* __restrict__ f', 'double* __restrict__ df'],
                                pf.f + pf.df,
                                [f'f[{i}]' for i in range(nz)] + [f'df[{i}]' for i in range(nz)])
    parts.append(code)
    code, ops['hvp'] = _emit_fn(pf, 'hvp', [cx, 'const double* __restrict__ dx',
                                           'const double* __restrict__ kb',
                                           'const double* __restrict__ dkb', cfc, cvp,
                                           'double* __restrict__ xb', 'double* __restrict__ dxb'],
                                pf.xb + pf.dxb,
                                [f'xb[{i}]' for i in range(nx)] + [f'dxb[{i}]' for i in range(nx)])
    parts.append(code)
    code, ops['vjp'] = _emit_fn(pf, 'vjp', [cx, 'const double* __restrict__ kb', cfc, cvp,
                                           'double* __restrict__ xb'],
                                pf.xb, [f'xb[{i}]' for i in range(nx)])
    parts.append(code)
    code, ops['fJW'] = _emit_fn(pf, 'fJW', [cx, 'const double* __restrict__ kb', cfc, cvp,
                                           'double* __restrict__ f', 'double* __restrict__ Jv',
                                           'double* __restrict__ Wv'],
                                pf.f + pf.J_nodes + pf.W_nodes,
                                [f'f[{i}]' for i in range(nz)]
                                + [f'Jv[{i}]' for i in range(len(pf.J))]
                                + [f'Wv[{i}]' for i in range(len(pf.W))])
    parts.append(code)
    code, ops['fJW_scatter'] = _emit_fn(
        pf, 'fJW_scatter', [cx, 'const double* __restrict__ kb', cfc, cvp, 'double* __restrict__ f',
                            'double* Jst', 'const int* __restrict__ tj', 'double* Hst',
                            'const int* __restrict__ th'],
        pf.f + pf.J_nodes + pf.W_nodes,
        [f'f[{i}]' for i in range(nz)] + [f'Jst[tj[{i}]]' for i in range(len(pf.J))]
        + [f'Hst[th[{i}]]' for i in range(len(pf.W))])
    parts.append(code)
    # ---- materialised form for the two-kernel shooting path (csrc/rk4_cells.cuh) ------------------
    # fJ_s / vjpW_s write the non-zeros of J and W straight into a strided scratch column: entries are stored in
    # pairs (2i, 2i+1) adjacent, pair i of a cell at double offset 2 * i * SCR_STRIDE, so that the direction
    # threads read two entries with one 16-byte shared-memory load (one load per two multiply-adds);
    # jmul / jtmul / wmul are the sparse products the direction threads run against those columns.
    code, ops['fJ_s'] = _emit_fn(pf, 'fJ_s', [cx, cfc, cvp, 'double* __restrict__ f', 'double* __restrict__ Js'],
                                 pf.f + pf.J_nodes,
                                 [f'f[{i}]' for i in range(nz)] + [f'Ja[{(i >> 1) * 2} * SCR_STRIDE + {i & 1}]' for i in range(len(pf.J))],
                                 prologue='    double* __restrict__ Ja = static_cast<double*>(__builtin_assume_aligned(Js, 16));\n')
    parts.append(code)
    code, ops['vjpW_s'] = _emit_fn(pf, 'vjpW_s', [cx, 'const double* __restrict__ kb', cfc, cvp,
                                                  'double* __restrict__ xb', 'double* __restrict__ Ws'],
                                   pf.xb + pf.W_nodes,
                                   [f'xb[{i}]' for i in range(nx)] + [f'Wa[{(i >> 1) * 2} * SCR_STRIDE + {i & 1}]' for i in range(len(pf.W))],
                                   prologue='    double* __restrict__ Wa = static_cast<double*>(__builtin_assume_aligned(Ws, 16));\n')
    parts.append(code)

    def sparse_product(fname, sig, lines_fn, nops, src, n_ent):
        loads = [f'const double2* __restrict__ P2 = reinterpret_cast<const double2*>({src});'] if n_ent else \
            [f'(void){src};']
        loads += [f'const double2 q{i} = P2[{i} * SCR_STRIDE];' for i in range((n_ent + 1) // 2)]
        body = '\n'.join('    ' + ln for ln in loads + lines_fn())
        return (f'  // {nops} multiply-adds\n  __device__ __forceinline__ static void {fname}({sig}) {{\n{body}\n  }}\n')

    def ent(e):
        return f'q{e >> 1}.{"xy"[e & 1]}'

    def jmul_lines():
        out = []
        for r in range(nz):
            terms = [f'{ent(e)} * dx[{c}]' for e, (rr, c) in enumerate(pf.J) if rr == r]
            out.append(f'dk[{r}] = ' + (' + '.join(terms) if terms else '0.0') + ';')
        return out

    def jtmul_lines():
        out = []
        for c in range(nx):
            terms = [f'{ent(e)} * dkb[{r}]' for e, (r, cc) in enumerate(pf.J) if cc == c]
            if terms:
                out.append(f'dxb[{c}] += ' + ' + '.join(terms) + ';')
        return out

    def wmul_lines():
        out = []
        for i in range(nx):
            terms = []
            for e, (r, c) in enumerate(pf.W):
                if r == i:
                    terms.append(f'{ent(e)} * dx[{c}]')
                elif c == i:
                    terms.append(f'{ent(e)} * dx[{r}]')
            out.append(f'dxb[{i}] = ' + (' + '.join(terms) if terms else '0.0') + ';')
        return out

    cJs, cWs = 'const double* __restrict__ Js', 'const double* __restrict__ Ws'
    parts.append(sparse_product('jmul', f'{cJs}, const double* __restrict__ dx, double* __restrict__ dk',
                                jmul_lines, len(pf.J), 'Js', len(pf.J)))
    parts.append(sparse_product('wmul', f'{cWs}, const double* __restrict__ dx, double* __restrict__ dxb',
                                wmul_lines, 2 * len(pf.W), 'Ws', len(pf.W)))
    parts.append(sparse_product('jtmul', f'{cJs}, const double* __restrict__ dkb, double* __restrict__ dxb',
                                jtmul_lines, len(pf.J), 'Js', len(pf.J)))

    def table(tname, vals):
        return f'  static constexpr signed char {tname}[{max(len(vals), 1)}] = {{' \
            + ', '.join(str(v) for v in (vals or [0])) + '};\n'

    hdr = (f'// GENERATED by aircraft_trajectory_optimization_b200/codegen.py -- do not edit.\n'
           f'// variant {name}: reference equations drone3d/dynamics/drone_models.py:47-123,249-292,\n'
           f'// point_model.py:28-75,149-213, rotations.py:44-102 (SURVEY.md App. B)\n'
           f'#pragma once\n\n'
           f'struct PF_{name} {{\n'
           f'  static constexpr int NZ = {nz}, NU = {nu}, NX = {nx}, NFC = {NFC}, '
           f'NVP = {len(variant.vp_names)};\n'
           f'  static constexpr int NJ = {len(pf.J)}, NW = {len(pf.W)};\n'
           f'  // shooting cells: CPB cells per 128-thread block of the direction kernel; the scratch of a block is\n'
           f'  // stored entry-major with the CPB cells of the block adjacent (stride of one entry = CPB doubles; J and W\n'
           f'  // entries in adjacent pairs, stride of one pair = 2 CPB doubles)\n'
           f'  static constexpr int CPB = 128 / ({nx} + 1), SCR_STRIDE = CPB;\n'
           f'  static constexpr bool QUAT = {"true" if variant.vehicle == "drone" and variant.orient == "quat" else "false"};\n'
           f'  static constexpr bool USES_FC = {"true" if variant.parametric else "false"};\n'
           + table('J_ROW', [r for r, _ in pf.J]) + table('J_COL', [c for _, c in pf.J])
           + table('W_ROW', [r for r, _ in pf.W]) + table('W_COL', [c for _, c in pf.W])
           + '\n' + '\n'.join(parts) + '};\n')
    with open(os.path.join(out_dir, f'pf_{name}.cuh'), 'w') as fh:
        fh.write(hdr)

    J_dep, H_dep = rk4_cell_tables(variant)
    meta = dict(name=name, nz=nz, nu=nu, nvp=len(variant.vp_names), quat=variant.vehicle == 'drone'
                and variant.orient == 'quat', J=pf.J, W=pf.W, ops=ops,
                rk4_J_dep=J_dep, rk4_H_dep=H_dep)
    with open(os.path.join(out_dir, f'pf_{name}.json'), 'w') as fh:
        json.dump(meta, fh)
    return meta


def generate_all(variants=None, out_dir=GEN_DIR, verbose=False):
    metas = {}
    for v in (variants or COMPILED_VARIANTS):
        metas[v.name] = generate_variant(v, out_dir)
        if verbose:
            m = metas[v.name]
            print(f'{v.name}: nnzJ={len(m["J"])} nnzW={len(m["W"])} ops={m["ops"]}')
    return metas


def load_meta(variant: Variant, out_dir=GEN_DIR):
    path = os.path.join(out_dir, f'pf_{variant.name}.json')
    if not os.path.exists(path):
        # generated files are build outputs (not tracked): regenerate on demand (a few seconds per variant)
        generate_variant(variant, out_dir)
    with open(path) as fh:
        return json.load(fh)


if __name__ == '__main__':
    generate_all(verbose=True)
