'''
Oracle: the five NLP functions CasADi's nlpsol would create for a restated raceline problem.
TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

[third party, not in the reference tree] `ca.nlpsol('solver','ipopt',{x,f,g})`
(drone3d/raceline/base_raceline.py:799) derives, by source-transformation AD of the SX graph,
    nlp_f(x)            -> f
    nlp_g(x)            -> g
    nlp_grad_f(x)       -> f, grad_f
    nlp_jac_g(x)        -> g, jac_g                 (CCS, ng x nw, rows ascending in each column)
    nlp_hess_l(x, lam_f, lam_g) -> triu( hess( lam_f f + lam_g' g ) )   (CCS, nw x nw)
with sparsity found by structural dependency propagation.  This module restates that on the
shared node store: forward sparse sweeps per g block for jac_g; per block, a reverse sweep of
lam' g_block followed by a forward sparse sweep of the adjoints for hess_l (the Lagrangian is a
sum over blocks, so its Hessian and its structural pattern are the sum / union over blocks).
'''
import ctypes
import os

import numpy as np

from aircraft_trajectory_optimization_b200 import symbolic as sx

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'libsxvm.so')
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(['make', '-C', _HERE, '-s'])
        lib = ctypes.CDLL(path)
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.sxvm_allocate.argtypes = [ctypes.c_int, ip, ip, ip, ctypes.POINTER(ctypes.c_ubyte), ip]
        lib.sxvm_allocate.restype = ctypes.c_int
        lib.sxvm_eval.argtypes = [ctypes.c_int, ip, ip, ip, ip, dp, dp, dp, ctypes.c_int, ip, dp]
        lib.sxvm_eval.restype = None
        lib.sxvm_eval_batch.argtypes = [ctypes.c_int, ip, ip, ip, ip, dp, ctypes.c_int, dp, ctypes.c_int,
                                        ctypes.c_int, ip, dp, ctypes.c_int, ctypes.c_int]
        lib.sxvm_eval_batch.restype = None
        lib.sxvm_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def max_threads():
    return int(_lib().sxvm_max_threads())


class Tape:
    ''' a flat instruction tape for a list of output nodes, run by oracle/sxvm.c '''

    def __init__(self, g: sx.Graph, outputs, n_in):
        nodes = g.reachable(outputs)
        n = len(nodes)
        pos = {node: i for i, node in enumerate(nodes)}
        op = np.fromiter((g.op[v] for v in nodes), dtype=np.int32, count=n)
        a = np.zeros(n, dtype=np.int32)
        b = np.zeros(n, dtype=np.int32)
        consts = []
        ga, gb, gop = g.a, g.b, g.op
        for i, v in enumerate(nodes):
            o = gop[v]
            if o == sx.CONST:
                a[i] = len(consts)
                consts.append(g.cval[v])
            elif o == sx.INPUT:
                a[i] = ga[v]
            else:
                a[i] = pos[ga[v]]
                if o <= sx.DIV:
                    b[i] = pos[gb[v]]
        keep = np.zeros(n, dtype=np.uint8)
        out_pos = np.fromiter((pos[o] for o in outputs), dtype=np.int64, count=len(outputs))
        keep[out_pos] = 1
        dst = np.zeros(n, dtype=np.int32)
        lib = _lib()
        ip = ctypes.POINTER(ctypes.c_int)
        self.nslots = lib.sxvm_allocate(n, op.ctypes.data_as(ip), a.ctypes.data_as(ip), b.ctypes.data_as(ip),
                                        keep.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                        dst.ctypes.data_as(ip))
        self.n, self.op, self.a, self.b, self.dst = n, op, a, b, dst
        self.consts = np.array(consts if consts else [0.0], dtype=np.float64)
        self.out_slot = dst[out_pos].astype(np.int32)
        self.n_in, self.n_out = n_in, len(outputs)
        self._w = np.zeros(max(self.nslots, 1))

    def _args(self):
        ip = ctypes.POINTER(ctypes.c_int)
        dp = ctypes.POINTER(ctypes.c_double)
        return (self.n, self.op.ctypes.data_as(ip), self.a.ctypes.data_as(ip), self.b.ctypes.data_as(ip),
                self.dst.ctypes.data_as(ip), self.consts.ctypes.data_as(dp))

    def __call__(self, inputs):
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        assert inputs.shape == (self.n_in,)
        out = np.empty(self.n_out)
        _lib().sxvm_eval(*self._args(), inputs.ctypes.data_as(dp), self._w.ctypes.data_as(dp),
                         self.n_out, self.out_slot.ctypes.data_as(ip), out.ctypes.data_as(dp))
        return out

    def batch(self, inputs, nthreads=0):
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        inputs = np.ascontiguousarray(inputs, dtype=np.float64)
        B = inputs.shape[0]
        assert inputs.shape == (B, self.n_in)
        out = np.empty((B, self.n_out))
        _lib().sxvm_eval_batch(*self._args(), self.n_in, inputs.ctypes.data_as(dp), self.nslots, self.n_out,
                               self.out_slot.ctypes.data_as(ip), out.ctypes.data_as(dp), B, nthreads)
        return out


def _ccs(entries, ncol):
    ''' entries: iterable of (row, col) -> (colind, row, order) sorted column-major, rows ascending '''
    ent = sorted(entries, key=lambda rc: (rc[1], rc[0]))
    colind = np.zeros(ncol + 1, dtype=np.int64)
    for _, c in ent:
        colind[c + 1] += 1
    colind = np.cumsum(colind)
    row = np.array([r for r, _ in ent], dtype=np.int64)
    return colind, row, ent


class OracleNLP:
    '''
    nlp_f / nlp_g / nlp_grad_f / nlp_jac_g / nlp_hess_l of a RefRaceline, CasADi conventions.
    Inputs of every tape: concat(x[nw], lam_g[ng], lam_f[1]).
    '''

    def __init__(self, prob, build_hess=True):
        g = prob.graph
        sx.set_graph(g)
        self.prob = prob
        self.nw = len(prob.w)
        self.ng = len(prob.g_flat)
        x_nodes = [e.i for e in prob.w]
        assert [g.a[i] for i in x_nodes] == list(range(self.nw)), 'x must own input slots 0..nw-1'
        wrt = {node: c for c, node in enumerate(x_nodes)}
        lam = [g.input(f'lam_g_{r}') for r in range(self.ng)]
        lam_f = g.input('lam_f')
        self.n_in = self.nw + self.ng + 1
        assert g.a[lam_f] == self.n_in - 1

        f_node = prob.J.i
        g_nodes = [sx._id(e) for e in prob.g_flat]

        # ---- jac_g ------------------------------------------------------------------------
        jac = {}
        row0 = 0
        for blk in prob.g:
            outs = [sx._id(e) for e in blk]
            nodes = g.reachable(outs)
            for r, d in enumerate(g.forward_sparse(outs, wrt, nodes)):
                for c, v in d.items():
                    jac[(row0 + r, c)] = v
            row0 += len(outs)
        self.jac_colind, self.jac_row, ent = _ccs(jac.keys(), self.nw)
        jac_nodes = [jac[e] for e in ent]

        # ---- grad_f -----------------------------------------------------------------------
        adj = g.reverse([f_node], [g.one])
        grad_nodes = [adj.get(n, g.zero) for n in x_nodes]

        self.t_f = Tape(g, [f_node], self.n_in)
        self.t_g = Tape(g, g_nodes, self.n_in)
        self.t_grad_f = Tape(g, [f_node] + grad_nodes, self.n_in)
        self.t_jac_g = Tape(g, g_nodes + jac_nodes, self.n_in)

        # ---- hess_l -----------------------------------------------------------------------
        self.t_hess_l = None
        if build_hess:
            hess = {}

            def add_block(outs, seeds):
                nodes = g.reachable(outs)
                badj = g.reverse(outs, seeds, nodes)
                vars_here = [n for n in nodes if g.op[n] == sx.INPUT and n in wrt]
                grads = [badj.get(n, g.zero) for n in vars_here]
                loc = {}
                for n, d in zip(vars_here, g.forward_sparse(grads, wrt)):
                    i = wrt[n]
                    for j, v in d.items():
                        if i <= j:
                            loc[(i, j)] = v           # upper-triangle copy wins
                        else:
                            loc.setdefault((j, i), v)  # mirrored copy only if the upper one is absent
                for key, v in loc.items():
                    hess.setdefault(key, []).append(v)

            add_block([f_node], [lam_f])
            row0 = 0
            for blk in prob.g:
                outs = [sx._id(e) for e in blk]
                add_block(outs, lam[row0:row0 + len(outs)])
                row0 += len(outs)
            final = {key: g.sum(lst) for key, lst in hess.items()}
            self.hess_colind, self.hess_row, ent = _ccs(final.keys(), self.nw)
            self.t_hess_l = Tape(g, [final[e] for e in ent], self.n_in)

    # ---- CasADi-shaped calls --------------------------------------------------------------
    def _in(self, x, lam_g=None, lam_f=1.0):
        v = np.zeros(self.n_in)
        v[:self.nw] = x
        if lam_g is not None:
            v[self.nw:self.nw + self.ng] = lam_g
        v[-1] = lam_f
        return v

    def nlp_f(self, x):
        return float(self.t_f(self._in(x))[0])

    def nlp_g(self, x):
        return self.t_g(self._in(x))

    def nlp_grad_f(self, x):
        out = self.t_grad_f(self._in(x))
        return float(out[0]), out[1:]

    def nlp_jac_g(self, x):
        out = self.t_jac_g(self._in(x))
        return out[:self.ng], out[self.ng:]

    def nlp_hess_l(self, x, lam_f, lam_g):
        return self.t_hess_l(self._in(x, lam_g, lam_f))

    @property
    def nnz_jac(self):
        return len(self.jac_row)

    @property
    def nnz_hess(self):
        return len(self.hess_row)

    def jac_csc(self, vals):
        import scipy.sparse as sp
        return sp.csc_matrix((vals, self.jac_row, self.jac_colind), shape=(self.ng, self.nw))

    def hess_csc(self, vals):
        import scipy.sparse as sp
        return sp.csc_matrix((vals, self.hess_row, self.hess_colind), shape=(self.nw, self.nw))
