'''
Structured problem descriptor: what the builders produce instead of the reference's SX graph.

The reference appends symbolic rows to `nlp['g']` in a fixed order and lets CasADi discover the
sparsity of jac_g / hess_l (drone3d/raceline/base_raceline.py:232-239, :625-646, :799).  Here the
builders (raceline.py) append *row records* in the same order to a `StructureBuilder`:

  * interval cells  -- rows produced by the dynamics of one interval (generated code on the GPU);
  * simple rows     -- affine rows and sums of squared affine forms (pure data).

`finalize()` then lays out jac_g (CCS, ng x nw) and hess_l (upper-triangular CCS) index-for-index
the way CasADi would, and builds the int32 slot tables that let every kernel
thread write its results directly at their CCS positions.

Decision vector (base_raceline.py:681-713): w = [H[0..N-1]; for n, k: Z[n,k], U[n,k], dU[n,k]].
'''
from dataclasses import dataclass, field
from typing import List

import numpy as np

RK4, COLLOC = 0, 1


@dataclass
class SimpleRow:
    row: int
    kind: int                 # 0 affine, 1 sum of squares
    vars: np.ndarray          # (nv,) indices into w
    A: np.ndarray             # (M, nv)
    c: np.ndarray             # (M,)
    scale_vp: int = -1        # -1: scale 1, k: scale = 1 / vp[k]^2
    mask: np.ndarray = None   # (M, nv) structural non-zeros of A


class StructureBuilder:
    def __init__(self, transcription, meta, N, K, R, dR, fc=None):
        self.transcription = transcription
        self.meta = meta
        self.variant = meta['name']
        self.N, self.K = N, K
        self.nz, self.nu = meta['nz'], meta['nu']
        self.nx = self.nz + self.nu
        self.S = self.nz + 2 * self.nu
        self.P = K + 1
        self.nw = N + N * self.P * self.S
        self.R = np.asarray(R, dtype=float)
        self.dR = np.asarray(dR, dtype=float)
        self.fc = None if fc is None else np.ascontiguousarray(fc, dtype=float).reshape(N * self.P, 13)
        self.lbg: List[float] = []
        self.ubg: List[float] = []
        self.simple: List[SimpleRow] = []
        self.tail = None            # tail.TailRows: end rows of open tracks (expression rows, evaluated from a tape)
        if transcription == RK4:
            nr = self.nx
            self.cell_row = -np.ones((N, nr), dtype=np.int32)
            self.cell_coef = np.zeros((N, nr))
            self.cell_partner = -np.ones((N, nr), dtype=np.int32)
            self.cell_pcoef = np.zeros((N, nr))
            self.cell_off = np.zeros((N, nr))
            self.cell_par = np.full((N, 1), 0.5)
        else:
            from .structure_colloc import init_colloc_cells
            init_colloc_cells(self)

    # ---- indices into w ---------------------------------------------------------------------
    def iH(self, n):
        return n

    def base(self, n, k=0):
        return self.N + (n * self.P + k) * self.S

    def iZ(self, n, k, i):
        return self.base(n, k) + i

    def iU(self, n, k, j):
        return self.base(n, k) + self.nz + j

    def idU(self, n, k, j):
        return self.base(n, k) + self.nx + j

    # ---- row allocation -----------------------------------------------------------------------
    @property
    def ng(self):
        return len(self.lbg)

    def alloc_rows(self, n, lb, ub):
        r0 = len(self.lbg)
        self.lbg += list(lb) if np.ndim(lb) else [lb] * n
        self.ubg += list(ub) if np.ndim(ub) else [ub] * n
        assert len(self.lbg) == r0 + n == len(self.ubg)
        return np.arange(r0, r0 + n)

    def add_affine(self, vars_, coefs, const, lb, ub, mask=None):
        '''
        one row  sum_i coefs[i] * w[vars[i]] + const.  mask: which entries the reference's expression
        structurally keeps (default: non-zero coefficients -- SX folds 0*x away at construction)
        '''
        vars_ = np.asarray(vars_, dtype=np.int64)
        coefs = np.asarray(coefs, dtype=float)
        keep = (coefs != 0) if mask is None else np.asarray(mask, dtype=bool)
        r = int(self.alloc_rows(1, lb, ub)[0])
        self.simple.append(SimpleRow(r, 0, vars_[keep], coefs[keep][None, :], np.array([float(const)]),
                                     mask=np.ones((1, int(keep.sum())), dtype=bool)))
        return r

    def add_squares(self, vars_, A, c, lb, ub, scale_vp=-1, mask=None):
        ''' one row  scale * sum_m (A[m] . w[vars] + c[m])^2 ; mask (M, nv): structural entries of A '''
        vars_ = np.asarray(vars_, dtype=np.int64)
        A = np.atleast_2d(np.asarray(A, dtype=float))
        mask = (A != 0) if mask is None else np.atleast_2d(np.asarray(mask, dtype=bool))
        keep = mask.any(axis=0)
        r = int(self.alloc_rows(1, lb, ub)[0])
        self.simple.append(SimpleRow(r, 1, vars_[keep], np.where(mask, A, 0.0)[:, keep],
                                     np.asarray(c, dtype=float), scale_vp, mask=mask[:, keep]))
        return r

    # ---- finalisation ---------------------------------------------------------------------------
    def finalize(self):
        if self.transcription == RK4:
            jent, hent = self._rk4_entries()
        else:
            from .structure_colloc import colloc_entries
            jent, hent = colloc_entries(self)
        return _assemble(self, jent, hent)

    def _rk4_entries(self):
        '''
        local slot layout of the RK4 cell (must match csrc/rk4_cells.cuh):
          jac : [c*NV + j] d(row of out_c)/d v_j ; [NZ*NV + c] partner of out row c ;
                [NZ*NV + NZ + 4*j + (0 u_j | 1 du_j | 2 h | 3 partner)] input row j
          hess: [i*NL + j], i <= j over local order (z, u, h, du)
        returns (rows, cols, cell, slot) arrays for jac and (r, c, cell, slot) for hess
        '''
        nz, nu, nx, N = self.nz, self.nu, self.nx, self.N
        NV, NL = nx + 1, nx + 1 + nu
        J_dep = np.array(self.meta['rk4_J_dep'], dtype=bool)            # (nz, NV)
        H_dep = [np.array(h, dtype=np.int64).reshape(-1, 2) for h in self.meta['rk4_H_dep']]
        self.cell_nj = nz * NV + nz + 4 * nu
        self.cell_nh = NL * NL
        jr, jc, jn, js = [], [], [], []
        hr, hc, hn, hs = [], [], [], []
        for n in range(N):
            lvar = np.array([*(self.base(n) + np.arange(nx)), self.iH(n),
                             *(self.base(n) + nx + np.arange(nu))], dtype=np.int64)
            hloc = np.zeros((NL, NL), dtype=bool)
            for c in range(nz):
                r = self.cell_row[n, c]
                if r < 0:
                    continue
                for j in np.nonzero(J_dep[c])[0]:
                    jr.append(r), jc.append(lvar[j]), jn.append(n), js.append(c * NV + j)
                if self.cell_partner[n, c] >= 0:
                    jr.append(r), jc.append(self.cell_partner[n, c]), jn.append(n), js.append(nz * NV + c)
                hloc[H_dep[c][:, 0], H_dep[c][:, 1]] = True
            for j in range(nu):
                r = self.cell_row[n, nz + j]
                if r >= 0:
                    b0 = nz * NV + nz + 4 * j
                    for k, col in enumerate((lvar[nz + j], lvar[NV + j], lvar[nx])):
                        jr.append(r), jc.append(col), jn.append(n), js.append(b0 + k)
                    if self.cell_partner[n, nz + j] >= 0:
                        jr.append(r), jc.append(self.cell_partner[n, nz + j]), jn.append(n), js.append(b0 + 3)
                    hloc[nx, NV + j] = True
                if self.R[j] != 0:
                    hloc[nz + j, nz + j] = True
                    hloc[nz + j, nx] = True
                if self.dR[j] != 0:
                    hloc[NV + j, NV + j] = True
                    hloc[nx, NV + j] = True
            ii, jj = np.nonzero(np.triu(hloc))
            for i, j in zip(ii, jj):
                a, b = lvar[i], lvar[j]
                hr.append(min(a, b)), hc.append(max(a, b)), hn.append(n), hs.append(i * NL + j)
        A = lambda v: np.array(v, dtype=np.int64)
        return (A(jr), A(jc), A(jn), A(js)), (A(hr), A(hc), A(hn), A(hs))


@dataclass
class NLPStructure:
    ''' everything the C ABI needs (rb_problem_desc) plus bounds '''
    transcription: int
    variant: str
    N: int
    K: int
    nz: int
    nu: int
    nw: int
    ng: int
    R: np.ndarray
    dR: np.ndarray
    fc: np.ndarray
    lbg: np.ndarray
    ubg: np.ndarray
    cell: dict
    srow: dict
    shess: dict
    jac_colind: np.ndarray
    jac_row: np.ndarray
    hess_colind: np.ndarray
    hess_row: np.ndarray
    lbw: np.ndarray = None
    ubw: np.ndarray = None
    w0: np.ndarray = None
    tail: dict = None       # open tracks: tape of the end rows (tail.py), None otherwise

    @property
    def nnz_jac(self):
        return len(self.jac_row)

    @property
    def nnz_hess(self):
        return len(self.hess_row)


def _ccs(rows, cols, nrow, ncol):
    ''' unique (row, col) pairs -> (colind, row, key array sorted) with keys col*nrow+row '''
    keys = np.unique(cols.astype(np.int64) * nrow + rows.astype(np.int64))
    c = keys // nrow
    r = keys - c * nrow
    colind = np.zeros(ncol + 1, dtype=np.int64)
    np.add.at(colind, c + 1, 1)
    return np.cumsum(colind), r, keys


def _assemble(sb: StructureBuilder, jent, hent) -> NLPStructure:
    nw, ng = sb.nw, sb.ng
    jr, jc, jn, js = jent
    hr, hc, hn, hs = hent

    # ---- simple rows: Jacobian entries and Hessian pairs ----------------------------------------
    s_rows = sb.simple
    sj_r = np.concatenate([np.full(len(s.vars), s.row, dtype=np.int64) for s in s_rows]) \
        if s_rows else np.zeros(0, dtype=np.int64)
    sj_c = np.concatenate([s.vars for s in s_rows]) if s_rows else np.zeros(0, dtype=np.int64)
    sh = {}   # (r, c) -> list of (row, coef, scale)
    for s in s_rows:
        if s.kind != 1:
            continue
        nzm = s.mask
        for a in range(len(s.vars)):
            for b in range(a, len(s.vars)):
                both = nzm[:, a] & nzm[:, b]
                if not both.any():
                    continue
                va, vb = int(s.vars[a]), int(s.vars[b])
                key = (min(va, vb), max(va, vb))
                coef = 2.0 * float(np.sum(s.A[both, a] * s.A[both, b]))
                sh.setdefault(key, []).append((s.row, coef, s.scale_vp))
    sh_keys = sorted(sh)
    sh_r = np.array([k[0] for k in sh_keys], dtype=np.int64)
    sh_c = np.array([k[1] for k in sh_keys], dtype=np.int64)

    # ---- expression rows of open tracks (tail.py) ---------------------------------------------------
    z64 = np.zeros(0, dtype=np.int64)
    (tj_r, tj_c), (th_r, th_c) = sb.tail.entries() if sb.tail is not None else ((z64, z64), (z64, z64))

    # ---- jac_g ------------------------------------------------------------------------------------
    all_r = np.concatenate([jr, sj_r, tj_r])
    all_c = np.concatenate([jc, sj_c, tj_c])
    assert len(np.unique(all_c * ng + all_r)) == len(all_r), 'duplicate Jacobian entry'
    jac_colind, jac_row, jkeys = _ccs(all_r, all_c, ng, nw)
    pos = np.searchsorted(jkeys, all_c * ng + all_r)
    cell_jslot = -np.ones((sb.N, sb.cell_nj), dtype=np.int32)
    cell_jslot[jn, js] = pos[:len(jr)]
    srow_jslot = pos[len(jr):len(jr) + len(sj_r)].astype(np.int32)
    tail_jslot = pos[len(jr) + len(sj_r):]

    # ---- hess_l -------------------------------------------------------------------------------------
    all_r = np.concatenate([hr, sh_r, th_r])
    all_c = np.concatenate([hc, sh_c, th_c])
    hess_colind, hess_row, hkeys = _ccs(all_r, all_c, nw, nw)
    assert len(np.unique(hc * nw + hr)) == len(hr), 'cell Hessian blocks overlap'
    hpos = np.searchsorted(hkeys, all_c * nw + all_r)
    cell_hslot = -np.ones((sb.N, sb.cell_nh), dtype=np.int32)
    cell_hslot[hn, hs] = hpos[:len(hr)]
    covered = np.zeros(len(hkeys), dtype=bool)
    covered[hpos[:len(hr)]] = True
    sh_slot = hpos[len(hr):len(hr) + len(sh_r)].astype(np.int32)
    tail = None
    if sb.tail is not None:
        tail_hslot = hpos[len(hr) + len(sh_r):]
        taken = covered.copy()
        taken[sh_slot] = True
        tail = sb.tail.tape(tail_jslot, tail_hslot, taken[tail_hslot])

    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    nsr = len(s_rows)
    srow = dict(
        n=nsr,
        row=i32([s.row for s in s_rows]),
        kind=i32([s.kind for s in s_rows]),
        scale=i32([s.scale_vp for s in s_rows]),
        var_ptr=i32(np.cumsum([0] + [len(s.vars) for s in s_rows])),
        var=i32(sj_c),
        jslot=i32(srow_jslot),
        form_ptr=i32(np.cumsum([0] + [len(s.c) for s in s_rows])),
        coef_ptr=i32(np.cumsum([0] + [s.A.size for s in s_rows])),
        A=f64(np.concatenate([s.A.ravel() for s in s_rows]) if nsr else np.zeros(0)),
        c=f64(np.concatenate([s.c for s in s_rows]) if nsr else np.zeros(0)),
    )
    shess = dict(
        n=len(sh_keys),
        slot=i32(sh_slot),
        add=i32(covered[sh_slot] if len(sh_keys) else []),
        ptr=i32(np.cumsum([0] + [len(sh[k]) for k in sh_keys])),
        row=i32([e[0] for k in sh_keys for e in sh[k]]),
        coef=f64([e[1] for k in sh_keys for e in sh[k]]),
        scale=i32([e[2] for k in sh_keys for e in sh[k]]),
    )
    cell = dict(row=i32(sb.cell_row), coef=f64(sb.cell_coef), partner=i32(sb.cell_partner),
                pcoef=f64(sb.cell_pcoef), off=f64(sb.cell_off), par=f64(sb.cell_par),
                jslot=i32(cell_jslot), hslot=i32(cell_hslot), nj=sb.cell_nj, nh=sb.cell_nh,
                ncp=sb.cell_par.shape[1])
    if sb.transcription == COLLOC:
        cell.update(tmpl_j=i32(sb.tmpl_j), tmpl_h=i32(sb.tmpl_h), C=f64(sb.C), D=f64(sb.D), B=f64(sb.B))
    return NLPStructure(
        transcription=sb.transcription, variant=sb.variant, N=sb.N, K=sb.K, nz=sb.nz, nu=sb.nu,
        nw=nw, ng=ng, R=f64(sb.R), dR=f64(sb.dR), fc=None if sb.fc is None else f64(sb.fc),
        lbg=f64(sb.lbg), ubg=f64(sb.ubg), cell=cell, srow=srow, shess=shess,
        jac_colind=jac_colind, jac_row=jac_row, hess_colind=hess_colind, hess_row=hess_row, tail=tail)
