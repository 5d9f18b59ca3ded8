'''
Block structure of the interior-point KKT system (host side of csrc/kkt_blocks.cuh).

Every interior-point iteration solves (reference: inside IPOPT, reached through
`self.solver(x0=..., ...)`, drone3d/raceline/base_raceline.py:160-165; the reference picks the
sparse symmetric-indefinite solver MA97 / MUMPS at :765-787)

        [ W + Sigma_x + delta_w I      J'   ] [dx]   [r_x]
        [ J                           -D    ] [dy] = [r_g]

with W the upper-triangular CCS values of hess_l, J the CCS values of jac_g, D >= 0 diagonal
(slack elimination of inequality rows + delta_c).  The unknowns are regrouped by interval:

  * stage triple n = [c_{n-1}, x_n; rows] with x_n = (Z[n,0], U[n,0]) the node variables, c_{n-1} =
    (H[n-1], dU[n-1,0], interior collocation points of interval n-1) the variables that produce x_n,
    and every row that is the identity on x_n (continuity / shooting rows n-1 -> n, which stay full
    rank inside the triple even through the quaternion renormalisation) or only touches the triple;
  * the triples are swept from n = N-1 down to 1: a Riccati-type recursion whose pivots are the
    stage-wise reduced Hessians (eliminating c_n together with its own node x_n, or the rows with the
    interval they leave instead of the node they define, makes the diagonal blocks singular or
    ill-conditioned -- both were tried);
  * the loop closes through triple 0 = [c_{N-1}, x_0; closure rows] (base_raceline.py:492-514,
    drone_raceline.py:47-104), which forms the dense *border* together with the phase-start step
    sizes of the global-frame equal-step rows (base_raceline.py:891-905).

In that order the matrix is block-tridiagonal (N-1 diagonal blocks S_n, couplings L_n between
consecutive triples) plus the border.  The kernel factors
it by a block forward sweep with dense, partially pivoted Gauss-Jordan inverses of the diagonal
Schur complements, carries 1 + nb right-hand sides (the residual and the border columns), sweeps
back, and finishes with the dense nb x nb border system.

All tables are int32 and problem-independent; values are gathered from (hess, jac, dx_diag, neg_D)
through `src` codes  kind << 28 | index   (kind 0 hess, 1 jac, 2 dx_diag, 3 neg_D).
'''
from dataclasses import dataclass

import numpy as np

K_HESS, K_JAC, K_DX, K_ND = 0, 1, 2, 3
_SHIFT = 28


def src_code(kind, idx):
    idx = np.asarray(idx, dtype=np.int64)
    assert (idx < (1 << _SHIFT)).all()
    return ((kind << _SHIFT) | idx).astype(np.int32)


@dataclass
class KKTStructure:
    nw: int
    ng: int
    N: int                  # number of diagonal blocks
    nb: int                 # border size
    bmax: int               # largest diagonal block
    mmax: int               # most coupling rows (rows of block n+1 touched by block n)
    qmax: int               # most coupling columns (variables of block n touched from block n+1)
    blk_ptr: np.ndarray     # [N+2] into unk (block N = border)
    unk: np.ndarray         # KKT index (i < nw variable, nw + r row) of every block-local unknown
    nvar: np.ndarray        # [N+1] number of variables at the head of every block
    # diagonal blocks: S[pos] = value(src); pos = r * bmax + c, both triangles listed
    dA_ptr: np.ndarray
    dA_src: np.ndarray
    dA_pos: np.ndarray
    # coupling L_n (compressed): rows cr (local in block n+1), columns cc (local in block n)
    cr_ptr: np.ndarray
    cr: np.ndarray
    cc_ptr: np.ndarray
    cc: np.ndarray
    cL_ptr: np.ndarray
    cL_src: np.ndarray
    cL_pos: np.ndarray      # rr * qmax + cc (compressed indices)
    # border columns: entry (local row in block n, border column j)
    bE_ptr: np.ndarray
    bE_src: np.ndarray
    bE_row: np.ndarray
    bE_col: np.ndarray
    # border-border
    bG_src: np.ndarray
    bG_pos: np.ndarray      # r * nb + c, both triangles listed
    # for tests / reference assembly: block id and local index of every KKT unknown
    blk: np.ndarray
    loc: np.ndarray
    # interface form of the border columns (csrc/kkt_chain.cuh).  Border unknowns are ordered by the first chain
    # block they couple to, so that at block n the columns 0 .. act[n]-1 are the ones that have appeared so far.
    # The forward-eliminated border column block Y_n = (L^-1 E)_n is non-zero only on the rows
    # sup_n = cr_{n-1} U rows(E_n) of block n ("support"); P_n = Y_n[sup_n, :act[n]] is what the kernels carry.
    act: np.ndarray = None      # [N] active border columns at block n (non-decreasing)
    sup_ptr: np.ndarray = None  # [N+1]
    sup: np.ndarray = None      # local row indices (ascending) of the support of block n
    crs: np.ndarray = None      # [cr_ptr[N]] position of cr_n[a] (a row of block n+1) inside sup_{n+1}
    bE_sup: np.ndarray = None   # [len(bE_src)] position of bE_row[e] inside sup_n
    p_off: np.ndarray = None    # [N+1] offset (doubles) of P_n in the factor storage: sup rows x ldq(n)
    q_off: np.ndarray = None    # [N+1] offset (doubles) of Q_n = S_n^-1[:, sup_n] P_n: b_n rows x ldq(n)
    amax: int = 0
    smax: int = 0

    @property
    def nk(self):
        return self.nw + self.ng

    @property
    def nrhs(self):
        return 1 + self.nb

    def tables(self):
        ''' the int32 arrays the C ABI takes, in rb_kkt_desc order '''
        return dict(blk_ptr=self.blk_ptr, unk=self.unk, nvar=self.nvar,
                    dA_ptr=self.dA_ptr, dA_src=self.dA_src, dA_pos=self.dA_pos,
                    cr_ptr=self.cr_ptr, cr=self.cr, cc_ptr=self.cc_ptr, cc=self.cc,
                    cL_ptr=self.cL_ptr, cL_src=self.cL_src, cL_pos=self.cL_pos,
                    bE_ptr=self.bE_ptr, bE_src=self.bE_src, bE_row=self.bE_row, bE_col=self.bE_col,
                    bG_src=self.bG_src, bG_pos=self.bG_pos,
                    act=self.act, sup_ptr=self.sup_ptr, sup=self.sup, crs=self.crs, bE_sup=self.bE_sup,
                    p_off=self.p_off, q_off=self.q_off)


def _var_stage(st):
    '''
    stage triple of every variable: node variables x_n = (Z[n,0], U[n,0]) belong to triple n, everything
    else of interval n (H[n], dU[n,0] and the interior collocation points) is c_n and belongs to triple
    n + 1 -- next to the node x_{n+1} it produces.  Returns (triple, is_node).
    '''
    S, P, N = st.nz + 2 * st.nu, st.K + 1, st.N
    i = np.arange(st.nw - N)
    n = i // (P * S)
    k = (i % (P * S)) // S
    j = i % S
    node = (k == 0) & (j < st.nz + st.nu)
    t = np.empty(st.nw, dtype=np.int64)
    t[:N] = np.arange(N) + 1
    t[N:] = np.where(node, n, n + 1)
    is_node = np.zeros(st.nw, dtype=bool)
    is_node[N:] = node
    return t, is_node


def _pattern(st):
    ''' (row, column) of every jac_g / hess_l entry '''
    jr = np.asarray(st.jac_row, dtype=np.int64)
    jc = np.repeat(np.arange(st.nw), np.diff(st.jac_colind))
    hr = np.asarray(st.hess_row, dtype=np.int64)
    hc = np.repeat(np.arange(st.nw), np.diff(st.hess_colind))
    return jr, jc, hr, hc


def _assign_blocks(st):
    ''' chain block (0 .. N-1 in sweep order, N = border) of every KKT unknown.  Returns (blk, N). '''
    nw, ng, N = st.nw, st.ng, st.N
    t, is_node = _var_stage(st)
    jr = np.asarray(st.jac_row, dtype=np.int64)
    jc = np.repeat(np.arange(nw), np.diff(st.jac_colind))
    hr = np.asarray(st.hess_row, dtype=np.int64)
    hc = np.repeat(np.arange(nw), np.diff(st.hess_colind))

    # ---- stage triples and border ---------------------------------------------------------------------
    # Triple n = [c_{n-1}, x_n; rows that are the identity on x_n].  The loop closes through triple 0 =
    # triple N = [c_{N-1}, x_0; closure rows], which becomes the border: x_0 is `low` (it couples to
    # triple 1 like any x_{n-1} does), c_{N-1} is `high` (anything touching it is a border row).
    # Collocation (K > 0): c_{N-1} holds the interior points of the last interval (~150 variables); it stays in
    # the chain as a block of its own, [c_{N-1}; defect / rate rows of the last interval], and only the closure
    # rows (which are the identity on x_0) join x_0 in the border -- otherwise the border would be as large as a
    # stage block.  Shooting (K = 0): c_{N-1} is five variables and goes to the border with its rows.
    chain_last = st.K > 0
    LOW, HIGH = -1, N
    top = N if chain_last else N - 1                 # triple index of the first chain block
    tt = t.copy()
    tt[(t == 0)] = LOW
    if not chain_last:
        tt[(t == N)] = HIGH
    # far couplings that are not part of the loop (the equal-step star rows of the global frame): the
    # variable at the low end of several far rows moves to the border as a `low` variable
    for _ in range(8):
        rmax = np.full(ng, LOW, dtype=np.int64)
        np.maximum.at(rmax, jr, tt[jc])
        far = (tt[jc] != LOW) & (tt[jc] < rmax[jr] - 1)
        hfar = (tt[hr] != LOW) & (tt[hc] != LOW) & (np.abs(tt[hr] - tt[hc]) > 1)
        if not far.any() and not hfar.any():
            break
        tt[jc[far]] = LOW
        lo = np.where(tt[hr] < tt[hc], hr, hc)
        tt[lo[hfar]] = LOW
    else:
        raise RuntimeError('could not arrange the KKT system into a block-tridiagonal chain + border')
    rmax = np.full(ng, LOW, dtype=np.int64)
    np.maximum.at(rmax, jr, tt[jc])
    row_t = np.where(rmax == LOW, HIGH, rmax)
    if chain_last:
        # rows of the last block that reach back to x_0 are the closure rows: border
        touches_low = np.zeros(ng, dtype=bool)
        touches_low[jr[tt[jc] == LOW]] = True
        row_t = np.where((row_t == N) & touches_low, HIGH + 1, row_t)
        BORDER_T = HIGH + 1
    else:
        BORDER_T = HIGH

    # chain blocks are the triples top, top-1, ..., 1 in that order (the sweep runs backward in time: a
    # Riccati recursion); block index Nc is the border
    Nc = top
    to_chain = lambda tv: np.where((tv == LOW) | (tv == BORDER_T), Nc, Nc - tv)
    blk = np.empty(nw + ng, dtype=np.int64)
    blk[:nw] = to_chain(tt)
    blk[nw:] = to_chain(np.where(rmax == LOW, BORDER_T, row_t))
    N = Nc

    # ---- structural rank of the equality rows inside every triple -------------------------------------
    # A triple must be able to satisfy its own equality rows with its own variables, otherwise its diagonal
    # block is singular whatever the values (the axial gate equation of the global frame pins the node
    # position, which the shooting rows and the equal-step row of that triple already determine).  Rows
    # that are structurally dependent on the earlier rows of their triple -- tested with random values on
    # the sparsity pattern -- move to the border.
    eq_row = np.asarray(st.lbg) == np.asarray(st.ubg)
    rng = np.random.default_rng(12345)
    jval = rng.uniform(0.5, 1.5, len(jr)) * rng.choice([-1.0, 1.0], len(jr))
    ent_ok = eq_row[jr] & (blk[nw + jr] < N) & (blk[jc] == blk[nw + jr])
    order_e = np.lexsort((jr[ent_ok], blk[nw + jr][ent_ok]))
    e_r, e_c, e_v, e_b = jr[ent_ok][order_e], jc[ent_ok][order_e], jval[ent_ok][order_e], blk[nw + jr][ent_ok][order_e]
    bstart = np.searchsorted(e_b, np.arange(N + 1))
    demoted = []
    for n in range(N):
        a, b = bstart[n], bstart[n + 1]
        if a == b:
            continue
        rows_u, ri = np.unique(e_r[a:b], return_inverse=True)
        cols_u, ci = np.unique(e_c[a:b], return_inverse=True)
        A = np.zeros((len(rows_u), len(cols_u)))
        A[ri, ci] = e_v[a:b]
        Q = np.zeros((0, len(cols_u)))
        for k in range(len(rows_u)):
            r = A[k] - (Q.T @ (Q @ A[k]) if len(Q) else 0.0)
            if np.linalg.norm(r) > 1e-8 * np.linalg.norm(A[k]):
                Q = np.vstack([Q, r / np.linalg.norm(r)])
            else:
                demoted.append(rows_u[k])
    # equality rows of a triple without any variable of that triple cannot be pivoted there either
    has_own = np.zeros(ng, dtype=bool)
    has_own[jr[blk[jc] == blk[nw + jr]]] = True
    demoted += list(np.nonzero(eq_row & (blk[nw:] < N) & ~has_own)[0])
    if demoted:
        blk[nw + np.array(demoted, dtype=np.int64)] = N
    return blk, N


def build_kkt_structure(st) -> KKTStructure:
    ''' st: NLPStructure (structure.py) '''
    blk, N = _assign_blocks(st)
    jr, jc, hr, hc = _pattern(st)
    ei = np.concatenate([hr, st.nw + jr])
    ej = np.concatenate([hc, jc])
    es = np.concatenate([src_code(K_HESS, np.arange(len(hr))), src_code(K_JAC, np.arange(len(jr)))])
    return _tables_from_blocks(st.nw, st.ng, N, blk, ei, ej, es)


def _tables_from_blocks(nw, ng, N, blk, ei, ej, es) -> KKTStructure:
    '''
    block tables of the symmetric matrix with entries (ei, ej, src es), every unordered pair listed once, whose
    unknowns are grouped by blk: 0 .. N-1 chain blocks (entries only inside a block or between neighbours), N the
    border, N + 1 unknowns that do not take part (eliminated beforehand: kkt_condensed.py)
    '''
    # local order inside a block: variables (ascending w index) then rows (ascending g index); border unknowns by the
    # first chain block they couple to (unknowns that only couple inside the border last)
    ei0, ej0 = ei, ej
    first = np.zeros(nw + ng, dtype=np.int64)
    fb = np.full(nw + ng, N, dtype=np.int64)
    m_i = (blk[ei0] == N) & (blk[ej0] < N)
    np.minimum.at(fb, ei0[m_i], blk[ej0][m_i])
    m_j = (blk[ej0] == N) & (blk[ei0] < N)
    np.minimum.at(fb, ej0[m_j], blk[ei0][m_j])
    first[blk == N] = fb[blk == N]
    order = np.lexsort((np.arange(nw + ng), first, blk))
    unk = order.astype(np.int32)
    counts = np.bincount(blk, minlength=N + 2)[:N + 1]
    blk_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    loc = np.empty(nw + ng, dtype=np.int64)
    loc[order] = np.arange(nw + ng) - blk_ptr[blk[order]]
    nvar = np.bincount(blk[:nw], minlength=N + 2)[:N + 1].astype(np.int32)
    bmax = int(counts[:N].max())
    nb = int(counts[N])

    # ---- all KKT entries (i, j, src) in the lower+upper sense: list each unordered pair once ------
    bi, bj = blk[ei], blk[ej]
    li, lj = loc[ei], loc[ej]

    in_chain = (bi < N) & (bj < N)
    assert (np.abs(bi - bj)[in_chain] <= 1).all()

    # diagonal blocks (both triangles; the diagonal once)
    dmask = in_chain & (bi == bj)
    d_b = bi[dmask]
    d_r, d_c, d_s = li[dmask], lj[dmask], es[dmask]
    off = d_r != d_c
    dA_b = np.concatenate([d_b, d_b[off]])
    dA_pos = np.concatenate([d_r * bmax + d_c, (d_c * bmax + d_r)[off]])
    dA_src = np.concatenate([d_s, d_s[off]])
    o = np.argsort(dA_b, kind='stable')
    dA_b, dA_pos, dA_src = dA_b[o], dA_pos[o], dA_src[o]
    dA_ptr = np.concatenate([[0], np.cumsum(np.bincount(dA_b, minlength=N))]).astype(np.int32)

    # couplings: entry between block n (column side) and block n+1 (row side)
    cmask = in_chain & (bi != bj)
    c_hi_is_i = bi[cmask] > bj[cmask]
    c_n = np.minimum(bi[cmask], bj[cmask])                       # the lower block
    c_row = np.where(c_hi_is_i, li[cmask], lj[cmask])            # local index in block n+1
    c_col = np.where(c_hi_is_i, lj[cmask], li[cmask])            # local index in block n
    c_src = es[cmask]
    cr_list, cc_list, cL_pos, cL_src_l = [], [], [], []
    cr_ptr, cc_ptr, cL_ptr = [0], [0], [0]
    mmax = qmax = 0
    o = np.argsort(c_n, kind='stable')
    c_n, c_row, c_col, c_src = c_n[o], c_row[o], c_col[o], c_src[o]
    starts = np.searchsorted(c_n, np.arange(N + 1))
    per_block = []
    for n in range(N):
        a, b = starts[n], starts[n + 1]
        rows_u, rr = np.unique(c_row[a:b], return_inverse=True)
        cols_u, cc_ = np.unique(c_col[a:b], return_inverse=True)
        per_block.append((rows_u, cols_u, rr, cc_, c_src[a:b]))
        mmax, qmax = max(mmax, len(rows_u)), max(qmax, len(cols_u))
    for rows_u, cols_u, rr, cc_, s in per_block:
        cr_list.append(rows_u), cc_list.append(cols_u)
        cL_pos.append(rr * qmax + cc_), cL_src_l.append(s)
        cr_ptr.append(cr_ptr[-1] + len(rows_u)), cc_ptr.append(cc_ptr[-1] + len(cols_u))
        cL_ptr.append(cL_ptr[-1] + len(s))
    cat = lambda v, dt=np.int32: (np.concatenate(v) if len(v) else np.zeros(0)).astype(dt)

    # border columns
    emask = (bi == N) ^ (bj == N)
    e_border_is_i = bi[emask] == N
    e_n = np.where(e_border_is_i, bj[emask], bi[emask])
    e_row = np.where(e_border_is_i, lj[emask], li[emask])
    e_col = np.where(e_border_is_i, li[emask], lj[emask])
    e_src = es[emask]
    o = np.argsort(e_n, kind='stable')
    e_n, e_row, e_col, e_src = e_n[o], e_row[o], e_col[o], e_src[o]
    bE_ptr = np.concatenate([[0], np.cumsum(np.bincount(e_n, minlength=N))]).astype(np.int32)

    gmask = (bi == N) & (bj == N)
    g_r, g_c, g_s = li[gmask], lj[gmask], es[gmask]
    off = g_r != g_c
    bG_pos = np.concatenate([g_r * max(nb, 1) + g_c, (g_c * max(nb, 1) + g_r)[off]])
    bG_src = np.concatenate([g_s, g_s[off]])

    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    # ---- interface form of the border columns -----------------------------------------------------------
    first_b = first[order[blk_ptr[N]:blk_ptr[N + 1]]]               # per border column, ascending
    act = np.searchsorted(first_b, np.arange(N), side='right').astype(np.int32)
    cr_ptr_a, cr_a = np.asarray(cr_ptr), cat(cr_list)
    sup_list, crs_list, bE_sup = [], [], np.zeros(len(e_src), dtype=np.int32)
    for n in range(N):
        rows_e = e_row[bE_ptr[n]:bE_ptr[n + 1]]
        prev = cr_a[cr_ptr_a[n - 1]:cr_ptr_a[n]] if n > 0 else np.zeros(0, dtype=np.int64)
        sup_n = np.union1d(prev, rows_e).astype(np.int64)
        sup_list.append(sup_n)
        bE_sup[bE_ptr[n]:bE_ptr[n + 1]] = np.searchsorted(sup_n, rows_e)
        if n > 0:
            crs_list.append(np.searchsorted(sup_n, prev))
    crs_list.append(np.zeros(cr_ptr_a[N] - cr_ptr_a[N - 1], dtype=np.int64))      # the last block couples to nothing
    sup_ptr = np.concatenate([[0], np.cumsum([len(v) for v in sup_list])]).astype(np.int32)
    ldq = (act + 1) // 2 * 2                                           # even row strides: 16-byte rows for TMA
    bsz = np.diff(blk_ptr)[:N]
    p_off = np.concatenate([[0], np.cumsum(np.diff(sup_ptr) * ldq)]).astype(np.int32)
    q_off = np.concatenate([[0], np.cumsum(bsz * ldq)]).astype(np.int32)
    return KKTStructure(
        act=act, sup_ptr=sup_ptr, sup=cat(sup_list), crs=cat(crs_list), bE_sup=bE_sup, p_off=p_off, q_off=q_off,
        amax=int(act.max()) if N else 0, smax=int(np.diff(sup_ptr).max()) if N else 0,
        nw=nw, ng=ng, N=N, nb=nb, bmax=bmax, mmax=mmax, qmax=qmax,
        blk_ptr=blk_ptr, unk=unk, nvar=nvar,
        dA_ptr=dA_ptr, dA_src=i32(dA_src), dA_pos=i32(dA_pos),
        cr_ptr=i32(cr_ptr), cr=cat(cr_list), cc_ptr=i32(cc_ptr), cc=cat(cc_list),
        cL_ptr=i32(cL_ptr), cL_src=cat(cL_src_l), cL_pos=cat(cL_pos),
        bE_ptr=bE_ptr, bE_src=i32(e_src), bE_row=i32(e_row), bE_col=i32(e_col),
        bG_src=i32(bG_src), bG_pos=i32(bG_pos), blk=blk, loc=loc)


# ---------------------------------------------------------------------------------------------------
# ctypes binding (device tensors in, device tensors out; no CPU fallback)
import ctypes

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class _KktDesc(ctypes.Structure):
    _fields_ = ([(k, ctypes.c_int) for k in ('nw', 'ng', 'N', 'nb', 'bmax', 'mmax', 'qmax', 'nnz_hess',
                                            'nnz_jac', 'n_bG')]
                + [(k, _i32p) for k in ('blk_ptr', 'unk', 'dA_ptr', 'dA_src', 'dA_pos', 'cr_ptr', 'cr',
                                        'cc_ptr', 'cc', 'cL_ptr', 'cL_src', 'cL_pos', 'bE_ptr', 'bE_src',
                                        'bE_row', 'bE_col', 'bG_src', 'bG_pos')]
                + [(k, ctypes.c_int) for k in ('amax', 'smax')]
                + [(k, _i32p) for k in ('act', 'sup_ptr', 'sup', 'crs', 'bE_sup', 'p_off', 'q_off')]
                + [(k, _i64p) for k in ('jac_colind', 'jac_row', 'hess_colind', 'hess_row')])


class _KktInteriorDesc(ctypes.Structure):
    _fields_ = ([(k, ctypes.c_int) for k in ('NI', 'amax', 'smax', 'n_aux', 'n_rsep')]
                + [(k, _i32p) for k in ('iu_ptr', 'iunk', 'su_ptr', 'sunk', 'iA_ptr', 'iA_pos', 'iA_src', 'iB_ptr',
                                        'iB_pos', 'iB_src', 'aux_orig', 'aux_c_ptr', 'aux_c_idx', 'rsep', 'r_c_ptr',
                                        'r_c_idx')])


def bind_kkt(lib):
    vp = ctypes.c_void_p
    lib.rb_kkt_set_interiors.argtypes = [vp, ctypes.POINTER(_KktInteriorDesc)]
    lib.rb_kkt_create.argtypes = [ctypes.POINTER(_KktDesc), ctypes.POINTER(vp)]
    lib.rb_kkt_destroy.argtypes = [vp]
    lib.rb_kkt_destroy.restype = None
    lib.rb_kkt_factor_bytes.argtypes = [vp, ctypes.c_int]
    lib.rb_kkt_factor_bytes.restype = ctypes.c_size_t
    lib.rb_kkt_factor_solve.argtypes = [vp, ctypes.c_int] + [vp] * 9
    lib.rb_kkt_resolve.argtypes = [vp, ctypes.c_int] + [vp] * 8
    lib.rb_kkt_matvec.argtypes = [vp, ctypes.c_int] + [vp] * 7
    lib.rb_kkt_resolve_rows.argtypes = [vp, ctypes.c_int, ctypes.c_int] + [vp] * 5


class KktSolver:
    '''
    Batched KKT factor / solve on the GPU for one problem structure.  All arguments are torch CUDA
    fp64 tensors (contiguous); see include/raceline_b200.h for shapes.
    '''

    def __init__(self, st, ks: KKTStructure = None, condensed=None):
        '''
        condensed: eliminate the interiors of collocation intervals first (kkt_condensed.py); default: whenever the
        transcription has interior collocation points (K > 0).  `ks` then is the structure of the reduced system.
        '''
        from .functions import load_library, _check
        self._check = _check
        self.lib = load_library()
        bind_kkt(self.lib)
        self.st = st
        auto = condensed is None
        if auto:
            condensed = ks is None and st.K > 0
        try:
            self._create(st, ks, condensed)
        except RuntimeError:
            if not (auto and condensed):
                raise
            # the reduced system of this structure does not fit the shared-memory chain kernels (wide couplings
            # next to a wide border): the uncondensed global-memory kernel solves it, slower
            self.__del__()
            self._create(st, None, False)
        self._factors = None
        self._factors_B = 0
        self._last_B = 0

    def _create(self, st, ks, condensed):
        _check = self._check
        self.handle = None
        self.cs = None
        if condensed:
            from .kkt_condensed import build_condensed_structure
            self.cs = build_condensed_structure(st)
            ks = self.cs.chain
        self.ks = ks or build_kkt_structure(st)
        ks = self.ks
        self._keep = []
        d = _KktDesc()
        d.nw, d.ng, d.N, d.nb, d.bmax, d.mmax, d.qmax = ks.nw, ks.ng, ks.N, ks.nb, ks.bmax, ks.mmax, ks.qmax
        d.nnz_hess, d.nnz_jac, d.n_bG = st.nnz_hess, st.nnz_jac, len(ks.bG_src)
        d.amax, d.smax = ks.amax, ks.smax
        for name, arr in ks.tables().items():
            if name == 'nvar':
                continue
            a = np.ascontiguousarray(arr, dtype=np.int32)
            self._keep.append(a)
            setattr(d, name, a.ctypes.data_as(_i32p))
        for name in ('jac_colind', 'jac_row', 'hess_colind', 'hess_row'):
            a = np.ascontiguousarray(getattr(st, name), dtype=np.int64)
            self._keep.append(a)
            setattr(d, name, a.ctypes.data_as(_i64p))
        self.handle = ctypes.c_void_p()
        _check(self.lib.rb_kkt_create(ctypes.byref(d), ctypes.byref(self.handle)), 'rb_kkt_create')
        if self.cs is not None:
            cs = self.cs
            di = _KktInteriorDesc()
            di.NI, di.amax, di.smax, di.n_aux, di.n_rsep = cs.NI, cs.amax, cs.smax, cs.n_aux, len(cs.rsep)
            for name, arr in cs.tables().items():
                a = np.ascontiguousarray(arr, dtype=np.int32)
                self._keep.append(a)
                setattr(di, name, a.ctypes.data_as(_i32p))
            _check(self.lib.rb_kkt_set_interiors(self.handle, ctypes.byref(di)), 'rb_kkt_set_interiors')

    def __del__(self):
        try:
            if getattr(self, 'handle', None) is not None and self.handle.value:
                self.lib.rb_kkt_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def _factor_buffer(self, B, device):
        import torch
        if self._factors is None or self._factors_B < B or self._factors.device != device:
            self._factors = None
            nbytes = self.lib.rb_kkt_factor_bytes(self.handle, B)
            self._factors = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
            self._factors_B = B
        return self._factors

    @staticmethod
    def _p(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    def _args(self, hess, jac, dx_diag, neg_d):
        import torch
        for t in (hess, jac, dx_diag, neg_d):
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float64
        B = hess.shape[0]
        assert hess.shape == (B, self.st.nnz_hess) and jac.shape == (B, self.st.nnz_jac)
        assert dx_diag.shape == (B, self.ks.nw) and neg_d.shape == (B, self.ks.ng)
        return B

    def factor_solve(self, hess, jac, dx_diag, neg_d, rhs, sol=None, status=None, stream=None):
        import torch
        B = self._args(hess, jac, dx_diag, neg_d)
        if sol is None:
            sol = torch.empty_like(rhs)
        if status is None:
            status = torch.zeros(B, 2, dtype=torch.int32, device=rhs.device)
        fac = self._factor_buffer(B, rhs.device)
        self._last_B = B
        if stream is None:
            stream = torch.cuda.current_stream(rhs.device).cuda_stream
        p = self._p
        self._check(self.lib.rb_kkt_factor_solve(self.handle, B, p(hess), p(jac), p(dx_diag), p(neg_d), p(rhs),
                                                 p(sol), p(fac), p(status), ctypes.c_void_p(stream)),
                    'rb_kkt_factor_solve')
        return sol, status

    def resolve(self, hess, jac, dx_diag, neg_d, rhs, sol=None, stream=None):
        import torch
        B = self._args(hess, jac, dx_diag, neg_d)
        if sol is None:
            sol = torch.empty_like(rhs)
        assert self._factors is not None and self._factors_B >= B, 'factor_solve first'
        if stream is None:
            stream = torch.cuda.current_stream(rhs.device).cuda_stream
        p = self._p
        self._check(self.lib.rb_kkt_resolve(self.handle, B, p(hess), p(jac), p(dx_diag), p(neg_d), p(rhs), p(sol),
                                            p(self._factors), ctypes.c_void_p(stream)), 'rb_kkt_resolve')
        return sol

    @property
    def can_resolve_rows(self):
        return self.cs is None and self.ks.bmax <= 64 and self.ks.nb <= 64

    def resolve_rows(self, inst, rhs, sol=None, stream=None):
        ''' solve with the stored factors for a subset: rhs row p belongs to the instance in factor slot inst[p]
        (int32 CUDA tensor) of the last factor_solve call '''
        import torch
        assert inst.dtype == torch.int32 and inst.is_cuda and inst.is_contiguous() and rhs.is_contiguous()
        assert self._factors is not None and self._last_B >= 1, 'factor_solve first'
        if sol is None:
            sol = torch.empty_like(rhs)
        if stream is None:
            stream = torch.cuda.current_stream(rhs.device).cuda_stream
        p = self._p
        self._check(self.lib.rb_kkt_resolve_rows(self.handle, rhs.shape[0], self._last_B, p(inst), p(rhs), p(sol),
                                                 p(self._factors), ctypes.c_void_p(stream)), 'rb_kkt_resolve_rows')
        return sol

    def matvec(self, hess, jac, dx_diag, neg_d, vec, out=None, stream=None):
        import torch
        B = self._args(hess, jac, dx_diag, neg_d)
        if out is None:
            out = torch.empty_like(vec)
        if stream is None:
            stream = torch.cuda.current_stream(vec.device).cuda_stream
        p = self._p
        self._check(self.lib.rb_kkt_matvec(self.handle, B, p(hess), p(jac), p(dx_diag), p(neg_d), p(vec), p(out),
                                           ctypes.c_void_p(stream)), 'rb_kkt_matvec')
        return out
