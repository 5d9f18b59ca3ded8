'''
Condensed form of the interior-point KKT system for collocation intervals (host side of csrc/kkt_condense.cuh).

The matrix is the one of kkt.py (what IPOPT hands to MA97 / MUMPS inside `self.solver(x0=...)`,
drone3d/raceline/base_raceline.py:160-165, :765-787).  With K = 7 collocation points per interval
(drone3d/raceline/base_raceline.py:398-490) a stage triple of kkt.py holds ~310 unknowns, and a chain of such blocks
is factorised one after the other by one CTA.  Almost all of a triple is *interior* to its interval: the variables
c_n = (H[n], dU[n,0], Z/U/dU[n,1..K]) and the rows that only involve interval n (collocation defects, input-rate
rows, path rows at interior points) see the rest of the problem only through the node x_n = (Z[n,0], U[n,0]) and the
continuity rows into node n+1.  Eliminating the interiors first,

    K = [ A   B ]      A = blockdiag(A_0 .. A_{N-1})  (interiors, independent of one another)
        [ B'  C ]      S = C - sum_n B_n' A_n^-1 B_n  (nodes, continuity / closure / gate rows, border)

turns the collocation problem into a multiple-shooting-shaped one: S is block tridiagonal with ~35 unknowns per
stage plus the loop-closure border -- exactly what the shared-memory chain kernels (csrc/kkt_chain.cuh) factorise --
and the N interior inversions run on N thread blocks at the same time instead of in sequence.  The inertia is the sum of
the inertias of the A_n and of S (Haynsworth), so IPOPT's regularisation logic is unchanged.

Tables (all int32, problem-independent):
  interiors     iu_ptr / iunk    KKT indices of the interior unknowns of interval n (variables first)
                su_ptr / sunk    the separator unknowns interval n touches, ascending KKT index
                iA_*             A_n entries: position r * amax + c (both triangles), value source (kkt.py codes)
                iB_*             B_n entries: position r * smax + c
  reduced       chain            KKTStructure of S; every value comes from the `aux` array (source kind 4)
                aux_orig         source of the entry of C at every aux slot (-1: pure fill)
                aux_c_ptr / idx  per aux slot: the entries n * smax^2 + a * smax + b of the T_n = B_n' A_n^-1 B_n that
                                 are subtracted from it (fixed order: deterministic sums)
                rsep / r_c_*     the same for the right-hand side: separator unknown, entries n * smax + a of B_n' y_n
'''
from dataclasses import dataclass

import numpy as np

from .kkt import (KKTStructure, _assign_blocks, _pattern, _tables_from_blocks, _var_stage, src_code, K_HESS, K_JAC)

K_AUX = 4


@dataclass
class CondensedStructure:
    nw: int
    ng: int
    NI: int                 # interiors (one per interval that has one)
    amax: int               # largest interior
    smax: int               # most separator unknowns touched by one interior
    iu_ptr: np.ndarray
    iunk: np.ndarray
    su_ptr: np.ndarray
    sunk: np.ndarray
    iA_ptr: np.ndarray
    iA_pos: np.ndarray
    iA_src: np.ndarray
    iB_ptr: np.ndarray
    iB_pos: np.ndarray
    iB_src: np.ndarray
    n_aux: int
    aux_orig: np.ndarray
    aux_c_ptr: np.ndarray
    aux_c_idx: np.ndarray
    rsep: np.ndarray
    r_c_ptr: np.ndarray
    r_c_idx: np.ndarray
    chain: KKTStructure

    @property
    def nk(self):
        return self.nw + self.ng

    def tables(self):
        return dict(iu_ptr=self.iu_ptr, iunk=self.iunk, su_ptr=self.su_ptr, sunk=self.sunk,
                    iA_ptr=self.iA_ptr, iA_pos=self.iA_pos, iA_src=self.iA_src,
                    iB_ptr=self.iB_ptr, iB_pos=self.iB_pos, iB_src=self.iB_src,
                    aux_orig=self.aux_orig, aux_c_ptr=self.aux_c_ptr, aux_c_idx=self.aux_c_idx,
                    rsep=self.rsep, r_c_ptr=self.r_c_ptr, r_c_idx=self.r_c_idx)


def _independent_rows(rows, cols, vals):
    ''' rows (in order of first appearance) that are linearly independent on the given sparse pattern with random values '''
    rows_u, ri = np.unique(rows, return_inverse=True)
    cols_u, ci = np.unique(cols, return_inverse=True)
    A = np.zeros((len(rows_u), len(cols_u)))
    A[ri, ci] = vals
    Q = np.zeros((0, len(cols_u)))
    keep = []
    for k in range(len(rows_u)):
        r = A[k] - (Q.T @ (Q @ A[k]) if len(Q) else 0.0)
        if np.linalg.norm(r) > 1e-8 * np.linalg.norm(A[k]):
            Q = np.vstack([Q, r / np.linalg.norm(r)])
            keep.append(rows_u[k])
    return np.array(keep, dtype=np.int64)


def build_condensed_structure(st) -> CondensedStructure:
    ''' st: NLPStructure of a collocation problem (K > 0) '''
    assert st.K > 0, 'shooting intervals have no interior worth condensing'
    nw, ng, NI = st.nw, st.ng, st.N
    nk = nw + ng
    blk, N = _assign_blocks(st)
    blk = blk.copy()
    t, is_node = _var_stage(st)
    jr, jc, hr, hc = _pattern(st)

    # ---- interiors -------------------------------------------------------------------------------------------
    # interior variables: everything of interval n but its node, unless the chain assignment sent it to the border
    # (phase-start step sizes of the equal-step rows, base_raceline.py:891-905)
    ivl_of_var = np.where(is_node, -1, t - 1)
    int_var = (~is_node) & (blk[:nw] < N)
    own = np.full(nk, -1, dtype=np.int64)            # interior the unknown belongs to (-1: separator)
    own[:nw][int_var] = ivl_of_var[int_var]
    # interior rows: chain rows all of whose variables are interior variables of ONE interval, that interval's node or
    # border variables (continuity rows touch the next node and stay in the separator: restricted to the interior they
    # lose rank through the quaternion renormalisation, drone_raceline.py:42-45)
    var_ivl = np.where(is_node, t, t - 1)
    foreign = np.zeros(ng, dtype=bool)               # touches a chain variable of another interval
    lo = np.full(ng, 10 ** 9, dtype=np.int64)
    hi = np.full(ng, -1, dtype=np.int64)
    chain_var = blk[:nw] < N
    m = chain_var[jc]
    np.minimum.at(lo, jr[m], var_ivl[jc[m]])
    np.maximum.at(hi, jr[m], var_ivl[jc[m]])
    has_int = np.zeros(ng, dtype=bool)
    has_int[jr[int_var[jc]]] = True
    cand = (blk[nw:] < N) & has_int & (lo == hi)
    row_ivl = np.where(cand, lo, -1)
    # structural rank of the candidate rows on the interior variables of their interval
    rng = np.random.default_rng(54321)
    jval = rng.uniform(0.5, 1.5, len(jr)) * rng.choice([-1.0, 1.0], len(jr))
    ent = (row_ivl[jr] >= 0) & int_var[jc] & (row_ivl[jr] == ivl_of_var[jc])
    for n in range(NI):
        e = ent & (row_ivl[jr] == n)
        if not e.any():
            continue
        rows_n = np.nonzero(row_ivl == n)[0]
        keep = _independent_rows(jr[e], jc[e], jval[e])
        row_ivl[np.setdiff1d(rows_n, keep)] = -1
    own[nw:] = row_ivl
    for _ in range(4):
        # separator unknowns every interior touches must sit in two neighbouring chain blocks or in the border
        ei = np.concatenate([hr, nw + jr])
        ej = np.concatenate([hc, jc])
        a_int, b_int = own[ei] >= 0, own[ej] >= 0
        assert not (a_int & b_int & (own[ei] != own[ej])).any(), 'interiors of different intervals are coupled'
        cross = a_int ^ b_int
        n_of = np.where(a_int, own[ei], own[ej])[cross]
        sep_of = np.where(a_int, ej, ei)[cross]
        moved = False
        for n in np.unique(n_of):
            su = np.unique(sep_of[n_of == n])
            cb = blk[su]
            cb = cb[cb < N]
            if len(cb) and cb.max() - cb.min() > 1:
                # keep the two most populated neighbouring blocks, send the rest to the border
                vals, cnt = np.unique(cb, return_counts=True)
                best = max(vals, key=lambda v: cnt[vals == v].sum() + cnt[vals == v + 1].sum())
                far = su[(blk[su] < N) & (blk[su] != best) & (blk[su] != best + 1)]
                blk[far] = N
                moved = True
        if not moved:
            break
    else:
        raise RuntimeError('could not arrange the condensed KKT system into a chain + border')

    # ---- per-interior tables ---------------------------------------------------------------------------------
    es = np.concatenate([src_code(K_HESS, np.arange(len(hr))), src_code(K_JAC, np.arange(len(jr)))])
    ivals = [n for n in range(NI) if (own == n).any()]
    remap = -np.ones(NI, dtype=np.int64)
    remap[ivals] = np.arange(len(ivals))
    own = np.where(own >= 0, remap[np.maximum(own, 0)], -1)
    NIc = len(ivals)
    order = np.lexsort((np.arange(nk), own))
    order = order[own[order] >= 0]
    iu_ptr = np.concatenate([[0], np.cumsum(np.bincount(own[own >= 0], minlength=NIc))]).astype(np.int64)
    iunk = order
    iloc = np.full(nk, -1, dtype=np.int64)
    iloc[iunk] = np.arange(len(iunk)) - iu_ptr[own[iunk]]
    amax = int(np.diff(iu_ptr).max())
    a_int, b_int = own[ei] >= 0, own[ej] >= 0
    # A entries
    both = a_int & b_int
    an, ar, ac, asrc = own[ei][both], iloc[ei][both], iloc[ej][both], es[both]
    off = ar != ac
    A_n = np.concatenate([an, an[off]])
    A_pos = np.concatenate([ar * amax + ac, (ac * amax + ar)[off]])
    A_src = np.concatenate([asrc, asrc[off]])
    o = np.argsort(A_n, kind='stable')
    A_n, A_pos, A_src = A_n[o], A_pos[o], A_src[o]
    iA_ptr = np.concatenate([[0], np.cumsum(np.bincount(A_n, minlength=NIc))])
    # B entries and the separator lists
    cross = a_int ^ b_int
    bn = np.where(a_int, own[ei], own[ej])[cross]
    br = np.where(a_int, iloc[ei], iloc[ej])[cross]
    bs = np.where(a_int, ej, ei)[cross]
    bsrc = es[cross]
    su_list, B_pos, B_src, B_cnt = [], [], [], []
    smax = 0
    for n in range(NIc):
        smax = max(smax, len(np.unique(bs[bn == n])))
    for n in range(NIc):
        e = bn == n
        su, ci = np.unique(bs[e], return_inverse=True)
        su_list.append(su)
        B_pos.append(br[e] * smax + ci)
        B_src.append(bsrc[e])
        B_cnt.append(int(e.sum()))
    su_ptr = np.concatenate([[0], np.cumsum([len(s) for s in su_list])])
    iB_ptr = np.concatenate([[0], np.cumsum(B_cnt)])

    # ---- reduced system: entries of C among separator unknowns + fill of every T_n --------------------------------
    sep = own < 0
    cmask = sep[ei] & sep[ej]
    ci_, cj_ = np.minimum(ei[cmask], ej[cmask]), np.maximum(ei[cmask], ej[cmask])
    keys = [ci_ * nk + cj_]
    for su in su_list:
        ia, ib = np.triu_indices(len(su))
        keys.append(su[ia] * nk + su[ib])
    allkeys = np.unique(np.concatenate(keys))
    n_aux = len(allkeys)
    aux_orig = -np.ones(n_aux, dtype=np.int64)
    aux_orig[np.searchsorted(allkeys, ci_ * nk + cj_)] = es[cmask]
    # contributions of the T_n
    c_slot, c_idx = [], []
    for n, su in enumerate(su_list):
        ia, ib = np.triu_indices(len(su))
        c_slot.append(np.searchsorted(allkeys, su[ia] * nk + su[ib]))
        c_idx.append(n * smax * smax + ia * smax + ib)
    c_slot, c_idx = np.concatenate(c_slot), np.concatenate(c_idx)
    o = np.argsort(c_slot, kind='stable')
    aux_c_ptr = np.concatenate([[0], np.cumsum(np.bincount(c_slot, minlength=n_aux))])
    aux_c_idx = c_idx[o]
    # right-hand side contributions
    r_unk = np.concatenate(su_list)
    r_idx = np.concatenate([n * smax + np.arange(len(su)) for n, su in enumerate(su_list)])
    rsep, rinv = np.unique(r_unk, return_inverse=True)
    o = np.argsort(rinv, kind='stable')
    r_c_ptr = np.concatenate([[0], np.cumsum(np.bincount(rinv, minlength=len(rsep)))])
    r_c_idx = r_idx[o]

    # chain blocks of the separator unknowns: the triples of kkt.py without their interiors, empty ones removed
    rblk = np.where(sep, blk, N + 1)
    used = np.unique(rblk[rblk < N])
    renum = -np.ones(N + 2, dtype=np.int64)
    renum[used] = np.arange(len(used))
    Nr = len(used)
    renum[N], renum[N + 1] = Nr, Nr + 1
    rblk = renum[rblk]
    ki, kj = allkeys // nk, allkeys % nk
    chain = _tables_from_blocks(nw, ng, Nr, rblk, ki, kj, src_code(K_AUX, np.arange(n_aux)))
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return CondensedStructure(
        nw=nw, ng=ng, NI=NIc, amax=amax, smax=int(smax), iu_ptr=i32(iu_ptr), iunk=i32(iunk), su_ptr=i32(su_ptr),
        sunk=i32(np.concatenate(su_list)), iA_ptr=i32(iA_ptr), iA_pos=i32(A_pos), iA_src=i32(A_src),
        iB_ptr=i32(iB_ptr), iB_pos=i32(np.concatenate(B_pos)), iB_src=i32(np.concatenate(B_src)),
        n_aux=n_aux, aux_orig=i32(aux_orig), aux_c_ptr=i32(aux_c_ptr), aux_c_idx=i32(aux_c_idx),
        rsep=i32(rsep), r_c_ptr=i32(r_c_ptr), r_c_idx=i32(r_c_idx), chain=chain)
