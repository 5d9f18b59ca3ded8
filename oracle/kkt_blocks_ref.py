'''
CPU checkers for the KKT solve.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

  * `kkt_matrix`   assembles the interior-point KKT matrix the way IPOPT hands it to its sparse
                   symmetric solver [third party; selected at drone3d/raceline/base_raceline.py:765-787]:
                   [[W + diag(dx_diag), J'], [J, -diag(D)]] from CCS values, as a scipy sparse matrix.
  * `sparse_solve` solves it with scipy's SuperLU (the independent answer).
  * `block_solve`  walks the product's block tables (aircraft_trajectory_optimization_b200/kkt.py)
                   in numpy, step for step like csrc/kkt_blocks.cuh, so that table errors and kernel
                   errors can be told apart.
'''
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def kkt_matrix(st, hess, jac, dx_diag, D):
    nw, ng = st.nw, st.ng
    W = sp.csc_matrix((hess, st.hess_row, st.hess_colind), shape=(nw, nw))
    W = W + sp.triu(W, 1).T + sp.diags(dx_diag)
    J = sp.csc_matrix((jac, st.jac_row, st.jac_colind), shape=(ng, nw))
    return sp.bmat([[W, J.T], [J, -sp.diags(D)]], format='csc')


def sparse_solve(st, hess, jac, dx_diag, D, rhs):
    K = kkt_matrix(st, hess, jac, dx_diag, D)
    return spla.splu(K).solve(rhs)


def _gather(ks, src, hess, jac, dx_diag, neg_d, aux=None):
    kind = src >> 28
    idx = src & ((1 << 28) - 1)
    out = np.empty(len(src))
    for k, arr in enumerate((hess, jac, dx_diag, neg_d, aux)):
        m = kind == k
        if m.any():
            out[m] = arr[idx[m]]
    return out


def sym_invert_bp(A):
    '''
    inverse and number of negative eigenvalues of the symmetric matrix A by Gauss-Jordan sweeps with
    Bunch-Kaufman pivoting -- the numpy twin of kkt_sym_invert in csrc/kkt_blocks.cuh
    '''
    M = np.array(A, dtype=float)
    b = M.shape[0]
    alpha = 0.6403882032022076
    unswept = np.ones(b, dtype=bool)
    neg = 0

    def argmax_col(c, skip):
        v = np.where(unswept, np.abs(M[:, c]), -1.0)
        v[skip] = -1.0
        i = int(np.argmax(v))
        return v[i], i

    while unswept.any():
        dg = np.where(unswept, np.abs(np.diag(M)), -1.0)
        k = int(np.argmax(dg))
        vk = dg[k]
        piv = (k,)
        if unswept.sum() > 1:
            lam, r = argmax_col(k, k)
            if not vk >= alpha * lam:
                sig, _ = argmax_col(r, r)
                if vk * sig >= alpha * lam * lam:
                    piv = (k,)
                elif abs(M[r, r]) >= alpha * sig:
                    piv = (r,)
                else:
                    piv = (min(k, r), max(k, r))
        if len(piv) == 1:
            k = piv[0]
            d = M[k, k]
            if not abs(d) > 1e-250:
                d = 1e-250
            neg += d < 0
            col = M[:, k].copy()
            row = M[k, :] / d
            M -= np.outer(col, row)
            M[k, :] = row
            M[:, k] = -col / d
            M[k, k] = 1.0 / d
            unswept[k] = False
        else:
            p, q = piv
            E = M[np.ix_([p, q], [p, q])]
            det = E[0, 0] * E[1, 1] - E[0, 1] ** 2
            if not abs(det) > 1e-250:
                det = -1e-250
            neg += 1 if det < 0 else (2 if E[0, 0] + E[1, 1] < 0 else 0)
            Ei = np.array([[E[1, 1], -E[0, 1]], [-E[0, 1], E[0, 0]]]) / det
            cols = M[:, [p, q]].copy()
            rows = Ei @ M[[p, q], :]
            M -= cols @ rows
            M[[p, q], :] = rows
            M[:, [p, q]] = -cols @ Ei
            M[np.ix_([p, q], [p, q])] = Ei
            unswept[[p, q]] = False
    return M, int(neg)


def block_solve(ks, hess, jac, dx_diag, D, rhs, with_inertia=False):
    ''' ks: KKTStructure.  Returns the solution in the original (w, g) ordering (and, on request, the
    number of negative eigenvalues summed over the block pivots, which equals that of K). '''
    neg = 0
    N, nb, bmax, qmax = ks.N, ks.nb, ks.bmax, ks.qmax
    nw = ks.nw
    neg_d = -np.asarray(D)
    nrhs = 1 + nb
    diag_of = lambda u: np.where(u < nw, dx_diag[np.minimum(u, nw - 1)], neg_d[np.maximum(u - nw, 0)])
    Z = [None] * N
    YL = [None] * N
    carry = None
    rcarry = None
    for n in range(N):
        u = ks.unk[ks.blk_ptr[n]:ks.blk_ptr[n + 1]]
        b = len(u)
        M = np.zeros((bmax, bmax))
        a, e = ks.dA_ptr[n], ks.dA_ptr[n + 1]
        M.ravel()[ks.dA_pos[a:e]] = _gather(ks, ks.dA_src[a:e], hess, jac, dx_diag, neg_d)
        M[np.arange(b), np.arange(b)] += diag_of(u)
        Y = np.zeros((b, nrhs))
        Y[:, 0] = rhs[u]
        a, e = ks.bE_ptr[n], ks.bE_ptr[n + 1]
        Y[ks.bE_row[a:e], 1 + ks.bE_col[a:e]] = _gather(ks, ks.bE_src[a:e], hess, jac, dx_diag, neg_d)
        if n > 0:
            cr = ks.cr[ks.cr_ptr[n - 1]:ks.cr_ptr[n]]
            M[np.ix_(cr, cr)] -= carry
            Y[cr] -= rcarry
        Sinv, ng_ = sym_invert_bp(M[:b, :b])
        neg += ng_
        Z[n] = Sinv @ Y
        if n < N - 1:
            cc = ks.cc[ks.cc_ptr[n]:ks.cc_ptr[n + 1]]
            m = ks.cr_ptr[n + 1] - ks.cr_ptr[n]
            Lc = np.zeros((max(m, 1), qmax))
            a, e = ks.cL_ptr[n], ks.cL_ptr[n + 1]
            Lc.ravel()[ks.cL_pos[a:e]] = _gather(ks, ks.cL_src[a:e], hess, jac, dx_diag, neg_d)
            Lc = Lc[:m, :len(cc)]
            YL[n] = Sinv[:, cc] @ Lc.T              # b x m
            carry = Lc @ YL[n][cc]                  # m x m
            rcarry = Lc @ Z[n][cc]                  # m x nrhs
    X = [None] * N
    X[N - 1] = Z[N - 1]
    for n in range(N - 2, -1, -1):
        cr = ks.cr[ks.cr_ptr[n]:ks.cr_ptr[n + 1]]
        X[n] = Z[n] - YL[n] @ X[n + 1][cr]
    sol = np.zeros(ks.nk)
    ub = ks.unk[ks.blk_ptr[N]:ks.blk_ptr[N + 1]]
    if nb:
        G = np.zeros((nb, nb))
        G.ravel()[ks.bG_pos] = _gather(ks, ks.bG_src, hess, jac, dx_diag, neg_d)
        G[np.arange(nb), np.arange(nb)] += diag_of(ub)
        rb = rhs[ub].copy()
        for n in range(N):
            a, e = ks.bE_ptr[n], ks.bE_ptr[n + 1]
            v = _gather(ks, ks.bE_src[a:e], hess, jac, dx_diag, neg_d)
            # E' X: row j of E' picks entries (row l of block n, column j)
            np.subtract.at(G, ks.bE_col[a:e], v[:, None] * X[n][ks.bE_row[a:e], 1:])
            np.subtract.at(rb, ks.bE_col[a:e], v * X[n][ks.bE_row[a:e], 0])
        Ginv, ng_ = sym_invert_bp(G)
        xb = Ginv @ rb
        neg += ng_
        sol[ub] = xb
    for n in range(N):
        u = ks.unk[ks.blk_ptr[n]:ks.blk_ptr[n + 1]]
        sol[u] = X[n][:, 0] - (X[n][:, 1:] @ xb if nb else 0.0)
    return (sol, neg) if with_inertia else sol


def chain_factor(ks, hess, jac, dx_diag, D, aux=None):
    '''
    numpy twin of kkt_factor_kernel (csrc/kkt_chain.cuh): block LDL' of the chain with the border last.  The border
    columns are carried in interface form: Y_n = (L^-1 E)_n lives on the support rows sup_n of block n only,
    P_n = Y_n[sup_n, :act_n];  Q_n = S_n^-1[:, sup_n] P_n;  border Schur complement G - sum_n P_n' Q_n[sup_n].
    Returns the factor dict and the number of negative eigenvalues.
    '''
    N, nb, bmax, qmax, nw = ks.N, ks.nb, ks.bmax, ks.qmax, ks.nw
    neg_d = -np.asarray(D)
    diag_of = lambda u: np.where(u < nw, dx_diag[np.minimum(u, nw - 1)], neg_d[np.maximum(u - nw, 0)])
    val = lambda src: _gather(ks, src, hess, jac, dx_diag, neg_d, aux)
    F = dict(Sinv=[None] * N, YL=[None] * N, P=[None] * N, Q=[None] * N)
    neg = 0
    carry = None
    Pnext = None
    Gacc = np.zeros((nb, nb))
    for n in range(N):
        u = ks.unk[ks.blk_ptr[n]:ks.blk_ptr[n + 1]]
        b = len(u)
        M = np.zeros((bmax, bmax))
        a, e = ks.dA_ptr[n], ks.dA_ptr[n + 1]
        M.ravel()[ks.dA_pos[a:e]] = val(ks.dA_src[a:e])
        M[np.arange(b), np.arange(b)] += diag_of(u)
        if n > 0:
            cr = ks.cr[ks.cr_ptr[n - 1]:ks.cr_ptr[n]]
            M[np.ix_(cr, cr)] -= carry
        Sinv, ng_ = sym_invert_bp(M[:b, :b])
        neg += ng_
        sup = ks.sup[ks.sup_ptr[n]:ks.sup_ptr[n + 1]]
        an = int(ks.act[n])
        P = np.zeros((len(sup), an))
        if n > 0:
            crs = ks.crs[ks.cr_ptr[n - 1]:ks.cr_ptr[n]]
            P[crs, :Pnext.shape[1]] = Pnext
        a, e = ks.bE_ptr[n], ks.bE_ptr[n + 1]
        np.add.at(P, (ks.bE_sup[a:e], ks.bE_col[a:e]), val(ks.bE_src[a:e]))
        Q = Sinv[:, sup] @ P
        Gacc[:an, :an] += P.T @ Q[sup]
        F['Sinv'][n], F['P'][n], F['Q'][n] = Sinv, P, Q
        if n < N - 1:
            cc = ks.cc[ks.cc_ptr[n]:ks.cc_ptr[n + 1]]
            m = ks.cr_ptr[n + 1] - ks.cr_ptr[n]
            Lc = np.zeros((max(m, 1), qmax))
            a, e = ks.cL_ptr[n], ks.cL_ptr[n + 1]
            Lc.ravel()[ks.cL_pos[a:e]] = val(ks.cL_src[a:e])
            Lc = Lc[:m, :len(cc)]
            YL = Sinv[:, cc] @ Lc.T
            carry = Lc @ YL[cc]
            Pnext = -YL[sup].T @ P
            F['YL'][n] = YL
    if nb:
        ub = ks.unk[ks.blk_ptr[N]:ks.blk_ptr[N + 1]]
        G = np.zeros((nb, nb))
        G.ravel()[ks.bG_pos] = val(ks.bG_src)
        G[np.arange(nb), np.arange(nb)] += diag_of(ub)
        F['SB'], ng_ = sym_invert_bp(G - Gacc)
        neg += ng_
    return F, int(neg)


def chain_solve(ks, F, rhs):
    ''' numpy twin of kkt_solve_kernel: one right-hand side with the factors of chain_factor '''
    N, nb = ks.N, ks.nb
    z = [None] * N
    rc = None
    racc = np.zeros(nb)
    for n in range(N):
        u = ks.unk[ks.blk_ptr[n]:ks.blk_ptr[n + 1]]
        y = rhs[u].copy()
        if n > 0:
            y[ks.cr[ks.cr_ptr[n - 1]:ks.cr_ptr[n]]] -= rc
        z[n] = F['Sinv'][n] @ y
        if n < N - 1:
            rc = F['YL'][n].T @ y
        sup = ks.sup[ks.sup_ptr[n]:ks.sup_ptr[n + 1]]
        racc[:ks.act[n]] += F['P'][n].T @ z[n][sup]
    sol = np.zeros(ks.nk)
    xb = np.zeros(0)
    if nb:
        ub = ks.unk[ks.blk_ptr[N]:ks.blk_ptr[N + 1]]
        xb = F['SB'] @ (rhs[ub] - racc)
        sol[ub] = xb
    x_next = None
    for n in range(N - 1, -1, -1):
        x = z[n] - F['Q'][n] @ xb[:ks.act[n]]
        if n < N - 1:
            x -= F['YL'][n] @ x_next[ks.cr[ks.cr_ptr[n]:ks.cr_ptr[n + 1]]]
        sol[ks.unk[ks.blk_ptr[n]:ks.blk_ptr[n + 1]]] = x
        x_next = x
    return sol


def condensed_factor(cs, hess, jac, dx_diag, D):
    '''
    numpy twin of csrc/kkt_condense.cuh + the chain kernels on the reduced system: interiors A_n inverted independently,
    T_n = B_n' A_n^-1 B_n gathered into the aux values of the reduced chain (aircraft_trajectory_optimization_b200/
    kkt_condensed.py).  Returns (factors, negative eigenvalues).
    '''
    nw = cs.nw
    neg_d = -np.asarray(D)
    diag_of = lambda u: np.where(u < nw, dx_diag[np.minimum(u, nw - 1)], neg_d[np.maximum(u - nw, 0)])
    val = lambda src: _gather(None, src, hess, jac, dx_diag, neg_d)
    amax, smax = cs.amax, cs.smax
    neg = 0
    Ainv, G = [], []
    Tbuf = np.zeros(cs.NI * smax * smax)
    for n in range(cs.NI):
        u = cs.iunk[cs.iu_ptr[n]:cs.iu_ptr[n + 1]]
        a = len(u)
        s = cs.su_ptr[n + 1] - cs.su_ptr[n]
        M = np.zeros((amax, amax))
        e0, e1 = cs.iA_ptr[n], cs.iA_ptr[n + 1]
        M.ravel()[cs.iA_pos[e0:e1]] = val(cs.iA_src[e0:e1])
        M[np.arange(a), np.arange(a)] += diag_of(u)
        Bm = np.zeros((amax, smax))
        e0, e1 = cs.iB_ptr[n], cs.iB_ptr[n + 1]
        Bm.ravel()[cs.iB_pos[e0:e1]] = val(cs.iB_src[e0:e1])
        Bm = Bm[:a, :s]
        Ai, ng_ = sym_invert_bp(M[:a, :a])
        neg += ng_
        Gn = Ai @ Bm
        T = np.zeros((smax, smax))
        T[:s, :s] = Bm.T @ Gn
        Tbuf[n * smax * smax:(n + 1) * smax * smax] = T.ravel()
        Ainv.append(Ai), G.append(Gn)
    aux = np.where(cs.aux_orig >= 0, val(np.maximum(cs.aux_orig, 0)), 0.0)
    aux -= np.add.reduceat(np.concatenate([Tbuf[cs.aux_c_idx], [0.0]]), cs.aux_c_ptr[:-1]) * (np.diff(cs.aux_c_ptr) > 0)
    F, ng_ = chain_factor(cs.chain, hess, jac, dx_diag, D, aux=aux)
    F['Ainv'], F['G'] = Ainv, G
    return F, neg + ng_


def condensed_solve(cs, F, rhs):
    smax = cs.smax
    tbuf = np.zeros(cs.NI * smax)
    y = []
    for n in range(cs.NI):
        u = cs.iunk[cs.iu_ptr[n]:cs.iu_ptr[n + 1]]
        s = cs.su_ptr[n + 1] - cs.su_ptr[n]
        y.append(F['Ainv'][n] @ rhs[u])
        tbuf[n * smax:n * smax + s] = F['G'][n].T @ rhs[u]
    rhs2 = rhs.copy()
    rhs2[cs.rsep] -= np.add.reduceat(np.concatenate([tbuf[cs.r_c_idx], [0.0]]), cs.r_c_ptr[:-1])
    sol = chain_solve(cs.chain, F, rhs2)
    for n in range(cs.NI):
        u = cs.iunk[cs.iu_ptr[n]:cs.iu_ptr[n + 1]]
        su = cs.sunk[cs.su_ptr[n]:cs.su_ptr[n + 1]]
        sol[u] = y[n] - F['G'][n] @ sol[su]
    return sol
