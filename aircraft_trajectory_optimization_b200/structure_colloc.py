'''
Collocation interval cell: local slot layout and CCS assembly (host side of csrc/colloc_cells.cuh).

Rows a cell owns (reference drone3d/raceline/base_raceline.py:398-434 ode rows, :460-490 /
:1132-1181 continuity to the next interval or loop closure through cont(sum_k D_k Z[n,k]);
drone3d/raceline/drone_raceline.py:42-45 quaternion renormalisation):

  rs(k)    k=0..K   sum_j C[j][k] Z[n,j][0] / H[n] >= 0            (parametric frame only)
  rf(k,i)  k=1..K   f_i(Z[n,k], U[n,k]) - sum_j C[j][k] Z[n,j][i] / H[n]
  rd(k,j)  k=0..K   dU[n,k][j] - sum_m C[m][k] U[n,m][j] / H[n]
  re(c)             coef * cont(sum_k D_k Z[n,k])[c] + pcoef * w[partner] + off
  reu(j)            coef * sum_k D_k U[n,k][j]       + pcoef * w[partner] + off

Every interior cell has the same local structure, so the kernel works from one *template*:
local contribution slot -> unique local entry id (uid); a per-cell table then maps uid -> CCS
position (-1 where the cell lacks the row).  uids are ordered like the CCS itself (column, then
row), so consecutive uids are (almost always) consecutive in global memory and the kernel's
copy-out from shared memory is coalesced.
'''
import numpy as np

KP = 8      # collocation points per interval (K = 7) -- the kernel is built for this


class CollocLayout:
    ''' offsets of the local slot groups; mirrored by struct CollocLayout in colloc_cells.cuh '''

    def __init__(self, nz, nu, NJ, NW):
        self.nz, self.nu, self.NJ, self.NW = nz, nu, NJ, NW
        nx = nz + nu
        self.S = nz + 2 * nu
        o = 0
        self.JS = o; o += KP * 9
        self.JF = o; o += 7 * NJ
        self.JC = o; o += 7 * nz * KP
        self.JH = o; o += 7 * nz
        self.JD = o; o += KP * nu * 10
        self.JE = o; o += nz * KP * 4
        self.JEP = o; o += nz
        self.JEU = o; o += nu * KP
        self.JEUP = o; o += nu
        self.NJS = o
        o = 0
        self.HW = o; o += 7 * NW
        self.HUU = o; o += KP * nu
        self.HDD = o; o += KP * nu
        self.HHZ = o; o += KP * nz
        self.HHU = o; o += KP * nu
        self.HHD = o; o += KP * nu
        self.HHH = o; o += 1
        self.HQ = o; o += 32 * 32
        self.NHS = o
        # local rows
        self.RS = 0
        self.RF = KP
        self.RD = self.RF + 7 * nz
        self.RE = self.RD + KP * nu
        self.REU = self.RE + nz
        self.NCR = self.REU + nu
        self.nx = nx

    # local variable ids: 0 = h, then point k: z, u, du
    def vz(self, k, i):
        return 1 + k * self.S + i

    def vu(self, k, j):
        return 1 + k * self.S + self.nz + j

    def vdu(self, k, j):
        return 1 + k * self.S + self.nx + j

    def rs(self, k):
        return self.RS + k

    def rf(self, k, i):
        return self.RF + (k - 1) * self.nz + i

    def rd(self, k, j):
        return self.RD + k * self.nu + j

    def re(self, c):
        return self.RE + c

    def reu(self, j):
        return self.REU + j


def init_colloc_cells(sb):
    if sb.K + 1 != KP:
        raise NotImplementedError('the collocation kernel is built for K = 7')
    meta = sb.meta
    sb.layout = CollocLayout(sb.nz, sb.nu, len(meta['J']), len(meta['W']))
    L = sb.layout
    N = sb.N
    sb.cell_row = -np.ones((N, L.NCR), dtype=np.int32)
    sb.cell_coef = np.zeros((N, L.NCR))
    sb.cell_partner = -np.ones((N, L.NCR), dtype=np.int32)
    sb.cell_pcoef = np.zeros((N, L.NCR))
    sb.cell_off = np.zeros((N, L.NCR))
    sb.cell_par = np.zeros((N, 1))

    def set_collocation(tau, B, C, D):
        sb.tau, sb.B, sb.C, sb.D = (np.asarray(a, dtype=float) for a in (tau, B, C, D))
    sb.set_collocation = set_collocation

    def sdot_row(n, k, row):
        sb.cell_row[n, L.rs(k)] = row
    sb.colloc_sdot_row = sdot_row

    def defect_rows(n, k, rows):
        for i, r in enumerate(rows):
            sb.cell_row[n, L.rf(k, i)] = r
    sb.colloc_defect_rows = defect_rows

    def du_rows(n, k, rows):
        for j, r in enumerate(rows):
            sb.cell_row[n, L.rd(k, j)] = r
    sb.colloc_du_rows = du_rows

    def end_row(n, c, row, coef, partner, pcoef, off):
        ''' c < nz: state component c of cont(sum D Z); c >= nz: input component c - nz '''
        lr = L.re(c) if c < sb.nz else L.reu(c - sb.nz)
        sb.cell_row[n, lr] = row
        sb.cell_coef[n, lr] = coef
        sb.cell_partner[n, lr] = partner
        sb.cell_pcoef[n, lr] = pcoef
        sb.cell_off[n, lr] = off
    sb.colloc_end_row = end_row


def _template(sb):
    '''
    (lrow, lvar) of every Jacobian slot and (lvar_a, lvar_b) of every Hessian slot, -1 where the
    reference's expression has no such term; plus the condition each Hessian slot needs.
    '''
    L = sb.layout
    nz, nu, NJ, NW = L.nz, L.nu, L.NJ, L.NW
    C, D, B = sb.C, sb.D, sb.B
    Jpat, Wpat = sb.meta['J'], sb.meta['W']
    quat = bool(sb.meta['quat'])
    jrow = -np.ones(L.NJS, dtype=np.int64)
    jvar = -np.ones(L.NJS, dtype=np.int64)

    def J(slot, r, v):
        jrow[slot], jvar[slot] = r, v

    for k in range(KP):
        for j in range(KP):
            if C[j][k] != 0:
                J(L.JS + k * 9 + j, L.rs(k), L.vz(j, 0))
        J(L.JS + k * 9 + 8, L.rs(k), 0)
    for k in range(1, KP):
        for e, (r, c) in enumerate(Jpat):
            J(L.JF + (k - 1) * NJ + e, L.rf(k, r), 1 + k * L.S + c)
        for i in range(nz):
            for j in range(KP):
                if C[j][k] != 0:
                    J(L.JC + ((k - 1) * nz + i) * KP + j, L.rf(k, i), L.vz(j, i))
            J(L.JH + (k - 1) * nz + i, L.rf(k, i), 0)
    for k in range(KP):
        for j in range(nu):
            b0 = L.JD + (k * nu + j) * 10
            for m in range(KP):
                if C[m][k] != 0:
                    J(b0 + m, L.rd(k, j), L.vu(m, j))
            J(b0 + 8, L.rd(k, j), L.vdu(k, j))
            J(b0 + 9, L.rd(k, j), 0)
    for c in range(nz):
        isq = quat and 3 <= c < 7
        for k in range(KP):
            if D[k] == 0:
                continue
            if isq:
                for b in range(4):
                    J(L.JE + (c * KP + k) * 4 + b, L.re(c), L.vz(k, 3 + b))
            else:
                J(L.JE + (c * KP + k) * 4, L.re(c), L.vz(k, c))
    for j in range(nu):
        for k in range(KP):
            if D[k] != 0:
                J(L.JEU + j * KP + k, L.reu(j), L.vu(k, j))
    # partner entries are not local variables: handled per cell (lvar = -2 marks them)
    for c in range(nz):
        J(L.JEP + c, L.re(c), -2)
    for j in range(nu):
        J(L.JEUP + j, L.reu(j), -2)

    ha = -np.ones(L.NHS, dtype=np.int64)
    hb = -np.ones(L.NHS, dtype=np.int64)
    hcond = np.zeros(L.NHS, dtype=np.int64)       # 0 always, 1 needs a quaternion end row

    def H(slot, a, b, cond=0):
        ha[slot], hb[slot], hcond[slot] = min(a, b), max(a, b), cond

    for k in range(1, KP):
        for e, (r, c) in enumerate(Wpat):
            H(L.HW + (k - 1) * NW + e, 1 + k * L.S + r, 1 + k * L.S + c)
    for k in range(KP):
        for j in range(nu):
            if sb.R[j] != 0 and B[k] != 0:
                H(L.HUU + k * nu + j, L.vu(k, j), L.vu(k, j))
            if sb.dR[j] != 0 and B[k] != 0:
                H(L.HDD + k * nu + j, L.vdu(k, j), L.vdu(k, j))
                H(L.HHD + k * nu + j, 0, L.vdu(k, j))
    parametric = sb.fc is not None
    for j in range(KP):
        for i in range(nz):
            if any(C[j][k] != 0 for k in range(1, KP)) or (parametric and i == 0):
                H(L.HHZ + j * nz + i, 0, L.vz(j, i))
        for jj in range(nu):
            H(L.HHU + j * nu + jj, 0, L.vu(j, jj))
    H(L.HHH, 0, 0)
    if quat:
        for k in range(KP):
            for a in range(4):
                for l in range(KP):
                    for b in range(4):
                        p, q = k * 4 + a, l * 4 + b
                        if p <= q and D[k] != 0 and D[l] != 0:
                            H(L.HQ + p * 32 + q, L.vz(k, 3 + a), L.vz(l, 3 + b), 1)
    return jrow, jvar, ha, hb, hcond


def colloc_entries(sb):
    ''' template uids + per-cell (row, col, cell, uid) entry lists for _assemble '''
    L = sb.layout
    N, nz, nu = sb.N, sb.nz, sb.nu
    jrow, jvar, ha, hb, hcond = _template(sb)

    # ---- Jacobian uids: unique (lvar, lrow), partner pseudo-variables last --------------------
    present = jrow >= 0
    lv = np.where(jvar == -2, 10 ** 6 + jrow, jvar)          # each partner entry is its own column
    key = np.where(present, lv * 10 ** 4 + jrow, -1)
    ukeys = np.unique(key[present])
    tmpl_j = -np.ones(L.NJS, dtype=np.int32)
    tmpl_j[present] = np.searchsorted(ukeys, key[present])
    nju = len(ukeys)
    u_row = (ukeys % 10 ** 4).astype(np.int64)
    u_var = (ukeys // 10 ** 4).astype(np.int64)

    # ---- Hessian uids: unique (lvar_b, lvar_a) -------------------------------------------------
    hp = ha >= 0
    hkey = np.where(hp, hb * 10 ** 4 + ha, -1)
    hukeys = np.unique(hkey[hp])
    tmpl_h = -np.ones(L.NHS, dtype=np.int32)
    tmpl_h[hp] = np.searchsorted(hukeys, hkey[hp])
    nhu = len(hukeys)
    hu_a = (hukeys % 10 ** 4).astype(np.int64)
    hu_b = (hukeys // 10 ** 4).astype(np.int64)
    hu_cond = np.zeros(nhu, dtype=np.int64)
    # an entry needs the quaternion rows only if ALL of its contributions do
    hu_cond[:] = 1
    np.minimum.at(hu_cond, tmpl_h[hp], hcond[hp])

    sb.tmpl_j, sb.tmpl_h = tmpl_j, tmpl_h
    sb.cell_nj, sb.cell_nh = nju, nhu

    jr, jc, jn, js = [], [], [], []
    hr, hc, hn, hs = [], [], [], []
    quat_rows = [L.re(c) for c in range(3, 7)] if sb.meta['quat'] else []
    for n in range(N):
        gvar = np.empty(1 + KP * L.S, dtype=np.int64)
        gvar[0] = sb.iH(n)
        gvar[1:] = sb.base(n, 0) + np.arange(KP * L.S)
        rows = sb.cell_row[n]
        grow = rows[u_row]
        ok = grow >= 0
        is_partner = u_var >= 10 ** 6
        gcol = np.where(is_partner, sb.cell_partner[n][np.where(is_partner, u_var - 10 ** 6, 0)],
                        gvar[np.where(is_partner, 0, u_var)])
        ok &= gcol >= 0
        idx = np.nonzero(ok)[0]
        jr.append(grow[idx]), jc.append(gcol[idx]), jn.append(np.full(len(idx), n)), js.append(idx)
        has_q = any(rows[r] >= 0 for r in quat_rows)
        okh = (hu_cond == 0) | has_q
        idx = np.nonzero(okh)[0]
        a, b = gvar[hu_a[idx]], gvar[hu_b[idx]]
        hr.append(np.minimum(a, b)), hc.append(np.maximum(a, b)), hn.append(np.full(len(idx), n)), hs.append(idx)
    cat = lambda v: np.concatenate(v).astype(np.int64)
    return (cat(jr), cat(jc), cat(jn), cat(js)), (cat(hr), cat(hc), cat(hn), cat(hs))
