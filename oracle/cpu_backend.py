'''
CPU backend for the interior-point driver.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Lets tests/ run aircraft_trajectory_optimization_b200.ipm.InteriorPoint on a machine without a GPU:
evaluations come from the oracle's tape interpreter (OracleNLP), the KKT solve from scipy's sparse
LU of the assembled matrix (kkt_blocks_ref.kkt_matrix).  It doubles as the CPU baseline of the
"converged solves/s" metric: it is what the reference's CPU path does per iteration (interpreted
SX evaluation + sparse direct solve), minus IPOPT itself, which is not installed.
'''
import numpy as np
import scipy.sparse.linalg as spla
import torch

from .kkt_blocks_ref import kkt_matrix


class OracleBackend:
    def __init__(self, nlp, st, check_inertia=False, ks=None):
        self.ks = ks            # KKTStructure: use the numpy block walk (it also yields the inertia)
        ''' nlp: OracleNLP; st: any object with the CCS pattern arrays + nw, ng (OracleNLP itself works) '''
        self.check_inertia = check_inertia
        self.nlp = nlp
        self.st = st
        self.ng = nlp.ng
        self.nw = nlp.nw

    def eval(self, x, lam_g, lam_f, derivs, idx=None, prev=None):
        ''' same contract as CudaBackend.eval: with idx, only those rows are evaluated and written into `prev` '''
        if idx is not None:
            sub = self.eval(x[idx], None if lam_g is None else lam_g[idx], None if lam_f is None else lam_f[idx], derivs)
            for k, v in sub.items():
                prev[k].index_copy_(0, idx, v)
            return dict(prev)
        B = x.shape[0]
        nlp = self.nlp
        IN = np.zeros((B, nlp.n_in))
        IN[:, :nlp.nw] = x.numpy()
        if lam_g is not None:
            IN[:, nlp.nw:nlp.nw + nlp.ng] = lam_g.numpy()
        IN[:, -1] = 1.0 if lam_f is None else lam_f.numpy()
        T = torch.from_numpy
        if not derivs:
            f = nlp.t_f.batch(IN)[:, 0]
            g = nlp.t_g.batch(IN)
            return dict(f=T(f.copy()), g=T(g))
        gf = nlp.t_grad_f.batch(IN)
        jg = nlp.t_jac_g.batch(IN)
        h = nlp.t_hess_l.batch(IN)
        return dict(f=T(gf[:, 0].copy()), grad_f=T(gf[:, 1:].copy()), g=T(jg[:, :nlp.ng].copy()),
                    jac=T(jg[:, nlp.ng:].copy()), hess=T(h))

    def select(self, keep):
        pass

    def eval_points(self, x_rows, inst):
        out = self.eval(x_rows, None, None, False)
        return dict(f=out['f'], g=out['g'])

    def kkt_matvec(self, hess, jac, dx_diag, neg_d, vec):
        out = np.empty(vec.shape)
        for b in range(vec.shape[0]):
            K = kkt_matrix(self.st, hess[b].numpy(), jac[b].numpy(), dx_diag[b].numpy(), -neg_d[b].numpy())
            out[b] = K @ vec[b].numpy()
        return torch.from_numpy(out)

    kkt_wave = 0       # instances per kernel wave (0: no speculative candidates); tests set it

    def kkt_solve_rows(self, hess, jac, idx, dx_diag, neg_d, rhs, refine_steps):
        ''' rows `idx` of hess / jac (repeats allowed) with dx_diag / neg_d / rhs given row by row '''
        return self.kkt_solve(hess[idx], jac[idx], dx_diag, neg_d, rhs, refine_steps)

    def kkt_solve(self, hess, jac, dx_diag, neg_d, rhs, refine_steps, idx=None):
        B = rhs.shape[0]
        sol = np.zeros(rhs.shape)
        status = np.zeros((B, 2), dtype=np.int32)
        status[:, 1] = self.ng
        for b in (range(B) if idx is None else idx.tolist()):
            K = kkt_matrix(self.st, hess[b].numpy(), jac[b].numpy(), dx_diag[b].numpy(), -neg_d[b].numpy())
            r = rhs[b].numpy()
            if self.ks is not None:
                from .kkt_blocks_ref import block_solve
                args = (self.ks, hess[b].numpy(), jac[b].numpy(), dx_diag[b].numpy(), -neg_d[b].numpy())
                try:
                    x, neg = block_solve(*args, r, with_inertia=True)
                    for _ in range(refine_steps):
                        rr = r - K @ x
                        if np.abs(rr).max() <= 1e-10 * max(1.0, np.abs(r).max()):      # like CudaBackend.refine_tol
                            break
                        x = x + block_solve(*args, rr)
                    sol[b], status[b, 1] = x, neg
                except np.linalg.LinAlgError:
                    sol[b] = np.nan
                    status[b, 0] = 1
                continue
            try:
                lu = spla.splu(K.tocsc())
                x = lu.solve(r)
                for _ in range(refine_steps):
                    rr = r - K @ x
                    if np.abs(rr).max() <= 1e-10 * max(1.0, np.abs(r).max()):
                        break
                    x = x + lu.solve(rr)
                sol[b] = x
                if self.check_inertia:
                    import scipy.linalg as sla
                    _, d, _ = sla.ldl(K.toarray())
                    ev = np.linalg.eigvalsh(d) if False else None
                    # block-diagonal d: count negative eigenvalues of its 1x1 / 2x2 blocks
                    neg = 0
                    i = 0
                    nk = d.shape[0]
                    while i < nk:
                        if i + 1 < nk and d[i + 1, i] != 0:
                            neg += int((np.linalg.eigvalsh(d[i:i + 2, i:i + 2]) < 0).sum())
                            i += 2
                        else:
                            neg += int(d[i, i] < 0)
                            i += 1
                    status[b, 1] = neg
            except RuntimeError:
                sol[b] = np.nan
                status[b, 0] = 1
        return torch.from_numpy(sol), torch.from_numpy(status)
