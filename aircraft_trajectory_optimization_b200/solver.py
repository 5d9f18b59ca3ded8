'''
nlpsol-shaped solver object: what `ca.nlpsol('solver', 'ipopt', prob, opts)` returns in the reference
(drone3d/raceline/base_raceline.py:752-799), backed by the batched interior-point driver (ipm.py)
and the CUDA library.  Call signature and result keys follow CasADi's nlpsol [third party]:

    sol = solver(x0=, lbx=, ubx=, lbg=, ubg=)      ->  {'x', 'f', 'g', 'lam_x', 'lam_g', 'lam_p'}
    solver.stats()                                  ->  {'success', 'return_status', 'iter_count',
                                                         't_wall_nlp_f', ..., 't_wall_total'}

`x0` may also be a (B, nw) array (and `p` a (B, nvp) array of vehicle parameters): B independent
instances are then solved in lock step on the GPU and every result gets a leading batch dimension.
There is no CPU fallback: without the CUDA library or a device the first call raises.
'''
import numpy as np

RETURN_STATUS = {0: 'Solve_Succeeded', 1: 'Solved_To_Acceptable_Level', 2: 'Maximum_Iterations_Exceeded',
                 3: 'Search_Direction_Becomes_Too_Small'}


class InteriorPointSolver:
    def __init__(self, functions_factory, max_iter=1000, verbose=True, options=None, device=None):
        self._factory = functions_factory
        self._functions = None
        self._backend = None
        self.max_iter = max_iter
        self.verbose = verbose
        self.options = options
        self.device = device
        self.result = None
        self._stats = dict(success=False, return_status='Not_Run', iter_count=0, t_wall_nlp_f=0.0,
                           t_wall_nlp_g=0.0, t_wall_nlp_grad_f=0.0, t_wall_nlp_hess_l=0.0,
                           t_wall_nlp_jac_g=0.0, t_wall_linear_solver=0.0, t_wall_total=0.0)

    @property
    def functions(self):
        if self._functions is None:
            self._functions = self._factory()
        return self._functions

    def stats(self):
        return dict(self._stats)

    def __call__(self, x0, lbx, ubx, lbg, ubg, p=None, lam_g0=None):
        import torch
        from .ipm import InteriorPoint, IpmOptions, CudaBackend
        F = self.functions
        if not torch.cuda.is_available():
            raise RuntimeError('the raceline interior-point solver runs only on a CUDA device')
        dev = torch.device('cuda', torch.cuda.current_device() if self.device is None else self.device)
        T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
        x0 = np.asarray(x0, dtype=np.float64)
        single = x0.ndim == 1
        X0 = T(np.atleast_2d(x0))
        vp = T(F.vp if p is None else p)
        if vp.dim() == 2 and vp.shape[0] != X0.shape[0]:
            raise ValueError('p must have one row of vehicle parameters per instance')
        if self._backend is None:
            self._backend = CudaBackend(F, vp)
        self._backend.vp = vp.contiguous()
        self._backend._vp_full = self._backend.vp
        opts = self.options or IpmOptions()
        opts.max_iter = self.max_iter
        opts.verbose = bool(self.verbose) and single
        ip = InteriorPoint(self._backend, opts)
        r = ip.solve(X0, T(lbx), T(ubx), T(lbg), T(ubg), None if lam_g0 is None else T(np.atleast_2d(lam_g0)))
        self.result = r
        st = r.status.cpu().numpy()
        self._stats.update(success=bool(r.success.all()), iter_count=int(r.iterations.max()),
                           return_status=RETURN_STATUS[int(st.max())] if single else [RETURN_STATUS[int(k)] for k in st],
                           success_each=r.success.cpu().numpy(), iterations_each=r.iterations.cpu().numpy(),
                           t_wall_nlp_hess_l=r.t_eval, t_wall_linear_solver=r.t_kkt, t_wall_total=r.t_total,
                           n_eval=r.n_eval, n_factor=r.n_factor, n_soc=getattr(r, 'n_soc', 0))
        out = {'x': r.x.cpu().numpy(), 'f': r.f.cpu().numpy(), 'g': r.g.cpu().numpy(),
               'lam_x': r.lam_x.cpu().numpy(), 'lam_g': r.lam_g.cpu().numpy(), 'lam_p': np.zeros(0)}
        if single:
            out = {k: (v[0] if k != 'lam_p' else v) for k, v in out.items()}
            out['f'] = float(out['f'])
        return out
