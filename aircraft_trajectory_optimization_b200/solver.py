'''
nlpsol-shaped solver object (placeholder until the batched interior-point driver lands).
'''


class InteriorPointSolver:
    def __init__(self, functions_factory, max_iter=1000, verbose=True):
        self._factory = functions_factory
        self._functions = None
        self.max_iter = max_iter
        self.verbose = verbose
        self._stats = dict(success=False, t_wall_nlp_f=0.0, t_wall_nlp_g=0.0, t_wall_nlp_grad_f=0.0,
                           t_wall_nlp_hess_l=0.0, t_wall_nlp_jac_g=0.0)

    @property
    def functions(self):
        if self._functions is None:
            self._functions = self._factory()
        return self._functions

    def stats(self):
        return dict(self._stats)

    def __call__(self, x0, lbx, ubx, lbg, ubg):
        raise NotImplementedError('interior-point driver not built yet')
