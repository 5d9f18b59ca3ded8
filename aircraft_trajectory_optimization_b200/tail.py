'''
End-of-track rows of OPEN (non-periodic) racelines: the handful of rows of g that are neither produced by an
interval cell nor affine / sums of squares.

Reference (global-frame racelines; the parametric ones cannot be opened in the reference either, their helper
functions would need the centerline spline inside the SX graph, base_centerline.py:117-154):

  * `_enforce_initial_constraints` / `_enforce_terminal_constraints`: vg'vg <= 0 at Z[0,0] and at the end state zF
    (base_raceline.py:516-543), the thrust axis R[:,2] = e3 and zero body rates for drones
    (drone_raceline.py:110-148), purely vertical thrust for the point mass (point_raceline.py:15-45);
  * the gate at the very end of the track, `_fix_gate(zF[:3], ...)` (base_raceline.py:914-918);
  * zF / uF: the RK4 step or the Legendre extrapolation of the LAST interval (base_raceline.py:322-348),
    through the quaternion renormalisation (drone_raceline.py:42-45).

These rows are nonlinear functions of the end state, i.e. compositions phi(F(z, u, h)) whose Jacobian and Hessian
entries CasADi derives from the whole SX graph.  They are a dozen rows per problem, so instead of a generated kernel
per (vehicle, transcription, K) they are differentiated here at problem-construction time with the same scalar
graph engine the code generator uses (symbolic.py: CasADi-SX folding rules, hence CasADi's structural pattern) and
handed to the library as a register-allocated *tape* that `tail_tape_kernel` (csrc/tail_tape.cuh) interprets, one
thread per problem instance, after the cell kernels.  The tape is ordered in three phases (g values; + Jacobian
entries; + Hessian entries) so that a g-only evaluation stops after the first.
'''
import numpy as np

from . import symbolic as sx
from .models import zdot, end_terms, NFC

# tape opcodes (csrc/tail_tape.cuh)
T_CONST, T_LOADX, T_LOADVP, T_LOADLAM = 0, 1, 2, 3
T_ADD, T_SUB, T_MUL, T_DIV, T_NEG, T_SQ, T_SQRT, T_SIN, T_COS, T_TAN = range(4, 14)
T_STORE_G, T_STORE_J, T_STORE_H, T_ADD_H = 14, 15, 16, 17
_OP_OF = {sx.ADD: T_ADD, sx.SUB: T_SUB, sx.MUL: T_MUL, sx.DIV: T_DIV, sx.NEG: T_NEG, sx.SQ: T_SQ,
          sx.SQRT: T_SQRT, sx.SIN: T_SIN, sx.COS: T_COS, sx.TAN: T_TAN}
K_X, K_VP, K_LAM = 0, 1, 2


class TailRows:
    ''' rows of g given as scalar expressions of a few decision variables (and the vehicle parameters) '''

    def __init__(self, sb, variant):
        self.sb = sb
        self.variant = variant
        self.g = sx.Graph()
        self.inputs = []            # input slot -> (kind, index)
        self._w = {}
        self._vp = {}
        self.blocks = []            # (rows, [node ids]) in the order the reference appends them

    # ---- symbols ---------------------------------------------------------------------------------------
    def use(self):
        sx.set_graph(self.g)

    def w(self, idx):
        idx = int(idx)
        if idx not in self._w:
            self._w[idx] = sx.SX(self.g.input(f'w{idx}'))
            self.inputs.append((K_X, idx))
        return self._w[idx]

    def vp(self):
        names = self.variant.vp_names
        for k in range(len(names)):
            if k not in self._vp:
                self._vp[k] = sx.SX(self.g.input(f'vp{k}'))
                self.inputs.append((K_VP, k))
        return [self._vp[k] for k in range(len(names))]

    def point(self, n, k):
        ''' (z, u, du) symbols of collocation point (n, k) '''
        sb = self.sb
        return ([self.w(sb.iZ(n, k, i)) for i in range(sb.nz)], [self.w(sb.iU(n, k, j)) for j in range(sb.nu)],
                [self.w(sb.idU(n, k, j)) for j in range(sb.nu)])

    # ---- end state of the last interval -------------------------------------------------------------------
    def _cont(self, z):
        ''' quaternion renormalisation (drone_raceline.py:42-45) '''
        v = self.variant
        if v.vehicle == 'drone' and v.orient == 'quat':
            z = list(z)
            nrm = sx.norm_2(z[3:7])
            z[3:7] = [e / nrm for e in z[3:7]]
        return z

    def zF(self, D=None):
        ''' base_raceline.py:322-336 '''
        self.use()
        sb = self.sb
        n = sb.N - 1
        nz = sb.nz
        if D is None:
            z, u, _ = self.point(n, 0)
            h = self.w(sb.iH(n))
            vp = self.vp()
            # frame constants of the interval's fixed path length (parametric models: numbers, folded like the
            # reference's numeric param_terms, base_raceline.py:1034-1048)
            fc = [0.0] * NFC if sb.fc is None else [float(v) for v in sb.fc[n * sb.P]]

            def f(zz):
                return zdot(self.variant, zz, u, fc, vp)
            k1 = f(z)
            k2 = f([z[i] + h / 2 * k1[i] for i in range(nz)])
            k3 = f([z[i] + h / 2 * k2[i] for i in range(nz)])
            k4 = f([z[i] + h * k3[i] for i in range(nz)])
            zn = [z[i] + h / 6 * (k1[i] + k2[i] * 2 + k3[i] * 2 + k4[i]) for i in range(nz)]
            return self._cont(zn)
        zn = [0] * nz
        for k in range(sb.K + 1):
            z, _, _ = self.point(n, k)
            zn = [zn[i] + z[i] * float(D[k]) for i in range(nz)]
        return self._cont(zn)

    def uF(self, D=None):
        ''' base_raceline.py:338-348 '''
        self.use()
        sb = self.sb
        n = sb.N - 1
        if D is None:
            _, u, du = self.point(n, 0)
            h = self.w(sb.iH(n))
            return [u[j] + du[j] * h for j in range(sb.nu)]
        un = [0] * sb.nu
        for k in range(sb.K + 1):
            _, u, _ = self.point(n, k)
            un = [un[j] + u[j] * float(D[k]) for j in range(sb.nu)]
        return un

    def terms(self, z, u):
        self.use()
        return end_terms(self.variant, z, u, self.vp())

    # ---- rows ------------------------------------------------------------------------------------------------
    def add_rows(self, exprs, lb, ub):
        ''' one block of rows (one `g += [...]` of the reference) '''
        self.use()
        exprs = list(exprs)
        rows = self.sb.alloc_rows(len(exprs), lb, ub)
        self.blocks.append((rows, [sx._id(e) for e in exprs]))
        return rows

    # ---- derivatives -----------------------------------------------------------------------------------------
    def entries(self):
        '''
        structural Jacobian / Hessian entries the way CasADi's whole-graph AD finds them (block by block: forward
        sparse sweep for jac_g; reverse sweep of lam' g_block followed by a forward sparse sweep of the adjoints,
        upper-triangular copy kept).  Returns (jr, jc), (hr, hc) as int64 arrays; the nodes are kept for tape().
        '''
        self.use()
        g = self.g
        wrt = {s.i: idx for idx, s in self._w.items()}
        self._lam = {}
        jac = {}
        hess = {}
        for rows, outs in self.blocks:
            nodes = g.reachable(outs)
            for r, d in zip(rows, g.forward_sparse(outs, wrt, nodes)):
                for c, v in d.items():
                    jac[(int(r), c)] = v
            seeds = []
            for r in rows:
                self._lam[int(r)] = g.input(f'lam{int(r)}')
                self.inputs.append((K_LAM, int(r)))
                seeds.append(self._lam[int(r)])
            badj = g.reverse(outs, seeds, nodes)
            vars_here = [n for n in nodes if g.op[n] == sx.INPUT and n in wrt]
            grads = [badj.get(n, g.zero) for n in vars_here]
            loc = {}
            for n, d in zip(vars_here, g.forward_sparse(grads, wrt)):
                i = wrt[n]
                for j, v in d.items():
                    if i <= j:
                        loc[(i, j)] = v
                    else:
                        loc.setdefault((j, i), v)
            for key, v in loc.items():
                hess.setdefault(key, []).append(v)
        self._jac_keys = sorted(jac, key=lambda rc: (rc[1], rc[0]))
        self._jac_nodes = [jac[k] for k in self._jac_keys]
        self._hess_keys = sorted(hess, key=lambda rc: (rc[1], rc[0]))
        self._hess_nodes = [g.sum(hess[k]) for k in self._hess_keys]
        A = lambda v: np.array(v, dtype=np.int64).reshape(-1, 2)
        jk, hk = A(self._jac_keys), A(self._hess_keys)
        return (jk[:, 0], jk[:, 1]), (hk[:, 0], hk[:, 1])

    def tape(self, jslot, hslot, hadd):
        '''
        jslot / hslot: CCS positions of the entries returned by entries(); hadd: which Hessian positions already hold
        a contribution of the cell / simple-row kernels (the tape adds there and assigns elsewhere).

        The tape is *levelised*: instructions of one level are independent of each other (the threads of a CTA run
        them side by side, one barrier per level), levels are grouped in three phases (g values; Jacobian entries;
        Hessian entries), each closed by a level of stores.  Work slots are reused once the level of the last reader
        has finished.  Returns dict(ins int32 [n][4] = (op, a, b, dst), lvl_ptr int32 [n_levels + 1], cval,
        n_slots, n_levels = (after g, after jac, after hess)).
        '''
        g = self.g
        op, a, b = g.op, g.a, g.b
        g_out = [(T_STORE_G, n, int(r)) for rows, outs in self.blocks for r, n in zip(rows, outs)]
        j_out = [(T_STORE_J, n, int(s)) for n, s in zip(self._jac_nodes, jslot)]
        h_out = [(T_ADD_H if ad else T_STORE_H, n, int(s)) for n, s, ad in zip(self._hess_nodes, hslot, hadd)]

        def operands(n):
            o = op[n]
            if o < sx.ADD:
                return ()
            return (a[n], b[n]) if o <= sx.DIV else (a[n],)

        # ---- levels: list of lists of ('n', node) | ('s', opcode, node, dest) ------------------------------------
        levels = []
        n_levels = []
        done = set()
        for outs in (g_out, j_out, h_out):
            lvl = {}
            by_level = {}
            for n in g.reachable([o[1] for o in outs]):
                if n in done:
                    continue
                done.add(n)
                lv = 0
                for s_ in operands(n):
                    if s_ in lvl:
                        lv = max(lv, lvl[s_] + 1)
                lvl[n] = lv
                by_level.setdefault(lv, []).append(('n', n))
            for lv in sorted(by_level):
                levels.append(by_level[lv])
            if outs:
                levels.append([('s', *o) for o in outs])
            n_levels.append(len(levels))

        # ---- liveness at level granularity ---------------------------------------------------------------------------
        last = {}
        for li, items in enumerate(levels):
            for e in items:
                if e[0] == 'n':
                    for s_ in operands(e[1]):
                        last[s_] = li
                else:
                    last[e[2]] = li
        slot = {}
        free = []
        n_slots = 0
        cval = []
        ins = []
        lvl_ptr = [0]
        for li, items in enumerate(levels):
            dying = set()
            # same opcodes next to each other: the 32 lanes of a warp then take the same branch of the interpreter
            items = sorted(items, key=lambda e: (1, e[1]) if e[0] == 's' else (0, op[e[1]]))
            for e in items:
                if e[0] == 's':
                    _, code, n, dest = e
                    ins.append((code, slot[n], dest, 0))
                    if last[n] == li:
                        dying.add(n)
                    continue
                n = e[1]
                srcs = operands(n)
                for s_ in srcs:
                    if last[s_] == li:
                        dying.add(s_)
                if free:
                    dst = free.pop()
                else:
                    dst = n_slots
                    n_slots += 1
                slot[n] = dst
                o = op[n]
                if o == sx.CONST:
                    ins.append((T_CONST, len(cval), 0, dst))
                    cval.append(g.cval[n])
                elif o == sx.INPUT:
                    kind, idx = self.inputs[a[n]]
                    ins.append(((T_LOADX, T_LOADVP, T_LOADLAM)[kind], idx, 0, dst))
                else:
                    ins.append((_OP_OF[o], slot[srcs[0]], slot[srcs[1]] if len(srcs) > 1 else 0, dst))
            for n in dying:
                free.append(slot.pop(n))
            lvl_ptr.append(len(ins))
        return dict(ins=np.ascontiguousarray(np.array(ins, dtype=np.int32).reshape(-1, 4)),
                    lvl_ptr=np.ascontiguousarray(np.array(lvl_ptr, dtype=np.int32)),
                    cval=np.ascontiguousarray(np.array(cval if cval else [0.0], dtype=np.float64)),
                    n_slots=max(n_slots, 1), n_levels=tuple(n_levels))

    # ---- numpy evaluation of the tape (host-side check of the emitter; the product path is the CUDA kernel) ------
    @staticmethod
    def run_tape(tp, x, vp, lam, g, jac, hess, phase=2):
        W = np.zeros(tp['n_slots'])
        for code, a_, b_, dst in tp['ins'][:tp['lvl_ptr'][tp['n_levels'][phase]]]:
            if code == T_CONST:
                W[dst] = tp['cval'][a_]
            elif code == T_LOADX:
                W[dst] = x[a_]
            elif code == T_LOADVP:
                W[dst] = vp[a_]
            elif code == T_LOADLAM:
                W[dst] = lam[a_] if lam is not None else 0.0
            elif code == T_ADD:
                W[dst] = W[a_] + W[b_]
            elif code == T_SUB:
                W[dst] = W[a_] - W[b_]
            elif code == T_MUL:
                W[dst] = W[a_] * W[b_]
            elif code == T_DIV:
                W[dst] = W[a_] / W[b_]
            elif code == T_NEG:
                W[dst] = -W[a_]
            elif code == T_SQ:
                W[dst] = W[a_] * W[a_]
            elif code == T_SQRT:
                W[dst] = np.sqrt(W[a_])
            elif code == T_SIN:
                W[dst] = np.sin(W[a_])
            elif code == T_COS:
                W[dst] = np.cos(W[a_])
            elif code == T_TAN:
                W[dst] = np.tan(W[a_])
            elif code == T_STORE_G:
                g[b_] = W[a_]
            elif code == T_STORE_J:
                jac[b_] = W[a_]
            elif code == T_STORE_H:
                hess[b_] = W[a_]
            elif code == T_ADD_H:
                hess[b_] += W[a_]
