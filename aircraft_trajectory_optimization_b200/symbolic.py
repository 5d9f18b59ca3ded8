'''
Scalar expression-DAG engine used at BUILD time.

What it is for
--------------
* the code generator (`codegen.py`) builds the vehicle ODE right-hand sides with it, differentiates
  them and emits straight-line fp64 CUDA for the sm_100a kernels;
* the CPU oracle under `oracle/` uses the same node store to restate the reference's fully
  unrolled NLP and to differentiate it as a whole graph.

The reference hands CasADi `SX` graphs to `ca.nlpsol` (drone3d/raceline/base_raceline.py:752-799);
CasADi is not installed here, so this module plays the role of the SX layer.  It follows the
construction-time folding rules CasADi's SX applies (0*x -> 0, x+0 -> x, 1*x -> x, x-x -> 0,
x/x -> 1, constants folded) because those rules decide the *structural* sparsity pattern that
`nlp_jac_g` / `nlp_hess_l` expose.

Nodes are plain integers into parallel Python lists (op, a, b); identical sub-expressions are
hash-consed, so common-subexpression elimination is automatic.  Ids are topologically ordered
(children < parent), which every sweep below relies on.
'''
import math

import numpy as _np

_ND = _np.ndarray

CONST, INPUT, ADD, SUB, MUL, DIV, NEG, SQ, SQRT, SIN, COS, TAN = range(12)
OP_NAMES = ['const', 'input', 'add', 'sub', 'mul', 'div', 'neg', 'sq', 'sqrt', 'sin', 'cos', 'tan']
_BINARY = (ADD, SUB, MUL, DIV)
_UNARY = (NEG, SQ, SQRT, SIN, COS, TAN)


class Graph:
    ''' node store with hash-consing and CasADi-SX-like folding '''

    def __init__(self):
        self.op = []
        self.a = []
        self.b = []
        self.cval = {}      # node id -> float (CONST nodes)
        self._ctab = {}     # float -> node id
        self._tab = {}      # (op, a, b) -> node id
        self.input_names = []   # INPUT node k has a == k
        self.input_ids = []
        self.zero = self.const(0.0)
        self.one = self.const(1.0)
        self.two = self.const(2.0)
        self.mone = self.const(-1.0)
        self.half = self.const(0.5)

    def __len__(self):
        return len(self.op)

    # ---- node creation -------------------------------------------------------------------
    def const(self, v):
        v = float(v)
        if v == 0.0:
            v = 0.0  # merge -0.0 with 0.0 (CasADi tests value == 0)
        i = self._ctab.get(v)
        if i is None:
            i = len(self.op)
            self.op.append(CONST)
            self.a.append(-1)
            self.b.append(-1)
            self.cval[i] = v
            self._ctab[v] = i
        return i

    def input(self, name):
        i = len(self.op)
        self.op.append(INPUT)
        self.a.append(len(self.input_names))
        self.b.append(-1)
        self.input_names.append(name)
        self.input_ids.append(i)
        return i

    def _node(self, op, a, b=-1):
        key = (op, a, b)
        i = self._tab.get(key)
        if i is None:
            i = len(self.op)
            self.op.append(op)
            self.a.append(a)
            self.b.append(b)
            self._tab[key] = i
        return i

    def is_const(self, i):
        return self.op[i] == CONST

    # ---- arithmetic with folding ---------------------------------------------------------
    def add(self, x, y):
        op = self.op
        if op[x] == CONST:
            if op[y] == CONST:
                return self.const(self.cval[x] + self.cval[y])
            if x == self.zero:
                return y
        elif op[y] == CONST and y == self.zero:
            return x
        if op[y] == NEG:
            return self.sub(x, self.a[y])
        if op[x] == NEG:
            return self.sub(y, self.a[x])
        if x > y:
            x, y = y, x
        return self._node(ADD, x, y)

    def sub(self, x, y):
        op = self.op
        if x == y:
            return self.zero
        if op[y] == CONST:
            if op[x] == CONST:
                return self.const(self.cval[x] - self.cval[y])
            if y == self.zero:
                return x
        elif op[x] == CONST and x == self.zero:
            return self.neg(y)
        if op[y] == NEG:
            return self.add(x, self.a[y])
        return self._node(SUB, x, y)

    def mul(self, x, y):
        op = self.op
        if op[x] == CONST:
            if op[y] == CONST:
                return self.const(self.cval[x] * self.cval[y])
            if x == self.zero:
                return x
            if x == self.one:
                return y
            if x == self.mone:
                return self.neg(y)
        elif op[y] == CONST:
            if y == self.zero:
                return y
            if y == self.one:
                return x
            if y == self.mone:
                return self.neg(x)
        if x == y:
            return self.sq(x)
        if op[x] == NEG and op[y] == NEG:
            x, y = self.a[x], self.a[y]
        if x > y:
            x, y = y, x
        return self._node(MUL, x, y)

    def div(self, x, y):
        op = self.op
        if op[x] == CONST:
            if op[y] == CONST:
                return self.const(self.cval[x] / self.cval[y])
            if x == self.zero:
                return x
        elif op[y] == CONST:
            if y == self.one:
                return x
            if y == self.mone:
                return self.neg(x)
        if x == y:
            return self.one
        return self._node(DIV, x, y)

    def neg(self, x):
        if self.op[x] == CONST:
            return self.const(-self.cval[x])
        if self.op[x] == NEG:
            return self.a[x]
        if self.op[x] == SUB:
            return self.sub(self.b[x], self.a[x])
        return self._node(NEG, x)

    def sq(self, x):
        if self.op[x] == CONST:
            return self.const(self.cval[x] * self.cval[x])
        if self.op[x] == NEG:
            x = self.a[x]
        return self._node(SQ, x)

    def sqrt(self, x):
        if self.op[x] == CONST:
            return self.const(math.sqrt(self.cval[x]))
        return self._node(SQRT, x)

    def sin(self, x):
        if self.op[x] == CONST:
            return self.const(math.sin(self.cval[x]))
        return self._node(SIN, x)

    def cos(self, x):
        if self.op[x] == CONST:
            return self.const(math.cos(self.cval[x]))
        return self._node(COS, x)

    def tan(self, x):
        if self.op[x] == CONST:
            return self.const(math.tan(self.cval[x]))
        return self._node(TAN, x)

    def powi(self, x, n):
        ''' integer power by squaring (x**2 -> sq, like SX) '''
        if n == 0:
            return self.one
        if n < 0:
            return self.div(self.one, self.powi(x, -n))
        if n == 1:
            return x
        if n % 2 == 0:
            return self.sq(self.powi(x, n // 2))
        return self.mul(x, self.powi(x, n - 1))

    def sum(self, ids):
        ''' left-to-right sum '''
        acc = self.zero
        for i in ids:
            acc = self.add(acc, i)
        return acc

    # ---- graph queries -------------------------------------------------------------------
    def reachable(self, outputs):
        ''' sorted (ascending = topological) list of nodes the outputs depend on '''
        op, a, b = self.op, self.a, self.b
        seen = set()
        stack = [o for o in outputs]
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            o = op[n]
            if o >= ADD:
                x = a[n]
                if x not in seen:
                    stack.append(x)
                if o <= DIV:
                    y = b[n]
                    if y not in seen:
                        stack.append(y)
        return sorted(seen)

    # ---- differentiation -----------------------------------------------------------------
    def partials(self, n):
        ''' (d n / d a, d n / d b) as node ids, CasADi-style (derivatives may reuse n itself) '''
        o = self.op[n]
        x = self.a[n]
        if o == ADD:
            return self.one, self.one
        if o == SUB:
            return self.one, self.mone
        y = self.b[n]
        if o == MUL:
            return y, x
        if o == DIV:
            return self.div(self.one, y), self.neg(self.div(n, y))
        if o == NEG:
            return self.mone, None
        if o == SQ:
            return self.mul(self.two, x), None
        if o == SQRT:
            return self.div(self.one, self.mul(self.two, n)), None
        if o == SIN:
            return self.cos(x), None
        if o == COS:
            return self.neg(self.sin(x)), None
        if o == TAN:
            return self.div(self.one, self.sq(self.cos(x))), None
        raise ValueError(f'no partials for op {o}')

    def reverse(self, outputs, seeds, nodes=None):
        '''
        reverse-mode sweep: returns {node id: adjoint node id} for every reachable node
        (callers pick the INPUT nodes they want).  adjoint of outputs[k] is seeded with seeds[k].
        '''
        if nodes is None:
            nodes = self.reachable(outputs)
        op, a, b = self.op, self.a, self.b
        adj = {}
        add, mul = self.add, self.mul
        for o, s in zip(outputs, seeds):
            adj[o] = add(adj[o], s) if o in adj else s
        zero = self.zero
        for n in reversed(nodes):
            o = op[n]
            if o < ADD:
                continue
            nb = adj.get(n)
            if nb is None or nb == zero:
                continue
            x = a[n]
            if o == ADD:
                y = b[n]
                adj[x] = add(adj[x], nb) if x in adj else nb
                adj[y] = add(adj[y], nb) if y in adj else nb
            elif o == SUB:
                y = b[n]
                adj[x] = add(adj[x], nb) if x in adj else nb
                m = self.neg(nb)
                adj[y] = add(adj[y], m) if y in adj else m
            else:
                da, db = self.partials(n)
                t = mul(nb, da)
                adj[x] = add(adj[x], t) if x in adj else t
                if db is not None:
                    y = b[n]
                    t = mul(nb, db)
                    adj[y] = add(adj[y], t) if y in adj else t
        return adj

    def forward_sparse(self, outputs, wrt, nodes=None, seeds=None):
        '''
        forward-mode sweep carrying a sparse gradient {column: node id} per node.
        wrt: {INPUT node id: column}.  seeds: optional {INPUT node id: seed node id} (default 1);
        mapping several inputs to one column with symbolic seeds gives a directional derivative.
        Returns a list of dicts, one per output.
        '''
        if nodes is None:
            nodes = self.reachable(outputs)
        op, a, b = self.op, self.a, self.b
        add, mul, sub, neg = self.add, self.mul, self.sub, self.neg
        D = {}
        empty = {}
        one = self.one
        for n in nodes:
            o = op[n]
            if o == CONST:
                continue
            if o == INPUT:
                c = wrt.get(n)
                if c is not None:
                    D[n] = {c: one if seeds is None else seeds.get(n, one)}
                continue
            x = a[n]
            dx = D.get(x, empty)
            if o <= DIV:
                y = b[n]
                dy = D.get(y, empty)
                if not dx and not dy:
                    continue
                if o == ADD:
                    if not dy:
                        D[n] = dx
                    elif not dx:
                        D[n] = dy
                    else:
                        r = dict(dx)
                        for c, v in dy.items():
                            r[c] = add(r[c], v) if c in r else v
                        D[n] = r
                elif o == SUB:
                    if not dy:
                        D[n] = dx
                    else:
                        r = dict(dx)
                        for c, v in dy.items():
                            r[c] = sub(r[c], v) if c in r else neg(v)
                        D[n] = r
                else:
                    da, db = self.partials(n)
                    r = {c: mul(da, v) for c, v in dx.items()}
                    for c, v in dy.items():
                        t = mul(db, v)
                        r[c] = add(r[c], t) if c in r else t
                    D[n] = r
            else:
                if not dx:
                    continue
                da, _ = self.partials(n)
                D[n] = {c: mul(da, v) for c, v in dx.items()}
        return [D.get(o, empty) for o in outputs]

    def dependencies(self, outputs, wrt, nodes=None):
        '''
        structural dependency propagation (what CasADi's sparsity sweeps do): returns, per output,
        a Python int used as a bitset over the columns in wrt ({INPUT node id: column}).
        '''
        if nodes is None:
            nodes = self.reachable(outputs)
        op, a, b = self.op, self.a, self.b
        dep = {}
        for n in nodes:
            o = op[n]
            if o == CONST:
                continue
            if o == INPUT:
                c = wrt.get(n)
                if c is not None:
                    dep[n] = 1 << c
                continue
            d = dep.get(a[n], 0)
            if o <= DIV:
                d |= dep.get(b[n], 0)
            if d:
                dep[n] = d
        return [dep.get(o, 0) for o in outputs]

    # ---- evaluation ----------------------------------------------------------------------
    def evaluate(self, outputs, input_values, nodes=None):
        '''
        interpret the graph.  input_values: sequence indexed by INPUT slot (node.a); entries may
        be floats or numpy arrays (broadcast).  Returns a list of values, one per output.
        '''
        import numpy as np
        if nodes is None:
            nodes = self.reachable(outputs)
        op, a, b, cval = self.op, self.a, self.b, self.cval
        val = {}
        for n in nodes:
            o = op[n]
            if o == CONST:
                val[n] = cval[n]
            elif o == INPUT:
                val[n] = input_values[a[n]]
            elif o == ADD:
                val[n] = val[a[n]] + val[b[n]]
            elif o == SUB:
                val[n] = val[a[n]] - val[b[n]]
            elif o == MUL:
                val[n] = val[a[n]] * val[b[n]]
            elif o == DIV:
                val[n] = val[a[n]] / val[b[n]]
            elif o == NEG:
                val[n] = -val[a[n]]
            elif o == SQ:
                v = val[a[n]]
                val[n] = v * v
            elif o == SQRT:
                val[n] = np.sqrt(val[a[n]])
            elif o == SIN:
                val[n] = np.sin(val[a[n]])
            elif o == COS:
                val[n] = np.cos(val[a[n]])
            elif o == TAN:
                val[n] = np.tan(val[a[n]])
        return [val[o] for o in outputs]

    # ---- code emission -------------------------------------------------------------------
    def emit_c(self, outputs, input_exprs, out_exprs, indent='  ', tmp='t', real='double'):
        '''
        straight-line C/CUDA for `outputs`.
        input_exprs: list indexed by INPUT slot -> C expression string to read that input.
        out_exprs:   list of C lvalue strings, one per output.
        Returns (code string, n_ops) where n_ops counts arithmetic instructions.
        '''
        nodes = self.reachable(outputs)
        op, a, b, cval = self.op, self.a, self.b, self.cval
        name = {}
        lines = []
        n_ops = 0

        def lit(v):
            if v == int(v) and abs(v) < 1e15:
                return f'{int(v)}.0'
            return repr(v)

        for n in nodes:
            o = op[n]
            if o == CONST:
                v = cval[n]
                name[n] = lit(v) if v >= 0 else f'({lit(v)})'
                continue
            if o == INPUT:
                name[n] = input_exprs[a[n]]
                continue
            x = name[a[n]]
            if o == ADD:
                e = f'{x} + {name[b[n]]}'
            elif o == SUB:
                e = f'{x} - {name[b[n]]}'
            elif o == MUL:
                e = f'{x} * {name[b[n]]}'
            elif o == DIV:
                e = f'{x} / {name[b[n]]}'
            elif o == NEG:
                e = f'-{x}'
            elif o == SQ:
                e = f'{x} * {x}'
            elif o == SQRT:
                e = f'sqrt({x})'
            elif o == SIN:
                e = f'sin({x})'
            elif o == COS:
                e = f'cos({x})'
            elif o == TAN:
                e = f'tan({x})'
            n_ops += 1
            name[n] = f'{tmp}{n}'
            lines.append(f'{indent}const {real} {tmp}{n} = {e};')
        for o, dst in zip(outputs, out_exprs):
            lines.append(f'{indent}{dst} = {name[o]};')
        return '\n'.join(lines), n_ops


# ------------------------------------------------------------------------------------------
# operator-overloading scalar handle (for writing model equations naturally)
# ------------------------------------------------------------------------------------------
_G = Graph()


def graph() -> Graph:
    ''' the current graph '''
    return _G


def new_graph() -> Graph:
    ''' start a fresh graph and make it current '''
    global _G
    _G = Graph()
    return _G


def set_graph(g: Graph):
    global _G
    _G = g


class SX:
    ''' scalar symbolic handle; numpy object arrays of these give vectors / matrices '''
    __slots__ = ('i',)

    def __init__(self, i):
        self.i = i

    @staticmethod
    def sym(name, n=None):
        import numpy as np
        if n is None:
            return SX(_G.input(name))
        return np.array([SX(_G.input(f'{name}_{k}')) for k in range(n)], dtype=object)

    @staticmethod
    def const(v):
        return SX(_G.const(v))

    def is_const(self):
        return _G.op[self.i] == CONST

    def value(self):
        return _G.cval[self.i]

    def __add__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.add(self.i, _id(o)))

    def __radd__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.add(_id(o), self.i))

    def __sub__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.sub(self.i, _id(o)))

    def __rsub__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.sub(_id(o), self.i))

    def __mul__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.mul(self.i, _id(o)))

    def __rmul__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.mul(_id(o), self.i))

    def __truediv__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.div(self.i, _id(o)))

    def __rtruediv__(self, o):
        if isinstance(o, _ND):
            return NotImplemented
        return SX(_G.div(_id(o), self.i))

    def __neg__(self):
        return SX(_G.neg(self.i))

    def __pos__(self):
        return self

    def __pow__(self, n):
        if isinstance(n, float) and n == 0.5:
            return SX(_G.sqrt(self.i))
        if int(n) != n:
            raise NotImplementedError('only integer powers')
        return SX(_G.powi(self.i, int(n)))

    def __repr__(self):
        if self.is_const():
            return f'SX({self.value()})'
        return f'SX(#{self.i}:{OP_NAMES[_G.op[self.i]]})'


def _id(o):
    if isinstance(o, SX):
        return o.i
    return _G.const(o)


def sqrt(x):
    return SX(_G.sqrt(_id(x)))


def sin(x):
    return SX(_G.sin(_id(x)))


def cos(x):
    return SX(_G.cos(_id(x)))


def tan(x):
    return SX(_G.tan(_id(x)))


def norm_2(v):
    ''' sqrt of the sum of squares, summed left to right '''
    acc = 0
    for e in v:
        acc = acc + e * e
    return sqrt(acc)


def ids(arr):
    ''' flatten an SX / array-of-SX / number container into a list of node ids '''
    import numpy as np
    if isinstance(arr, SX):
        return [arr.i]
    out = []
    for e in np.asarray(arr, dtype=object).ravel():
        out.append(_id(e))
    return out
