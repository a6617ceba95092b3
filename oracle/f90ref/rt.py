"""Runtime of the f90py-translated reference (TEST INFRASTRUCTURE).

Fortran semantics the generated code relies on: arrays with declared lower bounds and
bounds checking (FArr), sections (S), float32 default REAL, truncating integer division,
integer powers by repeated multiplication, blank-padded string comparison, and the two
processor-dependent intrinsic families, which are *bound from outside*:

* ``rng``  -- RANDOM_NUMBER.  The reference seeds the intrinsic generator from the wall
  clock (photon_mod.f90:68-87); the harness installs the oracle's per-packet Philox
  stream here so that both sides consume the same uniforms in the same order.
* ``math`` -- LOG/EXP/SIN/COS/ACOS/ATAN of default REAL.  ``use_libm()`` binds numpy's
  float32 routines (what a real build would call, up to the platform's libm);
  ``use_detmath(lib)`` binds the oracle's fully specified detmath (<= 1 ulp from libm)
  so that branch decisions can be compared bit for bit.
"""
from __future__ import annotations

import math
import sys

import numpy as np

f32 = np.float32
f64 = np.float64
c64 = np.complex64
c128 = np.complex128
ZERO32 = np.float32(0.0)
ZERO64 = np.float64(0.0)
UNINIT_INT = 0          # value of an uninitialised integer local (processor dependent)
QUIET = True            # drop PRINT output
_DT = {'i': np.int64, 'r': np.float32, 'd': np.float64, 'l': np.bool_, 'c': object, 't': object,
       'z': np.complex64, 'Z': np.complex128}


class FortranStop(Exception):
    def __init__(self, proc, line):
        super().__init__(f'STOP in {proc} at line {line}')
        self.proc, self.line = proc, line


class FortranBoundsError(IndexError):
    pass


class S:
    """array section triplet lo:hi:step (None = omitted)"""
    __slots__ = ('lo', 'hi', 'st')

    def __init__(self, lo, hi, st):
        self.lo, self.hi, self.st = lo, hi, st


class FArr:
    """numpy array + Fortran lower bounds.  Element access with integers returns a scalar
    (Python int for integer arrays, numpy scalar otherwise); any S subscript returns a view
    with lower bounds 1."""
    __slots__ = ('a', 'lb', 'isint', 'nd')

    def __init__(self, a, lb=None):
        self.a = a
        self.nd = a.ndim
        self.lb = tuple(lb) if lb is not None else (1,) * a.ndim
        self.isint = a.dtype.kind in 'iu'

    def _index(self, key):
        if type(key) is not tuple:
            key = (key,)
        if len(key) != self.nd:
            raise FortranBoundsError(f'rank mismatch: {len(key)} subscripts for rank {self.nd}')
        idx = []
        sec = False
        shape = self.a.shape
        for k, l, n in zip(key, self.lb, shape):
            if type(k) is S:
                sec = True
                st = 1 if k.st is None else int(k.st)
                if st <= 0:
                    raise FortranBoundsError('non-positive section stride not supported')
                lo = 0 if k.lo is None else int(k.lo) - l
                hi = n if k.hi is None else int(k.hi) - l + 1
                if hi > lo and (lo < 0 or hi > n):
                    raise FortranBoundsError(f'section {k.lo}:{k.hi} outside bounds {l}:{l + n - 1}')
                idx.append(slice(lo, max(hi, lo), st))
            else:
                i = k - l
                if i < 0 or i >= n:
                    raise FortranBoundsError(f'subscript {k} outside bounds {l}:{l + n - 1}')
                idx.append(i)
        return tuple(idx), sec

    def __getitem__(self, key):
        idx, sec = self._index(key)
        v = self.a[idx]
        if sec:
            return FArr(v)
        return int(v) if self.isint else v

    def __setitem__(self, key, val):
        idx, sec = self._index(key)
        self.a[idx] = val.a if type(val) is FArr else val

    def setall(self, val):
        self.a[...] = val.a if type(val) is FArr else val

    # elementwise arithmetic / comparison
    def _b(self, o, op, rev=False):
        ob = o.a if type(o) is FArr else o
        return FArr(op(ob, self.a) if rev else op(self.a, ob))

    def __add__(self, o): return self._b(o, np.add)
    def __radd__(self, o): return self._b(o, np.add, True)
    def __sub__(self, o): return self._b(o, np.subtract)
    def __rsub__(self, o): return self._b(o, np.subtract, True)
    def __mul__(self, o): return self._b(o, np.multiply)
    def __rmul__(self, o): return self._b(o, np.multiply, True)
    def __truediv__(self, o): return self._b(o, np.divide)
    def __rtruediv__(self, o): return self._b(o, np.divide, True)
    def __neg__(self): return FArr(-self.a)
    def __abs__(self): return FArr(np.abs(self.a))
    def __lt__(self, o): return self._b(o, np.less)
    def __le__(self, o): return self._b(o, np.less_equal)
    def __gt__(self, o): return self._b(o, np.greater)
    def __ge__(self, o): return self._b(o, np.greater_equal)
    def __eq__(self, o): return self._b(o, np.equal)
    def __ne__(self, o): return self._b(o, np.not_equal)
    __hash__ = None

    def __repr__(self):
        return f'FArr(lb={self.lb}, {self.a!r})'


def wrap(a, lb=None):
    """numpy array (Fortran order expected for rank>1) -> FArr sharing its memory"""
    return FArr(a, lb)


def alloc(kind, dims):
    shape = tuple(max(int(h) - int(l) + 1, 0) for l, h in dims)
    a = np.zeros(shape, dtype=_DT[kind], order='F')
    if kind == 'c':
        a[...] = ' '
    return FArr(a, tuple(int(l) for l, _ in dims))


def alloc_obj(cls, dims):
    shape = tuple(max(int(h) - int(l) + 1, 0) for l, h in dims)
    a = np.empty(shape, dtype=object, order='F')
    for i in np.ndindex(*shape):
        a[i] = cls()
    return FArr(a, tuple(int(l) for l, _ in dims))


def copy_arr(x):
    if x is None:
        return None
    if type(x) is FArr:
        return FArr(x.a.copy(order='F'), x.lb)
    return FArr(np.array(x, order='F'))


def rebase(x, lbs):
    """dummy array argument: same storage, the dummy's declared lower bounds"""
    if x is None:
        return None
    if x.lb == lbs:
        return x
    if len(lbs) != x.nd:
        raise FortranBoundsError('rank mismatch between actual and dummy array')
    return FArr(x.a, lbs)


def arrcon(items, kind):
    flat = []
    for it in items:
        if type(it) is FArr:
            flat.extend(it.a.ravel(order='F').tolist())
        else:
            flat.append(it)
    return FArr(np.array(flat, dtype=_DT[kind]))


def frange(lo, hi, st):
    lo, hi, st = int(lo), int(hi), int(st)
    if st > 0:
        return range(lo, hi + 1, st)
    return range(lo, hi - 1, st)


# ---- hooks installed by the harness -------------------------------------------------------
def _no_hook(*a):
    return None


loop_hook = _no_hook
proc_hook = _no_hook


class _NoRng:
    def next(self):
        raise RuntimeError('RANDOM_NUMBER called but no generator is installed (rt.rng)')


rng = _NoRng()


def random_number():
    return rng.next()


def random_fill(arr):
    flat = arr.a.reshape(-1, order='F')
    for i in range(flat.shape[0]):
        flat[i] = rng.next()
    arr.a[...] = flat.reshape(arr.a.shape, order='F')


def fprint(line, *items):
    if not QUIET:
        print(f'[ref:{line}]', *items, file=sys.stderr)


externs = {}      # name -> callable, supplied by the harness (see f90py.Gen externs)


def call_extern(name):
    if name not in externs:
        raise NotImplementedError(f'external routine {name} was not supplied by the harness')
    return externs[name]()


def mpi_allreduce_single(send, recv):
    """MPI_ALLREDUCE(sum) on one rank"""
    recv.setall(send)


io_log = {}       # unit -> list of records (tuples of the values of a list-directed WRITE)


def fwrite(unit, items):
    flat = []
    for it in items:             # a whole array is written element by element, in array element order
        if type(it) is FArr:
            flat.extend(int(v) if it.isint else v for v in it.a.ravel(order='F'))
        else:
            flat.append(it)
    io_log.setdefault(int(unit), []).append(tuple(flat))


# ---- list-directed READ ------------------------------------------------------------------------
# The harness binds a unit number to the text of the file the reference would OPEN on it
# (bind_unit); OPEN / CLOSE themselves do nothing.  A READ statement starts on a new record
# (line), takes as many values as it has items -- continuing over the following records when
# a line runs out -- and discards what is left of the last record it touched.  Values are separated
# by blanks or commas; '...' / "..." delimit character values; end of file raises FortranEOF, which
# the generated code turns into iostat = -1 (or lets propagate when the statement has no iostat).
class FortranEOF(Exception):
    pass


class _InUnit:
    def __init__(self, text):
        import shlex
        self.lines = []
        for raw in text.split('\n'):
            lex = shlex.shlex(raw, posix=True)
            lex.whitespace += ','
            lex.whitespace_split = True
            lex.commenters = ''
            try:
                toks = list(lex)
            except ValueError:
                toks = raw.replace(',', ' ').split()
            self.lines.append(toks)
        while self.lines and not self.lines[-1]:
            self.lines.pop()
        self.pos = 0


class _Reader:
    def __init__(self, u, line):
        self.u, self.line, self.col, self.started = u, line, 0, False

    def next(self, want):
        u = self.u
        while True:
            if u.pos >= len(u.lines):
                raise FortranEOF(f'end of file in READ at line {self.line}')
            row = u.lines[u.pos]
            self.started = True
            if self.col < len(row):
                tok = row[self.col]
                self.col += 1
                if want == 'c':
                    return tok
                return float(tok.lower().replace('d', 'e'))
            u.pos += 1
            self.col = 0

    def end(self):
        u = self.u
        if u.pos < len(u.lines) and (self.started or True):
            u.pos += 1                    # the rest of the record is skipped


io_in = {}        # unit -> _InUnit


def bind_unit(unit, text):
    io_in[int(unit)] = _InUnit(text)


def fread_begin(unit, line):
    u = io_in.get(int(unit))
    if u is None:
        raise NotImplementedError(f'READ on unit {unit} at line {line}: no file bound (rt.bind_unit)')
    return _Reader(u, line)


def fseek(unit, what):
    u = io_in.get(int(unit))
    if u is None:
        return
    if what == 'rewind':
        u.pos = 0
    elif u.pos > 0:
        u.pos -= 1


def fio(what):
    """OPEN / CLOSE: nothing to do, WRITE records are kept in io_log"""
    return None


def unsupported(what, line):
    raise NotImplementedError(f'untranslated statement reached at line {line}: {what}')


# ---- arithmetic helpers -----------------------------------------------------------------------
def idiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def ipow(x, n):
    """x**n for integer n the way gfortran expands it: repeated multiplication
    (x**2 = x*x, x**3 = x*x*x), reciprocal for negative n."""
    n = int(n)
    if n == 0:
        return x * 0 + 1
    m = abs(n)
    if m <= 3:
        r = x
        for _ in range(m - 1):
            r = r * x
    else:                       # __builtin_powi: square and multiply
        r = None
        base = x
        while m:
            if m & 1:
                r = base if r is None else r * base
            base = base * base
            m >>= 1
    if n < 0:
        return (x * 0 + 1) / r
    return r


def rpow(x, y):
    return np.power(x, y)


def f_int(x):
    if type(x) is FArr:
        return FArr(np.trunc(x.a).astype(np.int64))
    return int(x)


def f_nint(x):
    x = float(x)
    return int(x + 0.5) if x >= 0 else -int(-x + 0.5)


def f_floor(x):
    return int(math.floor(x))


def f_ceiling(x):
    return int(math.ceil(x))


def conv_i(x): return int(x)
def conv_r(x): return f32(x)
def conv_d(x): return f64(x)
def conv_l(x): return bool(x)


def _unw(x):
    return x.a if type(x) is FArr else x


def f_max(*a):
    if any(type(x) is FArr for x in a):
        r = _unw(a[0])
        for x in a[1:]:
            r = np.maximum(r, _unw(x))
        return FArr(r)
    r = a[0]
    for x in a[1:]:
        if x > r:
            r = x
    return r


def f_min(*a):
    if any(type(x) is FArr for x in a):
        r = _unw(a[0])
        for x in a[1:]:
            r = np.minimum(r, _unw(x))
        return FArr(r)
    r = a[0]
    for x in a[1:]:
        if x < r:
            r = x
    return r


def f_mod(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return a - b * idiv(a, b)
    return np.fmod(a, b)


def f_sign(a, b):
    return abs(a) if b >= 0 else -abs(a)


def f_size(x, dim=None):
    return int(x.a.size) if dim is None else int(x.a.shape[int(dim) - 1])


def f_lbound(x, dim):
    return x.lb[int(dim) - 1]


def f_ubound(x, dim):
    return x.lb[int(dim) - 1] + x.a.shape[int(dim) - 1] - 1


def _loc(x, dim, mask, fn):
    a = x.a
    if mask is not None:
        m = _unw(mask)
        if not m.any():
            return 0 if dim is not None else FArr(np.zeros(1, np.int64))
        idxs = np.flatnonzero(m)
        k = int(idxs[fn(a[idxs])])      # first occurrence among the masked elements
    else:
        if a.size == 0:
            return 0 if dim is not None else FArr(np.zeros(1, np.int64))
        k = int(fn(a))
    # result is relative to lower bound 1 whatever the array's bounds
    return k + 1 if dim is not None else FArr(np.array([k + 1], np.int64))


def f_minloc(x, dim=None, mask=None): return _loc(x, dim, mask, np.argmin)
def f_maxloc(x, dim=None, mask=None): return _loc(x, dim, mask, np.argmax)
def f_maxval(x): return x.a.max() if not x.isint else int(x.a.max())
def f_minval(x): return x.a.min() if not x.isint else int(x.a.min())


def f_sum(x):
    # sequential accumulation in the array's own precision, array element order
    flat = x.a.ravel(order='F')
    if x.isint:
        return int(flat.sum())
    acc = flat.dtype.type(0)
    for v in flat:
        acc = acc + v
    return acc


def f_product(x):
    flat = x.a.ravel(order='F')
    acc = flat.dtype.type(1)
    for v in flat:
        acc = acc * v
    return int(acc) if x.isint else acc


def f_any(x): return bool(_unw(x).any())
def f_all(x): return bool(_unw(x).all())
def f_count(x): return int(_unw(x).sum())
def f_trim(s): return s.rstrip(' ')
def f_adjustl(s): return s.lstrip(' ').ljust(len(s))
def f_len_trim(s): return len(s.rstrip(' '))
def f_len(s): return len(s)
def f_isnan(x): return bool(np.isnan(x))


def f_epsilon(x): return np.finfo(np.asarray(_unw(x)).dtype).eps
def f_huge(x):
    dt = np.asarray(_unw(x)).dtype
    return int(np.iinfo(np.int32).max) if dt.kind in 'iu' else np.finfo(dt).max
def f_tiny(x): return np.finfo(np.asarray(_unw(x)).dtype).tiny


def fstr(s, n):
    s = str(s)
    return s[:n] if len(s) >= n else s.ljust(n)


def substr(s, lo, hi):
    return s[int(lo) - 1:(len(s) if hi is None else int(hi))]


def strcmp(a, b, op):
    n = max(len(a), len(b))
    a, b = a.ljust(n), b.ljust(n)
    return {'==': a == b, '!=': a != b, '<': a < b, '<=': a <= b, '>': a > b, '>=': a >= b}[op]


# ---- processor-dependent math -------------------------------------------------------------------
class _Math:
    pass


math32 = _Math()
math64 = _Math()      # double-precision overrides (only what detmath defines), else numpy


def use_libm():
    """float32 elementary functions from numpy (the platform's single-precision routines)."""
    for n in ('log', 'exp', 'sin', 'cos', 'tan', 'log10', 'sinh', 'cosh', 'tanh'):
        setattr(math32, n, getattr(np, n))
    math32.acos, math32.asin, math32.atan = np.arccos, np.arcsin, np.arctan
    math32.name = 'libm'
    math64.__dict__.clear()


def use_detmath(lib):
    """bind LOG/SIN/COS/ACOS/ATAN/EXP of default REAL to oracle/detmath.h through the
    oracle's unit-test hook oracle_detmath (documented deviation 2 of the oracle)."""
    import ctypes as C

    fn = lib.oracle_detmath
    fn.argtypes = [C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int64]
    fn.restype = None

    def make(which):
        buf_in = (C.c_float * 1)()
        buf_out = (C.c_float * 1)()

        def f(x):
            buf_in[0] = x
            fn(which, buf_in, buf_out, 1)
            return np.float32(buf_out[0])
        return f
    use_libm()
    math32.log, math32.sin, math32.cos = make(0), make(1), make(2)
    math32.acos, math32.atan, math32.exp = make(3), make(4), make(5)
    math32.name = 'detmath'
    fd = lib.oracle_detmath_d
    fd.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64]
    fd.restype = None
    din, dout = (C.c_double * 1)(), (C.c_double * 1)()

    def exp64(x):
        din[0] = x
        fd(0, din, dout, 1)
        return np.float64(dout[0])
    math64.exp = exp64


use_libm()


def _m(name):
    npf = {'acos': np.arccos, 'asin': np.arcsin, 'atan': np.arctan}.get(name) or getattr(np, name)

    def f(x):
        if type(x) is np.float32:
            with np.errstate(all='ignore'):
                return np.float32(getattr(math32, name)(x))
        if type(x) is FArr:
            if x.a.dtype == np.float32:
                return FArr(np.array([getattr(math32, name)(v) for v in x.a.ravel(order='F')],
                                     np.float32).reshape(x.a.shape, order='F'))
            return FArr(npf(x.a))
        with np.errstate(all='ignore'):
            return getattr(math64, name, npf)(np.float64(x))
    return f


for _n in ('log', 'exp', 'sin', 'cos', 'tan', 'acos', 'asin', 'atan', 'log10', 'sinh', 'cosh', 'tanh'):
    globals()['m_' + _n] = _m(_n)


def m_sqrt(x):
    if type(x) is FArr:
        return FArr(np.sqrt(x.a))
    with np.errstate(all='ignore'):
        return np.sqrt(x)          # correctly rounded in the argument's precision


def m_atan2(y, x):
    return np.arctan2(y, x)
