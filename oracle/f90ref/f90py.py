"""f90py -- a small Fortran-90-subset -> Python translator (TEST INFRASTRUCTURE).

Why it exists: the reference (rwesson/mocassin) is Fortran 90 + MPI and no Fortran
compiler exists in this image, so the reference's own code could not be *run* to pin the
C oracle.  This translator turns the reference's hot-path modules, read where they lie
under /root/reference/source, into an executable Python module written to oracle/_ref/
(git-ignored; no reference source is copied into the repository).  The generated module is
the reference's own statements executed with Fortran semantics:

* default REAL arithmetic is IEEE float32 (numpy.float32 scalars), DOUBLE PRECISION float64,
  integer division truncates, x**n with integer n is repeated multiplication;
* arrays keep their declared lower bounds (rt.FArr), sections are views, whole-array
  assignment and elementwise arithmetic work as in Fortran; out-of-bounds access raises;
* derived types are classes with value-semantics assignment; user-defined operators
  dispatch through the module's interface blocks;
* dummy arguments are by reference: arrays and derived types are shared objects, scalar
  non-intent(in) dummies are returned to the caller and stored back into the actual
  argument; internal procedures see their host's variables (Python closures);
* DO loops leave the index at last+step on normal completion, EXIT/CYCLE/RETURN/STOP,
  SELECT CASE (strings compare blank-padded), optional arguments with PRESENT.

Only what the hot path uses is supported (no I/O beyond PRINT, no pointers, no WHERE /
FORALL, no GOTO); anything else raises at translation time, never silently.

The only things bound from outside are the intrinsics whose results the Fortran standard
leaves to the processor: RANDOM_NUMBER (bound to the oracle's Philox stream so that the
two sides draw the same uniforms) and LOG/SIN/COS/ACOS/ATAN/EXP (bound either to numpy's
libm or to the oracle's detmath), see rt.py.
"""
from __future__ import annotations

import keyword
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional

KW_DOTOPS = {'.and.', '.or.', '.not.', '.eqv.', '.neqv.', '.eq.', '.ne.', '.lt.', '.le.', '.gt.',
             '.ge.', '.true.', '.false.'}
REL_ALIAS = {'.eq.': '==', '.ne.': '/=', '.lt.': '<', '.le.': '<=', '.gt.': '>', '.ge.': '>='}


class TranslateError(Exception):
    pass


# ----------------------------------------------------------------------------------------
# logical lines
# ----------------------------------------------------------------------------------------
def logical_lines(text: str):
    """Yield (first line number, statement) with comments removed, continuations joined and
    everything outside string literals lower-cased."""
    out = []
    cur = []
    cur_line = 0
    in_str = None
    cont = False
    for ln, raw in enumerate(text.split('\n'), 1):
        line = raw.rstrip('\r')
        if cont:
            s = line.lstrip()
            if s.startswith('&'):
                line = s[1:]
            elif in_str is None:
                line = ' ' + s
            if in_str is None and not line.strip():
                continue  # blank / comment-only line inside a continuation
        else:
            cur_line = ln
        buf = []
        i = 0
        while i < len(line):
            ch = line[i]
            if in_str:
                buf.append(ch)
                if ch == in_str:
                    if i + 1 < len(line) and line[i + 1] == in_str:
                        buf.append(in_str)
                        i += 1
                    else:
                        in_str = None
            else:
                if ch in '"\'':
                    in_str = ch
                    buf.append(ch)
                elif ch == '!':
                    break
                else:
                    buf.append(ch.lower())
            i += 1
        s = ''.join(buf)
        st = s.rstrip()
        if cont and in_str is None and not st.strip():
            continue
        if st.endswith('&') :
            cur.append(st[:-1] if in_str else st[:-1].rstrip() + ' ')
            cont = True
            continue
        if in_str is not None and not st.endswith('&'):
            # unterminated string without continuation: give up on the literal
            in_str = None
        cur.append(s)
        stmt = ''.join(cur).strip()
        cur = []
        cont = False
        if stmt:
            for part in _split_semicolons(stmt):
                out.append((cur_line, part))
    return out


def _split_semicolons(s: str):
    if ';' not in s:
        return [s]
    parts, buf, q = [], [], None
    for ch in s:
        if q:
            buf.append(ch)
            if ch == q:
                q = None
        elif ch in '"\'':
            q = ch
            buf.append(ch)
        elif ch == ';':
            parts.append(''.join(buf).strip())
            buf = []
        else:
            buf.append(ch)
    parts.append(''.join(buf).strip())
    return [p for p in parts if p]


# ----------------------------------------------------------------------------------------
# tokens
# ----------------------------------------------------------------------------------------
TOKEN_RE = re.compile(r"""
 (?P<ws>\s+)
|(?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
|(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[ed][+-]?\d+)?(?:_\w+)?)
|(?P<dotop>\.[a-z]+\.)
|(?P<name>[a-z_]\w*)
|(?P<op>\*\*|//|==|/=|<=|>=|=>|::|\(/|/\)|[-+*/(),=<>%:\[\]])
""", re.X)
DOT_AHEAD = re.compile(r'[a-z]+\.')


@dataclass
class Tok:
    kind: str
    text: str


def tokenize(s: str) -> List[Tok]:
    toks = []
    i = 0
    n = len(s)
    while i < n:
        m = TOKEN_RE.match(s, i)
        if not m:
            raise TranslateError(f'cannot tokenize at {s[i:i+20]!r} in {s!r}')
        kind = m.lastgroup
        text = m.group(kind)
        end = m.end()
        if kind == 'num' and text.endswith('.') and not re.search(r'[ed]', text):
            m2 = DOT_AHEAD.match(s, end)
            if m2 and ('.' + m2.group(0)) in KW_DOTOPS:
                text = text[:-1]
                end -= 1
        if kind == 'num' and re.fullmatch(r'\d+', text) is None and re.fullmatch(r'\d+\.', text):
            pass
        if kind == 'op' and text == '(/' and toks and toks[-1].kind in ('name',) :
            # name(/ ... : '(' followed by '/=' never happens here; treat as '(' '/'
            toks.append(Tok('op', '('))
            i = m.start() + 1
            continue
        if kind == 'op' and text == '/)':
            # only an array-constructor close if one is open
            depth = 0
            for t in toks:
                if t.kind == 'op' and t.text == '(/':
                    depth += 1
                elif t.kind == 'op' and t.text == '/)':
                    depth -= 1
            if depth <= 0:
                toks.append(Tok('op', '/'))
                i = m.start() + 1
                continue
        if kind != 'ws':
            toks.append(Tok(kind, text))
        i = end
    return toks


# ----------------------------------------------------------------------------------------
# expression parser
# ----------------------------------------------------------------------------------------
class P:
    def __init__(self, toks: List[Tok], src: str = ''):
        self.t = toks
        self.i = 0
        self.src = src

    def peek(self, k=0) -> Optional[Tok]:
        j = self.i + k
        return self.t[j] if j < len(self.t) else None

    def at(self, text, k=0) -> bool:
        t = self.peek(k)
        return t is not None and t.text == text and t.kind != 'str'

    def at_name(self, text=None, k=0) -> bool:
        t = self.peek(k)
        return t is not None and t.kind == 'name' and (text is None or t.text == text)

    def eat(self, text=None) -> Tok:
        t = self.peek()
        if t is None or (text is not None and t.text != text):
            raise TranslateError(f'expected {text!r} at token {self.i} ({t}) in {self.src!r}')
        self.i += 1
        return t

    def done(self) -> bool:
        return self.i >= len(self.t)

    # --- precedence climbing ---
    def expr(self):
        left = self.eqv()
        while self.peek() is not None and self.peek().kind == 'dotop' and self.peek().text not in KW_DOTOPS:
            op = self.eat().text
            right = self.eqv()
            left = ('bin', op, left, right)
        return left

    def eqv(self):
        left = self.or_()
        while self.at('.eqv.') or self.at('.neqv.'):
            op = self.eat().text
            left = ('bin', op, left, self.or_())
        return left

    def or_(self):
        left = self.and_()
        while self.at('.or.'):
            self.eat()
            left = ('bin', '.or.', left, self.and_())
        return left

    def and_(self):
        left = self.not_()
        while self.at('.and.'):
            self.eat()
            left = ('bin', '.and.', left, self.not_())
        return left

    def not_(self):
        if self.at('.not.'):
            self.eat()
            return ('un', '.not.', self.not_())
        return self.rel()

    def rel(self):
        left = self.concat()
        t = self.peek()
        if t is not None and t.kind != 'str':
            op = REL_ALIAS.get(t.text, t.text)
            if op in ('==', '/=', '<', '<=', '>', '>='):
                self.eat()
                return ('bin', op, left, self.concat())
        return left

    def concat(self):
        left = self.add()
        while self.at('//'):
            self.eat()
            left = ('bin', '//', left, self.add())
        return left

    def add(self):
        if self.at('-') or self.at('+'):
            op = self.eat().text
            left = self.mul()
            if op == '-':
                left = ('un', '-', left)
        else:
            left = self.mul()
        while self.at('+') or self.at('-'):
            op = self.eat().text
            left = ('bin', op, left, self.mul())
        return left

    def mul(self):
        left = self.pow_()
        while self.at('*') or self.at('/'):
            op = self.eat().text
            left = ('bin', op, left, self.pow_())
        return left

    def pow_(self):
        base = self.unary_primary()
        if self.at('**'):
            self.eat()
            if self.at('-') or self.at('+'):
                op = self.eat().text
                e = self.pow_()
                if op == '-':
                    e = ('un', '-', e)
            else:
                e = self.pow_()
            return ('bin', '**', base, e)
        return base

    def unary_primary(self):
        if self.at('-') or self.at('+'):       # e.g. a*-b (extension)
            op = self.eat().text
            e = self.unary_primary()
            return ('un', '-', e) if op == '-' else e
        return self.primary()

    def primary(self):
        t = self.peek()
        if t is None:
            raise TranslateError(f'unexpected end of expression in {self.src!r}')
        if t.kind == 'num':
            self.eat()
            return _num(t.text)
        if t.kind == 'str':
            self.eat()
            q = t.text[0]
            return ('str', t.text[1:-1].replace(q + q, q))
        if t.kind == 'dotop' and t.text in ('.true.', '.false.'):
            self.eat()
            return ('log', t.text == '.true.')
        if t.text == '(/' or t.text == '[':
            close = '/)' if t.text == '(/' else ']'
            self.eat()
            items = []
            while not self.at(close):
                items.append(self.expr())
                if self.at(','):
                    self.eat()
            self.eat(close)
            return ('arrcon', items)
        if t.text == '(':
            self.eat()
            e = self.expr()
            self.eat(')')
            return ('paren', e)
        if t.kind == 'name':
            self.eat()
            node = ('name', t.text)
            return self.postfix(node)
        raise TranslateError(f'unexpected token {t} in {self.src!r}')

    def postfix(self, node):
        while True:
            if self.at('('):
                self.eat()
                args = self.args(')')
                node = ('call', node, args)
            elif self.at('%'):
                self.eat()
                node = ('comp', node, self.eat().text)
            else:
                return node

    def args(self, close):
        args = []
        while not self.at(close):
            args.append(self.arg())
            if self.at(','):
                self.eat()
            elif not self.at(close):
                raise TranslateError(f'bad argument list in {self.src!r}')
        self.eat(close)
        return args

    def arg(self):
        if self.at_name() and self.at('=', 1) :
            name = self.eat().text
            self.eat('=')
            return ('kw', name, self.expr())
        lo = None
        if not self.at(':'):
            lo = self.expr()
            if not self.at(':'):
                return lo
        self.eat(':')
        hi = None
        st = None
        if not (self.at(',') or self.at(')') or self.at(':')):
            hi = self.expr()
        if self.at(':'):
            self.eat()
            st = self.expr()
        return ('slice', lo, hi, st)


def _num(text: str):
    kindsfx = None
    if '_' in text:
        text, kindsfx = text.split('_', 1)
    if re.fullmatch(r'\d+', text):
        return ('num', text, 'i')
    if 'd' in text:
        return ('num', text.replace('d', 'e'), 'd')
    if kindsfx in ('8', 'dp'):
        return ('num', text, 'd')
    return ('num', text, 'r')


def parse_expr(s: str):
    p = P(tokenize(s), s)
    e = p.expr()
    if not p.done():
        raise TranslateError(f'trailing tokens in expression {s!r}')
    return e


# ----------------------------------------------------------------------------------------
# program structure
# ----------------------------------------------------------------------------------------
@dataclass
class Ty:
    base: str            # i r d l c t ?
    rank: int = 0
    tname: Optional[str] = None
    clen: Optional[int] = None

    def scalar(self):
        return Ty(self.base, 0, self.tname, self.clen)


@dataclass
class Decl:
    name: str
    ty: Ty
    dims: Optional[list] = None       # list of (lo_expr|None, hi_expr|None|'*') ; None entries = deferred
    attrs: set = field(default_factory=set)
    intent: Optional[str] = None
    init: object = None
    line: int = 0


@dataclass
class TypeDef:
    name: str
    comps: Dict[str, Decl] = field(default_factory=dict)
    order: List[str] = field(default_factory=list)


@dataclass
class Proc:
    kind: str                 # subroutine | function
    name: str
    args: List[str]
    result: Optional[str]
    decls: Dict[str, Decl] = field(default_factory=dict)
    body: list = field(default_factory=list)
    contains: Dict[str, 'Proc'] = field(default_factory=dict)
    parent: Optional['Proc'] = None
    module: Optional[str] = None
    line: int = 0
    prefix_ty: Optional[Ty] = None
    alias: Dict[str, str] = field(default_factory=dict)   # scalar dummy -> host variable it is bound to

    def out_scalars(self) -> List[str]:
        """dummies handed back to the caller: scalar, intrinsic type, not intent(in)."""
        res = []
        for a in self.args:
            d = self.decls.get(a)
            if d is None:
                raise TranslateError(f'{self.name}: dummy {a} has no declaration')
            if d.ty.rank == 0 and d.ty.base in 'irdlc' and d.intent != 'in' and a not in self.alias:
                res.append(a)
        return res


@dataclass
class Module:
    name: str
    uses: List[str] = field(default_factory=list)
    decls: Dict[str, Decl] = field(default_factory=dict)
    order: List[str] = field(default_factory=list)
    types: Dict[str, TypeDef] = field(default_factory=dict)
    procs: Dict[str, Proc] = field(default_factory=dict)
    operators: Dict[str, List[str]] = field(default_factory=dict)
    generics: Dict[str, List[str]] = field(default_factory=dict)
    failed: Dict[str, str] = field(default_factory=dict)      # procedures skipped by a lenient parse


TYPE_START = re.compile(r'^((integer|real|double\s*precision|double\s*complex|logical|character|complex)\b|type\s*\()')


def _split_top(toks: List[Tok], sep=','):
    parts, cur, depth = [], [], 0
    for t in toks:
        if t.kind == 'op' and t.text in ('(', '(/', '['):
            depth += 1
        elif t.kind == 'op' and t.text in (')', '/)', ']'):
            depth -= 1
        if depth == 0 and t.kind == 'op' and t.text == sep:
            parts.append(cur)
            cur = []
        else:
            cur.append(t)
    parts.append(cur)
    return parts


def _expr_from(toks: List[Tok], src=''):
    p = P(toks, src)
    e = p.expr()
    if not p.done():
        raise TranslateError(f'trailing tokens {toks[p.i:]} in {src!r}')
    return e


def _parse_dims(toks: List[Tok], src):
    dims = []
    for part in _split_top(toks):
        # forms:  :   n   lo:hi   lo:   *   lo:*
        depth = 0
        colon = None
        for k, t in enumerate(part):
            if t.kind == 'op' and t.text in ('(', '(/'):
                depth += 1
            elif t.kind == 'op' and t.text in (')', '/)'):
                depth -= 1
            elif depth == 0 and t.kind == 'op' and t.text == ':':
                colon = k
                break
        if colon is None:
            if len(part) == 1 and part[0].text == '*':
                dims.append((('num', '1', 'i'), '*'))
            else:
                dims.append((('num', '1', 'i'), _expr_from(part, src)))
        else:
            lo_t, hi_t = part[:colon], part[colon + 1:]
            lo = _expr_from(lo_t, src) if lo_t else None
            if not hi_t:
                hi = None
            elif len(hi_t) == 1 and hi_t[0].text == '*':
                hi = '*'
            else:
                hi = _expr_from(hi_t, src)
            dims.append((lo, hi))
    return dims


def parse_decl(stmt: str, line: int) -> List[Decl]:
    toks = tokenize(stmt)
    p = P(toks, stmt)
    t0 = p.eat().text
    base, tname, clen = None, None, None
    if t0 == 'integer':
        base = 'i'
    elif t0 == 'real':
        base = 'r'
    elif t0 == 'double':
        if p.at('complex'):
            p.eat()
            base = 'Z'                   # double complex
        else:
            p.eat('precision')
            base = 'd'
    elif t0 == 'doubleprecision':
        base = 'd'
    elif t0 == 'doublecomplex':
        base = 'Z'
    elif t0 == 'logical':
        base = 'l'
    elif t0 == 'character':
        base = 'c'
        clen = 1
    elif t0 == 'complex':
        base = 'z'                       # default (single precision) complex
    elif t0 == 'type':
        p.eat('(')
        tname = p.eat().text
        p.eat(')')
        base = 't'
    else:
        raise TranslateError('not a declaration: ' + stmt)
    # kind / len selector
    if base != 't' and (p.at('(') or p.at('*')):
        if p.at('*'):
            p.eat()
            if p.at('('):
                p.eat(); sel = [('*',)] if p.at('*') else [p.expr()]
                if p.at('*'):
                    p.eat()
                p.eat(')')
            else:
                sel = [_num(p.eat().text)]
        else:
            p.eat('(')
            sel = []
            while not p.at(')'):
                if p.at_name() and p.at('=', 1):
                    k = p.eat().text
                    p.eat('=')
                    if p.at('*'):
                        p.eat(); sel.append(('kwstar', k))
                    else:
                        sel.append(('kw', k, p.expr()))
                elif p.at('*'):
                    p.eat(); sel.append(('*',))
                else:
                    sel.append(p.expr())
                if p.at(','):
                    p.eat()
            p.eat(')')
        for s in sel:
            val = s[2] if s[0] == 'kw' else s
            if base == 'c':
                if s[0] in ('*', 'kwstar'):
                    clen = None
                elif val[0] == 'num':
                    clen = int(val[1])
                else:
                    clen = None
            elif base == 'r':
                if val[0] == 'num' and val[1] == '8':
                    base = 'd'
                elif val[0] == 'num' and val[1] == '4':
                    pass
                elif val[0] == 'name' and val[1] in ('dp', 'real64', 'double'):
                    base = 'd'
                else:
                    raise TranslateError('unsupported real kind in ' + stmt)
            elif base == 'z':
                if val[0] == 'num' and val[1] == '8':
                    base = 'Z'
                elif val[0] == 'num' and val[1] == '4':
                    pass
                else:
                    raise TranslateError('unsupported complex kind in ' + stmt)
            elif base == 'i':
                pass
    attrs = set()
    intent = None
    dim_attr = None
    while p.at(','):
        p.eat()
        a = p.eat().text
        if a == 'dimension':
            p.eat('(')
            depth = 1
            sub = []
            while True:
                t = p.eat()
                if t.kind == 'op' and t.text in ('(', '(/'):
                    depth += 1
                elif t.kind == 'op' and t.text in (')', '/)'):
                    depth -= 1
                    if depth == 0:
                        break
                sub.append(t)
            dim_attr = _parse_dims(sub, stmt)
        elif a == 'intent':
            p.eat('(')
            w = p.eat().text
            if w == 'in' and p.at_name('out'):
                p.eat(); w = 'inout'
            p.eat(')')
            intent = w
        else:
            attrs.add(a)
    if p.at('::'):
        p.eat()
    rest = toks[p.i:]
    decls = []
    for ent in _split_top(rest):
        if not ent:
            continue
        q = P(ent, stmt)
        name = q.eat().text
        dims = dim_attr
        if q.at('('):
            q.eat()
            depth = 1
            sub = []
            while True:
                t = q.eat()
                if t.kind == 'op' and t.text in ('(', '(/'):
                    depth += 1
                elif t.kind == 'op' and t.text in (')', '/)'):
                    depth -= 1
                    if depth == 0:
                        break
                sub.append(t)
            dims = _parse_dims(sub, stmt)
        elen = clen
        if q.at('*'):
            q.eat()
            if q.at('('):
                q.eat();
                if q.at('*'):
                    q.eat(); elen = None
                else:
                    elen = int(q.eat().text)
                q.eat(')')
            else:
                elen = int(q.eat().text)
        init = None
        if q.at('='):
            q.eat()
            init = q.expr()
        elif q.at('=>'):
            q.eat(); q.expr()
        if not q.done():
            raise TranslateError(f'bad entity {ent} in {stmt!r}')
        d = Decl(name, Ty(base, len(dims) if dims else 0, tname, elen), dims, set(attrs), intent, init, line)
        decls.append(d)
    return decls


# statement nodes are tuples: (kind, line, ...)
class Unit:
    """Parser of one source file into Module objects."""

    def __init__(self, text: str, fname: str = '', spec_only: bool = False, lenient: bool = False):
        """lenient: a statement the parser does not know becomes an ('untranslated', ...) node
        that raises if it is ever *reached* at run time, and a procedure whose block structure
        cannot be parsed is skipped (recorded in Module.failed); strict: both are errors."""
        self.lines = logical_lines(text)
        self.k = 0
        self.fname = fname
        self.spec_only = spec_only
        self.lenient = lenient
        self.modules: List[Module] = []
        while self.k < len(self.lines):
            ln, s = self.lines[self.k]
            m = re.match(r'^module\s+(\w+)$', s)
            if m and not s.startswith('module procedure'):
                self.k += 1
                self.modules.append(self.parse_module(m.group(1)))
            else:
                self.k += 1   # program units other than modules are ignored

    def err(self, msg, ln):
        raise TranslateError(f'{self.fname}:{ln}: {msg}')

    def parse_module(self, name) -> Module:
        mod = Module(name)
        in_contains = False
        while self.k < len(self.lines):
            ln, s = self.lines[self.k]
            if re.match(r'^end\s*module\b', s) or s == 'end':
                self.k += 1
                return mod
            if s == 'contains':
                in_contains = True
                self.k += 1
                if self.spec_only:
                    # skip to end module
                    while self.k < len(self.lines) and not re.match(r'^end\s*module\b', self.lines[self.k][1]):
                        self.k += 1
                continue
            if in_contains:
                pr = self.try_proc_header(s, ln)
                if pr is None:
                    self.err(f'expected a procedure, got {s!r}', ln)
                self.k += 1
                if self.parse_proc_or_skip(pr, mod):
                    pr.module = mod.name
                    mod.procs[pr.name] = pr
                continue
            self.k += 1
            self.spec_stmt(s, ln, mod.decls, mod, mod.order)
        return mod

    def spec_stmt(self, s, ln, decls, mod: Optional[Module], order=None) -> bool:
        """Handle one specification statement; return False if s is not one."""
        if s.startswith('use '):
            m = re.match(r'^use\s+(\w+)', s)
            if mod is not None:
                mod.uses.append(m.group(1))
            return True
        if re.match(r'^(implicit\b|public\b|private\b|save\b|external\b|include\b|intrinsic\b)', s):
            return True
        m = re.match(r'^interface\s*(.*)$', s)
        if m:
            what = m.group(1).strip()
            names = []
            while True:
                ln2, s2 = self.lines[self.k]
                self.k += 1
                if re.match(r'^end\s*interface', s2):
                    break
                m2 = re.match(r'^module\s+procedure\s+(.*)$', s2)
                if m2:
                    names += [x.strip() for x in m2.group(1).split(',')]
            if mod is not None:
                mo = re.match(r'^operator\s*\(\s*(.+?)\s*\)$', what)
                if mo:
                    mod.operators.setdefault(REL_ALIAS.get(mo.group(1), mo.group(1)), []).extend(names)
                elif what and not what.startswith('assignment'):
                    mod.generics.setdefault(what, []).extend(names)
            return True
        m = re.match(r'^type\s*(?:,\s*(?:public|private)\s*)?(?:::)?\s*(\w+)$', s)
        if m and not s.startswith('type('):
            td = TypeDef(m.group(1))
            while True:
                ln2, s2 = self.lines[self.k]
                self.k += 1
                if re.match(r'^end\s*type', s2):
                    break
                if s2 in ('sequence', 'private', 'public'):
                    continue
                for d in parse_decl(s2, ln2):
                    td.comps[d.name] = d
                    td.order.append(d.name)
            if mod is not None:
                mod.types[td.name] = td
            else:
                self.err('derived type defined inside a procedure is not supported', ln)
            return True
        if TYPE_START.match(s) and not re.match(r'^(real|integer|logical|character)\s*(\(.*?\))?\s*=', s):
            # "real function f(x)" is a procedure header, not a declaration
            if re.search(r'\bfunction\b', s.split('::')[0]) and '::' not in s:
                return False
            for d in parse_decl(s, ln):
                if d.name in decls:
                    # e.g. "dimension" given later; merge
                    pass
                decls[d.name] = d
                if order is not None:
                    order.append(d.name)
            return True
        if re.match(r'^parameter\s*\(', s) or re.match(r'^(dimension|data|common|equivalence|namelist)\b', s):
            self.err('unsupported specification statement: ' + s, ln)
        return False

    def try_proc_header(self, s, ln) -> Optional[Proc]:
        m = re.match(r'^(?:(?:recursive|pure|elemental)\s+)*(.*?)\b(subroutine|function)\s+(\w+)\s*(?:\((.*?)\))?\s*(?:result\s*\(\s*(\w+)\s*\))?$', s)
        if not m:
            return None
        prefix, kind, name, args, res = m.groups()
        prefix = prefix.strip()
        if prefix and not TYPE_START.match(prefix):
            return None
        if s.startswith('end'):
            return None
        argl = [a.strip() for a in args.split(',')] if args and args.strip() else []
        pr = Proc(kind, name, argl, (res or name) if kind == 'function' else None, line=ln)
        if prefix:
            pr.prefix_ty = parse_decl(prefix + ' :: ' + (res or name), ln)[0].ty
        return pr

    def parse_proc(self, pr: Proc, mod: Optional[Module]):
        # specification part
        while self.k < len(self.lines):
            ln, s = self.lines[self.k]
            if self.spec_stmt_in_proc(s, ln, pr):
                continue
            break
        if pr.kind == 'function' and pr.result not in pr.decls:
            if pr.prefix_ty is None:
                self.err(f'function {pr.name}: result type not declared', pr.line)
            pr.decls[pr.result] = Decl(pr.result, pr.prefix_ty, None, set(), None, None, pr.line)
        pr.body = self.parse_block(('end',), pr)
        ln, s = self.lines[self.k]
        if s == 'contains':
            self.k += 1
            while True:
                ln, s = self.lines[self.k]
                if re.match(r'^end\b', s):
                    break
                sub = self.try_proc_header(s, ln)
                if sub is None:
                    self.err(f'expected internal procedure, got {s!r}', ln)
                self.k += 1
                sub.parent = pr
                if self.parse_proc_or_skip(sub, mod):
                    pr.contains[sub.name] = sub
            ln, s = self.lines[self.k]
        if not re.match(r'^end(\s*(subroutine|function)(\s+\w+)?)?$', s):
            self.err(f'expected end of {pr.name}, got {s!r}', ln)
        self.k += 1

    def parse_proc_or_skip(self, pr: Proc, mod: Optional[Module]) -> bool:
        start = self.k
        if not self.lenient:
            self.parse_proc(pr, mod)
            return True
        try:
            self.parse_proc(pr, mod)
            return True
        except TranslateError as ex:
            # skip to "end subroutine|function <name>"
            self.k = start
            pat = re.compile(r'^end\s*(subroutine|function)\s+' + re.escape(pr.name) + r'$')
            while self.k < len(self.lines) and not pat.match(self.lines[self.k][1]):
                self.k += 1
            if self.k >= len(self.lines):
                raise
            self.k += 1
            if mod is not None:
                mod.failed[pr.name] = str(ex)
            return False

    def spec_stmt_in_proc(self, s, ln, pr: Proc) -> bool:
        save = self.k
        self.k += 1
        if self.spec_stmt(s, ln, pr.decls, None):
            return True
        self.k = save
        return False

    # ---- executable part ----
    def parse_block(self, terminators, pr) -> list:
        """Parse statements until a line whose first word(s) match one of `terminators`
        (not consumed)."""
        body = []
        while True:
            if self.k >= len(self.lines):
                self.err('unexpected end of file', self.lines[-1][0])
            ln, s = self.lines[self.k]
            if self.is_term(s, terminators):
                return body
            self.k += 1
            body.append(self.parse_stmt(s, ln, pr))

    @staticmethod
    def is_term(s, terms):
        for t in terms:
            if t == 'end':
                if s == 'contains' or re.match(r'^end(\s*(subroutine|function)(\s+\w+)?)?$', s):
                    return True
            elif t == 'endif':
                if re.match(r'^(end\s*if|else\s*if\b|else$|elseif\b)', s):
                    return True
            elif t == 'enddo':
                if re.match(r'^end\s*do$', s):
                    return True
            elif t == 'endselect':
                if re.match(r'^(end\s*select|case\b)', s):
                    return True
        return False

    def parse_stmt(self, s, ln, pr):
        toks = tokenize(s)
        first = toks[0].text
        # block if / one-line if
        if first == 'if' and len(toks) > 1 and toks[1].text == '(':
            close = _match_paren(toks, 1)
            cond = _expr_from(toks[2:close], s)
            rest = toks[close + 1:]
            if len(rest) == 1 and rest[0].text == 'then':
                branches = []
                body = self.parse_block(('endif',), pr)
                branches.append((cond, body))
                else_body = None
                while True:
                    ln2, s2 = self.lines[self.k]
                    self.k += 1
                    if re.match(r'^end\s*if$', s2):
                        break
                    m = re.match(r'^else\s*if\s*\(', s2)
                    if m:
                        t2 = tokenize(s2)
                        j = next(i for i, t in enumerate(t2) if t.text == '(')
                        c2 = _match_paren(t2, j)
                        cnd = _expr_from(t2[j + 1:c2], s2)
                        if not (len(t2) == c2 + 2 and t2[c2 + 1].text == 'then'):
                            self.err('bad else if: ' + s2, ln2)
                        branches.append((cnd, self.parse_block(('endif',), pr)))
                    elif s2 == 'else':
                        else_body = self.parse_block(('endif',), pr)
                    else:
                        self.err('bad if construct near ' + s2, ln2)
                return ('if', ln, branches, else_body)
            inner = self.parse_stmt(_untok(rest), ln, pr)
            return ('if', ln, [(cond, [inner])], None)
        if first == 'do':
            if len(toks) == 1:
                body = self.parse_block(('enddo',), pr)
                self.k += 1
                return ('doforever', ln, body)
            if toks[1].text == 'while':
                close = _match_paren(toks, 2)
                cond = _expr_from(toks[3:close], s)
                body = self.parse_block(('enddo',), pr)
                self.k += 1
                return ('dowhile', ln, cond, body)
            if toks[1].kind == 'name' and toks[2].text == '=':
                parts = _split_top(toks[3:])
                lo = _expr_from(parts[0], s)
                hi = _expr_from(parts[1], s)
                st = _expr_from(parts[2], s) if len(parts) > 2 else None
                body = self.parse_block(('enddo',), pr)
                self.k += 1
                return ('do', ln, toks[1].text, lo, hi, st, body)
            self.err('unsupported do statement: ' + s, ln)
        if first == 'select' and toks[1].text == 'case':
            close = _match_paren(toks, 2)
            sel = _expr_from(toks[3:close], s)
            cases = []
            # skip to first case
            while True:
                ln2, s2 = self.lines[self.k]
                self.k += 1
                if re.match(r'^end\s*select$', s2):
                    break
                if s2.replace(' ', '') == 'casedefault':
                    cases.append((None, self.parse_block(('endselect',), pr)))
                    continue
                m = re.match(r'^case\s*\(', s2)
                if not m:
                    self.err('bad select construct near ' + s2, ln2)
                t2 = tokenize(s2)
                c2 = _match_paren(t2, 1)
                vals = []
                for part in _split_top(t2[2:c2]):
                    q = P(part, s2)
                    vals.append(q.arg())
                cases.append((vals, self.parse_block(('endselect',), pr)))
            return ('select', ln, sel, cases)
        if first == 'call':
            name = toks[1].text
            args = []
            if len(toks) > 2:
                q = P(toks[2:], s)
                q.eat('(')
                args = q.args(')')
            return ('call', ln, name, args)
        if first in ('return', 'exit', 'cycle', 'continue') and len(toks) == 1:
            return (first, ln)
        if first == 'stop':
            return ('stop', ln)
        if first == 'print':
            items = _split_top(toks[1:])
            exprs = []
            for it in items[1:]:
                if it:
                    exprs.append(_expr_from(it, s))
            return ('print', ln, exprs)
        if first in ('open', 'close') and len(toks) > 1 and toks[1].text == '(':
            return ('ionoop', ln, s)
        if first == 'write' and len(toks) > 1 and toks[1].text == '(':
            # list-directed WRITE(unit,*) only; items may be implied-DO lists (expr, ..., v=lo,hi)
            try:
                close = _match_paren(toks, 1)
                ctl = _split_top(toks[2:close])
                if len(ctl) == 2 and len(ctl[1]) == 1 and ctl[1][0].text == '*':
                    unit = _expr_from(ctl[0], s)
                    items = []
                    for it in _split_top(toks[close + 1:]):
                        if it:
                            items.append(_io_item(it, s))
                    return ('write', ln, unit, items)
            except TranslateError:
                pass
            return ('io', ln, s)
        if first == 'read' and len(toks) > 1 and toks[1].text == '(':
            # list-directed READ(unit, *[, iostat=v]) items; items may be implied-DO lists
            try:
                close = _match_paren(toks, 1)
                unit, star, iostat = None, False, None
                for j, c in enumerate(_split_top(toks[2:close])):
                    if len(c) >= 3 and c[0].kind == 'name' and c[1].text == '=':
                        key = c[0].text.lower()
                        if key == 'unit':
                            unit = _expr_from(c[2:], s)
                        elif key == 'fmt':
                            star = len(c) == 3 and c[2].text == '*'
                        elif key == 'iostat':
                            iostat = _expr_from(c[2:], s)
                        else:
                            raise TranslateError('READ control item ' + key)
                    elif j == 0:
                        unit = _expr_from(c, s)
                    elif j == 1:
                        star = len(c) == 1 and c[0].text == '*'
                    else:
                        raise TranslateError('READ control list')
                if unit is not None and star:
                    items = [_io_item(it, s) for it in _split_top(toks[close + 1:]) if it]
                    return ('read', ln, unit, iostat, items)
            except TranslateError:
                pass
            return ('io', ln, s)
        if first in ('backspace', 'rewind') and len(toks) >= 2:
            try:
                inner = toks[2:_match_paren(toks, 1)] if toks[1].text == '(' else toks[1:]
                if len(inner) >= 3 and inner[0].kind == 'name' and inner[1].text == '=':
                    inner = inner[2:]
                return ('seek', ln, first, _expr_from(inner, s))
            except TranslateError:
                return ('io', ln, s)
        if first in ('write', 'read', 'open', 'close', 'rewind', 'backspace', 'inquire') and len(toks) > 1 and toks[1].text == '(':
            return ('io', ln, s)
        if first == 'allocate' and toks[1].text == '(':
            q = P(toks[1:], s)
            q.eat('(')
            args = q.args(')')
            return ('allocate', ln, args)
        if first == 'deallocate' and toks[1].text == '(':
            q = P(toks[1:], s)
            q.eat('(')
            args = q.args(')')
            return ('deallocate', ln, args)
        if first in ('goto', 'go', 'where', 'forall', 'nullify', 'format', 'data', 'entry'):
            if self.lenient:
                return ('untranslated', ln, s)
            self.err('unsupported statement: ' + s, ln)
        # assignment
        try:
            q = P(toks, s)
            lhs = q.primary()
            if q.at('='):
                q.eat()
                rhs = q.expr()
                if not q.done():
                    self.err('trailing tokens in assignment: ' + s, ln)
                return ('assign', ln, lhs, rhs)
            self.err('unsupported statement: ' + s, ln)
        except TranslateError:
            if self.lenient:
                return ('untranslated', ln, s)
            raise


def _io_item(toks, src):
    """an output item: an expression, or an implied-DO list ( item, item, ..., v = lo, hi [, st] )"""
    if toks[0].text == '(' and _match_paren(toks, 0) == len(toks) - 1:
        parts = _split_top(toks[1:-1])
        for k, pt in enumerate(parts):
            if len(pt) >= 3 and pt[0].kind == 'name' and pt[1].text == '=' and k >= 1 and len(parts) - k in (2, 3):
                var = pt[0].text
                lo = _expr_from(pt[2:], src)
                hi = _expr_from(parts[k + 1], src)
                st = _expr_from(parts[k + 2], src) if len(parts) - k == 3 else None
                return ('implied', [_io_item(x, src) for x in parts[:k]], var, lo, hi, st)
    return ('item', _expr_from(toks, src))


def _match_paren(toks, i):
    assert toks[i].text == '(' or toks[i].text == '(/', toks[i]
    depth = 0
    for j in range(i, len(toks)):
        t = toks[j]
        if t.kind == 'op' and t.text in ('(', '(/'):
            depth += 1
        elif t.kind == 'op' and t.text in (')', '/)'):
            depth -= 1
            if depth == 0:
                return j
    raise TranslateError('unbalanced parentheses')


def _untok(toks):
    return ' '.join(t.text for t in toks)


# ----------------------------------------------------------------------------------------
# code generation
# ----------------------------------------------------------------------------------------
INTRINSIC_SUBS = {'random_number', 'random_seed', 'date_and_time', 'cpu_time', 'system_clock'}
PY_RESERVED = set(keyword.kwlist) | {'None', 'True', 'False', 'print', 'int', 'abs', 'max', 'min', 'range', 'len',
                                      'float', 'bool', 'slice', 'type', 'id', 'input', 'sum', 'any', 'all'}


def pyname(n: str) -> str:
    if n in PY_RESERVED or n.startswith('_'):
        return n + '_v'
    return n


class Scope:
    def __init__(self, gen: 'Gen', proc: Optional[Proc]):
        self.gen = gen
        self.proc = proc

    def lookup_var(self, name):
        """-> (where, Decl, emitted python name) or None"""
        p = self.proc
        first = True
        while p is not None:
            if name in p.alias:              # dummy bound by name to a host variable
                name = p.alias[name]
            elif name in p.decls:
                return ('local' if first else 'host', p.decls[name], pyname(name))
            p = p.parent
            first = False
        d = self.gen.globals.get(name)
        if d is not None:
            return ('global', d, f'_G.{pyname(name)}')
        return None

    def lookup_proc(self, name) -> Optional[Proc]:
        p = self.proc
        while p is not None:
            if name in p.contains:
                return p.contains[name]
            p = p.parent
        return self.gen.procs.get(name)


class Gen:
    def __init__(self, modules: List[Module], loop_hooks=(), proc_hooks=(), skip_procs=(), lenient=False,
                 externs=()):
        """lenient: a statement that cannot be translated becomes a call that raises when it is
        reached (never a silent skip); self.untranslated maps each procedure to those
        statements so that a recipe can insist that the procedures it pins are complete."""
        self.modules = modules
        self.lenient = lenient
        self.untranslated: Dict[str, List[str]] = {}
        # argument-less subroutines the harness supplies (rt.externs[name]); used for a routine
        # whose *result* is an input of the pinned code (BoltGaunt -> contBoltz, gauntFF)
        self.externs = set(externs)
        self.globals: Dict[str, Decl] = {}
        self.types: Dict[str, TypeDef] = {}
        self.procs: Dict[str, Proc] = {}
        self.operators: Dict[str, List[str]] = {}
        self.generics: Dict[str, List[str]] = {}
        self.loop_hooks = set(loop_hooks)
        self.proc_hooks = set(proc_hooks)
        self.skip_procs = set(skip_procs)
        self.consts: Dict[tuple, str] = {}
        self.warnings: List[str] = []
        self.tmp = 0
        for m in modules:
            for n in m.order:
                self.globals.setdefault(n, m.decls[n])
            self.types.update(m.types)
            for n, p in m.procs.items():
                if n not in self.skip_procs:
                    self.procs[n] = p
            for k, v in m.operators.items():
                self.operators.setdefault(k, []).extend(v)
            for k, v in m.generics.items():
                self.generics.setdefault(k, []).extend(v)
        for pr in self.procs.values():
            self.bind_aliases(pr)

    def bind_aliases(self, host: Proc):
        """By-reference aliasing between a host variable and a dummy of an internal procedure.

        The reference passes host variables to an internal procedure and *also* assigns them
        through host association further down the call chain (energyPacketRun's rR/chType/gP
        are reRun/chTypeIn/gPIn, which pathSegment sets at photon_mod.f90:2863-2869).  With
        by-reference argument passing the dummy and the host variable are one storage unit,
        which a copy-in/copy-out translation of scalar dummies would break.  So: when every
        call of an internal procedure passes the same plain scalar variable of the host for a
        non-intent(in) scalar dummy, the dummy is bound to that variable by name."""
        if not host.contains:
            return
        sites: Dict[tuple, list] = {}
        for caller in [host] + list(host.contains.values()):
            sc = Scope(self, caller)
            for name, args in _walk_calls(caller.body):
                q = host.contains.get(name)
                if q is None or sc.lookup_var(name) is not None:
                    continue
                pos = 0
                slots = {}
                for a in args:
                    if a[0] == 'kw':
                        slots[a[1]] = a[2]
                    elif pos < len(q.args):
                        slots[q.args[pos]] = a
                        pos += 1
                for d in q.args:
                    dd = q.decls.get(d)
                    if dd is None or dd.ty.rank or dd.ty.base not in 'irdlc' or dd.intent == 'in':
                        continue
                    a = slots.get(d)
                    v = None
                    if a is not None and a[0] == 'name' and a[1] in host.decls and \
                            (caller is host or a[1] not in caller.decls) and \
                            host.decls[a[1]].ty.rank == 0 and 'parameter' not in host.decls[a[1]].attrs:
                        v = a[1]
                    sites.setdefault((q.name, d), []).append(v)
        for (qn, d), vs in sites.items():
            if vs and vs[0] is not None and all(v == vs[0] for v in vs):
                host.contains[qn].alias[d] = vs[0]

    # ---- helpers ----
    def const(self, text, kind):
        key = (text, kind)
        if key not in self.consts:
            self.consts[key] = f'_K{len(self.consts)}'
        return self.consts[key]

    def newtmp(self):
        self.tmp += 1
        return f'_t{self.tmp}'

    # ---- type of a declaration-level entity ----
    def comp_decl(self, tname, comp) -> Decl:
        td = self.types.get(tname)
        if td is None:
            raise TranslateError(f'unknown derived type {tname}')
        if comp not in td.comps:
            raise TranslateError(f'type {tname} has no component {comp}')
        return td.comps[comp]

    # ---- expressions ----
    def expr(self, e, sc: Scope, pre: list):
        """-> (python code, Ty)"""
        k = e[0]
        if k == 'num':
            if e[2] == 'i':
                return str(int(e[1])), Ty('i')
            return self.const(e[1], e[2]), Ty(e[2])
        if k == 'str':
            return repr(e[1]), Ty('c', 0, None, len(e[1]))
        if k == 'log':
            return ('True' if e[1] else 'False'), Ty('l')
        if k == 'paren':
            c, t = self.expr(e[1], sc, pre)
            return f'({c})', t
        if k == 'name':
            return self.name_ref(e[1], sc)
        if k == 'comp':
            bc, bt = self.expr(e[1], sc, pre)
            if bt.base != 't':
                raise TranslateError(f'component {e[2]} of non-derived expression {e[1]}')
            if bt.rank != 0:
                raise TranslateError('component of an array of derived type is not supported')
            d = self.comp_decl(bt.tname, e[2])
            return f'{bc}.{pyname(e[2])}', d.ty
        if k == 'call':
            return self.call_or_index(e, sc, pre)
        if k == 'un':
            c, t = self.expr(e[2], sc, pre)
            if e[1] == '-':
                return f'(-{c})', t
            if e[1] == '.not.':
                if t.rank:
                    raise TranslateError('.not. on arrays not supported')
                return f'(not {c})', Ty('l')
            raise TranslateError('unary ' + e[1])
        if k == 'bin':
            return self.binop(e, sc, pre)
        if k == 'arrcon':
            items = [self.expr(x, sc, pre) for x in e[1]]
            base = 'i'
            for c, t in items:
                base = _promote(base, t.base) if t.base in 'ird' else t.base
            return f'_rt.arrcon([{", ".join(c for c, _ in items)}], {base!r})', Ty(base, 1)
        raise TranslateError(f'cannot translate expression node {e}')

    def name_ref(self, name, sc: Scope):
        r = sc.lookup_var(name)
        if r is None:
            pr = sc.lookup_proc(name)
            if pr is not None and pr.kind == 'function' and not pr.args:
                raise TranslateError(f'function {name} referenced without ()')
            raise TranslateError(f'unresolved name {name!r} in {sc.proc.name if sc.proc else "<module>"}')
        where, d, ename = r
        return ename, d.ty

    def binop(self, e, sc, pre):
        op = e[1]
        a, ta = self.expr(e[2], sc, pre)
        if op in ('.and.', '.or.'):
            # keep the hoisted calls of the right operand unconditional (Fortran allows it)
            b, tb = self.expr(e[3], sc, pre)
            if ta.rank or tb.rank:
                raise TranslateError('array-valued .and./.or. not supported')
            return f'({a} {"and" if op == ".and." else "or"} {b})', Ty('l')
        b, tb = self.expr(e[3], sc, pre)
        rank = max(ta.rank, tb.rank)
        if ta.base == 't' or tb.base == 't' or op not in ('+', '-', '*', '/', '**', '==', '/=', '<', '<=', '>', '>=', '//', '.eqv.', '.neqv.'):
            cands = self.operators.get(op, [])
            for pn in cands:
                pr = self.procs.get(pn)
                if pr is None or len(pr.args) != 2:
                    continue
                d0, d1 = pr.decls[pr.args[0]].ty, pr.decls[pr.args[1]].ty
                if _ty_match(d0, ta) and _ty_match(d1, tb):
                    rt_ = pr.decls[pr.result].ty
                    return f'p_{pn}({a}, {b})', rt_
            raise TranslateError(f'no operator {op} for ({ta}, {tb})')
        if op in ('+', '-', '*'):
            return f'({a} {op} {b})', Ty(_promote(ta.base, tb.base), rank)
        if op == '/':
            if ta.base == 'i' and tb.base == 'i':
                if rank:
                    raise TranslateError('integer array division not supported')
                return f'_rt.idiv({a}, {b})', Ty('i')
            return f'({a} / {b})', Ty(_promote(ta.base, tb.base), rank)
        if op == '**':
            if tb.base == 'i':
                return f'_rt.ipow({a}, {b})', Ty(ta.base, rank)
            return f'_rt.rpow({a}, {b})', Ty(_promote(ta.base, tb.base), rank)
        if op in ('==', '/=', '<', '<=', '>', '>='):
            pop = '!=' if op == '/=' else op
            if ta.base == 'c' and tb.base == 'c':
                return f'_rt.strcmp({a}, {b}, {pop!r})', Ty('l')
            return f'({a} {pop} {b})', Ty('l', rank)
        if op == '//':
            return f'({a} + {b})', Ty('c', 0, None, (ta.clen or 0) + (tb.clen or 0) if ta.clen and tb.clen else None)
        if op == '.eqv.':
            return f'(bool({a}) == bool({b}))', Ty('l')
        if op == '.neqv.':
            return f'(bool({a}) != bool({b}))', Ty('l')
        raise TranslateError('binary ' + op)

    def index_code(self, args, sc, pre):
        """subscript list -> (python subscript text, number of section dims)"""
        parts = []
        nsec = 0
        for a in args:
            if a[0] == 'slice':
                lo = self.expr(a[1], sc, pre)[0] if a[1] is not None else 'None'
                hi = self.expr(a[2], sc, pre)[0] if a[2] is not None else 'None'
                st = self.expr(a[3], sc, pre)[0] if a[3] is not None else 'None'
                parts.append(f'_rt.S({lo}, {hi}, {st})')
                nsec += 1
            elif a[0] == 'kw':
                raise TranslateError('keyword in subscript list')
            else:
                c, t = self.expr(a, sc, pre)
                if t.rank:
                    raise TranslateError('vector subscripts are not supported')
                parts.append(c)
        if len(parts) == 1:
            return parts[0] + ',', nsec
        return ', '.join(parts), nsec

    def call_or_index(self, e, sc: Scope, pre):
        base, args = e[1], e[2]
        if base[0] == 'name':
            name = base[1]
            sf = getattr(self, 'stmt_funcs', {}).get((sc.proc.name if sc.proc else None, name))
            if sf is not None:
                rty, atys = sf
                cs = [self.conv(*self.expr(a, sc, pre), to) for a, to in zip(args, atys)]
                return f'sf_{pyname(name)}({", ".join(cs)})', rty
            r = sc.lookup_var(name)
            if r is not None:
                code, ty = self.name_ref(name, sc)
                return self.subscript(code, ty, args, sc, pre)
            pr = sc.lookup_proc(name)
            if pr is None and name in self.generics:
                pr = self.resolve_generic(name, args, sc, pre)
            if pr is not None:
                if pr.kind != 'function':
                    raise TranslateError(f'subroutine {name} used as a function')
                return self.call_proc(pr, args, sc, pre, as_function=True)
            if name in self.types:
                td = self.types[name]
                cs = []
                for a in args:
                    if a[0] == 'kw':
                        raise TranslateError('keyword structure constructor not supported')
                    cs.append(self.expr(a, sc, pre)[0])
                return f'T_{name}({", ".join(cs)})', Ty('t', 0, name)
            return self.intrinsic(name, args, sc, pre)
        if base[0] == 'comp':
            code, ty = self.expr(base, sc, pre)
            return self.subscript(code, ty, args, sc, pre)
        if base[0] == 'call':
            # e.g. a(i)(1:3) substring -- not supported
            raise TranslateError('substring / double subscript not supported')
        raise TranslateError(f'cannot translate reference {e}')

    def subscript(self, code, ty: Ty, args, sc, pre):
        if ty.rank == 0:
            if ty.base == 'c' and len(args) == 1 and args[0][0] == 'slice':
                a = args[0]
                lo = self.expr(a[1], sc, pre)[0] if a[1] is not None else '1'
                hi = self.expr(a[2], sc, pre)[0] if a[2] is not None else 'None'
                return f'_rt.substr({code}, {lo}, {hi})', Ty('c')
            raise TranslateError(f'subscript on scalar {code}')
        if len(args) != ty.rank:
            raise TranslateError(f'rank mismatch subscripting {code}: {len(args)} vs {ty.rank}')
        idx, nsec = self.index_code(args, sc, pre)
        return f'{code}[{idx}]', Ty(ty.base, nsec, ty.tname, ty.clen)

    def resolve_generic(self, name, args, sc, pre):
        cands = self.generics[name]
        scratch = []
        tys = [self.expr(a[2] if a[0] == 'kw' else a, sc, scratch)[1] for a in args if a[0] != 'slice']
        for pn in cands:
            pr = self.procs.get(pn)
            if pr is None or len(pr.args) < len(tys):
                continue
            if all(_ty_match(pr.decls[x].ty, t) for x, t in zip(pr.args, tys)):
                return pr
        raise TranslateError(f'cannot resolve generic {name}')

    def call_proc(self, pr: Proc, args, sc: Scope, pre, as_function):
        slots: Dict[str, object] = {}
        pos = 0
        for a in args:
            if a[0] == 'kw':
                if a[1] not in pr.args:
                    raise TranslateError(f'{pr.name} has no dummy {a[1]}')
                slots[a[1]] = a[2]
            else:
                if pos >= len(pr.args):
                    raise TranslateError(f'too many arguments calling {pr.name}')
                slots[pr.args[pos]] = a
                pos += 1
        codes = []
        for d in pr.args:
            if d in slots:
                a = slots[d]
                if a[0] == 'slice':
                    raise TranslateError('bare slice as actual argument')
                c, t = self.expr(a, sc, pre)
                dd = pr.decls[d]
                if dd.ty.rank == 0 and dd.ty.base in 'rd' and t.base != dd.ty.base and t.rank == 0:
                    self.warnings.append(f'{pr.name}: actual for {d} has type {t.base}, dummy {dd.ty.base}')
                codes.append(c)
            else:
                if 'optional' not in pr.decls[d].attrs:
                    raise TranslateError(f'missing argument {d} calling {pr.name}')
                codes.append('None')
        call = f'p_{pr.name}({", ".join(codes)})'
        outs = pr.out_scalars()
        rty = pr.decls[pr.result].ty if pr.kind == 'function' else None
        if not outs:
            return call, rty
        tmp = self.newtmp()
        pre.append(f'{tmp} = {call}')
        base = 1 if pr.kind == 'function' else 0
        for k, d in enumerate(outs):
            if d in slots and _is_lvalue(slots[d]):
                r = sc.lookup_var(_root_name(slots[d]))
                if r is None or 'parameter' in r[1].attrs:
                    continue
                dty = pr.decls[d].ty
                pre.extend(self.assign_lines(slots[d], f'{tmp}[{base + k}]', dty, sc, pre))
        if pr.kind == 'function':
            return f'{tmp}[0]', rty
        return tmp, None

    # ---- intrinsic functions ----
    def intrinsic(self, name, args, sc, pre):
        pos = []
        kw = {}
        for a in args:
            if a[0] == 'kw':
                kw[a[1]] = self.expr(a[2], sc, pre)
            elif a[0] == 'slice':
                raise TranslateError('slice as intrinsic argument')
            else:
                pos.append(self.expr(a, sc, pre))
        c = [x[0] for x in pos]
        t = [x[1] for x in pos]

        def prom():
            b = t[0].base
            for x in t[1:]:
                b = _promote(b, x.base)
            return b
        if name in ('abs', 'cabs', 'cdabs') and t[0].base in 'zZ':
            return f'abs({c[0]})', Ty('r' if t[0].base == 'z' else 'd', t[0].rank)
        if name == 'abs':
            return f'abs({c[0]})', t[0]
        if name in ('cmplx', 'dcmplx'):
            knd = c[2] if len(pos) > 2 else (kw['kind'][0] if 'kind' in kw else None)
            wide = name == 'dcmplx' or (knd is not None and knd.strip() == '8')
            im = c[1] if len(pos) > 1 else '0'
            if wide:
                return f'_rt.c128(complex(_rt.f64({c[0]}), _rt.f64({im})))', Ty('Z')
            if t[0].base in 'zZ':
                return f'_rt.c64({c[0]})', Ty('z')
            return f'_rt.c64(complex(_rt.f32({c[0]}), _rt.f32({im})))', Ty('z')
        if name in ('imag', 'aimag', 'dimag'):
            return f'({c[0]}).imag', Ty('r' if t[0].base == 'z' else 'd', t[0].rank)
        if name == 'conjg':
            return f'({c[0]}).conjugate()', t[0]
        if name in ('sqrt', 'log', 'exp', 'sin', 'cos', 'tan', 'acos', 'asin', 'atan', 'log10', 'sinh', 'cosh', 'tanh'):
            if t[0].base not in 'rd':
                raise TranslateError(f'{name} of non-real argument')
            return f'_rt.m_{name}({c[0]})', t[0]
        if name == 'atan2':
            return f'_rt.m_atan2({c[0]}, {c[1]})', t[0]
        if name in ('int', 'ifix', 'idint'):
            return f'_rt.f_int({c[0]})', Ty('i', t[0].rank)
        if name in ('nint', 'idnint'):
            return f'_rt.f_nint({c[0]})', Ty('i', t[0].rank)
        if name in ('floor', 'ceiling'):
            return f'_rt.f_{name}({c[0]})', Ty('i', t[0].rank)
        if name in ('real', 'float', 'sngl'):
            knd = None
            if len(pos) > 1:
                knd = c[1]
            if 'kind' in kw:
                knd = kw['kind'][0]
            if t[0].base in 'zZ':
                # REAL of a complex: the real part, in the kind of the argument unless a kind is given
                wide = (knd.strip() == '8') if knd is not None else t[0].base == 'Z'
                return (f'_rt.f64(({c[0]}).real)', Ty('d', t[0].rank)) if wide else (f'_rt.f32(({c[0]}).real)', Ty('r', t[0].rank))
            if knd is not None and knd.strip() == '8':
                return f'_rt.f64({c[0]})', Ty('d', t[0].rank)
            return f'_rt.f32({c[0]})', Ty('r', t[0].rank)
        if name == 'dble':
            return f'_rt.f64({c[0]})', Ty('d', t[0].rank)
        if name in ('max', 'min', 'amax1', 'amin1', 'max0', 'min0'):
            b = prom()
            fn = 'f_max' if 'max' in name else 'f_min'
            code = f'_rt.{fn}({", ".join(c)})'
            if any(x.base != b for x in t):
                code = f'_rt.conv_{b}({code})'
            return code, Ty(b, max(x.rank for x in t))
        if name == 'mod':
            return f'_rt.f_mod({c[0]}, {c[1]})', Ty(prom())
        if name == 'sign':
            return f'_rt.f_sign({c[0]}, {c[1]})', t[0]
        if name == 'size':
            dim = c[1] if len(c) > 1 else (kw['dim'][0] if 'dim' in kw else 'None')
            return f'_rt.f_size({c[0]}, {dim})', Ty('i')
        if name in ('lbound', 'ubound'):
            dim = c[1] if len(c) > 1 else (kw['dim'][0] if 'dim' in kw else None)
            if dim is None:
                raise TranslateError(name + ' without dim')
            return f'_rt.f_{name}({c[0]}, {dim})', Ty('i')
        if name in ('minloc', 'maxloc'):
            dim = c[1] if len(c) > 1 else (kw['dim'][0] if 'dim' in kw else None)
            mask = c[2] if len(c) > 2 else (kw['mask'][0] if 'mask' in kw else 'None')
            if dim is None:
                if t[0].rank != 1:
                    raise TranslateError(name + ' without dim on rank>1')
                return f'_rt.f_{name}({c[0]}, None, {mask})', Ty('i', 1)
            if t[0].rank != 1:
                raise TranslateError(name + ' with dim on rank>1')
            return f'_rt.f_{name}({c[0]}, {dim}, {mask})', Ty('i')
        if name in ('maxval', 'minval', 'sum', 'product'):
            if len(c) > 1 or kw:
                raise TranslateError(name + ' with dim/mask not supported')
            return f'_rt.f_{name}({c[0]})', t[0].scalar()
        if name in ('any', 'all'):
            return f'_rt.f_{name}({c[0]})', Ty('l')
        if name == 'count':
            return f'_rt.f_count({c[0]})', Ty('i')
        if name == 'present':
            return f'({c[0]} is not None)', Ty('l')
        if name == 'allocated':
            return f'({c[0]} is not None)', Ty('l')
        if name == 'trim':
            return f'_rt.f_trim({c[0]})', Ty('c')
        if name == 'adjustl':
            return f'_rt.f_adjustl({c[0]})', t[0]
        if name in ('len_trim', 'len'):
            return f'_rt.f_{name}({c[0]})', Ty('i')
        if name in ('epsilon', 'huge', 'tiny'):
            return f'_rt.f_{name}({c[0]})', t[0].scalar()
        if name == 'isnan':
            return f'_rt.f_isnan({c[0]})', Ty('l')
        raise TranslateError(f'unknown function or array {name!r} in {sc.proc.name if sc.proc else "<module>"}')

    # ---- statement functions:  f(a, b) = expr  with f a declared scalar ----
    def statement_function(self, s, sc: Scope, ind, out) -> bool:
        lhs = s[2]
        if lhs[0] != 'call' or lhs[1][0] != 'name' or sc.proc is None:
            return False
        name = lhs[1][1]
        d = sc.proc.decls.get(name)
        if d is None or d.ty.rank != 0 or d.ty.base in 'ct' or name in sc.proc.args:
            return False
        if not all(a[0] == 'name' for a in lhs[2]):
            return False
        # dummy arguments are typed by the host's declarations; the nested def shadows those locals
        params = []
        for a in lhs[2]:
            r = sc.lookup_var(a[1])
            if r is None:
                raise TranslateError(f'statement function {name}: dummy {a[1]} undeclared')
            params.append(r[2])
        spre: List[str] = []
        c, t = self.expr(s[3], sc, spre)
        if spre:
            raise TranslateError(f'statement function {name} needs a hoisted call')
        self.stmt_funcs = getattr(self, 'stmt_funcs', {})
        self.stmt_funcs[(sc.proc.name, name)] = (d.ty, [sc.lookup_var(a[1])[1].ty for a in lhs[2]])
        self.emit(out, ind, [f'def sf_{pyname(name)}({", ".join(params)}):', f'    return {self.conv(c, t, d.ty)}'])
        return True

    # ---- assignment ----
    def lvalue_ty(self, lhs, sc, pre):
        return self.expr(lhs, sc, pre)

    def assign_lines(self, lhs, rhs_code, rhs_ty: Optional[Ty], sc: Scope, pre) -> List[str]:
        if lhs[0] == 'paren':
            raise TranslateError('parenthesised lvalue')
        lcode, lty = self.expr(lhs, sc, pre)
        root = _root_name(lhs)
        r = sc.lookup_var(root)
        if r is None:
            raise TranslateError(f'assignment to unknown variable {root}')
        where, d, ename = r
        if 'parameter' in d.attrs:
            raise TranslateError(f'assignment to parameter {root}')
        if lty.rank > 0:
            if lhs[0] == 'call' and any(a[0] == 'slice' for a in lhs[2]):
                return [f'{lcode} = {rhs_code}']
            return [f'{lcode}.setall({rhs_code})']
        if lty.base == 't':
            return [f'{lcode}._assign({rhs_code})']
        val = self.conv(rhs_code, rhs_ty, lty)
        if lhs[0] == 'name' and where == 'host':
            self.nonlocals.add(ename)
        return [f'{lcode} = {val}']

    @staticmethod
    def conv(code, fr: Optional[Ty], to: Ty):
        if to.base == 'i':
            if fr is not None and fr.base == 'i':
                return code
            return f'_rt.f_int({code})'
        if to.base in 'rd' and fr is not None and fr.base in 'zZ':
            code = f'({code}).real'      # complex -> real assignment keeps the real part
        if to.base == 'r':
            if fr is not None and fr.base == 'r' and fr.rank == 0:
                return code
            return f'_rt.f32({code})'
        if to.base == 'd':
            if fr is not None and fr.base == 'd' and fr.rank == 0:
                return code
            return f'_rt.f64({code})'
        if to.base == 'z':
            if fr is not None and fr.base == 'z' and fr.rank == 0:
                return code
            return f'_rt.c64({code})'
        if to.base == 'Z':
            if fr is not None and fr.base == 'Z' and fr.rank == 0:
                return code
            return f'_rt.c128({code})'
        if to.base == 'c':
            if to.clen is None:
                return code
            return f'_rt.fstr({code}, {to.clen})'
        if to.base == 'l':
            return code
        return code

    # ---- statements ----
    def block(self, body, sc: Scope, ind: int, out: List[str], ctx):
        if not body:
            out.append('    ' * ind + 'pass')
            return
        for s in body:
            self.stmt(s, sc, ind, out, ctx)

    def emit(self, out, ind, lines):
        for l in lines:
            out.append('    ' * ind + l)

    def stmt(self, s, sc: Scope, ind: int, out: List[str], ctx):
        k, ln = s[0], s[1]
        pre: List[str] = []
        pname = sc.proc.name if sc.proc else ''
        if k == 'untranslated':
            self.untranslated.setdefault(pname, []).append(f'line {ln}: {s[2]}')
            self.emit(out, ind, [f'_rt.unsupported({s[2]!r}, {ln})'])
            return
        tmp: List[str] = []
        try:
            self._stmt(s, sc, ind, tmp, ctx, pre)
            out.extend(tmp)
        except TranslateError as ex:
            if self.lenient:
                self.untranslated.setdefault(pname, []).append(f'line {ln}: {ex}')
                self.emit(out, ind, [f'_rt.unsupported({str(ex)!r}, {ln})'])
                return
            if 'line ' not in str(ex):
                raise TranslateError(f'line {ln} ({pname}): {ex}') from None
            raise

    def _stmt(self, s, sc, ind, out, ctx, pre):
        k, ln = s[0], s[1]
        if k == 'assign' and self.statement_function(s, sc, ind, out):
            pass
        elif k == 'assign':
            rc, rt_ = self.expr(s[3], sc, pre)
            lines = self.assign_lines(s[2], rc, rt_, sc, pre)
            self.emit(out, ind, pre + lines)
        elif k == 'call':
            self.call_stmt(s, sc, ind, out, pre)
        elif k == 'if':
            branches, else_body = s[2], s[3]
            first = True
            nest = 0
            for cond, body in branches:
                cpre: List[str] = []
                cc, ct = self.expr(cond, sc, cpre)
                if first:
                    self.emit(out, ind, cpre + [f'if {cc}:'])
                    first = False
                elif cpre:
                    # hoisted calls in an else-if condition: nest instead of elif
                    self.emit(out, ind + nest, ['else:'])
                    nest += 1
                    self.emit(out, ind + nest, cpre + [f'if {cc}:'])
                else:
                    self.emit(out, ind + nest, [f'elif {cc}:'])
                self.block(body, sc, ind + nest + 1, out, ctx)
            if else_body is not None:
                self.emit(out, ind + nest, ['else:'])
                self.block(else_body, sc, ind + nest + 1, out, ctx)
        elif k == 'do':
            var, lo, hi, st, body = s[2], s[3], s[4], s[5], s[6]
            loc, lot = self.expr(lo, sc, pre)
            hic, hit = self.expr(hi, sc, pre)
            stc = self.expr(st, sc, pre)[0] if st is not None else '1'
            rng = self.newtmp()
            r = sc.lookup_var(var)
            if r is None:
                raise TranslateError(f'do variable {var} undeclared')
            where, d, vcode = r
            if where == 'host':
                self.nonlocals.add(vcode)
            self.emit(out, ind, pre + [f'{rng} = _rt.frange({loc}, {hic}, {stc})', f'for {vcode} in {rng}:'])
            hook = f'{sc.proc.name}.{var}' if sc.proc else var
            if hook in self.loop_hooks:
                self.emit(out, ind + 1, [f'_rt.loop_hook({hook!r}, {vcode})'])
            self.block(body, sc, ind + 1, out, ctx)
            self.emit(out, ind, ['else:', f'    {vcode} = {rng}.start + len({rng}) * {rng}.step'])
        elif k == 'dowhile':
            cpre: List[str] = []
            cc, _ = self.expr(s[2], sc, cpre)
            if cpre:
                raise TranslateError('function with out-arguments in a do-while condition')
            self.emit(out, ind, [f'while {cc}:'])
            self.block(s[3], sc, ind + 1, out, ctx)
        elif k == 'doforever':
            self.emit(out, ind, ['while True:'])
            self.block(s[2], sc, ind + 1, out, ctx)
        elif k == 'select':
            sel, cases = s[2], s[3]
            sc_, st_ = self.expr(sel, sc, pre)
            tmp = self.newtmp()
            self.emit(out, ind, pre + [f'{tmp} = {sc_}'])
            first = True
            default = None
            for vals, body in cases:
                if vals is None:
                    default = body
                    continue
                conds = []
                for v in vals:
                    if v[0] == 'slice':
                        parts = []
                        if v[1] is not None:
                            parts.append(f'{tmp} >= {self.expr(v[1], sc, pre)[0]}')
                        if v[2] is not None:
                            parts.append(f'{tmp} <= {self.expr(v[2], sc, pre)[0]}')
                        conds.append('(' + ' and '.join(parts) + ')')
                    else:
                        vc, vt = self.expr(v, sc, pre)
                        if st_.base == 'c':
                            conds.append(f"_rt.strcmp({tmp}, {vc}, '==')")
                        else:
                            conds.append(f'{tmp} == {vc}')
                self.emit(out, ind, [('if ' if first else 'elif ') + ' or '.join(conds) + ':'])
                first = False
                self.block(body, sc, ind + 1, out, ctx)
            if default is not None:
                if first:
                    self.emit(out, ind, ['if True:'])
                else:
                    self.emit(out, ind, ['else:'])
                self.block(default, sc, ind + 1, out, ctx)
        elif k == 'return':
            self.emit(out, ind, [ctx['return']])
        elif k == 'exit':
            self.emit(out, ind, ['break'])
        elif k == 'cycle':
            self.emit(out, ind, ['continue'])
        elif k == 'continue':
            self.emit(out, ind, ['pass'])
        elif k == 'stop':
            self.emit(out, ind, [f'raise _rt.FortranStop({(sc.proc.name if sc.proc else "")!r}, {ln})'])
        elif k == 'print':
            items = []
            for x in s[2]:
                try:
                    items.append(self.expr(x, sc, pre)[0])
                except TranslateError:
                    items.append(repr('<untranslated>'))
            self.emit(out, ind, pre + [f'_rt.fprint({ln}, {", ".join(items)})'])
        elif k == 'io':
            self.emit(out, ind, [f'_rt.unsupported({s[2]!r}, {ln})'])
        elif k == 'ionoop':
            self.emit(out, ind, [f'_rt.fio({s[2]!r})'])
        elif k == 'read':
            uc, _ = self.expr(s[2], sc, pre)
            body: List[str] = []
            self.read_items(s[4], sc, pre, body, 1 if s[3] is not None else 0)
            self.emit(out, ind, pre + [f'_rd = _rt.fread_begin({uc}, {ln})'])
            if s[3] is not None:             # iostat: 0, or -1 at the end of the file
                ok = self.assign_lines(s[3], '0', Ty('i'), sc, pre)
                eof = self.assign_lines(s[3], '-1', Ty('i'), sc, pre)
                self.emit(out, ind, ['try:'] + body + ['    _rd.end()'] + ['    ' + l for l in ok] +
                          ['except _rt.FortranEOF:'] + ['    ' + l for l in eof])
            else:
                self.emit(out, ind, body + ['_rd.end()'])
        elif k == 'seek':
            uc, _ = self.expr(s[3], sc, pre)
            self.emit(out, ind, pre + [f'_rt.fseek({uc}, {s[2]!r})'])
        elif k == 'write':
            uc, _ = self.expr(s[2], sc, pre)
            items = [self.io_item(it, sc, pre) for it in s[3]]
            self.emit(out, ind, pre + [f'_rt.fwrite({uc}, [{", ".join(items)}])'])
        elif k == 'allocate':
            lines = []
            for a in s[2]:
                if a[0] == 'kw':
                    if a[1] == 'stat':
                        lines += self.assign_lines(a[2], '0', Ty('i'), sc, pre)
                    continue
                if a[0] != 'call':
                    raise TranslateError('allocate of a scalar')
                target = a[1]
                tcode, tty = self.expr(target, sc, pre)
                dims = []
                for dsp in a[2]:
                    if dsp[0] == 'slice':
                        dims.append(f'({self.expr(dsp[1], sc, pre)[0]}, {self.expr(dsp[2], sc, pre)[0]})')
                    else:
                        dims.append(f'(1, {self.expr(dsp, sc, pre)[0]})')
                ctor = self.alloc_code(tty, dims)
                if target[0] == 'name':
                    where, d, en = sc.lookup_var(target[1])
                    if where == 'host':
                        self.nonlocals.add(en)
                lines.append(f'{tcode} = {ctor}')
            self.emit(out, ind, pre + lines)
        elif k == 'deallocate':
            lines = []
            for a in s[2]:
                if a[0] == 'kw':
                    if a[1] == 'stat':
                        lines += self.assign_lines(a[2], '0', Ty('i'), sc, pre)
                    continue
                tcode, tty = self.expr(a, sc, pre)
                if a[0] == 'name':
                    where, d, en = sc.lookup_var(a[1])
                    if where == 'host':
                        self.nonlocals.add(en)
                lines.append(f'{tcode} = None')
            self.emit(out, ind, pre + lines)
        else:
            raise TranslateError(f'statement kind {k}')

    def io_item(self, it, sc: Scope, pre) -> str:
        if it[0] == 'item':
            return self.expr(it[1], sc, pre)[0]
        _, inner, var, lo, hi, st = it
        r = sc.lookup_var(var)
        if r is None or r[0] != 'local':
            raise TranslateError(f'implied-DO variable {var} must be a local')
        vname = r[2]
        parts = ', '.join(self.io_item(x, sc, pre) for x in inner)
        loc, hic = self.expr(lo, sc, pre)[0], self.expr(hi, sc, pre)[0]
        stc = self.expr(st, sc, pre)[0] if st is not None else '1'
        # the comprehension binds the loop variable in its own scope, like the implied DO
        return f'*[_x for {vname} in _rt.frange({loc}, {hic}, {stc}) for _x in ({parts},)]'

    def read_items(self, items, sc: Scope, pre, out: List[str], ind: int):
        """input items of a list-directed READ: every scalar target takes the next value of the record
        stream; an implied DO becomes a loop over its (local) variable"""
        for it in items:
            if it[0] == 'item':
                lcode, lty = self.expr(it[1], sc, pre)
                if lty.rank > 0:
                    raise TranslateError('whole-array input item')
                want = 'c' if lty.base == 'c' else 'n'
                for l in self.assign_lines(it[1], f'_rd.next({want!r})', None, sc, pre):
                    out.append('    ' * ind + l)
                continue
            _, inner, var, lo, hi, st = it
            r = sc.lookup_var(var)
            if r is None or r[0] != 'local':
                raise TranslateError(f'implied-DO variable {var} must be a local')
            loc, hic = self.expr(lo, sc, pre)[0], self.expr(hi, sc, pre)[0]
            stc = self.expr(st, sc, pre)[0] if st is not None else '1'
            out.append('    ' * ind + f'for {r[2]} in _rt.frange({loc}, {hic}, {stc}):')
            self.read_items(inner, sc, pre, out, ind + 1)

    def alloc_code(self, ty: Ty, dims: List[str]) -> str:
        if ty.base == 't':
            return f'_rt.alloc_obj(T_{ty.tname}, [{", ".join(dims)}])'
        return f'_rt.alloc({ty.base!r}, [{", ".join(dims)}])'

    def call_stmt(self, s, sc: Scope, ind, out, pre):
        name, args = s[2], s[3]
        if name in INTRINSIC_SUBS and sc.lookup_proc(name) is None:
            if name == 'random_number':
                a = args[0][2] if args[0][0] == 'kw' else args[0]
                lc, lt = self.expr(a, sc, pre)
                if lt.rank:
                    self.emit(out, ind, pre + [f'_rt.random_fill({lc})'])
                else:
                    lines = self.assign_lines(a, '_rt.random_number()', Ty('r'), sc, pre)
                    self.emit(out, ind, pre + lines)
            elif name == 'random_seed':
                lines = []
                for a in args:
                    if a[0] == 'kw' and a[1] == 'size':
                        lines += self.assign_lines(a[2], '1', Ty('i'), sc, pre)
                    elif a[0] == 'kw' and a[1] == 'get':
                        lines += [f'{self.expr(a[2], sc, pre)[0]}.setall(0)']
                    elif a[0] != 'kw':
                        lines += self.assign_lines(a, '1', Ty('i'), sc, pre)
                self.emit(out, ind, pre + (lines or ['pass']))
            elif name == 'date_and_time':
                lines = []
                for a in args:
                    if a[0] == 'kw' and a[1] == 'values':
                        lines.append(f'{self.expr(a[2], sc, pre)[0]}.setall(0)')
                self.emit(out, ind, pre + (lines or ['pass']))
            else:
                self.emit(out, ind, ['pass'])
            return
        pr = sc.lookup_proc(name)
        if pr is None and name in self.generics:
            pr = self.resolve_generic(name, args, sc, pre)
        if pr is None and name in self.externs and not args:
            self.emit(out, ind, pre + [f'_rt.call_extern({name!r})'])
            return
        if pr is None and name == 'mpi_allreduce' and len(args) >= 2:
            # one rank: the sum over ranks of the send buffer is the send buffer
            a, _ = self.expr(args[0], sc, pre)
            b, _ = self.expr(args[1], sc, pre)
            self.emit(out, ind, pre + [f'_rt.mpi_allreduce_single({a}, {b})'])
            return
        if pr is None and name == 'mpi_barrier':
            self.emit(out, ind, ['pass'])
            return
        if pr is None:
            if name.startswith('mpi_'):
                self.emit(out, ind, [f'_rt.unsupported({name!r}, {s[1]})'])
                return
            raise TranslateError(f'call to unknown subroutine {name}')
        code, _ = self.call_proc(pr, args, sc, pre, as_function=False)
        if pr.out_scalars():
            self.emit(out, ind, pre)
        else:
            self.emit(out, ind, pre + [code])

    # ---- declarations -> initialisation code ----
    def zero_code(self, d: Decl, sc: Scope, pre) -> Optional[str]:
        ty = d.ty
        if ty.rank > 0:
            if 'allocatable' in d.attrs or 'pointer' in d.attrs:
                return 'None'
            dims = []
            for lo, hi in d.dims:
                if hi is None or hi == '*':
                    return None   # assumed shape / size: a dummy
                loc = self.expr(lo, sc, pre)[0] if lo is not None else '1'
                dims.append(f'({loc}, {self.expr(hi, sc, pre)[0]})')
            return self.alloc_code(ty, dims)
        if ty.base == 'i':
            return '_rt.UNINIT_INT'
        if ty.base == 'r':
            return '_rt.ZERO32'
        if ty.base == 'd':
            return '_rt.ZERO64'
        if ty.base == 'z':
            return '_rt.c64(0)'
        if ty.base == 'Z':
            return '_rt.c128(0)'
        if ty.base == 'l':
            return 'False'
        if ty.base == 'c':
            return repr(' ' * (ty.clen or 1))
        if ty.base == 't':
            return f'T_{ty.tname}()'
        raise TranslateError(f'cannot initialise {d.name}')

    def proc_code(self, pr: Proc, ind: int, out: List[str]):
        tmp: List[str] = []
        saved = getattr(self, 'nonlocals', set())
        try:
            self._proc_code(pr, ind, tmp)
            out.extend(tmp)
        except TranslateError as ex:
            if not self.lenient:
                raise
            self.nonlocals = saved
            self.untranslated.setdefault(pr.name, []).append(f'whole procedure: {ex}')
            self.emit(out, ind, [f'def p_{pr.name}(*a):', f'    _rt.unsupported({("procedure " + pr.name + ": " + str(ex))!r}, {pr.line})', ''])

    def _proc_code(self, pr: Proc, ind: int, out: List[str]):
        sc = Scope(self, pr)
        params = []
        for a in pr.args:
            if a not in pr.decls:
                raise TranslateError(f'{pr.name}: dummy {a} undeclared')
            # absent optional arguments are passed as None; a dummy bound by name to a host
            # variable (see bind_aliases) keeps its slot but is never read
            params.append('_bound_' + a if a in pr.alias else pyname(a))
        self.emit(out, ind, [f'def p_{pr.name}({", ".join(params)}):'])
        body: List[str] = []
        saved_nonlocals = getattr(self, 'nonlocals', set())
        self.nonlocals = set()
        outs = pr.out_scalars()
        res = pyname(pr.result) if pr.kind == 'function' else None
        if pr.kind == 'function':
            if outs:
                ret = f'return ({res}, {", ".join(pyname(o) for o in outs)},)'
            else:
                ret = f'return {res}'
        else:
            ret = f'return ({", ".join(pyname(o) for o in outs)},)' if outs else 'return None'
        ctx = {'return': ret}
        if pr.name in self.proc_hooks:
            self.emit(body, ind + 1, [f'_rt.proc_hook({pr.name!r})'])
        # dummies: rebase arrays to their declared lower bounds
        pre: List[str] = []
        for a in pr.args:
            d = pr.decls[a]
            if d.ty.rank > 0:
                lbs = []
                for lo, hi in d.dims:
                    lbs.append(self.expr(lo, sc, pre)[0] if lo is not None else '1')
                chk = f'if {pyname(a)} is not None: ' if 'optional' in d.attrs else ''
                pre.append(f'{chk}{pyname(a)} = _rt.rebase({pyname(a)}, ({", ".join(lbs)},))')
        # locals
        saved_vars: List[tuple] = []
        for n, d in pr.decls.items():
            if n in pr.args:
                continue
            if 'parameter' in d.attrs:
                c, t = self.expr(d.init, sc, pre)
                pre.append(f'{pyname(n)} = {self.conv(c, t, d.ty)}')
                continue
            z = self.zero_code(d, sc, pre)
            if z is None:
                raise TranslateError(f'{pr.name}: local {n} has assumed shape')
            if d.init is not None or 'save' in d.attrs:
                # SAVE (explicit, or implied by the initialiser): the value survives the call
                key = f'{pr.name}.{n}'
                saved_vars.append((key, pyname(n)))
                pre.append(f'if {key!r} in _SAVE:')
                pre.append(f'    {pyname(n)} = _SAVE[{key!r}]')
                pre.append('else:')
                pre.append(f'    {pyname(n)} = {z}')
                if d.init is not None:
                    ipre: List[str] = []
                    c, t = self.expr(d.init, sc, ipre)
                    if ipre:
                        raise TranslateError(f'{pr.name}: initialiser of {n} needs a call')
                    if d.ty.rank:
                        pre.append(f'    {pyname(n)}.setall({c})')
                    else:
                        pre.append(f'    {pyname(n)} = {self.conv(c, t, d.ty)}')
                continue
            pre.append(f'{pyname(n)} = {z}')
        self.emit(body, ind + 1, pre)
        inner: List[str] = []
        for sub in pr.contains.values():
            self.proc_code(sub, ind + 1, inner)
        stmts: List[str] = []
        if saved_vars:
            self.emit(stmts, ind + 1, ['try:'])
            self.block(pr.body, sc, ind + 2, stmts, ctx)
            self.emit(stmts, ind + 2, [ret])
            self.emit(stmts, ind + 1, ['finally:'] + [f'    _SAVE[{k!r}] = {v}' for k, v in saved_vars])
        else:
            self.block(pr.body, sc, ind + 1, stmts, ctx)
            self.emit(stmts, ind + 1, [ret])
        nl = sorted(self.nonlocals)
        if nl:
            if pr.parent is None:
                raise TranslateError(f'{pr.name}: nonlocal in a module procedure: {nl}')
            out.append('    ' * (ind + 1) + 'nonlocal ' + ', '.join(nl))
        out.extend(body)
        out.extend(inner)
        out.extend(stmts)
        out.append('')
        self.nonlocals = saved_nonlocals

    def type_code(self, td: TypeDef, out: List[str]):
        sc = Scope(self, None)
        names = [pyname(n) for n in td.order]
        out.append(f'class T_{td.name}:')
        out.append(f'    __slots__ = ({", ".join(repr(n) for n in names)},)')
        out.append('    def __init__(self, *a):')
        pre: List[str] = []
        lines = []
        for n in td.order:
            d = td.comps[n]
            z = self.zero_code(d, sc, pre)
            if z is None:
                raise TranslateError(f'type {td.name}: component {n} has assumed shape')
            lines.append(f'self.{pyname(n)} = {z}')
            if d.init is not None:
                c, t = self.expr(d.init, sc, pre)
                lines.append(f'self.{pyname(n)} = {self.conv(c, t, d.ty)}')
        for l in pre + lines:
            out.append('        ' + l)
        out.append('        if a:')
        out.append(f'            assert len(a) == {len(names)}')
        for i, n in enumerate(td.order):
            d = td.comps[n]
            if d.ty.rank > 0:
                out.append(f'            self.{pyname(n)} = _rt.copy_arr(a[{i}])')
            elif d.ty.base == 't':
                out.append(f'            self.{pyname(n)}._assign(a[{i}])')
            else:
                out.append(f'            self.{pyname(n)} = {self.conv(f"a[{i}]", None, d.ty)}')
        out.append('    def _assign(self, o):')
        for n in td.order:
            d = td.comps[n]
            p = pyname(n)
            if d.ty.rank > 0:
                if 'allocatable' in d.attrs or 'pointer' in d.attrs:
                    out.append(f'        self.{p} = _rt.copy_arr(o.{p})')
                else:
                    out.append(f'        self.{p}.setall(o.{p})')
            elif d.ty.base == 't':
                out.append(f'        self.{p}._assign(o.{p})')
            else:
                out.append(f'        self.{p} = o.{p}')
        out.append('    def __repr__(self):')
        out.append(f'        return "{td.name}(" + ", ".join(f"{{n}}={{getattr(self, n)!r}}" for n in self.__slots__) + ")"')
        out.append('')

    def generate(self, header: str = '') -> str:
        out: List[str] = []
        body: List[str] = []
        # types in dependency order
        done = set()

        def emit_type(td: TypeDef):
            if td.name in done:
                return
            done.add(td.name)
            for n in td.order:
                d = td.comps[n]
                if d.ty.base == 't' and d.ty.tname in self.types and d.ty.tname != td.name:
                    emit_type(self.types[d.ty.tname])
            self.type_code(td, body)
        for td in self.types.values():
            try:
                emit_type(td)
            except TranslateError as ex:
                self.warnings.append(f'type {td.name} skipped: {ex}')
                body.append(f'T_{td.name} = None  # not translated: {ex}')
        # module variables
        sc = Scope(self, None)
        body.append('def init_globals():')
        body.append('    _SAVE.clear()')
        body.append('    """module variables: PARAMETERs and initialised variables get their values, the')
        body.append('    rest the zero of their type (allocatables None)."""')
        for m in self.modules:
            body.append(f'    # module {m.name}')
            for n in m.order:
                d = m.decls[n]
                if self.globals.get(n) is not d:
                    continue
                try:
                    pre: List[str] = []
                    lines = []
                    z = self.zero_code(d, sc, pre)
                    if z is None:
                        raise TranslateError('assumed shape module variable')
                    lines.append(f'_G.{pyname(n)} = {z}')
                    if d.init is not None:
                        c, t = self.expr(d.init, sc, pre)
                        if d.ty.rank:
                            lines.append(f'_G.{pyname(n)}.setall({c})')
                        elif d.ty.base == 't':
                            lines.append(f'_G.{pyname(n)}._assign({c})')
                        else:
                            lines.append(f'_G.{pyname(n)} = {self.conv(c, t, d.ty)}')
                    for l in pre + lines:
                        body.append('    ' + l)
                except TranslateError as ex:
                    self.warnings.append(f'module variable {n} not initialised: {ex}')
                    body.append(f'    # {n}: not translated ({ex})')
        body.append('    return _G')
        body.append('')
        for pr in self.procs.values():
            self.proc_code(pr, 0, body)
        out.append(header)
        out.append('import numpy as _np')
        out.append('from oracle.f90ref import rt as _rt')
        out.append('')
        out.append('class _Globals:')
        out.append('    pass')
        out.append('_G = _Globals()')
        out.append('_SAVE = {}      # SAVEd local variables, "procedure.variable" -> value')
        out.append('')
        for (text, kind), name in self.consts.items():
            if kind == 'r':
                out.append(f'{name} = _np.float32({float(text)!r})  # {text}')
            else:
                out.append(f'{name} = _np.float64({float(text)!r})  # {text}d')
        out.append('')
        out.extend(body)
        return '\n'.join(out) + '\n'


def _walk_expr_calls(e):
    if not isinstance(e, tuple) or not e:
        return
    if e[0] == 'call' and isinstance(e[1], tuple) and e[1][0] == 'name':
        yield e[1][1], e[2]
    for x in e[1:]:
        if isinstance(x, tuple):
            yield from _walk_expr_calls(x)
        elif isinstance(x, list):
            for y in x:
                yield from _walk_expr_calls(y)


def _walk_calls(body):
    """(procedure name, actual argument list) of every CALL statement and every
    name(...) reference in a statement list"""
    for s in body:
        if s[0] == 'call':
            yield s[2], s[3]
            for a in s[3]:
                yield from _walk_expr_calls(a)
            continue
        for x in s[2:]:
            yield from _walk_any(x)


def _walk_any(x):
    if isinstance(x, tuple):
        if x and isinstance(x[0], str) and x[0] in ('assign', 'call', 'if', 'do', 'dowhile', 'doforever', 'select',
                                                     'print', 'allocate', 'deallocate') and len(x) > 1 and isinstance(x[1], int):
            yield from _walk_calls([x])
        else:
            yield from _walk_expr_calls(x)
            for y in x:
                if isinstance(y, list):
                    yield from _walk_any(y)
    elif isinstance(x, list):
        for y in x:
            yield from _walk_any(y)


def _promote(a, b):
    order = {'l': 0, 'i': 1, 'r': 2, 'd': 3}
    if a in 'zZ' or b in 'zZ':
        # mixed-mode arithmetic with a complex operand: complex of the wider real kind
        wide = 'Z' in (a, b) or 'd' in (a, b)
        return 'Z' if wide else 'z'
    if a not in order or b not in order:
        return a if a in order else b
    return a if order[a] >= order[b] else b


def _ty_match(dummy: Ty, actual: Ty) -> bool:
    if dummy.base != actual.base:
        return False
    if dummy.base == 't' and dummy.tname != actual.tname:
        return False
    return dummy.rank == actual.rank


def _is_lvalue(e) -> bool:
    if e[0] == 'name':
        return True
    if e[0] == 'comp':
        return _is_lvalue(e[1])
    if e[0] == 'call':
        return e[1][0] in ('name', 'comp') and _is_lvalue(e[1])
    return False


def _root_name(e) -> str:
    while e[0] != 'name':
        e = e[1]
    return e[1]
