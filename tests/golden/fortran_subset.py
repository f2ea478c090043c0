"""A small interpreter for the subset of Fortran 90 the reference's momentum-solver routines are written in.

TEST INFRASTRUCTURE.  It exists for one purpose: to EXECUTE THE REFERENCE'S OWN SOURCE TEXT in the build container -- where
no Fortran compiler exists -- so that golden vectors can be produced by the reference's statements
rather than by anybody's restatement of them.  tests/golden/make_reference_executed_golden.py drives it over
/root/reference/src/shared/*.F and stores inputs and outputs under tests/golden/refexec_*.npz; the tests replay those
fixtures through the oracle (CPU) and the CUDA library (GPU).  Neither the product nor the oracle imports this module,
and /root/reference is read at fixture-generation time only.

What is interpreted, and how it maps onto IEEE arithmetic:
* source handling: the C preprocessor conditionals of the reference (#ifdef / #if defined(..) || .. / #elif / #else /
  #endif) with a given set of defined macros (none: the plain CPU build), `!` comments (so every !$omp / !$acc directive),
  `&` continuation lines; names are case-insensitive;
* statements: assignment (scalars, array elements, whole arrays and `:` sections), pointer assignment `=>`, do / do while /
  exit / cycle, block and one-line if, select case, call, return, allocate / deallocate; declarations are read for
  explicit-shape local arrays and initialisers; `use m, only: a => b` renames are honoured;
* expressions: Fortran precedence (** right-associative, then * /, unary and binary + -, relational, .not., .and., .or.);
  every real operation is one Python float operation, i.e. one IEEE-754 double operation rounded to nearest, evaluated
  left to right as written -- no contraction, no re-association.  x**n with an integer n is expanded into
  multiplications the way compilers do (square-and-multiply: x**2 = x*x); integer / integer truncates;
* intrinsics: sqrt abs max min sign mod real dble int nint sum size sin cos tan asin acos atan atan2 exp log present
  associated trim;
* calls: subroutines found in the loaded source files are interpreted with Fortran argument association (scalars and
  array elements by reference, arrays by descriptor, keyword and absent optional arguments); MPAS framework calls are
  mapped onto the harness: MPAS_pool_get_array / _config / _dimension / _subpool bind a name to what the harness holds for
  (pool, name); timers, logging, halo exchanges of a single block are no-ops and have to be listed as such -- an unknown
  call is an error, never skipped.
Arrays are numpy arrays in this repository's convention (C order with the Fortran dimensions reversed, values 1-based).
"""
from __future__ import annotations

import math
import re

import numpy as np


class FortranError(RuntimeError):
    pass


# ----------------------------------------------------------------------------------------------------------- source

def preprocess(text, defined=()):
    """C-preprocessor conditionals only (the reference uses nothing else inside the routines of interest)."""
    defined = set(defined)

    def cond(expr):
        expr = expr.strip()
        expr = re.sub(r"defined\s*\(\s*(\w+)\s*\)", lambda m: "1" if m.group(1) in defined else "0", expr)
        expr = re.sub(r"defined\s+(\w+)", lambda m: "1" if m.group(1) in defined else "0", expr)
        expr = re.sub(r"\b([A-Za-z_]\w*)\b", lambda m: "1" if m.group(1) in defined else "0", expr)
        expr = expr.replace("||", " or ").replace("&&", " and ").replace("!", " not ")
        return bool(eval(expr, {"__builtins__": {}}, {}))           # digits, and / or / not, parentheses only

    out, stack = [], []          # stack of [taking, taken_before, parent_taking]
    for line in text.split("\n"):
        s = line.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            parent = all(f[0] for f in stack)
            if d.startswith("ifdef"):
                t = parent and d.split()[1] in defined
                stack.append([t, t, parent])
            elif d.startswith("ifndef"):
                t = parent and d.split()[1] not in defined
                stack.append([t, t, parent])
            elif d.startswith("if"):
                t = parent and cond(d[2:])
                stack.append([t, t, parent])
            elif d.startswith("elif"):
                f = stack[-1]
                t = f[2] and not f[1] and cond(d[4:])
                f[0] = t
                f[1] = f[1] or t
            elif d.startswith("else"):
                f = stack[-1]
                f[0] = f[2] and not f[1]
                f[1] = True
            elif d.startswith("endif"):
                stack.pop()
            out.append("")                      # keep line numbers
            continue
        out.append(line if all(f[0] for f in stack) else "")
    return "\n".join(out)


def _strip_comment(line):
    q = None
    for i, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:i]
    return line


def logical_lines(text):
    """[(first line number, statement text)] with comments removed and continuations joined."""
    out, cur, start = [], "", 0
    for no, raw in enumerate(text.split("\n"), 1):
        s = _strip_comment(raw).strip()
        if not s:
            continue
        if cur and s.startswith("&"):
            s = s[1:].lstrip()
        if not cur:
            start = no
        if s.endswith("&"):
            cur += s[:-1] + " "
            continue
        cur += s
        for part in _split_semicolons(cur):
            out.append((start, part.strip()))
        cur = ""
    return out


def _split_semicolons(s):
    if ";" not in s:
        return [s]
    parts, q, last = [], None, 0
    for i, ch in enumerate(s):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == ";":
            parts.append(s[last:i])
            last = i + 1
    parts.append(s[last:])
    return [p for p in parts if p.strip()]


# ------------------------------------------------------------------------------------------------------ expressions

_TOKEN = re.compile(r"""
    \s*(?:
      (?P<num>(?:\d+\.(?![A-Za-z]+\.)\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?(?:_\w+)?)
    | (?P<dot>\.(?:and|or|not|eqv|neqv|eq|ne|lt|le|gt|ge|true|false)\.)
    | (?P<name>[A-Za-z_]\w*)
    | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
    | (?P<op>\*\*|==|/=|<=|>=|=>|//|[-+*/(),:%<>=\[\]])
    )""", re.X | re.I)

_DOT_REL = {".eq.": "==", ".ne.": "/=", ".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">="}


def tokenize(s):
    toks, pos = [], 0
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m or m.end() == pos:
            if s[pos:].strip() == "":
                break
            raise FortranError("cannot tokenise %r at %r" % (s, s[pos:pos + 20]))
        pos = m.end()
        if m.group("num") is not None:
            t = m.group("num").lower()
            t = re.sub(r"_\w+$", "", t)
            if re.fullmatch(r"\d+", t):
                toks.append(("num", int(t)))
            else:
                toks.append(("num", float(t.replace("d", "e"))))
        elif m.group("dot") is not None:
            d = m.group("dot").lower()
            if d == ".true.":
                toks.append(("num", True))
            elif d == ".false.":
                toks.append(("num", False))
            else:
                toks.append(("op", _DOT_REL.get(d, d)))
        elif m.group("name") is not None:
            toks.append(("name", m.group("name").lower()))
        elif m.group("str") is not None:
            toks.append(("str", m.group("str")[1:-1]))
        else:
            toks.append(("op", m.group("op")))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("end", None)

    def take(self, kind=None, val=None):
        tok = self.peek()
        if (kind and tok[0] != kind) or (val is not None and tok[1] != val):
            raise FortranError("expected %s %s, found %s" % (kind, val, tok))
        self.i += 1
        return tok

    def at_op(self, *ops):
        tok = self.peek()
        return tok[0] == "op" and tok[1] in ops

    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.at_op(".or.", ".eqv.", ".neqv."):
            op = self.take()[1]
            a = ("bin", op, a, self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.at_op(".and."):
            self.take()
            a = ("bin", ".and.", a, self.p_not())
        return a

    def p_not(self):
        if self.at_op(".not."):
            self.take()
            return ("un", ".not.", self.p_not())
        return self.p_rel()

    def p_rel(self):
        a = self.p_add()
        if self.at_op("==", "/=", "<", "<=", ">", ">="):
            op = self.take()[1]
            a = ("bin", op, a, self.p_add())
        return a

    def p_add(self):
        if self.at_op("+", "-"):
            op = self.take()[1]
            a = ("un", op, self.p_mul())
        else:
            a = self.p_mul()
        while self.at_op("+", "-", "//"):
            op = self.take()[1]
            a = ("bin", op, a, self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while self.at_op("*", "/"):
            if self.peek()[1] == "/" and self.i + 1 < len(self.t) and self.t[self.i + 1] == ("op", ")"):
                break                                        # the closing "/)" of an array constructor
            op = self.take()[1]
            a = ("bin", op, a, self.p_pow())
        return a

    def p_pow(self):
        a = self.p_primary()
        if self.at_op("**"):
            self.take()
            if self.at_op("+", "-"):                       # x ** -n
                op = self.take()[1]
                b = ("un", op, self.p_pow())
            else:
                b = self.p_pow()                             # right-associative
            a = ("bin", "**", a, b)
        return a

    def p_args(self):
        """after '(' : list of (keyword or None, node); ':' sections become ('slice', lo, hi)"""
        args = []
        if self.at_op(")"):
            self.take()
            return args
        while True:
            kw = None
            if self.peek()[0] == "name" and self.i + 1 < len(self.t) and self.t[self.i + 1] == ("op", "=") and \
                    not (self.i + 2 < len(self.t) and self.t[self.i + 2] == ("op", "=")):
                kw = self.take()[1]
                self.take("op", "=")
            if self.at_op(":"):
                self.take()
                hi = None if self.at_op(",", ")") else self.expr()
                node = ("slice", None, hi)
            else:
                node = self.expr()
                if self.at_op(":"):
                    self.take()
                    hi = None if self.at_op(",", ")") else self.expr()
                    node = ("slice", node, hi)
            args.append((kw, node))
            if self.at_op(","):
                self.take()
                continue
            self.take("op", ")")
            return args

    def p_primary(self):
        tok = self.peek()
        if tok[0] == "num":
            self.take()
            return ("num", tok[1])
        if tok[0] == "str":
            self.take()
            return ("str", tok[1])
        if tok == ("op", "["):                                # [a, b, c]: the Fortran 2003 array constructor
            self.take()
            items = []
            while not self.at_op("]"):
                items.append(self.expr())
                if self.at_op(","):
                    self.take()
            self.take("op", "]")
            return ("array", items)
        if tok == ("op", "(") and self.i + 1 < len(self.t) and self.t[self.i + 1] == ("op", "/"):
            self.take()
            self.take()
            items = []
            while not self.at_op("/"):
                items.append(self.expr())
                if self.at_op(","):
                    self.take()
            self.take("op", "/")
            self.take("op", ")")
            return ("array", items)
        if tok == ("op", "("):
            self.take()
            a = self.expr()
            self.take("op", ")")
            return ("paren", a)
        if tok[0] == "name":
            self.take()
            node = ("name", tok[1])
            while True:
                if self.at_op("("):
                    self.take()
                    node = ("call", node, self.p_args())
                elif self.at_op("%"):
                    self.take()
                    node = ("comp", node, self.take("name")[1])
                else:
                    return node
        raise FortranError("unexpected token %s" % (tok,))


def parse_expr(s):
    p = Parser(tokenize(s))
    e = p.expr()
    if p.peek()[0] != "end":
        raise FortranError("trailing tokens in %r: %s" % (s, p.t[p.i:]))
    return e


# ------------------------------------------------------------------------------------------------------------ values

class FArray:
    """A Fortran array over a numpy array kept in this repository's layout (dimensions reversed), 1-based."""

    def __init__(self, a):
        self.a = a

    def _ix(self, idx):
        ix = []
        for i in reversed(idx):
            if isinstance(i, slice):
                ix.append(slice(None if i.start is None else i.start - 1, i.stop))
            else:
                ix.append(int(i) - 1)
        if len(ix) != self.a.ndim:
            raise FortranError("rank mismatch: %d subscripts for an array of rank %d" % (len(ix), self.a.ndim))
        for k, i in enumerate(ix):
            if not isinstance(i, slice) and not 0 <= i < self.a.shape[k]:
                raise FortranError("subscript %d out of bounds (extent %d)" % (i + 1, self.a.shape[k]))
        return tuple(ix)

    def get(self, idx):
        v = self.a[self._ix(idx)]
        if isinstance(v, np.ndarray):
            return FArray(v)
        return _scalar(v)

    def set(self, idx, val):
        if self.a.dtype == object:
            ix = self._ix(idx)
            if any(isinstance(i, slice) for i in ix):
                self.a[ix] = val.a if isinstance(val, FArray) else val
            else:
                self.a[ix] = val                   # an element of an array of objects / addresses: kept as it is
            return
        self.a[self._ix(idx)] = val.a if isinstance(val, FArray) else val


def _scalar(v):
    if isinstance(v, (np.floating, float)):
        return float(v)
    if not isinstance(v, (np.integer, int, np.bool_, bool)):
        return v                                   # an element of an array of derived-type objects
    if isinstance(v, (np.bool_, bool)):
        return bool(v)
    return int(v)


class ElemRef:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def get(self):
        return self.arr.get(self.idx)

    def set(self, v):
        self.arr.set(self.idx, v)


class VarRef:
    def __init__(self, frame, name):
        self.frame, self.name = frame, name

    def get(self):
        return self.frame.get(self.name)

    def set(self, v):
        self.frame.set(self.name, v)

    def bind(self, v):
        cur = self.frame.vars.get(self.name)
        if isinstance(cur, VarRef):
            cur.bind(v)
        else:
            self.frame.bind(self.name, v)


class CompRef:
    def __init__(self, obj, field):
        self.obj, self.field = obj, field

    def get(self):
        return getattr(self.obj, self.field)

    def set(self, v):
        setattr(self.obj, self.field, v)

    def bind(self, v):
        setattr(self.obj, self.field, v)


ABSENT = object()


def powi(x, n):
    """x**n for an integer n by square-and-multiply, as compilers expand it (x**2 = x*x, x**3 = x*(x*x))."""
    if n < 0:
        return 1.0 / powi(x, -n)
    y = x if n & 1 else (1 if isinstance(x, int) else 1.0)
    n >>= 1
    while n:
        x = x * x
        if n & 1:
            y = y * x
        n >>= 1
    return y


def _idiv(a, b):
    if isinstance(a, int) and isinstance(b, int) and not isinstance(a, bool):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    return a / b


def _sum(x, *rest):
    if isinstance(x, FArray):
        tot = 0.0 if x.a.dtype.kind == "f" else 0
        for v in x.a.T.reshape(-1):            # Fortran element order
            tot = tot + _scalar(v)
        return tot
    raise FortranError("sum() of a scalar")


def _size(x, dim=None):
    if dim is None:
        return int(x.a.size)
    return int(x.a.shape[x.a.ndim - int(dim)])


def _dot_product(a, b):
    s = 0.0
    for x, y in zip(a.a.reshape(-1), b.a.reshape(-1)):     # rank-1 arguments: one multiply and one add per element, in order
        s = s + float(x) * float(y)
    return s


def _spread(x, dim, ncopies):
    if x.a.ndim != 1:
        raise FortranError("spread() of a rank-%d array" % x.a.ndim)
    if dim == 2:                                            # result(i, j) = x(i): numpy layout [j][i]
        return FArray(np.tile(x.a, (int(ncopies), 1)))
    return FArray(np.tile(x.a[:, None], (1, int(ncopies))))  # dim = 1: result(i, j) = x(j): numpy layout [j][i]


def _matmul(a, b):
    """matmul of a rank-2 and a rank-1 (or two rank-2) arrays: result(i[,k]) = sum over j of a(i,j) * b(j[,k]), the
    products added in increasing j starting from zero -- the loop a compiler inlines for small fixed sizes."""
    A = a.a.T                                   # A[i, j] in Fortran index order
    if b.a.ndim == 1:
        out = np.zeros(A.shape[0])
        for i in range(A.shape[0]):
            s = 0.0
            for j in range(A.shape[1]):
                s = s + float(A[i, j]) * float(b.a[j])
            out[i] = s
        return FArray(out)
    B = b.a.T
    out = np.zeros((A.shape[0], B.shape[1]))
    for i in range(A.shape[0]):
        for k in range(B.shape[1]):
            s = 0.0
            for j in range(A.shape[1]):
                s = s + float(A[i, j]) * float(B[j, k])
            out[i, k] = s
    return FArray(out.T.copy())


def _minval(x, dim=None):
    if dim is None:
        return _scalar(x.a.min())
    return FArray(x.a.min(axis=x.a.ndim - int(dim)))


def _maxval(x, dim=None):
    if dim is None:
        return _scalar(x.a.max())
    return FArray(x.a.max(axis=x.a.ndim - int(dim)))


INTRINSICS = {
    "huge": lambda x: float(np.finfo(np.float64).max) if isinstance(x, float) else 2147483647,
    "tiny": lambda x: float(np.finfo(np.float64).tiny), "epsilon": lambda x: float(np.finfo(np.float64).eps),
    "transpose": lambda x: FArray(x.a.T.copy()),
    "merge": lambda a, b, mask: a if mask else b,
    "c_loc": lambda x: x,                                   # the array (or object) itself stands for its address
    "c_associated": lambda p, *q: p is not None and p != 0,
    "dot_product": _dot_product, "spread": _spread, "maxval": _maxval, "minval": _minval, "matmul": _matmul,
    "maxloc": lambda x: FArray(np.array([int(np.argmax(x.a.T.reshape(-1))) + 1], dtype=np.int64)),
    "any": lambda x: bool(np.any(x.a)) if isinstance(x, FArray) else bool(x),
    "all": lambda x: bool(np.all(x.a)) if isinstance(x, FArray) else bool(x),
    "sqrt": math.sqrt, "abs": abs, "max": max, "min": min,
    "sign": lambda a, b: math.copysign(abs(a), b) if isinstance(a, float) or isinstance(b, float) else (abs(a) if b >= 0 else -abs(a)),
    "mod": lambda a, b: math.fmod(a, b) if isinstance(a, float) else int(math.fmod(a, b)),
    "modulo": lambda a, b: a - b * math.floor(a / b) if isinstance(a, float) or isinstance(b, float) else a % b,
    "real": lambda a, *k: float(a), "dble": float, "int": lambda a, *k: int(a), "nint": lambda a: int(math.floor(abs(a) + 0.5)) * (1 if a >= 0 else -1),      # half away from zero (not Python's round)
    "sin": math.sin, "cos": math.cos, "tan": math.tan, "asin": math.asin, "acos": math.acos, "atan": math.atan,
    "atan2": math.atan2, "exp": math.exp, "log": math.log, "sum": _sum, "size": _size, "trim": lambda s: s.rstrip(),
}


# elementwise on whole arrays / sections (numpy's float64 sqrt, abs, maximum, minimum are the same IEEE operations)
ARRAY_INTRINSICS = {"sqrt": np.sqrt, "abs": np.abs, "max": np.maximum, "min": np.minimum}


class _Exit(Exception):
    pass


class _Cycle(Exception):
    pass


class _Return(Exception):
    pass


# --------------------------------------------------------------------------------------------------------- program

class Sub:
    def __init__(self, name, args, lines, file):
        self.name, self.args, self.file = name, args, file
        self.renames, self.decl_arrays, self.inits, self.body = {}, [], [], None
        self.lines = lines
        self.result = None         # name of the result variable of a function
        self.types = {}            # declared type of the names: "int" | "real" | "log" (assignment converts)
        self.decl_objects = []     # local variables of a derived type: (name, type name)


_DECL = re.compile(r"^(real\b|integer\b|logical\b|character\b|double\s+precision\b|type\s*\(|class\s*\()", re.I)


class Frame:
    def __init__(self, interp, sub):
        self.interp, self.sub, self.vars = interp, sub, {}

    def lookup(self, name):
        name = self.sub.renames.get(name, name) if name not in self.vars else name
        if name in self.vars:
            return self.vars[name]
        if name in self.interp.globals:
            return self.interp.globals[name]
        raise FortranError("%s: undefined variable %r" % (self.sub.name, name))

    def has(self, name):
        return name in self.vars or self.sub.renames.get(name, name) in self.interp.globals

    def get(self, name):
        v = self.lookup(name)
        if isinstance(v, (ElemRef, VarRef, CompRef)):
            return v.get()
        return v

    def set(self, name, val):
        t = self.sub.types.get(name)
        if t == "int" and isinstance(val, float):
            val = int(val)                                   # assignment to an integer truncates
        elif t == "real" and isinstance(val, int) and not isinstance(val, bool):
            val = float(val)
        if name in self.vars:
            cur = self.vars[name]
            if isinstance(cur, (ElemRef, VarRef, CompRef)):
                cur.set(val)
            elif isinstance(cur, FArray) and not isinstance(val, FArray):
                cur.a[...] = val
            elif isinstance(cur, FArray) and isinstance(val, FArray):
                cur.a[...] = val.a
            else:
                self.vars[name] = val
            return
        g = self.sub.renames.get(name, name)
        if g in self.interp.globals:
            cur = self.interp.globals[g]
            if isinstance(cur, FArray):
                cur.a[...] = val.a if isinstance(val, FArray) else val
            else:
                self.interp.globals[g] = val
            return
        self.vars[name] = val

    def bind(self, name, val):
        """pointer association / first definition: replaces whatever the name held.  A name the routine does not declare
        but its module does (a module-level pointer, e.g. special_boundaries.F:29-32) is associated at module level."""
        if name not in self.vars and name not in self.sub.types and name not in self.sub.args:
            g = self.sub.renames.get(name, name)
            if g in self.interp.globals:
                self.interp.globals[g] = val
                return
        self.vars[name] = val


class Interpreter:
    POOL_GETTERS = ("mpas_pool_get_array", "mpas_pool_get_config", "mpas_pool_get_dimension", "mpas_pool_get_subpool",
                    "mpas_pool_get_field", "mpas_pool_get_package")

    def __init__(self, defined=()):
        self.defined = tuple(defined)
        self.subs = {}
        self.globals = {"rkind": 8, "strkind": 512,   # module variables (lower-case names); the MPAS kind parameters
                        "c_int": 4, "c_double": 8, "c_char": 1, "c_null_ptr": None, "c_null_char": "\0"}   # iso_c_binding
        self.types = {}            # derived types: name -> [(component, kind, dims or None)], kind in int/real/log/ptr/char/type:<name>
        self.cfuncs = {}           # bind(C) interfaces: name -> dict(args=[...], decl={arg: (kind, value, out, array)}, result=kind)
        self.func_hooks = {}       # name -> python callable(interp, frame, args) -> value, for functions called in expressions
        self.module_objects = []   # module-level variables of derived type, created by resolve_constants
        self.pool = {}             # (pool name or None, variable name) -> value for the MPAS_pool_get_* calls
        self.noop = set()          # framework calls that do nothing on a single block
        self.hooks = {}            # name -> python callable(interp, frame, args) replacing a call
        self.trace = []            # names of the interpreted subroutines, in call order
        self.pending = []          # module-level initialisers not resolved yet
        self.component_types = {}  # lower-case component name -> "int" for integer components allocated by the source

    # ---- loading
    def load(self, path):
        text = preprocess(open(path).read(), self.defined)
        lines = logical_lines(text)
        self._derived_types(lines)
        self._module_constants(lines)
        i = 0
        in_interface = False
        while i < len(lines):
            no, s = lines[i]
            if re.match(r"^interface\b", s, re.I):
                in_interface = True
            elif re.match(r"^end\s*interface\b", s, re.I):
                in_interface = False
            if in_interface:
                mi = re.match(r"^(?:function|subroutine)\s+(\w+)\s*\((.*?)\)\s*(?:bind\s*\(.*?\))?\s*(?:result\s*\(\s*(\w+)\s*\))?\s*$", s, re.I)
                if mi:
                    j = i + 1
                    while not re.match(r"^end\s*(subroutine|function)\b", lines[j][1], re.I):
                        j += 1
                    self._c_interface(mi, lines[i + 1:j])
                    i = j
                i += 1
                continue
            m = re.match(r"^(?:recursive\s+|pure\s+|elemental\s+)*subroutine\s+(\w+)\s*(?:\((.*)\))?\s*$", s, re.I)
            mf = None if m else re.match(
                r"^(?:recursive\s+|pure\s+|elemental\s+|real\s*(?:\([^)]*\))?\s+|integer\s+|logical\s+|double\s+precision\s+)*"
                r"function\s+(\w+)\s*\((.*?)\)\s*(?:result\s*\(\s*(\w+)\s*\))?\s*$", s, re.I)
            if m or mf:
                mm = m or mf
                name = mm.group(1).lower()
                args = [a.strip().lower() for a in (mm.group(2) or "").split(",") if a.strip()]
                j = i + 1
                while not re.match(r"^end\s*(subroutine|function)\b", lines[j][1], re.I):
                    j += 1
                self.subs[name] = Sub(name, args, lines[i + 1:j], path)
                if mf:
                    self.subs[name].result = (mf.group(3) or name).lower()
                i = j
            i += 1

    @staticmethod
    def _decl_kind(attrs_l):
        a = attrs_l.strip()
        if a.startswith("integer"):
            return "int"
        if a.startswith("real") or a.startswith("double"):
            return "real"
        if a.startswith("logical"):
            return "log"
        if a.startswith("character"):
            return "char"
        m = re.match(r"^type\s*\(\s*(\w+)\s*\)", a)
        if m:
            return "ptr" if m.group(1) == "c_ptr" else "type:" + m.group(1)
        return "real"

    def _derived_types(self, lines):
        i = 0
        while i < len(lines):
            m = re.match(r"^type\s*(?:,[^:]*)?::\s*(\w+)\s*$", lines[i][1], re.I) or re.match(r"^type\s+(\w+)\s*$", lines[i][1], re.I)
            if m and not re.match(r"^type\s*\(", lines[i][1], re.I):
                name, comps, j = m.group(1).lower(), [], i + 1
                while not re.match(r"^end\s*type\b", lines[j][1], re.I):
                    d = lines[j][1]
                    if "::" in d:
                        attrs, items = d.split("::", 1)
                        kind = self._decl_kind(attrs.lower())
                        pointer = "pointer" in attrs.lower()
                        for part in _split_top(items):
                            mm = re.match(r"^(\w+)\s*(?:\((.*?)\))?", part.strip())
                            dims = mm.group(2) if (mm.group(2) and ":" not in mm.group(2)) else None
                            comps.append((mm.group(1).lower(), "ptr" if pointer else kind, dims))
                    j += 1
                self.types[name] = comps
                i = j
            i += 1

    def new_object(self, typename, fr=None):
        import types as _types
        obj = _types.SimpleNamespace(_type=typename)
        for comp, kind, dims in self.types[typename]:
            if dims is not None:
                n = [int(self.ev(parse_expr(x), fr or Frame(self, Sub("<type>", [], [], "")))) for x in _split_top(dims)]
                if kind in ("int", "real", "log"):
                    dt = {"int": np.int64, "real": np.float64, "log": np.bool_}[kind]
                    val = FArray(np.zeros(tuple(reversed(n)), dtype=dt))
                else:
                    val = FArray(np.array([None] * int(np.prod(n)), dtype=object).reshape(tuple(reversed(n))))
            elif kind.startswith("type:") and kind[5:] in self.types:
                val = self.new_object(kind[5:], fr)
            else:
                val = {"int": 0, "real": 0.0, "log": False, "char": ""}.get(kind)
            setattr(obj, comp, val)
        return obj

    def _c_interface(self, m, decl_lines):
        name = m.group(1).lower()
        args = [a.strip().lower() for a in m.group(2).split(",") if a.strip()]
        result = (m.group(3) or name).lower()
        decl = {}
        for no, d in decl_lines:
            if "::" not in d:
                continue
            attrs, items = d.split("::", 1)
            al = attrs.lower()
            kind = self._decl_kind(al)
            for part in _split_top(items):
                mm = re.match(r"^(\w+)\s*(\(.*\))?", part.strip())
                decl[mm.group(1).lower()] = dict(kind=kind, value="value" in al, out="intent(out)" in al.replace(" ", "") or
                                                 "intent(inout)" in al.replace(" ", ""),
                                                 array=bool(mm.group(2)) or "dimension" in al)
        self.cfuncs[name] = dict(args=args, decl=decl, result=decl.get(result, dict(kind="int"))["kind"])

    def _module_constants(self, lines):
        """Module-level declarations with an initialiser (`integer, parameter :: A = 1`, `real(kind=RKIND), parameter ::
        puny = 1.0e-11_RKIND`) become globals, evaluated from the reference text; an initialiser that names something not
        known yet is retried after the other files are loaded (resolve_constants)."""
        for no, s in lines:
            if re.match(r"^contains\b", s, re.I):
                break
            if not (_DECL.match(s) and "::" in s):
                continue
            attrs, items = s.split("::", 1)
            kind = self._decl_kind(attrs.lower())
            dim = re.search(r"dimension\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)", attrs.lower())
            for part in _split_top(items):
                m = re.match(r"^(\w+)\s*(?:\(.*?\))?\s*=(?!>)\s*(.*)$", part.strip())
                if m:
                    self.pending.append((m.group(1).lower(), m.group(2)))
                elif kind.startswith("type:") and kind[5:] in self.types and "pointer" not in attrs.lower():
                    mm = re.match(r"^(\w+)\s*(?:\((.*?)\))?", part.strip())
                    self.module_objects.append((mm.group(1).lower(), kind[5:], mm.group(2) or (dim.group(1) if dim else None)))
                else:
                    # a module variable without an initialiser: it exists (so that assignments inside routines reach it)
                    mm = re.match(r"^(\w+)\s*(\(.*\))?$", part.strip())
                    if mm and not mm.group(2) and dim is None and "allocatable" not in attrs.lower() and \
                            "pointer" not in attrs.lower() and mm.group(1).lower() not in self.globals:
                        self.globals[mm.group(1).lower()] = {"int": 0, "real": 0.0, "log": False, "char": ""}.get(kind)
        self.resolve_constants()

    def resolve_constants(self):
        fr = Frame(self, Sub("<module>", [], [], ""))
        progress = True
        while progress and self.pending:
            progress, rest = False, []
            for name, init in self.pending:
                try:
                    self.globals[name] = self.ev(parse_expr(init), fr)
                    progress = True
                except FortranError:
                    rest.append((name, init))
            self.pending = rest
        for name, typename, dims in self.module_objects:
            if name in self.globals:
                continue
            try:
                if dims is None:
                    self.globals[name] = self.new_object(typename, fr)
                else:
                    n = int(self.ev(parse_expr(dims), fr))
                    self.globals[name] = FArray(np.array([self.new_object(typename, fr) for _ in range(n)], dtype=object))
            except FortranError:
                pass

    def _prepare(self, sub):
        if sub.body is not None:
            return
        stmts = []
        for no, s in sub.lines:
            low = s.lower()
            if low.startswith("use "):
                for a, b in re.findall(r"(\w+)\s*=>\s*(\w+)", s):
                    sub.renames[a.lower()] = b.lower()
                continue
            if low.startswith("implicit ") or low.startswith("save") or low.startswith("private") or low.startswith("public"):
                continue
            if _DECL.match(s) and "::" in s and not re.match(r"^\w+\s*(\(.*\))?\s*=[^=]", s):
                self._declaration(sub, s)
                continue
            stmts.append((no, s))
        sub.body = self._block(sub, stmts, 0, ())[0]

    def _declaration(self, sub, s):
        attrs, items = s.split("::", 1)
        attrs_l = attrs.lower()
        dim = re.search(r"dimension\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)", attrs_l)
        deferred = "pointer" in attrs_l or "allocatable" in attrs_l
        kind = "int" if attrs_l.startswith("integer") else ("log" if attrs_l.startswith("logical") else "real")
        depth, cur, parts = 0, "", []
        for ch in items:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        for part in parts:
            part = part.strip()
            if not part:
                continue
            m = re.match(r"^(\w+)\s*(?:\((.*?)\))?\s*(?:=(?!>)\s*(.*))?$", part)
            if not m:
                continue
            name, own_dim, init = m.group(1).lower(), m.group(2), m.group(3)
            if not attrs_l.startswith(("type", "class", "character")):
                sub.types[name] = kind
            tk = self._decl_kind(attrs_l)
            if tk.startswith("type:") and tk[5:] in self.types and not deferred and name not in sub.args:
                sub.decl_objects.append((name, tk[5:]))
                continue
            shape = own_dim if own_dim else (dim.group(1) if dim else None)
            if shape is not None and not deferred and ":" not in shape and name not in sub.args:
                sub.decl_arrays.append((name, [parse_expr(x) for x in _split_top(shape)], kind))
            if init is not None and "parameter" in attrs_l:
                sub.inits.append((name, parse_expr(init)))

    # ---- statements -> tree
    def _block(self, sub, stmts, i, enders):
        """parse statements from i until one whose keyword is in `enders`; returns (list of nodes, index of the ender)"""
        nodes = []
        while i < len(stmts):
            no, s = stmts[i]
            norm = re.sub(r"\s+", " ", s).strip()                      # case kept: string literals are compared as written
            norm = re.sub(r"^\w+\s*:\s*(?=do\b|if\b)", "", norm, flags=re.I)   # construct names
            low = norm.lower()
            key = self._ender_key(low)
            if key in enders:
                return nodes, i
            m = re.match(r"^do\s+while\s*\((.*)\)$", norm, re.I)
            if m:
                body, j = self._block(sub, stmts, i + 1, ("enddo",))
                nodes.append(("dowhile", no, parse_expr(m.group(1)), body))
                i = j + 1
                continue
            m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", norm, re.I)
            if m:
                parts = _split_top(m.group(2))
                body, j = self._block(sub, stmts, i + 1, ("enddo",))
                nodes.append(("do", no, m.group(1).lower(), [parse_expr(p) for p in parts], body))
                i = j + 1
                continue
            if low == "do":
                body, j = self._block(sub, stmts, i + 1, ("enddo",))
                nodes.append(("dowhile", no, ("num", True), body))
                i = j + 1
                continue
            m = re.match(r"^if\s*\((.*)\)\s*then$", norm, re.I)
            if m:
                branches, cond = [], parse_expr(m.group(1))
                j = i
                while True:
                    body, j = self._block(sub, stmts, j + 1, ("elseif", "else", "endif"))
                    branches.append((cond, body))
                    n2 = re.sub(r"\s+", " ", stmts[j][1]).strip()
                    l2 = n2.lower()
                    k2 = self._ender_key(l2)
                    if k2 == "elseif":
                        cond = parse_expr(re.match(r"^else\s*if\s*\((.*)\)\s*then$", n2, re.I).group(1))
                    elif k2 == "else":
                        cond = ("num", True)
                    else:
                        break
                nodes.append(("if", no, branches))
                i = j + 1
                continue
            m = re.match(r"^select\s*case\s*\((.*)\)$", norm, re.I)
            if m:
                sel, cases, j = parse_expr(m.group(1)), [], i + 1
                while self._ender_key(re.sub(r"\s+", " ", stmts[j][1].lower()).strip()) != "endselect":
                    l2 = re.sub(r"\s+", " ", stmts[j][1]).strip()
                    mc = re.match(r"^case\s*(default|\((.*)\))$", l2, re.I)
                    if not mc:
                        raise FortranError("line %d: expected case, found %r" % (stmts[j][0], l2))
                    if mc.group(1).lower() == "default":
                        vals = None
                    else:                      # values and ranges lo:hi (either bound may be missing)
                        vals = []
                        for x in _split_top(mc.group(2)):
                            if ":" in x and not re.search(r"['\"]", x):
                                lo, hi = [y.strip() for y in x.split(":", 1)]
                                vals.append(("range", parse_expr(lo) if lo else None, parse_expr(hi) if hi else None))
                            else:
                                vals.append(parse_expr(x))
                    body, j = self._block(sub, stmts, j + 1, ("case", "endselect"))
                    cases.append((vals, body))
                nodes.append(("select", no, sel, cases))
                i = j + 1
                continue
            m = re.match(r"^if\s*\(", low)
            if m:
                close = _matching_paren(norm, norm.index("("))
                cond, rest = parse_expr(norm[norm.index("(") + 1:close]), s[_matching_paren(s, s.index("(")) + 1:].strip()
                nodes.append(("if", no, [(cond, [self._simple(sub, no, rest)])]))
                i += 1
                continue
            nodes.append(self._simple(sub, no, s))
            i += 1
        if enders:
            raise FortranError("%s: missing %s" % (sub.name, enders))
        return nodes, i

    @staticmethod
    def _ender_key(low):
        if re.match(r"^end\s*do\b", low):
            return "enddo"
        if re.match(r"^end\s*if\b", low):
            return "endif"
        if re.match(r"^else\s*if\s*\(", low):
            return "elseif"
        if re.match(r"^else\b", low):
            return "else"
        if re.match(r"^end\s*select\b", low):
            return "endselect"
        if re.match(r"^case\b", low):
            return "case"
        return None

    def _simple(self, sub, no, s):
        low = s.lower().strip()
        if low in ("exit", "cycle", "return", "continue"):
            return (low, no)
        m = re.match(r"^call\s+(\w+)\s*(?:\((.*)\))?$", s.strip(), re.I | re.S)
        if m:
            args = Parser(tokenize("(" + (m.group(2) or "") + ")"))
            args.take("op", "(")
            return ("call", no, m.group(1).lower(), args.p_args())
        m = re.match(r"^(allocate|deallocate)\s*\((.*)\)$", s.strip(), re.I)
        if m:
            items = [parse_expr(x) for x in _split_top(m.group(2)) if not re.match(r"^\s*stat\s*=", x, re.I)]
            return (m.group(1).lower(), no, items)
        if low.startswith("nullify") or low.startswith("write") or low.startswith("print") or low.startswith("stop"):
            return ("continue", no)
        # assignment or pointer assignment: split at the top-level '=' / '=>'
        depth, q = 0, None
        for i, ch in enumerate(s):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and s[i + 1:i + 2] != "=" and s[i - 1:i] not in ("=", "/", "<", ">"):
                if s[i + 1:i + 2] == ">":
                    return ("ptr", no, parse_expr(s[:i]), parse_expr(s[i + 2:]))
                return ("assign", no, parse_expr(s[:i]), parse_expr(s[i + 1:]))
        raise FortranError("%s line %d: cannot interpret %r" % (sub.name, no, s))

    # ---- evaluation
    def ev(self, node, fr):
        k = node[0]
        if k == "num" or k == "str":
            return node[1]
        if k == "paren":
            return self.ev(node[1], fr)
        if k == "array":
            vals = [self.ev(x, fr) for x in node[1]]
            if any(isinstance(v, str) for v in vals):
                return FArray(np.array(vals, dtype=object))
            return FArray(np.array(vals, dtype=np.float64 if any(isinstance(v, float) for v in vals) else np.int64))
        if k == "name":
            return fr.get(node[1])
        if k == "un":
            v = self.ev(node[2], fr)
            if isinstance(v, FArray):
                if node[1] == "-":
                    return FArray(-v.a)
                if node[1] == "+":
                    return v
                return FArray(~v.a)
            if node[1] == "-":
                return -v
            if node[1] == "+":
                return v
            return not v
        if k == "bin":
            op = node[1]
            if op == ".and.":
                return bool(self.ev(node[2], fr)) and bool(self.ev(node[3], fr))
            if op == ".or.":
                return bool(self.ev(node[2], fr)) or bool(self.ev(node[3], fr))
            a, b = self.ev(node[2], fr), self.ev(node[3], fr)
            if isinstance(a, FArray) or isinstance(b, FArray):
                return self._array_op(op, a, b)
            if op == "//":
                return str(a) + str(b)
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            if op == "/":
                return _idiv(a, b)
            if op == "**":
                if isinstance(b, int) and not isinstance(b, bool):
                    return powi(a, b)
                return math.pow(a, b)
            if op == "==":
                return a == b
            if op == "/=":
                return a != b
            if op == "<":
                return a < b
            if op == "<=":
                return a <= b
            if op == ">":
                return a > b
            if op == ">=":
                return a >= b
            if op == ".eqv.":
                return bool(a) == bool(b)
            if op == ".neqv.":
                return bool(a) != bool(b)
            raise FortranError("operator %s" % op)
        if k == "comp":
            base = self.ev(node[1], fr)
            if base is None:
                raise FortranError("component %s of a disassociated pointer" % node[2])
            return getattr(base, node[2])
        if k == "call":
            return self._call_expr(node, fr)
        raise FortranError("cannot evaluate %s" % (node,))

    @staticmethod
    def _array_op(op, a, b):
        """elementwise whole-array arithmetic (numpy float64 operations are the same IEEE operations)"""
        x = a.a if isinstance(a, FArray) else a
        y = b.a if isinstance(b, FArray) else b
        if op == "+":
            return FArray(x + y)
        if op == "-":
            return FArray(x - y)
        if op == "*":
            return FArray(x * y)
        if op == "/":
            return FArray(x / y)
        if op == "**" and isinstance(y, int) and not isinstance(y, bool):
            return FArray(powi(x, y))                        # elementwise multiplications
        rel = {"==": np.equal, "/=": np.not_equal, "<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal}
        if op in rel:
            return FArray(rel[op](x, y))
        raise FortranError("array operator %s" % op)

    def _index(self, args, fr):
        idx = []
        for kw, a in args:
            if a[0] == "slice":
                lo = None if a[1] is None else self.ev(a[1], fr)
                hi = None if a[2] is None else self.ev(a[2], fr)
                idx.append(slice(lo, hi))
            else:
                idx.append(self.ev(a, fr))
        return tuple(idx)

    def _call_expr(self, node, fr):
        base, args = node[1], node[2]
        if base[0] == "name":
            name = base[1]
            if fr.has(name):
                v = fr.get(name)
                if isinstance(v, FArray):
                    return v.get(self._index(args, fr))
            if name == "present":
                a = args[0][1][1]
                return a in fr.vars and fr.vars[a] is not ABSENT
            if name == "associated":
                vals = [self.ev(a, fr) for _, a in args]
                return vals[0] is not None if len(vals) == 1 else vals[0] is vals[1]
            if name in self.subs and self.subs[name].result:
                callee = self.invoke(name, args, fr)
                return callee.get(self.subs[name].result)
            if name in self.func_hooks:
                return self.func_hooks[name](self, fr, args)
            if name in INTRINSICS:
                vals = [self.ev(a, fr) for kw, a in args if kw is None]
                kws = {kw: self.ev(a, fr) for kw, a in args if kw is not None}
                if name in ARRAY_INTRINSICS and any(isinstance(v, FArray) for v in vals):
                    return FArray(ARRAY_INTRINSICS[name](*[v.a if isinstance(v, FArray) else v for v in vals]))
                return INTRINSICS[name](*vals, **kws)
            raise FortranError("%s: %r is neither an array nor a known function" % (fr.sub.name, name))
        v = self.ev(base, fr)
        if isinstance(v, FArray):
            return v.get(self._index(args, fr))
        raise FortranError("cannot index %s" % (base,))

    def _assign(self, lhs, val, fr):
        if lhs[0] == "name":
            fr.set(lhs[1], val)
        elif lhs[0] == "call":
            arr = self.ev(lhs[1], fr) if lhs[1][0] != "name" else fr.get(lhs[1][1])
            if isinstance(arr, str) and lhs[1][0] == "name":          # msg(i:j) = "..." : a substring of a character variable
                idx = self._index(lhs[2], fr)[0]
                lo, hi = (idx.start or 1, idx.stop or len(arr)) if isinstance(idx, slice) else (idx, idx)
                chars = list(arr.ljust(hi))
                chars[lo - 1:hi] = list(str(val).ljust(hi - lo + 1))[:hi - lo + 1]
                fr.set(lhs[1][1], "".join(chars))
                return
            if not isinstance(arr, FArray):
                raise FortranError("assignment to an element of %s, which is not an array" % (lhs[1],))
            arr.set(self._index(lhs[2], fr), val)
        elif lhs[0] == "comp":
            setattr(self.ev(lhs[1], fr), lhs[2], val)
        else:
            raise FortranError("cannot assign to %s" % (lhs,))

    # ---- execution
    def run_block(self, nodes, fr):
        for n in nodes:
            k = n[0]
            if k == "assign":
                self._assign(n[2], self.ev(n[3], fr), fr)
            elif k == "call":
                self._call_stmt(n, fr)
            elif k == "if":
                for cond, body in n[2]:
                    if self.ev(cond, fr):
                        self.run_block(body, fr)
                        break
            elif k == "do":
                lim = [self.ev(x, fr) for x in n[3]]
                if any(isinstance(x, float) for x in lim):
                    raise FortranError("%s line %d: real do-loop bounds %s" % (fr.sub.name, n[1], lim))
                lo, hi, st = lim[0], lim[1], (lim[2] if len(lim) > 2 else 1)
                count = max((hi - lo + st) // st, 0)
                v = lo
                try:
                    for _ in range(count):
                        fr.set(n[2], v)
                        try:
                            self.run_block(n[4], fr)
                        except _Cycle:
                            pass
                        v += st
                    fr.set(n[2], v)
                except _Exit:
                    pass
            elif k == "dowhile":
                try:
                    while self.ev(n[2], fr):
                        try:
                            self.run_block(n[3], fr)
                        except _Cycle:
                            pass
                except _Exit:
                    pass
            elif k == "select":
                sel = self.ev(n[2], fr)
                for vals, body in n[3]:
                    def hit(x):
                        if x[0] == "range":
                            return (x[1] is None or self.ev(x[1], fr) <= sel) and (x[2] is None or sel <= self.ev(x[2], fr))
                        return self.ev(x, fr) == sel
                    if vals is None or any(hit(x) for x in vals):
                        self.run_block(body, fr)
                        break
            elif k == "ptr":
                val = self.ev(n[3], fr)
                if n[2][0] == "name":
                    fr.bind(n[2][1], val)
                else:
                    self._assign(n[2], val, fr)
            elif k == "exit":
                raise _Exit()
            elif k == "cycle":
                raise _Cycle()
            elif k == "return":
                raise _Return()
            elif k == "allocate":
                for item in n[2]:
                    if item[0] in ("name", "comp"):        # allocate(p): a scalar of derived type
                        import types as _types
                        obj = _types.SimpleNamespace()
                        if item[0] == "name":
                            fr.bind(item[1], obj)
                        else:
                            setattr(self.ev(item[1], fr), item[2], obj)
                        continue
                    shape = [int(self.ev(a, fr)) for _, a in item[2]]
                    target = item[1]
                    kind = fr.sub.types.get(target[1], "real") if target[0] == "name" else self.component_types.get(target[2], "real")
                    dt = np.float64 if kind == "real" else (np.int64 if kind == "int" else np.bool_)
                    new = FArray(np.zeros(tuple(reversed(shape)), dtype=dt))
                    if target[0] == "comp":                # allocate(obj % field(n, m))
                        setattr(self.ev(target[1], fr), target[2], new)
                        continue
                    cur = fr.vars.get(target[1])
                    if isinstance(cur, (VarRef, CompRef)):   # an allocatable dummy: the caller's variable is allocated
                        cur.bind(new)
                    elif target[1] not in fr.vars and target[1] not in fr.sub.types:
                        self.globals[fr.sub.renames.get(target[1], target[1])] = new     # a module-level allocatable
                    else:
                        fr.bind(target[1], new)
            elif k in ("deallocate", "continue"):
                pass
            else:
                raise FortranError("statement %s" % (n,))

    def _call_stmt(self, n, fr):
        _, no, name, args = n
        if name in self.hooks:
            return self.hooks[name](self, fr, args)
        if name in self.noop:                       # (also overrides a routine of that name found in a loaded file)
            return None
        if name in self.POOL_GETTERS:
            return self._pool_get(name, args, fr)
        if name in self.subs:
            self.invoke(name, args, fr)
            return None
        raise FortranError("%s line %d: call of %r, which is neither loaded, mapped nor declared a no-op" % (fr.sub.name, no, name))

    def _pool_get(self, name, args, fr):
        pool = self.ev(args[0][1], fr) if args[0][1][0] == "name" and fr.has(args[0][1][1]) else None
        key = self.ev(args[1][1], fr)
        tnode = args[2][1]

        def bind(val):
            if tnode[0] == "name":
                fr.bind(tnode[1], val)
            else:                                          # MPAS_pool_get_array(pool, 'x', obj % field)
                setattr(self.ev(tnode[1], fr), tnode[2], val)
        if name == "mpas_pool_get_subpool":
            bind(key)
            return
        level = self.ev(args[3][1], fr) if len(args) > 3 else None
        for k in ((pool, key, level), (pool, key), (None, key, level), (None, key), key):
            if k in self.pool:
                bind(self.pool[k])
                return
        raise FortranError("%s: the harness holds nothing for %s(%r, %r)" % (fr.sub.name, name, pool, key))

    def invoke(self, name, args, caller):
        sub = self.subs[name]
        self._prepare(sub)
        self.trace.append(name)
        fr = Frame(self, sub)
        pos = [a for kw, a in args if kw is None]
        kws = {kw: a for kw, a in args if kw is not None}
        for i, d in enumerate(sub.args):
            node = pos[i] if i < len(pos) else kws.get(d)
            if node is None:
                fr.vars[d] = ABSENT
                continue
            fr.vars[d] = self._actual(node, caller)
        for nm, dims, kind in sub.decl_arrays:
            shape = tuple(int(self.ev(x, fr)) for x in dims)
            dt = np.float64 if kind == "real" else (np.int64 if kind == "int" else np.bool_)
            fr.vars[nm] = FArray(np.zeros(tuple(reversed(shape)), dtype=dt))
        for nm, typename in sub.decl_objects:
            fr.vars[nm] = self.new_object(typename, fr)
        for nm, init in sub.inits:
            fr.vars[nm] = self.ev(init, fr)
        try:
            self.run_block(sub.body, fr)
        except _Return:
            pass
        return fr

    def _actual(self, node, caller):
        if caller is None:
            return node                                    # already a value (entry call from Python)
        if node[0] == "name":
            if caller.has(node[1]):
                v = caller.lookup(node[1])
                if isinstance(v, (int, float, bool)) and not isinstance(v, (ElemRef, VarRef, CompRef)):
                    return VarRef(caller, node[1])
                return v
            return VarRef(caller, node[1])                  # an output the callee defines
        if node[0] == "call" and node[1][0] == "name" and caller.has(node[1][1]) and isinstance(caller.get(node[1][1]), FArray):
            arr = caller.get(node[1][1])
            idx = self._index(node[2], caller)
            if any(isinstance(i, slice) for i in idx):
                return arr.get(idx)                         # a section: a view of the same memory
            return ElemRef(arr, idx)
        if node[0] == "call" and node[1][0] == "comp":      # obj % array(i, j): an element (or section) of a component
            arr = self.ev(node[1], caller)
            if isinstance(arr, FArray):
                idx = self._index(node[2], caller)
                if any(isinstance(i, slice) for i in idx):
                    return arr.get(idx)
                return ElemRef(arr, idx)
        if node[0] == "comp":                               # obj % field: arrays and objects by descriptor, scalars by reference
            obj = self.ev(node[1], caller)
            val = getattr(obj, node[2], None)
            if isinstance(val, (int, float, bool)) or val is None:
                return CompRef(obj, node[2])
            return val
        return self.ev(node, caller)

    def call(self, name, *values, **kw):
        """Entry from Python: positional values (Python scalars, FArray, objects) for the dummy arguments."""
        name = name.lower()
        sub = self.subs[name]
        self._prepare(sub)
        self.trace.append(name)
        fr = Frame(self, sub)
        for i, d in enumerate(sub.args):
            if i < len(values):
                fr.vars[d] = values[i]
            elif d in kw:
                fr.vars[d] = kw[d]
            else:
                fr.vars[d] = ABSENT
        for nm, dims, kind in sub.decl_arrays:
            shape = tuple(int(self.ev(x, fr)) for x in dims)
            dt = np.float64 if kind == "real" else (np.int64 if kind == "int" else np.bool_)
            fr.vars[nm] = FArray(np.zeros(tuple(reversed(shape)), dtype=dt))
        for nm, typename in sub.decl_objects:
            fr.vars[nm] = self.new_object(typename, fr)
        for nm, init in sub.inits:
            fr.vars[nm] = self.ev(init, fr)
        try:
            self.run_block(sub.body, fr)
        except _Return:
            pass
        return fr


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def _matching_paren(s, i):
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise FortranError("unbalanced parentheses in %r" % s)
