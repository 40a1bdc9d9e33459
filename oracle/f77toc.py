#!/usr/bin/env python
"""oracle/f77toc.py -- mechanical FORTRAN 77 -> C translation of the reference's surfmodes/surfdisp96.f.

TEST INFRASTRUCTURE.  There is no Fortran compiler in the build image, and the reference ships neither golden
dispersion values nor tests; the hand-written restatement oracle/surfdisp96_ref.c would otherwise be pinned by
nothing but review.  This script reads the reference's own source WHERE IT LIES (nothing is copied into the
repository), translates it statement by statement into C with Fortran's semantics kept per construct, and
oracle/build_ref.sh compiles the result into the git-ignored oracle/_ref/libsurfdisp96_f2c.so.  The translation is
generic over the subset the file uses -- it knows nothing about dispersion curves -- so agreement between its
output and the restatement, bit for bit over thousands of random models (tests/test_oracle_vs_reference.py), is
evidence about the restatement that does not share its author's reading of the algorithm.

Semantics kept:
  * fixed form: column 6 continuation, labels in columns 1-5, `c`/`C`/`*`/`!` comment lines, trailing `!` comments,
    tabs, blanks insignificant, no 72-column limit (the reference needs -ffixed-line-length-0, src/makefile:54);
  * implicit typing (i-n integer, else default real) and `implicit double precision (a-h,o-z)` per program unit;
    real*4 / real*8 / double precision / integer[*4] / real(kind=8) / dimension / parameter / save / data;
  * every expression node is typed as Fortran types it (integer < real*4 < real*8); real literals without a `d`
    exponent are emitted as float literals (`0.01*ss1` multiplies by 0.01f), with one as double literals; float
    arithmetic stays float (x86-64 SSE: FLT_EVAL_METHOD 0; compiled -ffp-contract=off); `x**2` is `x*x`;
  * arguments by reference (expressions through C99 compound literals), arrays 1-based and column-major with the
    declared (possibly run-time) leading dimension, whole-array assignment, DO loops with the trip count fixed on
    entry and the terminal label acting as `continue`, `exit`, logical and block IF, GOTO, `save`d variables
    (thread-local, which is what the reference's `!$omp threadprivate` asks for);
  * intrinsics: dble sngl abs dabs sqrt dsqrt dsin dcos dexp dlog dsign dmin1 dmax1 (libm's sin/cos/exp/log/pow).

usage: f77toc.py /root/reference/surfmodes/surfdisp96.f out.c
"""
import re
import sys

INT, R4, R8, LOG = 1, 2, 3, 0
CT = {INT: "int", R4: "float", R8: "double", LOG: "int"}


# ---------------------------------------------------------------- source -> statements
def read_statements(path):
    """[(label or None, blank-free lowercase text, first source line number)]"""
    stmts = []
    for ln, raw in enumerate(open(path, errors="replace").read().split("\n"), 1):
        line = raw.rstrip("\r")
        if not line.strip():
            continue
        if line[0] in "cC*!":
            continue
        if line[0] == "\t":                      # tab form: tab + digit = continuation, else statement field
            rest = line[1:]
            line = ("     " + rest[0] + rest[1:]) if rest[:1].isdigit() and rest[:1] != "0" else "      " + rest
        line = line.replace("\t", " ")
        if line.lstrip().startswith("!"):        # whole-line ! comment (incl. !$omp directives)
            continue
        bang = line.find("!")                    # no character constants survive in this file's executable code
        if bang >= 0:
            line = line[:bang]
        if not line.strip():
            continue
        line = line.ljust(7)
        lab, cont, text = line[:5], line[5], line[6:]
        if cont not in " 0" and not lab.strip():
            if not stmts:
                raise SyntaxError(f"{ln}: continuation without a statement")
            stmts[-1][1] += text
            continue
        stmts.append([int(lab) if lab.strip() else None, text, ln])
    out = []
    for lab, text, ln in stmts:
        t = re.sub(r"\s+", "", text).lower()
        if t:
            out.append((lab, t, ln))
    return out


# ---------------------------------------------------------------- expressions
TOK = re.compile(r"""
    (?P<num>(\d+\.?\d*|\.\d+)([de][+-]?\d+)?)
  | (?P<dot>\.(lt|le|gt|ge|eq|ne|and|or|not|true|false)\.)
  | (?P<id>[a-z][a-z0-9_]*)
  | (?P<op>\*\*|==|/=|<=|>=|[-+*/(),<>=])
""", re.X)
DOTOPS = ("lt", "le", "gt", "ge", "eq", "ne", "and", "or", "not")


def tokenize(s):
    toks, i = [], 0
    while i < len(s):
        # a digit string followed by ".op." must not swallow the dot:  16.lt.x
        m = re.match(r"\d+(?=\.(%s)\.)" % "|".join(DOTOPS), s[i:])
        if m:
            toks.append(("num", m.group(0)))
            i += m.end()
            continue
        m = TOK.match(s, i)
        if not m:
            raise SyntaxError(f"cannot tokenize {s[i:]!r} in {s!r}")
        kind = m.lastgroup
        toks.append((kind, m.group(kind)))
        i = m.end()
    return toks


class Node:
    def __init__(self, kind, typ, c, **kw):
        self.kind, self.typ, self.c = kind, typ, c
        self.__dict__.update(kw)


class Unit:
    """One program unit (subroutine / function) being translated."""

    def __init__(self, tr, kind, name, args):
        self.tr, self.kind, self.name, self.args = tr, kind, name, args
        self.implicit_double = False
        self.types = {}      # name -> INT/R4/R8
        self.dims = {}       # name -> [dim expr strings]
        self.params = {}     # parameter constants: name -> C expression
        self.saved = set()
        self.data = {}       # name -> literal text
        self.locals = {}     # discovered local scalars: name -> type
        self.body = []
        self.tmp = 0
        self.do_stack = []

    def implicit(self, name):
        if name[0] in "ijklmn":
            return INT
        return R8 if self.implicit_double else R4

    def vtype(self, name):
        if name in self.types:
            return self.types[name]
        return self.implicit(name)

    def note(self, name):
        if name not in self.args and name not in self.dims and name not in self.params and name != self.name:
            self.locals.setdefault(name, self.vtype(name))

    # ---- C spelling of a variable reference
    def ref(self, name):
        if name in self.params:
            return name.upper()
        if name == self.name and self.kind == "function":
            return name + "_result"
        if name in self.args and name not in self.dims:
            return f"(*{name})"
        self.note(name)
        return name

    def addr(self, name):
        if name in self.args or name in self.dims:
            return name
        if name == self.name and self.kind == "function":
            return "&" + name + "_result"
        self.note(name)
        return "&" + name

    def index(self, name, idx):
        """idx: list of C int expressions -> flat 0-based offset, column-major, 1-based subscripts."""
        dims = self.dims[name]
        off = f"(({idx[0]})-1)"
        stride = None
        for k in range(1, len(idx)):
            d = self.expr_c(dims[k - 1], INT)
            stride = d if stride is None else f"({stride})*({d})"
            off += f"+({stride})*(({idx[k]})-1)"
        return off

    # ---- recursive descent with Fortran precedences
    def parse(self, text):
        saved = (getattr(self, "toks", None), getattr(self, "pos", 0))   # re-entrant: array bounds are parsed mid-expression
        self.toks = tokenize(text)
        self.pos = 0
        try:
            n = self.p_or()
            if self.pos != len(self.toks):
                raise SyntaxError(f"trailing tokens in {text!r}: {self.toks[self.pos:]}")
        finally:
            self.toks, self.pos = saved
        return n

    def peek(self):
        return self.toks[self.pos] if self.pos < len(self.toks) else (None, None)

    def take(self, val=None):
        k, v = self.peek()
        if val is not None and v != val:
            raise SyntaxError(f"expected {val!r}, got {v!r}")
        self.pos += 1
        return k, v

    def p_or(self):
        n = self.p_and()
        while self.peek()[1] == ".or.":
            self.take()
            r = self.p_and()
            n = Node("bin", LOG, f"({n.c} || {r.c})")
        return n

    def p_and(self):
        n = self.p_not()
        while self.peek()[1] == ".and.":
            self.take()
            r = self.p_not()
            n = Node("bin", LOG, f"({n.c} && {r.c})")
        return n

    def p_not(self):
        if self.peek()[1] == ".not.":
            self.take()
            r = self.p_not()
            return Node("un", LOG, f"(!{r.c})")
        return self.p_rel()

    REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "!=", "<": "<", "<=": "<=", ">": ">",
           ">=": ">=", "==": "==", "/=": "!="}

    def p_rel(self):
        n = self.p_add()
        v = self.peek()[1]
        if v in self.REL:
            self.take()
            r = self.p_add()
            a, b = self.promote(n, r)
            n = Node("rel", LOG, f"({a} {self.REL[v]} {b})")
        return n

    def promote(self, a, b):
        """C spellings of both operands converted to the common Fortran type (explicit casts: nothing is left to C)."""
        t = max(a.typ, b.typ)
        return self.cast(a, t), self.cast(b, t)

    def cast(self, n, t):
        if n.typ == t or n.typ == LOG:
            return n.c
        return f"(({CT[t]})({n.c}))"

    def p_add(self):
        k, v = self.peek()
        if v in ("+", "-"):            # unary sign binds the first TERM:  -a*b = -(a*b),  -a**2 = -(a**2)
            self.take()
            n = self.p_mul()
            n = Node("un", n.typ, f"(-{n.c})" if v == "-" else n.c)
        else:
            n = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            r = self.p_mul()
            a, b = self.promote(n, r)
            n = Node("bin", max(n.typ, r.typ), f"({a} {op} {b})")
        return n

    def p_mul(self):
        n = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            r = self.p_pow()
            a, b = self.promote(n, r)
            n = Node("bin", max(n.typ, r.typ), f"({a} {op} {b})")
        return n

    def p_pow(self):
        base = self.p_prim()
        if self.peek()[1] == "**":
            self.take()
            neg = False
            if self.peek()[1] == "-":
                self.take()
                neg = True
            ex = self.p_pow()              # right associative
            if neg:
                ex = Node("un", ex.typ, f"(-{ex.c})")
            return self.power(base, ex)
        return base

    def power(self, b, e):
        if e.typ == INT:
            if e.kind == "lit" and e.c == "2":      # what every Fortran compiler emits: one multiply
                f = {R8: "f_sq", R4: "f_sqf", INT: "f_sqi"}[b.typ]
                return Node("call", b.typ, f"{f}({b.c})")
            f = {R8: "f_powi", R4: "f_powif", INT: "f_powii"}[b.typ]
            return Node("call", b.typ, f"{f}({b.c}, {e.c})")
        t = max(b.typ, e.typ)
        f = "pow" if t == R8 else "powf"
        return Node("call", t, f"{f}({self.cast(b, t)}, {self.cast(e, t)})")

    def p_prim(self):
        k, v = self.take()
        if k == "num":
            if re.fullmatch(r"\d+", v):
                return Node("lit", INT, v)
            m = re.fullmatch(r"(\d*\.?\d*)(?:([de])([+-]?\d+))?", v)
            mant, ed, ex = m.group(1), m.group(2), m.group(3)
            if "." not in mant:
                mant += "."
            if mant.startswith("."):
                mant = "0" + mant
            if mant.endswith("."):
                mant += "0"
            lit = mant + (f"e{ex}" if ex else "")
            if ed == "d":
                return Node("lit", R8, lit)
            return Node("lit", R4, lit + "f")
        if v == "(":
            n = self.p_or()
            self.take(")")
            return Node("par", n.typ, f"({n.c})")
        if k == "dot" and v in (".true.", ".false."):
            return Node("lit", LOG, "1" if v == ".true." else "0")
        if k == "id":
            if self.peek()[1] == "(":
                self.take()
                args = []
                if self.peek()[1] != ")":
                    while True:
                        args.append(self.p_or())
                        if self.peek()[1] == ",":
                            self.take()
                            continue
                        break
                self.take(")")
                return self.call_or_index(v, args)
            return Node("var", self.vtype(v) if v not in self.params else self.ptype(v), self.ref(v), name=v)
        raise SyntaxError(f"unexpected token {v!r}")

    def ptype(self, name):
        return self.types.get(name, INT if name[0] in "ijklmn" else R4)

    INTRIN = {"dble": ("cast", R8), "sngl": ("cast", R4), "real": ("cast", R4), "float": ("cast", R4), "dfloat": ("cast", R8),
              "int": ("cast", INT)}

    def call_or_index(self, name, args):
        if name in self.dims:                                   # array element
            idx = [self.cast(a, INT) for a in args]
            return Node("elem", self.vtype(name), f"{name}[{self.index(name, idx)}]", name=name, idx=idx)
        if name in self.INTRIN:
            t = self.INTRIN[name][1]
            return Node("call", t, f"(({CT[t]})({args[0].c}))")
        if name in ("abs", "dabs"):
            a = args[0]
            f = {R8: "fabs", R4: "fabsf", INT: "abs"}[a.typ]
            return Node("call", a.typ, f"{f}({a.c})")
        if name in ("sqrt", "dsqrt"):
            a = args[0]
            return Node("call", a.typ, f"{'sqrt' if a.typ == R8 else 'sqrtf'}({a.c})")
        if name in ("dsin", "dcos", "dexp", "dlog"):
            return Node("call", R8, f"{name[1:]}({self.cast(args[0], R8)})")
        if name in ("sin", "cos", "exp", "log"):
            a = args[0]
            return Node("call", a.typ, f"{name}{'' if a.typ == R8 else 'f'}({a.c})")
        if name in ("dsign", "sign"):
            a, b = args
            t = max(a.typ, b.typ)
            return Node("call", t, f"{'copysign' if t == R8 else 'copysignf'}({self.cast(a, t)}, {self.cast(b, t)})")
        if name in ("dmin1", "dmax1", "amin1", "amax1", "min", "max"):
            t = max(a.typ for a in args)
            f = ("f_min" if "min" in name else "f_max") + {R8: "", R4: "f", INT: "i"}[t]
            c = self.cast(args[0], t)
            for a in args[1:]:
                c = f"{f}({c}, {self.cast(a, t)})"
            return Node("call", t, c)
        # user function
        self.tr.called.add(name)
        return Node("call", self.vtype(name), f"{name}_({self.actuals(args)})")

    def actuals(self, args):
        out = []
        for a in args:
            if a.kind == "var" and a.name not in self.params:
                out.append(f"(void*){self.addr(a.name)}")
            elif a.kind == "elem":
                out.append(f"(void*)&{a.c}")
            else:                                                # expression or constant: a temporary, by reference
                out.append(f"(void*)&({CT[a.typ]}){{{a.c}}}")
        return ", ".join(out)

    def expr_c(self, text, want=None):
        n = self.parse(text)
        return self.cast(n, want) if want is not None else n.c

    # ---- statements
    def emit(self, s):
        self.body.append(s)

    def assign(self, lhs, rhs):
        ln = self.parse(lhs)
        rn = self.parse(rhs)
        if ln.kind == "var" and ln.name in self.dims:            # whole-array assignment:  cp = 100.0
            n = " * ".join(f"({self.expr_c(d, INT)})" for d in self.dims[ln.name])
            t = self.vtype(ln.name)
            self.emit(f"{{ {CT[t]} v_ = {self.cast(rn, t)}; for (int i_ = 0; i_ < {n}; ++i_) {ln.name}[i_] = v_; }}")
            return
        self.emit(f"{ln.c} = {self.cast(rn, ln.typ)};")

    def statement(self, text, ln):
        t = text
        m = re.fullmatch(r"if\((.*)\)then", t)
        if m:
            self.emit(f"if ({self.parse(m.group(1)).c}) {{")
            return
        m = re.fullmatch(r"else?if\((.*)\)then", t)
        if m:
            self.emit(f"}} else if ({self.parse(m.group(1)).c}) {{")
            return
        if t == "else":
            self.emit("} else {")
            return
        if t == "endif":
            self.emit("}")
            return
        if t == "enddo":
            lab, var, step = self.do_stack.pop()
            assert lab is None, f"{ln}: enddo closes a labelled do"
            self.emit(f"}} }}")
            return
        if t == "exit":
            self.emit("break;")
            return
        if t == "continue":
            self.emit(";")
            return
        if t == "return":
            self.emit(f"return {self.name}_result;" if self.kind == "function" else "return;")
            return
        m = re.fullmatch(r"goto(\d+)", t)
        if m:
            self.emit(f"goto L{m.group(1)};")
            return
        m = re.fullmatch(r"do(\d*)([a-z][a-z0-9_]*)=(.*)", t)
        if m and self.top_level_commas(m.group(3)):
            parts = self.split_top(m.group(3))
            lab = int(m.group(1)) if m.group(1) else None
            var = m.group(2)
            e1, e2 = self.expr_c(parts[0], INT), self.expr_c(parts[1], INT)
            e3 = self.expr_c(parts[2], INT) if len(parts) > 2 else "1"
            v = self.ref(var)
            self.tmp += 1
            n = f"trip{self.tmp}_"
            # Fortran: iteration count fixed on entry = max(0, (e2 - e1 + e3) / e3); the variable keeps counting
            self.emit(f"{{ int st_ = {e3}; int {n} = (({e2}) - ({e1}) + st_) / st_; for ({v} = {e1}; {n} > 0; --{n}, {v} += st_) {{")
            self.do_stack.append((lab, var, e3))
            return
        m = re.fullmatch(r"call([a-z][a-z0-9_]*)\((.*)\)", t)
        if m:
            self.tr.called.add(m.group(1))
            args = [self.parse(a) for a in self.split_top(m.group(2))]
            self.emit(f"{m.group(1)}_({self.actuals(args)});")
            return
        m = re.match(r"if\(", t)
        if m:                                                    # logical IF: find the matching parenthesis
            depth, i = 0, 2
            while True:
                depth += t[i] == "("
                depth -= t[i] == ")"
                if depth == 0:
                    break
                i += 1
            cond, rest = t[3:i], t[i + 1:]
            self.emit(f"if ({self.parse(cond).c}) {{")
            self.statement(rest, ln)
            self.emit("}")
            return
        eq = self.top_level_eq(t)
        if eq > 0:
            self.assign(t[:eq], t[eq + 1:])
            return
        raise SyntaxError(f"line {ln}: cannot translate {text!r}")

    @staticmethod
    def split_top(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        out.append(cur)
        return out

    @classmethod
    def top_level_commas(cls, s):
        return len(cls.split_top(s)) >= 2

    @staticmethod
    def top_level_eq(s):
        depth = 0
        for i, ch in enumerate(s):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and s[i + 1:i + 2] != "=" and s[i - 1] not in "<>/=":
                return i
        return -1


# ---------------------------------------------------------------- declarations
DECL = [(r"doubleprecision(?:::)?(.*)", R8), (r"real\*8(?:::)?(.*)", R8), (r"real\(kind=8\)(?:::)?(.*)", R8),
        (r"real\*4(?:::)?(.*)", R4), (r"integer\*4(?:::)?(.*)", INT), (r"integer(?:::)?(.*)", INT), (r"real(?:::)?(.*)", R4)]


class Translator:
    def __init__(self):
        self.units = []
        self.called = set()

    def declare(self, u, text):
        """True when `text` is a specification statement (consumed)."""
        if text.startswith("implicitdoubleprecision"):
            u.implicit_double = True
            return True
        m = re.fullmatch(r"parameter\((.*)\)", text)
        if m:
            for item in Unit.split_top(m.group(1)):
                k, v = item.split("=", 1)
                u.params[k] = None
                u.params[k] = u.expr_c(v)
            return True
        m = re.fullmatch(r"save(.*)", text)
        if m:
            u.saved.update(x for x in m.group(1).split(",") if x)
            return True
        m = re.fullmatch(r"data([a-z0-9_,]+)/(.*)/", text)
        if m:
            for k, v in zip(m.group(1).split(","), m.group(2).split(",")):
                u.data[k] = u.parse(v).c
            return True
        m = re.fullmatch(r"dimension(.*)", text)
        if m:
            self.entities(u, m.group(1), None)
            return True
        for pat, typ in DECL:
            m = re.fullmatch(pat, text)
            if m and not re.match(r"[a-z0-9_]*=", m.group(1)):  # (an assignment to a variable called real... is not a declaration)
                self.entities(u, m.group(1), typ)
                return True
        return False

    def entities(self, u, text, typ):
        for ent in Unit.split_top(text):
            m = re.fullmatch(r"([a-z][a-z0-9_]*)(?:\((.*)\))?", ent)
            name, dims = m.group(1), m.group(2)
            if typ is not None:
                u.types[name] = typ
            if dims is not None:
                u.dims[name] = Unit.split_top(dims)

    def run(self, stmts):
        u = None
        in_spec = False
        for lab, text, ln in stmts:
            m = re.fullmatch(r"(subroutine|function)([a-z][a-z0-9_]*)\((.*)\)", text)
            if m:
                u = Unit(self, m.group(1), m.group(2), m.group(3).split(","))
                self.units.append(u)
                in_spec = True
                continue
            if text == "end":
                assert not u.do_stack, f"{ln}: unterminated do in {u.name}"
                u.emit(f"return {u.name}_result;" if u.kind == "function" else "return;")
                u = None
                continue
            if in_spec and lab is None and self.declare(u, text):
                continue
            in_spec = False
            if lab is not None:
                # terminal statement of labelled DO loops: the label sits INSIDE the loop body (jumping to it = next trip)
                u.emit(f"L{lab}: ;")
            closes = 0
            while u.do_stack and u.do_stack[-1][0] is not None and u.do_stack[-1][0] == lab:
                closes += 1
                u.do_stack.pop()
            if closes and text != "continue":
                u.statement(text, ln)
            elif not closes:
                u.statement(text, ln)
            for _ in range(closes):
                u.emit("} }")
        return self

    # ---- C output
    def c_source(self, src_path):
        o = [f"/* GENERATED by oracle/f77toc.py from {src_path} -- do not edit, do not commit (oracle/_ref/ is git-ignored). */",
             "#include <math.h>", "#include <stdlib.h>",
             "static inline double f_sq(double x) { return x * x; }", "static inline float f_sqf(float x) { return x * x; }",
             "static inline int f_sqi(int x) { return x * x; }",
             "static inline double f_powi(double x, int n) { double r = 1.0; int m = n < 0 ? -n : n; while (m--) r *= x; return n < 0 ? 1.0 / r : r; }",
             "static inline float f_powif(float x, int n) { float r = 1.0f; int m = n < 0 ? -n : n; while (m--) r *= x; return n < 0 ? 1.0f / r : r; }",
             "static inline int f_powii(int x, int n) { int r = 1; while (n-- > 0) r *= x; return r; }",
             "static inline double f_min(double a, double b) { return a < b ? a : b; }", "static inline double f_max(double a, double b) { return a > b ? a : b; }",
             "static inline float f_minf(float a, float b) { return a < b ? a : b; }", "static inline float f_maxf(float a, float b) { return a > b ? a : b; }",
             "static inline int f_mini(int a, int b) { return a < b ? a : b; }", "static inline int f_maxi(int a, int b) { return a > b ? a : b; }", ""]
        for u in self.units:                                   # prototypes: every dummy is an untyped address
            ret = CT[u.vtype(u.name)] if u.kind == "function" else "void"
            o.append(f"{ret} {u.name}_({', '.join('void* ' + a + '_a' for a in u.args)});")
        o.append("")
        for u in self.units:
            ret = CT[u.vtype(u.name)] if u.kind == "function" else "void"
            o.append(f"{ret} {u.name}_({', '.join('void* ' + a + '_a' for a in u.args)}) {{")
            for k, v in u.params.items():
                o.append(f"  enum {{ {k.upper()} = {v} }};")
            for a in u.args:
                o.append(f"  {CT[u.vtype(a)]}* {a} = ({CT[u.vtype(a)]}*){a}_a;")
            for name, dims in u.dims.items():
                if name in u.args:
                    continue
                n = " * ".join(f"({u.expr_c(d, INT)})" for d in dims)
                o.append(f"  {CT[u.vtype(name)]} {name}[{n}];")
            if u.kind == "function":
                o.append(f"  {CT[u.vtype(u.name)]} {u.name}_result = 0;")
            for name, typ in sorted(u.locals.items()):
                if name in u.saved:
                    o.append(f"  static __thread {CT[typ]} {name} = 0;")
                elif name in u.data:
                    o.append(f"  {CT[typ]} {name} = {u.data[name]};")
                else:
                    o.append(f"  {CT[typ]} {name} = 0;")
            o.extend("  " + s for s in u.body)
            o.append("}")
            o.append("")
        return "\n".join(o)


def main():
    src, out = sys.argv[1], sys.argv[2]
    tr = Translator().run(read_statements(src))
    open(out, "w").write(tr.c_source(src))
    names = [u.name for u in tr.units]
    missing = sorted(tr.called - set(names))
    print(f"f77toc: {len(names)} program units ({', '.join(names)}); unresolved externals: {missing or 'none'}")


if __name__ == "__main__":
    main()
