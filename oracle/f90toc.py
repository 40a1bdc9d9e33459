#!/usr/bin/env python
"""oracle/f90toc.py -- mechanical Fortran 90 -> C translation of the reference's fast-marching code: fm2d/fm2d_ttime.f90
(module traveltime: travel, fouds1, fouds2, addtree, downtree, updtree, bilinear), selected subroutines of
fm2d/fm2dray_cartesian.f90 (gridder, bsplrefine, srtimes, rpaths) and the body of modrays' source loop (MODRAYS_SOURCE below), with
the module variables of fm2d/fm2d_globalp.f90.

TEST INFRASTRUCTURE.  There is no Fortran compiler in the build image; the restatement oracle/fm2d_ref.c would
otherwise be pinned by physics only.  This script reads the reference's own source WHERE IT LIES (nothing is copied
into the repository) and translates it statement by statement with the expression machinery of oracle/f77toc.py (every
node typed as Fortran types it, default-real literals as float literals, integer powers as multiplications, arguments
by reference, 1-based column-major arrays).  oracle/build_ref.sh compiles the result together with the hand-written
driver oracle/ref_harness/fm2d_f90_harness.c into the git-ignored oracle/_ref/libfm2d_ttime_f2c.so;
tests/test_oracle_fm2d_vs_reference.py compares the restatement with it bit for bit, per routine and for whole calls.

What is added to f77toc's subset:
  * free form: `!` comments, `&` continuation lines, blank-insignificant matching after lower-casing (no character data
    outside WRITE statements, which are dropped -- they only print);
  * MODULE / USE / CONTAINS / IMPLICIT NONE; module variables become thread-local C globals (the reference marks them
    `!$omp threadprivate`); allocatable arrays (module or local) become a pointer plus extents and lower bounds (`x`, `x_d1`,
    `x_l1`, ...), set by ALLOCATE / DEALLOCATE / ALLOCATED and by array = array (the left-hand side is reallocated when the
    shapes differ: Fortran 2003, gfortran's default) or by the driver where it plays the part of modrays' own ALLOCATEs;
  * CYCLE, FLOOR, NINT, REAL(); a synthetic subroutine made of statement ranges of a larger one (run_ranges);
  * declarations with attributes (`REAL(KIND=i10), DIMENSION(2,2) :: vss`), kind parameters (i10 = c_double, i5 = single);
  * the derived type `backpointer` (a C struct), component references `btg(i)%px`, whole-structure assignment;
  * DO WHILE, EXIT, RETURN, subroutines without dummy arguments; STOP sets `f90_stopped` and returns (it only occurs in
    `travel` itself).

usage: f90toc.py /root/reference/fm2d/fm2d_globalp.f90 /root/reference/fm2d/fm2d_ttime.f90 \
                 /root/reference/fm2d/fm2dray_cartesian.f90:gridder,bsplrefine,srtimes,rpaths,@modrays_source out.c
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import f77toc as F                                                     # noqa: E402
from f77toc import INT, R4, R8, LOG, CT, Node, Unit                     # noqa: E402

STRUCT = 9
CT[STRUCT] = "backpointer"
F.TOK = re.compile(F.TOK.pattern.replace("[-+*/(),<>=]", "[-+*/(),<>=%]"), re.X)


def read_free_form(path):
    """[(blank-free lowercase statement text, source line number)]"""
    out, pending = [], None
    for ln, raw in enumerate(open(path, errors="replace").read().split("\n"), 1):
        line = raw.rstrip("\r").replace("\t", " ")
        s = line.strip()
        if not s or s.startswith("!"):
            continue
        if re.match(r"write\s*\(", s, re.I):            # prints only (its character data may hold `!`)
            continue
        bang = line.find("!")
        if bang >= 0:
            line = line[:bang]
        t = re.sub(r"\s+", "", line).lower()
        if not t:
            continue
        if t.startswith("&"):
            t = t[1:]
        if pending is not None:                          # the previous line ended in `&`
            t, ln = pending[0] + t, pending[1]
            pending = None
        if t.endswith("&"):
            pending = (t[:-1], ln)
            continue
        if re.fullmatch(r"\d+continue", t):             # a labelled CONTINUE nobody jumps to
            continue
        out.append((t, ln))
    return out


class Module:
    """module-scope names: scalars {name: type}, arrays {name: (type, rank)}, parameters {name: (type, C expr)}"""

    def __init__(self):
        self.scalars, self.arrays, self.params, self.kinds = {}, {}, {}, {}
        self.order = []

    def type_of(self, spec):
        if spec == "integer":
            return INT
        if spec == "logical":
            return LOG
        m = re.fullmatch(r"real\(kind=([a-z0-9_]+)\)", spec)
        if m:
            return self.kinds[m.group(1)]
        if spec == "type(backpointer)":
            return STRUCT
        raise SyntaxError(f"type {spec!r}")

    def declare(self, text):
        """a module-level (or local) declaration -> [(name, type, rank or None, dims or None, init or None)]"""
        m = re.fullmatch(r"(integer|logical|real\(kind=[a-z0-9_]+\)|type\(backpointer\))((?:,[a-z]+(?:\([^)]*\))?)*)(?:::)?(.*)", text)
        if not m:
            return None
        spec, attrs, ents = m.group(1), m.group(2), m.group(3)
        if re.match(r"[a-z0-9_]*=", ents) and "::" not in text:
            return None                                  # an assignment to a variable whose name starts like a type
        attr = [a for a in Unit.split_top(attrs[1:])] if attrs else []
        typ = self.type_of(spec)
        dim = next((a for a in attr if a.startswith("dimension(")), None)
        out = []
        for ent in Unit.split_top(ents):
            m2 = re.fullmatch(r"([a-z][a-z0-9_]*)(?:\((.*)\))?(?:=(.*))?", ent)
            name, dims, init = m2.group(1), m2.group(2), m2.group(3)
            if dims is None and dim:
                dims = dim[len("dimension("):-1]
            out.append((name, typ, Unit.split_top(dims) if dims else None, init, "parameter" in attr))
        return out

    def read(self, stmts):
        for text, ln in stmts:
            if re.fullmatch(r"(module[a-z0-9_]+|use[a-z0-9_]+|implicitnone|endmodule[a-z0-9_]*|contains)", text):
                continue
            d = self.declare(text)
            if d is None:
                raise SyntaxError(f"module line {ln}: {text!r}")
            for name, typ, dims, init, is_par in d:
                if is_par and typ == INT and init in ("c_double", "selected_real_kind(5,10)"):
                    self.kinds[name] = R8 if init == "c_double" else R4
                elif is_par:
                    u = Unit90(None, self, "subroutine", "_", [])
                    self.params[name] = (typ, u.cast(u.parse(init), typ))
                elif dims:
                    assert all(d_ == ":" for d_ in dims), (name, dims)
                    self.arrays[name] = (typ, len(dims))
                else:
                    self.scalars[name] = typ
                self.order.append(name)

    def c_decls(self):
        o = ["typedef struct { int px, pz; } backpointer;", "static __thread int f90_stopped;"]
        for name in self.order:
            if name in self.kinds:
                continue
            if name in self.params:
                t, c = self.params[name]
                o.append(f"static const {CT[t]} {name} = {c};")
            elif name in self.arrays:
                t, rank = self.arrays[name]
                o.append(f"static __thread {CT[t]}* {name}; static __thread int " + ", ".join(f"{name}_d{k + 1}, {name}_l{k + 1} = 1" for k in range(rank)) + ";")
            else:
                o.append(f"static __thread {CT[self.scalars[name]]} {name};")
        return o


class Unit90(Unit):
    def __init__(self, tr, mod, kind, name, args):
        super().__init__(tr, kind, name, args)
        self.mod = mod
        self.alloc = {}      # local allocatable arrays: name -> rank

    def is_mod(self, name):
        if name in self.types or name in self.args:      # a local declaration hides the module's
            return False
        m = self.mod
        return name in m.scalars or name in m.arrays or name in m.params or re.fullmatch(r"[a-z0-9_]+_[dl][123]", name) is not None

    def vtype(self, name):
        if name in self.types:
            return self.types[name]
        m = self.mod
        if name in m.scalars:
            return m.scalars[name]
        if name in m.arrays:
            return m.arrays[name][0]
        if name in m.params:
            return m.params[name][0]
        if re.fullmatch(r"[a-z0-9_]+_[dl][123]", name):
            return INT
        raise SyntaxError(f"{self.name}: {name!r} is not declared (IMPLICIT NONE)")

    def note(self, name):
        if not self.is_mod(name):
            super().note(name)

    def ref(self, name):
        if self.is_mod(name):
            return name
        return super().ref(name)

    def addr(self, name):
        if self.is_mod(name):
            return name if name in self.mod.arrays else "&" + name
        return super().addr(name)

    def rank_alloc(self, name):
        """rank of an allocatable array (local or of a module), else None"""
        if name in self.alloc:
            return self.alloc[name]
        if name in self.mod.arrays and name not in self.types:
            return self.mod.arrays[name][1]
        return None

    def dims_of(self, name):
        if name in self.dims:
            return self.dims[name]
        r = self.rank_alloc(name)
        if r is not None:
            return [f"{name}_d{k + 1}" for k in range(r)]
        return None

    def index(self, name, idx):
        dims = self.dims_of(name)
        lb = [f"{name}_l{k + 1}" for k in range(len(idx))] if self.rank_alloc(name) is not None else ["1"] * len(idx)
        off = f"(({idx[0]})-{lb[0]})"
        stride = None
        for k in range(1, len(idx)):
            d = self.expr_c(dims[k - 1], INT)
            stride = d if stride is None else f"({stride})*({d})"
            off += f"+({stride})*(({idx[k]})-{lb[k]})"
        return off

    def call_or_index(self, name, args):
        if self.dims_of(name) is not None:
            idx = [self.cast(a, INT) for a in args]
            return Node("elem", self.vtype(name), f"{name}[{self.index(name, idx)}]", name=name, idx=idx)
        if name == "allocated":
            return Node("call", LOG, f"({args[0].name} != 0)")
        if name == "floor":
            return Node("call", INT, f"((int)floor({self.cast(args[0], R8)}))")
        if name == "nint":
            return Node("call", INT, f"((int)lround({self.cast(args[0], R8)}))")
        return super().call_or_index(name, args)

    def p_prim(self):
        n = super().p_prim()
        while self.peek()[1] == "%":                     # component of a derived-type value
            self.take()
            k, field = self.take()
            assert n.typ == STRUCT and field in ("px", "pz"), (n.c, field)
            n = Node("elem", INT, f"{n.c}.{field}", name=None, idx=None)
        return n

    def assign(self, lhs, rhs):
        ln = self.parse(lhs)
        if ln.kind == "var" and self.dims_of(ln.name) is not None:   # whole-array assignment:  nsts = -1,  ttnr = ttn
            rn = self.parse(rhs)
            if rn.kind == "var" and self.dims_of(rn.name) is not None:
                a, b = ln.name, rn.name
                ra, rb = self.rank_alloc(a), self.rank_alloc(b)
                if ra is None and rb is None:           # two explicit-shape arrays of one shape:  vio = vi
                    assert self.dims_of(a) == self.dims_of(b), (lhs, rhs)
                    n = " * ".join(f"({self.expr_c(d, INT)})" for d in self.dims_of(a))
                    self.emit(f"for (int i_ = 0; i_ < {n}; ++i_) {a}[i_] = {b}[i_];")
                    return
                assert ra is not None and ra == rb, (lhs, rhs)
                # Fortran 2003 (gfortran's default, -frealloc-lhs): an allocatable left-hand side of another shape is reallocated
                differ = " || ".join(f"{a}_d{k + 1} != {b}_d{k + 1}" for k in range(ra))
                take = " ".join(f"{a}_d{k + 1} = {b}_d{k + 1}; {a}_l{k + 1} = {b}_l{k + 1};" for k in range(ra))
                n = " * ".join(f"{b}_d{k + 1}" for k in range(ra))
                ct = CT[self.vtype(a)]
                self.emit(f"if ({a} == 0 || {differ}) {{ free({a}); {take} {a} = ({ct}*)calloc((size_t)({n}), sizeof({ct})); }}")
                self.emit(f"for (int i_ = 0; i_ < {n}; ++i_) {a}[i_] = {b}[i_];")
                return
            n = " * ".join(f"({self.expr_c(d, INT)})" for d in self.dims_of(ln.name))
            t = self.vtype(ln.name)
            self.emit(f"{{ {CT[t]} v_ = {self.cast(rn, t)}; for (int i_ = 0; i_ < {n}; ++i_) {ln.name}[i_] = v_; }}")
            return
        rn = self.parse(rhs)
        assert (ln.typ == STRUCT) == (rn.typ == STRUCT), (lhs, rhs)
        self.emit(f"{ln.c} = {self.cast(rn, ln.typ)};")

    def statement(self, text, ln):
        t = text
        if t.startswith("rays(") or t.startswith("allocate(rays("):
            # rpaths hands every traced ray to the caller's container, an array of the derived type T_RAY of module m_fm2d
            # (allocatable component, type-bound procedure: outside the subset).  The translation hands the same data -- slot,
            # number of points, the (2, nrp) point array, source and receiver -- to the driver's f90_ray_store instead.
            if t.endswith("%points=praypts"):
                self.tr.uses_ray_store = True
                self.emit(f"f90_ray_store({self.expr_c('pnpts(1,2)', INT)}, {self.expr_c('nrp', INT)}, praypts, {self.expr_c('csid', INT)}, {self.expr_c('i', INT)});")
            return
        m = re.fullmatch(r"dowhile\((.*)\)", t)
        if m:
            self.emit(f"while ({self.parse(m.group(1)).c}) {{")
            self.do_stack.append(("while", None, None))
            return
        if t == "enddo" and self.do_stack and self.do_stack[-1][0] == "while":
            self.do_stack.pop()
            self.emit("}")
            return
        if t == "stop":
            self.emit("f90_stopped = 1; return;")
            return
        if t == "cycle":
            self.emit("continue;")
            return
        m = re.fullmatch(r"allocate\((.*)\)", t)
        if m:
            for item in self.split_top(m.group(1)):
                if item.startswith("stat="):
                    self.emit(f"{self.ref(item[5:])} = 0;")
                    continue
                m2 = re.fullmatch(r"([a-z][a-z0-9_]*)\((.*)\)", item)
                name, specs = m2.group(1), self.split_top(m2.group(2))
                assert self.rank_alloc(name) == len(specs), item
                n = []
                for k, sp in enumerate(specs):
                    lo, hi = sp.split(":") if ":" in sp else ("1", sp)
                    self.emit(f"{name}_l{k + 1} = {self.expr_c(lo, INT)}; {name}_d{k + 1} = ({self.expr_c(hi, INT)}) - {name}_l{k + 1} + 1;")
                    n.append(f"{name}_d{k + 1}")
                ct = CT[self.vtype(name)]
                # (+ 64 elements: modrays sizes btg by snb and never checks it; on tiny grids the narrow band outgrows it and the
                #  Fortran overruns the array -- the pad keeps that undefined behaviour from corrupting the test process)
                self.emit(f"{name} = ({ct}*)calloc((size_t)({' * '.join(n)}) + 64, sizeof({ct}));")
            return
        m = re.fullmatch(r"deallocate\((.*)\)", t)
        if m:
            for item in self.split_top(m.group(1)):
                if item.startswith("stat="):
                    self.emit(f"{self.ref(item[5:])} = 0;")
                else:
                    self.emit(f"free({item}); {item} = 0;")
            return
        m = re.fullmatch(r"call([a-z][a-z0-9_]*)", t)
        if m:
            self.tr.called.add(m.group(1))
            self.emit(f"{m.group(1)}_();")
            return
        super().statement(text, ln)


class Translator90:
    def __init__(self, mod):
        self.mod = mod
        self.units = []
        self.called = set()
        self.uses_ray_store = False

    def run(self, stmts, only=None):
        """only: names of the subroutines to translate (everything else in the file, module-level statements included, is skipped)"""
        u, in_spec, skipping = None, False, False
        for text, ln in stmts:
            if skipping:
                skipping = re.fullmatch(r"endsubroutine[a-z0-9_]*", text) is None
                continue
            if u is None:
                m = re.fullmatch(r"subroutine([a-z][a-z0-9_]*)(?:\((.*)\))?", text)
                if m and only is not None and m.group(1) not in only:
                    skipping = True
                    continue
                if not m and only is not None:
                    continue
                if m:
                    u = Unit90(self, self.mod, "subroutine", m.group(1), m.group(2).split(",") if m.group(2) else [])
                    self.units.append(u)
                    in_spec = True
                    continue
                if re.fullmatch(r"(module[a-z0-9_]+|use[a-z0-9_]+|implicitnone|contains|endmodule[a-z0-9_]*|typebackpointer|endtypebackpointer)", text):
                    continue
                d = self.mod.declare(text)               # module traveltime's own variables: ntr, btg; the type's components
                if d is None:
                    raise SyntaxError(f"line {ln}: {text!r} outside a subroutine")
                for name, typ, dims, init, is_par in d:
                    if name in ("px", "pz"):
                        continue
                    if dims:
                        self.mod.arrays[name] = (typ, len(dims))
                    else:
                        self.mod.scalars[name] = typ
                    self.mod.order.append(name)
                continue
            if re.fullmatch(r"endsubroutine[a-z0-9_]*", text):
                assert not u.do_stack, f"{ln}: unterminated do in {u.name}"
                u.emit("return;")
                u = None
                continue
            if in_spec:
                if text in ("implicitnone",) or re.fullmatch(r"use[a-z0-9_]+", text):
                    continue
                if text.startswith("type(t_ray)"):
                    continue                             # type(T_RAY) :: rays -- see Unit90.statement
                d = self.mod.declare(text)
                if d is not None:
                    for name, typ, dims, init, is_par in d:
                        u.types[name] = typ
                        if is_par:
                            u.params[name] = None
                            u.params[name] = u.cast(u.parse(init), typ)
                        elif dims and all(d_ == ":" for d_ in dims):
                            u.alloc[name] = len(dims)
                        elif dims:
                            u.dims[name] = dims
                        elif name not in u.args:
                            u.locals[name] = typ
                    continue
                in_spec = False
            u.statement(text, ln)
        return self

    def run_ranges(self, stmts, host, name, args, extra_decl, ranges):
        """A synthetic subroutine `name(args)` made of statement ranges of the subroutine `host`: its declarations are the host's
        (those of types outside the subset are skipped) plus extra_decl; every range is (first statement, last statement) as
        blank-free text, searched in order after the end of the previous one."""
        texts = [t for t, _ in stmts]
        h0 = next(i for i, t in enumerate(texts) if re.fullmatch(rf"subroutine{host}\(.*\)", t))
        h1 = next(i for i in range(h0, len(texts)) if texts[i] == f"endsubroutine{host}")
        u = Unit90(self, self.mod, "subroutine", name, args)
        self.units.append(u)
        decls = []
        for i in range(h0 + 1, h1):                     # the host's specification part
            t = texts[i]
            if t == "implicitnone" or re.fullmatch(r"use[a-z0-9_]+", t):
                continue
            if not re.match(r"(integer|real|type|logical|complex|character)", t) or ("::" not in t and "=" in t):
                break
            decls.append(t)
        for t in decls + list(extra_decl):
            try:
                d = self.mod.declare(t)
            except SyntaxError:
                d = None
            if d is None:
                continue                                 # logical, type(T_RAY): not used by the ranges (IMPLICIT NONE would tell)
            for nm, typ, dims, init, is_par in d:
                u.types[nm] = typ
                u.dims.pop(nm, None); u.alloc.pop(nm, None); u.locals.pop(nm, None)
                if dims and all(d_ == ":" for d_ in dims):
                    u.alloc[nm] = len(dims)
                elif dims:
                    u.dims[nm] = dims
                elif nm not in u.args:
                    u.locals[nm] = typ
        pos = h0
        for first, last in ranges:
            a = next(i for i in range(pos, h1) if texts[i] == first)
            if isinstance(last, tuple):                 # ("before", anchor, n): up to n + 1 statements before the anchor
                b = next(i for i in range(a, h1) if texts[i] == last[1]) - 1 - last[2]
            else:
                b = next(i for i in range(a, h1) if texts[i] == last)
            for i in range(a, b + 1):
                u.statement(texts[i], stmts[i][1])
            pos = b + 1
        assert not u.do_stack
        u.emit("return;")
        # locals that the ranges never mention need no storage
        used = " ".join(u.body)
        for nm in list(u.locals):
            if not re.search(rf"\b{nm}\b", used):
                del u.locals[nm]
        for nm in list(u.alloc):
            if not re.search(rf"\b{nm}\b", used):
                del u.alloc[nm]
        for nm in list(u.dims):
            if nm not in u.args and not re.search(rf"\b{nm}\b", used):
                del u.dims[nm]
        return self

    def c_source(self, paths):
        o = [f"/* GENERATED by oracle/f90toc.py from {' and '.join(paths)} -- do not edit, do not commit (oracle/_ref/ is git-ignored). */",
             "#include <math.h>", "#include <stdlib.h>",
             "static inline double f_sq(double x) { return x * x; }", "static inline float f_sqf(float x) { return x * x; }",
             "static inline int f_sqi(int x) { return x * x; }",
             "static inline double f_powi(double x, int n) { double r = 1.0; int m = n < 0 ? -n : n; while (m--) r *= x; return n < 0 ? 1.0 / r : r; }",
             "static inline double f_min(double a, double b) { return a < b ? a : b; }", "static inline double f_max(double a, double b) { return a > b ? a : b; }",
             "static inline int f_mini(int a, int b) { return a < b ? a : b; }", "static inline int f_maxi(int a, int b) { return a > b ? a : b; }", ""]
        o += self.mod.c_decls() + [""]
        if self.uses_ray_store:
            o.append("static void f90_ray_store(int slot, int nrp, const double* praypts, int csid, int revid); /* the driver's */")
        for u in self.units:
            o.append(f"static void {u.name}_({', '.join('void* ' + a + '_a' for a in u.args) or 'void'});")
        o.append("")
        for u in self.units:
            o.append(f"static void {u.name}_({', '.join('void* ' + a + '_a' for a in u.args) or 'void'}) {{")
            for a in u.args:
                if a in u.types:                         # (the ray container has no C type: it stays an untyped address)
                    o.append(f"  {CT[u.vtype(a)]}* {a} = ({CT[u.vtype(a)]}*){a}_a;")
            for k, v in u.params.items():
                o.append(f"  const {CT[u.types[k]]} {k.upper()} = {v};")
            for name, dims in u.dims.items():
                if name in u.args:
                    continue
                n = " * ".join(f"({u.expr_c(d, INT)})" for d in dims)
                o.append(f"  {CT[u.vtype(name)]} {name}[{n}];")
            for name, rank in u.alloc.items():
                o.append(f"  {CT[u.vtype(name)]}* {name} = 0; int " + ", ".join(f"{name}_d{k + 1} = 0, {name}_l{k + 1} = 1" for k in range(rank)) + ";")
            for name, typ in sorted(u.locals.items()):
                if name in u.dims or name in u.args or name in u.alloc:
                    continue
                o.append(f"  {CT[typ]} {name} = {{0}};" if typ == STRUCT else f"  {CT[typ]} {name} = 0;")
            o.extend("  " + s for s in u.body)
            o.append("}")
            o.append("")
        return "\n".join(o)


# What modrays (fm2dray_cartesian.f90:67-478) does for ONE source between `DO i=1,nsrc` and `call srtimes`, as a subroutine of
# its own: the source cell and the refinement window (:213-243), then the branch taken when `dynamic` is false -- it always is,
# the line before the test sets it (:251-252) -- i.e. source-grid refinement, the refined march, the mapping back onto the
# propagation grid, the completion of the narrow band and the second march, or the plain march (:283-438).  The statements
# in between (the `dynamic` restart: optional arguments, array sections, WHERE) are dead code and outside the subset.
MODRAYS_SOURCE = dict(
    host="modrays", name="modrays_source", args=["i", "nsrc", "sgs", "scx", "scz", "x", "z"],
    extra_decl=["real(kind=i10)::scx(nsrc),scz(nsrc)"],
    ranges=[("x=scx(i)", "if(vnb.gt.nnz)vnb=nnz"),
            # from the `else` branch of `if(dynamic)` to the ENDIF that closes IF(asgr.EQ.1): one ENDIF (that of `if(dynamic)`)
            # and the anchor itself are left out
            ("urg=0", ("before", "if(present(timefield))then", 1))])


def main():
    globalp, out = sys.argv[1], sys.argv[-1]
    mod = Module()
    mod.read(read_free_form(globalp))
    tr = Translator90(mod)
    paths = [globalp]
    for spec in sys.argv[2:-1]:                          # file  or  file:sub1,sub2  (`@modrays_source`: see MODRAYS_SOURCE)
        path, _, only = spec.partition(":")
        names = set(only.split(",")) if only else None
        stmts = read_free_form(path)
        if names and "@modrays_source" in names:
            names.discard("@modrays_source")
            tr.run(stmts, only=names)
            tr.run_ranges(stmts, **MODRAYS_SOURCE)
        else:
            tr.run(stmts, only=names)
        paths.append(spec)
    open(out, "w").write(tr.c_source(paths))
    names = [u.name for u in tr.units]
    missing = sorted(tr.called - set(names))
    print(f"f90toc: {len(names)} program units ({', '.join(names)}); unresolved externals: {missing or 'none'}")


if __name__ == "__main__":
    main()
