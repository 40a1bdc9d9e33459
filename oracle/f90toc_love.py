#!/usr/bin/env python
"""oracle/f90toc_love.py -- mechanical Fortran 90 -> C translation of the reference's generalized R/T secular functions:
surfmodes/Love.f90 (init_love, delete_love, EinvE_L, propdn_L, propup_L, SecFuns_L) and the units of surfmodes/Rayleigh.f90 a
column reaches from `surfmodes` (inv2, init_rayleigh, delete_rayleigh, startl, SecFunSurf, EinvE, propup; with a water layer
SecFunSt, Stoneley, EinvE_f, propdn_f), bisecim, det3 and sort
of util.f90, C_Interval / N_cf (C_interval.f90), C_Interval_L / N_cf_L (C_interval_L.f90), setup_grt (surfmodes.f90), the internal procedures FundaMode and check of
SearchLove.f90, FundaMode and StMode of SearchRayleigh.f90, CR0_Finder with its internal Rayhomo and St_Finder with its internal getSt, with
`csq`, the
parameters and the derived type T_GRT of surfmodes/GRT.f90.

TEST INFRASTRUCTURE, in the line of oracle/f77toc.py and oracle/f90toc.py: the sources are read where they lie, nothing is
copied.  What this subset adds to f90toc's:
  * COMPLEX*16 as C `double _Complex`; the generated file MUST be compiled with -fcx-fortran-rules, which makes gcc's middle end
    expand products and quotients exactly as it does for gfortran (plain product, range-reduced quotient, no NaN recovery);
    exp / sqrt of a complex go to libm's cexp / csqrt as gfortran's do; AIMAG, DBLE, DCMPLX, complex literals `(re,im)`;
  * fixed-shape arrays with lower bounds at module scope (`cs(0:1)`), whole arrays and sections with constant bounds as VALUES:
    an array expression is scalarised at translation time into its element expressions (column-major) -- array constructors
    `[a,b]`, sections `a22(:,2)` / `a44(2:3,3:4)` / `Rdu(:,:,j)` (leading extents of a rank-3 allocatable fixed by its
    ALLOCATE), array (op) scalar or array, MATMUL (k ascending, as gfortran's inline and library versions sum), RESHAPE,
    elementwise EXP, array-valued function results and array actual arguments (copy-in), pointers associated with a section
    (`pp=>a44(2:3,3:4); pp=-pp`: an alias) -- and an assignment evaluates every right-hand-side element into a temporary before the
    first store (Fortran's semantics: `b22 = b22/(2.*b22(1,1))` divides by the OLD b22(1,1));
  * the derived types T_GRT and T_MODES_PARA as C structs (allocatable components = pointer + extent + lower bound), passed by
    reference; whole component arrays as actual arguments, under MAXVAL / MINVAL (also of a run-time section) and in
    `GRT%mu = GRT%mu/mu0`; parameters re-defined by a later module (m_surfmodes' eps, pi) under their own C names;
  * internal procedures as units of their own, the host's variables they read at file scope (`host=kt:real,c:real,grt:t_grt`),
    a host's dummy arguments through pointers the host stores on entry (`v1:real*@cr0_finder`);
    assumed-shape dummies; a procedure name as actual argument (the callee's dummy procedure is bound by the driver);
  * FUNCTION units (scalar or array result), dummy procedures (bound to a routine of the driver), SELECT CASE on an integer, DO
    with a negative step, DO without a control, DO WHILE, CYCLE, MERGE, 1-D sections with run-time bounds copied through a
    temporary (`vvv(2:index0)=vvv(i:ii)`), whole work arrays passed by reference, USE ... ONLY / PRIVATE / PUBLIC.

usage: f90toc_love.py GRT.f90 Love.f90 util.f90:bisecim out.c
       f90toc_love.py GRT.f90 Rayleigh.f90:inv2,init_rayleigh,delete_rayleigh,startl,secfunsurf,einve,propup util.f90:bisecim out.c
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import f77toc as F                                                     # noqa: E402
import f90toc as G                                                     # noqa: E402
from f77toc import INT, R4, R8, LOG, CT, Node, Unit                     # noqa: E402

CPX, TGRT, TPARA = 4, 10, 11
CT[CPX] = "double _Complex"
CT[TGRT] = "T_GRT"
CT[TPARA] = "T_MODES_PARA"
STRUCT_OF = {TGRT: "t_grt", TPARA: "t_modes_para"}
F.TOK = re.compile(F.TOK.pattern.replace("[-+*/(),<>=%]", r"[-+*/(),<>=%:\[\]]"), re.X)

SPEC = r"(integer|logical|real\(kind=[a-z0-9_]+\)|real\*8|complex\*16|type\(t_grt\)|type\(t_modes_para\))"


class Arr:
    """an array VALUE: shape + element Nodes in column-major order"""

    def __init__(self, shape, elems):
        self.shape, self.elems = tuple(shape), list(elems)
        self.typ = max(e.typ for e in elems)
        self.kind = "arr"


class Mod:
    def __init__(self):
        self.kinds, self.params, self.scalars, self.fixed, self.alloc, self.struct = {}, {}, {}, {}, {}, {}
        self.structs = {"t_grt": self.struct}
        self.param_cname, self.param_decls = {}, []   # a parameter re-defined by a later module (m_surfmodes' eps, pi) gets a new C name
        self.order = []
        self.ints = {}           # integer parameters by value (array bounds)
        self.lead = {}           # rank-3 allocatable arrays: bounds of the two leading dimensions, fixed by their ALLOCATE
        self.func_types = {}     # function name -> scalar result type, or ("arr", type, shape)

    def type_of(self, spec):
        if spec == "integer":
            return INT
        if spec == "logical":
            return LOG
        if spec == "real*8":
            return R8
        if spec == "complex*16":
            return CPX
        if spec == "type(t_grt)":
            return TGRT
        if spec == "type(t_modes_para)":
            return TPARA
        m = re.fullmatch(r"real\(kind=([a-z0-9_]+)\)", spec)
        return self.kinds[m.group(1)]

    def declare(self, text):
        m = re.fullmatch(SPEC + r"((?:,[a-z]+(?:\([^)]*\))?)*)(?:::)?(.*)", text)
        if not m:
            return None
        spec, attrs, ents = m.group(1), m.group(2), m.group(3)
        if re.match(r"[a-z0-9_]*=", ents) and "::" not in text:
            return None
        attr = Unit.split_top(attrs[1:]) if attrs else []
        typ = self.type_of(spec)
        dim = next((a for a in attr if a.startswith("dimension(")), None)
        out = []
        for ent in Unit.split_top(ents):
            m2 = re.fullmatch(r"([a-z][a-z0-9_]*)(?:\(([^=]*)\))?(?:=(.*))?", ent)
            name, dims, init = m2.group(1), m2.group(2), m2.group(3)
            if dims is None and dim:
                dims = dim[len("dimension("):-1]
            out.append((name, typ, Unit.split_top(dims) if dims else None, init, "parameter" in attr))
        return out

    def const_int(self, text):
        u = UnitG(None, self, "subroutine", "_", [])
        return int(eval(u.expr_c(text, INT).replace("/", "//"), dict(self.ints)))

    def add(self, name, typ, dims, init, is_par, into=None):
        tgt = self if into is None else None
        if is_par and typ == INT and init == "c_double":
            self.kinds[name] = R8
        elif is_par:
            u = UnitG(None, self, "subroutine", "_", [])
            c = u.cast(u.parse(init), typ)
            cname = name if name not in self.params else f"{name}_{len(self.param_decls)}"
            self.params[name] = (typ, c)
            self.param_cname[name] = cname
            self.param_decls.append((cname, typ, c))
            if typ == INT:
                self.ints[name] = self.const_int(init)
        elif dims and all(d == ":" for d in dims):
            (self.alloc if into is None else into)[name] = (typ, len(dims)) if into is None else ("alloc", typ, len(dims))
            if tgt:
                self.order.append(name)
        elif dims:
            b = []
            for d in dims:
                lo, hi = d.split(":") if ":" in d else ("1", d)
                b.append((self.const_int(lo), self.const_int(hi)))
            (self.fixed if into is None else into)[name] = (typ, b) if into is None else ("fixed", typ, b)
            if tgt:
                self.order.append(name)
        else:
            (self.scalars if into is None else into)[name] = typ if into is None else ("scalar", typ)
            if tgt:
                self.order.append(name)

    def read_header(self, stmts):
        """module-level statements up to CONTAINS; returns the index after it"""
        in_type, cur = False, None
        for k, (text, ln) in enumerate(stmts):
            if text == "contains" and not in_type:
                return k + 1
            if re.fullmatch(r"(module[a-z0-9_]+|use[a-z0-9_,:]+|implicitnone|private|public::.*)", text):
                continue
            m = re.fullmatch(r"type(t_grt|t_modes_para)", text)
            if m:
                in_type = True
                cur = self.structs.setdefault(m.group(1), {})
                continue
            if text in ("endtypet_grt", "endtype", "endtypet_modes_para"):
                in_type = False
                continue
            d = self.declare(text)
            if d is None:
                raise SyntaxError(f"module line {ln}: {text!r}")
            for name, typ, dims, init, is_par in d:
                self.add(name, typ, dims, init, is_par, into=cur if in_type else None)
        return len(stmts)

    def c_decls(self):
        o = ["#include <complex.h>"]
        for sname, comps in self.structs.items():
            o.append("typedef struct {")
            for name, c in comps.items():
                if c[0] == "scalar":
                    o.append(f"  {CT[c[1]]} {name};")
                elif c[0] == "alloc":
                    o.append(f"  {CT[c[1]]}* {name}; int {name}_d1, {name}_l1;")
                else:
                    o.append(f"  {CT[c[1]]} {name}[{c[2][0][1] - c[2][0][0] + 1}];")
            o.append("} " + sname.upper() + ";")
        for cname, t, c in self.param_decls:
            o.append(f"static const {CT[t]} {cname} = {c};")
        for name in self.order:
            if name in self.alloc:
                t, rank = self.alloc[name]
                o.append(f"static __thread {CT[t]}* {name}; static __thread int " + ", ".join(f"{name}_d{k + 1}, {name}_l{k + 1} = 1" for k in range(rank)) + ";")
            elif name in self.fixed:
                t, b = self.fixed[name]
                n = 1
                for lo, hi in b:
                    n *= hi - lo + 1
                o.append(f"static __thread {CT[t]} {name}[{n}];")
            else:
                o.append(f"static __thread {CT[self.scalars[name]]} {name};")
        return o


class UnitG(Unit):
    def __init__(self, tr, mod, kind, name, args):
        super().__init__(tr, kind, name, args)
        self.mod = mod
        self.pre = []            # statements to run before the one being translated (array-function calls into temporaries)
        self.ptr = {}            # pointer name -> the array section it is associated with (an Arr of lvalues)
        self.result_shape = None
        self.case_open = []

    # ---- names
    def is_mod(self, name):
        if name in self.types or name in self.args or name == self.name:
            return False
        m = self.mod
        return name in m.scalars or name in m.fixed or name in m.alloc or name in m.params or re.fullmatch(r"[a-z0-9_]+_[dl][123]", name) is not None

    def vtype(self, name):
        if name in self.types:
            return self.types[name]
        m = self.mod
        for tab in (m.scalars,):
            if name in tab:
                return tab[name]
        if name in m.fixed:
            return m.fixed[name][0]
        if name in m.alloc:
            return m.alloc[name][0]
        if name in m.params:
            return m.params[name][0]
        if name in m.func_types and not isinstance(m.func_types[name], tuple):
            return m.func_types[name]
        if re.fullmatch(r"[a-z0-9_]+_[dl][123]", name):
            return INT
        if self.tr is not None and name in self.tr.host:
            return self.tr.host[name]
        if self.tr is not None and name in self.tr.host_ptr:
            return self.tr.host_ptr[name]
        raise SyntaxError(f"{self.name}: {name!r} is not declared")

    def note(self, name):
        if self.tr is not None and (name in self.tr.host or name in self.tr.host_ptr) and name not in self.args:
            return                                        # the host's variable (file scope)
        if not self.is_mod(name):
            super().note(name)

    def ref(self, name):
        if self.is_mod(name):
            return self.mod.param_cname.get(name, name) if name in self.mod.params else name
        if name in self.args and self.types.get(name) in STRUCT_OF:
            return name                                   # a pointer to the struct
        if self.tr is not None and name in self.tr.host_ptr and name not in self.args:
            return f"(*h_{name})"
        if self.tr is not None and name in self.tr.host and name not in self.args:
            return name
        return super().ref(name)

    def addr(self, name):
        if self.tr is not None and name in self.tr.host_ptr and name not in self.args:
            return f"h_{name}"
        if self.tr is not None and name in self.tr.host and name not in self.args:
            return "&" + name
        if self.is_mod(name):
            return name if (name in self.mod.fixed or name in self.mod.alloc) else "&" + name
        return super().addr(name)

    # ---- arrays: bounds known at translation time (fixed) or at run time (allocatable: rank 1 only here)
    def bounds(self, name):
        if name == self.name and self.result_shape:
            return [(1, n) for n in self.result_shape]
        if name in self.ptr:
            return [(1, n) for n in self.ptr[name].shape]
        if name in self.dims:
            try:
                return [(1, self.mod.const_int(d)) for d in self.dims[name]]
            except Exception:
                return None                               # a dummy array with run-time extents: plain element references only
        if name in self.mod.fixed and name not in self.types:
            return self.mod.fixed[name][1]
        if name in self.mod.lead and name not in self.types:
            return self.mod.lead[name] + [None]          # the last dimension is allocated at run time
        return None

    def elem(self, name, idx):
        """element Node of a fixed-shape array, idx = Python ints"""
        b = self.bounds(name)
        off, stride = 0, 1
        for (lo, hi), i in zip(b, idx):
            assert lo <= i <= hi, (name, idx)
            off += (i - lo) * stride
            stride *= hi - lo + 1
        cname = name + "_result" if name == self.name else name
        return Node("elem", self.vtype(name) if name != self.name else self.result_type, f"{cname}[{off}]", name=name, idx=None)

    def whole(self, name):
        if name in self.ptr:
            return self.ptr[name]
        b = self.bounds(name)
        shape = [hi - lo + 1 for lo, hi in b]
        elems = []

        def rec(k, idx):
            if k < 0:
                elems.append(self.elem(name, idx))
                return
            for i in range(b[k][0], b[k][1] + 1):
                rec(k - 1, [i] + idx) if False else None
        # column-major: first index fastest
        import itertools
        if shape and max(shape) > 64:                     # a work array passed whole (call sort(vvv,...)): by reference, never scalarised
            a = Arr([1], [Node("lit", self.vtype(name), "0")])
            a.shape, a.whole_name = tuple(shape), name
            return a
        for idx in itertools.product(*[range(lo, hi + 1) for lo, hi in reversed(b)]):
            elems.append(self.elem(name, list(reversed(idx))))
        a = Arr(shape, elems)
        a.whole_name = name
        return a

    # ---- expression parser extensions
    def binop(self, op, a, b):
        if isinstance(a, Arr) or isinstance(b, Arr):
            if isinstance(a, Arr) and isinstance(b, Arr):
                assert a.shape == b.shape, (a.shape, b.shape)
                return Arr(a.shape, [self.binop(op, x, y) for x, y in zip(a.elems, b.elems)])
            if isinstance(a, Arr):
                return Arr(a.shape, [self.binop(op, x, b) for x in a.elems])
            return Arr(b.shape, [self.binop(op, a, y) for y in b.elems])
        x, y = self.promote(a, b)
        return Node("bin", max(a.typ, b.typ), f"({x} {op} {y})")

    def p_add(self):
        k, v = self.peek()
        if v in ("+", "-"):
            self.take()
            n = self.p_mul()
            if v == "-":
                n = Arr(n.shape, [Node("un", e.typ, f"(-{e.c})") for e in n.elems]) if isinstance(n, Arr) else Node("un", n.typ, f"(-{n.c})")
        else:
            n = self.p_mul()
        while self.peek()[1] in ("+", "-"):
            op = self.take()[1]
            n = self.binop(op, n, self.p_mul())
        return n

    def p_mul(self):
        n = self.p_pow()
        while self.peek()[1] in ("*", "/"):
            op = self.take()[1]
            n = self.binop(op, n, self.p_pow())
        return n

    def p_rel(self):
        n = self.p_add()
        v = self.peek()[1]
        if v in self.REL:
            self.take()
            r = self.p_add()
            a, b = self.promote(n, r)
            n = Node("rel", LOG, f"({a} {self.REL[v]} {b})")
        return n

    def p_prim(self):
        k, v = self.peek()
        if v == "[":                                      # array constructor of scalars
            self.take()
            elems = [self.p_or()]
            while self.peek()[1] == ",":
                self.take()
                elems.append(self.p_or())
            self.take("]")
            return Arr([len(elems)], elems)
        if v == "(":                                      # parenthesis or complex literal (re,im)
            save = self.pos
            self.take()
            a = self.p_or()
            if self.peek()[1] == ",":
                self.take()
                b = self.p_or()
                self.take(")")
                return Node("lit", CPX, f"(({self.cast(a, R8)}) + ({self.cast(b, R8)}) * I)")
            self.pos = save
        if k == "id" and self.bounds(v) is not None:
            if self.pos + 1 >= len(self.toks) or self.toks[self.pos + 1][1] != "(":
                if v in self.mod.lead and v not in self.ptr:
                    return super().p_prim()               # a whole rank-3 allocatable (Rdu = 0): handled by assign
                self.take()
                return self.whole(v)
            self.take()
            self.take()
            b = self.bounds(v)
            subs = []                                     # per dimension: int index expression Node, or (lo, hi) for a section
            for d in range(len(b)):
                if self.peek()[1] == ":":
                    self.take()
                    assert b[d] is not None, f"{v}: a section over the run-time dimension"
                    subs.append(b[d])
                else:
                    e = self.p_or()
                    if self.peek()[1] == ":":
                        self.take()
                        hi = self.p_or()
                        subs.append((int(eval(e.c)), int(eval(hi.c))))
                    elif e.kind == "lit" and e.typ == INT:
                        subs.append(int(e.c))
                    else:
                        subs.append(e)
                if d + 1 < len(b):
                    self.take(",")
            self.take(")")
            if all(not isinstance(s, tuple) for s in subs):
                return self.dyn_elem(v, subs)
            import itertools
            rng = [range(s[0], s[1] + 1) if isinstance(s, tuple) else [None] for s in subs]
            shape = [len(r) for r, s in zip(rng, subs) if isinstance(s, tuple)]
            elems = []
            for idx in itertools.product(*reversed(rng)):
                idx = list(reversed(idx))
                full = [i if i is not None else s for i, s in zip(idx, subs)]
                elems.append(self.dyn_elem(v, full))
            return Arr(shape, elems)
        n = super().p_prim()
        while self.peek()[1] == "%":                     # component of a derived-type dummy
            self.take()
            _, comp = self.take()
            c = self.mod.structs[STRUCT_OF[n.typ]][comp]
            if c[0] == "scalar":
                n = Node("elem", c[1], f"{n.c}->{comp}", name=None, idx=None)
                continue
            lo = f"{n.c}->{comp}_l1" if c[0] == "alloc" else str(c[2][0][0])
            if self.peek()[1] != "(":                     # the whole component array: MAXVAL, an actual argument, comp = comp / x
                n = Node("comp", c[1], f"{n.c}->{comp}", ptr=f"{n.c}->{comp}", count=f"{n.c}->{comp}_d1", name=None)
                continue
            self.take("(")
            i = self.p_or()
            if self.peek()[1] == ":":                     # a section with run-time bounds (MINVAL(GRT%vs(ifs+1:nlayers)))
                self.take()
                hi = self.p_or()
                self.take(")")
                li = self.cast(i, INT)
                n = Node("comp", c[1], "?", ptr=f"({n.c}->{comp} + (({li}) - {lo}))", count=f"(({self.cast(hi, INT)}) - ({li}) + 1)", name=None)
                continue
            self.take(")")
            n = Node("elem", c[1], f"{n.c}->{comp}[({self.cast(i, INT)}) - {lo}]", name=None, idx=None)
        return n

    def dyn_elem(self, name, subs):
        """element of a fixed-shape array with (possibly run-time) scalar subscripts"""
        if name in self.ptr:                              # an element of the section a pointer is associated with
            a = self.ptr[name]
            assert all(isinstance(s, int) for s in subs), (name, subs)
            k, stride = 0, 1
            for n, s in zip(a.shape, subs):
                k += (s - 1) * stride
                stride *= n
            return a.elems[k]
        b = self.bounds(name)
        off, stride = "0", 1
        for bd, s in zip(b, subs):
            i = str(s) if isinstance(s, int) else self.cast(s, INT)
            if bd is None:                                # run-time lower bound of the last dimension
                off += f" + (({i}) - {name}_l{len(b)}) * {stride}"
                continue
            lo, hi = bd
            off += f" + (({i}) - ({lo})) * {stride}"
            stride *= hi - lo + 1
        cname = name + "_result" if name == self.name else name
        return Node("elem", self.result_type if name == self.name else self.vtype(name), f"{cname}[{off}]", name=name, idx=None)

    def call_or_index(self, name, args):
        m = self.mod
        if name in m.alloc and name not in self.types:   # rank-1 allocatable module array
            return Node("elem", m.alloc[name][0], f"{name}[({self.cast(args[0], INT)}) - {name}_l1]", name=name, idx=None)
        if name == "allocated":
            return Node("call", LOG, f"({args[0].name} != 0)")
        if name in ("maxval", "minval") and args[0].kind == "comp":
            return Node("call", args[0].typ, f"f_{name}({args[0].ptr}, {args[0].count})")
        if name == "floor":
            return Node("call", INT, f"((int)floor({self.cast(args[0], R8)}))")
        if name == "nint":
            return Node("call", INT, f"((int)lround({self.cast(args[0], R8)}))")
        if name == "merge":
            a, b, c = args
            t = max(a.typ, b.typ)
            return Node("call", t, f"(({c.c}) ? ({self.cast(a, t)}) : ({self.cast(b, t)}))")
        if name == "matmul":
            a, b = args
            (n, k1), (k2, p) = a.shape, b.shape
            assert k1 == k2
            elems = []
            for j in range(p):
                for i in range(n):
                    acc = None
                    for k in range(k1):                   # k ascending
                        t = self.binop("*", a.elems[i + n * k], b.elems[k + k1 * j])
                        acc = t if acc is None else self.binop("+", acc, t)
                    elems.append(acc)
            return Arr((n, p), elems)
        if name == "exp" and isinstance(args[0], Arr):
            return Arr(args[0].shape, [self.call_or_index("exp", [e]) for e in args[0].elems])
        if name == "reshape":
            return Arr(tuple(int(e.c) for e in args[1].elems), args[0].elems)
        if name == "exp" and args[0].typ == CPX:
            return Node("call", CPX, f"cexp({args[0].c})")
        if name == "sqrt" and args[0].typ == CPX:
            return Node("call", CPX, f"csqrt({args[0].c})")
        if name == "dcmplx":
            return Node("call", CPX, f"((double _Complex)({self.cast(args[0], R8)}))")
        if name in ("aimag", "dimag"):
            return Node("call", R8, f"cimag({args[0].c})")
        if name == "dble" and args[0].typ == CPX:
            return Node("call", R8, f"creal({args[0].c})")
        ft = m.func_types.get(name)
        if isinstance(ft, tuple):                        # array-valued function: call into a temporary first
            self.tmp += 1
            t = f"fres{self.tmp}_"
            n = 1
            for s in ft[2]:
                n *= s
            self.pre.append(f"{CT[ft[1]]} {t}[{n}]; {name}_({self.actuals(args)}, {t});")
            return Arr(ft[2], [Node("elem", ft[1], f"{t}[{i}]", name=None, idx=None) for i in range(n)])
        return super().call_or_index(name, args)

    def actuals(self, args):
        out = []
        for a in args:
            if isinstance(a, Arr) and getattr(a, "whole_name", None) and a.whole_name not in self.ptr:
                out.append(f"(void*){a.whole_name + '_result' if a.whole_name == self.name else a.whole_name}")   # a whole array: by reference
                continue
            if isinstance(a, Arr):                        # an array value (e.g. a section): copy-in to a contiguous temporary
                self.tmp += 1
                t = f"arg{self.tmp}_"
                self.pre.append(f"{CT[a.typ]} {t}[{len(a.elems)}] = {{" + ", ".join(self.cast(e, a.typ) for e in a.elems) + "};")
                out.append(f"(void*){t}")
                continue
            if a.kind == "comp":
                out.append(f"(void*){a.ptr}")
                continue
            if a.kind == "var" and a.name in self.mod.func_types and a.name not in self.types:
                out.append("(void*)0")                    # a procedure as actual argument: the translated callee calls the driver's binding
                continue
            if a.kind == "var" and self.types.get(a.name) in STRUCT_OF:
                out.append(f"(void*){a.name}")
            elif a.kind == "var" and a.name not in self.params:
                out.append(f"(void*){self.addr(a.name)}")
            elif a.kind == "elem":
                out.append(f"(void*)&{a.c}")
            else:
                out.append(f"(void*)&({CT[a.typ]}){{{a.c}}}")
        return ", ".join(out)

    # ---- statements
    def assign(self, lhs, rhs):
        m = re.fullmatch(r"([a-z][a-z0-9_]*%[a-z][a-z0-9_]*)/(.+)", rhs)
        if m and m.group(1) == lhs:                       # GRT%mu = GRT%mu/mu0: elementwise over the whole component
            a = self.parse(lhs)
            d = self.parse(m.group(2))
            self.emit(f"{{ {CT[a.typ]} q_ = {self.cast(d, a.typ)}; for (int i_ = 0; i_ < {a.count}; ++i_) {a.ptr}[i_] = {a.ptr}[i_] / q_; }}")
            return
        ln = self.parse(lhs)
        rn = self.parse(rhs)
        for p in self.pre:
            self.emit(p)
        self.pre = []
        if isinstance(ln, Arr):
            if not isinstance(rn, Arr):
                rn = Arr(ln.shape, [rn] * len(ln.elems))
            n_l, n_r = len(ln.elems), len(rn.elems)
            assert n_l == n_r, (lhs, rhs, ln.shape, rn.shape)
            t = ln.typ
            self.tmp += 1
            tv = f"av{self.tmp}_"
            self.emit("{ " + f"{CT[t]} {tv}[{n_l}];")
            for i, e in enumerate(rn.elems):
                self.emit(f"  {tv}[{i}] = {self.cast(e, t)};")
            for i, e in enumerate(ln.elems):
                self.emit(f"  {e.c} = {tv}[{i}];")
            self.emit("}")
            return
        if ln.kind == "var" and ln.name in self.mod.alloc and ln.name not in self.types:   # RduL = 0
            a = ln.name
            t = self.vtype(a)
            rank = self.mod.alloc[a][1]
            n = f"{a}_d1" if rank == 1 else f"4 * {a}_d3"
            self.emit(f"for (int i_ = 0; i_ < {n}; ++i_) {a}[i_] = {self.cast(rn, t)};")
            return
        self.emit(f"{ln.c} = {self.cast(rn, ln.typ)};")

    def statement(self, text, ln):
        t = text
        m = re.fullmatch(r"([a-z][a-z0-9_]*)=>(.*)", t)
        if m:                                             # pointer association with a section: an alias from here on
            self.ptr.pop(m.group(1), None)
            self.ptr[m.group(1)] = self.parse(m.group(2))
            return
        m = re.fullmatch(r"dowhile\((.*)\)", t)
        if m:
            self.emit(f"while ({self.parse(m.group(1)).c}) {{")
            self.do_stack.append(("while", None, None))
            return
        if t == "cycle":
            self.emit("continue;")
            return
        if t.startswith("stop"):                         # STOP, STOP 'message'
            self.emit("f90_stopped = 1; return" + (";" if self.kind == "subroutine" or self.result_shape else f" {self.name}_result;"))
            return
        m = re.fullmatch(r"([a-z][a-z0-9_]*)\(([^():,]+):([^():,]+)\)=([a-z][a-z0-9_]*)\(([^():,]+):([^():,]+)\)", t)
        if m and self.bounds(m.group(1)) is not None and len(self.bounds(m.group(1))) == 1:
            # a(l1:u1) = b(l2:u2) with run-time bounds (possibly overlapping): the right-hand side is taken first
            a, l1, u1, b, l2, u2 = m.groups()
            ct = CT[self.vtype(a)]
            lo_a, lo_b = self.bounds(a)[0][0], self.bounds(b)[0][0]
            self.emit(f"{{ int l1_ = {self.expr_c(l1, INT)}, n_ = ({self.expr_c(u1, INT)}) - l1_ + 1, l2_ = {self.expr_c(l2, INT)};")
            self.emit(f"  {ct}* t_ = ({ct}*)malloc(sizeof({ct}) * (size_t)(n_ > 0 ? n_ : 1));")
            self.emit(f"  for (int k_ = 0; k_ < n_; ++k_) t_[k_] = {b}[l2_ - ({lo_b}) + k_];")
            self.emit(f"  for (int k_ = 0; k_ < n_; ++k_) {a}[l1_ - ({lo_a}) + k_] = t_[k_];")
            self.emit("  free(t_); }")
            return
        if t == "do":                                     # DO without a control: left by EXIT / RETURN
            self.emit("for (;;) {")
            self.do_stack.append(("while", None, None))
            return
        if t == "enddo" and self.do_stack and self.do_stack[-1][0] == "while":
            self.do_stack.pop()
            self.emit("}")
            return
        m = re.fullmatch(r"selectcase\((.*)\)", t)
        if m:
            self.emit(f"switch ({self.expr_c(m.group(1), INT)}) {{")
            self.case_open.append(False)
            return
        m = re.fullmatch(r"case\((\d+)\)", t)
        if m:
            if self.case_open[-1]:
                self.emit("break;")
            self.case_open[-1] = True
            self.emit(f"case {m.group(1)}:")
            return
        if t == "endselect":
            self.case_open.pop()
            self.emit("break; }")
            return
        m = re.fullmatch(r"allocate\((.*)\)", t)
        if m:
            for item in self.split_top(m.group(1)):
                m2 = re.fullmatch(r"([a-z][a-z0-9_]*)\((.*)\)", item)
                name, sps = m2.group(1), self.split_top(m2.group(2))
                lead = 1
                if len(sps) == 3:                         # (2,2,lo:hi): the leading extents are constants
                    self.mod.lead[name] = [(1, self.mod.const_int(sps[0])), (1, self.mod.const_int(sps[1]))]
                    lead = self.mod.const_int(sps[0]) * self.mod.const_int(sps[1])
                k = len(sps)
                lo, hi = sps[-1].split(":") if ":" in sps[-1] else ("1", sps[-1])
                ct = CT[self.vtype(name)]
                self.emit(f"{name}_l{k} = {self.expr_c(lo, INT)}; {name}_d{k} = ({self.expr_c(hi, INT)}) - {name}_l{k} + 1;")
                self.emit(f"{name} = ({ct}*)calloc((size_t)({lead} * ({name}_d{k} > 0 ? {name}_d{k} : 0)) + 8, sizeof({ct}));")
            return
        m = re.fullmatch(r"deallocate\((.*)\)", t)
        if m:
            for item in self.split_top(m.group(1)):
                self.emit(f"free({item}); {item} = 0;")
            return
        m = re.fullmatch(r"call([a-z][a-z0-9_]*)", t)
        if m:
            self.emit(f"{m.group(1)}_();")
            return
        m = re.fullmatch(r"call([a-z][a-z0-9_]*)\((.*)\)", t)
        if m:
            args = [self.parse(a) for a in self.split_top(m.group(2))]
            act = self.actuals(args)
            for p in self.pre:
                self.emit(p)
            self.pre = []
            self.emit(f"{m.group(1)}_({act});")
            return
        m = re.fullmatch(r"if\((.*)\)then", t)
        if m and "==" in m.group(1):                      # (the == operator is in the base REL table; nothing special)
            pass
        super().statement(text, ln)


class TranslatorG:
    def __init__(self, mod):
        self.mod = mod
        self.units = []
        self.called = set()
        self.externs = {}        # dummy procedures (REAL*8, EXTERNAL :: f): name -> result type
        self.host = {}           # host-associated variables of translated internal procedures: name -> type (file-scope in C)
        self.host_ptr = {}       # ... that are DUMMY arguments of the host: name -> type; the host stores its pointer in h_<name>
        self.host_of = set()     # the hosts that do so

    def scan_functions(self, stmts):
        """result types of the FUNCTION units (needed before their callers are translated)"""
        cur = None
        for text, ln in stmts:
            m = re.fullmatch(r"(?:(real\*8|complex\*16|real\(kind=[a-z0-9_]+\)))?function([a-z][a-z0-9_]*)\((.*)\)", text)
            if m:
                cur = m.group(2)
                if m.group(1):
                    self.mod.func_types[cur] = self.mod.type_of(m.group(1))
                continue
            if cur and cur not in self.mod.func_types:
                d = self.mod.declare(text)
                if d:
                    for name, typ, dims, init, is_par in d:
                        if name == cur:
                            self.mod.func_types[cur] = ("arr", typ, tuple(int(x) for x in dims)) if dims else typ
            if text.startswith("endfunction"):
                cur = None

    def run(self, stmts, start, only=None):
        u, in_spec, skipping = None, False, False
        for text, ln in stmts[start:]:
            if skipping:
                m = re.fullmatch(r"(?:(?:real\*8|complex\*16|real\(kind=[a-z0-9_]+\)))?(subroutine|function)([a-z][a-z0-9_]*)(?:\((.*)\))?", text)
                if not (m and only is not None and m.group(2) in only):   # (an internal procedure of a skipped host may be selected)
                    skipping = re.fullmatch(r"end(subroutine|function)[a-z0-9_]*", text) is None
                    continue
                skipping = False
            if u is not None and text == "contains":      # the unit's own statements end here; its internal procedures follow
                assert not u.do_stack, f"{ln}: unterminated do in {u.name}"
                u.emit("return;" if u.kind == "subroutine" or u.result_shape else f"return {u.name}_result;")
                u = None
                continue
            if u is None:
                m = re.fullmatch(r"(?:(?:real\*8|complex\*16|real\(kind=[a-z0-9_]+\)))?(subroutine|function)([a-z][a-z0-9_]*)(?:\((.*)\))?", text)
                if not m and re.fullmatch(r"end(subroutine|function)[a-z0-9_]*", text):
                    continue                              # the end of a host whose body ended at CONTAINS
                if not m:
                    if re.fullmatch(r"endmodule[a-z0-9_]*", text) or only is not None:
                        continue                          # (with a selection, statements of unselected units in other forms)
                    raise SyntaxError(f"line {ln}: {text!r} outside a unit")
                if only is not None and m.group(2) not in only:
                    skipping = True
                    continue
                u = UnitG(self, self.mod, m.group(1), m.group(2), m.group(3).split(",") if m.group(3) else [])
                ft = self.mod.func_types.get(u.name)
                if isinstance(ft, tuple):
                    u.result_shape, u.result_type = ft[2], ft[1]
                elif ft is not None:
                    u.result_type = ft
                    u.types[u.name] = ft
                self.units.append(u)
                in_spec = True
                continue
            if re.fullmatch(r"end(subroutine|function)[a-z0-9_]*", text):
                assert not u.do_stack, f"{ln}: unterminated do in {u.name}"
                u.emit("return;" if u.kind == "subroutine" or u.result_shape else f"return {u.name}_result;")
                u = None
                continue
            if in_spec:
                if text == "implicitnone":
                    continue
                if re.fullmatch(r"use[a-z0-9_,:]+", text):
                    continue
                d = self.mod.declare(text)
                if d is not None:
                    for name, typ, dims, init, is_par in d:
                        if name == u.name:
                            continue                      # the function result: known from scan_functions
                        u.types[name] = typ
                        if is_par:
                            u.params[name] = None
                            u.params[name] = u.cast(u.parse(init), typ)
                        elif "external" in text.split("::")[0]:
                            self.externs[name] = typ      # a dummy procedure: the driver supplies name_
                        elif dims and all(d_ == ":" for d_ in dims) and name in u.args:
                            u.dims[name] = ["20000"] * len(dims)   # assumed shape: lower bound 1 (only rank 1 occurs: ccc)
                        elif dims:
                            u.dims[name] = dims
                        elif name not in u.args and name not in self.host:
                            u.locals[name] = typ                      # (a host-table name is the file-scope variable, in every unit)
                    continue
                in_spec = False
            u.statement(text, ln)
        return self

    def c_source(self, paths):
        o = [f"/* GENERATED by oracle/f90toc_love.py from {' and '.join(paths)} -- do not edit, do not commit (oracle/_ref/ is git-ignored). */",
             "/* compile with -fcx-fortran-rules -ffp-contract=off */",
             "#include <math.h>", "#include <stdlib.h>",
             "static inline double f_sq(double x) { return x * x; }",
             "static inline double f_powi(double x, int n) { double r = 1.0; int m = n < 0 ? -n : n; while (m--) r *= x; return n < 0 ? 1.0 / r : r; }",
             "static inline int f_mini(int a, int b) { return a < b ? a : b; }", "static inline int f_maxi(int a, int b) { return a > b ? a : b; }",
             "static inline double f_min(double a, double b) { return a < b ? a : b; }", "static inline double f_max(double a, double b) { return a > b ? a : b; }",
             "static inline double f_maxval(const double* a, int n) { double r = a[0]; for (int i = 1; i < n; ++i) if (a[i] > r) r = a[i]; return r; }",
             "static inline double f_minval(const double* a, int n) { double r = a[0]; for (int i = 1; i < n; ++i) if (a[i] < r) r = a[i]; return r; }",
             "static __thread int f90_stopped;", ""]
        o += self.mod.c_decls() + [""]

        def proto(u):
            ps = ["void* " + a + "_a" for a in u.args]
            if u.result_shape:
                ps.append(f"{CT[u.result_type]}* {u.name}_result")
            ret = "void" if u.kind == "subroutine" or u.result_shape else CT[u.result_type]
            return f"static {ret} {u.name}_({', '.join(ps) or 'void'})"
        for name, t in self.host.items():
            o.append(f"static __thread {CT[t]}{'*' if t in STRUCT_OF else ''} {name}; /* a host-associated variable of an internal procedure */")
        for name, t in self.host_ptr.items():
            o.append(f"static __thread {CT[t]}* h_{name}; /* a dummy argument of a host, read by its internal procedure */")
        for name, t in self.externs.items():
            if name not in [u.name for u in self.units]:
                o.append(f"static {CT[t]} {name}_(void*, void*, void*, void*); /* the driver's */")
        for u in self.units:
            o.append(proto(u) + ";")
        o.append("")
        for u in self.units:
            o.append(proto(u) + " {")
            if u.name in self.host_of:
                for a in u.args:
                    if a in self.host and self.host[a] in STRUCT_OF:
                        o.append(f"  {a} = ({CT[self.host[a]]}*){a}_a; /* the file-scope pointer the internal procedures read */")
            for a in u.args:
                o.append(f"  {CT[u.vtype(a)]}* {a} = ({CT[u.vtype(a)]}*){a}_a;")
            if u.name in self.host_of:
                for a in u.args:
                    if a in self.host_ptr:
                        o.append(f"  h_{a} = {a};")
            if u.kind == "function" and not u.result_shape:
                o.append(f"  {CT[u.result_type]} {u.name}_result = 0;")
            for k, v in u.params.items():
                o.append(f"  const {CT[u.types[k]]} {k.upper()} = {v};")
            for name, dims in u.dims.items():
                if name not in u.args and name != u.name:
                    n = " * ".join(str(self.mod.const_int(d)) for d in dims)
                    o.append(f"  {CT[u.vtype(name)]} {name}[{n}];")
            for name, typ in sorted(u.locals.items()):
                if name not in u.args and name != u.name and name not in self.externs:
                    o.append(f"  {CT[typ]} {name} = 0;")
            o.extend("  " + s for s in u.body)
            o.append("}")
            o.append("")
        return "\n".join(o)


def read(path):
    """free-form statements, `;` separators split (no character data in these files' executable statements)"""
    out = []
    for text, ln in G.read_free_form(path):
        for part in text.split(";"):
            if part:
                out.append((part, ln))
    return out


def main():
    """f90toc_love.py GRT.f90 file[:unit,unit,...] [file[:units] ...] out.c"""
    grt, out = sys.argv[1], sys.argv[-1]
    mod = Mod()
    tr = TranslatorG(mod)
    s1 = read(grt)
    k1 = mod.read_header(s1)
    tr.scan_functions(s1)
    tr.run(s1, k1, only={"csq"})
    for spec in sys.argv[2:-1]:
        if spec.startswith("host="):                      # host=kt:real,c:real,grt:t_grt
            for item in spec[5:].split(","):
                nm, ty = item.split(":")
                if "*@" in ty:                            # v1:real*@cr0_finder -- a dummy of that host
                    ty, hostname = ty.split("*@")
                    tr.host_ptr[nm] = {"real": R8, "integer": INT}[ty]
                    tr.host_of.add(hostname)
                elif "@" in ty:                           # grt:t_grt@st_finder -- a struct dummy of that host
                    ty, hostname = ty.split("@")
                    tr.host[nm] = {"t_grt": TGRT}[ty]
                    tr.host_of.add(hostname)
                else:
                    tr.host[nm] = {"real": R8, "integer": INT, "t_grt": TGRT}[ty]
            continue
        parts = spec.split(":")
        path, units = parts[0], (parts[1] if len(parts) > 1 else "")
        st = read(path)
        k = 0
        if st[0][0].startswith("module") and "nohdr" not in parts[2:]:
            k = mod.read_header(st)
        tr.scan_functions(st)
        tr.run(st, k, only=set(units.split(",")) if units else None)
    open(out, "w").write(tr.c_source([grt] + sys.argv[2:-1]))
    print(f"f90toc_love: {len(tr.units)} program units ({', '.join(u.name for u in tr.units)})")


if __name__ == "__main__":
    main()
