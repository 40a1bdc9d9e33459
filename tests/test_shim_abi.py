"""Static ABI check of fortran/mctomo_b200_shim.f90 against include/mctomo_b200.h.

There is no Fortran compiler in the image, so the shim's `bind(C)` interfaces and derived types cannot be compiled
here.  This test reads both files as text and holds every interface to the C prototype it binds: the symbol exists,
the number of dummy arguments equals the number of C parameters, a C scalar is passed `value` with a kind of the same
width, a C pointer is either `type(c_ptr), value` or a by-reference dummy, and the `bind(C)` derived types list the
same sequence of 32-bit integers and doubles as the C structs (the reference's T_GRID layout,
/root/reference/src/settings.f90:20-28, is what `mct_grid` mirrors).
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_SCALAR_KINDS = {
    "int": {"c_int", "c_int32_t"},
    "int32_t": {"c_int32_t", "c_int"},
    "int64_t": {"c_int64_t", "c_long_long"},
    "long long": {"c_long_long", "c_int64_t"},
    "double": {"c_double"},
}


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def _c_header():
    h = open(os.path.join(ROOT, "include", "mctomo_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", " ", h, flags=re.S)
    h = re.sub(r"//[^\n]*", " ", h)
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_][\w \*]*?)\b(mct_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, re.S):
        name, params = m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else _split_top(params)
        kinds = []
        for p in plist:
            if "*" in p or "[" in p:
                kinds.append(("ptr", p))
            else:
                base = re.sub(r"\bconst\b", "", p).strip()
                base = " ".join(base.split()[:-1])  # drop the parameter name
                kinds.append(("val", base))
        protos[name] = kinds
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\w*\s*\{(.*?)\}\s*(mct_\w+)\s*;", h, re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ty = decl.split()[0]
            n = len(_split_top(decl[len(ty):]))
            fields += [ty] * n
        structs[m.group(2)] = fields
    return protos, structs


def _shim_lines():
    raw = open(os.path.join(ROOT, "fortran", "mctomo_b200_shim.f90")).read().split("\n")
    lines, cur = [], ""
    for ln in raw:
        ln = re.sub(r"!.*$", "", ln).rstrip()  # the shim has no '!' inside character literals of the parsed parts
        if not ln.strip():
            continue
        if ln.rstrip().endswith("&"):
            cur += ln.rstrip()[:-1] + " "
            continue
        lines.append((cur + ln.strip().lstrip("&")).strip())
        cur = ""
    return lines


def _shim():
    lines = _shim_lines()
    ifaces, types = {}, {}
    i = 0
    while i < len(lines):
        ln = lines[i]
        mt = re.match(r"type\s*,\s*bind\(c\)\s*::\s*(\w+)", ln, re.I)
        if mt:
            fields = []
            i += 1
            while not re.match(r"end\s+type", lines[i], re.I):
                ty, names = lines[i].split("::")
                kind = re.search(r"\((\w+)\)", ty).group(1).lower()
                fields += [kind] * len(_split_top(names))
                i += 1
            types[mt.group(1)] = fields
        mf = re.search(r"\b(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*(?:result\(\w+\)\s*)?bind\(c\s*,\s*name\s*=\s*'(\w+)'\)", ln, re.I)
        if mf:
            args = [a.strip().lower() for a in mf.group(3).split(",") if a.strip()]
            decl = {}
            i += 1
            while not re.match(r"end\s+(function|subroutine)", lines[i], re.I):
                if "::" in lines[i] and not lines[i].lower().startswith("import"):
                    ty, names = lines[i].split("::")
                    for nm in _split_top(names):
                        decl[re.sub(r"\(.*\)", "", nm).strip().lower()] = ty.strip().lower()
                i += 1
            ifaces[mf.group(4)] = [(a, decl.get(a)) for a in args]
        i += 1
    return ifaces, types


def test_shim_interfaces_match_the_c_prototypes():
    protos, _ = _c_header()
    ifaces, _ = _shim()
    assert len(ifaces) >= 30, sorted(ifaces)
    problems = []
    for name, args in ifaces.items():
        if name not in protos:
            problems.append(f"{name}: bound in the shim, not declared in the header")
            continue
        cpar = protos[name]
        if len(cpar) != len(args):
            problems.append(f"{name}: {len(args)} dummies, {len(cpar)} C parameters")
            continue
        for (an, aty), (ck, cty) in zip(args, cpar):
            if aty is None:
                problems.append(f"{name}: dummy {an} is not declared")
                continue
            by_value = re.search(r",\s*value\b", aty) is not None
            kind = re.search(r"\((\w+)\)", aty).group(1)
            if ck == "val":
                ok = by_value and kind in C_SCALAR_KINDS.get(cty, set())
                if not ok:
                    problems.append(f"{name}: {an} is `{aty}`, C has scalar `{cty}`")
            else:
                # a C pointer: either an address passed by value, or any by-reference dummy
                if by_value and not aty.startswith("type(c_ptr)"):
                    problems.append(f"{name}: {an} is `{aty}` by value, C has pointer `{cty}`")
                if not by_value and aty.startswith("type(mct_"):
                    if kind not in cty:
                        problems.append(f"{name}: {an} is `{aty}`, C has `{cty}`")
    assert not problems, "\n".join(problems)


def test_shim_derived_types_match_the_c_structs():
    _, structs = _c_header()
    _, types = _shim()
    width = {"int32_t": "i4", "int": "i4", "double": "f8", "c_int32_t": "i4", "c_int": "i4", "c_double": "f8"}
    seen = 0
    for name, ffields in types.items():
        if name not in structs:
            continue
        seen += 1
        cf = [width[t] for t in structs[name]]
        ff = [width[t] for t in ffields]
        assert cf == ff, (name, structs[name], ffields)
    assert seen >= 3, (sorted(types), sorted(structs))


def test_shim_wrapper_calls_pass_the_interfaces_argument_counts():
    """Every call of a bound entry point inside the shim's wrapper routines passes as many actual arguments as the
    interface has dummies (none of them is optional)."""
    ifaces, _ = _shim()
    lines = _shim_lines()
    in_iface, calls, problems = False, 0, []
    for ln in lines:
        low = ln.lower()
        if re.match(r"interface\b", low):
            in_iface = True
        elif re.match(r"end\s+interface", low):
            in_iface = False
        if in_iface:
            continue
        for m in re.finditer(r"\b(mct_[a-z0-9_]+)\s*\(", low):
            name = m.group(1)
            if name not in ifaces or re.search(r"\b(function|subroutine)\s+" + name, low):
                continue
            depth, j = 1, m.end()
            while depth and j < len(low):
                depth += (low[j] == "(") - (low[j] == ")")
                j += 1
            assert depth == 0, ln
            inner = low[m.end():j - 1]
            n = len(_split_top(inner)) if inner.strip() else 0
            calls += 1
            if n != len(ifaces[name]):
                problems.append(f"{name}: called with {n} arguments, interface has {len(ifaces[name])}: {ln[:120]}")
    assert calls >= 15, calls
    assert not problems, "\n".join(problems)
