"""The Julia binding cannot run here (no `julia` in the image), so its `ccall`s are checked statically: every
`ccall((:gb_xxx, LIB), Ret, (ArgTypes...), args...)` in grid.jl_b200/julia/GridB200.jl is parsed and compared with the
prototype of gb_xxx in include/gridb200.h -- same symbol, same return class, same number of arguments, and the same
class (int / int64 / double / size_t / pointer) in every position, and as many actual arguments as declared types."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "grid.jl_b200", "julia", "GridB200.jl")
HDR = os.path.join(ROOT, "include", "gridb200.h")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def _c_class(t):
    t = t.strip()
    if "*" in t:
        return "ptr"
    base = re.sub(r"\b(const|unsigned)\b", "", t).split()
    base = base[0] if base else ""
    return {"int": "i32", "int64_t": "i64", "uint64_t": "i64", "double": "f64", "size_t": "size", "void": "void"}[base]


def header_prototypes():
    src = _strip_c_comments(open(HDR).read())
    protos = {}
    for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(gb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        args = []
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                # drop the parameter name (last identifier), keep the type incl. any '*'
                t = re.sub(r"[A-Za-z_][A-Za-z0-9_]*$", "", prm).strip() if not prm.endswith("*") else prm
                args.append(_c_class(t))
        protos[name] = ("ptr" if "*" in ret else _c_class(ret), args)
    return protos


def _split_top(s):
    """split on commas at nesting depth 0"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, i):
    """s[i] == '(' -> index one past its matching ')'"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j + 1
    raise AssertionError("unbalanced ccall")


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Cint": "i32", "Int64": "i64", "Cdouble": "f64", "Csize_t": "size"}[t]


def julia_ccalls():
    src = re.sub(r"#[^\n]*", "", open(JL).read())  # comments off (no '#' inside strings near the ccalls)
    calls = []
    for m in re.finditer(r"ccall\(", src):
        end = _balanced(src, m.end() - 1)
        parts = _split_top(src[m.end():end - 1])
        sym = re.match(r"\(\s*(:gb_[a-z0-9_]+|\$\(QuoteNode\(sym\)\))\s*,\s*LIB\s*\)", parts[0])
        assert sym, parts[0]
        ret = parts[1]
        assert parts[2].startswith("("), parts[2]
        types = _split_top(parts[2][1:-1])
        calls.append((sym.group(1), ret, types, parts[3:]))
    return calls


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    assert len(protos) >= 60 and "gb_eval" in protos and "gb_restrict3_sharded" in protos
    calls = julia_ccalls()
    assert len(calls) >= 25
    seen = set()
    for sym, ret, types, actual in calls:
        names = ["gb_restrictb", "gb_restrict_extrap"] if sym.startswith("$") else [sym[1:]]  # the @eval'd pair shares one ccall
        for name in names:
            assert name in protos, "%s is not declared in include/gridb200.h" % name
            cret, cargs = protos[name]
            assert _jl_class(ret) == cret, (name, ret)
            assert len(types) == len(cargs), "%s: %d Julia argument types, %d C parameters" % (name, len(types), len(cargs))
            for k, (jt, ct) in enumerate(zip(types, cargs)):
                assert _jl_class(jt) == ct, "%s argument %d: Julia %s vs C %s" % (name, k + 1, jt, ct)
            assert len(actual) == len(types), "%s: %d values passed for %d declared argument types" % (name, len(actual), len(types))
            seen.add(name)
    # the entry points the binding is expected to cover
    for name in ("gb_grid_create", "gb_eval", "gb_eval_outer", "gb_invert_opt", "gb_filledges", "gb_pad1", "gb_restrict", "gb_prolong", "gb_restrict3",
                 "gb_prolong3", "gb_restrict12", "gb_prolong12", "gb_prolongb", "gb_interp_irregular", "gb_host_register", "gb_comm_init_all",
                 "gb_grid_replicate", "gb_eval_sharded", "gb_restrict3_sharded", "gb_prolong3_sharded", "gb_grid_destroy", "gb_comm_destroy"):
        assert name in seen, name


def test_constants_match_the_header():
    hdr = _strip_c_comments(open(HDR).read())
    jl = open(JL).read()

    def c_enum(name):
        m = re.search(r"\b%s\s*=\s*(\d+)" % name, hdr)
        assert m, name
        return int(m.group(1))

    m = re.search(r"const GB_FAST, GB_TILED, GB_PLAIN, GB_SORTED = (\d+), (\d+), (\d+), (\d+)", jl)
    assert m and [int(x) for x in m.groups()] == [c_enum("GB_FAST"), c_enum("GB_TILED"), c_enum("GB_PLAIN"), c_enum("GB_SORTED")]
    m = re.search(r"const OUT_VAL, OUT_GRAD, OUT_HESS, OUT_PACKED = (\d+), (\d+), (\d+), (\d+)", jl)
    assert m and [int(x) for x in m.groups()] == [c_enum("GB_OUT_VAL"), c_enum("GB_OUT_GRAD"), c_enum("GB_OUT_HESS"), c_enum("GB_OUT_PACKED")]
    m = re.search(r"const GB_HOST, GB_DEVICE = (\d+), (\d+)", jl)
    assert m and [int(x) for x in m.groups()] == [c_enum("GB_HOST"), c_enum("GB_DEVICE")]
    # boundary-condition, interpolation-order and dtype codes
    for jname, cname in (("BCnil", "GB_BC_NIL"), ("BCnan", "GB_BC_NAN"), ("BCna", "GB_BC_NA"), ("BCreflect", "GB_BC_REFLECT"), ("BCperiodic", "GB_BC_PERIODIC"),
                         ("BCnearest", "GB_BC_NEAREST"), ("BCfill", "GB_BC_FILL")):
        m = re.search(r"bccode\(::Type\{%s\}\) = (\d+)" % jname, jl)
        assert m and int(m.group(1)) == c_enum(cname), jname
    for jname, cname in (("InterpNearest", "GB_NEAREST"), ("InterpLinear", "GB_LINEAR"), ("InterpQuadratic", "GB_QUADRATIC"), ("InterpCubic", "GB_CUBIC")):
        m = re.search(r"itcode\(::Type\{%s\}\) = (\d+)" % jname, jl)
        assert m and int(m.group(1)) == c_enum(cname), jname
    for jname, cname in (("Float32", "GB_F32"), ("Float64", "GB_F64")):
        m = re.search(r"dtcode\(::Type\{%s\}\) = (\d+)" % jname, jl)
        assert m and int(m.group(1)) == c_enum(cname), jname
