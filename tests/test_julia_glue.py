"""Static checks of julia/GPRsm100a.jl against include/gpr_sm100a.h and the reference's method table.

No `julia` binary exists in this image, so the glue cannot be executed here.  What CAN be machine-checked without one:
  * every `ccall` names an entry point the header declares, with the same return type, the same number of arguments and
    argument types that map one-to-one onto the C prototype (Cint <-> int, Int64 <-> int64_t, Ptr{Float64} <-> double*, ...),
    and passes exactly as many values as it declares types;
  * block keywords and `end` balance (a truncated edit of the file would not parse);
  * the overloads that sit next to a reference method of the same arity are typed so that Julia's dispatch cannot report an
    ambiguity (the round-1 review found a definition-time TypeError in this file; the desk-check is now a test).
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "julia", "GPRsm100a.jl")
HDR = os.path.join(ROOT, "include", "gpr_sm100a.h")

C2J = {
    "int": {"Cint"}, "int64_t": {"Int64"}, "double": {"Cdouble", "Float64"}, "char": {"Cchar"},
    "const double*": {"Ptr{Float64}", "Ref{Float64}"}, "double*": {"Ptr{Float64}", "Ref{Float64}"},
    "const int*": {"Ptr{Cint}"}, "int*": {"Ptr{Cint}", "Ref{Cint}"},
    "int64_t*": {"Ref{Int64}", "Ptr{Int64}"}, "const char*": {"Cstring"},
    "void*": {"Ptr{Cvoid}"}, "const void*": {"Ptr{Cvoid}"},
}
for opaque in ("gpr_ctx", "gpr_model", "gpr_mgpu", "gpr_mgpu_model"):
    C2J[opaque + "*"] = {"Ptr{Cvoid}"}
    C2J[opaque + "**"] = {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"}


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def header_prototypes():
    src = _strip_c_comments(open(HDR).read())
    protos = {}
    for m in re.finditer(r"\b(int64_t|int|const char\s*\*)\s+(gpr_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).replace(" ", ""), m.group(2), m.group(3)
        ret = {"int": "int", "int64_t": "int64_t", "constchar*": "const char*"}[ret]
        types = []
        if args.strip() != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                # drop the parameter name (last identifier), keep qualifiers and stars
                t = re.sub(r"\s*\b\w+$", "", a) if not a.endswith("*") else a
                t = t.replace(" *", "*").replace("* ", "*").strip()
                types.append(t)
        protos[name] = (ret, types)
    return protos


def _strip_julia(src):
    """Remove # comments, docstrings and string literals (keeps line structure)."""
    src = re.sub(r'"""(.|\n)*?"""', '""', src)
    out = []
    for line in src.split("\n"):
        buf, i, in_s = [], 0, False
        while i < len(line):
            ch = line[i]
            if in_s:
                if ch == "\\":
                    i += 2
                    continue
                if ch == '"':
                    in_s = False
                i += 1
                continue
            if ch == '"':
                in_s = True
                buf.append('""')
                i += 1
                continue
            if ch == "#":
                break
            buf.append(ch)
            i += 1
        out.append("".join(buf))
    return "\n".join(out)


def _split_top(s):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return parts


def julia_ccalls():
    src = _strip_julia(open(JL).read())
    calls = []
    for m in re.finditer(r"ccall\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        body = src[m.end():i - 1]
        parts = _split_top(body)
        name = re.match(r"\(:(\w+),\s*LIB\)", parts[0]).group(1)
        ret = parts[1]
        types = _split_top(parts[2].strip()[1:-1])
        calls.append((name, ret, types, parts[3:]))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 20
    ret_map = {"int": "Cint", "int64_t": "Int64", "const char*": "Cstring"}
    for name, ret, types, args in calls:
        assert name in protos, f"{name} is not declared in include/gpr_sm100a.h"
        cret, ctypes_ = protos[name]
        assert ret == ret_map[cret], f"{name}: return {ret} vs C {cret}"
        assert len(types) == len(ctypes_), f"{name}: {len(types)} ccall types vs {len(ctypes_)} C parameters"
        assert len(args) == len(types), f"{name}: {len(args)} values passed for {len(types)} declared types"
        for k, (jt, ct) in enumerate(zip(types, ctypes_)):
            assert ct in C2J, f"{name}: unmapped C type {ct!r}"
            assert jt in C2J[ct], f"{name} argument {k + 1}: Julia {jt} does not match C {ct}"


def test_entry_points_of_the_path_are_all_bound():
    bound = {c[0] for c in julia_ccalls()}
    need = {"gpr_ctx_create", "gpr_ctx_destroy", "gpr_last_error", "gpr_model_create", "gpr_model_destroy", "gpr_model_set_x",
            "gpr_model_set_y", "gpr_update_cache", "gpr_loss", "gpr_grad", "gpr_nlml_grad", "gpr_fetch", "gpr_predict",
            "gpr_split_predict", "gpr_integrate", "gpr_sample_mvn", "gpr_mgpu_create", "gpr_mgpu_destroy",
            "gpr_mgpu_model_create", "gpr_mgpu_model_destroy", "gpr_mgpu_nlml_grad", "gpr_mgpu_last_error", "gpr_device_count"}
    assert need <= bound, sorted(need - bound)


def test_blocks_balance():
    src = _strip_julia(open(JL).read())
    opens = closes = 0
    depth_sq = 0
    for tok in re.finditer(r"\[|\]|\b(?:mutable\s+struct|struct|function|module|if|for|while|let|begin|try|do|quote|end)\b", src):
        t = tok.group(0)
        if t == "[":
            depth_sq += 1
        elif t == "]":
            depth_sq -= 1
        elif t == "end":
            if depth_sq == 0:           # a[2:end] is an index, not a block end
                closes += 1
        else:
            # `x = cond ? a : b for ...` style generators live inside brackets / parentheses: only count statement forms
            if t in ("for", "if") and _inside_brackets(src, tok.start()):
                continue
            opens += 1
    assert opens == closes, (opens, closes)
    assert src.count("(") == src.count(")") and src.count("{") == src.count("}") and src.count("[") == src.count("]")


def _inside_brackets(src, pos):
    line_start = src.rfind("\n", 0, pos) + 1
    seg = src[line_start:pos]
    return seg.count("[") > seg.count("]") or seg.count("(") > seg.count(")")


def test_overloads_cannot_be_ambiguous_with_the_reference():
    """Reference methods of the same name and arity (cited lines of /root/reference, restated as signatures):
         update_cache!(pc::AbstractPredictCache, md::AbstractGPRModel)                src/predict.jl:29
         loss(::MarginalLikelihood, md::AbstractGPRModel, tc::AbstractCostCache)      src/cost.jl:113
         grad!(dL, ::MarginalLikelihood, md::AbstractGPRModel, tc::AbstractCostCache) src/cost.jl:119
         predict_mean!(mu, md::AbstractGPRModel, xp, pc::AbstractPredictCache)        src/predict.jl:36
         predict!(mu, S, md::AbstractGPRModel, xp, pc::AbstractPredictCache)          src/predict.jl:42
       An overload is unambiguous when it is at least as specific in EVERY position: the model argument must therefore be typed
       (SM100 / AbstractGPRModel), never left as Any, wherever the reference types it."""
    src = _strip_julia(open(JL).read())
    m = re.search(r"update_cache!\(pc::Union\{SM100PredictCache,SM100SplitPredictCache\},\s*md(::\w+)?\)", src)
    assert m and m.group(1) in ("::AbstractGPRModel", "::SM100"), "two-argument update_cache! must type md"
    for fn in ("loss", "grad!", "predict_mean!", "predict!"):
        sigs = re.findall(r"^function " + re.escape(fn) + r"\((.*?)\)\s*(?:#.*)?$", src, flags=re.M)
        assert sigs, fn
        for sig in sigs:
            assert "md::SM100" in sig, f"{fn}({sig}): model argument must be typed"
    # the wrapper's supertype must satisfy K<:AbstractKernel (src/models.jl:3-4): parameters are carried, not `Any`
    assert re.search(r"struct SM100\{K<:AbstractKernel,T,P<:AbstractArray\{T\},X<:AbstractArray\{T,2\},M<:GPRModel\{K,T,P,X\}\} <: AbstractGPRModel\{K,T,P,X\}", src)
    assert "using Random" in src
