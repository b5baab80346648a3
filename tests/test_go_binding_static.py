"""The Go binding (go/demodel_b200.go) cannot be compiled in this image (no Go toolchain), so it is checked
statically against the header it binds (include/demodel_b200.h, the boundary at start.go:197-204): every
C.dm_* call must name a declared entry point and pass as many arguments as the declaration takes, every
C.DM_* constant and every C.dm_* type must exist, and the struct fields the Go code touches must be fields of the
C structs.  That is what cgo itself would check first."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip_comments(src, line_comment="//"):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(re.escape(line_comment) + r"[^\n]*", " ", src)


def _split_args(s):
    """Top-level comma split of an argument list (no outer parentheses)."""
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


def _call_args(src, start):
    """src[start] is the '(' of a call: return the text between it and its matching ')'."""
    depth = 0
    for i in range(start, len(src)):
        if src[i] == "(":
            depth += 1
        elif src[i] == ")":
            depth -= 1
            if depth == 0:
                return src[start + 1:i]
    raise AssertionError("unbalanced call")


def _header():
    h = _strip_comments(open(os.path.join(ROOT, "include", "demodel_b200.h")).read())
    funcs = {}
    for m in re.finditer(r"\b(dm_\w+)\s*\(", h):
        name = m.group(1)
        before = h[:m.start()].rstrip()
        if before.endswith(("(", ",", "=")) or before.endswith("return"):
            continue                                   # not a declaration
        args = _call_args(h, m.end() - 1)
        tail = h[m.end() - 1 + len(args) + 2:].lstrip()
        if not tail.startswith(";"):
            continue
        a = _split_args(args)
        funcs[name] = 0 if a == ["void"] else len(a)
    consts = set(re.findall(r"#define\s+(DM_\w+)", h)) | set(re.findall(r"\b(DM_[A-Z0-9_]+)\s*=", h))
    types = set(re.findall(r"\}\s*(dm_\w+)\s*;", h)) | set(re.findall(r"typedef\s+struct\s+\w+\s+(dm_\w+)\s*;", h))
    fields = {}
    for m in re.finditer(r"typedef\s+struct\s*\w*\s*\{([^{}]*)\}\s*(dm_\w+)\s*;", h):
        fields[m.group(2)] = set(re.findall(r"(\w+)\s*(?:\[[^\]]*\])?\s*;", m.group(1)))
    return funcs, consts, types, fields


def test_go_calls_match_the_header():
    funcs, consts, types, fields = _header()
    assert len(funcs) >= 46 and "dm_stream_open" in funcs and funcs["dm_stream_open"] == 4, sorted(funcs)
    go = open(os.path.join(ROOT, "go", "demodel_b200.go")).read()
    go = re.sub(r'"(?:[^"\\\n]|\\.)*"', '""', go)                    # string literals may hold commas, parentheses and //
    go = _strip_comments(go)
    calls = 0
    for m in re.finditer(r"\bC\.(dm_\w+)\s*\(", go):
        name = m.group(1)
        assert name in funcs, f"go/demodel_b200.go calls C.{name}, which include/demodel_b200.h does not declare"
        n = len(_split_args(_call_args(go, m.end() - 1)))
        assert n == funcs[name], f"C.{name}: {n} arguments in Go, {funcs[name]} in the header"
        calls += 1
    assert calls >= 30, calls
    for name in set(re.findall(r"\bC\.(DM_\w+)", go)):
        assert name in consts, f"C.{name} is not defined by the header"
    used_types = {t for t in re.findall(r"\bC\.(dm_\w+)\b(?!\s*\()", go)}
    for t in used_types:
        assert t in types, f"C.{t} is not a type of the header"
    # struct fields written through a C struct variable (var c C.dm_config ... c.field = ...)
    for m in re.finditer(r"\bvar\s+(\w+)\s+C\.(dm_\w+)", go):
        var, typ = m.group(1), m.group(2)
        if typ not in fields:
            continue
        end = go.find("\n}\n", m.end())                             # the variable lives until its function ends
        scope = go[m.end():end if end > 0 else len(go)]
        for f in set(re.findall(r"\b" + re.escape(var) + r"\.(\w+)", scope)):
            f = f.lstrip("_")                                        # cgo spells a field that is a Go keyword with a leading _
            assert f in fields[typ], f"{typ} has no field {f} (used as {var}.{f})"


def test_every_exported_entry_point_is_declared_once():
    """The header is the boundary: no duplicate declarations, and the library exports exactly what it declares
    (tests/test_abi.py checks the export side against the built .so)."""
    h = _strip_comments(open(os.path.join(ROOT, "include", "demodel_b200.h")).read())
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int|void|uint32_t|size_t)\s*\**\s*(dm_\w+)\s*\(", h, flags=re.M)
    assert len(names) == len(set(names)), sorted(n for n in names if names.count(n) > 1)


def test_ctypes_table_matches_the_header():
    """The Python mirror's signature table (demodel_b200/_lib.py) against the same parse of the header: same entry
    points, same number of arguments each - a mismatch there corrupts the stack silently instead of failing."""
    from demodel_b200 import _lib
    funcs, _, _, fields = _header()
    assert set(_lib.SIGNATURES) == set(funcs), (sorted(set(funcs) - set(_lib.SIGNATURES)), sorted(set(_lib.SIGNATURES) - set(funcs)))
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == funcs[name], f"{name}: {len(argtypes)} ctypes arguments, {funcs[name]} in the header"
    for cls, typ in ((_lib.DmStats, "dm_stats"), (_lib.DmLayer, "dm_layer"), (_lib.DmCheckpoint, "dm_checkpoint"), (_lib.DmConfig, "dm_config")):
        assert [f[0] for f in cls._fields_] == [f for f in _ordered_fields(typ)], typ


def _ordered_fields(typ):
    h = _strip_comments(open(os.path.join(ROOT, "include", "demodel_b200.h")).read())
    body = re.search(r"typedef\s+struct\s*\w*\s*\{([^{}]*)\}\s*" + typ + r"\s*;", h).group(1)
    return re.findall(r"(\w+)\s*(?:\[[^\]]*\])?\s*;", body)


def test_documents_name_only_entry_points_that_exist():
    """INTEGRATION.md is what a maintainer of the Go proxy reads: every dm_* name in it (and in DESIGN.md / README.md)
    must be an entry point, a type, or a prefix family (`dm_cache_alias_*`) of the header."""
    funcs, _, types, _ = _header()
    known_tools = {"dm_sha256sum"}                                    # tools/dm_sha256sum.py
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"\b(dm_[a-z0-9_]+)\b", text)):
            ok = name in funcs or name in types or name in known_tools or any(f.startswith(name.rstrip("_")) for f in funcs)
            assert ok, f"{doc} mentions {name}, which include/demodel_b200.h does not declare"


def test_go_source_is_at_least_well_bracketed():
    """No compiler here: the cheapest syntax check there is.  Brackets balance outside strings, runes and comments, and
    the cgo preamble includes the header this repository ships."""
    go = open(os.path.join(ROOT, "go", "demodel_b200.go")).read()
    assert '#include "demodel_b200.h"' in go and re.search(r'^import\s+"C"', go, flags=re.M)
    code = re.sub(r'"(?:[^"\\\n]|\\.)*"|`[^`]*`|\'(?:[^\'\\\n]|\\.)+\'', '""', go)
    code = _strip_comments(code)
    stack = []
    pairs = {")": "(", "]": "[", "}": "{"}
    for ch in code:
        if ch in "([{":
            stack.append(ch)
        elif ch in ")]}":
            assert stack and stack.pop() == pairs[ch], "unbalanced " + ch
    assert not stack
