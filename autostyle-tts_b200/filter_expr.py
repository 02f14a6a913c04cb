"""Scalar filter expressions for `MilvusClient.search(..., filter=...)` / `query(filter=...)`.

The reference plumbs `filter=filter_expr` through every search wrapper but always passes None
(`/root/reference/milvus/RAG.py:368,387`, `/root/reference/milvus/search_json.py:251`).  This module
evaluates the commonly used subset of the Milvus boolean-expression grammar ON THE HOST, over the
scalar/dynamic fields the client keeps per row, and produces a row mask; the device applies it as a
bitmap inside the scan (`avs_set_filter`).

Supported: comparisons `== != < <= > >=`, `in [..]` / `not in [..]`, `like "prefix%"` (also `%suffix`,
`%infix%`), `and or not` (also `&& || !`), parentheses, arithmetic `+ - * / %` on numeric fields, string
and numeric literals, `true/false`, and `$meta["key"]` / `field["key"]` access into JSON fields.
A field that a row does not have makes every comparison on it false for that row (Milvus semantics
for missing dynamic fields).
"""
from __future__ import annotations

import ast
import re
from typing import Any, Callable, Dict

from .schema import MilvusException

_ALLOWED = (ast.Expression, ast.BoolOp, ast.And, ast.Or, ast.UnaryOp, ast.Not, ast.USub, ast.UAdd, ast.Compare, ast.Eq,
            ast.NotEq, ast.Lt, ast.LtE, ast.Gt, ast.GtE, ast.In, ast.NotIn, ast.Name, ast.Load, ast.Constant, ast.List,
            ast.Tuple, ast.BinOp, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Mod, ast.Call, ast.Subscript)


class _Missing:
    """Value of a field the row does not carry: compares false with everything."""
    def _f(self, *_):
        return False
    __eq__ = __ne__ = __lt__ = __le__ = __gt__ = __ge__ = __contains__ = _f
    __hash__ = object.__hash__

    def __getitem__(self, _):
        return self


_MISSING = _Missing()


def _like(value: Any, pattern: str) -> bool:
    if not isinstance(value, str):
        return False
    rx = "^" + ".*".join(re.escape(p) for p in pattern.split("%")) + "$"
    return re.match(rx, value, flags=re.S) is not None


def _translate(expr: str) -> str:
    out, i, n = [], 0, len(expr)
    while i < n:                                    # leave string literals untouched
        c = expr[i]
        if c in "\"'":
            j = i + 1
            while j < n and expr[j] != c:
                j += 2 if expr[j] == "\\" else 1
            out.append(expr[i:j + 1])
            i = j + 1
            continue
        if expr.startswith("&&", i):
            out.append(" and "); i += 2; continue
        if expr.startswith("||", i):
            out.append(" or "); i += 2; continue
        if c == "!" and not expr.startswith("!=", i):
            out.append(" not "); i += 1; continue
        out.append(c)
        i += 1
    s = "".join(out)
    s = re.sub(r"\$meta\b", "__meta__", s)
    s = re.sub(r"\btrue\b", "True", s, flags=re.I)
    s = re.sub(r"\bfalse\b", "False", s, flags=re.I)
    s = re.sub(r"\b(AND|OR|NOT|IN|LIKE)\b", lambda m: m.group(1).lower(), s)
    # `field like "pat"`  ->  __like__(field, "pat")
    s = re.sub(r"([A-Za-z_][\w]*(?:\[[^\]]+\])?)\s+like\s+(\"(?:[^\"\\]|\\.)*\"|'(?:[^'\\]|\\.)*')", r"__like__(\1, \2)", s)
    return s.strip()


def compile_filter(expr: str) -> Callable[[Dict[str, Any]], bool]:
    """-> predicate(row_fields) where row_fields maps field name -> value (dynamic fields flat, plus `$meta`)."""
    try:
        tree = ast.parse(_translate(expr), mode="eval")
    except SyntaxError as e:
        raise MilvusException(f"cannot parse expression: {expr}, error: {e.msg}") from e
    for node in ast.walk(tree):
        if not isinstance(node, _ALLOWED):
            raise MilvusException(f"cannot parse expression: {expr}, error: unsupported construct {type(node).__name__}")
        if isinstance(node, ast.Call) and not (isinstance(node.func, ast.Name) and node.func.id == "__like__"):
            raise MilvusException(f"cannot parse expression: {expr}, error: function calls are not supported")
    code = compile(tree, "<filter>", "eval")

    class _Env(dict):
        def __missing__(self, key):
            return _MISSING

    def predicate(fields: Dict[str, Any]) -> bool:
        env = _Env(fields)
        env["__like__"] = _like
        env["__meta__"] = _Env(fields)
        try:
            return bool(eval(code, {"__builtins__": {}}, env))
        except (TypeError, ValueError, ZeroDivisionError, KeyError, IndexError):
            return False
    return predicate
