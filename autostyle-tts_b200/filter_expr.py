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

The expression is tokenised ONCE: string literals are lifted out into named constants before any keyword
rewriting happens, so `speaker == "Tom AND Jerry"`, `text == "this is true"` or `speaker in ["IN"]` compare
against the literal exactly as written.  Arithmetic goes through a numeric-only helper: `"x" * 4000000000`
is a type error (false for the row), never an allocation.
"""
from __future__ import annotations

import ast
import re
from typing import Any, Callable, Dict, List, Tuple

from .schema import MilvusException

_ALLOWED = (ast.Expression, ast.BoolOp, ast.And, ast.Or, ast.UnaryOp, ast.Not, ast.USub, ast.UAdd, ast.Compare, ast.Eq,
            ast.NotEq, ast.Lt, ast.LtE, ast.Gt, ast.GtE, ast.In, ast.NotIn, ast.Name, ast.Load, ast.Constant, ast.List,
            ast.Tuple, ast.BinOp, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Mod, ast.Call, ast.Subscript)
_OPS = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/", ast.Mod: "%"}
_MAX_EXPR = 1 << 16
_MAX_LIST = 1 << 16


class _Missing:
    """Value of a field the row does not carry: compares false with everything."""
    def _f(self, *_):
        return False
    __eq__ = __ne__ = __lt__ = __le__ = __gt__ = __ge__ = __contains__ = _f
    __hash__ = object.__hash__

    def __getitem__(self, _):
        return self


_MISSING = _Missing()


def _like(value: Any, pattern: Any) -> bool:
    if not isinstance(value, str) or not isinstance(pattern, str):
        return False
    rx = "^" + ".*".join(re.escape(p) for p in pattern.split("%")) + "$"
    return re.match(rx, value, flags=re.S) is not None


def _arith(op: str, a: Any, b: Any) -> Any:
    """Numeric-only arithmetic: anything else (strings, lists, missing fields) is a type error -> row is false."""
    num = (int, float)
    if isinstance(a, bool) or isinstance(b, bool) or not isinstance(a, num) or not isinstance(b, num):
        raise TypeError("arithmetic on non-numeric operands")
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == "*":
        return a * b
    if op == "/":
        if isinstance(a, int) and isinstance(b, int):     # int64 / int64 truncates toward zero, as in the engine's C++
            q = abs(a) // abs(b)
            return q if (a >= 0) == (b >= 0) else -q
        return a / b
    return a % b


def _tokenise(expr: str) -> Tuple[str, List[str]]:
    """-> (expression with every string literal replaced by the name __lit_N__, the literal values).
    Operators `&& || !` become `and or not` in the same pass; nothing else is touched here."""
    out: List[str] = []
    lits: List[str] = []
    i, n = 0, len(expr)
    while i < n:
        c = expr[i]
        if c in "\"'":
            j = i + 1
            while j < n and expr[j] != c:
                j += 2 if expr[j] == "\\" else 1
            if j >= n:
                raise MilvusException(f"cannot parse expression: {expr}, error: unterminated string literal")
            try:
                val = ast.literal_eval(expr[i:j + 1])
            except (SyntaxError, ValueError) as e:
                raise MilvusException(f"cannot parse expression: {expr}, error: bad string literal") from e
            out.append(f" __lit_{len(lits)}__ ")
            lits.append(val)
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
    return "".join(out), lits


def _translate(expr: str) -> Tuple[str, List[str]]:
    if "__lit_" in expr or "__like__" in expr or "__arith__" in expr or "__meta__" in expr:
        raise MilvusException(f"cannot parse expression: {expr}, error: reserved identifier")
    s, lits = _tokenise(expr)                       # literal-free from here on: keyword rewrites cannot touch user text
    s = re.sub(r"\$meta\b", "__meta__", s)
    s = re.sub(r"\btrue\b", "True", s, flags=re.I)
    s = re.sub(r"\bfalse\b", "False", s, flags=re.I)
    s = re.sub(r"\b(AND|OR|NOT|IN|LIKE)\b", lambda m: m.group(1).lower(), s, flags=re.I)
    # `field like <literal>`  ->  __like__(field, <literal>)
    s = re.sub(r"([A-Za-z_][\w]*(?:\s*\[[^\]]+\])?)\s+like\s+(__lit_\d+__)", r"__like__(\1, \2)", s)
    return s.strip(), lits


class _ArithRewriter(ast.NodeTransformer):
    def visit_BinOp(self, node: ast.BinOp) -> ast.AST:
        self.generic_visit(node)
        op = _OPS.get(type(node.op))
        if op is None:
            raise MilvusException(f"unsupported operator {type(node.op).__name__}")
        call = ast.Call(func=ast.Name(id="__arith__", ctx=ast.Load()), args=[ast.Constant(op), node.left, node.right], keywords=[])
        return ast.copy_location(call, node)


def compile_filter(expr: str) -> Callable[[Dict[str, Any]], bool]:
    """-> predicate(row_fields) where row_fields maps field name -> value (dynamic fields flat, plus `$meta`)."""
    if len(expr) > _MAX_EXPR:
        raise MilvusException(f"cannot parse expression: expression longer than {_MAX_EXPR} characters")
    text, lits = _translate(expr)
    try:
        tree = ast.parse(text, mode="eval")
    except (SyntaxError, ValueError, RecursionError, MemoryError) as e:
        raise MilvusException(f"cannot parse expression: {expr}, error: {getattr(e, 'msg', type(e).__name__)}") from e
    for node in ast.walk(tree):
        if not isinstance(node, _ALLOWED):
            raise MilvusException(f"cannot parse expression: {expr}, error: unsupported construct {type(node).__name__}")
        if isinstance(node, ast.Call) and not (isinstance(node.func, ast.Name) and node.func.id == "__like__"):
            raise MilvusException(f"cannot parse expression: {expr}, error: function calls are not supported")
        if isinstance(node, (ast.List, ast.Tuple)) and len(node.elts) > _MAX_LIST:
            raise MilvusException(f"cannot parse expression: {expr}, error: list longer than {_MAX_LIST} elements")
        if isinstance(node, ast.Constant) and not isinstance(node.value, (int, float, bool)):
            raise MilvusException(f"cannot parse expression: {expr}, error: unsupported constant")
    tree = ast.fix_missing_locations(_ArithRewriter().visit(tree))
    code = compile(tree, "<filter>", "eval")
    consts = {f"__lit_{i}__": v for i, v in enumerate(lits)}
    consts["__like__"] = _like
    consts["__arith__"] = _arith

    class _Fields:
        """`$meta[...]` view of the current row: missing keys compare false."""
        __slots__ = ("row", "extra")

        def __getitem__(self, key):
            if key in self.extra:
                return self.extra[key]
            return self.row.get(key, _MISSING)

    class _Env(dict):
        """eval() locals: the literals / helpers live in the dict itself, field names fall through to the current
        row without copying it (one mapping for all rows of a bulk evaluation)."""
        __slots__ = ("view",)

        def __missing__(self, key):
            return self.view[key]

    env = _Env(consts)
    env.view = _Fields()
    env["__meta__"] = env.view
    glob = {"__builtins__": {}}

    def _eval_row(row: Dict[str, Any], extra: Dict[str, Any]) -> bool:
        env.view.row, env.view.extra = row, extra
        try:
            return bool(eval(code, glob, env))
        except (TypeError, ValueError, ZeroDivisionError, KeyError, IndexError, OverflowError, MemoryError):
            return False

    def predicate(fields: Dict[str, Any]) -> bool:
        return _eval_row(fields, {})

    def predicate_rows(metas, pk_name: str, pks):
        """Bulk form used by the client: one pass over the host-side rows, no per-row dict copies."""
        extra: Dict[str, Any] = {}
        out = [False] * len(pks)
        for r, pk in enumerate(pks):
            extra[pk_name] = pk
            out[r] = _eval_row(metas[r], extra)
        return out

    predicate.rows = predicate_rows
    return predicate
