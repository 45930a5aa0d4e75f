"""minimat -- a small interpreter for the subset of MATLAB that the reference's solver files use.  TEST INFRASTRUCTURE ONLY.

Why: the reference's hot path exists as MATLAB source (gqmap_gpu_mixture.m, gqmap_gpuSuper_mix_entropy.m, GaussHermite_2.m,
projsplx.m) and neither MATLAB nor Octave is available in the build container.  This module EXECUTES those files, unmodified and
read from /root/reference at run time, so that the oracle's C restatement can be compared with what the reference's own source
computes (tests/test_refsrc_parity.py, vectors under tests/golden/refsrc_*.npz made by tests/golden/make_refsrc_golden.py).
It is an independent implementation of MATLAB's documented language and built-in semantics (column-major N-d arrays, 1-based
indexing with `end`/colon/logical masks, implicit expansion, nested functions that share the parent workspace, multiple return
values, arrayfun, meshgrid/repmat/cat/circshift/sum/mean/min/max/...), not a restatement of the algorithm: it knows nothing about
optical flow.  gpuArray/gather are identities (the reference's GPU arrays hold the same IEEE doubles), `rand` is injectable so
that a run can be repeated by the oracle from the same initial state, and MEX calls are routed to callables supplied by the
caller (the tests pass oracle/refbin, i.e. the reference's own .mexw64 machine code).

Known, inherent differences to MATLAB proper: libm's exp/log/sqrt instead of MATLAB's (<= 1 ulp), NumPy's pairwise summation
order in sum/mean, LAPACK driver of eig.  All are at fp64 rounding level.

Only what the reference files need is implemented; anything else raises MatlabError loudly.
"""
import math
import os
import re

import numpy as np


class MatlabError(RuntimeError):
    ident = ""


# ======================================================================================================================
# lexer
# ======================================================================================================================
KEYWORDS = {"function", "end", "if", "elseif", "else", "while", "for", "break", "return", "continue"}
_num_re = re.compile(r"(\d+\.\d*|\d+|\.\d+)([eE][+-]?\d+)?")
_id_re = re.compile(r"[A-Za-z_]\w*")
_OPS2 = (".*", "./", ".^", ".'", "==", "~=", "<=", ">=", "&&", "||")
_OPS1 = "+-*/^<>=&|~:,;()[]{}@.'\\"


class Tok:
    __slots__ = ("kind", "val", "sp", "line")

    def __init__(self, kind, val, sp, line):
        self.kind, self.val, self.sp, self.line = kind, val, sp, line

    def __repr__(self):
        return "%s:%r@%d" % (self.kind, self.val, self.line)


def lex(src):
    toks, i, n, line, sp = [], 0, len(src), 1, False
    depth = 0                     # nesting of [] and {} (newlines inside are row separators -> ';')

    def prev_is_operand():
        if not toks:
            return False
        t = toks[-1]
        return t.kind in ("num", "id", "str") or (t.kind == "op" and t.val in (")", "]", "}", "'", ".'")) or \
            (t.kind == "kw" and t.val == "end")
    while i < n:
        c = src[i]
        if c in " \t\r":
            i += 1
            sp = True
            continue
        if c == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if src.startswith("...", i):
            while i < n and src[i] != "\n":
                i += 1
            i += 1
            line += 1
            sp = True
            continue
        if c == "\n":
            toks.append(Tok("op", ";" if depth else "\n", sp, line))
            i += 1
            line += 1
            sp = False
            continue
        if c == "'" and not (prev_is_operand() and not (sp and depth > 0)):
            j, out = i + 1, []                                   # string literal ('' is an escaped quote)
            while True:
                if j >= n or src[j] == "\n":
                    raise MatlabError("unterminated string at line %d" % line)
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        out.append("'")
                        j += 2
                        continue
                    break
                out.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(out), sp, line))
            i, sp = j + 1, False
            continue
        m = _num_re.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            text = m.group(0)
            if text.endswith(".") and m.end() < n and src[m.end()] in "*/^\\'":      # `2.*x` is 2 .* x
                text = text[:-1]
            toks.append(Tok("num", float(text), sp, line))
            i, sp = i + len(text), False
            continue
        m = _id_re.match(src, i)
        if m:
            w = m.group(0)
            toks.append(Tok("kw" if w in KEYWORDS else "id", w, sp, line))
            i, sp = m.end(), False
            continue
        for op in _OPS2:
            if src.startswith(op, i):
                toks.append(Tok("op", op, sp, line))
                i, sp = i + 2, False
                break
        else:
            if c in _OPS1:
                if c in "[{":
                    depth += 1
                elif c in "]}":
                    depth -= 1
                toks.append(Tok("op", c, sp, line))
                i, sp = i + 1, False
            else:
                raise MatlabError("unexpected character %r at line %d" % (c, line))
    toks.append(Tok("eof", None, False, line))
    return toks


# ======================================================================================================================
# parser  (AST = nested tuples)
# ======================================================================================================================
class FuncDef:
    def __init__(self, name, params, outs, body, nested, line):
        self.name, self.params, self.outs, self.body, self.nested, self.line = name, params, outs, body, nested, line
        self.parent = None
        self.vars = None           # names that are variables of this function's own workspace
        self.code = None


class Parser:
    def __init__(self, toks, end_terminated):
        self.t, self.i, self.end_terminated = toks, 0, end_terminated
        self.bracket = 0           # > 0 while parsing inside [ ] (whitespace separates elements)
        self.paren = 0             # > 0 while parsing inside ( ) nested in a [ ]

    # -- token helpers
    def peek(self, k=0):
        return self.t[min(self.i + k, len(self.t) - 1)]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def is_op(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == "op" and tok.val == v

    def is_kw(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == "kw" and tok.val == v

    def expect_op(self, v):
        tok = self.next()
        if tok.kind != "op" or tok.val != v:
            raise MatlabError("line %d: expected %r, got %r" % (tok.line, v, tok.val))
        return tok

    def skip_seps(self):
        while self.peek().kind == "op" and self.peek().val in ("\n", ";", ","):
            self.i += 1

    # -- file level
    def parse_file(self):
        funcs = []
        self.skip_seps()
        if self.peek().kind != "eof" and not self.is_kw("function"):          # a script: one body, its workspace is the result
            body = self.parse_block(())
            return [FuncDef("<script>", [], [], body, [], 1)]
        while self.peek().kind != "eof":
            if not self.is_kw("function"):
                raise MatlabError("line %d: expected `function`" % self.peek().line)
            funcs.append(self.parse_function())
            self.skip_seps()
        return funcs

    def parse_function(self):
        line = self.next().line                       # `function`
        outs = []
        if self.is_op("["):
            self.next()
            while not self.is_op("]"):
                if self.is_op(","):
                    self.next()
                    continue
                outs.append(self.next().val)
            self.next()
            self.expect_op("=")
            name = self.next().val
        else:
            name = self.next().val
            if self.is_op("="):
                self.next()
                outs = [name]
                name = self.next().val
        params = []
        if self.is_op("("):
            self.next()
            while not self.is_op(")"):
                if self.is_op(","):
                    self.next()
                    continue
                tok = self.next()
                params.append("~" if tok.val == "~" else tok.val)
            self.next()
        body, nested = self.parse_block(("end",) if self.end_terminated else (), in_function=True)
        if self.end_terminated:
            tok = self.next()
            if not (tok.kind == "kw" and tok.val == "end"):
                raise MatlabError("line %d: function %s is not closed by `end`" % (tok.line, name))
        return FuncDef(name, params, outs, body, nested, line)

    def parse_block(self, stop_kws, in_function=False):
        """statements up to (not including) one of stop_kws / eof (/ `function` when functions are not end-terminated)"""
        stmts, nested = [], []
        while True:
            self.skip_seps()
            tok = self.peek()
            if tok.kind == "eof":
                if stop_kws and not (in_function and not self.end_terminated):
                    raise MatlabError("unexpected end of file (missing `end`)")
                break
            if tok.kind == "kw" and tok.val in stop_kws:
                break
            if tok.kind == "kw" and tok.val == "function":
                if self.end_terminated and in_function:
                    nested.append(self.parse_function())
                    continue
                if not self.end_terminated and in_function:
                    break
                raise MatlabError("line %d: misplaced `function`" % tok.line)
            stmts.append(self.parse_statement())
        return (stmts, nested) if in_function else stmts

    # -- statements
    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "kw":
            if tok.val == "if":
                self.next()
                clauses, orelse = [], None
                cond = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                clauses.append((cond, body))
                while True:
                    k = self.next()
                    if k.val == "elseif":
                        cond = self.parse_expr()
                        clauses.append((cond, self.parse_block(("elseif", "else", "end"))))
                    elif k.val == "else":
                        orelse = self.parse_block(("end",))
                    elif k.val == "end":
                        break
                    else:
                        raise MatlabError("line %d: bad if" % k.line)
                return ("if", clauses, orelse, tok.line)
            if tok.val == "while":
                self.next()
                cond = self.parse_expr()
                body = self.parse_block(("end",))
                self.next()
                return ("while", cond, body, tok.line)
            if tok.val == "for":
                self.next()
                par = self.is_op("(")
                if par:
                    self.next()
                var = self.next().val
                self.expect_op("=")
                rng = self.parse_expr()
                if par:
                    self.expect_op(")")
                body = self.parse_block(("end",))
                self.next()
                return ("for", var, rng, body, tok.line)
            if tok.val in ("break", "return", "continue"):
                self.next()
                return (tok.val, tok.line)
            raise MatlabError("line %d: unexpected keyword %s" % (tok.line, tok.val))
        # multi-assignment  [a, b, ~] = f(...)
        if tok.kind == "op" and tok.val == "[":
            j, depth = self.i, 0
            while True:
                tj = self.t[j]
                if tj.kind == "eof":
                    break
                if tj.kind == "op" and tj.val == "[":
                    depth += 1
                if tj.kind == "op" and tj.val == "]":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            if self.t[j + 1].kind == "op" and self.t[j + 1].val == "=":
                self.next()
                lhs = []
                while not self.is_op("]"):
                    if self.is_op(","):
                        self.next()
                        continue
                    if self.is_op("~"):
                        self.next()
                        lhs.append(None)
                        continue
                    self.bracket += 1
                    lhs.append(self.parse_postfix())
                    self.bracket -= 1
                self.next()
                self.expect_op("=")
                return ("massign", lhs, self.parse_expr(), tok.line)
        e = self.parse_expr()
        if self.is_op("="):
            self.next()
            if e[0] not in ("name", "call", "field"):
                raise MatlabError("line %d: cannot assign to this expression" % tok.line)
            return ("assign", e, self.parse_expr(), tok.line)
        return ("expr", e, tok.line)

    # -- expressions (MATLAB precedence, lowest first)
    def parse_expr(self):
        return self.parse_oror()

    def _binary(self, sub, ops):
        a = sub()
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val in ops:
                if self.bracket and self.paren == 0 and tok.val in ("+", "-") and tok.sp and not self.peek(1).sp:
                    break                                  # `[a -b]`: a new element, not a subtraction
                self.next()
                a = ("bin", tok.val, a, sub())
            else:
                return a
        return a

    def parse_oror(self):
        return self._binary(self.parse_andand, ("||",))

    def parse_andand(self):
        return self._binary(self.parse_or, ("&&",))

    def parse_or(self):
        return self._binary(self.parse_and, ("|",))

    def parse_and(self):
        return self._binary(self.parse_cmp, ("&",))

    def parse_cmp(self):
        return self._binary(self.parse_range, ("==", "~=", "<", "<=", ">", ">="))

    def parse_range(self):
        a = self.parse_add()
        if self.is_op(":") and not self._colon_is_arg_end():
            self.next()
            b = self.parse_add()
            if self.is_op(":") and not self._colon_is_arg_end():
                self.next()
                c = self.parse_add()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def _colon_is_arg_end(self):
        nxt = self.peek(1)
        return nxt.kind == "op" and nxt.val in (")", ",")

    def parse_add(self):
        return self._binary(self.parse_mul, ("+", "-"))

    def parse_mul(self):
        return self._binary(self.parse_unary, ("*", "/", ".*", "./", "\\"))

    def parse_unary(self):
        tok = self.peek()
        if tok.kind == "op" and tok.val in ("-", "+", "~"):
            self.next()
            return ("un", tok.val, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        a = self.parse_postfix()
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val in ("^", ".^"):
                self.next()
                nt = self.peek()
                if nt.kind == "op" and nt.val in ("-", "+", "~"):       # 2^-1
                    self.next()
                    b = ("un", nt.val, self.parse_postfix())
                else:
                    b = self.parse_postfix()
                a = ("bin", tok.val, a, b)
            else:
                return a

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == "(" and not (self.bracket and self.paren == 0 and tok.sp):
                self.next()
                self.paren += 1
                saved, self.bracket = self.bracket, 0
                args = []
                while not self.is_op(")"):
                    if self.is_op(","):
                        self.next()
                        continue
                    if self.is_op(":") and self._colon_is_arg_end():
                        self.next()
                        args.append(("colon",))
                    else:
                        args.append(self.parse_expr())
                self.next()
                self.bracket = saved
                self.paren -= 1
                e = ("call", e, args)
            elif tok.kind == "op" and tok.val == "{" and not tok.sp:
                self.next()
                self.paren += 1
                saved, self.bracket = self.bracket, 0
                args = []
                while not self.is_op("}"):
                    if self.is_op(","):
                        self.next()
                        continue
                    args.append(self.parse_expr())
                self.next()
                self.bracket = saved
                self.paren -= 1
                e = ("cellidx", e, args)
            elif tok.kind == "op" and tok.val == "." and self.peek(1).kind == "id" and not tok.sp:
                self.next()
                e = ("field", e, self.next().val)
            elif tok.kind == "op" and tok.val in ("'", ".'") and not (self.bracket and self.paren == 0 and tok.sp):
                self.next()
                e = ("transpose", e)
            else:
                return e

    def parse_primary(self):
        tok = self.next()
        if tok.kind == "num":
            return ("num", tok.val)
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "id":
            return ("name", tok.val)
        if tok.kind == "kw" and tok.val == "end":
            return ("end",)
        if tok.kind == "op":
            if tok.val == "(":
                self.paren += 1
                saved, self.bracket = self.bracket, 0
                e = self.parse_expr()
                self.expect_op(")")
                self.bracket = saved
                self.paren -= 1
                return ("paren", e)
            if tok.val == "@":
                if self.is_op("("):                                  # anonymous function  @(a,b) expr
                    self.next()
                    params = []
                    while not self.is_op(")"):
                        if self.is_op(","):
                            self.next()
                            continue
                        params.append(self.next().val)
                    self.next()
                    saved_b, saved_p = self.bracket, self.paren
                    self.bracket, self.paren = 0, 0
                    body = self.parse_expr()
                    self.bracket, self.paren = saved_b, saved_p
                    return ("anon", params, body)
                return ("handle", self.next().val)
            if tok.val == "[":
                return self.parse_matrix()
            if tok.val == "{":
                m = self.parse_matrix(closer="}")
                if len(m[1]) > 1:
                    raise MatlabError("line %d: only 1 x n cell literals are supported" % tok.line)
                return ("cell", m[1][0] if m[1] else [])
        raise MatlabError("line %d: unexpected token %r" % (tok.line, tok.val))

    def parse_matrix(self, closer="]"):
        rows, row = [], []
        saved_b, saved_p = self.bracket, self.paren
        self.bracket, self.paren = 1, 0
        while True:
            tok = self.peek()
            if tok.kind == "op" and tok.val == closer:
                self.next()
                break
            if tok.kind == "op" and tok.val == ",":
                self.next()
                continue
            if tok.kind == "op" and tok.val in (";", "\n"):
                self.next()
                if row:
                    rows.append(row)
                row = []
                continue
            row.append(self.parse_expr())
        if row:
            rows.append(row)
        self.bracket, self.paren = saved_b, saved_p
        return ("matrix", rows)


def parse_source(src):
    toks = lex(src)
    try:
        return Parser(toks, True).parse_file()
    except MatlabError:
        return Parser(toks, False).parse_file()


# ======================================================================================================================
# values: Python float / bool for scalars, numpy arrays (>= 2-D, column-major semantics) otherwise, str, dict (struct)
# ======================================================================================================================
COLON = object()


class Cell(list):
    """1 x n cell array (only what `varargin` needs)"""


def norm(v):
    """canonical form of a computed value: 1x1 -> Python scalar, trailing singleton dimensions beyond 2 removed"""
    if isinstance(v, np.ndarray):
        if v.size == 1:
            x = v.reshape(-1)[0]
            return bool(x) if v.dtype == np.bool_ else float(x)
        if v.ndim < 2:
            return v.reshape((1, -1))
        if v.ndim > 2 and v.shape[-1] == 1:
            shp = list(v.shape)
            while len(shp) > 2 and shp[-1] == 1:
                shp.pop()
            return v.reshape(shp, order="F")
        return v
    if isinstance(v, np.generic):
        return v.item()
    return v


def arr(v):
    if isinstance(v, np.ndarray):
        return v
    if isinstance(v, (float, int, bool)):
        return np.array([[v]], dtype=np.bool_ if isinstance(v, bool) else np.float64)
    raise MatlabError("numeric value expected, got %r" % type(v))


def mshape(v):
    if isinstance(v, Cell):
        return (1, len(v)) if len(v) else (0, 0)
    if isinstance(v, np.ndarray):
        return v.shape if v.ndim >= 2 else (1,) + v.shape
    if isinstance(v, str):
        return (1, len(v)) if v else (0, 0)
    return (1, 1)


def flatF(a):
    return a.reshape(-1, order="F")


def first_nonsingleton(shape):
    for k, s in enumerate(shape):
        if s != 1:
            return k
    return 0


def _expand(a, b):
    """MATLAB implicit expansion: dimensions are aligned from the LEFT (missing trailing dimensions count as 1)"""
    if a.ndim < b.ndim:
        a = a.reshape(a.shape + (1,) * (b.ndim - a.ndim))
    elif b.ndim < a.ndim:
        b = b.reshape(b.shape + (1,) * (a.ndim - b.ndim))
    for x, y in zip(a.shape, b.shape):
        if x != y and x != 1 and y != 1:
            raise MatlabError("Matrix dimensions must agree (%s vs %s)" % (a.shape, b.shape))
    return a, b


def elementwise(op, a, b):
    a, b = _expand(arr(a), arr(b))
    return norm(op(a, b))


def is_scalar(v):
    return isinstance(v, (float, int, bool))


def truth(v):
    if is_scalar(v):
        return bool(v)
    a = arr(v)
    return a.size > 0 and bool(np.all(a != 0))


def _sub_to_index(s, dim):
    if s is COLON:
        return np.arange(dim)
    if is_scalar(s) and not isinstance(s, bool):
        k = int(s)
        if k != s or k < 1:
            raise MatlabError("Subscript indices must be positive integers (got %r)" % (s,))
        return np.array([k - 1])
    a = arr(s)
    if a.dtype == np.bool_:
        return np.nonzero(flatF(a))[0]
    f = flatF(a)
    k = f.astype(np.int64)
    if np.any(k != f) or np.any(k < 1):
        raise MatlabError("Subscript indices must be positive integers")
    return k - 1


def _fold_shape(shape, n):
    shape = tuple(shape)
    if n >= len(shape):
        return shape + (1,) * (n - len(shape))
    return shape[:n - 1] + (int(np.prod(shape[n - 1:])),)


def index_get(A, subs):
    if isinstance(A, str):
        r = index_get(np.array([[float(ord(c)) for c in A]]), subs)
        return "".join(chr(int(c)) for c in np.ravel(arr(r)))
    A = arr(A)
    n = len(subs)
    if n == 0:
        return norm(A)
    if n == 1:
        s = subs[0]
        flat = flatF(A)
        if is_scalar(s) and not isinstance(s, bool):
            k = int(s)
            if k != s or k < 1 or k > flat.size:
                raise MatlabError("Index exceeds matrix dimensions (linear index %r of %d)" % (s, flat.size))
            return flat[k - 1].item()
        if s is COLON:
            return norm(flat.reshape((-1, 1)).copy())
        sa = arr(s)
        idx = _sub_to_index(s, flat.size)
        if idx.size and idx.max() >= flat.size:
            raise MatlabError("Index exceeds matrix dimensions")
        out = flat[idx]
        if sa.dtype == np.bool_:
            shp = (1, -1) if (A.ndim == 2 and A.shape[0] == 1) else (-1, 1)
        elif A.ndim == 2 and min(A.shape) == 1 and sa.ndim == 2 and min(sa.shape) == 1:
            shp = (1, -1) if A.shape[0] == 1 else (-1, 1)          # vector indexed by vector: orientation of the source
        else:
            shp = sa.shape
        return norm(out.reshape(shp, order="F"))
    fs = _fold_shape(A.shape, n)
    if all(is_scalar(s) and not isinstance(s, bool) for s in subs):
        ks = []
        for s, d in zip(subs, fs):
            k = int(s)
            if k != s or k < 1 or k > d:
                raise MatlabError("Index exceeds matrix dimensions (%r of %r)" % (subs, fs))
            ks.append(k - 1)
        return A.reshape(fs, order="F").item(*ks)
    idx = [_sub_to_index(s, d) for s, d in zip(subs, fs)]
    for ix, d in zip(idx, fs):
        if ix.size and ix.max() >= d:
            raise MatlabError("Index exceeds matrix dimensions")
    return norm(np.asfortranarray(A.reshape(fs, order="F")[np.ix_(*idx)]))


def index_set(A, subs, value):
    """A(subs) = value with MATLAB value semantics (the array is copied, never modified in place)"""
    A = arr(A)
    v = value
    n = len(subs)
    if n == 1:
        s = subs[0]
        flat = flatF(A).copy()
        idx = _sub_to_index(s, flat.size)
        if idx.size and idx.max() >= flat.size:
            raise MatlabError("growing an array by indexed assignment is not supported")
        if not is_scalar(v) and arr(v).size != idx.size:
            raise MatlabError("In an assignment A(I) = B, the number of elements in B and I must be the same (%d vs %d)" % (arr(v).size, idx.size))
        flat[idx] = v if is_scalar(v) else flatF(arr(v))
        return flat.reshape(A.shape, order="F")
    empty = A.size == 0
    fs = _fold_shape(A.shape, n) if not empty else (0,) * n
    vshape = None if is_scalar(v) else arr(v).shape + (1,) * n
    if vshape is not None:                                   # value dimensions matched to the non-scalar subscripts, in order
        vdims = [d for d in vshape if d != 1]
    need, vi = [], 0
    for k, sub in enumerate(subs):
        if sub is COLON:
            if empty:
                if vshape is None:
                    raise MatlabError("A(:,...) = scalar on an undefined array")
                need.append(vshape[k])
            else:
                need.append(fs[k])
        else:
            ix = _sub_to_index(sub, fs[k])
            need.append(max(fs[k], int(ix.max()) + 1 if ix.size else 0))
    need = tuple(need)
    if empty or need != tuple(fs):                           # MATLAB grows the array, padding with zeros
        if not empty and n < A.ndim:
            raise MatlabError("growing an array through folded trailing dimensions is not supported")
        dt = A.dtype if not empty else (arr(v).dtype if not is_scalar(v) else np.float64)
        grown = np.zeros(need, dtype=dt, order="F")
        if not empty:
            grown[tuple(slice(0, d) for d in fs)] = A.reshape(fs, order="F")
        A, fs = grown, need
    idx = [_sub_to_index(s, d) for s, d in zip(subs, fs)]
    out = A.reshape(fs, order="F").copy(order="F")
    tgt = tuple(len(ix) for ix in idx)
    if is_scalar(v):
        out[np.ix_(*idx)] = v
    else:
        va = arr(v)
        if [d for d in va.shape if d != 1] != [d for d in tgt if d != 1]:
            raise MatlabError("Subscripted assignment dimension mismatch (%s into %s)" % (va.shape, tgt))
        out[np.ix_(*idx)] = va.reshape(tgt, order="F")
    return out.reshape(A.shape, order="F")


# ======================================================================================================================
# built-in functions: f(interp, nargout, *args) -> tuple of outputs
# ======================================================================================================================
def _dims(args):
    args = [a for a in args if not isinstance(a, str)]
    if len(args) == 1:
        a = args[0]
        if is_scalar(a):
            return (int(a), int(a))
        return tuple(int(x) for x in flatF(arr(a)))
    return tuple(int(a) for a in args)


def _reduce(fn, a, dim=None):
    a = arr(a)
    if a.dtype == np.bool_:
        a = a.astype(np.float64)
    ax = first_nonsingleton(a.shape) if dim is None else int(dim) - 1
    if ax >= a.ndim:
        return norm(a)
    return norm(fn(a, axis=ax, keepdims=True))


def _minmax(fn_red, fn_bin, nargout, args):
    if len(args) == 1:
        return (_reduce(fn_red, args[0]),)
    if len(args) == 2:
        a, b = args
        if is_scalar(a) and is_scalar(b):
            fa, fb = float(a), float(b)
            if fa != fa:
                return (fb,)
            if fb != fb:
                return (fa,)
            return (fn_bin(fa, fb),)
        return (elementwise(np.fmax if fn_bin is max else np.fmin, a, b),)
    raise MatlabError("min/max: unsupported call form")


def _repmat(a, reps):
    a = arr(a)
    nd = max(a.ndim, len(reps))
    a2 = a.reshape(a.shape + (1,) * (nd - a.ndim))
    reps = tuple(reps) + (1,) * (nd - len(reps))
    return norm(np.asfortranarray(np.tile(a2, reps)))


def _cat(dim, parts):
    dim = int(dim)
    parts = [arr(p) for p in parts if not (isinstance(p, np.ndarray) and p.size == 0)]
    nd = max([p.ndim for p in parts] + [dim])
    parts = [p.reshape(p.shape + (1,) * (nd - p.ndim)) for p in parts]
    return norm(np.asfortranarray(np.concatenate(parts, axis=dim - 1)))


def _circshift(a, k, dim=None):
    a = arr(a)
    ax = first_nonsingleton(a.shape) if dim is None else int(dim) - 1
    if ax >= a.ndim:
        return norm(a)
    return norm(np.asfortranarray(np.roll(a, int(k), axis=ax)))


def _size(nargout, a, dim=None):
    shp = mshape(a)
    if dim is not None:
        d = int(dim)
        return (float(shp[d - 1]) if d <= len(shp) else 1.0,)
    if nargout <= 1:
        return (np.array([[float(s) for s in shp]]),)
    out = [float(s) for s in shp[:nargout]] + [1.0] * max(0, nargout - len(shp))
    if nargout < len(shp):
        out[-1] = float(np.prod(shp[nargout - 1:]))
    return tuple(out)


def _arrayfun(interp, nargout, fh, *arrays):
    if not isinstance(fh, (FuncHandle, AnonFunc)):
        raise MatlabError("arrayfun: first argument must be a function handle")
    arrays = [arr(a) for a in arrays]
    # gpuArray arrayfun expands singleton dimensions of its inputs (the reference passes M x N index grids next to M x N x L state)
    nd = max(a.ndim for a in arrays)
    arrays = [a.reshape(a.shape + (1,) * (nd - a.ndim)) for a in arrays]
    shp = tuple(max(a.shape[d] for a in arrays) for d in range(nd))
    for a in arrays:
        if any(x != y and x != 1 for x, y in zip(a.shape, shp)):
            raise MatlabError("arrayfun: input sizes are not compatible (%s vs %s)" % (a.shape, shp))
    arrays = [np.asfortranarray(np.broadcast_to(a, shp)) for a in arrays]
    flats = [flatF(a).tolist() for a in arrays]
    k = max(nargout, 1)
    outs = [np.empty(len(flats[0])) for _ in range(k)]
    call = interp.call_handle
    for i, vals in enumerate(zip(*flats)):
        r = call(fh, vals, k)
        for o, x in zip(outs, r):
            if not is_scalar(x):
                raise MatlabError("arrayfun: function must return scalars")
            o[i] = x
    return tuple(norm(o.reshape(shp, order="F")) for o in outs)


def _eig(nargout, a):
    a = arr(a)
    if not np.allclose(a, a.T):
        raise MatlabError("eig: only symmetric matrices are supported")
    w, v = np.linalg.eigh(a)
    if nargout <= 1:
        return (w.reshape((-1, 1)),)
    return (np.asfortranarray(v), np.asfortranarray(np.diag(w)))


def _sort(nargout, a, mode="ascend"):
    a0 = arr(a)
    if not (a0.ndim == 2 and min(a0.shape) == 1) and a0.ndim != 3:
        raise MatlabError("sort: only vectors are supported")
    f = flatF(a0)
    idx = np.argsort(f, kind="stable")
    if mode == "descend":
        idx = np.argsort(-f, kind="stable")
    out = f[idx].reshape(a0.shape, order="F")
    ind = (idx + 1).astype(np.float64).reshape(a0.shape, order="F")
    return (norm(out), norm(ind))[:max(nargout, 1)]


def _diag(a, k=0):
    a = arr(a)
    k = int(k)
    if a.ndim == 2 and min(a.shape) == 1:
        return norm(np.asfortranarray(np.diag(flatF(a), k)))
    return norm(np.diag(a, k).reshape((-1, 1)))


def _repelem(a, *reps):
    a = arr(a)
    for ax, r in enumerate(reps):
        if ax < a.ndim:
            a = np.repeat(a, int(r), axis=ax)
    return norm(np.asfortranarray(a))


def _un(fn_scalar, fn_np):
    def f(interp, nargout, a):
        if is_scalar(a):
            return (fn_scalar(float(a)),)
        return (norm(fn_np(arr(a).astype(np.float64))),)
    return f


def _sqrt(x):
    if x < 0:
        raise MatlabError("sqrt of a negative number (complex results are not supported)")
    return math.sqrt(x)


def _log(x):
    if x < 0:
        raise MatlabError("log of a negative number (complex results are not supported)")
    return math.log(x) if x > 0 else -math.inf


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def _mod(a, b):
    if is_scalar(a) and is_scalar(b):
        return float(a) if b == 0 else float(a) - math.floor(float(a) / float(b)) * float(b)
    return elementwise(lambda x, y: x - np.floor(x / y) * y, a, b)


def _num2str(a):
    if isinstance(a, str):
        return a
    a = float(a)
    return "%d" % a if a == int(a) else "%.4g" % a


def _error(I, n, *a):
    ident = ""
    if len(a) >= 2 and isinstance(a[0], str) and ":" in a[0] and " " not in a[0]:       # error(id, msg, ...)
        ident, a = a[0], a[1:]
    msg = a[0] if a else "error"
    try:
        msg = msg.replace("%d", "%s") % tuple(a[1:]) if len(a) > 1 else msg
    except (TypeError, ValueError):
        pass
    e = MatlabError("error: " + str(msg))
    e.ident = ident
    raise e


def _struct(*a):
    if len(a) % 2:
        raise MatlabError("struct: field names and values must come in pairs")
    return {str(a[i]): a[i + 1] for i in range(0, len(a), 2)}


def _uint8(a):
    x = np.asarray(arr(a), dtype=np.float64)
    r = np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5))            # round half away from zero, then saturate
    return norm(np.asfortranarray(np.clip(np.nan_to_num(r, nan=0.0), 0, 255).astype(np.uint8)))


def _strfind(s, pat):
    hits = [i + 1.0 for i in range(len(s) - len(pat) + 1) if s.startswith(pat, i)] if pat else []
    return norm(np.array([hits])) if hits else np.zeros((0, 0))


_PREC = {"float32": "<f4", "single": "<f4", "int32": "<i4", "uint8": "u1", "double": "<f8", "float64": "<f8"}


def _fopen(I, n, name, mode="r"):
    try:
        f = open(name, {"r": "rb", "w": "wb"}[mode[0]])
    except OSError:
        return (-1.0,)
    fid = 3.0 + len(I.fids)
    I.fids[fid] = f
    return (fid,)


def _fread(I, n, fid, count, prec="uint8"):
    dt = np.dtype(_PREC[prec])
    k = -1 if math.isinf(float(count)) else int(count)
    data = np.frombuffer(I.fids[fid].read() if k < 0 else I.fids[fid].read(k * dt.itemsize), dtype=dt).astype(np.float64)
    return (norm(data.reshape((-1, 1))),) if data.size else (np.zeros((0, 0)),)


def _fwrite(I, n, fid, data, prec="uint8"):
    if isinstance(data, str):
        I.fids[fid].write(data.encode("latin-1"))
    else:
        I.fids[fid].write(flatF(arr(data)).astype(np.dtype(_PREC[prec])).tobytes())
    return ()


def _fclose(I, n, fid):
    I.fids.pop(fid).close()
    return ()


BUILTINS = {
    "clear": lambda I, n, *a: (),
    "save": lambda I, n, fn, *names: (I.on_save(fn, names, I.workspaces[-1]) if I.on_save else None, ())[1],
    "isfield": lambda I, n, s, f: (bool(isinstance(s, dict) and f in s),),
    "struct": lambda I, n, *a: (_struct(*a),),
    "ceil": _un(lambda x: float(math.ceil(x)) if math.isfinite(x) else x, np.ceil),
    "round": _un(lambda x: float(math.floor(abs(x) + 0.5)) * (1.0 if x >= 0 else -1.0) if math.isfinite(x) else x,
                 lambda v: np.where(v >= 0, np.floor(v + 0.5), -np.floor(-v + 0.5))),
    "onCleanup": lambda I, n, f: (Cleanup(f),),
    "error": _error,
    "uint8": lambda I, n, a: (_uint8(a),),
    "isnan": lambda I, n, a: (math.isnan(a) if is_scalar(a) else norm(np.isnan(arr(a).astype(np.float64))),),
    "atan2": lambda I, n, y, x: (math.atan2(y, x) if is_scalar(y) and is_scalar(x) else elementwise(np.arctan2, y, x),),
    "reshape": lambda I, n, a, *d: (norm(np.asfortranarray(arr(a).reshape(_dims(d), order="F"))),),
    "squeeze": lambda I, n, a: (a if is_scalar(a) or arr(a).ndim <= 2 else
                                norm(np.asfortranarray(arr(a).reshape([d for d in arr(a).shape if d != 1] or [1], order="F")
                                                       if sum(d != 1 for d in arr(a).shape) != 1 else
                                                       arr(a).reshape((-1, 1) if arr(a).shape[0] != 1 or arr(a).shape[1] == 1 else (1, -1),
                                                                      order="F"))),),
    "strfind": lambda I, n, s, p: (_strfind(s, p),),
    "strcmp": lambda I, n, a, b: (bool(isinstance(a, str) and isinstance(b, str) and a == b),),
    "fopen": _fopen, "fread": _fread, "fwrite": _fwrite, "fclose": _fclose,
    "gpuArray": lambda I, n, a: (a,),
    "gather": lambda I, n, *a: tuple(a[:max(n, 1)]),
    "double": lambda I, n, a: (norm(arr(a).astype(np.float64)),),
    "pi": lambda I, n: (math.pi,),
    "Inf": lambda I, n, *a: (math.inf,) if not a else (np.full(_dims(a), np.inf, order="F"),),
    "inf": lambda I, n, *a: (math.inf,) if not a else (np.full(_dims(a), np.inf, order="F"),),
    "NaN": lambda I, n, *a: (math.nan,) if not a else (norm(np.full(_dims(a), np.nan, order="F")),),
    "nan": lambda I, n, *a: (math.nan,) if not a else (norm(np.full(_dims(a), np.nan, order="F")),),
    "eps": lambda I, n: (2.0 ** -52,),
    "true": lambda I, n: (True,),
    "false": lambda I, n: (False,),
    "zeros": lambda I, n, *a: (norm(np.zeros(_dims(a), order="F")),),
    "ones": lambda I, n, *a: (norm(np.ones(_dims(a), order="F")),),
    "rand": lambda I, n, *a: (norm(I.rand(_dims(a) if [x for x in a if not isinstance(x, str)] else (1, 1))),),
    "size": lambda I, n, a, *d: _size(n, a, *d),
    "numel": lambda I, n, a: (float(np.prod(mshape(a))),),
    "length": lambda I, n, a: (float(max(mshape(a)) if np.prod(mshape(a)) else 0),),
    "isempty": lambda I, n, a: (bool(np.prod(mshape(a)) == 0),),
    "exp": _un(_exp, np.exp),
    "log": _un(_log, np.log),
    "sqrt": _un(_sqrt, np.sqrt),
    "floor": _un(lambda x: float(math.floor(x)) if math.isfinite(x) else x, np.floor),
    "abs": _un(abs, np.abs),
    "sum": lambda I, n, a, *d: (_reduce(np.sum, a, *d),),
    "mean": lambda I, n, a, *d: (_reduce(np.mean, a, *d),),
    "max": lambda I, n, *a: _minmax(np.nanmax, max, n, a),
    "min": lambda I, n, *a: _minmax(np.nanmin, min, n, a),
    "mod": lambda I, n, a, b: (_mod(a, b),),
    "meshgrid": lambda I, n, x, *y: tuple(np.asfortranarray(g) for g in np.meshgrid(flatF(arr(x)), flatF(arr(y[0] if y else x)))),
    "repmat": lambda I, n, a, *r: (_repmat(a, _dims(r)),),
    "cat": lambda I, n, d, *p: (_cat(d, p),),
    "circshift": lambda I, n, a, k, *d: (_circshift(a, k, *d),),
    "repelem": lambda I, n, a, *r: (_repelem(a, *r),),
    "arrayfun": _arrayfun,
    "eig": lambda I, n, a: _eig(n, a),
    "sort": lambda I, n, a, *m: _sort(n, a, *m),
    "diag": lambda I, n, a, *k: (_diag(a, *k),),
    "num2str": lambda I, n, a: (_num2str(a),),
    "fprintf": lambda I, n, *a: (I.on_fprintf(I.workspaces[-1]) if I.on_fprintf and I.workspaces else None, ())[1],
    "disp": lambda I, n, *a: (),
    "tic": lambda I, n: (),
    "toc": lambda I, n: (),
    "mkdir": lambda I, n, *a: (),
    "imwrite": lambda I, n, *a: (I.on_imwrite(*a) if I.on_imwrite else None, ())[1],
}


# ======================================================================================================================
# compiler: AST -> Python closures f(frame)
# ======================================================================================================================
class Frame:
    __slots__ = ("L", "P", "endctx", "nargout")

    def __init__(self, L, P, nargout):
        self.L, self.P, self.endctx, self.nargout = L, P, None, nargout


class FuncHandle:
    def __init__(self, fdef, penv):
        self.fdef, self.penv = fdef, penv


class AnonFunc:
    """@(params) expr -- the workspace is captured BY VALUE when the handle is created, as in MATLAB"""
    def __init__(self, params, body, captured, body0=None):
        self.params, self.body, self.captured, self.body0 = params, body, captured, body0


class Cleanup:
    """onCleanup(f): f runs when the function that holds the object returns (normally or through an error)"""
    def __init__(self, fn):
        self.fn = fn


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


def _assigned_names(stmts, out):
    for s in stmts:
        k = s[0]
        if k == "assign":
            e = s[1]
            while e[0] in ("call", "field"):
                e = e[1]
            out.add(e[1])
        elif k == "massign":
            for e in s[1]:
                if e is None:
                    continue
                while e[0] in ("call", "field"):
                    e = e[1]
                out.add(e[1])
        elif k == "if":
            for _, body in s[1]:
                _assigned_names(body, out)
            if s[2]:
                _assigned_names(s[2], out)
        elif k == "while":
            _assigned_names(s[2], out)
        elif k == "for":
            out.add(s[1])
            _assigned_names(s[3], out)
    return out


def _has_end(node):
    if not isinstance(node, tuple):
        return False
    if node[0] == "end":
        return True
    if node[0] == "call":                   # an `end` inside a nested index belongs to that index
        return _has_end(node[1])
    return any(_has_end(x) for x in node[1:] if isinstance(x, (tuple, list))) if node[0] != "matrix" else \
        any(_has_end(e) for row in node[1] for e in row)


_ARITH = {
    "+": (lambda a, b: a + b, np.add), "-": (lambda a, b: a - b, np.subtract),
    ".*": (lambda a, b: a * b, np.multiply),
    "==": (lambda a, b: a == b, np.equal), "~=": (lambda a, b: a != b, np.not_equal),
    "<": (lambda a, b: a < b, np.less), "<=": (lambda a, b: a <= b, np.less_equal),
    ">": (lambda a, b: a > b, np.greater), ">=": (lambda a, b: a >= b, np.greater_equal),
}


def _sdiv(a, b):
    a, b = float(a), float(b)
    if b == 0.0:
        if a == 0.0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _spow(a, b):
    a, b = float(a), float(b)
    if b == 2.0:
        return a * a
    try:
        r = a ** b
    except ZeroDivisionError:
        return math.inf
    except OverflowError:
        return math.inf
    if isinstance(r, complex):
        raise MatlabError("complex power results are not supported")
    return r


class Compiler:
    def __init__(self, interp, fdef, file_funcs):
        self.I, self.f, self.file_funcs = interp, fdef, file_funcs
        own = set(p for p in fdef.params if p != "~") | set(fdef.outs) | _assigned_names(fdef.body, set())
        pvars = fdef.parent.vars if fdef.parent is not None else set()
        fdef.vars = own                                    # what nested functions of THIS function may share
        self.local = set(fdef.params) | set(fdef.outs) | (own - pvars)
        self.shared = pvars - self.local
        self.siblings = {}
        if fdef.parent is not None:
            self.siblings = {n.name: n for n in fdef.parent.nested}
        self.children = {n.name: n for n in fdef.nested}

    # -- names
    def is_var(self, name):
        return name in self.local or name in self.shared

    def var_getter(self, name):
        if name in self.local:
            def g(fr):
                try:
                    return fr.L[name]
                except KeyError:
                    raise MatlabError("Undefined function or variable '%s' (in %s)" % (name, self.f.name))
            return g

        def g(fr):
            try:
                return fr.P[name]
            except KeyError:
                raise MatlabError("Undefined function or variable '%s' (in %s)" % (name, self.f.name))
        return g

    def var_setter(self, name):
        if name in self.local:
            def s(fr, v):
                fr.L[name] = v
        else:
            def s(fr, v):
                fr.P[name] = v
        return s

    def resolve_function(self, name):
        """-> callable(fr, args, nargout) -> tuple"""
        I = self.I
        if name in self.children:
            fd = self.children[name]
            return lambda fr, args, nargout: I.call_fdef(fd, args, nargout, fr.L)
        if name in self.siblings:
            fd = self.siblings[name]
            return lambda fr, args, nargout: I.call_fdef(fd, args, nargout, fr.P)
        if name in self.file_funcs:
            fd = self.file_funcs[name]
            return lambda fr, args, nargout: I.call_fdef(fd, args, nargout, None)
        if name in I.externals:
            ext = I.externals[name]
            return lambda fr, args, nargout: tuple(ext(nargout, *args))
        if name in BUILTINS:
            b = BUILTINS[name]
            return lambda fr, args, nargout: b(I, nargout, *args)
        fd = I.find_file_function(name)
        if fd is not None:
            return lambda fr, args, nargout: I.call_fdef(fd, args, nargout, None)
        raise MatlabError("Undefined function or variable '%s' (in %s, line %d)" % (name, self.f.name, self.f.line))

    # -- expressions
    def expr(self, node):
        k = node[0]
        if k == "num":
            v = node[1]
            return lambda fr: v
        if k == "str":
            v = node[1]
            return lambda fr: v
        if k == "paren":
            return self.expr(node[1])
        if k == "name":
            name = node[1]
            if self.is_var(name):
                return self.var_getter(name)
            fn_cache = []

            def callnoarg(fr):
                if not fn_cache:
                    fn_cache.append(self.resolve_function(name))
                r = fn_cache[0](fr, (), 1)
                if not r:
                    raise MatlabError("function %s returns no value" % name)
                return r[0]
            return callnoarg
        if k == "end":
            def endv(fr):
                A, pos, n = fr.endctx
                shp = _fold_shape(mshape(A), n)
                return float(shp[pos]) if n > 1 else float(np.prod(mshape(A)))
            return endv
        if k == "colon":
            return lambda fr: COLON
        if k == "handle":
            name = node[1]
            if name in self.children:
                fd = self.children[name]
                return lambda fr: FuncHandle(fd, fr.L)
            if name in self.siblings:
                fd = self.siblings[name]
                return lambda fr: FuncHandle(fd, fr.P)
            if name in self.file_funcs:
                fd = self.file_funcs[name]
                return lambda fr: FuncHandle(fd, None)
            raise MatlabError("@%s: only handles to functions of the same file are supported" % name)
        if k == "un":
            op, a = node[1], self.expr(node[2])
            if op == "-":
                def neg(fr):
                    v = a(fr)
                    return -v if is_scalar(v) else norm(-arr(v))
                return neg
            if op == "+":
                return a

            def lnot(fr):
                v = a(fr)
                return (not v) if is_scalar(v) else norm(arr(v) == 0)
            return lnot
        if k == "transpose":
            a = self.expr(node[1])

            def tr(fr):
                v = a(fr)
                if is_scalar(v):
                    return v
                v = arr(v)
                if v.ndim != 2:
                    raise MatlabError("transpose of an N-d array")
                return norm(np.asfortranarray(v.T))
            return tr
        if k == "range":
            a, st, b = self.expr(node[1]), (self.expr(node[2]) if node[2] is not None else None), self.expr(node[3])

            def rng(fr):
                lo, hi = float(a(fr)), float(b(fr))
                step = 1.0 if st is None else float(st(fr))
                if step == 0 or (step > 0 and lo > hi) or (step < 0 and lo < hi):
                    return np.zeros((1, 0))
                n = int(math.floor((hi - lo) / step + 1e-10)) + 1
                return (lo + step * np.arange(n)).reshape((1, -1)) if n != 1 else lo
            return rng
        if k == "bin":
            return self.binop(node)
        if k == "field":
            base, name = self.expr(node[1]), node[2]

            def fld(fr):
                s = base(fr)
                if not isinstance(s, dict):
                    raise MatlabError("field access .%s on a non-struct" % name)
                try:
                    return s[name]
                except KeyError:
                    raise MatlabError("Reference to non-existent field '%s'." % name)
            return fld
        if k == "matrix":
            rows = [[self.expr(e) for e in row] for row in node[1]]

            def mat(fr):
                vals = [[e(fr) for e in row] for row in rows]
                if not vals:
                    return np.zeros((0, 0))
                if any(isinstance(v, str) for row in vals for v in row):
                    if len(vals) != 1:
                        raise MatlabError("vertical concatenation of strings is not supported")
                    return "".join(v if isinstance(v, str) else chr(int(v)) for v in vals[0])
                rws = [_cat(2, row) if len(row) > 1 else row[0] for row in vals]
                return _cat(1, rws) if len(rws) > 1 else norm(rws[0])
            return mat
        if k == "call":
            return self.call(node, 1, single=True)
        if k == "cell":
            elems = [self.expr(e) for e in node[1]]
            return lambda fr: Cell(e(fr) for e in elems)
        if k == "anon":
            params, body_node = node[1], node[2]
            sub = Compiler.__new__(Compiler)
            sub.I, sub.f, sub.file_funcs = self.I, self.f, self.file_funcs
            sub.local = set(params) | self.local | self.shared            # everything resolves in the captured snapshot
            sub.shared = set()
            sub.siblings, sub.children = self.siblings, self.children
            sub._anon_parent = self
            body = sub.expr(body_node)
            body0 = sub.call(body_node, 0, single=False) if body_node[0] == "call" else None     # called as a statement: no output needed
            return lambda fr: AnonFunc(params, body, {**(fr.P or {}), **fr.L}, body0)
        if k == "cellidx":
            base, args = self.expr(node[1]), [self.expr(a) for a in node[2]]

            def cidx(fr):
                c = base(fr)
                if not isinstance(c, Cell) or len(args) != 1:
                    raise MatlabError("brace indexing is only supported on 1 x n cell arrays")
                i = int(args[0](fr))
                if i < 1 or i > len(c):
                    raise MatlabError("Index exceeds the number of cell elements")
                return c[i - 1]
            return cidx
        raise MatlabError("cannot compile expression node %r" % (k,))

    def binop(self, node):
        op, a, b = node[1], self.expr(node[2]), self.expr(node[3])
        if op in _ARITH:
            sf, nf = _ARITH[op]

            def f(fr):
                x, y = a(fr), b(fr)
                if is_scalar(x) and is_scalar(y):
                    return sf(x, y)
                return elementwise(nf, x, y)
            return f
        if op == "*":
            def mul(fr):
                x, y = a(fr), b(fr)
                if is_scalar(x) and is_scalar(y):
                    return x * y
                if is_scalar(x) or is_scalar(y):
                    return elementwise(np.multiply, x, y)
                xa, ya = arr(x), arr(y)
                if xa.ndim != 2 or ya.ndim != 2 or xa.shape[1] != ya.shape[0]:
                    raise MatlabError("inner matrix dimensions must agree")
                return norm(np.asfortranarray(xa @ ya))
            return mul
        if op in ("/", "./"):
            def div(fr):
                x, y = a(fr), b(fr)
                if is_scalar(x) and is_scalar(y):
                    return _sdiv(x, y)
                if op == "/" and not is_scalar(y):
                    raise MatlabError("matrix right division is not supported")
                with np.errstate(divide="ignore", invalid="ignore"):
                    return elementwise(np.divide, arr(x).astype(np.float64) if not is_scalar(x) else float(x),
                                       arr(y).astype(np.float64) if not is_scalar(y) else float(y))
            return div
        if op in ("^", ".^"):
            def pw(fr):
                x, y = a(fr), b(fr)
                if is_scalar(x) and is_scalar(y):
                    return _spow(x, y)
                if op == "^":
                    raise MatlabError("matrix power is not supported")
                if is_scalar(y) and float(y) == 2.0:
                    return elementwise(np.multiply, x, x)
                return elementwise(np.power, x, y)
            return pw
        if op == "&&":
            return lambda fr: truth(a(fr)) and truth(b(fr))
        if op == "||":
            return lambda fr: truth(a(fr)) or truth(b(fr))
        if op in ("&", "|"):
            fn = np.logical_and if op == "&" else np.logical_or

            def lg(fr):
                x, y = a(fr), b(fr)
                if is_scalar(x) and is_scalar(y):
                    return bool(fn(bool(x), bool(y)))
                return elementwise(lambda p, q: fn(p != 0, q != 0), x, y)
            return lg
        raise MatlabError("operator %s is not supported" % op)

    def args(self, arg_nodes, for_index):
        """-> evaluator(fr, A) -> list of values; sets the `end` context per argument when indexing"""
        comps = [self.expr(a) for a in arg_nodes]
        ends = [for_index and _has_end(a) for a in arg_nodes]
        n = len(comps)
        if not any(ends):
            return lambda fr, A: [c(fr) for c in comps]

        def ev(fr, A):
            out, saved = [], fr.endctx
            for pos, (c, he) in enumerate(zip(comps, ends)):
                if he:
                    fr.endctx = (A, pos, n)
                out.append(c(fr))
            fr.endctx = saved
            return out
        return ev

    def call(self, node, nargout, single):
        base, arg_nodes = node[1], node[2]
        if base[0] == "name" and not self.is_var(base[1]):
            name = base[1]
            if any(_has_end(a) for a in arg_nodes):
                raise MatlabError("`end` used in a call to function %s" % name)
            argev = self.args(arg_nodes, False)
            cache = []

            def docall(fr):
                if not cache:
                    cache.append(self.resolve_function(name))
                r = cache[0](fr, argev(fr, None), nargout)
                if single:
                    if not r:
                        raise MatlabError("function %s returns no value" % name)
                    return r[0]
                return r
            return docall
        getbase = self.expr(base)
        argev = self.args(arg_nodes, True)
        if len(arg_nodes) == 1 and arg_nodes[0][0] != "colon" and not _has_end(arg_nodes[0]):
            a0 = self.expr(arg_nodes[0])

            def idx1(fr):                           # hot path: X(k) with a scalar k
                A, s = getbase(fr), a0(fr)
                if isinstance(A, np.ndarray) and type(s) is float:
                    k = int(s)
                    if k == s and 1 <= k <= A.size:
                        r = A.reshape(-1, order="F")[k - 1].item()
                        return r if single else (r,)
                if isinstance(A, (FuncHandle, AnonFunc)):
                    r = self.I.call_handle(A, [s], nargout)
                    return r[0] if single else r
                r = index_get(A, [s])
                return r if single else (r,)
            return idx1

        def idx(fr):
            A = getbase(fr)
            if isinstance(A, (FuncHandle, AnonFunc)):
                r = self.I.call_handle(A, argev(fr, None), nargout)
                return (r[0] if r else None) if single else r
            r = index_get(A, argev(fr, A))
            return r if single else (r,)
        return idx

    # -- statements
    def lvalue_setter(self, e):
        """-> set(fr, value)"""
        if e[0] == "name":
            return self.var_setter(e[1])
        if e[0] == "call":
            base = e[1]
            if base[0] != "name":
                raise MatlabError("indexed assignment into a field or nested index is not supported")
            name = base[1]
            get, put = self.var_getter(name), self.var_setter(name)
            argev = self.args(e[2], True)

            isloc = name in self.local

            def seti(fr, v):
                env = fr.L if isloc else fr.P
                A = env[name] if name in env else np.zeros((0, 0))       # MATLAB creates the variable
                put(fr, norm(index_set(A, argev(fr, A), v)))
            return seti
        if e[0] == "field":
            base = e[1]
            if base[0] != "name":
                raise MatlabError("nested field assignment is not supported")
            name, fname = base[1], e[2]
            put = self.var_setter(name)
            isloc = name in self.local

            def setf(fr, v):
                env = fr.L if isloc else fr.P
                s = dict(env.get(name) or {})
                s[fname] = v
                put(fr, s)
            return setf
        raise MatlabError("bad assignment target")

    def block(self, stmts):
        comp = [self.stmt(s) for s in stmts]

        def run(fr):
            for c in comp:
                c(fr)
        return run

    def stmt(self, s):
        k = s[0]
        if k == "assign":
            put, val = self.lvalue_setter(s[1]), self.expr(s[2])
            line = s[3]

            def do(fr):
                try:
                    put(fr, val(fr))
                except MatlabError as e:
                    if "line " not in str(e)[:40]:
                        wrapped = MatlabError("%s line %d: %s" % (self.f.name, line, e))
                        wrapped.ident = e.ident
                        raise wrapped
                    raise
            return do
        if k == "massign":
            puts = [self.lvalue_setter(e) if e is not None else None for e in s[1]]
            rhs = s[2]
            n = len(puts)
            if rhs[0] != "call":
                raise MatlabError("multiple assignment needs a function call on the right-hand side")
            val = self.call(rhs, n, single=False)

            def do(fr):
                r = val(fr)
                if len(r) < n:
                    raise MatlabError("too many output arguments requested (%d of %d)" % (n, len(r)))
                for p, v in zip(puts, r):
                    if p is not None:
                        p(fr, v)
            return do
        if k == "expr":
            e = s[1]
            if e[0] == "call":
                c = self.call(e, 0, single=False)
                return lambda fr: c(fr)
            if e[0] == "name" and not self.is_var(e[1]):
                name, cache = e[1], []

                def do(fr):
                    if not cache:
                        cache.append(self.resolve_function(name))
                    cache[0](fr, (), 0)
                return do
            ev = self.expr(e)
            return lambda fr: ev(fr)
        if k == "if":
            clauses = [(self.expr(c), self.block(b)) for c, b in s[1]]
            orelse = self.block(s[2]) if s[2] else None

            def do(fr):
                for c, b in clauses:
                    if truth(c(fr)):
                        b(fr)
                        return
                if orelse:
                    orelse(fr)
            return do
        if k == "while":
            cond, body = self.expr(s[1]), self.block(s[2])

            def do(fr):
                while truth(cond(fr)):
                    try:
                        body(fr)
                    except _Break:
                        break
                    except _Continue:
                        continue
            return do
        if k == "for":
            put, rng, body = self.var_setter(s[1]), self.expr(s[2]), self.block(s[3])

            def do(fr):
                r = rng(fr)
                if is_scalar(r):
                    items = [r]
                else:
                    ra = arr(r)
                    if ra.ndim == 2 and ra.shape[0] == 1:
                        items = ra.reshape(-1).tolist()
                    else:
                        items = [norm(np.asfortranarray(ra.reshape((ra.shape[0], -1), order="F")[:, [c]])) for c in
                                 range(int(np.prod(ra.shape[1:])))]
                for it in items:
                    put(fr, it)
                    try:
                        body(fr)
                    except _Break:
                        break
                    except _Continue:
                        continue
            return do
        if k == "break":
            def do(fr):
                raise _Break()
            return do
        if k == "continue":
            def do(fr):
                raise _Continue()
            return do
        if k == "return":
            def do(fr):
                raise _Return()
            return do
        raise MatlabError("cannot compile statement %r" % (k,))


# ======================================================================================================================
# interpreter
# ======================================================================================================================
class Interp:
    """Interp(paths, externals={name: f(nargout, *args) -> tuple}, rand=f(shape) -> array, on_imwrite=f(img, filename),
    on_fprintf=f(workspace)): `on_fprintf` is called with the workspace (dict) of the innermost non-nested function whenever
    the program calls fprintf -- the reference's solver prints once per iteration, which makes it a per-iteration probe of
    state the function never returns (pn, rou, w, T)."""

    def __init__(self, paths, externals=None, rand=None, on_imwrite=None, on_fprintf=None):
        self.paths = list(paths)
        self.externals = dict(externals or {})
        self._rng = np.random.default_rng(0)
        self.rand = rand or (lambda shape: self._rng.random(int(np.prod(shape))).reshape(shape, order="F"))
        self.on_imwrite = on_imwrite
        self.on_fprintf = on_fprintf
        self.on_save = None            # f(filename, variable names, workspace) for `save(filename, 'a', 'b', ...)`
        self.workspaces = []           # workspaces of the active non-nested function calls, innermost last
        self.fids = {}                 # open files (fopen / fread / fwrite / fclose)
        self.files = {}                # function name -> FuncDef of the file's main function
        self.file_functions = {}       # function name -> all non-nested FuncDefs of that file
        self.last_workspace = None     # local workspace of the most recent top-level call (for state the function does not return)
        self.calls = 0

    def find_file_function(self, name):
        if name in self.files:
            return self.files[name]
        for p in self.paths:
            fn = os.path.join(p, name + ".m")
            if os.path.exists(fn):
                with open(fn, "r", encoding="latin-1") as f:
                    funcs = parse_source(f.read())
                file_funcs = {fd.name: fd for fd in funcs}

                def link(fd, parent):
                    fd.parent = parent
                    for nfd in fd.nested:
                        link(nfd, fd)
                for fd in funcs:
                    link(fd, None)
                    self._compile(fd, file_funcs)
                self.files[name] = funcs[0]
                self.file_functions[name] = funcs
                return funcs[0]
        return None

    def _compile(self, fd, file_funcs):
        c = Compiler(self, fd, file_funcs)            # sets fd.vars, which the nested functions' scope analysis needs
        fd.code = c.block(fd.body)
        for nfd in fd.nested:
            self._compile(nfd, file_funcs)

    def call_fdef(self, fd, args, nargout, penv):
        self.calls += 1
        L = {}
        if fd.params and fd.params[-1] == "varargin":
            nfix = len(fd.params) - 1
            L["varargin"] = Cell(args[nfix:])
            args = args[:nfix]
        elif len(args) > len(fd.params):
            raise MatlabError("%s: too many input arguments" % fd.name)
        for p, a in zip(fd.params, args):
            if p != "~":
                L[p] = a
        fr = Frame(L, penv, nargout)
        if penv is None:
            self.workspaces.append(L)
        try:
            fd.code(fr)
        except _Return:
            pass
        finally:
            if penv is None:
                self.workspaces.pop()
            for v in reversed(list(L.values())):
                if isinstance(v, Cleanup):
                    self.call_handle(v.fn, [], 0)
        k = max(nargout, 1) if fd.outs else 0
        if nargout > len(fd.outs):
            raise MatlabError("%s: too many output arguments" % fd.name)
        out = []
        for o in fd.outs[:k]:
            if o not in L:
                if nargout == 0:
                    break
                raise MatlabError("Output argument '%s' (and maybe others) not assigned during call to '%s'." % (o, fd.name))
            out.append(L[o])
        if penv is None:
            self.last_workspace = L
        return tuple(out)

    def call_handle(self, fh, args, nargout):
        if isinstance(fh, AnonFunc):
            if len(args) > len(fh.params):
                raise MatlabError("too many input arguments to an anonymous function")
            L = dict(fh.captured)
            L.update(zip(fh.params, args))
            if nargout == 0 and fh.body0 is not None:
                r = fh.body0(Frame(L, None, 0))
                return tuple(r) if isinstance(r, (tuple, list)) else ((r,) if r is not None else ())
            r = fh.body(Frame(L, None, nargout))
            return r if isinstance(r, tuple) else (r,)
        return self.call_fdef(fh.fdef, args, nargout, fh.penv)

    def run_script(self, path):
        """Execute a script file; returns its workspace (dict)."""
        with open(path, "r", encoding="latin-1") as f:
            funcs = parse_source(f.read())
        if len(funcs) != 1 or funcs[0].name != "<script>":
            raise MatlabError("%s is not a script" % path)
        fd = funcs[0]
        fd.parent = None
        self._compile(fd, {})
        self.call_fdef(fd, [], 0, None)
        return self.last_workspace

    def call(self, name, *args, nargout=1, local=None):
        """Call the main function of <name>.m found on the search path (or, with local=, another non-nested function of that
        file) with Python values (floats, NumPy arrays, dict structs)."""
        fd = self.find_file_function(name)
        if fd is None:
            raise MatlabError("no %s.m on the path %r" % (name, self.paths))
        if local is not None:
            cands = [f for f in self.file_functions[name] if f.name == local]
            if not cands:
                raise MatlabError("%s.m has no local function %s" % (name, local))
            fd = cands[0]

        def conv(v):
            if isinstance(v, dict):
                return {k: conv(x) for k, x in v.items()}
            if isinstance(v, np.ndarray):
                if v.dtype == np.bool_:
                    return norm(np.asfortranarray(v))
                return norm(np.asfortranarray(v, dtype=np.float64))
            if isinstance(v, (bool, str)):
                return v
            if isinstance(v, (int, float, np.generic)):
                return float(v)
            return v
        r = self.call_fdef(fd, [conv(a) for a in args], nargout, None)
        return r[0] if nargout == 1 else r
