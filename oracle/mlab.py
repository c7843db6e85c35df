"""mlab — a small interpreter for the MATLAB subset the reference is written in.

TEST INFRASTRUCTURE ONLY.  Neither MATLAB nor GNU Octave exists in the build image, so the
reference's ``.m`` files cannot be run by their own interpreter.  This module is the next best pin:
it parses and executes the **untouched reference source text** (``/root/reference/*.m``, read at run
time, never copied) statement by statement — control flow, indexing, ``end``, ``nargin``, anonymous
functions, cell arrays, multiple outputs, ``break`` semantics, 1-based ranges — so every reading
decision that a hand restatement has to make (stop operators, ``k = size(H,2)``, which ``alpha`` is
used at ``k == maxit``, what ``x`` is after a breakdown ...) is made by the source itself, not by us.
``tests/golden/make_reference_golden.py`` runs the reference through it and commits the outputs as
fixtures; ``oracle/solvers.py`` and the CUDA path are then checked against those fixtures.

What is NOT the reference here, and is stated as such in DESIGN.md: the built-ins.  ``*``, ``'``,
``norm``, ``\\``, ``svd``, ``eig``, ``sort`` ... are mapped to NumPy / SciPy (LAPACK), following
MATLAB's documented algorithm choices (``mldivide``: triangular check -> Cholesky for symmetric
matrices with positive diagonal -> LU; rectangular -> QR with column pivoting).  Rounding-level
differences against MathWorks' own kernels remain possible; the structure of the computation is
the reference's.

Only what the reference's non-plotting code uses is implemented; anything else raises
``MlabError`` naming the construct, so a silent mis-execution is not possible.
"""
from __future__ import annotations

import math
import os
import re

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

__all__ = ["Interp", "MlabError", "Cell"]


class MlabError(Exception):
    pass


# ======================================================================================
# tokenizer
# ======================================================================================
KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "break", "continue", "return",
            "switch", "case", "otherwise"}
OPS3 = ("...",)
OPS2 = ("==", "~=", "<=", ">=", "&&", "||", ".*", "./", ".\\", ".^", ".'")
OPS1 = "+-*/\\^'<>=&|~,;()[]{}:@."


class Tok:
    __slots__ = ("kind", "val", "sp", "line")

    def __init__(self, kind, val, sp, line):
        self.kind, self.val, self.sp, self.line = kind, val, sp, line

    def __repr__(self):
        return f"Tok({self.kind},{self.val!r},sp={self.sp},l{self.line})"


_num_re = re.compile(r"(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")
_id_re = re.compile(r"[A-Za-z_][A-Za-z_0-9]*")


def tokenize(src: str):
    src = src.replace("\r\n", "\n").replace("\r", "\n")
    toks = []
    i, n, line = 0, len(src), 1
    depth = []  # bracket stack: '(' '[' '{'
    sp = False

    def prev_is_value():
        if not toks:
            return False
        t = toks[-1]
        if t.kind in ("num", "str"):
            return True
        if t.kind == "id":
            return t.val not in KEYWORDS or t.val == "end"
        return t.kind == "op" and t.val in (")", "]", "}", "'", ".'")

    while i < n:
        c = src[i]
        if c in " \t":
            sp = True
            i += 1
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
            if depth and depth[-1] == "(":
                raise MlabError(f"line {line}: newline inside parentheses without '...'")
            toks.append(Tok("nl", "\n", sp, line))
            line += 1
            i += 1
            sp = False
            continue
        if c == "'":
            in_br = bool(depth) and depth[-1] in "[{"
            if prev_is_value() and not (in_br and sp):
                toks.append(Tok("op", "'", sp, line))
                i += 1
                sp = False
                continue
            j = i + 1
            buf = []
            while True:
                if j >= n or src[j] == "\n":
                    raise MlabError(f"line {line}: unterminated string")
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        buf.append("'")
                        j += 2
                        continue
                    break
                buf.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(buf), sp, line))
            i = j + 1
            sp = False
            continue
        if c == '"':
            j = src.index('"', i + 1)
            toks.append(Tok("str", src[i + 1:j], sp, line))
            i = j + 1
            sp = False
            continue
        m = _num_re.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            text = m.group(0)
            # "1.*x" / "1./x" / "1.^x": the dot belongs to the operator
            if text.endswith(".") and i + len(text) < n and src[i + len(text)] in "*/\\^'":
                text = text[:-1]
            toks.append(Tok("num", float(text), sp, line))
            i += len(text)
            sp = False
            continue
        m = _id_re.match(src, i)
        if m:
            toks.append(Tok("id", m.group(0), sp, line))
            i = m.end()
            sp = False
            continue
        two = src[i:i + 2]
        if two in OPS2:
            toks.append(Tok("op", two, sp, line))
            i += 2
            sp = False
            continue
        if c in OPS1:
            if c in "([{":
                depth.append(c)
            elif c in ")]}":
                if not depth:
                    raise MlabError(f"line {line}: unbalanced '{c}'")
                depth.pop()
            toks.append(Tok("op", c, sp, line))
            i += 1
            sp = False
            continue
        raise MlabError(f"line {line}: unexpected character {c!r}")
    toks.append(Tok("nl", "\n", False, line))
    toks.append(Tok("eof", None, False, line))
    return toks


# ======================================================================================
# parser  (AST = nested tuples)
# ======================================================================================
class Function:
    def __init__(self, name, params, outs, body, filename):
        self.name, self.params, self.outs, self.body, self.filename = name, params, outs, body, filename
        self.locals = {}  # sibling local functions of the same file


class Parser:
    def __init__(self, toks, filename="<string>"):
        self.t = toks
        self.p = 0
        self.filename = filename
        self.in_index = 0   # > 0 while parsing the arguments of x(...) / x{...}: `end` is a value
        self.in_matrix = 0  # > 0 inside [ ] / { } literals: blanks separate elements

    # -- helpers ------------------------------------------------------------------
    def peek(self, k=0):
        return self.t[self.p + k]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def is_op(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == "op" and tok.val == v

    def is_kw(self, v, k=0):
        tok = self.peek(k)
        return tok.kind == "id" and tok.val == v

    def expect_op(self, v):
        tok = self.next()
        if tok.kind != "op" or tok.val != v:
            raise MlabError(f"{self.filename}:{tok.line}: expected {v!r}, got {tok.val!r}")
        return tok

    def skip_nl(self):
        while self.peek().kind == "nl" or self.is_op(";") or self.is_op(","):
            self.next()

    # -- file ---------------------------------------------------------------------
    def parse_file(self):
        funcs, script = [], []
        self.skip_nl()
        while self.peek().kind != "eof":
            if self.is_kw("function"):
                funcs.append(self.parse_function())
            else:
                script.append(self.parse_statement())
            self.skip_nl()
        return funcs, script

    def parse_function(self):
        self.next()  # function
        outs = []
        # forms: function name(...) | function out = name(...) | function [o1,o2] = name(...)
        if self.is_op("["):
            self.next()
            while not self.is_op("]"):
                tok = self.next()
                if tok.kind == "id":
                    outs.append(tok.val)
                elif tok.kind == "op" and tok.val in (",",):
                    continue
                elif tok.kind == "op" and tok.val == "~":
                    outs.append("~")
                else:
                    raise MlabError(f"{self.filename}:{tok.line}: bad output list")
            self.next()
            self.expect_op("=")
            name = self.next().val
        else:
            first = self.next()
            if self.is_op("="):
                self.next()
                outs = [first.val]
                name = self.next().val
            else:
                name = first.val
        params = []
        if self.is_op("("):
            self.next()
            while not self.is_op(")"):
                tok = self.next()
                if tok.kind == "id":
                    params.append(tok.val)
                elif tok.kind == "op" and tok.val == "~":
                    params.append("~")
            self.next()
        # tolerate stray text after the signature on the same line (run_ptr_rtp_comparison.m:1)
        while self.peek().kind != "nl":
            self.next()
        body = self.parse_block(("end", "function"))
        if self.is_kw("end"):
            self.next()
        return Function(name, params, outs, body, self.filename)

    def parse_block(self, terminators):
        stmts = []
        while True:
            self.skip_nl()
            tok = self.peek()
            if tok.kind == "eof":
                return stmts
            if tok.kind == "id" and tok.val in terminators:
                return stmts
            stmts.append(self.parse_statement())

    # -- statements ---------------------------------------------------------------
    def end_stmt(self):
        """consume the statement terminator; returns True when output is suppressed"""
        tok = self.peek()
        if tok.kind == "op" and tok.val == ";":
            self.next()
            return True
        if tok.kind == "op" and tok.val == ",":
            self.next()
            return False
        if tok.kind in ("nl", "eof"):
            return False
        if tok.kind == "id" and tok.val in KEYWORDS:
            return False
        raise MlabError(f"{self.filename}:{tok.line}: unexpected {tok.val!r} after statement")

    def parse_statement(self):
        tok = self.peek()
        line = tok.line
        if tok.kind == "id":
            v = tok.val
            if v == "if":
                return self.parse_if()
            if v == "for":
                self.next()
                paren = self.is_op("(")
                if paren:
                    self.next()
                var = self.next().val
                self.expect_op("=")
                e = self.parse_expr()
                if paren:
                    self.expect_op(")")
                body = self.parse_block(("end",))
                self.next()
                return ("for", var, e, body, line)
            if v == "while":
                self.next()
                c = self.parse_expr()
                body = self.parse_block(("end",))
                self.next()
                return ("while", c, body, line)
            if v == "switch":
                self.next()
                e = self.parse_expr()
                self.skip_nl()
                cases, other = [], None
                while True:
                    self.skip_nl()
                    if self.is_kw("case"):
                        self.next()
                        ce = self.parse_expr()
                        cases.append((ce, self.parse_block(("case", "otherwise", "end"))))
                    elif self.is_kw("otherwise"):
                        self.next()
                        other = self.parse_block(("case", "otherwise", "end"))
                    elif self.is_kw("end"):
                        self.next()
                        break
                    else:
                        raise MlabError(f"{self.filename}:{self.peek().line}: bad switch")
                return ("switch", e, cases, other, line)
            if v in ("break", "continue", "return"):
                self.next()
                self.end_stmt()
                return (v, line)
            # command syntax: `hold on`, `clear all`, `close all`, `grid on`
            nxt = self.peek(1)
            if v not in KEYWORDS and nxt.kind == "id" and nxt.val not in KEYWORDS and nxt.sp:
                self.next()
                words = []
                while self.peek().kind not in ("nl", "eof") and not self.is_op(";") and not self.is_op(","):
                    words.append(str(self.next().val))
                self.end_stmt()
                return ("command", v, words, line)
        # multi-assignment  [a, b, ~] = f(...)
        if self.is_op("["):
            save = self.p
            lhs = self.try_parse_multi_lhs()
            if lhs is not None:
                rhs = self.parse_expr()
                self.end_stmt()
                return ("assign", lhs, rhs, line)
            self.p = save
        e = self.parse_expr()
        if self.is_op("="):
            self.next()
            if e[0] not in ("id", "index", "field"):
                raise MlabError(f"{self.filename}:{line}: cannot assign to this expression")
            rhs = self.parse_expr()
            self.end_stmt()
            return ("assign", [e], rhs, line)
        self.end_stmt()
        return ("expr", e, line)

    def try_parse_multi_lhs(self):
        # scan to the matching ']' and require '=' (not '==') right after it
        depth, q = 0, self.p
        while True:
            tok = self.t[q]
            if tok.kind in ("nl", "eof"):
                return None
            if tok.kind == "op" and tok.val in "([{":
                depth += 1
            elif tok.kind == "op" and tok.val in ")]}":
                depth -= 1
                if depth == 0:
                    break
            q += 1
        after = self.t[q + 1]
        if not (after.kind == "op" and after.val == "="):
            return None
        self.next()  # [
        lhs = []
        while not self.is_op("]"):
            if self.is_op(","):
                self.next()
                continue
            if self.is_op("~"):
                self.next()
                lhs.append(("tilde",))
                continue
            save_m = self.in_matrix
            self.in_matrix = 0
            lhs.append(self.parse_postfix())
            self.in_matrix = save_m
        self.next()  # ]
        self.expect_op("=")
        return lhs

    def parse_if(self):
        line = self.next().line  # if
        clauses = []
        cond = self.parse_expr()
        if self.is_op(","):
            self.next()
        body = self.parse_block(("elseif", "else", "end"))
        clauses.append((cond, body))
        other = None
        while True:
            if self.is_kw("elseif"):
                self.next()
                cond = self.parse_expr()
                if self.is_op(","):
                    self.next()
                clauses.append((cond, self.parse_block(("elseif", "else", "end"))))
            elif self.is_kw("else"):
                self.next()
                other = self.parse_block(("end",))
            elif self.is_kw("end"):
                self.next()
                break
            else:
                raise MlabError(f"{self.filename}:{self.peek().line}: unterminated if")
        return ("if", clauses, other, line)

    # -- expressions (MATLAB precedence, lowest first) ----------------------------
    def parse_expr(self):
        return self.parse_oror()

    def parse_oror(self):
        a = self.parse_andand()
        while self.is_op("||"):
            self.next()
            a = ("oror", a, self.parse_andand())
        return a

    def parse_andand(self):
        a = self.parse_or()
        while self.is_op("&&"):
            self.next()
            a = ("andand", a, self.parse_or())
        return a

    def parse_or(self):
        a = self.parse_and()
        while self.is_op("|"):
            self.next()
            a = ("bin", "|", a, self.parse_and())
        return a

    def parse_and(self):
        a = self.parse_cmp()
        while self.is_op("&"):
            self.next()
            a = ("bin", "&", a, self.parse_cmp())
        return a

    def parse_cmp(self):
        a = self.parse_range()
        while self.peek().kind == "op" and self.peek().val in ("==", "~=", "<", "<=", ">", ">="):
            if self.matrix_sep_here():
                break
            op = self.next().val
            a = ("bin", op, a, self.parse_range())
        return a

    def parse_range(self):
        a = self.parse_add()
        if self.is_op(":") and not self.range_colon_is_arg_end():
            self.next()
            b = self.parse_add()
            if self.is_op(":") and not self.range_colon_is_arg_end():
                self.next()
                c = self.parse_add()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def range_colon_is_arg_end(self):
        nxt = self.peek(1)
        return nxt.kind == "op" and nxt.val in (")", ",")

    def matrix_sep_here(self):
        """inside [ ]: `a -b` (blank before, none after a sign) starts a new element"""
        if not self.in_matrix:
            return False
        tok, nxt = self.peek(), self.peek(1)
        return tok.sp and not nxt.sp and tok.val in ("+", "-")

    def parse_add(self):
        a = self.parse_mul()
        while self.peek().kind == "op" and self.peek().val in ("+", "-"):
            if self.matrix_sep_here():
                break
            op = self.next().val
            a = ("bin", op, a, self.parse_mul())
        return a

    def parse_mul(self):
        a = self.parse_unary()
        while self.peek().kind == "op" and self.peek().val in ("*", "/", "\\", ".*", "./", ".\\"):
            op = self.next().val
            a = ("bin", op, a, self.parse_unary())
        return a

    def parse_unary(self):
        if self.peek().kind == "op" and self.peek().val in ("+", "-", "~"):
            op = self.next().val
            return ("un", op, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        a = self.parse_postfix()
        while self.peek().kind == "op" and self.peek().val in ("^", ".^"):
            op = self.next().val
            # the exponent may carry its own unary sign: 2^-1
            if self.peek().kind == "op" and self.peek().val in ("+", "-", "~"):
                u = self.next().val
                b = ("un", u, self.parse_postfix_power_operand())
            else:
                b = self.parse_postfix()
            a = ("bin", op, a, b)
        return a

    def parse_postfix_power_operand(self):
        return self.parse_postfix()

    def parse_postfix(self):
        a = self.parse_primary()
        while True:
            tok = self.peek()
            if tok.kind != "op":
                break
            if tok.val in ("(", "{"):
                if self.in_matrix and tok.sp:
                    break  # [a (1)] : two elements
                close = ")" if tok.val == "(" else "}"
                self.next()
                save_m, self.in_matrix = self.in_matrix, 0
                self.in_index += 1
                args = []
                while not self.is_op(close):
                    if self.is_op(","):
                        self.next()
                        continue
                    if self.is_op(":") and self.range_colon_is_arg_end():
                        self.next()
                        args.append(("colon",))
                    else:
                        args.append(self.parse_expr())
                self.next()
                self.in_index -= 1
                self.in_matrix = save_m
                a = ("index", a, args, tok.val)
            elif tok.val in ("'", ".'"):
                self.next()
                a = ("transpose", a)
            elif tok.val == "." and self.peek(1).kind == "id" and not tok.sp:
                self.next()
                a = ("field", a, self.next().val)
            elif tok.val == "." and self.is_op("(", 1) and not tok.sp:  # dynamic field  s.(expr)
                self.next()
                self.next()
                save_m, self.in_matrix = self.in_matrix, 0
                save_i, self.in_index = self.in_index, 0
                e = self.parse_expr()
                self.expect_op(")")
                self.in_matrix, self.in_index = save_m, save_i
                a = ("dynfield", a, e)
            else:
                break
        return a

    def parse_primary(self):
        tok = self.next()
        if tok.kind == "num":
            return ("num", tok.val)
        if tok.kind == "str":
            return ("str", tok.val)
        if tok.kind == "id":
            if tok.val == "end":
                if self.in_index:
                    return ("end",)
                raise MlabError(f"{self.filename}:{tok.line}: unexpected 'end'")
            if tok.val in KEYWORDS:
                raise MlabError(f"{self.filename}:{tok.line}: unexpected keyword {tok.val!r}")
            return ("id", tok.val)
        if tok.kind == "op":
            if tok.val == "(":
                save_m, self.in_matrix = self.in_matrix, 0
                save_i, self.in_index = self.in_index, self.in_index  # `end` keeps its meaning
                e = self.parse_expr()
                self.expect_op(")")
                self.in_matrix = save_m
                self.in_index = save_i
                return ("paren", e)
            if tok.val in ("[", "{"):
                close = "]" if tok.val == "[" else "}"
                self.in_matrix += 1
                rows, row = [], []
                while True:
                    t2 = self.peek()
                    if t2.kind == "op" and t2.val == close:
                        self.next()
                        break
                    if t2.kind == "nl" or (t2.kind == "op" and t2.val == ";"):
                        self.next()
                        if row:
                            rows.append(row)
                            row = []
                        continue
                    if t2.kind == "op" and t2.val == ",":
                        self.next()
                        continue
                    if t2.kind == "eof":
                        raise MlabError(f"{self.filename}:{tok.line}: unterminated bracket")
                    row.append(self.parse_expr())
                if row:
                    rows.append(row)
                self.in_matrix -= 1
                return ("matrix" if tok.val == "[" else "cellarr", rows)
            if tok.val == "@":
                if self.is_op("("):
                    self.next()
                    params = []
                    while not self.is_op(")"):
                        t2 = self.next()
                        if t2.kind == "id":
                            params.append(t2.val)
                        elif t2.kind == "op" and t2.val == "~":
                            params.append("~")
                    self.next()
                    save_m, self.in_matrix = self.in_matrix, 0
                    save_i, self.in_index = self.in_index, 0
                    body = self.parse_expr()
                    self.in_matrix, self.in_index = save_m, save_i
                    return ("anon", params, body)
                return ("fhandle", self.next().val)
            if tok.val == ":":
                return ("colon",)
        raise MlabError(f"{self.filename}:{tok.line}: unexpected token {tok.val!r}")


# ======================================================================================
# values
# ======================================================================================
class Cell:
    def __init__(self, r, c):
        self.a = np.empty((r, c), dtype=object)
        for i in range(r):
            for j in range(c):
                self.a[i, j] = np.zeros((0, 0))

    @property
    def shape(self):
        return self.a.shape


class FuncHandle:
    def __init__(self, fn, label):
        self.fn, self.label = fn, label

    def __call__(self, *args, nargout=1):
        return self.fn(list(args), nargout)


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


def _is_num(v):
    return isinstance(v, np.ndarray) or sp.issparse(v)


def mat(v):
    """anything numeric -> 2-D ndarray (or sparse matrix left alone)"""
    if sp.issparse(v):
        return v
    if isinstance(v, np.ndarray):
        if v.ndim == 2:
            return v
        if v.ndim == 0:
            return v.reshape(1, 1)
        if v.ndim == 1:
            return v.reshape(-1, 1)
        raise MlabError("N-d arrays are not supported")
    if isinstance(v, (bool, np.bool_)):
        return np.array([[bool(v)]])
    if isinstance(v, (int, float, np.integer, np.floating, complex, np.complexfloating)):
        return np.array([[v]], dtype=complex if isinstance(v, (complex, np.complexfloating)) else float)
    if isinstance(v, str):
        return np.array([[float(ord(ch)) for ch in v]]) if v else np.zeros((0, 0))
    raise MlabError(f"not a numeric value: {type(v).__name__}")


def numeric(v):
    """logical -> double for arithmetic"""
    v = mat(v)
    if not sp.issparse(v) and v.dtype == bool:
        return v.astype(float)
    return v


def is_scalar(v):
    return v.shape == (1, 1)


def scalar(v):
    v = mat(v)
    if sp.issparse(v):
        v = v.toarray()
    if v.size != 1:
        raise MlabError(f"expected a scalar, got {v.shape}")
    x = v.reshape(-1)[0]
    if np.iscomplexobj(x) and x.imag == 0:
        x = x.real
    return x


def truth(v):
    if isinstance(v, str):
        return len(v) > 0
    v = mat(v)
    if sp.issparse(v):
        v = v.toarray()
    return v.size > 0 and bool(np.all(v != 0))


def dense(v):
    return v.toarray() if sp.issparse(v) else v


# ======================================================================================
# interpreter
# ======================================================================================
class Interp:
    """``Interp(path=["/root/reference"]).call("hybrid_ba_gmres_rtp", [A,B,b,xt,tol,maxit,lam], nargout=4)``

    ``extra`` maps names to Python callables ``f(args, nargout) -> tuple`` standing in for
    functions that are neither MATLAB built-ins nor in the reference (Hansen's ``shaw``/``heat``/
    ``deriv2``: third-party and not vendored)."""

    def __init__(self, path=(), extra=None, hooks=None):
        self.path = list(path)
        self.files = {}     # name -> Function (primary function of name.m)
        self.extra = dict(extra or {})
        self.hooks = hooks or {}
        self.builtins = _make_builtins(self)
        self.calls = []     # trace of user-function calls (name, filename)
        self.last_ws = {}   # function name -> workspace at the end of its last call (H, Q, beta ...)
        self.out = []       # text written by fprintf / disp

    # -- loading ------------------------------------------------------------------
    def load_source(self, src, filename="<string>"):
        funcs, script = Parser(tokenize(src), filename).parse_file()
        table = {f.name: f for f in funcs}
        for f in funcs:
            f.locals = table
        return funcs, script

    def find_function(self, name, scope=None):
        if scope is not None and name in scope:
            return scope[name]
        if name in self.files:
            return self.files[name]
        for d in self.path:
            fn = os.path.join(d, name + ".m")
            if os.path.isfile(fn):
                with open(fn, "r", newline="") as fh:
                    funcs, script = self.load_source(fh.read(), fn)
                if not funcs:
                    raise MlabError(f"{fn}: script files cannot be called as functions")
                self.files[name] = funcs[0]
                return funcs[0]
        return None

    def local_function(self, filename_stem, name):
        """a local function of another file (e.g. compute_gcv_surface inside plot_gcv_surface.m)"""
        main = self.find_function(filename_stem)
        if main is None or name not in main.locals:
            raise MlabError(f"no local function {name} in {filename_stem}.m")
        return main.locals[name]

    def run_source(self, src, ws=None, filename="<script>"):
        """execute script text (e.g. a line range of one of the reference's driver scripts) in `ws`"""
        funcs, script = self.load_source(src, filename)
        ws = {} if ws is None else ws
        fr = {"ws": ws, "nargin": 0, "nargout": 0, "scope": {f.name: f for f in funcs}, "ends": [], "fn": None}
        self.exec_block(script, fr)
        return ws

    def run_lines(self, filename, first, last, ws=None):
        """execute lines first..last (1-based, inclusive) of a reference file as a script"""
        with open(filename, "r", newline="") as fh:
            lines = fh.read().replace("\r\n", "\n").split("\n")
        return self.run_source("\n".join(lines[first - 1:last]) + "\n", ws, f"{filename}:{first}-{last}")

    # -- calling ------------------------------------------------------------------
    def call(self, name, args, nargout=1, scope=None):
        f = name if isinstance(name, Function) else self.find_function(name, scope)
        if f is not None:
            return self.call_user(f, list(args), nargout)
        if name in self.extra:
            out = self.extra[name](list(args), nargout)
            return tuple(out) if isinstance(out, (tuple, list)) else (out,)
        if name in self.builtins:
            out = self.builtins[name](list(args), nargout)
            return out if isinstance(out, tuple) else (out,)
        raise MlabError(f"undefined function or variable {name!r}")

    def call_user(self, f, args, nargout):
        if len(args) > len(f.params):
            raise MlabError(f"{f.name}: too many input arguments")
        ws = {}
        for pname, a in zip(f.params, args):
            if pname != "~":
                ws[pname] = self._share(a)
        frame = {"ws": ws, "nargin": len(args), "nargout": nargout, "scope": f.locals, "ends": [], "fn": f}
        self.calls.append((f.name, f.filename))
        try:
            self.exec_block(f.body, frame)
        except _Return:
            pass
        self.last_ws[f.name] = ws
        outs = []
        for i, o in enumerate(f.outs[: max(nargout, 1)]):
            if o not in ws:
                if i < nargout or (nargout == 0 and i == 0 and False):
                    raise MlabError(f"{f.name}: output argument {o!r} not assigned (MATLAB raises the same error)")
                break
            outs.append(ws[o])
        return tuple(outs)

    @staticmethod
    def _share(v):
        if isinstance(v, np.ndarray):
            v.flags.writeable = False
        return v

    # -- statements ---------------------------------------------------------------
    def exec_block(self, stmts, fr):
        for s in stmts:
            self.exec_stmt(s, fr)

    def exec_stmt(self, s, fr):
        kind = s[0]
        try:
            if kind == "assign":
                self.exec_assign(s[1], s[2], fr)
            elif kind == "expr":
                e = s[1]
                # a bare call statement may return nothing
                if e[0] in ("id", "index"):
                    self.eval_multi(e, fr, 0)
                else:
                    self.eval(e, fr)
            elif kind == "if":
                for cond, body in s[1]:
                    if truth(self.eval(cond, fr)):
                        self.exec_block(body, fr)
                        break
                else:
                    if s[2] is not None:
                        self.exec_block(s[2], fr)
            elif kind == "for":
                rng = self.eval(s[2], fr)
                rng = dense(mat(rng)) if not isinstance(rng, Cell) else rng
                ncols = rng.shape[1] if rng.shape[0] > 0 else 0
                if ncols == 0:
                    fr["ws"][s[1]] = np.zeros((0, 0))
                for c in range(ncols):
                    col = rng[:, c:c + 1].copy()
                    fr["ws"][s[1]] = col
                    try:
                        self.exec_block(s[3], fr)
                    except _Break:
                        break
                    except _Continue:
                        continue
            elif kind == "while":
                while truth(self.eval(s[1], fr)):
                    try:
                        self.exec_block(s[2], fr)
                    except _Break:
                        break
                    except _Continue:
                        continue
            elif kind == "switch":
                v = self.eval(s[1], fr)
                done = False
                for ce, body in s[2]:
                    cv = self.eval(ce, fr)
                    cands = [cv] if not isinstance(cv, Cell) else list(cv.a.reshape(-1))
                    for c in cands:
                        if (isinstance(v, str) and isinstance(c, str) and v == c) or \
                                (not isinstance(v, str) and not isinstance(c, str) and scalar(v) == scalar(c)):
                            self.exec_block(body, fr)
                            done = True
                            break
                    if done:
                        break
                if not done and s[3] is not None:
                    self.exec_block(s[3], fr)
            elif kind == "break":
                raise _Break()
            elif kind == "continue":
                raise _Continue()
            elif kind == "return":
                raise _Return()
            elif kind == "command":
                self.exec_command(s[1], s[2], fr)
            else:
                raise MlabError(f"unknown statement {kind}")
        except MlabError as e:
            if not getattr(e, "located", False):
                e.located = True
                e.args = (f"{fr['fn'].filename if fr.get('fn') else '<script>'}:{s[-1]}: {e.args[0]}",)
            raise

    def exec_command(self, name, words, fr):
        if name in ("clear", "clc", "close", "hold", "grid", "format", "warning", "axis", "figure", "drawnow"):
            if name == "clear":
                fr["ws"].clear()
            return
        raise MlabError(f"command syntax {name!r} is not supported")

    def exec_assign(self, lhs, rhs, fr):
        if len(lhs) == 1:
            val = self.eval(rhs, fr)
            self.assign_to(lhs[0], val, fr, rhs)
            return
        vals = self.eval_multi(rhs, fr, len(lhs))
        if len(vals) < len([l for l in lhs if l[0] != "tilde"]) and len(vals) < len(lhs):
            raise MlabError("too many output arguments requested")
        for l, v in zip(lhs, vals):
            if l[0] != "tilde":
                self.assign_to(l, v, fr, None)

    def assign_to(self, target, val, fr, rhs_node):
        ws = fr["ws"]
        if target[0] == "id":
            if isinstance(val, np.ndarray) and any(val is o for o in ws.values()):
                val.flags.writeable = False  # alias: copy on the next indexed write
            ws[target[1]] = val
            if self.hooks and fr.get("fn") is not None:
                h = self.hooks.get((fr["fn"].name, target[1]))
                if h is not None:
                    h(val, ws)
            return
        if target[0] == "index":
            base = target[1]
            if base[0] != "id":
                raise MlabError("nested indexed assignment is not supported")
            name = base[1]
            cur = ws.get(name)
            if target[3] == "{":
                if cur is None:
                    cur = Cell(0, 0)
                if not isinstance(cur, Cell):
                    raise MlabError(f"{name} is not a cell array")
                idx = self.eval_indices(target[2], cur.a, fr)
                self._cell_store(cur, idx, val)
                ws[name] = cur
                return
            if isinstance(cur, Cell):
                raise MlabError("paren-assignment into a cell array is not supported")
            if cur is None:
                cur = np.zeros((0, 0))
            if sp.issparse(cur):
                cur = cur.toarray()
            idx = self.eval_indices(target[2], cur, fr)
            ws[name] = self._store(cur, idx, val)
            return
        raise MlabError("unsupported assignment target")

    @staticmethod
    def _cell_store(cell, idx, val):
        if len(idx) == 1:
            (i,) = idx
            i = np.asarray(i).reshape(-1)
            if i.size != 1:
                raise MlabError("cell brace assignment needs a scalar index")
            k = int(i[0])
            r, c = cell.a.shape
            if r * c <= k:
                # grow as a column unless it is a row
                new = Cell(1, k + 1) if r == 1 and c > 0 else Cell(k + 1, 1)
                new.a.reshape(-1, order="F")[: r * c] = cell.a.reshape(-1, order="F")
                flat_old = cell.a.reshape(-1, order="F")
                for t in range(r * c):
                    new.a[t if new.a.shape[1] == 1 else 0, 0 if new.a.shape[1] == 1 else t] = flat_old[t]
                cell.a = new.a
                r, c = cell.a.shape
            cell.a[k % r, k // r] = val
        else:
            i, j = (int(np.asarray(t).reshape(-1)[0]) for t in idx)
            cell.a[i, j] = val

    @staticmethod
    def _store(cur, idx, val):
        if isinstance(val, str):
            val = mat(val)
        val = dense(numeric(val)) if _is_num(val) or not isinstance(val, Cell) else val
        if val.dtype == complex and cur.dtype != complex:
            cur = cur.astype(complex)
        elif cur.dtype == bool and val.dtype != bool:
            cur = cur.astype(float)
        elif not cur.flags.writeable:
            cur = cur.copy()
        if cur.dtype not in (float, complex):
            cur = cur.astype(float)
        if len(idx) == 1:
            (i,) = idx
            if isinstance(i, np.ndarray) and i.dtype == bool:
                i = np.flatnonzero(i.reshape(-1, order="F"))
            i = np.asarray(i).reshape(-1)
            need = int(i.max()) + 1 if i.size else 0
            if need > cur.size:
                if cur.size == 0:
                    cur = np.zeros((1, need))
                elif cur.shape[0] == 1:
                    cur = np.concatenate([cur, np.zeros((1, need - cur.shape[1]))], axis=1)
                elif cur.shape[1] == 1:
                    cur = np.concatenate([cur, np.zeros((need - cur.shape[0], 1))], axis=0)
                else:
                    raise MlabError("linear index out of range in a matrix assignment")
            flat = cur.reshape(-1, order="F").copy() if not cur.flags.f_contiguous else None
            if val.size == 0 and i.size:
                raise MlabError("deleting elements with x(i) = [] is not supported")
            v = val.reshape(-1, order="F")
            if v.size != 1 and v.size != i.size:
                raise MlabError(f"assignment size mismatch: {i.size} elements <- {v.size}")
            r = cur.shape[0]
            cur[i % r, i // r] = v if v.size != 1 else v[0]
            return cur
        i, j = idx

        def norm_idx(t, dimlen):
            if isinstance(t, np.ndarray) and t.dtype == bool:
                return np.flatnonzero(t.reshape(-1, order="F"))
            return np.asarray(t).reshape(-1)

        i, j = norm_idx(i, cur.shape[0]), norm_idx(j, cur.shape[1])
        nr = max(cur.shape[0], int(i.max()) + 1 if i.size else 0)
        nc = max(cur.shape[1], int(j.max()) + 1 if j.size else 0)
        if (nr, nc) != cur.shape:
            big = np.zeros((nr, nc), dtype=cur.dtype)
            big[: cur.shape[0], : cur.shape[1]] = cur
            cur = big
        if val.size == 1:
            cur[np.ix_(i, j)] = val.reshape(-1)[0]
        else:
            if val.shape != (i.size, j.size):
                if val.size == i.size * j.size and (val.shape[0] == 1 or val.shape[1] == 1) and \
                        (i.size == 1 or j.size == 1):
                    val = val.reshape(i.size, j.size)
                else:
                    raise MlabError(f"assignment size mismatch: {(i.size, j.size)} <- {val.shape}")
            cur[np.ix_(i, j)] = val
        return cur

    # -- expressions --------------------------------------------------------------
    def eval_multi(self, node, fr, nargout):
        """evaluate a call-like node asking for `nargout` outputs; returns a tuple"""
        if node[0] == "paren":
            return (self.eval(node[1], fr),)
        if node[0] == "id":
            name = node[1]
            if name in fr["ws"]:
                return (fr["ws"][name],)
            return self.call_named(name, [], nargout, fr)
        if node[0] == "index" and node[3] == "(" and node[1][0] == "id" and node[1][1] not in fr["ws"]:
            name = node[1][1]
            args = [self.eval_arg(a, fr) for a in node[2]]
            return self.call_named(name, args, nargout, fr)
        if node[0] == "index" and node[3] == "(":
            base = self.eval(node[1], fr)
            if isinstance(base, FuncHandle):
                args = [self.eval_arg(a, fr) for a in node[2]]
                out = base.fn(args, nargout)
                return out if isinstance(out, tuple) else (out,)
        return (self.eval(node, fr),)

    def eval_arg(self, a, fr):
        if a[0] == "colon":
            return ":"
        return self.eval(a, fr)

    def call_named(self, name, args, nargout, fr):
        if name == "nargin":
            return (np.array([[float(fr["nargin"])]]),)
        if name == "nargout":
            return (np.array([[float(fr["nargout"])]]),)
        return self.call(name, args, nargout, fr.get("scope"))

    def eval(self, node, fr):
        k = node[0]
        if k == "num":
            return np.array([[node[1]]])
        if k == "str":
            return node[1]
        if k == "paren":
            return self.eval(node[1], fr)
        if k == "id":
            name = node[1]
            ws = fr["ws"]
            if name in ws:
                return ws[name]
            out = self.call_named(name, [], 1, fr)
            if not out:
                raise MlabError(f"{name}: no value returned")
            return out[0]
        if k == "index":
            return self.eval_index(node, fr)
        if k == "bin":
            return self.binop(node[1], self.eval(node[2], fr), self.eval(node[3], fr))
        if k == "un":
            v = self.eval(node[2], fr)
            if node[1] == "-":
                return -numeric(v)
            if node[1] == "+":
                return numeric(v)
            return ~(dense(mat(v)) != 0)
        if k == "transpose":
            v = self.eval(node[1], fr)
            if isinstance(v, str):
                raise MlabError("transpose of a string")
            v = mat(v)
            if sp.issparse(v):
                return v.T.tocsc()
            return np.conj(v.T) if np.iscomplexobj(v) else v.T.copy()
        if k == "range":
            a = scalar(self.eval(node[1], fr))
            b = scalar(self.eval(node[3], fr))
            st = 1.0 if node[2] is None else scalar(self.eval(node[2], fr))
            if st == 0 or (st > 0 and a > b) or (st < 0 and a < b):
                return np.zeros((1, 0))
            cnt = int(math.floor((b - a) / st * (1 + 2 * np.finfo(float).eps))) + 1
            return (a + st * np.arange(cnt, dtype=float)).reshape(1, -1)
        if k == "andand":
            return np.array([[truth(self.eval(node[1], fr)) and truth(self.eval(node[2], fr))]])
        if k == "oror":
            return np.array([[truth(self.eval(node[1], fr)) or truth(self.eval(node[2], fr))]])
        if k == "matrix":
            return self.build_matrix(node[1], fr)
        if k == "cellarr":
            rows = node[1]
            c = Cell(len(rows), max((len(r) for r in rows), default=0))
            for i, r in enumerate(rows):
                for j, e in enumerate(r):
                    c.a[i, j] = self.eval(e, fr)
            return c
        if k == "anon":
            params, body = node[1], node[2]
            captured = {n: self._share(v) for n, v in fr["ws"].items()}
            outer = fr

            def fn(args, nargout, _p=params, _b=body, _c=captured):
                ws = dict(_c)
                for pn, a in zip(_p, args):
                    if pn != "~":
                        ws[pn] = self._share(a)
                sub = {"ws": ws, "nargin": len(args), "nargout": nargout, "scope": outer.get("scope"),
                       "ends": [], "fn": outer.get("fn")}
                return self.eval_multi(_b, sub, nargout) if nargout > 1 else (self.eval(_b, sub),)
            return FuncHandle(fn, "@anon")
        if k == "fhandle":
            name = node[1]
            scope = fr.get("scope")
            return FuncHandle(lambda args, nargout: self.call(name, args, nargout, scope), "@" + name)
        if k == "end":
            if not fr["ends"]:
                raise MlabError("'end' outside of an index expression")
            arr, pos, npos = fr["ends"][-1]
            shp = arr.shape
            if npos == 1:
                return np.array([[float(shp[0] * shp[1])]])
            return np.array([[float(shp[pos])]])
        if k == "colon":
            return ":"
        if k in ("field", "dynfield"):
            base = self.eval(node[1], fr)
            name = node[2] if k == "field" else self.eval(node[2], fr)
            if isinstance(base, dict) and isinstance(name, str) and name in base:
                return base[name]
            raise MlabError(f"reference to non-existent field {name!r}")
        raise MlabError(f"cannot evaluate node {k}")

    def build_matrix(self, rows, fr):
        out_rows = []
        for r in rows:
            vals = [self.eval(e, fr) for e in r]
            if vals and all(isinstance(v, str) for v in vals):
                out_rows.append("".join(vals))
                continue
            vals = [mat(v) if isinstance(v, str) else numeric(v) for v in vals]
            vals = [v for v in vals if v.shape != (0, 0)]
            if not vals:
                continue
            if any(sp.issparse(v) for v in vals):
                out_rows.append(sp.hstack([sp.csc_matrix(v) for v in vals], format="csc"))
            else:
                h = {v.shape[0] for v in vals}
                if len(h) != 1:
                    raise MlabError(f"horizontal concatenation: inconsistent row counts {sorted(h)}")
                out_rows.append(np.concatenate(vals, axis=1))
        if not out_rows:
            return np.zeros((0, 0))
        if all(isinstance(r, str) for r in out_rows):
            if len(out_rows) == 1:
                return out_rows[0]
            raise MlabError("multi-row char arrays are not supported")
        out_rows = [mat(r) if isinstance(r, str) else r for r in out_rows]
        if len(out_rows) == 1:
            return out_rows[0]
        if any(sp.issparse(r) for r in out_rows):
            return sp.vstack([sp.csc_matrix(r) for r in out_rows], format="csc")
        w = {r.shape[1] for r in out_rows}
        if len(w) != 1:
            raise MlabError(f"vertical concatenation: inconsistent column counts {sorted(w)}")
        return np.concatenate(out_rows, axis=0)

    # -- indexing -----------------------------------------------------------------
    def eval_indices(self, arg_nodes, arr, fr):
        """-> list of 0-based index arrays (or bool masks), one per subscript"""
        n = len(arg_nodes)
        if n not in (1, 2):
            raise MlabError(f"{n}-subscript indexing is not supported")
        out = []
        for pos, a in enumerate(arg_nodes):
            if a[0] == "colon":
                dimlen = arr.shape[0] * arr.shape[1] if n == 1 else arr.shape[pos]
                out.append(np.arange(dimlen))
                continue
            fr["ends"].append((arr, pos, n))
            try:
                v = self.eval(a, fr)
            finally:
                fr["ends"].pop()
            if isinstance(v, str) and v == ":":
                dimlen = arr.shape[0] * arr.shape[1] if n == 1 else arr.shape[pos]
                out.append(np.arange(dimlen))
                continue
            v = dense(mat(v))
            if v.dtype == bool:
                out.append(v)
                continue
            vi = np.rint(v.real if np.iscomplexobj(v) else v).astype(np.int64)
            if np.any(np.abs(v - vi) > 0):
                raise MlabError("subscript indices must be integers")
            if np.any(vi < 1):
                raise MlabError("index must be a positive integer (MATLAB raises the same error)")
            idx = vi - 1
            idx_arr = idx  # keep the 2-D shape: x(rowvector) is a row
            out.append(idx_arr)
        return out

    def eval_index(self, node, fr):
        base_node, arg_nodes, br = node[1], node[2], node[3]
        # function call?
        if base_node[0] == "id" and base_node[1] not in fr["ws"] and br == "(":
            out = self.eval_multi(node, fr, 1)
            if not out:
                raise MlabError(f"{base_node[1]}: no value returned")
            return out[0]
        base = self.eval(base_node, fr)
        if isinstance(base, FuncHandle):
            args = [self.eval_arg(a, fr) for a in arg_nodes]
            out = base.fn(args, 1)
            out = out if isinstance(out, tuple) else (out,)
            return out[0]
        if isinstance(base, Cell):
            idx = self.eval_indices(arg_nodes, base.a, fr)
            if br == "{":
                if len(idx) == 1:
                    k = int(np.asarray(idx[0]).reshape(-1)[0])
                    r = base.a.shape[0]
                    return base.a[k % r, k // r]
                return base.a[int(np.asarray(idx[0]).reshape(-1)[0]), int(np.asarray(idx[1]).reshape(-1)[0])]
            sub = self._take(base.a, idx)
            c = Cell(0, 0)
            c.a = sub
            return c
        if br == "{":
            raise MlabError("brace indexing of a non-cell value")
        if isinstance(base, str):
            arr = mat(base)
            idx = self.eval_indices(arg_nodes, arr, fr)
            sub = self._take(arr, idx)
            return "".join(chr(int(c)) for c in sub.reshape(-1))
        arr = mat(base)
        if len(arg_nodes) == 1 and arg_nodes[0][0] == "colon":  # x(:) is always a column
            return dense(arr).reshape(-1, 1, order="F").copy()
        idx = self.eval_indices(arg_nodes, arr, fr)
        return self._take(arr, idx)

    @staticmethod
    def _take(arr, idx):
        if sp.issparse(arr):
            if len(idx) == 2:
                i, j = (np.flatnonzero(t.reshape(-1, order="F")) if t.dtype == bool else t.reshape(-1) for t in idx)
                return arr.tocsr()[i][:, j].tocsc()
            arr = arr.toarray()
        if len(idx) == 1:
            (i,) = idx
            if i.dtype == bool:
                flat = arr.reshape(-1, order="F")
                sel = flat[: i.size][i.reshape(-1, order="F")] if i.size <= flat.size else None
                if sel is None:
                    raise MlabError("logical index too long")
                return sel.reshape(-1, 1) if arr.shape[1] == 1 or arr.shape[0] != 1 else sel.reshape(1, -1)
            flat = arr.reshape(-1, order="F")
            if i.size and int(i.max()) >= flat.size:
                raise MlabError("index exceeds the number of array elements (MATLAB raises the same error)")
            sel = flat[i.reshape(-1, order="F")]
            if i.ndim == 2 and i.shape[0] != 1 and i.shape[1] != 1:
                return sel.reshape(i.shape, order="F")
            # vector index: result has the orientation of the index, except that indexing a
            # vector with a vector keeps the orientation of the source
            if arr.shape[0] == 1 and arr.shape[1] != 1:
                return sel.reshape(1, -1)
            if arr.shape[1] == 1 and arr.shape[0] != 1:
                return sel.reshape(-1, 1)
            return sel.reshape(i.shape if i.ndim == 2 else (-1, 1), order="F")
        i, j = idx
        i = np.flatnonzero(i.reshape(-1, order="F")) if i.dtype == bool else i.reshape(-1)
        j = np.flatnonzero(j.reshape(-1, order="F")) if j.dtype == bool else j.reshape(-1)
        if (i.size and int(i.max()) >= arr.shape[0]) or (j.size and int(j.max()) >= arr.shape[1]):
            raise MlabError("index exceeds array bounds (MATLAB raises the same error)")
        return arr[np.ix_(i, j)]

    # -- operators ----------------------------------------------------------------
    def binop(self, op, a, b):
        if isinstance(a, str) or isinstance(b, str):
            if op in ("==", "~=") and isinstance(a, str) and isinstance(b, str) and len(a) == len(b):
                eq = np.array([[x == y for x, y in zip(a, b)]])
                return eq if op == "==" else ~eq
            a = mat(a) if isinstance(a, str) else a
            b = mat(b) if isinstance(b, str) else b
        a, b = numeric(a), numeric(b)
        sa, sb = sp.issparse(a), sp.issparse(b)
        if op == "*":
            if is_scalar(a) or is_scalar(b):
                return self.elementwise("*", a, b)
            if a.shape[1] != b.shape[0]:
                raise MlabError(f"matrix multiply: inner dimensions {a.shape} * {b.shape}")
            r = a @ b
            if sp.issparse(r) and not (sa and sb):
                r = r.toarray()
            return np.asarray(r) if not sp.issparse(r) else r
        if op == "/":
            if is_scalar(b):
                return self.elementwise("/", a, b)
            # mrdivide: a/b = (b'\a')'
            return mldivide(dense(b).T, dense(a).T).T
        if op == "\\":
            if is_scalar(a):
                return self.elementwise("/", b, a)
            return mldivide(a, b)
        if op == "^":
            if is_scalar(a) and is_scalar(b):
                return np.array([[_pow(scalar(a), scalar(b))]])
            if is_scalar(b) and a.shape[0] == a.shape[1]:
                p = scalar(b)
                if p == int(p) and p >= 0:
                    m = dense(a)
                    return np.linalg.matrix_power(m, int(p))
            raise MlabError("matrix power with these operands is not supported")
        if op in (".*", "./", ".\\", ".^", "+", "-", "==", "~=", "<", "<=", ">", ">=", "&", "|"):
            return self.elementwise(op, a, b)
        raise MlabError(f"operator {op!r} is not supported")

    @staticmethod
    def elementwise(op, a, b):
        sa, sb = sp.issparse(a), sp.issparse(b)
        if sa or sb:
            # keep sparse results only where MATLAB would and it is cheap to do so
            if op in ("*", ".*") and (is_scalar(a) or is_scalar(b)):
                s, m_ = (a, b) if is_scalar(a) and not sa else (b, a)
                return (m_ * scalar(s)).tocsc()
            if op in ("/", "./") and sa and is_scalar(b):
                return (a / scalar(b)).tocsc()
            if op in ("+", "-") and sa and sb:
                return (a + b if op == "+" else a - b).tocsc()
            a, b = dense(a), dense(b)
        if a.shape != b.shape and not (is_scalar(a) or is_scalar(b)):
            ok = all(x == y or x == 1 or y == 1 for x, y in zip(a.shape, b.shape))
            if not ok:
                raise MlabError(f"arrays have incompatible sizes {a.shape} and {b.shape} for {op}")
        with np.errstate(all="ignore"):
            if op in ("*", ".*"):
                return a * b
            if op in ("/", "./"):
                return a / b
            if op == ".\\":
                return b / a
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == ".^":
                if np.iscomplexobj(a) or np.iscomplexobj(b) or (np.any(a < 0) and np.any(b != np.rint(b))):
                    return np.power(a.astype(complex), b)
                return np.power(a, b)
            if op == "==":
                return a == b
            if op == "~=":
                return a != b
            if op == "<":
                return a < b
            if op == "<=":
                return a <= b
            if op == ">":
                return a > b
            if op == ">=":
                return a >= b
            if op == "&":
                return (a != 0) & (b != 0)
            if op == "|":
                return (a != 0) | (b != 0)
        raise MlabError(f"operator {op!r}")


def _pow(a, b):
    with np.errstate(all="ignore"):
        if a < 0 and b != int(b):
            return complex(a) ** b
        return float(a) ** float(b) if not isinstance(a, complex) else a ** b


def mldivide(A, B):
    """MATLAB ``A\\B`` following the documented dense algorithm choice."""
    A_sp = sp.issparse(A)
    B = dense(numeric(B))
    if A.shape[0] != B.shape[0]:
        raise MlabError(f"mldivide: {A.shape} \\ {B.shape}")
    if A_sp:
        if A.shape[0] == A.shape[1]:
            import scipy.sparse.linalg as spla
            return np.asarray(spla.spsolve(A.tocsc(), B)).reshape(B.shape)
        A = A.toarray()
    A = np.asarray(A)
    m, n = A.shape
    with np.errstate(all="ignore"):
        if m != n:
            y, *_ = sla.lstsq(A, B, lapack_driver="gelsy", check_finite=False)
            return y
        if n == 0:
            return np.zeros((0, B.shape[1]))
        if not np.all(np.isfinite(A)) or not np.all(np.isfinite(B)):
            return np.full(B.shape, np.nan)
        if np.all(np.tril(A, -1) == 0):
            return _tri(A, B, lower=False)
        if np.all(np.triu(A, 1) == 0):
            return _tri(A, B, lower=True)
        if np.array_equal(A, A.T) and np.all(np.diag(A) > 0):
            try:
                c = sla.cho_factor(A, lower=False, check_finite=False)
                return sla.cho_solve(c, B, check_finite=False)
            except sla.LinAlgError:
                pass
        try:
            lu = sla.lu_factor(A, check_finite=False)
            return sla.lu_solve(lu, B, check_finite=False)
        except (sla.LinAlgError, ValueError):
            return np.full(B.shape, np.inf)


def _tri(A, B, lower):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if np.any(np.diag(A) == 0):
            with np.errstate(all="ignore"):
                # MATLAB warns "matrix is singular" and divides through (Inf/NaN results)
                n = A.shape[0]
                X = np.zeros_like(B, dtype=float)
                rng = range(n) if lower else range(n - 1, -1, -1)
                for i in rng:
                    s = B[i] - (A[i, :i] @ X[:i] if lower else A[i, i + 1:] @ X[i + 1:])
                    X[i] = s / A[i, i]
                return X
        return sla.solve_triangular(A, B, lower=lower, check_finite=False)


# ======================================================================================
# built-ins
# ======================================================================================
def _make_builtins(ip: Interp):
    B = {}

    def reg(name):
        def deco(f):
            B[name] = f
            return f
        return deco

    def dims(args):
        if not args:
            return (1, 1)
        if len(args) == 1:
            v = dense(mat(args[0]))
            if v.size == 1:
                k = int(scalar(v))
                return (k, k)
            v = v.reshape(-1)
            if v.size != 2:
                raise MlabError("N-d sizes are not supported")
            return (int(v[0]), int(v[1]))
        return (int(scalar(args[0])), int(scalar(args[1])))

    def shape_args(args):
        args = [a for a in args if not isinstance(a, str)]
        r, c = dims(args)
        return max(r, 0), max(c, 0)

    @reg("zeros")
    def _zeros(args, nargout):
        return np.zeros(shape_args(args))

    @reg("ones")
    def _ones(args, nargout):
        return np.ones(shape_args(args))

    @reg("nan")
    def _nan(args, nargout):
        return np.full(shape_args(args), np.nan)

    B["NaN"] = _nan

    @reg("inf")
    def _inf(args, nargout):
        return np.full(shape_args(args), np.inf)

    B["Inf"] = _inf

    @reg("eye")
    def _eye(args, nargout):
        r, c = shape_args(args)
        return np.eye(r, c)

    @reg("speye")
    def _speye(args, nargout):
        r, c = shape_args(args)
        return sp.eye(r, c, format="csc")

    @reg("cell")
    def _cell(args, nargout):
        r, c = shape_args(args)
        return Cell(r, c)

    @reg("true")
    def _true(args, nargout):
        return np.ones(shape_args(args), dtype=bool)

    @reg("false")
    def _false(args, nargout):
        return np.zeros(shape_args(args), dtype=bool)

    @reg("eps")
    def _eps(args, nargout):
        if args:
            return np.spacing(np.abs(dense(numeric(args[0]))))
        return np.array([[np.finfo(float).eps]])

    @reg("pi")
    def _pi(args, nargout):
        return np.array([[math.pi]])

    @reg("size")
    def _size(args, nargout):
        v = args[0]
        shp = (1, len(v)) if isinstance(v, str) else (v.shape if isinstance(v, Cell) else mat(v).shape)
        if len(args) == 2:
            d = int(scalar(args[1]))
            return np.array([[float(shp[d - 1] if d <= 2 else 1)]])
        if nargout <= 1:
            return np.array([[float(shp[0]), float(shp[1])]])
        return tuple(np.array([[float(s)]]) for s in shp[:nargout])

    @reg("numel")
    def _numel(args, nargout):
        v = args[0]
        if isinstance(v, str):
            return np.array([[float(len(v))]])
        shp = v.shape if isinstance(v, Cell) else mat(v).shape
        return np.array([[float(shp[0] * shp[1])]])

    @reg("length")
    def _length(args, nargout):
        v = args[0]
        if isinstance(v, str):
            return np.array([[float(len(v))]])
        shp = v.shape if isinstance(v, Cell) else mat(v).shape
        return np.array([[float(0 if 0 in shp else max(shp))]])

    @reg("isempty")
    def _isempty(args, nargout):
        v = args[0]
        if isinstance(v, str):
            return np.array([[len(v) == 0]])
        shp = v.shape if isinstance(v, Cell) else mat(v).shape
        return np.array([[0 in shp]])

    @reg("issparse")
    def _issparse(args, nargout):
        return np.array([[sp.issparse(args[0])]])

    @reg("sparse")
    def _sparse(args, nargout):
        if len(args) == 1:
            return sp.csc_matrix(dense(numeric(args[0])))
        if len(args) == 2:
            return sp.csc_matrix((int(scalar(args[0])), int(scalar(args[1]))))
        i = dense(mat(args[0])).reshape(-1).astype(np.int64) - 1
        j = dense(mat(args[1])).reshape(-1).astype(np.int64) - 1
        v = dense(numeric(args[2])).reshape(-1)
        if v.size == 1 and i.size != 1:
            v = np.full(i.size, v[0])
        shape = (int(scalar(args[3])), int(scalar(args[4]))) if len(args) >= 5 else (int(i.max()) + 1, int(j.max()) + 1)
        return sp.coo_matrix((v, (i, j)), shape=shape).tocsc()

    @reg("full")
    def _full(args, nargout):
        return dense(mat(args[0])).copy()

    @reg("nnz")
    def _nnz(args, nargout):
        v = mat(args[0])
        return np.array([[float(v.count_nonzero() if sp.issparse(v) else np.count_nonzero(v))]])

    @reg("norm")
    def _norm(args, nargout):
        v = numeric(args[0])
        kind = args[1] if len(args) > 1 else 2
        if isinstance(kind, str):
            if kind == "fro":
                if sp.issparse(v):
                    return np.array([[math.sqrt(float((v.data * v.data).sum()))]]) if v.nnz else np.zeros((1, 1))
                return np.array([[np.linalg.norm(v.reshape(-1))]])
            if kind in ("inf", "Inf"):
                kind = np.inf
            else:
                raise MlabError(f"norm(...,{kind!r}) is not supported")
        else:
            kind = scalar(kind)
        v = dense(v)
        if v.size == 0:
            return np.zeros((1, 1))
        if v.shape[0] == 1 or v.shape[1] == 1:
            x = v.reshape(-1)
            if kind == 2:
                return np.array([[np.linalg.norm(x)]])
            return np.array([[np.linalg.norm(x, kind)]])
        if kind == 2:
            return np.array([[np.linalg.svd(v, compute_uv=False)[0]]])
        return np.array([[np.linalg.norm(v, kind)]])

    def elementwise1(name, f):
        def g(args, nargout):
            v = dense(numeric(args[0]))
            with np.errstate(all="ignore"):
                return f(v)
        B[name] = g

    def _sqrt(v):
        if not np.iscomplexobj(v) and np.any(v < 0):
            return np.sqrt(v.astype(complex))
        return np.sqrt(v)

    def _log(v):
        if not np.iscomplexobj(v) and np.any(v < 0):
            return np.log(v.astype(complex))
        return np.log(v)

    elementwise1("sqrt", _sqrt)
    elementwise1("abs", np.abs)
    elementwise1("exp", np.exp)
    elementwise1("log", _log)
    elementwise1("log10", np.log10)
    elementwise1("sin", np.sin)
    elementwise1("cos", np.cos)
    elementwise1("floor", np.floor)
    elementwise1("ceil", np.ceil)
    elementwise1("real", lambda v: np.real(v).astype(float).copy())
    elementwise1("imag", lambda v: np.imag(v).astype(float).copy())
    elementwise1("isnan", np.isnan)
    elementwise1("isinf", np.isinf)
    elementwise1("isfinite", np.isfinite)
    elementwise1("round", lambda v: np.sign(v) * np.floor(np.abs(v) + 0.5))
    elementwise1("sign", np.sign)
    elementwise1("double", lambda v: v.astype(float))

    @reg("hypot")
    def _hypot(args, nargout):
        return np.hypot(dense(numeric(args[0])), dense(numeric(args[1])))

    @reg("mod")
    def _mod(args, nargout):
        a, b = dense(numeric(args[0])), dense(numeric(args[1]))
        with np.errstate(all="ignore"):
            return np.where(b == 0, a, np.mod(a, b))

    def reduce_dim(v, args_after):
        """dimension a reduction works along (0-based axis), MATLAB default rule"""
        if args_after:
            return int(scalar(args_after[0])) - 1
        return 0 if v.shape[0] != 1 else 1

    @reg("sum")
    def _sum(args, nargout):
        v = dense(numeric(args[0]))
        if v.size == 0:
            return np.zeros((1, 1))
        ax = reduce_dim(v, args[1:])
        return np.sum(v, axis=ax, keepdims=True)

    @reg("prod")
    def _prod(args, nargout):
        v = dense(numeric(args[0]))
        ax = reduce_dim(v, args[1:])
        return np.prod(v, axis=ax, keepdims=True)

    @reg("mean")
    def _mean(args, nargout):
        v = dense(numeric(args[0]))
        ax = reduce_dim(v, args[1:])
        return np.mean(v, axis=ax, keepdims=True)

    @reg("cumsum")
    def _cumsum(args, nargout):
        v = dense(numeric(args[0]))
        ax = reduce_dim(v, args[1:])
        return np.cumsum(v, axis=ax)

    @reg("any")
    def _any(args, nargout):
        v = dense(mat(args[0]))
        ax = reduce_dim(v, args[1:])
        return np.any(v != 0, axis=ax, keepdims=True)

    @reg("all")
    def _all(args, nargout):
        v = dense(mat(args[0]))
        ax = reduce_dim(v, args[1:])
        return np.all(v != 0, axis=ax, keepdims=True)

    def minmax(which):
        def f(args, nargout):
            a = dense(numeric(args[0]))
            if len(args) >= 2 and not (isinstance(args[1], np.ndarray) and args[1].size == 0):
                b = dense(numeric(args[1]))
                with np.errstate(all="ignore"):
                    # MATLAB ignores NaNs in min/max
                    return np.fmin(a, b) if which == "min" else np.fmax(a, b)
            if a.size == 0:
                return (np.zeros((0, 0)), np.zeros((0, 0)))[: max(nargout, 1)]
            ax = int(scalar(args[2])) - 1 if len(args) >= 3 else (0 if a.shape[0] != 1 else 1)
            with np.errstate(all="ignore"):
                allnan = np.all(np.isnan(a), axis=ax, keepdims=True)
                filled = np.where(np.isnan(a), np.inf if which == "min" else -np.inf, a)
                idx = (np.argmin if which == "min" else np.argmax)(filled, axis=ax)  # first occurrence
                val = np.take_along_axis(a, np.expand_dims(idx, ax), axis=ax)
                val = np.where(allnan, np.nan, val)
            idx = np.expand_dims(idx, ax).astype(float) + 1.0
            return (val, idx)[: max(nargout, 1)] if nargout > 1 else val
        return f

    B["min"] = minmax("min")
    B["max"] = minmax("max")

    @reg("diag")
    def _diag(args, nargout):
        v = dense(numeric(args[0]))
        k = int(scalar(args[1])) if len(args) > 1 else 0
        if v.shape[0] == 1 or v.shape[1] == 1:
            return np.diag(v.reshape(-1), k)
        return np.diag(v, k).reshape(-1, 1).copy()

    @reg("trace")
    def _trace(args, nargout):
        return np.array([[np.trace(dense(numeric(args[0])))]])

    @reg("svd")
    def _svd(args, nargout):
        v = dense(numeric(args[0]))
        econ = len(args) > 1
        if v.size == 0:
            z = np.zeros((0, 0))
            return (z, z, z)[: max(nargout, 1)] if nargout > 1 else np.zeros((0, 1))
        if nargout <= 1:
            return np.linalg.svd(v, compute_uv=False).reshape(-1, 1)
        U, s, Vt = np.linalg.svd(v, full_matrices=not econ)
        if econ:
            S = np.diag(s)
        else:
            S = np.zeros(v.shape)
            S[: s.size, : s.size] = np.diag(s)
        return (U, S, Vt.T.copy())[:nargout]

    @reg("eig")
    def _eig(args, nargout):
        v = dense(numeric(args[0]))
        if np.array_equal(v, np.conj(v.T)):
            w, V = np.linalg.eigh(v)
        else:
            w, V = np.linalg.eig(v)
            if np.iscomplexobj(w) and np.all(w.imag == 0):
                w, V = w.real, V.real
        if nargout <= 1:
            return w.reshape(-1, 1)
        return V, np.diag(w)

    @reg("sort")
    def _sort(args, nargout):
        v = dense(numeric(args[0]))
        desc = any(isinstance(a, str) and a == "descend" for a in args[1:])
        if v.shape[0] == 1 or v.shape[1] == 1:
            x = v.reshape(-1)
            key = x if not np.iscomplexobj(x) else np.abs(x)
            # stable; MATLAB's descending sort keeps equal elements in original order
            order = np.argsort(-key if desc else key, kind="stable")
            out = x[order].reshape(v.shape)
            idx = (order.astype(float) + 1).reshape(v.shape)
            return (out, idx)[: max(nargout, 1)] if nargout > 1 else out
        order = np.argsort(-v if desc else v, axis=0, kind="stable")
        out = np.take_along_axis(v, order, axis=0)
        return (out, order.astype(float) + 1) if nargout > 1 else out

    @reg("find")
    def _find(args, nargout):
        v = dense(mat(args[0]))
        idx = np.flatnonzero(v.reshape(-1, order="F") != 0).astype(float) + 1
        if len(args) > 1:
            idx = idx[: int(scalar(args[1]))]
        return idx.reshape(1, -1) if v.shape[0] == 1 and v.shape[1] != 1 else idx.reshape(-1, 1)

    @reg("linspace")
    def _linspace(args, nargout):
        n = int(scalar(args[2])) if len(args) > 2 else 100
        return np.linspace(scalar(args[0]), scalar(args[1]), n).reshape(1, -1)

    @reg("logspace")
    def _logspace(args, nargout):
        n = int(scalar(args[2])) if len(args) > 2 else 50
        return (10.0 ** np.linspace(scalar(args[0]), scalar(args[1]), n)).reshape(1, -1)

    @reg("strcmp")
    def _strcmp(args, nargout):
        return np.array([[isinstance(args[0], str) and isinstance(args[1], str) and args[0] == args[1]]])

    @reg("strcmpi")
    def _strcmpi(args, nargout):
        return np.array([[isinstance(args[0], str) and isinstance(args[1], str) and args[0].lower() == args[1].lower()]])

    @reg("lower")
    def _lower(args, nargout):
        return args[0].lower()

    @reg("upper")
    def _upper(args, nargout):
        return args[0].upper()

    @reg("num2str")
    def _num2str(args, nargout):
        return f"{scalar(args[0]):g}"

    @reg("error")
    def _error(args, nargout):
        raise MlabError("error(): " + " ".join(str(a) for a in args))

    def _noop(args, nargout):
        return ()

    for nm in ("warning", "tic", "figure", "drawnow", "clc"):
        B[nm] = _noop

    def _fmt_args(args):
        out = []
        for a in args:
            if isinstance(a, str):
                out.append(a)
            else:
                v = dense(numeric(a)).reshape(-1, order="F")
                out.extend(float(x) if not float(x).is_integer() else int(x) for x in np.real(v))
        return out

    @reg("sprintf")
    def _sprintf(args, nargout):
        fmt = args[0].replace("\\n", "\n").replace("\\t", "\t")
        vals = _fmt_args(args[1:])
        nspec = len(re.findall(r"%(?!%)", fmt))
        if nspec == 0:
            return fmt.replace("%%", "%")
        chunks = []
        for i in range(0, max(len(vals), 1), nspec):  # MATLAB recycles the format over the arguments
            part = vals[i:i + nspec]
            if len(part) < nspec:
                break
            part = [float(v) if isinstance(v, int) and re.search(r"%[-+0-9.]*[efg]", fmt) and False else v for v in part]
            try:
                chunks.append(fmt % tuple(part))
            except TypeError:
                chunks.append(fmt % tuple(float(v) if not isinstance(v, str) else v for v in part))
        return "".join(chunks)

    @reg("fprintf")
    def _fprintf(args, nargout):
        if args and not isinstance(args[0], str):
            args = args[1:]  # file id
        ip.out.append(_sprintf(args, 1))
        return ()

    @reg("disp")
    def _disp(args, nargout):
        ip.out.append(str(args[0]) + "\n")
        return ()

    @reg("load")
    def _load(args, nargout):
        import scipy.io as sio
        raw = sio.loadmat(args[0], squeeze_me=False, struct_as_record=False)
        out = {}
        for k, v in raw.items():
            if k.startswith("__"):
                continue
            if sp.issparse(v):
                out[k] = v.tocsc()
            elif isinstance(v, np.ndarray) and v.dtype.kind in "US":
                out[k] = str(v.reshape(-1)[0]) if v.size else ""
            elif isinstance(v, np.ndarray) and v.dtype.kind in "iufb":
                out[k] = np.atleast_2d(v).astype(float)
            else:
                out[k] = v
        return out

    @reg("isfield")
    def _isfield(args, nargout):
        return np.array([[isinstance(args[0], dict) and isinstance(args[1], str) and args[1] in args[0]]])

    @reg("arrayfun")
    def _arrayfun(args, nargout):
        f, x = args[0], dense(numeric(args[1]))
        vals = [scalar(f.fn([np.array([[v]])], 1)[0]) for v in x.reshape(-1, order="F")]
        return np.array(vals, dtype=float).reshape(x.shape, order="F")

    @reg("toc")
    def _toc(args, nargout):
        return np.zeros((1, 1))

    @reg("feval")
    def _feval(args, nargout):
        f = args[0]
        if isinstance(f, FuncHandle):
            return f.fn(args[1:], nargout)
        return ip.call(f, args[1:], nargout)

    @reg("isa")
    def _isa(args, nargout):
        return np.array([[args[1] == "function_handle" and isinstance(args[0], FuncHandle)]])

    @reg("optimset")
    def _optimset(args, nargout):
        return {str(args[i]): args[i + 1] for i in range(0, len(args) - 1, 2)}

    @reg("fminbnd")
    def _fminbnd(args, nargout):
        # MATLAB's fminbnd is a built-in (Forsythe-Malcolm-Moler golden section + parabolic
        # interpolation, documented); oracle/fminbnd.py restates it.
        from .fminbnd import fminbnd as fmb
        f = args[0]
        opts = args[3] if len(args) > 3 and isinstance(args[3], dict) else {}
        tolx = float(scalar(opts["TolX"])) if "TolX" in opts else 1e-4
        obj = lambda x: float(scalar(f.fn([np.array([[x]])], 1)[0]))
        res = fmb(obj, float(scalar(args[1])), float(scalar(args[2])), tolx)
        x, fval = res[0], res[1]
        outs = (np.array([[x]]), np.array([[fval]]), np.array([[1.0]]))
        return outs[: max(nargout, 1)] if nargout > 1 else outs[0]

    return B
