"""Parser of minijl: tokens → AST (plain tuples).  See lexer.py for what this is and is not.

AST nodes (first element = kind):
  ("num", v) ("str", parts) ("char", c) ("sym", name) ("name", id) ("bool", b) ("colon",) ("endidx",)
  ("call", f, args, kwargs) ("dotcall", f, args)  ("index", obj, idxs) ("field", obj, name) ("curly", obj, params)
  ("unop", op, x) ("binop", op, a, b) ("dotop", op, a, b) ("cmp", [operands], [ops]) ("and", a, b) ("or", a, b)
  ("ternary", c, a, b) ("range", a, b, step|None) ("tuple", items) ("vect", items) ("vcat", items) ("matrix", rows)
  ("comprehension", expr, var, iter) ("typed_vect", type, items) ("lambda", params, body) ("splat", x) ("decl", name|None, type)
  ("kw", name, value) ("adjoint", x) ("subtype", a, b) ("typevar_ub", type)  [`<:T` as a type parameter]
  ("assign", lhs, rhs) ("opassign", op, lhs, rhs) ("block", stmts) ("if", [(cond, block)...], else|None)
  ("for", var, iter, body) ("while", cond, body) ("return", x|None) ("break",) ("continue",)
  ("function", name|None, functor|None, params, kwparams, body) ("struct", name, tparams, super, fields, mutable)
  ("abstract", name, super) ("const", assign) ("global", stmt) ("local", stmt) ("using",) ("macrocall", name, args)
"""
from __future__ import annotations

from .lexer import JlSyntaxError, Tok, lex

BINARY_PREC = {
    "||": 4, "&&": 5,
    "|>": 8,
    ":": 10,
    "+": 11, "-": 11, "|": 11, ".+": 11, ".-": 11,
    "*": 12, "/": 12, "%": 12, "&": 12, "÷": 12, "\\": 12, ".*": 12, "./": 12,
    "<<": 13, ">>": 13,
    "^": 15, ".^": 15,
}
COMPARISONS = {"==", "!=", "<", "<=", ">", ">=", "<:", ">:", "in", "isa", "≤", "≥", "≠", "===", "!=="}
CMP_PREC = 7
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "^=", "%=", ".=", "÷="}
BLOCK_END = {"end", "else", "elseif", "catch", "finally"}


class Parser:
    def __init__(self, toks, filename="<string>"):
        self.toks = toks
        self.pos = 0
        self.filename = filename
        self.nl_skip = [False]      # newline-insensitive inside ( [ {, sensitive again inside block constructs
        self.in_bracket = [False]   # inside [ ]: whitespace separates elements and `end` is an index
        self.in_index = [False]

    # ---- token access -------------------------------------------------------------------------------------
    def _skip(self):
        if self.nl_skip[-1]:
            while self.toks[self.pos].kind == "nl":
                self.pos += 1

    def peek(self) -> Tok:
        self._skip()
        return self.toks[self.pos]

    def next(self) -> Tok:
        self._skip()
        t = self.toks[self.pos]
        self.pos += 1
        return t

    def err(self, msg, t=None):
        t = t or self.peek()
        raise JlSyntaxError(f"{self.filename}:{t.line}: {msg} (at {t.kind} {t.val!r})")

    def is_op(self, v):
        t = self.peek()
        return t.kind == "op" and t.val == v

    def is_kw(self, v):
        t = self.peek()
        return t.kind == "kw" and t.val == v

    def expect_op(self, v):
        t = self.next()
        if t.kind != "op" or t.val != v:
            self.err(f"expected {v!r}", t)
        return t

    def expect_kw(self, v):
        t = self.next()
        if t.kind != "kw" or t.val != v:
            self.err(f"expected {v!r}", t)
        return t

    def skip_terminators(self):
        while True:
            t = self.toks[self.pos]
            if t.kind == "nl" or (t.kind == "op" and t.val == ";"):
                self.pos += 1
            else:
                break

    # ---- program / blocks ----------------------------------------------------------------------------------
    def parse_program(self):
        stmts = []
        self.skip_terminators()
        while self.peek().kind != "eof":
            stmts.append(self.parse_statement())
            self.skip_terminators()
        return ("block", stmts)

    def parse_block(self, terminators=BLOCK_END):
        """Statements until one of the terminating keywords (not consumed).  Newlines are significant."""
        self.nl_skip.append(False)
        self.in_bracket.append(False)
        self.in_index.append(False)
        stmts = []
        self.skip_terminators()
        while not (self.toks[self.pos].kind == "kw" and self.toks[self.pos].val in terminators):
            if self.toks[self.pos].kind == "eof":
                self.err("unexpected end of input inside a block")
            stmts.append(self.parse_statement())
            self.skip_terminators()
        self.nl_skip.pop()
        self.in_bracket.pop()
        self.in_index.pop()
        return ("block", stmts)

    def parse_statement(self):
        t = self.peek()
        if t.kind == "kw":
            v = t.val
            if v in ("using", "import", "export"):
                while self.toks[self.pos].kind not in ("nl", "eof") and not (self.toks[self.pos].kind == "op" and self.toks[self.pos].val == ";"):
                    self.pos += 1
                return ("using",)
            if v == "abstract":
                self.next()
                self.expect_kw("type")
                name = self.parse_binary(CMP_PREC + 1)
                sup = None
                if self.is_op("<:"):
                    self.next()
                    sup = self.parse_binary(CMP_PREC + 1)
                self.expect_kw("end")
                return ("abstract", name, sup)
            if v in ("struct", "mutable"):
                return self.parse_struct()
            if v == "const":
                self.next()
                return ("const", self.parse_statement())
            if v == "global":
                self.next()
                return ("global", self.parse_statement())
            if v == "local":
                self.next()
                return ("local", self.parse_statement())
            if v == "return":
                self.next()
                nt = self.toks[self.pos]
                if nt.kind in ("nl", "eof") or (nt.kind == "op" and nt.val == ";") or (nt.kind == "kw" and nt.val in BLOCK_END):
                    return ("return", None)
                return ("return", self.parse_expr_stmt())
            if v == "break":
                self.next()
                return ("break",)
            if v == "continue":
                self.next()
                return ("continue",)
        return self.parse_expr_stmt()

    def parse_expr_stmt(self):
        """An expression statement; a top-level comma makes a tuple (`a, b = f()`, `"--opt", "-o"`)."""
        e = self.parse_expr(0)
        if self.is_op(",") and not self.nl_skip[-1]:
            items = [e]
            while self.is_op(","):
                self.next()
                # allow the continuation on the next line after a trailing comma
                while self.toks[self.pos].kind == "nl":
                    self.pos += 1
                items.append(self.parse_expr(1))
            e = ("tuple", items)
            if self.peek().kind == "op" and self.peek().val in ASSIGN_OPS:
                op = self.next().val
                rhs = self.parse_expr_stmt()
                return ("assign", e, rhs) if op == "=" else ("opassign", op[:-1], e, rhs)
        return e

    def parse_struct(self):
        mutable = False
        if self.is_kw("mutable"):
            self.next()
            mutable = True
        self.expect_kw("struct")
        head = self.parse_binary(CMP_PREC + 1)
        sup = None
        if self.is_op("<:"):
            self.next()
            sup = self.parse_binary(CMP_PREC + 1)
        tparams = []
        if head[0] == "curly":
            tparams = [p[1] if p[0] == "name" else p for p in head[2]]
            head = head[1]
        if head[0] != "name":
            self.err("bad struct name")
        body = self.parse_block()
        self.expect_kw("end")
        fields = []
        for s in body[1]:
            if s[0] == "decl":
                fields.append((s[1], s[2]))
            elif s[0] == "name":
                fields.append((s[1], None))
            else:
                self.err(f"unsupported struct member {s[0]}")
        return ("struct", head[1], tparams, sup, fields, mutable)

    # ---- expressions ----------------------------------------------------------------------------------------
    def parse_expr(self, min_prec=0):
        """min_prec 0: assignment allowed; 1: no assignment (call arguments handle `k=v` themselves)."""
        left = self.parse_ternary()
        while self.is_op("=>"):          # a => b  (a Pair)
            self.next()
            left = ("call", ("name", "Pair"), [left, self.parse_ternary()], [])
        if min_prec == 0:
            t = self.peek()
            if t.kind == "op" and t.val in ASSIGN_OPS and not self._array_space_break(t):
                self.next()
                while self.toks[self.pos].kind == "nl":   # `x =\n value`
                    self.pos += 1
                rhs = self.parse_expr_stmt() if not self.nl_skip[-1] else self.parse_expr(0)
                if t.val == "=":
                    if left[0] == "call" or (left[0] == "decl" and left[1] is not None and isinstance(left[1], tuple)):
                        return self.short_function(left, rhs)
                    return ("assign", left, rhs)
                return ("opassign", t.val[:-1], left, rhs)
        return left

    def short_function(self, sig, body):
        callee, args, kwargs = sig[1], sig[2], sig[3]
        if callee[0] == "paren":
            callee = callee[1]
        params = [self.to_param(a) for a in args]
        kwparams = [self.to_param(k) for k in kwargs]
        if callee[0] == "name":
            return ("function", callee[1], None, params, kwparams, body)
        if callee[0] == "decl":       # (dr::Type)(args) = ...   functor method
            return ("function", None, (callee[1], callee[2]), params, kwparams, body)
        if callee[0] == "curly":      # Name{T}(args) = ...
            return ("function", callee[1][1], None, params, kwparams, body)
        self.err("unsupported method definition")

    def to_param(self, a):
        """Parameter AST → (name|None, type|None, default|None, splat)"""
        if a[0] == "kw":
            inner = self.to_param(a[1]) if isinstance(a[1], tuple) else (a[1], None, None, False)
            return (inner[0], inner[1], a[2], False)
        if a[0] == "name":
            return (a[1], None, None, False)
        if a[0] == "decl":
            nm = a[1]
            if isinstance(nm, tuple):
                nm = nm[1]
            return (nm, a[2], None, False)
        if a[0] == "splat":
            p = self.to_param(a[1])
            return (p[0], p[1], None, True)
        self.err(f"unsupported parameter form {a[0]}")

    def skip_nl(self):
        """An operator at the end of a line continues the expression on the next line."""
        while self.toks[self.pos].kind == "nl":
            self.pos += 1

    def parse_ternary(self):
        cond = self.parse_arrow()
        if self.is_op("?") and not self._array_space_break(self.peek()):
            self.next()
            self.skip_nl()
            a = self.parse_ternary_branch()
            self.expect_op(":")
            self.skip_nl()
            b = self.parse_ternary_branch()
            return ("ternary", cond, a, b)
        return cond

    def parse_ternary_branch(self):
        self.no_range = getattr(self, "no_range", 0) + 1
        try:
            return self.parse_ternary()
        finally:
            self.no_range -= 1

    def parse_arrow(self):
        left = self.parse_binary(4)
        if self.is_op("->"):
            self.next()
            if left[0] == "tuple":
                params = [self.to_param(p) for p in left[1]]
            elif left[0] == "paren":
                params = [self.to_param(left[1])]
            else:
                params = [self.to_param(left)]
            saved = getattr(self, "no_range", 0)
            self.no_range = 0
            body = self.parse_expr(0)
            self.no_range = saved
            return ("lambda", params, body)
        if left[0] == "paren":
            return left[1]
        return left

    def _array_space_break(self, t: Tok) -> bool:
        """Inside [ ]: does whitespace before token t start a new element?  `a -b` → yes, `a - b` / `a-b` → no."""
        if not self.in_bracket[-1]:
            return False
        if not t.sp_before:
            return False
        if t.kind == "op" and t.val in ("+", "-", "'", "?", ":") and not t.sp_after:
            return True
        return False

    def parse_binary(self, min_prec):
        left = self.parse_unary()
        while True:
            t = self.peek()
            if t.kind == "kw" and t.val in ("in", "isa"):
                op = t.val
            elif t.kind == "op":
                op = t.val
            else:
                break
            if self._array_space_break(t):
                break
            if op in COMPARISONS:
                if CMP_PREC < min_prec:
                    break
                operands, ops = [left], []
                while True:
                    t = self.peek()
                    o = t.val if (t.kind == "op" or (t.kind == "kw" and t.val in ("in", "isa"))) else None
                    if o not in COMPARISONS:
                        break
                    self.next()
                    ops.append(o)
                    operands.append(self.parse_binary(CMP_PREC + 1))
                if ops == ["<:"]:
                    left = ("subtype", operands[0], operands[1])
                else:
                    left = ("cmp", operands, ops)
                continue
            if op == "&&" or op == "||":
                prec = BINARY_PREC[op]
                if prec < min_prec:
                    break
                self.next()
                self.skip_nl()
                right = self.parse_binary(prec + 1)
                left = ("and" if op == "&&" else "or", left, right)
                continue
            if op == ":":
                if getattr(self, "no_range", 0) or BINARY_PREC[":"] < min_prec:
                    break
                # a range; `a:b:c` = start:step:stop
                self.next()
                if self.in_index[-1] and (self.is_op(",") or self.is_op("]")):
                    self.err("unsupported open-ended range")
                b = self.parse_binary(BINARY_PREC[":"] + 1)
                if self.is_op(":") and not getattr(self, "no_range", 0):
                    self.next()
                    c = self.parse_binary(BINARY_PREC[":"] + 1)
                    left = ("range", left, c, b)
                else:
                    left = ("range", left, b, None)
                continue
            prec = BINARY_PREC.get(op)
            if prec is None or prec < min_prec:
                break
            self.next()
            self.skip_nl()
            if op in ("^", ".^"):
                right = self.parse_unary_pow()
            else:
                right = self.parse_binary(prec + 1)
            if op.startswith(".") and len(op) == 2:
                left = ("dotop", op[1], left, right)
            else:
                left = ("binop", op, left, right)
        return left

    def parse_unary_pow(self):
        # right operand of ^ : unary minus allowed, right associative
        t = self.peek()
        if t.kind == "op" and t.val in ("-", "+"):
            self.next()
            x = self.parse_unary_pow()
            return ("unop", t.val, x)
        base = self.parse_postfix()
        if self.is_op("^"):
            self.next()
            return ("binop", "^", base, self.parse_unary_pow())
        return base

    def parse_unary(self):
        t = self.peek()
        if t.kind == "op" and t.val in ("-", "+", "!"):
            self.next()
            # unary minus binds weaker than ^ but stronger than * /
            x = self.parse_unary()
            while self.is_op("^") or self.is_op(".^"):   # -x^2 = -(x^2)
                op = self.next().val
                x = ("binop", "^", x, self.parse_unary_pow()) if op == "^" else ("dotop", "^", x, self.parse_unary_pow())
            if t.val == "-" and x[0] == "num":
                return ("num", -x[1])
            return ("unop", t.val, x)
        if t.kind == "op" and t.val == "<:":     # `<:T` as a type parameter
            self.next()
            return ("typevar_ub", self.parse_postfix())
        if t.kind == "op" and t.val == "::":     # anonymous typed parameter `::T`
            self.next()
            return ("decl", None, self.parse_type_postfix())
        if t.kind == "op" and t.val == ":" :
            # a bare colon as an index / argument: `a[:, i]`, `reshape(x, 1, :)`
            nt = self.toks[self.pos + 1] if self.toks[self.pos].val == ":" else None
            self.next()
            return ("colon",)
        return self.parse_postfix()

    def parse_type_postfix(self):
        e = self.parse_primary()
        while True:
            t = self.peek()
            if t.kind == "op" and t.val == "{" and not t.sp_before:
                e = ("curly", e, self.parse_curly_params())
            elif t.kind == "op" and t.val == "." and not t.sp_before:
                self.next()
                e = ("field", e, self.next().val)
            else:
                return e

    def parse_curly_params(self):
        self.expect_op("{")
        self.nl_skip.append(True)
        self.in_bracket.append(False)
        self.in_index.append(False)
        params = []
        while not self.is_op("}"):
            params.append(self.parse_expr(1))
            if self.is_op(","):
                self.next()
        self.nl_skip.pop()
        self.in_bracket.pop()
        self.in_index.pop()
        self.expect_op("}")
        return params

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            t = self.toks[self.pos]      # postfix operators never follow a newline
            if self.nl_skip[-1]:
                t = self.peek()
            if t.kind != "op":
                break
            if t.val == "(" and not t.sp_before:
                args, kwargs = self.parse_call_args()
                e = ("call", e, args, kwargs)
                nt = self.toks[self.pos]
                if nt.kind == "kw" and nt.val == "do":     # f(args) do x ... end  ≡  f(x -> ..., args)
                    self.pos += 1
                    params = []
                    while self.toks[self.pos].kind not in ("nl",) and not (self.toks[self.pos].kind == "op" and self.toks[self.pos].val == ";"):
                        params.append(self.to_param(self.parse_binary(CMP_PREC + 1)))
                        if self.toks[self.pos].kind == "op" and self.toks[self.pos].val == ",":
                            self.pos += 1
                    body = self.parse_block()
                    self.expect_kw("end")
                    e = ("call", e[1], [("lambda", params, body)] + args, kwargs)
            elif t.val == "[" and not t.sp_before:
                self.next()
                idxs = self.parse_index_list()
                if e[0] in ("curly",) or (e[0] == "name" and e[1][:1].isupper() and self._looks_like_type(e)):
                    e = ("typed_vect", e, idxs)
                else:
                    e = ("index", e, idxs)
            elif t.val == "{" and not t.sp_before:
                e = ("curly", e, self.parse_curly_params())
            elif t.val == "." and not t.sp_before:
                nt = self.toks[self.pos + 1]
                if nt.kind == "op" and nt.val == "(":       # broadcast call f.(x)
                    self.next()
                    args, kwargs = self.parse_call_args()
                    e = ("dotcall", e, args)
                elif nt.kind in ("id", "kw"):
                    self.next()
                    self.next()
                    e = ("field", e, nt.val)
                else:
                    break
            elif t.val == "'" and not t.sp_before:
                self.next()
                e = ("adjoint", e)
            elif t.val == "::" :
                self.next()
                ty = self.parse_type_postfix()
                e = ("decl", e[1] if e[0] == "name" else e, ty)
            elif t.val == "...":
                self.next()
                e = ("splat", e)
            else:
                break
        return e

    def _looks_like_type(self, e):
        return e[1] in ("Any", "Float64", "Int", "Int64", "String", "Bool", "Real", "Vector", "Function")

    def parse_call_args(self):
        saved_nr, self.no_range = getattr(self, "no_range", 0), 0
        try:
            return self._parse_call_args()
        finally:
            self.no_range = saved_nr

    def _parse_call_args(self):
        self.expect_op("(")
        self.nl_skip.append(True)
        self.in_bracket.append(False)
        self.in_index.append(False)
        args, kwargs = [], []
        after_semi = False
        while not self.is_op(")"):
            if self.is_op(";"):
                self.next()
                after_semi = True
                continue
            a = self.parse_expr(1)
            if self.is_op("=") :
                self.next()
                val = self.parse_expr(1)
                if a[0] == "decl":
                    kwargs.append(("kw", a, val))
                else:
                    kwargs.append(("kw", a[1], val))
            elif after_semi:
                kwargs.append(("kw", a[1] if a[0] == "name" else a, None))
            else:
                args.append(a)
            if self.is_op(","):
                self.next()
        self.nl_skip.pop()
        self.in_bracket.pop()
        self.in_index.pop()
        self.expect_op(")")
        return args, kwargs

    def parse_index_list(self):
        saved_nr, self.no_range = getattr(self, "no_range", 0), 0
        try:
            return self._parse_index_list()
        finally:
            self.no_range = saved_nr

    def _parse_index_list(self):
        """After '[' of an indexing expression: comma-separated indices until ']'."""
        self.nl_skip.append(True)
        self.in_bracket.append(False)
        self.in_index.append(True)
        idxs = []
        while not self.is_op("]"):
            idxs.append(self.parse_expr(1))
            if self.is_op(",") or self.is_op(";"):   # `;`: the typed vertical concatenation T[a; b] of scalars
                self.next()
        self.nl_skip.pop()
        self.in_bracket.pop()
        self.in_index.pop()
        self.expect_op("]")
        return idxs

    def parse_array_literal(self):
        saved_nr, self.no_range = getattr(self, "no_range", 0), 0
        try:
            return self._parse_array_literal()
        finally:
            self.no_range = saved_nr

    def _parse_array_literal(self):
        """After '[': vect `[a, b]`, vcat `[a; b]`, hcat/matrix `[a b; c d]`, comprehension `[f(x) for x in xs]`."""
        self.nl_skip.append(True)
        self.in_bracket.append(True)
        self.in_index.append(False)
        try:
            if self.is_op("]"):
                self.next()
                return ("vect", [])
            first = self.parse_expr(1)
            if self.is_kw("for"):
                self.next()
                var = self.parse_binary(CMP_PREC + 1)
                t = self.next()
                if not ((t.kind == "op" and t.val == "=") or (t.kind == "kw" and t.val == "in")):
                    self.err("expected = or in", t)
                it = self.parse_expr(1)
                cond = None
                if self.is_kw("if"):
                    self.next()
                    cond = self.parse_expr(1)
                self.expect_op("]")
                return ("comprehension", first, var, it, cond)
            if self.is_op(","):
                items = [first]
                while self.is_op(","):
                    self.next()
                    if self.is_op("]"):
                        break
                    items.append(self.parse_expr(1))
                self.expect_op("]")
                return ("vect", items)
            rows, row = [], [first]
            while True:
                if self.is_op("]"):
                    self.next()
                    rows.append(row)
                    break
                if self.is_op(";"):
                    self.next()
                    rows.append(row)
                    row = []
                    if self.is_op("]"):
                        self.next()
                        break
                    row.append(self.parse_expr(1))
                    continue
                row.append(self.parse_expr(1))
            if all(len(r) == 1 for r in rows):
                if len(rows) == 1:
                    return ("vect", [rows[0][0]])
                return ("vcat", [r[0] for r in rows])
            return ("matrix", rows)
        finally:
            self.nl_skip.pop()
            self.in_bracket.pop()
            self.in_index.pop()

    def parse_primary(self):
        t = self.next()
        k = t.kind
        if k == "num":
            nt = self.toks[self.pos]
            if nt.kind == "id" and not nt.sp_before:      # numeric-literal coefficient: 3π = 3*π, binds tighter than * /
                self.pos += 1
                return ("binop", "*", ("num", t.val), ("name", nt.val))
            return ("num", t.val)
        if k == "str":
            parts = []
            for p in t.parts:
                if isinstance(p, tuple):
                    sub = Parser(lex(p[1]), self.filename)
                    sub.nl_skip = [True]
                    parts.append(sub.parse_expr(1))
                else:
                    parts.append(p)
            return ("str", parts)
        if k == "cmd":
            parts = []
            for p in t.parts:
                if isinstance(p, tuple):
                    sub = Parser(lex(p[1]), self.filename)
                    sub.nl_skip = [True]
                    parts.append(sub.parse_expr(1))
                else:
                    parts.append(p)
            return ("cmd", parts)
        if k == "char":
            return ("char", t.val)
        if k == "sym":
            return ("sym", t.val)
        if k == "id":
            nt = self.toks[self.pos]
            if t.val == "glob" and nt.kind == "str" and not nt.sp_before:
                # the string macro glob"pattern" of Glob.jl: a GlobMatch
                self.pos += 1
                return ("call", ("field", ("name", "Glob"), "GlobMatch"), [("str", list(nt.parts))], [])
            return ("name", t.val)
        if k == "macro":
            return self.parse_macro(t)
        if k == "kw":
            v = t.val
            if v == "true" or v == "false":
                return ("bool", v == "true")
            if v == "end" and self.in_index[-1]:
                return ("endidx",)
            if v == "begin":
                if self.in_index[-1]:
                    return ("num", 1)
                body = self.parse_block()
                self.expect_kw("end")
                return body
            if v == "if":
                return self.parse_if()
            if v == "for":
                return self.parse_for()
            if v == "while":
                self.nl_skip.append(False)
                cond = self.parse_expr(1)
                self.nl_skip.pop()
                body = self.parse_block()
                self.expect_kw("end")
                return ("while", cond, body)
            if v == "function":
                return self.parse_function()
            if v == "let":
                body = self.parse_block()
                self.expect_kw("end")
                return body
            if v == "return":
                nt = self.toks[self.pos]
                if nt.kind in ("nl", "eof") or (nt.kind == "op" and nt.val in (";", ")", ",")) or (nt.kind == "kw" and nt.val in BLOCK_END):
                    return ("return", None)
                return ("return", self.parse_expr(1))
            if v == "break":
                return ("break",)
            if v == "continue":
                return ("continue",)
            if v == "try":
                body = self.parse_block()
                catch_var, catch_body, fin = None, None, None
                if self.is_kw("catch"):
                    self.next()
                    nt = self.toks[self.pos]
                    if nt.kind == "id":
                        self.pos += 1
                        catch_var = nt.val
                    catch_body = self.parse_block()
                if self.is_kw("finally"):
                    self.next()
                    fin = self.parse_block()
                self.expect_kw("end")
                return ("try", body, catch_var, catch_body, fin)
            self.err(f"unexpected keyword {v!r}", t)
        if k == "op":
            v = t.val
            if v == "(":
                self.nl_skip.append(True)
                self.in_bracket.append(False)
                self.in_index.append(False)
                saved_nr, self.no_range = getattr(self, "no_range", 0), 0
                try:
                    if self.is_op(")"):
                        self.next()
                        return ("tuple", [])
                    first = self.parse_expr(0)
                    if self.is_op(";"):            # (a; b; c): a block whose value is its last expression
                        items = [first]
                        while self.is_op(";"):
                            self.next()
                            if self.is_op(")"):
                                break
                            items.append(self.parse_expr(0))
                        self.expect_op(")")
                        return ("paren", ("block", items)) if len(items) > 1 else ("paren", items[0])
                    if self.is_op(","):
                        items = [first]
                        while self.is_op(","):
                            self.next()
                            if self.is_op(")"):
                                break
                            items.append(self.parse_expr(0))
                        self.expect_op(")")
                        return ("tuple", items)
                    self.expect_op(")")
                    return ("paren", first)
                finally:
                    self.no_range = saved_nr
                    self.nl_skip.pop()
                    self.in_bracket.pop()
                    self.in_index.pop()
            if v == "[":
                return self.parse_array_literal()
            if v == ":":
                return ("colon",)
            if v == "$":
                return ("unop", "$", self.parse_postfix())
        self.err("unexpected token", t)

    def parse_if(self):
        clauses = []
        self.nl_skip.append(False)
        self.in_bracket.append(False)
        self.in_index.append(False)
        cond = self.parse_expr(1)
        self.nl_skip.pop(); self.in_bracket.pop(); self.in_index.pop()
        body = self.parse_block()
        clauses.append((cond, body))
        els = None
        while True:
            t = self.next()
            if t.kind == "kw" and t.val == "elseif":
                self.nl_skip.append(False)
                self.in_bracket.append(False)
                self.in_index.append(False)
                cond = self.parse_expr(1)
                self.nl_skip.pop(); self.in_bracket.pop(); self.in_index.pop()
                clauses.append((cond, self.parse_block()))
            elif t.kind == "kw" and t.val == "else":
                els = self.parse_block()
            elif t.kind == "kw" and t.val == "end":
                break
            else:
                self.err("expected elseif/else/end", t)
        return ("if", clauses, els)

    def parse_for(self):
        self.nl_skip.append(False)
        self.in_bracket.append(False)
        self.in_index.append(False)
        var = self.parse_binary(CMP_PREC + 1)
        t = self.next()
        if not ((t.kind == "op" and t.val == "=") or (t.kind == "kw" and t.val == "in")):
            self.err("expected = or in after the loop variable", t)
        it = self.parse_expr(1)
        loops = [(var, it)]
        while self.is_op(","):          # for a in A, b in B, …: nested loops, the first iterator outermost
            self.next()
            v2 = self.parse_binary(CMP_PREC + 1)
            t = self.next()
            if not ((t.kind == "op" and t.val == "=") or (t.kind == "kw" and t.val == "in")):
                self.err("expected = or in after the loop variable", t)
            loops.append((v2, self.parse_expr(1)))
        self.nl_skip.pop(); self.in_bracket.pop(); self.in_index.pop()
        body = self.parse_block()
        self.expect_kw("end")
        node = None
        for v, itx in reversed(loops):
            node = ("for", v, itx, body if node is None else ("block", [node]))
        return node

    def parse_function(self):
        # function name(args; kw) ... end   |   function (f::T)(args) ... end
        self.nl_skip.append(False)
        self.in_bracket.append(False)
        self.in_index.append(False)
        sig = self.parse_postfix()
        self.nl_skip.pop(); self.in_bracket.pop(); self.in_index.pop()
        if sig[0] != "call":
            self.err("unsupported function signature")
        body = self.parse_block()
        self.expect_kw("end")
        callee, args, kwargs = sig[1], sig[2], sig[3]
        params = [self.to_param(a) for a in args]
        kwparams = [self.to_param(k) for k in kwargs]
        if callee[0] == "name":
            return ("function", callee[1], None, params, kwparams, body)
        if callee[0] == "paren" and callee[1][0] == "decl":
            d = callee[1]
            return ("function", None, (d[1], d[2]), params, kwparams, body)
        if callee[0] == "decl":
            return ("function", None, (callee[1], callee[2]), params, kwparams, body)
        if callee[0] == "curly":
            return ("function", callee[1][1], None, params, kwparams, body)
        self.err("unsupported function signature")

    def parse_macro(self, t: Tok):
        name = t.val
        if name in ("inline", "inbounds", "simd", "fastmath", "views", "noinline", "propagate_inbounds", "eval", "everywhere"):
            return self.parse_statement() if not self.nl_skip[-1] else self.parse_expr(0)
        if name in ("__DIR__", "__FILE__"):
            return ("macrocall", name, [])
        nt = self.toks[self.pos]
        if nt.kind == "op" and nt.val == "(" and not nt.sp_before:
            args, _ = self.parse_call_args()
            return ("macrocall", name, args)
        if name == "add_arg_table!":
            settings = self.parse_postfix()
            self.expect_kw("begin")
            body = self.parse_block()
            self.expect_kw("end")
            return ("macrocall", name, [settings, body])
        if name == "show":   # @show expr — an assignment or a top-level tuple is one argument
            return ("macrocall", name, [self.parse_expr_stmt()])
        # space-separated arguments up to the end of the statement; commas make a tuple
        args = []
        while True:
            nt = self.toks[self.pos]
            if nt.kind in ("nl", "eof") or (nt.kind == "op" and nt.val in (";", ")", "]")) or (nt.kind == "kw" and nt.val in BLOCK_END):
                break
            args.append(self.parse_expr(1))
            if self.toks[self.pos].kind == "op" and self.toks[self.pos].val == ",":
                self.pos += 1
        return ("macrocall", name, args)


def parse(src: str, filename="<string>"):
    return Parser(lex(src), filename).parse_program()


def parse_expression(src: str):
    p = Parser(lex(src), "<Meta.parse>")
    p.nl_skip = [True]
    return p.parse_expr(0)
