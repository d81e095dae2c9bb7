"""Lexer of minijl — a small interpreter for the subset of Julia the reference's MCMC path is written in.

Test infrastructure only (tests/golden/make_ref_fixtures.py): it lets the UNMODIFIED reference sources under
/root/reference (inc/eap_chain.jl, inc/energy.jl, inc/acceptance.jl, inc/average.jl, mcmc_eap_chain.jl,
mcmc_clustering_eap_chain.jl) be executed in a container that has no Julia, so that the oracle can be pinned to
outputs of the reference itself.  Nothing in the product path imports it.
"""
from __future__ import annotations

import unicodedata

KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "begin", "return", "break", "continue",
            "struct", "mutable", "abstract", "type", "const", "global", "local", "using", "import", "true", "false",
            "in", "let", "do", "quote", "macro", "module", "export", "try", "catch", "finally", "primitive", "where",
            "isa"}

# longest first
OPERATORS = ["...", "===", "!==", "&&", "||", "==", "!=", "<=", ">=", "->", "::", "<:", ">:", "+=", "-=", "*=", "/=", "^=", "%=",
             ".+", ".-", ".*", "./", ".^", ".=", "=>", "|>", "<<", ">>", "÷=",
             "+", "-", "*", "/", "^", "%", "<", ">", "=", "!", "?", ":", ",", ";", "(", ")", "[", "]", "{", "}", ".",
             "&", "|", "\\", "÷", "'", "$", "≤", "≥", "≠"]


class Tok:
    __slots__ = ("kind", "val", "sp_before", "sp_after", "line", "parts")

    def __init__(self, kind, val, sp_before, line, parts=None):
        self.kind, self.val, self.sp_before, self.line, self.parts = kind, val, sp_before, line, parts
        self.sp_after = False

    def __repr__(self):
        return f"Tok({self.kind},{self.val!r},L{self.line})"


class JlSyntaxError(Exception):
    pass


def is_id_start(ch: str) -> bool:
    if ch == "_" or ch.isalpha():
        return True
    if ord(ch) < 128:
        return False
    cat = unicodedata.category(ch)
    return cat in ("Lu", "Ll", "Lt", "Lm", "Lo", "Nl", "Sc", "So") or ch in "∇∂∞√"


def is_id_char(ch: str) -> bool:
    if is_id_start(ch) or ch.isdigit() or ch == "!":
        return True
    if ord(ch) < 128:
        return False
    return unicodedata.category(ch) in ("Mn", "Mc", "Nd", "Pc", "Sk", "Me", "No") or ch in "′″‴"


def _scan_string(src: str, i: int, line: int, term: str = '"'):
    """src[i] is the opening quote (or backtick: term = "`").  Returns (parts, next index, line) — parts: str pieces and
    ('expr', text)."""
    assert src[i] == term
    i += 1
    parts, buf = [], []
    n = len(src)
    while True:
        if i >= n:
            raise JlSyntaxError(f"unterminated string (line {line})")
        ch = src[i]
        if ch == term:
            i += 1
            break
        if ch == "\\":
            nx = src[i + 1]
            buf.append({"n": "\n", "t": "\t", "r": "\r", "\\": "\\", '"': '"', "$": "$", "0": "\0", "'": "'"}.get(nx, "\\" + nx))
            i += 2
            continue
        if ch == "$":
            if buf:
                parts.append("".join(buf))
                buf = []
            if src[i + 1] == "(":
                depth, j = 0, i + 1
                while True:
                    c = src[j]
                    if c == '"':  # nested string literal inside the interpolation
                        _, j, line = _scan_string(src, j, line)
                        continue
                    if c == "(":
                        depth += 1
                    elif c == ")":
                        depth -= 1
                        if depth == 0:
                            break
                    j += 1
                parts.append(("expr", src[i + 2:j]))
                i = j + 1
            else:
                j = i + 1
                while j < n and is_id_char(src[j]) and src[j] != "!":
                    j += 1
                parts.append(("expr", src[i + 1:j]))
                i = j
            continue
        if ch == "\n":
            line += 1
        buf.append(ch)
        i += 1
    if buf or not parts:
        parts.append("".join(buf))
    return parts, i, line


def lex(src: str):
    src = src.replace("µ", "μ")  # Julia normalises the micro sign to Greek mu
    toks = []
    i, n, line = 0, len(src), 1
    sp = True

    def push(kind, val, parts=None):
        nonlocal sp
        t = Tok(kind, val, sp, line, parts)
        toks.append(t)
        sp = False
        return t

    def prev_is_value():
        if not toks:
            return False
        t = toks[-1]
        if t.kind in ("id", "num", "str", "char"):
            return True
        if t.kind == "kw" and t.val in ("end", "true", "false"):
            return True
        return t.kind == "op" and t.val in (")", "]", "}", "'")

    while i < n:
        ch = src[i]
        if ch == "\n":
            if toks:
                toks[-1].sp_after = True
            push("nl", "\n")
            line += 1
            i += 1
            sp = True
            continue
        if ch in " \t\r":
            if toks:
                toks[-1].sp_after = True
            sp = True
            i += 1
            continue
        if ch == "#":
            if src.startswith("#=", i):
                j = src.index("=#", i) + 2
                line += src.count("\n", i, j)
                i = j
            else:
                while i < n and src[i] != "\n":
                    i += 1
            if toks:
                toks[-1].sp_after = True
            sp = True
            continue
        if ch == '"':
            parts, j, line2 = _scan_string(src, i, line)
            push("str", None, parts)
            line = line2
            i = j
            continue
        if ch == "`":      # command literal: interpolation as in a string; split into words when it is run
            parts, j, line2 = _scan_string(src, i, line, "`")
            push("cmd", None, parts)
            line = line2
            i = j
            continue
        if ch.isdigit() or (ch == "." and i + 1 < n and src[i + 1].isdigit() and not prev_is_value()):
            if ch == "0" and i + 1 < n and src[i + 1] in "xX":   # hexadecimal literal (an unsigned integer in Julia)
                j = i + 2
                while j < n and (src[j] in "0123456789abcdefABCDEF_"):
                    j += 1
                push("num", int(src[i + 2:j].replace("_", ""), 16))
                i = j
                continue
            j = i
            while j < n and (src[j].isdigit() or src[j] == "_"):
                j += 1
            isfloat = False
            if j < n and src[j] == "." and not src.startswith("..", j) and not (j + 1 < n and (is_id_start(src[j + 1]) or src[j + 1] in "*/^+-=(")):
                isfloat = True
                j += 1
                while j < n and src[j].isdigit():
                    j += 1
            if j < n and src[j] in "eE" and (src[j + 1].isdigit() or (src[j + 1] in "+-" and src[j + 2].isdigit())):
                isfloat = True
                j += 2
                while j < n and src[j].isdigit():
                    j += 1
            text = src[i:j].replace("_", "")
            push("num", float(text) if isfloat else int(text))
            i = j
            continue
        if ch == "@":
            j = i + 1
            while j < n and (is_id_char(src[j]) or src[j] == "."):
                j += 1
            push("macro", src[i + 1:j])
            i = j
            continue
        if ch == "'" and not (prev_is_value() and not sp):
            # character literal
            if src[i + 1] == "\\":
                c = {"n": "\n", "t": "\t", "\\": "\\", "'": "'"}.get(src[i + 2], src[i + 2])
                j = i + 3
            else:
                c = src[i + 1]
                j = i + 2
            if src[j] != "'":
                raise JlSyntaxError(f"bad character literal (line {line})")
            push("char", c)
            i = j + 1
            continue
        if is_id_start(ch):
            j = i + 1
            while j < n and is_id_char(src[j]):
                # `!` belongs to the name only in `name!` (not in `a != b`)
                if src[j] == "!" and j + 1 < n and src[j + 1] == "=" and not (j + 2 < n and src[j + 2] == "="):
                    break
                j += 1
            word = src[i:j]
            if word in KEYWORDS and not (toks and toks[-1].kind == "op" and toks[-1].val == "." and not toks[-1].sp_before):
                push("kw", word)
            else:
                push("id", word)
            i = j
            continue
        if ch == ":" and i + 1 < n and (is_id_start(src[i + 1])) and not prev_is_value() and not src.startswith("::", i):
            j = i + 1
            while j < n and is_id_char(src[j]):
                j += 1
            push("sym", src[i + 1:j])
            i = j
            continue
        for op in OPERATORS:
            if src.startswith(op, i):
                # `.` followed by a digit was handled above; `.5` after a value is field access + number (never here)
                push("op", op)
                i += len(op)
                break
        else:
            raise JlSyntaxError(f"unexpected character {ch!r} (line {line})")
    push("nl", "\n")
    push("eof", None)
    return toks
