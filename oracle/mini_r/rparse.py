"""Tokenizer + Pratt parser for the subset of the R language the reference's R/*.R files are written in.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/mini_r/__init__.py). The grammar follows the R Language Definition,
section 10 ("Parser"): operator precedence and associativity as listed there, newlines terminate an expression unless
it is syntactically incomplete (open parenthesis / bracket, trailing binary operator), `else` may follow a newline
inside braces, `function(formals) body`, `if` / `for` / `while` / `repeat`, calls and the three index forms.

AST nodes are tuples:
  ("num", float) ("int", int) ("str", s) ("sym", name) ("null",) ("const", value)      leaves
  ("call", fn_ast, [(name|None, ast|None)...])      fn_ast is usually ("sym", name); a missing argument has ast None
  ("index", obj, args, double)                       x[...] (double False) or x[[...]] (double True)
  ("dollar", obj, name)
  ("binop", op, lhs, rhs)  ("unop", op, operand)
  ("assign", target, value, superassign)
  ("function", [(name, default_ast|None)...], body)
  ("block", [ast...])  ("if", cond, yes, no|None)  ("for", var, seq, body)  ("while", cond, body)
  ("repeat", body)  ("break",)  ("next",)  ("ns", pkg, name)  ("formula", lhs|None, rhs)
"""
from __future__ import annotations

import re

_TOKEN_RE = re.compile(r"""
    (?P<ws>[ \t\r\f]+)
  | (?P<comment>\#[^\n]*)
  | (?P<nl>\n)
  | (?P<num>(?:0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)L?)
  | (?P<str>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])*')
  | (?P<bt>`[^`]*`)
  | (?P<id>(?:[A-Za-z]|\.(?![0-9]))[A-Za-z0-9._]*|\.)
  | (?P<op>%[^%\n]*%|<<-|->>|:::|<-|->|<=|>=|==|!=|&&|\|\||::|\|>|[-+*/^<>=!&|~?:$@,;(){}\[\]])
""", re.X)

_ESC = {"n": "\n", "t": "\t", "\\": "\\", '"': '"', "'": "'", "0": "\0", "r": "\r"}


def _unescape(s):
    out, i = [], 0
    while i < len(s):
        if s[i] == "\\" and i + 1 < len(s):
            out.append(_ESC.get(s[i + 1], s[i + 1]))
            i += 2
        else:
            out.append(s[i])
            i += 1
    return "".join(out)


class RSyntaxError(SyntaxError):
    pass


def tokenize(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = _TOKEN_RE.match(src, pos)
        if not m:
            raise RSyntaxError("unexpected character %r at line %d" % (src[pos], line))
        pos = m.end()
        kind = m.lastgroup
        text = m.group()
        if kind in ("ws", "comment"):
            continue
        if kind == "nl":
            toks.append(("nl", "\n", line))
            line += 1
            continue
        if kind == "str":
            toks.append(("str", _unescape(text[1:-1]), line))
            line += text.count("\n")
        elif kind == "bt":
            toks.append(("id", text[1:-1], line))
        elif kind == "num":
            toks.append(("num", text, line))
        elif kind == "id":
            toks.append(("id", text, line))
        else:
            toks.append(("op", text, line))
    toks.append(("eof", "", line))
    return toks


# binding powers (R Language Definition 10.4.2, lowest to highest)
_BINARY = {
    "?": (1, 2),
    "=": (5, 4),                        # right
    "<-": (10, 9), "<<-": (10, 9),      # right
    "->": (12, 13), "->>": (12, 13),
    "~": (15, 16),
    "||": (20, 21), "|": (20, 21),
    "&&": (25, 26), "&": (25, 26),
    "==": (35, 36), "!=": (35, 36), "<": (35, 36), ">": (35, 36), "<=": (35, 36), ">=": (35, 36),
    "+": (40, 41), "-": (40, 41),
    "*": (45, 46), "/": (45, 46),
    "|>": (50, 51),
    ":": (55, 56),
    "^": (65, 64),                      # right
}
_SPECIAL_BP = (50, 51)                  # %any%
_UNARY = {"-": 60, "+": 60, "!": 30, "~": 15, "?": 1}
_POSTFIX_BP = 80
_KEYWORD_CONST = {"TRUE": True, "FALSE": False, "T": True, "F": False, "NA": float("nan"), "NA_real_": float("nan"),
                  "NA_integer_": float("nan"), "NA_character_": None, "Inf": float("inf"), "NaN": float("nan")}


class Parser:
    def __init__(self, src):
        self.toks = tokenize(src)
        self.i = 0
        self.depth = 0          # > 0 inside ( or [ : newlines are plain whitespace

    # ---- token helpers
    def peek(self, skip_nl=False):
        j = self.i
        if skip_nl or self.depth > 0:
            while self.toks[j][0] == "nl":
                j += 1
        return self.toks[j]

    def next(self, skip_nl=False):
        if skip_nl or self.depth > 0:
            while self.toks[self.i][0] == "nl":
                self.i += 1
        t = self.toks[self.i]
        self.i += 1
        return t

    def skip_newlines(self):
        while self.toks[self.i][0] == "nl" or (self.toks[self.i][0] == "op" and self.toks[self.i][1] == ";"):
            self.i += 1

    def expect(self, text, skip_nl=False):
        t = self.next(skip_nl)
        if t[1] != text or t[0] not in ("op", "id"):
            raise RSyntaxError("expected %r, got %r at line %d" % (text, t[1], t[2]))
        return t

    def at_op(self, text, skip_nl=False):
        t = self.peek(skip_nl)
        return t[0] == "op" and t[1] == text

    # ---- program
    def parse_program(self):
        out = []
        self.skip_newlines()
        while self.peek()[0] != "eof":
            out.append(self.parse_expr(0))
            t = self.peek()
            if t[0] not in ("nl", "eof") and not (t[0] == "op" and t[1] == ";"):
                raise RSyntaxError("unexpected %r at line %d" % (t[1], t[2]))
            self.skip_newlines()
        return out

    # ---- expressions
    def parse_expr(self, rbp):
        left = self.parse_prefix()
        while True:
            t = self.peek()
            if t[0] == "op":
                op = t[1]
                if op in ("(", "[", "$", "@") and _POSTFIX_BP > rbp:
                    left = self.parse_postfix(left)
                    continue
                if op in ("::", ":::"):
                    self.next()
                    name = self.next(True)
                    left = ("ns", left[1], name[1])
                    continue
                bp = _SPECIAL_BP if (op.startswith("%") and len(op) > 1) else _BINARY.get(op)
                if bp is None or bp[0] <= rbp:
                    break
                self.next()
                self.skip_only_newlines()
                right = self.parse_expr(bp[1])
                left = self.make_binary(op, left, right)
                continue
            break
        return left

    def skip_only_newlines(self):
        while self.toks[self.i][0] == "nl":
            self.i += 1

    def make_binary(self, op, left, right):
        if op in ("<-", "<<-", "="):
            return ("assign", left, right, op == "<<-")
        if op in ("->", "->>"):
            return ("assign", right, left, op == "->>")
        if op == "~":
            return ("formula", left, right)
        if op == "|>":
            if right[0] != "call":
                raise RSyntaxError("the pipe needs a call on its right-hand side")
            return ("call", right[1], [(None, left)] + right[2])
        return ("binop", op, left, right)

    def parse_prefix(self):
        t = self.next()
        kind, text, line = t
        if kind == "num":
            if text.endswith("L"):
                return ("int", int(text[:-1], 0))
            if text.lower().startswith("0x"):
                return ("num", float(int(text, 16)))
            return ("num", float(text))
        if kind == "str":
            return ("str", text)
        if kind == "id":
            if text == "function":
                return self.parse_function()
            if text == "if":
                return self.parse_if()
            if text == "for":
                return self.parse_for()
            if text == "while":
                self.expect("(")
                self.depth += 1
                cond = self.parse_expr(0)
                self.depth -= 1
                self.expect(")", True)
                self.skip_only_newlines()
                return ("while", cond, self.parse_expr(6))
            if text == "repeat":
                self.skip_only_newlines()
                return ("repeat", self.parse_expr(6))
            if text == "break":
                return ("break",)
            if text == "next":
                return ("next",)
            if text == "NULL":
                return ("null",)
            if text in _KEYWORD_CONST:
                return ("const", _KEYWORD_CONST[text])
            return ("sym", text)
        if kind == "op":
            if text == "(":
                self.depth += 1
                e = self.parse_expr(0)
                self.depth -= 1
                self.expect(")", True)
                return ("call", ("sym", "("), [(None, e)])
            if text == "{":
                return self.parse_block()
            if text in _UNARY:
                operand = self.parse_expr(_UNARY[text])
                if text == "~":
                    return ("formula", None, operand)
                return ("unop", text, operand)
        raise RSyntaxError("unexpected %r at line %d" % (text, line))

    def parse_block(self):
        saved, self.depth = self.depth, 0
        out = []
        self.skip_newlines()
        while not self.at_op("}"):
            if self.peek()[0] == "eof":
                raise RSyntaxError("unterminated block")
            out.append(self.parse_expr(0))
            t = self.peek()
            if not (t[0] == "nl" or (t[0] == "op" and t[1] in (";", "}"))):
                raise RSyntaxError("unexpected %r at line %d" % (t[1], t[2]))
            self.skip_newlines()
        self.expect("}")
        self.depth = saved
        return ("block", out)

    def parse_function(self):
        self.expect("(")
        self.depth += 1
        formals = []
        while not self.at_op(")"):
            name = self.next()
            if name[0] not in ("id", "str") and not (name[0] == "op" and name[1] == "..."):
                raise RSyntaxError("bad formal argument %r at line %d" % (name[1], name[2]))
            default = None
            if self.at_op("="):
                self.next()
                default = self.parse_expr(6)
            formals.append((name[1], default))
            if self.at_op(","):
                self.next()
        self.depth -= 1
        self.expect(")", True)
        self.skip_only_newlines()
        body = self.parse_expr(6)
        return ("function", formals, body)

    def parse_if(self):
        self.expect("(")
        self.depth += 1
        cond = self.parse_expr(0)
        self.depth -= 1
        self.expect(")", True)
        self.skip_only_newlines()
        yes = self.parse_expr(6)
        no = None
        # `else` may follow newlines (inside braces R allows it; at top level the reference never relies on the
        # difference)
        j = self.i
        while self.toks[j][0] == "nl":
            j += 1
        if self.toks[j][0] == "id" and self.toks[j][1] == "else":
            self.i = j + 1
            self.skip_only_newlines()
            no = self.parse_expr(6)
        return ("if", cond, yes, no)

    def parse_for(self):
        self.expect("(")
        self.depth += 1
        var = self.next()
        self.expect("in")
        seq = self.parse_expr(0)
        self.depth -= 1
        self.expect(")", True)
        self.skip_only_newlines()
        return ("for", var[1], seq, self.parse_expr(6))

    def parse_args(self, closer):
        """Arguments of a call or an index up to `closer` (")" or "]"); returns [(name|None, ast|None)]."""
        args = []
        self.depth += 1
        if self.at_op(closer):
            self.depth -= 1
            return args
        while True:
            if self.at_op(",") or self.at_op(closer):
                args.append((None, None))                      # empty argument: x[i, ]
            else:
                t = self.peek()
                t2 = None
                if t[0] in ("id", "str"):
                    j = self.i
                    while self.toks[j][0] == "nl":
                        j += 1
                    k = j + 1
                    while self.toks[k][0] == "nl":
                        k += 1
                    t2 = self.toks[k]
                if t2 is not None and t2[0] == "op" and t2[1] == "=" and t[1] not in ("function", "if"):
                    self.next()
                    self.next()
                    if self.at_op(",") or self.at_op(closer):
                        args.append((t[1], None))
                    else:
                        args.append((t[1], self.parse_expr(6)))
                else:
                    args.append((None, self.parse_expr(6)))
            if self.at_op(","):
                self.next()
                if self.at_op(closer):
                    args.append((None, None))
                    break
                continue
            break
        self.depth -= 1
        return args

    def parse_postfix(self, left):
        t = self.next()
        op = t[1]
        if op == "(":
            args = self.parse_args(")")
            self.expect(")", True)
            return ("call", left, args)
        if op == "[":
            double = False
            nt = self.toks[self.i]
            if nt[0] == "op" and nt[1] == "[":
                self.i += 1
                double = True
            args = self.parse_args("]")
            self.expect("]", True)
            if double:
                self.expect("]", True)
            return ("index", left, args, double)
        if op in ("$", "@"):
            name = self.next(True)
            if name[0] in ("id", "str"):
                return ("dollar", left, name[1])
            if name[0] == "op" and name[1] == "(":
                self.depth += 1
                e = self.parse_expr(0)
                self.depth -= 1
                self.expect(")", True)
                return ("dollar_expr", left, e)
            raise RSyntaxError("bad name after $ at line %d" % name[2])
        raise RSyntaxError("unexpected postfix %r" % op)


def parse(src):
    return Parser(src).parse_program()
