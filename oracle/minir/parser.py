"""TEST INFRASTRUCTURE ONLY — tokenizer and parser for the subset of the R language that the reference uses.

The reference (SDEtools) is R source and GNU R is not installed in the build image.  So that parity can be pinned to the
reference's OWN SOURCE TEXT rather than to a restatement of it, `oracle/minir` evaluates `R/sdetools.R` (and the fixture recipe
`tests/golden/make_reference_fixtures.R`, and the reference's `tests/testthat/test_solvers.R`) directly.  This file is the
front end: R's grammar with R's operator precedence and associativity (R-lang §10.4.2; `?Syntax`), which is exactly what a
restatement can get wrong — `-a*b`, `0.5*(gX+gY) %*% dB` (`%*%` binds tighter than `*`), left-associative `+`.

AST: tuples.  ("num", float, is_int) ("str", s) ("id", name) ("null",) ("call", fn, [(name|None, expr|None), ...])
("index", obj, args, double) ("dollar", obj, name) ("ns", pkg, name) ("function", [(name, default|None)], body, src)
("block", [exprs]) ("if", c, a, b|None) ("for", var, seq, body) ("while", c, body) ("repeat", body) ("break",) ("next",)
("assign", target, value, superassign)
Binary and unary operators are calls to the function of that name: ("call", ("id", "+"), [(None, l), (None, r)]).
"""
import re

_TOKEN = re.compile(r"""
    (?P<ws>[ \t\r\f]+)
  | (?P<comment>\#[^\n]*)
  | (?P<nl>\n)
  | (?P<num>0[xX][0-9a-fA-F]+L? | (?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?[Li]?)
  | (?P<str>"(?:\\.|[^"\\])*" | '(?:\\.|[^'\\])*')
  | (?P<bq>`[^`]*`)
  | (?P<id>(?:[A-Za-z]|\.(?![0-9]))[A-Za-z0-9._]* | \.\.\. | \.)
  | (?P<op><<-|->>|<-|->|<=|>=|==|!=|&&|\|\||\|>|::|%[^%\n]*%|[-+*/^<>=!&|~?:$@])
  | (?P<punct>\[\[|[()\[\]{},;])
""", re.X)

_ESC = {"n": "\n", "t": "\t", "\\": "\\", '"': '"', "'": "'", "0": "\0", "r": "\r"}


def tokenize(src):
    out, pos, line = [], 0, 1
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"minir: cannot tokenize at line {line}: {src[pos:pos + 30]!r}")
        pos = m.end()
        k = m.lastgroup
        if k in ("ws", "comment"):
            continue
        t = m.group()
        if k == "nl":
            out.append(("nl", "\n", line))
            line += 1
        elif k == "num":
            is_int = t.endswith("L")
            if is_int:
                t = t[:-1]
            v = float(int(t, 16)) if t[:2] in ("0x", "0X") else float(t)
            out.append(("num", (v, is_int), line))
        elif k == "str":
            out.append(("str", re.sub(r"\\(.)", lambda mm: _ESC.get(mm.group(1), mm.group(1)), t[1:-1]), line))
        elif k == "bq":
            out.append(("id", t[1:-1], line))
        elif k == "id":
            out.append(("id", t, line))
        else:
            out.append((k, t, line))
    out.append(("eof", "", line))
    return out


KEYWORDS = {"if", "else", "for", "while", "repeat", "function", "break", "next", "TRUE", "FALSE", "NULL", "NA", "NA_real_",
            "NA_integer_", "NA_character_", "Inf", "NaN", "in"}

# binary operators: (left binding power, right binding power); higher binds tighter.  R-lang 10.4.2, lowest to highest:
#  ?  |  = (right)  |  <- <<- (right)  |  -> ->> (left)  |  ~  |  || |  |  && &  |  !  |  comparison  |  + -  |  * /  |
#  %any% |>  |  :  |  unary + -  |  ^ (right)  |  $ @ [[ [ ( ::
BINARY = {
    "?": (1, 2), "=": (4, 3), "<-": (6, 5), "<<-": (6, 5), "->": (7, 8), "->>": (7, 8), "~": (9, 10),
    "||": (11, 12), "|": (11, 12), "&&": (13, 14), "&": (13, 14),
    "==": (17, 18), "!=": (17, 18), "<": (17, 18), ">": (17, 18), "<=": (17, 18), ">=": (17, 18),
    "+": (19, 20), "-": (19, 20), "*": (21, 22), "/": (21, 22), "|>": (23, 24), ":": (25, 26), "^": (30, 29),
}
BP_NOT, BP_UNARY_MINUS, BP_SPECIAL = 15, 27, (23, 24)


class Parser:
    def __init__(self, src):
        self.src = src
        self.tok = tokenize(src)
        self.i = 0
        self.skipnl = [False]                      # inside ( ) and [ ] newlines are white space; inside { } they end statements

    def peek(self):
        if self.skipnl[-1]:
            while self.tok[self.i][0] == "nl":
                self.i += 1
        return self.tok[self.i]

    def next(self):
        t = self.peek()
        self.i += 1
        return t

    def skip_newlines(self):
        while self.tok[self.i][0] == "nl":
            self.i += 1

    def at(self, kind, val=None):
        t = self.peek()
        return t[0] == kind and (val is None or t[1] == val)

    def expect(self, kind, val=None):
        t = self.next()
        if t[0] != kind or (val is not None and t[1] != val):
            raise SyntaxError(f"minir: line {t[2]}: expected {val or kind}, found {t[1]!r}")
        return t

    def program(self):
        out = []
        while True:
            while self.at("nl") or self.at("punct", ";"):
                self.next()
            if self.at("eof"):
                return out
            out.append(self.expr(0))
            if not (self.at("nl") or self.at("punct", ";") or self.at("eof")):
                t = self.peek()
                raise SyntaxError(f"minir: line {t[2]}: unexpected {t[1]!r}")

    def expr(self, minbp):
        left = self.prefix()
        while True:
            t = self.peek()
            if t[0] != "op":
                break
            op = t[1]
            if op.startswith("%") and op.endswith("%") and len(op) >= 2:
                lbp, rbp = BP_SPECIAL
            elif op in BINARY:
                lbp, rbp = BINARY[op]
            else:
                break
            if lbp < minbp:
                break
            self.next()
            self.skip_newlines()                   # a binary operator at the end of a line continues the expression
            right = self.expr(rbp)
            if op in ("<-", "<<-", "="):
                left = ("assign", left, right, op == "<<-")
            elif op in ("->", "->>"):
                left = ("assign", right, left, op == "->>")
            else:
                left = ("call", ("id", op), [(None, left), (None, right)])
        return left

    def prefix(self):
        t = self.peek()
        if t[0] == "op" and t[1] in ("-", "+"):
            self.next()
            return self.postfix_loop(("call", ("id", t[1]), [(None, self.expr(BP_UNARY_MINUS))]), allow=False)
        if t[0] == "op" and t[1] == "!":
            self.next()
            return ("call", ("id", "!"), [(None, self.expr(BP_NOT))])
        if t[0] == "op" and t[1] == "~":
            self.next()
            return ("call", ("id", "~"), [(None, self.expr(10))])
        return self.postfix_loop(self.primary())

    def postfix_loop(self, e, allow=True):
        while allow:
            t = self.peek()
            if t[0] == "punct" and t[1] == "(":
                e = ("call", e, self.arglist("(", ")"))
            elif t[0] == "punct" and t[1] == "[[":
                self.next()
                self.skipnl.append(True)
                args = self.args_until("]")
                self.expect("punct", "]")
                self.skipnl.pop()
                self.expect("punct", "]")
                e = ("index", e, args, True)
            elif t[0] == "punct" and t[1] == "[":
                e = ("index", e, self.arglist("[", "]"), False)
            elif t[0] == "op" and t[1] in ("$", "@"):
                self.next()
                n = self.next()
                if n[0] not in ("id", "str"):
                    raise SyntaxError(f"minir: line {n[2]}: name expected after $")
                e = ("dollar", e, n[1])
            elif t[0] == "op" and t[1] == "::":
                self.next()
                n = self.expect("id")
                e = ("ns", e[1], n[1])
            else:
                break
        return e

    def arglist(self, open_, close):
        self.expect("punct", open_)
        self.skipnl.append(True)
        args = self.args_until(close)
        self.skipnl.pop()
        self.expect("punct", close)
        return args

    def args_until(self, close):
        args = []
        if self.at("punct", close):
            return args
        while True:
            if self.at("punct", ",") or self.at("punct", close):
                args.append((None, None))          # an empty argument: X[i, ]
            else:
                t = self.peek()
                nxt = self.tok[self.i + 1]
                if t[0] in ("id", "str") and nxt[0] == "op" and nxt[1] == "=":
                    self.next()
                    self.next()
                    if self.at("punct", ",") or self.at("punct", close):
                        args.append((t[1], None))
                    else:
                        args.append((t[1], self.expr(5)))
                else:
                    args.append((None, self.expr(5)))
            if self.at("punct", ","):
                self.next()
                if self.at("punct", close):
                    args.append((None, None))
                    return args
                continue
            return args

    def primary(self):
        t = self.next()
        k, v = t[0], t[1]
        if k == "num":
            return ("num", v[0], v[1])
        if k == "str":
            return ("str", v)
        if k == "punct" and v == "(":
            self.skipnl.append(True)
            e = self.expr(0)
            self.skipnl.pop()
            self.expect("punct", ")")
            return ("call", ("id", "("), [(None, e)])
        if k == "punct" and v == "{":
            self.skipnl.append(False)
            out = []
            while True:
                while self.at("nl") or self.at("punct", ";"):
                    self.next()
                if self.at("punct", "}"):
                    break
                out.append(self.expr(0))
                if not (self.at("nl") or self.at("punct", ";") or self.at("punct", "}")):
                    tt = self.peek()
                    raise SyntaxError(f"minir: line {tt[2]}: unexpected {tt[1]!r} in block")
            self.skipnl.pop()
            self.expect("punct", "}")
            return ("block", out)
        if k == "id":
            if v == "function":
                start = self.tok[self.i - 1][2]
                self.expect("punct", "(")
                self.skipnl.append(True)
                params = []
                while not self.at("punct", ")"):
                    n = self.expect("id")
                    d = None
                    if self.at("op", "="):
                        self.next()
                        d = self.expr(5)
                    params.append((n[1], d))
                    if self.at("punct", ","):
                        self.next()
                self.skipnl.pop()
                self.expect("punct", ")")
                self.skip_newlines()
                body = self.expr(3)
                return ("function", params, body, start)
            if v == "if":
                self.expect("punct", "(")
                self.skipnl.append(True)
                c = self.expr(0)
                self.skipnl.pop()
                self.expect("punct", ")")
                self.skip_newlines()
                a = self.expr(3)
                b = None
                save = self.i
                # R accepts a newline before `else` only inside braces; every use in the evaluated sources is of that kind
                self.skip_newlines()
                if self.tok[self.i][0] == "id" and self.tok[self.i][1] == "else":
                    self.i += 1
                    self.skip_newlines()
                    b = self.expr(3)
                else:
                    self.i = save
                return ("if", c, a, b)
            if v == "for":
                self.expect("punct", "(")
                self.skipnl.append(True)
                var = self.expect("id")[1]
                self.expect("id", "in")
                seq = self.expr(0)
                self.skipnl.pop()
                self.expect("punct", ")")
                self.skip_newlines()
                return ("for", var, seq, self.expr(3))
            if v == "while":
                self.expect("punct", "(")
                self.skipnl.append(True)
                c = self.expr(0)
                self.skipnl.pop()
                self.expect("punct", ")")
                self.skip_newlines()
                return ("while", c, self.expr(3))
            if v == "repeat":
                self.skip_newlines()
                return ("repeat", self.expr(3))
            if v == "break":
                return ("break",)
            if v == "next":
                return ("next",)
            if v == "NULL":
                return ("null",)
            if v in ("TRUE", "T"):
                return ("lgl", True) if v == "TRUE" else ("id", v)
            if v == "FALSE":
                return ("lgl", False)
            if v in ("NA", "NA_real_", "NA_integer_", "NA_character_"):
                return ("na", v)
            if v == "Inf":
                return ("num", float("inf"), False)
            if v == "NaN":
                return ("num", float("nan"), False)
            return ("id", v)
        raise SyntaxError(f"minir: line {t[2]}: unexpected {v!r}")


def parse(src):
    return Parser(src).program()
