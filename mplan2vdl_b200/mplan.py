"""Python restatement of mplan2vdl's FRONT half: the mplan lexer / parser (Scanner.x, Parser.y) and the relational
IR builder (Mplan.hs), producing the IR classes vlite.py lowers.  With it the reference's own fixtures
(tests/tpch10noorder/NN.sql.mplan) become Voodoo programs without a Haskell toolchain:

    python -m mplan2vdl_b200.mplan /root/reference/tests/tpch10noorder/06.sql.mplan | python -m mplan2vdl_b200 --csv

Scope = what vlite.py lowers (Table / Select / GroupBy / Project / plain FK Join with range, equality and arithmetic
scalars); everything else raises NotImplementedError with the construct's name, like the reference's own `error`
calls.  Each function cites what it follows.  Validation: the programs generated from 01/03/05/06.sql.mplan equal
plans/q01,q03,q05,q06.vdl (tests/test_mplan_front.py; q06 is pinned by the reference README).
"""
from __future__ import annotations

import datetime
import re
import sys

from .vlite import Bin, CartesianProduct, Cast, GroupBy, Identity, IfThenElse, In, Join, Like, Lit, Project, Ref, Select, Table, Unary

DATE = ("date",)

# ------------------------------------------------------------------------------------------ lexer (Scanner.x:18-45)
_TOKEN = re.compile(r"""
    (?P<ws>[\s|]+)                                   # `|` is whitespace (Scanner.x:27)
  | (?P<str>"[A-Za-z0-9<>=!_%\-\ \#]*")              # value literals (Scanner.x:35)
  | (?P<multi>NOT\ NULL|no\ nil|!=)                  # multi-word keywords, `!=` before `!` (Scanner.x:41-45)
  | (?P<num>[0-9]+(?![A-Za-z0-9<>=!_%]))             # numbers (only inside type specs)
  | (?P<word>[A-Za-z0-9<>=!_%]+)                     # names: sys.>= , %TID% , l_quantity ... (Scanner.x:22-24)
  | (?P<punct>[\[\](),.;])
""", re.X)


def strip_comments(text: str) -> str:
    """MainFuns.hs:83-92: lines starting with # % -- [ are blanked."""
    out = []
    for line in text.splitlines():
        s = line.lstrip()
        out.append("" if s.startswith(("#", "%", "--", "[")) else line)
    return "\n".join(out)


def lex(text: str) -> list:
    toks, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError(f"mplan: cannot scan at {text[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        toks.append(("word" if kind == "multi" else kind, m.group()))
    return toks


# ------------------------------------------------------------------------------------------ parser (Parser.y:67-212)
ATTRS = {"NOT NULL", "ASC", "HASHCOL", "JOINIDX", "HASHIDX", "FETCH"}
INFIX = {"<": "Lt", ">": "Gt", "<=": "Leq", ">=": "Geq", "=": "Eq", "!=": "Neq", "or": "LogOr"}        # Mplan.hs:71-81
BINFUN = {"sql_add": "Add", "sql_sub": "Sub", "sql_mul": "Mul", "sql_div": "Div", "sql_min": "Min", "sql_max": "Max",
          "=": "Eq", "or": "LogOr", "and": "LogAnd", ">": "Gt", "<>": "Neq", "scale_down": "Div"}      # Mplan.hs:84-99


class P:
    """Parse-tree node kinds (Parser.y:230-284), as tuples:  ("ref", name, attrs) ("call", fname, args) ("cast", tspec, expr)
    ("lit", tspec, string) ("infix", op, l, r) ("interval", a, op1, m, op2, b) ("nested", [exprs]); an Expr is (node, alias)."""


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def take(self, value=None, kind=None):
        tok = self.peek()
        if (value is not None and tok[1] != value) or (kind is not None and tok[0] != kind):
            raise SyntaxError(f"mplan: expected {value or kind}, got {tok[1]!r} (token {self.i})")
        self.i += 1
        return tok

    def at(self, value):
        return self.peek()[1] == value and self.peek()[0] in ("word", "punct")

    # Tree: Leaf | Node
    def tree(self):
        if self.at("table"):
            self.take("table"); self.take("(")
            source = self.qname()
            self.take(")"); self.take("[")
            cols = self.expr_list("]")
            self.take("]"); self.take("COUNT")
            return ("leaf", source, cols)
        words = []
        while self.peek()[0] == "word" and not self.at("("):
            words.append(self.take()[1])
        self.take("(")
        children = [self.tree()]
        while self.at(","):
            self.take(",")
            children.append(self.tree())
        self.take(")")
        lists = []
        while self.at("["):
            self.take("[")
            lists.append(self.expr_list("]"))
            self.take("]")
        return ("node", " ".join(words), children, lists)

    def qname(self):
        parts = [self.take(kind="word")[1]]
        while self.at(".") and self.peek(1)[0] == "word":
            self.take(".")
            parts.append(self.take(kind="word")[1])
        if parts[0] == "sys" and len(parts) > 1:          # dropsys (Parser.y:100)
            parts = parts[1:]
        return ".".join(parts)

    def expr_list(self, closer):
        out = []
        if self.at(closer):
            return out
        out.append(self.expr())
        while self.at(","):
            self.take(",")
            out.append(self.expr())
        return out

    def expr(self):                                        # ExprNoComma (Parser.y:134-146)
        first = self.expr_bind()
        if self.peek()[0] == "word" and self.peek()[1] in INFIX:
            op1 = self.take()[1]
            mid = self.expr_bind()
            if self.peek()[0] == "word" and self.peek()[1] in INFIX:
                op2 = self.take()[1]
                last = self.expr_bind()
                return (("interval", first, op1, mid, op2, last), None)
            return (("infix", op1, first, mid), None)
        return first

    def expr_bind(self):                                   # BasicExpr [as QualifiedName]
        e = self.basic()
        alias = None
        if self.at("as"):
            self.take("as")
            alias = self.qname()
        if self.at("in"):                                  # InExpr (Parser.y:207-209)
            self.take("in"); self.take("(")
            items = self.expr_list(")")
            self.take(")")
            return (("in", (e, alias), items), None)
        if self.peek()[1] in ("FILTER", "!") and self.peek()[0] == "word":       # FilterExpr (Parser.y:202-206)
            negated = self.at("!")
            if negated:
                self.take("!")
            self.take("FILTER")
            oper = self.take(kind="word")[1]
            self.take("(")
            pattern = self.expr()
            self.take(",")
            escape = self.basic()
            self.take(")")
            return (("filter", oper, negated, (e, alias), pattern, escape), None)
        if self.at("notin"):                               # Parser.y:212; rejected by Mplan.hs:522
            raise NotImplementedError("implement this case of IN operator (Mplan.hs:522): notin")
        return (e, alias)

    def attrs(self):
        out = []
        while self.peek()[0] == "word" and self.peek()[1] in ATTRS:
            a = self.take()[1]
            out.append(("JOINIDX", self.qname()) if a == "JOINIDX" else (a, None))
        return out

    def basic(self):                                       # BasicExprBare (Parser.y:176-192)
        if self.at("("):
            self.take("(")
            es = self.expr_list(")")
            self.take(")")
            return ("nested", es)
        # TypeSpec '[' Expr ']'  |  TypeSpec literal   (a type spec is an identifier with optional (n, m))
        save = self.i
        if self.peek()[0] == "word":
            tname = self.take()[1]
            tparams = []
            ok = True
            if self.at("(") and self.peek(1)[0] == "num":
                self.take("(")
                tparams.append(int(self.take(kind="num")[1]))
                while self.at(","):
                    self.take(",")
                    tparams.append(int(self.take(kind="num")[1]))
                if self.at(")"):
                    self.take(")")
                else:
                    ok = False
            if ok and self.peek()[0] == "str":
                return ("lit", (tname, tparams), self.take()[1][1:-1])
            if ok and self.at("["):
                self.take("[")
                inner = self.expr()
                self.take("]")
                return ("cast", (tname, tparams), inner)
        self.i = save
        name = self.qname()
        if self.at("no nil"):
            self.take("no nil")
        if self.at("("):
            self.take("(")
            args = self.expr_list(")")
            self.take(")")
            self.attrs()
            return ("call", name, args)
        return ("ref", name, self.attrs())


def parse(text: str):
    p = Parser(lex(strip_comments(text)))
    tree = p.tree()
    if p.peek()[0] != "eof":
        raise SyntaxError(f"mplan: trailing input at token {p.i}: {p.peek()[1]!r}")
    return tree


# ------------------------------------------------------------------------------------------ Mplan.hs: scalars
def day_count(datestr: str) -> int:                        # Mplan.hs:46-57: days since 0000-01-01
    y, m, d = (int(x) for x in datestr.split("-"))
    return datetime.date(y, m, d).toordinal() + 365


def add_months_rollover(date: datetime.date, months: int) -> datetime.date:      # addGregorianMonthsRollOver
    y, m = divmod(date.year * 12 + (date.month - 1) + months, 12)
    m += 1
    first = datetime.date(y, m, 1)
    return first + datetime.timedelta(days=date.day - 1)   # a day past the month's end rolls over into the next month


class Front:
    def __init__(self, catalog, cross_product=False):
        self.cat = catalog
        self.cross_product = cross_product                  # Config.cross_product (Config.hs:150, 223)

    def _dtype_of_ref(self, name: str):
        """display type of a column reference, for typing char literals (Mplan.hs:441-446, 489-493)"""
        try:
            col = self.cat.column(name)
        except Exception:
            return None
        return ("str", name) if col.mtype.startswith(("char", "varchar")) else None

    def sc(self, node, ctx=None):
        kind = node[0]
        if kind == "ref":
            return Ref(node[1])
        if kind == "lit":                                   # Mplan.hs:461-484
            (tname, tparams), s = node[1], node[2]
            if tname == "date":
                return Lit(DATE, day_count(s))
            if tname == "decimal":
                return Lit(("dec", tparams[1]), int(s))
            if tname == "boolean":
                return Lit(("dec", 0), {"true": 1, "false": 0}[s])
            if tname in ("tinyint", "smallint", "int", "bigint"):
                return Lit(("dec", 0), int(s))
            if tname in ("char", "varchar"):
                if not ctx:
                    raise ValueError(f"need more information to assign type to char literal {s!r} (Mplan.hs:482)")
                return Lit(ctx, self.cat.dictionary[ctx[1]][s])
            raise NotImplementedError(f"mplan literal of type {tname}")
        if kind == "cast":                                  # Mplan.hs:453-459; only decimal casts change the value (Vlite.hs:939-956)
            (tname, tparams), (inner, _alias) = node[1], node[2]
            v = self.sc(inner, ctx)
            return Cast(tparams[1] if tname == "decimal" else None, v)
        if kind == "nested":
            return self.conjunction(node[1])
        if kind == "infix":                                 # Mplan.hs:486-496
            l = self.sc(node[2][0])
            newctx = self._dtype_of_ref(l.name) if isinstance(l, Ref) else None
            r = self.sc(node[3][0], newctx)
            return Bin(INFIX[node[1]], l, r)
        if kind == "interval":                              # Mplan.hs:498-512
            a, m, b = self.sc(node[1][0]), self.sc(node[3][0]), self.sc(node[5][0])
            return Bin("LogAnd", Bin(INFIX[node[2]], a, m), Bin(INFIX[node[4]], m, b))
        if kind == "in":                                    # Mplan.hs:514-522 (column IN literal list only)
            (lnode, _), items = node[1], node[2]
            if lnode[0] != "ref":
                raise NotImplementedError("implement this case of IN operator (Mplan.hs:522)")
            newctx = self._dtype_of_ref(lnode[1])
            return In(Ref(lnode[1]), [self.sc(e, newctx) for e, _ in items])
        if kind == "filter":                                # Mplan.hs:524-545: x [!] FILTER like (char[char(n) "pattern"], char "")
            _, oper, negated, (arg, _a), (pat, palias), escape = node
            ok = (oper == "like" and palias is None and pat[0] == "cast" and pat[1] == ("char", []) and pat[2][0][0] == "lit" and
                  pat[2][0][1][0] == "char" and len(pat[2][0][1][1]) == 1 and escape[0] == "lit" and escape[1] == ("char", []) and escape[2] == "")
            if not ok:
                raise NotImplementedError(f"mplan filter operator {oper} (Mplan.hs:547)")
            like = Like(self.sc(arg, ctx), pat[2][0][2])
            return Unary("Neg", like) if negated else like
        if kind == "call":
            fname, args = node[1], node[2]
            base = fname.split(".")[-1]
            if base == "ifthenelse" and len(args) == 3:         # Mplan.hs:441-451
                return IfThenElse(*(self.sc(a, ctx) for a, _ in args))
            if base == "like" and len(args) == 2:              # Mplan.hs:398-417: sys.like(x, char[char(n) "pattern"])
                pat = args[1][0]
                if not (pat[0] == "cast" and pat[1] == ("char", []) and pat[2][0][0] == "lit" and pat[2][0][1][0] == "char" and len(pat[2][0][1][1]) == 1):
                    raise NotImplementedError("implement this 'like' case (Mplan.hs:419)")
                return Like(self.sc(args[0][0], ctx), pat[2][0][2])
            if len(args) == 1 and base == "identity":          # Mplan.hs:392-396: a row id, whatever the argument
                return Identity()
            if len(args) == 1 and base in ("year", "sql_neg", "isnull"):      # Mplan.hs:106-112, 420-424
                return Unary({"year": "Year", "sql_neg": "Neg", "isnull": "IsNull"}[base], self.sc(args[0][0], ctx))
            if len(args) == 2:
                (x, _), (y, _) = args
                # date +/- interval folded into a date literal (Mplan.hs:368-388)
                if base in ("sql_add", "sql_sub") and x[0] == "lit" and x[1][0] == "date" and y[0] == "lit" and y[1][0] in ("month_interval", "sec_interval"):
                    y0, m0, d0 = (int(v) for v in x[2].split("-"))
                    date, num = datetime.date(y0, m0, d0), int(y[2]) * (-1 if base == "sql_sub" else 1)
                    if y[1][0] == "month_interval":
                        out = add_months_rollover(date, num)
                    else:
                        q = abs(num) // (1000 * 60 * 60 * 24)
                        out = date + datetime.timedelta(days=q if num >= 0 else -q)     # Haskell `quot` truncates toward zero
                    return Lit(DATE, out.toordinal() + 365)
                if base not in BINFUN:
                    raise NotImplementedError(f"mplan binary function {fname} (Mplan.hs:99)")
                l = self.sc(x)
                newctx = self._dtype_of_ref(l.name) if isinstance(l, Ref) else None
                return Bin(BINFUN[base], l, self.sc(y, newctx))
            raise NotImplementedError(f"mplan scalar function {fname}/{len(args)} (like, identity ... are outside the executor's scope)")
        raise NotImplementedError(f"mplan scalar {kind}")

    def conjunction(self, exprs):                           # Mplan.hs:549-559
        solved = [self.sc(e) for e, _ in exprs]
        if not solved:
            raise ValueError("empty conjunction list")
        out = solved[0]
        for e in solved[1:]:
            out = Bin("LogAnd", out, e)
        return out

    # -------------------------------------------------------------------------------------- Mplan.hs: relations
    def group_output(self, e):                              # solveGroupOutput (Mplan.hs:138-176)
        node, alias = e
        if node[0] == "ref":
            return (("FChoose", Ref(node[1])), alias)
        if node[0] == "call":
            base, args = node[1].split(".")[-1], node[2]
            if base == "count" and (not args or args[0][0][0] == "ref"):
                return (("Count",), alias)
            if len(args) == 1:
                inner = self.sc(args[0][0])
                op = {"sum": "FSum", "avg": "Avg", "max": "FMax", "min": "FMin"}.get(base)
                if op:
                    return ((op, inner), alias)
        raise NotImplementedError(f"group_by output expression {node[0]}")

    def solve(self, t):                                     # solve (Mplan.hs:227-332)
        if t[0] == "leaf":
            cols = []
            for node, alias in t[2]:
                if node[0] != "ref":
                    raise ValueError("table outputs should only have reference expressions")
                fk = [a[1] for a in node[2] if a[0] == "JOINIDX"]
                if len(fk) > 1:
                    raise ValueError("multiple fkey indices")
                if fk:
                    cols.append((fk[0], alias if alias is not None else node[1]))      # notice the reversal (Mplan.hs:240-251)
                else:
                    cols.append((node[1], alias))
            return Table(t[1], cols)
        _, relop, children, lists = t
        if relop == "project":
            if len(children) != 1 or len(lists) < 1:
                raise ValueError("project: one child, at least one output list")
            if len(lists) > 1 and lists[1]:
                raise NotImplementedError("order-by clauses (Mplan.hs:267-269)")
            return Project(self.solve(children[0]), [(self.sc(e), alias) for e, alias in lists[0]])
        if relop == "group by":
            keys, values = lists
            inputkeys = []
            for node, alias in keys:
                if node[0] != "ref":
                    raise ValueError("non-ref in group by key")
                inputkeys.append((node[1], alias))
            return GroupBy(self.solve(children[0]), inputkeys, [self.group_output(e) for e in values])
        if relop == "select":
            return Select(self.solve(children[0]), self.conjunction(lists[0]))
        if relop in ("join", "semijoin", "antijoin", "left outer join"):
            variant = {"join": "Plain", "semijoin": "LeftSemi", "antijoin": "LeftAnti", "left outer join": "LeftOuter"}[relop]   # classify_join (Mplan.hs:334-356)
            l, r = children
            if self.cross_product and relop == "join":      # --use_cross_product (Mplan.hs:309-313): plain joins only
                return Select(CartesianProduct(self.solve(l), self.solve(r)), self.conjunction(lists[0]))
            return Join(self.solve(l), self.solve(r), [self.sc(e) for e, _ in lists[0]], variant)
        raise NotImplementedError(f"relational operator {relop!r} (Mplan.hs:332)")


def relexpr_from_mplan(catalog, text: str, cross_product=False):
    """mplanFromParseTree (Mplan.hs:567-568) after Parser.fromString (Parser.y:301-304)."""
    return Front(catalog, cross_product).solve(parse(text))


def translate_mplan(catalog, text: str, agg_strategy="serial", cross_product=False, goffset=0, apply_cleanup_passes=True) -> str:
    """The whole translator (MainFuns.compile, 172-188) for the supported subset: mplan text -> Voodoo program text."""
    from . import vlite
    return vlite.translate(catalog, relexpr_from_mplan(catalog, text, cross_product), agg_strategy, goffset, apply_cleanup_passes)


if __name__ == "__main__":
    import argparse
    from .meta import builtin_catalog
    ap = argparse.ArgumentParser(prog="python -m mplan2vdl_b200.mplan", description="mplan -> Voodoo program (the reference's tpchrun)")
    ap.add_argument("mplanfile", nargs="?", default="-")
    g = ap.add_mutually_exclusive_group()                  # MainFuns.hs:61-65
    g.add_argument("--aggserial", action="store_true")
    g.add_argument("--agghierarchical", action="store_true")
    g.add_argument("--aggshuffle", action="store_true")
    ap.add_argument("-b", "--boundsfile", help="(table,col,min,max,count,trailing zeros) csv; with -t -s --dictionary replaces the built-in SF10 catalogue")
    ap.add_argument("-t", "--storagefile", help="output of 'select * from storage' in csv format")
    ap.add_argument("-s", "--schemafile", help="output of msqldump -D -d <dbname>")
    ap.add_argument("--dictionary", dest="dictionaryfile", help="dictionary to encode literal strings")
    ap.add_argument("--goffset", type=int, default=0, help="offset for synthesized group-by keys (MainFuns.hs:67)")
    ap.add_argument("-c", "--apply_cleanup_passes", type=lambda v: v.lower() not in ("false", "0", "no"), default=True,
                    help="after generating vdl identify and clean up known no-op patterns (default true)")
    ap.add_argument("--use_cross_product", action="store_true", help="plain joins as a selection over the cross product (MainFuns.hs:72)")
    ap.add_argument("-g", "--grainsize", type=int, default=8192, help="power of 2; only with --agghierarchical")
    a = ap.parse_args()
    if a.grainsize < 1 or a.grainsize & (a.grainsize - 1):
        sys.exit("grainsize must be a power of 2 (MainFuns.hs:112)")
    strategy = ("hierarchical", a.grainsize.bit_length() - 1) if a.agghierarchical else ("shuffle" if a.aggshuffle else "serial")
    src = sys.stdin.read() if a.mplanfile == "-" else open(a.mplanfile).read()
    files = (a.boundsfile, a.storagefile, a.schemafile, a.dictionaryfile)
    if any(files) and not all(files):
        sys.exit("usage: need a column bounds csv, a storage file, a schema file and a dictionary file, or none of them (MainFuns.hs:101-112)")
    from .meta import load_metadata_files
    catalog = load_metadata_files(*files) if all(files) else builtin_catalog()
    sys.stdout.write(translate_mplan(catalog, src, strategy, a.use_cross_product, a.goffset, a.apply_cleanup_passes))
