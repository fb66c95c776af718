"""A second opinion on random queries: the relational IR (tests/fuzz_plans.py builds it; mplan2vdl_b200/vlite.py lowers it
to a Voodoo program) evaluated DIRECTLY with numpy, relation by relation -- no Voodoo ops, no key packing, no
Scatter/Gather join lowering.  The plan interpreters (CPU oracle, CUDA paths) must give the same columns in the same order.

Semantics followed: selections keep row order; an FK join keeps the fact side's rows whose dimension row survives, in fact
order (Vlite.hs:1199-1209); a GROUP BY returns one row per distinct key tuple in ascending lexicographic order, first key
major (the packed key of makeCompositeKey, Vlite.hs:1123-1170, is monotone in that order) and no row at all for an
empty input; COUNT / AVG are FoldSum of 1 and integer Divide (Vlite.hs:1038-1046); arithmetic is int64."""
import numpy as np

from mplan2vdl_b200.vlite import Bin, Cast, GroupBy, IfThenElse, In, Join, Lit, Project, Ref, Select, Table

I64 = np.int64


class Frame:
    def __init__(self, base, rows, cols):
        self.base, self.rows, self.cols = base, rows, cols          # cols: [(name, array aligned with rows)], in order

    def get(self, name):
        for n, a in self.cols:
            if n == name:
                return a
        hits = [a for n, a in self.cols if n.endswith("." + name) or n.split(".")[0] == name]     # suffix lookup (Name.hs:94-112)
        if not hits:
            raise KeyError(name)
        return hits[0]


def tdiv(a, b):
    q = np.abs(a) // np.maximum(np.abs(b), 1)
    return np.where(b == 0, 0, np.where((a < 0) != (b < 0), -q, q)).astype(I64)


def expr(f: Frame, e):
    n = len(f.rows)
    if isinstance(e, Ref):
        return f.get(e.name)
    if isinstance(e, Lit):
        return np.full(n, e.n, I64)
    if isinstance(e, Cast):
        if e.point is None:
            return expr(f, e.arg)
        assert isinstance(e.arg, Lit) and e.arg.dtype[0] == "dec"           # the only representation-changing cast the fuzzer draws
        return np.full(n, e.arg.n * 10 ** (e.point - e.arg.dtype[1]), I64)
    if isinstance(e, In):
        return np.isin(expr(f, e.left), [x.n for x in e.set]).astype(I64)
    if isinstance(e, IfThenElse):
        return np.where(expr(f, e.if_) != 0, expr(f, e.then_), expr(f, e.else_)).astype(I64)
    assert isinstance(e, Bin), e
    a, b = expr(f, e.left), expr(f, e.right)
    with np.errstate(over="ignore"):
        return {"Add": lambda: a + b, "Sub": lambda: a - b, "Mul": lambda: a * b,
                "Lt": lambda: a < b, "Leq": lambda: a <= b, "Gt": lambda: a > b, "Geq": lambda: a >= b,
                "Eq": lambda: a == b, "Neq": lambda: a != b,
                "LogAnd": lambda: (a != 0) & (b != 0), "LogOr": lambda: (a != 0) | (b != 0)}[e.op]().astype(I64)


def rel(data: dict, r) -> Frame:
    if isinstance(r, Table):
        nrows = next(len(v) for k, v in data.items() if k.startswith(r.name + "."))
        cols = []
        for col, alias in r.columns:
            if col.endswith("%TID%"):
                cols.append((alias or col, np.arange(nrows, dtype=I64)))
            elif col in data:                       # (columns the query never uses are not among the plan's Loads)
                cols.append((alias or col, np.asarray(data[col], dtype=I64)))
        return Frame(r.name, np.arange(nrows, dtype=I64), cols)
    if isinstance(r, Select):
        f = rel(data, r.child)
        keep = expr(f, r.predicate) != 0
        return Frame(f.base, f.rows[keep], [(n, a[keep]) for n, a in f.cols])
    if isinstance(r, Join):
        l, rt = rel(data, r.left), rel(data, r.right)
        (c,) = r.conds
        names = (c.left.name, c.right.name)
        fkname = next(n for n in names if "%TID%" not in n)
        fact, dim = (l, rt) if any(n == fkname for n, _ in l.cols) else (rt, l)
        fk = fact.get(fkname)                                          # dimension base row of every fact row
        nbase = int(max(dim.rows.max(initial=-1), fk.max(initial=-1))) + 1
        pos = np.full(nbase, -1, I64)
        pos[dim.rows] = np.arange(len(dim.rows), dtype=I64)
        keep = pos[fk] >= 0
        at = pos[fk[keep]]
        return Frame(fact.base, fact.rows[keep], [(n, a[keep]) for n, a in fact.cols] + [(n, a[at]) for n, a in dim.cols])
    if isinstance(r, GroupBy):
        f = rel(data, r.child)
        n = len(f.rows)

        def outname(agg, alias):
            return alias if alias is not None else (agg[1].name if agg[0] == "FChoose" and isinstance(agg[1], Ref) else "?")
        if n == 0:
            return Frame(None, np.zeros(0, I64), [(outname(agg, alias), np.zeros(0, I64)) for agg, alias in r.outputaggs])
        keys = [f.get(k) for k, _ in r.inputkeys]
        if keys:
            order = np.lexsort(keys[::-1])                             # first key major, stable
            sk = [k[order] for k in keys]
            head = np.ones(n, bool)
            head[1:] = np.any([k[1:] != k[:-1] for k in sk], axis=0)
            gid = np.cumsum(head) - 1
        else:
            order, gid, head = np.arange(n), np.zeros(n, I64), np.arange(n) == 0
        ng = int(gid[-1]) + 1
        cnt = np.bincount(gid, minlength=ng).astype(I64)

        def fold(op, v):
            v = v[order]
            if op == "FSum":
                out = np.zeros(ng, I64)
                np.add.at(out, gid, v)
                return out
            out = np.full(ng, np.iinfo(I64).max if op == "FMin" else np.iinfo(I64).min, I64)
            (np.minimum if op == "FMin" else np.maximum).at(out, gid, v)
            return out
        cols = []
        for agg, alias in r.outputaggs:
            if agg[0] == "Count":
                v = cnt
            elif agg[0] == "Avg":
                v = tdiv(fold("FSum", expr(f, agg[1])), cnt)
            elif agg[0] == "FChoose":
                v = expr(f, agg[1])[order][head]
            else:
                v = fold(agg[0], expr(f, agg[1]))
            cols.append((outname(agg, alias), v))
        return Frame(None, np.arange(ng, dtype=I64), cols)
    assert isinstance(r, Project), r
    f = rel(data, r.child)
    return Frame(f.base, f.rows, [(alias or (e.name if isinstance(e, Ref) else "?"), expr(f, e)) for e, alias in r.projectout])


def evaluate(data: dict, query) -> list:
    """The query's output columns, in order."""
    return [a for _, a in rel(data, query).cols]
