"""A second opinion on random queries: the relational IR (tests/fuzz_plans.py builds it; mplan2vdl_b200/vlite.py lowers it
to a Voodoo program) evaluated DIRECTLY with numpy, relation by relation -- no Voodoo ops, no key packing, no
Scatter/Gather join lowering.  The plan interpreters (CPU oracle, CUDA paths) must give the same columns in the same order.

Semantics followed: selections keep row order; an FK join is evaluated BY VALUE (fact key = dimension key, the dimension
side being the one whose base column is unique) and keeps the fact side's rows whose dimension row survives, in fact
order (Vlite.hs:1199-1209); a semijoin keeps the LEFT side's rows that have a partner, an antijoin / a dim-side semijoin
additionally reproduce the two places where the reference's graph is not the SQL (see `join`); a join against a single
value broadcasts it (Vlite.hs:691-713); a GROUP BY returns one row per distinct key tuple in ascending lexicographic order, first key
major (the packed key of makeCompositeKey, Vlite.hs:1123-1170, is monotone in that order) and no row at all for an
empty input; COUNT / AVG are FoldSum of 1 and integer Divide (Vlite.hs:1038-1046); arithmetic is int64."""
import numpy as np

from mplan2vdl_b200.vlite import Bin, Cast, GroupBy, Identity, IfThenElse, In, Join, Like, Lit, Project, Ref, Select, Table, Unary

I64 = np.int64


class Frame:
    def __init__(self, base, rows, cols, origin=None):
        self.base, self.rows, self.cols = base, rows, cols          # cols: [(name, array aligned with rows)], in order
        self.origin = origin or {}                                  # column name -> base column it is a copy of (lineage)

    def index(self, name):
        for i, (n, _) in enumerate(self.cols):
            if n == name:
                return i
        hits = [i for i, (n, _) in enumerate(self.cols) if n.endswith("." + name) or n.split(".")[0] == name]     # suffix lookup (Name.hs:94-112)
        if not hits:
            raise KeyError(name)
        return hits[0]

    def get(self, name):
        return self.cols[self.index(name)][1]

    def origin_of(self, name):
        return self.origin.get(self.cols[self.index(name)][0])

    def take(self, keep):
        return Frame(self.base, self.rows[keep], [(n, a[keep]) for n, a in self.cols], self.origin)


def tdiv(a, b):
    q = np.abs(a) // np.maximum(np.abs(b), 1)
    return np.where(b == 0, 0, np.where((a < 0) != (b < 0), -q, q)).astype(I64)


def expr(f: Frame, e):
    n = len(f.rows)
    if isinstance(e, Ref):
        return f.get(e.name)
    if isinstance(e, Lit):
        return np.full(n, e.n, I64)
    if isinstance(e, Identity):
        return np.arange(n, dtype=I64)
    if isinstance(e, Like):                                                  # SQL LIKE as a regular expression over the decoded strings
        import re
        arg = e.arg
        while isinstance(arg, Cast):
            arg = arg.arg
        heap = HEAPS[f.origin_of(arg.name)]
        rx = re.compile("".join(".*" if c == "%" else ("." if c == "_" else re.escape(c)) for c in e.pattern), re.S)
        raw = heap.tobytes()
        memo = {}
        out = np.zeros(n, I64)
        for i, off in enumerate(expr(f, arg)):
            off = int(off)
            if off not in memo:
                memo[off] = 1 if rx.fullmatch(raw[off:raw.index(b"\0", off)].decode()) else 0
            out[i] = memo[off]
        return out
    if isinstance(e, Unary) and e.op == "Year":                              # the query's meaning: the calendar year (not the emitted approximation)
        import datetime
        days = expr(f, e.arg)
        u, inv = np.unique(days, return_inverse=True)
        return np.array([datetime.date.fromordinal(int(d) - 365).year for d in u], dtype=I64)[inv]
    if isinstance(e, Unary) and e.op == "Neg":                               # `!`: 1 - x (Vlite.hs:1016-1018)
        return (1 - expr(f, e.arg)).astype(I64)
    if isinstance(e, Cast):
        if e.point is None:
            return expr(f, e.arg)
        src = point_of(f, e.arg)
        v = expr(f, e.arg)
        if src is None or src == e.point:
            return v
        k = 10 ** abs(e.point - src)
        return (v * k if e.point > src else tdiv(v, np.full(n, k, I64))).astype(I64)
    if isinstance(e, In):
        def lit(x):
            while isinstance(x, Cast) and x.point is None:
                x = x.arg
            return x.n
        return np.isin(expr(f, e.left), [lit(x) for x in e.set]).astype(I64)
    if isinstance(e, IfThenElse) and isinstance(e.if_, Unary) and e.if_.op == "IsNull":
        return expr(f, e.else_)          # columns are statically NOT NULL: isnull(x) is false (Vlite.hs:996-1001)
    if isinstance(e, IfThenElse):
        return np.where(expr(f, e.if_) != 0, expr(f, e.then_), expr(f, e.else_)).astype(I64)
    assert isinstance(e, Bin), e
    a, b = expr(f, e.left), expr(f, e.right)
    with np.errstate(over="ignore"):
        return {"Add": lambda: a + b, "Sub": lambda: a - b, "Mul": lambda: a * b,
                "Lt": lambda: a < b, "Leq": lambda: a <= b, "Gt": lambda: a > b, "Geq": lambda: a >= b,
                "Eq": lambda: a == b, "Neq": lambda: a != b, "Div": lambda: tdiv(a, b),
                "Min": lambda: np.minimum(a, b), "Max": lambda: np.maximum(a, b),
                "LogAnd": lambda: (a != 0) & (b != 0), "LogOr": lambda: (a != 0) | (b != 0)}[e.op]().astype(I64)


HEAPS = {}       # base column -> its string heap (uint8), filled by evaluate() from the data's `<col>.heap` entries
POINTS = {}      # base column -> decimal point, filled by evaluate() from the catalogue when decimal casts matter


def point_of(f: Frame, e):
    """Decimal scale of a scalar expression (the reference tracks it as the display type; Vlite.hs:395-412, 939-956)."""
    if isinstance(e, Lit):
        return e.dtype[1] if e.dtype[0] == "dec" else None
    if isinstance(e, Ref):
        try:
            return f.points.get(f.cols[f.index(e.name)][0])
        except (KeyError, AttributeError):
            return None
    if isinstance(e, Cast):
        return e.point if e.point is not None else point_of(f, e.arg)
    if isinstance(e, IfThenElse):              # cond * then + !cond * else (Vlite.hs:237-245): the scale of `then`
        return point_of(f, e.then_) or 0
    if isinstance(e, Bin):
        a, b = point_of(f, e.left), point_of(f, e.right)
        if e.op == "Mul":
            return (a or 0) + (b or 0)
        if e.op == "Div":
            return (a or 0) - (b or 0)
        if e.op in ("Add", "Sub", "Min", "Max"):
            return a
        return 0
    return None


def join(data: dict, r) -> Frame:
    l, rt = rel(data, r.left), rel(data, r.right)
    eqs = [c for c in r.conds if isinstance(c, Bin) and c.op == "Eq" and isinstance(c.left, Ref) and isinstance(c.right, Ref)]
    if len(r.conds) == 1 and not _is_fk(data, l, rt, r.conds[0]):
        # a single value on one side: broadcast it, keep the other side's rows for which the condition holds
        (c,) = r.conds
        for mine, other, mine_left in ((l, rt, True), (rt, l, False)):
            if len(mine.cols) == 1 and len(mine.rows) <= 1:
                if len(mine.rows) == 0:
                    raise NotImplementedError("broadcast of an empty relation (the emitted Gather would fail)")
                both = Frame(other.base, other.rows, other.cols + [(mine.cols[0][0], np.full(len(other.rows), mine.cols[0][1][0], I64))], other.origin)
                both.points = dict(getattr(other, "points", {}), **getattr(mine, "points", {}))
                keep = expr(both, c) != 0
                out = other.take(keep)
                out.points = getattr(other, "points", {})
                return out
        raise NotImplementedError("join that is not a single complete FK")
    fk_conds = [c for c in eqs if _is_fk(data, l, rt, c)]
    extra = [c for c in r.conds if not any(c is q for q in fk_conds)]
    # which side is the dimension: the one whose key column is unique in the base data
    c0 = fk_conds[0]
    lname, rname = (c0.left.name, c0.right.name) if _has(l, c0.left.name) else (c0.right.name, c0.left.name)
    dim_is_right = _unique_origin(data, rt, rname)
    fact, dim = (l, rt) if dim_is_right else (rt, l)

    def keys(frame, left_side):
        cols = []
        for c in fk_conds:
            a, b = (c.left.name, c.right.name) if _has(l, c.left.name) else (c.right.name, c.left.name)
            cols.append(frame.get(a if left_side else b))
        return cols
    fkeys, dkeys = keys(fact, fact is l), keys(dim, dim is l)
    # value join on (possibly composite) keys: dim keys are unique
    def pack(cols):
        out = np.zeros(len(cols[0]), dtype=[("k%d" % i, I64) for i in range(len(cols))])
        for i, c in enumerate(cols):
            out["k%d" % i] = c
        return out
    dk, fkv = pack(dkeys), pack(fkeys)
    order = np.argsort(dk, kind="stable")
    sdk = dk[order]
    at = np.searchsorted(sdk, fkv)
    at = np.minimum(at, max(len(sdk) - 1, 0))
    hit = (sdk[at] == fkv) if len(sdk) else np.zeros(len(fkv), bool)
    dpos = order[at] if len(sdk) else at                 # position in dim of every fact row's partner (valid where hit)
    pts = dict(getattr(fact, "points", {}), **getattr(dim, "points", {}))
    if r.variant == "Plain":
        out = Frame(fact.base, fact.rows[hit], [(n, a[hit]) for n, a in fact.cols] + [(n, a[dpos[hit]]) for n, a in dim.cols],
                    dict(fact.origin, **dim.origin))
        out.points = pts
        for c in extra:                                   # 714-718: the other conditions select over the joined rows
            out2 = out.take(expr(out, c) != 0)
            out2.points = pts
            out = out2
        return out
    if extra:
        raise NotImplementedError("can only do this rewrite for plain joins")
    if r.variant == "LeftSemi":
        if fact is l:
            out = fact.take(hit)
        else:
            mark = np.zeros(len(dim.rows), bool)
            mark[dpos[hit]] = True
            # NOT the SQL: a fact row without a partner still scatters -- to slot 0 of the inverse index -- so the first
            # dimension row qualifies whenever one exists (the reference's graph, Vlite.hs:1214-1218, kept literally)
            if (~hit).any() and len(dim.rows):
                mark[0] = True
            out = dim.take(mark)
        out.points = getattr(l, "points", {})
        return out
    if r.variant == "LeftAnti":
        if fact is not l:
            raise NotImplementedError("anti join on the dimension side (Vlite.hs:1232)")
        # NOT the SQL (App. G13): the reference negates the POSITIONS of the matching rows instead of the boolean:
        # keep fact row j for every j < #matches whose j-th matching position is not 1
        m = np.nonzero(hit)[0]
        js = np.nonzero(m != 1)[0]
        out = fact.take(js)
        out.points = getattr(fact, "points", {})
        return out
    raise NotImplementedError(r.variant)


def _has(frame, name):
    try:
        frame.index(name)
        return True
    except KeyError:
        return False


def _unique_origin(data, frame, name):
    o = frame.origin_of(name)
    if o is None:
        return False
    if o.endswith("%TID%"):
        return True
    base = np.asarray(data[o])
    return len(np.unique(base)) == len(base)


def _is_fk(data, l, rt, c):
    if not (isinstance(c, Bin) and c.op == "Eq" and isinstance(c.left, Ref) and isinstance(c.right, Ref)):
        return False
    sides = [(l, c.left.name, rt, c.right.name), (l, c.right.name, rt, c.left.name)]
    for a, an, b, bn in sides:
        if _has(a, an) and _has(b, bn):
            oa, ob = a.origin_of(an), b.origin_of(bn)
            if oa is None or ob is None:
                return False
            return (_unique_origin(data, a, an) != _unique_origin(data, b, bn)) or oa.endswith("%TID%") or ob.endswith("%TID%") or oa == ob
    return False


def rel(data: dict, r) -> Frame:
    out = _rel(data, r)
    if not hasattr(out, "points"):
        out.points = {}
    return out


def _rel(data: dict, r) -> Frame:
    if isinstance(r, Table):
        nrows = next(len(v) for k, v in data.items() if k.startswith(r.name + "."))
        cols, origin, points = [], {}, {}
        for col, alias in r.columns:
            if col.endswith("%TID%"):
                cols.append((alias or col, np.arange(nrows, dtype=I64)))
            elif col in data:                       # (columns the query never uses are not among the plan's Loads)
                cols.append((alias or col, np.asarray(data[col], dtype=I64)))
            else:
                continue
            origin[alias or col] = col
            if col in POINTS:
                points[alias or col] = POINTS[col]
        out = Frame(r.name, np.arange(nrows, dtype=I64), cols, origin)
        out.points = points
        return out
    if isinstance(r, Select):
        f = rel(data, r.child)
        out = f.take(expr(f, r.predicate) != 0)
        out.points = f.points
        return out
    if isinstance(r, Join):
        return join(data, r)
    if isinstance(r, GroupBy):
        f = rel(data, r.child)
        n = len(f.rows)

        def outname(agg, alias):
            return alias if alias is not None else (agg[1].name if agg[0] == "FChoose" and isinstance(agg[1], Ref) else "?")
        origin, points = {}, {}
        for agg, alias in r.outputaggs:
            nm = outname(agg, alias)
            if agg[0] == "FChoose" and isinstance(agg[1], Ref) and _has(f, agg[1].name):
                origin[nm] = f.origin_of(agg[1].name)
            if agg[0] in ("FSum", "FMin", "FMax", "FChoose", "Avg"):
                pt = point_of(f, agg[1]) if not (isinstance(agg[1], Ref) and not _has(f, agg[1].name)) else None
                if pt is not None:
                    points[nm] = pt
        if n == 0:
            out = Frame(None, np.zeros(0, I64), [(outname(agg, alias), np.zeros(0, I64)) for agg, alias in r.outputaggs], origin)
            out.points = points
            return out
        # input keys may be aliased (ps_suppkey as L5.L5): the alias is in scope for the aggregates (Vlite.hs:626-634)
        extra = [(a, f.get(k)) for k, a in r.inputkeys if a is not None]
        if extra:
            f2 = Frame(f.base, f.rows, f.cols + extra, dict(f.origin, **{a: f.origin_of(k) for k, a in r.inputkeys if a is not None}))
            f2.points = f.points
            f = f2
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
            # a later aggregate may name an earlier output (sys.sum(...) as L1.L1, L1.L1 as L2.L2: Q11): such a reference is
            # already grouped (solveAgg, Vlite.hs:1050-1054)
            if agg[0] == "FChoose" and isinstance(agg[1], Ref) and any(nm == agg[1].name or nm.split(".")[0] == agg[1].name or nm.endswith("." + agg[1].name) for nm, _ in cols) and not _has(f, agg[1].name):
                prev = next(v for nm, v in cols if nm == agg[1].name or nm.split(".")[0] == agg[1].name or nm.endswith("." + agg[1].name))
                prevname = next(nm for nm, v in cols if nm == agg[1].name or nm.split(".")[0] == agg[1].name or nm.endswith("." + agg[1].name))
                cols.append((outname(agg, alias), prev))
                if prevname in points:
                    points[outname(agg, alias)] = points[prevname]
                continue
            if agg[0] == "Count":
                v = cnt
            elif agg[0] == "Avg":
                v = tdiv(fold("FSum", expr(f, agg[1])), cnt)
            elif agg[0] == "FChoose":
                v = expr(f, agg[1])[order][head]
            else:
                v = fold(agg[0], expr(f, agg[1]))
            cols.append((outname(agg, alias), v))
        out = Frame(None, np.arange(ng, dtype=I64), cols, origin)
        out.points = points
        return out
    assert isinstance(r, Project), r
    f = rel(data, r.child)
    cols, origin, points = [], {}, {}
    for e, alias in r.projectout:
        nm = alias or (e.name if isinstance(e, Ref) else "?")
        # later outputs may refer to earlier ones (Vlite.hs:595-597)
        scope = Frame(f.base, f.rows, f.cols + cols, dict(f.origin, **origin))
        scope.points = dict(f.points, **points)
        cols.append((nm, expr(scope, e)))
        if isinstance(e, Ref):
            origin[nm] = scope.origin_of(e.name)
        pt = point_of(scope, e)
        if pt is not None:
            points[nm] = pt
    out = Frame(f.base, f.rows, cols, origin)
    out.points = points
    return out


def set_catalog(cat):
    """Decimal scales of the base columns (schema.msqldump): needed where a query casts a computed decimal."""
    POINTS.clear()
    for t in cat.tables.values():
        for c in t.columns.values():
            if c.mtype == "decimal":
                POINTS[c.qualified] = c.scale


def evaluate(data: dict, query) -> list:
    """The query's output columns, in order."""
    HEAPS.clear()
    HEAPS.update({k[:-5]: v for k, v in data.items() if k.endswith(".heap")})
    return [a for _, a in rel(data, query).cols]


def base_columns(query) -> list:
    """Qualified names of every stored column the query's Table leaves mention (the plan itself may Load an FK index
    column instead of the key columns a value join needs)."""
    out = []

    def walk(r):
        if isinstance(r, Table):
            for col, _ in r.columns:
                if "%" not in col and col not in out:
                    out.append(col)
        for k in ("child", "left", "right"):
            if hasattr(r, k):
                walk(getattr(r, k))
    walk(query)
    return out
