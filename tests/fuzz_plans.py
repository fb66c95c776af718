"""Random relational IRs inside the feature set the translator restatement (mplan2vdl_b200/vlite.py) lowers:
Table / Select / GroupBy / Project / plain FK joins.  Each seed gives one query; `vlite.translate` turns it into the
Voodoo program the reference would print, which the parity tests run through the oracle, the fused GPU paths and
the op-at-a-time GPU path.  The point is to exercise the symbolic analyses (vdl_plan.cu `analyse`,
vdl_plan_join.inc `janalyse`) on shapes nobody wrote by hand."""
from __future__ import annotations

import random

from mplan2vdl_b200 import tpch_queries as Q
from mplan2vdl_b200.vlite import Bin, Cast, GroupBy, IfThenElse, In, Join, Lit, Project, Ref, Select, Table

DATE = ("date",)
# column -> (kind, lo, hi) used to draw literals inside the column's range (bounds.csv:59-79)
LI_PRED_COLS = {
    "lineitem.l_shipdate": ("date", 727564, 730089),
    "lineitem.l_discount": ("dec2", 0, 10),
    "lineitem.l_tax": ("dec2", 0, 8),
    "lineitem.l_quantity": ("dec2", 100, 5000),
    "lineitem.l_extendedprice": ("dec2", 90091, 10494950),
}
LI_DATES = ["lineitem.l_shipdate", "lineitem.l_commitdate", "lineitem.l_receiptdate"]
LI_KEYS = ["lineitem.l_returnflag", "lineitem.l_linestatus"]
LI_COLS = ("l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate", "l_commitdate", "l_receiptdate")
LI_VALS = ["lineitem.l_quantity", "lineitem.l_extendedprice", "lineitem.l_discount", "lineitem.l_tax"]


def lit_for(kind, v):
    return Lit(DATE, v) if kind == "date" else Lit(("dec", 2), v)


def range_pred(rng: random.Random, col: str, spec):
    kind, lo, hi = spec
    a, b = sorted(rng.randint(lo, hi) for _ in range(2))
    x = Ref(col)
    form = rng.randrange(5)
    if form == 0:
        return Bin("Leq", x, lit_for(kind, b))
    if form == 1:
        return Bin("Gt", x, lit_for(kind, a))
    if form == 2:
        return Bin("Lt", x, lit_for(kind, b))
    if form == 3:
        return Bin("Eq", x, lit_for(kind, rng.randint(lo, min(hi, lo + 20))))
    return Q.between(lit_for(kind, a), x, lit_for(kind, b), rng.choice(["Leq", "Lt"]), rng.choice(["Leq", "Lt"]))


def fancy_pred(rng: random.Random):
    """IN list / column-vs-column comparison / <> (what Q12 and Q19 use)."""
    form = rng.randrange(4)
    if form == 0:
        vals = rng.sample(range(0, 11), rng.randint(1, 3))
        return In(Ref("lineitem.l_discount"), [Lit(("dec", 2), v) for v in vals])
    if form == 1:
        a, b = rng.sample(LI_DATES, 2)
        return Bin(rng.choice(["Lt", "Leq", "Gt", "Geq", "Eq", "Neq"]), Ref(a), Ref(b))
    if form == 2:
        a, b, c = rng.sample(LI_DATES, 3)
        return Bin("LogAnd", Bin("Lt", Ref(a), Ref(b)), Bin("Lt", Ref(b), Ref(c)))         # a < b < c (Parser.y Interval)
    return Bin("Neq", Ref("lineitem.l_tax"), Lit(("dec", 2), rng.randint(0, 8)))


def case_expr(rng: random.Random):
    """sum(case when p then x else 0 end): ifthenelse, lowered to (1 - (p == 0)) * x + (p == 0) * 0 (Vlite.hs:237-245)."""
    c = rng.choice(sorted(LI_PRED_COLS))
    p = range_pred(rng, c, LI_PRED_COLS[c]) if rng.random() < 0.6 else fancy_pred(rng)
    then = Lit(("dec", 0), 1) if rng.random() < 0.5 else Ref(rng.choice(LI_VALS))
    return IfThenElse(p, then, Lit(("dec", 0), 0))


def value_expr(rng: random.Random):
    cols = rng.sample(LI_VALS, rng.randint(1, 2))
    e = Ref(cols[0])
    if len(cols) == 2:
        other = Ref(cols[1])
        if rng.random() < 0.6:
            other = Bin(rng.choice(["Sub", "Add"]), Q.ONE_2, other)          # (1 - d) / (1 + t) style
        e = Bin("Mul", e, other)
    return e


def aggs(rng: random.Random, keys):
    out = [(("FChoose", Ref(k)), None) for k in keys]
    names = []
    for i in range(rng.randint(1, 3)):
        kind = rng.choice(["FSum", "FSum", "Count", "Avg", "FMin", "FMax", "Case"])
        name = f"L{i + 1}"
        if kind == "Case":
            out.append((("FSum", case_expr(rng)), f"{name}.{name}"))
        elif kind == "Count":
            out.append((("Count",), f"{name}.{name}"))
        elif kind == "Avg":
            out.append((("Avg", Cast(None, Ref(rng.choice(LI_VALS)))), f"{name}.{name}"))
        elif kind in ("FMin", "FMax"):
            out.append(((kind, Ref(rng.choice(LI_VALS))), f"{name}.{name}"))
        else:
            out.append(((kind, value_expr(rng)), f"{name}.{name}"))
        names.append(name)
    return out, names


def single_table(seed: int):
    rng = random.Random(seed)
    keys = rng.sample(LI_KEYS, rng.randint(0, 2))
    preds = [range_pred(rng, c, LI_PRED_COLS[c]) for c in rng.sample(sorted(LI_PRED_COLS), rng.randint(0, 3))]
    if rng.random() < 0.4:
        preds.append(fancy_pred(rng))
    t = Table("lineitem", Q.li(*LI_COLS))
    child = Select(t, Q.conj(*preds)) if preds else t
    outaggs, names = aggs(rng, keys)
    g = GroupBy(child, [(k, None) for k in keys], outaggs)
    return Project(g, [(Ref(k), None) for k in keys] + [(Ref(n), f"{n}.out_{n.lower()}") for n in names])


def join_query(seed: int, catalog):
    """lineitem [sel] JOIN orders [sel] [JOIN customer [sel]] [JOIN supplier], grouped by a small key (or none)."""
    rng = random.Random(seed)
    lo, hi = 727563, 729968
    a, b = sorted(rng.randint(lo, hi) for _ in range(2))
    ocols = [("orders.o_orderdate", None), ("orders.o_shippriority", None), ("orders.%TID%", None), ("orders.orders_customer", "orders.%orders_customer")]
    orders = Table("orders", ocols)
    if rng.random() < 0.8:
        orders = Select(orders, Q.between(Lit(DATE, a), Ref("orders.o_orderdate"), Lit(DATE, b)))
    licols = Q.li(*LI_COLS) + [
        ("lineitem.lineitem_orders", "lineitem.%lineitem_orders"), ("lineitem.lineitem_supplier", "lineitem.%lineitem_supplier")]
    lineitem = Table("lineitem", licols)
    if rng.random() < 0.6:
        c = rng.choice(sorted(LI_PRED_COLS))
        pr = range_pred(rng, c, LI_PRED_COLS[c])
        if rng.random() < 0.4:
            pr = Bin("LogAnd", pr, fancy_pred(rng))
        lineitem = Select(lineitem, pr)
    if rng.random() < 0.5:
        j = Join(lineitem, orders, [Bin("Eq", Ref("lineitem.%lineitem_orders"), Ref("orders.%TID%"))])
    else:
        j = Join(orders, lineitem, [Bin("Eq", Ref("lineitem.%lineitem_orders"), Ref("orders.%TID%"))])
    keys = []
    if rng.random() < 0.5:
        seg = rng.choice(sorted(catalog.dictionary["customer.c_mktsegment"].values()))
        customer = Table("customer", [("customer.c_mktsegment", None), ("customer.c_nationkey", None), ("customer.%TID%", None)])
        if rng.random() < 0.7:
            customer = Select(customer, Bin("Eq", Ref("customer.c_mktsegment"), Lit(("str", "customer.c_mktsegment"), seg)))
        j = Join(j, customer, [Bin("Eq", Ref("orders.%orders_customer"), Ref("customer.%TID%"))])
        if rng.random() < 0.5:
            keys.append("customer.c_nationkey")
    if rng.random() < 0.4:
        supplier = Table("supplier", [("supplier.s_nationkey", None), ("supplier.%TID%", None)])
        j = Join(j, supplier, [Bin("Eq", Ref("lineitem.%lineitem_supplier"), Ref("supplier.%TID%"))])
        if not keys and rng.random() < 0.5:
            keys.append("supplier.s_nationkey")
    if not keys and rng.random() < 0.6:
        keys = rng.sample(LI_KEYS, rng.randint(1, 2))
    outaggs, names = aggs(rng, keys)
    g = GroupBy(j, [(k, None) for k in keys], outaggs)
    return Project(g, [(Ref(k), None) for k in keys] + [(Ref(n), f"{n}.out_{n.lower()}") for n in names])
