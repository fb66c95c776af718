"""Second opinion on the oracle: evaluate each query's *SQL* semantics directly in numpy.

TEST INFRASTRUCTURE ONLY (same rules as oracle/vdl_oracle.c).  This does not interpret the
Voodoo plan; it restates the SQL in the comment header of the reference's mplan fixtures
(tests/tpch10noorder/06.sql.mplan:1-9, 01.sql.mplan:1-17, ...) over the same integer
encodings the translator uses (dates as day counts Mplan.hs:46-57, decimals as scaled ints,
dictionary codes) so integer outputs must match the plan interpreter bit for bit.
All arithmetic is int64 with wraparound (numpy semantics), AVG is truncating integer division
as emitted (Vlite.hs:1038-1041).
"""
from __future__ import annotations

import numpy as np


def q6(c: dict) -> dict:
    """sum(l_extendedprice*l_discount) where 1994-01-01 <= shipdate < 1995-01-01, 0.05<=disc<=0.07, qty<24."""
    sd, d, q, ep = (c["lineitem." + k] for k in ("l_shipdate", "l_discount", "l_quantity", "l_extendedprice"))
    m = (sd >= 728294) & (sd < 728659) & (d >= 5) & (d <= 7) & (q < 2400)
    if not m.any():
        return {"revenue": np.zeros(0, np.int64)}          # G14: a fold over an empty selection has no runs
    with np.errstate(over="ignore"):
        return {"revenue": np.array([(ep[m].astype(np.int64) * d[m].astype(np.int64)).sum(dtype=np.int64)])}


def q1(c: dict) -> dict:
    """group by (returnflag, linestatus) where shipdate <= 1998-09-02 (729999)."""
    g = lambda k: c["lineitem." + k].astype(np.int64)
    sd, rf, ls, qty, ep, disc, tax = (g(k) for k in ("l_shipdate", "l_returnflag", "l_linestatus", "l_quantity",
                                                     "l_extendedprice", "l_discount", "l_tax"))
    m = sd <= 729999
    rf, ls, qty, ep, disc, tax = (a[m] for a in (rf, ls, qty, ep, disc, tax))
    # composite key as emitted: ((rf>>3)-2)<<2 | ((ls>>3)-2), & 31 (Vlite.hs:1123-1170; bounds.csv:67-68)
    key = ((((rf >> 3) - 2) << 2) | ((ls >> 3) - 2)) & 31
    names = ["l_returnflag__lineitem__l_returnflag", "l_linestatus__lineitem__l_linestatus", "sum_qty", "sum_base_price",
             "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc", "count_order"]
    out = {n: [] for n in names}
    with np.errstate(over="ignore"):
        dp = ep * (100 - disc)
        ch = dp * (100 + tax)
        for k in np.unique(key):
            s = key == k
            first = int(np.argmax(s))
            cnt = int(s.sum())
            vals = [rf[first], ls[first], qty[s].sum(dtype=np.int64), ep[s].sum(dtype=np.int64),
                    dp[s].sum(dtype=np.int64), ch[s].sum(dtype=np.int64)]
            vals += [_tdiv(vals[2], cnt), _tdiv(vals[3], cnt), _tdiv(disc[s].sum(dtype=np.int64), cnt), cnt]
            for n, v in zip(names, vals):
                out[n].append(int(v))
    return {n: np.array(v, dtype=np.int64) for n, v in out.items()}


def _tdiv(a, b) -> int:
    """C truncating division (G4)."""
    a, b = int(a), int(b)
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def q3(c: dict) -> dict:
    """03.sql.mplan: orders[o_orderdate < 1995-03-15] x customer[BUILDING] x lineitem[l_shipdate > 1995-03-15],
    group by (l_orderkey, o_shippriority, o_orderdate); groups in ascending composite-key order."""
    d = 728732
    cust_ok = c["customer.c_mktsegment"] == 16
    ord_ok = (c["orders.o_orderdate"] < d) & cust_ok[c["orders.orders_customer"]]
    fk = c["lineitem.lineitem_orders"]
    m = (c["lineitem.l_shipdate"] > d) & ord_ok[fk]
    ok = c["lineitem.l_orderkey"][m].astype(np.int64)
    od = c["orders.o_orderdate"][fk[m]].astype(np.int64)
    sp = c["orders.o_shippriority"][fk[m]].astype(np.int64)
    with np.errstate(over="ignore"):
        rev = c["lineitem.l_extendedprice"][m] * (100 - c["lineitem.l_discount"][m])
    key = ((((ok - 1) << 0) | (sp >> 33)) << 12) | (od - 727563)
    order = np.argsort(key, kind="stable")
    key, ok, od, sp, rev = key[order], ok[order], od[order], sp[order], rev[order]
    heads = np.r_[True, key[1:] != key[:-1]] if len(key) else np.zeros(0, bool)
    starts = np.nonzero(heads)[0]
    sums = np.add.reduceat(rev, starts) if len(starts) else np.zeros(0, np.int64)
    return {"l_orderkey__lineitem__l_orderkey": ok[starts], "revenue": sums.astype(np.int64),
            "o_orderdate__orders__o_orderdate": od[starts], "o_shippriority__orders__o_shippriority": sp[starts]}


def q5(c: dict) -> dict:
    """05.sql.mplan: revenue per nation for ASIA, orders of 1994, customer and supplier in the same nation."""
    o_ok = (c["orders.o_orderdate"] >= 728294) & (c["orders.o_orderdate"] < 728659)
    lo, ls = c["lineitem.lineitem_orders"], c["lineitem.lineitem_supplier"]
    cust_nation = c["customer.c_nationkey"][c["orders.orders_customer"][lo]]
    supp_nation_key = c["supplier.s_nationkey"][ls]
    nation_row = c["supplier.supplier_nation"][ls]
    region_ok = c["region.r_name"] == 64
    m = o_ok[lo] & (cust_nation == supp_nation_key) & region_ok[c["nation.nation_region"][nation_row]]
    name = c["nation.n_name"][nation_row[m]].astype(np.int64)
    with np.errstate(over="ignore"):
        rev = c["lineitem.l_extendedprice"][m] * (100 - c["lineitem.l_discount"][m])
    names = np.unique(name)
    return {"n_name__nation__n_name": names, "revenue": np.array([rev[name == n].sum(dtype=np.int64) for n in names], dtype=np.int64)}


def q12(c: dict) -> dict:
    """12.sql.mplan: per ship mode (MAIL = 40, SHIP = 160: dictionary.csv:86,89), lines received in 1994 that were
    committed late, split by order priority 1-URGENT / 2-HIGH (40 / 104: dictionary.csv:78-79) vs the rest."""
    L = lambda n: c["lineitem." + n]       # noqa: E731
    sel = (L("l_shipdate") < L("l_commitdate")) & (L("l_commitdate") < L("l_receiptdate")) & (L("l_receiptdate") >= 728294) & \
          (L("l_receiptdate") < 728659) & ((L("l_shipmode") == 40) | (L("l_shipmode") == 160))
    prio = c["orders.o_orderpriority"][L("lineitem_orders")]
    high = (prio == 40) | (prio == 104)
    modes = np.unique(L("l_shipmode")[sel]).astype(np.int64)
    return {"l_shipmode__lineitem__l_shipmode": modes,
            "high_line_count": np.array([int((sel & high & (L("l_shipmode") == m)).sum()) for m in modes], dtype=np.int64),
            "low_line_count": np.array([int((sel & ~high & (L("l_shipmode") == m)).sum()) for m in modes], dtype=np.int64)}


def q19(c: dict, dictionary: dict) -> dict:
    """19.sql.mplan: discounted revenue of lineitems delivered in person by air whose part matches one of three
    (brand, container set, quantity range, size range) combinations."""
    D = dictionary
    L = lambda n: c["lineitem." + n]                              # noqa: E731
    P = lambda n: c["part." + n][L("lineitem_part")]              # noqa: E731

    def inset(x, col, names):
        m = np.zeros(len(x), bool)
        for n in names:
            m |= x == D[col][n]
        return m
    q, size, brand, cont = L("l_quantity"), P("p_size"), P("p_brand"), P("p_container")
    base = (L("l_shipinstruct") == D["lineitem.l_shipinstruct"]["DELIVER IN PERSON"]) & inset(L("l_shipmode"), "lineitem.l_shipmode", ["AIR", "AIR REG"])
    sm, med, lg = ["SM CASE", "SM BOX", "SM PACK", "SM PKG"], ["MED BAG", "MED BOX", "MED PKG", "MED PACK"], ["LG CASE", "LG BOX", "LG PACK", "LG PKG"]
    pb = (size >= 1) & (size <= 15) & inset(brand, "part.p_brand", ["Brand#12", "Brand#23", "Brand#34"]) & inset(cont, "part.p_container", sm + med + lg)
    c1 = (brand == D["part.p_brand"]["Brand#12"]) & inset(cont, "part.p_container", sm) & (q >= 100) & (q <= 1100) & (size >= 1) & (size <= 5)
    c2 = (brand == D["part.p_brand"]["Brand#23"]) & inset(cont, "part.p_container", med) & (q >= 1000) & (q <= 2000) & (size >= 1) & (size <= 10)
    c3 = (brand == D["part.p_brand"]["Brand#34"]) & inset(cont, "part.p_container", lg) & (q >= 2000) & (q <= 3000) & (size >= 1) & (size <= 15)
    m = base & pb & (c1 | c2 | c3)
    if not m.any():
        return {"revenue": np.zeros(0, np.int64)}
    with np.errstate(over="ignore"):
        return {"revenue": np.array([(L("l_extendedprice")[m] * (100 - L("l_discount")[m])).sum(dtype=np.int64)], dtype=np.int64)}
