"""Synthetic TPC-H-shaped columns drawn to the reference's bounds metadata.

The reference ships metadata but no data, so throughput is reported on columns generated to
``tests/tpch10noorder/bounds.csv`` (min, max, trailing zeros, count; Config.hs:57, 114-120).
The recipe is *counter based*: a value depends only on (seed, stream, global row), so every
GPU shard generates its own row range in place (``vdl_column_fill_synthetic``) and the CPU
oracle generates identical data on the host -- nothing crosses PCIe.

    base = splitmix64(seed ^ (stream * 0x9E3779B97F4A7C15));  h = splitmix64(base + row)
    UNIFORM : vmin + stride * mulhi64(h, p0)         p0 = number of distinct values
    SEQ     : vmin + stride * row
    FKDENSE : vmin + stride * ((row * p0) // p1)     nondecreasing FK; p0 = dim rows, p1 = fact rows

Columns that must agree row by row (an FK index column and the key column it was built from,
e.g. lineitem_orders / l_orderkey) share one ``stream``.
"""
from __future__ import annotations

from dataclasses import dataclass

from .meta import Catalog

UNIFORM, SEQ, FKDENSE = 0, 1, 2

# lineitem cardinalities per TPC-H scale factor (SF10 is the fixture's own count, bounds.csv:59-79).
_LINEITEM_ROWS = {1: 6_001_215, 10: 59_986_052, 100: 600_037_902}
_FIXED_ROWS = {"nation": 25, "region": 5}


@dataclass(frozen=True)
class ColumnSpec:
    name: str       # "table.column"
    width: int      # 4 or 8 bytes
    kind: int
    vmin: int
    stride: int
    p0: int
    p1: int
    stream: int     # RNG stream id (64-bit)


def seed_for(sf: float) -> int:
    return 0x5EED ^ int(round(sf * 1000))


def _fnv1a(s: str) -> int:
    h = 0xCBF29CE484222325
    for b in s.encode():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def table_rows(cat: Catalog, table: str, sf: float) -> int:
    if table in _FIXED_ROWS:
        return _FIXED_ROWS[table]
    if table == "lineitem" and sf in _LINEITEM_ROWS:
        return _LINEITEM_ROWS[int(sf)]
    return max(1, int(round(cat.tables[table].rows * sf / 10.0)))


# columns whose values are dictionary codes with a non-power-of-two spacing, or whose real
# distribution is a sequence: (kind, vmin, stride, number of values or None = table rows)
_SPECIAL = {
    "lineitem.l_returnflag": (UNIFORM, 16, 24, 3),     # 16, 40, 64   (dictionary.csv:80-82)
    "lineitem.l_linestatus": (UNIFORM, 16, 24, 2),     # 16, 40
    "customer.c_mktsegment": (UNIFORM, -112, 32, 5),   # 5 codes, includes 16 = 'BUILDING' (dictionary.csv:74)
    "nation.n_name": (SEQ, 16, 24, None),              # distinct per nation, inside [16, 640]
    "region.r_name": (SEQ, 16, 24, None),              # 16, 40 'AMERICA', 64 'ASIA', 88, 112
    "region.r_regionkey": (SEQ, 0, 1, None),
    "nation.n_nationkey": (SEQ, 0, 1, None),
    "customer.c_custkey": (SEQ, 1, 1, None),
    "supplier.s_suppkey": (SEQ, 1, 1, None),
    "part.p_partkey": (SEQ, 1, 1, None),
    "orders.o_orderkey": (SEQ, 1, 4, None),
}

# key column -> (FK index column it mirrors, vmin, stride): value = vmin + stride * fk_index
_MIRRORS = {
    "lineitem.l_orderkey": ("lineitem_orders", 1, 4),
    "lineitem.l_suppkey": ("lineitem_supplier", 1, 1),
    "lineitem.l_partkey": ("lineitem_part", 1, 1),
    "orders.o_custkey": ("orders_customer", 1, 1),
    "customer.c_nationkey": ("customer_nation", 0, 1),
    "supplier.s_nationkey": ("supplier_nation", 0, 1),
    "nation.n_regionkey": ("nation_region", 0, 1),
    "partsupp.ps_partkey": ("partsupp_part", 1, 1),
    "partsupp.ps_suppkey": ("partsupp_supplier", 1, 1),
}


def column_spec(cat: Catalog, qualified: str, sf: float) -> ColumnSpec:
    table, cname = qualified.split(".", 1)
    tab = cat.tables[table]
    col = tab.columns[cname]
    rows = table_rows(cat, table, sf)
    fk = {f.name: f for f in tab.fkeys}
    stream = _fnv1a(qualified)

    def fk_spec(fkname: str, vmin: int, stride: int) -> ColumnSpec:
        dim_rows = table_rows(cat, fk[fkname].ref_table, sf)
        s = _fnv1a(f"{table}.{fkname}")
        if fkname == "lineitem_orders":   # clustered: lineitem is stored in l_orderkey order (storage.csv:188)
            return ColumnSpec(qualified, col.width, FKDENSE, vmin, stride, dim_rows, rows, s)
        return ColumnSpec(qualified, col.width, UNIFORM, vmin, stride, dim_rows, 0, s)

    if cname in fk:
        return fk_spec(cname, 0, 1)
    if qualified in _MIRRORS:
        fkname, vmin, stride = _MIRRORS[qualified]
        return fk_spec(fkname, vmin, stride)
    if qualified in _SPECIAL:
        kind, vmin, stride, n = _SPECIAL[qualified]
        return ColumnSpec(qualified, col.width, kind, vmin, stride, n if n is not None else rows, 0, stream)
    if col.vmin == -(1 << 63):            # pkey pseudo-columns carry nil bounds (bounds.csv:4): row ids
        return ColumnSpec(qualified, col.width, SEQ, 0, 1, rows, 0, stream)
    tz = col.trailing_zeros if col.vmax > col.vmin else 0
    nvals = ((col.vmax - col.vmin) >> tz) + 1
    return ColumnSpec(qualified, col.width, UNIFORM, col.vmin, 1 << tz, nvals, 0, stream)
