"""Loaders for the metadata files mplan2vdl consumes, in the reference's own formats.

The executor needs the same four files the translator takes on its command line
(``tpchrun`` passes ``-b bounds.csv -s schema.msqldump -t storage.csv --dictionary
dictionary.csv``; reference tpchrun:1-4, MainFuns.hs:140-153):

* ``bounds.csv``     ``table,column,min,max,count,trailing_zeros``           Config.hs:57, 137-147
* ``storage.csv``    ``schema,table,column,type,location,count,typewidth,
                       columnsize,heapsize,hashes,imprints,sorted``           Config.hs:60-72
* ``dictionary.csv`` ``table,column,"string",code``                           Config.hs:75-86
* ``schema.msqldump`` ``CREATE TABLE .. CONSTRAINT .. PRIMARY KEY .. FOREIGN KEY .. REFERENCES``
                                                                              SchemaParser.y:62-141

They give the executor the physical width of every column (Types.hs:84-87, 129-140), the
value bounds used to size group-key domains and to draw synthetic data, FK targets for the
replicated dimension tables, and the dictionary used to decode string-coded outputs
(resolve.py:34-94).
"""
from __future__ import annotations

import csv
import json
import os
import re
from dataclasses import dataclass, field

# Monet type -> storage width in bytes (Types.hs:129-140 getSTypeOfMType, 84-87 sizeOf).
_WIDTH = {
    "int": 4, "date": 4, "smallint": 4, "tinyint": 4,
    "oid": 8, "bigint": 8, "char": 8, "varchar": 8, "decimal": 8,
}


@dataclass
class Column:
    table: str
    name: str
    vmin: int
    vmax: int
    count: int
    trailing_zeros: int
    mtype: str = ""      # Monet type name from storage.csv ("" if the column is only in bounds.csv)
    width: int = 8       # bytes in this executor's storage model
    is_sorted: bool = False
    scale: int = 0       # DECIMAL(p, s) of the schema: the column's display type DDecimal{point = s} (Types.hs:143-153)

    @property
    def qualified(self) -> str:
        return f"{self.table}.{self.name}"


@dataclass
class ForeignKey:
    name: str            # constraint name == name of the FK index column (Mplan.hs:240-251 JOINIDX)
    table: str
    columns: list
    ref_table: str
    ref_columns: list


@dataclass
class Table:
    name: str
    rows: int = 0
    columns: dict = field(default_factory=dict)
    pkey: list = field(default_factory=list)
    pkey_name: str = ""
    fkeys: list = field(default_factory=list)


@dataclass
class Catalog:
    tables: dict = field(default_factory=dict)
    dictionary: dict = field(default_factory=dict)   # "table.column" -> {string: code}

    def column(self, qualified: str) -> Column:
        t, c = qualified.split(".", 1)
        return self.tables[t].columns[c]

    def decode(self, qualified: str, code: int):
        """Inverse dictionary lookup, as resolve.py:64-94 does for ``name__table__col`` outputs."""
        for s, k in self.dictionary.get(qualified, {}).items():
            if k == code:
                return s
        return None

    def to_json(self) -> dict:
        out = {"tables": {}, "dictionary": self.dictionary}
        for t in self.tables.values():
            out["tables"][t.name] = {
                "rows": t.rows, "pkey": t.pkey, "pkey_name": t.pkey_name,
                "fkeys": [vars(f) for f in t.fkeys],
                "columns": {c.name: {"min": c.vmin, "max": c.vmax, "count": c.count, "tz": c.trailing_zeros,
                                     "mtype": c.mtype, "width": c.width, "sorted": c.is_sorted, "scale": c.scale}
                            for c in t.columns.values()},
            }
        return out

    @staticmethod
    def from_json(obj: dict) -> "Catalog":
        cat = Catalog(dictionary=obj.get("dictionary", {}))
        for tn, t in obj["tables"].items():
            tab = Table(name=tn, rows=t["rows"], pkey=t["pkey"], pkey_name=t.get("pkey_name", ""))
            tab.fkeys = [ForeignKey(**f) for f in t["fkeys"]]
            for cn, c in t["columns"].items():
                tab.columns[cn] = Column(tn, cn, c["min"], c["max"], c["count"], c["tz"], c["mtype"], c["width"],
                                         c.get("sorted", False), c.get("scale", 0))
            cat.tables[tn] = tab
        return cat


def read_bounds(path: str) -> list:
    """bounds.csv rows -> Column (no header; 6 fields; Config.hs:57)."""
    cols = []
    with open(path, newline="") as f:
        for rec in csv.reader(f):
            if not rec:
                continue
            if len(rec) != 6:
                raise ValueError(f"{path}: bounds record needs 6 fields (Config.hs:57), got {rec}")
            t, c, lo, hi, cnt, tz = rec
            cols.append(Column(t, c, int(lo), int(hi), int(cnt), int(tz)))
    return cols


def read_storage(path: str) -> dict:
    """storage.csv -> {(table, column): (mtype, count, sorted)} (12 fields; Config.hs:60-72)."""
    out = {}
    with open(path, newline="") as f:
        for rec in csv.reader(f):
            if not rec:
                continue
            if len(rec) != 12:
                raise ValueError(f"{path}: storage record needs 12 fields (Config.hs:60-72), got {len(rec)}")
            _schema, t, c, typ, _loc, cnt, _tw, _cs, _hs, _h, _i, srt = rec
            out[(t, c)] = (typ, int(cnt), srt.strip().lower() == "true")
    return out


def read_dictionary(path: str) -> dict:
    """dictionary.csv -> {"table.column": {string: code}} (Config.hs:75-86)."""
    out: dict = {}
    with open(path, newline="") as f:
        for rec in csv.reader(f):
            if not rec:
                continue
            t, c, s, code = rec
            out.setdefault(f"{t}.{c}", {})[s] = int(code)
    return out


_RE_TABLE = re.compile(r'CREATE TABLE\s+"(\w+)"\."(\w+)"\s*\(', re.I)
_RE_PKEY = re.compile(r'CONSTRAINT\s+"(\w+)"\s+PRIMARY KEY\s*\(([^)]*)\)', re.I)
_RE_COLDEF = re.compile(r'^\s*"(\w+)"\s+([A-Za-z]+)\s*(?:\(\s*(\d+)\s*(?:,\s*(\d+)\s*)?\))?')
_RE_FKEY = re.compile(r'CONSTRAINT\s+"(\w+)"\s+FOREIGN KEY\s*\(([^)]*)\)\s+REFERENCES\s+"(\w+)"\."(\w+)"\s*\(([^)]*)\)', re.I)


def _names(s: str) -> list:
    return [x.strip().strip('"') for x in s.split(",") if x.strip()]


def read_schema(path: str) -> dict:
    """``msqldump -D`` DDL -> {table: (pkey_name, pkey_cols, [ForeignKey], {column: decimal scale})} (SchemaParser.y:62-141)."""
    out = {}
    cur = None
    with open(path) as f:
        for line in f:
            m = _RE_TABLE.search(line)
            if m:
                cur = m.group(2)
                out[cur] = ["", [], [], {}]
                continue
            if cur is None:
                continue
            m = _RE_FKEY.search(line)
            if m:
                out[cur][2].append(ForeignKey(m.group(1), cur, _names(m.group(2)), m.group(4), _names(m.group(5))))
                continue
            m = _RE_PKEY.search(line)
            if m:
                out[cur][0], out[cur][1] = m.group(1), _names(m.group(2))
                continue
            m = _RE_COLDEF.match(line)
            if m and m.group(2).lower() == "decimal" and m.group(4) is not None:
                out[cur][3][m.group(1)] = int(m.group(4))          # column type specs (getTspecs, Config.hs:187-188)
    return out


def load_metadata(directory: str) -> Catalog:
    """Read the four files of a metadata directory such as the reference's tests/tpch10noorder/."""
    dpath = os.path.join(directory, "dictionary.csv")
    return load_metadata_files(os.path.join(directory, "bounds.csv"), os.path.join(directory, "storage.csv"),
                               os.path.join(directory, "schema.msqldump"), dpath if os.path.exists(dpath) else None)


def load_metadata_files(bounds: str, storage_path: str, schema: str, dictionary: str | None) -> Catalog:
    """makeConfig (Config.hs:149-170) from the reference's four inputs: -b bounds, -t storage, -s schema, --dictionary."""
    cat = Catalog()
    storage = read_storage(storage_path)
    for col in read_bounds(bounds):
        tab = cat.tables.setdefault(col.table, Table(col.table))
        st = storage.get((col.table, col.name))
        if st:
            col.mtype, _cnt, col.is_sorted = st
            if col.mtype not in _WIDTH:
                raise ValueError(f"storage type {col.mtype} of {col.qualified} not expected (Types.hs:129-140)")
            col.width = _WIDTH[col.mtype]
        tab.columns[col.name] = col
        tab.rows = max(tab.rows, col.count)
    for t, (pk_name, pk_cols, fks, scales) in read_schema(schema).items():
        if t in cat.tables:
            cat.tables[t].pkey_name, cat.tables[t].pkey, cat.tables[t].fkeys = pk_name, pk_cols, fks
            for c, sc in scales.items():
                if c in cat.tables[t].columns:
                    cat.tables[t].columns[c].scale = sc
    if dictionary:
        cat.dictionary = read_dictionary(dictionary)
    return cat


_BUILTIN = os.path.join(os.path.dirname(__file__), "catalog", "tpch10noorder.json")


def builtin_catalog() -> Catalog:
    """The SF10 catalogue derived from the reference's tests/tpch10noorder metadata by
    tools/make_catalog.py (committed so the GPU box, which has no /root/reference, can use it)."""
    with open(_BUILTIN) as f:
        return Catalog.from_json(json.load(f))
