/*
 * vdl_oracle.c -- CPU oracle for the Voodoo dataflow graphs emitted by mplan2vdl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under mplan2vdl_b200/ (the product) may
 * import, link or execute this file.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs (as the checker / the timed
 * CPU baseline, never as the product path).
 *
 * PARITY UNPINNED: the reference (orm011/mplan2vdl) is a text->text compiler that
 * ships NO executor (the Voodoo server it POSTed plans to is not in the repo:
 * eval_query.sh:21-24; tests/Tests.hs:17-18 "runs nothing"), and no Haskell
 * toolchain exists in this image, so there is neither a golden result vector nor a
 * reference binary to pin the op semantics against.  What IS pinned by the
 * reference: the plan text (README.md:40-52 golden lines, checked in
 * tests/test_plan_text.py), the op vocabulary and argument order (Vdl.hs:32-44,
 * 110-131, 410-453), literal encodings (Mplan.hs:46-57) and type widths
 * (Types.hs:84-87).  The semantics below are the ones SURVEY.md section 2.3 / App. G
 * infers from how Vlite.hs *uses* each op; each function cites the usage site.
 * A second, independent opinion (SQL semantics over the same columns in numpy) lives
 * in oracle/sqlref.py and must agree bit-exactly.
 *
 * Model: dense vectors of int64 (G1).  All arithmetic wraps mod 2^64 (G5).
 *
 *   Load,<table.col>                  bind a registered column (int32 columns sign-extend)  Vdl.hs:161-168,419-420
 *   Project,<out>,Id n,<in>           rename; identity on data                              Vdl.hs:422-423
 *   RangeV,val,from,Id v,step         out[i]=from+i*step, len=len(v)                        Vdl.hs:428-431; Vlite.hs:176-191
 *   RangeC,val,from,count,step        explicit length                                       Vdl.hs:433-434
 *   <Binop>,val,Id a,val,Id b,val     elementwise; comparisons/logicals give 0/1           Vdl.hs:436-439, 136-157
 *       BitShift: b>=0 arithmetic right shift, b<0 left shift by -b                         Vlite.hs:205-208 (G7)
 *       Divide: C truncation; x/0 := 0; INT64_MIN/-1 wraps (G4)  Modulo: C remainder; x%0 := 0
 *       LogicalAnd/Or: any non-zero is true (G8)
 *   FoldSelect,val,Id fold,val,Id p   ascending positions i with p[i]!=0 (fold must be pos_) Vlite.hs:721-730
 *   Gather,Id src,Id pos,val          out[i]=src[pos[i]]                                     Vlite.hs:86-87
 *   Scatter,Id src,Id fold,val,Id pos out[pos[i]]=src[i], unwritten=0, length = index-space
 *                                     of pos (G2: see domain_of); positions unique, or        Vlite.hs:1057-1059,1268-1275
 *                                     duplicates that carry the same value (semijoin marks)   Vlite.hs:1214-1222
 *   Partition,val,Id data,val,Id piv  destination positions of the STABLE sort of rows by
 *                                     bucket(data) = #pivots < data[i]  (G3)                  Vlite.hs:1082-1098,1057-1060
 *   Fold{Sum,Min,Max,Choose,Count}    one output per run of equal consecutive `groups`
 *       ,val,Id groups,val,Id data    values, in run order; Choose = first of run (G6);
 *                                     empty input -> empty output (G14)                      Vlite.hs:1048-1070; Vdl.hs:255-264
 *                                     2-level folds (--agghierarchical, Vlite.hs:1181-1192) fold the level-1 RESULTS by
 *                                     the row-space groups: in the dense model the data then has one entry per
 *                                     level-1 run and is grouped by the groups value at that run's first row
 *   Shuffle,Id n                      identity (order-destroying hint)                        Vdl.hs:449-450
 *   MaterializeCompact,Id n           query output; name = the Project's <out>               Vdl.hs:278-292,452-453
 *   CrossProductOuter / Inner,Id l,Id r   positions into l (i / |r|) resp. r (i % |r|) of the |l| x |r| pairs                Vlite.hs:89-93,283-292
 *   Like,val,Id d,val,Id heap,val,pat out[i] = 1 if the NUL-terminated string at byte offset d[i] of the column's string
 *                                     heap (`Load,<table>.<col>.heap`, a byte vector) matches the SQL LIKE pattern
 *                                     (% any run, _ any one byte, no escape, case-sensitive), else 0          Vdl.hs:244-247,444-447
 *   Semisort                          rejected (VliteFormat only: Vlite.hs:1061-1064)
 *
 * Threads: OpenMP, row-range split per op, deterministic combine.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <omp.h>
#include <time.h>

typedef int64_t i64;
typedef uint64_t u64;

/* ------------------------------------------------------------------ vectors */
enum { K_DENSE = 0, K_COL32 = 1, K_RANGE = 2, K_COL64 = 3, K_BYTES = 4 };
typedef struct {
  int kind;
  i64 n;
  i64 *d;            /* K_DENSE (owned) or K_COL64 (borrowed) */
  const int32_t *d32;/* K_COL32 (borrowed) */
  const unsigned char *d8; /* K_BYTES (borrowed): a string heap, only Like reads it */
  i64 from, step;    /* K_RANGE */
  i64 domain;        /* length of the vector these values index into; -1 unknown */
  int valid;
} vec;

static inline i64 vget(const vec *v, i64 i) {
  switch (v->kind) {
    case K_COL32: return (i64)v->d32[i];
    case K_RANGE: return (i64)((u64)v->from + (u64)i * (u64)v->step);
    default: return v->d[i];
  }
}

/* ------------------------------------------------------------------ plan */
enum {
  OP_LOAD, OP_PROJECT, OP_RANGEV, OP_RANGEC,
  OP_LAND, OP_LOR, OP_BAND, OP_BOR, OP_SHIFT, OP_EQ, OP_ADD, OP_SUB, OP_GT, OP_MUL, OP_DIV, OP_MOD,
  OP_FCHOOSE, OP_FSELECT, OP_FMAX, OP_FSUM, OP_FMIN, OP_FCOUNT,
  OP_GATHER, OP_SCATTER, OP_PARTITION, OP_SHUFFLE, OP_MATERIALIZE, OP_LIKE, OP_CROSS_OUTER, OP_CROSS_INNER, OP_UNSUPPORTED
};
static const struct { const char *name; int op; } OPNAMES[] = {
  {"Load", OP_LOAD}, {"Project", OP_PROJECT}, {"RangeV", OP_RANGEV}, {"RangeC", OP_RANGEC},
  {"LogicalAnd", OP_LAND}, {"LogicalOr", OP_LOR}, {"BitwiseAnd", OP_BAND}, {"BitwiseOr", OP_BOR},
  {"BitShift", OP_SHIFT}, {"Equals", OP_EQ}, {"Add", OP_ADD}, {"Subtract", OP_SUB},
  {"Greater", OP_GT}, {"Multiply", OP_MUL}, {"Divide", OP_DIV}, {"Modulo", OP_MOD},
  {"FoldChoose", OP_FCHOOSE}, {"FoldSelect", OP_FSELECT}, {"FoldMax", OP_FMAX},
  {"FoldSum", OP_FSUM}, {"FoldMin", OP_FMIN}, {"FoldCount", OP_FCOUNT},
  {"Gather", OP_GATHER}, {"Scatter", OP_SCATTER}, {"Partition", OP_PARTITION},
  {"Shuffle", OP_SHUFFLE}, {"MaterializeCompact", OP_MATERIALIZE},
  {"Like", OP_LIKE}, {"CrossProductOuter", OP_CROSS_OUTER},
  {"CrossProductInner", OP_CROSS_INNER}, {"Semisort", OP_UNSUPPORTED}, {NULL, 0}};

typedef struct {
  int id, op;
  int a, b, c;        /* argument node ids (0 = none) */
  i64 k0, k1, k2;     /* RangeV: from, step; RangeC: from, count, step */
  char name[160];     /* Load: column; Project: out name; Like: pattern */
  int lastuse;        /* index of last statement that reads this node */
} stmt;

#define MAXCOLS 256
typedef struct { char name[128]; const void *data; int width; i64 rows; } colbind;
typedef struct { char name[160]; i64 n; i64 *d; } output;

typedef struct orc_env {
  colbind cols[MAXCOLS]; int ncols;
  output *outs; int nouts;
  char err[512];
  double seconds;
  int nstmts;
} orc_env;

static int fail(orc_env *e, const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(e->err, sizeof e->err, fmt, ap); va_end(ap);
  return -1;
}

orc_env *orc_env_new(void) { return (orc_env *)calloc(1, sizeof(orc_env)); }
static void clear_outputs(orc_env *e) {
  for (int i = 0; i < e->nouts; i++) free(e->outs[i].d);
  free(e->outs); e->outs = NULL; e->nouts = 0;
}
void orc_env_free(orc_env *e) { if (!e) return; clear_outputs(e); free(e); }
const char *orc_last_error(orc_env *e) { return e->err; }
double orc_last_seconds(orc_env *e) { return e->seconds; }
int orc_num_statements(orc_env *e) { return e->nstmts; }
int orc_num_outputs(orc_env *e) { return e->nouts; }
const char *orc_output_name(orc_env *e, int i) { return e->outs[i].name; }
i64 orc_output_len(orc_env *e, int i) { return e->outs[i].n; }
const i64 *orc_output_data(orc_env *e, int i) { return e->outs[i].d; }

/* Columns are borrowed, not copied.  width = 4 (int/date, Types.hs:84-87,129-140) or 8. */
int orc_bind_column(orc_env *e, const char *name, const void *data, int width, i64 rows) {
  if (width != 4 && width != 8 && width != 1) return fail(e, "column %s: width %d not 4/8 (or 1 for a string heap)", name, width);
  for (int i = 0; i < e->ncols; i++)
    if (!strcmp(e->cols[i].name, name)) {
      e->cols[i].data = data; e->cols[i].width = width; e->cols[i].rows = rows; return 0;
    }
  if (e->ncols == MAXCOLS) return fail(e, "too many columns");
  colbind *c = &e->cols[e->ncols++];
  snprintf(c->name, sizeof c->name, "%s", name);
  c->data = data; c->width = width; c->rows = rows;
  return 0;
}

/* ---- parser: VdlFormat lines "id,Op,fields..." (Vdl.hs:410-453, printLine 455-477) */
static int parse_ref(const char *s, int *out) { /* "Id 12" */
  if (strncmp(s, "Id ", 3)) return -1;
  char *end; long v = strtol(s + 3, &end, 10);
  if (*end || v <= 0) return -1;
  *out = (int)v; return 0;
}
static int parse_int(const char *s, i64 *out) {
  char *end; long long v = strtoll(s, &end, 10);
  if (end == s || *end) return -1;
  *out = (i64)v; return 0;
}

static int parse_plan(orc_env *e, const char *text, stmt **out_stmts, int *out_n) {
  int cap = 64, n = 0;
  stmt *st = (stmt *)calloc(cap, sizeof(stmt));
  const char *p = text;
  int lineno = 0;
  while (*p) {
    const char *eol = strchr(p, '\n');
    size_t len = eol ? (size_t)(eol - p) : strlen(p);
    char line[1024];
    if (len >= sizeof line) { free(st); return fail(e, "line %d too long", lineno + 1); }
    memcpy(line, p, len); line[len] = 0;
    p = eol ? eol + 1 : p + len;
    lineno++;
    char *meta = strstr(line, " ;; ");           /* --metadata suffix, Vdl.hs:463-466 */
    if (meta) *meta = 0;
    size_t l = strlen(line);
    while (l && (line[l - 1] == '\r' || line[l - 1] == ' ')) line[--l] = 0;
    if (!l) continue;
    char *f[16]; int nf = 0;
    char *q = line;
    while (nf < 16) { f[nf++] = q; char *c = strchr(q, ','); if (!c) break; *c = 0; q = c + 1; }
    if (nf < 2) { free(st); return fail(e, "line %d: too few fields", lineno); }
    if (n == cap) { cap *= 2; st = (stmt *)realloc(st, cap * sizeof(stmt)); memset(st + n, 0, (cap - n) * sizeof(stmt)); }
    stmt *s = &st[n];
    i64 idv;
    if (parse_int(f[0], &idv) || idv != n + 1) { free(st); return fail(e, "line %d: ids must be 1,2,3,... (Vdl.hs:297-311)", lineno); }
    s->id = (int)idv; s->op = -1;
    for (int k = 0; OPNAMES[k].name; k++) if (!strcmp(OPNAMES[k].name, f[1])) { s->op = OPNAMES[k].op; break; }
    if (s->op < 0) { free(st); return fail(e, "line %d: unknown op %s", lineno, f[1]); }
    int bad = 0;
#define NEED(k) do { if (nf != (k)) { bad = 1; goto done; } } while (0)
#define VAL(i)  do { if (strcmp(f[i], "val")) { bad = 1; goto done; } } while (0)
    switch (s->op) {
      case OP_LOAD: NEED(3); snprintf(s->name, sizeof s->name, "%s", f[2]); break;
      case OP_PROJECT: NEED(5); snprintf(s->name, sizeof s->name, "%s", f[2]); bad = parse_ref(f[3], &s->a); break;
      case OP_RANGEV: NEED(6); VAL(2); bad = parse_int(f[3], &s->k0) | parse_ref(f[4], &s->a) | parse_int(f[5], &s->k1); break;
      case OP_RANGEC: NEED(6); VAL(2); bad = parse_int(f[3], &s->k0) | parse_int(f[4], &s->k1) | parse_int(f[5], &s->k2); break;
      case OP_GATHER: NEED(5); VAL(4); bad = parse_ref(f[2], &s->a) | parse_ref(f[3], &s->b); break;
      case OP_SCATTER: /* id,Scatter,src,fold,val,pos,val (Vdl.hs:441-442) */
        NEED(7); VAL(4); VAL(6); bad = parse_ref(f[2], &s->a) | parse_ref(f[3], &s->b) | parse_ref(f[5], &s->c); break;
      case OP_SHUFFLE: case OP_MATERIALIZE: NEED(3); bad = parse_ref(f[2], &s->a); break;
      case OP_CROSS_OUTER: case OP_CROSS_INNER: /* id,CrossProductOuter,Id left,Id right (Vdl.hs:412-416) */
        NEED(4); bad = parse_ref(f[2], &s->a) | parse_ref(f[3], &s->b); break;
      case OP_LIKE: { /* id,Like,val,Id data,val,Id heap,val,pattern (Vdl.hs:444-447); a pattern may contain commas */
        if (nf < 8) { bad = 1; goto done; }
        VAL(2); VAL(4); VAL(6); bad = parse_ref(f[3], &s->a) | parse_ref(f[5], &s->b);
        size_t w = 0; s->name[0] = 0;
        for (int k = 7; k < nf; k++) w += (size_t)snprintf(s->name + w, sizeof s->name - w, "%s%s", k > 7 ? "," : "", f[k]);
        break;
      }
      case OP_UNSUPPORTED: free(st); return fail(e, "line %d: op %s is out of scope for the oracle", lineno, f[1]);
      default: /* binary ops and folds: Op,val,Id a,val,Id b,val */
        NEED(7); VAL(2); VAL(4); VAL(6); bad = parse_ref(f[3], &s->a) | parse_ref(f[5], &s->b); break;
    }
  done:
    if (bad) { free(st); return fail(e, "line %d: malformed %s statement", lineno, f[1]); }
    int args[3] = {s->a, s->b, s->c};
    for (int k = 0; k < 3; k++) {
      if (args[k] >= s->id) { free(st); return fail(e, "line %d: forward reference Id %d", lineno, args[k]); }
      if (args[k] > 0) st[args[k] - 1].lastuse = n;
    }
    n++;
  }
  *out_stmts = st; *out_n = n;
  return 0;
}

/* ------------------------------------------------------------------ ops */
static vec new_dense(i64 n) {
  vec v; memset(&v, 0, sizeof v);
  v.kind = K_DENSE; v.n = n; v.domain = -1; v.valid = 1;
  v.d = (i64 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(i64));
  return v;
}
static void vfree(vec *v) { if (v->valid && v->kind == K_DENSE) free(v->d); v->valid = 0; v->d = NULL; }

static inline i64 binop_apply(int op, i64 a, i64 b) {
  switch (op) {
    case OP_LAND: return (a != 0) && (b != 0);
    case OP_LOR: return (a != 0) || (b != 0);
    case OP_BAND: return a & b;
    case OP_BOR: return a | b;
    case OP_SHIFT:                       /* Vlite.hs:205-208: sign encodes direction */
      if (b >= 0) return b >= 64 ? (a < 0 ? -1 : 0) : (a >> b);
      return b <= -64 ? 0 : (i64)((u64)a << (-b));
    case OP_EQ: return a == b;
    case OP_ADD: return (i64)((u64)a + (u64)b);
    case OP_SUB: return (i64)((u64)a - (u64)b);
    case OP_GT: return a > b;
    case OP_MUL: return (i64)((u64)a * (u64)b);
    case OP_DIV: if (b == 0) return 0; if (b == -1) return (i64)(0 - (u64)a); return a / b;
    case OP_MOD: if (b == 0 || b == -1) return 0; return a % b;
  }
  return 0;
}

static int op_binary(orc_env *e, int op, const vec *a, const vec *b, vec *out) {
  if (a->n != b->n) return fail(e, "elementwise op on lengths %lld vs %lld", (long long)a->n, (long long)b->n);
  *out = new_dense(a->n);
  i64 n = a->n; i64 *o = out->d;
#define LOOP(GA, GB) _Pragma("omp parallel for schedule(static)") for (i64 i = 0; i < n; i++) o[i] = binop_apply(op, GA, GB);
  /* specialise the common operand kinds so the inner loop has no per-element dispatch */
  int ka = a->kind == K_COL64 ? K_DENSE : a->kind, kb = b->kind == K_COL64 ? K_DENSE : b->kind;
  if (ka == K_DENSE && kb == K_DENSE) { const i64 *x = a->d, *y = b->d; LOOP(x[i], y[i]) }
  else if (ka == K_DENSE && kb == K_RANGE && b->step == 0) { const i64 *x = a->d; i64 y = b->from; LOOP(x[i], y) }
  else if (ka == K_RANGE && a->step == 0 && kb == K_DENSE) { i64 x = a->from; const i64 *y = b->d; LOOP(x, y[i]) }
  else if (ka == K_COL32 && kb == K_RANGE && b->step == 0) { const int32_t *x = a->d32; i64 y = b->from; LOOP((i64)x[i], y) }
  else if (ka == K_RANGE && a->step == 0 && kb == K_COL32) { i64 x = a->from; const int32_t *y = b->d32; LOOP(x, (i64)y[i]) }
  else { LOOP(vget(a, i), vget(b, i)) }
#undef LOOP
  /* positions `p % k`, k a positive constant: the scatter size hint (addScatterSizeHint, Vlite.hs:1117-1120; G9) --
     they index a space of k slots (G2), whatever p indexed (an empty dim' still receives the slot-0 writes of the
     partnerless fact rows, Vlite.hs:1214-1218) */
  if (op == OP_MOD && b->kind == K_RANGE && b->step == 0 && b->from > 0) out->domain = b->from;
  return 0;
}

/* FoldSelect: Vlite.hs:721-730 -- idx = Fold FSel (pos_ p) p.  Dense model (G1): global stable compaction. */
static int op_fold_select(orc_env *e, const vec *fold, const vec *pred, vec *out) {
  if (!(fold->kind == K_RANGE && fold->from == 0 && fold->step == 1 && fold->n == pred->n))
    return fail(e, "FoldSelect: fold argument must be pos_ of the predicate (Vlite.hs:726-727)");
  i64 n = pred->n;
  int nt = omp_get_max_threads();
  i64 *cnt = (i64 *)calloc((size_t)nt + 1, sizeof(i64));
#pragma omp parallel num_threads(nt)
  {
    int t = omp_get_thread_num();
    i64 lo = n * t / nt, hi = n * (t + 1) / nt, c = 0;
    for (i64 i = lo; i < hi; i++) c += vget(pred, i) != 0;
    cnt[t + 1] = c;
  }
  for (int t = 0; t < nt; t++) cnt[t + 1] += cnt[t];
  *out = new_dense(cnt[nt]);
  out->domain = n;
  i64 *o = out->d;
#pragma omp parallel num_threads(nt)
  {
    int t = omp_get_thread_num();
    i64 lo = n * t / nt, hi = n * (t + 1) / nt, w = cnt[t];
    for (i64 i = lo; i < hi; i++) if (vget(pred, i) != 0) o[w++] = i;
  }
  free(cnt);
  return 0;
}

/* Gather: Vlite.hs:86-87 (@@), 1264, 1276-1277.  out[i] = src[pos[i]]. */
static int op_gather(orc_env *e, const vec *src, const vec *pos, vec *out) {
  *out = new_dense(pos->n);
  out->domain = src->domain;
  i64 n = pos->n, m = src->n, bad = 0; i64 *o = out->d;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (i64 i = 0; i < n; i++) {
    i64 p = vget(pos, i);
    if (p < 0 || p >= m) { bad++; o[i] = 0; } else o[i] = vget(src, p);
  }
  if (bad) { vfree(out); return fail(e, "Gather: %lld positions out of range [0,%lld)", (long long)bad, (long long)m); }
  return 0;
}

/* Scatter: Vlite.hs:1057-1059 (group sort), 1268-1275 (dim validity / inverse index).
 * G2: output length = index space of `pos` (domain), else max(pos)+1.  Unwritten slots = 0. */
static int op_scatter(orc_env *e, const vec *src, const vec *pos, vec *out) {
  if (src->n != pos->n) return fail(e, "Scatter: source length %lld != positions length %lld", (long long)src->n, (long long)pos->n);
  i64 n = src->n, len = pos->domain;
  if (len < 0) { len = 0; for (i64 i = 0; i < n; i++) { i64 p = vget(pos, i); if (p + 1 > len) len = p + 1; } }
  *out = new_dense(len);
  out->domain = src->domain;
  i64 *o = out->d, bad = 0;
  memset(o, 0, (size_t)len * sizeof(i64));
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (i64 i = 0; i < n; i++) {
    i64 p = vget(pos, i);
    if (p < 0 || p >= len) bad++; else o[p] = vget(src, i);
  }
  if (bad) { vfree(out); return fail(e, "Scatter: %lld positions out of range [0,%lld)", (long long)bad, (long long)len); }
  return 0;
}

/* CrossProductOuter / CrossProductInner (Vlite.hs:89-93, 283-292; joins under --use_cross_product, Mplan.hs:309-313): the
 * positions into `left` resp. `right` of the |left| x |right| pairs, left-major: "0,1,2,3 X 0,1 = 0,0,1,1,2,2,3,3 (outer),
 * 0,1,0,1,0,1,0,1 (inner)".  Their index space is the side they point into. */
static int op_cross(orc_env *e, const vec *left, const vec *right, int inner, vec *out) {
  i64 nl = left->n, nr = right->n;
  if (nl < 0 || nr < 0 || (nr > 0 && nl > (((i64)1 << 33) / nr))) return fail(e, "CrossProduct: %lld x %lld rows", (long long)nl, (long long)nr);
  i64 n = nl * nr;
  *out = new_dense(n);
  out->domain = inner ? nr : nl;
  i64 *o = out->d;
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; i++) o[i] = inner ? i % nr : i / nr;
  return 0;
}

/* Partition: Vlite.hs:1082-1098 emits Partition(key, RangeC(min,1,max-min+1)); the result is used as the
 * scatter positions that sort rows by key (1057-1060, "Assumes the fgroups are already sorted" 1172).
 * G3: bucket(v) = number of pivots strictly below v; result[i] = rank of row i in the stable bucket sort. */
static int op_partition(orc_env *e, const vec *data, const vec *piv, vec *out) {
  i64 n = data->n, np = piv->n;
  if (np < 1) return fail(e, "Partition: empty pivot vector");
  int range_piv = piv->kind == K_RANGE && piv->step > 0;
  if (!range_piv) for (i64 j = 1; j < np; j++) if (vget(piv, j - 1) > vget(piv, j)) return fail(e, "Partition: pivots must be ascending");
  u64 *key = (u64 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(u64));
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; i++) {
    i64 v = vget(data, i); u64 b;
    if (range_piv) {               /* pivots from, from+step, ...: #pivots < v */
      if (v <= piv->from) b = 0;
      else { u64 d = (u64)v - (u64)piv->from; b = (d + (u64)piv->step - 1) / (u64)piv->step; if (b > (u64)np) b = (u64)np; }
    } else {
      i64 lo = 0, hi = np;         /* first index with piv[idx] >= v */
      while (lo < hi) { i64 mid = (lo + hi) / 2; if (vget(piv, mid) < v) lo = mid + 1; else hi = mid; }
      b = (u64)lo;
    }
    key[i] = b;
  }
  /* keys already in bucket order: the stable sort is the identity (lineitem clustered on its order key: Q3's group-by) */
  i64 descents = 0;
#pragma omp parallel for schedule(static) reduction(+ : descents)
  for (i64 i = 1; i < n; i++) descents += key[i - 1] > key[i];
  *out = new_dense(n);
  out->domain = n;
  i64 *o = out->d;
  if (descents == 0) {
#pragma omp parallel for schedule(static)
    for (i64 j = 0; j < n; j++) o[j] = j;
    free(key);
    return 0;
  }
  /* stable LSD radix sort of row ids by key, 11 bits per pass over the bits in use; every pass in parallel: each thread
   * counts the digits of its contiguous slice, the offsets are the prefix over (digit, thread), each thread places its slice */
  u64 maxk = 0;
#pragma omp parallel for schedule(static) reduction(max : maxk)
  for (i64 i = 0; i < n; i++) if (key[i] > maxk) maxk = key[i];
  int bits = 0; while (bits < 64 && (maxk >> bits)) bits++;
  i64 *ord = (i64 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(i64)), *tmp = (i64 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(i64));
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < n; i++) ord[i] = i;
  int nt = omp_get_max_threads();
  if (n < 65536) nt = 1;
  i64 *hist = (i64 *)malloc((size_t)nt * 2048 * sizeof(i64));
  for (int sh = 0; sh < bits; sh += 11) {
#pragma omp parallel num_threads(nt)
    {
      int t = omp_get_thread_num();
      i64 lo = n * t / nt, hi = n * (t + 1) / nt, *h = hist + (size_t)t * 2048;
      memset(h, 0, 2048 * sizeof(i64));
      for (i64 i = lo; i < hi; i++) h[(key[ord[i]] >> sh) & 2047]++;
    }
    i64 run = 0;
    for (int d = 0; d < 2048; d++)
      for (int t = 0; t < nt; t++) { i64 c = hist[(size_t)t * 2048 + d]; hist[(size_t)t * 2048 + d] = run; run += c; }
#pragma omp parallel num_threads(nt)
    {
      int t = omp_get_thread_num();
      i64 lo = n * t / nt, hi = n * (t + 1) / nt, *h = hist + (size_t)t * 2048;
      for (i64 i = lo; i < hi; i++) tmp[h[(key[ord[i]] >> sh) & 2047]++] = ord[i];
    }
    i64 *sw = ord; ord = tmp; tmp = sw;
  }
#pragma omp parallel for schedule(static)
  for (i64 j = 0; j < n; j++) o[ord[j]] = j;
  free(key); free(ord); free(tmp); free(hist);
  return 0;
}

/* Folds: Vlite.hs:1048-1070, 1179; Vdl.hs:255-264.  One output per run of equal consecutive group values. */
static inline i64 fold_apply(int op, i64 acc, i64 v) {
  switch (op) {
    case OP_FSUM: return (i64)((u64)acc + (u64)v);
    case OP_FCOUNT: return acc + 1;
    case OP_FMIN: return v < acc ? v : acc;
    case OP_FMAX: return v > acc ? v : acc;
    default: return acc; /* FoldChoose keeps the first */
  }
}
static inline i64 fold_init(int op, i64 v) { return op == OP_FCOUNT ? 1 : v; }
static inline i64 fold_merge(int op, i64 acc, i64 part) {
  if (op == OP_FCOUNT) return acc + part;
  return fold_apply(op, acc, part);
}

static int op_fold(orc_env *e, int op, const vec *groups, const vec *data, vec *out) {
  if (groups->n != data->n) return fail(e, "Fold: groups length %lld != data length %lld", (long long)groups->n, (long long)data->n);
  i64 n = data->n;
  int nt = omp_get_max_threads();
  if (n < 4096) nt = 1;
  i64 *heads = (i64 *)calloc((size_t)nt + 1, sizeof(i64));
  i64 *lead = (i64 *)calloc((size_t)nt, sizeof(i64));
  char *has_lead = (char *)calloc((size_t)nt, 1);
#pragma omp parallel num_threads(nt)
  {
    int t = omp_get_thread_num();
    i64 lo = n * t / nt, hi = n * (t + 1) / nt, c = 0;
    for (i64 i = lo; i < hi; i++) c += (i == 0) || vget(groups, i) != vget(groups, i - 1);
    heads[t + 1] = c;
  }
  for (int t = 0; t < nt; t++) heads[t + 1] += heads[t];
  i64 nruns = heads[nt];
  *out = new_dense(nruns);
  i64 *o = out->d;
#pragma omp parallel num_threads(nt)
  {
    int t = omp_get_thread_num();
    i64 lo = n * t / nt, hi = n * (t + 1) / nt, r = heads[t] - 1;
    int in_lead = 1; i64 acc = 0; int have = 0;
    for (i64 i = lo; i < hi; i++) {
      int head = (i == 0) || vget(groups, i) != vget(groups, i - 1);
      i64 v = vget(data, i);
      if (head) {
        if (!in_lead) o[r] = acc;
        else if (have) { lead[t] = acc; has_lead[t] = 1; }
        in_lead = 0; r++; acc = fold_init(op, v); have = 1;
      } else if (have) acc = fold_apply(op, acc, v);
      else { acc = fold_init(op, v); have = 1; }
    }
    if (have) { if (!in_lead) o[r] = acc; else { lead[t] = acc; has_lead[t] = 1; } }
  }
  /* merge the leading partial of each chunk into the run it continues (sequential, deterministic) */
  for (int t = 1; t < nt; t++)
    if (has_lead[t] && op != OP_FCHOOSE) { i64 r = heads[t] - 1; o[r] = fold_merge(op, o[r], lead[t]); }
  free(heads); free(lead); free(has_lead);
  return 0;
}

/* Like: Vlite.hs:1010-1014 -> Vdl.hs:244-247.  SQL LIKE without escape: `%` matches any run of bytes (also none), `_`
 * exactly one byte, everything else itself.  Greedy match with one backtrack point (the last `%`). */
static int like_match(const unsigned char *s, const unsigned char *end, const char *p) {
  const unsigned char *star_s = NULL; const char *star_p = NULL;
  while (s < end && *s) {
    if (*p == '%') { star_p = ++p; star_s = s; }
    else if (*p && (*p == '_' || (unsigned char)*p == *s)) { p++; s++; }
    else if (star_p) { p = star_p; s = ++star_s; }
    else return 0;
  }
  while (*p == '%') p++;
  return *p == 0;
}
static int op_like(orc_env *e, const vec *data, const vec *heap, const char *pattern, vec *out) {
  if (heap->kind != K_BYTES) return fail(e, "Like: the dictionary must be a string heap (Load,<table>.<col>.heap)");
  if (data->kind == K_BYTES) return fail(e, "Like: the data must be a vector of heap offsets");
  i64 n = data->n, hl = heap->n, bad = 0;
  *out = new_dense(n);
  i64 *o = out->d;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (i64 i = 0; i < n; i++) {
    i64 off = vget(data, i);
    if (off < 0 || off >= hl) { bad++; o[i] = 0; } else o[i] = like_match(heap->d8 + off, heap->d8 + hl, pattern);
  }
  if (bad) { vfree(out); return fail(e, "Like: %lld offsets outside the heap [0,%lld)", (long long)bad, (long long)hl); }
  return 0;
}

/* ------------------------------------------------------------------ interpreter */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int orc_run(orc_env *e, const char *plan_text, int nthreads) {
  stmt *st = NULL; int n = 0;
  e->err[0] = 0;
  clear_outputs(e);
  if (parse_plan(e, plan_text, &st, &n)) return -1;
  e->nstmts = n;
  /* a Fold over the results of a Fold (make2LevelFold, Vlite.hs:1181-1192) re-reads the inner Fold's groups */
  for (int i = 0; i < n; i++) {
    int op = st[i].op;
    if ((op == OP_FCHOOSE || op == OP_FMAX || op == OP_FSUM || op == OP_FMIN || op == OP_FCOUNT) && st[i].b > 0) {
      stmt *inner = &st[st[i].b - 1];
      while (inner->op == OP_PROJECT || inner->op == OP_SHUFFLE) inner = &st[inner->a - 1];
      int iop = inner->op;
      if ((iop == OP_FCHOOSE || iop == OP_FMAX || iop == OP_FSUM || iop == OP_FMIN || iop == OP_FCOUNT) && inner->a > 0 && st[inner->a - 1].lastuse < i)
        st[inner->a - 1].lastuse = i;
    }
  }
  if (nthreads > 0) omp_set_num_threads(nthreads);
  vec *val = (vec *)calloc((size_t)n + 1, sizeof(vec));
  e->outs = (output *)calloc((size_t)n + 1, sizeof(output));
  /* storage ownership: Project/Shuffle/MaterializeCompact alias their argument's storage; a dense
     vector is freed after the last statement that reads it or any of its aliases */
  int *owner = (int *)calloc((size_t)n + 1, sizeof(int)), *lastown = (int *)calloc((size_t)n + 1, sizeof(int));
  for (int i = 0; i < n; i++) {
    int id = st[i].id, op = st[i].op;
    owner[id] = (op == OP_PROJECT || op == OP_SHUFFLE || op == OP_MATERIALIZE) ? owner[st[i].a] : id;
  }
  for (int i = 0; i < n; i++) { int o = owner[st[i].id]; if (st[i].lastuse > lastown[o]) lastown[o] = st[i].lastuse; }
  int rc = 0;
  double t0 = now_s();
  for (int i = 0; i < n && !rc; i++) {
    stmt *s = &st[i];
    vec *A = s->a ? &val[s->a] : NULL, *B = s->b ? &val[s->b] : NULL, *C = s->c ? &val[s->c] : NULL;
    vec r; memset(&r, 0, sizeof r); r.domain = -1;
    if (s->op != OP_LIKE && s->op != OP_PROJECT && s->op != OP_LOAD &&
        ((A && A->kind == K_BYTES) || (B && B->kind == K_BYTES) || (C && C->kind == K_BYTES))) {
      rc = fail(e, "statement %d: a string heap can only be the dictionary of a Like", s->id); break;
    }
    switch (s->op) {
      case OP_LOAD: {
        colbind *cb = NULL;
        for (int k = 0; k < e->ncols; k++) if (!strcmp(e->cols[k].name, s->name)) cb = &e->cols[k];
        if (!cb) { rc = fail(e, "Load: column %s is not bound", s->name); break; }
        r.valid = 1; r.n = cb->rows;
        if (cb->width == 1) { r.kind = K_BYTES; r.d8 = (const unsigned char *)cb->data; }
        else if (cb->width == 4) { r.kind = K_COL32; r.d32 = (const int32_t *)cb->data; }
        else { r.kind = K_COL64; r.d = (i64 *)cb->data; }
        break;
      }
      case OP_PROJECT: case OP_SHUFFLE: r = *A; if (r.kind == K_DENSE) { /* share storage: alias, never freed twice */
          r.kind = K_COL64; } break;
      case OP_RANGEV: r.valid = 1; r.kind = K_RANGE; r.n = A->n; r.from = s->k0; r.step = s->k1;
        if (s->k0 == 0 && s->k1 == 1) r.domain = A->n;   /* pos_ v indexes v's space */
        break;
      case OP_RANGEC: r.valid = 1; r.kind = K_RANGE; r.n = s->k1; r.from = s->k0; r.step = s->k2; break;
      case OP_FSELECT: rc = op_fold_select(e, A, B, &r); break;
      case OP_GATHER: rc = op_gather(e, A, B, &r); break;
      case OP_SCATTER: rc = op_scatter(e, A, C, &r); break;
      case OP_PARTITION: rc = op_partition(e, A, B, &r); break;
      case OP_LIKE: rc = op_like(e, A, B, s->name, &r); break;
      case OP_CROSS_OUTER: case OP_CROSS_INNER: rc = op_cross(e, A, B, s->op == OP_CROSS_INNER, &r); break;
      case OP_FCHOOSE: case OP_FMAX: case OP_FSUM: case OP_FMIN: case OP_FCOUNT: {
        if (A->n != B->n) {   /* level 2 of a hierarchical fold: group the level-1 results by the groups at each run's head */
          stmt *inner = &st[s->b - 1];
          while (inner->op == OP_PROJECT || inner->op == OP_SHUFFLE) inner = &st[inner->a - 1];
          int iop = inner->op;
          vec *g1 = (iop == OP_FCHOOSE || iop == OP_FMAX || iop == OP_FSUM || iop == OP_FMIN || iop == OP_FCOUNT) ? &val[inner->a] : NULL;
          if (g1 && g1->valid && g1->n == A->n) {
            vec eff; memset(&eff, 0, sizeof eff);
            rc = op_fold(e, OP_FCHOOSE, g1, A, &eff);
            if (!rc) { rc = op_fold(e, s->op, &eff, B, &r); vfree(&eff); }
            break;
          }
        }
        rc = op_fold(e, s->op, A, B, &r); break;
      }
      case OP_MATERIALIZE: {
        output *o = &e->outs[e->nouts++];
        const char *nm = st[s->a - 1].op == OP_PROJECT ? st[s->a - 1].name : "val";
        snprintf(o->name, sizeof o->name, "%s", nm);
        o->n = A->n; o->d = (i64 *)malloc((size_t)(A->n > 0 ? A->n : 1) * sizeof(i64));
        {
          i64 *od = o->d; const i64 an = A->n;
#pragma omp parallel for schedule(static)
          for (i64 k = 0; k < an; k++) od[k] = vget(A, k);
        }
        r = *A; if (r.kind == K_DENSE) r.kind = K_COL64;
        break;
      }
      default: rc = op_binary(e, s->op, A, B, &r); break;
    }
    if (rc) break;
    val[s->id] = r;
    int args[3] = {s->a, s->b, s->c};
    for (int k = 0; k < 3; k++)
      if (args[k] > 0) { int o = owner[args[k]]; if (lastown[o] == i) vfree(&val[o]); }
  }
  e->seconds = now_s() - t0;
  for (int i = 1; i <= n; i++) vfree(&val[i]);
  free(val); free(st); free(owner); free(lastown);
  return rc;
}

/* ------------------------------------------------------------------ synthetic columns
 * Counter-based generator shared (as a specification, not as code) with the CUDA library
 * (mplan2vdl_b200/csrc/vdl_synth.cuh): value depends only on (seed, colid, global row).
 *   base = splitmix64(seed ^ (colid * 0x9E3779B97F4A7C15));  h = splitmix64(base + row)
 *   kind 0 UNIFORM : vmin + stride * mulhi64(h, p0)           p0 = number of distinct values
 *   kind 1 SEQ     : vmin + stride * row
 *   kind 2 FKDENSE : vmin + stride * ((row * p0) / p1)        nondecreasing FK, p0 = dim rows, p1 = fact rows
 * Recipe per column: SURVEY.md Appendix D (bounds from tests/tpch10noorder/bounds.csv). */
static inline u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ULL;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
void orc_gen_column(void *out, int width, i64 rows, i64 row_offset, u64 seed, u64 colid, int kind,
                    i64 vmin, i64 stride, i64 p0, i64 p1, int nthreads) {
  u64 base = splitmix64(seed ^ (colid * 0x9E3779B97F4A7C15ULL));
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
  for (i64 i = 0; i < rows; i++) {
    u64 row = (u64)(row_offset + i);
    i64 v;
    if (kind == 0) {
      u64 h = splitmix64(base + row);
      u64 idx = (u64)(((unsigned __int128)h * (unsigned __int128)(u64)p0) >> 64);
      v = (i64)((u64)vmin + (u64)stride * idx);
    } else if (kind == 1) {
      v = (i64)((u64)vmin + (u64)stride * row);
    } else {
      v = (i64)((u64)vmin + (u64)stride * ((row * (u64)p0) / (u64)p1));
    }
    if (width == 4) ((int32_t *)out)[i] = (int32_t)v; else ((i64 *)out)[i] = v;
  }
}

int orc_max_threads(void) { return omp_get_max_threads(); }
