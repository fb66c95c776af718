/*
 * q6_host.c -- a host with no Python and no torch in it: plain C over the C ABI of libvdl_cuda (include/vdl_cuda.h), the way
 * a compiled host such as the reference's Haskell binary would drive it.
 *
 *   cc -std=c11 -O2 -Iinclude examples/q6_host.c -o q6_host -Lmplan2vdl_b200 -lvdl_cuda -Wl,-rpath,$PWD/mplan2vdl_b200
 *   ./q6_host plans/q06.vdl [rows] [device]
 *
 * It registers four lineitem columns (random values inside the reference's bounds.csv ranges, stored in the reference's
 * widths: three decimals as int64, the date as int32 -- Types.hs:129-140), hands the library the program text mplan2vdl
 * prints for TPC-H Q6 (README.md:40-52), runs it, and checks the one output against the query evaluated by a scalar loop in
 * this file (what the SQL says: sum(l_extendedprice * l_discount) over 1994, discount 0.05..0.07, quantity < 24).
 * Exit code 0 = the library's answer is that number; any failure of the library is reported with vdl_last_error and a
 * non-zero exit code -- there is no CPU fallback.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vdl_cuda.h"

static uint64_t rng_state = 0x5EEDull;
static uint64_t rnd(void) {                     /* splitmix64 */
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static int64_t between(int64_t lo, int64_t hi) { return lo + (int64_t)(rnd() % (uint64_t)(hi - lo + 1)); }

static char *slurp(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *s = (char *)malloc((size_t)n + 1);
  if (s && fread(s, 1, (size_t)n, f) != (size_t)n) { free(s); s = NULL; }
  if (s) s[n] = 0;
  fclose(f);
  return s;
}

#define CHECK(call)                                                                                 \
  do {                                                                                              \
    int rc_ = (call);                                                                               \
    if (rc_) { fprintf(stderr, "q6_host: %s failed (%d): %s\n", #call, rc_, vdl_last_error(ctx)); return 2; } \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s plans/q06.vdl [rows] [device]\n", argv[0]); return 64; }
  const int64_t rows = argc > 2 ? atoll(argv[2]) : 1000000;
  const int device = argc > 3 ? atoi(argv[3]) : 0;
  char *text = slurp(argv[1]);
  if (!text) { fprintf(stderr, "q6_host: cannot read %s\n", argv[1]); return 66; }

  /* columns in the reference's stored representation and value ranges (tests/tpch10noorder/bounds.csv): decimals are scaled
   * integers -- quantity 1.00..50.00 = 100..5000 in steps of 100, extendedprice in cents, discount 0.00..0.10 = 0..10 --
   * and dates are day numbers (Mplan.hs:46-57: 1992-01-01 = 727564 ... 1998-12-01 = 730089) */
  int64_t *qty = (int64_t *)malloc((size_t)rows * 8), *price = (int64_t *)malloc((size_t)rows * 8), *disc = (int64_t *)malloc((size_t)rows * 8);
  int32_t *ship = (int32_t *)malloc((size_t)rows * 4);
  if (!qty || !price || !disc || !ship) { fprintf(stderr, "q6_host: out of host memory\n"); return 71; }
  int64_t want = 0;
  for (int64_t i = 0; i < rows; i++) {
    qty[i] = 100 * between(1, 50);
    price[i] = between(90091, 10494950);
    disc[i] = between(0, 10);
    ship[i] = (int32_t)between(727564, 730089);
    /* Q6: l_shipdate >= 1994-01-01 (728294) and < 1995-01-01 (728659), l_discount between 0.06 -+ 0.01, l_quantity < 24 */
    if (ship[i] >= 728294 && ship[i] < 728659 && disc[i] >= 5 && disc[i] <= 7 && qty[i] < 24 * 100) want += price[i] * disc[i];
  }

  vdl_ctx *ctx = NULL;
  int rc = vdl_ctx_create(device, &ctx);
  if (rc) { fprintf(stderr, "q6_host: vdl_ctx_create(%d) failed (%d): %s\n", device, rc, vdl_last_error(NULL)); return 2; }
  struct { const char *name; int dtype; const void *data; } cols[4] = {
      {"lineitem.l_quantity", VDL_I64, qty}, {"lineitem.l_extendedprice", VDL_I64, price},
      {"lineitem.l_discount", VDL_I64, disc}, {"lineitem.l_shipdate", VDL_I32, ship}};
  for (int c = 0; c < 4; c++) {
    vdl_vec v;
    CHECK(vdl_column_alloc(ctx, cols[c].name, cols[c].dtype, rows, &v));
    CHECK(vdl_column_upload(ctx, v, cols[c].data, rows));
  }
  vdl_plan *plan = NULL;
  CHECK(vdl_plan_load(ctx, text, VDL_PLAN_FUSE, &plan));
  CHECK(vdl_plan_run(plan));
  int statements = 0, nodes = 0, scans = 0;
  int64_t launches = 0;
  CHECK(vdl_plan_stats(plan, &statements, &nodes, &scans, &launches));
  const char *name = NULL;
  const int64_t *data = NULL;
  int64_t len = 0;
  CHECK(vdl_plan_output(plan, 0, &name, &data, &len));
  /* no row selected: the Fold of an empty vector is an empty vector (the sparse model has no run to put a sum in) */
  const int64_t got = len == 1 ? data[0] : 0;
  printf("{\"plan\": \"%s\", \"rows\": %lld, \"statements\": %d, \"fused_scans\": %d, \"launches\": %lld, \"%s\": %lld, \"expected\": %lld}\n", argv[1],
         (long long)rows, statements, scans, (long long)launches, name ? name : "?", (long long)got, (long long)want);
  const int ok = (len == 1 && got == want) || (len == 0 && want == 0);
  vdl_plan_destroy(plan);
  vdl_ctx_destroy(ctx);
  free(qty); free(price); free(disc); free(ship); free(text);
  if (!ok) { fprintf(stderr, "q6_host: MISMATCH\n"); return 1; }
  return 0;
}
