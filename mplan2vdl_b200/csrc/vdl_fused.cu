// Fused select -> map -> fold: ONE launch that streams the columns of a table shard through shared
// memory with TMA bulk copies and folds products of affine column terms per group key.
//
// What it replaces in the emitted graph (reference Vlite.hs):
//   idx = FoldSelect(pos_ p, p); c_i @@ idx          solve' Select          Vlite.hs:721-730
//   key = ((c >> tz) - min) << bits | ... & hint     makeCompositeKey       Vlite.hs:1123-1170, 1111-1115
//   Partition(key) -> Scatter(x, .) -> Fold(op, ., .) solveAgg / getScatterMask Vlite.hs:1048-1098
// In the dense model elementwise ops commute with the Gather by the selection, the
// stable Partition+Scatter only orders rows by key, and Fold emits one row per existing key in ascending
// order -- so the whole chain is "for every row that passes p: acc[key][j] op= value_j(row)", followed
// by dropping the keys that saw no row.  No intermediate vector ever reaches HBM.
//
// Kernel shape (B200, sm_100a): persistent, one CTA per SM.  Warp 0 is the producer: one elected lane
// issues `cp.async.bulk` (TMA, 1-D) copies of TILE rows of every column into a ring of shared-memory
// stages, completion tracked by mbarriers (expect_tx).  The other 8 warps are consumers: they wait on
// the stage's "full" barrier, evaluate predicate / key / products straight from shared memory (column
// chosen by runtime index = just an address, so ONE kernel serves every plan), and release the stage
// through its "empty" barrier.  Ungrouped folds accumulate in registers; grouped folds accumulate in
// lane-private shared-memory tables (plain LDS/STS read-modify-write, no atomics, no bank conflicts)
// over a per-CTA compacted slot map so that sparse key domains (Q1: 6 of 32 slots) stay small.
// HBM traffic = the columns, once.  Roofline: HBM.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "vdl_internal.h"

#define K_MAX_ACC 10
#define K_MAX_CHOOSE 6

// soff / w4 are derived by the host from col: byte offset of the column inside a staged tile, 4-byte flag.
struct KAffine { int32_t col, shr; i64 a, b; int32_t soff, w4; int32_t narrow, pad; };   // narrow: host-side, see FF_NARROW
struct KPred { int32_t col, shr; i64 lo; u64 span; int32_t soff, w4; int32_t lo32; uint32_t span32; };
struct KKey { KAffine e; int32_t shl, pad; };
// op: 0 sum, 1 min, 2 max.  chain: value = value of the previous accumulator x own factors (prefix sharing:
// ep, ep*(100-d), ep*(100-d)*(100+t) evaluate each factor once).
struct KAcc { int32_t op, nfac, chain, pad; KAffine fac[VDL_MAX_FACTORS]; };

struct KDesc {
  i64 rows, row_base, key_mask, domain, ntiles;
  const void *col[VDL_MAX_COLS];
  int32_t width[VDL_MAX_COLS], soff[VDL_MAX_COLS];
  int32_t ncols, npreds, nkeys, nacc, nchoose, cnt_idx, first_idx;
  int32_t tile_rows, stages, stage_bytes, stage_tx, gmax, pad0;
  KPred pred[VDL_MAX_PREDS];
  KKey key[VDL_MAX_KEYS];
  KAcc acc[K_MAX_ACC];
  KAcc choose[K_MAX_CHOOSE];
  i64 *table;                    // [nacc + nchoose][domain]
  int *errflag;
  // Epilogue run by the LAST CTA to finish (ticket counter `done`): 0 none, 1 FoldChoose values only (the partial
  // table is then complete for an external all-gather), 2 FoldChoose + finalize + table reset (single GPU: the scan
  // is the only launch of a step), 3 the same with the peer-memory exchange of the tables before the finalize.
  unsigned int *done;
  int32_t epilogue, pad1;
};

struct FinDesc {
  const i64 *parts;              // nranks tables back to back, each part_stride int64
  i64 part_stride, domain;
  int32_t nranks, nacc, nchoose, cnt_idx, first_idx, nout;
  int32_t acc_op[K_MAX_ACC];
  int32_t out_kind[VDL_MAX_AGGS], out_idx[VDL_MAX_AGGS];   // kind 0: accumulator, 1: choose
  i64 *out[VDL_MAX_AGGS];
  int32_t npost, pad;
  vdl_post_op post[VDL_MAX_POSTS];
  i64 *post_out[VDL_MAX_POSTS];
  i64 *ngroups;                  // [0] number of groups, [1] snapshot of the context's error counter
  const int *errflag;
  i64 *hmirror;                  // mapped pinned host copy of the whole result buffer (same layout as out[0]...), or null
  i64 *reset_table;              // this rank's partial table, re-initialised for the next launch after the merge, or null
};

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier; L2 evict-first policy
// because every byte of a scan is touched exactly once.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <int NC>
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

__device__ __forceinline__ i64 acc_identity(int op) { return op == 0 ? 0 : (op == 1 ? INT64_MAX : INT64_MIN); }
__device__ __forceinline__ i64 acc_combine(int op, i64 a, i64 v) {
  return op == 0 ? (i64)((u64)a + (u64)v) : (op == 1 ? (v < a ? v : a) : (v > a ? v : a));
}
__device__ __forceinline__ void acc_global(int op, i64 *p, i64 v) {
  if (op == 0) atomicAdd((unsigned long long *)p, (unsigned long long)v);
  else if (op == 1) atomicMin((long long *)p, (long long)v);
  else atomicMax((long long *)p, (long long)v);
}
__device__ __forceinline__ i64 warp_reduce(int op, i64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = acc_combine(op, v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- descriptor-driven evaluation over a staged tile -------------------------------------------
// Every loop over the descriptor (predicates, key parts, accumulators, factors) is unrolled as a chain of
// nested uniform `if (I < n)` tests, so each descriptor field is a constant-bank immediate operand of the
// instruction that uses it (no loads, no dependent latency) and unused slots cost one uniform branch.
//
// SHAPES.  The same code is instantiated over a shape-traits class S.  GenericShape answers every structural
// question (how many predicates / key parts / accumulators / factors, 4- or 8-byte column, shift or not, b == 1,
// a == 0, fits int32 ...) from the descriptor at run time, so one kernel runs any plan.  A static shape
// (vdl_shapes.cuh) answers them at compile time: the uniform tests fold away and arithmetic narrows to 32 bits
// where the column statistics allow (the executor's use of the reference's bound inference, Vlite.hs:417-467),
// leaving straight-line code -- while every constant (bounds, offsets, a, b, shifts, masks) stays a run-time
// descriptor field.  The host launches a static instantiation only when the prepared descriptor satisfies every
// assumption the shape makes (shape_matches), otherwise the generic one: identical results either way.

// factor flags (what the static code may assume about a KAffine)
#define FF_W4 1        // 4-byte column (else 8-byte)
#define FF_SHR0 2      // shr == 0
#define FF_B1 4        // b == 1
#define FF_BM1 8       // b == -1
#define FF_A0 16       // a == 0
#define FF_CONST 32    // col == -1
#define FF_ROWID 64    // col == -2
#define FF_NARROW 128  // leaf and a + b*leaf fit int32 (column statistics)
// how a register-slot kernel keeps an accumulator per thread (see vdl_shapes.cuh)
#define RK_WIDE 0
#define RK_N32 1
#define RK_FIRST 2
#define RK_MADW 3

struct GenericShape {
  static constexpr bool kStatic = false;
  static constexpr int NPREDS = 0, NKEYS = 0, NACC = 0, KEY32 = 0;
  static constexpr int PRED_MODE[VDL_MAX_PREDS] = {}, PRED_SHR0[VDL_MAX_PREDS] = {};
  static constexpr int KEY_FLAGS[VDL_MAX_KEYS] = {}, KEY_SHL0[VDL_MAX_KEYS] = {};
  static constexpr int ACC_OP[K_MAX_ACC] = {}, ACC_CHAIN[K_MAX_ACC] = {}, ACC_NFAC[K_MAX_ACC] = {};
  static constexpr int FAC[K_MAX_ACC][VDL_MAX_FACTORS] = {};
  static constexpr int RS_G = 0;
  static constexpr int ACC_RK[K_MAX_ACC] = {};
};

__device__ __forceinline__ i64 tile_leaf(const KAffine &A, const unsigned char *tile, int r, i64 grow) {
  if (A.col == -2) return grow;
  i64 v = A.w4 ? (i64)((const int32_t *)(tile + A.soff))[r] : ((const i64 *)(tile + A.soff))[r];
  if (A.shr) v >>= A.shr;
  return v;
}
// a + b * leaf; col -1: the constant a; col -2: leaf = global row id
__device__ __forceinline__ i64 tile_affine(const KAffine &A, const unsigned char *tile, int r, i64 grow) {
  if (A.col == -1) return A.a;
  i64 leaf = tile_leaf(A, tile, r, grow);
  if (A.b != 1) leaf = (i64)((u64)A.b * (u64)leaf);
  return (i64)((u64)A.a + (u64)leaf);
}
// the same with the structure known at compile time; 32-bit arithmetic when FF_NARROW
template <int FL>
__device__ __forceinline__ int32_t affine32(const KAffine &A, const unsigned char *tile, int r) {
  int32_t v = ((const int32_t *)(tile + A.soff))[(FL & FF_W4) ? r : 2 * r];   // low word of an 8-byte value
  if (!(FL & FF_SHR0)) v >>= A.shr;
  if (FL & FF_BM1) v = -v;
  else if (!(FL & FF_B1)) v *= (int32_t)A.b;
  if (!(FL & FF_A0)) v += (int32_t)A.a;
  return v;
}
template <int FL>
__device__ __forceinline__ i64 affine_static(const KAffine &A, const unsigned char *tile, int r, i64 grow) {
  if constexpr (FL & FF_CONST) return A.a;
  else if constexpr (FL & FF_NARROW) return (i64)affine32<FL>(A, tile, r);
  else {
    i64 v;
    if constexpr (FL & FF_ROWID) v = grow;
    else {
      v = (FL & FF_W4) ? (i64)((const int32_t *)(tile + A.soff))[r] : ((const i64 *)(tile + A.soff))[r];
      if (!(FL & FF_SHR0)) v >>= A.shr;
    }
    if (FL & FF_BM1) v = (i64)(0 - (u64)v);
    else if (!(FL & FF_B1)) v = (i64)((u64)A.b * (u64)v);
    if (!(FL & FF_A0)) v = (i64)((u64)A.a + (u64)v);
    return v;
  }
}

template <class S, int J, int F>
__device__ __forceinline__ i64 factor_chain(const KAcc &A, i64 v, const unsigned char *tile, int r, i64 grow) {
  if constexpr (F < VDL_MAX_FACTORS) {
    if constexpr (S::kStatic) {
      if constexpr (F < S::ACC_NFAC[J]) {
        i64 x = affine_static<S::FAC[J][F]>(A.fac[F], tile, r, grow);
        v = (F == 0 && !S::ACC_CHAIN[J]) ? x : (i64)((u64)v * (u64)x);
        return factor_chain<S, J, F + 1>(A, v, tile, r, grow);
      }
    } else {
      if (F < A.nfac) {
        i64 x = tile_affine(A.fac[F], tile, r, grow);
        v = (F == 0 && !A.chain) ? x : (i64)((u64)v * (u64)x);
        return factor_chain<S, J, F + 1>(A, v, tile, r, grow);
      }
    }
  }
  return v;
}
// generic (looping) form for the rare global-atomic path
__device__ __noinline__ i64 acc_value_slow(const KDesc &d, int j, const unsigned char *tile, int r, i64 grow) {
  i64 v = 1;
  int j0 = j;
  while (d.acc[j0].chain) j0--;
  for (int q = j0; q <= j; q++)
    for (int f = 0; f < d.acc[q].nfac; f++) v = (i64)((u64)v * (u64)tile_affine(d.acc[q].fac[f], tile, r, grow));
  return v;
}

template <class S, int I, int NC, int R>
__device__ __forceinline__ void pred_chain(const KDesc &d, const unsigned char *tile, int ctid, unsigned &pass) {
  if constexpr (I < VDL_MAX_PREDS && (!S::kStatic || I < S::NPREDS)) {
    if (S::kStatic || I < d.npreds) {
      const KPred &P = d.pred[I];
      const int mode = S::kStatic ? S::PRED_MODE[I] : P.w4;
      const bool shr0 = S::kStatic ? (bool)S::PRED_SHR0[I] : (P.shr == 0);
      if (mode) {      // 32-bit compare against bounds clamped to int32 by the host: a 4-byte column (mode 1), or
                       // the low words of an 8-byte column whose values all fit int32 per its statistics (mode 2)
        const int32_t *p = (const int32_t *)(tile + P.soff);
#pragma unroll
        for (int k = 0; k < R; k++) {
          int32_t v = p[(ctid + k * NC) * mode];
          if (!shr0) v >>= P.shr;
          if ((uint32_t)v - (uint32_t)P.lo32 > P.span32) pass &= ~(1u << k);
        }
      } else {
        const i64 *p = (const i64 *)(tile + P.soff);
#pragma unroll
        for (int k = 0; k < R; k++) {
          i64 v = p[ctid + k * NC];
          if (!shr0) v >>= P.shr;
          if ((u64)v - (u64)P.lo > P.span) pass &= ~(1u << k);
        }
      }
      pred_chain<S, I + 1, NC, R>(d, tile, ctid, pass);
    }
  }
}

template <class S, int Q>
__device__ __forceinline__ i64 key_chain(const KDesc &d, i64 key, const unsigned char *tile, int r, i64 grow) {
  if constexpr (Q < VDL_MAX_KEYS && (!S::kStatic || Q < S::NKEYS)) {
    if (S::kStatic || Q < d.nkeys) {
      i64 x;
      if constexpr (S::kStatic) x = affine_static<S::KEY_FLAGS[Q]>(d.key[Q].e, tile, r, grow);
      else x = tile_affine(d.key[Q].e, tile, r, grow);
      if (!(S::kStatic && S::KEY_SHL0[Q]) && d.key[Q].shl) x = (i64)((u64)x << d.key[Q].shl);
      return key_chain<S, Q + 1>(d, key | x, tile, r, grow);
    }
  }
  return key;
}
// all key parts narrow: the whole key in 32-bit arithmetic
template <class S, int Q>
__device__ __forceinline__ int32_t key_chain32(const KDesc &d, int32_t key, const unsigned char *tile, int r) {
  if constexpr (Q < S::NKEYS) {
    int32_t x = affine32<S::KEY_FLAGS[Q]>(d.key[Q].e, tile, r);
    if (!S::KEY_SHL0[Q]) x <<= d.key[Q].shl;
    return key_chain32<S, Q + 1>(d, key | x, tile, r);
  }
  return key;
}

// lane-private read-modify-write of accumulator J and all following ones
template <class S, int J, int NC>
__device__ __forceinline__ void acc_chain(const KDesc &d, i64 *t, i64 prev, const unsigned char *tile, int r, i64 grow) {
  if constexpr (J < K_MAX_ACC && (!S::kStatic || J < S::NACC)) {
    if (S::kStatic || J < d.nacc) {
      const bool chain = S::kStatic ? (bool)S::ACC_CHAIN[J] : (bool)d.acc[J].chain;
      const int op = S::kStatic ? S::ACC_OP[J] : d.acc[J].op;
      i64 v = factor_chain<S, J, 0>(d.acc[J], chain ? prev : 1, tile, r, grow);
      t[J * NC] = acc_combine(op, t[J * NC], v);
      acc_chain<S, J + 1, NC>(d, t, v, tile, r, grow);
    }
  }
}

// ---- register slots (G > 0): per-thread accumulators of the first G keys of the CTA live in registers ----------
// All indexing below is static after unrolling, so the arrays are plain registers and entries of the kind an
// accumulator does not use are never materialised.
template <class S, int G>
struct RegAcc {
  static constexpr int NA = (S::kStatic && S::NACC > 0) ? S::NACC : 1;
  static constexpr int NG = G > 0 ? G : 1;
  i64 w[NG][NA];
  int32_t n[NG][NA];
};
template <class S, int G, int J>
__device__ __forceinline__ void rs_init(RegAcc<S, G> &ra) {
  if constexpr (G > 0 && J < S::NACC) {
#pragma unroll
    for (int g = 0; g < G; g++) {
      constexpr int op = S::ACC_OP[J];
      if constexpr (S::ACC_RK[J] == RK_WIDE || S::ACC_RK[J] == RK_MADW) ra.w[g][J] = acc_identity(op);
      else if constexpr (S::ACC_RK[J] == RK_N32) ra.n[g][J] = 0;
      else ra.n[g][J] = INT32_MAX;
    }
    rs_init<S, G, J + 1>(ra);
  }
}
// product of own factors F0 .. F1-1 of accumulator J onto v
template <class S, int J, int F, int F1>
__device__ __forceinline__ i64 factor_range(const KAcc &A, i64 v, bool have, const unsigned char *tile, int r, i64 grow) {
  if constexpr (F < F1) {
    i64 x = affine_static<S::FAC[J][F]>(A.fac[F], tile, r, grow);
    return factor_range<S, J, F + 1, F1>(A, have ? (i64)((u64)v * (u64)x) : x, true, tile, r, grow);
  }
  return v;
}
// values of every accumulator for one row (prefix-shared exactly like acc_chain).  RK_MADW accumulators are kept as
// the pair (a, b) with value a * b, both int32 by the host's proof, so that the update is one `mad.wide.s32`.
template <class S, int J>
__device__ __forceinline__ void rs_values(const KDesc &d, i64 *v, int32_t *va, int32_t *vb, i64 prev, const unsigned char *tile, int r, i64 grow) {
  if constexpr (J < S::NACC) {
    constexpr bool chain = (bool)S::ACC_CHAIN[J];
    constexpr int nfac = S::ACC_NFAC[J];
    if constexpr (S::ACC_RK[J] == RK_FIRST) {
      v[J] = 0;   // the CTA-local row index is supplied by the caller
      rs_values<S, J + 1>(d, v, va, vb, prev, tile, r, grow);
    } else if constexpr (S::ACC_RK[J] == RK_MADW) {
      i64 a = factor_range<S, J, 0, nfac - 1>(d.acc[J], chain ? prev : 1, chain, tile, r, grow);
      va[J] = (int32_t)a;
      vb[J] = (int32_t)affine_static<S::FAC[J][nfac - 1]>(d.acc[J].fac[nfac - 1], tile, r, grow);
      v[J] = (i64)va[J] * (i64)vb[J];
      rs_values<S, J + 1>(d, v, va, vb, v[J], tile, r, grow);
    } else {
      v[J] = factor_chain<S, J, 0>(d.acc[J], chain ? prev : 1, tile, r, grow);
      rs_values<S, J + 1>(d, v, va, vb, v[J], tile, r, grow);
    }
  }
}
// Predicated updates (`@p add` / `@p min`): the slot test never branches, so a warp whose 32 rows hit 6 different
// slots runs ONE straight instruction stream instead of 6 divergent switch arms (the compiler turns an if-chain
// over `slot == g` into a jump table, which serialises the arms and stalls on every indirect branch).
__device__ __forceinline__ void pred_add64(i64 &acc, i64 v, int s, int g) {
  // split add with carry: ptxas keeps the low add predicated (a predicated add.s64 / mad.wide becomes add + 2 SEL)
  uint32_t lo = (uint32_t)acc, hi = (uint32_t)((u64)acc >> 32);
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %4, %5;\n @p add.cc.u32 %0, %0, %2;\n @p addc.u32 %1, %1, %3;\n}"
      : "+r"(lo), "+r"(hi) : "r"((uint32_t)v), "r"((uint32_t)((u64)v >> 32)), "r"(s), "r"(g));
  acc = (i64)(((u64)hi << 32) | lo);
}
__device__ __forceinline__ void pred_min64(i64 &acc, i64 v, int s, int g) {
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %2, %3;\n @p min.s64 %0, %0, %1;\n}" : "+l"(acc) : "l"(v), "r"(s), "r"(g));
}
__device__ __forceinline__ void pred_max64(i64 &acc, i64 v, int s, int g) {
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %2, %3;\n @p max.s64 %0, %0, %1;\n}" : "+l"(acc) : "l"(v), "r"(s), "r"(g));
}
__device__ __forceinline__ void pred_madw(i64 &acc, int32_t a, int32_t b, int s, int g) {
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %3, %4;\n @p mad.wide.s32 %0, %1, %2, %0;\n}" : "+l"(acc) : "r"(a), "r"(b), "r"(s), "r"(g));
}
__device__ __forceinline__ void pred_add32(int32_t &acc, int32_t v, int s, int g) {
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %2, %3;\n @p add.s32 %0, %0, %1;\n}" : "+r"(acc) : "r"(v), "r"(s), "r"(g));
}
__device__ __forceinline__ void pred_min32(int32_t &acc, int32_t v, int s, int g) {
  asm("{\n .reg .pred p;\n setp.eq.s32 p, %2, %3;\n @p min.s32 %0, %0, %1;\n}" : "+r"(acc) : "r"(v), "r"(s), "r"(g));
}
// slot Q's accumulators take the row's values iff s == Q
template <class S, int G, int Q, int J>
__device__ __forceinline__ void rs_apply(RegAcc<S, G> &ra, int s, const i64 *v, const int32_t *va, const int32_t *vb, int lrow) {
  if constexpr (J < S::NACC) {
    constexpr int op = S::ACC_OP[J];
    if constexpr (S::ACC_RK[J] == RK_MADW) {
      // acc += (slot matches ? a : 0) * b as ONE 32x32+64 multiply-add (IMAD.WIDE.U32, fma pipe) after one select that
      // accumulators sharing `a` share; 0 <= a, b < 2^31 by the host's proof.  (A predicated 64-bit add costs 3.)
      const uint32_t am = (s == Q) ? (uint32_t)va[J] : 0u;
      ra.w[Q][J] = (i64)((u64)ra.w[Q][J] + (u64)am * (u64)(uint32_t)vb[J]);
    }
    else if constexpr (S::ACC_RK[J] == RK_WIDE) {
      if constexpr (op == 0) pred_add64(ra.w[Q][J], v[J], s, Q);
      else if constexpr (op == 1) pred_min64(ra.w[Q][J], v[J], s, Q);
      else pred_max64(ra.w[Q][J], v[J], s, Q);
    } else if constexpr (S::ACC_RK[J] == RK_N32) pred_add32(ra.n[Q][J], (int32_t)v[J], s, Q);
    else pred_min32(ra.n[Q][J], lrow, s, Q);
    rs_apply<S, G, Q, J + 1>(ra, s, v, va, vb, lrow);
  }
}
template <class S, int G, int GL, int Q>
__device__ __forceinline__ void rs_apply_slots(RegAcc<S, G> &ra, int s, const i64 *v, const int32_t *va, const int32_t *vb, int lrow) {
  if constexpr (Q < GL) {
    rs_apply<S, G, Q, 0>(ra, s, v, va, vb, lrow);
    rs_apply_slots<S, G, GL, Q + 1>(ra, s, v, va, vb, lrow);
  }
}

// end of kernel: warp-reduce every (slot, accumulator) and let lane 0 store it at out[g * NACC + J]
template <class S, int G, int J>
__device__ __forceinline__ void rs_flush(const KDesc &d, const RegAcc<S, G> &ra, i64 *out, int lane) {
  if constexpr (G > 0 && J < S::NACC) {
#pragma unroll
    for (int g = 0; g < G; g++) {
      i64 v;
      if constexpr (S::ACC_RK[J] == RK_WIDE || S::ACC_RK[J] == RK_MADW) v = ra.w[g][J];
      else if constexpr (S::ACC_RK[J] == RK_N32) v = (i64)ra.n[g][J];
      else {
        // CTA-local row index (iteration * tile_rows + row in tile) -> global row id; monotonic within the CTA
        const int idx = ra.n[g][J];
        v = INT64_MAX;
        if (idx != INT32_MAX) {
          const i64 tile = (i64)blockIdx.x + (i64)(idx / d.tile_rows) * gridDim.x;
          v = d.row_base + tile * d.tile_rows + idx % d.tile_rows;
        }
      }
      constexpr int op = S::ACC_OP[J];
      v = warp_reduce(op, v);
      if (lane == 0) out[g * S::NACC + J] = v;
    }
    rs_flush<S, G, J + 1>(d, ra, out, lane);
  }
}

// Per-CTA group state: key -> compact slot (lane-private accumulator tables are indexed by slot).
struct GroupState {
  int32_t *slotmap;   // [domain]  -1 unseen, -2 being claimed, -3 overflow (stays on the global-atomic path), >=0 slot
  int32_t *slotkey;   // [gmax]
  int32_t *nslots;
  i64 *tbl;           // [gmax][nacc][NC]
};

// Phase 2 (the Gathers + elementwise map + Fold of the plan): fold one selected row into the lane-private tables.
// GL (register slots only): the slot tests cover slots 0 .. GL-1 (= G today; a slot >= GL takes the global-atomic path).
template <class S, int NC, int G, int GL>
__device__ __forceinline__ void fold_row(const KDesc &d, const unsigned char *tile, int r, i64 grow, const GroupState &g, int ctid,
                                         RegAcc<S, G> &ra, int lrow) {
  i64 key;
  if constexpr (S::kStatic && S::KEY32) key = (i64)(key_chain32<S, 0>(d, 0, tile, r) & (int32_t)d.key_mask);
  else key = key_chain<S, 0>(d, 0, tile, r, grow) & d.key_mask;
  if ((u64)key >= (u64)d.domain) {   // the planner proves key < domain (mask); never expected
    atomicAdd(d.errflag, 1);
    return;
  }
  const int s = ((volatile int32_t *)g.slotmap)[key];
  if constexpr (G > 0) {
    i64 v[RegAcc<S, G>::NA];
    int32_t va[RegAcc<S, G>::NA], vb[RegAcc<S, G>::NA];
    rs_values<S, 0>(d, v, va, vb, 1, tile, r, grow);
    rs_apply_slots<S, G, GL, 0>(ra, s, v, va, vb, lrow);     // s < 0 (key without a slot yet) matches none
  }
  if (s >= 0 && (G == 0 || s < GL)) {
    if constexpr (G == 0) acc_chain<S, 0, NC>(d, g.tbl + (size_t)s * d.nacc * NC + ctid, 1, tile, r, grow);
  } else {
    // first rows of a key in this CTA: fold straight into the global table and claim a slot for the rest
#pragma unroll 1
    for (int j = 0; j < d.nacc; j++) acc_global(d.acc[j].op, d.table + (size_t)j * d.domain + key, acc_value_slow(d, j, tile, r, grow));
    if (s == -1 && atomicCAS(&g.slotmap[key], -1, -2) == -1) {
      int ns = atomicAdd(g.nslots, 1);
      if (ns < d.gmax) {
        g.slotkey[ns] = (int32_t)key;
        __threadfence_block();
        atomicExch(&g.slotmap[key], ns);
      } else {
        atomicExch(&g.slotmap[key], -3);
      }
    }
  }
}

// Dense tile, register slots: the R rows of a thread as ONE branch-free block so that their shared-memory loads and
// dependent chains overlap.  A row that did not pass the selection, or whose key has no register slot yet, gets
// slot -1 (matches no predicate); the latter rows are then redone by fold_row, which owns the claim / global path.
template <class S, int NC, int R, int G, int GL>
__device__ __forceinline__ void fold_dense(const KDesc &d, const unsigned char *tile, unsigned pass, int ctid, i64 grow0, const GroupState &g,
                                           RegAcc<S, G> &ra, int lrow0) {
  if constexpr (S::kStatic && S::KEY32) {
    int sl[R];
    unsigned redo = 0;
#pragma unroll
    for (int k = 0; k < R; k++) {
      // the host checked 0 <= key_mask < domain for a KEY32 shape: the masked key needs no range test
      const int key = key_chain32<S, 0>(d, 0, tile, ctid + k * NC) & (int32_t)d.key_mask;
      const int s = ((volatile int32_t *)g.slotmap)[key];
      const bool on = (pass >> k) & 1, have = (unsigned)s < (unsigned)GL;
      sl[k] = (on && have) ? s : -1;
      if (on && !have) redo |= 1u << k;
    }
#pragma unroll
    for (int k = 0; k < R; k++) {
      const int r = ctid + k * NC;
      i64 v[RegAcc<S, G>::NA];
      int32_t va[RegAcc<S, G>::NA], vb[RegAcc<S, G>::NA];
      rs_values<S, 0>(d, v, va, vb, 1, tile, r, grow0 + r);
      rs_apply_slots<S, G, GL, 0>(ra, sl[k], v, va, vb, lrow0 + r);
    }
    if (redo) {
#pragma unroll 1
      for (int k = 0; k < R; k++)
        if ((redo >> k) & 1) fold_row<S, NC, G, GL>(d, tile, ctid + k * NC, grow0 + ctid + k * NC, g, ctid, ra, lrow0 + ctid + k * NC);
    }
  } else {
#pragma unroll
    for (int k = 0; k < R; k++)
      if ((pass >> k) & 1) fold_row<S, NC, G, GL>(d, tile, ctid + k * NC, grow0 + ctid + k * NC, g, ctid, ra, lrow0 + ctid + k * NC);
  }
}
template <class S, int NC, int R, int G, int GL>
__device__ __forceinline__ void fold_queue(const KDesc &d, const unsigned char *buf, const uint16_t *queue, int e0, int nsel, i64 grow0,
                                           const GroupState &g, int ctid, RegAcc<S, G> &ra, int lrow0) {
  for (int e = e0; e < nsel; e += NC) {
    const int r = queue[e];
    fold_row<S, NC, G, GL>(d, buf, r, grow0 + r, g, ctid, ra, lrow0 + r);
  }
}

// Phase 1 of a tile (the plan's FoldSelect, Vlite.hs:721-730, done in shared memory): thread ctid evaluates the
// predicates of rows ctid + k*NC, k < R, together (R independent shared-memory loads in flight) and the rows
// that pass are compacted CTA-wide into `queue` with one warp-aggregated shared atomic per warp and k.
template <class S, int NC, int R, int G>
__device__ __forceinline__ void select_rows(const KDesc &d, const unsigned char *tile, int nvalid, int ctid, int *qcount, uint16_t *queue,
                                            i64 grow0, const GroupState &g, RegAcc<S, G> &ra, int lrow0) {
  unsigned pass = 0;
#pragma unroll
  for (int k = 0; k < R; k++)
    if (ctid + k * NC < nvalid) pass |= 1u << k;
  pred_chain<S, 0, NC, R>(d, tile, ctid, pass);
  const int lane = ctid & 31;
  const unsigned any = __ballot_sync(0xffffffffu, pass != 0);
  if (!any) return;
  if (__popc(any) >= 24) {
    // dense selection (most lanes own a selected row): compaction would buy nothing, fold the rows where they are
    if constexpr (G > 0) {   // straight-line predicated code: let the R rows of a thread overlap
      fold_dense<S, NC, R, G, G>(d, tile, pass, ctid, grow0, g, ra, lrow0);
    } else {
#pragma unroll 1
      for (int k = 0; k < R; k++)
        if ((pass >> k) & 1) fold_row<S, NC, G, 0>(d, tile, ctid + k * NC, grow0 + ctid + k * NC, g, ctid, ra, lrow0 + ctid + k * NC);
    }
    return;
  }
  if (__popc(any) <= 4) {
    // few selected rows in this warp (selective predicates): the owning lanes append on their own
    if (pass) {
      int base = atomicAdd(qcount, __popc(pass));
#pragma unroll
      for (int k = 0; k < R; k++)
        if ((pass >> k) & 1) queue[base++] = (uint16_t)(ctid + k * NC);
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < R; k++) {
    const bool p = (pass >> k) & 1;
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(qcount, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (p) queue[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)(ctid + k * NC);
    }
  }
}

__device__ __forceinline__ void choose_keys(const KDesc &d, i64 k0, i64 kstride);
template <int NT>
__device__ __forceinline__ void finalize_block(const FinDesc &f, const i64 *parts, int nranks, int tid, int *warp_cnt, i64 *running);
template <int NT>
__device__ __forceinline__ const i64 *exchange_block(const XDesc &x, const i64 *table, int *errflag, int tid);

template <class S, int NC, int R, int G>
__global__ void __launch_bounds__(NC + 32, 1) fused_scan_fold_kernel(const __grid_constant__ KDesc d, const __grid_constant__ FinDesc fd, const __grid_constant__ XDesc xd) {
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [ring: stages * stage_bytes][full[stages]][empty[stages]][sel[stages]][qcount[stages]]
  //         [queue[stages][tile_rows]][nslots][slotkey[gmax]][slotmap[domain]][tables]
  constexpr int NW = NC / 32;                 // consumer warps
  unsigned char *ring = smem;
  uint64_t *full = (uint64_t *)(smem + (size_t)d.stages * d.stage_bytes);
  uint64_t *empty = full + d.stages;
  uint64_t *sel = empty + d.stages;
  int *qcount = (int *)(sel + d.stages);
  uint16_t *queue = (uint16_t *)(qcount + d.stages + (d.stages & 1));
  GroupState g;
  g.nslots = (int32_t *)(queue + (size_t)d.stages * (NC * R));
  g.slotkey = g.nslots + 2;
  g.slotmap = g.slotkey + d.gmax;
  size_t tbl_off = (size_t)((unsigned char *)(g.slotmap + d.domain) - smem);
  tbl_off = (tbl_off + 15) & ~(size_t)15;
  g.tbl = (i64 *)(smem + tbl_off);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool dense = d.domain <= d.gmax;     // every key has its own slot from the start

  if (tid == 0) {
    for (int s = 0; s < d.stages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
      mbar_init(&sel[s], NW);
      qcount[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *g.nslots = dense ? (int)d.domain : 0;
  }
  for (i64 k = tid; k < d.domain; k += NC + 32) g.slotmap[k] = dense ? (int32_t)k : -1;
  if (dense && tid < d.gmax) g.slotkey[tid] = tid;
  if (G == 0 && tid >= 32) {
    const int ctid = tid - 32;
    for (int s = 0; s < d.gmax; s++)
      for (int j = 0; j < d.nacc; j++) g.tbl[((size_t)s * d.nacc + j) * NC + ctid] = acc_identity(d.acc[j].op);
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const uint64_t policy = policy_evict_first();
      int st = 0;
      uint32_t ph = 0;
      for (i64 tile = blockIdx.x; tile < d.ntiles; tile += gridDim.x) {
        mbar_wait(&empty[st], ph ^ 1);       // every consumer warp folded the stage's previous tile
        qcount[st] = 0;                      // (published to the consumers by the arrive below)
        mbar_expect_tx(&full[st], (uint32_t)d.stage_tx);
        unsigned char *dst = ring + (size_t)st * d.stage_bytes;
#pragma unroll 1
        for (int c = 0; c < d.ncols; c++) {
          uint32_t bytes = (uint32_t)(d.tile_rows * d.width[c]);
          bulk_g2s(dst + d.soff[c], (const char *)d.col[c] + (size_t)tile * bytes, bytes, &full[st], policy);
        }
        if (++st == d.stages) { st = 0; ph ^= 1; }
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  // Software-pipelined by one tile and free of CTA-wide barriers in steady state: iteration `it` SELECTS tile it
  // (predicates -> the stage's compaction queue, then a non-blocking arrive on sel[stage]) and FOLDS tile it-1
  // (wait on its sel barrier -- normally long complete --, fold this warp's share of the queue, release the stage).
  const int ctid = tid - 32, cw = warp - 1;
  RegAcc<S, G> ra;
  rs_init<S, G, 0>(ra);
  // The rows past the last full tile form one more (partial) tile, owned by the CTA next in the round-robin; it
  // is staged with plain loads into the (by then idle) next ring stage and goes through the same code.
  const i64 ntiles_all = d.ntiles + (d.ntiles * d.tile_rows < d.rows ? 1 : 0);
  int st = 0, pst = 0, it = 0;
  uint32_t ph = 0, pph = 0;
  i64 ptile = -1;
  for (i64 tile = blockIdx.x;; tile += gridDim.x, it++) {
    const bool have = tile < ntiles_all;
    if (have) {
      unsigned char *buf = ring + (size_t)st * d.stage_bytes;
      int nvalid = d.tile_rows;
      if (tile < d.ntiles) {
        mbar_wait(&full[st], ph);
      } else {
        const i64 tail0 = d.ntiles * d.tile_rows;
        nvalid = (int)(d.rows - tail0);
        consumer_barrier<NC>();      // every warp is done with this stage's previous tile (folded >= 1 iteration ago)
        if (ctid == 0) qcount[st] = 0;
#pragma unroll 1
        for (int c = 0; c < d.ncols; c++) {
          if (d.width[c] == 4) {
            const int32_t *src = (const int32_t *)d.col[c] + tail0;
            for (int r = ctid; r < nvalid; r += NC) ((int32_t *)(buf + d.soff[c]))[r] = src[r];
          } else {
            const i64 *src = (const i64 *)d.col[c] + tail0;
            for (int r = ctid; r < nvalid; r += NC) ((i64 *)(buf + d.soff[c]))[r] = src[r];
          }
        }
        consumer_barrier<NC>();
      }
      select_rows<S, NC, R, G>(d, buf, nvalid, ctid, &qcount[st], queue + (size_t)st * (NC * R), d.row_base + tile * d.tile_rows, g, ra,
                               it * d.tile_rows);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sel[st]);
    }
    if (ptile >= 0) {
      mbar_wait(&sel[pst], pph);
      const int nsel = ((volatile int *)qcount)[pst];
      const unsigned char *buf = ring + (size_t)pst * d.stage_bytes;
      const i64 grow0 = d.row_base + ptile * d.tile_rows;
      // queue entries in chunks of 32, dealt to the warps starting at a warp that rotates with the tile
      int chunk = cw - ((it - 1) % NW);
      if (chunk < 0) chunk += NW;
      const uint16_t *q = queue + (size_t)pst * (NC * R);
      const int e0 = chunk * 32 + lane, lrow0 = (it - 1) * d.tile_rows;
      fold_queue<S, NC, R, G, G>(d, buf, q, e0, nsel, grow0, g, ctid, ra, lrow0);
      __syncwarp();
      if (lane == 0 && ptile < d.ntiles) mbar_arrive(&empty[pst]);
    }
    if (!have) break;
    ptile = tile; pst = st; pph = ph;
    if (++st == d.stages) { st = 0; ph ^= 1; }
  }

  consumer_barrier<NC>();
  int ns = *((volatile int32_t *)g.nslots);
  if (ns > d.gmax) ns = d.gmax;
  if constexpr (G > 0) {
    // register slots -> warp shuffles -> one row per warp in the (now idle) ring -> one global atomic per (CTA, slot, accumulator)
    i64 *red = (i64 *)ring;                   // [NW][G * NACC]
    rs_flush<S, G, 0>(d, ra, red + (size_t)cw * (G * S::NACC), lane);
    consumer_barrier<NC>();
    for (int p = ctid; p < ns * S::NACC; p += NC) {
      const int j = p % S::NACC, op = d.acc[j].op;
      i64 v = acc_identity(op);
      for (int w = 0; w < NW; w++) v = acc_combine(op, v, red[(size_t)w * (G * S::NACC) + p]);
      if (v != acc_identity(op)) acc_global(op, d.table + (size_t)j * d.domain + g.slotkey[p / S::NACC], v);
    }
  } else {
    // lane-private tables -> one global atomic per (warp, slot, accumulator)
    for (int p = cw; p < ns * d.nacc; p += NW) {
      const int s = p / d.nacc, j = p % d.nacc, op = d.acc[j].op;
      i64 v = acc_identity(op);
      for (int t = lane; t < NC; t += 32) v = acc_combine(op, v, g.tbl[((size_t)s * d.nacc + j) * NC + t]);
      v = warp_reduce(op, v);
      if (lane == 0) acc_global(op, d.table + (size_t)j * d.domain + g.slotkey[s], v);
    }
  }

  // ---- epilogue by the last CTA to get here: the table is complete, finish the step without another launch
  if (d.epilogue) {
    __threadfence();                       // this CTA's atomics are ordered before its ticket
    consumer_barrier<NC>();
    if (ctid == 0) g.nslots[1] = atomicAdd(d.done, 1u) == gridDim.x - 1;
    consumer_barrier<NC>();
    if (g.nslots[1]) {
      __threadfence();
      if (d.nchoose) choose_keys(d, ctid, NC);
      if (d.epilogue >= 2) {
        __threadfence();
        consumer_barrier<NC>();
        const i64 *parts = fd.parts;
        int nranks = 1;
        if (d.epilogue == 3) {             // combine across GPUs through peer memory first
          parts = exchange_block<NC>(xd, d.table, d.errflag, ctid);
          nranks = xd.world;
        }
        // scratch in the idle ring: [running][warp counts]
        finalize_block<NC>(fd, parts, nranks, ctid, (int *)(ring + 16), (i64 *)ring);
      }
      if (ctid == 0) *d.done = 0;
    }
  }
}

// identity-initialise the global table ([nacc][domain] by op, choose section zero)
__global__ void fused_init_kernel(const __grid_constant__ KDesc d) {
  i64 n = (i64)(d.nacc + d.nchoose) * d.domain;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    int j = (int)(i / d.domain);
    d.table[i] = j < d.nacc ? acc_identity(d.acc[j].op) : 0;
  }
}

// FoldChoose = first row of the run (App. G6) = the expression at the smallest selected row of the key.
// Reads the table with cache-global loads: other CTAs built it with atomics at L2.
__device__ __forceinline__ void choose_keys(const KDesc &d, i64 k0, i64 kstride) {
  for (i64 k = k0; k < d.domain; k += kstride) {
    if (__ldcg(&d.table[(size_t)d.cnt_idx * d.domain + k]) <= 0) continue;
    i64 grow = __ldcg(&d.table[(size_t)d.first_idx * d.domain + k]);
    i64 r = grow - d.row_base;
    for (int c = 0; c < d.nchoose; c++) {
      const KAcc &A = d.choose[c];
      i64 v = 1;
      for (int f = 0; f < A.nfac; f++) {
        const KAffine &F = A.fac[f];
        i64 x;
        if (F.col == -1) x = F.a;
        else {
          i64 leaf = F.col == -2 ? grow : ((d.width[F.col] == 4 ? (i64)((const int32_t *)d.col[F.col])[r] : ((const i64 *)d.col[F.col])[r]) >> F.shr);
          x = (i64)((u64)F.a + (u64)F.b * (u64)leaf);
        }
        v = (i64)((u64)v * (u64)x);
      }
      d.table[(size_t)(d.nacc + c) * d.domain + k] = v;
    }
  }
}
__global__ void fused_choose_kernel(const __grid_constant__ KDesc d) {
  choose_keys(d, (i64)blockIdx.x * blockDim.x + threadIdx.x, (i64)gridDim.x * blockDim.x);
}

// Merge the per-rank tables, drop empty keys, emit one dense vector per fold in ascending key order, run the post
// ops, optionally re-initialise this rank's table.  One thread block of NT threads (threads tid 0..NT-1; the
// barrier is the named barrier 2 so that the scan kernel's consumer warps can run it without the producer warp).
// Store this rank's table into every rank's exchange buffer, publish the epoch, wait for every rank's epoch.
// Returns the [world][stride] block of this rank's own buffer that now holds all tables of the step.
template <int NT>
__device__ __forceinline__ const i64 *exchange_block(const XDesc &x, const i64 *table, int *errflag, int tid) {
  auto bar = []() { asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); };
  const int par = (int)(x.epoch & 1);
  const size_t slot = ((size_t)par * x.world + x.rank) * x.stride, flags = (size_t)2 * x.world * x.stride;
  for (int p = 0; p < x.world; p++) {                // NVLink stores (plain st.global to the peer mapping)
    i64 *dst = x.peer[p] + slot;
    for (i64 i = tid; i < x.stride; i += NT) dst[i] = __ldcg(&table[i]);
  }
  __threadfence_system();
  bar();
  if (tid < x.world) {
    st_release_sys((u64 *)(x.peer[tid] + flags) + (size_t)par * x.world + x.rank, x.epoch);
    const u64 *mine = (const u64 *)(x.peer[x.rank] + flags) + (size_t)par * x.world + tid;
    const u64 t0 = global_timer_ns();
    while (ld_acquire_sys(mine) < x.epoch) {
      if (global_timer_ns() - t0 > x.timeout_ns) { atomicAdd(errflag, 1 << 20); break; }   // a peer never arrived: fail, do not hang
      __nanosleep(64);
    }
  }
  bar();
  return x.peer[x.rank] + (size_t)par * x.world * x.stride;
}

template <int NT>
__device__ __forceinline__ void finalize_block(const FinDesc &f, const i64 *parts, int nranks, int tid, int *warp_cnt, i64 *running) {
  constexpr int NWF = NT / 32;
  auto bar = []() { asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); };
  const int lane = tid & 31, warp = tid >> 5;
  // out[] / post_out[] / ngroups live in one device buffer that starts at out[0]; hmirror has the same layout
  auto mirror = [&](i64 *dev) { return f.hmirror + (dev - f.out[0]); };
  if (tid == 0) *running = 0;
  bar();
  for (i64 base = 0; base < f.domain; base += NT) {
    i64 k = base + tid;
    i64 cnt = 0;
    if (k < f.domain)
      for (int r = 0; r < nranks; r++) cnt += __ldcg(&parts[(size_t)r * f.part_stride + (size_t)f.cnt_idx * f.domain + k]);
    bool exists = cnt > 0;
    unsigned m = __ballot_sync(0xffffffffu, exists);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    bar();
    int before = 0, total = 0;
    for (int w = 0; w < NWF; w++) {
      if (w < warp) before += warp_cnt[w];
      total += warp_cnt[w];
    }
    if (exists) {
      i64 pos = *running + before + __popc(m & ((1u << lane) - 1));
      int best = 0;   // rank holding the first row of this key
      if (f.nchoose > 0) {
        i64 bf = INT64_MAX;
        for (int r = 0; r < nranks; r++) {
          i64 fr = __ldcg(&parts[(size_t)r * f.part_stride + (size_t)f.first_idx * f.domain + k]);
          if (fr < bf) { bf = fr; best = r; }
        }
      }
      i64 ov[VDL_MAX_AGGS], pv[VDL_MAX_POSTS];
      for (int o = 0; o < f.nout; o++) {
        i64 v;
        if (f.out_kind[o] == 1) {
          v = __ldcg(&parts[(size_t)best * f.part_stride + (size_t)(f.nacc + f.out_idx[o]) * f.domain + k]);
        } else {
          int j = f.out_idx[o], op = f.acc_op[j];
          v = acc_identity(op);
          for (int r = 0; r < nranks; r++) v = acc_combine(op, v, __ldcg(&parts[(size_t)r * f.part_stride + (size_t)j * f.domain + k]));
        }
        f.out[o][pos] = v;
        if (f.hmirror) mirror(f.out[o])[pos] = v;
        ov[o] = v;
      }
      // elementwise epilogue over the fold results (AVG's Divide, ...)
      for (int q = 0; q < f.npost; q++) {
        const vdl_post_op &P = f.post[q];
        i64 a = P.a_kind == VDL_POST_CONST ? P.a : (P.a_kind == VDL_POST_FOLD ? ov[P.a] : pv[P.a]);
        i64 b = P.b_kind == VDL_POST_CONST ? P.b : (P.b_kind == VDL_POST_FOLD ? ov[P.b] : pv[P.b]);
        pv[q] = binop_apply(P.op, a, b);
        f.post_out[q][pos] = pv[q];
        if (f.hmirror) mirror(f.post_out[q])[pos] = pv[q];
      }
    }
    bar();
    if (tid == 0) *running += total;
    bar();
  }
  if (tid == 0) {
    const i64 ng = *running, err = *f.errflag;
    f.ngroups[0] = ng; f.ngroups[1] = err;
    if (f.hmirror) { mirror(f.ngroups)[0] = ng; mirror(f.ngroups)[1] = err; }
  }
  if (f.reset_table) {      // every thread has read what it needed (barriers above): identity-initialise for the next launch
    const i64 n = (i64)(f.nacc + f.nchoose) * f.domain;
    for (i64 i = tid; i < n; i += NT) {
      int j = (int)(i / f.domain);
      f.reset_table[i] = j < f.nacc ? acc_identity(f.acc_op[j]) : 0;
    }
  }
}

__global__ void __launch_bounds__(256, 1) fused_finalize_kernel(const __grid_constant__ FinDesc f) {
  __shared__ int warp_cnt[8];
  __shared__ i64 running;
  finalize_block<256>(f, f.parts, f.nranks, threadIdx.x, warp_cnt, &running);
}

// A rank with nothing to scan still takes part in the exchange: its (identity) table, then the common finalize.
__global__ void __launch_bounds__(256, 1) fused_exchange_kernel(const __grid_constant__ KDesc d, const __grid_constant__ FinDesc f, const __grid_constant__ XDesc x) {
  __shared__ int warp_cnt[8];
  __shared__ i64 running;
  const i64 *parts = exchange_block<256>(x, d.table, d.errflag, threadIdx.x);
  finalize_block<256>(f, parts, x.world, threadIdx.x, warp_cnt, &running);
}

// ------------------------------------------------------------------------------ host side
typedef void (*scan_kernel_fn)(const KDesc, const FinDesc, const XDesc);
struct vdl_fused {
  vdl_ctx *ctx = nullptr;
  KDesc kd;
  FinDesc fd;
  vdl_vec table = 0;
  vdl_vec out[VDL_MAX_AGGS] = {0};
  int nout = 0;
  i64 *d_outbuf = nullptr;       // [nout][domain] fold results, then [ngroups, errflag]: fetched with ONE copy
  i64 *h_outbuf = nullptr;       // pinned mirror
  i64 ngroups = -1;
  bool finalized = false, always_false = false, rs = false;
  bool table_clean = false;       // the partial table holds the fold identities (init kernel or a finalize that reset it)
  unsigned int *d_done = nullptr; // ticket counter of the scan kernel's last-CTA epilogue
  i64 *h_mapped = nullptr;        // device address of h_outbuf (mapped pinned memory)
  XDesc xd;                       // peer exchange (world == 0: not configured)
  // register slots: kernels by slot count; the launch picks the smallest count that covers the groups the previous
  // run of this scan produced (more slots = more predicated work per row; too few = keys on the slow global path)
  scan_kernel_fn rs_kernel[9] = {nullptr};
  int rs_gmax = 0;
  i64 groups_seen = -1;
  size_t smem_bytes = 0;
  int grid = 1, nc = 256, r = 4;
  scan_kernel_fn kernel = nullptr;
  const char *shape = "generic";
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // identity of the columns at prepare time (the proofs above hold for exactly this data)
  int ncols = 0;
  vdl_vec col_handle[VDL_MAX_COLS] = {0};
  u64 col_gen[VDL_MAX_COLS] = {0};
};

bool vdl_fused_current(vdl_fused *f) {
  for (int c = 0; c < f->ncols; c++) {
    u64 g;
    if (!vec_identity(f->ctx, f->col_handle[c], &g) || g != f->col_gen[c]) return false;
  }
  return true;
}

#include "vdl_shapes.cuh"

template <class S>
static scan_kernel_fn scan_kernel_for(int nc, int r) {
  if (nc == 512) return r == 4 ? fused_scan_fold_kernel<S, 512, 4, 0> : (r == 2 ? fused_scan_fold_kernel<S, 512, 2, 0> : fused_scan_fold_kernel<S, 512, 1, 0>);
  return r == 4 ? fused_scan_fold_kernel<S, 256, 4, 0> : (r == 2 ? fused_scan_fold_kernel<S, 256, 2, 0> : fused_scan_fold_kernel<S, 256, 1, 0>);
}
// register-slot instantiations (geometries: rs_geometries) for g of the shape's RS_G slots in registers
template <class S, int G>
static scan_kernel_fn rs_kernel_for_g(int nc, int r) {
  // consumer warps + the producer warp are dealt round-robin to the 4 SM sub-partitions (16 K registers each):
  // 11 + 1 warps -> 3 per sub-partition -> 168 registers per thread; 15 + 1 -> 4 -> 128; 7 + 1 -> 2 -> 255
  if (nc == 352 && r == 4) return fused_scan_fold_kernel<S, 352, 4, G>;
  if (nc == 352 && r == 2) return fused_scan_fold_kernel<S, 352, 2, G>;
  if (nc == 480 && r == 2) return fused_scan_fold_kernel<S, 480, 2, G>;
  return nullptr;
}
template <class S>
static scan_kernel_fn rs_kernel_for(int nc, int r, int g) {
  if constexpr (S::RS_G > 0) {
    static_assert(S::RS_G == 8, "instantiate the slot counts this shape allows");
    if (g == 8) return rs_kernel_for_g<S, 8>(nc, r);
    if (g == 6) return rs_kernel_for_g<S, 6>(nc, r);
    if (g == 4) return rs_kernel_for_g<S, 4>(nc, r);
  }
  return nullptr;
}
static const int rs_slot_counts[] = {4, 6, 8};
static const int rs_geometries[][2] = {{352, 4}, {352, 2}, {480, 2}};

// does the prepared descriptor satisfy every assumption static shape S compiles in?
static bool flags_ok(int fl, const KAffine &a) {
  if (((fl & FF_CONST) != 0) != (a.col == -1) || ((fl & FF_ROWID) != 0) != (a.col == -2)) return false;
  if (a.col >= 0 && ((fl & FF_W4) != 0) != (a.w4 != 0)) return false;
  if ((fl & FF_SHR0) && a.shr != 0) return false;
  if ((fl & FF_B1) && a.b != 1) return false;
  if ((fl & FF_BM1) && a.b != -1) return false;
  if ((fl & FF_A0) && a.a != 0) return false;
  if ((fl & FF_NARROW) && !a.narrow) return false;
  return true;
}
template <class S>
static bool shape_matches(const KDesc &k) {
  if (k.npreds != S::NPREDS || k.nkeys != S::NKEYS || k.nacc != S::NACC) return false;
  for (int i = 0; i < k.npreds; i++)
    if (k.pred[i].w4 != S::PRED_MODE[i] || (S::PRED_SHR0[i] && k.pred[i].shr != 0)) return false;
  for (int i = 0; i < k.nkeys; i++) {
    if (!flags_ok(S::KEY_FLAGS[i], k.key[i].e) || (S::KEY_SHL0[i] && k.key[i].shl != 0)) return false;
    if (S::KEY32 && !k.key[i].e.narrow) return false;
  }
  if (S::KEY32 && (k.key_mask < 0 || k.key_mask > INT32_MAX || k.key_mask >= k.domain)) return false;
  for (int j = 0; j < k.nacc; j++) {
    if (k.acc[j].op != S::ACC_OP[j] || k.acc[j].chain != S::ACC_CHAIN[j] || k.acc[j].nfac != S::ACC_NFAC[j]) return false;
    for (int t = 0; t < k.acc[j].nfac; t++)
      if (!flags_ok(S::FAC[j][t], k.acc[j].fac[t])) return false;
  }
  return true;
}

static bool affine_ok(const vdl_affine &a, int ncols) { return a.column >= -2 && a.column < ncols && a.shr >= 0 && a.shr < 64; }
static KAffine to_k(const vdl_affine &a) { return KAffine{a.column, a.shr, a.a, a.b, 0, 0, 0, 0}; }

extern "C" int vdl_fused_prepare(vdl_ctx *ctx, const vdl_fused_desc *desc, vdl_fused **out) {
  if (!ctx || !desc || !out) return VDL_EINVAL;
  *out = nullptr;
  if (desc->ncolumns < 1 || desc->ncolumns > VDL_MAX_COLS) return vdl_fail(ctx, VDL_EINVAL, "fused scan: %d columns (1..%d)", desc->ncolumns, VDL_MAX_COLS);
  if (desc->npreds < 0 || desc->npreds > VDL_MAX_PREDS || desc->nkeys < 0 || desc->nkeys > VDL_MAX_KEYS || desc->nfolds < 1 || desc->nfolds > VDL_MAX_AGGS)
    return vdl_fail(ctx, VDL_EINVAL, "fused scan: preds/keys/folds out of range");
  if (desc->domain < 1 || desc->domain > (1 << 20)) return vdl_fail(ctx, VDL_EUNSUPPORTED, "fused scan: key domain %lld not in [1, 2^20]", (long long)desc->domain);
  if (desc->nkeys == 0 && desc->domain != 1) return vdl_fail(ctx, VDL_EINVAL, "fused scan: no key parts but domain %lld", (long long)desc->domain);
  if (desc->nposts < 0 || desc->nposts > VDL_MAX_POSTS) return vdl_fail(ctx, VDL_EINVAL, "fused scan: %d post ops (0..%d)", desc->nposts, VDL_MAX_POSTS);
  for (int q = 0; q < desc->nposts; q++) {
    const vdl_post_op &P = desc->post[q];
    auto ok = [&](int kind, i64 v) { return kind == VDL_POST_CONST || (kind == VDL_POST_FOLD && v >= 0 && v < desc->nfolds) || (kind == VDL_POST_POST && v >= 0 && v < q); };
    if (P.op < 0 || P.op > VDL_MODULO || !ok(P.a_kind, P.a) || !ok(P.b_kind, P.b)) return vdl_fail(ctx, VDL_EINVAL, "fused scan: bad post op %d", q);
  }
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));

  vdl_fused *f = new vdl_fused();
  f->ctx = ctx;
  KDesc &k = f->kd;
  memset(&k, 0, sizeof k);
  memset(&f->fd, 0, sizeof f->fd);
  memset(&f->xd, 0, sizeof f->xd);
  k.rows = desc->rows;
  k.row_base = desc->row_base;
  k.key_mask = desc->key_mask;
  k.domain = desc->domain;
  k.ncols = desc->ncolumns;
  k.errflag = ctx->d_errflag;
  int rowbytes = 0;
  for (int c = 0; c < desc->ncolumns; c++) {
    Vec *v = vec_get(ctx, desc->column[c]);
    if (!v || v->is_range) { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: column %d is not a stored column", c); }
    if (v->len < desc->rows) { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: column %s has %lld rows < %lld", v->name.c_str(), (long long)v->len, (long long)desc->rows); }
    k.col[c] = v->ptr;
    k.width[c] = v->dtype;
    rowbytes += v->dtype;
    f->col_handle[c] = desc->column[c];
    f->col_gen[c] = v->gen;
  }
  f->ncols = desc->ncolumns;
  for (int i = 0; i < desc->npreds; i++) {
    const vdl_range_pred &p = desc->pred[i];
    if (p.column < 0 || p.column >= desc->ncolumns || p.shr < 0 || p.shr > 63) { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: bad predicate %d", i); }
    if (p.lo > p.hi) f->always_false = true;
    k.pred[k.npreds++] = KPred{p.column, p.shr, p.lo, (u64)p.hi - (u64)p.lo, 0, 0, 0, 0};
  }
  for (int i = 0; i < desc->nkeys; i++) {
    if (!affine_ok(desc->key[i].e, desc->ncolumns) || desc->key[i].shl < 0 || desc->key[i].shl > 63) { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: bad key part %d", i); }
    k.key[k.nkeys++] = KKey{to_k(desc->key[i].e), desc->key[i].shl, 0};
  }
  // folds -> scan-time accumulators (+ count, + first row when a FoldChoose is present)
  f->nout = desc->nfolds;
  for (int i = 0; i < desc->nfolds; i++) {
    const vdl_fold_spec &s = desc->fold[i];
    if (s.nfactors < 0 || s.nfactors > VDL_MAX_FACTORS) { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: fold %d has %d factors", i, s.nfactors); }
    for (int t = 0; t < s.nfactors; t++)
      if (!affine_ok(s.factor[t], desc->ncolumns)) { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: fold %d factor %d invalid", i, t); }
    KAcc a;
    memset(&a, 0, sizeof a);
    a.nfac = s.nfactors;
    for (int t = 0; t < s.nfactors; t++) a.fac[t] = to_k(s.factor[t]);
    if (s.op == VDL_FOLD_SUM || s.op == VDL_FOLD_MIN || s.op == VDL_FOLD_MAX) {
      a.op = s.op == VDL_FOLD_SUM ? 0 : (s.op == VDL_FOLD_MIN ? 1 : 2);
      f->fd.out_kind[i] = 0;
      f->fd.out_idx[i] = k.nacc;
      k.acc[k.nacc++] = a;
    } else if (s.op == VDL_FOLD_CHOOSE) {
      if (k.nchoose == K_MAX_CHOOSE) { delete f; return vdl_fail(ctx, VDL_EUNSUPPORTED, "fused scan: more than %d FoldChoose", K_MAX_CHOOSE); }
      f->fd.out_kind[i] = 1;
      f->fd.out_idx[i] = k.nchoose;
      k.choose[k.nchoose++] = a;
    } else if (s.op == VDL_FOLD_COUNT) {
      f->fd.out_kind[i] = 0;
      f->fd.out_idx[i] = -1;   // patched to cnt_idx below
    } else { delete f; return vdl_fail(ctx, VDL_EINVAL, "fused scan: fold op %d", s.op); }
  }
  KAcc cnt;
  memset(&cnt, 0, sizeof cnt);
  k.cnt_idx = k.nacc;
  k.acc[k.nacc++] = cnt;                  // SUM of the empty product = row count per key
  k.first_idx = -1;
  if (k.nchoose) {
    KAcc first;
    memset(&first, 0, sizeof first);
    first.op = 1;
    first.nfac = 1;
    first.fac[0] = KAffine{-2, 0, 0, 1, 0, 0, 0, 0};  // MIN over the global row id
    k.first_idx = k.nacc;
    k.acc[k.nacc++] = first;
  }
  for (int i = 0; i < desc->nfolds; i++)
    if (f->fd.out_kind[i] == 0 && f->fd.out_idx[i] < 0) f->fd.out_idx[i] = k.cnt_idx;

  // geometry: consumer threads NC, rows per thread and tile R, ring depth, lane-private tables (or none: register slots)
  const int smem_max = ctx->smem_optin > 0 ? ctx->smem_optin : 232448;
  if ((size_t)desc->domain * 4 > 64 * 1024) { delete f; return vdl_fail(ctx, VDL_EUNSUPPORTED, "fused scan: key domain %lld too large for the shared-memory slot map", (long long)desc->domain); }
  auto set_geometry = [&](int nc, int r, int gmax, bool tables, int min_stages) -> bool {
    int tile_rows = nc * r, off = 0;
    size_t fixed = (tables ? (size_t)gmax * k.nacc * nc * 8 : 0) + (size_t)desc->domain * 4 + (size_t)gmax * 4 + 16 + 256;
    size_t per_stage = 3 * 8 + 4 + 4 + 2 * (size_t)tile_rows;   // barriers, queue count, queue
    for (int c = 0; c < k.ncols; c++) { k.soff[c] = off; off += ((tile_rows * k.width[c] + 127) / 128) * 128; }
    int stages = (int)(((long)smem_max - (long)fixed) / (long)(off + per_stage));
    if (stages < min_stages) return false;
    k.tile_rows = tile_rows; k.stage_bytes = off; k.stage_tx = tile_rows * rowbytes; k.stages = std::min(stages, 12);
    k.gmax = gmax; f->nc = nc; f->r = r;
    f->smem_bytes = (size_t)k.stages * (k.stage_bytes + per_stage) + fixed;
    k.ntiles = desc->rows / k.tile_rows;
    f->grid = (int)std::max<i64>(1, std::min<i64>(ctx->sm_count, k.ntiles));
    return true;
  };
  auto default_geometry = [&]() -> bool {
    for (int nc = 512; nc >= 256; nc /= 2) {
      // slots with lane-private tables: as many as the domain needs, up to what ~100 KB holds
      int gmax = 1;
      while (gmax < desc->domain && gmax < 64 && (size_t)(gmax * 2) * k.nacc * nc * 8 <= 120 * 1024) gmax *= 2;
      if (nc == 512 && gmax < desc->domain && gmax < 8) continue;     // too few slots: halve the consumers instead
      for (int r = 2048 / nc; r >= 1; r /= 2) {
        if (r > 4) continue;
        if (set_geometry(nc, r, gmax, true, (r == 1 && nc == 256) ? 2 : 3)) return true;
      }
    }
    return false;
  };
  if (!default_geometry()) { delete f; return vdl_fail(ctx, VDL_EUNSUPPORTED, "fused scan: %d columns x %d accumulators do not fit shared memory", k.ncols, k.nacc); }

  // prefix sharing between consecutive accumulators (ep, ep*(100-d), ep*(100-d)*(100+t) evaluate each factor once)
  for (int j = 1; j < k.nacc; j++) {
    KAcc &a = k.acc[j];
    if (j == k.cnt_idx || j == k.first_idx || j - 1 == k.cnt_idx) continue;
    // full factor list of the previous accumulator (its own chain expanded) a prefix of this one?
    std::vector<KAffine> prev;
    int q = j - 1;
    std::vector<int> chain_members;
    while (true) { chain_members.push_back(q); if (!k.acc[q].chain) break; q--; }
    for (int m = (int)chain_members.size() - 1; m >= 0; m--)
      for (int t = 0; t < k.acc[chain_members[m]].nfac; t++) prev.push_back(k.acc[chain_members[m]].fac[t]);
    bool prefix = !prev.empty() && (int)prev.size() < a.nfac;
    for (size_t t = 0; prefix && t < prev.size(); t++)
      prefix = prev[t].col == a.fac[t].col && prev[t].shr == a.fac[t].shr && prev[t].a == a.fac[t].a && prev[t].b == a.fac[t].b;
    if (prefix) {
      int np = (int)prev.size();
      for (int t = np; t < a.nfac; t++) a.fac[t - np] = a.fac[t];
      a.nfac -= np;
      a.chain = 1;
    }
  }

  // derived descriptor fields (depend on the geometry): staged offsets / width flags, 32-bit bounds for 4-byte
  // predicate columns, value bounds of every affine term from the column statistics
  struct Bound { __int128 lo, hi; };
  std::map<const KAffine *, Bound> bound;
  int place_rc = VDL_OK;
  auto place = [&](KAffine &a) {
    if (a.col == -1) { bound[&a] = Bound{a.a, a.a}; return; }
    if (a.col == -2) {
      __int128 v0 = (__int128)a.a + (__int128)a.b * k.row_base, v1 = (__int128)a.a + (__int128)a.b * (k.row_base + k.rows);
      bound[&a] = Bound{std::min(v0, v1), std::max(v0, v1)};
      return;
    }
    a.soff = k.soff[a.col];
    a.w4 = k.width[a.col] == 4;
    // narrow: leaf and a + b*leaf provably fit int32 given the column's exact min/max (cf. inferBounds, Vlite.hs:417-467)
    i64 cmin, cmax;
    int rc2 = vdl_column_analyze(ctx, desc->column[a.col], &cmin, &cmax);
    if (rc2) { place_rc = rc2; return; }
    __int128 l0 = cmin >> a.shr, l1 = cmax >> a.shr;
    __int128 v0 = (__int128)a.a + (__int128)a.b * l0, v1 = (__int128)a.a + (__int128)a.b * l1;
    auto fits = [](__int128 x) { return x >= INT32_MIN && x <= INT32_MAX; };
    a.narrow = cmin <= cmax && fits(l0) && fits(l1) && fits(v0) && fits(v1) && fits(a.a) && fits(a.b);
    bound[&a] = Bound{std::min(v0, v1), std::max(v0, v1)};
  };
  auto derive = [&]() -> int {
    for (int i = 0; i < k.npreds; i++) {
      KPred &p = k.pred[i];
      p.soff = k.soff[p.col];
      p.w4 = k.width[p.col] == 4;
      if (!p.w4) {   // 8-byte column: when the column statistics say every (shifted) value fits int32, read low words only
        i64 cmin, cmax;
        int rc2 = vdl_column_analyze(ctx, desc->column[p.col], &cmin, &cmax);
        if (rc2) return rc2;
        if ((cmin >> p.shr) >= INT32_MIN && (cmax >> p.shr) <= INT32_MAX) p.w4 = 2;
      }
      if (p.w4) {   // values of a 4-byte column (shifted or not) lie in int32: clamp the bounds, compare in 32 bits
        i64 lo = std::max<i64>(p.lo, INT32_MIN), hi = std::min<i64>((i64)((u64)p.lo + p.span), INT32_MAX);
        if (lo > hi) f->always_false = true;
        p.lo32 = (int32_t)lo;
        p.span32 = (uint32_t)(hi - lo);
      }
    }
    for (int i = 0; i < k.nkeys; i++) place(k.key[i].e);
    for (int j = 0; j < k.nacc; j++)
      for (int t = 0; t < k.acc[j].nfac; t++) place(k.acc[j].fac[t]);
    return place_rc;
  };
  { int rc2 = derive(); if (rc2) { vdl_fused_destroy(f); return rc2; } }

  // largest |value| accumulator j can take on one row (chain expanded), saturating
  auto acc_maxabs = [&](int j) -> __int128 {
    __int128 m = 1;
    const __int128 cap = (__int128)1 << 100;
    for (int q = j;; q--) {
      for (int t = 0; t < k.acc[q].nfac; t++) {
        const Bound &b = bound[&k.acc[q].fac[t]];
        __int128 x = std::max(b.lo < 0 ? -b.lo : b.lo, b.hi < 0 ? -b.hi : b.hi);
        m = m * x;
        if (m > cap) m = cap;
      }
      if (!k.acc[q].chain) break;
    }
    return m;
  };
  // the proofs a register-slot instantiation of shape S needs under the CURRENT geometry (vdl_shapes.cuh)
  auto rs_proofs_hold = [&](const int *rk) -> bool {
    const i64 ntiles_all = k.ntiles + (k.ntiles * k.tile_rows < k.rows ? 1 : 0);
    const i64 tiles_per_cta = (ntiles_all + f->grid - 1) / f->grid;
    const __int128 rows_per_thread = (__int128)tiles_per_cta * 2 * f->r;   // own rows of a dense tile + a share of the queue
    for (int j = 0; j < k.nacc; j++) {
      if (rk[j] == RK_N32 && (k.acc[j].op != 0 || rows_per_thread * acc_maxabs(j) > INT32_MAX)) return false;
      if (rk[j] == RK_MADW) {          // value = a * b with a = everything but the last own factor, b = that factor
        if (k.acc[j].op != 0 || k.acc[j].nfac < 1) return false;
        const Bound &bb = bound[&k.acc[j].fac[k.acc[j].nfac - 1]];
        __int128 bmax = std::max(bb.lo < 0 ? -bb.lo : bb.lo, bb.hi < 0 ? -bb.hi : bb.hi);
        if (bmax > INT32_MAX || (bmax > 0 && acc_maxabs(j) / bmax > INT32_MAX) || acc_maxabs(j) >= ((__int128)1 << 100)) return false;
        for (int q = j;; q--) {        // every factor nonnegative: the multiply-add is unsigned
          for (int t = 0; t < k.acc[q].nfac; t++)
            if (bound[&k.acc[q].fac[t]].lo < 0) return false;
          if (!k.acc[q].chain) break;
        }
      }
      if (rk[j] == RK_FIRST && (k.acc[j].op != 1 || (__int128)tiles_per_cta * k.tile_rows >= INT32_MAX)) return false;
    }
    return true;
  };

  // pick the kernel: a static shape whose assumptions all hold, else the generic one; a shape with register slots
  // runs in that mode when the key domain's slots and the 32-bit proofs allow it
  f->kernel = scan_kernel_for<GenericShape>(f->nc, f->r);
  auto try_shape = [&](auto shape_tag) -> int {
    using S = decltype(shape_tag);
    if (!shape_matches<S>(k)) return 0;
    f->kernel = scan_kernel_for<S>(f->nc, f->r);
    f->shape = S::kName;
    if constexpr (S::RS_G > 0) {
      if (getenv("VDL_NO_REGISTER_SLOTS")) return 1;
      int want_nc = 0, want_r = 0;
      if (const char *e = getenv("VDL_RS_GEOMETRY")) sscanf(e, "%d,%d", &want_nc, &want_r);
      for (auto &geo : rs_geometries) {
        if (want_nc && (geo[0] != want_nc || geo[1] != want_r)) continue;
        if (!rs_kernel_for<S>(geo[0], geo[1], S::RS_G) || !set_geometry(geo[0], geo[1], S::RS_G, false, 3)) continue;
        int rc2 = derive();
        if (rc2) return -rc2;
        if (shape_matches<S>(k) && rs_proofs_hold(S::ACC_RK)) {
          for (int g : rs_slot_counts) f->rs_kernel[g] = rs_kernel_for<S>(geo[0], geo[1], g);
          f->kernel = f->rs_kernel[S::RS_G];
          f->rs = true;
          f->rs_gmax = S::RS_G;
          return 1;
        }
      }
      default_geometry();               // no register-slot geometry qualified: back to the shared-memory tables
      int rc2 = derive();
      if (rc2) return -rc2;
    }
    return 1;
  };
  if (!getenv("VDL_GENERIC_ONLY")) {
    int m = try_shape(ShapeSel3Sum2{});
    if (m == 0) m = try_shape(ShapeSel1Key2Sum5{});
    if (m < 0) { vdl_fused_destroy(f); return -m; }
  }
  if (getenv("VDL_DEBUG_SHAPE")) {
    fprintf(stderr, "[vdl] fused scan: shape=%s%s nc=%d r=%d stages=%d gmax=%d npreds=%d nkeys=%d nacc=%d smem=%zu\n", f->shape, f->rs ? " (register slots)" : "", f->nc, f->r, k.stages, k.gmax, k.npreds, k.nkeys, k.nacc, f->smem_bytes);
    for (int i = 0; i < k.npreds; i++) fprintf(stderr, "[vdl]   pred %d: mode=%d shr=%d\n", i, k.pred[i].w4, k.pred[i].shr);
    for (int i = 0; i < k.nkeys; i++) fprintf(stderr, "[vdl]   key %d: col=%d w4=%d shr=%d a=%lld b=%lld shl=%d narrow=%d\n", i, k.key[i].e.col, k.key[i].e.w4, k.key[i].e.shr, (long long)k.key[i].e.a, (long long)k.key[i].e.b, k.key[i].shl, k.key[i].e.narrow);
    for (int j = 0; j < k.nacc; j++) {
      fprintf(stderr, "[vdl]   acc %d: op=%d chain=%d nfac=%d", j, k.acc[j].op, k.acc[j].chain, k.acc[j].nfac);
      for (int t = 0; t < k.acc[j].nfac; t++) fprintf(stderr, " [col=%d w4=%d shr=%d a=%lld b=%lld narrow=%d]", k.acc[j].fac[t].col, k.acc[j].fac[t].w4, k.acc[j].fac[t].shr, (long long)k.acc[j].fac[t].a, (long long)k.acc[j].fac[t].b, k.acc[j].fac[t].narrow);
      fprintf(stderr, "\n");
    }
  }

  // device buffers
  int rc = vec_new(ctx, VDL_I64, (i64)(k.nacc + k.nchoose) * k.domain, &f->table);
  if (rc) { delete f; return rc; }
  k.table = (i64 *)ctx->vecs[f->table].ptr;
  {
    size_t nb = ((size_t)(f->nout + desc->nposts) * k.domain + 2) * sizeof(i64);
    if (cudaMalloc(&f->d_outbuf, nb) != cudaSuccess || cudaHostAlloc(&f->h_outbuf, nb, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&f->h_mapped, f->h_outbuf, 0) != cudaSuccess || cudaMalloc(&f->d_done, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemsetAsync(f->d_done, 0, sizeof(unsigned int), ctx->stream) != cudaSuccess) {
      vdl_fused_destroy(f);
      return vdl_fail(ctx, VDL_ENOMEM, "fused scan: result buffers");
    }
    k.done = f->d_done;
  }
  for (int i = 0; i < f->nout; i++) {   // the fold results are views into the one result buffer
    rc = vec_new_range(ctx, 0, 0, 0, &f->out[i]);
    if (rc) { vdl_fused_destroy(f); return rc; }
    Vec &v = ctx->vecs[f->out[i]];
    v.is_range = false;
    v.ptr = f->d_outbuf + (size_t)i * k.domain;
    v.dtype = VDL_I64;
    v.cap_rows = k.domain;
    v.domain = -1;
    f->fd.out[i] = (i64 *)v.ptr;
  }
  cudaEventCreate(&f->ev0);
  cudaEventCreate(&f->ev1);
  f->fd.domain = k.domain;
  f->fd.nacc = k.nacc;
  f->fd.nchoose = k.nchoose;
  f->fd.cnt_idx = k.cnt_idx;
  f->fd.first_idx = k.first_idx;
  f->fd.nout = f->nout;
  f->fd.part_stride = (i64)(k.nacc + k.nchoose) * k.domain;
  f->fd.npost = desc->nposts;
  for (int q = 0; q < desc->nposts; q++) {
    f->fd.post[q] = desc->post[q];
    f->fd.post_out[q] = f->d_outbuf + (size_t)(f->nout + q) * k.domain;
  }
  f->fd.ngroups = f->d_outbuf + (size_t)(f->nout + desc->nposts) * k.domain;
  f->fd.errflag = ctx->d_errflag;
  f->fd.hmirror = f->h_mapped;
  for (int j = 0; j < k.nacc; j++) f->fd.acc_op[j] = k.acc[j].op;

  cudaError_t e = cudaFuncSetAttribute(f->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
  if (e != cudaSuccess) { vdl_fused_destroy(f); return vdl_cuda_fail(ctx, e, "cudaFuncSetAttribute(fused_scan_fold_kernel)"); }
  for (int g : rs_slot_counts)
    if (f->rs && f->rs_kernel[g] && (e = cudaFuncSetAttribute(f->rs_kernel[g], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max)) != cudaSuccess) {
      vdl_fused_destroy(f);
      return vdl_cuda_fail(ctx, e, "cudaFuncSetAttribute(fused_scan_fold_kernel, register slots)");
    }
  *out = f;
  return VDL_OK;
}

// self_finalize != 0: single-GPU step -- the scan kernel's last CTA also runs FoldChoose, the finalize and the table
// reset, so a step is ONE launch and the results are on the host (mapped pinned memory) when the stream drains.
// self_finalize == 0: the scan leaves the complete partial table (FoldChoose values included) for an external
// combine across ranks; vdl_fused_finalize() then merges the gathered tables.
extern "C" int vdl_fused_launch_ex(vdl_fused *f, int self_finalize) {
  if (!f) return VDL_EINVAL;
  vdl_ctx *ctx = f->ctx;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!vdl_fused_current(f))
    return vdl_fail(ctx, VDL_ESTALE, "fused scan: a column was rewritten or dropped after vdl_fused_prepare (its statistics proofs no longer hold): prepare again");
  f->finalized = false;
  f->ngroups = -1;
  if (!f->table_clean) {
    int nb = (int)std::min<i64>(ctx->sm_count, ((i64)(f->kd.nacc + f->kd.nchoose) * f->kd.domain + 255) / 256);
    fused_init_kernel<<<std::max(nb, 1), 256, 0, ctx->stream>>>(f->kd);
    ctx->launches++;
    f->table_clean = true;
  }
  VDL_CUDA(ctx, cudaEventRecord(f->ev0, ctx->stream));
  if (f->rs) {
    int g = f->rs_gmax;
    if (f->groups_seen >= 0 && !getenv("VDL_RS_ALL_SLOTS"))
      for (int c : rs_slot_counts)
        if (c >= f->groups_seen && c < g && f->rs_kernel[c]) g = c;
    f->kernel = f->rs_kernel[g];
    f->kd.gmax = g;
  }
  const bool scan = f->kd.rows > 0 && !f->always_false;
  if (self_finalize == 2) {
    if (f->xd.world < 1) return vdl_fail(ctx, VDL_EINVAL, "fused scan: peer exchange requested before vdl_fused_set_peers");
    f->xd.epoch++;
  }
  f->fd.parts = f->kd.table;
  f->fd.nranks = 1;
  f->fd.reset_table = f->kd.table;
  if (scan) {
    f->kd.epilogue = self_finalize == 2 ? 3 : (self_finalize ? 2 : 1);
    f->kernel<<<f->grid, f->nc + 32, f->smem_bytes, ctx->stream>>>(f->kd, f->fd, f->xd);
    ctx->launches++;
    f->table_clean = self_finalize != 0;
  } else if (self_finalize == 2) {       // nothing to scan here, but the other ranks wait for this rank's table
    fused_exchange_kernel<<<1, 256, 0, ctx->stream>>>(f->kd, f->fd, f->xd);
    ctx->launches++;
  } else if (self_finalize) {            // nothing to scan: the (identity) table finalizes to zero groups
    fused_finalize_kernel<<<1, 256, 0, ctx->stream>>>(f->fd);
    ctx->launches++;
  }
  VDL_CUDA(ctx, cudaEventRecord(f->ev1, ctx->stream));
  f->timed = true;
  VDL_CUDA(ctx, cudaGetLastError());
  if (self_finalize) f->finalized = true;
  return VDL_OK;
}

// step counter of the peer exchange; the plan carries it over when a scan is re-prepared on the same buffers
u64 vdl_fused_epoch(vdl_fused *f) { return f->xd.epoch; }
void vdl_fused_set_epoch(vdl_fused *f, u64 e) { f->xd.epoch = e; }

extern "C" int vdl_fused_exchange_bytes(vdl_fused *f, int world, int64_t *bytes) {
  if (!f || !bytes || world < 1 || world > VDL_MAX_RANKS) return VDL_EINVAL;
  *bytes = ((int64_t)2 * world * f->fd.part_stride + 2 * world) * (int64_t)sizeof(i64);
  return VDL_OK;
}

extern "C" int vdl_fused_set_peers(vdl_fused *f, int rank, int world, void *const *peer_buffers) {
  if (!f || !peer_buffers || world < 1 || world > VDL_MAX_RANKS || rank < 0 || rank >= world) return VDL_EINVAL;
  memset(&f->xd, 0, sizeof f->xd);
  f->xd.rank = rank;
  f->xd.world = world;
  f->xd.stride = f->fd.part_stride;
  f->xd.timeout_ns = 10000000000ull;                     // 10 s; VDL_PEER_TIMEOUT_MS overrides (tests)
  if (const char *e = getenv("VDL_PEER_TIMEOUT_MS")) f->xd.timeout_ns = (u64)atoll(e) * 1000000ull;
  for (int r = 0; r < world; r++) {
    if (!peer_buffers[r]) return vdl_fail(f->ctx, VDL_EINVAL, "set_peers: buffer of rank %d is null", r);
    f->xd.peer[r] = (i64 *)peer_buffers[r];
  }
  return VDL_OK;
}

extern "C" int vdl_fused_launch(vdl_fused *f) { return vdl_fused_launch_ex(f, 0); }

extern "C" int vdl_fused_partials(vdl_fused *f, void **device_ptr, int64_t *n_int64) {
  if (!f || !device_ptr || !n_int64) return VDL_EINVAL;
  *device_ptr = f->kd.table;
  *n_int64 = f->fd.part_stride;
  return VDL_OK;
}

extern "C" int vdl_fused_finalize(vdl_fused *f, const void *all_partials, int nranks) {
  if (!f) return VDL_EINVAL;
  vdl_ctx *ctx = f->ctx;
  if (nranks < 1 || (nranks > 1 && !all_partials)) return vdl_fail(ctx, VDL_EINVAL, "finalize: nranks %d without gathered partials", nranks);
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  f->fd.parts = all_partials ? (const i64 *)all_partials : f->kd.table;
  f->fd.nranks = nranks;
  f->fd.reset_table = f->kd.table;      // the gathered copies (or this one launch) are the last readers of the table
  fused_finalize_kernel<<<1, 256, 0, ctx->stream>>>(f->fd);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  f->table_clean = true;
  f->finalized = true;
  f->ngroups = -1;
  return VDL_OK;
}

// One device->host copy brings the group count, the error counter and every fold result.
static int fused_fetch(vdl_fused *f) {
  vdl_ctx *ctx = f->ctx;
  if (!f->finalized) return vdl_fail(ctx, VDL_EINVAL, "fused scan not finalized");
  if (f->ngroups >= 0) return VDL_OK;
  size_t n = (size_t)(f->nout + f->fd.npost) * f->kd.domain + 2;
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // the finalize wrote the results straight into h_outbuf (mapped)
  i64 ng = f->h_outbuf[n - 2], err = f->h_outbuf[n - 1];
  if (err) {
    cudaMemsetAsync(ctx->d_errflag, 0, sizeof(int), ctx->stream);
    if (err >= (1 << 20)) return vdl_fail(ctx, VDL_ECUDA, "fused scan: a peer GPU never delivered its partial table (exchange timed out)");
    return vdl_fail(ctx, VDL_ERANGE, "fused scan: %lld rows produced a group key outside the key domain", (long long)err);
  }
  f->ngroups = ng;
  f->groups_seen = ng;
  for (int i = 0; i < f->nout; i++) ctx->vecs[f->out[i]].len = ng;
  return VDL_OK;
}

extern "C" int vdl_fused_num_groups(vdl_fused *f, int64_t *ngroups) {
  if (!f || !ngroups) return VDL_EINVAL;
  VDL_TRY(fused_fetch(f));
  *ngroups = f->ngroups;
  return VDL_OK;
}

// Host copy of fold `fold_index`'s result (valid until the next launch); no device work beyond the one fetch.
extern "C" int vdl_fused_result_host(vdl_fused *f, int fold_index, const int64_t **data, int64_t *len) {
  if (!f || !data || !len || fold_index < 0 || fold_index >= f->nout) return VDL_EINVAL;
  VDL_TRY(fused_fetch(f));
  *data = f->h_outbuf + (size_t)fold_index * f->kd.domain;
  *len = f->ngroups;
  return VDL_OK;
}

extern "C" int vdl_fused_post_host(vdl_fused *f, int post_index, const int64_t **data, int64_t *len) {
  if (!f || !data || !len || post_index < 0 || post_index >= f->fd.npost) return VDL_EINVAL;
  VDL_TRY(fused_fetch(f));
  *data = f->h_outbuf + (size_t)(f->nout + post_index) * f->kd.domain;
  *len = f->ngroups;
  return VDL_OK;
}

extern "C" int vdl_fused_result(vdl_fused *f, int fold_index, vdl_vec *out) {
  if (!f || !out || fold_index < 0 || fold_index >= f->nout) return VDL_EINVAL;
  int64_t n;
  VDL_TRY(vdl_fused_num_groups(f, &n));
  *out = f->out[fold_index];
  return VDL_OK;
}

extern "C" int vdl_fused_last_kernel_ms(vdl_fused *f, float *ms) {
  if (!f || !ms) return VDL_EINVAL;
  if (!f->timed) return vdl_fail(f->ctx, VDL_EINVAL, "fused scan not launched yet");
  VDL_CUDA(f->ctx, cudaEventSynchronize(f->ev1));
  VDL_CUDA(f->ctx, cudaEventElapsedTime(ms, f->ev0, f->ev1));
  return VDL_OK;
}

extern "C" const char *vdl_fused_shape_name(vdl_fused *f) { return f ? f->shape : ""; }

extern "C" int vdl_fused_destroy(vdl_fused *f) {
  if (!f) return VDL_EINVAL;
  vdl_ctx *ctx = f->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (f->table) vdl_vec_free(ctx, f->table);
  for (int i = 0; i < f->nout; i++)
    if (f->out[i]) vdl_vec_free(ctx, f->out[i]);
  if (f->d_outbuf) cudaFree(f->d_outbuf);
  if (f->d_done) cudaFree(f->d_done);
  if (f->h_outbuf) cudaFreeHost(f->h_outbuf);
  if (f->ev0) cudaEventDestroy(f->ev0);
  if (f->ev1) cudaEventDestroy(f->ev1);
  delete f;
  return VDL_OK;
}
