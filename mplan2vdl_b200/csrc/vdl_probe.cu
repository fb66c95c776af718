// Fused FK-join probe: ONE pass over a fact-table shard that follows foreign-key index columns into
// dimension columns (replicated in HBM, mostly L2-resident), applies the selection predicates of the whole
// join chain, and either folds per group key (Q5-class plans) or emits the surviving rows' expressions as
// dense vectors in row order (Q3-class plans, whose high-cardinality group-by continues op-at-a-time on the
// few surviving rows).
//
// What it replaces in the emitted graph (reference Vlite.hs): handleGatherJoin / deduceMasks
// (Vlite.hs:1199-1232, 1248-1280; diagram 1420-1447) lower an FK join to
//     fk'      = Gather(Load fact.fk_idx, factmask)
//     valid    = Scatter(ones -> dimmask positions)         inv = Scatter(pos -> dimmask positions)
//     boolean  = Gather(valid, fk')                         gmask = Gather(inv, fk')
//     selmask  = FoldSelect(pos_ boolean, boolean)
//     fact cols, gmask <- Gather(., selmask)                dim cols <- Gather(Gather(col, dimmask), gmask)
// and chain that per join.  In the dense model every vector of the joined row space is a
// function of the underlying fact row i:  dim column e -> e[fk(i)],  nested joins -> e[fk2[fk1(i)]], and the
// selection is the conjunction of the fact predicates and the dimension predicates evaluated at fk(i).  The
// planner (vdl_plan.cu, analyse) proves that normal form; this kernel evaluates it:
//   leaf      value(i) = column[ parent leaf's value(i) ]   (parent < 0: column[i] of the fact table)
//   term      a + b * (leaf >> shr), or a constant, or the row id
//   predicate lo <= term <= hi, or term == term; evaluated in chain order, first failure rejects the row
// No validity vector, inverse index, position vector or gathered copy is ever materialised.
//
// Kernel shape: persistent thread blocks take 4096-row tiles in order from a ticket counter and evaluate the
// predicates stage by stage, compacting the survivors in row order after each (see "staged evaluation" below; the
// fact-column loads of a stage are coalesced, and the lineitem->orders index is clustered, so the first dimension
// gather is nearly sequential too).  Fold mode accumulates into a shared-memory table per block (flushed with
// global atomics); emit mode places a tile's survivors with a decoupled look-back across tiles (one 64-bit
// status word per tile), so the output vectors are exactly the dense FoldSelect order.  Bound: HBM for the fact columns + L2 latency for the dimension gathers.
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "vdl_internal.h"

#include "vdl_probe_kernel.cuh"

// value of leaf l at fact row `row` (local to the shard): walk the parent chain, then load top-down
__device__ __forceinline__ i64 leaf_value(const PDesc &d, int l, i64 row, bool &ok) {
  int chain[P_MAX_DEPTH];
  int n = 0;
#pragma unroll
  for (int k = 0; k < P_MAX_DEPTH; k++)
    if (l >= 0) { chain[k] = l; n = k + 1; l = d.leaf[l].parent; }
  i64 idx = row;
#pragma unroll
  for (int k = P_MAX_DEPTH - 1; k >= 0; k--) {
    if (k < n) {
      const PLeaf &L = d.leaf[chain[k]];
      if ((u64)idx >= (u64)L.len) { ok = false; return 0; }
      idx = L.w4 ? (i64)__ldg((const int32_t *)L.ptr + idx) : __ldg((const i64 *)L.ptr + idx);
    }
  }
  return idx;
}
__device__ __forceinline__ i64 plain_term_value(const PDesc &d, const PTerm &t, i64 row, bool &ok) {
  if (t.leaf == -1) return t.a;
  i64 v = t.leaf == -2 ? d.row_base + row : leaf_value(d, t.leaf, row, ok);
  if (t.shr) v >>= t.shr;
  return (i64)((u64)t.a + (u64)t.b * (u64)v);
}
// does predicate P hold at the row?  (its terms are plain: no indicators inside predicates)
__device__ __forceinline__ bool pred_holds(const PDesc &d, const PPred &P, i64 row, bool &ok) {
  const i64 t = plain_term_value(d, P.t, row, ok);
  if (P.kind == 0) {
    if ((u64)t - (u64)P.lo <= P.span) return true;
    for (int k = 0; k < P.nmore; k++)
      if ((u64)t - (u64)P.lo_more[k] <= P.span_more[k]) return true;
    return false;
  }
  const i64 u = plain_term_value(d, P.u, row, ok);
  switch (P.cmp) {
    case VDL_CMP_EQ: return t == u;
    case VDL_CMP_NE: return t != u;
    case VDL_CMP_GT: return t > u;
    case VDL_CMP_GE: return t >= u;
    case VDL_CMP_LT: return t < u;
    default: return t <= u;
  }
}
__device__ __forceinline__ i64 term_value(const PDesc &d, const PTerm &t, i64 row, bool &ok) {
  if (t.leaf <= -3) {                       // indicator of a predicate as a 0/1 value
    const i64 v = pred_holds(d, d.ind[-3 - t.leaf], row, ok) ? 1 : 0;
    return (i64)((u64)t.a + (u64)t.b * (u64)v);
  }
  return plain_term_value(d, t, row, ok);
}
__device__ __forceinline__ i64 prod_value(const PDesc &d, const PProd &p, i64 row, bool &ok) {
  i64 v = 1;
  for (int f = 0; f < p.nfac; f++) v = (i64)((u64)v * (u64)term_value(d, p.f[f], row, ok));
  return v;
}
__device__ __forceinline__ bool row_passes(const PDesc &d, i64 row, bool &ok) {
  for (int q = 0; q < d.npreds; q++) {
    if (!pred_holds(d, d.pred[q], row, ok) || !ok) return false;
  }
  return true;
}

// ---- staged evaluation ---------------------------------------------------------------------------------------------
// A tile goes through the predicates one STAGE at a time; after every stage the surviving rows are compacted (in row
// order) into a shared-memory queue, so the next stage runs with full warps on survivors only and its lookups are
// independent loads of different rows (memory-level parallelism instead of divergent lanes waiting on a long
// dependent chain).  A range predicate on a lookup chain of depth <= 4 -- every predicate the FK-join lowering
// produces -- runs as straight-line code: the chain (pointers, lengths, widths) is read from the descriptor once per
// tile and stage, not per row.
struct StageRegs { const void *ptr[4]; i64 len[4]; int w4mask, shr, depth; i64 a, b, lo; u64 span; };

// S.ptr[0] is the leaf's own column (the LAST load), S.ptr[D-1] the fact column (the first): all indexing static.
// The first hop is indexed by the fact row itself (in range: prepare checks the fact columns' lengths); PLAIN: the
// term is the bare leaf (a = 0, b = 1), as in every predicate the lowering emits.
template <int D, bool PLAIN>
__device__ __forceinline__ bool chain_test(const StageRegs &S, i64 row, i64 row_base, bool &ok) {
  i64 v = row;
  if (D == 0) v = row_base + row;
#pragma unroll
  for (int k = D - 1; k >= 0; k--) {
    if (k != D - 1 && (u64)v >= (u64)S.len[k]) { ok = false; return false; }
    v = ((S.w4mask >> k) & 1) ? (i64)__ldg((const int32_t *)S.ptr[k] + v) : __ldg((const i64 *)S.ptr[k] + v);
  }
  v >>= S.shr;
  if (!PLAIN) v = (i64)((u64)S.a + (u64)S.b * (u64)v);
  return (u64)v - (u64)S.lo <= S.span;
}

#ifndef P_SUB
#define P_SUB 4      // rows per thread and round of a stage (8 was best while every round ended in two barriers)
#endif

__global__ void __launch_bounds__(P_THREADS, P_BLOCKS) probe_kernel(const __grid_constant__ PDesc d) {
  extern __shared__ __align__(16) unsigned char psm[];
  __shared__ uint16_t queue[2][P_TILE];
  __shared__ int s_cnt[3];                        // survivors appended by stage q: s_cnt[(q + 1) % 3]
  __shared__ unsigned int s_bits[P_TILE / 32];    // emit mode: the tile's survivors as a bitmap, to put them back in row order
  __shared__ int wsum[P_THREADS / 32];
  __shared__ i64 s_off;
  __shared__ unsigned int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool folding = d.nemits == 0;
  const int nacc = d.nfolds + 2;
  i64 *stab = (i64 *)psm;                         // [nacc][domain] when smem_table
  if (folding && d.smem_table) {
    for (i64 i = tid; i < (i64)nacc * d.domain; i += P_THREADS) {
      int j = (int)(i / d.domain);
      stab[i] = j < d.nfolds ? p_identity(d.fold_op[j]) : (j == d.nfolds ? 0 : INT64_MAX);
    }
  }
  __syncthreads();
  i64 *tab = (folding && d.smem_table) ? stab : d.table;

  for (;;) {
    if (tid == 0) { s_tile = atomicAdd(d.ticket, 1u); s_cnt[1] = 0; }
    __syncthreads();
    const i64 tile = s_tile;
    if (tile >= d.ntiles) break;
    const i64 base = tile * P_TILE;
    if (d.npf) {
      const i64 b2 = base + (i64)d.pf_dist * P_TILE;
      const i64 n2 = min((i64)P_TILE, d.rows - b2);
      for (int c = 0; c < d.npf && n2 > 0; c++) {
        const unsigned char *p0 = d.pf_ptr[c] + (b2 << d.pf_shift[c]);
        const i64 bytes = n2 << d.pf_shift[c];
        for (i64 o = (i64)tid * 128; o < bytes; o += P_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
      }
    }
    bool ok = true;
    int n_in = (int)min((i64)P_TILE, d.rows - base);      // stage input: all rows of the tile, then the previous stage's survivors
    int cur = 0;
    for (int q = 0; q < d.npreds && n_in > 0; q++) {
      const PPred &P = d.pred[q];
      // the chain of this stage's term, if it has the fast form
      StageRegs S;
      S.depth = -1;
      if (P.kind == 0 && P.nmore == 0 && P.t.leaf != -1) {
        int l = P.t.leaf, n = 0;
        S.w4mask = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          S.ptr[k] = nullptr; S.len[k] = 0;
          if (l >= 0) {
            const PLeaf &L = d.leaf[l];
            S.ptr[k] = L.ptr; S.len[k] = L.len; S.w4mask |= (L.w4 ? 1 : 0) << k;
            l = L.parent; n = k + 1;
          }
        }
        if (l < 0) { S.depth = n; S.shr = P.t.shr; S.a = P.t.a; S.b = P.t.b; S.lo = P.lo; S.span = P.span; if (P.t.a == 0 && P.t.b == 1) S.depth += 8; }   // +8: plain
      }
      for (int j0 = 0; j0 < n_in; j0 += P_SUB * P_THREADS) {
        bool f[P_SUB];
        int r[P_SUB];
#pragma unroll
        for (int k = 0; k < P_SUB; k++) {
          const int j = j0 + k * P_THREADS + tid;
          r[k] = j < n_in ? (q == 0 ? j : (int)queue[cur][j]) : -1;
        }
#define STAGE_CASE(DEPTH, PL)                                                                                                   \
  _Pragma("unroll") for (int k = 0; k < P_SUB; k++) f[k] = r[k] >= 0 && chain_test<DEPTH, PL>(S, base + r[k], d.row_base, ok)
        if (S.depth == 8 + 2) { STAGE_CASE(2, true);
        } else if (S.depth == 8 + 1) { STAGE_CASE(1, true);
        } else if (S.depth == 8 + 3) { STAGE_CASE(3, true);
        } else if (S.depth == 8 + 4) { STAGE_CASE(4, true);
        } else if (S.depth == 1) { STAGE_CASE(1, false);
        } else if (S.depth == 2) { STAGE_CASE(2, false);
        } else if (S.depth == 3) { STAGE_CASE(3, false);
        } else if (S.depth == 4) { STAGE_CASE(4, false);
        } else if (S.depth == 0 || S.depth == 8) { STAGE_CASE(0, false);
        } else {      // generic: constants, column-vs-column comparisons, range sets, deeper chains
#pragma unroll 1
          for (int k = 0; k < P_SUB; k++) f[k] = r[k] >= 0 && pred_holds(d, P, base + r[k], ok);
        }
        // append this round's survivors to the next stage's queue.  Queue order is irrelevant to a stage (and to a
        // fold); emit mode restores row order once, at the end.  One shared atomic per warp and round, no barrier.
        unsigned m[P_SUB];
        int wtotal = 0;
#pragma unroll
        for (int k = 0; k < P_SUB; k++) { m[k] = __ballot_sync(0xffffffffu, f[k]); wtotal += __popc(m[k]); }
        if (wtotal) {
          int at = 0;
          if (lane == 0) at = atomicAdd(&s_cnt[(q + 1) % 3], wtotal);
          at = __shfl_sync(0xffffffffu, at, 0);
#pragma unroll
          for (int k = 0; k < P_SUB; k++) {
            if (f[k]) queue[cur ^ 1][at + __popc(m[k] & ((1u << lane) - 1))] = (uint16_t)r[k];
            at += __popc(m[k]);
          }
        }
      }
      // the counter two stages ahead is idle: every thread read it (as its n_in) before the previous stage's barrier
      if (tid == 0) s_cnt[(q + 2) % 3] = 0;
      __syncthreads();
      n_in = s_cnt[(q + 1) % 3];
      cur ^= 1;
    }
    const bool implicit = d.npreds == 0;            // no predicate at all: every row of the tile survives
    // ---- survivors: fold into the table, or emit in order
    if (folding) {
      for (int j = tid; j < n_in; j += P_THREADS) {
        const i64 row = base + (implicit ? j : (int)queue[cur][j]);
        i64 key = 0;
        for (int q = 0; q < d.nkeys; q++) key |= (i64)((u64)term_value(d, d.key[q], row, ok) << d.key_shl[q]);
        key &= d.key_mask;
        if ((u64)key >= (u64)d.domain) { ok = false; continue; }
        for (int j2 = 0; j2 < d.nfolds; j2++) {
          const int op = d.fold_op[j2];
          if (op == VDL_FOLD_CHOOSE || op == VDL_FOLD_COUNT) continue;     // from the first row / the row count
          table_update(op, tab + (size_t)j2 * d.domain + key, prod_value(d, d.fold[j2], row, ok));
        }
        atomicAdd((unsigned long long *)(tab + (size_t)d.nfolds * d.domain + key), 1ull);
        atomicMin((long long *)(tab + (size_t)(d.nfolds + 1) * d.domain + key), (long long)(d.row_base + row));
      }
    } else {
      if (!implicit) {
        // survivors back into row order: bitmap of the tile, one 32-row word per thread, block scan of the word counts
        static_assert(P_TILE / 32 == P_THREADS, "one bitmap word per thread");
        s_bits[tid] = 0;
        __syncthreads();
        for (int j = tid; j < n_in; j += P_THREADS) { const unsigned rr = queue[cur][j]; atomicOr(&s_bits[rr >> 5], 1u << (rr & 31)); }
        __syncthreads();
        unsigned w = s_bits[tid];
        const int c = __popc(w);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int at = incl - c;
        for (int ww = 0; ww < warp; ww++) at += wsum[ww];
        while (w) { const int b = __ffs(w) - 1; w &= w - 1; queue[cur ^ 1][at++] = (uint16_t)(tid * 32 + b); }
        cur ^= 1;
        __syncthreads();
      }
      // tile offset by decoupled look-back over the tiles before this one
      if (tid == 0) {
        const unsigned long long T = (unsigned long long)n_in;
        i64 excl = 0;
        if (tile == 0) {
          atomicExch(d.tile_state, (2ull << 62) | T);
        } else {
          atomicExch(d.tile_state + tile, (1ull << 62) | T);
          for (i64 j = tile - 1;; j--) {
            unsigned long long st;
            do { st = *((volatile unsigned long long *)(d.tile_state + j)); } while ((st >> 62) == 0);
            excl += (i64)(st & ((1ull << 62) - 1));
            if ((st >> 62) == 2) break;
          }
          atomicExch(d.tile_state + tile, (2ull << 62) | (unsigned long long)(excl + (i64)T));
        }
        s_off = excl;
        if (tile == d.ntiles - 1) *d.total = excl + (i64)T;
      }
      __syncthreads();
      const i64 off = s_off;
      for (int j = tid; j < n_in; j += P_THREADS) {
        const i64 row = base + (implicit ? j : (int)queue[cur][j]);
        for (int e = 0; e < d.nemits; e++) d.emit_out[e][off + j] = prod_value(d, d.emit[e], row, ok);
      }
    }
    if (!ok) atomicAdd(d.errflag, 1);
    __syncthreads();
  }
  if (folding && d.smem_table) {
    __syncthreads();
    for (i64 i = tid; i < (i64)nacc * d.domain; i += P_THREADS) {
      int j = (int)(i / d.domain);
      const i64 v = stab[i];
      if (j < d.nfolds) { if (v != p_identity(d.fold_op[j])) table_update(d.fold_op[j], d.table + i, v); }
      else if (j == d.nfolds) { if (v) atomicAdd((unsigned long long *)(d.table + i), (unsigned long long)v); }
      else if (v != INT64_MAX) atomicMin((long long *)(d.table + i), (long long)v);
    }
  }
}

__global__ void probe_init_kernel(const __grid_constant__ PDesc d) {
  const i64 n = (i64)(d.nfolds + 2) * d.domain;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    int j = (int)(i / d.domain);
    d.table[i] = j < d.nfolds ? p_identity(d.fold_op[j]) : (j == d.nfolds ? 0 : INT64_MAX);
  }
}

// External combine (several ranks): every rank evaluates FoldChoose at ITS first row of each key and stores the value in
// the fold's (otherwise unused) table row, so the tables alone carry everything the merge needs.
__global__ void probe_choose_kernel(const __grid_constant__ PDesc d) {
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < d.domain; k += (i64)gridDim.x * blockDim.x) {
    if (d.table[(size_t)d.nfolds * d.domain + k] <= 0) continue;
    const i64 first = d.table[(size_t)(d.nfolds + 1) * d.domain + k] - d.row_base;
    bool ok = true;
    for (int j = 0; j < d.nfolds; j++)
      if (d.fold_op[j] == VDL_FOLD_CHOOSE) d.table[(size_t)j * d.domain + k] = prod_value(d, d.fold[j], first, ok);
    if (!ok) atomicAdd(d.errflag, 1);
  }
}

// One dense vector per fold in ascending key order (keys without rows dropped: G14), FoldChoose evaluated at the
// key's first row, then the post ops; results mirrored into mapped host memory.
__global__ void __launch_bounds__(256, 1) probe_finalize_kernel(const __grid_constant__ PDesc d, const __grid_constant__ PFin f) {
  __shared__ int warp_cnt[8];
  __shared__ i64 running;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) running = 0;
  __syncthreads();
  for (i64 base = 0; base < f.domain; base += 256) {
    const i64 k = base + tid;
    i64 cnt = 0;
    if (k < f.domain)
      for (int r = 0; r < f.nranks; r++) cnt += f.table[(size_t)r * f.stride + (size_t)f.nfolds * f.domain + k];
    const bool exists = cnt > 0;
    unsigned m = __ballot_sync(0xffffffffu, exists);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 8; w++) { if (w < warp) before += warp_cnt[w]; total += warp_cnt[w]; }
    if (exists) {
      const i64 pos = running + before + __popc(m & ((1u << lane) - 1));
      int best = 0;                 // rank that holds the key's first row
      i64 firstg = INT64_MAX;
      for (int r = 0; r < f.nranks; r++) {
        const i64 fr = f.table[(size_t)r * f.stride + (size_t)(f.nfolds + 1) * f.domain + k];
        if (fr < firstg) { firstg = fr; best = r; }
      }
      const i64 first = firstg - d.row_base;
      i64 ov[VDL_MAX_AGGS], pv[VDL_MAX_POSTS];
      bool ok = true;
      for (int j = 0; j < f.nfolds; j++) {
        const int op = f.fold_op[j];
        i64 v;
        if (op == VDL_FOLD_COUNT) v = cnt;
        else if (op == VDL_FOLD_CHOOSE) v = f.precomputed_choose ? f.table[(size_t)best * f.stride + (size_t)j * f.domain + k] : prod_value(d, d.fold[j], first, ok);
        else {
          v = p_identity(op);
          for (int r = 0; r < f.nranks; r++) {
            const i64 x = f.table[(size_t)r * f.stride + (size_t)j * f.domain + k];
            v = op == VDL_FOLD_MIN ? (x < v ? x : v) : (op == VDL_FOLD_MAX ? (x > v ? x : v) : (i64)((u64)v + (u64)x));
          }
        }
        ov[j] = v;
        f.out[(size_t)j * f.domain + pos] = v;
        if (f.hmirror) f.hmirror[(size_t)j * f.domain + pos] = v;
      }
      for (int q = 0; q < f.npost; q++) {
        const vdl_post_op &P = f.post[q];
        i64 a = P.a_kind == VDL_POST_CONST ? P.a : (P.a_kind == VDL_POST_FOLD ? ov[P.a] : pv[P.a]);
        i64 b = P.b_kind == VDL_POST_CONST ? P.b : (P.b_kind == VDL_POST_FOLD ? ov[P.b] : pv[P.b]);
        pv[q] = binop_apply(P.op, a, b);
        f.out[(size_t)(f.nfolds + q) * f.domain + pos] = pv[q];
        if (f.hmirror) f.hmirror[(size_t)(f.nfolds + q) * f.domain + pos] = pv[q];
      }
    }
    __syncthreads();
    if (tid == 0) running += total;
    __syncthreads();
  }
  if (tid == 0) {
    const size_t tail = (size_t)(f.nfolds + f.npost) * f.domain;
    const i64 ng = running, err = *f.errflag;
    f.out[tail] = ng; f.out[tail + 1] = err;
    if (f.hmirror) {
      f.hmirror[tail] = ng; f.hmirror[tail + 1] = err;
      __threadfence_system();
      ((volatile i64 *)f.hmirror)[tail + 2] = f.seq;
    }
  }
}

// ------------------------------------------------------------------------------ host side
// Multi-GPU combine over peer memory, as in the fused scan (XDesc, vdl_internal.h): one block stores this rank's partial
// table (FoldChoose values included) into slot [parity][rank] of EVERY rank's exchange buffer over NVLink, publishes the
// step's epoch and waits for the other ranks' flags; the finalize kernel that follows merges the `world` tables that
// arrived in this rank's own buffer.  A peer that never arrives: error flag after the timeout, never a hang.
__global__ void __launch_bounds__(256, 1) probe_exchange_kernel(const __grid_constant__ XDesc x, const i64 *table, int *errflag) {
  const int tid = threadIdx.x;
  const int par = (int)(x.epoch & 1);
  const size_t slot = ((size_t)par * x.world + x.rank) * x.stride, flags = (size_t)2 * x.world * x.stride;
  for (int p = 0; p < x.world; p++) {
    i64 *dst = x.peer[p] + slot;
    for (i64 i = tid; i < x.stride; i += 256) dst[i] = __ldcg(&table[i]);
  }
  __threadfence_system();
  __syncthreads();
  if (tid < x.world) {
    st_release_sys((u64 *)(x.peer[tid] + flags) + (size_t)par * x.world + x.rank, x.epoch);
    const u64 *mine = (const u64 *)(x.peer[x.rank] + flags) + (size_t)par * x.world + tid;
    const u64 t0 = global_timer_ns();
    while (ld_acquire_sys(mine) < x.epoch) {
      if (global_timer_ns() - t0 > x.timeout_ns) { atomicAdd(errflag, 1 << 20); break; }
      __nanosleep(64);
    }
  }
}

#include <string>
#include "vdl_embedded.inc"

// ---- run-time specialisation of the probe (NVRTC) ---------------------------------------------------------------------
// probe_kernel above INTERPRETS the descriptor: per row and stage it walks parent chains through the leaf table, dispatches
// on predicate kinds and reads bounds from the constant bank -- 160 thread-instructions per lineitem row on Q5, a handful of
// them loads (profiles/r01e_q05_sf10_dominant_kernel.txt).  For tables of a few hundred thousand rows and up the descriptor
// is instead PRINTED as CUDA C: one straight-line function per predicate stage (typed loads along the lookup chain, bounds
// and constants as literals, out-of-range lookups turned into a flag and a clamped index so that no branch separates the
// loads), one for the fold / emit of a surviving row, and the same tile loop around them.  The P_SUB rows a thread handles
// per round are independent straight-line chains, so their loads overlap (memory-level parallelism instead of one
// dependent chain per warp).  Same results as the interpreter (tests/test_gpu_probe.py and the fuzz plans run both).
struct ProbeGen {
  const PDesc &d;
  std::string body;                 // statements of the function being generated
  bool have[VDL_MAX_LEAVES];
  int tmp = 0;
  explicit ProbeGen(const PDesc &desc) : d(desc) { reset(); }
  void reset() { body.clear(); memset(have, 0, sizeof have); tmp = 0; }
  static std::string lit(i64 v) { char b[48]; snprintf(b, sizeof b, "((i64)0x%llxull)", (unsigned long long)v); return b; }
  static std::string ulit(u64 v) { char b[48]; snprintf(b, sizeof b, "0x%llxull", (unsigned long long)v); return b; }
  std::string leaf(int l) {
    char nm[16], b[320];
    snprintf(nm, sizeof nm, "L%d", l);
    if (have[l]) return nm;
    have[l] = true;
    const PLeaf &L = d.leaf[l];
    const char *ld = L.w4 ? "(i64)__ldg((const int *)d.leaf[%d].ptr + %s)" : "__ldg((const i64 *)d.leaf[%d].ptr + %s)";
    char load[200];
    if (L.parent < 0) {
      snprintf(load, sizeof load, ld, l, "row");
      snprintf(b, sizeof b, "  const i64 %s = %s;\n", nm, load);
    } else {
      const std::string p = leaf(L.parent);
      char idx[64];
      snprintf(idx, sizeof idx, "(in%d ? %s : 0)", l, p.c_str());
      snprintf(load, sizeof load, ld, l, idx);
      snprintf(b, sizeof b, "  const bool in%d = (u64)%s < (u64)d.leaf[%d].len; ok = ok && in%d;\n  const i64 %s = %s;\n", l, p.c_str(), l, l, nm, load);
    }
    body += b;
    return nm;
  }
  std::string term(const PTerm &t) {
    if (t.leaf == -1) return lit(t.a);
    std::string v;
    if (t.leaf == -2) v = "(d.row_base + row)";
    else if (t.leaf <= -3) v = "(i64)(" + pred(d.ind[-3 - t.leaf]) + " ? 1 : 0)";
    else {
      v = leaf(t.leaf);
      if (t.shr) v = "(" + v + " >> " + std::to_string(t.shr) + ")";
    }
    if (t.a == 0 && t.b == 1) return v;
    return "(i64)((u64)" + lit(t.a) + " + (u64)" + lit(t.b) + " * (u64)" + v + ")";
  }
  std::string pred(const PPred &P) {
    char nm[16];
    snprintf(nm, sizeof nm, "T%d", tmp++);
    body += std::string("  const i64 ") + nm + " = " + term(P.t) + ";\n";
    if (P.kind == 0) {
      std::string e = std::string("((u64)") + nm + " - (u64)" + lit(P.lo) + " <= " + ulit(P.span);
      for (int k = 0; k < P.nmore; k++) e += std::string(" || (u64)") + nm + " - (u64)" + lit(P.lo_more[k]) + " <= " + ulit(P.span_more[k]);
      return e + ")";
    }
    char un[16];
    snprintf(un, sizeof un, "T%d", tmp++);
    body += std::string("  const i64 ") + un + " = " + term(P.u) + ";\n";
    static const char *CMP[] = {"==", "!=", ">", ">=", "<", "<="};
    return std::string("(") + nm + " " + CMP[P.cmp] + " " + un + ")";
  }
  std::string product(const PProd &p) {
    if (p.nfac == 0) return lit(1);
    std::string e = "(i64)(";
    for (int f = 0; f < p.nfac; f++) e += std::string(f ? " * " : "") + "(u64)" + term(p.f[f]);
    return e + ")";
  }
};

static const char *PROBE_JIT_TAIL = R"SKEL(
#define STAGE_LOOP(Q, FN)                                                                                          \
  if (n_in > 0) {                                                                                                  \
    for (int j0 = 0; j0 < n_in; j0 += P_SUB * P_THREADS) {                                                         \
      bool f[P_SUB];                                                                                               \
      int r[P_SUB];                                                                                                \
      _Pragma("unroll") for (int k = 0; k < P_SUB; k++) {                                                          \
        const int j = j0 + k * P_THREADS + tid;                                                                    \
        r[k] = j < n_in ? ((Q) == 0 ? j : (int)queue[cur][j]) : -1;                                                \
      }                                                                                                            \
      /* every slot evaluates a row (its own, or the tile's row 0 as a stand-in): straight-line, loads overlap */ \
      _Pragma("unroll") for (int k = 0; k < P_SUB; k++) {                                                          \
        bool okk = true;                                                                                           \
        const bool pass = FN(d, base + (r[k] >= 0 ? r[k] : 0), okk);                                               \
        f[k] = r[k] >= 0 && pass && okk;                                                                           \
        if (r[k] >= 0 && !okk) ok = false;                                                                         \
      }                                                                                                            \
      unsigned m[P_SUB];                                                                                           \
      int wtotal = 0;                                                                                              \
      _Pragma("unroll") for (int k = 0; k < P_SUB; k++) { m[k] = __ballot_sync(0xffffffffu, f[k]); wtotal += __popc(m[k]); } \
      if (wtotal) {                                                                                                \
        int at = 0;                                                                                                \
        if (lane == 0) at = atomicAdd(&s_cnt[((Q) + 1) % 3], wtotal);                                              \
        at = __shfl_sync(0xffffffffu, at, 0);                                                                      \
        _Pragma("unroll") for (int k = 0; k < P_SUB; k++) {                                                        \
          if (f[k]) queue[cur ^ 1][at + __popc(m[k] & ((1u << lane) - 1))] = (uint16_t)r[k];                       \
          at += __popc(m[k]);                                                                                      \
        }                                                                                                          \
      }                                                                                                            \
    }                                                                                                              \
  }                                                                                                                \
  if (tid == 0) s_cnt[((Q) + 2) % 3] = 0;                                                                          \
  __syncthreads();                                                                                                 \
  n_in = s_cnt[((Q) + 1) % 3];                                                                                     \
  cur ^= 1;

extern "C" __global__ void __launch_bounds__(P_THREADS, P_JIT_BLOCKS) vdl_probe_jit(const __grid_constant__ PDesc d) {
  extern __shared__ __align__(16) unsigned char psm[];
  __shared__ uint16_t queue[2][P_TILE];
  __shared__ int s_cnt[3];
  __shared__ unsigned int s_bits[P_TILE / 32];
  __shared__ int wsum[P_THREADS / 32];
  __shared__ i64 s_off;
  __shared__ unsigned int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool folding = P_NEMITS == 0;
  constexpr int nacc = P_NFOLDS + 2;
  i64 *stab = (i64 *)psm;
  if (folding && d.smem_table) {
    for (i64 i = tid; i < (i64)nacc * d.domain; i += P_THREADS) {
      int j = (int)(i / d.domain);
      stab[i] = j < P_NFOLDS ? p_identity(d.fold_op[j]) : (j == P_NFOLDS ? 0 : INT64_MAX);
    }
  }
  __syncthreads();
  i64 *tab = (folding && d.smem_table) ? stab : d.table;
  for (;;) {
    if (tid == 0) { s_tile = atomicAdd(d.ticket, 1u); s_cnt[1] = 0; }
    __syncthreads();
    const i64 tile = s_tile;
    if (tile >= d.ntiles) break;
    const i64 base = tile * P_TILE;
    if (d.npf) {
      const i64 b2 = base + (i64)d.pf_dist * P_TILE;
      const i64 n2 = min((i64)P_TILE, d.rows - b2);
      for (int c = 0; c < d.npf && n2 > 0; c++) {
        const unsigned char *p0 = d.pf_ptr[c] + (b2 << d.pf_shift[c]);
        const i64 bytes = n2 << d.pf_shift[c];
        for (i64 o = (i64)tid * 128; o < bytes; o += P_THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
      }
    }
    bool ok = true;
    int n_in = (int)min((i64)P_TILE, d.rows - base);
    int cur = 0;
    P_STAGES
    constexpr bool implicit = P_NPREDS == 0;
    if (folding) {
      for (int j = tid; j < n_in; j += P_THREADS) probe_fold_row(d, base + (implicit ? j : (int)queue[cur][j]), tab, ok);
    } else {
      if (!implicit) {
        s_bits[tid] = 0;
        __syncthreads();
        for (int j = tid; j < n_in; j += P_THREADS) { const unsigned rr = queue[cur][j]; atomicOr(&s_bits[rr >> 5], 1u << (rr & 31)); }
        __syncthreads();
        unsigned w = s_bits[tid];
        const int c = __popc(w);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int at = incl - c;
        for (int ww = 0; ww < warp; ww++) at += wsum[ww];
        while (w) { const int b = __ffs(w) - 1; w &= w - 1; queue[cur ^ 1][at++] = (uint16_t)(tid * 32 + b); }
        cur ^= 1;
        __syncthreads();
      }
      if (tid == 0) {
        const unsigned long long T = (unsigned long long)n_in;
        i64 excl = 0;
        if (tile == 0) {
          atomicExch(d.tile_state, (2ull << 62) | T);
        } else {
          atomicExch(d.tile_state + tile, (1ull << 62) | T);
          for (i64 j = tile - 1;; j--) {
            unsigned long long st;
            do { st = *((volatile unsigned long long *)(d.tile_state + j)); } while ((st >> 62) == 0);
            excl += (i64)(st & ((1ull << 62) - 1));
            if ((st >> 62) == 2) break;
          }
          atomicExch(d.tile_state + tile, (2ull << 62) | (unsigned long long)(excl + (i64)T));
        }
        s_off = excl;
        if (tile == d.ntiles - 1) *d.total = excl + (i64)T;
      }
      __syncthreads();
      const i64 off = s_off;
      for (int j = tid; j < n_in; j += P_THREADS) probe_emit_row(d, base + (implicit ? j : (int)queue[cur][j]), off + j, ok);
    }
    if (!ok) atomicAdd(d.errflag, 1);
    __syncthreads();
  }
  if (folding && d.smem_table) {
    __syncthreads();
    for (i64 i = tid; i < (i64)nacc * d.domain; i += P_THREADS) {
      int j = (int)(i / d.domain);
      const i64 v = stab[i];
      if (j < P_NFOLDS) { if (v != p_identity(d.fold_op[j])) table_update(d.fold_op[j], d.table + i, v); }
      else if (j == P_NFOLDS) { if (v) atomicAdd((unsigned long long *)(d.table + i), (unsigned long long)v); }
      else if (v != INT64_MAX) atomicMin((long long *)(d.table + i), (long long)v);
    }
  }
}
)SKEL";

static std::string probe_jit_source(const PDesc &d, int p_sub, int blocks) {
  ProbeGen g(d);
  char b[256];
  snprintf(b, sizeof b, "#include \"vdl_probe_kernel.cuh\"\n#define P_SUB %d\n#define P_JIT_BLOCKS %d\n#define P_NPREDS %d\n#define P_NFOLDS %d\n#define P_NEMITS %d\n",
           p_sub, blocks, d.npreds, d.nfolds, d.nemits);
  std::string s = b, stages;
  for (int q = 0; q < d.npreds; q++) {
    g.reset();
    const std::string e = g.pred(d.pred[q]);
    snprintf(b, sizeof b, "__device__ __forceinline__ bool probe_stage_%d(const PDesc &d, const i64 row, bool &ok) {\n", q);
    s += b + g.body + "  return " + e + ";\n}\n";
    snprintf(b, sizeof b, "STAGE_LOOP(%d, probe_stage_%d) ", q, q);
    stages += b;
  }
  // fold of one surviving row (fold mode) / its emitted expressions (emit mode)
  g.reset();
  s += "__device__ __forceinline__ void probe_fold_row(const PDesc &d, const i64 row, i64 *tab, bool &ok) {\n";
  if (d.nemits == 0) {
    std::string keyexpr = "  i64 key = 0;\n";
    std::string pre;
    for (int q = 0; q < d.nkeys; q++) {
      const std::string t = g.term(d.key[q]);
      keyexpr += "  key |= (i64)((u64)" + t + " << " + std::to_string(d.key_shl[q]) + ");\n";
    }
    std::string vals;
    for (int j = 0; j < d.nfolds; j++) {
      if (d.fold_op[j] == VDL_FOLD_CHOOSE || d.fold_op[j] == VDL_FOLD_COUNT) continue;
      snprintf(b, sizeof b, "  table_update(%d, tab + (size_t)%d * d.domain + key, ", d.fold_op[j], j);
      vals += b + g.product(d.fold[j]) + ");\n";
    }
    s += g.body + keyexpr + "  key &= d.key_mask;\n  if ((u64)key >= (u64)d.domain) { ok = false; return; }\n" + vals;
    snprintf(b, sizeof b, "  atomicAdd((unsigned long long *)(tab + (size_t)%d * d.domain + key), 1ull);\n  atomicMin((long long *)(tab + (size_t)%d * d.domain + key), (long long)(d.row_base + row));\n",
             d.nfolds, d.nfolds + 1);
    s += b;
  }
  s += "}\n";
  g.reset();
  s += "__device__ __forceinline__ void probe_emit_row(const PDesc &d, const i64 row, const i64 at, bool &ok) {\n";
  {
    std::string st;
    for (int e = 0; e < d.nemits; e++) { snprintf(b, sizeof b, "  d.emit_out[%d][at] = ", e); st += b + g.product(d.emit[e]) + ";\n"; }
    s += g.body + st;
  }
  s += "}\n#define P_STAGES " + stages + "\n";
  return s + PROBE_JIT_TAIL;
}

static cudaKernel_t probe_jit_compile(vdl_ctx *ctx, const PDesc &d, size_t smem, bool *ok, std::string *log) {
  int p_sub = 8, blocks = 8;       // measured on B200 (Q5 / Q3 / Q12 / Q19 SF10): 8 rows per thread and round, 8 blocks x 64 registers
  if (const char *e = getenv("VDL_PROBE_JIT_SUB")) p_sub = std::max(1, std::min(16, atoi(e)));
  if (const char *e = getenv("VDL_PROBE_JIT_BLOCKS")) blocks = std::max(1, std::min(16, atoi(e)));
  const std::string src = probe_jit_source(d, p_sub, blocks);
  static const char *const headers[] = {EMB_vdl_cuda_h, EMB_vdl_device_cuh, EMB_vdl_probe_kernel_cuh};
  static const char *const names[] = {"vdl_cuda.h", "vdl_device.cuh", "vdl_probe_kernel.cuh"};
  if (getenv("VDL_DEBUG_JIT")) fprintf(stderr, "[vdl jit] probe:\n%s\n", src.substr(0, src.find("#define STAGE_LOOP")).c_str());
  cudaKernel_t kh = vdl_jit_kernel(ctx, "probe|" + src, src, "vdl_probe_jit", 3, headers, names, ok, log);
  if (kh && smem > 48 * 1024 && cudaFuncSetAttribute((const void *)kh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return kh;
}

// Host-only check (no GPU): print a two-stage descriptor with a lookup chain, a range set, a column comparison, an
// indicator and two folds as CUDA C and compile it with NVRTC for sm_100a; then an emit-mode one.
extern "C" int vdl_probe_jit_selftest(char *log, int log_capacity) {
  if (log && log_capacity > 0) log[0] = 0;
  PDesc d;
  memset(&d, 0, sizeof d);
  d.nleaves = 4;
  d.leaf[0] = PLeaf{nullptr, 10, 0, -1}; d.leaf[1] = PLeaf{nullptr, 10, 1, 0}; d.leaf[2] = PLeaf{nullptr, 10, 0, 1}; d.leaf[3] = PLeaf{nullptr, 10, 1, -1};
  d.npreds = 2;
  d.pred[0].kind = 0; d.pred[0].t = PTerm{1, 0, 0, 1}; d.pred[0].lo = 5; d.pred[0].span = 7; d.pred[0].nmore = 1; d.pred[0].lo_more[0] = 40; d.pred[0].span_more[0] = 0;
  d.pred[1].kind = 1; d.pred[1].cmp = VDL_CMP_GE; d.pred[1].t = PTerm{2, 3, -1, 2}; d.pred[1].u = PTerm{3, 0, 0, 1};
  d.nind = 1; d.ind[0] = d.pred[0];
  d.nkeys = 1; d.key[0] = PTerm{2, 3, -2, 1}; d.key_shl[0] = 1; d.key_mask = 31; d.domain = 32;
  d.nfolds = 3; d.fold_op[0] = VDL_FOLD_SUM; d.fold[0].nfac = 2; d.fold[0].f[0] = PTerm{3, 0, 0, 1}; d.fold[0].f[1] = PTerm{-3, 0, 0, 1};
  d.fold_op[1] = VDL_FOLD_MIN; d.fold[1].nfac = 1; d.fold[1].f[0] = PTerm{-2, 0, 0, 1};
  d.fold_op[2] = VDL_FOLD_COUNT;
  std::string l;
  bool ok1 = false, ok2 = false;
  vdl_jit_kernel(nullptr, "", probe_jit_source(d, 4, 12), "vdl_probe_jit", 3, (const char *const[]){EMB_vdl_cuda_h, EMB_vdl_device_cuh, EMB_vdl_probe_kernel_cuh},
                 (const char *const[]){"vdl_cuda.h", "vdl_device.cuh", "vdl_probe_kernel.cuh"}, &ok1, &l);
  if (l == "NVRTC is not installed") return VDL_ENOTFOUND;
  d.nfolds = 0; d.nkeys = 0; d.nemits = 2; d.emit[0] = d.fold[0]; d.emit[1].nfac = 1; d.emit[1].f[0] = PTerm{2, 0, 0, 1};
  if (ok1) vdl_jit_kernel(nullptr, "", probe_jit_source(d, 4, 12), "vdl_probe_jit", 3, (const char *const[]){EMB_vdl_cuda_h, EMB_vdl_device_cuh, EMB_vdl_probe_kernel_cuh},
                          (const char *const[]){"vdl_cuda.h", "vdl_device.cuh", "vdl_probe_kernel.cuh"}, &ok2, &l);
  if (log && log_capacity > 1) snprintf(log, (size_t)log_capacity, "%s", l.c_str());
  return ok1 && ok2 ? VDL_OK : VDL_ECUDA;
}

struct vdl_probe {
  vdl_ctx *ctx = nullptr;
  PDesc pd;
  PFin pf;
  XDesc xd;                            // world 0: no peer exchange configured
  bool folding = true, ran = false, fetched = false, finalized = false;
  vdl_vec table = 0;
  i64 *d_out = nullptr, *h_out = nullptr, *h_mapped = nullptr;
  i64 *d_total = nullptr;              // [0] survivors (emit mode)
  unsigned int *d_ticket = nullptr;
  unsigned long long *d_state = nullptr;
  vdl_vec emit_vec[VDL_MAX_EMITS] = {0};
  i64 ngroups = -1, nselected = -1;
  int grid = 1;
  size_t smem = 0;
  // event pairs around the dominant kernel of the last VDL_EVENT_RING launches: durations are read AFTER a timed loop
  cudaEvent_t ev0[VDL_EVENT_RING] = {nullptr}, ev1[VDL_EVENT_RING] = {nullptr};
  long nlaunch = 0;
  cudaKernel_t jit_kernel = nullptr;   // the descriptor printed as CUDA C and compiled at run time (probe_jit_source)
  // identity of the leaf columns at prepare time (device pointers and lengths are baked into the descriptor)
  int nleaves = 0;
  vdl_vec leaf_handle[VDL_MAX_LEAVES] = {0};
  u64 leaf_gen[VDL_MAX_LEAVES] = {0};
};

bool vdl_probe_current(vdl_probe *p) {
  for (int l = 0; l < p->nleaves; l++) {
    u64 g;
    if (!vec_identity(p->ctx, p->leaf_handle[l], &g) || g != p->leaf_gen[l]) return false;
  }
  return true;
}

static bool term_ok(const vdl_term &t, int nleaves, int nind = 0) { return t.leaf >= -2 - nind && t.leaf < nleaves && t.shr >= 0 && t.shr < 64; }
static PTerm to_p(const vdl_term &t) { return PTerm{t.leaf, t.shr, t.a, t.b}; }

extern "C" int vdl_abi_sizeof_probe_desc(void) { return (int)sizeof(vdl_probe_desc); }

extern "C" int vdl_probe_prepare(vdl_ctx *ctx, const vdl_probe_desc *desc, vdl_probe **out) {
  if (!ctx || !desc || !out) return VDL_EINVAL;
  *out = nullptr;
  if (desc->nleaves < 1 || desc->nleaves > VDL_MAX_LEAVES || desc->npreds < 0 || desc->npreds > VDL_MAX_PROBE_PREDS ||
      desc->nkeys < 0 || desc->nkeys > VDL_MAX_KEYS || desc->nfolds < 0 || desc->nfolds > VDL_MAX_AGGS || desc->nemits < 0 ||
      desc->nemits > VDL_MAX_EMITS || desc->nposts < 0 || desc->nposts > VDL_MAX_POSTS)
    return vdl_fail(ctx, VDL_EINVAL, "probe: descriptor counts out of range");
  if ((desc->nfolds > 0) == (desc->nemits > 0)) return vdl_fail(ctx, VDL_EINVAL, "probe: exactly one of folds / emits must be given");
  if (desc->rows < 0) return vdl_fail(ctx, VDL_EINVAL, "probe: negative row count");
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  vdl_probe *p = new vdl_probe();
  memset(&p->xd, 0, sizeof p->xd);
  p->ctx = ctx;
  PDesc &d = p->pd;
  memset(&d, 0, sizeof d);
  memset(&p->pf, 0, sizeof p->pf);
  auto fail = [&](int rc) { vdl_probe_destroy(p); return rc; };
  d.rows = desc->rows;
  d.row_base = desc->row_base;
  d.nleaves = desc->nleaves;
  for (int l = 0; l < desc->nleaves; l++) {
    Vec *v = vec_get(ctx, desc->leaf[l].column);
    if (!v || v->is_range) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: leaf %d is not a stored vector", l));
    int par = desc->leaf[l].parent;
    if (par >= l || par < -1) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: leaf %d has parent %d (must precede it)", l, par));
    if (par < 0 && v->len < desc->rows) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: fact column %s has %lld rows < %lld", v->name.c_str(), (long long)v->len, (long long)desc->rows));
    int depth = 1;
    for (int q = par; q >= 0; q = desc->leaf[q].parent) depth++;
    if (depth > P_MAX_DEPTH) return fail(vdl_fail(ctx, VDL_EUNSUPPORTED, "probe: lookup chain deeper than %d", P_MAX_DEPTH));
    d.leaf[l] = PLeaf{v->ptr, v->len, v->dtype == VDL_I32, par};
    p->leaf_handle[l] = desc->leaf[l].column;
    p->leaf_gen[l] = v->gen;
    p->nleaves = l + 1;
  }
  auto to_pred = [&](const vdl_probe_pred &s, PPred *o) -> bool {
    if ((s.kind != 0 && s.kind != 1) || !term_ok(s.t, desc->nleaves) || (s.kind == 1 && (!term_ok(s.u, desc->nleaves) || s.cmp < VDL_CMP_EQ || s.cmp > VDL_CMP_LE)) ||
        s.nmore < 0 || s.nmore > VDL_MAX_MORE_RANGES)
      return false;
    memset(o, 0, sizeof *o);
    o->kind = s.kind; o->cmp = s.cmp; o->t = to_p(s.t); o->u = to_p(s.u);
    if (s.kind == 0) {
      // empty ranges are dropped; no range left: never true (0 in [1, 1])
      i64 lo[1 + VDL_MAX_MORE_RANGES], hi[1 + VDL_MAX_MORE_RANGES];
      int n = 0;
      if (s.lo <= s.hi) { lo[n] = s.lo; hi[n++] = s.hi; }
      for (int k = 0; k < s.nmore; k++) if (s.lo_more[k] <= s.hi_more[k]) { lo[n] = s.lo_more[k]; hi[n++] = s.hi_more[k]; }
      if (n == 0) { o->t = PTerm{-1, 0, 0, 0}; o->lo = 1; o->span = 0; return true; }
      o->lo = lo[0]; o->span = (u64)hi[0] - (u64)lo[0];
      o->nmore = n - 1;
      for (int k = 1; k < n; k++) { o->lo_more[k - 1] = lo[k]; o->span_more[k - 1] = (u64)hi[k] - (u64)lo[k]; }
    }
    return true;
  };
  d.npreds = desc->npreds;
  for (int q = 0; q < desc->npreds; q++)
    if (!to_pred(desc->pred[q], &d.pred[q])) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: bad predicate %d", q));
  if (desc->nindicators < 0 || desc->nindicators > VDL_MAX_INDICATORS) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: %d indicators", desc->nindicators));
  d.nind = desc->nindicators;
  for (int q = 0; q < desc->nindicators; q++)
    if (!to_pred(desc->indicator[q], &d.ind[q])) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: bad indicator %d", q));
  d.nkeys = desc->nkeys;
  for (int q = 0; q < desc->nkeys; q++) {
    if (!term_ok(desc->key[q], desc->nleaves, desc->nindicators) || desc->key_shl[q] < 0 || desc->key_shl[q] > 63) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: bad key part %d", q));
    d.key[q] = to_p(desc->key[q]);
    d.key_shl[q] = desc->key_shl[q];
  }
  auto prod = [&](const vdl_product &s, PProd *o) {
    if (s.nfactors < 0 || s.nfactors > VDL_MAX_FACTORS) return false;
    o->nfac = s.nfactors;
    for (int t = 0; t < s.nfactors; t++) { if (!term_ok(s.factor[t], desc->nleaves, desc->nindicators)) return false; o->f[t] = to_p(s.factor[t]); }
    return true;
  };
  p->folding = desc->nfolds > 0;
  d.nfolds = desc->nfolds;
  d.nemits = desc->nemits;
  d.ntiles = (desc->rows + P_TILE - 1) / P_TILE;
  d.errflag = ctx->d_errflag;
  if (cudaMalloc(&p->d_ticket, sizeof(unsigned int)) != cudaSuccess || cudaMalloc(&p->d_total, 2 * sizeof(i64)) != cudaSuccess)
    return fail(vdl_fail(ctx, VDL_ENOMEM, "probe: counters"));
  d.ticket = p->d_ticket;
  d.total = p->d_total;
  if (p->folding) {
    if (desc->nkeys == 0 ? desc->domain != 1 : (desc->domain < 1 || desc->domain > (1 << 22)))
      return fail(vdl_fail(ctx, VDL_EUNSUPPORTED, "probe: key domain %lld not supported", (long long)desc->domain));
    d.key_mask = desc->key_mask;
    d.domain = desc->domain;
    for (int j = 0; j < desc->nfolds; j++) {
      int op = desc->fold[j].op;
      if (op < VDL_FOLD_SUM || op > VDL_FOLD_COUNT || !prod(desc->fold[j].value, &d.fold[j])) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: bad fold %d", j));
      d.fold_op[j] = op;
    }
    int rc = vec_new(ctx, VDL_I64, (i64)(d.nfolds + 2) * d.domain, &p->table);
    if (rc) return fail(rc);
    d.table = (i64 *)ctx->vecs[p->table].ptr;
    size_t tb = (size_t)(d.nfolds + 2) * d.domain * 8;
    d.smem_table = tb <= P_SMEM_TABLE_BYTES;
    p->smem = d.smem_table ? tb : 0;
    for (int q = 0; q < desc->nposts; q++) {
      const vdl_post_op &P = desc->post[q];
      auto okk = [&](int kind, i64 v) { return kind == VDL_POST_CONST || (kind == VDL_POST_FOLD && v >= 0 && v < desc->nfolds) || (kind == VDL_POST_POST && v >= 0 && v < q); };
      if (P.op < 0 || P.op > VDL_MODULO || !okk(P.a_kind, P.a) || !okk(P.b_kind, P.b)) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: bad post op %d", q));
      p->pf.post[q] = P;
    }
    size_t nb = ((size_t)(d.nfolds + desc->nposts) * d.domain + 3) * sizeof(i64);
    if (cudaMalloc(&p->d_out, nb) != cudaSuccess || cudaHostAlloc(&p->h_out, nb, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(&p->h_mapped, p->h_out, 0) != cudaSuccess)
      return fail(vdl_fail(ctx, VDL_ENOMEM, "probe: result buffers"));
    memset(p->h_out, 0, nb);
    p->pf.domain = d.domain; p->pf.nfolds = d.nfolds; p->pf.npost = desc->nposts;
    for (int j = 0; j < d.nfolds; j++) p->pf.fold_op[j] = d.fold_op[j];
    p->pf.table = d.table; p->pf.nranks = 1; p->pf.stride = (i64)(d.nfolds + 2) * d.domain; p->pf.out = p->d_out; p->pf.hmirror = p->h_mapped; p->pf.errflag = ctx->d_errflag;
  } else {
    for (int e = 0; e < desc->nemits; e++)
      if (!prod(desc->emit[e], &d.emit[e])) return fail(vdl_fail(ctx, VDL_EINVAL, "probe: bad emit %d", e));
    if (cudaMalloc(&p->d_state, (size_t)std::max<i64>(1, d.ntiles) * 8) != cudaSuccess) return fail(vdl_fail(ctx, VDL_ENOMEM, "probe: tile states"));
    d.tile_state = p->d_state;
    if (cudaHostAlloc(&p->h_out, 2 * sizeof(i64), cudaHostAllocDefault) != cudaSuccess) return fail(vdl_fail(ctx, VDL_ENOMEM, "probe: host counter"));
  }
  // prefetch list: fact-table leaves (no parent) a predicate stage starts from.  Distance 0 = this tile's own columns,
  // pulled in while its first stage runs (measured best: Q12 -9 %, Q19 -10 %, Q3 -8 %, Q5 -4 % kernel time; a few tiles
  // ahead is as good, 256+ tiles ahead thrashes L2).  VDL_PROBE_PREFETCH=off disables it, =N looks N tiles ahead.
  d.npf = 0;
  d.pf_dist = 0;
  const char *pfenv = getenv("VDL_PROBE_PREFETCH");
  if (pfenv && *pfenv >= '0' && *pfenv <= '9') d.pf_dist = atoi(pfenv);
  if (!(pfenv && !strcmp(pfenv, "off"))) {
    std::vector<char> want(desc->nleaves, 0);
    auto root = [&](int l) { while (l >= 0 && desc->leaf[l].parent >= 0) l = desc->leaf[l].parent; return l; };
    for (int q = 0; q < desc->npreds; q++)
      for (int l : {desc->pred[q].t.leaf, desc->pred[q].kind == 1 ? desc->pred[q].u.leaf : -1}) { int r = root(l); if (r >= 0) want[r] = 1; }
    for (int l = 0; l < desc->nleaves; l++)
      if (want[l] && d.leaf[l].ptr) { d.pf_ptr[d.npf] = (const unsigned char *)d.leaf[l].ptr; d.pf_shift[d.npf] = d.leaf[l].w4 ? 2 : 3; d.npf++; }
  }
  for (int i = 0; i < VDL_EVENT_RING; i++) { cudaEventCreate(&p->ev0[i]); cudaEventCreate(&p->ev1[i]); }
  int per_sm = P_BLOCKS;
  p->grid = (int)std::max<i64>(1, std::min<i64>((i64)ctx->sm_count * per_sm, d.ntiles));
  if (p->smem > 48 * 1024) cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
  {   // load every kernel a step may launch NOW (CUDA loads lazily, and loading can wait for running kernels: a first
      // launch of the finalize kernel behind a peer's spinning exchange kernel would stall the host until the exchange
      // times out -- seen when one host thread drives several ranks)
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, probe_kernel); cudaFuncGetAttributes(&fa, probe_init_kernel); cudaFuncGetAttributes(&fa, probe_choose_kernel);
    cudaFuncGetAttributes(&fa, probe_exchange_kernel); cudaFuncGetAttributes(&fa, probe_finalize_kernel);
    cudaGetLastError();
  }
  {   // run-time specialisation (VDL_PROBE_JIT=0 switches it off; VDL_PROBE_JIT_MIN_ROWS, default 262144)
    const char *e = getenv("VDL_PROBE_JIT");
    i64 min_rows = 1 << 18;
    if (const char *m = getenv("VDL_PROBE_JIT_MIN_ROWS")) min_rows = atoll(m);
    bool lens_ok = true;             // the generated code clamps a failed lookup to index 0: every looked-up column needs a row
    for (int l = 0; l < d.nleaves; l++) if (d.leaf[l].parent >= 0 && d.leaf[l].len < 1) lens_ok = false;
    if (!(e && !strcmp(e, "0")) && d.rows >= min_rows && lens_ok) p->jit_kernel = probe_jit_compile(ctx, d, p->smem, nullptr, nullptr);
  }
  *out = p;
  return VDL_OK;
}

extern "C" int vdl_probe_run(vdl_probe *p) { return vdl_probe_run_ex(p, 1); }

extern "C" int vdl_probe_exchange_bytes(vdl_probe *p, int world, int64_t *bytes) {
  if (!p || !bytes || !p->folding || world < 1 || world > VDL_MAX_RANKS) return VDL_EINVAL;
  const int64_t stride = (int64_t)(p->pd.nfolds + 2) * p->pd.domain;
  *bytes = ((int64_t)2 * world * stride + 2 * world) * (int64_t)sizeof(i64);
  return VDL_OK;
}

extern "C" int vdl_probe_set_peers(vdl_probe *p, int rank, int world, void *const *peer_buffers) {
  if (!p || !p->folding || !peer_buffers || world < 1 || world > VDL_MAX_RANKS || rank < 0 || rank >= world) return VDL_EINVAL;
  memset(&p->xd, 0, sizeof p->xd);
  p->xd.rank = rank;
  p->xd.world = world;
  p->xd.stride = (i64)(p->pd.nfolds + 2) * p->pd.domain;
  p->xd.timeout_ns = 10000000000ull;                     // 10 s; VDL_PEER_TIMEOUT_MS overrides (tests)
  if (const char *e = getenv("VDL_PEER_TIMEOUT_MS")) p->xd.timeout_ns = (u64)atoll(e) * 1000000ull;
  for (int r = 0; r < world; r++) {
    if (!peer_buffers[r]) return vdl_fail(p->ctx, VDL_EINVAL, "set_peers: buffer of rank %d is null", r);
    p->xd.peer[r] = (i64 *)peer_buffers[r];
  }
  return VDL_OK;
}

// step counter of the peer exchange; the plan carries it over when a probe is re-prepared on the same buffers
u64 vdl_probe_epoch(vdl_probe *p) { return p->xd.epoch; }
void vdl_probe_set_epoch(vdl_probe *p, u64 e) { p->xd.epoch = e; }

extern "C" int vdl_probe_partials(vdl_probe *p, void **device_ptr, int64_t *n_int64) {
  if (!p || !device_ptr || !n_int64 || !p->folding) return VDL_EINVAL;
  *device_ptr = p->pd.table;
  *n_int64 = (int64_t)(p->pd.nfolds + 2) * p->pd.domain;
  return VDL_OK;
}

// Merge `nranks` partial tables laid out back to back (an all-gather result; NULL and 1: this rank's own) and finalize.
extern "C" int vdl_probe_finalize(vdl_probe *p, const void *all_partials, int nranks) {
  if (!p || !p->folding) return VDL_EINVAL;
  vdl_ctx *ctx = p->ctx;
  if (nranks < 1 || (nranks > 1 && !all_partials)) return vdl_fail(ctx, VDL_EINVAL, "probe finalize: nranks %d without gathered partials", nranks);
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  p->pf.table = all_partials ? (const i64 *)all_partials : p->pd.table;
  p->pf.nranks = nranks;
  p->pf.stride = (i64)(p->pd.nfolds + 2) * p->pd.domain;
  p->pf.precomputed_choose = 1;
  p->pf.seq++;
  probe_finalize_kernel<<<1, 256, 0, ctx->stream>>>(p->pd, p->pf);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  p->fetched = false;
  p->finalized = true;
  return VDL_OK;
}

// finalize != 0: single GPU, results on their way when this returns.  0 (fold mode): leave the complete partial
// table (FoldChoose values included) for an external combine, then vdl_probe_finalize().
extern "C" int vdl_probe_run_ex(vdl_probe *p, int finalize) {
  if (!p) return VDL_EINVAL;
  vdl_ctx *ctx = p->ctx;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!vdl_probe_current(p))
    return vdl_fail(ctx, VDL_ESTALE, "probe: a column was rewritten or dropped after vdl_probe_prepare: prepare again");
  PDesc &d = p->pd;
  p->fetched = false;
  p->ngroups = p->nselected = -1;
  VDL_CUDA(ctx, cudaMemsetAsync(p->d_ticket, 0, sizeof(unsigned int), ctx->stream));
  if (p->folding) {
    int nb = (int)std::min<i64>(ctx->sm_count, ((i64)(d.nfolds + 2) * d.domain + 255) / 256);
    probe_init_kernel<<<std::max(nb, 1), 256, 0, ctx->stream>>>(d);
    ctx->launches++;
  } else {
    VDL_CUDA(ctx, cudaMemsetAsync(p->d_state, 0, (size_t)std::max<i64>(1, d.ntiles) * 8, ctx->stream));
    VDL_CUDA(ctx, cudaMemsetAsync(p->d_total, 0, 2 * sizeof(i64), ctx->stream));
    // emit capacity: one vector of `rows` int64 per emitted expression (the survivors are usually far fewer; the
    // pool keeps the pages), released by the caller through the vector handles
    for (int e = 0; e < d.nemits; e++) {
      if (p->emit_vec[e]) { vdl_vec_free(ctx, p->emit_vec[e]); p->emit_vec[e] = 0; }
      VDL_TRY(vec_new(ctx, VDL_I64, d.rows, &p->emit_vec[e]));
      d.emit_out[e] = (i64 *)ctx->vecs[p->emit_vec[e]].ptr;
    }
  }
  VDL_CUDA(ctx, cudaEventRecord(p->ev0[p->nlaunch % VDL_EVENT_RING], ctx->stream));
  if (d.rows > 0) {
    if (p->jit_kernel) {
      void *args[] = {(void *)&d};
      VDL_CUDA(ctx, cudaLaunchKernel((const void *)p->jit_kernel, dim3(p->grid), dim3(P_THREADS), args, p->smem, ctx->stream));
    } else {
      probe_kernel<<<p->grid, P_THREADS, p->smem, ctx->stream>>>(d);
    }
    ctx->launches++;
  }
  VDL_CUDA(ctx, cudaEventRecord(p->ev1[p->nlaunch % VDL_EVENT_RING], ctx->stream));
  p->nlaunch++;
  p->finalized = !p->folding || finalize != 0;
  if (p->folding && finalize == 2) {       // peer-memory combine: every rank ends with the global result, no host round trip
    if (p->xd.world < 1) return vdl_fail(ctx, VDL_EINVAL, "probe: peer exchange requested before vdl_probe_set_peers");
    p->xd.epoch++;
    int cb = (int)std::min<i64>(ctx->sm_count, (d.domain + 255) / 256);
    probe_choose_kernel<<<std::max(cb, 1), 256, 0, ctx->stream>>>(d);
    probe_exchange_kernel<<<1, 256, 0, ctx->stream>>>(p->xd, d.table, ctx->d_errflag);
    p->pf.table = p->xd.peer[p->xd.rank] + (size_t)(p->xd.epoch & 1) * p->xd.world * p->xd.stride;
    p->pf.nranks = p->xd.world; p->pf.stride = p->xd.stride; p->pf.precomputed_choose = 1;
    p->pf.seq++;
    probe_finalize_kernel<<<1, 256, 0, ctx->stream>>>(d, p->pf);
    ctx->launches += 3;
  } else if (p->folding && finalize) {
    p->pf.table = d.table; p->pf.nranks = 1; p->pf.stride = (i64)(d.nfolds + 2) * d.domain; p->pf.precomputed_choose = 0;
    p->pf.seq++;
    probe_finalize_kernel<<<1, 256, 0, ctx->stream>>>(d, p->pf);
    ctx->launches++;
  } else if (p->folding) {
    int cb = (int)std::min<i64>(ctx->sm_count, (d.domain + 255) / 256);
    probe_choose_kernel<<<std::max(cb, 1), 256, 0, ctx->stream>>>(d);
    ctx->launches++;
  } else {
    VDL_CUDA(ctx, cudaMemcpyAsync(p->h_out, p->d_total, sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
  }
  VDL_CUDA(ctx, cudaGetLastError());
  p->ran = true;
  return VDL_OK;
}

static int probe_fetch(vdl_probe *p) {
  vdl_ctx *ctx = p->ctx;
  if (!p->ran) return vdl_fail(ctx, VDL_EINVAL, "probe has not run");
  if (!p->finalized) return vdl_fail(ctx, VDL_EINVAL, "probe ran for an external combine: call vdl_probe_finalize first");
  if (p->fetched) return VDL_OK;
  if (p->folding) VDL_TRY(wait_published(ctx, p->h_out + (size_t)(p->pd.nfolds + p->pf.npost) * p->pd.domain + 2, p->pf.seq));
  else VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (p->folding) {
    size_t tail = (size_t)(p->pd.nfolds + p->pf.npost) * p->pd.domain;
    i64 err = p->h_out[tail + 1];
    if (err) {
      cudaMemsetAsync(ctx->d_errflag, 0, sizeof(int), ctx->stream);
      if (err >= (1 << 20)) return vdl_fail(ctx, VDL_ECUDA, "probe: a peer GPU never delivered its partial table (exchange timed out)");
      return vdl_fail(ctx, VDL_ERANGE, "probe: %lld rows had a lookup index or group key out of range", (long long)err);
    }
    p->ngroups = p->h_out[tail];
  } else {
    p->nselected = p->h_out[0];
    for (int e = 0; e < p->pd.nemits; e++) {
      Vec &v = ctx->vecs[p->emit_vec[e]];
      v.len = p->nselected;
    }
  }
  p->fetched = true;
  return VDL_OK;
}

extern "C" int vdl_probe_result_host(vdl_probe *p, int index, const int64_t **data, int64_t *len) {
  if (!p || !data || !len || !p->folding || index < 0 || index >= p->pd.nfolds + p->pf.npost) return VDL_EINVAL;
  VDL_TRY(probe_fetch(p));
  *data = p->h_out + (size_t)index * p->pd.domain;
  *len = p->ngroups;
  return VDL_OK;
}

// Emit mode: the k-th emitted vector (length = number of surviving rows).  Ownership passes to the caller.
extern "C" int vdl_probe_emit_take(vdl_probe *p, int k, vdl_vec *out) {
  if (!p || !out || p->folding || k < 0 || k >= p->pd.nemits) return VDL_EINVAL;
  VDL_TRY(probe_fetch(p));
  if (!p->emit_vec[k]) return vdl_fail(p->ctx, VDL_EINVAL, "probe: emitted vector %d already taken", k);
  *out = p->emit_vec[k];
  p->emit_vec[k] = 0;
  return VDL_OK;
}

extern "C" int vdl_probe_last_kernel_ms(vdl_probe *p, float *ms) {
  if (!p || !ms || !p->ran) return VDL_EINVAL;
  const int i = (int)((p->nlaunch - 1) % VDL_EVENT_RING);
  VDL_CUDA(p->ctx, cudaEventSynchronize(p->ev1[i]));
  VDL_CUDA(p->ctx, cudaEventElapsedTime(ms, p->ev0[i], p->ev1[i]));
  return VDL_OK;
}

extern "C" int vdl_probe_kernel_ms_stats(vdl_probe *p, int n, float *mean_ms, float *min_ms) {
  if (!p || n < 1 || !p->ran) return VDL_EINVAL;
  n = (int)std::min<long>(std::min<long>(n, VDL_EVENT_RING), p->nlaunch);
  double sum = 0;
  float mn = 1e30f;
  for (int k = 1; k <= n; k++) {
    const int i = (int)((p->nlaunch - k) % VDL_EVENT_RING);
    float ms = 0;
    VDL_CUDA(p->ctx, cudaEventSynchronize(p->ev1[i]));
    VDL_CUDA(p->ctx, cudaEventElapsedTime(&ms, p->ev0[i], p->ev1[i]));
    sum += ms;
    mn = std::min(mn, ms);
  }
  if (mean_ms) *mean_ms = (float)(sum / n);
  if (min_ms) *min_ms = mn;
  return VDL_OK;
}

extern "C" int vdl_probe_destroy(vdl_probe *p) {
  if (!p) return VDL_EINVAL;
  vdl_ctx *ctx = p->ctx;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (p->table) vdl_vec_free(ctx, p->table);
  for (int e = 0; e < VDL_MAX_EMITS; e++)
    if (p->emit_vec[e]) vdl_vec_free(ctx, p->emit_vec[e]);
  if (p->d_out) cudaFree(p->d_out);
  if (p->h_out) cudaFreeHost(p->h_out);
  if (p->d_total) cudaFree(p->d_total);
  if (p->d_ticket) cudaFree(p->d_ticket);
  if (p->d_state) cudaFree(p->d_state);
  for (int i = 0; i < VDL_EVENT_RING; i++) { if (p->ev0[i]) cudaEventDestroy(p->ev0[i]); if (p->ev1[i]) cudaEventDestroy(p->ev1[i]); }
  delete p;
  return VDL_OK;
}
