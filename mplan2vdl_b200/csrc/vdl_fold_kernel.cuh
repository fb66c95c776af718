// Fold by runs in ONE pass (included by vdl_ops.cu).
// FoldSum/Min/Max/Choose/Count (Vlite.hs:1048-1070, 1179; Vdl.hs:255-264): one output per run of equal consecutive
// `groups` values, in run order.
//
// A block of 256 threads takes 2048-row tiles in ticket order; every thread owns 8 consecutive rows, so a run that starts and ends inside a
// thread never leaves its registers.  The tile's summary (run heads, does it contain a head, aggregate of its trailing open
// run) is the element of a segmented scan; the exclusive prefix over the earlier tiles -- how many runs started before the
// tile, and what the run that reaches into it has accumulated so far -- comes from a decoupled look-back (two status records
// per tile: its own aggregate, then its inclusive prefix).  Every run is written exactly once, by the thread that holds its
// LAST row: no initialisation pass, no atomics, groups and data read once.
#pragma once

#define FL_R 8
#define FL_THREADS 256
#define FL_WARPS (FL_THREADS / 32)
#define FL_TILE (FL_THREADS * FL_R)
struct FoldTileState { unsigned long long a0, a1, p0, p1; };   // word 0: ready << 63 | has_head << 62 | heads; word 1: open-run aggregate
struct FoldPart { i64 agg; int cnt; int has; };

template <int OP>
__device__ __forceinline__ i64 fold_ident() { return OP == VDL_FOLD_MIN ? INT64_MAX : (OP == VDL_FOLD_MAX ? INT64_MIN : 0); }
template <int OP>
__device__ __forceinline__ i64 fold_comb(i64 a, i64 b) {
  if (OP == VDL_FOLD_MIN) return b < a ? b : a;
  if (OP == VDL_FOLD_MAX) return b > a ? b : a;
  return (i64)((u64)a + (u64)b);
}
// segmented-scan operator: x covers the rows before y's
template <int OP>
__device__ __forceinline__ FoldPart fold_join(const FoldPart &x, const FoldPart &y) {
  FoldPart r;
  r.cnt = x.cnt + y.cnt;
  r.has = x.has | y.has;
  r.agg = y.has ? y.agg : fold_comb<OP>(x.agg, y.agg);
  return r;
}

// A record is one aligned 16-byte word pair, stored and loaded by ONE 128-bit access (single transaction: status and
// aggregate can never be seen apart, so no fence separates them).
__device__ __forceinline__ void fold_publish(unsigned long long *rec, i64 cnt, bool has, i64 agg) {
  const unsigned long long w0 = (1ull << 63) | ((unsigned long long)has << 62) | (unsigned long long)cnt;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(rec), "l"(w0), "l"((unsigned long long)agg) : "memory");
}
__device__ __forceinline__ void fold_peek(const unsigned long long *rec, unsigned long long *w0, unsigned long long *w1) {
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(*w0), "=l"(*w1) : "l"(rec) : "memory");
}

template <int OP>
__global__ void __launch_bounds__(FL_THREADS) fold_lookback_kernel(Operand groups, Operand data, i64 n, i64 ntiles, FoldTileState *state,
                                                            unsigned int *ticket, i64 *__restrict__ out, i64 *total) {
  __shared__ i64 s_x[FL_WARPS][FL_R * 32 + 32];
  __shared__ i64 s_agg[FL_WARPS];
  __shared__ int s_cnt[FL_WARPS], s_has[FL_WARPS];
  __shared__ i64 s_lb_c[FL_WARPS], s_lb_a[FL_WARPS];
  __shared__ int s_lb_f[FL_WARPS];
  __shared__ unsigned int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    __syncthreads();                                  // the previous tile's shared state is no longer read
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const i64 tile = s_tile;
    if (tile >= ntiles) break;
    // coalesced loads (row = warp's first row + 32 k + lane), then a warp-private transpose through shared memory so that
    // every thread ends up with its 8 CONSECUTIVE rows (slot r + r / 8: the blocked reads are conflict-free)
    const i64 i0 = tile * FL_TILE + (i64)tid * FL_R;
    i64 g[FL_R], v[FL_R];
    {
      const i64 wbase = tile * FL_TILE + (i64)warp * (FL_R * 32);
      i64 *xw = s_x[warp];
#pragma unroll
      for (int k = 0; k < FL_R; k++) {
        const i64 i = wbase + k * 32 + lane;
        g[k] = i < n ? op_ld(groups, i) : 0;
        v[k] = i < n ? (OP == VDL_FOLD_COUNT ? 1 : op_ld(data, i)) : fold_ident<OP>();
      }
#pragma unroll
      for (int k = 0; k < FL_R; k++) { const int r = k * 32 + lane; xw[r + (r >> 3)] = g[k]; }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < FL_R; k++) g[k] = xw[lane * (FL_R + 1) + k];
      __syncwarp();
      if (OP != VDL_FOLD_COUNT) {
#pragma unroll
        for (int k = 0; k < FL_R; k++) { const int r = k * 32 + lane; xw[r + (r >> 3)] = v[k]; }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < FL_R; k++) v[k] = xw[lane * (FL_R + 1) + k];
        __syncwarp();
      } else {
#pragma unroll
        for (int k = 0; k < FL_R; k++) v[k] = i0 + k < n ? 1 : 0;
      }
    }
    i64 gprev = __shfl_up_sync(0xffffffffu, g[FL_R - 1], 1), gnext = __shfl_down_sync(0xffffffffu, g[0], 1);
    if (lane == 0 && i0 > 0 && i0 < n) gprev = op_ld(groups, i0 - 1);
    if (lane == 31 && i0 + FL_R < n) gnext = op_ld(groups, i0 + FL_R);
    unsigned heads = 0, ends = 0;
#pragma unroll
    for (int k = 0; k < FL_R; k++) {
      const i64 i = i0 + k;
      if (i < n && (i == 0 || g[k] != (k ? g[k - 1] : gprev))) heads |= 1u << k;
      if (i < n && (i + 1 >= n || g[k] != (k + 1 < FL_R ? g[k + 1] : gnext))) ends |= 1u << k;
    }
    // this thread's summary, then an inclusive segmented scan over the warp and the block
    FoldPart mine;
    mine.cnt = __popc(heads); mine.has = heads != 0; mine.agg = fold_ident<OP>();
    if (OP != VDL_FOLD_CHOOSE) {
#pragma unroll
      for (int k = 0; k < FL_R; k++) mine.agg = (heads >> k & 1) ? v[k] : fold_comb<OP>(mine.agg, v[k]);
    }
    FoldPart inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      FoldPart y;
      y.agg = __shfl_up_sync(0xffffffffu, inc.agg, o);
      y.cnt = __shfl_up_sync(0xffffffffu, inc.cnt, o);
      y.has = __shfl_up_sync(0xffffffffu, inc.has, o);
      if (lane >= o) inc = fold_join<OP>(y, inc);
    }
    if (lane == 31) { s_agg[warp] = inc.agg; s_cnt[warp] = inc.cnt; s_has[warp] = inc.has; }
    FoldPart exc;                                     // rows of this warp before this thread
    exc.agg = __shfl_up_sync(0xffffffffu, inc.agg, 1);
    exc.cnt = __shfl_up_sync(0xffffffffu, inc.cnt, 1);
    exc.has = __shfl_up_sync(0xffffffffu, inc.has, 1);
    if (lane == 0) { exc.agg = fold_ident<OP>(); exc.cnt = 0; exc.has = 0; }
    __syncthreads();
    FoldPart before;                                  // rows of this tile before this warp
    before.agg = fold_ident<OP>(); before.cnt = 0; before.has = 0;
    for (int w = 0; w < warp; w++) { FoldPart y; y.agg = s_agg[w]; y.cnt = s_cnt[w]; y.has = s_has[w]; before = fold_join<OP>(before, y); }
    // tile summary (every thread computes it), then the look-back: ALL threads poll, thread i the tile i + 1 places back, so
    // one round covers more tiles than are usually in flight (4 blocks per SM) -- at full bandwidth a hundred tiles start per
    // microsecond, and a 32-wide window would need several dependent round trips through L2 per tile
    FoldPart T;
    T.agg = fold_ident<OP>(); T.cnt = 0; T.has = 0;
    for (int w = 0; w < FL_WARPS; w++) { FoldPart y; y.agg = s_agg[w]; y.cnt = s_cnt[w]; y.has = s_has[w]; T = fold_join<OP>(T, y); }
    i64 base = 0, carry = fold_ident<OP>();
    FoldTileState *me = state + tile;
    if (tile > 0) {
      if (tid == 0) fold_publish(&me->a0, T.cnt, T.has, T.agg);
      bool closed = false, found = false;
      for (i64 j0 = tile - 1; !found; j0 -= FL_THREADS) {
        const i64 j = j0 - tid;
        unsigned long long w0 = (1ull << 63) | (1ull << 62), w1 = (unsigned long long)fold_ident<OP>();
        bool prefix = true;                          // before the first tile: an empty, closed prefix
        if (j >= 0) {
          const FoldTileState *st = state + j;
          for (;;) {                                 // both records in flight at once: the prefix wins when it is there
            unsigned long long a0, a1;
            fold_peek(&st->p0, &w0, &w1);
            fold_peek(&st->a0, &a0, &a1);
            if (w0 >> 63) { prefix = true; break; }
            if (a0 >> 63) { prefix = false; w0 = a0; w1 = a1; break; }
            __nanosleep(40);
          }
        }
        const unsigned pm = __ballot_sync(0xffffffffu, prefix);
        const int plane = pm ? __ffs(pm) - 1 : 32;                   // nearest inclusive prefix: nothing beyond it counts
        const bool counted = lane <= plane;
        const unsigned hm = __ballot_sync(0xffffffffu, counted && ((w0 >> 62) & 1));
        const int hlane = hm ? __ffs(hm) - 1 : 32;                   // nearest tile with a head: the open run starts there
        i64 c = counted ? (i64)(w0 & ((1ull << 62) - 1)) : 0;
        i64 a = (counted && lane <= hlane) ? (i64)w1 : fold_ident<OP>();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          c += __shfl_xor_sync(0xffffffffu, c, o);
          a = fold_comb<OP>(a, __shfl_xor_sync(0xffffffffu, a, o));
        }
        __syncthreads();                             // s_lb_* of the previous round (and s_agg above) are no longer read
        if (lane == 0) { s_lb_c[warp] = c; s_lb_a[warp] = a; s_lb_f[warp] = (pm ? 1 : 0) | (hm ? 2 : 0); }
        __syncthreads();
        for (int w = 0; w < FL_WARPS && !found; w++) {               // warp w polled the tiles 32 w .. 32 w + 31 places back
          base += s_lb_c[w];
          if (!closed) carry = fold_comb<OP>(s_lb_a[w], carry);
          closed = closed || (s_lb_f[w] & 2);
          found = s_lb_f[w] & 1;
        }
      }
    }
    // inclusive prefix of this tile: heads so far; its open run continues the carry unless the tile has a head of its own
    if (tid == 0) {
      fold_publish(&me->p0, base + T.cnt, true, T.has ? T.agg : fold_comb<OP>(carry, T.agg));
      if (tile == ntiles - 1) *total = base + T.cnt;
    }
    const FoldPart pre = fold_join<OP>(before, exc);  // rows of the tile before this thread
    i64 acc = pre.has ? pre.agg : fold_comb<OP>(carry, pre.agg);
    i64 rid = base + pre.cnt - 1;
#pragma unroll
    for (int k = 0; k < FL_R; k++) {
      const bool h = heads >> k & 1;
      rid += h;
      if (OP == VDL_FOLD_CHOOSE) { if (h) out[rid] = v[k]; }
      else {
        acc = h ? v[k] : fold_comb<OP>(acc, v[k]);
        if (ends >> k & 1) out[rid] = acc;
      }
    }
  }
}

// Constant groups (Vlite.hs:636-638: an empty group-by folds over zeros_) are ONE run: a plain grid reduction, 8 loads in
// flight per thread, one atomic per block.
template <int OP>
__global__ void __launch_bounds__(256) fold_all_kernel(Operand data, i64 n, i64 *__restrict__ out) {
  __shared__ i64 wred[8];
  const i64 stride = (i64)gridDim.x * blockDim.x;
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  i64 v = fold_ident<OP>();
  for (; i + 7 * stride < n; i += 8 * stride) {
    i64 x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = op_ld(data, i + k * stride);
#pragma unroll
    for (int k = 0; k < 8; k++) v = fold_comb<OP>(v, x[k]);
  }
  for (; i < n; i += stride) v = fold_comb<OP>(v, op_ld(data, i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fold_comb<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) v = fold_comb<OP>(v, wred[w]);
    if (OP == VDL_FOLD_MIN) atomicMin((long long *)out, (long long)v);
    else if (OP == VDL_FOLD_MAX) atomicMax((long long *)out, (long long)v);
    else atomicAdd((unsigned long long *)out, (unsigned long long)v);
  }
}
__global__ void fold_all_init_kernel(int op, Operand data, i64 n, i64 *out) {
  *out = op == VDL_FOLD_MIN ? INT64_MAX : (op == VDL_FOLD_MAX ? INT64_MIN : (op == VDL_FOLD_CHOOSE ? op_ld(data, 0) : (op == VDL_FOLD_COUNT ? n : 0)));
}
