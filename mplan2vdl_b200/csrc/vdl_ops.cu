// Per-op kernels: one entry point per Voodoo op (reference op set Vdl.hs:32-44, 110-131), so that any
// emitted graph runs op-at-a-time when the fusion pass does not apply, and for the per-op sweep.
// Semantics: SURVEY.md section 2.3 / Appendix G (dense vectors, int64 wraparound).
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "vdl_internal.h"

__device__ __forceinline__ i64 op_ld(const Operand &o, i64 i) {
  if (o.kind == 0) return ((const i64 *)o.p)[i];
  if (o.kind == 1) return (i64)((const int32_t *)o.p)[i];
  return (i64)((u64)o.from + (u64)i * (u64)o.step);
}

// ---------------------------------------------------------------------------------- RangeV / RangeC
extern "C" int vdl_op_range(vdl_ctx *ctx, int64_t from, int64_t step, int64_t len, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  return vec_new_range(ctx, from, step, len, out);   // virtual: consumers read from + i*step
}

// ---------------------------------------------------------------------------------- elementwise
// op semantics: binop_apply in vdl_internal.h
template <int OP>
__global__ void __launch_bounds__(256) binary_kernel(Operand a, Operand b, i64 *__restrict__ out, i64 n) {
  // two elements per thread and step: 16-byte stores, 2 x (8 or 4)-byte loads per operand in flight
  i64 stride = (i64)gridDim.x * blockDim.x * 2;
  for (i64 i = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
    if (i + 1 < n) {
      longlong2 r;
      r.x = binop_apply(OP, op_ld(a, i), op_ld(b, i));
      r.y = binop_apply(OP, op_ld(a, i + 1), op_ld(b, i + 1));
      *(longlong2 *)(out + i) = r;
    } else {
      out[i] = binop_apply(OP, op_ld(a, i), op_ld(b, i));
    }
  }
}

typedef void (*binary_fn)(Operand, Operand, i64 *, i64);
static binary_fn binary_table[12] = {
    binary_kernel<0>, binary_kernel<1>, binary_kernel<2>, binary_kernel<3>, binary_kernel<4>,  binary_kernel<5>,
    binary_kernel<6>, binary_kernel<7>, binary_kernel<8>, binary_kernel<9>, binary_kernel<10>, binary_kernel<11>};

extern "C" int vdl_op_binary(vdl_ctx *ctx, int op, vdl_vec a, vdl_vec b, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  if (op < 0 || op > VDL_MODULO) return vdl_fail(ctx, VDL_EINVAL, "binary op %d unknown", op);
  Vec *va = vec_get(ctx, a), *vb = vec_get(ctx, b);
  if (!va || !vb) return VDL_EINVAL;
  if (va->len != vb->len) return vdl_fail(ctx, VDL_EINVAL, "elementwise op on lengths %lld vs %lld", (long long)va->len, (long long)vb->len);
  i64 n = va->len;
  Operand oa = operand_of(*va), ob = operand_of(*vb);
  // positions `p % k` with a constant k > 0 -- the scatter size hint of addScatterSizeHint (Vlite.hs:1117-1120) -- index
  // a space of k slots (App. G2 / G9)
  i64 dom = -1;
  if (op == VDL_MODULO && vb->is_range && vb->step == 0 && vb->from > 0) dom = vb->from;
  VDL_TRY(vec_new(ctx, VDL_I64, n, out));
  ctx->vecs[*out].domain = dom;
  if (n == 0) return VDL_OK;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  i64 want = (n / 2 + 255) / 256;
  int blocks = (int)std::max<i64>(1, std::min<i64>(want, (i64)ctx->sm_count * 16));
  binary_table[op]<<<blocks, 256, 0, ctx->stream>>>(oa, ob, (i64 *)ctx->vecs[*out].ptr, n);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- Like over a string heap
// Vlite.hs:1010-1014 -> Vdl.hs:244-247: the data vector holds byte offsets into the column's string heap.  One thread per
// row reads its (NUL-terminated) string and runs the glob match with one backtrack point (the last `%`): sequential byte
// reads of a string that is 8-byte aligned and a few tens of bytes long, so a row costs its 8-byte offset plus one or two
// 32-byte sectors of the heap (L2-resident for dictionary-like columns such as p_type).  Bound: HBM / L2 bandwidth.
struct LikeArgs { char pat[VDL_LIKE_MAX_PATTERN]; };
__global__ void __launch_bounds__(256) like_kernel(Operand data, const unsigned char *__restrict__ heap, i64 heap_len, const __grid_constant__ LikeArgs a,
                                                   i64 *__restrict__ out, i64 n, int *errflag) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    const i64 off = op_ld(data, i);
    if ((u64)off >= (u64)heap_len) { atomicAdd(errflag, 1); out[i] = 0; continue; }
    const unsigned char *s = heap + off, *end = heap + heap_len, *star_s = nullptr;
    int p = 0, star_p = -1;
    bool ok = true;
    while (s < end && *s) {
      const char c = a.pat[p];
      if (c == '%') { star_p = ++p; star_s = s; }
      else if (c && (c == '_' || (unsigned char)c == *s)) { p++; s++; }
      else if (star_p >= 0) { p = star_p; s = ++star_s; }
      else { ok = false; break; }
    }
    if (ok) {
      while (a.pat[p] == '%') p++;
      ok = a.pat[p] == 0;
    }
    out[i] = ok ? 1 : 0;
  }
}

extern "C" int vdl_op_like(vdl_ctx *ctx, vdl_vec data, vdl_vec heap, const char *pattern, vdl_vec *out) {
  if (!ctx || !out || !pattern) return VDL_EINVAL;
  Vec *vd = vec_get(ctx, data), *vh = vec_get_any(ctx, heap);
  if (!vd || !vh) return VDL_EINVAL;
  if (vh->dtype != VDL_U8) return vdl_fail(ctx, VDL_EINVAL, "Like: the dictionary must be a string heap (a VDL_U8 vector; Load,<table>.<col>.heap)");
  if (strlen(pattern) >= VDL_LIKE_MAX_PATTERN) return vdl_fail(ctx, VDL_EUNSUPPORTED, "Like: pattern longer than %d bytes", VDL_LIKE_MAX_PATTERN - 1);
  const i64 n = vd->len;
  const Operand od = operand_of(*vd);
  const unsigned char *hp = (const unsigned char *)vh->ptr;
  const i64 hl = vh->len;
  VDL_TRY(vec_new(ctx, VDL_I64, n, out));
  if (n == 0) return VDL_OK;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  LikeArgs a;
  memset(&a, 0, sizeof a);
  strcpy(a.pat, pattern);
  int blocks = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 16));
  like_kernel<<<blocks, 256, 0, ctx->stream>>>(od, hp, hl, a, (i64 *)ctx->vecs[*out].ptr, n, ctx->d_errflag);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- map (expression tree in one launch)
// The op-at-a-time remainder of a plan is mostly long chains of elementwise ops over short vectors (Q19's OR of ANDs over
// the join's survivors: ~150 launches of a few microseconds each).  One launch interprets the whole tree per row: the
// program is the same for every thread (no divergence), the register file lives in local memory (L1), and each input is
// read from HBM once instead of once per consumer.
#define VDL_MAP_JIT_MIN_ROWS (1 << 16)
__global__ void __launch_bounds__(256) map_kernel(const __grid_constant__ MapArgs m, i64 *__restrict__ out, i64 n, int *errflag) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const int nt = m.d.ninstrs;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    i64 r[VDL_MAP_MAX_REGS];
    int last = 0;
    for (int t = 0; t < nt; t++) {
      const vdl_map_instr ins = m.d.instr[t];
      i64 v;
      if (ins.op == VDL_MAP_LOAD) {
        v = op_ld(m.in[ins.b], i);
      } else if (ins.op == VDL_MAP_RANGE) {
        v = (i64)((u64)m.d.imm[ins.a] + (u64)i * (u64)m.d.imm[ins.b]);
      } else if (ins.op == VDL_MAP_GATHER) {
        const i64 a = r[ins.a];
        if ((u64)a >= (u64)m.tab_len[ins.b]) { atomicAdd(errflag, 1); v = 0; }
        else v = op_ld(m.tab[ins.b], a);
      } else {
        v = binop_apply(ins.op, r[ins.a], r[ins.b]);
      }
      r[ins.dst] = v;
      last = ins.dst;
    }
    out[i] = r[last];
  }
}

extern "C" int vdl_abi_sizeof_map_desc(void) { return (int)sizeof(vdl_map_desc); }

extern "C" int vdl_op_map(vdl_ctx *ctx, const vdl_map_desc *d, const vdl_vec *inputs, const vdl_vec *tables, vdl_vec *out) {
  if (!ctx || !d || !out || !inputs || (d->ntables > 0 && !tables)) return VDL_EINVAL;
  if (d->ninputs < 1 || d->ninputs > VDL_MAP_MAX_INPUTS || d->ntables < 0 || d->ntables > VDL_MAP_MAX_TABLES || d->ninstrs < 1 ||
      d->ninstrs > VDL_MAP_MAX_INSTRS || d->nimms < 0 || d->nimms > VDL_MAP_MAX_IMMS)
    return vdl_fail(ctx, VDL_EINVAL, "map: %d inputs, %d tables, %d instructions, %d immediates out of range", d->ninputs, d->ntables, d->ninstrs, d->nimms);
  MapArgs *mp = new MapArgs();          // ~3 KB: off the stack
  std::unique_ptr<MapArgs> hold(mp);
  MapArgs &m = *mp;
  m.d = *d;
  i64 n = -1;
  for (int k = 0; k < d->ninputs; k++) {
    Vec *v = vec_get(ctx, inputs[k]);
    if (!v) return VDL_EINVAL;
    if (n >= 0 && v->len != n) return vdl_fail(ctx, VDL_EINVAL, "elementwise op on lengths %lld vs %lld", (long long)n, (long long)v->len);
    n = v->len;
    m.in[k] = operand_of(*v);
  }
  for (int k = 0; k < d->ntables; k++) {
    Vec *v = vec_get(ctx, tables[k]);
    if (!v) return VDL_EINVAL;
    m.tab[k] = operand_of(*v);
    m.tab_len[k] = v->len;
  }
  i64 domain = -1;
  i64 regconst[VDL_MAP_MAX_REGS];      // constant value of a register, when isconst
  bool isconst[VDL_MAP_MAX_REGS] = {false};
  unsigned written = 0;
  auto reg_ok = [&](int r, bool read) { return r >= 0 && r < VDL_MAP_MAX_REGS && (!read || (written >> r & 1)); };
  for (int t = 0; t < d->ninstrs; t++) {
    const vdl_map_instr &ins = d->instr[t];
    bool ok = reg_ok(ins.dst, false);
    if (ins.op == VDL_MAP_LOAD) ok = ok && ins.b >= 0 && ins.b < d->ninputs;
    else if (ins.op == VDL_MAP_RANGE) ok = ok && ins.a >= 0 && ins.a < d->nimms && ins.b >= 0 && ins.b < d->nimms;
    else if (ins.op == VDL_MAP_GATHER) ok = ok && reg_ok(ins.a, true) && ins.b >= 0 && ins.b < d->ntables;
    else ok = ok && ins.op >= 0 && ins.op <= VDL_MODULO && reg_ok(ins.a, true) && reg_ok(ins.b, true);
    if (!ok) return vdl_fail(ctx, VDL_EINVAL, "map: instruction %d (op %d dst %d a %d b %d) is malformed or reads a register nothing wrote", t, ins.op, ins.dst, ins.a, ins.b);
    written |= 1u << ins.dst;
    // index space of the result (App. G2): gathering positions keeps their space; `p % k` is the scatter size hint
    i64 dm = -1, cv = 0;
    bool ic = false;
    if (ins.op == VDL_MAP_GATHER) dm = ctx->vecs[tables[ins.b]].domain;
    else if (ins.op == VDL_MAP_LOAD) dm = ctx->vecs[inputs[ins.b]].domain;
    else if (ins.op == VDL_MAP_RANGE) { ic = d->imm[ins.b] == 0; cv = d->imm[ins.a]; }
    else if (ins.op == VDL_MODULO && isconst[ins.b] && regconst[ins.b] > 0) dm = regconst[ins.b];
    regconst[ins.dst] = cv; isconst[ins.dst] = ic;
    if (t == d->ninstrs - 1) domain = dm;
  }
  VDL_TRY(vec_new(ctx, VDL_I64, n, out));
  ctx->vecs[*out].domain = domain;
  if (n == 0) return VDL_OK;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  int blocks = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 8));
  // long vectors: a kernel specialised to this program (vdl_jit.cu); short ones are launch-bound either way
  int jitted = n >= VDL_MAP_JIT_MIN_ROWS ? vdl_jit_map_launch(ctx, m, (i64 *)ctx->vecs[*out].ptr, n, blocks) : 0;
  if (jitted < 0) return VDL_ECUDA;
  if (!jitted) map_kernel<<<blocks, 256, 0, ctx->stream>>>(m, (i64 *)ctx->vecs[*out].ptr, n, ctx->d_errflag);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- scan helpers
// Exclusive scan of `n` int64 counters in place by ONE block (n is #blocks or 256 * #blocks: small next to
// the data), total written to *total.  Warp-shuffle scan, 1024 elements per step, running carry.
__global__ void __launch_bounds__(1024, 1) exclusive_scan_kernel(i64 *data, i64 n, i64 *total) {
  __shared__ i64 warp_sum[32];
  __shared__ i64 carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (i64 base = 0; base < n; base += 1024) {
    i64 i = base + tid;
    i64 v = i < n ? data[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      i64 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      i64 w = warp_sum[lane], s = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        i64 y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sum[lane] = s - w;   // exclusive prefix of the warp sums
    }
    __syncthreads();
    i64 c = carry;
    if (i < n) data[i] = c + warp_sum[warp] + x - v;
    __syncthreads();
    if (tid == 1023) carry = c + warp_sum[31] + x;
    __syncthreads();
  }
  if (tid == 0) *total = carry;
}

// Large inputs (radix histograms: 256 counters per 4096 rows; a billion-row FoldSelect: 244 K block counts): scan
// 4096-element chunks in parallel, scan the chunk totals with the one-block kernel, add the chunk offsets back.
#define SCAN_CHUNK 4096
__global__ void __launch_bounds__(1024) scan_chunks_kernel(i64 *data, i64 n, i64 *chunk_sum) {
  __shared__ i64 warp_sum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 base = (i64)blockIdx.x * SCAN_CHUNK + tid * 4;
  i64 v[4], t = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = base + k < n ? data[base + k] : 0; t += v[k]; }
  i64 x = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { i64 y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_sum[warp] = x;
  __syncthreads();
  if (warp == 0) {
    i64 w = warp_sum[lane], q = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { i64 y = __shfl_up_sync(0xffffffffu, q, o); if (lane >= o) q += y; }
    warp_sum[lane] = q - w;
    if (lane == 31) chunk_sum[blockIdx.x] = q;
  }
  __syncthreads();
  i64 run = warp_sum[warp] + x - t;
#pragma unroll
  for (int k = 0; k < 4; k++) { if (base + k < n) data[base + k] = run; run += v[k]; }
}
__global__ void __launch_bounds__(1024) add_chunk_offsets_kernel(i64 *data, i64 n, const i64 *chunk_off) {
  const i64 off = chunk_off[blockIdx.x], base = (i64)blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
#pragma unroll
  for (int k = 0; k < 4; k++) if (base + k < n) data[base + k] += off;
}
// elements of scratch a scan of n counters needs, counters included: [n counters][total][chunk sums][their total]
static size_t scan_elems(i64 n) { return (size_t)n + 1 + 2 * (size_t)((n + SCAN_CHUNK - 1) / SCAN_CHUNK) + 64; }
// exclusive scan of data[0..n) in place, total to data[n]; data must have scan_elems(n) elements
static void device_exclusive_scan(vdl_ctx *ctx, i64 *data, i64 n) {
  if (n <= 4 * SCAN_CHUNK) {
    exclusive_scan_kernel<<<1, 1024, 0, ctx->stream>>>(data, n, data + n);
    ctx->launches++;
    return;
  }
  const i64 nc = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  i64 *sums = data + n + 1;
  scan_chunks_kernel<<<(unsigned)nc, 1024, 0, ctx->stream>>>(data, n, sums);
  device_exclusive_scan(ctx, sums, nc);                        // recursion depth <= 2 for any realistic n
  add_chunk_offsets_kernel<<<(unsigned)nc, 1024, 0, ctx->stream>>>(data, n, sums);
  cudaMemcpyAsync(data + n, sums + nc, 8, cudaMemcpyDeviceToDevice, ctx->stream);
  ctx->launches += 2;
}

#define SEL_TILE 4096   // rows per block of the flag-count / compaction kernels (16 steps of 256)

// flag(i): FoldSelect -> pred[i] != 0 ; Fold head -> i == 0 || g[i] != g[i-1]
template <bool HEADS>
__device__ __forceinline__ bool flag_at(const Operand &o, i64 i) {
  if (!HEADS) return op_ld(o, i) != 0;
  return i == 0 || op_ld(o, i) != op_ld(o, i - 1);
}

template <bool HEADS>
__global__ void __launch_bounds__(256) flag_count_kernel(Operand o, i64 n, i64 *__restrict__ block_count) {
  __shared__ int wsum[8];
  i64 base = (i64)blockIdx.x * SEL_TILE;
  int c = 0;
  for (int s = 0; s < SEL_TILE / 256; s++) {
    i64 i = base + s * 256 + threadIdx.x;
    c += (i < n) && flag_at<HEADS>(o, i);
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) c += __shfl_xor_sync(0xffffffffu, c, k);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; w++) t += wsum[w];
    block_count[blockIdx.x] = t;
  }
}

// Rank of every flagged element inside its block step, by warp ballots (stream compaction without atomics).
struct StepRank {
  int rank;    // number of flagged elements before this one in the block so far (valid when flagged)
  int incl;    // flagged elements up to and including this one (valid for every element)
};
__device__ __forceinline__ StepRank step_rank(bool flag, int *wcnt /*[8]*/, int &run) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) wcnt[warp] = __popc(m);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    int c = wcnt[w];
    if (w < warp) before += c;
    total += c;
  }
  StepRank r;
  r.rank = run + before + __popc(m & ((1u << lane) - 1));
  r.incl = r.rank + (flag ? 1 : 0);
  run += total;
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------- FoldSelect
// Vlite.hs:721-730: idx = Fold FSel (pos_ p) p.  Dense model: global stable compaction of positions.
static int flag_scan(vdl_ctx *ctx, bool heads, const Operand &o, i64 n, i64 **block_off, i64 *total) {
  i64 nb = (n + SEL_TILE - 1) / SEL_TILE;
  VDL_TRY(scratch_reserve(ctx, scan_elems(nb) * 8));
  i64 *cnt = (i64 *)ctx->scratch;
  if (heads) flag_count_kernel<true><<<(unsigned)nb, 256, 0, ctx->stream>>>(o, n, cnt);
  else flag_count_kernel<false><<<(unsigned)nb, 256, 0, ctx->stream>>>(o, n, cnt);
  ctx->launches++;
  device_exclusive_scan(ctx, cnt, nb);
  VDL_TRY(read_scalar(ctx, cnt + nb, total, 8));
  *block_off = cnt;
  return VDL_OK;
}

// Single pass: a block takes 4096-element tiles in ticket order, reads its predicates ONCE (16 flags per thread kept
// as a bit mask), gets the tile's output offset by a decoupled look-back over one 64-bit status word per tile
// ((status << 62) | count; 1 = tile aggregate, 2 = inclusive prefix) and writes the positions in order.
__global__ void __launch_bounds__(256) select_lookback_kernel(Operand o, i64 n, i64 ntiles, unsigned long long *state, unsigned int *ticket,
                                                              i64 *__restrict__ out, i64 *total) {
  __shared__ int wcnt[8];
  __shared__ i64 s_off;
  __shared__ unsigned int s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const i64 tile = s_tile;
    if (tile >= ntiles) break;
    const i64 base = tile * SEL_TILE;
    unsigned flags = 0;
#pragma unroll
    for (int s = 0; s < SEL_TILE / 256; s++) {
      const i64 i = base + s * 256 + tid;
      if (i < n && op_ld(o, i) != 0) flags |= 1u << s;
    }
    int c = __popc(flags);
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) c += __shfl_xor_sync(0xffffffffu, c, k);
    if (lane == 0) wcnt[warp] = c;
    __syncthreads();
    if (tid == 0) {
      unsigned long long T = 0;
      for (int w = 0; w < 8; w++) T += (unsigned long long)wcnt[w];
      i64 excl = 0;
      if (tile == 0) atomicExch(state, (2ull << 62) | T);
      else {
        atomicExch(state + tile, (1ull << 62) | T);
        for (i64 j = tile - 1;; j--) {
          unsigned long long st;
          do { st = *((volatile unsigned long long *)(state + j)); } while ((st >> 62) == 0);
          excl += (i64)(st & ((1ull << 62) - 1));
          if ((st >> 62) == 2) break;
        }
        atomicExch(state + tile, (2ull << 62) | (unsigned long long)(excl + (i64)T));
      }
      s_off = excl;
      if (tile == ntiles - 1) *total = excl + (i64)T;
    }
    __syncthreads();
    i64 off = s_off;
    // order inside the tile: step-major, then thread: rank by ballots per step
#pragma unroll 1
    for (int s = 0; s < SEL_TILE / 256; s++) {
      const bool f = (flags >> s) & 1;
      const unsigned m = __ballot_sync(0xffffffffu, f);
      if (lane == 0) wcnt[warp] = __popc(m);
      __syncthreads();
      int before = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < 8; w++) { const int cw = wcnt[w]; if (w < warp) before += cw; tot += cw; }
      if (f) out[off + before + __popc(m & ((1u << lane) - 1))] = base + s * 256 + tid;
      off += tot;
      __syncthreads();
    }
  }
}

extern "C" int vdl_op_fold_select(vdl_ctx *ctx, vdl_vec pred, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  Vec *vp = vec_get(ctx, pred);
  if (!vp) return VDL_EINVAL;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  i64 n = vp->len;
  Operand o = operand_of(*vp);
  if (n == 0) {
    VDL_TRY(vec_new(ctx, VDL_I64, 0, out));
    ctx->vecs[*out].domain = 0;
    return VDL_OK;
  }
  const i64 ntiles = (n + SEL_TILE - 1) / SEL_TILE;
  VDL_TRY(scratch_reserve(ctx, (size_t)(ntiles + 4) * 8));
  unsigned long long *state = (unsigned long long *)ctx->scratch;          // [ntiles] status words, then total, ticket
  i64 *d_total = (i64 *)(state + ntiles);
  unsigned int *ticket = (unsigned int *)(state + ntiles + 1);
  VDL_CUDA(ctx, cudaMemsetAsync(state, 0, (size_t)(ntiles + 2) * 8, ctx->stream));
  VDL_TRY(vec_new(ctx, VDL_I64, n, out));                                   // capacity n; the length is known after the pass
  const int grid = (int)std::min<i64>(ntiles, (i64)ctx->sm_count * 8);
  select_lookback_kernel<<<grid, 256, 0, ctx->stream>>>(o, n, ntiles, state, ticket, (i64 *)ctx->vecs[*out].ptr, d_total);
  ctx->launches++;
  i64 total = 0;
  VDL_TRY(read_scalar(ctx, d_total, &total, 8));
  Vec &v = ctx->vecs[*out];
  v.len = total;
  v.domain = n;   // these positions index the predicate's row space (App. G2)
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- Gather / Scatter
// Gather (Vlite.hs:86-87 `@@`, FK fetch 1264, 1276-1277): out[i] = src[pos[i]].
__global__ void __launch_bounds__(256) gather_kernel(Operand src, i64 src_len, Operand pos, i64 n, i64 *__restrict__ out, int *errflag) {
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    i64 p = op_ld(pos, i);
    if ((u64)p >= (u64)src_len) { atomicAdd(errflag, 1); out[i] = 0; }
    else out[i] = op_ld(src, p);
  }
}

extern "C" int vdl_op_gather(vdl_ctx *ctx, vdl_vec src, vdl_vec pos, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  Vec *vs = vec_get(ctx, src), *vp = vec_get(ctx, pos);
  if (!vs || !vp) return VDL_EINVAL;
  i64 n = vp->len, m = vs->len, dom = vs->domain;
  Operand os = operand_of(*vs), op = operand_of(*vp);
  VDL_TRY(vec_new(ctx, VDL_I64, n, out));
  ctx->vecs[*out].domain = dom;   // gathering positions keeps their index space
  if (n == 0) return VDL_OK;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  int blocks = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 16));
  gather_kernel<<<blocks, 256, 0, ctx->stream>>>(os, m, op, n, (i64 *)ctx->vecs[*out].ptr, ctx->d_errflag);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}

// Scatter (Vlite.hs:1057-1059 group sort, 1268-1275 dim validity / inverse index): out[pos[i]] = src[i],
// unwritten slots 0.  Positions are unique in every use the translator emits (permutations, Unique masks).
__global__ void __launch_bounds__(256) scatter_kernel(Operand src, Operand pos, i64 n, i64 *__restrict__ out, i64 out_len, int *errflag) {
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    i64 p = op_ld(pos, i);
    if ((u64)p >= (u64)out_len) atomicAdd(errflag, 1);
    else out[p] = op_ld(src, i);
  }
}

extern "C" int vdl_op_scatter(vdl_ctx *ctx, vdl_vec src, vdl_vec pos, int64_t out_len, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  Vec *vs = vec_get(ctx, src), *vp = vec_get(ctx, pos);
  if (!vs || !vp) return VDL_EINVAL;
  if (vs->len != vp->len) return vdl_fail(ctx, VDL_EINVAL, "Scatter: source length %lld != positions length %lld", (long long)vs->len, (long long)vp->len);
  if (out_len < 0) return vdl_fail(ctx, VDL_EINVAL, "Scatter: negative output length");
  i64 n = vs->len, dom = vs->domain;
  Operand os = operand_of(*vs), op = operand_of(*vp);
  VDL_TRY(vec_new(ctx, VDL_I64, out_len, out));
  ctx->vecs[*out].domain = dom;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  if (out_len > 0) VDL_CUDA(ctx, cudaMemsetAsync(ctx->vecs[*out].ptr, 0, (size_t)out_len * 8, ctx->stream));
  if (n == 0) return VDL_OK;
  int blocks = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 16));
  scatter_kernel<<<blocks, 256, 0, ctx->stream>>>(os, op, n, (i64 *)ctx->vecs[*out].ptr, out_len, ctx->d_errflag);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- Partition
// Vlite.hs:1082-1098 emits Partition(key, RangeC(min,1,max-min+1)); its result is the scatter position that
// sorts rows stably by key (1057-1060, 1172).  bucket(v) = number of pivots below v (App. G3); the
// permutation is a stable LSD radix sort of row ids by bucket, 8 bits per pass over the bits the pivot
// count needs (Q1: 6 bits = 1 pass; Q3's 38-bit composite key = 5 passes).
#define RDX_TILE 4096

__device__ __forceinline__ u64 bucket_of(i64 v, i64 pfrom, i64 pstep, i64 pcount) {
  if (v <= pfrom) return 0;
  u64 d = (u64)v - (u64)pfrom;
  u64 b = pstep == 1 ? d : (d + (u64)pstep - 1) / (u64)pstep;     // every emitted Partition has unit-step pivots (Vlite.hs:1088-1091)
  return b > (u64)pcount ? (u64)pcount : b;
}

__global__ void __launch_bounds__(256) bucket_kernel(Operand data, i64 n, i64 pfrom, i64 pstep, i64 pcount, u64 *__restrict__ key, i64 *__restrict__ idx) {
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    i64 v = op_ld(data, i);
    u64 b;
    if (v <= pfrom) b = 0;
    else {
      u64 d = (u64)v - (u64)pfrom;
      b = (d + (u64)pstep - 1) / (u64)pstep;
      if (b > (u64)pcount) b = (u64)pcount;
    }
    key[i] = b;
    idx[i] = i;
  }
}

// digit histogram per block, stored digit-major: hist[d * nblocks + block]
__global__ void __launch_bounds__(256) radix_hist_kernel(const u64 *__restrict__ key, i64 n, int shift, i64 nblocks, i64 *__restrict__ hist) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  i64 base = (i64)blockIdx.x * RDX_TILE;
  for (int s = 0; s < RDX_TILE / 256; s++) {
    i64 i = base + s * 256 + threadIdx.x;
    if (i < n) atomicAdd(&h[(key[i] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[(i64)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// stable scatter: rank inside the block = elements with the same digit earlier in the block
__global__ void __launch_bounds__(256) radix_scatter_kernel(const u64 *__restrict__ key_in, const i64 *__restrict__ idx_in, i64 n, int shift,
                                                            i64 nblocks, const i64 *__restrict__ offs, u64 *__restrict__ key_out,
                                                            i64 *__restrict__ idx_out) {
  __shared__ int wcount[8][256];   // per-warp count of each digit in the current step
  __shared__ int wbase[8][256];    // per-warp base of each digit in the current step
  __shared__ int run[256];         // elements of each digit in earlier steps of this block
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = 0; w < 8; w++) wcount[w][tid] = 0;
  run[tid] = 0;
  __syncthreads();
  i64 base = (i64)blockIdx.x * RDX_TILE;
  for (int s = 0; s < RDX_TILE / 256; s++) {
    i64 i = base + s * 256 + tid;
    bool valid = i < n;
    u64 k = valid ? key_in[i] : 0;
    int d = valid ? (int)((k >> shift) & 255) : 256 + lane;   // invalid lanes never match anybody
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int rank_in_warp = __popc(peers & ((1u << lane) - 1));
    if (valid && rank_in_warp == 0) wcount[warp][d] = __popc(peers);
    __syncthreads();
    {   // thread `tid` owns digit `tid`: per-warp counts -> per-warp bases; counts are cleared for the next step
      int off = run[tid];
#pragma unroll
      for (int w = 0; w < 8; w++) {
        int c = wcount[w][tid];
        wcount[w][tid] = 0;
        wbase[w][tid] = off;
        off += c;
      }
      run[tid] = off;
    }
    __syncthreads();
    if (valid) {
      i64 dst = offs[(i64)d * nblocks + blockIdx.x] + wbase[warp][d] + rank_in_warp;
      key_out[dst] = k;
      idx_out[dst] = idx_in[i];
    }
  }
}

// One-pass variant for <= 256 buckets (every low-cardinality group-by: Q1's 32, Q5's 128): the same histogram / scan,
// then the destination of row i is written straight to out[i] -- no key / index ping-pong, no inversion pass.
__global__ void __launch_bounds__(256) bucket_hist_kernel(Operand data, i64 n, i64 pfrom, i64 pstep, i64 pcount, i64 nblocks, i64 *__restrict__ hist) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  i64 base = (i64)blockIdx.x * RDX_TILE;
  for (int s = 0; s < RDX_TILE / 256; s++) {
    i64 i = base + s * 256 + threadIdx.x;
    if (i < n) atomicAdd(&h[(int)bucket_of(op_ld(data, i), pfrom, pstep, pcount)], 1);
  }
  __syncthreads();
  hist[(i64)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}
__global__ void __launch_bounds__(256) bucket_place_kernel(Operand data, i64 n, i64 pfrom, i64 pstep, i64 pcount, i64 nblocks,
                                                           const i64 *__restrict__ offs, i64 *__restrict__ out) {
  __shared__ int wcount[8][256];
  __shared__ int wbase[8][256];
  __shared__ int run[256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int w = 0; w < 8; w++) wcount[w][tid] = 0;
  run[tid] = 0;
  __syncthreads();
  i64 base = (i64)blockIdx.x * RDX_TILE;
  for (int s = 0; s < RDX_TILE / 256; s++) {
    i64 i = base + s * 256 + tid;
    bool valid = i < n;
    int d = valid ? (int)bucket_of(op_ld(data, i), pfrom, pstep, pcount) : 256 + lane;
    unsigned peers = __match_any_sync(0xffffffffu, d);
    int rank_in_warp = __popc(peers & ((1u << lane) - 1));
    if (valid && rank_in_warp == 0) wcount[warp][d] = __popc(peers);
    __syncthreads();
    {
      int off = run[tid];
#pragma unroll
      for (int w = 0; w < 8; w++) {
        int c = wcount[w][tid];
        wcount[w][tid] = 0;
        wbase[w][tid] = off;
        off += c;
      }
      run[tid] = off;
    }
    __syncthreads();
    if (valid) out[i] = offs[(i64)d * nblocks + blockIdx.x] + wbase[warp][d] + rank_in_warp;
  }
}

__global__ void __launch_bounds__(256) invert_perm_kernel(const i64 *__restrict__ order, i64 n, i64 *__restrict__ out) {
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) out[order[j]] = j;
}

// rows already ordered by bucket (e.g. lineitem clustered on l_orderkey, storage.csv `sorted`)?  counts the descents
__global__ void __launch_bounds__(256) descents_kernel(Operand data, i64 n, i64 pfrom, i64 pstep, i64 pcount, int *descents) {
  // 4 consecutive elements per thread (+ the next one): independent loads in flight, every element read ~1.25 times
  const i64 stride = (i64)gridDim.x * blockDim.x * 4;
  int bad = 0;
  for (i64 i0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += stride) {
    u64 b[5];
#pragma unroll
    for (int k = 0; k < 5; k++) b[k] = i0 + k < n ? bucket_of(op_ld(data, i0 + k), pfrom, pstep, pcount) : ~0ull;
#pragma unroll
    for (int k = 0; k < 4; k++) bad |= b[k] > b[k + 1] && i0 + k + 1 < n;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAdd(descents, 1);
}

extern "C" int vdl_op_partition(vdl_ctx *ctx, vdl_vec data, int64_t pfrom, int64_t pstep, int64_t pcount, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  Vec *vd = vec_get(ctx, data);
  if (!vd) return VDL_EINVAL;
  if (pcount < 1 || pstep < 1) return vdl_fail(ctx, VDL_EINVAL, "Partition: pivots must be an ascending range (count %lld step %lld)", (long long)pcount, (long long)pstep);
  i64 n = vd->len;
  Operand od = operand_of(*vd);
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n > 0) {
    // one cheap pass first: keys that are already in bucket order sort to the identity permutation, which stays a
    // virtual range -- no radix passes, and a Scatter by it is the source vector itself
    VDL_TRY(scratch_reserve(ctx, 64));
    int *d_desc = (int *)ctx->scratch, h_desc = 1;
    VDL_CUDA(ctx, cudaMemsetAsync(d_desc, 0, sizeof(int), ctx->stream));
    int g0 = (int)std::max<i64>(1, std::min<i64>((n + 1023) / 1024, (i64)ctx->sm_count * 16));
    descents_kernel<<<g0, 256, 0, ctx->stream>>>(od, n, pfrom, pstep, pcount, d_desc);
    ctx->launches++;
    VDL_TRY(read_scalar(ctx, d_desc, &h_desc, sizeof(int)));
    if (h_desc == 0) return vec_new_range(ctx, 0, 1, n, out);      // (its domain is n: positions 0..n-1)
  }
  VDL_TRY(vec_new(ctx, VDL_I64, n, out));
  ctx->vecs[*out].domain = n;
  if (n == 0) return VDL_OK;
  int bits = 0;
  while (bits < 63 && ((u64)pcount >> bits)) bits++;
  i64 nb = (n + RDX_TILE - 1) / RDX_TILE;
  if (pcount < 256) {                       // buckets 0..pcount fit one digit
    VDL_TRY(scratch_reserve(ctx, scan_elems(256 * nb) * 8));
    i64 *h = (i64 *)ctx->scratch;
    bucket_hist_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(od, n, pfrom, pstep, pcount, nb, h);
    device_exclusive_scan(ctx, h, 256 * nb);
    bucket_place_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(od, n, pfrom, pstep, pcount, nb, h, (i64 *)ctx->vecs[*out].ptr);
    ctx->launches += 2;
    VDL_CUDA(ctx, cudaGetLastError());
    return VDL_OK;
  }
  // temporaries: key/idx ping-pong + histogram
  vdl_vec tk[2], ti[2];
  for (int k = 0; k < 2; k++) { VDL_TRY(vec_new(ctx, VDL_I64, n, &tk[k])); VDL_TRY(vec_new(ctx, VDL_I64, n, &ti[k])); }
  VDL_TRY(scratch_reserve(ctx, scan_elems(256 * nb) * 8));
  i64 *hist = (i64 *)ctx->scratch;
  u64 *key[2] = {(u64 *)ctx->vecs[tk[0]].ptr, (u64 *)ctx->vecs[tk[1]].ptr};
  i64 *idx[2] = {(i64 *)ctx->vecs[ti[0]].ptr, (i64 *)ctx->vecs[ti[1]].ptr};
  int grid = (int)std::max<i64>(1, std::min<i64>((n + 255) / 256, (i64)ctx->sm_count * 16));
  bucket_kernel<<<grid, 256, 0, ctx->stream>>>(od, n, pfrom, pstep, pcount, key[0], idx[0]);
  ctx->launches++;
  int cur = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(key[cur], n, shift, nb, hist);
    device_exclusive_scan(ctx, hist, 256 * nb);
    radix_scatter_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(key[cur], idx[cur], n, shift, nb, hist, key[cur ^ 1], idx[cur ^ 1]);
    ctx->launches += 2;
    cur ^= 1;
  }
  invert_perm_kernel<<<grid, 256, 0, ctx->stream>>>(idx[cur], n, (i64 *)ctx->vecs[*out].ptr);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  for (int k = 0; k < 2; k++) { VDL_TRY(vdl_vec_free(ctx, tk[k])); VDL_TRY(vdl_vec_free(ctx, ti[k])); }
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- Fold by runs
// FoldSum/Min/Max/Choose/Count (Vlite.hs:1048-1070, 1179; Vdl.hs:255-264): one output per run of equal
// consecutive `groups` values, in run order.  Run id = (number of run heads up to the element) - 1; inside a
// warp equal run ids are contiguous, so a segmented shuffle reduction leaves one global atomic per
// (warp, run) -- long runs (the common case after a Partition sort) cost almost no atomics.
__device__ __forceinline__ i64 fold_identity(int op) { return op == VDL_FOLD_MIN ? INT64_MAX : (op == VDL_FOLD_MAX ? INT64_MIN : 0); }
__device__ __forceinline__ i64 fold_combine(int op, i64 a, i64 b) {
  if (op == VDL_FOLD_MIN) return b < a ? b : a;
  if (op == VDL_FOLD_MAX) return b > a ? b : a;
  return (i64)((u64)a + (u64)b);
}

__global__ void __launch_bounds__(256) fold_init_kernel(i64 *out, i64 n, i64 v) {
  i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = v;
}

__global__ void __launch_bounds__(256) fold_runs_kernel(int op, Operand groups, Operand data, i64 n, const i64 *__restrict__ block_off, i64 *__restrict__ out) {
  __shared__ int wcnt[8];
  const int lane = threadIdx.x & 31;
  i64 base = (i64)blockIdx.x * SEL_TILE;
  i64 off = block_off[blockIdx.x];   // run heads before this block
  if (block_off[blockIdx.x + 1] == off && op != VDL_FOLD_CHOOSE) {
    // no run starts inside this tile: all of it continues run off-1 (long runs: a single-group fold, a sorted low-
    // cardinality key) -> plain block reduction, ONE atomic per tile instead of one per warp and step
    __shared__ i64 wred[8];
    i64 v = fold_identity(op);
    for (int s = 0; s < SEL_TILE / 256; s++) {
      i64 i = base + s * 256 + threadIdx.x;
      if (i < n) v = fold_combine(op, v, op == VDL_FOLD_COUNT ? 1 : op_ld(data, i));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fold_combine(op, v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) wred[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++) v = fold_combine(op, v, wred[w]);
      i64 *dst = &out[off - 1];
      if (op == VDL_FOLD_MIN) atomicMin((long long *)dst, (long long)v);
      else if (op == VDL_FOLD_MAX) atomicMax((long long *)dst, (long long)v);
      else atomicAdd((unsigned long long *)dst, (unsigned long long)v);
    }
    return;
  }
  int run = 0;
  for (int s = 0; s < SEL_TILE / 256; s++) {
    i64 i = base + s * 256 + threadIdx.x;
    bool valid = i < n;
    bool head = valid && flag_at<true>(groups, i);
    StepRank r = step_rank(head, wcnt, run);
    i64 rid = off + r.incl - 1;        // run id of this element
    if (op == VDL_FOLD_CHOOSE) {
      if (head) out[rid] = op_ld(data, i);
      continue;
    }
    i64 v = valid ? (op == VDL_FOLD_COUNT ? 1 : op_ld(data, i)) : fold_identity(op);
    if (!valid) rid = -1 - lane;       // never equal to a neighbour
    // segmented inclusive scan over lanes with equal rid (contiguous)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      i64 pv = __shfl_up_sync(0xffffffffu, v, o);
      i64 pr = __shfl_up_sync(0xffffffffu, rid, o);
      if (lane >= o && pr == rid) v = fold_combine(op, v, pv);
    }
    i64 nr = __shfl_down_sync(0xffffffffu, rid, 1);
    bool last = lane == 31 || nr != rid;
    // a run whose head AND end are inside this warp's 32 elements belongs to this lane alone: plain store
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const int head_lane = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));     // lane of this element's run head, -1 if in an earlier warp
    const bool own = valid && last && lane < 31 && nr != rid && head_lane >= 0;
    if (own) { out[rid] = v; continue; }
    if (valid && last) {
      if (op == VDL_FOLD_MIN) atomicMin((long long *)&out[rid], (long long)v);
      else if (op == VDL_FOLD_MAX) atomicMax((long long *)&out[rid], (long long)v);
      else atomicAdd((unsigned long long *)&out[rid], (unsigned long long)v);
    }
  }
}

extern "C" int vdl_op_fold(vdl_ctx *ctx, int fold_op, vdl_vec groups, vdl_vec data, vdl_vec *out) {
  if (!ctx || !out) return VDL_EINVAL;
  if (fold_op < VDL_FOLD_SUM || fold_op > VDL_FOLD_COUNT) return vdl_fail(ctx, VDL_EINVAL, "fold op %d unknown", fold_op);
  Vec *vg = vec_get(ctx, groups), *vd = vec_get(ctx, data);
  if (!vg || !vd) return VDL_EINVAL;
  if (vg->len != vd->len) {
    // Level 2 of a hierarchical fold (make2LevelFold, Vlite.hs:1181-1192): `data` holds one result per level-1 run and
    // `groups` is still row-aligned.  Dense model: result k belongs to the group value at the first row of level-1 run k,
    // i.e. FoldChoose(level-1 groups, groups).  Anything else of unequal lengths is an error.
    const vdl_vec g1 = vd->fold_groups;
    Vec *v1 = g1 > 0 && (size_t)g1 < ctx->vecs.size() && ctx->vecs[g1].live ? &ctx->vecs[g1] : nullptr;
    if (!v1 || v1->gen != vd->fold_groups_gen || v1->len != vg->len)
      return vdl_fail(ctx, VDL_EINVAL, "Fold: groups length %lld != data length %lld", (long long)vg->len, (long long)vd->len);
    vdl_vec heads = 0;
    VDL_TRY(vdl_op_fold(ctx, VDL_FOLD_CHOOSE, g1, groups, &heads));
    int rc = vdl_op_fold(ctx, fold_op, heads, data, out);
    vdl_vec_free(ctx, heads);
    return rc;
  }
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  const u64 groups_gen = vg->gen;
  i64 n = vd->len;
  Operand og = operand_of(*vg), od = operand_of(*vd);
  i64 total = 0, *off = nullptr;
  if (n > 0) VDL_TRY(flag_scan(ctx, true, og, n, &off, &total));
  VDL_TRY(vec_new(ctx, VDL_I64, total, out));
  ctx->vecs[*out].fold_groups = groups;
  ctx->vecs[*out].fold_groups_gen = groups_gen;
  if (total == 0) return VDL_OK;
  i64 *o = (i64 *)ctx->vecs[*out].ptr;
  int grid = (int)std::max<i64>(1, std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16));
  fold_init_kernel<<<grid, 256, 0, ctx->stream>>>(o, total, fold_op == VDL_FOLD_MIN ? INT64_MAX : (fold_op == VDL_FOLD_MAX ? INT64_MIN : 0));
  fold_runs_kernel<<<(unsigned)((n + SEL_TILE - 1) / SEL_TILE), 256, 0, ctx->stream>>>(fold_op, og, od, n, off, o);
  ctx->launches += 2;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}
