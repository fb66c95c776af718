// Declarations shared by the device code of libvdl_cuda, written so that NVRTC can compile them too (the fused scan is
// specialised at run time: vdl_fused_jit.cu): no host headers, fixed-width types spelled out under __CUDACC_RTC__.
#pragma once

#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
typedef unsigned long uintptr_t;
#define INT32_MAX 2147483647
#define INT32_MIN (-2147483647 - 1)
#define INT64_MAX 9223372036854775807LL
#define INT64_MIN (-9223372036854775807LL - 1)
#else
#include <stdint.h>
#endif

#include "vdl_cuda.h"

typedef int64_t i64;
typedef uint64_t u64;

// Peer-memory exchange of the partial tables (one buffer per rank, addressable by all ranks):
//   data  [2 (epoch parity)][world][stride] int64   rank r's table of the step lands in slot [parity][r] of EVERY buffer
//   flags [2][world] uint64                         epoch of the last step whose table rank r has fully stored
struct XDesc {
  int32_t rank, world;
  u64 epoch;                      // this step's number (1, 2, ...); parity double-buffers against a rank running ahead
  u64 timeout_ns;                 // give up waiting for a peer after this long: error flag, never a hang
  i64 stride;                     // int64 per table
  i64 *peer[VDL_MAX_RANKS];       // base of every rank's buffer as seen from this GPU
};



#ifdef __CUDACC__
// system-scope release / acquire and a wall clock for the peer-memory exchange (vdl_fused.cu, vdl_probe.cu)
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 global_timer_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Elementwise op semantics (Vdl.hs:136-157, 209-231).  Comparisons / logicals give 0/1; BitShift: +k arithmetic right,
// -k left (Vlite.hs:205-208); Divide truncates, x/0 := 0, INT64_MIN/-1 wraps; Modulo is the C remainder, x%0 := 0.
__device__ __forceinline__ i64 binop_apply(int op, i64 a, i64 b) {
  switch (op) {
    case VDL_LOGICAL_AND: return (a != 0) && (b != 0);
    case VDL_LOGICAL_OR: return (a != 0) || (b != 0);
    case VDL_BITWISE_AND: return a & b;
    case VDL_BITWISE_OR: return a | b;
    case VDL_BITSHIFT:
      if (b >= 0) return b >= 64 ? (a < 0 ? -1 : 0) : (a >> b);
      return b <= -64 ? 0 : (i64)((u64)a << (-b));
    case VDL_EQUALS: return a == b;
    case VDL_ADD: return (i64)((u64)a + (u64)b);
    case VDL_SUBTRACT: return (i64)((u64)a - (u64)b);
    case VDL_GREATER: return a > b;
    case VDL_MULTIPLY: return (i64)((u64)a * (u64)b);
    case VDL_DIVIDE:
      if (b == 0) return 0;
      if (b == -1) return (i64)(0 - (u64)a);
      return a / b;
    case VDL_MODULO:
      if (b == 0 || b == -1) return 0;
      return a % b;
  }
  return 0;
}
#endif
