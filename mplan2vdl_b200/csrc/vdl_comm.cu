// Row-range sharded execution over several GPUs of one box from ONE process, behind the C ABI (SURVEY.md section 8 b:
// `vdl_comm_init_all(n)`): what mplan2vdl_b200/dist.py does for one-process-per-GPU launches (torchrun + CUDA IPC), for a
// host -- the Haskell `Exec` module, hs/Exec.hs -- that drives all GPUs itself.  A communicator owns one context per rank,
// enables peer access between the devices and wires the exchange buffers of a plan's partial aggregate tables, so that
// vdl_comm_plan_run() is one scan-kernel launch per GPU whose last thread block stores the rank's table into every peer's
// buffer over NVLink, waits for the peers' epoch flags, merges and finalizes (exchange_block, vdl_fused_kernel.cuh): no
// collective library, no host round trip, and every rank ends with the global result.
#include <stdio.h>

#include <vector>

#include "vdl_internal.h"

struct vdl_comm {
  std::vector<vdl_ctx *> ctx;
  std::vector<int> device;
  std::string err;
};

struct vdl_comm_plan {
  vdl_comm *comm = nullptr;
  std::vector<vdl_plan *> plan;          // one per rank
  std::vector<std::vector<void *>> bufs; // [partial table][rank] exchange buffers
  bool wired = false;
};

static int comm_fail(vdl_comm *c, int code, const char *msg) {
  if (c) c->err = msg;
  return code;
}

extern "C" int vdl_comm_init_all(int nranks, const int *devices, vdl_comm **out) {
  if (!out || nranks < 1 || nranks > VDL_MAX_RANKS) return VDL_EINVAL;
  *out = nullptr;
  vdl_comm *c = new vdl_comm();
  for (int r = 0; r < nranks; r++) {
    const int dev = devices ? devices[r] : r;
    vdl_ctx *ctx = nullptr;
    int rc = vdl_ctx_create(dev, &ctx);
    if (rc) {
      for (auto *x : c->ctx) vdl_ctx_destroy(x);
      delete c;
      return rc;
    }
    c->ctx.push_back(ctx);
    c->device.push_back(dev);
  }
  // every pair of distinct devices must be able to address each other's memory (NVLink / NVSwitch)
  for (int a = 0; a < nranks; a++)
    for (int b = 0; b < nranks; b++) {
      if (c->device[a] == c->device[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, c->device[a], c->device[b]);
      if (!can) {
        for (auto *x : c->ctx) vdl_ctx_destroy(x);
        delete c;
        return VDL_EUNSUPPORTED;
      }
      cudaSetDevice(c->device[a]);
      cudaError_t e = cudaDeviceEnablePeerAccess(c->device[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        for (auto *x : c->ctx) vdl_ctx_destroy(x);
        delete c;
        return VDL_ECUDA;
      }
      cudaGetLastError();
    }
  *out = c;
  return VDL_OK;
}

extern "C" int vdl_comm_size(vdl_comm *c) { return c ? (int)c->ctx.size() : 0; }
extern "C" vdl_ctx *vdl_comm_ctx(vdl_comm *c, int rank) { return (c && rank >= 0 && rank < (int)c->ctx.size()) ? c->ctx[rank] : nullptr; }
extern "C" const char *vdl_comm_last_error(vdl_comm *c) { return c ? c->err.c_str() : "no communicator"; }

extern "C" int vdl_comm_destroy(vdl_comm *c) {
  if (!c) return VDL_EINVAL;
  for (auto *x : c->ctx) vdl_ctx_destroy(x);
  delete c;
  return VDL_OK;
}

// The same program on every rank; the caller has registered each rank's shard of the fact table (and the replicated
// dimension tables) with that rank's context (vdl_comm_ctx) under the same column names.  row_base[r] = global row id of
// rank r's first fact row (NULL: every shard starts at 0, which only matters to plans that use row ids).
extern "C" int vdl_comm_plan_load(vdl_comm *c, const char *vdl_text, int flags, const int64_t *row_base, vdl_comm_plan **out) {
  if (!c || !vdl_text || !out) return VDL_EINVAL;
  *out = nullptr;
  vdl_comm_plan *cp = new vdl_comm_plan();
  cp->comm = c;
  for (size_t r = 0; r < c->ctx.size(); r++) {
    vdl_plan *p = nullptr;
    int rc = vdl_plan_load(c->ctx[r], vdl_text, flags, &p);
    if (!rc && row_base) rc = vdl_plan_set_row_base(p, row_base[r]);
    if (rc) {
      c->err = vdl_last_error(c->ctx[r]);
      if (p) vdl_plan_destroy(p);
      for (auto *q : cp->plan) vdl_plan_destroy(q);
      delete cp;
      return rc;
    }
    cp->plan.push_back(p);
  }
  *out = cp;
  return VDL_OK;
}

extern "C" vdl_plan *vdl_comm_plan_rank(vdl_comm_plan *cp, int rank) { return (cp && rank >= 0 && rank < (int)cp->plan.size()) ? cp->plan[rank] : nullptr; }

static void free_bufs(vdl_comm_plan *cp) {
  for (auto &per : cp->bufs)
    for (size_t r = 0; r < per.size(); r++)
      if (per[r]) vdl_ipc_free(cp->comm->ctx[r], per[r]);
  cp->bufs.clear();
}

// One step: every rank's launches are issued first (asynchronous on each context's stream), then every rank's results are
// awaited -- the peer exchange inside the kernels needs all ranks in flight.  The first call runs the shards once without
// the exchange to prepare the scans (their table sizes are only known then), allocates the exchange buffers and wires them.
extern "C" int vdl_comm_plan_run(vdl_comm_plan *cp) {
  if (!cp) return VDL_EINVAL;
  vdl_comm *c = cp->comm;
  const int n = (int)cp->plan.size();
  if (n == 1) {
    int rc = vdl_plan_run(cp->plan[0]);
    if (rc) c->err = vdl_last_error(c->ctx[0]);
    return rc;
  }
  if (!cp->wired) {
    for (int r = 0; r < n; r++) {
      if (vdl_plan_num_emits(cp->plan[r]) > 0)
        return comm_fail(c, VDL_EUNSUPPORTED, "vdl_comm_plan_run: plans whose probe passes emit vectors are combined by the caller (vdl_plan_emit / vdl_plan_emit_replace)");
      if (vdl_plan_num_partials(cp->plan[r]) < 1)
        return comm_fail(c, VDL_EUNSUPPORTED, "vdl_comm_plan_run: a plan without a fused scan or a probe fold group cannot be row-sharded");
      int rc = vdl_plan_run_local(cp->plan[r]);
      if (!rc) rc = vdl_ctx_synchronize(c->ctx[r]);
      if (rc) { c->err = vdl_last_error(c->ctx[r]); return rc; }
    }
    const int np = vdl_plan_num_partials(cp->plan[0]);
    cp->bufs.assign(np, std::vector<void *>(n, nullptr));
    for (int i = 0; i < np; i++)
      for (int r = 0; r < n; r++) {
        int64_t bytes = 0;
        int rc = vdl_plan_exchange_bytes(cp->plan[r], i, n, &bytes);
        if (!rc) rc = vdl_ipc_alloc(c->ctx[r], bytes, &cp->bufs[i][r]);
        if (rc) { c->err = vdl_last_error(c->ctx[r]); free_bufs(cp); return rc; }
      }
    for (int i = 0; i < np; i++)
      for (int r = 0; r < n; r++) {
        int rc = vdl_plan_set_peers(cp->plan[r], i, r, n, cp->bufs[i].data());
        if (rc) { c->err = vdl_last_error(c->ctx[r]); free_bufs(cp); return rc; }
      }
    cp->wired = true;
  }
  int first_rc = VDL_OK;
  for (int r = 0; r < n; r++) {
    int rc = vdl_plan_launch(cp->plan[r]);
    if (rc && !first_rc) { first_rc = rc; c->err = vdl_last_error(c->ctx[r]); }
  }
  for (int r = 0; r < n; r++) {
    int rc = vdl_plan_finish(cp->plan[r], nullptr, 1);
    if (rc && !first_rc) { first_rc = rc; c->err = vdl_last_error(c->ctx[r]); }
  }
  return first_rc;
}

extern "C" int vdl_comm_plan_destroy(vdl_comm_plan *cp) {
  if (!cp) return VDL_EINVAL;
  for (auto *p : cp->plan) vdl_plan_destroy(p);
  free_bufs(cp);
  delete cp;
  return VDL_OK;
}
