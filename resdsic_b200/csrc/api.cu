// C-ABI entry points: dtype dispatch, the program executor and its CUDA-graph form.
// (The per-op entry points of entropy.cu / layout.cu are defined next to their kernels.)
#include <stdlib.h>

#include <new>
#include <vector>

#include "common.cuh"

int rdsic_conv_forward_f32(const rdsic_conv_desc* d, cudaStream_t stream);
int rdsic_conv_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream);
int rdsic_conv_gdn_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream);
int rdsic_attn_forward_f32(const rdsic_attn_desc* d, cudaStream_t stream);
int rdsic_attn_forward_tc(const rdsic_attn_desc* d, cudaStream_t stream);

extern "C" {

int rdsic_abi_version(void) { return RDSIC_ABI_VERSION; }

const char* rdsic_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case RDSIC_E_ARG: return "resdsic_b200: invalid argument or unsupported shape";
    case RDSIC_E_ALIGN: return "resdsic_b200: misaligned pointer or stride";
    case RDSIC_E_UNSUPPORTED: return "resdsic_b200: unsupported configuration";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "resdsic_b200: unknown error";
  }
}

int rdsic_sizeof(int what) {
  switch (what) {
    case 0: return (int)sizeof(rdsic_op);
    case 1: return (int)sizeof(rdsic_conv_desc);
    case 2: return (int)sizeof(rdsic_attn_desc);
    case 3: return (int)sizeof(rdsic_eb_desc);
    case 4: return (int)sizeof(rdsic_gc_desc);
    case 5: return (int)sizeof(rdsic_copy_desc);
    case 6: return (int)sizeof(rdsic_view);
    case 7: return (int)sizeof(rdsic_ln_desc);
    case 8: return (int)sizeof(rdsic_patch_desc);
    case 9: return (int)sizeof(rdsic_mask_desc);
    default: return -1;
  }
}

int rdsic_conv_forward(const rdsic_conv_desc* d, rdsic_stream_t stream) {
  if (!d) return RDSIC_E_ARG;
  if (d->tail_mode) return d->w_dtype == RDSIC_BF16 ? rdsic_conv_gdn_forward_bf16(d, (cudaStream_t)stream) : RDSIC_E_UNSUPPORTED;
  if (d->w_dtype == RDSIC_BF16) return rdsic_conv_forward_bf16(d, (cudaStream_t)stream);
  return rdsic_conv_forward_f32(d, (cudaStream_t)stream);
}

int rdsic_attn_forward(const rdsic_attn_desc* d, rdsic_stream_t stream) {
  if (!d) return RDSIC_E_ARG;
  // bf16 activations: tensor-core kernel for the reference's two configurations; everything else (and the
  // fp32 mode) runs the register-resident fp32 kernel.  Both are CUDA kernels of this library.
  static const bool use_tc = !(getenv("RDSIC_ATTN_TC") && atoi(getenv("RDSIC_ATTN_TC")) == 0);
  if (use_tc && d->qkv.dtype == RDSIC_BF16 && d->qkv.ptr && d->out.ptr && d->bias_table && d->ws > 0 &&
      d->H % d->ws == 0 && d->W % d->ws == 0 && d->shift >= 0 && d->shift < d->ws && !d->qkv.nchw && !d->out.nchw) {
    const int rc = rdsic_attn_forward_tc(d, (cudaStream_t)stream);
    if (rc != RDSIC_E_UNSUPPORTED) return rc;
  }
  return rdsic_attn_forward_f32(d, (cudaStream_t)stream);
}

static int run_one(const rdsic_op* op, rdsic_stream_t stream) {
  switch (op->kind) {
    case RDSIC_OP_CONV: return rdsic_conv_forward(&op->u.conv, stream);
    case RDSIC_OP_ATTN: return rdsic_attn_forward(&op->u.attn, stream);
    case RDSIC_OP_EB: return rdsic_eb_forward(&op->u.eb, stream);
    case RDSIC_OP_GC: return rdsic_gc_forward(&op->u.gc, stream);
    case RDSIC_OP_COPY: return rdsic_copy_forward(&op->u.copy, stream);
    case RDSIC_OP_LN: return rdsic_ln_forward(&op->u.ln, stream);
    case RDSIC_OP_PATCH: return rdsic_patch_forward(&op->u.patch, stream);
    case RDSIC_OP_MASK: return rdsic_mask_forward(&op->u.mask, stream);
    default: return RDSIC_E_ARG;
  }
}

// lanes == nullptr: every lane runs in program order on `mainS` (serial, always correct).
// Otherwise lanes[0] == mainS and lanes[1..] are capture-time side streams.
static int run_lanes(const rdsic_op* ops, int n_ops, cudaStream_t mainS, cudaStream_t* lanes, int* n_launched,
                     int* failed_op) {
  int launched = 0, rc = 0, i = 0;
  std::vector<cudaEvent_t> events;
  bool used[RDSIC_MAX_LANES] = {};
  cudaEvent_t named[RDSIC_MAX_EVENTS] = {};
  auto sync = [&](int waiter, int src) -> int {
    cudaEvent_t e = nullptr;
    int r = (int)cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (r) return r;
    events.push_back(e);
    r = (int)cudaEventRecord(e, lanes[src]);
    if (!r) r = (int)cudaStreamWaitEvent(lanes[waiter], e, 0);
    return r;
  };
  for (; i < n_ops && !rc; ++i) {
    const rdsic_op* op = &ops[i];
    if (op->lane < 0 || op->lane >= RDSIC_MAX_LANES) { rc = RDSIC_E_ARG; break; }
    if (op->kind == RDSIC_OP_RECORD || op->kind == RDSIC_OP_WAIT) {
      const int id = op->u.sync.event;
      if (id < 0 || id >= RDSIC_MAX_EVENTS) { rc = RDSIC_E_ARG; break; }
      if (!lanes) continue;
      used[op->lane] = true;
      if (op->kind == RDSIC_OP_RECORD) {
        if (!named[id]) {
          rc = (int)cudaEventCreateWithFlags(&named[id], cudaEventDisableTiming);
          if (rc) break;
          events.push_back(named[id]);
        }
        rc = (int)cudaEventRecord(named[id], lanes[op->lane]);
      } else {
        if (!named[id]) { rc = RDSIC_E_ARG; break; }  // WAIT before its RECORD in program order
        rc = (int)cudaStreamWaitEvent(lanes[op->lane], named[id], 0);
      }
      continue;
    }
    if (op->kind == RDSIC_OP_FORK || op->kind == RDSIC_OP_JOIN) {
      const int src = op->u.sync.src;
      if (src < 0 || src >= RDSIC_MAX_LANES || src == op->lane) { rc = RDSIC_E_ARG; break; }
      if (lanes) {
        used[op->lane] = used[src] = true;
        rc = sync(op->lane, src);
      }
      continue;
    }
    if (lanes) used[op->lane] = true;
    rc = run_one(op, (rdsic_stream_t)(lanes ? lanes[op->lane] : mainS));
    if (!rc) ++launched;  // every compute op is exactly one kernel launch
  }
  if (lanes && !rc)
    for (int l = 1; l < RDSIC_MAX_LANES && !rc; ++l)
      if (used[l]) rc = sync(0, l);  // nothing may stay un-joined when the capture ends
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  if (rc && failed_op) *failed_op = i < n_ops ? i : n_ops - 1;
  if (n_launched) *n_launched = launched;
  return rc;
}

int rdsic_run_program(const rdsic_op* ops, int n_ops, rdsic_stream_t stream, int* n_launched, int* failed_op) {
  if (!ops || n_ops < 0) return RDSIC_E_ARG;
  return run_lanes(ops, n_ops, (cudaStream_t)stream, nullptr, n_launched, failed_op);
}

struct rdsic_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int n_kernels = 0;
};

int rdsic_graph_create(const rdsic_op* ops, int n_ops, rdsic_stream_t stream, rdsic_graph** out) {
  if (!ops || n_ops <= 0 || !out) return RDSIC_E_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  rdsic_graph* g = new (std::nothrow) rdsic_graph();
  if (!g) return (int)cudaErrorMemoryAllocation;
  // side lanes exist during capture only: they become parallel branches of the graph
  cudaStream_t lanes[RDSIC_MAX_LANES] = {};
  lanes[0] = s;
  cudaError_t e = cudaSuccess;
  for (int l = 1; l < RDSIC_MAX_LANES && e == cudaSuccess; ++l) e = cudaStreamCreateWithFlags(&lanes[l], cudaStreamNonBlocking);
  auto drop_lanes = [&]() {
    for (int l = 1; l < RDSIC_MAX_LANES; ++l)
      if (lanes[l]) cudaStreamDestroy(lanes[l]);
  };
  if (e != cudaSuccess) { drop_lanes(); delete g; return (int)e; }
  e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { drop_lanes(); delete g; return (int)e; }
  int launched = 0, failed = -1;
  int rc = run_lanes(ops, n_ops, s, lanes, &launched, &failed);
  e = cudaStreamEndCapture(s, &g->graph);
  drop_lanes();
  if (rc || e != cudaSuccess) {
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return rc ? rc : (int)e;
  }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g->graph);
    delete g;
    return (int)e;
  }
  g->n_kernels = launched;
  *out = g;
  return 0;
}

int rdsic_graph_launch(rdsic_graph* g, rdsic_stream_t stream) {
  if (!g || !g->exec) return RDSIC_E_ARG;
  cudaError_t e = cudaGraphLaunch(g->exec, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : (int)e;
}

int rdsic_graph_num_kernels(const rdsic_graph* g) { return g ? g->n_kernels : 0; }

void rdsic_graph_destroy(rdsic_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

}  // extern "C"
