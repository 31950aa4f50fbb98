// C-ABI entry points: dtype dispatch, the program executor and its CUDA-graph form.
// (The per-op entry points of entropy.cu / layout.cu are defined next to their kernels.)
#include <new>

#include "common.cuh"

int rdsic_conv_forward_f32(const rdsic_conv_desc* d, cudaStream_t stream);
int rdsic_conv_forward_bf16(const rdsic_conv_desc* d, cudaStream_t stream);
int rdsic_attn_forward_f32(const rdsic_attn_desc* d, cudaStream_t stream);

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
    default: return -1;
  }
}

int rdsic_conv_forward(const rdsic_conv_desc* d, rdsic_stream_t stream) {
  if (!d) return RDSIC_E_ARG;
  if (d->w_dtype == RDSIC_BF16) return rdsic_conv_forward_bf16(d, (cudaStream_t)stream);
  return rdsic_conv_forward_f32(d, (cudaStream_t)stream);
}

int rdsic_attn_forward(const rdsic_attn_desc* d, rdsic_stream_t stream) {
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
    default: return RDSIC_E_ARG;
  }
}

// side == nullptr: every lane runs in program order on `mainS` (serial, always correct)
static int run_lanes(const rdsic_op* ops, int n_ops, cudaStream_t mainS, cudaStream_t side, int* n_launched,
                     int* failed_op) {
  int launched = 0, rc = 0, i = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (side) {
    for (int k = 0; k < 2 && !rc; ++k) rc = (int)cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
  }
  for (; i < n_ops && !rc; ++i) {
    const rdsic_op* op = &ops[i];
    if (op->kind == RDSIC_OP_FORK || op->kind == RDSIC_OP_JOIN) {
      if (!side) continue;
      cudaStream_t from = op->kind == RDSIC_OP_FORK ? mainS : side, to = op->kind == RDSIC_OP_FORK ? side : mainS;
      cudaEvent_t e = ev[op->kind == RDSIC_OP_FORK ? 0 : 1];
      rc = (int)cudaEventRecord(e, from);
      if (!rc) rc = (int)cudaStreamWaitEvent(to, e, 0);
      continue;
    }
    rc = run_one(op, (rdsic_stream_t)((side && op->lane == 1) ? side : mainS));
    if (!rc) ++launched;  // every compute op is exactly one kernel launch
  }
  for (int k = 0; k < 2; ++k)
    if (ev[k]) cudaEventDestroy(ev[k]);
  if (rc && failed_op) *failed_op = i - 1;
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
  cudaStream_t side = nullptr;  // lane 1 during capture only: becomes parallel graph branches
  cudaError_t e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete g; return (int)e; }
  e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { cudaStreamDestroy(side); delete g; return (int)e; }
  int launched = 0, failed = -1;
  int rc = run_lanes(ops, n_ops, s, side, &launched, &failed);
  e = cudaStreamEndCapture(s, &g->graph);
  cudaStreamDestroy(side);
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
