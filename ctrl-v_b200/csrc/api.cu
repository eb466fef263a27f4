// Library plumbing: thread-local error string, device check, TMA tensor-map encoding through
// the driver entry point (no link-time dependency on libcuda).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

static thread_local char g_err[512] = {0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CTRLV_PDL");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;  // opt-in: measured neutral inside the captured step graph
  }
  return v != 0;
}

// ---- launch plans --------------------------------------------------------------------------------------
// A plan is the ordered list of everything the library was asked to do on the recording thread between
// ctrlv_plan_create and ctrlv_plan_finish: kernel launches (entry point, grid, block, shared memory, cluster
// size, a copy of the argument bytes — CUtensorMaps included), ctrlv_memset_zero calls, and fork / join marks
// of a second stream.  ctrlv_plan_run replays it on a stream; launches recorded on any stream other than the
// plan's main stream go to the plan's own side stream, ordered by the recorded fork / join events.
constexpr int kPlanMaxArgs = 24;
struct PlanOp {
  int kind;  // 0 launch, 1 memset, 2 fork (side waits for main), 3 join (main waits for side)
  int side;  // 1: issued on the side stream
  const void* func;
  dim3 grid, block;
  size_t smem;
  int cluster_x;
  int nargs;
  size_t arg_off[kPlanMaxArgs];  // offsets into the argument arena
  void* ptr;          // memset
  size_t bytes;
  cudaEvent_t ev;     // fork / join
};
struct Plan {
  std::vector<PlanOp> ops;
  std::vector<unsigned char> arena;  // argument bytes, each blob 128-byte aligned relative to arena_base()
  cudaStream_t main = nullptr, side = nullptr;
  bool finished = false, broken = false;
  long long launches = 0;
  unsigned char* arena_base() { return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(arena.data()) + 127) & ~uintptr_t(127)); }
};
static thread_local Plan* g_rec = nullptr;

bool plan_recording() { return g_rec != nullptr; }

void plan_record_launch(const void* func, dim3 grid, dim3 block, size_t smem, int cluster_x, cudaStream_t stream,
                        void** args, const size_t* sizes, int nargs) {
  Plan* pl = g_rec;
  if (!pl) return;
  if (nargs > kPlanMaxArgs) {  // (no kernel of this library has that many; fail loudly at ctrlv_plan_finish)
    pl->broken = true;
    set_last_error("plan: a launch with %d arguments cannot be recorded", nargs);
    return;
  }
  PlanOp op;
  memset(&op, 0, sizeof(op));
  op.kind = 0;
  op.side = (stream != pl->main) ? 1 : 0;
  op.func = func; op.grid = grid; op.block = block; op.smem = smem; op.cluster_x = cluster_x; op.nargs = nargs;
  // arena offsets are kept relative (the vector may grow); alignment is restored against arena_base() at run time
  for (int i = 0; i < nargs; ++i) {
    size_t off = (pl->arena.size() + 127) / 128 * 128;
    op.arg_off[i] = off;
    pl->arena.resize(off + sizes[i] + 128);
    memcpy(pl->arena.data() + off, args[i], sizes[i]);
  }
  pl->ops.push_back(op);
  pl->launches++;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return CTRLV_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return CTRLV_ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return CTRLV_OK;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  int rc = load_encode();
  if (rc) return rc;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error(
        "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] "
        "strides [%llu %llu %llu] box [%u %u %u %u] base %p",
        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
        (unsigned long long)(rank > 2 ? strides_bytes[1] : 0),
        (unsigned long long)(rank > 3 ? strides_bytes[2] : 0), box[0], rank > 1 ? box[1] : 0,
        rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return CTRLV_ERR_CUDA;
  }
  return CTRLV_OK;
}

}  // namespace ctrlv

extern "C" const char* ctrlv_last_error(void) { return ctrlv::g_err; }

extern "C" int64_t ctrlv_launch_count(void) { return (int64_t)ctrlv::g_launches.load(std::memory_order_relaxed); }

struct ctrlv_plan { ctrlv::Plan p; };

extern "C" int ctrlv_plan_create(void* main_stream, ctrlv_plan** out) {
  CTRLV_CHECK_ARG(out != nullptr, "plan_create: null output");
  CTRLV_CHECK_ARG(ctrlv::g_rec == nullptr, "plan_create: this thread is already recording a plan");
  ctrlv_plan* pl = new ctrlv_plan();
  pl->p.main = reinterpret_cast<cudaStream_t>(main_stream);
  cudaError_t e = cudaStreamCreateWithFlags(&pl->p.side, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete pl;
    ctrlv::set_last_error("plan_create: cudaStreamCreate failed: %s", cudaGetErrorString(e));
    return CTRLV_ERR_CUDA;
  }
  ctrlv::g_rec = &pl->p;
  *out = pl;
  return CTRLV_OK;
}

static int plan_mark(int kind) {
  ctrlv::Plan* pl = ctrlv::g_rec;
  if (!pl) return CTRLV_OK;  // not recording: the caller's own stream ordering is all there is
  ctrlv::PlanOp op;
  memset(&op, 0, sizeof(op));
  op.kind = kind;
  CTRLV_CUDA(cudaEventCreateWithFlags(&op.ev, cudaEventDisableTiming));
  pl->ops.push_back(op);
  return CTRLV_OK;
}
extern "C" int ctrlv_plan_fork(void) { return plan_mark(2); }
extern "C" int ctrlv_plan_join(void) { return plan_mark(3); }

extern "C" int ctrlv_plan_finish(ctrlv_plan* plan) {
  CTRLV_CHECK_ARG(plan != nullptr && ctrlv::g_rec == &plan->p, "plan_finish: this plan is not being recorded on this thread");
  ctrlv::g_rec = nullptr;
  CTRLV_CHECK_ARG(!plan->p.broken, "plan_finish: a launch could not be recorded (%s)", ctrlv::g_err);
  plan->p.finished = true;
  // move the argument bytes so that every blob is 128-byte aligned in memory (CUtensorMap needs 64)
  std::vector<unsigned char> al(plan->p.arena.size() + 256);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(al.data()) + 127) & ~uintptr_t(127));
  memcpy(base, plan->p.arena.data(), plan->p.arena.size());
  plan->p.arena.swap(al);
  return CTRLV_OK;
}

extern "C" int64_t ctrlv_plan_size(const ctrlv_plan* plan) { return plan ? (int64_t)plan->p.launches : 0; }

extern "C" int ctrlv_plan_run(ctrlv_plan* plan, void* stream_) {
  CTRLV_CHECK_ARG(plan != nullptr && plan->p.finished, "plan_run: plan is null or still recording");
  ctrlv::Plan& pl = plan->p;
  cudaStream_t main = reinterpret_cast<cudaStream_t>(stream_);
  unsigned char* base = pl.arena_base();
  for (const ctrlv::PlanOp& op : pl.ops) {
    cudaStream_t st = op.side ? pl.side : main;
    if (op.kind == 0) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = op.grid; cfg.blockDim = op.block; cfg.dynamicSmemBytes = op.smem; cfg.stream = st;
      cudaLaunchAttribute attr[2];
      int na = 0;
      if (op.cluster_x > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)op.cluster_x; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
      }
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = ctrlv::pdl_enabled() ? 1 : 0;
      ++na;
      cfg.attrs = attr; cfg.numAttrs = (unsigned)na;
      void* argv[ctrlv::kPlanMaxArgs];
      for (int i = 0; i < op.nargs; ++i) argv[i] = base + op.arg_off[i];
      ctrlv::count_launch();
      CTRLV_CUDA(cudaLaunchKernelExC(&cfg, op.func, argv));
    } else if (op.kind == 1) {
      CTRLV_CUDA(cudaMemsetAsync(op.ptr, 0, op.bytes, st));
    } else if (op.kind == 2) {
      CTRLV_CUDA(cudaEventRecord(op.ev, main));
      CTRLV_CUDA(cudaStreamWaitEvent(pl.side, op.ev, 0));
    } else {
      CTRLV_CUDA(cudaEventRecord(op.ev, pl.side));
      CTRLV_CUDA(cudaStreamWaitEvent(main, op.ev, 0));
    }
  }
  return CTRLV_OK;
}

extern "C" int ctrlv_plan_destroy(ctrlv_plan* plan) {
  if (!plan) return CTRLV_OK;
  if (ctrlv::g_rec == &plan->p) ctrlv::g_rec = nullptr;
  for (ctrlv::PlanOp& op : plan->p.ops)
    if (op.ev) cudaEventDestroy(op.ev);
  if (plan->p.side) cudaStreamDestroy(plan->p.side);
  delete plan;
  return CTRLV_OK;
}

extern "C" int ctrlv_memset_zero(void* ptr, int64_t bytes, void* stream_) {
  CTRLV_CHECK_ARG(ptr != nullptr && bytes >= 0, "memset_zero: bad arguments");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (ctrlv::Plan* pl = ctrlv::g_rec) {
    ctrlv::PlanOp op;
    memset(&op, 0, sizeof(op));
    op.kind = 1; op.side = (stream != pl->main) ? 1 : 0; op.ptr = ptr; op.bytes = (size_t)bytes;
    pl->ops.push_back(op);
  }
  CTRLV_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, stream));
  return CTRLV_OK;
}

extern "C" const char* ctrlv_version(void) { return "ctrlv_b200 0.1 (sm_100a)"; }

extern "C" int ctrlv_device_check(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    ctrlv::set_last_error("no CUDA device visible (%s)", cudaGetErrorString(e));
    return CTRLV_ERR_CUDA;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    ctrlv::set_last_error("device compute capability %d.x is not sm_100 (B200)", major);
    return CTRLV_ERR_UNSUPPORTED;
  }
  return CTRLV_OK;
}
