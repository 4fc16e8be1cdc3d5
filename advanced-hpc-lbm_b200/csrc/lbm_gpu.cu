// lbm_gpu.cu -- host side of liblbm_b200.so: the C-ABI declared in include/lbm_gpu.h.
//
// Replaces the step loop of the reference's main() (d2q9-bgk.c:180-201) and the device
// side of initialise()/finalise() (d2q9-bgk.c:2787-2857, :2871-2890).  No PyTorch, no
// CPU fallback: every entry point either runs the CUDA path or returns an error.
#include "lbm_gpu.h"
#include "lbm_kernels.cuh"
#include "lbm_cluster.cuh"
#include "lbm_tb2.cuh"
#include "lbm_pairs.cuh"

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

struct CachedWindow { int device; size_t bytes; char* ptr; };
std::mutex g_window_mutex;
std::vector<CachedWindow> g_window_cache;     // windows of destroyed LBM_GPU_POOL lattices
struct CachedIpcMapping { cudaIpcMemHandle_t handle; int device; void* ptr; };
std::vector<CachedIpcMapping> g_ipc_cache;    // neighbours' windows opened by LBM_GPU_POOL lattices (kept mapped)

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return 1;
}

struct CudaError {
  std::string what;
};

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      char b_[512];                                                                           \
      snprintf(b_, sizeof b_, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e_),        \
               cudaGetErrorString(e_), __FILE__, __LINE__, #call);                            \
      throw CudaError{b_};                                                                    \
    }                                                                                         \
  } while (0)

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

constexpr int kPersistScalarBlocks = 296;      // measured, profiles/r01_small_grids.md
constexpr long long kPersistMaxCells = 1280 * 1024;   // K5 up to here (1024^2: 109 GLUPS against 63 for K7); from 1280^2 on K7 wins (profiles/r02_kernel_variants.md)
constexpr size_t kStagingBytes = 64u << 20;   // device staging for AoS<->SoA / mask / fields

// descriptor exchanged between processes (lbm_gpu_ipc_export / _connect)
struct IpcDesc {
  uint32_t magic;
  int32_t elem_size;
  int32_t device;
  int32_t pid;
  int32_t nx, pitch, rows;
  int32_t pad_;
  long long row0;
  unsigned long long win_addr;        // only meaningful inside the exporting process
  unsigned long long win_bytes;
  unsigned long long off_sync;        // byte offset of the sync words inside the window
  unsigned long long off_gmask;       // byte offset of the ghost mask rows inside the window
  unsigned long long steps_done;
  unsigned long long passes_done;     // kernel passes so far: the unit of the flag protocol
  long long local_free_cells;
  int32_t cur;                        // lattice buffer holding the current state
  int32_t tb2_ok;                     // this slab could run the two-step kernel
  cudaIpcMemHandle_t handle;          // of the window (the lattice itself is never shared)
  char gpu_uuid[16];                  // physical GPU: two flag-ordered slabs must not share one
};
static_assert(sizeof(IpcDesc) <= LBM_GPU_IPC_DESC_BYTES, "descriptor too large");
constexpr uint32_t kIpcMagic = 0x4c424d31u;   // "LBM1"

enum SyncWord { kFlagFromBelow = 0, kFlagFromAbove = 1, kBoundaryDone = 2, kScratch0 = 3, kScratch1 = 4,
                kScratch2 = 5, kGridBarrier = 6, kAbort = 7, kSyncError = 8 /* must follow kAbort */,
                kAvScratch = 16 /* kAvWords words */, kSyncWords = 16 + 128 };
constexpr int kTb2MinRows = 8;                             // per slab, for the two-step kernel
constexpr int kAvWords = LBM_AV_STRIDE * LBM_AV_SLOTS;     // words of one step's |u| sums (1 KiB)
constexpr int kMaxSegmentSteps = 1 << 16;                  // 64 MiB of sums per run segment

struct GridBase {
  virtual ~GridBase() {}
  virtual bool is_f64() const = 0;
};

template <typename real>
struct Slab {
  int device = 0;
  long long row0 = 0;   // first global row
  int rows = 0;         // local rows
  int accel_row = LBM_NO_ROW;   // local row of global row ny-2, or LBM_NO_ROW
  char* base = nullptr; // lattice[2] | side[2] | mask (private to this GPU)
  size_t bytes = 0;
  size_t off_lattice[2] = {0, 0}, off_side[2] = {0, 0}, off_mask = 0;
  real* lattice[2] = {nullptr, nullptr};
  real* side[2] = {nullptr, nullptr};
  uint32_t* mask = nullptr;
  // window: ghost rows [parity 2][direction 2][depth 2][9 planes][pitch], ghost mask rows, sync words.
  // The only memory neighbours (other GPUs / processes) read or write.
  char* win = nullptr;
  size_t win_bytes = 0, off_sync = 0, off_gmask = 0;
  unsigned long long* sync = nullptr;
  uint32_t* ghost_mask = nullptr;     // inside the window: mask words of row -1, then of row `rows`
  unsigned long long* av = nullptr;   // per step LBM_AV_SLOTS x {sum of low halves, sum of high halves}
  unsigned long long* av_compact = nullptr;
  size_t av_cap = 0;                  // steps
  void* staging = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;            // host <-> staging copies, overlapped with the kernels on `stream`
  cudaEvent_t stage_filled[2] = {nullptr, nullptr};   // a staging half is ready for its consumer
  cudaEvent_t stage_drained[2] = {nullptr, nullptr};  // ... and has been consumed
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t step_ev[2] = {nullptr, nullptr};   // "step t finished", alternating by parity
  long long free_cells = 0;
  // neighbour views: the windows of the slab above (holds global row row0+rows) and below
  char* up_win = nullptr;
  char* dn_win = nullptr;
  unsigned long long* up_flag = nullptr;      // neighbour above's kFlagFromBelow
  unsigned long long* dn_flag = nullptr;      // neighbour below's kFlagFromAbove
  unsigned long long* up_abort = nullptr;     // the neighbours' kAbort words
  unsigned long long* dn_abort = nullptr;
  uint32_t* up_gmask = nullptr;               // neighbour above's copy of the mask of its row -1
  uint32_t* dn_gmask = nullptr;               // neighbour below's copy of the mask of its row `rows`
  void* ipc_mapped[2] = {nullptr, nullptr};   // pointers to close on destroy
  bool pooled = false;                        // base/staging came from the stream-ordered pool
  alignas(64) CUtensorMap tmap[2];            // K1c: (x, row, plane) view of lattice[0], lattice[1], row boxes
  alignas(64) CUtensorMap tmap_halo[2];       // same view, 4-element boxes (the x halo of a tile)
};

template <typename real> struct ParamT;
template <> struct ParamT<float> { typedef lbm_param type; };
template <> struct ParamT<double> { typedef lbm_param_f64 type; };

template <typename real>
class Grid : public GridBase {
 public:
  typedef typename ParamT<real>::type Param;
  Param prm;
  unsigned flags = 0;
  int kernel = 0;
  int persistent_vec = 4;      // cells per thread of the persistent kernel (4, or 1 for tiny grids)
  int cluster_ctas = 0;        // K6: CTAs of the cluster (16 or 8), 0 = not applicable
  int cluster_rows = 0;        // K6: rows per CTA
  size_t cluster_smem = 0;     // K6: dynamic shared memory per CTA
  int pitch = 0, mask_pitch = 0;
  bool slab_mode = false;      // one process per GPU: neighbours are other processes
  bool connected = false;      // neighbour views are set
  bool multi = false;          // more than one slab in the whole grid
  bool use_flags = false;      // cross-slab ordering by device-side flags (else: CUDA events)
  long long global_free_cells = -1;
  long long steps_done = 0;
  long long passes_done = 0;   // kernel passes (one or two timesteps each): the unit of the flag protocol
  int cur = 0;                 // lattice buffer / window parity holding the current state (flips per pass)
  bool tb2 = false;            // two timesteps per pass (K7) wherever two steps remain
  int tb2_strips = 0, tb2_wout = 0, tb2_seg_rows = 0, tb2_span = 0, tb2_threads = 0;
  int tb2_tail_rows = 16, tb2_tail_segs = 0;   // K7: height and number of the short segments that end a launch
  int pairs_tile_rows = 0;     // K9: rows per block
  bool failed = false;         // a neighbour never arrived: the lattice contents are void
  unsigned long long timeout_ns = 10000000000ULL;
  long long launches = 0;
  double last_run_ms = 0.0, last_step_ms = 0.0;
  std::vector<Slab<real>> slabs;

  bool is_f64() const override { return sizeof(real) == 8; }

  ~Grid() override {
    for (auto& s : slabs) {
      cudaSetDevice(s.device);
      if (s.stream) cudaStreamSynchronize(s.stream);
      for (int i = 0; i < 2; i++)
        if (s.ipc_mapped[i]) cudaIpcCloseMemHandle(s.ipc_mapped[i]);
      pool_free(s.av, s);
      pool_free(s.staging, s);
      pool_free(s.base, s);
      if (s.stream) cudaStreamSynchronize(s.stream);
      window_free(s);
      if (s.ev0) cudaEventDestroy(s.ev0);
      if (s.ev1) cudaEventDestroy(s.ev1);
      for (int i = 0; i < 2; i++) {
        if (s.stage_filled[i]) cudaEventDestroy(s.stage_filled[i]);
        if (s.stage_drained[i]) cudaEventDestroy(s.stage_drained[i]);
      }
      if (s.copy_stream) { cudaStreamSynchronize(s.copy_stream); cudaStreamDestroy(s.copy_stream); }
      for (int i = 0; i < 2; i++)
        if (s.step_ev[i]) cudaEventDestroy(s.step_ev[i]);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
  }

  // LBM_GPU_POOL: the lattice and the staging buffer come from the device's stream-ordered
  // memory pool with an unlimited release threshold, so destroying a lattice and creating
  // the next one of a similar size (parameter sweeps, bench.py's end-to-end leg) reuses
  // the pages instead of paying cudaFree + cudaMalloc of ~20 GB each time (25-80 ms).
  // Not the default: the pool's first growth is slower than one cudaMalloc (0.4 s for
  // 19 GB), which a one-shot CLI run would pay for nothing, and the memory stays with the
  // process after lbm_gpu_destroy.  The window is always a plain allocation (CUDA IPC).
  bool use_pool() const { return (flags & LBM_GPU_POOL) != 0; }
  void pool_alloc(void** p, size_t bytes, Slab<real>& s) {
    if (use_pool()) {
      int supported = 0;
      CK(cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, s.device));
      if (supported) {
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, s.device));
        unsigned long long keep = ~0ULL;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        CK(cudaMallocAsync(p, bytes, s.stream));
        s.pooled = true;
        return;
      }
    }
    CK(cudaMalloc(p, bytes));
  }
  static void pool_free(void* p, Slab<real>& s) {
    if (!p) return;
    if (s.pooled) cudaFreeAsync(p, s.stream);
    else cudaFree(p);
  }
  // The window must be a plain cudaMalloc allocation (CUDA IPC), and a plain cudaFree
  // is a device-wide synchronisation that was measured to take up to 0.6 s when it follows
  // large transfers.  With LBM_GPU_POOL a destroyed lattice parks its window in a small
  // per-process cache instead and the next lattice of the same width picks it up.
  void window_alloc(Slab<real>& s) {
    if (use_pool()) {
      std::lock_guard<std::mutex> lock(g_window_mutex);
      for (size_t i = 0; i < g_window_cache.size(); i++)
        if (g_window_cache[i].device == s.device && g_window_cache[i].bytes == s.win_bytes) {
          s.win = g_window_cache[i].ptr;
          g_window_cache.erase(g_window_cache.begin() + i);
          return;
        }
    }
    CK(cudaMalloc((void**)&s.win, s.win_bytes));
  }
  void window_free(Slab<real>& s) {
    if (!s.win) return;
    if (use_pool()) {
      std::lock_guard<std::mutex> lock(g_window_mutex);
      g_window_cache.push_back({s.device, s.win_bytes, s.win});
    } else {
      cudaFree(s.win);
    }
    s.win = nullptr;
  }

  // ---------------------------------------------------------------- allocation ----
  void alloc_slab(Slab<real>& s) {
    CK(cudaSetDevice(s.device));
    const long long plane = (long long)s.rows * pitch;
    const size_t lat = (size_t)9 * plane * sizeof(real);
    const size_t side = (size_t)6 * pitch * sizeof(real);
    const size_t maskb = (size_t)s.rows * mask_pitch * sizeof(uint32_t);
    size_t off = 0;
    for (int i = 0; i < 2; i++) { s.off_lattice[i] = off; off = round_up(off + lat, 256); }
    for (int i = 0; i < 2; i++) { s.off_side[i] = off; off = round_up(off + side, 256); }
    s.off_mask = off; off = round_up(off + maskb, 256);
    s.bytes = off;
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CK(cudaEventCreateWithFlags(&s.stage_filled[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&s.stage_drained[i], cudaEventDisableTiming));
    }
    pool_alloc((void**)&s.base, s.bytes, s);
    for (int i = 0; i < 2; i++) {
      s.lattice[i] = (real*)(s.base + s.off_lattice[i]);
      s.side[i] = (real*)(s.base + s.off_side[i]);
    }
    s.mask = (uint32_t*)(s.base + s.off_mask);
    // the window gets its own allocation, a multiple of 2 MiB so that it never shares a
    // driver block with anything else (it is exported over CUDA IPC)
    s.off_gmask = round_up((size_t)LBM_GHOST_PLANE_ROWS * pitch * sizeof(real), 256);
    s.off_sync = round_up(s.off_gmask + (size_t)2 * mask_pitch * sizeof(uint32_t), 256);
    s.win_bytes = round_up(s.off_sync + kSyncWords * sizeof(unsigned long long), 2u << 20);
    window_alloc(s);
    s.ghost_mask = (uint32_t*)(s.win + s.off_gmask);
    s.sync = (unsigned long long*)(s.win + s.off_sync);
    CK(cudaEventCreate(&s.ev0));
    CK(cudaEventCreate(&s.ev1));
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&s.step_ev[i], cudaEventDisableTiming));
    CK(cudaMemsetAsync(s.base + s.off_side[0], 0, s.bytes - s.off_side[0], s.stream));
    CK(cudaMemsetAsync(s.win, 0, s.win_bytes, s.stream));
    void* st = nullptr;
    pool_alloc(&st, kStagingBytes, s);
    s.staging = st;
  }

  // K1c: one 3-D tensor map (x, row, plane) per lattice buffer, box = one row of a tile
  void make_tensor_maps(Slab<real>& s) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) throw CudaError{"cuTensorMapEncodeTiled is not available in this driver"};
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)s.rows, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(real), (cuuint64_t)plane_stride(s) * sizeof(real)};
    const cuuint32_t box[3] = {LBM_TMA_BOX, 1, 1};
    const cuuint32_t hbox[3] = {4, 1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int b = 0; b < 2; b++) {
      for (int h = 0; h < 2; h++) {
        const CUresult r = ((EncodeFn)fn)(h ? &s.tmap_halo[b] : &s.tmap[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s.lattice[b],
                                          dims, strides, h ? hbox : box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          char msg[96];
          snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
          throw CudaError{msg};
        }
      }
    }
  }

  // ghost rows of a window: parity b, direction d (0 = "from below", 1 = "from above")
  real* win_section(char* win, int b, int d) const { return (real*)win + lbm::ghost_offset(pitch, b, d); }

  long long plane_stride(const Slab<real>& s) const { return (long long)s.rows * pitch; }

  void setup_geometry(int nx) {
    pitch = (int)round_up(nx, 32);
    mask_pitch = pitch / 32;
    if (const char* e = getenv("LBM_GPU_SYNC_TIMEOUT_MS")) {      // longest wait for a neighbouring slab
      const double ms = atof(e);
      if (ms > 0) timeout_ns = (unsigned long long)(ms * 1e6);
    }
    if (flags & LBM_GPU_KERNEL_SCALAR) kernel = LBM_GPU_KERNEL_SCALAR;
    else if (flags & LBM_GPU_KERNEL_VEC4) kernel = LBM_GPU_KERNEL_VEC4;
    else if (flags & LBM_GPU_KERNEL_PERSISTENT) kernel = LBM_GPU_KERNEL_PERSISTENT;
    else if (flags & LBM_GPU_KERNEL_TB2) kernel = LBM_GPU_KERNEL_VEC4;         // decided in choose_tb2()
    else if (flags & LBM_GPU_KERNEL_PAIRS) kernel = LBM_GPU_KERNEL_VEC4;       // decided in choose_pairs()
    else if (flags & LBM_GPU_KERNEL_CLUSTER) kernel = LBM_GPU_KERNEL_VEC4;     // decided in choose_kernel()
    else kernel = LBM_GPU_KERNEL_VEC4;                         // refined in choose_kernel()
    if (flags & LBM_GPU_KERNEL_TMA) {
      if (sizeof(real) != 4) throw CudaError{"the TMA kernel is built for single precision only"};
      kernel = LBM_GPU_KERNEL_TMA;
    }
  }

  // launch shape shared by the step kernels: blockDim (bx, by), tiles of bx*vec x by cells
  void tile_shape(int& vec, int& bx, int& by) const {
    vec = (kernel == LBM_GPU_KERNEL_SCALAR || (kernel == LBM_GPU_KERNEL_PERSISTENT && persistent_vec == 1)) ? 1 : 4;
    const int nxv = (prm.nx + vec - 1) / vec;
    bx = (int)std::min<long long>(LBM_BLOCK_THREADS, round_up(nxv, 32));
    if (kernel == LBM_GPU_KERNEL_TMA) bx = LBM_BLOCK_THREADS;     // one row of LBM_TMA_TILE cells per block
    by = LBM_BLOCK_THREADS / bx;
  }

  // blocks of lbm_steps_persistent that can be resident at once on the slab's GPU
  const void* persistent_fn() const {
    const bool strict = (flags & LBM_GPU_STRICT) != 0;
    if (persistent_vec == 1)
      return strict ? (const void*)lbm::lbm_steps_persistent<real, true, 1> : (const void*)lbm::lbm_steps_persistent<real, false, 1>;
    return strict ? (const void*)lbm::lbm_steps_persistent<real, true, 4> : (const void*)lbm::lbm_steps_persistent<real, false, 4>;
  }

  int persistent_capacity(const Slab<real>& s) {
    int per_sm = 0, sms = 0, coop = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_fn(), LBM_BLOCK_THREADS, 0));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device));
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s.device));
    return coop ? per_sm * sms : 0;
  }

  // After the slabs exist: grids small enough to be launch-latency bound (they live in
  // L2) run all their steps in one persistent cooperative kernel.
  const void* cluster_fn() const {
    return (flags & LBM_GPU_STRICT) ? (const void*)lbm::lbm_steps_cluster<true> : (const void*)lbm::lbm_steps_cluster<false>;
  }

  // K6 applies to a single fp32 slab whose double-buffered lattice fits in the shared memory
  // of one cluster (16 CTAs if the device schedules such a cluster, else 8).
  bool cluster_fits() {
    if (sizeof(real) != 4 || slabs.size() != 1 || slab_mode) return false;
    Slab<real>& s = slabs[0];
    CK(cudaSetDevice(s.device));
    int max_smem = 0;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s.device));
    for (int c : {16, 8}) {
      const int rows = (s.rows + c - 1) / c;
      const size_t smem = (size_t)2 * 9 * rows * prm.nx * sizeof(float);
      if (smem > (size_t)max_smem) continue;
      if (cudaFuncSetAttribute(cluster_fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
          cudaFuncSetAttribute(cluster_fn(), cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(c);
      cfg.blockDim = dim3(LBM_CLUSTER_THREADS);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, cluster_fn(), &cfg) != cudaSuccess || nclusters < 1) {
        cudaGetLastError();
        continue;
      }
      cluster_ctas = c;
      cluster_rows = rows;
      cluster_smem = smem;
      return true;
    }
    return false;
  }

  void choose_kernel() {
    const bool forced = flags & (LBM_GPU_KERNEL_SCALAR | LBM_GPU_KERNEL_VEC4 | LBM_GPU_KERNEL_PERSISTENT |
                                 LBM_GPU_KERNEL_TMA | LBM_GPU_KERNEL_CLUSTER | LBM_GPU_KERNEL_TB2 |
                                 LBM_GPU_KERNEL_PAIRS);
    // K6 is opt-in only: 16 SMs doing all the arithmetic are no faster than K5 spreading it
    // over the whole chip (profiles/r01_small_grids.md).
    if (flags & LBM_GPU_KERNEL_CLUSTER) {
      if (cluster_fits()) { kernel = LBM_GPU_KERNEL_CLUSTER; return; }
      throw CudaError{"the cluster kernel needs a single-GPU fp32 lattice that fits in one cluster's shared memory"};
    }
    const bool want = (kernel == LBM_GPU_KERNEL_PERSISTENT);
    if (!want && (forced || slabs.size() != 1 || slab_mode)) return;
    if (slabs.size() != 1 || slab_mode) throw CudaError{"the persistent kernel handles a single slab only"};
    CK(cudaSetDevice(slabs[0].device));
    const int saved = kernel;
    kernel = LBM_GPU_KERNEL_PERSISTENT;
    auto tiles_for = [&](int pv) {
      persistent_vec = pv;
      int vec, bx, by;
      tile_shape(vec, bx, by);
      return (long long)(((prm.nx + vec - 1) / vec + bx - 1) / bx) * ((slabs[0].rows + by - 1) / by);
    };
    // tiny grids: one cell per thread (4x the threads on the step's dependent-latency chain)
    // as long as that does not put more than kPersistScalarBlocks blocks on the grid barrier
    long long tiles = tiles_for(1);
    int cap = persistent_capacity(slabs[0]);
    long long scalar_limit = std::min<long long>(cap, kPersistScalarBlocks);
    if (const char* e = getenv("LBM_PERSIST_VEC"))      // measurement knob of profiles/r01_small_grids.md
      scalar_limit = (atoi(e) == 1) ? cap : 0;
    if (tiles > scalar_limit) {
      tiles = tiles_for(4);
      cap = persistent_capacity(slabs[0]);
    }
    if (cap < 1) {
      if (want) throw CudaError{"cooperative launch is not available on this device"};
      kernel = saved;
    } else if (!want && (long long)prm.nx * slabs[0].rows > kPersistMaxCells) {
      kernel = saved;             // beyond L2: bandwidth bound, one launch per pass is free
    }
  }

  // K7 (two timesteps per pass) needs fp32, a width the 16-byte bulk copies can cut (multiple
  // of 4), slabs tall enough for two-row ghost zones, and no
  // kernel forced by the caller.  Every slab of the grid must come to the same answer: the
  // ghost-row pushes and the flag protocol count passes, not timesteps.
  bool tb2_possible(int rows_min) const {
    if (sizeof(real) != 4) return false;
    if (flags & (LBM_GPU_KERNEL_SCALAR | LBM_GPU_KERNEL_VEC4 | LBM_GPU_KERNEL_PERSISTENT | LBM_GPU_KERNEL_TMA |
                 LBM_GPU_KERNEL_CLUSTER | LBM_GPU_KERNEL_PAIRS)) return false;
    if (getenv("LBM_GPU_NO_TB2") && getenv("LBM_GPU_NO_TB2")[0] == '1') return false;
    return (prm.nx % 4 == 0) && prm.nx >= 32 && rows_min >= kTb2MinRows;
  }

  void enable_tb2() {
    if constexpr (sizeof(real) == 4) {
      for (auto& s : slabs) {
        CK(cudaSetDevice(s.device));
        for (const void* fn : {(const void*)lbm::lbm_step2_tb<false, false>, (const void*)lbm::lbm_step2_tb<false, true>,
                               (const void*)lbm::lbm_step2_tb<true, false>, (const void*)lbm::lbm_step2_tb<true, true>})
          CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(lbm::Tb2Smem)));
      }
      if (prm.nx + 2 * LBM_TB2_PAD <= LBM_TB2_SPAN) {      // narrow grid: one strip, the whole width plus halo
        tb2_strips = 1;
        tb2_wout = prm.nx;
        tb2_span = prm.nx + 2 * LBM_TB2_PAD;
      } else {
        tb2_strips = (prm.nx + LBM_TB2_MAX_WOUT - 1) / LBM_TB2_MAX_WOUT;
        tb2_wout = (int)round_up((prm.nx + tb2_strips - 1) / tb2_strips, 4);
        tb2_span = LBM_TB2_SPAN;
      }
      tb2_threads = (int)round_up(tb2_span / 4, 32);
      // Segment height.  Every segment recomputes two rows of the first sub-step, so tall is
      // cheap; but the blocks of a launch run in "waves" of the resident blocks and a last wave
      // that is nearly empty leaves the SMs idle (16384 rows in 64-row segments: 19.03 waves).
      // Score each height by (rows / (rows + 2)) x (waves / ceil(waves)) and take the best; grids
      // with few blocks prefer short segments so that there are several waves at all.
      int sms = 148, rows_max = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, slabs[0].device);
      for (auto& s : slabs) rows_max = std::max(rows_max, s.rows);
      const double resident = (double)LBM_TB2_MIN_BLOCKS * sms;
      double best = -1.0;
      tb2_seg_rows = 64;
      for (int h = 16; h <= 128; h++) {
        const double waves = (double)tb2_strips * ((rows_max + h - 1) / h) / resident;
        const double score = (double)h / (h + 2) * waves / std::ceil(waves) * (waves >= 4.0 ? 1.0 : 0.25 * waves);
        if (score > best + 1e-9) { best = score; tb2_seg_rows = h; }
      }
      if (const char* e = getenv("LBM_TB2_SEG_ROWS")) tb2_seg_rows = std::max(2, atoi(e));   // tuning knob
      // A block lives (segment height) x ~2.6 us; the last blocks of a launch to be scheduled
      // leave the other SMs idle for about half of that.  So the launch ENDS with short
      // segments -- one set of resident blocks' worth, given the highest block indices -- and the
      // SMs drain together (tall slabs only: the short segments pay 2 halo rows per 16).
      tb2_tail_segs = (int)((resident + tb2_strips - 1) / tb2_strips) + 1;
      if (const char* e = getenv("LBM_TB2_TAIL_ROWS")) tb2_tail_rows = atoi(e);                // tuning knob, 0 = off
      if (tb2_tail_rows < 2) tb2_tail_segs = 0;
      tb2 = true;
      kernel = LBM_GPU_KERNEL_TB2;
    }
  }

  // segments of a slab for K7: n_big segments of equal height, then (tall slabs) n_tail short
  // ones; the last segment keeps at least two rows (the rows that push into the neighbour
  // above must sit in an edge segment)
  struct Tb2Segments { int seg_rows, n_big, big_rows, seg_rows_tail, n_tail; };
  Tb2Segments tb2_segments(int rows) const {
    Tb2Segments g;
    // measured (profiles/r02_kernel_variants.md): +1.2 % on 16384 rows, nothing on 8192, -2 % on 2048
    const long long min_rows = (long long)tb2_tail_segs * tb2_tail_rows * (tb2_tail_rows >= 8 ? 32 : 4);
    g.n_tail = (tb2_tail_segs > 0 && min_rows <= rows) ? tb2_tail_segs : 0;
    g.seg_rows_tail = tb2_tail_rows;
    g.big_rows = rows - g.n_tail * g.seg_rows_tail;
    int nsegs = std::max(1, (g.big_rows + tb2_seg_rows - 1) / tb2_seg_rows);
    for (;;) {
      g.seg_rows = (g.big_rows + nsegs - 1) / nsegs;
      nsegs = (g.big_rows + g.seg_rows - 1) / g.seg_rows;
      if (nsegs == 1 || g.n_tail > 0 || g.big_rows - (nsegs - 1) * g.seg_rows >= 2) break;
      nsegs--;
    }
    g.n_big = nsegs;
    return g;
  }

  // K9 (two timesteps per grid barrier) for the small grids the persistent kernel K5 would
  // take: single GPU, fp32, nx a multiple of 4 and at most 256, every block resident.
  const void* pairs_fn() const {
    if constexpr (sizeof(real) == 4)
      return (flags & LBM_GPU_STRICT) ? (const void*)lbm::lbm_steps_pairs<true> : (const void*)lbm::lbm_steps_pairs<false>;
    return nullptr;
  }
  dim3 pairs_block() const { return dim3((unsigned)round_up(prm.nx / 4, 32), (unsigned)(pairs_tile_rows + 2)); }
  size_t pairs_smem() const { return (size_t)(pairs_tile_rows + 2) * 9 * prm.nx * sizeof(float); }

  void choose_pairs() {
    const bool want = (flags & LBM_GPU_KERNEL_PAIRS) != 0;
    const bool small = (kernel == LBM_GPU_KERNEL_PERSISTENT) && !(flags & LBM_GPU_KERNEL_PERSISTENT);
    static const bool off = getenv("LBM_GPU_NO_PAIRS") && getenv("LBM_GPU_NO_PAIRS")[0] == '1';
    if (!want && (!small || off)) return;
    bool ok = sizeof(real) == 4 && slabs.size() == 1 && !slab_mode && prm.nx % 4 == 0 && prm.nx >= 4 &&
              prm.nx <= LBM_PAIRS_MAX_NX && slabs[0].rows >= 4;
    if (ok) {
      Slab<real>& s = slabs[0];
      CK(cudaSetDevice(s.device));
      int sms = 0, coop = 0, per_sm = 0;
      CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.device));
      CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s.device));
      // one block per SM if the grid allows it: the fewest rows per block that still fits
      pairs_tile_rows = std::min(LBM_PAIRS_MAX_TILE_ROWS, std::max(1, (s.rows + sms - 1) / sms));
      if (const char* e = getenv("LBM_PAIRS_TILE_ROWS"))                      // tuning knob
        pairs_tile_rows = std::min(LBM_PAIRS_MAX_TILE_ROWS, std::max(1, atoi(e)));
      pairs_tile_rows = std::min(pairs_tile_rows, s.rows - 2);
      const dim3 b = pairs_block();
      ok = coop && cudaFuncSetAttribute(pairs_fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pairs_smem()) == cudaSuccess &&
           cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pairs_fn(), (int)(b.x * b.y), pairs_smem()) == cudaSuccess &&
           (long long)per_sm * sms >= (s.rows + pairs_tile_rows - 1) / pairs_tile_rows;
      cudaGetLastError();
    }
    if (ok) kernel = LBM_GPU_KERNEL_PAIRS;
    else if (want) throw CudaError{"the pairs kernel needs a single-GPU fp32 lattice with nx a multiple of 4 and <= 256 whose "
                                   "blocks are all resident"};
  }

  // single-process form: decide once all slabs exist
  void choose_tb2() {
    int rows_min = slabs[0].rows;
    for (auto& s : slabs) rows_min = std::min(rows_min, s.rows);
    const bool want = (flags & LBM_GPU_KERNEL_TB2) != 0;
    const bool big = (kernel == LBM_GPU_KERNEL_VEC4);        // not taken by the persistent kernel
    if (tb2_possible(rows_min) && (want || big)) enable_tb2();
    else if (want) throw CudaError{"the two-step kernel needs fp32, nx a multiple of 4 and >= 32, and >= 8 rows per slab"};
  }

  // ------------------------------------------------------------- lattice input ----
  // cells: AoS rows for this slab (or NULL -> rest state); obstacles: rows for this slab
  void load_slab(Slab<real>& s, const real* cells_aos, const void* obstacles, int cur) {
    CK(cudaSetDevice(s.device));
    const int nx = prm.nx;
    if (cells_aos) {
      upload_cells(s, cells_aos, cur);
    } else {
      const real w0 = prm.density * (real)4 / (real)9;     // d2q9-bgk.c:2802-2804
      const real w1 = prm.density / (real)9;
      const real w2 = prm.density / (real)36;
      const long long total = plane_stride(s);
      lbm::lbm_init_rest<real><<<(unsigned)((total + 255) / 256), 256, 0, s.stream>>>(
          s.lattice[cur], plane_stride(s), pitch, nx, s.rows, w0, w1, w2);
      CK(cudaGetLastError());
      launches++;
    }
    // mask
    unsigned long long* counter = s.sync + kScratch0;
    CK(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s.stream));
    if (obstacles) {
      // host rows -> one half of the staging buffer (copy stream) -> packed into the mask
      // (kernel stream); the copy of chunk i+1 overlaps the packing of chunk i, one
      // synchronisation per call
      const bool bits = (flags & LBM_GPU_OBST_BITS) != 0;
      const int wpr = (nx + 31) / 32;
      const size_t row_bytes = bits ? (size_t)wpr * 4 : (size_t)nx * 4;
      const size_t half = kStagingBytes / 2;
      if (row_bytes > half) throw CudaError{"nx too large for the obstacle staging buffer"};
      const int chunk_rows = (int)(half / row_bytes);
      CK(cudaEventRecord(s.stage_drained[0], s.stream));       // orders the first copies behind the memsets above
      CK(cudaEventRecord(s.stage_drained[1], s.stream));
      int i = 0;
      for (int r = 0; r < s.rows; r += chunk_rows, i++) {
        const int n = std::min(chunk_rows, s.rows - r);
        const int b = i & 1;
        char* st = (char*)s.staging + (size_t)b * half;
        CK(cudaStreamWaitEvent(s.copy_stream, s.stage_drained[b], 0));
        CK(cudaMemcpyAsync(st, (const char*)obstacles + (size_t)r * row_bytes, (size_t)n * row_bytes,
                           cudaMemcpyHostToDevice, s.copy_stream));
        CK(cudaEventRecord(s.stage_filled[b], s.copy_stream));
        CK(cudaStreamWaitEvent(s.stream, s.stage_filled[b], 0));
        if (bits) {
          const long long words = (long long)n * wpr;
          lbm::lbm_copy_mask_bits<<<(unsigned)((words + 255) / 256), 256, 0, s.stream>>>(
              (const uint32_t*)st, s.mask, mask_pitch, nx, r, n, counter);
        } else {
          const long long threads = (long long)n * wpr * 32;
          lbm::lbm_pack_mask<<<(unsigned)((threads + 255) / 256), 256, 0, s.stream>>>(
              (const int*)st, s.mask, mask_pitch, nx, r, n, counter);
        }
        CK(cudaGetLastError());
        launches++;
        CK(cudaEventRecord(s.stage_drained[b], s.stream));
      }
    }
    unsigned long long blocked = 0;
    CK(cudaMemcpyAsync(&blocked, counter, sizeof blocked, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    s.free_cells = (long long)s.rows * nx - (long long)blocked;
  }

  void upload_cells(Slab<real>& s, const real* cells_aos, int cur) {
    CK(cudaSetDevice(s.device));
    const int nx = prm.nx;
    const size_t row_bytes = (size_t)nx * 9 * sizeof(real);
    const int chunk_rows = (int)std::max<size_t>(1, kStagingBytes / row_bytes);
    if (row_bytes > kStagingBytes) throw CudaError{"nx too large for the AoS staging buffer"};
    for (int r = 0; r < s.rows; r += chunk_rows) {
      const int n = std::min(chunk_rows, s.rows - r);
      const long long ncells = (long long)n * nx;
      CK(cudaMemcpyAsync(s.staging, (const char*)cells_aos + (size_t)r * row_bytes, (size_t)n * row_bytes,
                         cudaMemcpyHostToDevice, s.stream));
      lbm::lbm_aos_to_soa<real><<<(unsigned)((ncells * 9 + 255) / 256), 256, 0, s.stream>>>(
          (const real*)s.staging, s.lattice[cur], plane_stride(s), pitch, nx, r, ncells);
      CK(cudaGetLastError());
      launches++;
      CK(cudaStreamSynchronize(s.stream));
    }
  }

  // ------------------------------------------------------------------ wiring ----
  // all slabs live in this process: neighbours are reached through peer access
  void connect_local() {
    const int n = (int)slabs.size();
    for (int i = 0; i < n; i++) {
      Slab<real>& s = slabs[i];
      Slab<real>& up = slabs[(i + 1) % n];
      Slab<real>& dn = slabs[(i + n - 1) % n];
      CK(cudaSetDevice(s.device));
      for (Slab<real>* o : {&up, &dn}) {
        if (o->device != s.device) {
          int can = 0;
          CK(cudaDeviceCanAccessPeer(&can, s.device, o->device));
          if (!can) throw CudaError{"peer access between the selected GPUs is not available"};
          cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
          cudaGetLastError();
        }
      }
      s.up_win = up.win;
      s.dn_win = dn.win;
      s.up_flag = up.sync + kFlagFromBelow;
      s.dn_flag = dn.sync + kFlagFromAbove;
      s.up_abort = up.sync + kAbort;
      s.dn_abort = dn.sync + kAbort;
      s.up_gmask = up.ghost_mask;
      s.dn_gmask = dn.ghost_mask + mask_pitch;
    }
    multi = n > 1;
    // Slabs of one process are ordered by the device-side flag protocol (edge blocks wait,
    // interior blocks never do) whenever every slab has a GPU of its own; slabs that share a
    // GPU (tests on a one-GPU box) cannot wait on one another inside kernels and are ordered
    // by CUDA events, as is everything when LBM_GPU_SYNC_EVENTS asks for it.
    bool distinct = true;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++)
        if (slabs[i].device == slabs[j].device) distinct = false;
    if ((flags & LBM_GPU_SYNC_FLAGS) && multi && !distinct)
      throw CudaError{"LBM_GPU_SYNC_FLAGS needs every slab on its own GPU (kernels that wait on "
                      "one another must not share a device)"};
    use_flags = multi && distinct && !(flags & LBM_GPU_SYNC_EVENTS);
    connected = true;
  }

  void prepare() {
    if (!connected) throw CudaError{"lattice is not connected to its neighbours yet"};
    for (auto& s : slabs) {
      CK(cudaSetDevice(s.device));
      lbm::PrepareArgs<real> a;
      a.cur = s.lattice[cur];
      a.side_cur = s.side[cur];
      a.mask = s.mask;
      a.push_up = win_section(s.up_win, cur, 0);
      a.push_dn = win_section(s.dn_win, cur, 1);
      a.mask_up = s.up_gmask;
      a.mask_dn = s.dn_gmask;
      a.plane_stride = plane_stride(s);
      a.nx = prm.nx; a.rows = s.rows; a.pitch = pitch; a.mask_pitch = mask_pitch; a.accel_row = s.accel_row;
      a.deep = tb2 ? 1 : 0;
      a.aw1 = prm.density * prm.accel / (real)9;      // d2q9-bgk.c:230-231
      a.aw2 = prm.density * prm.accel / (real)36;
      lbm::lbm_prepare<real><<<(std::max(prm.nx, mask_pitch) + 127) / 128, 128, 0, s.stream>>>(a);
      CK(cudaGetLastError());
      launches++;
    }
    for (auto& s : slabs) { CK(cudaSetDevice(s.device)); CK(cudaStreamSynchronize(s.stream)); }
  }

  // -------------------------------------------------------------------- run ----
  template <bool STRICT, bool MULTI>
  void launch_step(Slab<real>& s, const lbm::StepArgs<real>& a, dim3 grid, dim3 block, int src) {
    if (kernel == LBM_GPU_KERNEL_SCALAR)
      lbm::lbm_step_scalar<real, STRICT, MULTI><<<grid, block, 0, s.stream>>>(a);
    else if (kernel == LBM_GPU_KERNEL_TMA)
      launch_tma<STRICT, MULTI>(s, a, grid, block, src);
    else
      launch_vec4<STRICT, MULTI>(s, a, grid, block);
  }
  // Step kernels go out with programmatic stream serialization (PDL): the next pass's grid
  // is set up while the current one drains, which hides the launch gap on mid-size grids.
  static bool use_pdl() {
    static const bool pdl = !(getenv("LBM_GPU_NO_PDL") && getenv("LBM_GPU_NO_PDL")[0] == '1');
    return pdl;
  }
  template <typename Kernel, typename Args>
  void launch_pdl(Slab<real>& s, Kernel k, const Args& a, dim3 grid, dim3 block, size_t smem) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k, a));
  }
  template <bool STRICT, bool MULTI>
  void launch_vec4(Slab<real>& s, const lbm::StepArgs<real>& a, dim3 grid, dim3 block) {
    launch_pdl(s, lbm::lbm_step_vec4<real, STRICT, MULTI>, a, grid, block, 0);
  }
  template <bool STRICT, bool MULTI>
  void launch_tma(Slab<real>& s, const lbm::StepArgs<real>& a, dim3 grid, dim3 block, int src) {
    if constexpr (sizeof(real) == 4)
      lbm::lbm_step_tma<STRICT, MULTI><<<grid, block, 0, s.stream>>>(a, s.tmap[src], s.tmap_halo[src]);
  }
  template <bool STRICT, bool MULTI>
  void launch_tb2(Slab<real>& s, const lbm::Tb2Args& ta, dim3 grid) {
    if constexpr (sizeof(real) == 4)
      launch_pdl(s, lbm::lbm_step2_tb<STRICT, MULTI>, ta, grid, dim3(tb2_threads), sizeof(lbm::Tb2Smem));
  }

  void launch_cluster(int n_steps, bool strict) {
    if constexpr (sizeof(real) == 4) {
      Slab<real>& s = slabs[0];
      CK(cudaSetDevice(s.device));
      lbm::ClusterArgs a;
      a.src = s.lattice[cur];
      a.dst = s.lattice[(cur + n_steps) & 1];
      a.mask = s.mask;
      a.av = s.av;
      a.plane_stride = plane_stride(s);
      a.nx = prm.nx; a.ny = s.rows; a.pitch = pitch; a.mask_pitch = mask_pitch;
      a.rows_per_cta = cluster_rows;
      a.n_steps = n_steps;
      a.omega = prm.omega;
      a.aw1 = prm.density * prm.accel / 9.f;
      a.aw2 = prm.density * prm.accel / 36.f;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cluster_ctas);
      cfg.blockDim = dim3(LBM_CLUSTER_THREADS);
      cfg.dynamicSmemBytes = cluster_smem;
      cfg.stream = s.stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cluster_ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (strict) CK(cudaLaunchKernelEx(&cfg, lbm::lbm_steps_cluster<true>, a));
      else CK(cudaLaunchKernelEx(&cfg, lbm::lbm_steps_cluster<false>, a));
      launches++;
    }
    (void)n_steps; (void)strict;
  }

  // The per-step sums take 1 KiB of device memory per step, so very long runs are cut into
  // segments of kMaxSegmentSteps (one sync + one read-back per segment, no other effect).
  void run(int n_steps, double* sums_out) {
    if (!connected) throw CudaError{"lattice is not connected to its neighbours yet"};
    if (failed) throw CudaError{"an earlier run of this lattice was abandoned (a neighbouring slab never arrived); "
                                "its contents are void"};
    double ms = 0.0;
    const int total = n_steps;
    for (int done = 0; done < total; done += kMaxSegmentSteps) {
      const int n = std::min(kMaxSegmentSteps, total - done);
      run_segment(n, sums_out ? sums_out + done : nullptr);
      ms += last_run_ms;
    }
    if (total > 0) {
      last_run_ms = ms;
      last_step_ms = ms / total;
    }
  }

  // arguments common to every kernel of one pass on one slab: reads buffer `cur`, writes cur^1
  void fill_step_args(lbm::StepArgs<real>& a, Slab<real>& s, int t_in_segment) {
    const int src = cur, dst = cur ^ 1;
    a.src = s.lattice[src];
    a.dst = s.lattice[dst];
    a.side_src = s.side[src];
    a.side_dst = s.side[dst];
    a.mask = s.mask;
    a.av = s.av + (size_t)t_in_segment * kAvWords;
    a.ghost_s = win_section(s.win, src, 0);
    a.ghost_n = win_section(s.win, src, 1);
    a.push_up = win_section(s.up_win, dst, 0);
    a.push_dn = win_section(s.dn_win, dst, 1);
    a.flag_from_below = s.sync + kFlagFromBelow;
    a.flag_from_above = s.sync + kFlagFromAbove;
    a.up_flag = s.up_flag;
    a.dn_flag = s.dn_flag;
    a.boundary_done = s.sync + kBoundaryDone;
    a.abort_word = s.sync + kAbort;
    a.up_abort = s.up_abort;
    a.dn_abort = s.dn_abort;
    a.pass = (unsigned long long)passes_done;
    a.timeout_ns = timeout_ns;
    a.plane_stride = plane_stride(s);
    a.nx = prm.nx; a.rows = s.rows; a.pitch = pitch; a.mask_pitch = mask_pitch;
    a.accel_row = s.accel_row;
    a.tiles_x = a.tiles_y = 0;
    a.edge_tiles = 1;
    a.deep = 0;
    a.omega = prm.omega;
    a.aw1 = prm.density * prm.accel / (real)9;       // d2q9-bgk.c:230-231
    a.aw2 = prm.density * prm.accel / (real)36;
  }

  // one pass over every slab: one timestep (K1a/K1b/K1c), or two (K7) when `two`
  void launch_pass(int t_in_segment, int pass_in_segment, bool two) {
    const bool strict = (flags & LBM_GPU_STRICT) != 0;
    const int nslabs = (int)slabs.size();
    int vec, bx, by;
    tile_shape(vec, bx, by);
    const int nxv = (prm.nx + vec - 1) / vec;
    for (int si = 0; si < nslabs; si++) {
      Slab<real>& s = slabs[si];
      CK(cudaSetDevice(s.device));
      if (multi && !use_flags && pass_in_segment > 0) {
        // slabs sharing a GPU: pass p needs the neighbours' pass p-1 (their pushes into this
        // slab's ghost rows, and their reads of the ghost rows this pass overwrites)
        const int up = (si + 1) % nslabs, dn = (si + nslabs - 1) % nslabs;
        CK(cudaStreamWaitEvent(s.stream, slabs[up].step_ev[(pass_in_segment - 1) & 1], 0));
        if (dn != up) CK(cudaStreamWaitEvent(s.stream, slabs[dn].step_ev[(pass_in_segment - 1) & 1], 0));
      }
      lbm::StepArgs<real> a;
      fill_step_args(a, s, t_in_segment);
      if (two) {
        if constexpr (sizeof(real) == 4) {
          lbm::Tb2Args ta;
          const Tb2Segments sg = tb2_segments(s.rows);
          const int nsegs = sg.n_big + sg.n_tail;
          a.tiles_x = tb2_strips;
          a.tiles_y = nsegs;
          a.edge_tiles = 1;
          a.deep = 1;
          ta.s = a;
          ta.ghost_mask = s.ghost_mask;
          ta.av2 = a.av + kAvWords;
          ta.wout = tb2_wout;
          ta.span = tb2_span;
          ta.seg_rows = sg.seg_rows;
          ta.n_big = sg.n_big;
          ta.big_rows = sg.big_rows;
          ta.seg_rows_tail = sg.seg_rows_tail;
          const dim3 grid((unsigned)((long long)tb2_strips * nsegs));
          if (strict) { if (use_flags) launch_tb2<true, true>(s, ta, grid); else launch_tb2<true, false>(s, ta, grid); }
          else        { if (use_flags) launch_tb2<false, true>(s, ta, grid); else launch_tb2<false, false>(s, ta, grid); }
        }
      } else {
        a.tiles_x = (nxv + bx - 1) / bx;
        a.tiles_y = (s.rows + by - 1) / by;
        a.deep = tb2 ? 1 : 0;                 // a lone step between two-step passes keeps their ghost rows filled
        a.edge_tiles = tb2 ? (2 + by - 1) / by : 1;   // the tile rows that hold the slab's two edge rows
        const dim3 grid((unsigned)((long long)a.tiles_x * a.tiles_y));
        const dim3 block(bx, by);
        if (strict) { if (use_flags) launch_step<true, true>(s, a, grid, block, cur); else launch_step<true, false>(s, a, grid, block, cur); }
        else        { if (use_flags) launch_step<false, true>(s, a, grid, block, cur); else launch_step<false, false>(s, a, grid, block, cur); }
      }
      CK(cudaGetLastError());
      launches++;
      if (multi && !use_flags) CK(cudaEventRecord(s.step_ev[pass_in_segment & 1], s.stream));
    }
    passes_done++;
    cur ^= 1;
  }

  void run_segment(int n_steps, double* sums_out) {
    if (n_steps <= 0) return;
    const bool strict = (flags & LBM_GPU_STRICT) != 0;
    for (auto& s : slabs) {
      CK(cudaSetDevice(s.device));
      if (s.av_cap < (size_t)n_steps) {
        pool_free(s.av, s);
        s.av = nullptr;
        s.av_cap = std::max<size_t>((size_t)n_steps, 1024);
        void* av = nullptr;
        pool_alloc(&av, s.av_cap * (kAvWords + 2) * sizeof(unsigned long long), s);
        s.av = (unsigned long long*)av;
        s.av_compact = s.av + s.av_cap * kAvWords;      // {low, high} per step, filled after the run
      }
      CK(cudaMemsetAsync(s.av, 0, (size_t)n_steps * kAvWords * sizeof(unsigned long long), s.stream));
    }
    for (auto& s : slabs) { CK(cudaSetDevice(s.device)); CK(cudaEventRecord(s.ev0, s.stream)); }

    if (kernel == LBM_GPU_KERNEL_CLUSTER) {
      launch_cluster(n_steps, strict);
      cur = (cur + n_steps) & 1;
    } else if (kernel == LBM_GPU_KERNEL_PERSISTENT) {
      int vec, bx, by;
      tile_shape(vec, bx, by);
      const int nxv = (prm.nx + vec - 1) / vec;
      Slab<real>& s = slabs[0];
      CK(cudaSetDevice(s.device));
      lbm::PersistArgs<real> pa;
      memset(&pa, 0, sizeof pa);
      fill_step_args(pa.s, s, 0);
      lbm::StepArgs<real>& a = pa.s;
      a.tiles_x = (nxv + bx - 1) / bx;
      a.tiles_y = (s.rows + by - 1) / by;
      for (int b = 0; b < 2; b++) { pa.lattice[b] = s.lattice[b]; pa.side[b] = s.side[b]; }
      pa.window = (real*)s.win;
      pa.av = s.av;
      pa.barrier = s.sync + kGridBarrier;
      pa.first_parity = cur;
      pa.n_steps = n_steps;
      pa.n_tiles = a.tiles_x * a.tiles_y;
      const int cap = persistent_capacity(s);
      const int rounds = (pa.n_tiles + cap - 1) / cap;
      const int nblocks = (pa.n_tiles + rounds - 1) / rounds;
      CK(cudaMemsetAsync(pa.barrier, 0, sizeof(unsigned long long), s.stream));
      void* kargs[] = {(void*)&pa};
      CK(cudaLaunchCooperativeKernel(persistent_fn(), dim3(nblocks), dim3(bx, by), kargs, 0, s.stream));
      launches++;
      cur = (cur + n_steps) & 1;
    } else if (kernel == LBM_GPU_KERNEL_PAIRS) {
      const int n_pairs = n_steps / 2;
      if constexpr (sizeof(real) == 4) {
        if (n_pairs > 0) {
          Slab<real>& s = slabs[0];
          CK(cudaSetDevice(s.device));
          lbm::PairsArgs pa;
          memset(&pa, 0, sizeof pa);
          fill_step_args(pa.s, s, 0);
          for (int b = 0; b < 2; b++) { pa.lattice[b] = s.lattice[b]; pa.side[b] = s.side[b]; }
          pa.window = (real*)s.win;
          pa.av = s.av;
          pa.barrier = s.sync + kGridBarrier;
          pa.first_parity = cur;
          pa.n_pairs = n_pairs;
          pa.tile_rows = pairs_tile_rows;
          CK(cudaMemsetAsync(pa.barrier, 0, sizeof(unsigned long long), s.stream));
          void* kargs[] = {(void*)&pa};
          CK(cudaLaunchCooperativeKernel(pairs_fn(), dim3((unsigned)((s.rows + pairs_tile_rows - 1) / pairs_tile_rows)),
                                         pairs_block(), kargs, pairs_smem(), s.stream));
          launches++;
          cur = (cur + n_pairs) & 1;
          passes_done += n_pairs;
        }
      }
      if (n_steps & 1) launch_pass(n_steps - 1, 0, false);     // the odd last step: K1a
    } else {
      int t = 0, p = 0;
      while (t < n_steps) {
        const bool two = tb2 && (n_steps - t >= 2);
        launch_pass(t, p, two);
        t += two ? 2 : 1;
        p++;
      }
      if (use_flags) {
        // end of the run: wait (on the device, bounded) until both neighbours are as far
        for (auto& s : slabs) {
          CK(cudaSetDevice(s.device));
          lbm::StepArgs<real> a;
          fill_step_args(a, s, 0);
          lbm::lbm_wait_neighbours<real><<<1, 32, 0, s.stream>>>(a);
          CK(cudaGetLastError());
          launches++;
        }
      }
    }
    double ms_max = 0.0;
    for (auto& s : slabs) { CK(cudaSetDevice(s.device)); CK(cudaEventRecord(s.ev1, s.stream)); }
    for (auto& s : slabs) {
      CK(cudaSetDevice(s.device));
      CK(cudaStreamSynchronize(s.stream));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
      ms_max = std::max(ms_max, (double)ms);
    }
    last_run_ms = ms_max;
    last_step_ms = ms_max / n_steps;
    steps_done += n_steps;
    if (use_flags) check_sync_errors();
    if (kernel == LBM_GPU_KERNEL_CLUSTER) prepare();      // side row + ghost rows of the new state

    if (sums_out) {
      std::vector<unsigned long long> words((size_t)n_steps * 2);
      std::vector<unsigned __int128> tot(n_steps, 0);
      std::vector<char> bad(n_steps, 0);
      for (auto& s : slabs) {
        CK(cudaSetDevice(s.device));
        lbm::lbm_compact_av<<<(n_steps + 255) / 256, 256, 0, s.stream>>>(s.av, s.av_compact, n_steps);
        CK(cudaGetLastError());
        launches++;
        CK(cudaMemcpyAsync(words.data(), s.av_compact, words.size() * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        for (int t = 0; t < n_steps; t++) {
          const unsigned long long lo = words[2 * (size_t)t], hi = words[2 * (size_t)t + 1];
          if (hi & LBM_NONFINITE_MARK) bad[t] = 1;
          tot[t] += ((unsigned __int128)(hi & ~LBM_NONFINITE_MARK) << 32) + lo;
        }
      }
      for (int t = 0; t < n_steps; t++)
        sums_out[t] = bad[t] ? std::nan("") : (double)((long double)tot[t] / (long double)LBM_FIX_SCALE);
    }
  }

  // The flag protocol gave up (see boundary_wait): say why, in the reference's die() spirit.
  void check_sync_errors() {
    for (auto& s : slabs) {
      CK(cudaSetDevice(s.device));
      unsigned long long w[2] = {0, 0};
      CK(cudaMemcpyAsync(w, s.sync + kAbort, sizeof w, cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
      if (w[0] == 0ULL) continue;
      failed = true;
      const unsigned long long why = w[1] >> 56, pass = w[1] & 0xffffffffffffffULL;
      char msg[256];
      if (why == LBM_SYNC_TIMEOUT)
        snprintf(msg, sizeof msg, "rows [%lld,%lld): a neighbouring slab did not complete pass %llu within %.1f s "
                 "(its process died, or it was asked for a different number of steps); run abandoned",
                 s.row0, s.row0 + s.rows, pass, (double)timeout_ns * 1e-9);
      else
        snprintf(msg, sizeof msg, "rows [%lld,%lld): run abandoned because a neighbouring slab gave up%s",
                 s.row0, s.row0 + s.rows, why == LBM_SYNC_ABORTED ? " (abort word set)" : "");
      throw CudaError{msg};
    }
  }

  long long local_free_cells() const {
    long long n = 0;
    for (auto& s : slabs) n += s.free_cells;
    return n;
  }
  long long divisor() const { return global_free_cells >= 0 ? global_free_cells : local_free_cells(); }

  // ---------------------------------------------------------------- outputs ----
  Slab<real>& slab_of_row(long long grow) {
    for (auto& s : slabs)
      if (grow >= s.row0 && grow < s.row0 + s.rows) return s;
    throw CudaError{"row is not held by this process"};
  }

  void download_rows(long long row0, long long nrows, real* out) {
    const int nx = prm.nx;
    const size_t row_bytes = (size_t)nx * 9 * sizeof(real);
    if (row_bytes > kStagingBytes) throw CudaError{"nx too large for the AoS staging buffer"};
    const int chunk_rows = (int)std::max<size_t>(1, kStagingBytes / row_bytes);
    long long g = row0;
    while (g < row0 + nrows) {
      Slab<real>& s = slab_of_row(g);
      CK(cudaSetDevice(s.device));
      const int n = (int)std::min<long long>({(long long)chunk_rows, s.row0 + s.rows - g, row0 + nrows - g});
      const long long ncells = (long long)n * nx;
      lbm::lbm_soa_to_aos<real><<<(unsigned)((ncells * 9 + 255) / 256), 256, 0, s.stream>>>(
          s.lattice[cur], (real*)s.staging, plane_stride(s), pitch, nx, (int)(g - s.row0), ncells);
      CK(cudaGetLastError());
      launches++;
      CK(cudaMemcpyAsync((char*)out + (size_t)(g - row0) * row_bytes, s.staging, (size_t)n * row_bytes,
                         cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
      g += n;
    }
  }

  // Fields of rows [row0, row0+nrows): computed chunk by chunk into one half of the staging
  // buffer (kernel stream) and copied out from the other half (copy stream): the kernel of
  // chunk i+1 overlaps the device-to-host copies of chunk i, one synchronisation per call.
  void final_fields(long long row0, long long nrows, real* ux, real* uy, real* u, real* p) {
    const int nx = prm.nx;
    const size_t row_bytes = (size_t)nx * sizeof(real);
    const size_t half = kStagingBytes / 2;
    if (4 * row_bytes > half) throw CudaError{"nx too large for the fields staging buffer"};
    const int chunk_rows = (int)(half / (4 * row_bytes));
    std::vector<Slab<real>*> used;
    long long g = row0;
    int i = 0;
    while (g < row0 + nrows) {
      Slab<real>& s = slab_of_row(g);
      CK(cudaSetDevice(s.device));
      if (used.empty() || used.back() != &s) {
        used.push_back(&s);
        i = 0;
      }
      const int n = (int)std::min<long long>({(long long)chunk_rows, s.row0 + s.rows - g, row0 + nrows - g});
      const long long ncells = (long long)n * nx;
      const int b = i & 1;
      real* st = (real*)((char*)s.staging + (size_t)b * half);
      real* d_ux = ux ? st : nullptr;
      real* d_uy = uy ? st + ncells : nullptr;
      real* d_u = u ? st + 2 * ncells : nullptr;
      real* d_p = p ? st + 3 * ncells : nullptr;
      if (i >= 2) CK(cudaStreamWaitEvent(s.stream, s.stage_drained[b], 0));
      lbm::lbm_fields<real><<<(unsigned)((ncells + 255) / 256), 256, 0, s.stream>>>(
          s.lattice[cur], s.mask, plane_stride(s), pitch, mask_pitch, nx, (int)(g - s.row0), n,
          prm.density, d_ux, d_uy, d_u, d_p, nullptr);
      CK(cudaGetLastError());
      launches++;
      CK(cudaEventRecord(s.stage_filled[b], s.stream));
      CK(cudaStreamWaitEvent(s.copy_stream, s.stage_filled[b], 0));
      const size_t off = (size_t)(g - row0) * nx;
      if (ux) CK(cudaMemcpyAsync(ux + off, d_ux, ncells * sizeof(real), cudaMemcpyDeviceToHost, s.copy_stream));
      if (uy) CK(cudaMemcpyAsync(uy + off, d_uy, ncells * sizeof(real), cudaMemcpyDeviceToHost, s.copy_stream));
      if (u) CK(cudaMemcpyAsync(u + off, d_u, ncells * sizeof(real), cudaMemcpyDeviceToHost, s.copy_stream));
      if (p) CK(cudaMemcpyAsync(p + off, d_p, ncells * sizeof(real), cudaMemcpyDeviceToHost, s.copy_stream));
      CK(cudaEventRecord(s.stage_drained[b], s.copy_stream));
      g += n;
      i++;
    }
    for (Slab<real>* s : used) {
      CK(cudaSetDevice(s->device));
      CK(cudaStreamSynchronize(s->copy_stream));
      CK(cudaStreamSynchronize(s->stream));
    }
  }

  double av_velocity_sum() {
    unsigned __int128 tot = 0;
    for (auto& s : slabs) {
      CK(cudaSetDevice(s.device));
      unsigned long long* av = s.sync + kAvScratch;       // 128-byte aligned inside the sync block
      CK(cudaMemsetAsync(av, 0, kAvWords * sizeof(unsigned long long), s.stream));
      const long long ncells = (long long)s.rows * prm.nx;
      lbm::lbm_fields<real><<<(unsigned)((ncells + 255) / 256), 256, 0, s.stream>>>(
          s.lattice[cur], s.mask, plane_stride(s), pitch, mask_pitch, prm.nx, 0, s.rows, prm.density,
          nullptr, nullptr, nullptr, nullptr, av);
      CK(cudaGetLastError());
      launches++;
      unsigned long long w[kAvWords];
      CK(cudaMemcpyAsync(w, av, sizeof w, cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
      for (int k = 0; k < LBM_AV_SLOTS; k++)
        tot += ((unsigned __int128)w[LBM_AV_STRIDE * k + 1] << 32) + w[LBM_AV_STRIDE * k];
    }
    return (double)((long double)tot / (long double)LBM_FIX_SCALE);
  }

  // exact digest of the rows held by this process (see lbm_digest)
  void digest(double* total_density, unsigned long long* checksum) {
    unsigned long long mass = 0, sum = 0;
    for (auto& s : slabs) {
      CK(cudaSetDevice(s.device));
      unsigned long long* out = s.sync + kScratch1;
      CK(cudaMemsetAsync(out, 0, 2 * sizeof(unsigned long long), s.stream));
      const long long ncells = (long long)s.rows * prm.nx;
      lbm::lbm_digest<real><<<(unsigned)((ncells + 255) / 256), 256, 0, s.stream>>>(
          s.lattice[cur], plane_stride(s), pitch, prm.nx, 0, s.rows, s.row0, out);
      CK(cudaGetLastError());
      launches++;
      unsigned long long w[2];
      CK(cudaMemcpyAsync(w, out, sizeof w, cudaMemcpyDeviceToHost, s.stream));
      CK(cudaStreamSynchronize(s.stream));
      mass += w[0];
      sum += w[1];
    }
    if (total_density) *total_density = (double)(long long)mass / 4294967296.0;
    if (checksum) *checksum = sum;
  }
};

// split ny rows over n slabs: remainder spread over the first slabs
void split_rows(long long ny, int n, std::vector<long long>& row0, std::vector<int>& rows) {
  row0.resize(n);
  rows.resize(n);
  const long long base = ny / n, rem = ny % n;
  long long r = 0;
  for (int i = 0; i < n; i++) {
    rows[i] = (int)(base + (i < rem ? 1 : 0));
    row0[i] = r;
    r += rows[i];
  }
}

template <typename real>
int create_impl(const typename ParamT<real>::type* params, const real* cells_aos, const void* obstacles,
                int n_gpus, const int* device_ids, unsigned flags, lbm_gpu** out) {
  if (!params || !out) return fail("lbm_gpu_create: NULL argument");
  *out = nullptr;
  if (params->nx < 1 || params->ny < 2) return fail("lbm_gpu_create: need nx >= 1 and ny >= 2 (got %d x %d)", params->nx, params->ny);
  if (n_gpus < 1) return fail("lbm_gpu_create: n_gpus must be >= 1");
  if (params->ny < n_gpus) return fail("lbm_gpu_create: fewer rows (%d) than GPUs (%d)", params->ny, n_gpus);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return fail("lbm_gpu_create: no CUDA device available (this library has no CPU fallback)");
  }
  std::unique_ptr<Grid<real>> g(new Grid<real>());
  try {
    g->prm = *params;
    g->flags = flags;
    g->setup_geometry(params->nx);
    std::vector<long long> row0;
    std::vector<int> rows;
    split_rows(params->ny, n_gpus, row0, rows);
    g->slabs.resize(n_gpus);
    const size_t cell_row = (size_t)params->nx * 9;
    const size_t obst_row_bytes = (flags & LBM_GPU_OBST_BITS) ? (size_t)((params->nx + 31) / 32) * 4 : (size_t)params->nx * 4;
    for (int i = 0; i < n_gpus; i++) {
      Slab<real>& s = g->slabs[i];
      s.device = device_ids ? device_ids[i] : i;
      if (s.device < 0 || s.device >= ndev) return fail("lbm_gpu_create: device %d not available (%d visible)", s.device, ndev);
      s.row0 = row0[i];
      s.rows = rows[i];
      const long long ar = (long long)params->ny - 2;
      s.accel_row = (ar >= s.row0 && ar < s.row0 + s.rows) ? (int)(ar - s.row0) : LBM_NO_ROW;
      g->alloc_slab(s);
      if (g->kernel == LBM_GPU_KERNEL_TMA) g->make_tensor_maps(s);
      g->load_slab(s, cells_aos ? cells_aos + (size_t)s.row0 * cell_row : nullptr,
                   obstacles ? (const char*)obstacles + (size_t)s.row0 * obst_row_bytes : nullptr, 0);
    }
    g->connect_local();
    g->choose_kernel();
    g->choose_pairs();
    g->choose_tb2();
    g->prepare();
  } catch (const CudaError& e) {
    return fail("lbm_gpu_create: %s", e.what.c_str());
  } catch (const std::exception& e) {      // e.g. std::bad_alloc: never let it cross the C boundary
    return fail("lbm_gpu_create: %s", e.what());
  }
  *out = reinterpret_cast<lbm_gpu*>(static_cast<GridBase*>(g.release()));
  return 0;
}

template <typename real>
Grid<real>* as_grid(lbm_gpu* h, const char* fn) {
  if (!h) { fail("%s: NULL handle", fn); return nullptr; }
  GridBase* b = reinterpret_cast<GridBase*>(h);
  if (b->is_f64() != (sizeof(real) == 8)) {
    fail("%s: handle was created for %s", fn, b->is_f64() ? "double precision (use the _f64 entry points)" : "single precision");
    return nullptr;
  }
  return static_cast<Grid<real>*>(b);
}

template <typename real, typename F>
int guarded(lbm_gpu* h, const char* fn, F&& body) {
  Grid<real>* g = as_grid<real>(h, fn);
  if (!g) return 1;
  try {
    body(*g);
  } catch (const CudaError& e) {
    return fail("%s: %s", fn, e.what.c_str());
  } catch (const std::exception& e) {
    return fail("%s: %s", fn, e.what());
  }
  return 0;
}

template <typename real>
int run_impl(lbm_gpu* h, int n_steps, real* av_out, double* sums_out, const char* fn) {
  if (n_steps < 0) return fail("%s: negative step count", fn);
  return guarded<real>(h, fn, [&](Grid<real>& g) {
    std::vector<double> sums;
    double* sp = sums_out;
    if (!sp && av_out) { sums.resize(std::max(n_steps, 1)); sp = sums.data(); }
    g.run(n_steps, sp);
    if (av_out) {
      const double div = (double)g.divisor();
      for (int t = 0; t < n_steps; t++) av_out[t] = (real)(sp[t] / div);
    }
  });
}

}  // namespace

// ======================================================================= C-ABI ====
extern "C" {

int lbm_gpu_abi_version(void) { return 2; }

const char* lbm_gpu_last_error(void) { return g_error.c_str(); }

int lbm_gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

int lbm_gpu_create(const lbm_param* params, const float* cells_aos, const void* obstacles, int n_gpus,
                   const int* device_ids, unsigned flags, lbm_gpu** out) {
  return create_impl<float>(params, cells_aos, obstacles, n_gpus, device_ids, flags, out);
}
int lbm_gpu_create_f64(const lbm_param_f64* params, const double* cells_aos, const void* obstacles, int n_gpus,
                       const int* device_ids, unsigned flags, lbm_gpu** out) {
  return create_impl<double>(params, cells_aos, obstacles, n_gpus, device_ids, flags, out);
}

int lbm_gpu_create_slab(const lbm_param* params, long long row0, long long nrows, int device,
                        const float* cells_aos_rows, const void* obstacles_rows, unsigned flags, lbm_gpu** out) {
  if (!params || !out) return fail("lbm_gpu_create_slab: NULL argument");
  *out = nullptr;
  if (params->nx < 1 || params->ny < 2) return fail("lbm_gpu_create_slab: need nx >= 1 and ny >= 2");
  if (row0 < 0 || nrows < 1 || row0 + nrows > params->ny) return fail("lbm_gpu_create_slab: rows [%lld,%lld) outside the grid", row0, row0 + nrows);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    return fail("lbm_gpu_create_slab: no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= ndev) return fail("lbm_gpu_create_slab: device %d not available (%d visible)", device, ndev);
  std::unique_ptr<Grid<float>> g(new Grid<float>());
  try {
    g->prm = *params;
    g->flags = flags;
    g->slab_mode = true;
    g->setup_geometry(params->nx);
    g->slabs.resize(1);
    Slab<float>& s = g->slabs[0];
    s.device = device;
    s.row0 = row0;
    s.rows = (int)nrows;
    const long long ar = (long long)params->ny - 2;
    s.accel_row = (ar >= row0 && ar < row0 + nrows) ? (int)(ar - row0) : LBM_NO_ROW;
    g->alloc_slab(s);
    if (g->kernel == LBM_GPU_KERNEL_TMA) g->make_tensor_maps(s);
    g->load_slab(s, cells_aos_rows, obstacles_rows, 0);
    if (g->kernel == LBM_GPU_KERNEL_PERSISTENT) throw CudaError{"the persistent kernel handles a single slab only"};
    if (nrows == params->ny) { g->connect_local(); g->choose_tb2(); g->prepare(); }   // whole grid in one slab
  } catch (const CudaError& e) {
    return fail("lbm_gpu_create_slab: %s", e.what.c_str());
  } catch (const std::exception& e) {
    return fail("lbm_gpu_create_slab: %s", e.what());
  }
  *out = reinterpret_cast<lbm_gpu*>(static_cast<GridBase*>(g.release()));
  return 0;
}

int lbm_gpu_ipc_export(lbm_gpu* h, void* desc) {
  if (!desc) return fail("lbm_gpu_ipc_export: NULL descriptor");
  return guarded<float>(h, "lbm_gpu_ipc_export", [&](Grid<float>& g) {
    if (g.slabs.size() != 1) throw CudaError{"only a one-slab handle can be exported"};
    Slab<float>& s = g.slabs[0];
    CK(cudaSetDevice(s.device));
    IpcDesc d;
    memset(&d, 0, sizeof d);
    d.magic = kIpcMagic;
    d.elem_size = 4;
    d.device = s.device;
    d.pid = (int32_t)getpid();
    d.nx = g.prm.nx; d.pitch = g.pitch; d.rows = s.rows;
    d.row0 = s.row0;
    d.win_addr = (unsigned long long)(uintptr_t)s.win;
    d.win_bytes = s.win_bytes;
    d.off_sync = s.off_sync;
    d.off_gmask = s.off_gmask;
    d.steps_done = (unsigned long long)g.steps_done;
    d.passes_done = (unsigned long long)g.passes_done;
    d.local_free_cells = s.free_cells;
    d.cur = g.cur;
    d.tb2_ok = g.tb2_possible(s.rows) ? 1 : 0;
    CK(cudaIpcGetMemHandle(&d.handle, s.win));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, s.device));
    memcpy(d.gpu_uuid, prop.uuid.bytes, 16);
    memset(desc, 0, LBM_GPU_IPC_DESC_BYTES);
    memcpy(desc, &d, sizeof d);
  });
}

namespace {
// wire a one-slab handle to the slabs below and above it; `ring_tb2`: every slab of the ring
// can run (and therefore will run) the two-step kernel
void ipc_connect_impl(Grid<float>& g, const IpcDesc& dn, const IpcDesc& up, bool ring_tb2) {
  if (g.slabs.size() != 1) throw CudaError{"only a one-slab handle can be connected"};
  Slab<float>& s = g.slabs[0];
  CK(cudaSetDevice(s.device));
  const long long ny = g.prm.ny;
  for (const IpcDesc* d : {&dn, &up}) {
    if (d->magic != kIpcMagic || d->elem_size != 4) throw CudaError{"bad neighbour descriptor"};
    if (d->nx != g.prm.nx || d->pitch != g.pitch) throw CudaError{"neighbour descriptor is for another grid width"};
    if (d->steps_done != (unsigned long long)g.steps_done || d->passes_done != (unsigned long long)g.passes_done ||
        d->cur != g.cur)
      throw CudaError{"neighbour is at a different timestep"};
  }
  cudaDeviceProp my_prop;
  CK(cudaGetDeviceProperties(&my_prop, s.device));
  for (const IpcDesc* d : {&dn, &up})
    if (memcmp(d->gpu_uuid, my_prop.uuid.bytes, 16) == 0 &&
        !(d->pid == (int32_t)getpid() && d->win_addr == (unsigned long long)(uintptr_t)s.win))
      throw CudaError{"a neighbouring slab runs on the same physical GPU: kernels that wait on one another "
                      "must not share a device (use lbm_gpu_create with n_gpus slabs in one process instead)"};
  if ((dn.row0 + dn.rows) % ny != s.row0 % ny) throw CudaError{"descriptor 'below' does not hold row0-1"};
  if ((s.row0 + s.rows) % ny != up.row0 % ny) throw CudaError{"descriptor 'above' does not hold row0+nrows"};
  auto map = [&](const IpcDesc& d, int slot) -> char* {
    if (d.pid == (int32_t)getpid()) {
      // exported by this very process (tests, or a host that drives several GPUs through
      // slab handles): the address is valid here, no IPC mapping needed
      char* p = (char*)(uintptr_t)d.win_addr;
      if (p == s.win) return p;
      if (d.device == s.device)
        throw CudaError{"two slabs ordered by device-side flags must not share a GPU"};
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, s.device, d.device));
      if (!can) throw CudaError{"peer access between the selected GPUs is not available"};
      cudaError_t e = cudaDeviceEnablePeerAccess(d.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
      cudaGetLastError();
      return p;
    }
    // LBM_GPU_POOL: the neighbour parks its window in its cache and exports the same
    // allocation again next time, so the mapping is kept too (cudaIpcCloseMemHandle is
    // as slow as cudaFree: up to 0.4 s measured)
    if (g.use_pool()) {
      std::lock_guard<std::mutex> lock(g_window_mutex);
      for (auto& c : g_ipc_cache)
        if (c.device == s.device && memcmp(&c.handle, &d.handle, sizeof d.handle) == 0) return (char*)c.ptr;
    }
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, d.handle, cudaIpcMemLazyEnablePeerAccess));
    if (g.use_pool()) {
      std::lock_guard<std::mutex> lock(g_window_mutex);
      g_ipc_cache.push_back({d.handle, s.device, p});
    } else {
      s.ipc_mapped[slot] = p;
    }
    return (char*)p;
  };
  for (const IpcDesc* d : {&dn, &up})
    if (d->win_bytes != s.win_bytes || d->off_sync != s.off_sync || d->off_gmask != s.off_gmask)
      throw CudaError{"neighbour window layout differs"};
  s.dn_win = map(dn, 0);
  s.up_win = (memcmp(&dn.handle, &up.handle, sizeof dn.handle) == 0 && dn.pid == up.pid) ? s.dn_win : map(up, 1);
  unsigned long long* dn_sync = (unsigned long long*)(s.dn_win + dn.off_sync);
  unsigned long long* up_sync = (unsigned long long*)(s.up_win + up.off_sync);
  s.dn_flag = dn_sync + kFlagFromAbove;
  s.up_flag = up_sync + kFlagFromBelow;
  s.dn_abort = dn_sync + kAbort;
  s.up_abort = up_sync + kAbort;
  s.dn_gmask = (uint32_t*)(s.dn_win + dn.off_gmask) + g.mask_pitch;
  s.up_gmask = (uint32_t*)(s.up_win + up.off_gmask);
  g.multi = true;
  g.use_flags = true;      // neighbours are other processes: device-side flags
  if (g.flags & LBM_GPU_KERNEL_TB2) {
    if (!ring_tb2) throw CudaError{"the two-step kernel needs lbm_gpu_ipc_connect_all and fp32, nx a multiple of 4 "
                                   "and >= 32, >= 8 rows on every rank"};
  }
  if (ring_tb2) g.enable_tb2();
  g.connected = true;
}
}  // namespace

int lbm_gpu_ipc_connect(lbm_gpu* h, const void* desc_below, const void* desc_above) {
  if (!desc_below || !desc_above) return fail("lbm_gpu_ipc_connect: NULL descriptor");
  return guarded<float>(h, "lbm_gpu_ipc_connect", [&](Grid<float>& g) {
    IpcDesc dn, up;
    memcpy(&dn, desc_below, sizeof dn);
    memcpy(&up, desc_above, sizeof up);
    ipc_connect_impl(g, dn, up, false);
  });
}

int lbm_gpu_ipc_connect_all(lbm_gpu* h, const void* descs, int n) {
  if (!descs || n < 1) return fail("lbm_gpu_ipc_connect_all: bad argument");
  return guarded<float>(h, "lbm_gpu_ipc_connect_all", [&](Grid<float>& g) {
    if (g.slabs.size() != 1) throw CudaError{"only a one-slab handle can be connected"};
    Slab<float>& s = g.slabs[0];
    std::vector<IpcDesc> all(n);
    for (int i = 0; i < n; i++) memcpy(&all[i], (const char*)descs + (size_t)i * LBM_GPU_IPC_DESC_BYTES, sizeof(IpcDesc));
    const long long ny = g.prm.ny;
    long long rows_total = 0, free_total = 0;
    int below = -1, above = -1;
    bool all_tb2 = true;
    for (int i = 0; i < n; i++) {
      const IpcDesc& d = all[i];
      if (d.magic != kIpcMagic) throw CudaError{"bad descriptor in the list"};
      rows_total += d.rows;
      free_total += d.local_free_cells;
      all_tb2 = all_tb2 && d.tb2_ok && d.rows >= kTb2MinRows;
      if ((d.row0 + d.rows) % ny == s.row0 % ny) below = i;
      if ((s.row0 + s.rows) % ny == d.row0 % ny) above = i;
    }
    if (rows_total != ny) throw CudaError{"the descriptors do not cover the grid's rows exactly once"};
    if (below < 0 || above < 0) throw CudaError{"no descriptor holds the rows next to this slab"};
    ipc_connect_impl(g, all[below], all[above], all_tb2);
    g.global_free_cells = free_total;
  });
}

int lbm_gpu_ipc_prepare(lbm_gpu* h) {
  return guarded<float>(h, "lbm_gpu_ipc_prepare", [&](Grid<float>& g) { g.prepare(); });
}

int lbm_gpu_run(lbm_gpu* h, int n_steps, float* av_vels_out) {
  return run_impl<float>(h, n_steps, av_vels_out, nullptr, "lbm_gpu_run");
}
int lbm_gpu_run_f64(lbm_gpu* h, int n_steps, double* av_vels_out) {
  return run_impl<double>(h, n_steps, av_vels_out, nullptr, "lbm_gpu_run_f64");
}
int lbm_gpu_run_sums(lbm_gpu* h, int n_steps, double* sums_out) {
  if (!h) return fail("lbm_gpu_run_sums: NULL handle");
  if (reinterpret_cast<GridBase*>(h)->is_f64()) return run_impl<double>(h, n_steps, nullptr, sums_out, "lbm_gpu_run_sums");
  return run_impl<float>(h, n_steps, nullptr, sums_out, "lbm_gpu_run_sums");
}

int lbm_gpu_set_global_free_cells(lbm_gpu* h, long long free_cells) {
  if (!h) return fail("lbm_gpu_set_global_free_cells: NULL handle");
  if (free_cells < 1) return fail("lbm_gpu_set_global_free_cells: count must be positive");
  GridBase* b = reinterpret_cast<GridBase*>(h);
  if (b->is_f64()) static_cast<Grid<double>*>(b)->global_free_cells = free_cells;
  else static_cast<Grid<float>*>(b)->global_free_cells = free_cells;
  return 0;
}

int lbm_gpu_download(lbm_gpu* h, float* out) {
  if (!out) return fail("lbm_gpu_download: NULL output");
  return guarded<float>(h, "lbm_gpu_download", [&](Grid<float>& g) {
    long long rows = 0;
    for (auto& s : g.slabs) rows += s.rows;
    g.download_rows(g.slabs[0].row0, rows, out);
  });
}
int lbm_gpu_download_f64(lbm_gpu* h, double* out) {
  if (!out) return fail("lbm_gpu_download_f64: NULL output");
  return guarded<double>(h, "lbm_gpu_download_f64", [&](Grid<double>& g) {
    long long rows = 0;
    for (auto& s : g.slabs) rows += s.rows;
    g.download_rows(g.slabs[0].row0, rows, out);
  });
}
int lbm_gpu_download_rows(lbm_gpu* h, long long row0, long long nrows, float* out) {
  if (!out || nrows < 0) return fail("lbm_gpu_download_rows: bad argument");
  return guarded<float>(h, "lbm_gpu_download_rows", [&](Grid<float>& g) { g.download_rows(row0, nrows, out); });
}

int lbm_gpu_final_fields(lbm_gpu* h, long long row0, long long nrows, float* ux, float* uy, float* u, float* p) {
  if (nrows < 0) return fail("lbm_gpu_final_fields: bad argument");
  return guarded<float>(h, "lbm_gpu_final_fields", [&](Grid<float>& g) { g.final_fields(row0, nrows, ux, uy, u, p); });
}
int lbm_gpu_final_fields_f64(lbm_gpu* h, long long row0, long long nrows, double* ux, double* uy, double* u, double* p) {
  if (nrows < 0) return fail("lbm_gpu_final_fields_f64: bad argument");
  return guarded<double>(h, "lbm_gpu_final_fields_f64", [&](Grid<double>& g) { g.final_fields(row0, nrows, ux, uy, u, p); });
}

int lbm_gpu_av_velocity(lbm_gpu* h, float* av_out) {
  if (!av_out) return fail("lbm_gpu_av_velocity: NULL output");
  return guarded<float>(h, "lbm_gpu_av_velocity", [&](Grid<float>& g) {
    *av_out = (float)(g.av_velocity_sum() / (double)g.divisor());
  });
}
int lbm_gpu_av_velocity_f64(lbm_gpu* h, double* av_out) {
  if (!av_out) return fail("lbm_gpu_av_velocity_f64: NULL output");
  return guarded<double>(h, "lbm_gpu_av_velocity_f64", [&](Grid<double>& g) {
    *av_out = g.av_velocity_sum() / (double)g.divisor();
  });
}

int lbm_gpu_digest(lbm_gpu* h, double* total_density, unsigned long long* checksum) {
  if (!h) return fail("lbm_gpu_digest: NULL handle");
  if (reinterpret_cast<GridBase*>(h)->is_f64())
    return guarded<double>(h, "lbm_gpu_digest", [&](Grid<double>& g) { g.digest(total_density, checksum); });
  return guarded<float>(h, "lbm_gpu_digest", [&](Grid<float>& g) { g.digest(total_density, checksum); });
}

int lbm_gpu_upload(lbm_gpu* h, const float* cells_aos) {
  if (!cells_aos) return fail("lbm_gpu_upload: NULL input");
  return guarded<float>(h, "lbm_gpu_upload", [&](Grid<float>& g) {
    const int cur = g.cur;
    const long long base = g.slabs[0].row0;
    for (auto& s : g.slabs) g.upload_cells(s, cells_aos + (size_t)(s.row0 - base) * g.prm.nx * 9, cur);
    if (!g.slab_mode || g.slabs[0].rows == g.prm.ny) g.prepare();
  });
}
int lbm_gpu_upload_f64(lbm_gpu* h, const double* cells_aos) {
  if (!cells_aos) return fail("lbm_gpu_upload_f64: NULL input");
  return guarded<double>(h, "lbm_gpu_upload_f64", [&](Grid<double>& g) {
    const int cur = g.cur;
    for (auto& s : g.slabs) g.upload_cells(s, cells_aos + (size_t)s.row0 * g.prm.nx * 9, cur);
    g.prepare();
  });
}

int lbm_gpu_get_info(lbm_gpu* h, lbm_gpu_info* info) {
  if (!h || !info) return fail("lbm_gpu_get_info: NULL argument");
  memset(info, 0, sizeof *info);
  GridBase* b = reinterpret_cast<GridBase*>(h);
  auto fill = [&](auto& g) {
    info->nx = g.prm.nx; info->ny = g.prm.ny;
    info->n_gpus = (int)g.slabs.size();
    info->is_f64 = g.is_f64();
    info->kernel = g.kernel;
    info->pitch = g.pitch;
    info->local_free_cells = g.local_free_cells();
    info->free_cells = g.divisor();
    info->local_row0 = g.slabs[0].row0;
    long long rows = 0;
    size_t bytes = 0;
    for (auto& s : g.slabs) { rows += s.rows; bytes = std::max(bytes, s.bytes + kStagingBytes); }
    info->local_rows = rows;
    info->steps_done = g.steps_done;
    info->kernel_launches = g.launches;
    info->last_run_device_ms = g.last_run_ms;
    info->last_step_kernel_ms = g.last_step_ms;
    info->device_bytes = bytes;
  };
  if (b->is_f64()) fill(*static_cast<Grid<double>*>(b));
  else fill(*static_cast<Grid<float>*>(b));
  return 0;
}

int lbm_gpu_host_alloc(size_t bytes, void** out) {
  if (!out) return fail("lbm_gpu_host_alloc: NULL argument");
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail("lbm_gpu_host_alloc: %s", cudaGetErrorString(e));
  }
  return 0;
}

void lbm_gpu_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

void lbm_gpu_destroy(lbm_gpu* h) {
  if (!h) return;
  delete reinterpret_cast<GridBase*>(h);
}

}  // extern "C"
