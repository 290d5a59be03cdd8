// art_b200.cu -- the C ABI of libart_b200.so (include/art_b200.h): argument checking, the
// host-side lowering of scene descriptions, launch configuration and the host-buffer entry point.
// All ray arithmetic happens in the kernels of art_kernels.cuh; there is no CPU ray path here.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "art_kernels.cuh"
#include "art_lowering.h"

using namespace art;

// -------------------------------------------------------------------------------------------------
// error plumbing
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int32_t fail(int32_t code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define ART_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return fail(ART_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
  } while (0)
#define ART_LAUNCHED()                                                                       \
  do {                                                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) return fail(ART_E_CUDA, std::string("launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// -------------------------------------------------------------------------------------------------
// the chain object
// -------------------------------------------------------------------------------------------------
struct HostWorkspace {
  double* cols = nullptr;     // 7 input + 8 output columns of cap_n doubles each
  uint8_t* alive = nullptr;
  size_t cap_n = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev[8] = {};
  double* pinned = nullptr;   // staging: moments | central | detector | peer epoch, status | caller's detector (in)
  double* src_state = nullptr;  // device source state of art_run_source_host (SRC_STATE_LEN doubles)
  // art_run_source_host replays its launch sequence as a CUDA graph once the same call has been seen twice
  cudaGraphExec_t src_graph = nullptr;
  std::string src_key;        // the arguments the cached graph was captured for
  std::string src_last_key;   // arguments of the previous call
  int src_exchanges = 0;      // peer exchanges inside the cached graph
};

struct ArtChain {
  int device = 0;
  int sm_count = 148;
  int n_elements = 0, n_variants = 0, n_defects = 0;
  ElemDev* d_elems = nullptr;
  double* d_ztab = nullptr;
  int* d_zoff = nullptr;
  MapDev* d_maps = nullptr;
  int ztab_len = 0;
  size_t smem_bytes = 0;
  bool has_defects = false;
  int surfs = 0;                 // SURFS_* class of the chain's surfaces
  double* d_partials = nullptr;
  size_t partial_rows = 0;
  int unfolded_rows = 0;         // partial rows a NO_FOLD launch left for art_peer_exchange_fold (0: none)
  int unfolded_kind = -1;        // 0: central rows (PLEN_TRACE), 1: moments rows (PLEN_DET)
  double* d_central = nullptr;   // n_variants x ART_CENTRAL_LEN scratch (sweep, host run)
  double* d_moments = nullptr;   // n_variants x ART_MOMENTS_LEN scratch
  ArtDetector* d_det = nullptr;  // n_variants scratch
  HostWorkspace ws;
};

#ifndef ART_GRID_PER_SM
#define ART_GRID_PER_SM 2  // blocks per SM of the grid-stride kernels = the resident count: a persistent grid (measured 8 -> 2: cfg2 step 0.414 -> 0.397 ms, fewer block prologues and partial rows)
#endif
static int blocks_per_variant(const ArtChain* c, long long n, int n_variants, int per_thread = RPT, int tpb = TPB) {
  const long long npairs = (n + per_thread - 1) / per_thread;
  long long maxb = (npairs + tpb - 1) / tpb;
  if (maxb < 1) maxb = 1;
  // one variant: a persistent grid; several variants share the SMs block by block, so keep enough blocks
  // for the last wave to be full (1024 variants x 1 block left the tail at 46 %: 87.7 -> 92.7 ms on cfg5)
  long long target = (long long)c->sm_count * (n_variants > 1 ? 8 : ART_GRID_PER_SM);
  long long bpv = (target + n_variants - 1) / n_variants;
  if (bpv > maxb) bpv = maxb;
  if (bpv < 1) bpv = 1;
  return (int)bpv;
}

static BundleDev to_dev(const ArtBundleView* b) {
  BundleDev d;
  if (!b) {
    d.px = d.py = d.pz = d.ux = d.uy = d.uz = d.path = d.inc = d.inten = nullptr;
    d.alive = nullptr;
    return d;
  }
  d.px = b->px; d.py = b->py; d.pz = b->pz;
  d.ux = b->ux; d.uy = b->uy; d.uz = b->uz;
  d.path = b->path; d.inc = b->incidence; d.inten = b->intensity;
  d.alive = b->alive;
  return d;
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static const char* check_columns(const ArtBundleView* b, bool need_pu, bool uniform_point = false) {
  const double* cols[9] = {b->px, b->py, b->pz, b->ux, b->uy, b->uz, b->path, b->incidence, b->intensity};
  int have = 0;
  for (int i = 0; i < 6; ++i) have += cols[i] != nullptr;
  if (need_pu && have != 6) return "the point and vector columns (px..uz) are required";
  if (have != 0 && have != 6) return "point / vector columns must be given all six or not at all";
  for (int i = uniform_point ? 3 : 0; i < 9; ++i)  // a uniform point is three single doubles
    if (cols[i] && !aligned16(cols[i])) return "ray columns must be 16-byte aligned";
  return nullptr;
}

// -------------------------------------------------------------------------------------------------
// plain queries
// -------------------------------------------------------------------------------------------------
extern "C" int32_t art_version(void) { return ART_B200_VERSION; }
extern "C" const char* art_last_error(void) { return g_err.c_str(); }
extern "C" int64_t art_launch_count(void) { return g_launches.load(); }

extern "C" int32_t art_abi_sizes(int32_t sizes_out[6]) {
  if (!sizes_out) return fail(ART_E_INVALID, "sizes_out is NULL");
  sizes_out[0] = (int32_t)sizeof(ArtElementDesc);
  sizes_out[1] = (int32_t)sizeof(ArtZernikeDesc);
  sizes_out[2] = (int32_t)sizeof(ArtBundleView);
  sizes_out[3] = (int32_t)sizeof(ArtDetector);
  sizes_out[4] = (int32_t)sizeof(ArtGridMapDesc);
  sizes_out[5] = (int32_t)sizeof(ArtSourceDesc);
  return ART_OK;
}

extern "C" int32_t art_device_count(int32_t* count) {
  if (!count) return fail(ART_E_INVALID, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(ART_E_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = n;
  return ART_OK;
}

extern "C" int32_t art_element_rotation(const double normal[3], const double majoraxis[3], double rot_out[9]) {
  if (!normal || !majoraxis || !rot_out) return fail(ART_E_INVALID, "NULL argument");
  element_rotation(normal, majoraxis, rot_out);
  return ART_OK;
}

extern "C" int32_t art_detector_make(const double centre[3], const double normal[3], const double refpoint[3],
                                     double l0, ArtDetector* det_out) {
  if (!centre || !normal || !refpoint || !det_out) return fail(ART_E_INVALID, "NULL argument");
  const double nn = std::sqrt(normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2]);
  if (!(nn > 0.0)) return fail(ART_E_INVALID, "detector normal must be non-zero");
  // Detector.normal setter normalises (ART/ModuleDetector.py:62-70)
  const double nrm[3] = {normal[0] / nn, normal[1] / nn, normal[2] / nn};
  const double cv[3] = {-nrm[0], -nrm[1], -nrm[2]};
  detector_fill(centre, nrm, refpoint, cv, l0, 0.0, det_out);
  return ART_OK;
}

// -------------------------------------------------------------------------------------------------
// chain construction
// -------------------------------------------------------------------------------------------------
extern "C" int32_t art_chain_destroy(ArtChain* c) {
  if (!c) return ART_OK;
  cudaSetDevice(c->device);
  cudaFree(c->d_elems);
  cudaFree(c->d_ztab);
  cudaFree(c->d_zoff);
  cudaFree(c->d_maps);
  cudaFree(c->d_partials);
  cudaFree(c->d_central);
  cudaFree(c->d_moments);
  cudaFree(c->d_det);
  cudaFree(c->ws.cols);
  cudaFree(c->ws.alive);
  cudaFree(c->ws.src_state);
  if (c->ws.src_graph) cudaGraphExecDestroy(c->ws.src_graph);
  if (c->ws.pinned) cudaFreeHost(c->ws.pinned);
  for (auto& e : c->ws.ev)
    if (e) cudaEventDestroy(e);
  if (c->ws.stream) cudaStreamDestroy(c->ws.stream);
  if (c->ws.copy_stream) cudaStreamDestroy(c->ws.copy_stream);
  delete c;
  return ART_OK;
}

template <typename K>
static cudaError_t allow_smem(K kernel, size_t bytes) {
  if (bytes <= 32 * 1024) return cudaSuccess;  // static + dynamic stay below the default 48 KB limit
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

extern "C" int32_t art_chain_create(const ArtElementDesc* elements, int32_t n_elements, int32_t n_variants,
                                    const ArtZernikeDesc* defects, int32_t n_defects, const ArtGridMapDesc* gridmaps,
                                    int32_t n_gridmaps, ArtChain** chain_out) {
  if (!chain_out) return fail(ART_E_INVALID, "chain_out is NULL");
  *chain_out = nullptr;
  if (!elements || n_elements < 1 || n_elements > ART_MAX_ELEMENTS)
    return fail(ART_E_INVALID, "n_elements must be in [1, " + std::to_string(ART_MAX_ELEMENTS) + "]");
  if (n_variants < 1) return fail(ART_E_INVALID, "n_variants must be >= 1");
  if (n_defects < 0 || (n_defects > 0 && !defects)) return fail(ART_E_INVALID, "bad defect list");
  if (n_gridmaps < 0 || (n_gridmaps > 0 && !gridmaps)) return fail(ART_E_INVALID, "bad grid-map list");

  std::vector<ElemDev> h((size_t)n_elements * n_variants);
  bool any_def = false;
  for (int v = 0; v < n_variants; ++v)
    for (int k = 0; k < n_elements; ++k) {
      const ArtElementDesc& d = elements[(size_t)v * n_elements + k];
      std::string why = lower_element(d, h[(size_t)v * n_elements + k]);
      if (!why.empty())
        return fail(ART_E_INVALID, "variant " + std::to_string(v) + " element " + std::to_string(k) + ": " + why);
      if (d.n_defects > 0) {
        any_def = true;
        if (d.first_defect + d.n_defects > n_defects)
          return fail(ART_E_INVALID, "element " + std::to_string(k) + " refers to defects beyond the list");
      }
      if (d.n_gridmaps > 0) {
        any_def = true;
        if (d.first_gridmap + d.n_gridmaps > n_gridmaps)
          return fail(ART_E_INVALID, "element " + std::to_string(k) + " refers to grid maps beyond the list");
      }
      if (v > 0) {
        const ArtElementDesc& d0 = elements[k];
        if (d0.surface != d.surface || d0.support != d.support || d0.n_defects != d.n_defects ||
            d0.first_defect != d.first_defect || d0.n_gridmaps != d.n_gridmaps || d0.first_gridmap != d.first_gridmap)
          return fail(ART_E_INVALID, "variants must share surface / support kinds and defects");
      }
    }
  for (int v = 0; v < n_variants; ++v)  // hand-over maps between consecutive elements of each variant
    for (int k = 0; k + 1 < n_elements; ++k) link_elements(h[(size_t)v * n_elements + k], h[(size_t)v * n_elements + k + 1]);
  std::vector<double> ztab;
  std::vector<int> zoff;
  for (int i = 0; i < n_defects; ++i) {
    std::vector<double> t;
    std::string why = build_zernike_table(defects[i], t);
    if (!why.empty()) return fail(ART_E_INVALID, "defect " + std::to_string(i) + ": " + why);
    zoff.push_back((int)ztab.size());
    ztab.insert(ztab.end(), t.begin(), t.end());
  }

  std::vector<MapDev> hmaps((size_t)n_gridmaps);
  for (int i = 0; i < n_gridmaps; ++i) {
    std::string why = lower_gridmap(gridmaps[i], hmaps[i]);
    if (!why.empty()) return fail(ART_E_INVALID, "grid map " + std::to_string(i) + ": " + why);
  }

  ArtChain* c = new (std::nothrow) ArtChain();
  if (!c) return fail(ART_E_NOMEM, "out of host memory");
  auto bail = [&](int32_t code, const std::string& msg) {
    art_chain_destroy(c);
    return fail(code, msg);
  };
  cudaError_t e;
#define CK(call)                                                                  \
  if ((e = (call)) != cudaSuccess) return bail(ART_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e))
  CK(cudaGetDevice(&c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10)
    return bail(ART_E_UNSUPPORTED, std::string("libart_b200 is built for sm_100a only; device is ") + prop.name);
  c->sm_count = prop.multiProcessorCount;
  c->n_elements = n_elements;
  c->n_variants = n_variants;
  c->n_defects = n_defects;
  c->has_defects = any_def;
  c->ztab_len = (int)ztab.size();
  c->smem_bytes = sizeof(ElemDev) * ART_MAX_ELEMENTS + sizeof(double) * ztab.size() + sizeof(int) * zoff.size();
  c->smem_bytes = (c->smem_bytes + 15) & ~size_t(15);
  if (c->smem_bytes > 200 * 1024) return bail(ART_E_UNSUPPORTED, "Zernike tables exceed 200 KB of shared memory");
  {
    // dynamic shared memory: tables | moment slots (fused detector) | cp.async input stages (plain trace)
#define ART_ALLOW(DEFS, SURF)                                                      \
  fused = c->smem_bytes + smem_moments_bytes(trace_block_threads(DEFS, SURF));     \
  plain = c->smem_bytes + stage_bytes(trace_block_threads(DEFS, SURF));            \
  CK(allow_smem(trace_kernel<true, false, DEFS, SURF, false>, plain));             \
  CK(allow_smem(trace_kernel<false, false, DEFS, SURF, false>, plain));            \
  CK(allow_smem(trace_kernel<true, true, DEFS, SURF, false>, fused));              \
  CK(allow_smem(trace_kernel<false, true, DEFS, SURF, false>, fused));             \
  CK(allow_smem(trace_kernel<true, false, DEFS, SURF, true>, plain));              \
  CK(allow_smem(trace_kernel<false, false, DEFS, SURF, true>, plain));             \
  CK(allow_smem(trace_kernel<true, true, DEFS, SURF, true>, fused));               \
  CK(allow_smem(trace_kernel<false, true, DEFS, SURF, true>, fused));
    size_t fused = 0, plain = 0;
    ART_ALLOW(true, SURFS_ANY)
    ART_ALLOW(false, SURFS_ANY)
    ART_ALLOW(false, SURFS_TOROID)
    ART_ALLOW(false, SURFS_QUADRIC)
#undef ART_ALLOW
    CK(allow_smem(detector_kernel, DET_STAGE_BYTES));
  }
  {
    bool tor = false, quad = false;
    for (int k = 0; k < n_elements; ++k) {
      const int sf = elements[k].surface;
      tor = tor || sf == ART_SURF_TOROIDAL;
      quad = quad || (sf == ART_SURF_SPHERICAL || sf == ART_SURF_PARABOLIC || sf == ART_SURF_ELLIPSOIDAL ||
                      sf == ART_SURF_CYLINDRICAL);
    }
    c->surfs = (tor && quad) ? SURFS_ANY : (tor ? SURFS_TOROID : SURFS_QUADRIC);
  }

  CK(cudaMalloc(&c->d_elems, h.size() * sizeof(ElemDev)));
  CK(cudaMemcpy(c->d_elems, h.data(), h.size() * sizeof(ElemDev), cudaMemcpyHostToDevice));
  if (!ztab.empty()) {
    CK(cudaMalloc(&c->d_ztab, ztab.size() * sizeof(double)));
    CK(cudaMemcpy(c->d_ztab, ztab.data(), ztab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_zoff, zoff.size() * sizeof(int)));
    CK(cudaMemcpy(c->d_zoff, zoff.data(), zoff.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (!hmaps.empty()) {
    CK(cudaMalloc(&c->d_maps, hmaps.size() * sizeof(MapDev)));
    CK(cudaMemcpy(c->d_maps, hmaps.data(), hmaps.size() * sizeof(MapDev), cudaMemcpyHostToDevice));
  }
  c->partial_rows = (size_t)c->sm_count * 8 + (size_t)n_variants + 8;
  CK(cudaMalloc(&c->d_partials, c->partial_rows * PLEN_FUSED * sizeof(double)));
  CK(cudaMalloc(&c->d_central, (size_t)n_variants * ART_CENTRAL_LEN * sizeof(double)));
  CK(cudaMalloc(&c->d_moments, (size_t)n_variants * ART_MOMENTS_LEN * sizeof(double)));
  CK(cudaMalloc(&c->d_det, (size_t)n_variants * sizeof(ArtDetector)));
#undef CK
  *chain_out = c;
  return ART_OK;
}

// -------------------------------------------------------------------------------------------------
// trace launches
// -------------------------------------------------------------------------------------------------
static int32_t launch_trace(ArtChain* c, int variant_first, int n_variants, const ArtBundleView* in,
                            const ArtBundleView* out_final, const ArtBundleView* out_history, uint32_t flags,
                            const ArtDetector* det, double* x_out, double* y_out, double* l_out,
                            double* central_out, double* moments_out, cudaStream_t st, bool keep_l2 = false,
                            double place_distance = 0.0, ArtDetector* place_det = nullptr, int chunk_blocks = 0,
                            int chunk_row0 = 0, const double* wstate = nullptr, double wcoef = 0.0) {
  // chunk_blocks > 0: this launch is one chunk of a pipelined trace -- it uses exactly chunk_blocks
  // blocks, writes its partial rows at chunk_row0 and leaves the fold to the caller
  if (!c) return fail(ART_E_INVALID, "chain is NULL");
  if (!in) return fail(ART_E_INVALID, "input bundle is NULL");
  if (variant_first < 0 || n_variants < 1 || variant_first + n_variants > c->n_variants)
    return fail(ART_E_INVALID, "variant range outside the chain's variants");
  if (in->n < 0) return fail(ART_E_INVALID, "negative ray count");
  // the trace kernel indexes the rays of one variant with 32 bits (ART_INDEX32, art_kernels.cuh); 180 GB of HBM
  // hold fewer than 2^32 rays (48 B per source ray)
  if (in->n >= (int64_t)0xFFFFFFFEll) return fail(ART_E_INVALID, "more than 2^32 - 2 rays in one bundle");
  const bool uniform_point = (flags & ART_TRACE_UNIFORM_POINT) != 0;
  if (const char* why = check_columns(in, true, uniform_point))
    return fail(ART_E_INVALID, std::string("input bundle: ") + why);
  if (out_final) {
    if (const char* why = check_columns(out_final, false))
      return fail(ART_E_INVALID, std::string("output bundle: ") + why);
    if (out_final->n != in->n * (int64_t)n_variants)
      return fail(ART_E_INVALID, "output bundle must hold n_variants * n rays");
  }
  if (det && !moments_out) return fail(ART_E_INVALID, "moments_out is NULL");

  TraceArgs a;
  a.elems = c->d_elems;
  a.n_elements = c->n_elements;
  a.variant_first = variant_first;
  a.ztab = c->d_ztab;
  a.zoff = c->d_zoff;
  a.ztab_len = c->ztab_len;
  a.n_defects = c->n_defects;
  a.maps = c->d_maps;
  a.in = to_dev(in);
  a.out = to_dev(out_final);
  a.has_out = out_final != nullptr;
  a.has_hist = out_history != nullptr;
  bool want_inc = false;
  for (int k = 0; k < ART_MAX_ELEMENTS; ++k) {
    if (out_history && k < c->n_elements) {
      if (const char* why = check_columns(&out_history[k], false))
        return fail(ART_E_INVALID, std::string("history bundle: ") + why);
      if (out_history[k].n != in->n * (int64_t)n_variants)
        return fail(ART_E_INVALID, "history bundles must hold n_variants * n rays");
      a.hist[k] = to_dev(&out_history[k]);
      want_inc = want_inc || out_history[k].incidence != nullptr;
    } else {
      a.hist[k] = to_dev(nullptr);
    }
  }
  want_inc = want_inc || (out_final && out_final->incidence);
  if (flags & ART_TRACE_NO_INCIDENCE) want_inc = false;
  a.n = in->n;
  a.flags = flags;
  a.partials = c->d_partials;
  a.det = det;
  a.x_out = x_out;
  a.y_out = y_out;
  a.l_out = l_out;

  const int surfs_class = c->has_defects ? SURFS_ANY : c->surfs;
  const int bt = trace_block_threads(c->has_defects, surfs_class);  // threads per block of this chain's kernels
  const int bpv = chunk_blocks > 0 ? chunk_blocks : blocks_per_variant(c, in->n, n_variants, ART_RPT, bt);
  if ((size_t)bpv * n_variants + chunk_row0 > c->partial_rows)
    return fail(ART_E_INVALID, "internal: partial buffer too small");
  if (chunk_blocks > 0) a.partials = c->d_partials + (size_t)chunk_row0 * PLEN_TRACE;
  const dim3 grid(bpv, n_variants);
  a.moments_smem_offset = (int)c->smem_bytes;
  a.stage_smem_offset = (int)c->smem_bytes;
  a.keep_l2 = keep_l2 ? 1 : 0;
  a.uniform_point = uniform_point ? 1 : 0;
  a.wstate = wstate;
  a.wcoef = wcoef;
  if (wstate && !in->intensity) return fail(ART_E_INVALID, "internal: computed weights need an intensity column");
  const size_t sm = c->smem_bytes + (det ? (size_t)smem_moments_bytes(bt) : (size_t)stage_bytes(bt, uniform_point));
  // Zernike chains run the general kernel; defect-free chains one specialised for their surface class
#define ART_TRACE_LAUNCH2(INC, DET, UPT)                                                                      \
  do {                                                                                                       \
    if (c->has_defects) trace_kernel<INC, DET, true, SURFS_ANY, UPT><<<grid, bt, sm, st>>>(a);              \
    else if (c->surfs == SURFS_TOROID) trace_kernel<INC, DET, false, SURFS_TOROID, UPT><<<grid, bt, sm, st>>>(a);   \
    else if (c->surfs == SURFS_QUADRIC) trace_kernel<INC, DET, false, SURFS_QUADRIC, UPT><<<grid, bt, sm, st>>>(a); \
    else trace_kernel<INC, DET, false, SURFS_ANY, UPT><<<grid, bt, sm, st>>>(a);                            \
  } while (0)
#define ART_TRACE_LAUNCH(INC, DET)                       \
  do {                                                   \
    if (uniform_point) ART_TRACE_LAUNCH2(INC, DET, true); \
    else ART_TRACE_LAUNCH2(INC, DET, false);             \
  } while (0)
  if (det) {
    if (want_inc) ART_TRACE_LAUNCH(true, true);
    else ART_TRACE_LAUNCH(false, true);
    ART_LAUNCHED();
    fold_kernel<<<n_variants, TPB, 0, st>>>(c->d_partials, bpv, 1, central_out, moments_out);
    ART_LAUNCHED();
  } else {
    if (want_inc) ART_TRACE_LAUNCH(true, false);
    else ART_TRACE_LAUNCH(false, false);
    ART_LAUNCHED();
    if ((flags & ART_TRACE_NO_FOLD) && chunk_blocks == 0) {
      if (n_variants != 1) return fail(ART_E_INVALID, "ART_TRACE_NO_FOLD is for one variant");
      c->unfolded_rows = bpv;
      c->unfolded_kind = 0;
    } else if (central_out && chunk_blocks == 0) {
      // place_det: fold the central sums and place the variant's detector in the same launch
      fold_kernel<<<n_variants, TPB, 0, st>>>(c->d_partials, bpv, place_det ? 3 : 0, central_out, nullptr,
                                              place_distance, place_det);
      ART_LAUNCHED();
    }
  }
#undef ART_TRACE_LAUNCH
#undef ART_TRACE_LAUNCH2
  return ART_OK;
}

extern "C" int32_t art_trace(ArtChain* chain, int32_t variant_first, int32_t n_variants, const ArtBundleView* in,
                             const ArtBundleView* out_final, const ArtBundleView* out_history, uint32_t flags,
                             double* central_out, void* stream) {
  return launch_trace(chain, variant_first, n_variants, in, out_final, out_history, flags, nullptr, nullptr,
                      nullptr, nullptr, central_out, nullptr, (cudaStream_t)stream);
}

extern "C" int32_t art_trace_detect(ArtChain* chain, int32_t variant_first, int32_t n_variants,
                                    const ArtBundleView* in, const ArtBundleView* out_final, uint32_t flags,
                                    const ArtDetector* det, double* x_out, double* y_out, double* l_out,
                                    double* central_out, double* moments_out, void* stream) {
  if (!det) return fail(ART_E_INVALID, "det is NULL");
  return launch_trace(chain, variant_first, n_variants, in, out_final, nullptr, flags, det, x_out, y_out, l_out,
                      central_out, moments_out, (cudaStream_t)stream);
}

extern "C" int32_t art_detector_autoplace(const double* central, double distance, int32_t n_variants,
                                          ArtDetector* det_out, void* stream) {
  if (!central || !det_out || n_variants < 1) return fail(ART_E_INVALID, "bad argument");
  autoplace_kernel<<<(n_variants + 63) / 64, 64, 0, (cudaStream_t)stream>>>(central, distance, n_variants, det_out);
  ART_LAUNCHED();
  return ART_OK;
}

// reduction scratch for calls that come without a chain (stand-alone Detector objects): one lazily
// grown buffer per (device, stream), so concurrent callers on different streams or host threads never
// share partial rows; the table itself is guarded by a mutex.  Buffers live until the process ends.
struct ScratchSlot {
  double* ptr = nullptr;
  size_t rows = 0;
};
static std::mutex g_scratch_mutex;
static std::map<std::pair<int, void*>, ScratchSlot> g_scratch;
static int32_t global_scratch(size_t rows, void* stream, double** out) {
  int dev = 0;
  ART_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_scratch_mutex);
  ScratchSlot& s = g_scratch[std::make_pair(dev, stream)];
  if (s.rows < rows) {
    if (s.ptr) {
      // kernels still reading the old buffer were launched on this very stream: order the free behind them
      ART_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
      ART_CUDA(cudaFree(s.ptr));
    }
    s.ptr = nullptr;
    s.rows = 0;
    ART_CUDA(cudaMalloc(&s.ptr, rows * PLEN_FUSED * sizeof(double)));
    s.rows = rows;
  }
  *out = s.ptr;
  return ART_OK;
}

extern "C" int32_t art_detector_moments(ArtChain* chain, const ArtBundleView* bundle, int32_t n_variants,
                                        const ArtDetector* det, double* x_out, double* y_out, double* l_out,
                                        double* moments_out, void* stream) {
  if (!bundle || !det || n_variants < 1) return fail(ART_E_INVALID, "bad argument");
  if (!moments_out && (!chain || n_variants != 1))
    return fail(ART_E_INVALID, "moments_out may be NULL only with a chain and one variant (art_peer_exchange_fold)");
  ArtChain tmp;  // launch-shape defaults when no chain lends its scratch
  if (!chain) {
    // no chain has opted this device in to the kernel's shared-memory size yet
    ART_CUDA(cudaFuncSetAttribute(detector_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DET_STAGE_BYTES));
    int dev = 0;
    ART_CUDA(cudaGetDevice(&dev));
    int sms = 148;
    ART_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    tmp.sm_count = sms;
    tmp.partial_rows = (size_t)sms * 8 + (size_t)n_variants + 8;
    int32_t rc = global_scratch(tmp.partial_rows, stream, &tmp.d_partials);
    if (rc) return rc;
    chain = &tmp;
  }
  if (const char* why = check_columns(bundle, true)) return fail(ART_E_INVALID, std::string("bundle: ") + why);
  if (bundle->n < 0 || bundle->n % n_variants != 0)
    return fail(ART_E_INVALID, "bundle must hold n_variants * n rays");
  DetArgs a;
  a.b = to_dev(bundle);
  a.n = bundle->n / n_variants;
  a.det = det;
  a.x_out = x_out;
  a.y_out = y_out;
  a.l_out = l_out;
  a.partials = chain->d_partials;
  const int bpv = blocks_per_variant(chain, a.n, n_variants);
  if ((size_t)bpv * n_variants > chain->partial_rows)
    return fail(ART_E_INVALID, "n_variants exceeds the chain's variant count");
  cudaStream_t st = (cudaStream_t)stream;
  // Bulk-copy kernel (copy engine + mbarrier ring) whenever the rows of every variant keep the 16-byte alignment
  // the engine needs and there is at least one full tile; the LDGSTS kernel otherwise (ART_B200_DET_LEGACY=1
  // forces it, for A/B measurements).
  static const bool legacy = std::getenv("ART_B200_DET_LEGACY") != nullptr;
  const bool rows_aligned = n_variants == 1 || a.n % 16 == 0;
  const bool flags_aligned = !a.b.alive || aligned16(a.b.alive);
  if (!legacy && rows_aligned && flags_aligned && a.n >= DB_TILE) {
    static std::mutex attr_mutex;
    static bool attr_done[64] = {};
    int dev = 0;
    ART_CUDA(cudaGetDevice(&dev));
    {
      std::lock_guard<std::mutex> lock(attr_mutex);
      if (dev >= 0 && dev < 64 && !attr_done[dev]) {
        ART_CUDA(cudaFuncSetAttribute(detector_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_SMEM_BYTES));
        attr_done[dev] = true;
      }
    }
    // one resident block per SM (ART_DB_MINB): a persistent grid over the tiles
    const long long tiles = a.n / DB_TILE;
    const long long target = (long long)chain->sm_count * ART_DB_MINB * (n_variants > 1 ? 4 : 1);
    long long bulk_bpv = (target + n_variants - 1) / n_variants;
    if (bulk_bpv > tiles) bulk_bpv = tiles;
    if (bulk_bpv < 1) bulk_bpv = 1;
    if ((size_t)bulk_bpv * n_variants > chain->partial_rows)
      return fail(ART_E_INVALID, "n_variants exceeds the chain's variant count");
    detector_bulk_kernel<<<dim3((unsigned)bulk_bpv, n_variants), DB_THREADS, DB_SMEM_BYTES, st>>>(a);
    ART_LAUNCHED();
    if (!moments_out) {
      chain->unfolded_rows = (int)bulk_bpv;
      chain->unfolded_kind = 1;
      return ART_OK;
    }
    fold_kernel<<<n_variants, TPB, 0, st>>>(chain->d_partials, (int)bulk_bpv, 2, nullptr, moments_out);
    ART_LAUNCHED();
    return ART_OK;
  }
  detector_kernel<<<dim3(bpv, n_variants), TPB, DET_STAGE_BYTES, st>>>(a);
  ART_LAUNCHED();
  if (!moments_out) {
    chain->unfolded_rows = bpv;
    chain->unfolded_kind = 1;
    return ART_OK;
  }
  fold_kernel<<<n_variants, TPB, 0, st>>>(chain->d_partials, bpv, 2, nullptr, moments_out);
  ART_LAUNCHED();
  return ART_OK;
}

extern "C" int32_t art_detector_histogram(const ArtBundleView* bundle, const ArtDetector* det, const double* moments,
                                          int32_t nx, int32_t ny, int32_t nt, double wscale, int64_t* hist_out,
                                          void* stream) {
  if (!bundle || !det || !moments || !hist_out) return fail(ART_E_INVALID, "NULL argument");
  if (nx < 1 || ny < 1 || nt < 1 || (int64_t)nx * ny > (int64_t)1 << 26 || nt > 1 << 26)
    return fail(ART_E_INVALID, "bin counts must be in [1, 2^26]");
  if (!(wscale > 0.0)) return fail(ART_E_INVALID, "wscale must be positive");
  if (const char* why = check_columns(bundle, true)) return fail(ART_E_INVALID, std::string("bundle: ") + why);
  if (bundle->n < 0) return fail(ART_E_INVALID, "negative ray count");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t len = (size_t)3 * nx * ny + (size_t)2 * nt;
  ART_CUDA(cudaMemsetAsync(hist_out, 0, len * sizeof(int64_t), st));
  if (bundle->n == 0) return ART_OK;
  HistArgs a;
  a.b = to_dev(bundle);
  a.n = bundle->n;
  a.det = det;
  a.moments = moments;
  a.nx = nx;
  a.ny = ny;
  a.nt = nt;
  a.wscale = wscale;
  a.hist = reinterpret_cast<long long*>(hist_out);
  int dev = 0, sms = 148;
  ART_CUDA(cudaGetDevice(&dev));
  ART_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = hist_smem_bytes(nx, ny, nt);
  if (smem <= (size_t)100 * 1024) {
    // block-private bins in shared memory, two blocks per SM; enough blocks that none sees more than
    // HIST_MAX_RAYS_PER_BLOCK rays (32-bit partial sums)
    if (cudaError_t e = allow_smem(histogram_smem_kernel, smem + 1024)) return fail(ART_E_CUDA, cudaGetErrorString(e));
    long long blocks = (long long)sms * 2;
    const long long need = (bundle->n + HIST_MAX_RAYS_PER_BLOCK - 1) / HIST_MAX_RAYS_PER_BLOCK;
    if (blocks < need) blocks = need;
    const long long most = (bundle->n + 2 * ART_HIST_TPB - 1) / (2 * ART_HIST_TPB);
    if (blocks > most) blocks = most;
    histogram_smem_kernel<<<(unsigned)blocks, ART_HIST_TPB, smem, st>>>(a);
  } else {
    long long blocks = (bundle->n + 255) / 256;
    if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
    histogram_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  }
  ART_LAUNCHED();
  return ART_OK;
}

extern "C" int32_t art_detector_scan_moments(ArtChain* chain, const ArtBundleView* bundle, int32_t n_variants,
                                             const ArtDetector* det, double* scan_out, void* stream) {
  if (!bundle || !det || !scan_out || n_variants < 1) return fail(ART_E_INVALID, "bad argument");
  if (const char* why = check_columns(bundle, true)) return fail(ART_E_INVALID, std::string("bundle: ") + why);
  if (bundle->n < 0 || bundle->n % n_variants != 0)
    return fail(ART_E_INVALID, "bundle must hold n_variants * n rays");
  ArtChain tmp;
  if (!chain) {
    int dev = 0;
    ART_CUDA(cudaGetDevice(&dev));
    int sms = 148;
    ART_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    tmp.sm_count = sms;
    tmp.partial_rows = (size_t)sms * 8 + (size_t)n_variants + 8;
    int32_t rc = global_scratch(tmp.partial_rows, stream, &tmp.d_partials);
    if (rc) return rc;
    chain = &tmp;
  }
  DetArgs a;
  a.b = to_dev(bundle);
  a.n = bundle->n / n_variants;
  a.det = det;
  a.x_out = a.y_out = a.l_out = nullptr;
  a.partials = chain->d_partials;
  const int bpv = blocks_per_variant(chain, a.n, n_variants, 1);
  if ((size_t)bpv * n_variants > chain->partial_rows)
    return fail(ART_E_INVALID, "n_variants exceeds the chain's variant count");
  cudaStream_t st = (cudaStream_t)stream;
  scan_kernel<<<dim3(bpv, n_variants), TPB, 0, st>>>(a);
  ART_LAUNCHED();
  fold_kernel<<<n_variants, TPB, 0, st>>>(chain->d_partials, bpv, 4, nullptr, scan_out);
  ART_LAUNCHED();
  return ART_OK;
}

extern "C" int32_t art_moments_merge(const double* rows, int32_t n_ranks, int32_t n_variants, double* out,
                                     void* stream) {
  if (!rows || !out || n_ranks < 1 || n_variants < 1) return fail(ART_E_INVALID, "bad argument");
  const int total = n_variants * ART_MOMENTS_LEN;
  merge_moments_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rows, n_ranks, n_variants, out);
  ART_LAUNCHED();
  return ART_OK;
}

static int32_t peer_exchange_impl(const uint64_t* peer_bufs, int32_t rank, int32_t world, int32_t kind,
                                  int32_t n_variants, double* rows, double distance, ArtDetector* det_out,
                                  void* stream, const double* partials, int n_partials) {
  if (!peer_bufs || !rows) return fail(ART_E_INVALID, "NULL argument");
  if (world < 1 || world > ART_PEER_MAX_RANKS || rank < 0 || rank >= world)
    return fail(ART_E_INVALID, "rank / world out of range");
  if (kind < 0 || kind > 2) return fail(ART_E_INVALID, "kind must be 0 (central), 1 (moments) or 2 (source extents)");
  if (n_variants < 1 || n_variants > ART_PEER_MAX_VARIANTS)
    return fail(ART_E_INVALID, "n_variants must be in [1, ART_PEER_MAX_VARIANTS]");
  PeerArgs a;
  for (int r = 0; r < ART_PEER_MAX_RANKS; ++r) a.bufs[r] = r < world ? (unsigned long long)peer_bufs[r] : 0ull;
  for (int r = 0; r < world; ++r)
    if (!a.bufs[r]) return fail(ART_E_INVALID, "peer buffer address is NULL");
  a.rank = rank;
  a.world = world;
  a.kind = kind;
  a.n_variants = n_variants;
  a.rows = rows;
  a.distance = distance;
  a.det_out = kind == 0 ? det_out : nullptr;
  a.spin_limit = 10000000ull;  // polling rounds of ~1 us each (system-scope loads of the pending cells): ~10 s
  a.partials = partials;
  a.n_partials = n_partials;
  peer_exchange_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(a);
  ART_LAUNCHED();
  return ART_OK;
}

extern "C" int32_t art_peer_exchange(const uint64_t* peer_bufs, int32_t rank, int32_t world, int32_t kind,
                                     int32_t n_variants, double* rows, double distance, ArtDetector* det_out,
                                     void* stream) {
  return peer_exchange_impl(peer_bufs, rank, world, kind, n_variants, rows, distance, det_out, stream, nullptr, 0);
}

extern "C" int32_t art_peer_exchange_fold(ArtChain* chain, const uint64_t* peer_bufs, int32_t rank, int32_t world,
                                          int32_t kind, double* rows, double distance, ArtDetector* det_out,
                                          void* stream) {
  if (!chain) return fail(ART_E_INVALID, "chain is NULL");
  if (kind != 0 && kind != 1) return fail(ART_E_INVALID, "kind must be 0 (central) or 1 (moments)");
  if (chain->unfolded_rows < 1 || chain->unfolded_kind != kind)
    return fail(ART_E_INVALID, "the chain holds no unfolded rows of this kind (art_trace with ART_TRACE_NO_FOLD / "
                               "art_detector_moments with moments_out == NULL must be the chain's last launch)");
  const int n_partials = chain->unfolded_rows;
  chain->unfolded_rows = 0;
  chain->unfolded_kind = -1;
  return peer_exchange_impl(peer_bufs, rank, world, kind, 1, rows, distance, det_out, stream, chain->d_partials,
                            n_partials);
}

// the u64 words behind the 16-byte cells of an exchange buffer (peer_exchange_kernel, art_kernels.cuh)
static inline uint64_t* peer_words(uint64_t buf, int world) {
  return reinterpret_cast<uint64_t*>(buf + (uint64_t)16 * 2 * (uint64_t)world * PEER_MAX_DOUBLES);
}

extern "C" int32_t art_peer_status(const uint64_t* peer_bufs, int32_t rank, int32_t world, uint64_t* status_out,
                                   void* stream) {
  if (!peer_bufs || !status_out || world < 1 || rank < 0 || rank >= world) return fail(ART_E_INVALID, "bad argument");
  const uint64_t* words = peer_words(peer_bufs[rank], world);
  ART_CUDA(cudaMemcpyAsync(status_out, words + world + 1, sizeof(uint64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  ART_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return ART_OK;
}

extern "C" int32_t art_peer_stats(const uint64_t* peer_bufs, int32_t rank, int32_t world, uint64_t* stats_out,
                                  int32_t reset, void* stream) {
  if (!peer_bufs || !stats_out || world < 1 || rank < 0 || rank >= world) return fail(ART_E_INVALID, "bad argument");
  uint64_t* words = peer_words(peer_bufs[rank], world);
  cudaStream_t st = (cudaStream_t)stream;
  ART_CUDA(cudaMemcpyAsync(stats_out, words + world + 2, ART_PEER_STATS * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (reset) ART_CUDA(cudaMemsetAsync(words + world + 2, 0, ART_PEER_STATS * sizeof(uint64_t), st));
  ART_CUDA(cudaStreamSynchronize(st));
  return ART_OK;
}

// The sweep traces every ray twice (pass 1: central sums, fold + autoplace; pass 2: fused trace +
// detector) and stores nothing per ray.  The alternative -- trace once, keep each variant's final
// bundle (57 B/ray) in L2 and run the detector kernel on it -- was built and measured slower on B200
// (26.3 vs 22.9 ms for 256 variants x 10^6 rays): a variant's rows plus the source bundle do not fit
// L2 together, and per-variant launches leave the GPU idle between four short dependent kernels.
extern "C" int32_t art_sweep(ArtChain* chain, int32_t variant_first, int32_t n_variants, const ArtBundleView* in,
                             uint32_t flags, double distance, double* central_out, ArtDetector* det_out,
                             double* moments_out, void* stream) {
  if (!chain || !det_out || !moments_out || !in) return fail(ART_E_INVALID, "bad argument");
  double* central = central_out ? central_out : chain->d_central;
  const uint32_t f = flags | ART_TRACE_NO_INCIDENCE;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t rc = launch_trace(chain, variant_first, n_variants, in, nullptr, nullptr, f, nullptr, nullptr, nullptr,
                            nullptr, central, nullptr, st, false, distance, det_out);
  if (rc) return rc;
  return launch_trace(chain, variant_first, n_variants, in, nullptr, nullptr, f, det_out, nullptr, nullptr, nullptr,
                      nullptr, moments_out, st);
}

extern "C" int32_t art_delays(const double* l, const uint8_t* alive, int64_t n, int32_t n_variants,
                              const ArtDetector* det, const double* moments, double* delays_out, void* stream) {
  if (!l || !det || !moments || !delays_out || n < 0 || n_variants < 1) return fail(ART_E_INVALID, "bad argument");
  long long bx = (n + TPB - 1) / TPB;
  if (bx < 1) bx = 1;
  if (bx > 148 * 16) bx = 148 * 16;
  delays_kernel<<<dim3((unsigned)bx, n_variants), TPB, 0, (cudaStream_t)stream>>>(l, alive, n, det, moments,
                                                                                 delays_out);
  ART_LAUNCHED();
  return ART_OK;
}

// -------------------------------------------------------------------------------------------------
// sources
// -------------------------------------------------------------------------------------------------
extern "C" int32_t art_source_generate(int32_t kind, int64_t n_total, int64_t first, int64_t count, int64_t stride,
                                       double rho, const double axis[3], const double origin[3],
                                       int64_t n_point_sources, int64_t rays_per_source, double source_radius,
                                       const ArtBundleView* bundle, void* stream) {
  if (!bundle || !axis || !origin) return fail(ART_E_INVALID, "NULL argument");
  if (kind < 0 || kind > 2)
    return fail(ART_E_INVALID, "kind must be 0 (point source), 1 (plane wave) or 2 (extended source)");
  if (kind == 2 && (n_point_sources < 1 || rays_per_source < 1 || n_point_sources * rays_per_source != n_total))
    return fail(ART_E_INVALID, "extended source: n_total must equal n_point_sources * rays_per_source");
  if (n_total < 1 || first < 0 || count < 0 || stride < 1 || bundle->n < count ||
      (count > 0 && first + (count - 1) * stride >= n_total))
    return fail(ART_E_INVALID, "bad index range");
  const bool no_points = !bundle->px && !bundle->py && !bundle->pz;  // point source kept as a uniform origin
  if (!(no_points && kind == 0 && bundle->ux && bundle->uy && bundle->uz)) {
    if (const char* why = check_columns(bundle, true)) return fail(ART_E_INVALID, std::string("bundle: ") + why);
  } else if (!aligned16(bundle->ux) || !aligned16(bundle->uy) || !aligned16(bundle->uz) ||
             (bundle->path && !aligned16(bundle->path))) {
    return fail(ART_E_INVALID, "bundle: ray columns must be 16-byte aligned");
  }
  if (bundle->alive && (reinterpret_cast<uintptr_t>(bundle->alive) & 1))
    return fail(ART_E_INVALID, "bundle: the alive flags must be 2-byte aligned");
  SourceArgs a;
  a.kind = kind;
  a.n_total = n_total;
  a.first = first;
  a.count = count;
  a.stride = stride;
  a.per = kind == 2 ? rays_per_source : 1;
  a.n_ps = kind == 2 ? n_point_sources : 1;
  a.ps_radius = source_radius;
  a.rho = rho;
  const double ez[3] = {0.0, 0.0, 1.0};
  rotation_from_to(ez, axis, a.rot);  // RotationRayList(RayList, ez, Axis), ART/ModuleSource.py:79,167
  for (int i = 0; i < 3; ++i) a.origin[i] = origin[i];
  a.b = to_dev(bundle);
  a.partials = nullptr;
  a.origin_out = nullptr;
  a.origin_stride = 0;
  long long bx = ((count + 1) / 2 + TPB - 1) / TPB;  // two rays per thread
  if (bx < 1) bx = 1;
  if (bx > 148 * 16) bx = 148 * 16;
  source_kernel<<<(unsigned)bx, TPB, 0, (cudaStream_t)stream>>>(a);
  ART_LAUNCHED();
  return ART_OK;
}

static const int kIntensityBlocks = 148 * 4;
static double* g_ext_partials[64] = {};  // per device scratch of kIntensityBlocks x 2 doubles

extern "C" int32_t art_source_extents(const ArtBundleView* bundle, const double axis[3], double* extents_out,
                                      void* stream) {
  if (!bundle || !axis || !extents_out) return fail(ART_E_INVALID, "NULL argument");
  if (!bundle->ux || !bundle->uy || !bundle->uz) return fail(ART_E_INVALID, "bundle: the vector columns are required");
  int dev = 0;
  ART_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(ART_E_UNSUPPORTED, "device index above 63");
  if (!g_ext_partials[dev]) ART_CUDA(cudaMalloc(&g_ext_partials[dev], sizeof(double) * 2 * kIntensityBlocks));
  IntensityArgs a;
  a.b = to_dev(bundle);
  a.n = bundle->n;
  for (int i = 0; i < 3; ++i) a.axis[i] = axis[i];
  a.pass = 0;
  a.mode = 0;
  a.scale = 1.0;
  a.lnf = 0.0;
  a.state = nullptr;
  a.partials = g_ext_partials[dev];
  cudaStream_t st = (cudaStream_t)stream;
  intensity_kernel<<<kIntensityBlocks, TPB, 0, st>>>(a);
  ART_LAUNCHED();
  extents_fold_kernel<<<1, 32, 0, st>>>(g_ext_partials[dev], kIntensityBlocks, extents_out);
  ART_LAUNCHED();
  return ART_OK;
}

extern "C" int32_t art_source_intensity(const ArtBundleView* bundle, const double axis[3], int32_t mode,
                                        double scale, double fraction, void* stream) {
  if (!bundle || !axis) return fail(ART_E_INVALID, "NULL argument");
  if (!bundle->intensity) return fail(ART_E_INVALID, "bundle has no intensity column");
  if (!bundle->ux || !bundle->uy || !bundle->uz) return fail(ART_E_INVALID, "bundle: the vector columns are required");
  if (mode == 1 && !bundle->px) return fail(ART_E_INVALID, "bundle: mode 1 needs the point columns");
  if (mode != 0 && mode != 1) return fail(ART_E_INVALID, "mode must be 0 or 1");
  if (!(fraction > 0.0 && fraction < 1.0)) fraction = 0.1353352832366127;  // 1/e^2, ART/ModuleSource.py:233-238
  IntensityArgs a;
  a.b = to_dev(bundle);
  a.n = bundle->n;
  for (int i = 0; i < 3; ++i) a.axis[i] = axis[i];
  a.pass = 1;
  a.mode = mode;
  a.scale = scale;
  a.lnf = -0.5 * std::log(fraction);
  a.state = nullptr;
  a.partials = nullptr;
  intensity_kernel<<<kIntensityBlocks, TPB, 0, (cudaStream_t)stream>>>(a);
  ART_LAUNCHED();
  return ART_OK;
}

// -------------------------------------------------------------------------------------------------
// host-buffer end-to-end entry point
// -------------------------------------------------------------------------------------------------
static int32_t ensure_workspace(ArtChain* c, size_t n) {
  HostWorkspace& w = c->ws;
  if (!w.stream) {
    ART_CUDA(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    ART_CUDA(cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
    for (auto& e : w.ev) ART_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ART_CUDA(cudaMallocHost(&w.pinned, sizeof(double) * (ART_MOMENTS_LEN + ART_CENTRAL_LEN) + 2 * sizeof(ArtDetector) +
                                          2 * sizeof(uint64_t)));
  }
  if (n > w.cap_n || !w.cols) {
    cudaFree(w.cols);
    cudaFree(w.alive);
    w.cols = nullptr;
    w.alive = nullptr;
    w.cap_n = 0;
    const size_t cap = ((n > 0 ? n : 1) + 1023) & ~size_t(1023);  // n == 0 (an empty shard) still gets columns
    ART_CUDA(cudaMalloc(&w.cols, sizeof(double) * 15 * cap));
    ART_CUDA(cudaMalloc(&w.alive, cap));
    w.cap_n = cap;
  }
  return ART_OK;
}

// The part of a host-level run that follows the trace, in two halves so that the first can be captured in a
// CUDA graph: statistics_enqueue issues the all-reduce of the central sums + Detector.autoplace (or takes the
// caller's detector from the pinned staging area), the detector moments of the stored final bundle, the merge
// over the ranks and the copies of the results (and optionally of the final bundle) to the host;
// statistics_finish waits for the stream, checks the peer exchanges and hands the results out.
// c->d_central holds this rank's folded central sums on entry.
static ArtDetector* pinned_det_in(HostWorkspace& w) {
  double* pc = w.pinned + ART_MOMENTS_LEN;
  ArtDetector* pd = reinterpret_cast<ArtDetector*>(pc + ART_CENTRAL_LEN);
  uint64_t* pw = reinterpret_cast<uint64_t*>(pd + 1);
  return reinterpret_cast<ArtDetector*>(pw + 2);
}

static int32_t statistics_enqueue(ArtChain* c, const ArtBundleView* dout_p, size_t n, bool want_inc,
                                  const ArtBundleView* out_final_host, double distance, bool manual_det,
                                  const uint64_t* peer_bufs, int32_t rank, int32_t world) {
  HostWorkspace& w = c->ws;
  const size_t cap = w.cap_n;
  auto col = [&](int j) { return w.cols + (size_t)j * cap; };
  cudaStream_t st = w.stream;
  const ArtBundleView& dout = *dout_p;
  int32_t rc = ART_OK;
  // sharded bundle: the central sums of all ranks are added (and the detector placed) inside one kernel over
  // peer memory, likewise the moments rows below
  if (manual_det) {
    ART_CUDA(cudaMemcpyAsync(c->d_det, pinned_det_in(w), sizeof(ArtDetector), cudaMemcpyHostToDevice, st));
    if (peer_bufs) {
      rc = art_peer_exchange(peer_bufs, rank, world, 0, 1, c->d_central, distance, nullptr, st);
      if (rc) return rc;
    }
  } else if (peer_bufs) {
    rc = art_peer_exchange(peer_bufs, rank, world, 0, 1, c->d_central, distance, c->d_det, st);
    if (rc) return rc;
  } else {
    rc = art_detector_autoplace(c->d_central, distance, 1, c->d_det, st);
    if (rc) return rc;
  }
  rc = art_detector_moments(c, &dout, 1, c->d_det, nullptr, nullptr, nullptr, c->d_moments, st);
  if (rc) return rc;
  if (peer_bufs) {
    rc = art_peer_exchange(peer_bufs, rank, world, 1, 1, c->d_moments, 0.0, nullptr, st);
    if (rc) return rc;
  }
  double* pm = w.pinned;
  double* pc = pm + ART_MOMENTS_LEN;
  ArtDetector* pd = reinterpret_cast<ArtDetector*>(pc + ART_CENTRAL_LEN);
  ART_CUDA(cudaMemcpyAsync(pm, c->d_moments, sizeof(double) * ART_MOMENTS_LEN, cudaMemcpyDeviceToHost, st));
  ART_CUDA(cudaMemcpyAsync(pc, c->d_central, sizeof(double) * ART_CENTRAL_LEN, cudaMemcpyDeviceToHost, st));
  ART_CUDA(cudaMemcpyAsync(pd, c->d_det, sizeof(ArtDetector), cudaMemcpyDeviceToHost, st));
  if (peer_bufs) {
    uint64_t* pw = reinterpret_cast<uint64_t*>(pd + 1);  // {epoch, status} of this rank's exchange buffer
    const uint64_t* words = peer_words(peer_bufs[rank], world);
    ART_CUDA(cudaMemcpyAsync(pw, words + world, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  if (out_final_host && n) {
    double* hdst[8] = {out_final_host->px, out_final_host->py, out_final_host->pz, out_final_host->ux,
                       out_final_host->uy, out_final_host->uz, out_final_host->path,
                       want_inc ? out_final_host->incidence : nullptr};
    for (int j = 0; j < 8; ++j)
      if (hdst[j]) ART_CUDA(cudaMemcpyAsync(hdst[j], col(7 + j), sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (out_final_host->alive)
      ART_CUDA(cudaMemcpyAsync(out_final_host->alive, w.alive, n, cudaMemcpyDeviceToHost, st));
  }
  return ART_OK;
}

static int32_t statistics_finish(ArtChain* c, double* moments_host, double* central_host, ArtDetector* det_host,
                                 bool peers, int exchanges_before) {
  HostWorkspace& w = c->ws;
  double* pm = w.pinned;
  double* pc = pm + ART_MOMENTS_LEN;
  ArtDetector* pd = reinterpret_cast<ArtDetector*>(pc + ART_CENTRAL_LEN);
  uint64_t* pw = reinterpret_cast<uint64_t*>(pd + 1);
  ART_CUDA(cudaStreamSynchronize(w.stream));
  // A peer that did not arrive leaves the rows of this rank unreduced (peer_exchange_kernel): the status word
  // then holds the epoch that timed out.  This call used the epochs pw[0] - 1 (central sums) and pw[0] (moments)
  // and `exchanges_before` earlier ones.
  if (peers && pw[1] != 0 && pw[1] + 1 + (uint64_t)exchanges_before >= pw[0])
    return fail(ART_E_PEER_TIMEOUT, "peer-memory exchange timed out at epoch " + std::to_string(pw[1]) +
                                        ": a rank did not arrive; the statistics of this call are not reduced");
  if (moments_host)
    for (int j = 0; j < ART_MOMENTS_LEN; ++j) moments_host[j] = pm[j];
  if (central_host)
    for (int j = 0; j < ART_CENTRAL_LEN; ++j) central_host[j] = pc[j];
  if (det_host) *det_host = *pd;
  return ART_OK;
}

static int32_t statistics_tail(ArtChain* c, const ArtBundleView* dout_p, size_t n, bool want_inc,
                               const ArtBundleView* out_final_host, double distance, const ArtDetector* manual_det,
                               double* moments_host, double* central_host, ArtDetector* det_host,
                               const uint64_t* peer_bufs, int32_t rank, int32_t world, int exchanges_before = 0) {
  if (manual_det) *pinned_det_in(c->ws) = *manual_det;
  int32_t rc = statistics_enqueue(c, dout_p, n, want_inc, out_final_host, distance, manual_det != nullptr, peer_bufs,
                                  rank, world);
  if (rc) return rc;
  return statistics_finish(c, moments_host, central_host, det_host, peer_bufs != nullptr, exchanges_before);
}

static int32_t run_host_impl(ArtChain* c, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                             uint32_t flags, double distance, const ArtDetector* manual_det, double* moments_host,
                             double* central_host, ArtDetector* det_host, const uint64_t* peer_bufs, int32_t rank,
                             int32_t world) {
  if (!c || !in_host) return fail(ART_E_INVALID, "NULL argument");
  if (in_host->n < 0) return fail(ART_E_INVALID, "negative ray count");
  if (!in_host->px || !in_host->py || !in_host->pz || !in_host->ux || !in_host->uy || !in_host->uz)
    return fail(ART_E_INVALID, "the point and vector columns (px..uz) are required");
  if (out_final_host && out_final_host->n != in_host->n)
    return fail(ART_E_INVALID, "output bundle must hold n rays");
  ART_CUDA(cudaSetDevice(c->device));
  const size_t n = (size_t)in_host->n;
  int32_t rc = ensure_workspace(c, n);
  if (rc) return rc;
  HostWorkspace& w = c->ws;
  const size_t cap = w.cap_n;
  auto col = [&](int j) { return w.cols + (size_t)j * cap; };
  cudaStream_t st = w.stream;

  if (in_host->path || in_host->alive)
    return fail(ART_E_UNSUPPORTED, "art_run_host starts from a fresh source bundle (no path / alive columns)");
  const bool uniform_point = (flags & ART_TRACE_UNIFORM_POINT) != 0;
  const double* src[7] = {in_host->px, in_host->py, in_host->pz, in_host->ux,
                          in_host->uy, in_host->uz, in_host->intensity};
  ArtBundleView din = {};
  din.n = (int64_t)n;
  double** dst[7] = {&din.px, &din.py, &din.pz, &din.ux, &din.uy, &din.uz, &din.intensity};
  for (int j = 0; j < 7; ++j)
    if (src[j]) *dst[j] = col(j);

  ArtBundleView dout = {};
  dout.n = (int64_t)n;
  dout.px = col(7); dout.py = col(8); dout.pz = col(9);
  dout.ux = col(10); dout.uy = col(11); dout.uz = col(12);
  dout.path = col(13);
  const bool want_inc = out_final_host && out_final_host->incidence && !(flags & ART_TRACE_NO_INCIDENCE);
  dout.incidence = want_inc ? col(14) : nullptr;
  dout.alive = w.alive;
  dout.intensity = din.intensity;

  // Pipelined upload: the bundle goes over PCIe in chunks on the copy stream while the compute stream
  // traces the chunks that have arrived; the chunks' partial rows are folded once at the end.
  const int n_chunks = n >= (size_t)1 << 20 ? 8 : 1;
  size_t chunk = ((n + n_chunks - 1) / n_chunks + 1) & ~size_t(1);  // even: pairs stay aligned
  if (chunk == 0) chunk = 2;
  int chunk_blocks = c->sm_count * 8 / n_chunks;
  if (chunk_blocks < 1) chunk_blocks = 1;
  if (uniform_point)
    for (int j = 0; j < 3; ++j)
      ART_CUDA(cudaMemcpyAsync(col(j), src[j], sizeof(double), cudaMemcpyHostToDevice, w.copy_stream));
  int launched_chunks = 0;
  for (int k = 0; k < n_chunks; ++k) {
    const size_t off = (size_t)k * chunk;
    if (off >= n && !(n == 0 && k == 0)) break;
    const size_t len = n == 0 ? 0 : (off + chunk <= n ? chunk : n - off);
    for (int j = uniform_point ? 3 : 0; j < 7; ++j)
      if (src[j] && len)
        ART_CUDA(cudaMemcpyAsync(col(j) + off, src[j] + off, sizeof(double) * len, cudaMemcpyHostToDevice, w.copy_stream));
    ART_CUDA(cudaEventRecord(w.ev[k], w.copy_stream));
    ART_CUDA(cudaStreamWaitEvent(st, w.ev[k], 0));
    ArtBundleView cin = din, cout = dout;
    cin.n = cout.n = (int64_t)len;
    double** ci[4] = {&cin.ux, &cin.uy, &cin.uz, &cin.intensity};
    for (auto p : ci)
      if (*p) *p += off;
    if (!uniform_point) { cin.px += off; cin.py += off; cin.pz += off; }
    double** co[8] = {&cout.px, &cout.py, &cout.pz, &cout.ux, &cout.uy, &cout.uz, &cout.path, &cout.incidence};
    for (auto p : co)
      if (*p) *p += off;
    cout.alive += off;
    cout.intensity = nullptr;
    rc = launch_trace(c, 0, 1, &cin, &cout, nullptr, flags, nullptr, nullptr, nullptr, nullptr, c->d_central,
                      nullptr, st, false, 0.0, nullptr, chunk_blocks, k * chunk_blocks);
    if (rc) return rc;
    ++launched_chunks;
  }
  fold_kernel<<<1, TPB, 0, st>>>(c->d_partials, launched_chunks * chunk_blocks, 0, c->d_central, nullptr);
  ART_LAUNCHED();
  return statistics_tail(c, &dout, n, want_inc, out_final_host, distance, manual_det, moments_host, central_host,
                         det_host, peer_bufs, rank, world);
}

extern "C" int32_t art_run_host(ArtChain* c, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                                uint32_t flags, double distance, const ArtDetector* manual_det,
                                double* moments_host, double* central_host, ArtDetector* det_host) {
  return run_host_impl(c, in_host, out_final_host, flags, distance, manual_det, moments_host, central_host, det_host,
                       nullptr, 0, 1);
}

extern "C" int32_t art_run_host_sharded(ArtChain* c, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                                        uint32_t flags, double distance, const ArtDetector* manual_det,
                                        double* moments_host, double* central_host, ArtDetector* det_host,
                                        const uint64_t* peer_bufs, int32_t rank, int32_t world) {
  if (!peer_bufs) return fail(ART_E_INVALID, "peer_bufs is NULL (use art_run_host for an unsharded bundle)");
  if (world < 1 || world > ART_PEER_MAX_RANKS || rank < 0 || rank >= world)
    return fail(ART_E_INVALID, "rank / world out of range");
  return run_host_impl(c, in_host, out_final_host, flags, distance, manual_det, moments_host, central_host, det_host,
                       peer_bufs, rank, world);
}

// -------------------------------------------------------------------------------------------------
// descriptor-driven end-to-end entry point: the reference's real host input is SourceProperties
// (ART/ModuleProcessing.py:58-79), not ray columns
// -------------------------------------------------------------------------------------------------
// the launch sequence of art_run_source_host on the workspace stream (no synchronisation, no pageable copies:
// capturable in a CUDA graph); *exchanges_out = peer exchanges issued before the statistics tail
static int32_t run_source_enqueue(ArtChain* c, const ArtSourceDesc* src, uint32_t flags, double distance,
                                  bool manual_det, const uint64_t* peer_bufs, int32_t rank, int32_t world,
                                  int* exchanges_out) {
  HostWorkspace& w = c->ws;
  const size_t n = (size_t)src->count;
  const size_t cap = w.cap_n;
  auto col = [&](int j) { return w.cols + (size_t)j * cap; };
  cudaStream_t st = w.stream;
  int32_t rc = ART_OK;

  const bool point = src->kind == 0;  // one origin for all rays: three single doubles instead of three columns
  const bool weighted = src->intensity != 0;
  ArtBundleView din = {};
  din.n = (int64_t)n;
  din.px = col(0); din.py = col(1); din.pz = col(2);
  din.ux = col(3); din.uy = col(4); din.uz = col(5);
  din.intensity = weighted ? col(6) : nullptr;
  if (point) flags |= ART_TRACE_UNIFORM_POINT;
  else flags &= ~ART_TRACE_UNIFORM_POINT;

  // K0: the Vogel-spiral bundle in closed form, straight into the workspace columns, with the sums of the
  // directions that ApplyGaussianIntensityToRayList's axis needs
  SourceArgs a;
  a.kind = src->kind;
  a.n_total = src->n_total;
  a.first = src->first;
  a.count = src->count;
  a.stride = src->stride;
  a.per = src->kind == 2 ? src->rays_per_source : 1;
  a.n_ps = src->kind == 2 ? src->n_point_sources : 1;
  a.ps_radius = src->source_radius;
  a.rho = src->rho;
  const double ez[3] = {0.0, 0.0, 1.0};
  rotation_from_to(ez, src->axis, a.rot);
  for (int i = 0; i < 3; ++i) a.origin[i] = src->origin[i];
  a.b = to_dev(&din);
  a.b.inten = nullptr;
  a.origin_out = nullptr;
  if (point) {
    a.b.px = a.b.py = a.b.pz = nullptr;
    a.origin_out = col(0);  // the one origin of a point source: written by the generator into col(0..2)[0]
    a.origin_stride = (long long)cap;
  }
  long long bx = (((long long)n + 1) / 2 + TPB - 1) / TPB;  // two rays per thread
  if (bx < 1) bx = 1;
  if (bx > (long long)c->sm_count * 8) bx = (long long)c->sm_count * 8;
  a.partials = weighted ? c->d_partials : nullptr;
  source_kernel<<<(unsigned)bx, TPB, 0, st>>>(a);
  ART_LAUNCHED();
  int exchanges = 0;
  double fraction = 0.0;
  if (weighted) {
    // axis = normalised mean direction of the WHOLE bundle; then the largest angle to it / largest |P|
    fold_kernel<<<1, TPB, 0, st>>>(c->d_partials, (int)bx, 0, c->d_central, nullptr);
    ART_LAUNCHED();
    if (peer_bufs) {
      rc = art_peer_exchange(peer_bufs, rank, world, 0, 1, c->d_central, 0.0, nullptr, st);
      if (rc) return rc;
      ++exchanges;
    }
    source_axis_kernel<<<1, 32, 0, st>>>(c->d_central, w.src_state);
    ART_LAUNCHED();
    ExtentArgs ea;
    ea.src = a;
    ea.state = w.src_state;
    ea.partials = c->d_partials;
    const int iblocks = c->sm_count * 2;
    source_extents_kernel<<<iblocks, TPB, 0, st>>>(ea);
    ART_LAUNCHED();
    source_extents_fold<<<1, 32, 0, st>>>(c->d_partials, iblocks, w.src_state);
    ART_LAUNCHED();
    if (peer_bufs) {
      rc = art_peer_exchange(peer_bufs, rank, world, 2, 1, w.src_state + 3, 0.0, nullptr, st);
      if (rc) return rc;
      ++exchanges;
    }
    fraction = src->intensity_fraction;
    if (!(fraction > 0.0 && fraction < 1.0)) fraction = 0.1353352832366127;  // 1/e^2, ART/ModuleSource.py:233-238
  }

  ArtBundleView dout = {};
  dout.n = (int64_t)n;
  dout.px = col(7); dout.py = col(8); dout.pz = col(9);
  dout.ux = col(10); dout.uy = col(11); dout.uz = col(12);
  dout.path = col(13);
  dout.alive = w.alive;
  ArtBundleView tout = dout;
  dout.intensity = din.intensity;
  // the Gaussian weights are computed inside the trace kernel from the device source state (and left in the
  // intensity column for the detector kernel)
  rc = launch_trace(c, 0, 1, &din, &tout, nullptr, flags | ART_TRACE_NO_INCIDENCE, nullptr, nullptr, nullptr, nullptr,
                    c->d_central, nullptr, st, false, 0.0, nullptr, 0, 0, weighted ? w.src_state : nullptr,
                    weighted ? std::log(fraction) : 0.0);
  if (rc) return rc;
  *exchanges_out = exchanges;
  return statistics_enqueue(c, &dout, n, false, nullptr, distance, manual_det, peer_bufs, rank, world);
}

extern "C" int32_t art_run_source_host(ArtChain* c, const ArtSourceDesc* src, uint32_t flags, double distance,
                                       const ArtDetector* manual_det, double* moments_host, double* central_host,
                                       ArtDetector* det_host, const uint64_t* peer_bufs, int32_t rank, int32_t world) {
  if (!c || !src) return fail(ART_E_INVALID, "NULL argument");
  if (src->kind < 0 || src->kind > 2)
    return fail(ART_E_INVALID, "kind must be 0 (point source), 1 (plane wave) or 2 (extended source)");
  if (src->kind == 2 && (src->n_point_sources < 1 || src->rays_per_source < 1 ||
                         src->n_point_sources * src->rays_per_source != src->n_total))
    return fail(ART_E_INVALID, "extended source: n_total must equal n_point_sources * rays_per_source");
  if (src->n_total < 1 || src->first < 0 || src->count < 0 || src->stride < 1 ||
      (src->count > 0 && src->first + (src->count - 1) * src->stride >= src->n_total))
    return fail(ART_E_INVALID, "bad index range");
  if (peer_bufs && (world < 1 || world > ART_PEER_MAX_RANKS || rank < 0 || rank >= world))
    return fail(ART_E_INVALID, "rank / world out of range");
  ART_CUDA(cudaSetDevice(c->device));
  int32_t rc = ensure_workspace(c, (size_t)src->count);
  if (rc) return rc;
  HostWorkspace& w = c->ws;
  if (!w.src_state) ART_CUDA(cudaMalloc(&w.src_state, sizeof(double) * SRC_STATE_LEN));
  if (manual_det) *pinned_det_in(w) = *manual_det;  // the graph reads the caller's detector from pinned memory

  // The sequence is a dozen short dependent launches: the second call with the same arguments captures it in a
  // CUDA graph (the first one runs eagerly and sets every function attribute), later calls replay the graph.
  // The key holds everything the launches depend on.
  std::string key(reinterpret_cast<const char*>(src), sizeof(ArtSourceDesc));
  {
    const uint64_t extra[6] = {(uint64_t)flags, 0, manual_det ? 1u : 0u, (uint64_t)(uintptr_t)peer_bufs,
                               (uint64_t)(uint32_t)rank << 32 | (uint32_t)world, (uint64_t)w.cap_n};
    key.append(reinterpret_cast<const char*>(extra), sizeof(extra));
    key.append(reinterpret_cast<const char*>(&distance), sizeof(distance));
    if (peer_bufs) key.append(reinterpret_cast<const char*>(peer_bufs), sizeof(uint64_t) * world);
  }
  static const bool no_graph = std::getenv("ART_B200_NO_GRAPH") != nullptr;
  int exchanges = 0;
  if (!no_graph && w.src_graph && key == w.src_key) {
    ART_CUDA(cudaGraphLaunch(w.src_graph, w.stream));
    exchanges = w.src_exchanges;
  } else if (!no_graph && key == w.src_last_key) {
    if (w.src_graph) {
      cudaGraphExecDestroy(w.src_graph);
      w.src_graph = nullptr;
      w.src_key.clear();
    }
    cudaGraph_t graph = nullptr;
    ART_CUDA(cudaStreamBeginCapture(w.stream, cudaStreamCaptureModeThreadLocal));
    rc = run_source_enqueue(c, src, flags, distance, manual_det != nullptr, peer_bufs, rank, world, &exchanges);
    cudaError_t e = cudaStreamEndCapture(w.stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (e != cudaSuccess) return fail(ART_E_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&w.src_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      w.src_graph = nullptr;
      return fail(ART_E_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    w.src_key = key;
    w.src_exchanges = exchanges;
    ART_CUDA(cudaGraphLaunch(w.src_graph, w.stream));
  } else {
    rc = run_source_enqueue(c, src, flags, distance, manual_det != nullptr, peer_bufs, rank, world, &exchanges);
    if (rc) return rc;
  }
  w.src_last_key = key;
  return statistics_finish(c, moments_host, central_host, det_host, peer_bufs != nullptr, exchanges);
}

// RayTracingCalculation for a host caller: host columns in, the bundle after every element (and / or
// the final one) back in host columns.  Allocates and frees its device staging per call.
extern "C" int32_t art_trace_host(ArtChain* c, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                                  const ArtBundleView* out_history_host, uint32_t flags) {
  if (!c || !in_host) return fail(ART_E_INVALID, "NULL argument");
  if (in_host->n < 0) return fail(ART_E_INVALID, "negative ray count");
  if (!in_host->px || !in_host->py || !in_host->pz || !in_host->ux || !in_host->uy || !in_host->uz)
    return fail(ART_E_INVALID, "the point and vector columns (px..uz) are required");
  const int K = c->n_elements;
  const size_t n = (size_t)in_host->n;
  const int n_out = (out_history_host ? K : 0) + (out_final_host ? 1 : 0);
  for (int k = 0; k < n_out; ++k) {
    const ArtBundleView* o = (out_history_host && k < K) ? &out_history_host[k] : out_final_host;
    if (o->n != in_host->n) return fail(ART_E_INVALID, "output bundles must hold n rays");
  }
  ART_CUDA(cudaSetDevice(c->device));
  const size_t cap = (n + 31) & ~size_t(31);
  double* dcols = nullptr;
  uint8_t* dflags = nullptr;
  const size_t ncols = 8 + 8 * (size_t)n_out;
  if (cap) {
    ART_CUDA(cudaMalloc(&dcols, sizeof(double) * ncols * cap));
    if (cudaMalloc(&dflags, (size_t)(1 + n_out) * cap) != cudaSuccess) {
      cudaFree(dcols);
      return fail(ART_E_NOMEM, "out of device memory for the host staging buffers");
    }
  }
  auto cleanup = [&]() {
    cudaFree(dcols);
    cudaFree(dflags);
  };
  auto col = [&](size_t j) { return dcols + j * cap; };
  cudaStream_t st = nullptr;  // legacy default stream: the copies below are synchronous with the host
  ArtBundleView din = {};
  din.n = (int64_t)n;
  const double* src[8] = {in_host->px, in_host->py, in_host->pz, in_host->ux,
                          in_host->uy, in_host->uz, in_host->path, in_host->intensity};
  double** dst[8] = {&din.px, &din.py, &din.pz, &din.ux, &din.uy, &din.uz, &din.path, &din.intensity};
  cudaError_t e = cudaSuccess;
  const bool uniform_point = (flags & ART_TRACE_UNIFORM_POINT) != 0;
  for (int j = 0; j < 8 && e == cudaSuccess; ++j) {
    if (!src[j]) continue;
    *dst[j] = col(j);
    const size_t cnt = (uniform_point && j < 3) ? (cap ? 1 : 0) : n;
    if (cnt) e = cudaMemcpyAsync(col(j), src[j], sizeof(double) * cnt, cudaMemcpyHostToDevice, st);
  }
  if (e == cudaSuccess && in_host->alive) {
    din.alive = dflags;
    if (n) e = cudaMemcpyAsync(dflags, in_host->alive, n, cudaMemcpyHostToDevice, st);
  }
  if (e != cudaSuccess) {
    cleanup();
    return fail(ART_E_CUDA, std::string("H2D copy: ") + cudaGetErrorString(e));
  }
  std::vector<ArtBundleView> dviews(n_out);
  for (int k = 0; k < n_out; ++k) {
    ArtBundleView& d = dviews[k];
    d = ArtBundleView{};
    d.n = (int64_t)n;
    double** f[8] = {&d.px, &d.py, &d.pz, &d.ux, &d.uy, &d.uz, &d.path, &d.incidence};
    for (int j = 0; j < 8; ++j) *f[j] = col(8 + 8 * (size_t)k + j);
    d.alive = dflags + (size_t)(1 + k) * cap;
  }
  const ArtBundleView* dhist = out_history_host ? dviews.data() : nullptr;
  const ArtBundleView* dfinal = out_final_host ? &dviews[n_out - 1] : nullptr;
  int32_t rc = launch_trace(c, 0, 1, &din, dfinal, dhist, flags, nullptr, nullptr, nullptr, nullptr, nullptr,
                            nullptr, st);
  if (rc) {
    cleanup();
    return rc;
  }
  for (int k = 0; k < n_out && e == cudaSuccess && n; ++k) {
    const ArtBundleView* o = (out_history_host && k < K) ? &out_history_host[k] : out_final_host;
    double* h[8] = {o->px, o->py, o->pz, o->ux, o->uy, o->uz, o->path, o->incidence};
    for (int j = 0; j < 8 && e == cudaSuccess; ++j)
      if (h[j]) e = cudaMemcpyAsync(h[j], col(8 + 8 * (size_t)k + j), sizeof(double) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && o->alive)
      e = cudaMemcpyAsync(o->alive, dflags + (size_t)(1 + k) * cap, n, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) return fail(ART_E_CUDA, std::string("D2H copy: ") + cudaGetErrorString(e));
  return ART_OK;
}

// -------------------------------------------------------------------------------------------------
// probes
// -------------------------------------------------------------------------------------------------
extern "C" int32_t art_probe_fp64(double* flops_per_second) {
  if (!flops_per_second) return fail(ART_E_INVALID, "NULL argument");
  int dev = 0;
  ART_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  ART_CUDA(cudaGetDeviceProperties(&prop, dev));
  double* d = nullptr;
  ART_CUDA(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  ART_CUDA(cudaEventCreate(&e0));
  ART_CUDA(cudaEventCreate(&e1));
  const int blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    ART_CUDA(cudaEventRecord(e0));
    probe_fp64_kernel<<<blocks, TPB>>>(d, iters, 1.0 + rep);
    ART_LAUNCHED();
    ART_CUDA(cudaEventRecord(e1));
    ART_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ART_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * TPB / (ms * 1e-3);
    if (rep > 0 && fl > best) best = fl;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *flops_per_second = best;
  return ART_OK;
}

extern "C" int32_t art_probe_hbm(double* bytes_per_second) {
  if (!bytes_per_second) return fail(ART_E_INVALID, "NULL argument");
  const long long n2 = 1ll << 26;  // 64 Mi double2 = 1 GiB per buffer
  double2 *a = nullptr, *b = nullptr;
  ART_CUDA(cudaMalloc(&a, sizeof(double2) * n2));
  ART_CUDA(cudaMalloc(&b, sizeof(double2) * n2));
  ART_CUDA(cudaMemset(a, 0, sizeof(double2) * n2));
  cudaEvent_t e0, e1;
  ART_CUDA(cudaEventCreate(&e0));
  ART_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    ART_CUDA(cudaEventRecord(e0));
    probe_copy_kernel<<<148 * 16, TPB>>>(a, b, n2);
    ART_LAUNCHED();
    ART_CUDA(cudaEventRecord(e1));
    ART_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    ART_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double bw = 2.0 * sizeof(double2) * (double)n2 / (ms * 1e-3);
    if (rep > 0 && bw > best) best = bw;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  *bytes_per_second = best;
  return ART_OK;
}
