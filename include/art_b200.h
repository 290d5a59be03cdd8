/*
 * art_b200.h -- C ABI of libart_b200.so: the B200-native replacement for the ray-bundle hot path
 * of ART (mightymightys/AttosecondRaytracing v0.93).
 *
 * Plain C: pointers, sizes and POD structs only; no torch / C++ types cross this boundary.
 * Every function returns 0 on success and a negative ART_E_* code on failure; the message of the
 * last failure on the calling thread is available from art_last_error().  No function
 * synchronises the device or allocates device memory unless its comment says so.  All ray
 * columns are DEVICE pointers owned by the caller (in the Python host: torch CUDA tensors) except
 * in the *_host entry points, where they are host pointers.  `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * repository root).
 */
#ifndef ART_B200_H
#define ART_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ART_B200_VERSION 100 /* 0.1.0 */

/* error codes */
#define ART_OK 0
#define ART_E_INVALID (-1)  /* bad argument */
#define ART_E_CUDA (-2)     /* CUDA runtime error (message holds cudaGetErrorString) */
#define ART_E_NOMEM (-3)
#define ART_E_UNSUPPORTED (-4)
#define ART_E_PEER_TIMEOUT (-5) /* multi-GPU: a rank did not arrive at a peer-memory exchange */

/* limits */
#define ART_MAX_ELEMENTS 16 /* elements per chain (longer chains: chain the calls) */

/* surface kinds: the `type` of an optic, ART/ModuleMirror.py + ART/ModuleMask.py */
enum {
  ART_SURF_PLANE = 0,       /* MirrorPlane       ART/ModuleMirror.py:42   params: -                */
  ART_SURF_SPHERICAL = 1,   /* MirrorSpherical   ART/ModuleMirror.py:117  params: |Radius|         */
  ART_SURF_PARABOLIC = 2,   /* MirrorParabolic   ART/ModuleMirror.py:212  params: p (semi latus r.) */
  ART_SURF_TOROIDAL = 3,    /* MirrorToroidal    ART/ModuleMirror.py:391  params: R major, r minor */
  ART_SURF_ELLIPSOIDAL = 4, /* MirrorEllipsoidal ART/ModuleMirror.py:565  params: a, b             */
  ART_SURF_CYLINDRICAL = 5, /* MirrorCylindrical ART/ModuleMirror.py:781  params: |Radius|         */
  ART_SURF_MASK = 6         /* Mask              ART/ModuleMask.py:21     params: -                */
};

/* support (aperture) kinds, ART/ModuleSupport.py; params as the reference constructors take them */
enum {
  ART_SUPP_ROUND = 0,         /* SupportRound(Radius)                                   :46  */
  ART_SUPP_ROUND_HOLE = 1,    /* SupportRoundHole(Radius, RadiusHole, cx, cy)           :109 */
  ART_SUPP_RECT = 2,          /* SupportRectangle(DimX, DimY)                           :200 */
  ART_SUPP_RECT_HOLE = 3,     /* SupportRectangleHole(DimX, DimY, RadiusHole, cx, cy)   :273 */
  ART_SUPP_RECT_RECT_HOLE = 4 /* SupportRectangleRectHole(DimX, DimY, HoleX, HoleY, cx, cy) :373 */
};

/*
 * One optical element = ART/ModuleOpticalElement.py:23 OpticalElement(Type, Position, Normal,
 * MajorAxis) plus the numbers the tracer needs from `Type` (the optic): its surface kind and
 * parameters, `Type.get_centre()`, its support and its Zernike defects.
 * The lab->element rotation is derived inside the library from (normal, majoraxis) exactly as
 * ART/ModuleProcessing.py:289-294 + ART/ModuleGeometry.py:333-343 do (Kahan angle, 1e-10
 * thresholds, the MINUS-identity branch at pi).
 */
typedef struct ArtElementDesc {
  int32_t surface;           /* ART_SURF_*                                              */
  int32_t support;           /* ART_SUPP_*                                              */
  double surface_params[4];  /* see ART_SURF_*                                          */
  double support_params[6];  /* see ART_SUPP_*                                          */
  double centre[3];          /* Type.get_centre(), element frame                        */
  double position[3];        /* OpticalElement.position, lab frame                      */
  double normal[3];          /* OpticalElement.normal                                   */
  double majoraxis[3];       /* OpticalElement.majoraxis                                */
  int32_t n_defects;         /* Zernike defects in DeformedMirror.DeformationList       */
  int32_t first_defect;      /* index of this element's first Zernike defect in the list */
  int32_t n_gridmaps;        /* gridded defects (MeasuredMap / Fourrier) of this element */
  int32_t first_gridmap;     /* index of the first one in the grid-map list              */
} ArtElementDesc;

/*
 * One Zernike defect = ART/ModuleDefects.py:149 Zernike(Support, coefficients):
 * `radius` = Support._CircumCirc(), coefficient c[i] belongs to the reference key (n[i], m[i]),
 * 0 <= m <= n (ART/recursive_zernike_generator.py index convention: m < n/2 sine-like, m > n/2
 * cosine-like, (1,0) = y, (1,1) = x).
 */
typedef struct ArtZernikeDesc {
  double radius;
  int32_t n_coefficients;
  const int32_t* n;
  const int32_t* m;
  const double* c;
} ArtZernikeDesc;

/*
 * One gridded defect = ART/ModuleDefects.py:34 MeasuredMap or :69 Fourrier: a height map and its two
 * slope maps on the regular grid X = linspace(x0, x1, nx), Y = linspace(y0, y1, ny), evaluated by
 * bilinear interpolation exactly like the scipy RegularGridInterpolator(method="linear") objects the
 * reference builds (:45-47, :108-110; points outside the grid are clamped to its edge cell).
 * h, dx, dy: DEVICE pointers owned by the caller, nx*ny doubles each, value at (ix, iy) in [ix*ny + iy]
 * (the arrays the reference hands to the interpolators: np.transpose(deformation) etc.).
 * get_normal of these classes is (dX, dY, 1)/norm -- not negated, unlike Zernike's.
 */
typedef struct ArtGridMapDesc {
  int32_t nx, ny;
  double x0, x1, y0, y1;
  const double* h;
  const double* dx;
  const double* dy;
} ArtGridMapDesc;

/*
 * Structure-of-arrays FP64 ray bundle: the `list[Ray]` of ART/ModuleOpticalRay.py:11.
 * Ray i has point (px,py,pz)[i], unit vector (ux,uy,uz)[i], path[i] = np.sum(Ray.path),
 * incidence[i], intensity[i]; Ray.number is the index i (or number[i] on the host side, the
 * kernels never need it).  alive[i] == 0 marks a ray the reference would have dropped from the
 * list (ART/ModuleMirror.py:932, ART/ModuleMask.py:132); its other columns are unspecified.
 * Nullable columns: path (input: 0), incidence (not produced), intensity, alive (input: all alive).
 * Columns must be 16-byte aligned.
 */
typedef struct ArtBundleView {
  double* px; double* py; double* pz;
  double* ux; double* uy; double* uz;
  double* path;
  double* incidence;
  double* intensity;
  uint8_t* alive;
  int64_t n;
} ArtBundleView;

/* Detector plane, ART/ModuleDetector.py:25.  `rot` takes lab vectors into the detector frame
 * (normal -> ez, RotationPoint semantics) as get_PointList2D does (:212-234); `l0` is the pivot
 * subtracted from optical path lengths before they are squared. */
typedef struct ArtDetector {
  double centre[3];
  double normal[3];
  double refpoint[3];
  double cvec[3]; /* central unit vector of the bundle (autoplace: -normal), the reference axis of
                     ReturnNumericalAperture, ART/ModuleProcessing.py:536-566 */
  double rot[9];
  double l0;
  double n_rays; /* rays the central ray was averaged over (0: detector undefined) */
} ArtDetector;

/* central-ray sums over the surviving final rays, one row of ART_CENTRAL_LEN doubles:
 * sum ux,uy,uz, sum px,py,pz, sum path, count, sum intensity (survivors), sum intensity (all
 * source rays; the denominator of getETransmission, ART/ModuleAnalysisAndPlots.py:62-77) */
enum {
  ART_C_SUX = 0, ART_C_SUY, ART_C_SUZ, ART_C_SPX, ART_C_SPY, ART_C_SPZ, ART_C_SPATH, ART_C_N,
  ART_C_SW_OUT, ART_C_SW_IN,
  ART_CENTRAL_LEN = 10
};

/* detector moments, one row of ART_MOMENTS_LEN doubles (x, y in the detector plane relative to
 * Detector.centre, d = L - l0 with L the total optical path to the plane, w = intensity) */
enum {
  ART_M_N = 0, ART_M_SX, ART_M_SY, ART_M_SXX, ART_M_SYY, ART_M_SD, ART_M_SDD,
  ART_M_SW, ART_M_SWX, ART_M_SWY, ART_M_SWXX, ART_M_SWYY, ART_M_SWD, ART_M_SWDD,
  ART_M_XMIN, ART_M_XMAX, ART_M_YMIN, ART_M_YMAX, ART_M_DMIN, ART_M_DMAX,
  ART_M_TMAX, /* max over rays of tan^2(angle(ray, central vector)/2): ReturnNumericalAperture,
                 ART/ModuleProcessing.py:536-566, NA = sin(2 atan(sqrt(TMAX))) */
  ART_MOMENTS_LEN = 24 /* [21..23] reserved, written as 0 */
};

/* scan sums for the detector-distance optimiser, one row of ART_SCAN_LEN doubles.  Moving the
 * detector by s along the beam (Detector.shiftByDistance(s), ART/ModuleDetector.py:175) moves a ray's
 * in-plane point to (x + s ax, y + s ay) and its path to L + s (1 + gp), with ax, ay the in-plane
 * components of u / (cvec.u) and gp = 1/(cvec.u) - 1, so spot and duration variances are quadratics in
 * s whose coefficients are these sums (d = L - l0; second block: the same weighted by intensity): */
enum {
  ART_S_N = 0, ART_S_SW = 1,
  ART_S_X = 2, ART_S_Y, ART_S_AX, ART_S_AY, ART_S_XX, ART_S_YY, ART_S_AXAX, ART_S_AYAY, ART_S_XAX, ART_S_YAY,
  ART_S_D, ART_S_G, ART_S_DD, ART_S_GG, ART_S_DG,
  ART_S_WEIGHTED = 17, /* ART_S_WEIGHTED + (k - ART_S_X) = intensity-weighted sum k */
  ART_SCAN_LEN = 32
};

/* trace flags */
#define ART_TRACE_IGNORE_DEFECTS 1u /* IgnoreDefects=True of RayTracingCalculation (its default) */
#define ART_TRACE_NO_INCIDENCE 2u   /* do not compute Ray.incidence (saves an atan2 per ray)      */
#define ART_TRACE_UNIFORM_POINT 4u  /* the INPUT bundle's px, py, pz each point to ONE double shared by
                                       all rays (a point source, ART/ModuleSource.py:54: every ray starts
                                       at S): 24 B/ray less to move and to read                         */

#define ART_TRACE_NO_FOLD 8u        /* art_trace: leave the central sums as per-block partial rows in the chain's
                                       scratch (central_out is not written); art_peer_exchange_fold reduces
                                       them inside the exchange kernel -- one launch fewer per step on a
                                       multi-GPU run.  One variant only.                                    */

typedef struct ArtChain ArtChain;

int32_t art_version(void);
const char* art_last_error(void);

/* sizeof(ArtElementDesc), sizeof(ArtZernikeDesc), sizeof(ArtBundleView), sizeof(ArtDetector),
 * sizeof(ArtGridMapDesc), sizeof(ArtSourceDesc) as this library was compiled -- lets a binding in another language verify its
 * struct layouts. */
int32_t art_abi_sizes(int32_t sizes_out[6]);

/* CUDA device count (plumbing for the host; no reference counterpart). */
int32_t art_device_count(int32_t* count);

/*
 * The lab->element rotation of one element, row-major 3x3 (host arithmetic, no GPU needed).
 * Replaces the per-ray RotationRayList(..., n, ez) / RotationRayList(..., mPrime, ex) pair of
 * ART/ModuleProcessing.py:290-294.
 */
int32_t art_element_rotation(const double normal[3], const double majoraxis[3], double rot_out[9]);

/*
 * Build an immutable chain on the current CUDA device: `n_variants` x `n_elements` element
 * descriptions (variant-major; variants share surface kinds but may differ in pose -- the
 * OpticalChain lists of ART/ModuleOpticalChain.py:533 get_OE_loop_list) and the Zernike defects
 * they refer to.  Allocates a few KB of device memory and copies synchronously.
 * Replaces the `optical_elements` argument of RayTracingCalculation, ART/ModuleProcessing.py:250.
 */
int32_t art_chain_create(const ArtElementDesc* elements, int32_t n_elements, int32_t n_variants,
                         const ArtZernikeDesc* defects, int32_t n_defects, const ArtGridMapDesc* gridmaps,
                         int32_t n_gridmaps, ArtChain** chain_out);
int32_t art_chain_destroy(ArtChain* chain);

/*
 * The trace: ART/ModuleProcessing.py:250-313 RayTracingCalculation(source_rays,
 * optical_elements, IgnoreDefects) for variants [variant_first, variant_first + n_variants).
 *   in          source bundle (shared by all variants).
 *   out_final   bundle after the last element; for variant v its rows are [v*in->n, (v+1)*in->n)
 *               of the columns (v counted from variant_first).  May be NULL.
 *   out_history NULL, or `n_elements` views: bundle after element k (same row layout).
 *   central_out NULL, or device pointer to n_variants x ART_CENTRAL_LEN doubles receiving the
 *               sums over the surviving final rays that FindCentralRay needs
 *               (ART/ModuleProcessing.py:464-482), reduced in a fixed order (deterministic).
 * One fused kernel launch; every element of the chain is applied per ray in registers.  in->n < 2^32 - 2
 * (the kernel indexes the rays of one variant with 32 bits; ART_E_INVALID otherwise).
 */
int32_t art_trace(ArtChain* chain, int32_t variant_first, int32_t n_variants, const ArtBundleView* in,
                  const ArtBundleView* out_final, const ArtBundleView* out_history, uint32_t flags,
                  double* central_out, void* stream);

/*
 * Detector.autoplace, ART/ModuleDetector.py:109-137, for n_variants detectors at once, entirely
 * on the device: central (device, n_variants x ART_CENTRAL_LEN) -> det_out (device, n_variants ArtDetector).
 */
int32_t art_detector_autoplace(const double* central, double distance, int32_t n_variants,
                               ArtDetector* det_out, void* stream);

/* Fill an ArtDetector (host struct) from centre / normal / refpoint: computes rot, cvec = -normal;
 * l0 as given.  For manually placed detectors (ARTmain.py:113 setup_detector, ManualDetector
 * branch).  Host arithmetic only. */
int32_t art_detector_make(const double centre[3], const double normal[3], const double refpoint[3],
                          double l0, ArtDetector* det_out);

/*
 * Detector response + statistics sums: Detector.get_PointList3D/2D, get_Delays
 * (ART/ModuleDetector.py:191-279) and the sums behind StandardDeviation /
 * WeightedStandardDeviation / getETransmission / DiameterPointList
 * (ART/ModuleProcessing.py:485-532, ART/ModuleAnalysisAndPlots.py:62, ART/ModuleGeometry.py:164).
 *   bundle   n_variants x n rays (row layout as art_trace's out_final; intensity, if non-NULL,
 *            has n entries shared by all variants).
 *   det      device, n_variants detectors.
 *   x_out, y_out, l_out  NULL or device columns (n_variants x n): in-plane coordinates relative
 *            to Detector.centre and total optical path length L (mm) of each alive ray.
 *   moments_out  device, n_variants x ART_MOMENTS_LEN; NULL (with a chain, one variant): the per-block rows are
 *            left unfolded in the chain's scratch for art_peer_exchange_fold.
 *   chain    lends its reduction scratch (a chain serves one stream at a time); NULL: a
 *            per-device scratch inside the library is used (allocated on first use / growth).
 */
int32_t art_detector_moments(ArtChain* chain, const ArtBundleView* bundle, int32_t n_variants,
                             const ArtDetector* det, double* x_out, double* y_out, double* l_out,
                             double* moments_out, void* stream);

/*
 * Sums behind FindOptimalDistance, ART/ModuleProcessing.py:317-460: one pass over the stored bundle
 * gives the ART_SCAN_LEN sums from which the spot size and duration standard deviations at ANY
 * detector shift follow in closed form, so the reference's scan (4 decades x 20 detector positions,
 * each recomputing every hit) becomes host arithmetic on 32 numbers -- over all rays instead of the
 * 1000-ray random subsample of ARTmain.py:168.  Arguments as art_detector_moments; scan_out: device,
 * n_variants x ART_SCAN_LEN.  Rows are additive over ranks (all-reduce SUM).
 */
int32_t art_detector_scan_moments(ArtChain* chain, const ArtBundleView* bundle, int32_t n_variants,
                                  const ArtDetector* det, double* scan_out, void* stream);

/*
 * Binned detector response -- the data behind SpotDiagram and DelayGraph
 * (ART/ModuleAnalysisAndPlots.py:133-250, 360-440, which scatter-plot every ray of the list; with 1e7+ rays
 * the plots are drawn from these bins instead).  One pass over the stored bundle (one chain variant):
 * every alive ray is intersected with the detector as in art_detector_moments and binned
 *   - by its in-plane point over the bounding box [XMIN,XMAX] x [YMIN,YMAX] of the moments row
 *     (nx x ny uniform bins, numpy.histogram2d's rule: right edges belong to the last bin), and
 *   - by its path-length deviation d = L - l0 over [DMIN,DMAX] (nt bins; delay in fs = (d - SD/N)/c*1e15).
 * hist_out (device, ART_HIST_LEN(nx,ny,nt) int64, zeroed by the call), consecutive blocks:
 *   [nx*ny] ray counts (index ix*ny + iy)       [nx*ny] sum of round(min(w/wscale,1) * 2^26)
 *   [nx*ny] sum of round((d-DMIN)/(DMAX-DMIN) * 2^26)
 *   [nt]    ray counts                            [nt]    sum of round(min(w/wscale,1) * 2^26)
 * Integer accumulation: the result does not depend on the order of the atomic adds, and the histograms of
 * the shards of a bundle add exactly (all-reduce SUM of int64) provided every rank passes the same detector
 * and the same -- merged -- moments row.  moments: device, one row of ART_MOMENTS_LEN doubles.
 */
#define ART_HIST_FIXED_ONE 67108864.0 /* 2^26 */
#define ART_HIST_LEN(nx, ny, nt) (3 * (int64_t)(nx) * (int64_t)(ny) + 2 * (int64_t)(nt))
int32_t art_detector_histogram(const ArtBundleView* bundle, const ArtDetector* det, const double* moments,
                               int32_t nx, int32_t ny, int32_t nt, double wscale, int64_t* hist_out,
                               void* stream);

/* Multi-GPU: merge the moments rows that an all-gather collected from every rank
 * (rows: device, n_ranks x n_variants x ART_MOMENTS_LEN) into out (n_variants x ART_MOMENTS_LEN):
 * sums added in rank order, extents by min / max.  The statistics of the sharded bundle then follow
 * exactly as for a single device. */
int32_t art_moments_merge(const double* rows, int32_t n_ranks, int32_t n_variants, double* out, void* stream);

/*
 * Multi-GPU (one node, NVLink / NVSwitch): the exchange of the central sums or of the moments rows between
 * the ranks done INSIDE one kernel over peer memory instead of an NCCL collective plus a follow-up kernel.
 * Every rank calls it with the same arguments (except rank / rows); peer_bufs[r] is the device address,
 * valid on THIS device, of rank r's exchange buffer -- ART_PEER_BUFFER_BYTES(world) bytes of symmetric /
 * peer-mapped memory (torch.distributed._symmetric_memory, cudaIpc or cuMem handles), zero-filled once
 * before the first call and owned by the caller.
 *   kind 0  rows = n_variants x ART_CENTRAL_LEN, summed over the ranks in rank order (replaces the
 *           all-reduce after art_trace); with det_out non-NULL Detector.autoplace at `distance` follows in
 *           the same kernel (replaces art_detector_autoplace).
 *   kind 1  rows = n_variants x ART_MOMENTS_LEN, merged as art_moments_merge does (replaces the all-gather
 *           + art_moments_merge after art_detector_moments).
 *   kind 2  rows = n_variants x 2 (largest angle to the axis, largest |P| of a source bundle), maximum over
 *           the ranks (the normalisation of ApplyGaussianIntensityToRayList for a sharded source).
 * rows are reduced in place; every rank ends up with bit-identical rows.  The call sequence must be the same
 * on all ranks.  No host synchronisation, CUDA-graph capturable (the epoch is kept in the buffer).  A peer
 * that does not arrive within ~10 s leaves rows unreduced and sets the buffer's status word
 * (art_peer_status reads it back; 0 = fine).
 */
#define ART_PEER_MAX_RANKS 16
#define ART_PEER_MAX_VARIANTS 64
#define ART_PEER_STATS 4 /* exchanges, ns polling for the peers' rows, ns inside the kernel, reserved */
#define ART_PEER_BUFFER_BYTES(world)                                                       \
  ((int64_t)16 * ((int64_t)2 * (world) * ART_PEER_MAX_VARIANTS * ART_MOMENTS_LEN) + \
   (int64_t)8 * ((world) + 2 + ART_PEER_STATS))
int32_t art_peer_exchange(const uint64_t* peer_bufs, int32_t rank, int32_t world, int32_t kind, int32_t n_variants,
                          double* rows, double distance, ArtDetector* det_out, void* stream);
/*
 * art_peer_exchange for ONE variant whose rows have not been folded yet: `chain` ran art_trace with
 * ART_TRACE_NO_FOLD (kind 0) or art_detector_moments with moments_out == NULL (kind 1) as its last launch on
 * this stream; the exchange kernel first folds the chain's per-block partial rows into `rows` (the arithmetic and
 * order of the stand-alone fold) and then proceeds as art_peer_exchange.  Saves the fold launch in front of
 * each of the two exchanges of a multi-GPU step.
 */
int32_t art_peer_exchange_fold(ArtChain* chain, const uint64_t* peer_bufs, int32_t rank, int32_t world, int32_t kind,
                               double* rows, double distance, ArtDetector* det_out, void* stream);
/* Synchronises the stream and returns the status word of this rank's buffer in *status_out. */
int32_t art_peer_status(const uint64_t* peer_bufs, int32_t rank, int32_t world, uint64_t* status_out, void* stream);
/* Synchronises the stream and copies the ART_PEER_STATS counters of this rank's buffer to stats_out: number of
 * exchanges, nanoseconds spent polling for the peers' rows (the skew between the ranks plus the NVLink latency)
 * and nanoseconds inside the exchange kernel, accumulated since the buffer was zeroed -- the breakdown a
 * timeline of the multi-GPU step would show.  With reset != 0 the counters are zeroed afterwards. */
int32_t art_peer_stats(const uint64_t* peer_bufs, int32_t rank, int32_t world, uint64_t* stats_out, int32_t reset,
                       void* stream);

/*
 * Trace and detector in ONE kernel (K1 with K2 as its epilogue), for detectors that are known
 * before the trace (manual detectors, the second pass of a sweep): as art_trace, and in addition
 * every surviving final ray is intersected with det[v] and the moments are reduced; no per-ray
 * bundle has to be stored (out_final may be NULL).  det: device, n_variants detectors.
 */
int32_t art_trace_detect(ArtChain* chain, int32_t variant_first, int32_t n_variants, const ArtBundleView* in,
                         const ArtBundleView* out_final, uint32_t flags, const ArtDetector* det,
                         double* x_out, double* y_out, double* l_out, double* central_out,
                         double* moments_out, void* stream);

/*
 * The batched misalignment sweep: the loop of ARTmain.py:326-332 main() over the chains that
 * ART/ModuleOpticalChain.py:533 get_OE_loop_list builds, each followed by Detector.autoplace at
 * `distance` and GetResultSummary.  Two passes over the (L2-resident) source bundle -- trace +
 * central sums, autoplace on the device, trace + detector moments -- and no per-ray output.
 *   central_out  device, n_variants x ART_CENTRAL_LEN   (nullable)
 *   det_out      device, n_variants ArtDetector
 *   moments_out  device, n_variants x ART_MOMENTS_LEN
 */
int32_t art_sweep(ArtChain* chain, int32_t variant_first, int32_t n_variants, const ArtBundleView* in,
                  uint32_t flags, double distance, double* central_out, ArtDetector* det_out,
                  double* moments_out, void* stream);

/* Per-ray delays in fs relative to the unweighted mean path, Detector.get_Delays
 * ART/ModuleDetector.py:254-279: delay = (l - (l0 + SD/N)) / c * 1e15 for the alive rays.
 * l, alive (nullable), delays_out: n_variants x n device columns; det / moments as returned by
 * art_detector_moments (moments may have been all-reduced over ranks in between). */
int32_t art_delays(const double* l, const uint8_t* alive, int64_t n, int32_t n_variants, const ArtDetector* det,
                   const double* moments, double* delays_out, void* stream);

/*
 * Synthetic source bundles in closed form (device): rays first, first + stride, first + 2 stride, ...
 * (count of them) of the n_total-ray Vogel-spiral bundle written to bundle[0 .. count).  stride = 1
 * gives a contiguous index range; stride = number of ranks deals the rays round-robin, which
 * balances the work when an aperture blocks a contiguous range of spiral indices.
 *   kind 0  PointSource(origin, axis, Divergence, n_total)  ART/ModuleSource.py:54-81, rho = tan(Divergence)
 *   kind 1  PlaneWaveDisk(origin, axis, Radius, n_total)    ART/ModuleSource.py:135-169, rho = Radius
 *           (the reference emits rays 0 .. n_total-2 of the n_total-point spiral)
 *   kind 2  ExtendedSource(origin, axis, Diameter, Divergence, ...)  ART/ModuleSource.py:85-131:
 *           n_point_sources point sources on a Vogel spiral of radius source_radius, each emitting
 *           rays_per_source cone rays (rho = tan(Divergence)); n_total = their product, ray number
 *           k * rays_per_source + l.  (For kinds 0 and 1 the three extra arguments are ignored.)
 * axis: the bundle is rotated ez -> axis with RotationPoint semantics.
 */
int32_t art_source_generate(int32_t kind, int64_t n_total, int64_t first, int64_t count, int64_t stride,
                            double rho, const double axis[3], const double origin[3],
                            int64_t n_point_sources, int64_t rays_per_source, double source_radius,
                            const ArtBundleView* bundle, void* stream);
/* ApplyGaussianIntensityToRayList, ART/ModuleSource.py:219-261, in two calls so that sharded
 * bundles can all-reduce in between: art_source_extents writes {max angle(axis, u), max |P|} of
 * the bundle to extents_out (device, 2 doubles); art_source_intensity then fills bundle->intensity
 * with exp(-2 (q/scale)^2 * (-0.5 ln fraction)), q = tan(angle) (mode 0) or |P| (mode 1). */
int32_t art_source_extents(const ArtBundleView* bundle, const double axis[3], double* extents_out, void* stream);
int32_t art_source_intensity(const ArtBundleView* bundle, const double axis[3], int32_t mode, double scale,
                             double fraction, void* stream);

/*
 * End-to-end convenience with HOST buffers: copies the source bundle to the device, traces
 * variant 0, autoplaces the detector at `distance` (or uses *manual_det if non-NULL), reduces the
 * moments and copies back moments (ART_MOMENTS_LEN), central sums (ART_CENTRAL_LEN), the detector
 * and -- where the corresponding host pointers are non-NULL -- the final bundle.  Synchronises.
 * Uses an internal device workspace that grows to the largest n seen (allocates on growth).
 * This is the call a host without device buffers makes in place of
 * OpticalChain.get_output_rays() + GetResultSummary (ARTmain.py:248-290 run_ART).
 */
int32_t art_run_host(ArtChain* chain, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                     uint32_t flags, double distance, const ArtDetector* manual_det,
                     double* moments_host, double* central_host, ArtDetector* det_host);

/*
 * art_run_host for ONE SHARD of a bundle that is spread over the GPUs of a node (one process / thread per
 * GPU, every rank calling with its own shard): the central sums and the moments rows of all ranks are
 * combined inside the call over peer memory (art_peer_exchange; peer_bufs / rank / world as there), so every
 * rank places the identical detector and returns the statistics of the WHOLE bundle.  An empty shard (n = 0)
 * still takes part in both exchanges.  Returns ART_E_PEER_TIMEOUT when a rank did not arrive at one of this
 * call's exchanges (the outputs are then not written).
 */
int32_t art_run_host_sharded(ArtChain* chain, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                             uint32_t flags, double distance, const ArtDetector* manual_det,
                             double* moments_host, double* central_host, ArtDetector* det_host,
                             const uint64_t* peer_bufs, int32_t rank, int32_t world);

/*
 * The source bundle as the reference's host hands it over: a DESCRIPTION, not rays.  OEPlacement
 * (ART/ModuleProcessing.py:58-79) turns SourceProperties {Divergence, SourceSize, NumberRays, Wavelength}
 * into PointSource / PlaneWaveDisk / ExtendedSource (+ ApplyGaussianIntensityToRayList); this struct
 * carries the arguments of those generators (art_source_generate / art_source_intensity have the details).
 */
typedef struct ArtSourceDesc {
  int32_t kind;               /* 0 PointSource, 1 PlaneWaveDisk, 2 ExtendedSource                     */
  int32_t intensity;          /* 0: every ray has intensity 1; 1: ApplyGaussianIntensityToRayList     */
  int64_t n_total;            /* rays of the WHOLE bundle (PlaneWaveDisk: NbRays; it emits NbRays-1)  */
  int64_t first, count, stride; /* this call's share: rays first, first+stride, ... (count of them)  */
  double rho;                 /* tan(Divergence) (kinds 0, 2) or the disk radius (kind 1)             */
  double axis[3];             /* direction of the bundle                                              */
  double origin[3];           /* S / Centre                                                           */
  int64_t n_point_sources;    /* kind 2 only                                                          */
  int64_t rays_per_source;    /* kind 2 only                                                          */
  double source_radius;       /* kind 2 only                                                          */
  double intensity_fraction;  /* relative intensity at the edge; outside (0,1): 1/e^2                 */
} ArtSourceDesc;

/*
 * End to end from the source DESCRIPTION: generates this call's share of the synthetic bundle on the device
 * (K0, closed form), weights it, traces variant 0, autoplaces the detector at `distance` (or uses
 * *manual_det), reduces the moments and copies moments / central sums / detector back.  The host moves a
 * ~150-byte descriptor in and ~450 bytes out; no ray ever crosses PCIe.  Synchronises; uses the chain's
 * internal workspace like art_run_host.  This is what ARTmain.py:248-290 run_ART does for a config whose
 * source is given by SourceProperties.
 * peer_bufs non-NULL: the bundle is spread over `world` GPUs (every rank passes its own first/count/stride
 * of the same n_total-ray bundle); the axis and extent of the intensity profile, the central sums and the
 * moments are combined over peer memory inside the call as in art_run_host_sharded.
 */
int32_t art_run_source_host(ArtChain* chain, const ArtSourceDesc* source, uint32_t flags, double distance,
                            const ArtDetector* manual_det, double* moments_host, double* central_host,
                            ArtDetector* det_host, const uint64_t* peer_bufs, int32_t rank, int32_t world);

/*
 * ART/ModuleProcessing.py:250 RayTracingCalculation for a caller that holds HOST arrays (the
 * reference's list[Ray] flattened to columns): copies the source bundle to the device, traces
 * variant 0 and copies back the bundle after EVERY element (out_history_host: n_elements views,
 * may be NULL) and / or the final bundle (out_final_host, may be NULL).  Output columns that are
 * NULL are skipped; `alive` tells which rays the reference would have kept.  Synchronises;
 * allocates and frees its device staging buffers per call.
 */
int32_t art_trace_host(ArtChain* chain, const ArtBundleView* in_host, const ArtBundleView* out_final_host,
                       const ArtBundleView* out_history_host, uint32_t flags);

/* FP64 FMA throughput probe for the roofline denominator: runs a register-resident DFMA loop on
 * the whole device and returns measured FLOP/s (2 per FMA).  Synchronises. */
int32_t art_probe_fp64(double* flops_per_second);
/* HBM copy-bandwidth probe (read+write bytes per second of a device-to-device stream copy kernel). */
int32_t art_probe_hbm(double* bytes_per_second);

/* Number of kernel launches issued by this library on this process so far (for bench.py's
 * `gpu_launches`). */
int64_t art_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ART_B200_H */
