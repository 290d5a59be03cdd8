"""Device-side engine: a thin, typed wrapper over the C ABI of libart_b200.so.

`DeviceChain` owns one `ArtChain*` (the packed element table of one or more chain variants on the
current CUDA device) and exposes the entry points with torch tensors as buffers:

    trace()     RayTracingCalculation            ART/ModuleProcessing.py:250
    autoplace() Detector.autoplace               ART/ModuleDetector.py:109
    moments()   Detector response + statistics   ART/ModuleDetector.py:191-279, ModuleProcessing.py:485-532
    sweep()     the loop over misaligned chains  ARTmain.py:326-332
    run_host()  host buffers in, statistics out  ARTmain.py:248 run_ART

Everything runs on `torch.cuda.current_stream()`.  There is no CPU path: without a CUDA device
(or without the built library) these calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._lowering import LoweredChain
from .ModuleOpticalRay import RayBundle

OUT_COLUMNS = ("px", "py", "pz", "ux", "uy", "uz", "path", "incidence")
OUT_COLUMNS_NO_INC = OUT_COLUMNS[:-1]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("attosecondraytracing_b200 traces rays on a CUDA device (B200, sm_100a) only; "
                           "no CUDA device is available and there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


class DeviceChain:
    """Immutable packed chain (possibly many pose variants) resident on one CUDA device."""

    def __init__(self, variants, device=None):
        """variants: list of OpticalElement lists (one per variant), or a single OpticalElement list."""
        if variants and not isinstance(variants[0], (list, tuple)):
            variants = [variants]
        self.device = require_cuda(device)
        self.lowered = LoweredChain(variants, map_device=self.device)
        self.n_elements = self.lowered.n_elements
        self.n_variants = self.lowered.n_variants
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_chain_create(self.lowered.elements, self.n_elements, self.n_variants,
                                                     self.lowered.defects, self.lowered.n_defects,
                                                     self.lowered.gridmaps, self.lowered.n_gridmaps,
                                                     C.byref(self._handle)))

    def close(self):
        if getattr(self, "_handle", None):
            _cabi.lib().art_chain_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _check_bundle(self, bundle):
        if bundle.device.type != "cuda":
            raise RuntimeError("the ray bundle must live on the CUDA device (RayBundle.to('cuda'))")

    def _new_out(self, src, n_variants, want_incidence):
        out = RayBundle(src.n * n_variants, device=self.device,
                        columns=OUT_COLUMNS if want_incidence else OUT_COLUMNS_NO_INC, with_alive=True,
                        wavelength=src.wavelength)
        return out

    def new_output(self, bundle, n_variants=1, want_incidence=True):
        """A reusable output bundle for trace(..., out=...): n_variants x bundle.n rows."""
        return self._new_out(bundle, n_variants, want_incidence)

    def trace(self, bundle, ignore_defects=True, history=False, want_incidence=True, want_central=True,
              variant_first=0, n_variants=None, store_final=True, out=None, central=None, fold=True):
        """Trace `bundle` through the chain.  Returns (bundles, central):
        bundles = list of RayBundle after each element (history=True) or [final bundle];
        for several variants the rows of variant v are [v*n, (v+1)*n).  central = tensor
        (n_variants, 10) of the central-ray sums, or None.
        fold=False (one variant, multi-GPU): the central sums stay as per-block rows in the chain's scratch and
        `central` is NOT written -- PeerExchange.all_reduce_central(..., chain=self) folds them inside the exchange."""
        self._check_bundle(bundle)
        nv = self.n_variants - variant_first if n_variants is None else n_variants
        flags = ((_cabi.TRACE_IGNORE_DEFECTS if ignore_defects else 0) | (0 if want_incidence else _cabi.TRACE_NO_INCIDENCE)
                 | bundle.trace_flags() | (0 if fold else _cabi.TRACE_NO_FOLD))
        outs = []
        hist_arr = None
        final_view = None
        if history:
            outs = [self._new_out(bundle, nv, want_incidence) for _ in range(self.n_elements)]
            hist_arr = (_cabi.ArtBundleView * self.n_elements)(*[o.view() for o in outs])
        elif store_final:
            outs = [out if out is not None else self._new_out(bundle, nv, want_incidence)]
            final_view = C.byref(outs[0].view())
        if central is None and want_central:
            central = torch.empty((nv, _cabi.CENTRAL_LEN), dtype=torch.float64, device=self.device)
        vin = bundle.view()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_trace(self._handle, variant_first, nv, C.byref(vin), final_view, hist_arr,
                                              flags, _ptr(central), _stream()))
        for o in outs:
            if bundle.has("intensity") and nv == 1:
                o.shared_intensity = bundle.col("intensity")
            if bundle.number is not None and nv == 1:
                o.number = bundle.number
            o.invalidate()
        return outs, central

    def count_entering(self, bundle, ignore_defects=True, variant_first=0, n_variants=None):
        """Rays ENTERING each element, per variant -- the interaction count of the metric
        (sum over elements of the rays that reach them).  Returns an int64 tensor (n_variants, n_elements).
        Runs one trace that stores only the per-element alive flags."""
        self._check_bundle(bundle)
        nv = self.n_variants - variant_first if n_variants is None else n_variants
        n, K = bundle.n, self.n_elements
        alive = torch.empty((K, nv * n), dtype=torch.uint8, device=self.device)
        views = (_cabi.ArtBundleView * K)()
        for k in range(K):
            views[k].alive = alive[k].data_ptr()
            views[k].n = nv * n
        flags = (_cabi.TRACE_IGNORE_DEFECTS if ignore_defects else 0) | _cabi.TRACE_NO_INCIDENCE | bundle.trace_flags()
        vin = bundle.view()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_trace(self._handle, variant_first, nv, C.byref(vin), None, views, flags, None,
                                              _stream()))
        surv = alive.reshape(K, nv, n).sum(dim=2, dtype=torch.int64).T  # (nv, K): rays LEAVING element k
        n_in = int(bundle.alive.sum()) if bundle.alive is not None else n
        first = torch.full((nv, 1), n_in, dtype=torch.int64, device=self.device)
        return torch.cat([first, surv[:, :-1]], dim=1), surv[:, -1]

    def autoplace(self, central, distance, det=None):
        """Detector.autoplace for every variant row of `central`; returns an (n_variants, 23) tensor of ArtDetector."""
        nv = central.shape[0]
        if det is None:
            det = torch.empty((nv, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_detector_autoplace(_ptr(central), float(distance), nv, _ptr(det), _stream()))
        return det

    def moments(self, bundle, det, intensity=None, want_points=False, out=None, fold=True):
        """Detector moments of a stored bundle (n_variants x n rows).  Returns (moments, x, y, l).
        fold=False (one variant, multi-GPU): the rows stay unfolded in the chain's scratch (moments is None) for
        PeerExchange.all_reduce_moments(..., chain=self)."""
        self._check_bundle(bundle)
        bundle = bundle.materialize()  # the detector kernel reads per-ray points
        nv = det.shape[0]
        mom = None
        if fold:
            mom = out if out is not None else torch.empty((nv, _cabi.MOMENTS_LEN), dtype=torch.float64, device=self.device)
        x = y = l = None
        if want_points:
            x = torch.full((bundle.n,), float("nan"), dtype=torch.float64, device=self.device)
            y = torch.full_like(x, float("nan"))
            l = torch.full_like(x, float("nan"))
        v = bundle.view()
        if intensity is not None:
            v.intensity = intensity.data_ptr()
        elif not bundle.has("intensity"):
            shared = getattr(bundle, "shared_intensity", None)
            v.intensity = shared.data_ptr() if shared is not None else None
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_detector_moments(self._handle, C.byref(v), nv, _ptr(det), _ptr(x), _ptr(y),
                                                         _ptr(l), _ptr(mom), _stream()))
        return mom, x, y, l

    def histogram(self, bundle, det, moments, bins=(64, 64), delay_bins=128, intensity=None, wscale=1.0, out=None):
        """Binned detector response of a stored bundle (art_detector_histogram): an int64 tensor of
        hist_len(nx, ny, nt) entries -- see `split_histogram`.  `moments`: the (merged) moments row whose
        extents span the bins; every rank of a sharded bundle must pass the same row and detector, then the
        tensors add exactly (all-reduce SUM)."""
        self._check_bundle(bundle)
        bundle = bundle.materialize()
        nx, ny = int(bins[0]), int(bins[1])
        nt = int(delay_bins)
        hist = out if out is not None else torch.empty((_cabi.hist_len(nx, ny, nt),), dtype=torch.int64,
                                                       device=self.device)
        v = bundle.view()
        if intensity is not None:
            v.intensity = intensity.data_ptr()
        elif not bundle.has("intensity"):
            shared = getattr(bundle, "shared_intensity", None)
            v.intensity = shared.data_ptr() if shared is not None else None
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_detector_histogram(C.byref(v), _ptr(det), _ptr(moments), nx, ny, nt,
                                                           float(wscale), _ptr(hist), _stream()))
        return hist

    def scan_moments(self, bundle, det):
        """The ART_SCAN_LEN sums per variant from which spot / duration SDs at any detector shift follow
        (FindOptimalDistance in closed form).  Returns an (n_variants, 32) tensor."""
        self._check_bundle(bundle)
        bundle = bundle.materialize()
        nv = det.shape[0]
        out = torch.empty((nv, _cabi.SCAN_LEN), dtype=torch.float64, device=self.device)
        v = bundle.view()
        if not bundle.has("intensity"):
            shared = getattr(bundle, "shared_intensity", None)
            v.intensity = shared.data_ptr() if shared is not None else None
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_detector_scan_moments(self._handle, C.byref(v), nv, _ptr(det), _ptr(out),
                                                              _stream()))
        return out

    def delays(self, l, alive, det, moments, n_variants=1):
        """Per-ray delays (fs) relative to the unweighted mean path; NaN for dead rays."""
        out = torch.full_like(l, float("nan"))
        n = l.numel() // n_variants
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_delays(_ptr(l), _ptr(alive), n, n_variants, _ptr(det), _ptr(moments),
                                               _ptr(out), _stream()))
        return out

    def trace_detect(self, bundle, det, ignore_defects=True, variant_first=0, want_points=False):
        """Fused trace + detector for known detectors; returns (moments, central, x, y, l)."""
        self._check_bundle(bundle)
        nv = det.shape[0]
        flags = (_cabi.TRACE_IGNORE_DEFECTS if ignore_defects else 0) | _cabi.TRACE_NO_INCIDENCE | bundle.trace_flags()
        mom = torch.empty((nv, _cabi.MOMENTS_LEN), dtype=torch.float64, device=self.device)
        central = torch.empty((nv, _cabi.CENTRAL_LEN), dtype=torch.float64, device=self.device)
        x = y = l = None
        if want_points:
            x = torch.full((bundle.n * nv,), float("nan"), dtype=torch.float64, device=self.device)
            y = torch.full_like(x, float("nan"))
            l = torch.full_like(x, float("nan"))
        vin = bundle.view()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_trace_detect(self._handle, variant_first, nv, C.byref(vin), None, flags,
                                                     _ptr(det), _ptr(x), _ptr(y), _ptr(l), _ptr(central), _ptr(mom),
                                                     _stream()))
        return mom, central, x, y, l

    def sweep(self, bundle, distance, ignore_defects=True, variant_first=0, n_variants=None, out=None):
        """Trace every variant, autoplace its detector at `distance`, reduce its moments.
        Returns (moments (nv,24), central (nv,10), det (nv,23)) device tensors (`out` = such a
        triple to reuse)."""
        self._check_bundle(bundle)
        nv = self.n_variants - variant_first if n_variants is None else n_variants
        flags = (_cabi.TRACE_IGNORE_DEFECTS if ignore_defects else 0) | bundle.trace_flags()
        if out is not None:
            mom, central, det = out
        else:
            mom = torch.empty((nv, _cabi.MOMENTS_LEN), dtype=torch.float64, device=self.device)
            central = torch.empty((nv, _cabi.CENTRAL_LEN), dtype=torch.float64, device=self.device)
            det = torch.empty((nv, _cabi.DETECTOR_DOUBLES), dtype=torch.float64, device=self.device)
        vin = bundle.view()
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_sweep(self._handle, variant_first, nv, C.byref(vin), flags, float(distance),
                                              _ptr(central), _ptr(det), _ptr(mom), _stream()))
        return mom, central, det

    def run_host(self, host_bundle, distance, ignore_defects=True, out_host=None, manual_det=None, peer=None):
        """Host buffers in (pinned or pageable torch CPU tensors), statistics out: H2D, trace,
        autoplace, moments, D2H inside one synchronous C call.  Returns (moments, central, det) numpy.
        peer: a distributed.PeerExchange -- host_bundle is then this rank's shard of a bundle spread over the
        GPUs of the node and the statistics are those of the whole bundle (art_run_host_sharded)."""
        if host_bundle.device.type != "cpu":
            raise RuntimeError("run_host takes a host-resident bundle")
        flags = (_cabi.TRACE_IGNORE_DEFECTS if ignore_defects else 0) | host_bundle.trace_flags()
        if out_host is None or not out_host.has("incidence"):
            flags |= _cabi.TRACE_NO_INCIDENCE
        mom = np.empty(_cabi.MOMENTS_LEN)
        cen = np.empty(_cabi.CENTRAL_LEN)
        det = _cabi.ArtDetector()
        vin = host_bundle.view()
        vout = C.byref(out_host.view()) if out_host is not None else None
        dp = _cabi.c_double_p
        md = C.byref(manual_det) if manual_det is not None else None
        with torch.cuda.device(self.device):
            if peer is None:
                _cabi.check(_cabi.lib().art_run_host(self._handle, C.byref(vin), vout, flags, float(distance), md,
                                                     mom.ctypes.data_as(dp), cen.ctypes.data_as(dp), C.byref(det)))
            else:
                _cabi.check(_cabi.lib().art_run_host_sharded(self._handle, C.byref(vin), vout, flags, float(distance),
                                                             md, mom.ctypes.data_as(dp), cen.ctypes.data_as(dp),
                                                             C.byref(det), peer._ptrs, peer.rank, peer.world))
        if out_host is not None:
            out_host.invalidate()
        return mom, cen, det

    def run_source(self, source_desc, distance, ignore_defects=True, manual_det=None, peer=None):
        """Source DESCRIPTION in (an _cabi.ArtSourceDesc, see ModuleSource.source_descriptor), statistics out:
        the library generates the bundle on the device, weights it, traces, autoplaces the detector and
        reduces the moments inside one synchronous C call (art_run_source_host) -- what run_ART does for a
        config whose source is given by SourceProperties (ARTmain.py:248-290).  Returns (moments, central,
        det).  peer: a distributed.PeerExchange when the bundle is spread over the GPUs of the node (every
        rank passes its own first / count / stride)."""
        flags = _cabi.TRACE_IGNORE_DEFECTS if ignore_defects else 0
        mom = np.empty(_cabi.MOMENTS_LEN)
        cen = np.empty(_cabi.CENTRAL_LEN)
        det = _cabi.ArtDetector()
        dp = _cabi.c_double_p
        md = C.byref(manual_det) if manual_det is not None else None
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().art_run_source_host(
                self._handle, C.byref(source_desc), flags, float(distance), md, mom.ctypes.data_as(dp),
                cen.ctypes.data_as(dp), C.byref(det), peer._ptrs if peer is not None else None,
                peer.rank if peer is not None else 0, peer.world if peer is not None else 1))
        return mom, cen, det


# ----------------------------------------------------------------------------------------------
# statistics from a moments row (host arithmetic on ~24 numbers)
# ----------------------------------------------------------------------------------------------
LIGHTSPEED = 299792458000  # mm/s, ART/ModuleDetector.py:21


def summary_from_moments(m, central=None):
    """Dict of the bundle statistics the reference computes with StandardDeviation /
    WeightedStandardDeviation / GetResultSummary / getETransmission / DiameterPointList /
    ReturnNumericalAperture, from one moments row (and optionally the central-sum row)."""
    m = np.asarray(m, dtype=np.float64)
    n = m[_cabi.M_N]
    out = {"n_rays": int(n)}
    if not n > 0:
        for k in ("SpotSizeSD", "DurationSD", "SpotSizeSD_w", "DurationSD_w", "Diameter", "NA", "mean_path_offset"):
            out[k] = float("nan")
        return out
    fs_per_mm = 1e15 / LIGHTSPEED

    def var(s1, s2, w):
        return max(s2 / w - (s1 / w) ** 2, 0.0)

    vx, vy = var(m[_cabi.M_SX], m[_cabi.M_SXX], n), var(m[_cabi.M_SY], m[_cabi.M_SYY], n)
    out["SpotSizeSD"] = float(np.sqrt(vx + vy))                                    # ModuleProcessing.py:485-507
    out["DurationSD"] = float(np.sqrt(var(m[_cabi.M_SD], m[_cabi.M_SDD], n)) * fs_per_mm)
    sw = m[_cabi.M_SW]
    wvx, wvy = var(m[_cabi.M_SWX], m[_cabi.M_SWXX], sw), var(m[_cabi.M_SWY], m[_cabi.M_SWYY], sw)
    out["SpotSizeSD_w"] = float(np.sqrt(wvx + wvy))                                # ModuleProcessing.py:510-532
    # weighted SD of the delays: the delays are referenced to the UNWEIGHTED mean but the
    # weighted variance is taken about the weighted mean, so the reference point drops out
    out["DurationSD_w"] = float(np.sqrt(var(m[_cabi.M_SWD], m[_cabi.M_SWDD], sw)) * fs_per_mm)
    out["Diameter"] = float(max(abs(m[_cabi.M_XMAX] - m[_cabi.M_XMIN]), abs(m[_cabi.M_YMAX] - m[_cabi.M_YMIN])))
    out["bbox_centre"] = (0.5 * (m[_cabi.M_XMAX] + m[_cabi.M_XMIN]), 0.5 * (m[_cabi.M_YMAX] + m[_cabi.M_YMIN]))
    out["NA"] = float(np.sin(2 * np.arctan(np.sqrt(m[_cabi.M_TMAX]))))             # ModuleProcessing.py:536-566
    out["mean_path_offset"] = float(m[_cabi.M_SD] / n)
    out["delay_min_fs"] = float((m[_cabi.M_DMIN] - m[_cabi.M_SD] / n) * fs_per_mm)
    out["delay_max_fs"] = float((m[_cabi.M_DMAX] - m[_cabi.M_SD] / n) * fs_per_mm)
    if central is not None:
        c = np.asarray(central, dtype=np.float64)
        out["ETransmission"] = float(100 * c[_cabi.C_SW_OUT] / c[_cabi.C_SW_IN])   # ModuleAnalysisAndPlots.py:62-77
    return out


def split_histogram(hist, moments_row, bins=(64, 64), delay_bins=128, wscale=1.0):
    """The int64 vector of `DeviceChain.histogram` (numpy, possibly summed over ranks) as the binned data
    of the reference's SpotDiagram / DelayGraph (ART/ModuleAnalysisAndPlots.py:133, 360):
      x_edges, y_edges  mm, centred on the bounding-box midpoint like get_PointList2DCentre
      spot_count        (nx, ny) rays per bin;  spot_intensity: sum of intensities per bin;
      spot_delay        mean delay (fs, relative to the mean path) of the rays of a bin, NaN where empty
      delay_edges       fs;  delay_count, delay_intensity: (nt,)"""
    nx, ny = int(bins[0]), int(bins[1])
    nt = int(delay_bins)
    h = np.asarray(hist, dtype=np.int64)
    m = np.asarray(moments_row, dtype=np.float64)
    nxy = nx * ny
    one = _cabi.HIST_FIXED_ONE
    cnt = h[:nxy].reshape(nx, ny)
    wsum = h[nxy:2 * nxy].reshape(nx, ny) * (wscale / one)
    dmin, dmax = m[_cabi.M_DMIN], m[_cabi.M_DMAX]
    mean_d = m[_cabi.M_SD] / m[_cabi.M_N] if m[_cabi.M_N] > 0 else 0.0
    to_fs = 1e15 / 299792458000.0  # LightSpeed in mm/s, ART/ModuleDetector.py:21
    with np.errstate(invalid="ignore", divide="ignore"):
        dmean = dmin + (dmax - dmin) * (h[2 * nxy:3 * nxy].reshape(nx, ny) / one) / cnt
    xmid = 0.5 * (m[_cabi.M_XMIN] + m[_cabi.M_XMAX])
    ymid = 0.5 * (m[_cabi.M_YMIN] + m[_cabi.M_YMAX])
    return {
        "x_edges": np.linspace(m[_cabi.M_XMIN], m[_cabi.M_XMAX], nx + 1) - xmid,
        "y_edges": np.linspace(m[_cabi.M_YMIN], m[_cabi.M_YMAX], ny + 1) - ymid,
        "spot_count": cnt,
        "spot_intensity": wsum,
        "spot_delay": (dmean - mean_d) * to_fs,
        "delay_edges": (np.linspace(dmin, dmax, nt + 1) - mean_d) * to_fs,
        "delay_count": h[3 * nxy:3 * nxy + nt].copy(),
        "delay_intensity": h[3 * nxy + nt:3 * nxy + 2 * nt] * (wscale / one),
    }


def scan_statistics(scan, s, weighted=False):
    """(spot size SD in mm, duration SD in fs) with the detector moved by `s` mm along the beam, from one
    row of scan sums (include/art_b200.h ART_S_*): the variances of x + s ax, y + s ay and d + s gp."""
    m = np.asarray(scan, dtype=np.float64)
    o = _cabi.S_WEIGHTED - _cabi.S_X if weighted else 0
    W = m[_cabi.S_SW] if weighted else m[_cabi.S_N]

    def var(k_a, k_b, k_aa, k_bb, k_ab):
        s1 = m[k_a + o] + s * m[k_b + o]
        s2 = m[k_aa + o] + 2.0 * s * m[k_ab + o] + s * s * m[k_bb + o]
        return max(s2 / W - (s1 / W) ** 2, 0.0)

    vx = var(_cabi.S_X, _cabi.S_AX, _cabi.S_XX, _cabi.S_AXAX, _cabi.S_XAX)
    vy = var(_cabi.S_Y, _cabi.S_AY, _cabi.S_YY, _cabi.S_AYAY, _cabi.S_YAY)
    vd = var(_cabi.S_D, _cabi.S_G, _cabi.S_DD, _cabi.S_GG, _cabi.S_DG)
    return float(np.sqrt(vx + vy)), float(np.sqrt(vd) * 1e15 / LIGHTSPEED)


def optimal_shift_from_scan(scan, first_distance, spot_sd0, numerical_aperture, OptFor="intensity", Amplitude=None,
                            Precision=3, IntensityWeighted=False):
    """The search schedule of FindOptimalDistance / _FindOptimalDistanceBIS
    (ART/ModuleProcessing.py:317-460) evaluated on the closed-form statistics: Precision+1 refinements
    of 2*Amplitude/Step detector positions each, fitness SpotSize^2 * Duration ("intensity"),
    Duration or SpotSize; the first minimum wins.  Returns (shift s in mm, OptSizeSpot, OptDuration,
    Amplitude used)."""
    if OptFor not in ("intensity", "size", "spotsize", "duration"):
        raise NameError("I don`t recognize what you want to optimize the detector distance for. OptFor must be "
                        "either 'intensity', 'size' or 'duration'.")
    if Amplitude is None:  # :431-434
        Amplitude = min(4 * np.ceil(2 * spot_sd0 / np.tan(np.arcsin(numerical_aperture))), first_distance)
    Step = Amplitude / 10
    s = 0.0
    opt_spot = opt_dur = float("nan")
    for k in range(Precision + 1):
        amp_k, step_k = Amplitude * 0.1**k, Step * 0.1**k
        s -= amp_k
        n = int(2 * amp_k / step_k)
        spots, durs, fits = [], [], []
        for _ in range(n):
            spot, dur = scan_statistics(scan, s, IntensityWeighted)
            spots.append(spot)
            durs.append(dur)
            fits.append(spot**2 * dur if OptFor == "intensity" else (dur if OptFor == "duration" else spot))
            s += step_k
        ind = fits.index(min(fits))
        opt_spot = spots[ind] if OptFor != "duration" else float("nan")
        opt_dur = durs[ind] if OptFor in ("intensity", "duration") else float("nan")
        s -= (n - ind) * step_k
    return s, opt_spot, opt_dur, Amplitude


def detector_from_row(row):
    """numpy view of one ArtDetector row (23 doubles) as a dict."""
    r = np.asarray(row, dtype=np.float64)
    return {"centre": r[0:3], "normal": r[3:6], "refpoint": r[6:9], "cvec": r[9:12], "rot": r[12:21].reshape(3, 3),
            "l0": r[21], "n_rays": r[22]}
