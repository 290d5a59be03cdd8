"""
TEST INFRASTRUCTURE ONLY -- the scene catalogue shared by the golden generator, the tests and
the benchmark.  A scene is plain data (JSON-able): which optics, at which distances / incidence
angles (the arguments of the reference's `OEPlacement`, ART/ModuleProcessing.py:133), which
misalignments are applied afterwards, and where the detector goes.

The five BASELINE configs are `cfg1` .. `cfg5`; the rest cover every surface
and support class of the path.
"""
import numpy as np


def toroidal_radii(focal, incidence_deg):
    """ReturnOptimalToroidalRadii, ART/ModuleMirror.py:533-561."""
    a = incidence_deg * np.pi / 180
    return 2 * focal * (1 / np.cos(a) - np.cos(a)), 2 * focal * np.cos(a)


def zernike_table(max_order=20, seed=0, scale=1e-4):
    """The documented cfg4 coefficient table: every (n,m), 2<=n<=max_order,
    c_nm = default_rng(seed).normal() * scale / (n+1) mm, drawn in (n, m) lexicographic order."""
    rng = np.random.default_rng(seed)
    coeffs = []
    for n in range(2, max_order + 1):
        for m in range(0, n + 1):
            coeffs.append([n, m, float(rng.normal() * scale / (n + 1))])
    return coeffs


_TOR = toroidal_radii(500, 80)
_TOR600 = toroidal_radii(600, 80)


def _tor(support):
    return {"kind": "toroidal", "majorradius": _TOR[0], "minorradius": _TOR[1], "support": support}


SCENES = {
    # ---- BASELINE config 1: examples/CONFIG_singleparabola.py:20-56 (as shipped) -------------
    "cfg1_par": {
        "source": {"Divergence": 0, "SourceSize": 50, "Wavelength": 800e-6, "NumberRays": 1000},
        "optics": [{"kind": "parabolic", "feff": 100, "offaxisangle_deg": 90,
                    "support": ["roundhole", 30, 5, 10, 5]}],
        "distances": [200], "incidences": [0.0], "plane_angles": [0],
        "post": [{"op": "rotate_roll_by", "element": 0, "value": float(np.rad2deg(50e-6))}],
        "detector_distance": 100,
    },
    # ---- BASELINE config 2: examples/CONFIG_toroidal2f-2f.py:18-66, aligned chain ------------
    "cfg2_tor2f": {
        "source": {"Divergence": 30e-3 / 2, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 1000},
        "optics": [_tor(["rect", 300, 50])],
        "distances": [1000], "incidences": [80], "plane_angles": [0],
        "post": [], "detector_distance": 1000,
    },
    # list entry 0 of the file's roll sweep (roll -0.5 deg)
    "cfg2_tor2f_roll": {
        "source": {"Divergence": 30e-3 / 2, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 1000},
        "optics": [_tor(["rect", 300, 50])],
        "distances": [1000], "incidences": [80], "plane_angles": [0],
        "post": [{"op": "rotate_roll_by", "element": 0, "value": -0.5}], "detector_distance": 1000,
    },
    # ---- BASELINE config 3: examples/CONFIG_2toroidals_f-x-f.py:19-50 with d3 = 500 ----------
    "cfg3_2tor": {
        "source": {"Divergence": 50e-3 / 2, "SourceSize": 0, "Wavelength": 80e-6, "NumberRays": 1000},
        "optics": [{"kind": "mask", "support": ["roundhole", 20, 7, 0, 0]},
                   _tor(["rect", 150, 32]), _tor(["rect", 150, 32])],
        "distances": [400, 100, 500], "incidences": [0, 80, -80], "plane_angles": [0, 0, 0],
        "post": [], "detector_distance": 500,
    },
    # ---- BASELINE config 4: examples/CONFIG_deformed.py:19-57 geometry, Zernike defect -------
    "cfg4_zern": {
        "source": {"Divergence": 0, "SourceSize": 100, "Wavelength": 800e-6, "NumberRays": 1000},
        "optics": [{"kind": "parabolic", "feff": 25.4, "offaxisangle_deg": 0, "support": ["rect", 40, 40],
                    "defects": [{"kind": "zernike", "coefficients": zernike_table(20)}]}],
        "distances": [15], "incidences": [0], "plane_angles": [0],
        "post": [], "detector_distance": 25.4,
    },
    # ---- BASELINE config 5: examples/CONFIG_CollimatingTelescope.py:15-42 --------------------
    "cfg5_tele": {
        "source": {"Divergence": 2.2e-3, "SourceSize": 0, "Wavelength": 780e-6, "NumberRays": 1000},
        "optics": [{"kind": "spherical", "radius_signed": -1500, "support": ["round", 25]},
                   {"kind": "spherical", "radius_signed": 2500, "support": ["round", 25]},
                   {"kind": "parabolic", "feff": 100, "offaxisangle_deg": 90, "support": ["round", 25]}],
        "distances": [5000, 598, 1000], "incidences": [5, 3.4, 0.04], "plane_angles": [0, 0, 0],
        "post": [], "detector_distance": 100,
    },
    # ---- coverage of the remaining surfaces / supports -------------
    "ell_ab": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "ellipsoidal", "SemiMajorAxis": 1000, "SemiMinorAxis": 173.64817766693042,
                    "support": ["rect", 200, 40]}],
        "distances": [1000], "incidences": [80], "plane_angles": [0], "post": [], "detector_distance": 1000,
    },
    "ell_offaxis": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "ellipsoidal", "OffAxisAngle": 20, "f_object": 600, "f_image": 1400,
                    "support": ["rect", 200, 40]}],
        "distances": [600], "incidences": [80], "plane_angles": [0], "post": [], "detector_distance": 1400,
    },
    "plane_round": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "plane", "support": ["round", 30]}],
        "distances": [500], "incidences": [45], "plane_angles": [0], "post": [], "detector_distance": 300,
    },
    "plane_round_clip": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "plane", "support": ["round", 8]}],
        "distances": [500], "incidences": [45], "plane_angles": [0], "post": [], "detector_distance": 300,
    },
    "cyl_cc": {
        "source": {"Divergence": 0, "SourceSize": 20, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "cylindrical", "radius_signed": 2000, "support": ["rect", 60, 60]}],
        "distances": [300], "incidences": [30], "plane_angles": [0], "post": [], "detector_distance": 500,
    },
    "cyl_cx": {
        "source": {"Divergence": 0, "SourceSize": 20, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "cylindrical", "radius_signed": -2000, "support": ["rect", 60, 60]}],
        "distances": [300], "incidences": [30], "plane_angles": [0], "post": [], "detector_distance": 500,
    },
    "sph_recthole": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "spherical", "radius_signed": 1000, "support": ["recthole", 60, 60, 5, 9, -8]}],
        "distances": [1000], "incidences": [2], "plane_angles": [0], "post": [], "detector_distance": 1000,
    },
    "mask_rrh_plane": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "mask", "support": ["rectrecthole", 30, 20, 10, 6, 2, -1]},
                   {"kind": "plane", "support": ["rect", 60, 30]}],
        "distances": [500, 300], "incidences": [0, 30], "plane_angles": [0, 0], "post": [],
        "detector_distance": 200,
    },
    "tor_twisted": {
        "source": {"Divergence": 25e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 500},
        "optics": [{"kind": "toroidal", "majorradius": _TOR600[0], "minorradius": _TOR600[1],
                    "support": ["rect", 200, 30]},
                   {"kind": "toroidal", "majorradius": _TOR600[0], "minorradius": _TOR600[1],
                    "support": ["rect", 200, 30]}],
        "distances": [600, 600], "incidences": [80, -80], "plane_angles": [0, 30], "post": [],
        "detector_distance": 600,
    },
    # misaligned telescope variants (the cfg5 sweep axis and two other degrees of freedom)
    "tele_pitch": {
        "base": "cfg5_tele", "post": [{"op": "rotate_pitch_by", "element": 2, "value": 0.02}],
    },
    "tele_yaw_shift": {
        "base": "cfg5_tele", "post": [{"op": "rotate_yaw_by", "element": 1, "value": 1.5},
                                      {"op": "shift_along_cross", "element": 0, "value": 0.3},
                                      {"op": "shift_along_normal", "element": 2, "value": -0.2},
                                      {"op": "shift_along_major", "element": 1, "value": 0.1}],
    },
    # two stacked Zernike defects on a sphere with a round support (normal_add chain)
    "sph_zern2": {
        "source": {"Divergence": 20e-3, "SourceSize": 0, "Wavelength": 50e-6, "NumberRays": 300},
        "optics": [{"kind": "spherical", "radius_signed": 1000, "support": ["round", 25],
                    "defects": [{"kind": "zernike", "coefficients": zernike_table(6, seed=1, scale=2e-4)},
                                {"kind": "zernike", "coefficients": [[1, 0, 3e-5], [3, 1, -1e-4], [4, 2, 5e-5]]}]}],
        "distances": [1000], "incidences": [3], "plane_angles": [0], "post": [], "detector_distance": 1000,
    },
}


def measured_map(nx, ny, amplitude):
    """A deterministic synthetic 'measured' height map (mm), shape (nx, ny) as MeasuredMap takes it."""
    i = np.arange(nx)[:, None] / (nx - 1)
    j = np.arange(ny)[None, :] / (ny - 1)
    return amplitude * (np.sin(5.1 * i + 0.3) * np.cos(3.7 * j - 0.2) + 0.5 * np.sin(9.0 * i * j) + 0.25 * (i - 0.5) ** 2)


SCENES.update({
    # gridded defects: ModuleDefects.Fourrier (+ a Zernike term) on the cfg4 geometry.
    "par_fourier": {
        "source": {"Divergence": 0, "SourceSize": 100, "Wavelength": 800e-6, "NumberRays": 800},
        "optics": [{"kind": "parabolic", "feff": 25.4, "offaxisangle_deg": 0, "support": ["rect", 40, 40],
                    "defects": [{"kind": "fourier", "rms": 1e-4, "slope": -2, "smallest": 1.0, "seed": 3},
                                {"kind": "zernike", "coefficients": [[2, 1, 5e-5], [3, 0, -4e-5]]}]}],
        "distances": [15], "incidences": [0], "plane_angles": [0], "post": [], "detector_distance": 25.4,
    },
})


SCENES.update({
    # ModuleDefects.MeasuredMap on the same geometry: a deterministic synthetic height map.  The stock reference
    # cannot construct it under numpy >= 2; pinned through the documented compatibility patch (oracle/make_ref.py,
    # oracle/gen_golden_gridmap.py).  Square: the reference hands (X[nx], Y[ny]) grids but TRANSPOSED (ny, nx) value
    # arrays to its interpolators (ART/ModuleDefects.py:45-47), which only fits for nx == ny
    "par_measured": {
        "source": {"Divergence": 0, "SourceSize": 100, "Wavelength": 800e-6, "NumberRays": 800},
        "optics": [{"kind": "parabolic", "feff": 25.4, "offaxisangle_deg": 0, "support": ["rect", 40, 40],
                    "defects": [{"kind": "measuredmap", "nx": 53, "ny": 53, "amplitude": 2e-4}]}],
        "distances": [15], "incidences": [0], "plane_angles": [0], "post": [], "detector_distance": 25.4,
    },
})


def resolve(name):
    """Return the full scene dict for `name` (following 'base')."""
    s = dict(SCENES[name])
    if "base" in s:
        base = dict(SCENES[s.pop("base")])
        base.update(s)
        s = base
    s["name"] = name
    return s


# Full-size workloads of the BASELINE configs: scene + ray count (+ sweep) -- used by the
# scale-subset goldens, the -m gpu tests and bench.py.
WORKLOADS = {
    "cfg2": {"scene": "cfg2_tor2f", "rays": 10_000_000},
    "cfg3": {"scene": "cfg3_2tor", "rays": 100_000_000},
    "cfg4": {"scene": "cfg4_zern", "rays": 50_000_000},
    "cfg5": {"scene": "cfg5_tele", "rays": 1_000_000,
             "sweep": {"element": 2, "axis": "pitch", "lo": -0.05, "hi": 0.05, "n": 1024}},
}
