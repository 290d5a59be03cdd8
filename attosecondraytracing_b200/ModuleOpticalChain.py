"""OpticalChain -- the source bundle plus the list of optical elements (ART/ModuleOpticalChain.py:27).

`get_output_rays()` is the drop-in boundary: it lowers the chain, runs the fused CUDA trace and
returns one lazy RayBundle per optical element.  The source bundle may be a RayBundle (device
columns) or a list of Ray objects (converted once).  Loop lists (`get_OE_loop_list` etc.) share the
source bundle instead of deep-copying it per variant, and `trace_variants` / `sweep_statistics`
evaluate a whole list of misaligned chains in one batched launch.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import ModuleGeometry as mgeo
from . import ModuleOpticalElement as moe
from . import ModuleProcessing as mp
from .ModuleOpticalRay import Ray, RayBundle


class OpticalChain:
    def __init__(self, source_rays, optical_elements, description="", loop_variable_name=None,
                 loop_variable_value=None):
        self.source_rays = source_rays
        # the elements are copied so later edits outside do not change this chain (:118-120); the
        # source bundle is immutable device data and is shared
        self.optical_elements = copy.deepcopy(optical_elements)
        self.description = description
        self.loop_variable_name = loop_variable_name
        self.loop_variable_value = loop_variable_value
        self._output_rays = None
        self._last_key = None

    # ---- properties ----------------------------------------------------------------------------------
    @property
    def source_rays(self):
        return self._source_rays

    @source_rays.setter
    def source_rays(self, source_rays):
        if isinstance(source_rays, RayBundle):
            self._source_rays = source_rays
        elif type(source_rays) == list and all(isinstance(x, Ray) for x in source_rays):
            self._source_rays = RayBundle.from_rays(source_rays)
        else:
            raise TypeError("Source_rays must be list of Ray-objects.")

    @property
    def optical_elements(self):
        return self._optical_elements

    @optical_elements.setter
    def optical_elements(self, optical_elements):
        if not (type(optical_elements) == list and all(isinstance(x, moe.OpticalElement) for x in optical_elements)):
            raise TypeError("Optical_elements must be list of OpticalElement-objects.")
        self._optical_elements = optical_elements

    @property
    def loop_variable_name(self):
        return self._loop_variable_name

    @loop_variable_name.setter
    def loop_variable_name(self, loop_variable_name):
        if not (type(loop_variable_name) == str or loop_variable_name is None):
            raise TypeError("loop_variable_name must be a string.")
        self._loop_variable_name = loop_variable_name

    @property
    def loop_variable_value(self):
        return self._loop_variable_value

    @loop_variable_value.setter
    def loop_variable_value(self, loop_variable_value):
        if not (type(loop_variable_value) in [int, float, np.float64] or loop_variable_value is None):
            raise TypeError("loop_variable_value must be a number of types int or float.")
        self._loop_variable_value = loop_variable_value

    # ---- tracing -----------------------------------------------------------------------------------
    def copy_chain(self):
        return OpticalChain(self.source_rays, self.optical_elements, self.description)

    def _key(self, kwargs):
        return (self.source_rays.content_key(), tuple(oe._pose_key() for oe in self.optical_elements),
                tuple(sorted(kwargs.items())))

    def get_output_rays(self, **kwargs):
        """List of RayBundles, one per optical element; re-traced only when the source bundle, an
        element pose or the keyword arguments (IgnoreDefects) changed (:183-202).  Unlike the
        reference the cache key is O(#elements), not a hash over every ray."""
        key = self._key(kwargs)
        if key != self._last_key or self._output_rays is None:
            self._output_rays = mp.RayTracingCalculation(self.source_rays, self.optical_elements, **kwargs)
            self._last_key = key
        return self._output_rays

    # ---- (mis-)alignment of the source (:219-369) ------------------------------------------------------
    def _source_axes(self):
        cv = mp.FindCentralRay(self.source_rays).vector
        for oe in self.optical_elements:
            if "Mirror" in oe.type.type and np.linalg.norm(np.cross(cv, oe.normal)) > 1e-10:
                return cv, oe.normal
        raise Exception("There doesn't seem to be a non-normal-incidence mirror in this optical chain, "
                        "so you should rather give 'axis' as a numpy-array of length 3.")

    def _replace_source(self, P=None, U=None):
        old = self.source_rays.materialize()
        new = RayBundle(old.n, device=old.device, columns=old._names, wavelength=old.wavelength,
                        storage=old._storage.clone())
        new.number = old.number
        new.alive = old.alive
        if P is not None:
            for i, c in enumerate(("px", "py", "pz")):
                new.col(c).copy_(P[:, i])
        if U is not None:
            for i, c in enumerate(("ux", "uy", "uz")):
                new.col(c).copy_(U[:, i])
        self._source_rays = new

    def shift_source(self, axis, distance: float):
        """Translate the source by `distance` mm along "vert" / "horiz" / "random" or a 3-vector."""
        if type(distance) not in [int, float, np.float64]:
            raise ValueError('The "distance"-argument must be an int or float number.')
        cv, oen = self._source_axes()
        if type(axis) == np.ndarray and len(axis) == 3:
            t = axis
        else:
            perp = np.cross(cv, oen)
            horiz = np.cross(perp, cv)
            if axis == "vert":
                t = perp
            elif axis == "horiz":
                t = horiz
            elif axis == "random":
                t = np.random.uniform(-1, 1, 1) * perp + np.random.uniform(-1, 1, 1) * horiz
            else:
                raise ValueError('The shift direction must be specified by "axis" as one of ["vert", "horiz", "random"].')
        b = self.source_rays
        d = torch.as_tensor(distance * mgeo.Normalize(t), dtype=torch.float64, device=b.device)
        P = torch.stack([b.col("px"), b.col("py"), b.col("pz")], dim=1) + d
        self._replace_source(P=P)

    def tilt_source(self, axis, angle: float):
        """Rotate the source by `angle` degrees about "in_plane" / "out_plane" / "random" or a 3-vector
        through the lab origin."""
        if type(angle) not in [int, float, np.float64]:
            raise ValueError('The "angle"-argument must be an int or float number.')
        cv, oen = self._source_axes()
        if type(axis) == np.ndarray and len(axis) == 3:
            rot_axis = axis
        else:
            a_in = np.cross(cv, oen)
            a_out = np.cross(a_in, cv)
            if axis == "in_plane":
                rot_axis = a_in
            elif axis == "out_plane":
                rot_axis = a_out
            elif axis == "random":
                rot_axis = np.random.uniform(-1, 1, 1) * a_in + np.random.uniform(-1, 1, 1) * a_out
            else:
                raise ValueError('The tilt axis must be specified by as one of ["in_plane", "out_plane", "random"] '
                                 "or as a numpy-array of length 3.")
        b = self.source_rays
        M = torch.as_tensor(mgeo.RotationMatrixAroundAxis(rot_axis, np.deg2rad(angle)), dtype=torch.float64,
                            device=b.device)
        P = torch.stack([b.col("px"), b.col("py"), b.col("pz")], dim=1) @ M.T
        U = torch.stack([b.col("ux"), b.col("uy"), b.col("uz")], dim=1) @ M.T
        self._replace_source(P=P, U=U / torch.linalg.norm(U, dim=1, keepdim=True))

    def get_source_loop_list(self, axis: str, loop_variable_values):
        names = {"tilt_in_plane": "source tilt in-plane (deg)", "tilt_out_plane": "source tilt out-of-plane (deg)",
                 "tilt_random": "source tilt random axis (deg)", "shift_vert": "source shift vertical (mm)",
                 "shift_horiz": "source shift horizontal (mm)", "shift_random": "source shift random-direction (mm)"}
        if axis not in names:
            raise ValueError("For automatic loop-list generation, the axis must be one of " + str(list(names)) + ".")
        if type(loop_variable_values) not in [list, np.ndarray]:
            raise ValueError("For automatic loop-list generation, the loop_variable_values must be a list or a numpy-array.")
        chains = []
        for x in loop_variable_values:
            ch = self.copy_chain()
            ch.loop_variable_name = names[axis]
            ch.loop_variable_value = x
            if axis.startswith("tilt_"):
                ch.tilt_source(axis[5:], x)
            else:
                ch.shift_source(axis[6:], x)
            chains.append(ch)
        return chains

    # ---- (mis-)alignment of the optical elements (:449-657) ---------------------------------------------
    def rotate_OE(self, OEindx: int, axis: str, angle: float):
        if abs(OEindx) > len(self.optical_elements):
            raise ValueError('The "OEnumber"-argument is out of range compared to the length of OpticalChain.optical_elements.')
        if type(angle) not in [int, float, np.float64]:
            raise ValueError('The "angle"-argument must be an int or float number.')
        oe = self.optical_elements[OEindx]
        if axis == "pitch":
            oe.rotate_pitch_by(angle)
        elif axis == "roll":
            oe.rotate_roll_by(angle)
        elif axis == "yaw":
            oe.rotate_yaw_by(angle)
        elif axis in ("random", "rotate_random"):
            oe.rotate_random_by(angle)
        else:
            raise ValueError('The "axis"-argument must be a string out of ["pitch", "roll", "yaw", "random"].')

    def shift_OE(self, OEindx: int, axis: str, distance: float):
        if abs(OEindx) > len(self.optical_elements):
            raise ValueError('The "OEnumber"-argument is out of range compared to the length of OpticalChain.optical_elements.')
        if type(distance) not in [int, float, np.float64]:
            raise ValueError('The "dist"-argument must be an int or float number.')
        oe = self.optical_elements[OEindx]
        if axis == "normal":
            oe.shift_along_normal(distance)
        elif axis == "major":
            oe.shift_along_major(distance)
        elif axis == "cross":
            oe.shift_along_cross(distance)
        elif axis == "random":
            oe.shift_along_random(distance)
        else:
            raise ValueError('The "axis"-argument must be a string out of ["normal", "major", "cross", "random"].')

    def get_OE_loop_list(self, OEindx: int, axis: str, loop_variable_values):
        """One chain per value with element OEindx rotated ("pitch", "roll", "yaw", "rotate_random")
        or shifted ("shift_normal", "shift_major", "shift_cross", "shift_random") by it (:533-614).
        All chains share this chain's source bundle."""
        if abs(OEindx) > len(self.optical_elements):
            raise ValueError('The "OEnumber"-argument is out of range compared to the length of OpticalChain.optical_elements.')
        name = self.optical_elements[OEindx].type.type + "_idx_" + str(OEindx)
        names = {"pitch": " pitch rotation (deg)", "roll": " roll rotation (deg)", "yaw": " yaw rotation (deg)",
                 "rotate_random": " random rotation (deg)", "shift_normal": " shift along normal axis (mm)",
                 "shift_major": " shift along major axis (mm)",
                 "shift_cross": " shift along (normal x major)-direction (mm)",
                 "shift_random": " shift along random axis (mm)"}
        if axis not in names:
            raise ValueError("For automatic loop-list generation, the axis must be one of " + str(list(names)) + ".")
        if type(loop_variable_values) not in [list, np.ndarray]:
            raise ValueError("For automatic loop-list generation, the loop_variable_values must be a list or a numpy-array.")
        chains = []
        for x in loop_variable_values:
            ch = self.copy_chain()
            ch.loop_variable_name = name + names[axis]
            ch.loop_variable_value = x
            if axis in ("pitch", "roll", "yaw", "rotate_random"):
                ch.rotate_OE(OEindx, axis, float(x))
            else:
                ch.shift_OE(OEindx, axis[6:], float(x))
            chains.append(ch)
        return chains

    def get_OE_random_loop_list(self, rotate_std: float, shift_std: float, number_sims: int):
        """number_sims chains with every element randomly rotated / shifted (np.random, :616-657)."""
        name = ("all optical elements randomly rotated with std=" + str(rotate_std)
                + "deg and and shifted with Std=" + str(shift_std) + "mm")
        chains = []
        for i in range(number_sims):
            ch = self.copy_chain()
            ch.loop_variable_name = name
            ch.loop_variable_value = i
            for j in range(len(self.optical_elements)):
                ch.rotate_OE(j, "random", float(np.random.normal(loc=0, scale=rotate_std)))
                ch.shift_OE(j, "random", float(np.random.normal(loc=0, scale=shift_std)))
            chains.append(ch)
        return chains


# --------------------------------------------------------------------------------------------------
# batched evaluation of a loop list: the per-chain loop of ARTmain.py:326-332 as ONE launch
# --------------------------------------------------------------------------------------------------
def sweep_statistics(OpticalChainList, DetectorDistance, IgnoreDefects=True, group=None):
    """Trace every chain of the list (they must share the source bundle and the optics, differing
    only in element poses), autoplace a detector at DetectorDistance behind each final bundle and
    return the statistics per chain: a list of dicts (SpotSizeSD, DurationSD, ETransmission, ...).

    With torch.distributed initialised and `group` given (True = default group) the variants are
    split over the ranks and the rows are all-gathered."""
    from .engine import DeviceChain, summary_from_moments
    import torch.distributed as dist
    chains = list(OpticalChainList)
    src = chains[0].source_rays
    if any(c.source_rays is not src for c in chains):
        raise ValueError("sweep_statistics needs chains that share one source bundle (use get_OE_loop_list)")
    src = mp._as_device_bundle(src)
    nv = len(chains)
    use_dist = group is not None and dist.is_available() and dist.is_initialized()
    grp = None if group is True else group
    rank, world = (dist.get_rank(grp), dist.get_world_size(grp)) if use_dist else (0, 1)
    per = (nv + world - 1) // world
    lo, hi = min(rank * per, nv), min((rank + 1) * per, nv)
    rows = torch.zeros((per * world, 24 + 10), dtype=torch.float64, device=src.device)
    if hi > lo:
        dc = DeviceChain([c.optical_elements for c in chains[lo:hi]], device=src.device)
        mom, central, det = dc.sweep(src, DetectorDistance, ignore_defects=IgnoreDefects)
        rows[rank * per: rank * per + (hi - lo), :24] = mom
        rows[rank * per: rank * per + (hi - lo), 24:] = central
        torch.cuda.current_stream().synchronize()
        dc.close()
    if use_dist:
        dist.all_reduce(rows, op=dist.ReduceOp.SUM, group=grp)  # disjoint rows: an all-gather
    rows = rows.cpu().numpy()[:nv]
    return [summary_from_moments(r[:24], r[24:]) for r in rows]
