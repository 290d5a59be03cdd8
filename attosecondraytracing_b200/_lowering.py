"""Scene objects -> the POD arrays of the C ABI (ArtElementDesc / ArtZernikeDesc)."""
from __future__ import annotations

import ctypes as C

from . import _cabi
from .ModuleDefects import GridDefect, Zernike
from .ModuleMirror import DeformedMirror


class LoweredChain:
    """ctypes arrays describing `n_variants` x `n_elements` elements (variant-major) and their defects."""

    def __init__(self, variants, map_device=None):
        """variants: list (one per variant) of lists of OpticalElement; all variants must share optics.
        map_device: where the arrays of gridded defects live (a torch CUDA device; None = host memory,
        for the CPU check of the device code)."""
        if not variants or not variants[0]:
            raise ValueError("an optical chain needs at least one optical element")
        self.n_variants = len(variants)
        self.n_elements = len(variants[0])
        if self.n_elements > _cabi.ART_MAX_ELEMENTS:
            raise ValueError(f"at most {_cabi.ART_MAX_ELEMENTS} optical elements per chain")
        self.elements = (_cabi.ArtElementDesc * (self.n_variants * self.n_elements))()
        self._keep = []
        zern, maps = [], []
        first_defect, first_map, counts = [], [], []
        for oe in variants[0]:
            optic = oe.type
            first_defect.append(len(zern))
            first_map.append(len(maps))
            if isinstance(optic, DeformedMirror):
                for d in optic.DeformationList:
                    if isinstance(d, Zernike):
                        zern.append(d)
                    elif isinstance(d, GridDefect):
                        maps.append(d)
                    else:
                        raise NotImplementedError(f"{type(d).__name__} defects are not supported by the CUDA path")
            counts.append((len(zern) - first_defect[-1], len(maps) - first_map[-1]))
        for v, oes in enumerate(variants):
            if len(oes) != self.n_elements:
                raise ValueError("all chain variants must have the same number of optical elements")
            for k, oe in enumerate(oes):
                optic = oe.type
                if not (hasattr(optic, "type") and ("Mirror" in optic.type or optic.type == "Mask")):
                    # same failure as ART/ModuleProcessing.py:302-303
                    raise NameError("I don`t recognize the type of optical element " + str(getattr(optic, "type", optic)) + ".")
                d = self.elements[v * self.n_elements + k]
                d.surface, sp = optic._lower()
                d.support, ap = optic.support._lower()
                for i in range(4):
                    d.surface_params[i] = float(sp[i])
                for i in range(6):
                    d.support_params[i] = float(ap[i])
                ctr = optic.get_centre()
                for i in range(3):
                    d.centre[i] = float(ctr[i])
                    d.position[i] = float(oe.position[i])
                    d.normal[i] = float(oe.normal[i])
                    d.majoraxis[i] = float(oe.majoraxis[i])
                nd, nm = counts[k]
                d.n_defects = nd
                d.first_defect = first_defect[k] if nd else 0
                d.n_gridmaps = nm
                d.first_gridmap = first_map[k] if nm else 0
        self.n_gridmaps = len(maps)
        self.gridmaps = (_cabi.ArtGridMapDesc * max(1, self.n_gridmaps))()
        for i, g in enumerate(maps):
            nx, ny, x0, x1, y0, y1, ph, pdx, pdy, keep = g._lower(map_device)
            self._keep.append(keep)
            m = self.gridmaps[i]
            m.nx, m.ny, m.x0, m.x1, m.y0, m.y1, m.h, m.dx, m.dy = nx, ny, x0, x1, y0, y1, ph, pdx, pdy
        self.n_defects = len(zern)
        self.defects = (_cabi.ArtZernikeDesc * max(1, self.n_defects))()
        for i, z in enumerate(zern):
            radius, n, m, c = z._lower()
            an = (C.c_int32 * len(n))(*n)
            am = (C.c_int32 * len(m))(*m)
            ac = (C.c_double * len(c))(*c)
            self._keep += [an, am, ac]
            self.defects[i].radius = radius
            self.defects[i].n_coefficients = len(n)
            self.defects[i].n = an
            self.defects[i].m = am
            self.defects[i].c = ac
