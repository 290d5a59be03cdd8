"""Rays: the per-ray record `Ray` of the reference (ART/ModuleOpticalRay.py:11) and the
structure-of-arrays FP64 `RayBundle` that replaces `list[Ray]` on the GPU.

A RayBundle owns torch tensors (the device buffers the CUDA kernels read and write):
columns px, py, pz, ux, uy, uz, path (sum of the reference's path tuple), incidence, intensity, a
one-byte `alive` flag per ray (0 = the reference would have dropped the ray from its list) and
optionally explicit ray numbers (default: number = index).  It behaves like the list the
reference returns -- len(), indexing, iteration give `Ray` objects of the surviving rays in
order -- but materialises them lazily, so 10^8-ray bundles never become Python objects.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

_COLUMNS = ("px", "py", "pz", "ux", "uy", "uz", "path", "incidence", "intensity")


class Ray:
    """One ray (ART/ModuleOpticalRay.py:11): point, unit vector, path tuple, number, wavelength,
    incidence, intensity.  The vector is normalised whenever it is set, as in the reference."""

    __slots__ = ("_point", "_vector", "_path", "_number", "_wavelength", "_incidence", "_intensity")

    def __init__(self, Point, Vector, Path=(0.0,), Number=None, Wavelength=None, Incidence=None, Intensity=None):
        self.point = Point
        self.vector = Vector
        self._path = Path
        self._wavelength = Wavelength
        self._incidence = Incidence
        self._intensity = Intensity
        if not (type(Number) == int or Number is None):
            raise TypeError("Ray Number must be an integer.")
        self._number = Number

    @property
    def point(self):
        return self._point

    @point.setter
    def point(self, Point):
        if not (type(Point) == np.ndarray and len(Point) == 3):
            raise TypeError("Ray Point must be a 3D numpy.ndarray, but it is  %s." % type(Point))
        self._point = Point

    @property
    def vector(self):
        return self._vector

    @vector.setter
    def vector(self, Vector):
        if not (type(Vector) == np.ndarray and len(Vector) == 3 and np.linalg.norm(Vector) > 1e-9):
            raise TypeError("Ray Vector must be a 3D numpy.ndarray with finite length.")
        self._vector = Vector / np.linalg.norm(Vector)

    @property
    def path(self):
        return self._path

    @path.setter
    def path(self, Path):
        self._path = Path

    @property
    def number(self):
        return self._number

    @property
    def wavelength(self):
        return self._wavelength

    @wavelength.setter
    def wavelength(self, Wavelength):
        if not (type(Wavelength) in (int, float, np.float64) or Wavelength is None):
            raise TypeError("Ray Wavelength must be int or float or None.")
        self._wavelength = Wavelength

    @property
    def incidence(self):
        return self._incidence

    @incidence.setter
    def incidence(self, Incidence):
        if not (type(Incidence) in (int, float, np.float64) or Incidence is None):
            raise TypeError("Ray Incidence must be a float or None.")
        self._incidence = Incidence

    @property
    def intensity(self):
        return self._intensity

    @intensity.setter
    def intensity(self, Intensity):
        if not (type(Intensity) in (int, float, np.float64) or Intensity is None):
            raise TypeError("Ray Intensity must be int or float or None.")
        self._intensity = Intensity

    def copy_ray(self):
        return Ray(self.point, self.vector, self.path, self.number, self.wavelength, self.incidence, self.intensity)

    def __hash__(self):
        return hash(tuple(self.point.ravel()) + tuple(self.vector.ravel())
                    + (self.path, self.number, self.wavelength, self.incidence, self.intensity))


def _padded(n):
    return (n + 31) & ~31


class RayBundle:
    """Structure-of-arrays FP64 ray bundle (see module docstring)."""

    def __init__(self, n, device="cpu", columns=_COLUMNS, with_alive=False, wavelength=None, storage=None):
        self.n = int(n)
        self.device = torch.device(device)
        self.wavelength = wavelength
        self.number = None  # optional int64 tensor of explicit ray numbers
        self.origin = None  # point source: ONE (3,) point shared by all rays instead of px/py/pz columns
        self._names = tuple(columns)
        pad = max(_padded(self.n), 32)
        if storage is None:
            storage = torch.empty((len(self._names), pad), dtype=torch.float64, device=self.device)
        self._storage = storage  # rows are 256-byte aligned -> 128-bit column accesses are legal
        # the trace kernel writes every ray's flag, so no initialisation pass is needed on the device
        self.alive = ((torch.empty if self.device.type == "cuda" else torch.ones)(pad, dtype=torch.uint8, device=self.device)[: self.n]
                      if with_alive else None)
        self._index = None  # cached indices of the alive rays
        self.version = 0    # bumped by whoever rewrites the columns (cache key of OpticalChain)

    # ---- columns ----------------------------------------------------------------------------
    _POINT = ("px", "py", "pz")

    def has(self, name):
        return name in self._names or (self.origin is not None and name in self._POINT)

    def col(self, name):
        if name not in self._names and self.origin is not None and name in self._POINT:
            return self.origin[self._POINT.index(name)].expand(self.n)  # stride-0 read-only view
        return self._storage[self._names.index(name), : self.n]

    def trace_flags(self):
        """Flags the trace entry points need for this bundle's layout (ART_TRACE_UNIFORM_POINT)."""
        return _cabi.TRACE_UNIFORM_POINT if self.origin is not None else 0

    def materialize(self):
        """A bundle with real px/py/pz columns (no-op unless this is a uniform-origin point source)."""
        if self.origin is None:
            return self
        names = self._POINT + self._names
        b = RayBundle(self.n, device=self.device, columns=names, wavelength=self.wavelength)
        for i, c in enumerate(self._POINT):
            b.col(c).fill_(float(self.origin[i]))
        for c in self._names:
            b.col(c).copy_(self.col(c))
        b.alive, b.number = self.alive, self.number
        return b

    def __getattr__(self, name):
        if name in _COLUMNS:
            if name in self.__dict__.get("_names", ()) or (self.__dict__.get("origin") is not None
                                                            and name in self._POINT):
                return self.col(name)
            return None
        raise AttributeError(name)

    def view(self):
        """ArtBundleView over the tensors (pointers stay valid while this bundle lives)."""
        v = _cabi.ArtBundleView()
        base, stride = self._storage.data_ptr(), self._storage.stride(0) * 8
        for name in _COLUMNS:  # row pointers from the base: an empty slice has no data_ptr of its own
            setattr(v, name, base + self._names.index(name) * stride if name in self._names else None)
        if self.origin is not None:  # uniform point: px, py, pz each point to one double
            v.px, v.py, v.pz = (self.origin.data_ptr() + 8 * i for i in range(3))
        v.alive = self.alive.untyped_storage().data_ptr() if self.alive is not None else None
        v.n = self.n
        return v

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, P, U, path=None, intensity=None, number=None, wavelength=None, device="cpu",
                   normalize=True):
        """Bundle from (n,3) arrays of points and directions; directions are normalised like the
        Ray.vector setter does (ART/ModuleOpticalRay.py:85-90)."""
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1, 3)
        U = np.ascontiguousarray(U, dtype=np.float64).reshape(-1, 3)
        if normalize:
            U = U / np.linalg.norm(U, axis=1)[:, None]
        n = P.shape[0]
        names = ["px", "py", "pz", "ux", "uy", "uz"]
        data = [P[:, 0], P[:, 1], P[:, 2], U[:, 0], U[:, 1], U[:, 2]]
        if path is not None:
            names.append("path")
            data.append(np.asarray(path, dtype=np.float64))
        if intensity is not None:
            names.append("intensity")
            data.append(np.asarray(intensity, dtype=np.float64))
        b = cls(n, device="cpu", columns=names, wavelength=wavelength)
        for i, d in enumerate(data):
            b._storage[i, :n] = torch.from_numpy(np.ascontiguousarray(d))
        if number is not None:
            b.number = torch.from_numpy(np.ascontiguousarray(number, dtype=np.int64))
        return b.to(device)

    @classmethod
    def from_rays(cls, rays, device="cpu"):
        """Bundle from a list of Ray objects (the reference's source_rays)."""
        n = len(rays)
        P = np.array([r.point for r in rays], dtype=np.float64).reshape(n, 3)
        U = np.array([r.vector for r in rays], dtype=np.float64).reshape(n, 3)
        path = np.array([float(np.sum(r.path)) for r in rays], dtype=np.float64)
        inten = None
        if n and all(r.intensity is not None for r in rays):
            inten = np.array([r.intensity for r in rays], dtype=np.float64)
        numbers = [r.number for r in rays]
        number = None
        if any(k is None for k in numbers):
            pass
        elif numbers != list(range(n)):
            number = np.array(numbers, dtype=np.int64)
        wl = rays[0].wavelength if n else None
        return cls.from_numpy(P, U, path=path if np.any(path != 0) else None, intensity=inten, number=number,
                              wavelength=wl, device=device)

    def to(self, device):
        device = torch.device(device)
        if device == self.device:
            return self
        b = RayBundle(self.n, device=device, columns=self._names, wavelength=self.wavelength,
                      storage=self._storage.to(device))
        b.alive = None if self.alive is None else self.alive.to(device)
        b.number = None if self.number is None else self.number.to(device)
        b.origin = None if self.origin is None else self.origin.to(device)
        return b

    def pin_memory(self):
        self._storage = self._storage.pin_memory()
        if self.origin is not None:
            self.origin = self.origin.pin_memory()
        return self

    # ---- the list facade ----------------------------------------------------------------------
    def alive_index(self):
        """int64 tensor with the indices of the surviving rays, in order."""
        if self._index is None:
            if self.alive is None:
                self._index = torch.arange(self.n, device=self.device)
            else:
                self._index = torch.nonzero(self.alive, as_tuple=False).reshape(-1)
        return self._index

    def invalidate(self):
        self._index = None
        self.version += 1

    def __len__(self):
        return int(self.alive_index().numel())

    def numbers(self):
        """Ray numbers of the surviving rays (int64 tensor)."""
        idx = self.alive_index()
        return idx if self.number is None else self.number[idx]

    def to_numpy(self):
        """Compacted copy of the surviving rays as numpy arrays: number, P, U, path, incidence, intensity."""
        idx = self.alive_index()
        out = {"number": self.numbers().cpu().numpy()}

        def get(name):
            return self.col(name)[idx].cpu().numpy() if self.has(name) else None

        out["P"] = np.stack([get("px"), get("py"), get("pz")], axis=1)
        out["U"] = np.stack([get("ux"), get("uy"), get("uz")], axis=1)
        out["path"] = get("path") if self.has("path") else np.zeros(idx.numel())
        out["incidence"] = get("incidence")
        out["intensity"] = get("intensity")
        return out

    def _ray(self, i):
        g = int(self.alive_index()[i])
        row = self._storage[:, g].cpu().numpy()
        vals = dict(zip(self._names, row))
        if self.origin is not None:
            vals.update(zip(self._POINT, self.origin.cpu().numpy()))
        num = g if self.number is None else int(self.number[g])
        inc = vals.get("incidence")
        return Ray(np.array([vals["px"], vals["py"], vals["pz"]]), np.array([vals["ux"], vals["uy"], vals["uz"]]),
                   Path=(0.0, float(vals.get("path", 0.0))), Number=num, Wavelength=self.wavelength,
                   Incidence=None if inc is None or np.isnan(inc) else float(inc),
                   Intensity=None if "intensity" not in vals else float(vals["intensity"]))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._ray(j) for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError("ray index out of range")
        return self._ray(i)

    def __iter__(self):
        d = self.to_numpy()
        for j in range(d["number"].size):
            inc = None if d["incidence"] is None or np.isnan(d["incidence"][j]) else float(d["incidence"][j])
            yield Ray(d["P"][j].copy(), d["U"][j].copy(), Path=(0.0, float(d["path"][j])), Number=int(d["number"][j]),
                      Wavelength=self.wavelength, Incidence=inc,
                      Intensity=None if d["intensity"] is None else float(d["intensity"][j]))

    # ---- persistence: bundles pickle as host columns (ModuleProcessing.save_compressed) -----------------
    def __getstate__(self):
        return {"n": self.n, "names": self._names, "wavelength": self.wavelength,
                "storage": self._storage[:, : self.n].cpu().numpy(),
                "alive": None if self.alive is None else self.alive.cpu().numpy(),
                "number": None if self.number is None else self.number.cpu().numpy(),
                "origin": None if self.origin is None else self.origin.cpu().numpy()}

    def __setstate__(self, st):
        self.__init__(st["n"], device="cpu", columns=st["names"], wavelength=st["wavelength"])
        self._storage[:, : self.n] = torch.from_numpy(st["storage"])
        self.alive = None if st["alive"] is None else torch.from_numpy(st["alive"])
        self.number = None if st["number"] is None else torch.from_numpy(st["number"])
        self.origin = None if st["origin"] is None else torch.from_numpy(st["origin"])

    def content_key(self):
        """Cheap identity of the bundle's content for OpticalChain's result cache."""
        return (id(self._storage), self.n, self.version, None if self.origin is None else tuple(self.origin.tolist()))
