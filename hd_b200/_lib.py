"""ctypes binding of libhdb200.so (C ABI: include/hdb200.h).

The library is built in-tree by hd_b200/csrc/Makefile (see __graft_entry__.build).  Loading
fails loudly when it is missing; compute entry points fail loudly (HDB_ERR_NO_DEVICE) when no
CUDA device is present.  There is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhdb200.so")

OK, ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_INPUT, ERR_SMALL_BUFFER = 0, -1, -2, -3, -4, -5, -6, -7
STATUS_OK, STATUS_NONFINITE, STATUS_SINGULAR, STATUS_IGNORE_LOAD, STATUS_EIGEN, STATUS_PATTERN = range(6)
Q_ELIMINATED, Q_IGNORE, Q_CSR_ROWPTR, Q_CSR_COLIND, Q_SCATTER, Q_FREE_INDEX, Q_PERM, Q_L_COLPTR, Q_L_ROWIND = range(9)

# every symbol include/hdb200.h declares (tests check that the library exports each of them)
SYMBOLS = [
    "hdb_topology_create", "hdb_topology_destroy", "hdb_topology_get_info", "hdb_topology_query",
    "hdb_check_model", "hdb_check_loads", "hdb_check_modal", "hdb_assemble_node_loads",
    "hdb_static_batch", "hdb_static_batch_dev", "hdb_static_batch_geom", "hdb_static_batch_geom_dev", "hdb_assemble_batch", "hdb_element_batch",
    "hdb_modal_mode_count", "hdb_modal_batch", "hdb_modal_batch_dev",
    "hdb_buckling_max_free", "hdb_buckling_batch", "hdb_buckling_batch_dev", "hdb_uniform_loads_dev",
    "hdb_host_alloc", "hdb_host_free", "hdb_set_profiling", "hdb_last_kernel_ms", "hdb_launch_count",
    "hdb_device_count", "hdb_error_string", "hdb_last_error_detail", "hdb_version",
    "hdb_multi_create", "hdb_multi_destroy", "hdb_multi_device_count", "hdb_multi_shard", "hdb_multi_topology",
    "hdb_multi_static_batch", "hdb_multi_static_batch_dev", "hdb_multi_synchronize", "hdb_multi_gather_summary",
    "hdb_shard_range",
]


class HdbError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        L = lib()
        msg = L.hdb_error_string(code).decode()
        detail = L.hdb_last_error_detail().decode()
        super().__init__(f"{where}: {msg} ({code})" + (f" — {detail}" if detail else ""))


BEAM_DTYPE = np.dtype([("n0", np.int64), ("n1", np.int64), ("A", np.float64), ("J", np.float64), ("E", np.float64)])
LOAD_DTYPE = np.dtype([("n", np.int64), ("forces", np.float64, (3,))])
MASS_DTYPE = np.dtype([("n", np.int64), ("mass", np.float64)])
assert BEAM_DTYPE.itemsize == 40 and LOAD_DTYPE.itemsize == 32 and MASS_DTYPE.itemsize == 16


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("ordering", C.c_int32), ("reserved", C.c_int32 * 6)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nP", "nB", "dof", "nFree", "nElim", "nIgnore", "nSupported", "nnzK",
                                         "nnzKff", "nnzL", "nnzR", "nUpdates", "device")] + [("reserved", C.c_int32 * 3)]


_lib = None


def lib():
    """Load libhdb200.so; raises when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C hd_b200/csrc). hd_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.hdb_topology_create.argtypes = [i32, vp, i32, vp, vp, vp, C.POINTER(Options), C.POINTER(vp)]
        L.hdb_topology_destroy.argtypes = [vp]
        L.hdb_topology_destroy.restype = None
        L.hdb_topology_get_info.argtypes = [vp, C.POINTER(Info)]
        L.hdb_topology_query.argtypes = [vp, i32, vp, i64]
        L.hdb_topology_query.restype = i64
        L.hdb_check_model.argtypes = [i32, vp, i32, vp, i32, vp, i32, vp, C.c_char_p, i32]
        L.hdb_check_loads.argtypes = [i32, i32, vp, C.c_char_p, i32]
        L.hdb_check_modal.argtypes = [i32, i32, vp, i32, vp, C.c_char_p, i32]
        L.hdb_assemble_node_loads.argtypes = [i32, i32, vp, vp]
        L.hdb_static_batch.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]
        L.hdb_static_batch_dev.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.hdb_static_batch_geom.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp]
        L.hdb_static_batch_geom_dev.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.hdb_assemble_batch.argtypes = [vp, i32, vp, vp, vp]
        L.hdb_element_batch.argtypes = [vp, i32, vp, vp, vp, vp]
        L.hdb_modal_mode_count.argtypes = [vp, i32, vp]
        L.hdb_modal_batch.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp]
        L.hdb_modal_batch_dev.argtypes = [vp, i32, vp, i32, vp, vp, i32, vp, vp, vp, vp, vp]
        L.hdb_buckling_max_free.argtypes = [vp]
        L.hdb_buckling_batch.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp]
        L.hdb_buckling_batch_dev.argtypes = [vp, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp]
        L.hdb_uniform_loads_dev.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp]
        L.hdb_host_alloc.argtypes = [C.POINTER(vp), i64]
        L.hdb_host_free.argtypes = [vp]
        L.hdb_set_profiling.argtypes = [vp, i32]
        L.hdb_last_kernel_ms.argtypes = [vp, vp, i32]
        L.hdb_launch_count.restype = i64
        L.hdb_multi_create.argtypes = [i32, vp, i32, vp, vp, vp, C.POINTER(Options), i32, vp, C.POINTER(vp)]
        L.hdb_multi_destroy.argtypes = [vp]
        L.hdb_multi_destroy.restype = None
        L.hdb_multi_device_count.argtypes = [vp]
        L.hdb_multi_shard.argtypes = [vp, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.hdb_shard_range.argtypes = [i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.hdb_multi_topology.argtypes = [vp, i32]
        L.hdb_multi_topology.restype = vp
        L.hdb_multi_static_batch.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]
        L.hdb_multi_static_batch_dev.argtypes = [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.hdb_multi_synchronize.argtypes = [vp]
        L.hdb_multi_gather_summary.argtypes = [vp, vp, vp, vp, i32, vp]
        L.hdb_error_string.argtypes = [i32]
        L.hdb_error_string.restype = C.c_char_p
        L.hdb_last_error_detail.restype = C.c_char_p
        _lib = L
    return _lib


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ptr(x):
    """numpy array -> host pointer; torch tensor (any device) -> data_ptr; int -> raw address."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, int):
        return C.c_void_p(x)
    return C.c_void_p(x.data_ptr())


def check(code, where):
    if code < 0:
        raise HdbError(code, where)
    return code


def beams_array(beamN, sections):
    beamN = np.asarray(beamN, dtype=np.int64).reshape(-1, 2)
    sections = np.asarray(sections, dtype=np.float64).reshape(-1, 3)
    b = np.zeros(beamN.shape[0], dtype=BEAM_DTYPE)
    b["n0"], b["n1"] = beamN[:, 0], beamN[:, 1]
    b["A"], b["J"], b["E"] = sections[:, 0], sections[:, 1], sections[:, 2]
    return b


class Topology:
    """Handle of the symbolic phase (hdb_topology_create) and its batched compute calls."""

    def __init__(self, points, beamN, sections, pins, supports, ordering=0, device=-1):
        L = lib()
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.beams = beams_array(beamN, sections)
        nP, nB = self.points.shape[0], self.beams.shape[0]
        self.pins = None
        if pins is not None and len(pins):
            self.pins = np.ascontiguousarray(pins, dtype=np.uint8).reshape(nB, 6)
        self.supports = np.ascontiguousarray(supports, dtype=np.uint8).reshape(nP, 3)
        opts = Options(device=device, ordering=ordering)
        h = C.c_void_p()
        check(L.hdb_topology_create(nP, _np_ptr(self.points), nB, _np_ptr(self.beams), _np_ptr(self.pins),
                                    _np_ptr(self.supports), C.byref(opts), C.byref(h)), "hdb_topology_create")
        self._h = h
        info = Info()
        check(L.hdb_topology_get_info(h, C.byref(info)), "hdb_topology_get_info")
        self.info = info
        self.nP, self.nB, self.dof = info.nP, info.nB, info.dof

    def close(self):
        if getattr(self, "_h", None):
            lib().hdb_topology_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, which):
        L = lib()
        n = check(L.hdb_topology_query(self._h, which, None, 0), "hdb_topology_query")
        out = np.zeros(n, dtype=np.int32)
        check(L.hdb_topology_query(self._h, which, _np_ptr(out), out.nbytes), "hdb_topology_query")
        return out

    # ---- LinearStatic ----
    def static_batch(self, sections, loads, want=("disp", "forces", "react")):
        """Host numpy in, host numpy out (goes through the chunked H2D/D2H pipeline)."""
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        loads = np.ascontiguousarray(loads, dtype=np.float64).reshape(nM, -1, self.dof)
        nLC = loads.shape[1]
        disp = np.empty((nM, nLC, self.nP, 3)) if "disp" in want else None
        forces = np.empty((nM, nLC, self.nB, 6)) if "forces" in want else None
        react = np.empty((nM, nLC, self.nP, 3)) if "react" in want else None
        status = np.zeros(nM, dtype=np.int32)
        fail = np.zeros(nM, dtype=np.int32)
        check(lib().hdb_static_batch(self._h, nM, _np_ptr(sections), nLC, _np_ptr(loads), _np_ptr(disp),
                                     _np_ptr(forces), _np_ptr(react), _np_ptr(status), _np_ptr(fail)),
              "hdb_static_batch")
        return status, fail, disp, forces, react

    def static_batch_geom(self, points, sections, loads, want=("disp", "forces", "react")):
        """static_batch with per-model coordinates points[nModels][nP][2] (hdb_static_batch_geom)."""
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        points = np.ascontiguousarray(points, dtype=np.float64).reshape(nM, self.nP, 2)
        loads = np.ascontiguousarray(loads, dtype=np.float64).reshape(nM, -1, self.dof)
        nLC = loads.shape[1]
        disp = np.empty((nM, nLC, self.nP, 3)) if "disp" in want else None
        forces = np.empty((nM, nLC, self.nB, 6)) if "forces" in want else None
        react = np.empty((nM, nLC, self.nP, 3)) if "react" in want else None
        status = np.zeros(nM, dtype=np.int32)
        fail = np.zeros(nM, dtype=np.int32)
        check(lib().hdb_static_batch_geom(self._h, nM, _np_ptr(points), _np_ptr(sections), nLC, _np_ptr(loads), _np_ptr(disp),
                                          _np_ptr(forces), _np_ptr(react), _np_ptr(status), _np_ptr(fail)),
              "hdb_static_batch_geom")
        return status, fail, disp, forces, react

    def static_batch_geom_dev(self, nModels, points, sections, nLC, loads, disp, forces, react, status, fail, stream=0):
        """Device tensors (torch) or raw device addresses; asynchronous on `stream`."""
        check(lib().hdb_static_batch_geom_dev(self._h, nModels, _ptr(points), _ptr(sections), nLC, _ptr(loads), _ptr(disp),
                                              _ptr(forces), _ptr(react), _ptr(status), _ptr(fail), C.c_void_p(stream)),
              "hdb_static_batch_geom_dev")

    def static_batch_raw(self, nModels, sections, nLC, loads, disp, forces, react, status, fail):
        """Any mix of pinned host buffers / device tensors accepted by hdb_static_batch."""
        check(lib().hdb_static_batch(self._h, nModels, _ptr(sections), nLC, _ptr(loads), _ptr(disp), _ptr(forces),
                                     _ptr(react), _ptr(status), _ptr(fail)), "hdb_static_batch")

    def static_batch_dev(self, nModels, sections, nLC, loads, disp, forces, react, status, fail, stream=0):
        """Device tensors (torch) or raw device addresses; asynchronous on `stream`."""
        check(lib().hdb_static_batch_dev(self._h, nModels, _ptr(sections), nLC, _ptr(loads), _ptr(disp), _ptr(forces),
                                         _ptr(react), _ptr(status), _ptr(fail), C.c_void_p(stream)),
              "hdb_static_batch_dev")

    def assemble_batch(self, sections):
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        kval = np.empty((nM, self.info.nnzK))
        status = np.zeros(nM, dtype=np.int32)
        check(lib().hdb_assemble_batch(self._h, nM, _np_ptr(sections), _np_ptr(kval), _np_ptr(status)),
              "hdb_assemble_batch")
        return status, kval

    def element_batch(self, sections):
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        k = np.empty((nM, self.nB, 6, 6))
        t = np.empty((nM, self.nB, 6, 6))
        kg = np.empty((nM, self.nB, 6, 6))
        check(lib().hdb_element_batch(self._h, nM, _np_ptr(sections), _np_ptr(k), _np_ptr(t), _np_ptr(kg)),
              "hdb_element_batch")
        return k, t, kg

    # ---- Modal ----
    def modal_mode_count(self, massN):
        massN = np.ascontiguousarray(massN, dtype=np.int64)
        return check(lib().hdb_modal_mode_count(self._h, len(massN), _np_ptr(massN)), "hdb_modal_mode_count")

    def modal_batch(self, sections, massN, mass, want_shapes=True):
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        massN = np.ascontiguousarray(massN, dtype=np.int64)
        mass = np.ascontiguousarray(mass, dtype=np.float64).reshape(nM, len(massN))
        mo = max(self.modal_mode_count(massN), 1)
        nModes = np.zeros(nM, dtype=np.int32)
        hz = np.zeros((nM, mo))
        shapes = np.zeros((nM, mo, self.nP, 3)) if want_shapes else None
        status = np.zeros(nM, dtype=np.int32)
        check(lib().hdb_modal_batch(self._h, nM, _np_ptr(sections), len(massN), _np_ptr(massN), _np_ptr(mass), mo,
                                    _np_ptr(nModes), _np_ptr(hz), _np_ptr(shapes), _np_ptr(status)), "hdb_modal_batch")
        return status, nModes, hz, shapes

    def modal_batch_dev(self, nModels, sections, massN, mass, maxModes, nModes, hz, shapes, status, stream=0):
        massN = np.ascontiguousarray(massN, dtype=np.int64)
        check(lib().hdb_modal_batch_dev(self._h, nModels, _ptr(sections), len(massN), _np_ptr(massN), _ptr(mass),
                                        maxModes, _ptr(nModes), _ptr(hz), _ptr(shapes), _ptr(status),
                                        C.c_void_p(stream)), "hdb_modal_batch_dev")

    # ---- mod.LoadUniform + assemblyNodeLoad on the device ----
    def uniform_loads_dev(self, nModels, nLC, ubeam, uproj, q, nodes, f, loads, stream=0):
        """q, f, loads: device tensors; ubeam, uproj, nodes: host arrays."""
        ubeam = np.ascontiguousarray(ubeam, dtype=np.int32)
        uproj = np.ascontiguousarray(uproj, dtype=np.uint8)
        nodes = np.ascontiguousarray(nodes, dtype=np.int64)
        check(lib().hdb_uniform_loads_dev(self._h, nModels, nLC, len(ubeam), _np_ptr(ubeam), _np_ptr(uproj), _ptr(q),
                                          len(nodes), _np_ptr(nodes), _ptr(f), _ptr(loads), C.c_void_p(stream)),
              "hdb_uniform_loads_dev")

    # ---- linear buckling (model.go:491-601) ----
    def buckling_max_free(self) -> int:
        return int(lib().hdb_buckling_max_free(self._h))

    def buckling_batch(self, sections, forces, amount, want_shapes=True):
        """forces[nModels][nLC][nB][6] (from static_batch) -> status[nM], nOut[nM][nLC], factors[nM][nLC][amount],
        shapes[nM][nLC][amount][nP][3]."""
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        forces = np.ascontiguousarray(forces, dtype=np.float64).reshape(nM, -1, self.nB, 6)
        nLC = forces.shape[1]
        mo = max(int(amount), 1)
        nOut = np.zeros((nM, nLC), dtype=np.int32)
        factors = np.zeros((nM, nLC, mo))
        shapes = np.zeros((nM, nLC, mo, self.nP, 3)) if want_shapes else None
        status = np.zeros(nM, dtype=np.int32)
        check(lib().hdb_buckling_batch(self._h, nM, _np_ptr(sections), nLC, _np_ptr(forces), int(amount), mo,
                                       _np_ptr(nOut), _np_ptr(factors), _np_ptr(shapes), _np_ptr(status)),
              "hdb_buckling_batch")
        return status, nOut, factors, shapes

    # ---- measurement ----
    def set_profiling(self, on=True):
        check(lib().hdb_set_profiling(self._h, int(on)), "hdb_set_profiling")

    def last_kernel_ms(self):
        ms = np.zeros(8, dtype=np.float32)
        n = check(lib().hdb_last_kernel_ms(self._h, _np_ptr(ms), 8), "hdb_last_kernel_ms")
        return ms[:n].tolist()


def shard_range(n_models: int, n_shards: int, shard: int):
    """(first, count) of shard `shard` of `n_shards`: the library's contiguous-range rule (hdb_shard_range); needs no GPU."""
    first, count = C.c_int32(0), C.c_int32(0)
    check(lib().hdb_shard_range(n_models, n_shards, shard, C.byref(first), C.byref(count)), "hdb_shard_range")
    return first.value, count.value


def launch_count() -> int:
    return int(lib().hdb_launch_count())


def device_count() -> int:
    return int(lib().hdb_device_count())


def pinned_empty(shape, dtype=np.float64):
    """numpy array over cudaHostAlloc memory (hdb_host_alloc). Keep the returned owner alive."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    check(lib().hdb_host_alloc(C.byref(p), n), "hdb_host_alloc")
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib().hdb_host_free(self.ptr)
            except Exception:
                pass

    return arr, _Owner(p)


class MultiTopology:
    """One process, several GPUs of one box (hdb_multi_*): contiguous model ranges per device, no data-path collective."""

    def __init__(self, points, beamN, sections, pins, supports, devices, ordering=0):
        L = lib()
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        self.beams = beams_array(beamN, sections)
        self.nP, self.nB = self.points.shape[0], self.beams.shape[0]
        self.dof = 3 * self.nP
        self.pins = None
        if pins is not None and len(pins):
            self.pins = np.ascontiguousarray(pins, dtype=np.uint8).reshape(self.nB, 6)
        self.supports = np.ascontiguousarray(supports, dtype=np.uint8).reshape(self.nP, 3)
        self.devices = np.ascontiguousarray(devices, dtype=np.int32)
        opts = Options(device=-1, ordering=ordering)
        h = C.c_void_p()
        check(L.hdb_multi_create(self.nP, _np_ptr(self.points), self.nB, _np_ptr(self.beams), _np_ptr(self.pins),
                                 _np_ptr(self.supports), C.byref(opts), len(self.devices), _np_ptr(self.devices), C.byref(h)),
              "hdb_multi_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().hdb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, n_models, d):
        a, b = C.c_int32(), C.c_int32()
        check(lib().hdb_multi_shard(self._h, n_models, d, C.byref(a), C.byref(b)), "hdb_multi_shard")
        return a.value, b.value

    def static_batch(self, sections, loads, want=("disp", "forces", "react")):
        """Host numpy in, host numpy out; every device works on its contiguous range of models."""
        sections = np.ascontiguousarray(sections, dtype=np.float64).reshape(-1, self.nB, 3)
        nM = sections.shape[0]
        loads = np.ascontiguousarray(loads, dtype=np.float64).reshape(nM, -1, self.dof)
        nLC = loads.shape[1]
        disp = np.empty((nM, nLC, self.nP, 3)) if "disp" in want else None
        forces = np.empty((nM, nLC, self.nB, 6)) if "forces" in want else None
        react = np.empty((nM, nLC, self.nP, 3)) if "react" in want else None
        status = np.zeros(nM, dtype=np.int32)
        fail = np.zeros(nM, dtype=np.int32)
        check(lib().hdb_multi_static_batch(self._h, nM, _np_ptr(sections), nLC, _np_ptr(loads), _np_ptr(disp), _np_ptr(forces),
                                           _np_ptr(react), _np_ptr(status), _np_ptr(fail)), "hdb_multi_static_batch")
        return status, fail, disp, forces, react

    @staticmethod
    def _ptr_array(items):
        a = (C.c_void_p * len(items))()
        for i, x in enumerate(items):
            p = _ptr(x)
            a[i] = p.value if p is not None else None
        return a

    def static_batch_dev(self, counts, sections, nLC, loads, disp, forces, react, status, fail, streams=None):
        """Resident shards: lists (one entry per device) of device tensors / addresses; asynchronous."""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        arr = lambda xs: None if xs is None else self._ptr_array(xs)
        keep = [arr(x) for x in (sections, loads, disp, forces, react, status, fail)]
        st = None if streams is None else (C.c_void_p * len(streams))(*streams)
        check(lib().hdb_multi_static_batch_dev(self._h, _np_ptr(counts), keep[0], nLC, keep[1], keep[2], keep[3], keep[4], keep[5],
                                               keep[6], st), "hdb_multi_static_batch_dev")

    def synchronize(self):
        check(lib().hdb_multi_synchronize(self._h), "hdb_multi_synchronize")

    def gather_summary(self, counts, status, disp, nLC):
        """NCCL all-gather of the per-shard reduced results: array [nDev][4] = failed, max |d|, models, 0."""
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        out = np.zeros((len(self.devices), 4))
        check(lib().hdb_multi_gather_summary(self._h, _np_ptr(counts), self._ptr_array(status),
                                             None if disp is None else self._ptr_array(disp), nLC, _np_ptr(out)),
              "hdb_multi_gather_summary")
        return out
