"""The reference's own header-only arithmetic, compiled from /root/reference (oracle/_ref/).

`oracle/ref_headers.cpp` wraps `src/particles/particles_utils/ShapeFactors.H`,
`src/particles/pusher/PushPlasmaParticles.H` and `src/utils/DualNumbers.H` of the reference --
included from where they lie, with `oracle/ref_shim/` standing in for the three AMReX names they
use -- behind C entry points.  tests/test_oracle_refheaders.py holds the NumPy restatement to it.

THIS IS TEST INFRASTRUCTURE (see the header of hipace_oracle.py).  The .so is built in the build
container only (the GPU box has no /root/reference; the built file travels with the snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, '_ref')
_LIB = os.path.join(_OUT, 'libhipace_refhdr.so')
REF_SRC = '/root/reference/src'


def make_refhdr(force=False):
    """g++ on our wrapper + the reference headers in place; returns the path or None when neither
    the reference tree nor a previously built library is available."""
    have_ref = os.path.exists(os.path.join(REF_SRC, 'particles/particles_utils/ShapeFactors.H'))
    if os.path.exists(_LIB) and not (force and have_ref):
        src_m = os.path.getmtime(os.path.join(_HERE, 'ref_headers.cpp'))
        if not have_ref or os.path.getmtime(_LIB) >= src_m:
            return _LIB
    if not have_ref:
        return None
    os.makedirs(_OUT, exist_ok=True)
    # -ffp-contract=off: the reference's CPU build does not fuse either (GCC default for ISO C++)
    subprocess.run(['g++', '-std=c++17', '-O1', '-ffp-contract=off', '-shared', '-fPIC',
                    '-I', os.path.join(_HERE, 'ref_shim'), '-I', REF_SRC,
                    '-o', _LIB, os.path.join(_HERE, 'ref_headers.cpp')], check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = make_refhdr()
        if path is None:
            return None
        _lib = C.CDLL(path)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_shape(order, xmid):
    """-> (array variant [order+1, n], single non-branchless, single branchless, leftmost cell)"""
    L = lib()
    x = np.ascontiguousarray(xmid, dtype=np.float64)
    n = x.size
    s = np.zeros((3 * (order + 1), n))
    cell = np.zeros(n, dtype=np.int64)
    assert L.ref_shape(C.c_int(order), C.c_long(n), _dp(x), _dp(s), _dp(cell)) == 0
    m = order + 1
    return s[:m], s[m:2 * m], s[2 * m:], cell


def ref_dshape(dtype, order, xmid):
    L = lib()
    x = np.ascontiguousarray(xmid, dtype=np.float64)
    n = x.size
    m = order + dtype + 1
    s, ds = np.zeros((m, n)), np.zeros((m, n))
    cell = np.zeros(n, dtype=np.int64)
    assert L.ref_dshape(C.c_int(dtype), C.c_int(order), C.c_long(n), _dp(x), _dp(s), _dp(ds),
                        _dp(cell)) == 0
    return s, ds, cell


def _ptrs(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def ref_momentum_push(inputs, clight_inv, qmc):
    """inputs: 12 arrays (ux, uy, psi_inv, ExmBy, EypBx, Ez, Bx_c, By_c, Bz, A, ADx, ADy)"""
    L = lib()
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in inputs]
    n = a[0].size
    out = np.zeros((3, n))
    L.ref_momentum_push(C.c_long(n), _ptrs(a), C.c_double(clight_inv), C.c_double(qmc), _dp(out))
    return out


def ref_momentum_push_dual(inputs, eps, clight_inv, qmc):
    L = lib()
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in inputs]
    e = [np.ascontiguousarray(v, dtype=np.float64) for v in eps]
    n = a[0].size
    out = np.zeros((6, n))
    L.ref_momentum_push_dual(C.c_long(n), _ptrs(a), _ptrs(e), C.c_double(clight_inv),
                             C.c_double(qmc), _dp(out))
    return out[:3], out[3:]


def ref_gather(order, xp, yp, planes, g, comps, dx_inv, dy_inv, x_off, y_off):
    """doGatherShapeN (runtime-order overload) on planes[ncomp, ny+2g, nx+2g] -> [6, n]"""
    L = lib()
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    yp = np.ascontiguousarray(yp, dtype=np.float64)
    planes = np.ascontiguousarray(planes, dtype=np.float64)
    cm = np.ascontiguousarray(comps, dtype=np.int32)
    out = np.zeros((6, xp.size))
    rc = L.ref_gather(C.c_int(order), C.c_long(xp.size), _dp(xp), _dp(yp), _dp(planes),
                      C.c_int(planes.shape[2]), C.c_int(planes.shape[1]), C.c_int(g), _dp(cm),
                      C.c_double(dx_inv), C.c_double(dy_inv), C.c_double(x_off), C.c_double(y_off),
                      _dp(out))
    assert rc == 0
    return out


def ref_laser_gather(order, xp, yp, plane, g, dx_inv, dy_inv, x_off, y_off):
    """doLaserGatherShapeN<order>, both overloads -> (A, ADx, ADy, A of the value-only overload)"""
    L = lib()
    xp = np.ascontiguousarray(xp, dtype=np.float64)
    yp = np.ascontiguousarray(yp, dtype=np.float64)
    plane = np.ascontiguousarray(plane, dtype=np.float64)
    out = np.zeros((4, xp.size))
    rc = L.ref_laser_gather(C.c_int(order), C.c_long(xp.size), _dp(xp), _dp(yp), _dp(plane),
                            C.c_int(plane.shape[1]), C.c_int(plane.shape[0]), C.c_int(g),
                            C.c_double(dx_inv), C.c_double(dy_inv), C.c_double(x_off),
                            C.c_double(y_off), _dp(out))
    assert rc == 0
    return out


def ref_open_boundary(s, x, y, monopole, xd, yd):
    """fields/OpenBoundary.H: GetMultipoleCoeffs summed over the sources, GetFieldMultipole at the
    points (all coordinates scaled as Fields::SetBoundaryCondition scales them)"""
    L = lib()
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (s, x, y, xd, yd)]
    out = np.zeros(a[3].size)
    L.ref_open_boundary(C.c_long(a[0].size), _dp(a[0]), _dp(a[1]), _dp(a[2]), C.c_int(int(monopole)),
                        C.c_long(a[3].size), _dp(a[3]), _dp(a[4]), _dp(out))
    return out
