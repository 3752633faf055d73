"""TEST INFRASTRUCTURE ONLY -- ctypes view of ``oracle/liboracle_collate.so`` (plain-C restatement, see
``oracle/collate_oracle.c`` for the reference file:line map)."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle_collate.so')
        if not os.path.exists(path):
            subprocess.check_call(['make', '-C', _HERE, 'liboracle_collate.so'])
        lib = ctypes.CDLL(path)
        lib.oracle_grid_subsampling.restype = ctypes.c_int64
        lib.oracle_grid_subsampling.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_radius_neighbors.restype = ctypes.c_int64
        lib.oracle_radius_neighbors.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float,
                                                ctypes.c_void_p, ctypes.c_int64]
        _LIB = lib
    return _LIB


def grid_subsampling(points, lengths, voxel_size):
    points = points.contiguous().float().cpu()
    lengths = lengths.contiguous().long().cpu()
    n, b = points.shape[0], lengths.shape[0]
    s_points = torch.zeros((n, 3), dtype=torch.float32)
    s_lengths = torch.zeros((b,), dtype=torch.int64)
    total = _lib().oracle_grid_subsampling(points.data_ptr(), n, lengths.data_ptr(), b, float(voxel_size),
                                           s_points.data_ptr(), s_lengths.data_ptr())
    return [s_points[:total].clone(), s_lengths]


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius):
    q_points = q_points.contiguous().float().cpu()
    s_points = s_points.contiguous().float().cpu()
    q_lengths = q_lengths.contiguous().long().cpu()
    s_lengths = s_lengths.contiguous().long().cpu()
    nq, ns, b = q_points.shape[0], s_points.shape[0], q_lengths.shape[0]
    args = (q_points.data_ptr(), nq, s_points.data_ptr(), ns, q_lengths.data_ptr(), s_lengths.data_ptr(), b,
            float(radius))
    width = _lib().oracle_radius_neighbors(*args, None, 0)
    out = torch.zeros((nq, width), dtype=torch.int64)
    _lib().oracle_radius_neighbors(*args, out.data_ptr(), width)
    return out
