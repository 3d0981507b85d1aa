"""pyftk.synthesizers (ref: python/pyftk.cpp:144-179, include/ftk/ndarray/synthetic.hh).

Host-side numpy generators with the reference's closed forms; returned arrays use the reference's
numpy convention (shape (nv, W, H, T) whose buffer is read dim-0-fastest by the trackers).
"""
import numpy as np


def _grid(DW, DH):
    j, i = np.meshgrid(np.arange(DH, dtype=np.float64), np.arange(DW, dtype=np.float64), indexing="ij")
    return i, j   # arrays of shape (DH, DW): memory order with the W index fastest


def woven_snapshot(DW, DH, t, scaling_factor=15.0):
    """synthetic_woven_2D (synthetic.hh:32-48); memory-order array (DH, DW)"""
    i, j = _grid(DW, DH)
    x = ((i / (DW - 1)) - 0.5) * scaling_factor
    y = ((j / (DH - 1)) - 0.5) * scaling_factor
    return np.cos(x * np.cos(t) - y * np.sin(t)) * np.sin(x * np.sin(t) + y * np.cos(t))


def spiral_woven(DW, DH, DT):
    """synthetic_woven_2Dt (synthetic.hh:90-109): t = k / (DT - 1) + 1e-4"""
    snaps = [woven_snapshot(DW, DH, float(k) / (DT - 1) + 1e-4) for k in range(DT)]
    return np.stack(snaps).reshape(-1).reshape(1, DW, DH, DT)


def double_gyre_snapshot(DW, DH, time, A=0.1, omega=2 * np.pi, eps=0.25):
    """synthetic_double_gyre (synthetic.hh:130-150,193-217); memory-order array (DH, DW, 2)"""
    i, j = _grid(DW, DH)
    x = (i / (DW - 1)) * 2
    y = j / (DH - 1)
    a = eps * np.sin(omega * time)
    b = 1 - 2 * eps * np.sin(omega * time)
    f = a * x * x + b * x
    dfdx = 2 * a * x + b
    u = -np.pi * A * np.sin(np.pi * f) * np.cos(np.pi * y)
    v = np.pi * A * np.cos(np.pi * f) * np.sin(np.pi * y) * dfdx
    return np.stack([u, v], axis=-1)


def double_gyre_flow(DW, DH, DT):
    if DT < 1:
        raise RuntimeError("DT must be an integer greater than 1")
    snaps = [double_gyre_snapshot(DW, DH, i * 0.1) for i in range(DT)]
    return np.stack(snaps).reshape(-1).reshape(2, DW, DH, DT)


def moving_extremum_snapshot(dims, x0, direction, t):
    """synthetic_moving_extremum (synthetic.hh:332-354); memory-order array ([D,]H,W)"""
    nd = len(dims)
    grids = np.meshgrid(*[np.arange(d, dtype=np.float64) for d in reversed(dims)], indexing="ij")
    out = np.zeros(tuple(reversed(dims)))
    for q in range(nd):
        xc = x0[q] + direction[q] * t
        e = grids[nd - 1 - q] - xc
        out = out + e * e
    return out


def moving_extremum(DW, DH, DT, x0, y0, dir_x, dir_y):
    snaps = [moving_extremum_snapshot([DW, DH], [x0, y0], [dir_x, dir_y], float(k)) for k in range(DT)]
    return np.stack(snaps).reshape(-1).reshape(1, DW, DH, DT)
