"""pyftk.trackers (ref: python/pyftk.cpp:92-142)."""
import numpy as np

from .tracker import track, critical_point_type_to_string


def track_critical_points_2d_scalar(array, device=0):
    """array: float64, shape (1, width, height, time) -- same contract as the reference, including its
    memory reinterpretation: the buffer is read dim-0-fastest, i.e. snapshot k is the k-th contiguous
    chunk of W*H values with the W index fastest (python/pyftk.cpp:97-99, no transpose).
    Returns [{'length': n, 'trace': [{'x','y','t','type','scalar'}, ...]}, ...]."""
    a = np.ascontiguousarray(array, dtype=np.float64)
    if a.ndim != 4:
        raise RuntimeError("Number of dimensions must be 4: (1, width, height, time)")
    DW, DH, DT = a.shape[1], a.shape[2], a.shape[3]
    flat = a.reshape(-1)
    snaps = [flat[k * DW * DH:(k + 1) * DW * DH].reshape(DH, DW) for k in range(DT)]
    tr = track(snaps, [DW, DH], field="scalar", device=device)
    result = []
    for pts, _loop in tr.get_traced_critical_points():
        trace = [{"x": float(p["x"][0]), "y": float(p["x"][1]), "t": float(p["t"]),
                  "type": critical_point_type_to_string(2, int(p["cp_type"]), True), "scalar": float(p["scalar"])} for p in pts]
        result.append({"length": len(trace), "trace": trace})
    tr.close()
    return result
