"""pyftk.extractors (ref: python/pyftk.cpp:15-90): single-snapshot extraction (ordinal sweep only)."""
import numpy as np

from .tracker import make_tracker, critical_point_type_to_string


def extract_critical_points_2d_scalar(array, device=0):
    a = np.ascontiguousarray(array, dtype=np.float64)
    if a.ndim != 4:
        raise RuntimeError("Number of dimensions must be 4 (1, width, height, 1)")
    # ftk::ndarray(array) keeps the numpy shape with dim 0 fastest: DW = shape[0], DH = shape[1] (pyftk.cpp:20-22)
    DW, DH = a.shape[0], a.shape[1]
    tr = make_tracker([DW, DH], field="scalar", device=device)
    tr.push_scalar_field_snapshot(a.reshape(-1)[:DW * DH].reshape(DH, DW))
    tr.update_timestep()
    out = [{"x": float(p["x"][0]), "y": float(p["x"][1]), "t": float(p["t"]),
            "type": critical_point_type_to_string(2, int(p["cp_type"]), True), "scalar": float(p["scalar"])}
           for p in tr.get_critical_points()]
    tr.close()
    return out


def extract_critical_points_2d_vector(array, device=0):
    a = np.ascontiguousarray(array, dtype=np.float64)
    if a.ndim != 4:
        raise RuntimeError("Number of dimensions must be 4 (2, width, height, 1)")
    if a.shape[0] != 2:
        raise RuntimeError("The first dimension must be 2")
    DW, DH = a.shape[1], a.shape[2]
    from .tracker import SOURCE_NONE, SOURCE_DERIVED, SOURCE_GIVEN
    # the reference configures vector SOURCE_DERIVED here but pushes the vector itself (pyftk.cpp:63-72);
    # the pushed array is what the sweep reads, which GIVEN expresses
    tr = make_tracker([DW, DH], field="vector", device=device, scalar_source=SOURCE_NONE, vector_source=SOURCE_GIVEN,
                      jacobian_source=SOURCE_DERIVED, jacobian_symmetric=False)
    tr.push_vector_field_snapshot(a.reshape(-1)[:2 * DW * DH].reshape(DH, DW, 2))
    tr.update_timestep()
    out = [{"x": float(p["x"][0]), "y": float(p["x"][1]), "t": float(p["t"]),
            "type": critical_point_type_to_string(2, int(p["cp_type"]), False)} for p in tr.get_critical_points()]
    tr.close()
    return out
