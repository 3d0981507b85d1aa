"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's streaming trajectories.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Restates, container for container, with Python dicts / sorted lists where the reference uses
std::map / std::set of elements:
  * critical_point_tracker::trace_critical_points_online  include/ftk/filters/critical_point_tracker.hh:522-641
  * the grow() call sites                                  critical_point_tracker_2d_regular.hh:288-329,
                                                           critical_point_tracker_3d_regular.hh:173-200
  * extract_connected_components                           include/ftk/algorithms/cca.hh:91-116
  * union_find (weighted, no same-set check)               include/ftk/basic/union_find.hh:15-107
  * connected_component_to_linear_components, is_loop      include/ftk/geometry/cc2curves.hh:10-122
  * feature_curve_set_t::add (ids in insertion order)      include/ftk/features/feature_curve_set.hh:458-465
An element is the tuple (x, y, z, t, type); tuple order = the reference's element order
(corner compared dimension 0 first, then type; simplicial_regular_mesh.hh:327-337).  Mesh facets and
cofaces come from the C oracle's tables (cp_oracle.c, simplicial_regular_mesh.hh:620-831).

Pinned: tests/golden/stream/*.npz hold the trajectories of the UNMODIFIED reference run with
set_enable_streaming_trajectories(true) (oracle/ref_harness.cpp --out-stream); tests/test_streaming.py
checks this restatement against every one of them.
"""
import ctypes as C

from . import cp_oracle as O


class _Mesh:
    _cache = {}

    def __init__(self, nd):
        L = O.lib()
        self.nd = nd          # spatial dimensionality n; the mesh is (n+1)-dimensional
        ndm = nd + 1

        def table(fn, k, t):
            buf = (C.c_int32 * (64 * (ndm + 1)))()
            cnt = fn(ndm, k, t, buf)
            return [(buf[i * (ndm + 1)], tuple(buf[i * (ndm + 1) + 1 + j] for j in range(ndm))) for i in range(cnt)]

        self.side_of = [table(L.cpo_mesh_side_of, nd, t) for t in range(L.cpo_mesh_ntypes(ndm, nd, 0))]
        self.sides = [table(L.cpo_mesh_sides, ndm, t) for t in range(L.cpo_mesh_ntypes(ndm, ndm, 0))]

    @classmethod
    def get(cls, nd):
        if nd not in cls._cache:
            cls._cache[nd] = cls(nd)
        return cls._cache[nd]

    def neighbors(self, e):
        """std::set of every side of every cell e is a side of (e included), ascending"""
        nd = self.nd
        c = (e[0], e[1], e[3]) if nd == 2 else (e[0], e[1], e[2], e[3])
        out = set()
        for ct, coff in self.side_of[e[4]]:
            cc = tuple(c[j] + coff[j] for j in range(nd + 1))
            for st, soff in self.sides[ct]:
                sc = tuple(cc[j] + soff[j] for j in range(nd + 1))
                out.add((sc[0], sc[1], 0, sc[2], st) if nd == 2 else (sc[0], sc[1], sc[2], sc[3], st))
        return sorted(out)


class _UnionFind:
    """basic/union_find.hh:15-107"""

    def __init__(self):
        self.eles, self.parent, self.sz = set(), {}, {}

    def add(self, i):
        if i in self.eles:
            return
        self.eles.add(i)
        self.parent[i] = i
        self.sz[i] = 1

    def find(self, i):
        # union_find.hh:57-68: parent_i is a C++ reference to the ORIGINAL element's slot
        orig = i
        while i != self.parent[orig]:
            self.parent[i] = self.parent[self.parent[orig]]
            i = self.parent[i]
            self.parent[orig] = self.parent[i]
        return i

    def unite(self, i, j):
        i, j = self.find(i), self.find(j)
        if self.sz[i] < self.sz[j]:
            self.parent[i] = j
            self.sz[j] = (self.sz[j] + self.sz[i]) & 0xFFFFFFFFFFFFFFFF   # size_t
        else:
            self.parent[j] = i
            self.sz[i] = (self.sz[i] + self.sz[j]) & 0xFFFFFFFFFFFFFFFF

    def get_sets(self):
        root2set = {}
        for e in sorted(self.eles):
            root2set.setdefault(self.find(e), []).append(e)
        return [root2set[r] for r in sorted(root2set)]


def extract_connected_components(neighbors, qualified):
    """cca.hh:91-116"""
    uf = _UnionFind()
    q = sorted(qualified)
    for e in q:
        uf.add(e)
    for e in q:
        for nb in neighbors(e):
            if nb in uf.eles:
                uf.unite(e, nb)
    return uf.get_sets()


def connected_component_to_linear_components(component, neighbors):
    """cc2curves.hh:10-108"""
    comp = set(component)
    ordinary, special = set(), set()
    for node in sorted(comp):
        valid = [nb for nb in neighbors(node) if nb != node and nb in comp]
        (special if len(valid) > 2 else ordinary).add(node)
    cc = extract_connected_components(lambda node: [nb for nb in neighbors(node) if nb in comp and nb not in special], ordinary)
    out = []
    for c in cc:
        cset = set(c)
        seed = c[0]
        visited = {seed}
        trace = [seed]
        seed_neighbors = [nb for nb in neighbors(seed) if nb != seed and nb in ordinary]
        for direction in range(2):
            if not seed_neighbors:
                break
            current = seed_neighbors[0] if direction == 0 else seed_neighbors[-1]
            while True:
                if current not in visited:
                    if direction == 0:
                        trace.append(current)
                    else:
                        trace.insert(0, current)
                    visited.add(current)
                found = False
                for nb in neighbors(current):
                    if nb != current and nb in cset and nb not in visited:
                        found, current = True, nb
                        break
                if not found:
                    break
            if len(seed_neighbors) == 1:
                break
        out.append(trace)
    return out


def is_loop(graph, neighbors):
    """cc2curves.hh:113-122"""
    return len(graph) > 1 and graph[-1] in neighbors(graph[0])


class OnlineTracer:
    """trajectories: list (index = id) of dicts {elements: [...], loop, complete}"""

    def __init__(self, nd):
        self.mesh = _Mesh.get(nd)
        self.trajectories = []

    def grow(self, discrete):
        """critical_point_tracker.hh:522-641; `discrete` = iterable of elements found since the last grow"""
        neighbors = self.mesh.neighbors
        discrete = set(discrete)
        for traj in self.trajectories:
            if traj["complete"]:
                continue
            continued = False
            current = traj["elements"][-1]
            while True:
                nxt = next((i for i in neighbors(current) if i in discrete), None)
                if nxt is None:
                    break
                current = nxt
                traj["elements"].append(current)
                discrete.discard(current)
                continued = True
            current = traj["elements"][0]
            while True:
                nxt = next((i for i in neighbors(current) if i in discrete), None)
                if nxt is None:
                    break
                current = nxt
                traj["elements"].insert(0, current)
                discrete.discard(current)
                continued = True
            if not continued:
                traj["complete"] = True
        for component in extract_connected_components(neighbors, discrete):
            for graph in connected_component_to_linear_components(component, neighbors):
                self.trajectories.append({"elements": list(graph), "loop": is_loop(graph, neighbors), "complete": False})


def trace_streaming(nd, corner, simplex_type, timestep, T):
    """The reference's call order for a T-step run (push k; advance_timestep for k >= 1; update_timestep at the end):
    the sweep of timestep j (ordinal j + interval j..j+1) is followed by grow() for j = 0 .. T-2; the last ordinal
    sweep (timestep T-1) is not.  Points are rows of the golden arrays; returns [(point indices, loop, complete)]."""
    elems = [tuple(int(v) for v in corner[i]) + (int(simplex_type[i]),) for i in range(len(corner))]
    index = {e: i for i, e in enumerate(elems)}
    tr = OnlineTracer(nd)
    for j in range(T - 1):
        tr.grow(e for e, ts in zip(elems, timestep) if int(ts) == j)
    return [([index[e] for e in t["elements"]], t["loop"], t["complete"]) for t in tr.trajectories]
