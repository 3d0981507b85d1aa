// Implicit space-time simplicial mesh: unit-simplex type tables.
//
// What the reference builds at start-up with recursive cube subdivision and std::set bookkeeping
// (ref: include/ftk/mesh/simplicial_regular_mesh.hh:620-831) is derived here from the closed-form
// description of the Kuhn triangulation: a k-simplex type of the d-cube is a strictly increasing
// chain of vertex subsets  {} = v0 < v1 < ... < vk  of {0..d-1} (bit j of a vertex mask = offset 1
// in dimension j; dimension d-1 is time).  Types are numbered in lexicographic order of their vertex
// lists with dimension 0 compared first, which is the order the reference's std::set produces.
#pragma once
#include <cstdint>
#include <vector>

namespace ftkb {

struct TypeOffset {
  int type;
  int off[4];
};

struct MeshTables {
  int nd = 0;                                            // mesh dimensionality: 3 (2D+t) or 4 (3D+t)
  std::vector<std::vector<std::vector<uint8_t>>> unit;   // unit[k][type] -> k+1 vertex masks, ascending
  std::vector<std::vector<int>> ordinal_types, interval_types;  // [k] -> type list
  std::vector<std::vector<uint8_t>> is_ordinal;          // [k][type]
  std::vector<std::vector<std::vector<TypeOffset>>> sides;    // [k][type] -> facets (type of k-1, offset)
  std::vector<std::vector<std::vector<TypeOffset>>> side_of;  // [k][type] -> cofaces (type of k+1, offset)

  int ntypes(int k) const { return (int)unit[k].size(); }
};

// nd_mesh = 3 or 4; tables are built once and cached
const MeshTables &mesh_tables(int nd_mesh);

// Flat tables for the device (constant memory), n-simplices of the (n+1)-D mesh only.
struct DeviceMeshTables {
  int32_t nd;                 // mesh dimensionality (n + 1)
  int32_t ntypes;             // 12 (2D+t) or 60 (3D+t)
  uint8_t vmask[60][4];       // vertex masks of each type (n+1 used)
  uint8_t ordinal[60];
  uint8_t tmask[60];          // OR of the vertex masks (which cube vertices' dimensions a type touches)
  uint8_t n_nb[60];           // number of neighbour candidates
  int8_t nb_type[60][8];      // neighbour candidates = other facets of the two cofaces
  int8_t nb_off[60][8][4];
};

void fill_device_tables(int nd_mesh, DeviceMeshTables *out);

}  // namespace ftkb
