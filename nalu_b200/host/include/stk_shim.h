// stk_shim.h -- the sliver of the STK mesh vocabulary the assembly path's API is written in.
//
// The reference's Kernel / ElemDataRequests / AssembleElemSolverAlgorithm / LinearSystem signatures
// mention stk::mesh::{BulkData, MetaData, FieldBase, Field, Part, Entity} and stk::topology
// (e.g. include/kernel/MomentumAdvDiffElemKernel.h:32-37, include/ElemDataRequests.h:106-126).
// Trilinos/STK is not available here, and the hot path only ever needs, per rank: node ids, node
// owners, Hex8 connectivity and flat nodal / element fields.  This header provides exactly that
// behind the same names so that host code written against the reference's interface compiles
// unchanged.  It is NOT an STK re-implementation: no buckets, selectors, ghosting or I/O.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define STK_ThrowRequireMsg(cond, msg)                                                                     \
  do {                                                                                                     \
    if (!(cond)) {                                                                                         \
      std::ostringstream _oss;                                                                             \
      _oss << msg;                                                                                         \
      throw std::logic_error(_oss.str());                                                                  \
    }                                                                                                      \
  } while (0)
#include <sstream>

namespace stk {

struct topology {
  enum rank_t { NODE_RANK = 0, EDGE_RANK = 1, FACE_RANK = 2, ELEM_RANK = 3, ELEMENT_RANK = 3, INVALID_RANK = 256 };
  enum topology_t { INVALID_TOPOLOGY = 0, NODE, HEX_8, HEX_27, QUAD_4_2D, QUAD_4, HEXAHEDRON_8 = HEX_8 };
  topology(topology_t t = INVALID_TOPOLOGY) : t_(t) {}
  operator topology_t() const { return t_; }
  unsigned num_nodes() const
  {
    return t_ == HEX_8 ? 8 : (t_ == HEX_27 ? 27 : ((t_ == QUAD_4_2D || t_ == QUAD_4) ? 4 : (t_ == NODE ? 1 : 0)));
  }
  rank_t rank() const { return t_ == NODE ? NODE_RANK : (t_ == QUAD_4 ? FACE_RANK : ELEM_RANK); }
  std::string name() const
  {
    return t_ == HEX_8 ? "HEXAHEDRON_8" : (t_ == HEX_27 ? "HEXAHEDRON_27" : (t_ == QUAD_4 ? "QUADRILATERAL_4" : "UNSUPPORTED"));
  }
  // HEX_8 side table (the side_node_ordinals of HexSCS, src/master_element/Hex8CVFEM.C:414-422)
  template <typename It>
  void face_node_ordinals(unsigned face, It out) const
  {
    static const int t[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
    for (int i = 0; i < 4; ++i) *out++ = t[face][i];
  }
  topology_t t_;
};

namespace mesh {

typedef topology::rank_t EntityRank;
enum FieldState { StateNone = 0, StateNew = 0, StateNP1 = 0, StateOld = 1, StateN = 1, StateNM1 = 2 };

struct Entity {
  EntityRank rank = topology::INVALID_RANK;
  int64_t index = -1; // local index within its rank
  bool is_valid() const { return index >= 0; }
};

class BulkData;
class MetaData;

class Part {
public:
  Part(const std::string &name, topology t) : name_(name), topo_(t) {}
  const std::string &name() const { return name_; }
  stk::topology topology() const { return topo_; }
  // A node-set / side-set part (what the boundary-condition parts of an Exodus mesh are): the local node
  // indices it holds.  A part without an explicit list (an element block) holds every node of the mesh.
  void set_nodes(const std::vector<int32_t> &local_nodes)
  {
    nodes_ = local_nodes;
    explicit_nodes_ = true;
  }
  bool has_explicit_nodes() const { return explicit_nodes_; }
  const std::vector<int32_t> &nodes() const { return nodes_; }
  // A side-set part: its locally owned faces as (local element, face ordinal) -- what bulk.begin_elements(face) /
  // begin_element_ordinals(face) give in the reference (include/AssembleFaceElemSolverAlgorithm.h:91-103) -- numbered
  // first_face() .. first_face() + n in the rank's face numbering (the index of a FACE_RANK field).
  void set_faces(const std::vector<int32_t> &elems, const std::vector<int32_t> &ordinals, size_t first_face)
  {
    face_elem_ = elems;
    face_ord_ = ordinals;
    first_face_ = first_face;
  }
  const std::vector<int32_t> &face_elems() const { return face_elem_; }
  const std::vector<int32_t> &face_ordinals() const { return face_ord_; }
  size_t first_face() const { return first_face_; }
private:
  std::string name_;
  stk::topology topo_;
  std::vector<int32_t> nodes_, face_elem_, face_ord_;
  size_t first_face_ = 0;
  bool explicit_nodes_ = false;
};
typedef std::vector<Part *> PartVector;

// One state of a (possibly multi-state) field: flat [nEntities][scalarsPerEntity] doubles on the host
// plus a device mirror.  Host writes must be followed by modify_on_host(); the assembly algorithm
// calls sync_to_device() (a no-op when nothing changed) before it launches.
class FieldBase {
public:
  const std::string &name() const { return name_; }
  EntityRank entity_rank() const { return rank_; }
  unsigned mesh_meta_data_ordinal() const { return ordinal_; }
  unsigned number_of_states() const { return nstates_; }
  FieldState state() const { return state_; }
  unsigned max_size() const { return scalars_; }
  FieldBase &field_of_state(FieldState s) const
  {
    STK_ThrowRequireMsg((unsigned)s < nstates_, "field " << name_ << " has no state " << (int)s);
    return *states_[s];
  }
  double *host_data() { return host_.data(); }
  const double *host_data() const { return host_.data(); }
  size_t num_entities() const { return scalars_ ? host_.size() / scalars_ : 0; }
  void modify_on_host() { host_dirty_ = true; }
  void sync_to_device();           // host -> device when modified (allocates on first use)
  const double *device_data();     // syncs first
  void set_scalars_per_entity(unsigned n) { scalars_ = n; if (nent_) host_.assign(nent_ * n, 0.0); }
  void set_num_entities(size_t n) { nent_ = n; host_.assign(n * scalars_, 0.0); host_dirty_ = true; }
  void set_device_data(double *p)  // adopt caller-owned device memory (stays device resident)
  {
    release_device();
    dev_ = p;
    dev_owned_ = false;
    host_dirty_ = false;
  }
  ~FieldBase() { release_device(); }

protected:
  FieldBase() = default;
private:
  friend class MetaData;
  friend class BulkData;
  void release_device();
  std::string name_;
  EntityRank rank_ = topology::NODE_RANK;
  unsigned ordinal_ = 0, nstates_ = 1, scalars_ = 0;
  FieldState state_ = StateNone;
  FieldBase *states_[3] = {nullptr, nullptr, nullptr};
  std::vector<double> host_;
  size_t nent_ = 0;
  double *dev_ = nullptr;
  bool dev_owned_ = false, host_dirty_ = true;
};

template <typename T>
class Field : public FieldBase {
  friend class MetaData;
  Field() = default;
public:
  Field<T> &field_of_state(FieldState s) const { return static_cast<Field<T> &>(FieldBase::field_of_state(s)); }
};

class MetaData {
public:
  explicit MetaData(unsigned spatial_dim = 3) : dim_(spatial_dim) {}
  ~MetaData()
  {
    for (FieldBase *f : fields_) delete f;
    for (Part *p : parts_) delete p;
  }
  unsigned spatial_dimension() const { return dim_; }
  EntityRank side_rank() const { return topology::FACE_RANK; }
  Part *get_part(const std::string &name) const
  {
    for (Part *p : parts_)
      if (p->name() == name) return p;
    return nullptr;
  }
  Part &declare_part_with_topology(const std::string &name, stk::topology t)
  {
    parts_.push_back(new Part(name, t));
    return *parts_.back();
  }
  // declare_field<double>(rank, name, numberOfStates); the returned handle is StateNP1
  template <typename T>
  Field<T> &declare_field(EntityRank rank, const std::string &name, unsigned nstates = 1)
  {
    STK_ThrowRequireMsg(nstates >= 1 && nstates <= 3, "declare_field: 1..3 states");
    Field<T> *st[3] = {nullptr, nullptr, nullptr};
    for (unsigned s = 0; s < nstates; ++s) {
      Field<T> *f = new Field<T>();
      f->name_ = name;
      f->rank_ = rank;
      f->ordinal_ = (unsigned)fields_.size();
      f->nstates_ = nstates;
      f->state_ = (FieldState)s;
      fields_.push_back(f);
      st[s] = f;
    }
    for (unsigned s = 0; s < nstates; ++s)
      for (unsigned q = 0; q < nstates; ++q) st[s]->states_[q] = st[q];
    return *st[0];
  }
  // get_field<double>(rank, name): nullptr when absent, like the reference's MetaData
  template <typename T>
  Field<T> *get_field(EntityRank rank, const std::string &name) const
  {
    for (FieldBase *f : fields_)
      if (f->rank_ == rank && f->name_ == name && f->state_ == StateNP1) return static_cast<Field<T> *>(f);
    return nullptr;
  }
  const std::vector<FieldBase *> &get_fields() const { return fields_; }

private:
  friend class BulkData;
  unsigned dim_;
  std::vector<FieldBase *> fields_;
  std::vector<Part *> parts_;
};

// put_field_on_mesh(field, part, scalarsPerEntity, init): sizes every state of the field
template <typename T>
void put_field_on_mesh(Field<T> &field, const Part &, unsigned scalars, const T * = nullptr);
typedef Field<double> ScalarFieldTypeBase;

// Rank-local Hex8 mesh (flat arrays; all elements are locally owned)
class BulkData {
public:
  BulkData(MetaData &meta, int rank = 0, int nranks = 1) : meta_(meta), rank_(rank), nranks_(nranks) {}
  MetaData &mesh_meta_data() const { return meta_; }
  int parallel_rank() const { return rank_; }
  int parallel_size() const { return nranks_; }
  // Ioss `generated:nx x ny x nz` piece of this rank under the z-slab decomposition
  // (unit_tests/UnitTestUtils.C:30-35): ids 1+i+(nx+1)j+(nx+1)(ny+1)k, unit spacing
  void fill_generated_mesh(int nx, int ny, int nz, Part &block, const std::string &coords_name = "coordinates");
  // `|sideset:xXyYzZ` of the Ioss generated mesh: one QUAD_4 part "surface_<k>" per letter, in the order given (x / X: low /
  // high x side, ...), holding this rank's faces of that side, plus "all_surfaces" = their union; sizes the FACE_RANK
  // fields.  Call after fill_generated_mesh.
  void create_generated_sidesets(const std::string &sides);
  // nx x ny x nz HEX_27 elements on the (2nx+1)(2ny+1)(2nz+1) lattice of spacing 1/2 (one rank): ids 1 + lexicographic
  // index, connectivity in STK's Hex27 node order (what promoting the generated Hex8 box to P2 produces, up to the numbering)
  void fill_generated_hex27_mesh(int nx, int ny, int nz, Part &block, const std::string &coords_name = "coordinates");
  int nodes_per_elem() const { return npe_; }
  size_t num_faces() const { return num_faces_; }
  size_t num_nodes() const { return node_gid_.size(); }
  size_t num_elems() const { return conn_.size() / size_t(npe_); }
  const std::vector<int64_t> &node_ids() const { return node_gid_; }    // naluGlobalId, 1-based
  const std::vector<int32_t> &node_owners() const { return node_owner_; }
  const std::vector<int32_t> &connectivity() const { return conn_; }   // [nElem][8] (Hex27 meshes: [nElem][27]), local node indices
  const int32_t *begin_nodes(Entity elem) const { return &conn_[size_t(elem.index) * size_t(npe_)]; }
  int64_t identifier(Entity e) const { return e.rank == topology::NODE_RANK ? node_gid_[size_t(e.index)] : elem_gid0_ + e.index; }
  Entity get_entity(EntityRank rank, int64_t id) const;
  const int *box() const { return box_; } // local element counts when the mesh is a generated box
  Part *block() const { return block_; }

private:
  MetaData &meta_;
  int rank_, nranks_;
  std::vector<int64_t> node_gid_;
  std::vector<int32_t> node_owner_, conn_;
  int64_t elem_gid0_ = 1;
  int npe_ = 8;
  int box_[3] = {0, 0, 0};
  int k0_ = 0, nz_global_ = 0;
  size_t num_faces_ = 0;
  Part *block_ = nullptr;
};

inline double *field_data(FieldBase &f, Entity e) { return f.host_data() + size_t(e.index) * f.max_size(); }

// scalars per entity for every state of the field; the storage is sized when the mesh is filled
// (BulkData::fill_generated_mesh) or immediately if it already is
template <typename T>
void put_field_on_mesh(Field<T> &field, const Part &, unsigned scalars, const T *)
{
  for (unsigned s = 0; s < field.number_of_states(); ++s) field.FieldBase::field_of_state((FieldState)s).set_scalars_per_entity(scalars);
}

} // namespace mesh
} // namespace stk
