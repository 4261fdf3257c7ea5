// host.cpp -- host side of the drop-in boundary: the reference's registration API lowered onto the C ABI.
#include "../include/AssembleElemSolverAlgorithm.h"
#include "../include/AssembleFaceElemSolverAlgorithm.h"
#include "../include/kernel/ContinuityAdvElemKernel.h"
#include "../include/kernel/MassElemKernels.h"
#include "../include/kernel/MomentumAdvDiffElemKernel.h"
#include "../include/kernel/MomentumBuoyancySrcElemKernel.h"
#include "../include/kernel/SSTSrcElemKernels.h"
#include "../include/kernel/ScalarAdvDiffElemKernel.h"
#include "../include/kernel/UpwAdvDiffElemKernels.h"
#include "../include/nso/MomentumNSOElemKernel.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace {
void check(int rc, const char *what)
{
  if (rc != NALU_B200_OK)
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(rc) + "): " + nalu_b200_last_error());
}
} // namespace

// ================================================================================================
// stk shim
// ================================================================================================
namespace stk {
namespace mesh {

void FieldBase::release_device()
{
  if (dev_ && dev_owned_) nalu_b200_free(dev_);
  dev_ = nullptr;
  dev_owned_ = false;
}

void FieldBase::sync_to_device()
{
  if (!dev_) {
    void *p = nullptr;
    check(nalu_b200_malloc(&p, host_.size() * sizeof(double)), "field device allocation");
    dev_ = static_cast<double *>(p);
    dev_owned_ = true;
    host_dirty_ = true;
  }
  if (host_dirty_ && dev_owned_) {
    check(nalu_b200_memcpy_h2d(dev_, host_.data(), host_.size() * sizeof(double), nullptr), "field host->device copy");
    check(nalu_b200_stream_sync(nullptr), "stream sync");
  }
  host_dirty_ = false;
}

const double *FieldBase::device_data()
{
  sync_to_device();
  return dev_;
}

void BulkData::fill_generated_mesh(int nx, int ny, int nzg, Part &block, const std::string &coords_name)
{
  // z-slab decomposition: extra layers to the low ranks; a shared plane is owned by the lower rank
  const int base = nzg / nranks_, rem = nzg % nranks_;
  int k0 = 0;
  for (int r = 0; r < rank_; ++r) k0 += base + (r < rem ? 1 : 0);
  const int lz = base + (rank_ < rem ? 1 : 0);
  STK_ThrowRequireMsg(lz > 0, "generated mesh: more ranks than element layers");
  const int NX = nx + 1, NY = ny + 1, NZ = lz + 1;
  const int64_t sxy = int64_t(NX) * NY;
  node_gid_.clear(); node_owner_.clear(); conn_.clear();
  for (int z = 0; z < NZ; ++z)
    for (int y = 0; y < NY; ++y)
      for (int x = 0; x < NX; ++x) {
        node_gid_.push_back(1 + x + int64_t(NX) * y + sxy * (z + k0));
        node_owner_.push_back((rank_ > 0 && z == 0) ? rank_ - 1 : rank_);
      }
  for (int k = 0; k < lz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int32_t n0 = i + NX * (j + NY * k), s = int32_t(sxy);
        const int32_t c[8] = {n0, n0 + 1, n0 + 1 + NX, n0 + NX, n0 + s, n0 + 1 + s, n0 + 1 + NX + s, n0 + NX + s};
        conn_.insert(conn_.end(), c, c + 8);
      }
  elem_gid0_ = 1 + int64_t(nx) * ny * k0;
  box_[0] = nx; box_[1] = ny; box_[2] = lz;
  k0_ = k0;
  nz_global_ = nzg;
  block_ = &block;
  for (FieldBase *f : meta_.fields_)
    f->set_num_entities(f->entity_rank() == topology::NODE_RANK ? num_nodes()
                                                                 : (f->entity_rank() == topology::FACE_RANK ? num_faces_ : num_elems()));
  if (Field<double> *c = meta_.get_field<double>(topology::NODE_RANK, coords_name)) {
    STK_ThrowRequireMsg(c->max_size() == 3, "coordinates field must have 3 scalars per node");
    double *x = c->host_data();
    size_t n = 0;
    for (int z = 0; z < NZ; ++z)
      for (int y = 0; y < NY; ++y)
        for (int xx = 0; xx < NX; ++xx, ++n) {
          x[3 * n + 0] = xx;
          x[3 * n + 1] = y;
          x[3 * n + 2] = z + k0;
        }
    c->modify_on_host();
  }
}

void BulkData::fill_generated_hex27_mesh(int nx, int ny, int nz, Part &block, const std::string &coords_name)
{
  STK_ThrowRequireMsg(nranks_ == 1, "fill_generated_hex27_mesh: one rank (partitioned Hex27 meshes come through the C ABI)");
  // tensor-grid node (i, j, k) of an element -> STK Hex27 ordinal (HexahedralP2Element::stkNodeMap_, Hex27CVFEM.C:36-47)
  static const int stk27[27] = {0, 8, 1, 11, 21, 9, 3, 10, 2, 12, 25, 13, 23, 20, 24, 15, 26, 14, 4, 16, 5, 19, 22, 17, 7, 18, 6};
  const int NX = 2 * nx + 1, NY = 2 * ny + 1, NZ = 2 * nz + 1;
  node_gid_.clear(); node_owner_.clear(); conn_.clear();
  for (int64_t n = 0; n < int64_t(NX) * NY * NZ; ++n) {
    node_gid_.push_back(n + 1);
    node_owner_.push_back(0);
  }
  npe_ = 27;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        int32_t c[27];
        for (int cz = 0; cz < 3; ++cz)
          for (int b = 0; b < 3; ++b)
            for (int a = 0; a < 3; ++a) c[stk27[a + 3 * (b + 3 * cz)]] = (2 * i + a) + NX * ((2 * j + b) + NY * (2 * k + cz));
        conn_.insert(conn_.end(), c, c + 27);
      }
  elem_gid0_ = 1;
  box_[0] = box_[1] = box_[2] = 0;
  k0_ = 0;
  nz_global_ = nz;
  block_ = &block;
  for (FieldBase *f : meta_.fields_)
    f->set_num_entities(f->entity_rank() == topology::NODE_RANK ? num_nodes()
                                                                 : (f->entity_rank() == topology::FACE_RANK ? num_faces_ : num_elems()));
  if (Field<double> *c = meta_.get_field<double>(topology::NODE_RANK, coords_name)) {
    STK_ThrowRequireMsg(c->max_size() == 3, "coordinates field must have 3 scalars per node");
    double *x = c->host_data();
    size_t n = 0;
    for (int z = 0; z < NZ; ++z)
      for (int y = 0; y < NY; ++y)
        for (int xx = 0; xx < NX; ++xx, ++n) {
          x[3 * n + 0] = 0.5 * xx;
          x[3 * n + 1] = 0.5 * y;
          x[3 * n + 2] = 0.5 * z;
        }
    c->modify_on_host();
  }
}

void BulkData::create_generated_sidesets(const std::string &sides)
{
  STK_ThrowRequireMsg(box_[0] > 0, "create_generated_sidesets: fill_generated_mesh first");
  const int nx = box_[0], ny = box_[1], lz = box_[2];
  std::vector<int32_t> all_e, all_o;
  int k = 0;
  for (char c : sides) {
    ++k;
    std::vector<int32_t> fe, fo;
    auto add = [&](int i, int j, int kk, int ord) {
      fe.push_back(i + nx * (j + ny * kk));
      fo.push_back(ord);
    };
    // element side facing -x: 3, +x: 1, -y: 0, +y: 2, -z: 4, +z: 5 (Hex8CVFEM.C:414-422)
    for (int kk = 0; kk < lz; ++kk)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          if (c == 'x' && i == 0) add(i, j, kk, 3);
          if (c == 'X' && i == nx - 1) add(i, j, kk, 1);
          if (c == 'y' && j == 0) add(i, j, kk, 0);
          if (c == 'Y' && j == ny - 1) add(i, j, kk, 2);
          if (c == 'z' && kk == 0 && k0_ == 0) add(i, j, kk, 4);
          if (c == 'Z' && kk == lz - 1 && k0_ + lz == nz_global_) add(i, j, kk, 5);
        }
    Part &p = meta_.declare_part_with_topology("surface_" + std::to_string(k), stk::topology::QUAD_4);
    p.set_faces(fe, fo, all_e.size());
    all_e.insert(all_e.end(), fe.begin(), fe.end());
    all_o.insert(all_o.end(), fo.begin(), fo.end());
  }
  Part &all = meta_.declare_part_with_topology("all_surfaces", stk::topology::QUAD_4);
  all.set_faces(all_e, all_o, 0);
  num_faces_ = all_e.size();
  for (FieldBase *f : meta_.fields_)
    if (f->entity_rank() == topology::FACE_RANK) f->set_num_entities(num_faces_);
}

Entity BulkData::get_entity(EntityRank rank, int64_t id) const
{
  Entity e;
  e.rank = rank;
  if (rank == topology::NODE_RANK) {
    auto it = std::lower_bound(node_gid_.begin(), node_gid_.end(), id);
    e.index = (it != node_gid_.end() && *it == id) ? int64_t(it - node_gid_.begin()) : -1;
  } else {
    e.index = id - elem_gid0_;
    if (e.index < 0 || e.index >= int64_t(num_elems())) e.index = -1;
  }
  return e;
}

} // namespace mesh
} // namespace stk

namespace sierra {
namespace nalu {

// ================================================================================================
// ElemDataRequests (src/ElemDataRequests.C:13-102)
// ================================================================================================
void ElemDataRequests::add_gathered_nodal_field(const stk::mesh::FieldBase &field, unsigned scalarsPerNode)
{
  STK_ThrowRequireMsg(field.entity_rank() == stk::topology::NODE_RANK,
                      "ElemDataRequests ERROR, add_gathered_nodal_field called with field "
                          << field.name() << " which has entity-rank==" << field.entity_rank()
                          << " but is expected to have rank==NODE_RANK");
  FieldInfo fieldInfo(&field, scalarsPerNode);
  auto iter = fields.find(fieldInfo);
  if (iter == fields.end())
    fields.insert(fieldInfo);
  else
    STK_ThrowRequireMsg(iter->scalarsDim1 == scalarsPerNode,
                        "ElemDataRequests ERROR, gathered-nodal-field " << field.name() << " requested with scalarsPerNode=="
                                                                       << scalarsPerNode
                                                                       << ", but previously requested with scalarsPerNode=="
                                                                       << iter->scalarsDim1);
}
void ElemDataRequests::add_gathered_nodal_field(const stk::mesh::FieldBase &field, unsigned d1, unsigned d2)
{
  STK_ThrowRequireMsg(field.entity_rank() == stk::topology::NODE_RANK,
                      "ElemDataRequests ERROR, add_gathered_nodal_field called with field "
                          << field.name() << " which has entity-rank==" << field.entity_rank()
                          << " but is expected to have rank==NODE_RANK");
  FieldInfo fieldInfo(&field, d1, d2);
  auto iter = fields.find(fieldInfo);
  if (iter == fields.end())
    fields.insert(fieldInfo);
  else
    STK_ThrowRequireMsg(iter->scalarsDim1 == d1 && iter->scalarsDim2 == d2,
                        "ElemDataRequests ERROR, gathered-nodal-field " << field.name() << " requested with tensorDim1==" << d1
                                                                       << ",tensorDim2==" << d2 << ", but previously requested with tensorDim1=="
                                                                       << iter->scalarsDim1 << ",tensorDim2==" << iter->scalarsDim2);
}
void ElemDataRequests::add_ip_field(const stk::mesh::FieldBase &field, unsigned scalarsPerElement)
{
  FieldInfo fieldInfo(&field, scalarsPerElement);
  auto iter = fields.find(fieldInfo);
  if (iter == fields.end())
    fields.insert(fieldInfo);
  else
    STK_ThrowRequireMsg(iter->scalarsDim1 == scalarsPerElement,
                        "ElemDataRequests ERROR, ip-field " << field.name() << " requested with scalarsPerElement==" << scalarsPerElement
                                                           << ", but previously requested with scalarsPerElement==" << iter->scalarsDim1);
}
void ElemDataRequests::add_face_field(const stk::mesh::FieldBase &field, unsigned scalarsPerFace)
{
  STK_ThrowRequireMsg(field.entity_rank() == stk::topology::FACE_RANK || field.entity_rank() == stk::topology::EDGE_RANK,
                      "ElemDataRequests ERROR, add_face_field called with field " << field.name() << " which has entity-rank=="
                                                                                 << field.entity_rank());
  add_ip_field(field, scalarsPerFace);
}
void ElemDataRequests::add_face_field(const stk::mesh::FieldBase &field, unsigned d1, unsigned d2)
{
  STK_ThrowRequireMsg(field.entity_rank() == stk::topology::FACE_RANK || field.entity_rank() == stk::topology::EDGE_RANK,
                      "ElemDataRequests ERROR, add_face_field called with field " << field.name() << " which has entity-rank=="
                                                                                 << field.entity_rank());
  add_ip_field(field, d1, d2);
}
void ElemDataRequests::add_element_field(const stk::mesh::FieldBase &field, unsigned scalarsPerElement)
{
  STK_ThrowRequireMsg(field.entity_rank() == stk::topology::ELEM_RANK,
                      "ElemDataRequests ERROR, add_element_field called with field "
                          << field.name() << " which has entity-rank==" << field.entity_rank()
                          << " but is expected to have rank==ELEM_RANK");
  add_ip_field(field, scalarsPerElement);
}
void ElemDataRequests::add_ip_field(const stk::mesh::FieldBase &field, unsigned d1, unsigned d2)
{
  FieldInfo fieldInfo(&field, d1, d2);
  auto iter = fields.find(fieldInfo);
  if (iter == fields.end())
    fields.insert(fieldInfo);
  else
    STK_ThrowRequireMsg(iter->scalarsDim1 == d1 && iter->scalarsDim2 == d2,
                        "ElemDataRequests ERROR, ip-field " << field.name() << " requested with different tensor dimensions");
}
void ElemDataRequests::add_element_field(const stk::mesh::FieldBase &field, unsigned d1, unsigned d2)
{
  STK_ThrowRequireMsg(field.entity_rank() == stk::topology::ELEM_RANK,
                      "ElemDataRequests ERROR, add_element_field called with field " << field.name() << " of the wrong rank");
  add_ip_field(field, d1, d2);
}
void ElemDataRequests::add_coordinates_field(const stk::mesh::FieldBase &field, unsigned scalarsPerNode, COORDS_TYPES cType)
{
  coordsFields_[cType] = &field;
  add_gathered_nodal_field(field, scalarsPerNode);
}

// ================================================================================================
// Master elements: constant tables (Hex8CVFEM.C:75-108,255-345; MasterElementWork.F:2344-2402)
// ================================================================================================
namespace {
double sgx(int n) { return ((n ^ (n >> 1)) & 1) ? 1.0 : -1.0; }
double sgy(int n) { return ((n >> 1) & 1) ? 1.0 : -1.0; }
double sgz(int n) { return ((n >> 2) & 1) ? 1.0 : -1.0; }
void hex_shape(const double *loc, int npts, double *shp)
{
  for (int p = 0; p < npts; ++p)
    for (int n = 0; n < 8; ++n)
      shp[p * 8 + n] = (0.5 + sgx(n) * loc[3 * p]) * (0.5 + sgy(n) * loc[3 * p + 1]) * (0.5 + sgz(n) * loc[3 * p + 2]);
}
const int kLrscv[24] = {0, 1, 1, 2, 2, 3, 0, 3, 4, 5, 5, 6, 6, 7, 4, 7, 0, 4, 1, 5, 2, 6, 3, 7};
void scs_loc(double q, double *loc)
{
  for (int ip = 0; ip < 12; ++ip) {
    const int L = kLrscv[2 * ip], R = kLrscv[2 * ip + 1];
    // the ip sits on the facet through the midpoint of edge (L,R): average of the two nodes' signs, scaled
    loc[3 * ip + 0] = q * (sgx(L) + sgx(R)) * 0.5;
    loc[3 * ip + 1] = q * (sgy(L) + sgy(R)) * 0.5;
    loc[3 * ip + 2] = q * (sgz(L) + sgz(R)) * 0.5;
  }
}
void scv_loc(double q, double *loc)
{
  for (int ip = 0; ip < 8; ++ip) {
    loc[3 * ip + 0] = q * sgx(ip);
    loc[3 * ip + 1] = q * sgy(ip);
    loc[3 * ip + 2] = q * sgz(ip);
  }
}
} // namespace

HexSCS::HexSCS()
{
  nodesPerElement_ = 8;
  numIntPoints_ = 12;
  std::memcpy(lrscv_, kLrscv, sizeof(kLrscv));
  static const int faceNodes[24] = {0, 1, 5, 4, 1, 2, 6, 5, 2, 3, 7, 6, 0, 4, 7, 3, 0, 3, 2, 1, 4, 5, 6, 7};
  std::memcpy(faceNodes_, faceNodes, sizeof(faceNodes));
  // the node across the element from face node k: the edge (opposing, nearest) is the one that leaves the face
  for (int f = 0; f < 6; ++f)
    for (int k = 0; k < 4; ++k) {
      const int n = faceNodes[4 * f + k];
      int opp = -1;
      for (int e = 0; e < 12 && opp < 0; ++e) {
        const int a = kLrscv[2 * e], b = kLrscv[2 * e + 1];
        if (a != n && b != n) continue;
        const int o = a == n ? b : a;
        bool on_face = false;
        for (int q = 0; q < 4; ++q) on_face = on_face || faceNodes[4 * f + q] == o;
        if (!on_face) opp = o;
      }
      oppNode_[4 * f + k] = opp;
    }
}
Quad3DSCS::Quad3DSCS()
{
  nodesPerElement_ = 4;
  numIntPoints_ = 4;
  for (int i = 0; i < 4; ++i) ipNodeMap_[i] = i;
}
namespace {
void quad_shape(double q, double *shp)
{
  static const double sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
  for (int p = 0; p < 4; ++p)
    for (int n = 0; n < 4; ++n) shp[4 * p + n] = (0.5 + sx[n] * sx[p] * q) * (0.5 + sy[n] * sy[p] * q);
}
} // namespace
void Quad3DSCS::shape_fcn(double *shpfc) { quad_shape(0.25, shpfc); }
void Quad3DSCS::shifted_shape_fcn(double *shpfc) { quad_shape(0.5, shpfc); }
void HexSCS::shape_fcn(double *shpfc)
{
  double loc[36];
  scs_loc(0.25, loc);
  hex_shape(loc, 12, shpfc);
}
void HexSCS::shifted_shape_fcn(double *shpfc)
{
  double loc[36];
  scs_loc(0.5, loc);
  hex_shape(loc, 12, shpfc);
}
HexSCV::HexSCV()
{
  nodesPerElement_ = 8;
  numIntPoints_ = 8;
  for (int i = 0; i < 8; ++i) ipNodeMap_[i] = i;
}
void HexSCV::shape_fcn(double *shpfc)
{
  double loc[24];
  scv_loc(0.25, loc);
  hex_shape(loc, 8, shpfc);
}
void HexSCV::shifted_shape_fcn(double *shpfc)
{
  double loc[24];
  scv_loc(0.5, loc);
  hex_shape(loc, 8, shpfc);
}
// ---- Hex27 (src/master_element/Hex27CVFEM.C) ---------------------------------------------------------------------------
HexahedralP2Element::HexahedralP2Element()
{
  static const int m[27] = {0, 8, 1, 11, 21, 9, 3, 10, 2, 12, 25, 13, 23, 20, 24, 15, 26, 14, 4, 16, 5, 19, 22, 17, 7, 18, 6};
  for (int q = 0; q < 27; ++q) stkNodeMap_[q] = m[q];
  nodesPerElement_ = 27;
  scsDist_ = std::sqrt(3.0) / 3.0;
  scsEndLoc_[0] = -1.0; scsEndLoc_[1] = -scsDist_; scsEndLoc_[2] = scsDist_; scsEndLoc_[3] = 1.0;
  gaussAbscissae_[0] = -std::sqrt(3.0) / 3.0; gaussAbscissae_[1] = std::sqrt(3.0) / 3.0;
  const double sh[6] = {-1.0, -1.0, 0.0, 0.0, +1.0, +1.0};
  for (int q = 0; q < 6; ++q) gaussAbscissaeShift_[q] = sh[q];
}
double HexahedralP2Element::gauss_point_location(int nodeOrdinal, int gaussPointOrdinal) const
{
  const double b = scsEndLoc_[nodeOrdinal + 1], a = scsEndLoc_[nodeOrdinal]; // isoparametric_mapping(b, a, xi)
  return gaussAbscissae_[gaussPointOrdinal] * (b - a) / 2.0 + (a + b) / 2.0;
}
double HexahedralP2Element::shifted_gauss_point_location(int nodeOrdinal, int gaussPointOrdinal) const
{
  return gaussAbscissaeShift_[nodeOrdinal * 2 + gaussPointOrdinal];
}
// the 27 formulas of Hex27CVFEM.C:355-421 are the products l_a(s) l_b(t) l_c(u) of the 1-D quadratic Lagrange basis on
// {-1, 0, 1}, stored at the STK ordinal of tensor node (a, b, c)
void HexahedralP2Element::hex27_shape_fcn(int numIntPoints, const double *intgLoc, double *shpfc) const
{
  for (int ip = 0; ip < numIntPoints; ++ip) {
    double l[3][3];
    for (int d = 0; d < 3; ++d) {
      const double x = intgLoc[3 * ip + d];
      l[d][0] = -0.5 * x * (1.0 - x); l[d][1] = (1.0 - x) * (1.0 + x); l[d][2] = 0.5 * x * (1.0 + x);
    }
    for (int c = 0; c < 3; ++c)
      for (int b = 0; b < 3; ++b)
        for (int a = 0; a < 3; ++a) shpfc[27 * ip + tensor_product_node_map(a, b, c)] = l[0][a] * l[1][b] * l[2][c];
  }
}
Hex27SCS::Hex27SCS()
{
  numIntPoints_ = 216;
  const double scsLoc[2] = {-scsDist_, scsDist_};
  // u, t, s surfaces in this order; per surface 3 x 3 node segments x 2 x 2 Gauss points (Hex27CVFEM.C:912-1047)
  for (int dir = 2; dir >= 0; --dir)
    for (int m = 0; m < 2; ++m)
      for (int l = 0; l < 3; ++l)
        for (int k = 0; k < 3; ++k) {
          int lo[3], hi[3];
          if (dir == 2) { lo[0] = k; lo[1] = l; lo[2] = m; }
          else if (dir == 1) { lo[0] = k; lo[1] = m; lo[2] = l; }
          else { lo[0] = m; lo[1] = k; lo[2] = l; }
          hi[0] = lo[0]; hi[1] = lo[1]; hi[2] = lo[2];
          hi[dir] += 1;
          const int nlo = tensor_product_node_map(lo[0], lo[1], lo[2]), nhi = tensor_product_node_map(hi[0], hi[1], hi[2]);
          for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i) {
              lrscv_.push_back(m == 0 ? nlo : nhi);
              lrscv_.push_back(m == 0 ? nhi : nlo);
              double x[3], xs[3];
              const double g1 = gauss_point_location(k, i), g2 = gauss_point_location(l, j);
              const double s1 = shifted_gauss_point_location(k, i), s2 = shifted_gauss_point_location(l, j);
              if (dir == 2) { x[0] = g1; x[1] = g2; x[2] = scsLoc[m]; xs[0] = s1; xs[1] = s2; xs[2] = scsLoc[m]; }
              else if (dir == 1) { x[0] = g1; x[1] = scsLoc[m]; x[2] = g2; xs[0] = s1; xs[1] = scsLoc[m]; xs[2] = s2; }
              else { x[0] = scsLoc[m]; x[1] = g1; x[2] = g2; xs[0] = scsLoc[m]; xs[1] = s1; xs[2] = s2; }
              intgLoc_.insert(intgLoc_.end(), x, x + 3);
              intgLocShift_.insert(intgLocShift_.end(), xs, xs + 3);
            }
        }
}
Hex27SCV::Hex27SCV()
{
  numIntPoints_ = 216;
  for (int n = 0; n < 3; ++n)
    for (int m = 0; m < 3; ++m)
      for (int l = 0; l < 3; ++l)
        for (int k = 0; k < 2; ++k)
          for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i) {
              const double x[3] = {gauss_point_location(l, i), gauss_point_location(m, j), gauss_point_location(n, k)};
              const double xs[3] = {shifted_gauss_point_location(l, i), shifted_gauss_point_location(m, j),
                                    shifted_gauss_point_location(n, k)};
              intgLoc_.insert(intgLoc_.end(), x, x + 3);
              intgLocShift_.insert(intgLocShift_.end(), xs, xs + 3);
              ipNodeMap_.push_back(tensor_product_node_map(l, m, n));
            }
}

MasterElement *MasterElementRepo::get_surface_master_element(const stk::topology &theTopo)
{
  static Hex27SCS scs27;
  if (theTopo.t_ == stk::topology::HEX_27) return &scs27;
  static HexSCS scs;
  static Quad3DSCS quad;
  if (theTopo.t_ == stk::topology::QUAD_4) return &quad;
  if (theTopo.t_ != stk::topology::HEX_8) throw std::runtime_error("only Hex8 / Quad4-face master elements are on this path");
  return &scs;
}
MasterElement *MasterElementRepo::get_volume_master_element(const stk::topology &theTopo)
{
  static HexSCV scv;
  static Hex27SCV scv27;
  if (theTopo.t_ == stk::topology::HEX_27) return &scv27;
  if (theTopo.t_ != stk::topology::HEX_8) throw std::runtime_error("only Hex8 master elements are on this path");
  return &scv;
}

// ================================================================================================
// Kernels
// ================================================================================================
void Kernel::clear(nalu_b200_kernel_desc &d)
{
  std::memset(&d, 0, sizeof(d));
  d.om_dens_fac = 1.0;
  d.dt = 1.0;
  d.gamma1 = 1.0;
  d.gamma2 = -1.0;
}
void Kernel::bind(nalu_b200_kernel_desc &d, int slot, stk::mesh::FieldBase *f)
{
  if (!f) throw std::runtime_error("kernel field not found in MetaData");
  d.nodal[slot] = f->device_data();
}

template <typename AlgTraits>
MomentumAdvDiffElemKernel<AlgTraits>::MomentumAdvDiffElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                               VectorFieldType *velocity, ScalarFieldType *viscosity,
                                                               ElemDataRequests &dataPreReqs)
    : Kernel(), viscosity_(viscosity),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes()),
      includeDivU_(solnOpts.includeDivU_), shiftedGradOp_(solnOpts.get_shifted_grad_op(velocity->name()))
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  velocityNp1_ = &(velocity->field_of_state(stk::mesh::StateNP1));
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  massFlowRate_ = metaData.get_field<double>(stk::topology::ELEMENT_RANK, "mass_flow_rate_scs");
  STK_ThrowRequireMsg(coordinates_ && massFlowRate_, "MomentumAdvDiffElemKernel: coordinates / mass_flow_rate_scs not registered");
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
  skewSymmetric_ = solnOpts.get_skew_symmetric(velocity->name());
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { skewSymmetric_ ? meSCS->shifted_shape_fcn(ptr) : meSCS->shape_fcn(ptr); },
                                   v_adv_shape_function_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*velocityNp1_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  dataPreReqs.add_element_field(*massFlowRate_, AlgTraits::numScsIp_);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_GRAD_OP : SCS_GRAD_OP, CURRENT_COORDINATES);
}
template <typename AlgTraits>
bool MomentumAdvDiffElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_MOMENTUM_ADV_DIFF;
  d.options = (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0) | (skewSymmetric_ ? NALU_OPT_SKEW_SYMMETRIC : 0);
  d.include_divu = includeDivU_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  bind(d, NALU_F_VISCOSITY, viscosity_);
  d.mass_flow_rate = massFlowRate_->device_data();
  return true;
}

template <typename AlgTraits>
ContinuityAdvElemKernel<AlgTraits>::ContinuityAdvElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                           ElemDataRequests &dataPreReqs)
    : Kernel(), meshMotion_(solnOpts.does_mesh_move()), shiftMdot_(solnOpts.cvfemShiftMdot_),
      shiftPoisson_(solnOpts.get_shifted_grad_op("pressure")), reducedSensitivities_(solnOpts.cvfemReducedSensPoisson_),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  const std::string velocity_name = meshMotion_ ? "velocity_rtm" : "velocity";
  velocityRTM_ = metaData.get_field<double>(stk::topology::NODE_RANK, velocity_name);
  Gpdx_ = metaData.get_field<double>(stk::topology::NODE_RANK, "dpdx");
  pressure_ = metaData.get_field<double>(stk::topology::NODE_RANK, "pressure");
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  STK_ThrowRequireMsg(velocityRTM_ && Gpdx_ && pressure_ && density, "ContinuityAdvElemKernel: required fields not registered");
  densityNp1_ = &(density->field_of_state(stk::mesh::StateNP1));
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*velocityRTM_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_gathered_nodal_field(*pressure_, 1);
  dataPreReqs.add_gathered_nodal_field(*Gpdx_, AlgTraits::nDim_);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  if (!shiftPoisson_ || !reducedSensitivities_) dataPreReqs.add_master_element_call(SCS_GRAD_OP, CURRENT_COORDINATES);
  if (shiftPoisson_ || reducedSensitivities_) dataPreReqs.add_master_element_call(SCS_SHIFTED_GRAD_OP, CURRENT_COORDINATES);
  if (shiftMdot_)
    get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shifted_shape_fcn(ptr); }, v_shape_function_);
  else
    get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
}
template <typename AlgTraits>
void ContinuityAdvElemKernel<AlgTraits>::setup(const TimeIntegrator &timeIntegrator)
{
  dt_ = timeIntegrator.get_time_step();
  gamma1_ = timeIntegrator.get_gamma1();
  projTimeScale_ = dt_ / gamma1_;
}
template <typename AlgTraits>
bool ContinuityAdvElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_CONTINUITY_ADV;
  d.options = (shiftMdot_ ? NALU_OPT_SHIFT_MDOT : 0) | (shiftPoisson_ ? NALU_OPT_SHIFT_POISSON : 0) |
              (reducedSensitivities_ ? NALU_OPT_REDUCED_SENS : 0);
  d.dt = dt_;
  d.gamma1 = gamma1_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityRTM_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_PRESSURE, pressure_);
  bind(d, NALU_F_DPDX, Gpdx_);
  return true;
}

template <typename AlgTraits>
ScalarAdvDiffElemKernel<AlgTraits>::ScalarAdvDiffElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                           ScalarFieldType *scalarQ, ScalarFieldType *diffFluxCoeff,
                                                           ElemDataRequests &dataPreReqs)
    : Kernel(), scalarQ_(scalarQ), diffFluxCoeff_(diffFluxCoeff),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes()),
      shiftedGradOp_(solnOpts.get_shifted_grad_op(scalarQ->name()))
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  massFlowRate_ = metaData.get_field<double>(stk::topology::ELEMENT_RANK, "mass_flow_rate_scs");
  STK_ThrowRequireMsg(coordinates_ && massFlowRate_, "ScalarAdvDiffElemKernel: coordinates / mass_flow_rate_scs not registered");
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
  skewSymmetric_ = solnOpts.get_skew_symmetric(scalarQ->name());
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { skewSymmetric_ ? meSCS->shifted_shape_fcn(ptr) : meSCS->shape_fcn(ptr); },
                                   v_adv_shape_function_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*scalarQ_, 1);
  dataPreReqs.add_gathered_nodal_field(*diffFluxCoeff_, 1);
  dataPreReqs.add_element_field(*massFlowRate_, AlgTraits::numScsIp_);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_GRAD_OP : SCS_GRAD_OP, CURRENT_COORDINATES);
}
template <typename AlgTraits>
bool ScalarAdvDiffElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_SCALAR_ADV_DIFF;
  d.options = (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0) | (skewSymmetric_ ? NALU_OPT_SKEW_SYMMETRIC : 0);
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_SCALARQ, scalarQ_);
  bind(d, NALU_F_VISCOSITY, diffFluxCoeff_);
  d.mass_flow_rate = massFlowRate_->device_data();
  return true;
}

template <typename AlgTraits>
ScalarDiffElemKernel<AlgTraits>::ScalarDiffElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                     ScalarFieldType *scalarQ, ScalarFieldType *diffFluxCoeff,
                                                     ElemDataRequests &dataPreReqs)
    : Kernel(), scalarQ_(scalarQ), diffFluxCoeff_(diffFluxCoeff),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes()),
      shiftedGradOp_(solnOpts.get_shifted_grad_op(scalarQ_->name()))
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(coordinates_, "ScalarDiffElemKernel: coordinates not registered");
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*scalarQ_, 1);
  dataPreReqs.add_gathered_nodal_field(*diffFluxCoeff_, 1);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_GRAD_OP : SCS_GRAD_OP, CURRENT_COORDINATES);
}
template <typename AlgTraits>
bool ScalarDiffElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_SCALAR_DIFF;
  d.options = shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_SCALARQ, scalarQ_);
  bind(d, NALU_F_VISCOSITY, diffFluxCoeff_);
  return true;
}

namespace {
// StateN / StateNM1 selection of the mass kernels (src/kernel/MomentumMassElemKernel.C:46-56): with two
// states NM1 aliases N
template <typename F>
void pick_states(F *f, F *&nm1, F *&n, F *&np1)
{
  n = &(f->field_of_state(stk::mesh::StateN));
  np1 = &(f->field_of_state(stk::mesh::StateNP1));
  nm1 = f->number_of_states() == 2 ? n : &(f->field_of_state(stk::mesh::StateNM1));
}
} // namespace

template <typename AlgTraits>
MomentumMassElemKernel<AlgTraits>::MomentumMassElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                         ElemDataRequests &dataPreReqs, const bool lumpedMass)
    : Kernel(), lumpedMass_(lumpedMass), densFac_(solnOpts.balancedForce_ ? 1.0 : 0.0),
      om_densFac_(solnOpts.balancedForce_ ? 0.0 : 1.0),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  VectorFieldType *velocity = metaData.get_field<double>(stk::topology::NODE_RANK, "velocity");
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  STK_ThrowRequireMsg(velocity && density, "MomentumMassElemKernel: velocity / density not registered");
  STK_ThrowRequireMsg(velocity->number_of_states() >= 2 && density->number_of_states() >= 2,
                      "MomentumMassElemKernel: velocity and density need at least two states");
  pick_states(velocity, velocityNm1_, velocityN_, velocityNp1_);
  pick_states(density, densityNm1_, densityN_, densityNp1_);
  Gjp_ = metaData.get_field<double>(stk::topology::NODE_RANK, "dpdx");
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(Gjp_ && coordinates_, "MomentumMassElemKernel: dpdx / coordinates not registered");
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  if (lumpedMass_)
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shifted_shape_fcn(ptr); }, v_shape_function_);
  else
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*densityNm1_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityN_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_gathered_nodal_field(*velocityNm1_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*velocityN_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*velocityNp1_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*Gjp_, AlgTraits::nDim_);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
}
template <typename AlgTraits>
void MomentumMassElemKernel<AlgTraits>::setup(const TimeIntegrator &ti)
{
  dt_ = ti.get_time_step();
  gamma1_ = ti.get_gamma1();
  gamma2_ = ti.get_gamma2();
  gamma3_ = ti.get_gamma3();
}
template <typename AlgTraits>
bool MomentumMassElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_MOMENTUM_MASS;
  d.options = lumpedMass_ ? NALU_OPT_LUMPED_MASS : 0;
  d.dens_fac = densFac_;
  d.om_dens_fac = om_densFac_;
  d.dt = dt_; d.gamma1 = gamma1_; d.gamma2 = gamma2_; d.gamma3 = gamma3_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  bind(d, NALU_F_VELOCITY_N, velocityN_);
  bind(d, NALU_F_VELOCITY_NM1, velocityNm1_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_DENSITY_N, densityN_);
  bind(d, NALU_F_DENSITY_NM1, densityNm1_);
  bind(d, NALU_F_DPDX, Gjp_);
  return true;
}

template <typename AlgTraits>
ContinuityMassElemKernel<AlgTraits>::ContinuityMassElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                             ElemDataRequests &dataPreReqs, const bool lumpedMass)
    : Kernel(), lumpedMass_(lumpedMass), ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  STK_ThrowRequireMsg(density && density->number_of_states() >= 2, "ContinuityMassElemKernel: density needs at least two states");
  pick_states(density, densityNm1_, densityN_, densityNp1_);
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(coordinates_, "ContinuityMassElemKernel: coordinates not registered");
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  if (lumpedMass_)
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shifted_shape_fcn(ptr); }, v_shape_function_);
  else
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*densityNm1_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityN_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
}
template <typename AlgTraits>
void ContinuityMassElemKernel<AlgTraits>::setup(const TimeIntegrator &ti)
{
  dt_ = ti.get_time_step();
  gamma1_ = ti.get_gamma1();
  gamma2_ = ti.get_gamma2();
  gamma3_ = ti.get_gamma3();
}
template <typename AlgTraits>
bool ContinuityMassElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_CONTINUITY_MASS;
  d.options = lumpedMass_ ? NALU_OPT_LUMPED_MASS : 0;
  d.dt = dt_; d.gamma1 = gamma1_; d.gamma2 = gamma2_; d.gamma3 = gamma3_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_DENSITY_N, densityN_);
  bind(d, NALU_F_DENSITY_NM1, densityNm1_);
  return true;
}

template <typename AlgTraits>
ScalarMassElemKernel<AlgTraits>::ScalarMassElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                     ScalarFieldType *scalarQ, ElemDataRequests &dataPreReqs, const bool lumpedMass)
    : Kernel(), lumpedMass_(lumpedMass), ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  STK_ThrowRequireMsg(scalarQ->number_of_states() >= 2, "ScalarMassElemKernel: scalarQ needs at least two states");
  pick_states(scalarQ, scalarQNm1_, scalarQN_, scalarQNp1_);
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  STK_ThrowRequireMsg(density && density->number_of_states() >= 2, "ScalarMassElemKernel: density needs at least two states");
  pick_states(density, densityNm1_, densityN_, densityNp1_);
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(coordinates_, "ScalarMassElemKernel: coordinates not registered");
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  if (lumpedMass_)
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shifted_shape_fcn(ptr); }, v_shape_function_);
  else
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*scalarQNm1_, 1);
  dataPreReqs.add_gathered_nodal_field(*scalarQN_, 1);
  dataPreReqs.add_gathered_nodal_field(*scalarQNp1_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNm1_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityN_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
}
template <typename AlgTraits>
void ScalarMassElemKernel<AlgTraits>::setup(const TimeIntegrator &ti)
{
  dt_ = ti.get_time_step();
  gamma1_ = ti.get_gamma1();
  gamma2_ = ti.get_gamma2();
  gamma3_ = ti.get_gamma3();
}
template <typename AlgTraits>
bool ScalarMassElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_SCALAR_MASS;
  d.options = lumpedMass_ ? NALU_OPT_LUMPED_MASS : 0;
  d.dt = dt_; d.gamma1 = gamma1_; d.gamma2 = gamma2_; d.gamma3 = gamma3_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_SCALARQ, scalarQNp1_);
  bind(d, NALU_F_SCALARQ_N, scalarQN_);
  bind(d, NALU_F_SCALARQ_NM1, scalarQNm1_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_DENSITY_N, densityN_);
  bind(d, NALU_F_DENSITY_NM1, densityNm1_);
  return true;
}

// ---- src/kernel/MomentumBuoyancySrcElemKernel.C:26-54 / MomentumBuoyancyBoussinesqSrcElemKernel.C:26-59
template <typename AlgTraits>
MomentumBuoyancySrcElemKernel<AlgTraits>::MomentumBuoyancySrcElemKernel(const stk::mesh::BulkData &bulkData,
                                                                       const SolutionOptions &solnOpts,
                                                                       ElemDataRequests &dataPreReqs)
    : Kernel(), rhoRef_(solnOpts.referenceDensity_),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  STK_ThrowRequireMsg(density, "MomentumBuoyancySrcElemKernel: density not registered");
  densityNp1_ = &(density->field_of_state(stk::mesh::StateNP1));
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(coordinates_, "MomentumBuoyancySrcElemKernel: coordinates not registered");
  const std::array<double, 3> g = solnOpts.get_gravity_vector();
  for (int i = 0; i < AlgTraits::nDim_; ++i) gravity_[i] = g[i];
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
}
template <typename AlgTraits>
bool MomentumBuoyancySrcElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_MOMENTUM_BUOYANCY_SRC;
  d.rho_ref = rhoRef_;
  for (int i = 0; i < 3; ++i) d.gravity[i] = gravity_[i];
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  return true;
}

template <typename AlgTraits>
MomentumBuoyancyBoussinesqSrcElemKernel<AlgTraits>::MomentumBuoyancyBoussinesqSrcElemKernel(
    const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts, ElemDataRequests &dataPreReqs)
    : Kernel(), rhoRef_(solnOpts.referenceDensity_), tRef_(solnOpts.referenceTemperature_),
      beta_(solnOpts.thermalExpansionCoeff_),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  ScalarFieldType *temperature = metaData.get_field<double>(stk::topology::NODE_RANK, "temperature");
  STK_ThrowRequireMsg(temperature, "MomentumBuoyancyBoussinesqSrcElemKernel: temperature not registered");
  temperatureNp1_ = &(temperature->field_of_state(stk::mesh::StateNP1));
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(coordinates_, "MomentumBuoyancyBoussinesqSrcElemKernel: coordinates not registered");
  const std::array<double, 3> g = solnOpts.get_gravity_vector();
  for (int i = 0; i < AlgTraits::nDim_; ++i) gravity_[i] = g[i];
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*temperatureNp1_, 1);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
}
template <typename AlgTraits>
bool MomentumBuoyancyBoussinesqSrcElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_MOMENTUM_BUOYANCY_BOUSSINESQ_SRC;
  d.rho_ref = rhoRef_;
  d.t_ref = tRef_;
  d.beta = beta_;
  for (int i = 0; i < 3; ++i) d.gravity[i] = gravity_[i];
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_TEMPERATURE, temperatureNp1_);
  return true;
}

// ---- src/kernel/TurbKineticEnergySSTSrcElemKernel.C:28-90, SpecificDissipationRateSSTSrcElemKernel.C:26-97
namespace {
template <typename AlgTraits, typename K>
void sst_common_registrations(K &k, const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                              ElemDataRequests &dataPreReqs, bool lumpedMass, bool shiftedGradOp, ScalarFieldType *&tkeNp1,
                              ScalarFieldType *&sdrNp1, ScalarFieldType *&densityNp1, VectorFieldType *&velocityNp1,
                              ScalarFieldType *&tvisc, VectorFieldType *&coordinates, ScalarFieldType **fOneBlend)
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  ScalarFieldType *tke = metaData.get_field<double>(stk::topology::NODE_RANK, "turbulent_ke");
  ScalarFieldType *sdr = metaData.get_field<double>(stk::topology::NODE_RANK, "specific_dissipation_rate");
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  VectorFieldType *velocity = metaData.get_field<double>(stk::topology::NODE_RANK, "velocity");
  tvisc = metaData.get_field<double>(stk::topology::NODE_RANK, "turbulent_viscosity");
  coordinates = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(tke && sdr && density && velocity && tvisc && coordinates, "SST source kernel: field not registered");
  tkeNp1 = &tke->field_of_state(stk::mesh::StateNP1);
  sdrNp1 = &sdr->field_of_state(stk::mesh::StateNP1);
  densityNp1 = &density->field_of_state(stk::mesh::StateNP1);
  velocityNp1 = &velocity->field_of_state(stk::mesh::StateNP1);
  if (fOneBlend) {
    *fOneBlend = metaData.get_field<double>(stk::topology::NODE_RANK, "sst_f_one_blending");
    STK_ThrowRequireMsg(*fOneBlend, "SST source kernel: sst_f_one_blending not registered");
  }
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  if (lumpedMass)
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shifted_shape_fcn(ptr); }, k.v_shape_function_);
  else
    get_scv_shape_fn_data<AlgTraits>([&](double *ptr) { meSCV->shape_fcn(ptr); }, k.v_shape_function_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*tkeNp1, 1);
  dataPreReqs.add_gathered_nodal_field(*sdrNp1, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNp1, 1);
  dataPreReqs.add_gathered_nodal_field(*velocityNp1, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*tvisc, 1);
  if (fOneBlend) dataPreReqs.add_gathered_nodal_field(**fOneBlend, 1);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp ? SCV_SHIFTED_GRAD_OP : SCV_GRAD_OP, CURRENT_COORDINATES);
}
} // namespace

template <typename AlgTraits>
TurbKineticEnergySSTSrcElemKernel<AlgTraits>::TurbKineticEnergySSTSrcElemKernel(const stk::mesh::BulkData &bulkData,
                                                                               const SolutionOptions &solnOpts,
                                                                               ElemDataRequests &dataPreReqs,
                                                                               const bool lumpedMass)
    : Kernel(), lumpedMass_(lumpedMass), shiftedGradOp_(solnOpts.get_shifted_grad_op("velocity")),
      betaStar_(solnOpts.get_turb_model_constant(TM_betaStar)),
      tkeProdLimitRatio_(solnOpts.get_turb_model_constant(TM_tkeProdLimitRatio)),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  sst_common_registrations<AlgTraits>(*this, bulkData, solnOpts, dataPreReqs, lumpedMass_, shiftedGradOp_, tkeNp1_, sdrNp1_,
                                      densityNp1_, velocityNp1_, tvisc_, coordinates_, nullptr);
}
template <typename AlgTraits>
bool TurbKineticEnergySSTSrcElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_TKE_SST_SRC;
  d.options = (lumpedMass_ ? NALU_OPT_LUMPED_MASS : 0) | (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0);
  d.turb[NALU_TM_BETA_STAR] = betaStar_;
  d.turb[NALU_TM_TKE_PROD_LIMIT_RATIO] = tkeProdLimitRatio_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_TKE, tkeNp1_);
  bind(d, NALU_F_SDR, sdrNp1_);
  bind(d, NALU_F_TVISC, tvisc_);
  return true;
}

template <typename AlgTraits>
SpecificDissipationRateSSTSrcElemKernel<AlgTraits>::SpecificDissipationRateSSTSrcElemKernel(
    const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts, ElemDataRequests &dataPreReqs, const bool lumpedMass)
    : Kernel(), lumpedMass_(lumpedMass), shiftedGradOp_(solnOpts.get_shifted_grad_op("velocity")),
      betaStar_(solnOpts.get_turb_model_constant(TM_betaStar)), sigmaWTwo_(solnOpts.get_turb_model_constant(TM_sigmaWTwo)),
      betaOne_(solnOpts.get_turb_model_constant(TM_betaOne)), betaTwo_(solnOpts.get_turb_model_constant(TM_betaTwo)),
      gammaOne_(solnOpts.get_turb_model_constant(TM_gammaOne)), gammaTwo_(solnOpts.get_turb_model_constant(TM_gammaTwo)),
      tkeProdLimitRatio_(solnOpts.get_turb_model_constant(TM_tkeProdLimitRatio)),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  sst_common_registrations<AlgTraits>(*this, bulkData, solnOpts, dataPreReqs, lumpedMass_, shiftedGradOp_, tkeNp1_, sdrNp1_,
                                      densityNp1_, velocityNp1_, tvisc_, coordinates_, &fOneBlend_);
}
template <typename AlgTraits>
bool SpecificDissipationRateSSTSrcElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_SDR_SST_SRC;
  d.options = (lumpedMass_ ? NALU_OPT_LUMPED_MASS : 0) | (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0);
  d.turb[NALU_TM_BETA_STAR] = betaStar_;
  d.turb[NALU_TM_TKE_PROD_LIMIT_RATIO] = tkeProdLimitRatio_;
  d.turb[NALU_TM_SIGMA_W_TWO] = sigmaWTwo_;
  d.turb[NALU_TM_BETA_ONE] = betaOne_;
  d.turb[NALU_TM_BETA_TWO] = betaTwo_;
  d.turb[NALU_TM_GAMMA_ONE] = gammaOne_;
  d.turb[NALU_TM_GAMMA_TWO] = gammaTwo_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_TKE, tkeNp1_);
  bind(d, NALU_F_SDR, sdrNp1_);
  bind(d, NALU_F_TVISC, tvisc_);
  bind(d, NALU_F_FONE_BLEND, fOneBlend_);
  return true;
}

// ---- src/kernel/TurbKineticEnergySSTDESSrcElemKernel.C:27-99
template <typename AlgTraits>
TurbKineticEnergySSTDESSrcElemKernel<AlgTraits>::TurbKineticEnergySSTDESSrcElemKernel(const stk::mesh::BulkData &bulkData,
                                                                                     const SolutionOptions &solnOpts,
                                                                                     ElemDataRequests &dataPreReqs,
                                                                                     const bool lumpedMass)
    : Kernel(), lumpedMass_(lumpedMass), shiftedGradOp_(solnOpts.get_shifted_grad_op("velocity")),
      betaStar_(solnOpts.get_turb_model_constant(TM_betaStar)),
      tkeProdLimitRatio_(solnOpts.get_turb_model_constant(TM_tkeProdLimitRatio)),
      cDESke_(solnOpts.get_turb_model_constant(TM_cDESke)), cDESkw_(solnOpts.get_turb_model_constant(TM_cDESkw)),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  sst_common_registrations<AlgTraits>(*this, bulkData, solnOpts, dataPreReqs, lumpedMass_, shiftedGradOp_, tkeNp1_, sdrNp1_,
                                      densityNp1_, velocityNp1_, tvisc_, coordinates_, &fOneBlend_);
  maxLengthScale_ = bulkData.mesh_meta_data().get_field<double>(stk::topology::NODE_RANK, "sst_max_length_scale");
  STK_ThrowRequireMsg(maxLengthScale_, "SST-DES source kernel: sst_max_length_scale not registered");
  dataPreReqs.add_gathered_nodal_field(*maxLengthScale_, 1);
}
template <typename AlgTraits>
bool TurbKineticEnergySSTDESSrcElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_TKE_SST_DES_SRC;
  d.options = (lumpedMass_ ? NALU_OPT_LUMPED_MASS : 0) | (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0);
  d.turb[NALU_TM_BETA_STAR] = betaStar_;
  d.turb[NALU_TM_TKE_PROD_LIMIT_RATIO] = tkeProdLimitRatio_;
  d.turb[NALU_TM_CDES_KE] = cDESke_;
  d.turb[NALU_TM_CDES_KW] = cDESkw_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_TKE, tkeNp1_);
  bind(d, NALU_F_SDR, sdrNp1_);
  bind(d, NALU_F_TVISC, tvisc_);
  bind(d, NALU_F_FONE_BLEND, fOneBlend_);
  bind(d, NALU_F_SST_MAX_LENGTH_SCALE, maxLengthScale_);
  return true;
}

// ---- src/kernel/TurbKineticEnergyKsgsSrcElemKernel.C:27-72
template <typename AlgTraits>
TurbKineticEnergyKsgsSrcElemKernel<AlgTraits>::TurbKineticEnergyKsgsSrcElemKernel(const stk::mesh::BulkData &bulkData,
                                                                                 const SolutionOptions &solnOpts,
                                                                                 ElemDataRequests &dataPreReqs, double lrksgsfac)
    : Kernel(), lrksgsfac_(lrksgsfac), tkeProdLimitRatio_(solnOpts.get_turb_model_constant(TM_tkeProdLimitRatio)),
      ipNodeMap_(MasterElementRepo::get_volume_master_element(AlgTraits::topo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  auto scalar = [&](const char *n) { return metaData.get_field<double>(stk::topology::NODE_RANK, n); };
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  tkeNp1_ = scalar("turbulent_ke");
  densityNp1_ = scalar("density");
  tvisc_ = scalar("turbulent_viscosity");
  dualNodalVolume_ = scalar("dual_nodal_volume");
  cEps_ = scalar("c_epsilon");
  Gju_ = metaData.get_field<double>(stk::topology::NODE_RANK, "dudx");
  visc_ = scalar("viscosity");
  // a field the low-Re form alone reads is bound to one that always exists otherwise (:51-55)
  dsqrtkSq_ = lrksgsfac > 0.0 ? scalar("dsqrtk_dx_sq") : visc_;
  STK_ThrowRequireMsg(coordinates_ && tkeNp1_ && densityNp1_ && tvisc_ && dualNodalVolume_ && cEps_ && Gju_ && visc_ && dsqrtkSq_,
                      "Ksgs source kernel: field not registered");
  MasterElement *meSCV = MasterElementRepo::get_volume_master_element(AlgTraits::topo_);
  dataPreReqs.add_cvfem_volume_me(meSCV);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*tkeNp1_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_gathered_nodal_field(*tvisc_, 1);
  dataPreReqs.add_gathered_nodal_field(*dualNodalVolume_, 1);
  dataPreReqs.add_gathered_nodal_field(*cEps_, 1);
  dataPreReqs.add_gathered_nodal_field(*Gju_, AlgTraits::nDim_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*visc_, 1);
  dataPreReqs.add_gathered_nodal_field(*dsqrtkSq_, 1);
  dataPreReqs.add_master_element_call(SCV_VOLUME, CURRENT_COORDINATES);
}
template <typename AlgTraits>
bool TurbKineticEnergyKsgsSrcElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_TKE_KSGS_SRC;
  d.turb[NALU_TM_TKE_PROD_LIMIT_RATIO] = tkeProdLimitRatio_;
  d.turb[NALU_TM_LRKSGS_FAC] = lrksgsfac_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_TKE, tkeNp1_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_TVISC, tvisc_);
  bind(d, NALU_F_DUAL_VOLUME, dualNodalVolume_);
  bind(d, NALU_F_C_EPSILON, cEps_);
  bind(d, NALU_F_GJU, Gju_);
  bind(d, NALU_F_VISCOSITY, visc_);
  bind(d, NALU_F_DSQRTK_DX_SQ, dsqrtkSq_);
  return true;
}

// ---- src/nso/MomentumNSOElemKernel.C:27-125
template <typename AlgTraits>
MomentumNSOElemKernel<AlgTraits>::MomentumNSOElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                       VectorFieldType *velocity, GenericFieldType *Gju, ScalarFieldType *viscosity,
                                                       const double fourthFac, const double altResFac, ElemDataRequests &dataPreReqs)
    : Kernel(), viscosity_(viscosity), Gju_(Gju),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes()), fourthFac_(fourthFac),
      altResFac_(altResFac), om_altResFac_(1.0 - altResFac), includeDivU_(solnOpts.includeDivU_),
      shiftedGradOp_(solnOpts.get_shifted_grad_op(velocity->name()))
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  ScalarFieldType *density = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  STK_ThrowRequireMsg(velocity && Gju && viscosity && density, "MomentumNSOElemKernel: field not registered");
  pick_states(velocity, velocityNm1_, velocityN_, velocityNp1_);
  pick_states(density, densityNm1_, densityN_, densityNp1_);
  velocityRTM_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.does_mesh_move() ? "velocity_rtm" : "velocity");
  pressure_ = metaData.get_field<double>(stk::topology::NODE_RANK, "pressure");
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  STK_ThrowRequireMsg(velocityRTM_ && pressure_ && coordinates_, "MomentumNSOElemKernel: velocity_rtm / pressure / coordinates not registered");
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_gathered_nodal_field(*Gju_, AlgTraits::nDim_, AlgTraits::nDim_);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*velocityNm1_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*velocityN_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*velocityNp1_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*velocityRTM_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*densityNm1_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityN_, 1);
  dataPreReqs.add_gathered_nodal_field(*densityNp1_, 1);
  dataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  dataPreReqs.add_gathered_nodal_field(*pressure_, 1);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_GRAD_OP : SCS_GRAD_OP, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(SCS_GIJ, CURRENT_COORDINATES);
}
template <typename AlgTraits>
void MomentumNSOElemKernel<AlgTraits>::setup(const TimeIntegrator &ti)
{
  dt_ = ti.get_time_step();
  gamma1_ = ti.get_gamma1();
  gamma2_ = ti.get_gamma2();
  gamma3_ = ti.get_gamma3(); // gamma3 may be zero
}
template <typename AlgTraits>
bool MomentumNSOElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_MOMENTUM_NSO;
  d.options = shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0;
  d.include_divu = includeDivU_;
  d.dt = dt_;
  d.gamma1 = gamma1_;
  d.gamma2 = gamma2_;
  d.gamma3 = gamma3_;
  d.nso_fourth_fac = fourthFac_;
  d.nso_alt_res_fac = altResFac_;
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_GJU, Gju_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  bind(d, NALU_F_VELOCITY_N, velocityN_);
  bind(d, NALU_F_VELOCITY_NM1, velocityNm1_);
  if (velocityRTM_ != velocityNp1_) bind(d, NALU_F_VELOCITY_RTM, velocityRTM_);
  bind(d, NALU_F_DENSITY, densityNp1_);
  bind(d, NALU_F_DENSITY_N, densityN_);
  bind(d, NALU_F_DENSITY_NM1, densityNm1_);
  bind(d, NALU_F_VISCOSITY, viscosity_);
  bind(d, NALU_F_PRESSURE, pressure_);
  return true;
}

// ---- upwind adv-diff kernels --------------------------------------------------------------------
namespace {
// EquationSystem::create_peclet_function (include/EquationSystem.h:306-325), captured as data
PecletFunctionSpec make_peclet_spec(EquationSystem *eqSystem, const std::string &dofName)
{
  STK_ThrowRequireMsg(eqSystem, "upwind kernel: an EquationSystem is needed to create the Peclet function");
  PecletFunctionSpec p;
  Realm &realm = eqSystem->realm_;
  if ("classic" == realm.get_tanh_functional_form(dofName)) {
    p.tanh_form = false;
    p.hybridFactor = realm.get_hybrid_factor(dofName);
  } else {
    p.tanh_form = true;
    p.c1 = realm.get_tanh_trans(dofName);
    p.c2 = realm.get_tanh_width(dofName);
  }
  return p;
}
void upw_fill(nalu_b200_kernel_desc &d, const PecletFunctionSpec &p, double alpha, double alphaUpw, double hoUpwind, bool useLimiter)
{
  d.upw[NALU_UPW_ALPHA] = alpha;
  d.upw[NALU_UPW_ALPHA_UPW] = alphaUpw;
  d.upw[NALU_UPW_HO_UPWIND] = hoUpwind;
  d.upw[NALU_UPW_HYBRID_FACTOR] = p.hybridFactor;
  d.upw[NALU_UPW_TANH_TRANS] = p.c1;
  d.upw[NALU_UPW_TANH_WIDTH] = p.c2;
  if (useLimiter) d.options |= NALU_OPT_USE_LIMITER;
  if (p.tanh_form) d.options |= NALU_OPT_PECLET_TANH;
}
} // namespace

template <typename AlgTraits>
ScalarUpwAdvDiffElemKernel<AlgTraits>::ScalarUpwAdvDiffElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                                 EquationSystem *eqSystem, ScalarFieldType *scalarQ,
                                                                 VectorFieldType *Gjq, ScalarFieldType *diffFluxCoeff,
                                                                 ElemDataRequests &dataPreReqs)
    : Kernel(), solnOpts_(solnOpts), scalarQ_(scalarQ), Gjq_(Gjq), diffFluxCoeff_(diffFluxCoeff),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes()), dofName_(scalarQ->name()),
      alpha_(solnOpts.get_alpha_factor(dofName_)), alphaUpw_(solnOpts.get_alpha_upw_factor(dofName_)),
      hoUpwind_(solnOpts.get_upw_factor(dofName_)), useLimiter_(solnOpts.primitive_uses_limiter(dofName_)),
      om_alpha_(1.0 - alpha_), om_alphaUpw_(1.0 - alphaUpw_), shiftedGradOp_(solnOpts.get_shifted_grad_op(scalarQ->name())),
      pecletFunction_(make_peclet_spec(eqSystem, dofName_))
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  massFlowRate_ = metaData.get_field<double>(stk::topology::ELEMENT_RANK, "mass_flow_rate_scs");
  density_ = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  velocityRTM_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.does_mesh_move() ? "velocity_rtm" : "velocity");
  STK_ThrowRequireMsg(coordinates_ && massFlowRate_ && density_ && velocityRTM_ && Gjq_,
                      "ScalarUpwAdvDiffElemKernel: required fields not registered");
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
  skewSymmetric_ = solnOpts.get_skew_symmetric(scalarQ->name());
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { skewSymmetric_ ? meSCS->shifted_shape_fcn(ptr) : meSCS->shape_fcn(ptr); },
                                   v_adv_shape_function_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_gathered_nodal_field(*velocityRTM_, AlgTraits::nDim_);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*Gjq_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*scalarQ_, 1);
  dataPreReqs.add_gathered_nodal_field(*density_, 1);
  dataPreReqs.add_gathered_nodal_field(*diffFluxCoeff_, 1);
  dataPreReqs.add_element_field(*massFlowRate_, AlgTraits::numScsIp_);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_GRAD_OP : SCS_GRAD_OP, CURRENT_COORDINATES);
}
template <typename AlgTraits>
void ScalarUpwAdvDiffElemKernel<AlgTraits>::setup(const TimeIntegrator &)
{
  alpha_ = solnOpts_.get_alpha_factor(dofName_);
  alphaUpw_ = solnOpts_.get_alpha_upw_factor(dofName_);
  hoUpwind_ = solnOpts_.get_upw_factor(dofName_);
  useLimiter_ = solnOpts_.primitive_uses_limiter(dofName_);
  om_alpha_ = 1.0 - alpha_;
  om_alphaUpw_ = 1.0 - alphaUpw_;
}
template <typename AlgTraits>
bool ScalarUpwAdvDiffElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_SCALAR_UPW_ADV_DIFF;
  d.options = (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0) | (skewSymmetric_ ? NALU_OPT_SKEW_SYMMETRIC : 0);
  upw_fill(d, pecletFunction_, alpha_, alphaUpw_, hoUpwind_, useLimiter_);
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityRTM_);
  bind(d, NALU_F_GJQ, Gjq_);
  bind(d, NALU_F_SCALARQ, scalarQ_);
  bind(d, NALU_F_DENSITY, density_);
  bind(d, NALU_F_VISCOSITY, diffFluxCoeff_);
  d.mass_flow_rate = massFlowRate_->device_data();
  return true;
}

template <typename AlgTraits>
MomentumUpwAdvDiffElemKernel<AlgTraits>::MomentumUpwAdvDiffElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                                     EquationSystem *eqSystem, VectorFieldType *velocity,
                                                                     ScalarFieldType *viscosity, GenericFieldType *Gju,
                                                                     ElemDataRequests &dataPreReqs)
    : Kernel(), solnOpts_(solnOpts), viscosity_(viscosity), Gju_(Gju),
      lrscv_(MasterElementRepo::get_surface_master_element(AlgTraits::topo_)->adjacentNodes()), dofName_(velocity->name()),
      alpha_(solnOpts.get_alpha_factor(dofName_)), alphaUpw_(solnOpts.get_alpha_upw_factor(dofName_)),
      hoUpwind_(solnOpts.get_upw_factor(dofName_)), useLimiter_(solnOpts.primitive_uses_limiter(dofName_)),
      om_alpha_(1.0 - alpha_), om_alphaUpw_(1.0 - alphaUpw_), includeDivU_(solnOpts.includeDivU_),
      shiftedGradOp_(solnOpts.get_shifted_grad_op(velocity->name())), pecletFunction_(make_peclet_spec(eqSystem, dofName_))
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  velocityRTM_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.does_mesh_move() ? "velocity_rtm" : "velocity");
  velocityNp1_ = &(velocity->field_of_state(stk::mesh::StateNP1));
  coordinates_ = metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name());
  density_ = metaData.get_field<double>(stk::topology::NODE_RANK, "density");
  massFlowRate_ = metaData.get_field<double>(stk::topology::ELEMENT_RANK, "mass_flow_rate_scs");
  STK_ThrowRequireMsg(coordinates_ && massFlowRate_ && density_ && velocityRTM_ && Gju_,
                      "MomentumUpwAdvDiffElemKernel: required fields not registered");
  MasterElement *meSCS = MasterElementRepo::get_surface_master_element(AlgTraits::topo_);
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { meSCS->shape_fcn(ptr); }, v_shape_function_);
  skewSymmetric_ = solnOpts.get_skew_symmetric(velocity->name());
  get_scs_shape_fn_data<AlgTraits>([&](double *ptr) { skewSymmetric_ ? meSCS->shifted_shape_fcn(ptr) : meSCS->shape_fcn(ptr); },
                                   v_adv_shape_function_);
  dataPreReqs.add_cvfem_surface_me(meSCS);
  dataPreReqs.add_gathered_nodal_field(*Gju_, AlgTraits::nDim_, AlgTraits::nDim_);
  dataPreReqs.add_coordinates_field(*coordinates_, AlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*velocityRTM_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*velocityNp1_, AlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*density_, 1);
  dataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  dataPreReqs.add_element_field(*massFlowRate_, AlgTraits::numScsIp_);
  dataPreReqs.add_master_element_call(SCS_AREAV, CURRENT_COORDINATES);
  dataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_GRAD_OP : SCS_GRAD_OP, CURRENT_COORDINATES);
}
template <typename AlgTraits>
void MomentumUpwAdvDiffElemKernel<AlgTraits>::setup(const TimeIntegrator &)
{
  alpha_ = solnOpts_.get_alpha_factor(dofName_);
  alphaUpw_ = solnOpts_.get_alpha_upw_factor(dofName_);
  hoUpwind_ = solnOpts_.get_upw_factor(dofName_);
  useLimiter_ = solnOpts_.primitive_uses_limiter(dofName_);
  om_alpha_ = 1.0 - alpha_;
  om_alphaUpw_ = 1.0 - alphaUpw_;
}
template <typename AlgTraits>
bool MomentumUpwAdvDiffElemKernel<AlgTraits>::device_descriptor(nalu_b200_kernel_desc &d)
{
  clear(d);
  d.kind = NALU_K_MOMENTUM_UPW_ADV_DIFF;
  d.options = (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0) | (skewSymmetric_ ? NALU_OPT_SKEW_SYMMETRIC : 0);
  d.include_divu = includeDivU_;
  upw_fill(d, pecletFunction_, alpha_, alphaUpw_, hoUpwind_, useLimiter_);
  bind(d, NALU_F_COORDINATES, coordinates_);
  bind(d, NALU_F_VELOCITY, velocityNp1_);
  if (velocityRTM_ != velocityNp1_) bind(d, NALU_F_VELOCITY_RTM, velocityRTM_);
  bind(d, NALU_F_VISCOSITY, viscosity_);
  bind(d, NALU_F_DENSITY, density_);
  bind(d, NALU_F_GJU, Gju_);
  d.mass_flow_rate = massFlowRate_->device_data();
  return true;
}

NALU_B200_DECLARE_KERNEL(TurbKineticEnergySSTSrcElemKernel)
NALU_B200_DECLARE_KERNEL(TurbKineticEnergySSTDESSrcElemKernel)
NALU_B200_DECLARE_KERNEL(TurbKineticEnergyKsgsSrcElemKernel)
NALU_B200_DECLARE_KERNEL(MomentumNSOElemKernel)
NALU_B200_DECLARE_KERNEL(ScalarUpwAdvDiffElemKernel)
NALU_B200_DECLARE_KERNEL(MomentumUpwAdvDiffElemKernel)
NALU_B200_DECLARE_KERNEL(SpecificDissipationRateSSTSrcElemKernel)
NALU_B200_DECLARE_KERNEL(MomentumBuoyancySrcElemKernel)
NALU_B200_DECLARE_KERNEL(MomentumBuoyancyBoussinesqSrcElemKernel)
NALU_B200_DECLARE_KERNEL(MomentumAdvDiffElemKernel)
NALU_B200_DECLARE_KERNEL(ContinuityAdvElemKernel)
NALU_B200_DECLARE_KERNEL(ScalarAdvDiffElemKernel)
NALU_B200_DECLARE_KERNEL(ScalarDiffElemKernel)
NALU_B200_DECLARE_KERNEL(MomentumMassElemKernel)
NALU_B200_DECLARE_KERNEL(ContinuityMassElemKernel)
NALU_B200_DECLARE_KERNEL(ScalarMassElemKernel)
// Hex27 (BASELINE config 5): the momentum list and the one-dof kernels assemble_hex27.cu serves
template class MomentumAdvDiffElemKernel<AlgTraitsHex27>;
template class MomentumMassElemKernel<AlgTraitsHex27>;
template class ContinuityAdvElemKernel<AlgTraitsHex27>;
template class ScalarAdvDiffElemKernel<AlgTraitsHex27>;


// ================================================================================================
// Face/element boundary-condition kernels and AssembleFaceElemSolverAlgorithm
// ================================================================================================
void FaceElemKernel::clear(nalu_b200_face_kernel_desc &d) { std::memset(&d, 0, sizeof(d)); d.dt = 1.0; d.gamma1 = 1.0; }
void FaceElemKernel::bind(nalu_b200_face_kernel_desc &d, int slot, stk::mesh::FieldBase *f)
{
  if (!f) throw std::runtime_error("face kernel field not found in MetaData");
  d.nodal[slot] = f->device_data();
}
void FaceElemKernel::bind_face(nalu_b200_face_kernel_desc &d, int slot, stk::mesh::FieldBase *f, size_t first_face)
{
  if (!f) throw std::runtime_error("face kernel face field not found in MetaData");
  d.face[slot] = f->device_data() + first_face * f->max_size();
}
namespace {
void peclet_fill(nalu_b200_face_kernel_desc &d, const PecletFunctionSpec &p, double alphaUpw, double hoUpwind)
{
  d.upw[NALU_UPW_ALPHA_UPW] = alphaUpw;
  d.upw[NALU_UPW_HO_UPWIND] = hoUpwind;
  d.upw[NALU_UPW_HYBRID_FACTOR] = p.hybridFactor;
  d.upw[NALU_UPW_TANH_TRANS] = p.c1;
  d.upw[NALU_UPW_TANH_WIDTH] = p.c2;
  if (p.tanh_form) d.options |= NALU_OPT_PECLET_TANH;
}
template <typename T>
T *need(T *f, const char *kernel, const char *field)
{
  STK_ThrowRequireMsg(f, kernel << ": field " << field << " not registered");
  return f;
}
} // namespace

// ---- src/kernel/MomentumSymmetryElemKernel.C:23-63
template <typename BcAlgTraits>
MomentumSymmetryElemKernel<BcAlgTraits>::MomentumSymmetryElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts,
                                                                   VectorFieldType *velocity, ScalarFieldType *viscosity,
                                                                   ElemDataRequests &faceDataPreReqs, ElemDataRequests &elemDataPreReqs)
    : viscosity_(viscosity), includeDivU_(solnOpts.includeDivU_), shiftedGradOp_(solnOpts.get_shifted_grad_op(velocity->name())),
      meSCS_(MasterElementRepo::get_surface_master_element(BcAlgTraits::elemTopo_))
{
  velocityNp1_ = &(velocity->field_of_state(stk::mesh::StateNP1));
  coordinates_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name()), name(), "coordinates");
  exposedAreaVec_ = need(metaData.get_field<double>(metaData.side_rank(), "exposed_area_vector"), name(), "exposed_area_vector");
  MasterElement *meFC = MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_);
  faceDataPreReqs.add_cvfem_face_me(meFC);
  elemDataPreReqs.add_cvfem_surface_me(meSCS_);
  faceDataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  faceDataPreReqs.add_face_field(*exposedAreaVec_, BcAlgTraits::numFaceIp_, BcAlgTraits::nDim_);
  elemDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  elemDataPreReqs.add_gathered_nodal_field(*velocityNp1_, BcAlgTraits::nDim_);
  elemDataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_FACE_GRAD_OP : SCS_FACE_GRAD_OP, CURRENT_COORDINATES);
  get_face_shape_fn_data<BcAlgTraits>([&](double *ptr) { meFC->shape_fcn(ptr); }, vf_shape_function_);
}
template <typename BcAlgTraits>
bool MomentumSymmetryElemKernel<BcAlgTraits>::device_face_descriptor(nalu_b200_face_kernel_desc &d, size_t first_face)
{
  clear(d);
  d.kind = NALU_FK_MOMENTUM_SYMMETRY;
  d.options = shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0;
  d.include_divu = includeDivU_;
  bind(d, NALU_FN_COORDINATES, coordinates_);
  bind(d, NALU_FN_VELOCITY, velocityNp1_);
  bind(d, NALU_FN_VISCOSITY, viscosity_);
  bind_face(d, NALU_FF_EXPOSED_AREA_VEC, exposedAreaVec_, first_face);
  return true;
}

// ---- src/kernel/MomentumOpenAdvDiffElemKernel.C:26-100
template <typename BcAlgTraits>
MomentumOpenAdvDiffElemKernel<BcAlgTraits>::MomentumOpenAdvDiffElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts,
                                                                         EquationSystem *eqSystem, VectorFieldType *velocity,
                                                                         GenericFieldType *Gjui, ScalarFieldType *viscosity,
                                                                         ElemDataRequests &faceDataPreReqs,
                                                                         ElemDataRequests &elemDataPreReqs)
    : viscosity_(viscosity), Gjui_(Gjui), alphaUpw_(solnOpts.get_alpha_upw_factor("velocity")), om_alphaUpw_(1.0 - alphaUpw_),
      hoUpwind_(solnOpts.get_upw_factor("velocity")), includeDivU_(solnOpts.includeDivU_),
      meshVelocityCorrection_(solnOpts.does_mesh_move() ? 1.0 : 0.0), shiftedGradOp_(solnOpts.get_shifted_grad_op(velocity->name())),
      faceIpNodeMap_(MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_)->ipNodeMap()),
      meSCS_(MasterElementRepo::get_surface_master_element(BcAlgTraits::elemTopo_)),
      pecletFunction_(make_peclet_spec(eqSystem, velocity->name()))
{
  velocityNp1_ = &(velocity->field_of_state(stk::mesh::StateNP1));
  if (solnOpts.does_mesh_move()) {
    velocityRTM_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "velocity_rtm"), name(), "velocity_rtm");
    meshVelocity_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "mesh_velocity"), name(), "mesh_velocity");
  } else {
    velocityRTM_ = velocityNp1_;
    meshVelocity_ = velocityNp1_;
  }
  coordinates_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name()), name(), "coordinates");
  density_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "density"), name(), "density");
  exposedAreaVec_ = need(metaData.get_field<double>(metaData.side_rank(), "exposed_area_vector"), name(), "exposed_area_vector");
  openMassFlowRate_ = need(metaData.get_field<double>(metaData.side_rank(), "open_mass_flow_rate"), name(), "open_mass_flow_rate");
  velocityBc_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "open_velocity_bc"), name(), "open_velocity_bc");
  MasterElement *meFC = MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_);
  faceDataPreReqs.add_cvfem_face_me(meFC);
  elemDataPreReqs.add_cvfem_surface_me(meSCS_);
  faceDataPreReqs.add_gathered_nodal_field(*density_, 1);
  faceDataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  faceDataPreReqs.add_gathered_nodal_field(*velocityNp1_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_gathered_nodal_field(*meshVelocity_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_gathered_nodal_field(*velocityBc_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_gathered_nodal_field(*Gjui_, BcAlgTraits::nDim_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  faceDataPreReqs.add_face_field(*exposedAreaVec_, BcAlgTraits::numFaceIp_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_face_field(*openMassFlowRate_, BcAlgTraits::numFaceIp_);
  elemDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  elemDataPreReqs.add_gathered_nodal_field(*velocityNp1_, BcAlgTraits::nDim_);
  elemDataPreReqs.add_gathered_nodal_field(*velocityRTM_, BcAlgTraits::nDim_);
  elemDataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  elemDataPreReqs.add_gathered_nodal_field(*density_, 1);
  elemDataPreReqs.add_master_element_call(shiftedGradOp_ ? SCS_SHIFTED_FACE_GRAD_OP : SCS_FACE_GRAD_OP, CURRENT_COORDINATES);
  get_face_shape_fn_data<BcAlgTraits>([&](double *ptr) { meFC->shape_fcn(ptr); }, vf_shape_function_);
  skewSymmetric_ = solnOpts.get_skew_symmetric(velocity->name());
  get_face_shape_fn_data<BcAlgTraits>([&](double *ptr) { skewSymmetric_ ? meFC->shifted_shape_fcn(ptr) : meFC->shape_fcn(ptr); },
                                      vf_adv_shape_function_);
}
template <typename BcAlgTraits>
bool MomentumOpenAdvDiffElemKernel<BcAlgTraits>::device_face_descriptor(nalu_b200_face_kernel_desc &d, size_t first_face)
{
  clear(d);
  d.kind = NALU_FK_MOMENTUM_OPEN_ADV_DIFF;
  d.options = (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0) | (skewSymmetric_ ? NALU_OPT_SKEW_SYMMETRIC : 0) |
              (meshVelocityCorrection_ != 0.0 ? NALU_OPT_MESH_MOVES : 0);
  d.include_divu = includeDivU_;
  peclet_fill(d, pecletFunction_, alphaUpw_, hoUpwind_);
  bind(d, NALU_FN_COORDINATES, coordinates_);
  bind(d, NALU_FN_VELOCITY, velocityNp1_);
  if (meshVelocityCorrection_ != 0.0) {
    bind(d, NALU_FN_VELOCITY_RTM, velocityRTM_);
    bind(d, NALU_FN_MESH_VELOCITY, meshVelocity_);
  }
  bind(d, NALU_FN_VISCOSITY, viscosity_);
  bind(d, NALU_FN_DENSITY, density_);
  bind(d, NALU_FN_GJUI, Gjui_);
  bind(d, NALU_FN_VELOCITY_BC, velocityBc_);
  bind_face(d, NALU_FF_EXPOSED_AREA_VEC, exposedAreaVec_, first_face);
  bind_face(d, NALU_FF_OPEN_MASS_FLOW_RATE, openMassFlowRate_, first_face);
  return true;
}

// ---- src/kernel/ScalarOpenAdvElemKernel.C:25-82
template <typename BcAlgTraits>
ScalarOpenAdvElemKernel<BcAlgTraits>::ScalarOpenAdvElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts,
                                                             EquationSystem *eqSystem, ScalarFieldType *scalarQ,
                                                             ScalarFieldType *bcScalarQ, VectorFieldType *Gjq,
                                                             ScalarFieldType *diffFluxCoeff, ElemDataRequests &faceDataPreReqs,
                                                             ElemDataRequests &elemDataPreReqs)
    : scalarQ_(scalarQ), bcScalarQ_(bcScalarQ), Gjq_(Gjq), diffFluxCoeff_(diffFluxCoeff),
      alphaUpw_(solnOpts.get_alpha_upw_factor(scalarQ->name())), om_alphaUpw_(1.0 - alphaUpw_),
      hoUpwind_(solnOpts.get_upw_factor(scalarQ->name())),
      faceIpNodeMap_(MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_)->ipNodeMap()),
      meSCS_(MasterElementRepo::get_surface_master_element(BcAlgTraits::elemTopo_)),
      pecletFunction_(make_peclet_spec(eqSystem, scalarQ->name()))
{
  velocityRTM_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.does_mesh_move() ? "velocity_rtm" : "velocity"),
                      name(), "velocity");
  coordinates_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name()), name(), "coordinates");
  density_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "density"), name(), "density");
  openMassFlowRate_ = need(metaData.get_field<double>(metaData.side_rank(), "open_mass_flow_rate"), name(), "open_mass_flow_rate");
  MasterElement *meFC = MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_);
  faceDataPreReqs.add_cvfem_face_me(meFC);
  elemDataPreReqs.add_cvfem_surface_me(meSCS_);
  faceDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  faceDataPreReqs.add_gathered_nodal_field(*Gjq_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_gathered_nodal_field(*scalarQ_, 1);
  faceDataPreReqs.add_gathered_nodal_field(*bcScalarQ_, 1);
  faceDataPreReqs.add_face_field(*openMassFlowRate_, BcAlgTraits::numFaceIp_);
  elemDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  elemDataPreReqs.add_gathered_nodal_field(*velocityRTM_, BcAlgTraits::nDim_);
  elemDataPreReqs.add_gathered_nodal_field(*diffFluxCoeff_, 1);
  elemDataPreReqs.add_gathered_nodal_field(*density_, 1);
  skewSymmetric_ = solnOpts.get_skew_symmetric(scalarQ_->name());
  get_face_shape_fn_data<BcAlgTraits>([&](double *ptr) { skewSymmetric_ ? meFC->shifted_shape_fcn(ptr) : meFC->shape_fcn(ptr); },
                                      vf_adv_shape_function_);
}
template <typename BcAlgTraits>
bool ScalarOpenAdvElemKernel<BcAlgTraits>::device_face_descriptor(nalu_b200_face_kernel_desc &d, size_t first_face)
{
  clear(d);
  d.kind = NALU_FK_SCALAR_OPEN_ADV;
  d.options = skewSymmetric_ ? NALU_OPT_SKEW_SYMMETRIC : 0;
  peclet_fill(d, pecletFunction_, alphaUpw_, hoUpwind_);
  bind(d, NALU_FN_COORDINATES, coordinates_);
  bind(d, NALU_FN_VELOCITY, velocityRTM_);
  bind(d, NALU_FN_VISCOSITY, diffFluxCoeff_);
  bind(d, NALU_FN_DENSITY, density_);
  bind(d, NALU_FN_SCALARQ, scalarQ_);
  bind(d, NALU_FN_BC_SCALARQ, bcScalarQ_);
  bind(d, NALU_FN_GJQ, Gjq_);
  bind_face(d, NALU_FF_OPEN_MASS_FLOW_RATE, openMassFlowRate_, first_face);
  return true;
}

// ---- src/kernel/ContinuityOpenElemKernel.C:25-97
template <typename BcAlgTraits>
ContinuityOpenElemKernel<BcAlgTraits>::ContinuityOpenElemKernel(const stk::mesh::MetaData &metaData, const SolutionOptions &solnOpts,
                                                               ElemDataRequests &faceDataPreReqs, ElemDataRequests &elemDataPreReqs)
    : shiftedGradOp_(solnOpts.get_shifted_grad_op("pressure")), reducedSensitivities_(solnOpts.cvfemReducedSensPoisson_),
      shiftMdot_(solnOpts.cvfemShiftMdot_), meSCS_(MasterElementRepo::get_surface_master_element(BcAlgTraits::elemTopo_))
{
  velocityRTM_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.does_mesh_move() ? "velocity_rtm" : "velocity"),
                      name(), "velocity");
  Gpdx_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "dpdx"), name(), "dpdx");
  coordinates_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name()), name(), "coordinates");
  pressure_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "pressure"), name(), "pressure");
  pressureBc_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "pressure_bc"), name(), "pressure_bc");
  density_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "density"), name(), "density");
  exposedAreaVec_ = need(metaData.get_field<double>(metaData.side_rank(), "exposed_area_vector"), name(), "exposed_area_vector");
  dynamicPressure_ = need(metaData.get_field<double>(metaData.side_rank(), "dynamic_pressure"), name(), "dynamic_pressure");
  MasterElement *meFC = MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_);
  faceDataPreReqs.add_cvfem_face_me(meFC);
  elemDataPreReqs.add_cvfem_surface_me(meSCS_);
  faceDataPreReqs.add_gathered_nodal_field(*pressure_, 1);
  faceDataPreReqs.add_gathered_nodal_field(*pressureBc_, 1);
  faceDataPreReqs.add_gathered_nodal_field(*density_, 1);
  faceDataPreReqs.add_gathered_nodal_field(*velocityRTM_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_gathered_nodal_field(*Gpdx_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_face_field(*exposedAreaVec_, BcAlgTraits::numFaceIp_, BcAlgTraits::nDim_);
  faceDataPreReqs.add_face_field(*dynamicPressure_, BcAlgTraits::numFaceIp_);
  faceDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  elemDataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  elemDataPreReqs.add_gathered_nodal_field(*pressure_, 1);
  if (!shiftedGradOp_ || !reducedSensitivities_) elemDataPreReqs.add_master_element_call(SCS_FACE_GRAD_OP, CURRENT_COORDINATES);
  if (shiftedGradOp_ || reducedSensitivities_) elemDataPreReqs.add_master_element_call(SCS_SHIFTED_FACE_GRAD_OP, CURRENT_COORDINATES);
  get_face_shape_fn_data<BcAlgTraits>([&](double *ptr) { shiftMdot_ ? meFC->shifted_shape_fcn(ptr) : meFC->shape_fcn(ptr); },
                                      vf_shape_function_);
}
template <typename BcAlgTraits>
void ContinuityOpenElemKernel<BcAlgTraits>::setup(const TimeIntegrator &timeIntegrator)
{
  dt_ = timeIntegrator.get_time_step();
  gamma1_ = timeIntegrator.get_gamma1();
}
template <typename BcAlgTraits>
bool ContinuityOpenElemKernel<BcAlgTraits>::device_face_descriptor(nalu_b200_face_kernel_desc &d, size_t first_face)
{
  clear(d);
  d.kind = NALU_FK_CONTINUITY_OPEN;
  d.options = (shiftedGradOp_ ? NALU_OPT_SHIFTED_GRAD_OP : 0) | (reducedSensitivities_ ? NALU_OPT_REDUCED_SENS : 0) |
              (shiftMdot_ ? NALU_OPT_SHIFT_MDOT : 0);
  d.dt = dt_;
  d.gamma1 = gamma1_;
  bind(d, NALU_FN_COORDINATES, coordinates_);
  bind(d, NALU_FN_VELOCITY, velocityRTM_);
  bind(d, NALU_FN_DENSITY, density_);
  bind(d, NALU_FN_PRESSURE, pressure_);
  bind(d, NALU_FN_PRESSURE_BC, pressureBc_);
  bind(d, NALU_FN_DPDX, Gpdx_);
  bind_face(d, NALU_FF_EXPOSED_AREA_VEC, exposedAreaVec_, first_face);
  bind_face(d, NALU_FF_DYNAMIC_PRESSURE, dynamicPressure_, first_face);
  return true;
}
// ---- src/kernel/MomentumWallFunctionElemKernel.C:26-66
template <typename BcAlgTraits>
MomentumWallFunctionElemKernel<BcAlgTraits>::MomentumWallFunctionElemKernel(const stk::mesh::BulkData &bulkData, const SolutionOptions &solnOpts,
                                                                           ElemDataRequests &dataPreReqs)
    : elog_(solnOpts.get_turb_model_constant(TM_elog)), kappa_(solnOpts.get_turb_model_constant(TM_kappa)),
      yplusCrit_(solnOpts.get_turb_model_constant(TM_yplus_crit)),
      ipNodeMap_(MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_)->ipNodeMap())
{
  const stk::mesh::MetaData &metaData = bulkData.mesh_meta_data();
  VectorFieldType *velocity = need(metaData.get_field<double>(stk::topology::NODE_RANK, "velocity"), name(), "velocity");
  velocityNp1_ = &(velocity->field_of_state(stk::mesh::StateNP1));
  bcVelocity_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "wall_velocity_bc"), name(), "wall_velocity_bc");
  density_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "density"), name(), "density");
  viscosity_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, "viscosity"), name(), "viscosity");
  exposedAreaVec_ = need(metaData.get_field<double>(metaData.side_rank(), "exposed_area_vector"), name(), "exposed_area_vector");
  wallFrictionVelocityBip_ =
      need(metaData.get_field<double>(metaData.side_rank(), "wall_friction_velocity_bip"), name(), "wall_friction_velocity_bip");
  wallNormalDistanceBip_ =
      need(metaData.get_field<double>(metaData.side_rank(), "wall_normal_distance_bip"), name(), "wall_normal_distance_bip");
  coordinates_ = need(metaData.get_field<double>(stk::topology::NODE_RANK, solnOpts.get_coordinates_name()), name(), "coordinates");
  MasterElement *meFC = MasterElementRepo::get_surface_master_element(BcAlgTraits::faceTopo_);
  get_face_shape_fn_data<BcAlgTraits>([&](double *ptr) { meFC->shape_fcn(ptr); }, vf_shape_function_);
  dataPreReqs.add_cvfem_face_me(meFC);
  dataPreReqs.add_coordinates_field(*coordinates_, BcAlgTraits::nDim_, CURRENT_COORDINATES);
  dataPreReqs.add_gathered_nodal_field(*velocityNp1_, BcAlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*bcVelocity_, BcAlgTraits::nDim_);
  dataPreReqs.add_gathered_nodal_field(*density_, 1);
  dataPreReqs.add_gathered_nodal_field(*viscosity_, 1);
  dataPreReqs.add_face_field(*exposedAreaVec_, BcAlgTraits::numFaceIp_, BcAlgTraits::nDim_);
  dataPreReqs.add_face_field(*wallFrictionVelocityBip_, BcAlgTraits::numFaceIp_);
  dataPreReqs.add_face_field(*wallNormalDistanceBip_, BcAlgTraits::numFaceIp_);
}
template <typename BcAlgTraits>
bool MomentumWallFunctionElemKernel<BcAlgTraits>::device_face_descriptor(nalu_b200_face_kernel_desc &d, size_t first_face)
{
  clear(d);
  d.kind = NALU_FK_MOMENTUM_WALL_FUNCTION;
  d.wall[NALU_WALL_ELOG] = elog_;
  d.wall[NALU_WALL_KAPPA] = kappa_;
  d.wall[NALU_WALL_YPLUS_CRIT] = yplusCrit_;
  bind(d, NALU_FN_COORDINATES, coordinates_);
  bind(d, NALU_FN_VELOCITY, velocityNp1_);
  bind(d, NALU_FN_WALL_VELOCITY_BC, bcVelocity_);
  bind(d, NALU_FN_DENSITY, density_);
  bind(d, NALU_FN_VISCOSITY, viscosity_);
  bind_face(d, NALU_FF_EXPOSED_AREA_VEC, exposedAreaVec_, first_face);
  bind_face(d, NALU_FF_WALL_FRICTION_VELOCITY_BIP, wallFrictionVelocityBip_, first_face);
  bind_face(d, NALU_FF_WALL_NORMAL_DISTANCE_BIP, wallNormalDistanceBip_, first_face);
  return true;
}
template class MomentumWallFunctionElemKernel<AlgTraitsQuad4Hex8>;
template class MomentumSymmetryElemKernel<AlgTraitsQuad4Hex8>;
template class MomentumOpenAdvDiffElemKernel<AlgTraitsQuad4Hex8>;
template class ScalarOpenAdvElemKernel<AlgTraitsQuad4Hex8>;
template class ContinuityOpenElemKernel<AlgTraitsQuad4Hex8>;

// ---- src/AssembleFaceElemSolverAlgorithm.C:48-108
AssembleFaceElemSolverAlgorithm::AssembleFaceElemSolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem,
                                                                 unsigned nodesPerFace, unsigned nodesPerElem, bool interleaveMEViews)
    : SolverAlgorithm(realm, part, eqSystem), numDof_(eqSystem->linsys_->numDof()), nodesPerFace_(nodesPerFace),
      nodesPerElem_(nodesPerElem), rhsSize_(nodesPerFace * eqSystem->linsys_->numDof()), interleaveMEViews_(interleaveMEViews)
{
}
AssembleFaceElemSolverAlgorithm::~AssembleFaceElemSolverAlgorithm()
{
  for (nalu_b200_face_set *fs : faceSets_) nalu_b200_face_set_destroy(fs);
}
void AssembleFaceElemSolverAlgorithm::initialize_connectivity() { eqSystem_->linsys_->buildFaceElemToNodeGraph(partVec_); }
void AssembleFaceElemSolverAlgorithm::execute()
{
  for (Kernel *k : activeKernels_) k->setup(*realm_.timeIntegrator_);
  DeviceCsrLinearSystem *ls = dynamic_cast<DeviceCsrLinearSystem *>(eqSystem_->linsys_);
  if (!ls) throw std::runtime_error("AssembleFaceElemSolverAlgorithm: the device path needs a DeviceCsrLinearSystem");
  STK_ThrowRequireMsg(nodesPerFace_ == 4 && nodesPerElem_ == 8,
                      "AssembleFaceElemSolverAlgorithm expected nodesPerEntity_ = 4 (Quad4 faces of Hex8 elements), got " << nodesPerFace_);
  if (activeKernels_.empty()) return;
  if (faceSets_.empty()) {
    STK_ThrowRequireMsg(ls->graph(), "AssembleFaceElemSolverAlgorithm::execute before finalizeLinearSystem");
    for (stk::mesh::Part *part : partVec_) {
      nalu_b200_face_set *fs = nullptr;
      check(nalu_b200_face_set_create(&fs, ls->graph(), int64_t(part->face_elems().size()), part->face_elems().data(),
                                      part->face_ordinals().data()),
            "face_set_create");
      faceSets_.push_back(fs);
    }
  }
  for (size_t ip = 0; ip < partVec_.size(); ++ip) {
    std::vector<nalu_b200_face_kernel_desc> descs(activeKernels_.size());
    for (size_t i = 0; i < activeKernels_.size(); ++i) {
      FaceElemKernel *fk = dynamic_cast<FaceElemKernel *>(activeKernels_[i]);
      if (!fk || !fk->device_face_descriptor(descs[i], partVec_[ip]->first_face()))
        throw std::runtime_error(std::string("kernel ") + activeKernels_[i]->name() +
                                 " has no face/element device form; this path has no CPU fallback");
    }
    ls->assembleFaceElemDevice(faceSets_[ip], descs.data(), int(descs.size()));
  }
}

// ================================================================================================
// Algorithms
// ================================================================================================
Algorithm::~Algorithm()
{
  for (Kernel *k : activeKernels_) delete k;
}
EquationSystem::~EquationSystem() { delete linsys_; }

AssembleElemSolverAlgorithm::AssembleElemSolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem,
                                                         stk::mesh::EntityRank entityRank, unsigned nodesPerEntity,
                                                         bool interleaveMEViews)
    : SolverAlgorithm(realm, part, eqSystem), entityRank_(entityRank), nodesPerEntity_(nodesPerEntity),
      rhsSize_(nodesPerEntity * eqSystem->linsys_->numDof()), interleaveMEViews_(interleaveMEViews)
{
}
void AssembleElemSolverAlgorithm::initialize_connectivity() { eqSystem_->linsys_->buildElemToNodeGraph(partVec_); }

void AssembleElemSolverAlgorithm::execute()
{
  // set any data
  for (Kernel *k : activeKernels_) k->setup(*realm_.timeIntegrator_);
  DeviceCsrLinearSystem *ls = dynamic_cast<DeviceCsrLinearSystem *>(eqSystem_->linsys_);
  if (!ls) throw std::runtime_error("AssembleElemSolverAlgorithm: the device path needs a DeviceCsrLinearSystem");
  STK_ThrowRequireMsg((nodesPerEntity_ == 8 || nodesPerEntity_ == 27) && entityRank_ == stk::topology::ELEM_RANK &&
                          nodesPerEntity_ == unsigned(realm_.bulk_data().nodes_per_elem()),
                      "AssembleElemSolverAlgorithm expected nodesPerEntity_ = 8 (Hex8) or 27 (Hex27) matching the mesh, got " << nodesPerEntity_);
  std::vector<nalu_b200_kernel_desc> descs(activeKernels_.size());
  for (size_t i = 0; i < activeKernels_.size(); ++i)
    if (!activeKernels_[i]->device_descriptor(descs[i]))
      throw std::runtime_error(std::string("kernel ") + activeKernels_[i]->name() + " has no device form; this path has no CPU fallback");
  if (descs.empty()) return;
  ls->assembleElemDevice(descs.data(), int(descs.size()), firstAlgorithmAfterZero_);
}

AssembleNodeSolverAlgorithm::AssembleNodeSolverAlgorithm(Realm &realm, stk::mesh::Part *part, EquationSystem *eqSystem)
    : SolverAlgorithm(realm, part, eqSystem)
{
}
void AssembleNodeSolverAlgorithm::initialize_connectivity() { eqSystem_->linsys_->buildNodeGraph(partVec_); }
void AssembleNodeSolverAlgorithm::execute()
{
  DeviceCsrLinearSystem *ls = dynamic_cast<DeviceCsrLinearSystem *>(eqSystem_->linsys_);
  if (!ls) throw std::runtime_error("AssembleNodeSolverAlgorithm: the device path needs a DeviceCsrLinearSystem");
  const stk::mesh::MetaData &meta = realm_.meta_data();
  const TimeIntegrator &ti = *realm_.timeIntegrator_;
  nalu_b200_kernel_desc d;
  std::memset(&d, 0, sizeof(d));
  d.kind = massKind_;
  d.om_dens_fac = 1.0;
  d.dt = ti.get_time_step(); d.gamma1 = ti.get_gamma1(); d.gamma2 = ti.get_gamma2(); d.gamma3 = ti.get_gamma3();
  auto get = [&](const char *name) {
    stk::mesh::Field<double> *f = meta.get_field<double>(stk::topology::NODE_RANK, name);
    if (!f) throw std::runtime_error(std::string("node mass algorithm: field ") + name + " not registered");
    return f;
  };
  auto states = [&](stk::mesh::Field<double> *f, int np1, int n, int nm1) {
    stk::mesh::Field<double> *a, *b, *c;
    pick_states(f, a, b, c);
    d.nodal[np1] = c->device_data();
    d.nodal[n] = b->device_data();
    d.nodal[nm1] = a->device_data();
  };
  states(get("density"), NALU_F_DENSITY, NALU_F_DENSITY_N, NALU_F_DENSITY_NM1);
  d.nodal[NALU_F_DUAL_VOLUME] = get("dual_nodal_volume")->device_data();
  if (massKind_ == NALU_K_MOMENTUM_MASS) {
    states(get("velocity"), NALU_F_VELOCITY, NALU_F_VELOCITY_N, NALU_F_VELOCITY_NM1);
    d.nodal[NALU_F_DPDX] = get("dpdx")->device_data();
  } else if (massKind_ == NALU_K_SCALAR_MASS) {
    if (!scalarQ_) throw std::runtime_error("scalar node mass algorithm: scalarQ_ not set");
    states(static_cast<stk::mesh::Field<double> *>(scalarQ_), NALU_F_SCALARQ, NALU_F_SCALARQ_N, NALU_F_SCALARQ_NM1);
  }
  check(nalu_b200_assemble_node_mass(ls->graph(), &d, ls->deviceValues(), ls->deviceRhs(), nullptr), "assemble_node_mass");
}

// ================================================================================================
// LinearSystem
// ================================================================================================
LinearSystem::LinearSystem(Realm &realm, const unsigned numDof, EquationSystem *eqSys, LinearSolver *linearSolver)
    : realm_(realm), eqSys_(eqSys), inConstruction_(false), numDof_(numDof), eqSysName_(eqSys ? eqSys->name_ : ""),
      linearSolver_(linearSolver)
{
}
LinearSystem *LinearSystem::create(Realm &realm, const unsigned numDof, EquationSystem *eqSys, LinearSolver *linearSolver)
{
  // src/LinearSystem.C:85-97 switches on the solver type; this build has exactly one
  return new DeviceCsrLinearSystem(realm, numDof, eqSys, linearSolver);
}

DeviceCsrLinearSystem::DeviceCsrLinearSystem(Realm &realm, const unsigned numDof, EquationSystem *eqSys, LinearSolver *linearSolver)
    : LinearSystem(realm, numDof, eqSys, linearSolver)
{
  STK_ThrowRequireMsg(numDof == 1 || numDof == 3, "DeviceCsrLinearSystem: numDof must be 1 or 3");
}
DeviceCsrLinearSystem::~DeviceCsrLinearSystem()
{
  if (d_values_) nalu_b200_free(d_values_);
  if (d_rhs_) nalu_b200_free(d_rhs_);
  for (double *p : sendBufs_) nalu_b200_free(p);
  for (double *p : recvBufs_) nalu_b200_free(p);
  if (graph_) nalu_b200_graph_destroy(graph_);
}
void DeviceCsrLinearSystem::unsupported(const char *what) const
{
  throw std::runtime_error(std::string("DeviceCsrLinearSystem::") + what + " is outside the element-assembly path");
}
void DeviceCsrLinearSystem::beginLinearSystemConstruction()
{
  if (inConstruction_) return;
  inConstruction_ = true;
}
void DeviceCsrLinearSystem::buildNodeGraph(const stk::mesh::PartVector &)
{
  beginLinearSystemConstruction(); // every node is connected to itself by any element graph; nothing to add
}
void DeviceCsrLinearSystem::buildElemToNodeGraph(const stk::mesh::PartVector &)
{
  beginLinearSystemConstruction();
  elemGraphRequested_ = true;
}
void DeviceCsrLinearSystem::buildFaceElemToNodeGraph(const stk::mesh::PartVector &)
{
  // TpetraLinearSystem::buildFaceElemToNodeGraph (src/TpetraLinearSystem.C:487-521) connects the nodes of the element
  // attached to each face: a subset of the element graph, which this system always holds
  beginLinearSystemConstruction();
  elemGraphRequested_ = true;
}
void DeviceCsrLinearSystem::assembleFaceElemDevice(const nalu_b200_face_set *faces, const nalu_b200_face_kernel_desc *descs, int n)
{
  check(nalu_b200_assemble_face_elem(graph_, faces, descs, n, d_values_, d_rhs_, nullptr), "assemble_face_elem");
}
void DeviceCsrLinearSystem::finalizeLinearSystem()
{
  STK_ThrowRequireMsg(inConstruction_, "finalizeLinearSystem called before any build*Graph");
  STK_ThrowRequireMsg(elemGraphRequested_, "DeviceCsrLinearSystem needs buildElemToNodeGraph (node-only graphs are not on this path)");
  STK_ThrowRequireMsg(!graph_, "finalizeLinearSystem called twice");
  inConstruction_ = false;
  stk::mesh::BulkData &bulk = realm_.bulk_data();
  const int *box = bulk.box();
  if (bulk.nodes_per_elem() == 8)
    check(nalu_b200_graph_create(&graph_, bulk.parallel_rank(), bulk.parallel_size(), int(numDof_), int(bulk.num_nodes()),
                                 bulk.node_ids().data(), bulk.node_owners().data(), int64_t(bulk.num_elems()),
                                 bulk.connectivity().data(), box[0], box[1], box[2]),
          "graph_create");
  else // HEX_27 parts
    check(nalu_b200_graph_create_topo(&graph_, bulk.parallel_rank(), bulk.parallel_size(), int(numDof_), bulk.nodes_per_elem(),
                                      int(bulk.num_nodes()), bulk.node_ids().data(), bulk.node_owners().data(),
                                      int64_t(bulk.num_elems()), bulk.connectivity().data()),
          "graph_create_topo");
  if (bulk.parallel_size() > 1) {
    STK_ThrowRequireMsg(bool(transport_.exchange_graph), "multi-rank DeviceCsrLinearSystem needs a HaloTransport (setTransport)");
    std::map<int, std::vector<int64_t>> send;
    for (int p = 0; p < bulk.parallel_size(); ++p) {
      if (p == bulk.parallel_rank()) continue;
      int64_t nbytes = 0;
      check(nalu_b200_graph_send_size(graph_, p, &nbytes), "graph_send_size");
      if (!nbytes) continue;
      std::vector<int64_t> buf(size_t(nbytes / 8));
      check(nalu_b200_graph_send_data(graph_, p, buf.data()), "graph_send_data");
      send[p] = std::move(buf);
    }
    for (auto &kv : transport_.exchange_graph(send))
      check(nalu_b200_graph_recv_data(graph_, kv.first, kv.second.data(), int64_t(kv.second.size() * 8)), "graph_recv_data");
  }
  check(nalu_b200_graph_finalize(graph_), "graph_finalize");
  check(nalu_b200_graph_get_info(graph_, &info_), "graph_get_info");
  void *p = nullptr;
  check(nalu_b200_malloc(&p, size_t(info_.nnz_owned + info_.nnz_shared) * sizeof(double)), "values allocation");
  d_values_ = static_cast<double *>(p);
  check(nalu_b200_malloc(&p, size_t(info_.n_owned_rows + info_.n_shared_rows) * sizeof(double)), "rhs allocation");
  d_rhs_ = static_cast<double *>(p);
  node_row_lid_.resize(bulk.num_nodes());
  check(nalu_b200_graph_copy_node_lids(graph_, node_row_lid_.data(), nullptr), "copy_node_lids");
  zeroSystem();
}
void DeviceCsrLinearSystem::zeroSystem() { check(nalu_b200_zero_system(graph_, d_values_, d_rhs_, nullptr), "zero_system"); }

void DeviceCsrLinearSystem::assembleElemDevice(const nalu_b200_kernel_desc *descs, int n, bool overwrite)
{
  check(nalu_b200_assemble_elem(graph_, descs, n, d_values_, d_rhs_, NALU_ASSEMBLE_AUTO, overwrite ? 1 : 0, nullptr), "assemble_elem");
  last_family_ = nalu_b200_last_assemble_mode();
}

int DeviceCsrLinearSystem::entityToLID(stk::mesh::Entity node) const { return node_row_lid_[size_t(node.index)]; }

void DeviceCsrLinearSystem::sumInto(unsigned numEntities, const stk::mesh::Entity *entities, const double *rhs, const double *lhs,
                                    int *, int *, const char *)
{
  std::vector<int32_t> nodes(numEntities);
  for (unsigned i = 0; i < numEntities; ++i) nodes[i] = int32_t(entities[i].index);
  check(nalu_b200_sum_into(graph_, int(numEntities), nodes.data(), rhs, lhs, d_values_, d_rhs_, nullptr), "sum_into");
}
void DeviceCsrLinearSystem::sumInto(const std::vector<stk::mesh::Entity> &sym_meshobj, std::vector<int> &, std::vector<double> &,
                                    const std::vector<double> &rhs, const std::vector<double> &lhs, const char *tag)
{
  sumInto(unsigned(sym_meshobj.size()), sym_meshobj.data(), rhs.data(), lhs.data(), nullptr, nullptr, tag);
}

// TpetraLinearSystem::applyDirichletBCs (src/TpetraLinearSystem.C:1579-1653).  The selector of the reference --
// (locally owned | globally shared) & union(parts) & has the solution field & active -- is, on the flat mesh,
// the union of the parts' node lists; rows and residual entries are rewritten on the device.
void DeviceCsrLinearSystem::applyDirichletBCs(stk::mesh::FieldBase *solutionField, stk::mesh::FieldBase *bcValuesField,
                                              const stk::mesh::PartVector &parts, const unsigned beginPos, const unsigned endPos)
{
  STK_ThrowRequireMsg(solutionField && bcValuesField, "applyDirichletBCs: solution / bc fields required");
  STK_ThrowRequireMsg(solutionField->max_size() == numDof_ && bcValuesField->max_size() == numDof_,
                      "applyDirichletBCs: fieldSize == numDof_ (TpetraLinearSystem.C:1603)");
  const size_t nNodes = realm_.bulk_data().num_nodes();
  std::vector<char> in(nNodes, 0);
  for (const stk::mesh::Part *p : parts) {
    if (!p->has_explicit_nodes()) std::fill(in.begin(), in.end(), 1);
    else
      for (int32_t n : p->nodes()) {
        STK_ThrowRequireMsg(n >= 0 && size_t(n) < nNodes, "applyDirichletBCs: part node out of range");
        in[size_t(n)] = 1;
      }
  }
  std::vector<int32_t> nodes;
  for (size_t n = 0; n < nNodes; ++n)
    if (in[n]) nodes.push_back(int32_t(n));
  if (nodes.empty()) return;
  void *d_nodes = nullptr;
  check(nalu_b200_malloc(&d_nodes, nodes.size() * sizeof(int32_t)), "applyDirichletBCs node list");
  check(nalu_b200_memcpy_h2d(d_nodes, nodes.data(), nodes.size() * sizeof(int32_t), nullptr), "applyDirichletBCs node list copy");
  const int rc = nalu_b200_apply_dirichlet_bcs(graph_, int32_t(nodes.size()), static_cast<const int32_t *>(d_nodes),
                                               solutionField->device_data(), bcValuesField->device_data(), int32_t(beginPos),
                                               int32_t(endPos), d_values_, d_rhs_, nullptr);
  nalu_b200_stream_sync(nullptr);
  nalu_b200_free(d_nodes);
  check(rc, "apply_dirichlet_bcs");
}

// TpetraLinearSystem::resetRows (src/TpetraLinearSystem.C:1706-1755)
void DeviceCsrLinearSystem::resetRows(std::vector<stk::mesh::Entity> nodeList, const unsigned beginPos, const unsigned endPos)
{
  if (nodeList.empty()) return;
  std::vector<int32_t> nodes(nodeList.size());
  for (size_t i = 0; i < nodeList.size(); ++i) nodes[i] = int32_t(nodeList[i].index);
  void *d_nodes = nullptr;
  check(nalu_b200_malloc(&d_nodes, nodes.size() * sizeof(int32_t)), "resetRows node list");
  check(nalu_b200_memcpy_h2d(d_nodes, nodes.data(), nodes.size() * sizeof(int32_t), nullptr), "resetRows node list copy");
  const int rc = nalu_b200_reset_rows(graph_, int32_t(nodes.size()), static_cast<const int32_t *>(d_nodes), int32_t(beginPos),
                                      int32_t(endPos), d_values_, d_rhs_, nullptr);
  nalu_b200_stream_sync(nullptr);
  nalu_b200_free(d_nodes);
  check(rc, "reset_rows");
}

void DeviceCsrLinearSystem::loadComplete()
{
  int32_t ns = 0, nr = 0;
  check(nalu_b200_halo_num_peers(graph_, &ns, &nr), "halo_num_peers");
  if (ns == 0 && nr == 0) return;
  STK_ThrowRequireMsg(bool(transport_.exchange_rows), "loadComplete on more than one rank needs a HaloTransport");
  std::vector<int> sp(ns), rp(nr);
  std::vector<int64_t> sc(ns), rc(nr);
  for (int i = 0; i < ns; ++i) check(nalu_b200_halo_send_peer(graph_, i, &sp[i], &sc[i]), "halo_send_peer");
  for (int i = 0; i < nr; ++i) check(nalu_b200_halo_recv_peer(graph_, i, &rp[i], &rc[i]), "halo_recv_peer");
  if (sendBufs_.empty() && recvBufs_.empty()) {
    for (int i = 0; i < ns; ++i) {
      void *p = nullptr;
      check(nalu_b200_malloc(&p, size_t(sc[i]) * 8), "halo buffer");
      sendBufs_.push_back(static_cast<double *>(p));
    }
    for (int i = 0; i < nr; ++i) {
      void *p = nullptr;
      check(nalu_b200_malloc(&p, size_t(rc[i]) * 8), "halo buffer");
      recvBufs_.push_back(static_cast<double *>(p));
    }
  }
  for (int i = 0; i < ns; ++i) check(nalu_b200_halo_pack(graph_, i, d_values_, d_rhs_, sendBufs_[i], nullptr), "halo_pack");
  check(nalu_b200_stream_sync(nullptr), "stream sync");
  transport_.exchange_rows(sp, sendBufs_, sc, rp, recvBufs_, rc);
  for (int i = 0; i < nr; ++i) check(nalu_b200_halo_unpack_add(graph_, i, recvBufs_[i], d_values_, d_rhs_, nullptr), "halo_unpack_add");
}

std::vector<uint64_t> DeviceCsrLinearSystem::ownedRowPtr() const
{
  std::vector<uint64_t> v(size_t(info_.n_owned_rows) + 1);
  check(nalu_b200_graph_copy_rowptr(graph_, 0, v.data()), "copy_rowptr");
  return v;
}
std::vector<int32_t> DeviceCsrLinearSystem::ownedColIdx() const
{
  std::vector<int32_t> v(size_t(info_.nnz_owned));
  check(nalu_b200_graph_copy_colidx(graph_, 0, v.data()), "copy_colidx");
  return v;
}
std::vector<int64_t> DeviceCsrLinearSystem::ownedRowGids() const
{
  std::vector<int64_t> v(size_t(info_.n_owned_rows));
  check(nalu_b200_graph_copy_row_gids(graph_, 0, v.data()), "copy_row_gids");
  return v;
}
std::vector<int64_t> DeviceCsrLinearSystem::colGids() const
{
  std::vector<int64_t> v(size_t(info_.n_cols));
  check(nalu_b200_graph_copy_col_gids(graph_, v.data()), "copy_col_gids");
  return v;
}
std::vector<double> DeviceCsrLinearSystem::ownedValues() const
{
  std::vector<double> v(size_t(info_.nnz_owned));
  check(nalu_b200_memcpy_d2h(v.data(), d_values_, v.size() * 8, nullptr), "values device->host");
  check(nalu_b200_stream_sync(nullptr), "stream sync");
  return v;
}
std::vector<double> DeviceCsrLinearSystem::ownedRhs() const
{
  std::vector<double> v(size_t(info_.n_owned_rows));
  check(nalu_b200_memcpy_d2h(v.data(), d_rhs_, v.size() * 8, nullptr), "rhs device->host");
  check(nalu_b200_stream_sync(nullptr), "stream sync");
  return v;
}
void DeviceCsrLinearSystem::writeToFile(const char *filename, bool)
{
  // MatrixMarket coordinate format with global ids, like Tpetra::MatrixMarket::Writer output of the owned matrix
  const auto rp = ownedRowPtr();
  const auto ci = ownedColIdx();
  const auto rg = ownedRowGids();
  const auto cg = colGids();
  const auto v = ownedValues();
  FILE *f = std::fopen(filename, "w");
  if (!f) throw std::runtime_error(std::string("writeToFile: cannot open ") + filename);
  std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n");
  int64_t maxg = 0;
  for (int64_t g : cg) maxg = std::max(maxg, g);
  std::fprintf(f, "%lld %lld %zu\n", (long long)maxg, (long long)maxg, v.size());
  for (size_t r = 0; r + 1 < rp.size(); ++r)
    for (uint64_t q = rp[r]; q < rp[r + 1]; ++q)
      std::fprintf(f, "%lld %lld %.17g\n", (long long)rg[r], (long long)cg[size_t(ci[q])], v[q]);
  std::fclose(f);
}

} // namespace nalu
} // namespace sierra
