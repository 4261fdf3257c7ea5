// ScratchViews.h -- host-side mirror of the reference's per-element scratch: include/ScratchViews.h:50-175 (classes),
// :209-398 (master-element views), :584-704 (field views and byte counts), src/ScratchViews.C:408-484 (fill_pre_req_data),
// include/SharedMemData.h:23-54.
//
// On this path the element loop runs inside the CUDA kernels, which see flat nodal / element arrays and keep their
// scratch in registers and shared memory (nalu_b200/csrc/general_elem.cuh ElemScratch, box5_tile.cuh records); no
// ScratchViews object exists on the device.  The class is kept for host code written against the reference's API --
// a Kernel that wants to look at the gathered data of one element, a unit test, a CPU cross-check: same constructor,
// same lookups by field ordinal, same view shapes (LayoutRight), same gather semantics, master-element views filled by
// the same operators the device runs (csrc/hex8_dev.cuh compiled for the host).  Kokkos' team scratch is replaced by
// views that own their storage; TeamHandleType is an empty tag.
#pragma once
#include "ElemDataRequests.h"
#include "master_element/MasterElement.h"
#include <set>
#include <vector>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define NALU_SV_UNDEF_HOST
#endif
#ifndef __device__
#define __device__
#define NALU_SV_UNDEF_DEVICE
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#define NALU_SV_UNDEF_FORCEINLINE
#endif
#endif
#include "../../csrc/hex8_dev.cuh"
#ifdef NALU_SV_UNDEF_HOST
#undef __host__
#undef NALU_SV_UNDEF_HOST
#endif
#ifdef NALU_SV_UNDEF_DEVICE
#undef __device__
#undef NALU_SV_UNDEF_DEVICE
#endif
#ifdef NALU_SV_UNDEF_FORCEINLINE
#undef __forceinline__
#undef NALU_SV_UNDEF_FORCEINLINE
#endif

namespace sierra {
namespace nalu {

struct TeamHandleType {};

// SharedMemView<T*> .. SharedMemView<T****>: row-major (Kokkos::LayoutRight, include/KokkosInterface.h:33-34), owning
template <typename P>
class SharedMemView;
namespace detail {
template <typename T, int RANK>
class ViewN {
public:
  typedef T value_type;
  ViewN() = default;
  ViewN(int n0, int n1 = 1, int n2 = 1, int n3 = 1) { resize(n0, n1, n2, n3); }
  void resize(int n0, int n1 = 1, int n2 = 1, int n3 = 1)
  {
    n_[0] = n0; n_[1] = n1; n_[2] = n2; n_[3] = n3;
    d_.assign(size_t(n0) * n1 * n2 * n3, T());
  }
  int extent(int d) const { return d < RANK ? n_[d] : 1; }
  int extent_int(int d) const { return extent(d); }
  size_t size() const { return d_.size(); }
  T *data() { return d_.data(); }
  const T *data() const { return d_.data(); }
  T &operator()(int i) { return d_[i]; }
  T &operator()(int i, int j) { return d_[size_t(i) * n_[1] + j]; }
  T &operator()(int i, int j, int k) { return d_[(size_t(i) * n_[1] + j) * n_[2] + k]; }
  T &operator()(int i, int j, int k, int l) { return d_[((size_t(i) * n_[1] + j) * n_[2] + k) * n_[3] + l]; }
  const T &operator()(int i) const { return d_[i]; }
  const T &operator()(int i, int j) const { return d_[size_t(i) * n_[1] + j]; }
  const T &operator()(int i, int j, int k) const { return d_[(size_t(i) * n_[1] + j) * n_[2] + k]; }
private:
  int n_[4] = {0, 1, 1, 1};
  std::vector<T> d_;
};
} // namespace detail
template <typename T> class SharedMemView<T *> : public detail::ViewN<T, 1> { using detail::ViewN<T, 1>::ViewN; };
template <typename T> class SharedMemView<T **> : public detail::ViewN<T, 2> { using detail::ViewN<T, 2>::ViewN; };
template <typename T> class SharedMemView<T ***> : public detail::ViewN<T, 3> { using detail::ViewN<T, 3>::ViewN; };
template <typename T> class SharedMemView<T ****> : public detail::ViewN<T, 4> { using detail::ViewN<T, 4>::ViewN; };

template <typename T>
class MasterElementViews {
public:
  typedef T value_type;
  // include/ScratchViews.h:209-344: one view per requested quantity, sized by the topology; returns the scalars
  int create_master_element_views(const TeamHandleType &, const std::set<ELEM_DATA_NEEDED> &dataEnums, int nDim, int /*nodesPerFace*/,
                                  int nodesPerElem, int /*numFaceIp*/, int numScsIp, int numScvIp, int /*numFemIp*/)
  {
    int n = 0;
    bool needDeriv = false, needDerivScv = false, needDetj = false, needDetjScv = false;
    for (ELEM_DATA_NEEDED d : dataEnums) {
      switch (d) {
      case SCS_AREAV: scs_areav.resize(numScsIp, nDim); n += numScsIp * nDim; break;
      case SCS_GRAD_OP: dndx.resize(numScsIp, nodesPerElem, nDim); n += nodesPerElem * numScsIp * nDim; needDeriv = needDetj = true; break;
      case SCS_SHIFTED_GRAD_OP: dndx_shifted.resize(numScsIp, nodesPerElem, nDim); n += nodesPerElem * numScsIp * nDim; needDeriv = needDetj = true; break;
      case SCS_GIJ: gijUpper.resize(numScsIp, nDim, nDim); gijLower.resize(numScsIp, nDim, nDim); n += 2 * numScsIp * nDim * nDim; needDeriv = true; break;
      case SCV_VOLUME: scv_volume.resize(numScvIp); n += numScvIp; break;
      case SCV_GRAD_OP: dndx_scv.resize(numScvIp, nodesPerElem, nDim); n += nodesPerElem * numScvIp * nDim; needDerivScv = needDetjScv = true; break;
      case SCV_SHIFTED_GRAD_OP: dndx_scv_shifted.resize(numScvIp, nodesPerElem, nDim); n += nodesPerElem * numScvIp * nDim; needDerivScv = needDetjScv = true; break;
      default: STK_ThrowRequireMsg(false, "MasterElementViews: this master-element quantity is not served on the Hex8 element path"); break;
      }
    }
    if (needDeriv) { deriv.resize(numScsIp, nodesPerElem, nDim); n += nodesPerElem * numScsIp * nDim; }
    if (needDerivScv) { deriv_scv.resize(numScvIp, nodesPerElem, nDim); n += nodesPerElem * numScvIp * nDim; }
    if (needDetj) { det_j.resize(numScsIp); n += numScsIp; }
    if (needDetjScv) { det_j_scv.resize(numScvIp); n += numScvIp; }
    return n;
  }
  // include/ScratchViews.h:346-398 with the Hex8 operators of csrc/hex8_dev.cuh (HexSCS::determinant / grad_op /
  // shifted_grad_op, HexSCV::determinant / grad_op: src/master_element/Hex8CVFEM.C:149-178,523-581)
  void fill_master_element_views(const std::set<ELEM_DATA_NEEDED> &dataEnums, SharedMemView<double **> *coordsView, MasterElement *,
                                 MasterElement *, MasterElement *, MasterElement *, MasterElement *, int = 0)
  {
    namespace h8 = ::nalu_b200::hex8;
    double c[8][3];
    for (int n = 0; n < 8; ++n)
      for (int d = 0; d < 3; ++d) c[n][d] = (*coordsView)(n, d);
    auto grad = [&](SharedMemView<T ***> &out, SharedMemView<T *> &det, double q, bool scv, int nip) {
      for (int ip = 0; ip < nip; ++ip) {
        double s1, s2, s3, g[8][3];
        if (scv) h8::scv_intg_loc(ip, q, s1, s2, s3);
        else h8::scs_intg_loc(ip, q, s1, s2, s3);
        det(ip) = h8::grad_op_ip(c, s1, s2, s3, g);
        for (int n = 0; n < 8; ++n)
          for (int d = 0; d < 3; ++d) out(ip, n, d) = g[n][d];
      }
    };
    for (ELEM_DATA_NEEDED d : dataEnums) {
      switch (d) {
      case SCS_AREAV:
        for (int ip = 0; ip < 12; ++ip) {
          double a[3];
          h8::scs_areav_ip(c, ip, a);
          for (int k = 0; k < 3; ++k) scs_areav(ip, k) = a[k];
        }
        break;
      case SCS_GRAD_OP: grad(dndx, det_j, 0.25, false, 12); break;
      case SCS_SHIFTED_GRAD_OP: grad(dndx_shifted, det_j, 0.5, false, 12); break;
      case SCS_GIJ: // HexSCS::gij: always the unshifted points (src/master_element/Hex8CVFEM.C:800-808)
        for (int ip = 0; ip < 12; ++ip) {
          double s1, s2, s3, G[12];
          h8::scs_intg_loc(ip, 0.25, s1, s2, s3);
          h8::gij_ip(c, s1, s2, s3, G);
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
              gijUpper(ip, i, j) = h8::sym3(G, i, j);
              gijLower(ip, i, j) = h8::sym3(G + 6, i, j);
            }
        }
        break;
      case SCV_VOLUME:
        for (int ip = 0; ip < 8; ++ip) scv_volume(ip) = h8::scv_volume_ip(c, ip);
        break;
      case SCV_GRAD_OP: grad(dndx_scv, det_j_scv, 0.25, true, 8); break;
      case SCV_SHIFTED_GRAD_OP: grad(dndx_scv_shifted, det_j_scv, 0.5, true, 8); break;
      default: break;
      }
    }
  }

  SharedMemView<T **> fc_areav, scs_areav;
  SharedMemView<T ***> dndx_fc_elem, dndx_shifted_fc_elem, dndx, dndx_shifted, dndx_scv, dndx_scv_shifted, deriv_fc, deriv, deriv_scv;
  SharedMemView<T *> det_j_fc, det_j, det_j_scv, scv_volume;
  SharedMemView<T ***> gijUpper, gijLower;
};

// what a field view may be (include/ScratchViews.h:31-48 ViewHolder / ViewT)
struct ViewHolder {
  virtual ~ViewHolder() {}
  int dim_ = 0;
};
template <typename V>
struct ViewT : public ViewHolder {
  ViewT(V v, int dim) : view_(v) { dim_ = dim; }
  V view_;
};

template <typename T>
class ScratchViews {
public:
  typedef T value_type;
  // include/ScratchViews.h:584-704: nodal scalar -> [npe], nodal vector -> [npe][d], nodal tensor -> [npe][d1][d2],
  // element field -> [d] or [d1][d2]; looked up by field.mesh_meta_data_ordinal()
  ScratchViews(const TeamHandleType &team, const stk::mesh::BulkData &bulkData, int nodesPerEntity, const ElemDataRequests &dataNeeded)
  {
    const stk::mesh::MetaData &meta = bulkData.mesh_meta_data();
    const int nDim = (int)meta.spatial_dimension();
    fieldViews.assign(meta.get_fields().size(), nullptr);
    int numScalars = 0;
    for (const FieldInfo &fi : dataNeeded.get_fields()) {
      const unsigned ord = fi.field->mesh_meta_data_ordinal();
      const bool nodal = fi.field->entity_rank() == stk::topology::NODE_RANK;
      const unsigned d1 = fi.scalarsDim1, d2 = fi.scalarsDim2;
      if (nodal) {
        if (d2 == 0) {
          if (d1 == 1) { fieldViews[ord] = new ViewT<SharedMemView<T *>>(SharedMemView<T *>(nodesPerEntity), 1); numScalars += nodesPerEntity; }
          else { fieldViews[ord] = new ViewT<SharedMemView<T **>>(SharedMemView<T **>(nodesPerEntity, (int)d1), 2); numScalars += nodesPerEntity * d1; }
        } else {
          fieldViews[ord] = new ViewT<SharedMemView<T ***>>(SharedMemView<T ***>(nodesPerEntity, (int)d1, (int)d2), 3);
          numScalars += nodesPerEntity * d1 * d2;
        }
      } else {
        if (d2 == 0) { fieldViews[ord] = new ViewT<SharedMemView<T *>>(SharedMemView<T *>((int)d1), 1); numScalars += d1; }
        else { fieldViews[ord] = new ViewT<SharedMemView<T **>>(SharedMemView<T **>((int)d1, (int)d2), 2); numScalars += d1 * d2; }
      }
    }
    // master-element views per coordinates type (include/ScratchViews.h:600-628); Hex8: 12 SCS / 8 SCV points
    const int numScsIp = dataNeeded.get_cvfem_surface_me() ? 12 : 0, numScvIp = dataNeeded.get_cvfem_volume_me() ? 8 : 0;
    for (auto it = dataNeeded.get_coordinates_map().begin(); it != dataNeeded.get_coordinates_map().end(); ++it) {
      hasCoordField[it->first] = true;
      numScalars += meViews[it->first].create_master_element_views(team, dataNeeded.get_data_enums(it->first), nDim, 0, nodesPerEntity, 0,
                                                                  numScsIp, numScvIp, 0);
    }
    num_bytes_required = numScalars * (int)sizeof(T);
  }
  ScratchViews(const ScratchViews &) = delete;
  virtual ~ScratchViews()
  {
    for (ViewHolder *vh : fieldViews) delete vh;
  }

  SharedMemView<T *> &get_scratch_view_1D(const stk::mesh::FieldBase &field) { return get<SharedMemView<T *>>(field, 1); }
  SharedMemView<T **> &get_scratch_view_2D(const stk::mesh::FieldBase &field) { return get<SharedMemView<T **>>(field, 2); }
  SharedMemView<T ***> &get_scratch_view_3D(const stk::mesh::FieldBase &field) { return get<SharedMemView<T ***>>(field, 3); }
  SharedMemView<T ****> &get_scratch_view_4D(const stk::mesh::FieldBase &field) { return get<SharedMemView<T ****>>(field, 4); }
  MasterElementViews<T> &get_me_views(const COORDS_TYPES cType)
  {
    STK_ThrowRequireMsg(hasCoordField[cType], "ScratchViews::get_me_views: coordinates type not registered");
    return meViews[cType];
  }
  bool has_coord_field(const COORDS_TYPES cType) const { return hasCoordField[cType]; }
  int total_bytes() const { return num_bytes_required; }
  const std::vector<ViewHolder *> &get_field_views() const { return fieldViews; }

  const int32_t *elemNodes = nullptr; // local node indices of the element (the shim's Entity list)

private:
  template <typename V>
  V &get(const stk::mesh::FieldBase &field, int dim)
  {
    ViewHolder *vh = field.mesh_meta_data_ordinal() < fieldViews.size() ? fieldViews[field.mesh_meta_data_ordinal()] : nullptr;
    STK_ThrowRequireMsg(vh != nullptr, "ScratchViews ERROR, trying to get " << dim << "D scratch-view for field " << field.name()
                                                                              << " which wasn't declared as pre-req field.");
    STK_ThrowRequireMsg(vh->dim_ == dim, "ScratchViews ERROR, field " << field.name() << " was registered with another rank");
    return static_cast<ViewT<V> *>(vh)->view_;
  }
  std::vector<ViewHolder *> fieldViews;
  MasterElementViews<T> meViews[MAX_COORDS_TYPES];
  bool hasCoordField[MAX_COORDS_TYPES] = {false, false};
  int num_bytes_required = 0;
};

// src/ScratchViews.C:408-484: gather the registered fields of one element (node rank: the field data of its nodes in
// connectivity order; element rank: the element's own data), then the master-element views from the gathered coordinates
inline void fill_pre_req_data(ElemDataRequests &dataNeeded, const stk::mesh::BulkData &bulkData, stk::mesh::Entity elem,
                              ScratchViews<double> &prereqData, bool fillMEViews = true)
{
  const int32_t *nodes = bulkData.begin_nodes(elem);
  const int npe = 8;
  prereqData.elemNodes = nodes;
  for (const FieldInfo &fi : dataNeeded.get_fields()) {
    const stk::mesh::FieldBase &f = *fi.field;
    const unsigned d1 = fi.scalarsDim1, d2 = fi.scalarsDim2;
    const double *base = f.host_data();
    if (f.entity_rank() == stk::topology::NODE_RANK) {
      if (d2 == 0 && d1 == 1) {
        auto &v = prereqData.get_scratch_view_1D(f);
        for (int n = 0; n < npe; ++n) v(n) = base[size_t(nodes[n]) * f.max_size()];
      } else if (d2 == 0) {
        auto &v = prereqData.get_scratch_view_2D(f);
        for (int n = 0; n < npe; ++n)
          for (unsigned d = 0; d < d1; ++d) v(n, (int)d) = base[size_t(nodes[n]) * f.max_size() + d];
      } else {
        auto &v = prereqData.get_scratch_view_3D(f);
        for (int n = 0; n < npe; ++n)
          for (unsigned a = 0; a < d1; ++a)
            for (unsigned b = 0; b < d2; ++b) v(n, (int)a, (int)b) = base[size_t(nodes[n]) * f.max_size() + a * d2 + b];
      }
    } else {
      if (d2 == 0) {
        auto &v = prereqData.get_scratch_view_1D(f);
        for (unsigned d = 0; d < d1; ++d) v((int)d) = base[size_t(elem.index) * f.max_size() + d];
      } else {
        auto &v = prereqData.get_scratch_view_2D(f);
        for (unsigned a = 0; a < d1; ++a)
          for (unsigned b = 0; b < d2; ++b) v((int)a, (int)b) = base[size_t(elem.index) * f.max_size() + a * d2 + b];
      }
    }
  }
  if (!fillMEViews) return;
  for (auto it = dataNeeded.get_coordinates_map().begin(); it != dataNeeded.get_coordinates_map().end(); ++it) {
    auto &coords = prereqData.get_scratch_view_2D(*it->second);
    prereqData.get_me_views(it->first).fill_master_element_views(dataNeeded.get_data_enums(it->first), &coords, nullptr,
                                                                 dataNeeded.get_cvfem_surface_me(), dataNeeded.get_cvfem_volume_me(),
                                                                 nullptr, nullptr);
  }
}

} // namespace nalu
} // namespace sierra
