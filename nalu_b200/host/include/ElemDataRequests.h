// ElemDataRequests.h -- the registration half of the reference's gather machinery, kept verbatim in
// meaning: include/ElemDataRequests.h:26-180, src/ElemDataRequests.C:13-102.
// Kernels announce which nodal / element fields and which master-element quantities they need; the
// algorithm takes the union.  On this path the union decides what the fused CUDA launch gathers.
#pragma once
#include "stk_shim.h"
#include <array>
#include <map>
#include <set>
#include <string>

namespace sierra {
namespace nalu {

class MasterElement;

enum ELEM_DATA_NEEDED {
  FC_AREAV = 0, SCS_AREAV, SCS_FACE_GRAD_OP, SCS_SHIFTED_FACE_GRAD_OP, SCS_GRAD_OP, SCS_SHIFTED_GRAD_OP, SCS_GIJ,
  SCV_VOLUME, SCV_GRAD_OP, SCV_SHIFTED_GRAD_OP, FEM_GRAD_OP, FEM_SHIFTED_GRAD_OP, FEM_DET_J, FEM_NORMAL,
  FEM_FACE_NORMAL, FEM_FACE_GRAD_OP, FEM_FACE_DET_J, FEM_GIJ,
};
enum COORDS_TYPES { CURRENT_COORDINATES = 0, MODEL_COORDINATES, MAX_COORDS_TYPES };
static const std::string CoordinatesTypeNames[] = {"current_coordinates", "model_coordinates"};

struct FieldInfo {
  FieldInfo(const stk::mesh::FieldBase *fld, unsigned scalars) : field(fld), scalarsDim1(scalars), scalarsDim2(0) {}
  FieldInfo(const stk::mesh::FieldBase *fld, unsigned d1, unsigned d2) : field(fld), scalarsDim1(d1), scalarsDim2(d2) {}
  const stk::mesh::FieldBase *field;
  unsigned scalarsDim1, scalarsDim2;
};
struct FieldInfoLess {
  bool operator()(const FieldInfo &a, const FieldInfo &b) const
  {
    return a.field->mesh_meta_data_ordinal() < b.field->mesh_meta_data_ordinal();
  }
};
typedef std::set<FieldInfo, FieldInfoLess> FieldSet;

class ElemDataRequests {
public:
  void add_master_element_call(ELEM_DATA_NEEDED data, COORDS_TYPES cType = CURRENT_COORDINATES)
  {
    STK_ThrowRequireMsg(coordsFields_.find(cType) != coordsFields_.end(),
                        "ElemDataRequests::add_master_element_call: Coordinates field ("
                            << CoordinatesTypeNames[cType]
                            << ") must be registered to ElemDataRequests before registering MasterElement calls");
    dataEnums[cType].insert(data);
  }
  void add_gathered_nodal_field(const stk::mesh::FieldBase &field, unsigned scalarsPerNode);
  void add_gathered_nodal_field(const stk::mesh::FieldBase &field, unsigned tensorDim1, unsigned tensorDim2);
  void add_face_field(const stk::mesh::FieldBase &field, unsigned scalarsPerFace);
  void add_face_field(const stk::mesh::FieldBase &field, unsigned tensorDim1, unsigned tensorDim2);
  void add_ip_field(const stk::mesh::FieldBase &field, unsigned scalarsPerElement);
  void add_element_field(const stk::mesh::FieldBase &field, unsigned scalarsPerElement);
  void add_ip_field(const stk::mesh::FieldBase &field, unsigned tensorDim1, unsigned tensorDim2);
  void add_element_field(const stk::mesh::FieldBase &field, unsigned tensorDim1, unsigned tensorDim2);
  void add_coordinates_field(const stk::mesh::FieldBase &field, unsigned scalarsPerNode, COORDS_TYPES cType);
  void add_cvfem_face_me(MasterElement *me) { meFC_ = me; }
  void add_cvfem_volume_me(MasterElement *me) { meSCV_ = me; }
  void add_cvfem_surface_me(MasterElement *me) { meSCS_ = me; }
  void add_fem_face_me(MasterElement *me) { meFCFEM_ = me; }
  void add_fem_volume_me(MasterElement *me) { meFEM_ = me; }

  const std::set<ELEM_DATA_NEEDED> &get_data_enums(const COORDS_TYPES cType) const { return dataEnums[cType]; }
  const stk::mesh::FieldBase *get_coordinates_field(const COORDS_TYPES cType) const
  {
    auto it = coordsFields_.find(cType);
    STK_ThrowRequireMsg(it != coordsFields_.end(), "ElemDataRequests::get_coordinates_field: Coordinates field ("
                                                       << CoordinatesTypeNames[cType]
                                                       << ") must be registered to ElemDataRequests before access");
    return it->second;
  }
  const std::map<COORDS_TYPES, const stk::mesh::FieldBase *> &get_coordinates_map() const { return coordsFields_; }
  const FieldSet &get_fields() const { return fields; }
  MasterElement *get_cvfem_face_me() const { return meFC_; }
  MasterElement *get_cvfem_volume_me() const { return meSCV_; }
  MasterElement *get_cvfem_surface_me() const { return meSCS_; }
  MasterElement *get_fem_volume_me() const { return meFEM_; }
  MasterElement *get_fem_face_me() const { return meFCFEM_; }

private:
  std::array<std::set<ELEM_DATA_NEEDED>, MAX_COORDS_TYPES> dataEnums;
  std::map<COORDS_TYPES, const stk::mesh::FieldBase *> coordsFields_;
  FieldSet fields;
  MasterElement *meFC_ = nullptr, *meSCS_ = nullptr, *meSCV_ = nullptr, *meFCFEM_ = nullptr, *meFEM_ = nullptr;
};

} // namespace nalu
} // namespace sierra
