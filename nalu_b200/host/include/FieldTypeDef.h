// FieldTypeDef.h -- the reference's field typedefs (include/FieldTypeDef.h) on the shim
#pragma once
#include "stk_shim.h"
namespace sierra {
namespace nalu {
typedef stk::mesh::Field<double> ScalarFieldType;
typedef stk::mesh::Field<double> VectorFieldType;
typedef stk::mesh::Field<double> GenericFieldType;
typedef stk::mesh::Field<double> TensorFieldType;
} // namespace nalu
} // namespace sierra
