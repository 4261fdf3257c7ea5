// MasterElement.h -- the master-element objects the kernels' constructors query
// (include/master_element/MasterElement.h, MasterElementFactory.h; Hex8CVFEM.C:75-108,255-345,832-847).
// Only the constant tables live on the host (lrscv, ipNodeMap, shape functions at the integration
// points); the geometric operators (determinant, grad_op, ...) run inside the CUDA kernels
// (nalu_b200/csrc/hex8_dev.cuh, box_tile.cuh) and asking the host object for them throws, like the
// reference's unimplemented MasterElement virtuals (include/master_element/MasterElement.h:56-272).
#pragma once
#include "../stk_shim.h"
#include <stdexcept>
#include <vector>

namespace sierra {
namespace nalu {

class MasterElement {
public:
  virtual ~MasterElement() {}
  virtual const int *adjacentNodes() { throw std::runtime_error("adjacentNodes not implemented"); }
  virtual const int *ipNodeMap(int /*ordinal*/ = 0) { throw std::runtime_error("ipNodeMap not implemented"); }
  virtual const int *side_node_ordinals(int /*sideOrdinal*/) { throw std::runtime_error("side_node_ordinals not implemented"); }
  virtual int opposingNodes(const int /*ordinal*/, const int /*node*/) { throw std::runtime_error("opposingNodes not implemented"); }
  virtual void shape_fcn(double *) { throw std::runtime_error("shape_fcn not implemented"); }
  virtual void shifted_shape_fcn(double *) { throw std::runtime_error("shifted_shape_fcn not implemented"); }
  virtual void determinant(const int, const double *, double *, double *)
  {
    throw std::runtime_error("determinant: geometric operators run on the device on this path");
  }
  virtual void grad_op(const int, const double *, double *, double *, double *, double *)
  {
    throw std::runtime_error("grad_op: geometric operators run on the device on this path");
  }
  int nDim_ = 3, nodesPerElement_ = 0, numIntPoints_ = 0;
};

class HexSCS : public MasterElement {
public:
  HexSCS();
  const int *adjacentNodes() override { return lrscv_; }
  // boundary tables (Hex8CVFEM.C:288-300,381-393,414-422): face ip -> nearest node, side -> node ordinals, opposing node
  const int *ipNodeMap(int ordinal = 0) override { return &faceNodes_[ordinal * 4]; }
  const int *side_node_ordinals(int ordinal) override { return &faceNodes_[ordinal * 4]; }
  int opposingNodes(const int ordinal, const int node) override { return oppNode_[ordinal * 4 + node]; }
  void shape_fcn(double *shpfc) override;         // [12][8], intgLoc_
  void shifted_shape_fcn(double *shpfc) override; // [12][8], intgLocShift_
private:
  int lrscv_[24];
  int faceNodes_[24], oppNode_[24];
};

// Quad3DSCS (src/master_element/Quad43DCVFEM.C:20-148): the face master element of a Hex8 side
class Quad3DSCS : public MasterElement {
public:
  Quad3DSCS();
  const int *ipNodeMap(int = 0) override { return ipNodeMap_; }
  void shape_fcn(double *shpfc) override;         // [4][4], intgLoc_ (+-1/4)
  void shifted_shape_fcn(double *shpfc) override; // [4][4], the nodes
private:
  int ipNodeMap_[4];
};

class HexSCV : public MasterElement {
public:
  HexSCV();
  const int *ipNodeMap(int = 0) override { return ipNodeMap_; }
  void shape_fcn(double *shpfc) override;         // [8][8]
  void shifted_shape_fcn(double *shpfc) override; // identity: lumping
private:
  int ipNodeMap_[8];
};

// HexahedralP2Element tables (src/master_element/Hex27CVFEM.C:25-75,192-262,355-421): STK node map, sub-control-surface end
// locations, the 2-point rule per segment
class HexahedralP2Element : public MasterElement {
public:
  HexahedralP2Element();
  int tensor_product_node_map(int i, int j, int k) const { return stkNodeMap_[i + 3 * (j + 3 * k)]; }
  double gauss_point_location(int nodeOrdinal, int gaussPointOrdinal) const;
  double shifted_gauss_point_location(int nodeOrdinal, int gaussPointOrdinal) const;
  void hex27_shape_fcn(int numIntPoints, const double *intgLoc, double *shpfc) const;
protected:
  int stkNodeMap_[27];
  double scsDist_, scsEndLoc_[4], gaussAbscissae_[2], gaussAbscissaeShift_[6];
  std::vector<double> intgLoc_, intgLocShift_;
};
// Hex27SCS (Hex27CVFEM.C:851-1050): 216 points, lrscv
class Hex27SCS : public HexahedralP2Element {
public:
  Hex27SCS();
  const int *adjacentNodes() override { return lrscv_.data(); }
  void shape_fcn(double *shpfc) override { hex27_shape_fcn(numIntPoints_, intgLoc_.data(), shpfc); }                 // [216][27]
  void shifted_shape_fcn(double *shpfc) override { hex27_shape_fcn(numIntPoints_, intgLocShift_.data(), shpfc); }
private:
  std::vector<int> lrscv_;
};
// Hex27SCV (Hex27CVFEM.C:634-722): 216 points, 8 per node
class Hex27SCV : public HexahedralP2Element {
public:
  Hex27SCV();
  const int *ipNodeMap(int = 0) override { return ipNodeMap_.data(); }
  void shape_fcn(double *shpfc) override { hex27_shape_fcn(numIntPoints_, intgLoc_.data(), shpfc); }
  void shifted_shape_fcn(double *shpfc) override { hex27_shape_fcn(numIntPoints_, intgLocShift_.data(), shpfc); }   // nodal: lumping
private:
  std::vector<int> ipNodeMap_;
};

struct MasterElementRepo {
  static MasterElement *get_surface_master_element(const stk::topology &theTopo);
  static MasterElement *get_volume_master_element(const stk::topology &theTopo);
};

} // namespace nalu
} // namespace sierra
