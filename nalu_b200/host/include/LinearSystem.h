// LinearSystem.h -- the reference's LinearSystem interface (include/LinearSystem.h:39-167) and its
// B200 implementation.  DeviceCsrLinearSystem is what LinearSystem::create returns on this path: the
// role TpetraLinearSystem plays in the reference (src/TpetraLinearSystem.C), with the CSR graph built
// once per mesh by the C ABI (nalu_b200_graph_*) and values / rhs resident in HBM.
#pragma once
#include "../../../include/nalu_b200.h"
#include "stk_shim.h"
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace sierra {
namespace nalu {

class EquationSystem;
class Realm;
class LinearSolver;

class LinearSystem {
public:
  LinearSystem(Realm &realm, const unsigned numDof, EquationSystem *eqSys, LinearSolver *linearSolver);
  virtual ~LinearSystem() {}
  static LinearSystem *create(Realm &realm, const unsigned numDof, EquationSystem *eqSys, LinearSolver *linearSolver);

  // Graph/Matrix Construction
  virtual void buildNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildFaceToNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildEdgeToNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildElemToNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildReducedElemToNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildFaceElemToNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildNonConformalNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void buildOversetNodeGraph(const stk::mesh::PartVector &parts) = 0;
  virtual void finalizeLinearSystem() = 0;

  // Matrix Assembly
  virtual void zeroSystem() = 0;
  /** sumInto for host callers (node algorithms, BCs): entities, rhs[n], lhs[n][n] row-major, n = numEntities*numDof */
  virtual void sumInto(unsigned numEntities, const stk::mesh::Entity *entities, const double *rhs, const double *lhs,
                       int *localIds, int *sortPermutation, const char *trace_tag) = 0;
  virtual void sumInto(const std::vector<stk::mesh::Entity> &sym_meshobj, std::vector<int> &scratchIds,
                       std::vector<double> &scratchVals, const std::vector<double> &rhs, const std::vector<double> &lhs,
                       const char *trace_tag = 0) = 0;
  virtual void applyDirichletBCs(stk::mesh::FieldBase *solutionField, stk::mesh::FieldBase *bcValuesField,
                                 const stk::mesh::PartVector &parts, const unsigned beginPos, const unsigned endPos) = 0;
  virtual void prepareConstraints(const unsigned beginPos, const unsigned endPos) = 0;
  virtual void resetRows(std::vector<stk::mesh::Entity> nodeList, const unsigned beginPos, const unsigned endPos) = 0;

  // Solve
  virtual int solve(stk::mesh::FieldBase *linearSolutionField) = 0;
  virtual void loadComplete() = 0;
  virtual void writeToFile(const char *filename, bool useOwned = true) = 0;
  virtual void writeSolutionToFile(const char *filename, bool useOwned = true) = 0;

  unsigned numDof() const { return numDof_; }
  const std::string name() { return eqSysName_; }

protected:
  Realm &realm_;
  EquationSystem *eqSys_;
  bool inConstruction_;
  const unsigned numDof_;
  const std::string eqSysName_;
  LinearSolver *linearSolver_;
};

// How the shared-row exchange of loadComplete reaches the other ranks.  The C ABI does the pack and
// the owner-side add on the device; moving the packed doubles between processes belongs to whatever
// the host application already has (MPI in the reference, NCCL through torch.distributed in bench.py).
struct HaloTransport {
  // graph handshake, once per mesh: send[dest] -> returns recv[src]
  std::function<std::map<int, std::vector<int64_t>>(const std::map<int, std::vector<int64_t>> &)> exchange_graph;
  // loadComplete: device buffers; send_bufs[i] goes to send_peers[i], recv_bufs[i] comes from recv_peers[i]
  std::function<void(const std::vector<int> &send_peers, const std::vector<double *> &send_bufs,
                     const std::vector<int64_t> &send_counts, const std::vector<int> &recv_peers,
                     const std::vector<double *> &recv_bufs, const std::vector<int64_t> &recv_counts)>
      exchange_rows;
};

class DeviceCsrLinearSystem : public LinearSystem {
public:
  DeviceCsrLinearSystem(Realm &realm, const unsigned numDof, EquationSystem *eqSys, LinearSolver *linearSolver);
  ~DeviceCsrLinearSystem();

  void buildNodeGraph(const stk::mesh::PartVector &parts) override;        // diagonal only: implied by any element graph
  void buildElemToNodeGraph(const stk::mesh::PartVector &parts) override;  // TpetraLinearSystem.C:392-413
  void buildFaceToNodeGraph(const stk::mesh::PartVector &) override { unsupported("buildFaceToNodeGraph"); }
  void buildEdgeToNodeGraph(const stk::mesh::PartVector &) override { unsupported("buildEdgeToNodeGraph"); }
  void buildReducedElemToNodeGraph(const stk::mesh::PartVector &) override { unsupported("buildReducedElemToNodeGraph"); }
  // TpetraLinearSystem.C:487-521: the element attached to each face of the parts; all elements of the rank mesh are
  // connected through the element graph already, so this adds nothing to it
  void buildFaceElemToNodeGraph(const stk::mesh::PartVector &parts) override;
  void buildNonConformalNodeGraph(const stk::mesh::PartVector &) override { unsupported("buildNonConformalNodeGraph"); }
  void buildOversetNodeGraph(const stk::mesh::PartVector &) override { unsupported("buildOversetNodeGraph"); }
  void finalizeLinearSystem() override;                                     // TpetraLinearSystem.C:1257-1353

  void zeroSystem() override;                                               // :1355-1372
  void sumInto(unsigned numEntities, const stk::mesh::Entity *entities, const double *rhs, const double *lhs,
               int *localIds, int *sortPermutation, const char *trace_tag) override; // :1457-1522
  void sumInto(const std::vector<stk::mesh::Entity> &sym_meshobj, std::vector<int> &scratchIds,
               std::vector<double> &scratchVals, const std::vector<double> &rhs, const std::vector<double> &lhs,
               const char *trace_tag = 0) override;                         // :1524-1577
  void applyDirichletBCs(stk::mesh::FieldBase *solutionField, stk::mesh::FieldBase *bcValuesField,
                         const stk::mesh::PartVector &parts, const unsigned beginPos, const unsigned endPos) override; // :1579-1653
  void prepareConstraints(const unsigned, const unsigned) override { unsupported("prepareConstraints"); }
  void resetRows(std::vector<stk::mesh::Entity> nodeList, const unsigned beginPos, const unsigned endPos) override; // :1706-1755
  int solve(stk::mesh::FieldBase *) override { unsupported("solve (the linear solver is a consumer of this path)"); return 1; }
  void loadComplete() override;                                             // :1757-1778
  void writeToFile(const char *filename, bool useOwned = true) override;   // MatrixMarket coordinate file
  void writeSolutionToFile(const char *, bool = true) override { unsupported("writeSolutionToFile"); }

  // ---- what the element algorithm and the parity tests use -----------------------------------
  /** batched device scatter: every active kernel of one AssembleElemSolverAlgorithm in one launch */
  void assembleElemDevice(const nalu_b200_kernel_desc *descs, int n, bool overwrite);
  /** every active kernel of one AssembleFaceElemSolverAlgorithm over the faces of a face set */
  void assembleFaceElemDevice(const nalu_b200_face_set *faces, const nalu_b200_face_kernel_desc *descs, int n);
  void setTransport(const HaloTransport &t) { transport_ = t; }
  const nalu_b200_graph *graph() const { return graph_; }
  int numOwnedRows() const { return info_.n_owned_rows; }
  int numSharedNotOwnedRows() const { return info_.n_shared_rows; }
  int64_t nnzOwned() const { return info_.nnz_owned; }
  std::vector<uint64_t> ownedRowPtr() const;
  std::vector<int32_t> ownedColIdx() const;
  std::vector<int64_t> ownedRowGids() const;
  std::vector<int64_t> colGids() const;
  std::vector<double> ownedValues() const; // device -> host copy
  std::vector<double> ownedRhs() const;
  double *deviceValues() const { return d_values_; }
  double *deviceRhs() const { return d_rhs_; }
  int lastAssembleKernelFamily() const { return last_family_; } // 1 general (atomic), 2 tiled write-once, 3 tiled + general
  int entityToLID(stk::mesh::Entity node) const;

private:
  void beginLinearSystemConstruction();
  [[noreturn]] void unsupported(const char *what) const;
  nalu_b200_graph *graph_ = nullptr;
  nalu_b200_graph_info info_{};
  double *d_values_ = nullptr, *d_rhs_ = nullptr;
  std::vector<int32_t> node_row_lid_;
  bool elemGraphRequested_ = false;
  int last_family_ = 0;
  HaloTransport transport_;
  std::vector<double *> sendBufs_, recvBufs_;
};

} // namespace nalu
} // namespace sierra
