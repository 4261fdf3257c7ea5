// assemble_hex27.cu -- MomentumAdvDiffElemKernel<AlgTraitsHex27> + sumInto on Hex27 (P2 CVFEM) meshes: BASELINE config 5,
// the stress test of the per-element block size (27 nodes, 216 sub-control-surface points, an 81 x 81 element matrix).
//
// Reference path: src/kernel/MomentumAdvDiffElemKernel.C:95-186 instantiated with AlgTraitsHex27 (include/AlgTraits.h:27-34),
// the master element src/master_element/Hex27CVFEM.C (HexahedralP2Element tables :25-75,192-262,355-632, Hex27SCS
// :887-1050 integration points / lrscv / weights, :1428-1470 area vectors, generic_grad_op
// include/master_element/MasterElementFunctions.h:52-108) and apply_coeff -> sumInto (src/TpetraLinearSystem.C:1457-1522).
//
// The reference fills a 52 KB scratch matrix per element (and 216 x 27 x 3 gradient weights) and then searches 81 CSR rows.
// Here a CTA owns an element and the element matrix never exists:
//   phase B  thread = integration point: J = sum_n x_n (x) dN_n/dxi and, with the same table pass, U' = sum_n u_n (x) dN_n/dxi;
//            the area vector of a constant-xi_d surface is +-w cof(J)[:, d] (x_t x x_u etc. ARE cofactor columns), so one
//            cofactor matrix gives the area vector, G = -mu_ip cof(J)/det J and div u = (cof/det) : U'.  16 doubles per point
//            go to shared memory (G, A, mdot, the divU stress) -- 27 KB instead of the 52 KB matrix + 140 KB of dndx.
//   phase C  thread = (row node r, three column nodes): walks the 12-24 points whose left or right sub-control volume is r's
//            and accumulates its three 3 x 3 blocks in registers: h = G dN_c/dxi (= -mu dN_c/dx), block += h (x) A,
//            diag += N_c mdot + h.A; the residual is -(row . u) - the divU stress (the reference's term-by-term sums
//            regrouped: its advective flux mdot * u_ip IS sum_c (N_c mdot) u_c).
//   scatter  elements are launched colour by colour (graph.cpp: elements of a colour share no row), every CSR entry receives
//            exactly one value-less reduction (red.global.add.f64) per launch: deterministic, no atomics in the contended
//            sense, no search (slot map built with the graph).
// Reference-element tables (tensor products of the 1-D quadratic Lagrange basis at the Gauss points of each sub-surface
// segment) are built on the host once, in two layouts: point-fastest for phase B, node-fastest for phase C.
#include "internal.h"

#include <cmath>
#include <mutex>
#include <vector>

namespace nalu_b200 {
namespace hex27 {

constexpr int NPE = 27, NIP = 216, NREC = 16;

struct Tables {
  double *dN_ipfast = nullptr;  // [27][3][216]
  double *dN_nodefast = nullptr; // [216][27][3]
  double *N_ipfast = nullptr;   // [27][216]
  double *N_nodefast = nullptr; // [216][27]
  double *wdir = nullptr;       // [216] signed weight: A = wdir * cof(J)[:, dir]
  int *dir = nullptr;           // [216]
  int *lr = nullptr;            // [216][2]
  int *inc_ptr = nullptr;       // [28]
  int *inc = nullptr;           // [432] ip << 1 | (1 if the node is the point's right node)
  double *pt = nullptr;         // [216][3] reference coordinates of the points
  int *stk = nullptr;           // [27] tensor-grid node -> STK ordinal
  double *scv_pt = nullptr;     // [216][3] Hex27SCV points, 8 per node, nodes in STK order (Hex27CVFEM.C:660-707 up to that order)
  double *scv_w = nullptr;      // [216]
};

struct HostTables {
  std::vector<double> dN, N, w, pt; // [216][27][3], [216][27], [216], [216][3]
  std::vector<int> dir, lr;
  std::vector<double> scv_pt, scv_w; // [216][3], [216]: point q of node n (STK ordinal) at index 8 n + q
};

// node (i, j, k) of the 3 x 3 x 3 tensor grid -> STK Hex27 ordinal (Hex27CVFEM.C:36-47)
static const int kStkNodeMap[27] = {0, 8, 1, 11, 21, 9, 3, 10, 2, 12, 25, 13, 23, 20, 24, 15, 26, 14, 4, 16, 5, 19, 22, 17, 7, 18, 6};

static void lagrange3(double x, double *l, double *d)
{
  l[0] = -0.5 * x * (1.0 - x); l[1] = (1.0 - x) * (1.0 + x); l[2] = 0.5 * x * (1.0 + x);
  d[0] = x - 0.5; d[1] = -2.0 * x; d[2] = x + 0.5;
}

static HostTables build_host_tables()
{
  HostTables T;
  T.dN.assign(size_t(NIP) * NPE * 3, 0.0);
  T.N.assign(size_t(NIP) * NPE, 0.0);
  T.w.assign(NIP, 0.0);
  T.pt.assign(3 * NIP, 0.0);
  T.dir.assign(NIP, 0);
  T.lr.assign(2 * NIP, 0);
  const double scsDist = std::sqrt(3.0) / 3.0;
  const double endLoc[4] = {-1.0, -scsDist, scsDist, 1.0};
  const double gaussAbs[2] = {-std::sqrt(3.0) / 3.0, std::sqrt(3.0) / 3.0};
  auto gp = [&](int node, int q) { // Gauss point q of the segment of 1-D node `node`
    const double a = endLoc[node], b = endLoc[node + 1];
    return gaussAbs[q] * (b - a) / 2.0 + (a + b) / 2.0;
  };
  auto seg = [&](int node) { return endLoc[node + 1] - endLoc[node]; };
  const double scsLoc[2] = {-scsDist, scsDist};
  const double orient[2] = {-1.0, 1.0};
  // area vector = w * (x_s1 x x_s2) with (s1, s2) = (t, s), (s, u), (t, u) for the u-, t-, s-direction
  // (Hex27CVFEM.C:1428-1470) = sign * cof(J)[:, d], sign = -1, -1, +1
  const double cofsign[3] = {+1.0, -1.0, -1.0}; // indexed by direction s = 0, t = 1, u = 2
  int ip = 0;
  for (int pass = 0; pass < 3; ++pass) { // u, t, s surfaces in this order (Hex27CVFEM.C:912-1047)
    const int d = 2 - pass;
    for (int m = 0; m < 2; ++m)
      for (int l = 0; l < 3; ++l)
        for (int k = 0; k < 3; ++k) {
          int lo[3], hi[3]; // tensor indices of the nodes on the two sides of surface m
          if (d == 2) { lo[0] = k; lo[1] = l; lo[2] = m; }
          else if (d == 1) { lo[0] = k; lo[1] = m; lo[2] = l; }
          else { lo[0] = m; lo[1] = k; lo[2] = l; }
          hi[0] = lo[0]; hi[1] = lo[1]; hi[2] = lo[2];
          hi[d] += 1;
          const int nlo = kStkNodeMap[lo[0] + 3 * (lo[1] + 3 * lo[2])], nhi = kStkNodeMap[hi[0] + 3 * (hi[1] + 3 * hi[2])];
          for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i, ++ip) {
              T.lr[2 * ip] = m == 0 ? nlo : nhi;
              T.lr[2 * ip + 1] = m == 0 ? nhi : nlo;
              double x[3];
              if (d == 2) { x[0] = gp(k, i); x[1] = gp(l, j); x[2] = scsLoc[m]; }
              else if (d == 1) { x[0] = gp(k, i); x[1] = scsLoc[m]; x[2] = gp(l, j); }
              else { x[0] = scsLoc[m]; x[1] = gp(k, i); x[2] = gp(l, j); }
              const double w = seg(k) * seg(l) * 0.25 * (d == 0 ? -orient[m] : orient[m]);
              T.w[ip] = w * cofsign[d];
              T.dir[ip] = d;
              for (int a = 0; a < 3; ++a) T.pt[3 * ip + a] = x[a];
              double L[3][3], D[3][3];
              for (int a = 0; a < 3; ++a) lagrange3(x[a], L[a], D[a]);
              for (int c = 0; c < 3; ++c)
                for (int b = 0; b < 3; ++b)
                  for (int a = 0; a < 3; ++a) {
                    const int n = kStkNodeMap[a + 3 * (b + 3 * c)];
                    T.N[size_t(ip) * NPE + n] = L[0][a] * L[1][b] * L[2][c];
                    T.dN[(size_t(ip) * NPE + n) * 3 + 0] = D[0][a] * L[1][b] * L[2][c];
                    T.dN[(size_t(ip) * NPE + n) * 3 + 1] = L[0][a] * D[1][b] * L[2][c];
                    T.dN[(size_t(ip) * NPE + n) * 3 + 2] = L[0][a] * L[1][b] * D[2][c];
                  }
            }
        }
  }
  // Hex27SCV::set_interior_info (Hex27CVFEM.C:660-707): 2 x 2 x 2 Gauss points in the sub-control volume of every node; stored
  // node-major in STK order so that the 8 points of a node are 8 neighbouring threads
  T.scv_pt.assign(3 * NIP, 0.0);
  T.scv_w.assign(NIP, 0.0);
  for (int c = 0; c < 3; ++c)
    for (int b = 0; b < 3; ++b)
      for (int a = 0; a < 3; ++a) {
        const int n = kStkNodeMap[a + 3 * (b + 3 * c)];
        for (int k = 0; k < 2; ++k)
          for (int j = 0; j < 2; ++j)
            for (int i = 0; i < 2; ++i) {
              const int q = 8 * n + i + 2 * (j + 2 * k);
              T.scv_pt[3 * q + 0] = gp(a, i);
              T.scv_pt[3 * q + 1] = gp(b, j);
              T.scv_pt[3 * q + 2] = gp(c, k);
              T.scv_w[q] = seg(a) * seg(b) * seg(c) * 0.125;
            }
      }
  return T;
}

template <class V> static int to_device(V **d, const std::vector<V> &h)
{
  NALU_CUDA(cudaMalloc((void **)d, h.size() * sizeof(V)));
  NALU_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
  return NALU_B200_OK;
}

// one table set per device, built on first use (never freed: 0.4 MB)
static int get_tables(const Tables **out)
{
  static std::mutex mu;
  static Tables tab[64];
  static bool have[64] = {false};
  int dev = 0;
  NALU_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) return NALU_B200_ERR_STATE;
  if (!have[dev]) {
    const HostTables H = build_host_tables();
    std::vector<double> dNi(size_t(NPE) * 3 * NIP), Ni(size_t(NPE) * NIP);
    for (int ip = 0; ip < NIP; ++ip)
      for (int n = 0; n < NPE; ++n) {
        Ni[size_t(n) * NIP + ip] = H.N[size_t(ip) * NPE + n];
        for (int j = 0; j < 3; ++j) dNi[(size_t(n) * 3 + j) * NIP + ip] = H.dN[(size_t(ip) * NPE + n) * 3 + j];
      }
    std::vector<int> ptr(NPE + 1, 0), inc(2 * NIP);
    for (int q = 0; q < 2 * NIP; ++q) ++ptr[H.lr[q] + 1];
    for (int n = 0; n < NPE; ++n) ptr[n + 1] += ptr[n];
    std::vector<int> pos(ptr.begin(), ptr.end() - 1);
    for (int q = 0; q < 2 * NIP; ++q) inc[pos[H.lr[q]]++] = q; // q = ip << 1 | side, ascending ip per node
    Tables &t = tab[dev];
    int rc;
    if ((rc = to_device(&t.dN_ipfast, dNi)) || (rc = to_device(&t.dN_nodefast, H.dN)) || (rc = to_device(&t.N_ipfast, Ni)) ||
        (rc = to_device(&t.N_nodefast, H.N)) || (rc = to_device(&t.wdir, H.w)) || (rc = to_device(&t.dir, H.dir)) ||
        (rc = to_device(&t.lr, H.lr)) || (rc = to_device(&t.inc_ptr, ptr)) || (rc = to_device(&t.inc, inc)) ||
        (rc = to_device(&t.pt, H.pt)) || (rc = to_device(&t.scv_pt, H.scv_pt)) || (rc = to_device(&t.scv_w, H.scv_w)) || (rc = to_device(&t.stk, std::vector<int>(kStkNodeMap, kStkNodeMap + 27))))
      return rc;
    have[dev] = true;
  }
  *out = &tab[dev];
  return NALU_B200_OK;
}

__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// cofactor matrix C of J (C[i][j] = d det / d J[i][j]) and det J
__device__ __forceinline__ double cofactor(const double (&J)[3][3], double (&C)[3][3])
{
  C[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  C[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  C[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  C[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
  C[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
  C[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
  C[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
  C[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
  C[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  return J[0][0] * C[0][0] + J[1][0] * C[1][0] + J[2][0] * C[2][0];
}

struct Params {
  Tables T;
  const int *elem_list; // elements of this colour (null: all elements, atomic scatter)
  long long n;
  const int *elem_nodes;        // [nElem][27]
  const unsigned char *elem_slot; // [nElem][27][27]
  const int *node_row;
  const long long *browptr;
  const double *coords, *vel, *visc, *mdot; // mdot: [nElem][216]
  double include_divu;
  // fused lumped MomentumMassElemKernel (src/kernel/MomentumMassElemKernel.C:98-171 with the shifted SCV shape functions)
  const double *vel_n, *vel_nm1, *rho, *rho_n, *rho_nm1, *gp;
  double g1_dt, g2_dt, g3_dt, dens_fac, om_dens_fac;
  double *values, *rhs;
  int *err;
};

constexpr int kThreads = 256;

// tensor-grid node (a, b, c) -> STK ordinal, usable as a compile-time constant in unrolled device loops
__host__ __device__ constexpr int stk_of(int q)
{
  constexpr int m[27] = {0, 8, 1, 11, 21, 9, 3, 10, 2, 12, 25, 13, 23, 20, 24, 15, 26, 14, 4, 16, 5, 19, 22, 17, 7, 18, 6};
  return m[q];
}
// row nodes in the order the threads of phase C take them: a node is the left or right node of 24 (body centre), 20 (face
// centres), 16 (edge mids) or 12 (corners) points; a warp holds 3.5 rows, so heavy and light rows alternate
// (face, corner, edge, edge: 64 points per 4 rows = the average) and every warp reaches the barrier at about the same time
__constant__ int c_row_order[27] = {21, 0, 8, 9, 22, 1, 10, 11, 23, 2, 12, 13, 24, 3, 14, 15, 25, 4, 16, 17, 26, 5, 18, 19, 20, 6, 7};

struct Smem {
  double rec[NREC][NIP];  // per point: G (9), A (3), mdot, divU stress (3)
  double oned[18][NIP];   // per point: l_a(xi_d) at [3 d + a], l_a'(xi_d) at [9 + 3 d + a]
  double part[NPE * 9][3];
  double x[NPE][3], u[NPE][3], mu[NPE];
  double vol[NPE]; // fused lumped mass: the part of each node's dual volume inside this element
  int row[NPE], inc_ptr[NPE + 1], inc[2 * NIP];
};

__device__ __forceinline__ void lagrange3_dev(double x, double (&l)[3], double (&d)[3])
{
  l[0] = -0.5 * x * (1.0 - x); l[1] = (1.0 - x) * (1.0 + x); l[2] = 0.5 * x * (1.0 + x);
  d[0] = x - 0.5; d[1] = -2.0 * x; d[2] = x + 0.5;
}

// weight * det J at one Hex27SCV point (Hex27SCV::determinant, Hex27CVFEM.C:743-763, :809-845); its own function so that its
// registers are not live together with phase B's
__device__ __noinline__ double scv_point_volume(const double (*x)[3], const double *pt, double w)
{
  double Lv[3][3], Dv[3][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) lagrange3_dev(pt[d], Lv[d], Dv[d]);
  double Jv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double t1 = Lv[1][b] * Lv[2][c], t2 = Dv[1][b] * Lv[2][c], t3 = Lv[1][b] * Dv[2][c];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int n = stk_of(a + 3 * (b + 3 * c));
        const double d0 = Dv[0][a] * t1, d1 = Lv[0][a] * t2, d2 = Lv[0][a] * t3;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double xn = x[n][i] - x[0][i];
          Jv[i][0] += d0 * xn; Jv[i][1] += d1 * xn; Jv[i][2] += d2 * xn;
        }
      }
    }
  double Cv[3][3];
  return w * cofactor(Jv, Cv);
}

// The 27 shape functions are products of three 1-D quadratic Lagrange polynomials (the factorisation visible in
// hex27_shape_fcn, Hex27CVFEM.C:355-421), so a point needs 18 numbers, not the 27 x 4 table entries the reference stores:
// phase B keeps them in registers (its thread IS the point), phase C reads them from shared memory.
template <bool MASS>
__global__ void __launch_bounds__(kThreads, 2) assemble_hex27_kernel(const __grid_constant__ Params P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem &S = *reinterpret_cast<Smem *>(smem_raw);

  const int t = threadIdx.x;
  const long long e = P.elem_list ? P.elem_list[blockIdx.x] : (long long)blockIdx.x;
  const int *en = P.elem_nodes + e * NPE;
  if (t < NPE) {
    const int n = en[t];
    S.row[t] = P.node_row[n];
    S.mu[t] = P.visc[n];
  }
  if (t < 81) {
    const int n = en[t / 3];
    S.x[t / 3][t % 3] = P.coords[(size_t)n * 3 + t % 3];
  } else if (t >= 96 && t < 96 + 81) {
    const int q = t - 96, n = en[q / 3];
    S.u[q / 3][q % 3] = P.vel[(size_t)n * 3 + q % 3];
  } else if (t >= 192 && t < 192 + NPE + 1) {
    S.inc_ptr[t - 192] = P.T.inc_ptr[t - 192];
  }
  for (int q = t; q < 2 * NIP; q += kThreads) S.inc[q] = P.T.inc[q];
  __syncthreads();

  if (MASS && t < NIP) {
    // fused lumped mass: thread = SCV point 8 n + q; the 8 points of a node are summed over 8 neighbouring lanes
    double v = scv_point_volume(S.x, P.T.scv_pt + 3 * t, P.T.scv_w[t]);
    if (!(v > 0.0)) atomicExch(P.err, 1);
    const unsigned mask = t >= 192 ? 0x00FFFFFFu : 0xFFFFFFFFu; // the last active warp holds points 192..215
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 4);
    if ((t & 7) == 0) S.vol[t >> 3] = v;
  }
  // ---- phase B: thread = integration point ---------------------------------------------------------------------
  if (t < NIP) {
    double L[3][3], D[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lagrange3_dev(P.T.pt[3 * t + d], L[d], D[d]);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        S.oned[3 * d + a][t] = L[d][a];
        S.oned[9 + 3 * d + a][t] = D[d][a];
      }
    }
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, U[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double mu = 0.0;
    // coordinates relative to the element's first node: the Jacobian is translation invariant, its rounding error is not
    const double x0 = S.x[0][0], y0 = S.x[0][1], z0 = S.x[0][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double t1 = L[1][b] * L[2][c], t2 = D[1][b] * L[2][c], t3 = L[1][b] * D[2][c];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const int n = stk_of(a + 3 * (b + 3 * c));
          const double d0 = D[0][a] * t1, d1 = L[0][a] * t2, d2 = L[0][a] * t3;
          const double xn[3] = {S.x[n][0] - x0, S.x[n][1] - y0, S.x[n][2] - z0};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            J[i][0] += d0 * xn[i]; J[i][1] += d1 * xn[i]; J[i][2] += d2 * xn[i];
            U[i][0] += d0 * S.u[n][i]; U[i][1] += d1 * S.u[n][i]; U[i][2] += d2 * S.u[n][i];
          }
          mu += (L[0][a] * t1) * S.mu[n];
        }
      }
    double C[3][3];
    const double det = cofactor(J, C);
    if (!(det > 0.0)) atomicExch(P.err, 1);
    const double inv = 1.0 / det;
    const int d = P.T.dir[t];
    const double w = P.T.wdir[t];
    double A[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) A[i] = w * (d == 0 ? C[i][0] : d == 1 ? C[i][1] : C[i][2]);
    double divu = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) divu += C[i][j] * U[i][j];
    divu *= inv;
    const double f = -mu * inv;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) S.rec[3 * i + j][t] = f * C[i][j];
    const double ds = 2.0 / 3.0 * mu * divu * P.include_divu;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      S.rec[9 + i][t] = A[i];
      S.rec[13 + i][t] = ds * A[i];
    }
    S.rec[12][t] = P.mdot[e * NIP + t];
  }
  __syncthreads();

  // ---- phase C: thread = (row node, the three column nodes (0..2, b, c) of one tensor line) -----------------------
  const int rr = t / 9, g = t - 9 * rr;
  if (rr < NPE) {
    const int r = c_row_order[rr];
    const int gb = g % 3, gc = g / 3;
    double blk[3][3][3], dg[3] = {0.0, 0.0, 0.0}, rs[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) blk[m][i][j] = 0.0;
    const int q0 = S.inc_ptr[r], q1 = S.inc_ptr[r + 1];
#pragma unroll 1
    for (int q = q0; q < q1; ++q) {
      const int code = S.inc[q], ip = code >> 1;
      const double sg = (code & 1) ? -1.0 : 1.0; // the left node gains what leaves through the point, the right node loses it
      double G[3][3], A[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) G[i][j] = S.rec[3 * i + j][ip];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        A[i] = sg * S.rec[9 + i][ip];
        rs[i] += sg * S.rec[13 + i][ip];
      }
      const double md = sg * S.rec[12][ip];
      const double l1 = S.oned[3 + gb][ip], d1b = S.oned[12 + gb][ip], l2 = S.oned[6 + gc][ip], d2c = S.oned[15 + gc][ip];
      const double t1 = l1 * l2, t2 = d1b * l2, t3 = l1 * d2c;
      // h = G grad_xi N_c with grad_xi N_c = (l0' t1, l0 t2, l0 t3): the parts common to the three columns of the line first
      double ga[3], gb[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        ga[i] = G[i][0] * t1;
        gb[i] = G[i][1] * t2 + G[i][2] * t3;
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const double l0 = S.oned[m][ip], d0a = S.oned[9 + m][ip];
        double h[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) h[i] = ga[i] * d0a + gb[i] * l0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) blk[m][i][j] += h[i] * A[j];
        dg[m] += (l0 * t1) * md + h[0] * A[0] + h[1] * A[1] + h[2] * A[2];
      }
    }
    // residual part of this thread's three blocks, then the blocks leave
    const int rn = S.row[r];
    const long long b0 = P.browptr[rn];
    const int nb = (int)(P.browptr[rn + 1] - b0);
    double *rowv = P.values + 9 * b0;
    const unsigned char *slots = P.elem_slot + (e * NPE + r) * NPE;
    double pr[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int c = P.T.stk[m + 3 * g];
#pragma unroll
      for (int i = 0; i < 3; ++i) blk[m][i][i] += dg[m];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) pr[i] -= blk[m][i][j] * S.u[c][j];
      double *dst = rowv + 3 * (int)slots[c];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (P.elem_list) red_add(dst + (long long)i * nb * 3 + j, blk[m][i][j]);
          else atomicAdd(dst + (long long)i * nb * 3 + j, blk[m][i][j]);
        }
    }
    if (g == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) pr[i] -= rs[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) S.part[t][i] = pr[i];
  }
  __syncthreads();
  if (t < 81) {
    const int q = t / 3, i = t - 3 * q;
    double v = 0.0;
#pragma unroll
    for (int gg = 0; gg < 9; ++gg) v += S.part[q * 9 + gg][i];
    const int r = c_row_order[q];
    if (MASS) {
      // shifted (nodal) SCV shape functions: every quantity is the row node's own
      const int nd = en[r];
      const double V = S.vol[r];
      const double rp = P.rho[nd], ui = S.u[r][i];
      v += -(P.g1_dt * rp * ui + P.g2_dt * P.rho_n[nd] * P.vel_n[(size_t)nd * 3 + i] + P.g3_dt * P.rho_nm1[nd] * P.vel_nm1[(size_t)nd * 3 + i]) * V -
           P.gp[(size_t)nd * 3 + i] * V * (rp * P.dens_fac + P.om_dens_fac);
      const int rn = S.row[r];
      const long long b0 = P.browptr[rn];
      const int nb = (int)(P.browptr[rn + 1] - b0);
      double *dd = P.values + 9 * b0 + (long long)i * nb * 3 + 3 * (int)P.elem_slot[(e * NPE + r) * NPE + r] + i;
      const double m = P.g1_dt * rp * V;
      if (P.elem_list) red_add(dd, m);
      else atomicAdd(dd, m);
    }
    double *dst = P.rhs + (long long)S.row[r] * 3 + i;
    if (P.elem_list) red_add(dst, v);
    else atomicAdd(dst, v);
  }
}

// ---- one-dof systems on Hex27: ContinuityAdvElemKernel (src/kernel/ContinuityAdvElemKernel.C:95-175) and ScalarAdvDiffElemKernel
// (src/kernel/ScalarAdvDiffElemKernel.C:69-127) with AlgTraitsHex27, default options ------------------------------------------
// Both add, per point, (b . grad_xi N_c + a N_c) to the left node's row and subtract it from the right node's:
//   continuity  b = -cof(J)^T A / det J (so that b . grad_xi N_c = -dN_c/dx . A), a = 0, flux f = mdot / tau with
//               mdot = (rho u)_ip . A - tau (grad p . A - (Gp)_ip . A) and grad p . A = -b . sum_c p_c grad_xi N_c
//   scalar      b = -Gamma_ip cof(J)^T A / det J, a = mdot_ip, residual = -(row . q)
// Same phases as the momentum kernel with a 5-double record per point.
struct OneDofParams {
  Tables T;
  const int *elem_list;
  const int *elem_nodes;
  const unsigned char *elem_slot;
  const int *node_row;
  const long long *browptr;
  const double *coords;
  const double *f0, *f1, *f2, *f3; // continuity: pressure, density, velocity, Gpdx; scalar: scalarQ, diffFluxCoeff
  const double *mdot;              // scalar: [nElem][216]
  double tau;                      // continuity: projTimeScale = dt / gamma1
  double *values, *rhs;
  int *err;
};
struct SmemOneDof {
  double rec[5][NIP];    // b (3), a, f
  double oned[18][NIP];
  double part[NPE * 9];
  double x[NPE][3], u[NPE][3], gp[NPE][3], p[NPE], rho[NPE];
  int row[NPE], inc_ptr[NPE + 1], inc[2 * NIP];
};

template <int KIND> // 0 continuity, 1 scalar adv-diff
__global__ void __launch_bounds__(kThreads, 2) assemble_hex27_onedof_kernel(const __grid_constant__ OneDofParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemOneDof &S = *reinterpret_cast<SmemOneDof *>(smem_raw);
  const int t = threadIdx.x;
  const long long e = P.elem_list ? P.elem_list[blockIdx.x] : (long long)blockIdx.x;
  const int *en = P.elem_nodes + e * NPE;
  if (t < NPE) {
    const int n = en[t];
    S.row[t] = P.node_row[n];
    S.p[t] = P.f0[n];   // pressure | scalarQ
    S.rho[t] = P.f1[n]; // density | diffFluxCoeff
  }
  if (t < 81) {
    const int n = en[t / 3];
    S.x[t / 3][t % 3] = P.coords[(size_t)n * 3 + t % 3];
  } else if (KIND == 0 && t >= 96 && t < 96 + 81) {
    const int q = t - 96, n = en[q / 3];
    S.u[q / 3][q % 3] = P.f2[(size_t)n * 3 + q % 3];
    S.gp[q / 3][q % 3] = P.f3[(size_t)n * 3 + q % 3];
  } else if (t >= 192 && t < 192 + NPE + 1) {
    S.inc_ptr[t - 192] = P.T.inc_ptr[t - 192];
  }
  for (int q = t; q < 2 * NIP; q += kThreads) S.inc[q] = P.T.inc[q];
  __syncthreads();

  if (t < NIP) {
    double L[3][3], D[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lagrange3_dev(P.T.pt[3 * t + d], L[d], D[d]);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        S.oned[3 * d + a][t] = L[d][a];
        S.oned[9 + 3 * d + a][t] = D[d][a];
      }
    }
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double pg[3] = {0, 0, 0}, ru[3] = {0, 0, 0}, gpi[3] = {0, 0, 0}, gam = 0.0;
    const double x0 = S.x[0][0], y0 = S.x[0][1], z0 = S.x[0][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const double t1 = L[1][b] * L[2][c], t2 = D[1][b] * L[2][c], t3 = L[1][b] * D[2][c];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const int n = stk_of(a + 3 * (b + 3 * c));
          const double d0 = D[0][a] * t1, d1 = L[0][a] * t2, d2 = L[0][a] * t3, sf = L[0][a] * t1;
          const double xn[3] = {S.x[n][0] - x0, S.x[n][1] - y0, S.x[n][2] - z0};
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            J[i][0] += d0 * xn[i]; J[i][1] += d1 * xn[i]; J[i][2] += d2 * xn[i];
          }
          if (KIND == 0) {
            pg[0] += d0 * S.p[n]; pg[1] += d1 * S.p[n]; pg[2] += d2 * S.p[n];
            const double sr = sf * S.rho[n];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              ru[i] += sr * S.u[n][i];
              gpi[i] += sf * S.gp[n][i];
            }
          } else {
            gam += sf * S.rho[n];
          }
        }
      }
    double C[3][3];
    const double det = cofactor(J, C);
    if (!(det > 0.0)) atomicExch(P.err, 1);
    const double inv = 1.0 / det;
    const int d = P.T.dir[t];
    const double w = P.T.wdir[t];
    double A[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) A[i] = w * (d == 0 ? C[i][0] : d == 1 ? C[i][1] : C[i][2]);
    const double scale = KIND == 0 ? -inv : -gam * inv;
#pragma unroll
    for (int k = 0; k < 3; ++k) b[k] = scale * (C[0][k] * A[0] + C[1][k] * A[1] + C[2][k] * A[2]);
    S.rec[0][t] = b[0]; S.rec[1][t] = b[1]; S.rec[2][t] = b[2];
    if (KIND == 0) {
      const double dpdxA = -(b[0] * pg[0] + b[1] * pg[1] + b[2] * pg[2]);
      const double mdot = (ru[0] * A[0] + ru[1] * A[1] + ru[2] * A[2]) -
                          P.tau * (dpdxA - (gpi[0] * A[0] + gpi[1] * A[1] + gpi[2] * A[2]));
      S.rec[3][t] = 0.0;
      S.rec[4][t] = mdot / P.tau;
    } else {
      S.rec[3][t] = P.mdot[e * NIP + t];
      S.rec[4][t] = 0.0;
    }
  }
  __syncthreads();

  const int rr = t / 9, g = t - 9 * rr;
  if (rr < NPE) {
    const int r = c_row_order[rr];
    const int gb = g % 3, gc = g / 3;
    double acc[3] = {0.0, 0.0, 0.0}, rs = 0.0;
    const int q0 = S.inc_ptr[r], q1 = S.inc_ptr[r + 1];
#pragma unroll 1
    for (int q = q0; q < q1; ++q) {
      const int code = S.inc[q], ip = code >> 1;
      const double sg = (code & 1) ? -1.0 : 1.0;
      const double l1 = S.oned[3 + gb][ip], d1b = S.oned[12 + gb][ip], l2 = S.oned[6 + gc][ip], d2c = S.oned[15 + gc][ip];
      const double t1 = l1 * l2, t2 = d1b * l2, t3 = l1 * d2c;
      const double ba = sg * S.rec[0][ip] * t1;
      const double bb = sg * (S.rec[1][ip] * t2 + S.rec[2][ip] * t3 + S.rec[3][ip] * t1);
      rs += sg * S.rec[4][ip];
#pragma unroll
      for (int m = 0; m < 3; ++m) acc[m] += ba * S.oned[9 + m][ip] + bb * S.oned[m][ip];
    }
    const int rn = S.row[r];
    double *rowv = P.values + P.browptr[rn];
    const unsigned char *slots = P.elem_slot + (e * NPE + r) * NPE;
    double pr = 0.0;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int c = P.T.stk[m + 3 * g];
      if (KIND == 1) pr -= acc[m] * S.p[c];
      if (P.elem_list) red_add(rowv + (int)slots[c], acc[m]);
      else atomicAdd(rowv + (int)slots[c], acc[m]);
    }
    if (KIND == 0 && g == 0) pr -= rs;
    S.part[t] = pr;
  }
  __syncthreads();
  if (t < NPE) {
    double v = 0.0;
#pragma unroll
    for (int gg = 0; gg < 9; ++gg) v += S.part[t * 9 + gg];
    double *dst = P.rhs + S.row[c_row_order[t]];
    if (P.elem_list) red_add(dst, v);
    else atomicAdd(dst, v);
  }
}

// mass_flow_rate_scs of a Hex27 element: mdot_ip = (sum_n N_n rho_n u_n) . A_ip (the unit tests' calc_mass_flow_rate_scs,
// unit_tests/kernels/UnitTestKernelUtils.C:547-628, instantiated with AlgTraitsHex27); thread = (element, point)
// with `pressure` non-null: ComputeMdotElemAlgorithm (src/ComputeMdotElemAlgorithm.C:85-254, default options) on AlgTraitsHex27,
// mdot_ip = (rho u_ip - tau (grad p_ip - (Gp)_ip)) . A_ip -- the flux ContinuityAdvElemKernel forms for its residual
__global__ void __launch_bounds__(kThreads) mdot_hex27_kernel(Tables T, long long e0, long long n_elems, const int *elem_nodes, const double *coords,
                                                               const double *density, const double *vel, const double *pressure,
                                                               const double *gpdx, double tau, double *mdot)
{
  const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (tid >= n_elems * NIP) return;
  const long long e = e0 + tid / NIP;
  const int ip = (int)(tid % NIP);
  const int *en = elem_nodes + e * NPE;
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, ru[3] = {0, 0, 0}, pg[3] = {0, 0, 0}, gp[3] = {0, 0, 0};
  const int n0 = en[0];
  const double x0[3] = {coords[(size_t)n0 * 3], coords[(size_t)n0 * 3 + 1], coords[(size_t)n0 * 3 + 2]};
  for (int n = 0; n < NPE; ++n) {
    const int nd = en[n];
    const double sf = T.N_ipfast[n * NIP + ip] * density[nd];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double x = coords[(size_t)nd * 3 + i] - x0[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) J[i][j] += T.dN_ipfast[(n * 3 + j) * NIP + ip] * x;
      ru[i] += sf * vel[(size_t)nd * 3 + i];
    }
    if (pressure) {
      const double p = pressure[nd], sn = T.N_ipfast[n * NIP + ip];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        pg[j] += T.dN_ipfast[(n * 3 + j) * NIP + ip] * p;
        gp[j] += sn * gpdx[(size_t)nd * 3 + j];
      }
    }
  }
  double C[3][3];
  const double det = cofactor(J, C);
  const int d = T.dir[ip];
  const double w = T.wdir[ip];
  double A[3], md = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[i] = w * (d == 0 ? C[i][0] : d == 1 ? C[i][1] : C[i][2]);
    md += ru[i] * A[i];
  }
  if (pressure) {
    // grad p . A = sum_k (cof(J)^T A)_k / det * (sum_c p_c dN_c/dxi_k)
    double dpdxA = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) dpdxA += (C[0][k] * A[0] + C[1][k] * A[1] + C[2][k] * A[2]) * pg[k];
    dpdxA /= det;
    md -= tau * (dpdxA - (gp[0] * A[0] + gp[1] * A[1] + gp[2] * A[2]));
  }
  mdot[e * NIP + ip] = md;
}

// AssembleNodalGradElemAlgorithm / AssembleNodalGradUElemAlgorithm (src/AssembleNodalGradElemAlgorithm.C:64-150,
// src/AssembleNodalGradUElemAlgorithm.C:120-195) on Hex27: Green-Gauss contributions q_ip A_ip / V to the left node and minus that
// to the right node of every point; thread = (element, point), atomics as for Hex8
template <int NCOMP>
__global__ void __launch_bounds__(kThreads) nodal_grad_hex27_kernel(Tables T, long long n_elems, const int *elem_nodes, const double *coords,
                                                                     const double *field, const double *dual_volume, double *grad)
{
  const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (tid >= n_elems * NIP) return;
  const long long e = tid / NIP;
  const int ip = (int)(tid - e * NIP);
  const int *en = elem_nodes + e * NPE;
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, q[NCOMP];
#pragma unroll
  for (int i = 0; i < NCOMP; ++i) q[i] = 0.0;
  const int n0 = en[0];
  const double x0[3] = {coords[(size_t)n0 * 3], coords[(size_t)n0 * 3 + 1], coords[(size_t)n0 * 3 + 2]};
  for (int n = 0; n < NPE; ++n) {
    const int nd = en[n];
    const double sf = T.N_ipfast[n * NIP + ip];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double x = coords[(size_t)nd * 3 + i] - x0[i];
#pragma unroll
      for (int j = 0; j < 3; ++j) J[i][j] += T.dN_ipfast[(n * 3 + j) * NIP + ip] * x;
    }
#pragma unroll
    for (int i = 0; i < NCOMP; ++i) q[i] += sf * field[(size_t)nd * NCOMP + i];
  }
  double C[3][3];
  cofactor(J, C);
  const int d = T.dir[ip];
  const double w = T.wdir[ip];
  const int nl = en[T.lr[2 * ip]], nr = en[T.lr[2 * ip + 1]];
  const double ivl = 1.0 / dual_volume[nl], ivr = 1.0 / dual_volume[nr];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double a = w * (d == 0 ? C[j][0] : d == 1 ? C[j][1] : C[j][2]);
#pragma unroll
    for (int i = 0; i < NCOMP; ++i) {
      const double fac = q[i] * a;
      atomicAdd(grad + (size_t)nl * NCOMP * 3 + i * 3 + j, fac * ivl);
      atomicAdd(grad + (size_t)nr * NCOMP * 3 + i * 3 + j, -fac * ivr);
    }
  }
}

} // namespace hex27

// [MomentumAdvDiffElemKernel] or [MomentumMassElemKernel(lumped), MomentumAdvDiffElemKernel] in either order
static bool classify_hex27(const nalu_b200_kernel_desc *k, int nk, int &ia, int &im)
{
  ia = im = -1;
  for (int i = 0; i < nk; ++i) {
    if (k[i].kind == NALU_K_MOMENTUM_ADV_DIFF && k[i].options == 0 && ia < 0) ia = i;
    else if (k[i].kind == NALU_K_MOMENTUM_MASS && k[i].options == NALU_OPT_LUMPED_MASS && im < 0) im = i;
    else return false;
  }
  return ia >= 0;
}

static bool hex27_onedof_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (g.npe != 27 || g.ndof != 1 || nk != 1 || !g.d_elem_nodes27 || !g.d_elem_slot27 || k[0].options != 0 || !k[0].nodal[NALU_F_COORDINATES])
    return false;
  if (k[0].kind == NALU_K_CONTINUITY_ADV)
    return k[0].nodal[NALU_F_VELOCITY] && k[0].nodal[NALU_F_DENSITY] && k[0].nodal[NALU_F_PRESSURE] && k[0].nodal[NALU_F_DPDX];
  if (k[0].kind == NALU_K_SCALAR_ADV_DIFF) return k[0].nodal[NALU_F_SCALARQ] && k[0].nodal[NALU_F_VISCOSITY] && k[0].mass_flow_rate;
  return false;
}

bool hex27_supported(const Graph &g, const nalu_b200_kernel_desc *k, int nk)
{
  if (hex27_onedof_supported(g, k, nk)) return true;
  int ia, im;
  if (g.npe != 27 || g.ndof != 3 || nk < 1 || nk > 2 || !g.d_elem_nodes27 || !g.d_elem_slot27 || !classify_hex27(k, nk, ia, im)) return false;
  const nalu_b200_kernel_desc &a = k[ia];
  if (!a.nodal[NALU_F_COORDINATES] || !a.nodal[NALU_F_VELOCITY] || !a.nodal[NALU_F_VISCOSITY] || !a.mass_flow_rate) return false;
  if (im >= 0) {
    const nalu_b200_kernel_desc &m = k[im];
    // one velocity / coordinates field for both kernels (velocity NP1 is the dof of both)
    if (m.nodal[NALU_F_COORDINATES] != a.nodal[NALU_F_COORDINATES] || m.nodal[NALU_F_VELOCITY] != a.nodal[NALU_F_VELOCITY]) return false;
    for (int f : {NALU_F_VELOCITY_N, NALU_F_VELOCITY_NM1, NALU_F_DENSITY, NALU_F_DENSITY_N, NALU_F_DENSITY_NM1, NALU_F_DPDX})
      if (!m.nodal[f]) return false;
  }
  return true;
}

int launch_assemble_hex27(const Graph &g, const nalu_b200_kernel_desc *kernels, int nk, double *d_values, double *d_rhs,
                          cudaStream_t stream)
{
  using namespace hex27;
  if (hex27_onedof_supported(g, kernels, nk)) {
    const nalu_b200_kernel_desc &k = kernels[0];
    const Tables *T = nullptr;
    int rc = get_tables(&T);
    if (rc) return rc;
    OneDofParams P;
    P.T = *T;
    P.elem_nodes = g.d_elem_nodes27;
    P.elem_slot = g.d_elem_slot27;
    P.node_row = g.d_node_row;
    P.browptr = (const long long *)g.d_browptr;
    P.coords = k.nodal[NALU_F_COORDINATES];
    const bool cont = k.kind == NALU_K_CONTINUITY_ADV;
    P.f0 = cont ? k.nodal[NALU_F_PRESSURE] : k.nodal[NALU_F_SCALARQ];
    P.f1 = cont ? k.nodal[NALU_F_DENSITY] : k.nodal[NALU_F_VISCOSITY];
    P.f2 = cont ? k.nodal[NALU_F_VELOCITY] : nullptr;
    P.f3 = cont ? k.nodal[NALU_F_DPDX] : nullptr;
    P.mdot = k.mass_flow_rate;
    P.tau = k.dt / k.gamma1;
    P.values = d_values;
    P.rhs = d_rhs;
    P.err = g.d_err;
    auto kern = cont ? assemble_hex27_onedof_kernel<0> : assemble_hex27_onedof_kernel<1>;
    NALU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemOneDof)));
    if (g.d_color_elems && !g.color_ptr.empty()) {
      for (size_t c = 0; c + 1 < g.color_ptr.size(); ++c) {
        const long long n = g.color_ptr[c + 1] - g.color_ptr[c];
        if (n <= 0) continue;
        P.elem_list = g.d_color_elems + g.color_ptr[c];
        kern<<<(unsigned)n, kThreads, sizeof(SmemOneDof), stream>>>(P);
        count_launch();
      }
    } else if (g.n_elems > 0) {
      P.elem_list = nullptr;
      kern<<<(unsigned)g.n_elems, kThreads, sizeof(SmemOneDof), stream>>>(P);
      count_launch();
    }
    NALU_CUDA(cudaGetLastError());
    return NALU_B200_OK;
  }
  int ia, im;
  if (!classify_hex27(kernels, nk, ia, im)) return NALU_B200_ERR_ARG;
  const nalu_b200_kernel_desc &k = kernels[ia];
  const Tables *T = nullptr;
  int rc = get_tables(&T);
  if (rc) return rc;
  Params P;
  P.T = *T;
  P.elem_nodes = g.d_elem_nodes27;
  P.elem_slot = g.d_elem_slot27;
  P.node_row = g.d_node_row;
  P.browptr = (const long long *)g.d_browptr;
  P.coords = k.nodal[NALU_F_COORDINATES];
  P.vel = k.nodal[NALU_F_VELOCITY];
  P.visc = k.nodal[NALU_F_VISCOSITY];
  P.mdot = k.mass_flow_rate;
  P.include_divu = k.include_divu;
  P.values = d_values;
  P.rhs = d_rhs;
  P.err = g.d_err;
  const bool mass = im >= 0;
  P.vel_n = P.vel_nm1 = P.rho = P.rho_n = P.rho_nm1 = P.gp = nullptr;
  P.g1_dt = P.g2_dt = P.g3_dt = P.dens_fac = P.om_dens_fac = 0.0;
  if (mass) {
    const nalu_b200_kernel_desc &m = kernels[im];
    P.vel_n = m.nodal[NALU_F_VELOCITY_N];
    P.vel_nm1 = m.nodal[NALU_F_VELOCITY_NM1];
    P.rho = m.nodal[NALU_F_DENSITY];
    P.rho_n = m.nodal[NALU_F_DENSITY_N];
    P.rho_nm1 = m.nodal[NALU_F_DENSITY_NM1];
    P.gp = m.nodal[NALU_F_DPDX];
    P.g1_dt = m.gamma1 / m.dt;
    P.g2_dt = m.gamma2 / m.dt;
    P.g3_dt = m.gamma3 / m.dt;
    P.dens_fac = m.dens_fac;
    P.om_dens_fac = m.om_dens_fac;
  }
  auto kern = mass ? assemble_hex27_kernel<true> : assemble_hex27_kernel<false>;
  NALU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  if (g.d_color_elems && !g.color_ptr.empty()) {
    const size_t ngroups = g.color_ptr.size() - 1;
    for (size_t c = 0; c < ngroups; ++c) {
      const long long n = g.color_ptr[c + 1] - g.color_ptr[c];
      if (n <= 0) continue;
      P.elem_list = g.d_color_elems + g.color_ptr[c];
      P.n = n;
      kern<<<(unsigned)n, kThreads, sizeof(Smem), stream>>>(P);
      count_launch();
    }
  } else if (g.n_elems > 0) { // an element that reaches one row through two of its nodes: contended atomics
    P.elem_list = nullptr;
    P.n = g.n_elems;
    kern<<<(unsigned)g.n_elems, kThreads, sizeof(Smem), stream>>>(P);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

int launch_mdot_hex27(const Graph &g, const double *coords, const double *density, const double *velocity, double *mdot,
                      cudaStream_t stream, const double *pressure, const double *gpdx, double tau, long long e0, long long e1)
{
  using namespace hex27;
  const Tables *T = nullptr;
  int rc = get_tables(&T);
  if (rc) return rc;
  if (e1 < 0 || e1 > g.n_elems) e1 = g.n_elems;
  if (e0 < 0) e0 = 0;
  const long long ne = e1 - e0, nt = ne * NIP;
  if (nt > 0) {
    mdot_hex27_kernel<<<(unsigned)((nt + kThreads - 1) / kThreads), kThreads, 0, stream>>>(*T, e0, ne, g.d_elem_nodes27, coords, density, velocity,
                                                                                          pressure, gpdx, tau, mdot);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}

} // namespace nalu_b200

namespace nalu_b200 {
int launch_nodal_grad_hex27(const Graph &g, const double *coords, const double *field, int ncomp, const double *dual_volume, double *grad,
                            cudaStream_t stream)
{
  using namespace hex27;
  if (ncomp != 1 && ncomp != 3) {
    set_error("assemble_nodal_grad_elem: ncomp must be 1 (scalar) or 3 (vector)");
    return NALU_B200_ERR_ARG;
  }
  const Tables *T = nullptr;
  int rc = get_tables(&T);
  if (rc) return rc;
  const long long nt = g.n_elems * NIP;
  if (nt > 0) {
    const unsigned nb = (unsigned)((nt + kThreads - 1) / kThreads);
    if (ncomp == 1) nodal_grad_hex27_kernel<1><<<nb, kThreads, 0, stream>>>(*T, g.n_elems, g.d_elem_nodes27, coords, field, dual_volume, grad);
    else nodal_grad_hex27_kernel<3><<<nb, kThreads, 0, stream>>>(*T, g.n_elems, g.d_elem_nodes27, coords, field, dual_volume, grad);
    count_launch();
  }
  NALU_CUDA(cudaGetLastError());
  return NALU_B200_OK;
}
} // namespace nalu_b200
