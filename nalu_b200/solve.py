"""Solver hand-off: the consumer on the other side of the assembly path.

The reference hands the assembled Tpetra matrix to a Belos / MueLu solver (TpetraLinearSystem::solve,
src/TpetraLinearSystem.C:1841-1926 -> TpetraLinearSolver::solve, include/LinearSolver.h:184-253).  Here the device-resident
CSR that the assembly kernels wrote (owned rows: rowptr / colidx from the graph, values from the kernels) is wrapped -- values
without a copy -- as a torch sparse CSR tensor, whose products run on cuSPARSE, and solved with a Jacobi (diagonal)
preconditioned BiCGStab: the role the reference's momentum / scalar solver decks give to GMRES + SGS.  It is a consumer of
the path, not part of it: it exists to show that the matrix the kernels leave in HBM is usable in place, and it gives the
parity tests an end-to-end statement (the solution of the device-assembled system equals the solution of the oracle's).

One rank only: on several ranks the owned matrix has columns for shared and remote nodes, so a product needs the
neighbours' vector entries (Tpetra's Import); that exchange is not built.
"""
import torch


class DeviceCsrOperator:
    """y = A x with the owned rows of a DeviceCsrLinearSystem (cuSPARSE through torch.sparse_csr)."""

    def __init__(self, ls):
        if ls.mesh.nranks != 1:
            raise NotImplementedError("DeviceCsrOperator: the vector import of a multi-rank product is not built")
        rowptr, colidx = ls.device_graph_tensors(False)
        n = ls.n_owned_rows
        self.n = n
        self.values = ls.owned_values()
        # torch wants both index arrays in one integer type; the values tensor is the assembly's own storage
        self.A = torch.sparse_csr_tensor(rowptr, colidx.to(torch.int64), self.values, size=(n, n))
        rows = torch.repeat_interleave(torch.arange(n, device=rowptr.device), rowptr[1:] - rowptr[:-1])
        diag = torch.zeros(n, dtype=torch.float64, device=rowptr.device)
        on_diag = colidx.to(torch.int64) == rows
        diag[rows[on_diag]] = self.values[on_diag]
        self.diag = diag

    def matvec(self, x):
        return torch.mv(self.A, x)


def bicgstab(op, b, x0=None, tol=1e-10, max_iter=1000, precond_diag=None):
    """Right-preconditioned BiCGStab, M = diag(A).  Returns (x, iterations, relative residual ||b - A x|| / ||b||).
    All vectors stay on the device; the scalars of the recurrence are read back once per iteration for the tests of
    convergence and breakdown."""
    minv = 1.0 / (op.diag if precond_diag is None else precond_diag)
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b - op.matvec(x)
    bnorm = float(torch.linalg.vector_norm(b))
    if bnorm == 0.0:
        return torch.zeros_like(b), 0, 0.0
    rhat = r.clone()
    rho = alpha = omega = 1.0
    v = torch.zeros_like(b)
    p = torch.zeros_like(b)
    res = float(torch.linalg.vector_norm(r)) / bnorm
    it = 0
    while res > tol and it < max_iter:
        it += 1
        rho_new = float(torch.dot(rhat, r))
        if rho_new == 0.0:
            break
        beta = (rho_new / rho) * (alpha / omega)
        p = r + beta * (p - omega * v)
        y = minv * p
        v = op.matvec(y)
        alpha = rho_new / float(torch.dot(rhat, v))
        s = r - alpha * v
        if float(torch.linalg.vector_norm(s)) / bnorm <= tol:
            x = x + alpha * y
            r = s
            res = float(torch.linalg.vector_norm(r)) / bnorm
            break
        z = minv * s
        t = op.matvec(z)
        omega = float(torch.dot(t, s)) / float(torch.dot(t, t))
        x = x + alpha * y + omega * z
        r = s - omega * t
        rho = rho_new
        res = float(torch.linalg.vector_norm(r)) / bnorm
    # true residual, not the recurrence's
    res = float(torch.linalg.vector_norm(b - op.matvec(x))) / bnorm
    return x, it, res


def solve(ls, tol=1e-10, max_iter=1000):
    """LinearSystem::solve for a one-rank DeviceCsrLinearSystem: A delta = rhs on the device."""
    op = DeviceCsrOperator(ls)
    return bicgstab(op, ls.owned_rhs(), tol=tol, max_iter=max_iter)
