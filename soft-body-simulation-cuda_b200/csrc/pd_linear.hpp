// Linear back-ends of the reference's IPC (double) solver (pd_linear.cu): LinearSolver<double> of
// src/simulation/solver/linear/linear.h:55-71 with the PCGJacobiSolver / CGSolver implementations behind it.
#pragma once
#include <memory>

namespace pdb200 {

constexpr int LIN_CG_IC0 = 1;        // CGSolver<double>, linear/cg.cu (IPCSolver::linearSolver[1], IPC/ipc.cu:96)
constexpr int LIN_PCG_JACOBI = 2;    // PCGJacobiSolver<double>, linear/pcgJacobi.cu (IPCSolver::linearSolver[2], IPC/ipc.cu:97)

class LinearSolver {
public:
    // maxIter / tol <= 0: the reference's defaults (2000, 1e-5 for PCG-Jacobi; 100, 1e-6 for CG + IC(0))
    LinearSolver(int kind, int N, int maxIter, double tol, int device);
    ~LinearSolver();
    LinearSolver(const LinearSolver&) = delete;
    LinearSolver& operator=(const LinearSolver&) = delete;
    // LinearSolver<double>::Solve: DEVICE pointers; A / rowIdx / colIdx = COO with duplicates (summed), any order; guess may be null (x0 = 0)
    void solveDevice(int N, const double* b, double* x, const double* A, int nz, const int* rowIdx, const int* colIdx, const double* guess);
    void solveHost(int N, const double* b, double* x, const double* A, int nz, const int* rowIdx, const int* colIdx, const double* guess);
    void stats(int* iterations, double* residual, int* nnz) const;

private:
    struct Impl;
    std::unique_ptr<Impl> d_;
};

}  // namespace pdb200
