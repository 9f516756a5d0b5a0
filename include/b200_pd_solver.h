/*
 * b200_pd_solver.h -- the reference-side adapter: `class B200PdSolver : public Solver<float>`.
 *
 * Header-only C++17, compiled INSIDE the reference tree (it includes the reference's own
 * src/def.h and src/simulation/solver/solver.h) and linked against libpd_b200.so through the C ABI
 * of include/pd_b200.h.  It replaces `PdSolver` at its single construction site
 *     src/simulation/simulationContext.cu:120   solver = std::make_unique<PdSolver>(threadsPerBlock, data);
 * and is driven by the unchanged caller
 *     src/simulation/simulationContext.cpp:86   impl.solver->Update(impl.data, impl.params);
 * INTEGRATION.md shows the three-line patch.  Same names, argument meaning and error behaviour as
 * PdSolver (src/simulation/solver/projective/pdSolver.h:12-44): void returns, nothing thrown across
 * Update(); failures are printed like the reference's CHECK_* macros do (linear.h:16-49) and leave
 * SolverData untouched.
 *
 * Data flow per Update:
 *   first Update after construction/Reset : D2H of the rest data (X0, Tet, mass, mu, DBC) -> pd_scene_from_desc ->
 *       pd_create (device layout is built once), DBCX <- X0 like SolverPrepare (pdSolver.cu:134)
 *   every Update : pd_set_params (CopyUIToParams runs before every Update, simulationContext.cpp:85)
 *       -> pd_update_device(engine, 1, data.X, data.V, data.XTilde): AoS glm::vec3 import, one PD step (with
 *       SolverParams::handleCollision: the mesh-mesh collision pass of pdSolver.cu:218-225 on data.Tri / dev_TriFathers, inside
 *       the engine, before the fixed bodies -- as in PdSolver::Update), AoS export, all on the engine's stream, synchronised on
 *       return (the renderer reads data.X next).
 *       While data.mouseSelection.dragging: pd_set_drag_device(data.moreDBC, data.OffsetX, mouseSelection.target) first.
 *       (data.DBCX is not written back: nothing but the PD solver itself reads it.)
 * The fixed bodies are passed as plain structs; B200_FIXED_BODY_FROM shows how to fill one from a FixedBody*.
 */
#pragma once

#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include <def.h>                          /* SolverData, SolverParams, indexType  (src/def.h) */
#include <simulation/solver/solver.h>     /* template<typename Scalar> class Solver (solver.h:11-28) */

#include "pd_b200.h"

/* Fill a pd_fixed_body from one of the reference's FixedBody subclasses (collision/rigid/*.h):
 *   B200_FIXED_BODY_FROM(fb, PD_PLANE,    plane->m_model,    0.f)
 *   B200_FIXED_BODY_FROM(fb, PD_SPHERE,   sphere->m_model,   sphere->m_radius)
 *   B200_FIXED_BODY_FROM(fb, PD_CYLINDER, cylinder->m_model, cylinder->m_radius)
 * glm::mat4 is column-major float[16], which is exactly pd_fixed_body::model. */
#define B200_FIXED_BODY_FROM(out, kind, glm_mat4_model, r)                          \
    do {                                                                            \
        (out).type = (kind);                                                        \
        const float* m__ = &(glm_mat4_model)[0][0];                                 \
        for (int i__ = 0; i__ < 16; ++i__) (out).model[i__] = m__[i__];             \
        (out).radius = (r);                                                         \
    } while (0)

class B200PdSolver : public Solver<float> {
public:
    /* PdSolver::SolverType (pdSolver.h:15-18); PCGJacobi is the extra back-end of linear/pcgJacobi.cu.
     * EigenCholesky runs on the engine's own sparse Cholesky like CuSolverCholesky does.  One difference from the reference, kept
     * on purpose: its Eigen branch never updates `err` (pdSolver.cu:178-185 vs :186-192), so there the PD loop always runs
     * numIterations; here both direct modes leave the loop once sqrt(err) < tol, as the reference's cuSOLVER branch does --
     * set SolverParams::tol = 0 to get the Eigen branch's iteration count. */
    enum class SolverType { Jacobi = PD_JACOBI, CuSolverCholesky = PD_CHOLESKY, EigenCholesky = PD_CHOLESKY, PCGJacobi = PD_PCG_JACOBI };

    /* Same leading arguments as PdSolver(int, const SolverData<float>&) (pdSolver.cu:22-27).  Unlike
     * FEMSolver's ctor (femSolver.cu:6-17) nothing is allocated inside `solverData`: V0/DmInv live
     * in the engine's own tile stream; SolverData::V0/DmInv/ExtForce stay null (only the IPC and
     * explicit solvers read them). */
    B200PdSolver(int threadsPerBlock, const SolverData<float>& solverData, std::vector<pd_fixed_body> fixedBodies = {},
                 int device = 0)
        : Solver<float>(threadsPerBlock), fixed_(std::move(fixedBodies)), device_(device)
    {
        (void)solverData;
        /* the four named counters of pdSolver.cu:26 */
        performanceData = {{"local step", 0.f}, {"global step", 0.f}, {"collision handling(fixed)", 0.f}, {"collision handling(mesh)", 0.f}};
    }
    ~B200PdSolver() override { if (engine_) pd_destroy(engine_); }
    B200PdSolver(const B200PdSolver&) = delete;
    B200PdSolver& operator=(const B200PdSolver&) = delete;

    void SetGlobalSolver(SolverType val) { solverType_ = static_cast<int>(val); }
    /* Call after SimulationCUDAContext::UpdateSoftBodyAttr has refilled SolverData::mu (simulationContext.cu:165-176): PdSolver
     * reads data.mu inside computeLocal every iteration, the engine keeps |V0|*mu in its tile stream.  Like the reference,
     * matrix_diag and the assembled matrix follow only after Reset(). */
    void MuChanged(const SolverData<float>& solverData)
    {
        if (engine_ && solverData.mu && pd_update_mu_device(engine_, solverData.mu) != PD_OK) std::fprintf(stderr, "B200PdSolver: %s\n", pd_last_error());
    }
    /* number of PD iterations per Update; the reference keeps it in SolverParams::numIterations */

    void Update(SolverData<float>& solverData, const SolverParams<float>& solverParams) override
    {
        if (!solverReady) {
            SolverPrepare(solverData, solverParams);
            if (!engine_) return;                    /* printed already; SolverData untouched */
            solverReady = true;
        }
        SolverStep(solverData, solverParams);
    }

    void Reset() override
    {   /* Solver::Reset (solver.h:39-43) + PdSolver::Reset (pdSolver.cu:234-241): SimulationCUDAContext::Reset has
         * already restored X/XTilde/V from X0 in SolverData; the next Update re-prepares (dt, mu may have changed). */
        Solver<float>::Reset();
        for (auto& kv : performanceData) kv.second = 0.f;
        if (engine_) { pd_destroy(engine_); engine_ = nullptr; }
        dragSeen_ = false;
    }

protected:
    void SolverPrepare(SolverData<float>& d, const SolverParams<float>& sp) override
    {
        const size_t nV = (size_t)d.numVerts, nT = (size_t)d.numTets;
        std::vector<float> X(3 * nV), mass(nV), mu(nT), dbc(nV, 0.f);
        std::vector<uint32_t> tet(4 * nT);
        static_assert(sizeof(indexType) == sizeof(uint32_t), "indexType is unsigned int (def.h:4)");
        bool ok = cudaMemcpy(X.data(), d.X0, 12 * nV, cudaMemcpyDeviceToHost) == cudaSuccess;
        ok = ok && cudaMemcpy(tet.data(), d.Tet, 16 * nT, cudaMemcpyDeviceToHost) == cudaSuccess;
        ok = ok && cudaMemcpy(mass.data(), d.mass, 4 * nV, cudaMemcpyDeviceToHost) == cudaSuccess;
        ok = ok && cudaMemcpy(mu.data(), d.mu, 4 * nT, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (d.DBC) ok = ok && cudaMemcpy(dbc.data(), d.DBC, 4 * nV, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (d.DBCX) ok = ok && cudaMemcpy(d.DBCX, d.X0, 12 * nV, cudaMemcpyDeviceToDevice) == cudaSuccess;   /* pdSolver.cu:134 */
        if (!ok) { std::fprintf(stderr, "B200PdSolver: reading SolverData failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
        pd_scene_desc desc{};
        desc.num_verts = d.numVerts; desc.num_tets = d.numTets;
        desc.X = X.data(); desc.Tet = tet.data(); desc.mass = mass.data(); desc.mu = mu.data(); desc.DBC = dbc.data();
        desc.num_fixed = (int)fixed_.size(); desc.fixed = fixed_.data();
        /* surface triangles and their father bodies for the mesh-mesh collision pass (SolverData::Tri / dev_TriFathers,
         * dataLoader.cu:343-369): with them the engine runs DetectCollision + CCDKernel itself (pdSolver.cu:218-225) */
        std::vector<uint32_t> tri, father;
        collisionOk_ = false;
        if (d.numTris > 0 && d.Tri) {
            tri.resize(3 * (size_t)d.numTris); father.assign((size_t)d.numTris, 0u);
            bool okT = cudaMemcpy(tri.data(), d.Tri, 12 * (size_t)d.numTris, cudaMemcpyDeviceToHost) == cudaSuccess;
            if (okT && d.dev_TriFathers) okT = cudaMemcpy(father.data(), d.dev_TriFathers, 4 * (size_t)d.numTris, cudaMemcpyDeviceToHost) == cudaSuccess;
            if (okT) { desc.num_tris = d.numTris; desc.Tri = tri.data(); desc.TriFathers = father.data(); collisionOk_ = true; }
            else cudaGetLastError();
        }
        pd_params p; toParams(sp, &p);
        noteCollision(p);
        pd_scene* scene = pd_scene_from_desc(&desc, &p);
        if (!scene) { std::fprintf(stderr, "B200PdSolver: %s\n", pd_last_error()); return; }
        pd_engine_options opt; pd_default_options(&opt);
        opt.device = device_;
        engine_ = pd_create(scene, &opt);
        pd_scene_free(scene);
        if (!engine_) { std::fprintf(stderr, "B200PdSolver: %s\n", pd_last_error()); return; }
        pd_set_perf(engine_, perf ? 1 : 0);
    }

    bool SolverStep(SolverData<float>& d, const SolverParams<float>& sp) override
    {
        pd_params p; toParams(sp, &p);
        noteCollision(p);
        pd_set_perf(engine_, perf ? 1 : 0);
        /* mouse drag (README "PD solver supports interactive object dragging"): while MouseSelection::dragging the
         * caller's Control_Kernel keeps SolverData::moreDBC / OffsetX up to date (simulationContext.cu:202-231) and
         * RayIntersect the target (:196); when dragging ends main.cpp zeroes moreDBC (ResetMoreDBC(true), main.cpp:95-114).
         * PdSolver reads all three in every SolverStep (pdSolver.cu:156-157,171,194,206). */
        const bool drag = d.mouseSelection.dragging && d.moreDBC && d.OffsetX;
        if (drag || dragSeen_) {
            const int rc = drag ? pd_set_drag_device(engine_, d.moreDBC, &d.OffsetX[0].x, &d.mouseSelection.target.x)
                                : pd_set_drag_device(engine_, nullptr, nullptr, nullptr);
            if (rc != PD_OK) { std::fprintf(stderr, "B200PdSolver: %s\n", pd_last_error()); return false; }
            dragSeen_ = drag;
        }
        if (pd_set_params(engine_, &p) != PD_OK || pd_update_device(engine_, 1, &d.X[0].x, &d.V[0].x, &d.XTilde[0].x) != PD_OK) {
            std::fprintf(stderr, "B200PdSolver: %s\n", pd_last_error());
            return false;
        }
        if (perf) {
            pd_perf pf;
            if (pd_get_perf(engine_, &pf) == PD_OK) {
                performanceData[0].second = pf.local_step_ms; performanceData[1].second = pf.global_step_ms;
                performanceData[2].second = pf.collision_fixed_ms; performanceData[3].second = pf.collision_mesh_ms;
            }
        }
        return true;
    }

private:
    /* SolverParams::handleCollision (the GUI default, context.h:44) asks for the mesh-mesh collision pass of
     * PdSolver::Update (pdSolver.cu:218-225); the engine runs it on SolverData::Tri / dev_TriFathers (SolverPrepare above).
     * Only a SolverData WITHOUT surface triangles cannot: say so ONCE and step with the flag off -- X is then already
     * XTilde on return (pdSolver.cu:227), so the caller cannot run DetectCollision / CCDKernel afterwards either. */
    void noteCollision(pd_params& p)
    {
        if (!p.handle_collision || collisionOk_) return;
        if (!collisionNoted_) {
            std::fprintf(stderr, "B200PdSolver: handleCollision=true but no surface triangles were given to the engine; "
                                 "mesh-mesh collision is OFF for this solver (fixed bodies still respond)\n");
            collisionNoted_ = true;
        }
        p.handle_collision = 0;
    }
    void toParams(const SolverParams<float>& sp, pd_params* p) const
    {
        pd_default_params(p);
        p->dt = sp.dt; p->gravity = sp.gravity; p->muN = sp.muN; p->muT = sp.muT; p->rho = sp.rho; p->tol = sp.tol; p->damp = sp.damp;
        p->num_iterations = (int)sp.numIterations;
        p->global_solver = solverType_;
        p->handle_collision = sp.handleCollision ? 1 : 0;
        p->threads_per_block = threadsPerBlock;
    }

    std::vector<pd_fixed_body> fixed_;
    pd_engine* engine_ = nullptr;
    int solverType_ = PD_JACOBI;
    int device_ = 0;
    bool dragSeen_ = false;
    bool collisionOk_ = false, collisionNoted_ = false;
};
