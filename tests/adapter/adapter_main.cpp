// Drop-in check of include/b200_pd_solver.h: drives `B200PdSolver` exactly the way
// SimulationCUDAContext drives PdSolver (simulationContext.cu:120-122, simulationContext.cpp:85-86,
// simulationContext.cu:233-243), on device arrays laid out like DataLoader::AllocData leaves them
// (dataLoader.cu:291-378), and compares X/XTilde/V with the plain C-ABI path (pd_create/pd_step).
// Built against the REFERENCE's own def.h / solver.h by `make -C oracle adapter` (output under
// oracle/_ref, next to the other artefacts that need /root/reference at build time).
//   adapter_test [cells] [steps]   -> prints "max_abs_diff <d>" and exits 0 iff d == 0
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "b200_pd_solver.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

int main(int argc, char** argv)
{
    const int cells = argc > 1 ? std::atoi(argv[1]) : 6, steps = argc > 2 ? std::atoi(argv[2]) : 3;
    const float origin[3] = {0.f, 2.f, 0.f};
    pd_scene* sc = pd_scene_kuhn_grid(cells, cells, cells, 1.0f, 0.05f, 7u, origin, 1.0f, 2e5f);
    if (!sc) { std::fprintf(stderr, "%s\n", pd_last_error()); return 2; }
    pd_fixed_body floor{};
    floor.type = PD_PLANE;
    for (int i = 0; i < 4; ++i) floor.model[5 * i] = 1.f;
    pd_scene_add_fixed(sc, &floor);
    int nV = 0, nT = 0;
    pd_scene_counts(sc, &nV, &nT, nullptr, nullptr);
    std::vector<float> X(3 * (size_t)nV), mass(nV), mu(nT), dbc(nV);
    std::vector<uint32_t> tet(4 * (size_t)nT);
    pd_scene_get(sc, X.data(), tet.data(), mass.data(), mu.data(), dbc.data(), nullptr, nullptr);

    SolverParams<float> params;
    params.numIterations = 20; params.dt = 1.0f / 60.0f; params.gravity = 98.f; params.handleCollision = false;

    // --- path 1: the C ABI directly
    pd_params p; pd_default_params(&p);
    p.dt = params.dt; p.gravity = params.gravity; p.num_iterations = (int)params.numIterations; p.muN = params.muN; p.muT = params.muT;
    p.rho = params.rho; p.tol = params.tol; p.damp = params.damp; p.handle_collision = 0;
    pd_scene_set_params(sc, &p);
    pd_engine* eng = pd_create(sc, nullptr);
    if (!eng) { std::fprintf(stderr, "%s\n", pd_last_error()); return 2; }
    std::vector<float> X1(X.size()), V1(X.size()), T1(X.size());
    if (pd_step(eng, steps) || pd_download(eng, X1.data(), V1.data(), T1.data())) { std::fprintf(stderr, "%s\n", pd_last_error()); return 2; }
    // ... then two more steps with a mouse drag (a handful of vertices held at target + offset), through pd_set_drag
    std::vector<float> more(nV, 0.f), off(3 * (size_t)nV, 0.f);
    const int pick = nV / 2;
    const float target[3] = {X1[3 * pick] + 0.4f, X1[3 * pick + 1] + 0.3f, X1[3 * pick + 2]};
    int nDrag = 0;
    for (int i = 0; i < nV; ++i) {
        float d2 = 0;
        for (int k = 0; k < 3; ++k) { off[3 * i + k] = X1[3 * i + k] - X1[3 * pick + k]; d2 += off[3 * i + k] * off[3 * i + k]; }
        if (d2 < 1.5f) { more[i] = 10.f; ++nDrag; }
    }
    std::vector<float> X1d(X.size()), V1d(X.size()), T1d(X.size());
    if (pd_set_drag(eng, more.data(), off.data(), target) || pd_step(eng, 2) || pd_download(eng, X1d.data(), V1d.data(), T1d.data())) { std::fprintf(stderr, "%s\n", pd_last_error()); return 2; }
    pd_destroy(eng);
    pd_scene_free(sc);

    // --- path 2: SolverData on the device + the Solver<float> plugin interface
    SolverData<float> d;
    d.numVerts = nV; d.numTets = nT;
    CK(cudaMalloc((void**)&d.X, 12 * (size_t)nV)); CK(cudaMalloc((void**)&d.X0, 12 * (size_t)nV)); CK(cudaMalloc((void**)&d.XTilde, 12 * (size_t)nV));
    CK(cudaMalloc((void**)&d.V, 12 * (size_t)nV)); CK(cudaMalloc((void**)&d.DBCX, 12 * (size_t)nV));
    CK(cudaMalloc((void**)&d.Tet, 16 * (size_t)nT)); CK(cudaMalloc((void**)&d.mass, 4 * (size_t)nV)); CK(cudaMalloc((void**)&d.mu, 4 * (size_t)nT));
    CK(cudaMalloc((void**)&d.DBC, 4 * (size_t)nV));
    CK(cudaMemcpy(d.X, X.data(), 12 * (size_t)nV, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d.X0, X.data(), 12 * (size_t)nV, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.XTilde, X.data(), 12 * (size_t)nV, cudaMemcpyHostToDevice)); CK(cudaMemset(d.V, 0, 12 * (size_t)nV));
    CK(cudaMemcpy(d.Tet, tet.data(), 16 * (size_t)nT, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d.mass, mass.data(), 4 * (size_t)nV, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.mu, mu.data(), 4 * (size_t)nT, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d.DBC, dbc.data(), 4 * (size_t)nV, cudaMemcpyHostToDevice));

    std::unique_ptr<Solver<float>> solver = std::make_unique<B200PdSolver>(128, d, std::vector<pd_fixed_body>{floor});
    solver->SetPerf(true);
    // one throw-away Update + Reset, as the GUI's reset button does (simulationContext.cu:233-243)
    solver->Update(d, params);
    CK(cudaMemcpy(d.X, d.X0, 12 * (size_t)nV, cudaMemcpyDeviceToDevice)); CK(cudaMemcpy(d.XTilde, d.X0, 12 * (size_t)nV, cudaMemcpyDeviceToDevice));
    CK(cudaMemset(d.V, 0, 12 * (size_t)nV));
    solver->Reset();
    for (int s = 0; s < steps; ++s) solver->Update(d, params);
    std::vector<float> X2(X.size()), V2(X.size()), T2(X.size());
    CK(cudaMemcpy(X2.data(), d.X, 12 * (size_t)nV, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(V2.data(), d.V, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(T2.data(), d.XTilde, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
    double worst = 0, moved = 0;
    for (size_t i = 0; i < X.size(); ++i) {
        worst = std::fmax(worst, std::fabs((double)X1[i] - X2[i]));
        worst = std::fmax(worst, std::fabs((double)V1[i] - V2[i]));
        worst = std::fmax(worst, std::fabs((double)T1[i] - T2[i]));
        moved = std::fmax(moved, std::fabs((double)X2[i] - X[i]));
    }
    // the same drag through SolverData (what Control_Kernel / RayIntersect leave there, simulationContext.cu:177-231)
    CK(cudaMalloc((void**)&d.moreDBC, 4 * (size_t)nV)); CK(cudaMalloc((void**)&d.OffsetX, 12 * (size_t)nV));
    CK(cudaMemcpy(d.moreDBC, more.data(), 4 * (size_t)nV, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d.OffsetX, off.data(), 12 * (size_t)nV, cudaMemcpyHostToDevice));
    d.mouseSelection.dragging = true; d.mouseSelection.select_v = pick; d.mouseSelection.target = glm::vec3(target[0], target[1], target[2]);
    for (int s = 0; s < 2; ++s) solver->Update(d, params);
    CK(cudaMemcpy(X2.data(), d.X, 12 * (size_t)nV, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(V2.data(), d.V, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(T2.data(), d.XTilde, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
    double worstDrag = 0, held = 0;
    for (size_t i = 0; i < X.size(); ++i) {
        worstDrag = std::fmax(worstDrag, std::fabs((double)X1d[i] - X2[i]));
        worstDrag = std::fmax(worstDrag, std::fabs((double)V1d[i] - V2[i]));
        worstDrag = std::fmax(worstDrag, std::fabs((double)T1d[i] - T2[i]));
    }
    for (int i = 0; i < nV; ++i)
        if (more[i] > 0.f)
            for (int k = 0; k < 3; ++k) held = std::fmax(held, std::fabs((double)X2[3 * i + k] - (double)(target[k] + off[3 * i + k])) + std::fabs((double)V2[3 * i + k]));
    std::printf("drag: %d vertices held, max_abs_diff %.9g, distance of the held vertices from target+offset %.9g\n", nDrag, worstDrag, held);
    // release (main.cpp:95-99: SetDragging(false) + ResetMoreDBC(true)) and one more Update: must run, nothing held any more
    d.mouseSelection.dragging = false;
    CK(cudaMemset(d.moreDBC, 0, 4 * (size_t)nV));
    solver->Update(d, params);
    CK(cudaMemcpy(V2.data(), d.V, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
    double vfree = 0;
    for (int i = 0; i < nV; ++i) if (more[i] > 0.f) vfree = std::fmax(vfree, std::fabs((double)V2[3 * i + 1]));
    std::printf("release: |V.y| of a formerly held vertex %.4f\n", vfree);
    worst = std::fmax(worst, worstDrag);
    if (held != 0.0 || nDrag == 0 || !(vfree > 0.0)) worst = std::fmax(worst, 1.0);
    // --- SolverParams::handleCollision = true, the reference's default (context.h:44, def.h:79): the adapter hands SolverData::Tri /
    // dev_TriFathers to the engine, which runs the mesh-mesh pass itself; it must step (it used to leave the engine
    // null) and give exactly what pd_step gives with pd_params.handle_collision = 1
    {
        pd_scene* sc2 = pd_scene_kuhn_grid(cells, cells, cells, 1.0f, 0.05f, 7u, origin, 1.0f, 2e5f);
        pd_scene_add_fixed(sc2, &floor);
        pd_params pc = p; pc.handle_collision = 1;
        pd_scene_set_params(sc2, &pc);
        int nTri = 0;
        pd_scene_get_surface(sc2, &nTri, nullptr, nullptr);
        std::vector<uint32_t> tri(3 * (size_t)nTri), father((size_t)nTri);
        pd_scene_get_surface(sc2, &nTri, tri.data(), father.data());
        pd_engine* e2 = pd_create(sc2, nullptr);
        if (!e2) { std::fprintf(stderr, "%s\n", pd_last_error()); return 2; }
        std::vector<float> Xc(X.size()), Vc(X.size()), Tc(X.size());
        if (pd_step(e2, 2) || pd_download(e2, Xc.data(), Vc.data(), Tc.data())) { std::fprintf(stderr, "%s\n", pd_last_error()); return 2; }
        pd_destroy(e2); pd_scene_free(sc2);
        SolverData<float> dc = d;
        dc.moreDBC = nullptr; dc.OffsetX = nullptr; dc.mouseSelection.dragging = false;
        dc.numTris = nTri;
        CK(cudaMalloc((void**)&dc.Tri, 12 * (size_t)nTri)); CK(cudaMalloc((void**)&dc.dev_TriFathers, 4 * (size_t)nTri));
        CK(cudaMemcpy(dc.Tri, tri.data(), 12 * (size_t)nTri, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dc.dev_TriFathers, father.data(), 4 * (size_t)nTri, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dc.X, X.data(), 12 * (size_t)nV, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dc.XTilde, X.data(), 12 * (size_t)nV, cudaMemcpyHostToDevice));
        CK(cudaMemset(dc.V, 0, 12 * (size_t)nV));
        SolverParams<float> pcol = params; pcol.handleCollision = true;
        std::unique_ptr<Solver<float>> sc3 = std::make_unique<B200PdSolver>(128, dc, std::vector<pd_fixed_body>{floor});
        for (int s = 0; s < 2; ++s) sc3->Update(dc, pcol);
        CK(cudaMemcpy(X2.data(), dc.X, 12 * (size_t)nV, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(V2.data(), dc.V, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(T2.data(), dc.XTilde, 12 * (size_t)nV, cudaMemcpyDeviceToHost));
        double wc = 0, mv = 0;
        for (size_t i = 0; i < X.size(); ++i) {
            wc = std::fmax(wc, std::fabs((double)Xc[i] - X2[i])); wc = std::fmax(wc, std::fabs((double)Vc[i] - V2[i])); wc = std::fmax(wc, std::fabs((double)Tc[i] - T2[i]));
            mv = std::fmax(mv, std::fabs((double)X2[i] - X[i]));
        }
        std::printf("handleCollision=true: %d surface triangles, moved %.4f, max_abs_diff %.9g\n", nTri, mv, wc);
        if (!(mv > 0.0)) wc = std::fmax(wc, 1.0);
        worst = std::fmax(worst, wc);
    }
    const auto& perf = solver->GetPerformanceData();
    std::printf("nV %d nT %d steps %d moved %.4f perf[%s]=%.3f ms perf[%s]=%.3f ms\n", nV, nT, steps, moved, perf[0].first.c_str(), perf[0].second,
                perf[1].first.c_str(), perf[1].second);
    std::printf("max_abs_diff %.9g\n", worst);
    return (worst == 0.0 && moved > 0.0 && perf.size() == 4) ? 0 : 1;
}
