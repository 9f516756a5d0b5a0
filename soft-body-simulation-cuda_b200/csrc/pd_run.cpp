// pd_run -- headless C++ host driver over the C ABI (include/pd_b200.h), nothing else.
//
// The reference's frame loop is main.cpp:mainLoop -> Context::Update -> SimulationCUDAContext::Update
// (context.cpp:546-555, simulationContext.cpp:79-86) inside a GLFW/ImGui window.  This is the same loop without the
// window: load a context of context.json (same schema, same TetGen assets), step it, optionally drag a vertex the way
// the mouse does (simulationContext.cu:177-231), and dump frames as TetGen .node files (the asset format the reference's
// own loader reads back, dataLoader.cu:131-173) or raw float32.
//
//   pd_run --json context.json [--context NAME] [--assets DIR] [--steps N] [--iterations K] [--dt DT]
//          [--solver jacobi|cholesky|pcg] [--device D] [--perf] [--out PREFIX] [--every M] [--format node|bin]
//          [--drag V,TX,TY,TZ,FROM,TO] [--info]
//
// --info loads the scene and builds the device layout on the host only (no GPU needed) and prints the counts.
// One JSON line with the timing goes to stdout at the end; errors go to stderr with a non-zero exit code.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pd_b200.h"

namespace {

struct Args {
    std::string json, context, assets, out, format = "node", solver;
    int steps = 100, iterations = -1, device = 0, every = 0;
    float dt = -1.f;
    bool perf = false, info = false;
    bool drag = false; int dragV = -1, dragFrom = 0, dragTo = 0; float dragTarget[3] = {0, 0, 0};
};

int usage(const char* why)
{
    if (why) std::fprintf(stderr, "pd_run: %s\n", why);
    std::fprintf(stderr,
                 "usage: pd_run --json context.json [--context NAME] [--assets DIR] [--steps N] [--iterations K] [--dt DT]\n"
                 "              [--solver jacobi|cholesky|pcg] [--device D] [--perf] [--out PREFIX] [--every M] [--format node|bin]\n"
                 "              [--drag V,TX,TY,TZ,FROM,TO] [--info]\n");
    return 64;
}

bool write_node(const std::string& path, const std::vector<float>& X)
{   // TetGen .node: "<n> 3 0 0" then "<index> x y z" (1-based, like the shipped assets)
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    const size_t n = X.size() / 3;
    std::fprintf(f, "%zu  3  0  0\n", n);
    for (size_t i = 0; i < n; ++i) std::fprintf(f, "%zu  %.9g  %.9g  %.9g\n", i + 1, X[3 * i], X[3 * i + 1], X[3 * i + 2]);
    return std::fclose(f) == 0;
}

bool write_bin(const std::string& path, const std::vector<float>& X, const std::vector<float>& V)
{   // raw little-endian float32: X[3n] then V[3n]
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(X.data(), 4, X.size(), f) == X.size() && std::fwrite(V.data(), 4, V.size(), f) == V.size();
    return (std::fclose(f) == 0) && ok;
}

}  // namespace

int main(int argc, char** argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        const std::string k = argv[i];
        auto val = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : nullptr; };
        const char* v = nullptr;
        if (k == "--perf") a.perf = true;
        else if (k == "--info") a.info = true;
        else if (k == "--help" || k == "-h") return usage(nullptr);
        else if (!(v = val())) return usage(("missing value for " + k).c_str());
        else if (k == "--json") a.json = v;
        else if (k == "--context") a.context = v;
        else if (k == "--assets") a.assets = v;
        else if (k == "--out") a.out = v;
        else if (k == "--format") a.format = v;
        else if (k == "--solver") a.solver = v;
        else if (k == "--steps") a.steps = std::atoi(v);
        else if (k == "--iterations") a.iterations = std::atoi(v);
        else if (k == "--device") a.device = std::atoi(v);
        else if (k == "--every") a.every = std::atoi(v);
        else if (k == "--dt") a.dt = (float)std::atof(v);
        else if (k == "--drag") {
            if (std::sscanf(v, "%d,%f,%f,%f,%d,%d", &a.dragV, &a.dragTarget[0], &a.dragTarget[1], &a.dragTarget[2], &a.dragFrom, &a.dragTo) != 6)
                return usage("--drag wants V,TX,TY,TZ,FROM,TO");
            a.drag = true;
        } else return usage(("unknown option " + k).c_str());
    }
    if (a.json.empty()) return usage("--json is required");
    if (a.steps < 0 || (a.format != "node" && a.format != "bin")) return usage("bad --steps / --format");
    int solver = -1;
    if (a.solver == "jacobi") solver = PD_JACOBI;
    else if (a.solver == "cholesky") solver = PD_CHOLESKY;
    else if (a.solver == "pcg") solver = PD_PCG_JACOBI;
    else if (!a.solver.empty()) return usage("--solver is one of jacobi, cholesky, pcg");

    // Context::LoadSimContext (context.cpp:319-387): soft bodies, fixed bodies, dt / gravity / friction of the named context
    pd_scene* sc = pd_scene_load_json(a.json.c_str(), a.context.empty() ? nullptr : a.context.c_str(), a.assets.empty() ? nullptr : a.assets.c_str());
    if (!sc) { std::fprintf(stderr, "pd_run: %s\n", pd_last_error()); return 1; }
    pd_params p;
    pd_scene_get_params(sc, &p);
    if (a.iterations > 0) p.num_iterations = a.iterations;
    if (a.dt > 0.f) p.dt = a.dt;
    if (solver >= 0) p.global_solver = solver;
    p.handle_collision = 0;                       // mesh-mesh BVH/CCD is outside the PD hot path (context.json has no switch for it)
    if (pd_scene_set_params(sc, &p) != PD_OK) { std::fprintf(stderr, "pd_run: %s\n", pd_last_error()); return 1; }
    int nV = 0, nT = 0, nF = 0, nB = 0;
    pd_scene_counts(sc, &nV, &nT, &nF, &nB);

    if (a.info) {
        pd_layout* L = pd_layout_build(sc, 1);
        if (!L) { std::fprintf(stderr, "pd_run: %s\n", pd_last_error()); return 1; }
        int nTiles = 0, maxLocal = 0; uint32_t nSlots = 0; size_t recBytes = 0;
        pd_layout_counts(L, &nTiles, &nSlots, &recBytes, &maxLocal);
        std::printf("{\"num_verts\": %d, \"num_tets\": %d, \"num_fixed\": %d, \"num_bodies\": %d, \"dt\": %.9g, \"gravity\": %.9g, "
                    "\"num_iterations\": %d, \"global_solver\": %d, \"num_tiles\": %d, \"num_slots\": %u, \"tile_stream_bytes\": %zu}\n",
                    nV, nT, nF, nB, p.dt, p.gravity, p.num_iterations, p.global_solver, nTiles, nSlots, recBytes);
        pd_layout_free(L);
        pd_scene_free(sc);
        return 0;
    }
    if (a.drag && (a.dragV < 0 || a.dragV >= nV)) { pd_scene_free(sc); return usage("--drag: vertex out of range"); }

    pd_scene_free(sc);
    pd_engine_options opt;
    pd_default_options(&opt);
    opt.device = a.device;
    // Impl::Init + PdSolver ctor (simulationContext.cu:34-123); refuses contexts whose "precision" is not float (those run IPC)
    pd_engine* e = pd_create_from_json(a.json.c_str(), a.context.empty() ? nullptr : a.context.c_str(), a.assets.empty() ? nullptr : a.assets.c_str(), &opt);
    if (!e) { std::fprintf(stderr, "pd_run: %s\n", pd_last_error()); return 2; }
    if (pd_set_params(e, &p) != PD_OK) { std::fprintf(stderr, "pd_run: %s\n", pd_last_error()); pd_destroy(e); return 2; }   // CopyUIToParams before the first Update
    pd_set_perf(e, a.perf ? 1 : 0);

    std::vector<float> X(3 * (size_t)nV), V(3 * (size_t)nV);
    auto dump = [&](int frame) -> bool {
        if (a.out.empty()) return true;
        if (pd_download(e, X.data(), V.data(), nullptr) != PD_OK) return false;
        char name[32];
        std::snprintf(name, sizeof name, ".%05d.%s", frame, a.format.c_str());
        return a.format == "node" ? write_node(a.out + name, X) : write_bin(a.out + name, X, V);
    };
    bool ok = dump(0);
    bool dragging = false;
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < a.steps && ok; ++s) {
        if (a.drag) {       // mouse press at step FROM (Control_Kernel on the current X), release at step TO (ResetMoreDBC(true))
            if (s == a.dragFrom) { ok = pd_drag_select(e, a.dragV, 10.0f, a.dragTarget) == PD_OK; dragging = ok; }
            if (s == a.dragTo && dragging) { ok = pd_set_drag(e, nullptr, nullptr, nullptr) == PD_OK; dragging = false; }
        }
        ok = ok && pd_step(e, 1) == PD_OK;        // SimulationCUDAContext::Update
        if (ok && a.every > 0 && (s + 1) % a.every == 0) ok = dump(s + 1);
    }
    ok = ok && pd_synchronize(e) == PD_OK;
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (ok && (a.every <= 0 || a.steps % a.every != 0)) ok = dump(a.steps);
    pd_perf pf{};
    ok = ok && pd_get_perf(e, &pf) == PD_OK;
    if (!ok) { std::fprintf(stderr, "pd_run: %s\n", pd_last_error()); pd_destroy(e); return 3; }
    ok = pd_download(e, X.data(), V.data(), nullptr) == PD_OK;
    double ymin = 1e30, vmax = 0; bool finite = true;
    for (size_t i = 0; i < X.size(); ++i) {
        if (!(X[i] == X[i]) || X[i] > 1e30f || X[i] < -1e30f) finite = false;
        if (i % 3 == 1 && X[i] < ymin) ymin = X[i];
        const double av = V[i] < 0 ? -V[i] : V[i];
        if (av > vmax) vmax = av;
    }
    const double its = (double)pf.pd_iterations;
    std::printf("{\"num_verts\": %d, \"num_tets\": %d, \"steps\": %d, \"pd_iterations\": %lld, \"inner_iterations\": %lld, \"kernel_launches\": %lld, "
                "\"wall_s\": %.6f, \"ms_per_step\": %.4f, \"mtet_updates_per_s\": %.1f, \"finite\": %s, \"y_min\": %.6g, \"v_max\": %.6g, "
                "\"perf_ms\": {\"local step\": %.3f, \"global step\": %.3f, \"collision handling(fixed)\": %.3f, \"collision handling(mesh)\": %.3f}}\n",
                nV, nT, a.steps, pf.pd_iterations, pf.inner_iterations, pf.kernel_launches, wall, a.steps ? 1e3 * wall / a.steps : 0.0,
                wall > 0 ? (double)nT * its / wall / 1e6 : 0.0, finite ? "true" : "false", ymin, vmax,
                pf.local_step_ms, pf.global_step_ms, pf.collision_fixed_ms, pf.collision_mesh_ms);
    pd_destroy(e);
    return (ok && finite) ? 0 : 4;
}
