// TEST INFRASTRUCTURE ONLY.  The product's kernels (csrc/pd_kernels.cuh, rotation.cuh) compiled for the HOST with
// -DPD_HOST_EMU (tests/emu/host_emu.hpp) and driven the way Engine::enqueuePredict / enqueueIteration / enqueueFinish
// (csrc/pd_engine.cu) drive them on the GPU, on the device layout built by the product's own layout.cpp.  One or several
// ranks in lock step (the halo push is a host copy along the plan's push lists, as k_halo_push does it).  Used by
// tests/test_kernel_emulation.py to check the kernels' logic against the oracle without a GPU, including the
// experiments (PD_H_PLANES at compile time, PD_DIST_TRIM at run time).  Nothing here is linked into libpd_b200.so.
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../soft-body-simulation-cuda_b200/csrc/pd_body_kernel.cuh"

using namespace pdb200;

namespace {

struct Rank {
    Layout L;
    RankPlan P;
    int nV = 0, nOwn = 0;
    std::vector<uint32_t> tileTab;
    std::vector<float4> Pslots, q[3], b0, X, V, XT, X0, dbcx, offX;
    std::vector<float2> cc;
    std::vector<float> mass, dbc, md, more;
    // exchange state of the in-kernel halo push (concurrent mode): what Engine::setPeers builds in the ranks' windows
    std::vector<unsigned long long> flags;        // written by the peers, indexed by rank
    unsigned long long epoch = 0;
    unsigned int ticket = 0, status = 0;
    std::vector<uint32_t> pushNbr;                // neighbour slot of every push-list entry
    std::vector<float4*> peerQ;                   // [3 * nNbr]: buffer k of neighbour j
    std::vector<unsigned long long*> peerFlag;    // [nNbr]: this rank's entry in neighbour j's flag array
    long long phase = 0;
};

struct Emu {
    int nVg = 0, world = 1, rotMode = 0, grid = 3, numDBC = 0;
    bool drag = false;
    float target[3] = {0, 0, 0};
    float dt2Prepared = 0.f;
    bool ready = false;
    std::vector<float> fb;
    DevFixedBodies dfb{};
    std::vector<std::unique_ptr<Rank>> ranks;
    std::string err;
    bool bodyMode = false;          // PD_BODY_KERNEL experiment: one emulated CTA per body, the whole step in one launch
    BodyBatch bb;
};

void push(Emu& e, int buf)
{   // k_halo_push: q_peer[dst] = q[src]
    if (e.world == 1) return;
    for (auto& r : e.ranks)
        for (size_t i = 0; i < r->P.pushSrc.size(); ++i)
            e.ranks[(size_t)r->P.pushRank[i]]->q[buf][r->P.pushDst[i]] = r->q[buf][r->P.pushSrc[i]];
}

template <int RM, bool BASE>
void step_once(Emu& e, float dt, float gravity, float rho, float muN, float muT, int iters)
{
    const float dtInv = 1.0f / dt, wdbc = 1e6f * (dtInv * dtInv);
    const int vb = 256;
    for (auto& rp : e.ranks) {
        Rank& r = *rp;
        const unsigned vg = (unsigned)((r.nOwn + vb - 1) / vb);
        const DragArgs dr{r.more.data(), r.offX.data(), r.dbcx.data(), e.target[0], e.target[1], e.target[2], e.numDBC > 0 ? 1 : 0};
        if (e.drag) pd_emu::launch_flat(vg, vb, k_predict<true>, r.nOwn, r.X.data(), r.V.data(), r.mass.data(), r.dbc.data(), r.md.data(), dt, e.dt2Prepared,
                                        gravity, r.q[0].data(), r.q[2].data(), r.b0.data(), r.cc.data(), dr, (unsigned long long*)nullptr);
        else pd_emu::launch_flat(vg, vb, k_predict<false>, r.nOwn, r.X.data(), r.V.data(), r.mass.data(), r.dbc.data(), r.md.data(), dt, e.dt2Prepared,
                                 gravity, r.q[0].data(), r.q[2].data(), r.b0.data(), r.cc.data(), DragArgs{}, (unsigned long long*)nullptr);
    }
    push(e, 0); push(e, 2);
    float omega = 1.0f;
    for (int i = 0; i < iters; ++i) {
        const int ic = i % 3, in = (i + 1) % 3, ip = (i + 2) % 3;
        if (i <= 10) omega = 1;
        else if (i == 11) omega = 2 / (2 - rho * rho);
        else omega = 4 / (4 - rho * rho * omega);
        for (auto& rp : e.ranks) {
            Rank& r = *rp;
            const unsigned vg = (unsigned)((r.nOwn + vb - 1) / vb);
            const unsigned grid = (unsigned)std::min(r.L.nTiles, e.grid);
            DistWait dw{};                  // lock step: the pushes are host copies between the launches, nobody waits
            dw.firstTile = 0x7fffffff;
            pd_emu::launch(grid, (unsigned)TILE_T, LOCAL_SMEM_BYTES, k_local<RM, true, false>, (const uint8_t*)r.L.records.data(), (const uint32_t*)r.tileTab.data(),
                           r.L.nTiles, (const uint32_t*)r.L.vstage.data(), (const uint32_t*)r.L.vlist.data(), (const float4*)r.q[ic].data(),
                           (const float4*)r.b0.data(), r.Pslots.data(), (unsigned long long*)nullptr, dw);
            if (e.drag) pd_emu::launch_flat(vg, vb, k_vertex_jacobi<BASE, true>, r.nOwn, (const float4*)r.q[ic].data(), (const float4*)r.q[ip].data(), r.q[in].data(),
                                            (const float4*)r.dbcx.data(), (const float4*)r.b0.data(), (const float2*)r.cc.data(), (const uint32_t*)r.L.vslotPtr.data(),
                                            (const uint32_t*)r.L.vslot.data(), (const float4*)r.Pslots.data(), omega, wdbc, 0, (unsigned long long*)nullptr);
            else pd_emu::launch_flat(vg, vb, k_vertex_jacobi<BASE, false>, r.nOwn, (const float4*)r.q[ic].data(), (const float4*)r.q[ip].data(), r.q[in].data(),
                                     (const float4*)r.dbcx.data(), (const float4*)r.b0.data(), (const float2*)r.cc.data(), (const uint32_t*)r.L.vslotPtr.data(),
                                     (const uint32_t*)r.L.vslot.data(), (const float4*)r.Pslots.data(), omega, wdbc, 0, (unsigned long long*)nullptr);
        }
        push(e, in);
    }
    for (auto& rp : e.ranks) {
        Rank& r = *rp;
        const unsigned vg = (unsigned)((r.nOwn + vb - 1) / vb);
        if (e.drag) pd_emu::launch_flat(vg, vb, k_finish<true>, r.nOwn, (const float4*)r.q[iters % 3].data(), dtInv, r.X.data(), r.XT.data(), r.V.data(), e.dfb, muT, muN,
                                        (const float*)r.more.data());
        else pd_emu::launch_flat(vg, vb, k_finish<false>, r.nOwn, (const float4*)r.q[iters % 3].data(), dtInv, r.X.data(), r.XT.data(), r.V.data(), e.dfb, muT, muN,
                                 (const float*)nullptr);
    }
}

}  // namespace

// One rank's whole step with the halo push INSIDE the local kernel (DistWait: push-list slices by all CTAs, every position with
// the phase's tag in its fourth lane; the tag checked on the staged ghosts of the boundary tiles), as Engine::step drives it on one GPU per process.  The ranks run
// concurrently, one driving thread each, all CTAs of a launch resident at once.
template <int RM, bool BASE>
void rank_step_concurrent(Emu& e, Rank& r, float dt, float gravity, float rho, float muN, float muT, int iters)
{
    const float dtInv = 1.0f / dt, wdbc = 1e6f * (dtInv * dtInv);
    const int vb = 256;
    const unsigned vg = (unsigned)((r.nOwn + vb - 1) / vb);
    const int base = (int)(r.phase % 3);
    pd_emu::launch_flat(vg, vb, k_predict<false>, r.nOwn, r.X.data(), r.V.data(), r.mass.data(), r.dbc.data(), r.md.data(), dt, e.dt2Prepared,
                        gravity, r.q[base].data(), r.q[(base + 2) % 3].data(), r.b0.data(), r.cc.data(), DragArgs{}, &r.epoch);
    ++r.phase;
    float omega = 1.0f;
    const int nNbr = (int)r.P.neighbours.size();
    for (int i = 0; i < iters; ++i) {
        const int ic = (base + i) % 3, in = (base + i + 1) % 3, ip = (base + i + 2) % 3;
        if (i <= 10) omega = 1;
        else if (i == 11) omega = 2 / (2 - rho * rho);
        else omega = 4 / (4 - rho * rho * omega);
        DistWait dw{};
        dw.epoch = &r.epoch; dw.nNbr = nNbr; dw.nOwn = r.nOwn; dw.firstTile = r.P.nInteriorTiles;
        dw.status = &r.status; dw.nPush = (int)r.P.pushSrc.size(); dw.pushSrc = r.P.pushSrc.data(); dw.pushDst = r.P.pushDst.data();
        dw.pushNbr = r.pushNbr.data(); dw.peerQ = r.peerQ.data() + (size_t)ic * nNbr;
        const unsigned grid = (unsigned)std::min(r.L.nTiles, e.grid);
        pd_emu::launch_resident(grid, (unsigned)TILE_T, LOCAL_SMEM_BYTES, k_local<RM, true, false>, (const uint8_t*)r.L.records.data(), (const uint32_t*)r.tileTab.data(),
                                r.L.nTiles, (const uint32_t*)r.L.vstage.data(), (const uint32_t*)r.L.vlist.data(), (const float4*)r.q[ic].data(),
                                (const float4*)r.b0.data(), r.Pslots.data(), (unsigned long long*)nullptr, dw);
        pd_emu::launch_flat(vg, vb, k_vertex_jacobi<BASE, false>, r.nOwn, (const float4*)r.q[ic].data(), (const float4*)r.q[ip].data(), r.q[in].data(),
                            (const float4*)r.dbcx.data(), (const float4*)r.b0.data(), (const float2*)r.cc.data(), (const uint32_t*)r.L.vslotPtr.data(),
                            (const uint32_t*)r.L.vslot.data(), (const float4*)r.Pslots.data(), omega, wdbc, 0, &r.epoch);
        ++r.phase;
    }
    pd_emu::launch_flat(vg, vb, k_finish<false>, r.nOwn, (const float4*)r.q[(base + iters) % 3].data(), dtInv, r.X.data(), r.XT.data(), r.V.data(), e.dfb, muT, muN,
                        (const float*)nullptr);
}

extern "C" {

const char* emu_variant(void) { return "default"; }

// planes: 6 floats each (p0, up); spheres: 4 (c, r); cylinders: 7 (c, axis, r) -- the arrays the collision kernels consume
void* emu_create(int nV, int nT, const float* X, const uint32_t* Tet, const float* mass, const float* mu, const float* DBC,
                 int nPlanes, const float* planes, int nSpheres, const float* spheres, int nCyls, const float* cyls,
                 int rotMode, int reorder, int world, int trim, int grid)
{
    auto e = std::make_unique<Emu>();
    try {
        e->nVg = nV; e->world = world; e->rotMode = rotMode; e->grid = grid > 0 ? grid : 3;
        for (int i = 0; i < nV; ++i) if (DBC && DBC[i] > 0.f) e->numDBC++;
        e->fb.insert(e->fb.end(), planes, planes + 6 * (size_t)nPlanes);
        e->fb.insert(e->fb.end(), spheres, spheres + 4 * (size_t)nSpheres);
        e->fb.insert(e->fb.end(), cyls, cyls + 7 * (size_t)nCyls);
        e->fb.push_back(0.f);
        e->dfb.nPlanes = nPlanes; e->dfb.nSpheres = nSpheres; e->dfb.nCyls = nCyls;
        e->dfb.planes = e->fb.data(); e->dfb.spheres = e->fb.data() + 6 * (size_t)nPlanes; e->dfb.cyls = e->dfb.spheres + 4 * (size_t)nSpheres;
        Layout G;
        build_layout(nV, nT, X, Tet, mu, reorder != 0, G);
        for (int rk = 0; rk < world; ++rk) {
            auto r = std::make_unique<Rank>();
            if (world == 1) { r->L = G; r->nOwn = G.nV; }
            else {
                build_rank_plan(G, world, rk, r->P, trim != 0);
                extract_rank_layout(G, r->P, r->L);
                r->nOwn = r->P.nOwn;
            }
            const int n = r->nV = r->L.nV;
            build_tile_table(r->L, r->tileTab);
            r->Pslots.assign((size_t)r->L.nTiles * TILE_NLMAX, make_float4(0, 0, 0, 0));
            for (auto& qb : r->q) qb.assign((size_t)n, make_float4(0, 0, 0, 0));
            r->b0.assign((size_t)n, make_float4(0, 0, 0, 0)); r->V = r->b0; r->offX = r->b0;
            r->X.resize((size_t)n); r->cc.assign((size_t)n, float2{0, 0});
            r->mass.resize((size_t)n); r->dbc.resize((size_t)n); r->more.assign((size_t)n, 0.f);
            for (int v = 0; v < n; ++v) {
                const uint32_t o = r->L.vertOrder[(size_t)v];
                r->mass[(size_t)v] = mass[o]; r->dbc[(size_t)v] = DBC ? DBC[o] : 0.f;
                r->X[(size_t)v] = make_float4(X[3 * (size_t)o], X[3 * (size_t)o + 1], X[3 * (size_t)o + 2], 0.f);
            }
            r->XT = r->X; r->X0 = r->X; r->dbcx = r->X;
            e->ranks.push_back(std::move(r));
        }
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "emu_create: %s\n", ex.what());
        return nullptr;
    }
    return e.release();
}

// world > 1: every rank steps on its own thread with the in-kernel halo exchange (instead of emu_step's lock step)
int emu_step_concurrent(void* h, float dt, float gravity, float rho, float muN, float muT, int iters, int nSteps)
{
    Emu& e = *static_cast<Emu*>(h);
    if (e.world < 2 || e.drag || e.bodyMode) return 1;
    if (!e.ready) {
        for (auto& r : e.ranks) matrix_diag_host(r->L, r->md);
        e.dt2Prepared = dt * dt;
        e.ready = true;
    }
    for (auto& rp : e.ranks) {          // Engine::setPeers
        Rank& r = *rp;
        if (!r.flags.empty()) continue;
        r.flags.assign((size_t)e.world, 0ull);
        const int nNbr = (int)r.P.neighbours.size();
        std::vector<int> slotOfRank((size_t)e.world, -1);
        for (int j = 0; j < nNbr; ++j) slotOfRank[(size_t)r.P.neighbours[(size_t)j]] = j;
        r.pushNbr.resize(r.P.pushSrc.size());
        for (size_t i = 0; i < r.P.pushSrc.size(); ++i) r.pushNbr[i] = (uint32_t)slotOfRank[(size_t)r.P.pushRank[i]];
        r.peerQ.assign((size_t)3 * std::max(nNbr, 1), nullptr);
        r.peerFlag.assign((size_t)std::max(nNbr, 1), nullptr);
    }
    for (auto& rp : e.ranks) {          // (second pass: every rank's flag array exists now)
        Rank& r = *rp;
        const int nNbr = (int)r.P.neighbours.size();
        for (int j = 0; j < nNbr; ++j) {
            Rank& peer = *e.ranks[(size_t)r.P.neighbours[(size_t)j]];
            for (int k = 0; k < 3; ++k) r.peerQ[(size_t)k * nNbr + j] = peer.q[k].data();
            r.peerFlag[(size_t)j] = &peer.flags[(size_t)r.P.rank];
        }
    }
    std::vector<std::thread> drivers;
    std::vector<int> rc((size_t)e.world, 0);
    for (int rk = 0; rk < e.world; ++rk)
        drivers.emplace_back([&, rk] {
            try {
                for (int s = 0; s < nSteps; ++s) {
                    if (e.rotMode == 1) rank_step_concurrent<1, false>(e, *e.ranks[(size_t)rk], dt, gravity, rho, muN, muT, iters);
                    else rank_step_concurrent<0, true>(e, *e.ranks[(size_t)rk], dt, gravity, rho, muN, muT, iters);
                }
            } catch (const std::exception& ex) {
                std::fprintf(stderr, "emu_step_concurrent: %s\n", ex.what());
                rc[(size_t)rk] = std::string(ex.what()).find("cannot create") != std::string::npos ? 77 : 1;
            }
        });
    for (auto& t : drivers) t.join();
    for (int v : rc) if (v) return v;
    for (auto& r : e.ranks) if (r->status) return 2;      // a halo wait gave up
    return 0;
}

// PD_BODY_KERNEL experiment: bodyVertStart[nBodies] = first ORIGINAL vertex id of every body (single rank only)
int emu_enable_body_kernel(void* h, int nV, int nT, const float* X, const uint32_t* Tet, const float* mu, int nBodies, const int* bodyVertStart)
{
    Emu& e = *static_cast<Emu*>(h);
    if (e.world != 1) return 1;
    try {
        build_body_batch(nV, nT, X, Tet, mu, std::vector<int>(bodyVertStart, bodyVertStart + nBodies), e.ranks[0]->L.vertNewOfOld.data(), e.bb);
        e.bodyMode = true;
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "emu_enable_body_kernel: %s\n", ex.what());
        return 2;
    }
    return 0;
}

void emu_destroy(void* h) { delete static_cast<Emu*>(h); }

int emu_step(void* h, float dt, float gravity, float rho, float muN, float muT, int iters, int nSteps)
{
    Emu& e = *static_cast<Emu*>(h);
    try {
        if (!e.ready) {     // Engine::prepare
            for (auto& r : e.ranks) matrix_diag_host(r->L, r->md);
            e.dt2Prepared = dt * dt;
            e.ready = true;
        }
        for (int s = 0; s < nSteps && e.bodyMode; ++s) {
            Rank& r = *e.ranks[0];
            const float dtInv = 1.0f / dt, wdbc = 1e6f * (dtInv * dtInv);
            const size_t smem = body_smem_bytes(e.bb.nVmax, e.bb.nTmax);
            auto run = [&](auto kernel) {
                pd_emu::launch((unsigned)e.bb.bodies.size(), 512u, smem, kernel, (const BodyDesc*)e.bb.bodies.data(), (const uint32_t*)e.bb.verts.data(),
                               (const uint8_t*)e.bb.rec.data(), (const uint32_t*)e.bb.incPtr.data(), (const uint16_t*)e.bb.inc.data(), (const float*)e.bb.md.data(),
                               e.bb.nVmax, e.bb.nTmax, r.X.data(), r.V.data(), r.XT.data(), (const float*)r.mass.data(), (const float*)r.dbc.data(),
                               (const float4*)r.dbcx.data(), dt, e.dt2Prepared, gravity, iters, rho, wdbc, e.dfb, muT, muN, (unsigned long long*)nullptr);
            };
            if (e.rotMode == 1) run(k_body_step<1>); else run(k_body_step<0>);
        }
        for (int s = 0; s < nSteps && !e.bodyMode; ++s) {
            if (e.rotMode == 1) step_once<1, false>(e, dt, gravity, rho, muN, muT, iters);
            else step_once<0, true>(e, dt, gravity, rho, muN, muT, iters);
        }
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "emu_step: %s\n", ex.what());
        return std::string(ex.what()).find("cannot create") != std::string::npos ? 77 : 1;      // 77: the host cannot run the emulation (skip)
    }
    return 0;
}

static void gather3(const Emu& e, const std::vector<float4> Rank::*field, float* out)
{
    if (!out) return;
    for (const auto& r : e.ranks)
        for (int v = 0; v < r->nOwn; ++v) {
            const size_t o = r->L.vertOrder[(size_t)v];
            const float4 a = ((*r).*field)[(size_t)v];
            out[3 * o] = a.x; out[3 * o + 1] = a.y; out[3 * o + 2] = a.z;
        }
}
void emu_get(void* h, float* X, float* V, float* XT)
{
    const Emu& e = *static_cast<Emu*>(h);
    gather3(e, &Rank::X, X); gather3(e, &Rank::V, V); gather3(e, &Rank::XT, XT);
}
void emu_set(void* h, const float* X, const float* V, const float* XT)
{
    Emu& e = *static_cast<Emu*>(h);
    for (auto& r : e.ranks)
        for (int v = 0; v < r->nV; ++v) {
            const size_t o = r->L.vertOrder[(size_t)v];
            if (X) r->X[(size_t)v] = make_float4(X[3 * o], X[3 * o + 1], X[3 * o + 2], 0.f);
            if (V) r->V[(size_t)v] = make_float4(V[3 * o], V[3 * o + 1], V[3 * o + 2], 0.f);
            if (XT) r->XT[(size_t)v] = make_float4(XT[3 * o], XT[3 * o + 1], XT[3 * o + 2], 0.f);
        }
}
// mouse drag, single rank only (like the engine); more == NULL clears
int emu_set_drag(void* h, const float* more, const float* off, const float* target)
{
    Emu& e = *static_cast<Emu*>(h);
    if (e.world != 1) return 1;
    Rank& r = *e.ranks[0];
    e.drag = false;
    for (int v = 0; v < r.nV; ++v) {
        const size_t o = r.L.vertOrder[(size_t)v];
        r.more[(size_t)v] = more ? more[o] : 0.f;
        if (more && off) r.offX[(size_t)v] = make_float4(off[3 * o], off[3 * o + 1], off[3 * o + 2], 0.f);
        e.drag = e.drag || r.more[(size_t)v] > 0.f;
    }
    if (target) for (int k = 0; k < 3; ++k) e.target[k] = target[k];
    return 0;
}
// Control_Kernel (k_drag_select) on the current X; selectV = original vertex id or -1; single rank only
int emu_drag_select(void* h, int selectV, float controlMag, const float* target)
{
    Emu& e = *static_cast<Emu*>(h);
    if (e.world != 1) return 1;
    Rank& r = *e.ranks[0];
    int flag = 0;
    const int sel = selectV >= 0 ? (int)r.L.vertNewOfOld[(size_t)selectV] : -1;
    pd_emu::launch_flat((unsigned)((r.nV + 255) / 256), 256u, k_drag_select, r.nV, (const float4*)r.X.data(), (const float*)r.dbc.data(), sel, controlMag,
                        r.more.data(), r.offX.data(), &flag);
    e.drag = flag != 0;
    if (target) for (int k = 0; k < 3; ++k) e.target[k] = target[k];
    return 0;
}
void emu_get_drag(void* h, float* more, float* off)
{
    const Emu& e = *static_cast<Emu*>(h);
    const Rank& r = *e.ranks[0];
    for (int v = 0; v < r.nV; ++v) {
        const size_t o = r.L.vertOrder[(size_t)v];
        if (more) more[o] = r.more[(size_t)v];
        if (off) { off[3 * o] = r.offX[(size_t)v].x; off[3 * o + 1] = r.offX[(size_t)v].y; off[3 * o + 2] = r.offX[(size_t)v].z; }
    }
}
void emu_info(void* h, long long* tetsEvaluated, long long* tiles, long long* ghosts)
{
    const Emu& e = *static_cast<Emu*>(h);
    long long t = 0, ti = 0, g = 0;
    for (const auto& r : e.ranks) { t += r->L.nT; ti += r->L.nTiles; g += r->nV - r->nOwn; }
    if (tetsEvaluated) *tetsEvaluated = t;
    if (tiles) *tiles = ti;
    if (ghosts) *ghosts = g;
}

}  // extern "C"
