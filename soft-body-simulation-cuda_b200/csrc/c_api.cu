// extern "C" boundary (include/pd_b200.h).  No exceptions cross it.
#include "../../include/pd_b200.h"

#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include <cuda_runtime.h>

#include "layout.hpp"
#include "pd_engine.hpp"
#include "collision.hpp"
#include "pd_linear.hpp"
#include "scene.hpp"

using namespace pdb200;

struct pd_scene { Scene s; };
struct pd_layout { Layout L; };
struct pd_engine { Engine* e; };
struct pd_rank_plan { RankPlan P; };

static thread_local std::string g_err;
static int fail(int code, const std::string& m) { g_err = m; return code; }

#define PD_TRY try {
#define PD_CATCH_INT                                                                                   \
    } catch (const std::exception& ex) {                                                              \
        const std::string m = ex.what();                                                               \
        int code = PD_ERR_INVALID;                                                                     \
        if (m.find("CUDA") != std::string::npos) code = PD_ERR_CUDA;                                   \
        else if (m.find("open") != std::string::npos || m.find("json") != std::string::npos) code = PD_ERR_IO; \
        else if (m.find("outside the PD hot path") != std::string::npos || m.find("out of scope") != std::string::npos) code = PD_ERR_UNSUPPORTED; \
        return fail(code, m);                                                                          \
    } catch (...) { return fail(PD_ERR_INVALID, "unknown exception"); }
#define PD_CATCH_PTR                                                                                   \
    } catch (const std::exception& ex) { g_err = ex.what(); return nullptr;                           \
    } catch (...) { g_err = "unknown exception"; return nullptr; }

static void to_c(const SolverParams& p, pd_params* o)
{
    o->dt = p.dt; o->gravity = p.gravity; o->muN = p.muN; o->muT = p.muT; o->rho = p.rho; o->tol = p.tol; o->damp = p.damp;
    o->num_iterations = p.numIterations; o->global_solver = p.globalSolver; o->pcg_max_iter = p.pcgMaxIter;
    o->pcg_tol = p.pcgTol; o->handle_collision = p.handleCollision; o->threads_per_block = p.threadsPerBlock;
}
static void from_c(const pd_params* i, SolverParams& p)
{
    p.dt = i->dt; p.gravity = i->gravity; p.muN = i->muN; p.muT = i->muT; p.rho = i->rho; p.damp = i->damp;
    p.tol = i->tol < 1e-6f ? 1e-6f : i->tol;     // CopyUIToParams clamp, simulationContext.cpp:28-31
    p.numIterations = i->num_iterations; p.globalSolver = i->global_solver; p.pcgMaxIter = i->pcg_max_iter;
    p.pcgTol = i->pcg_tol; p.handleCollision = i->handle_collision; p.threadsPerBlock = i->threads_per_block;
}

extern "C" {

const char* pd_last_error(void) { return g_err.c_str(); }
const char* pd_version(void) { return "pd_b200 0.1 (sm_100a)"; }

void pd_default_params(pd_params* p) { SolverParams d; to_c(d, p); }
void pd_default_options(pd_engine_options* o) { o->device = 0; o->rot_mode = PD_ROT_AUTO; o->reorder = 1; o->use_graph = 1; o->ctas_per_sm = 0; o->rank = 0; o->world = 1; o->body_kernel = -1; }

// ------------------------------------------------------------------ scene
pd_scene* pd_scene_load_json(const char* json_path, const char* context_name, const char* asset_root)
{
    PD_TRY
    if (!json_path) { g_err = "json_path is NULL"; return nullptr; }
    pd_scene* s = new pd_scene;
    try { s->s = load_context_json(json_path, context_name ? context_name : "", asset_root ? asset_root : ""); }
    catch (...) { delete s; throw; }
    return s;
    PD_CATCH_PTR
}

pd_scene* pd_scene_from_desc(const pd_scene_desc* d, const pd_params* params)
{
    PD_TRY
    if (!d || d->num_verts <= 0 || d->num_tets <= 0 || !d->X || !d->Tet || !d->mass || !d->mu) { g_err = "invalid scene description"; return nullptr; }
    for (size_t i = 0; i < 4 * (size_t)d->num_tets; ++i)
        if (d->Tet[i] >= (uint32_t)d->num_verts) { g_err = "tet references vertex out of range"; return nullptr; }
    pd_scene* s = new pd_scene;
    Scene& sc = s->s;
    sc.name = "desc";
    sc.numVerts = d->num_verts; sc.numTets = d->num_tets;
    sc.X.assign(d->X, d->X + 3 * (size_t)d->num_verts);
    sc.Tet.assign(d->Tet, d->Tet + 4 * (size_t)d->num_tets);
    sc.mass.assign(d->mass, d->mass + d->num_verts);
    sc.mu.assign(d->mu, d->mu + d->num_tets);
    sc.lambda.assign((size_t)d->num_tets, 0.f);
    if (d->DBC) sc.DBC.assign(d->DBC, d->DBC + d->num_verts); else sc.DBC.assign((size_t)d->num_verts, 0.f);
    sc.bodyVertStart = {0}; sc.bodyTetStart = {0}; sc.bodyNames = {"desc"}; sc.bodyHasTri = {0};
    if (d->num_tris > 0 && d->Tri) {      // SolverData::Tri / dev_TriFathers as the caller holds them (dataLoader.cu:343-369)
        sc.Tri.assign(d->Tri, d->Tri + 3 * (size_t)d->num_tris);
        if (d->TriFathers) sc.triFather.assign(d->TriFathers, d->TriFathers + d->num_tris); else sc.triFather.assign((size_t)d->num_tris, 0u);
        sc.triWholeScene = true;
    }
    for (int i = 0; i < d->num_fixed; ++i) {
        FixedBody f; f.type = d->fixed[i].type; std::memcpy(f.model, d->fixed[i].model, 64); f.radius = d->fixed[i].radius;
        sc.fixed.push_back(f);
    }
    if (params) from_c(params, sc.params);
    return s;
    PD_CATCH_PTR
}

pd_scene* pd_scene_kuhn_grid(int nx, int ny, int nz, float h, float jitter, uint32_t seed, const float origin[3], float mass, float mu)
{
    PD_TRY
    if (nx <= 0 || ny <= 0 || nz <= 0) { g_err = "grid dimensions must be positive"; return nullptr; }
    pd_scene* s = new pd_scene;
    const float o0[3] = {0, 0, 0};
    s->s = make_kuhn_grid(nx, ny, nz, h, jitter, seed, origin ? origin : o0, mass, mu);
    return s;
    PD_CATCH_PTR
}

// Batch of independent contexts -> one scene: bodies of different contexts share no tets, so the system matrix is
// block diagonal and one step of the merged scene is one step of every context (SURVEY.md 8e "free decoupling").
pd_scene* pd_scene_merge(const pd_scene* const* scenes, int n)
{
    PD_TRY
    if (!scenes || n <= 0 || !scenes[0]) { g_err = "no scenes to merge"; return nullptr; }
    pd_scene* out = new pd_scene;
    Scene& m = out->s;
    m.name = "batch"; m.params = scenes[0]->s.params; m.fixed = scenes[0]->s.fixed; m.precision = scenes[0]->s.precision;
    for (int i = 0; i < n; ++i) {
        if (!scenes[i]) { delete out; g_err = "scene is NULL"; return nullptr; }
        const Scene& a = scenes[i]->s;
        const SolverParams &p = a.params, &q = m.params;
        if (p.dt != q.dt || p.gravity != q.gravity || p.muN != q.muN || p.muT != q.muT || p.rho != q.rho || p.numIterations != q.numIterations ||
            a.fixed.size() != m.fixed.size()) { delete out; g_err = "contexts of a batch must share solver parameters and fixed bodies"; return nullptr; }
        const uint32_t vOff = (uint32_t)m.numVerts;
        const uint32_t bOff = (uint32_t)m.bodyVertStart.size();
        if (a.triWholeScene) {       // a caller-given surface: keep it, under the merged numbering, as explicit triangles of its bodies
            if (a.bodyVertStart.size() != 1) { delete out; g_err = "cannot merge a multi-body scene that carries caller-given surface triangles"; return nullptr; }
            for (uint32_t v : a.Tri) m.Tri.push_back(v + vOff);
            m.triFather.insert(m.triFather.end(), a.triFather.size(), bOff);
        } else {
            for (uint32_t v : a.Tri) m.Tri.push_back(v + vOff);
            for (uint32_t f : a.triFather) m.triFather.push_back(f + bOff);
        }
        for (size_t b = 0; b < a.bodyVertStart.size(); ++b) {
            m.bodyVertStart.push_back(a.bodyVertStart[b] + m.numVerts); m.bodyTetStart.push_back(a.bodyTetStart[b] + m.numTets);
            m.bodyNames.push_back(a.bodyNames[b] + "#" + std::to_string(i));
            m.bodyHasTri.push_back(a.triWholeScene ? 1 : (b < a.bodyHasTri.size() ? a.bodyHasTri[b] : 0));
        }
        m.X.insert(m.X.end(), a.X.begin(), a.X.end());
        for (uint32_t v : a.Tet) m.Tet.push_back(v + vOff);
        m.mass.insert(m.mass.end(), a.mass.begin(), a.mass.end()); m.DBC.insert(m.DBC.end(), a.DBC.begin(), a.DBC.end());
        m.mu.insert(m.mu.end(), a.mu.begin(), a.mu.end()); m.lambda.insert(m.lambda.end(), a.lambda.begin(), a.lambda.end());
        m.numVerts += a.numVerts; m.numTets += a.numTets;
    }
    return out;
    PD_CATCH_PTR
}

void pd_scene_free(pd_scene* s) { delete s; }

int pd_scene_counts(const pd_scene* s, int* nv, int* nt, int* nf, int* nb)
{
    if (!s) return fail(PD_ERR_INVALID, "scene is NULL");
    if (nv) *nv = s->s.numVerts;
    if (nt) *nt = s->s.numTets;
    if (nf) *nf = (int)s->s.fixed.size();
    if (nb) *nb = (int)s->s.bodyVertStart.size();
    return PD_OK;
}

int pd_scene_get(const pd_scene* s, float* X, uint32_t* Tet, float* mass, float* mu, float* DBC, pd_fixed_body* fixed, int* bvs)
{
    if (!s) return fail(PD_ERR_INVALID, "scene is NULL");
    const Scene& sc = s->s;
    if (X) std::memcpy(X, sc.X.data(), sc.X.size() * 4);
    if (Tet) std::memcpy(Tet, sc.Tet.data(), sc.Tet.size() * 4);
    if (mass) std::memcpy(mass, sc.mass.data(), sc.mass.size() * 4);
    if (mu) std::memcpy(mu, sc.mu.data(), sc.mu.size() * 4);
    if (DBC) std::memcpy(DBC, sc.DBC.data(), sc.DBC.size() * 4);
    if (fixed)
        for (size_t i = 0; i < sc.fixed.size(); ++i) {
            fixed[i].type = sc.fixed[i].type; std::memcpy(fixed[i].model, sc.fixed[i].model, 64); fixed[i].radius = sc.fixed[i].radius;
        }
    if (bvs) for (size_t i = 0; i < sc.bodyVertStart.size(); ++i) bvs[i] = sc.bodyVertStart[i];
    return PD_OK;
}

int pd_scene_get_surface(const pd_scene* s, int* num_tris, uint32_t* tri, uint32_t* father)
{
    PD_TRY
    if (!s || !num_tris) return fail(PD_ERR_INVALID, "NULL argument");
    std::vector<uint32_t> t, f;
    scene_surface(s->s, t, f);
    *num_tris = (int)f.size();
    if (tri) std::memcpy(tri, t.data(), t.size() * 4);
    if (father) std::memcpy(father, f.data(), f.size() * 4);
    return PD_OK;
    PD_CATCH_INT
}

int pd_scene_get_params(const pd_scene* s, pd_params* out)
{
    if (!s || !out) return fail(PD_ERR_INVALID, "NULL argument");
    to_c(s->s.params, out);
    return PD_OK;
}
int pd_scene_set_params(pd_scene* s, const pd_params* in)
{
    if (!s || !in) return fail(PD_ERR_INVALID, "NULL argument");
    from_c(in, s->s.params);
    return PD_OK;
}
int pd_scene_add_fixed(pd_scene* s, const pd_fixed_body* fb)
{
    if (!s || !fb) return fail(PD_ERR_INVALID, "NULL argument");
    if (fb->type < 0 || fb->type > 2) return fail(PD_ERR_INVALID, "unknown fixed body type");
    FixedBody f; f.type = fb->type; std::memcpy(f.model, fb->model, 64); f.radius = fb->radius;
    s->s.fixed.push_back(f);
    return PD_OK;
}
int pd_scene_write_tetgen(const pd_scene* s, const char* np, const char* ep)
{
    PD_TRY
    if (!s || !np || !ep) return fail(PD_ERR_INVALID, "NULL argument");
    write_tetgen(s->s, np, ep);
    return PD_OK;
    PD_CATCH_INT
}

int pd_load_node(const char* path, int centralize, float** X, int* nv)
{
    PD_TRY
    if (!path || !X || !nv) return fail(PD_ERR_INVALID, "NULL argument");
    std::vector<float> v = load_node_file(path, centralize != 0);
    *X = (float*)std::malloc(v.size() * 4);
    std::memcpy(*X, v.data(), v.size() * 4);
    *nv = (int)(v.size() / 3);
    return PD_OK;
    PD_CATCH_INT
}
int pd_load_ele(const char* path, int start_index, uint32_t** T, int* nt)
{
    PD_TRY
    if (!path || !T || !nt) return fail(PD_ERR_INVALID, "NULL argument");
    std::vector<uint32_t> v = load_ele_file(path, start_index);
    *T = (uint32_t*)std::malloc(v.size() * 4);
    std::memcpy(*T, v.data(), v.size() * 4);
    *nt = (int)(v.size() / 4);
    return PD_OK;
    PD_CATCH_INT
}
void pd_free(void* p) { std::free(p); }

void pd_model_matrix(const float pos[3], const float rot[3], const float scale[3], int sbo, float M[16]) { model_matrix(pos, rot, scale, sbo != 0, M); }
void pd_transform_vertices(float* X, int nv, const float M[16]) { transform_vertices(X, nv, M); }
void pd_plane_up(const float M[16], float up[3]) { plane_up(M, up); }

// ------------------------------------------------------------------ layout
pd_layout* pd_layout_build(const pd_scene* s, int reorder)
{
    PD_TRY
    if (!s) { g_err = "scene is NULL"; return nullptr; }
    pd_layout* l = new pd_layout;
    try { build_layout(s->s.numVerts, s->s.numTets, s->s.X.data(), s->s.Tet.data(), s->s.mu.data(), reorder != 0, l->L); }
    catch (...) { delete l; throw; }
    return l;
    PD_CATCH_PTR
}
void pd_layout_free(pd_layout* l) { delete l; }
int pd_layout_counts(const pd_layout* l, int* nTiles, uint32_t* nSlots, size_t* recBytes, int* maxLocal)
{
    if (!l) return fail(PD_ERR_INVALID, "layout is NULL");
    if (nTiles) *nTiles = l->L.nTiles;
    if (nSlots) *nSlots = l->L.nSlots;
    if (recBytes) *recBytes = l->L.records.size();
    if (maxLocal) *maxLocal = l->L.maxLocal;
    return PD_OK;
}
int pd_layout_get(const pd_layout* l, uint32_t* tetOrder, uint32_t* vertOrder, uint32_t* tetNew, uint32_t* tileTetStart,
                  uint64_t* tileRecOff, uint8_t* records, uint32_t* vslotPtr, uint32_t* vslot, uint32_t* vlist)
{
    if (!l) return fail(PD_ERR_INVALID, "layout is NULL");
    const Layout& L = l->L;
    if (tetOrder) std::memcpy(tetOrder, L.tetOrder.data(), L.tetOrder.size() * 4);
    if (vertOrder) std::memcpy(vertOrder, L.vertOrder.data(), L.vertOrder.size() * 4);
    if (tetNew) std::memcpy(tetNew, L.tetNew.data(), L.tetNew.size() * 4);
    if (tileTetStart) std::memcpy(tileTetStart, L.tileTetStart.data(), L.tileTetStart.size() * 4);
    if (tileRecOff) std::memcpy(tileRecOff, L.tileRecOff.data(), L.tileRecOff.size() * 8);
    if (records) std::memcpy(records, L.records.data(), L.records.size());
    if (vslotPtr) std::memcpy(vslotPtr, L.vslotPtr.data(), L.vslotPtr.size() * 4);
    if (vslot) std::memcpy(vslot, L.vslot.data(), L.vslot.size() * 4);
    if (vlist) std::memcpy(vlist, L.vlist.data(), L.vlist.size() * 4);
    return PD_OK;
}
int pd_layout_tile_table(const pd_layout* l, uint32_t* table)
{
    PD_TRY
    if (!l || !table) return fail(PD_ERR_INVALID, "bad argument");
    std::vector<uint32_t> m;
    build_tile_table(l->L, m);
    std::memcpy(table, m.data(), m.size() * 4);
    return PD_OK;
    PD_CATCH_INT
}
int pd_layout_get_vstage(const pd_layout* l, uint32_t* vstage)
{
    if (!l || !vstage) return fail(PD_ERR_INVALID, "bad argument");
    std::memcpy(vstage, l->L.vstage.data(), l->L.vstage.size() * 4);
    return PD_OK;
}
int pd_layout_matrix_diag(const pd_layout* l, float* md)
{
    PD_TRY
    if (!l || !md) return fail(PD_ERR_INVALID, "NULL argument");
    std::vector<float> v;
    matrix_diag_host(l->L, v);
    std::memcpy(md, v.data(), v.size() * sizeof(float));
    return PD_OK;
    PD_CATCH_INT
}
int pd_morton_keys(const float* X, const uint32_t* Tet, int nT, uint32_t* keys)
{
    PD_TRY
    if (!X || !Tet || !keys || nT < 0) return fail(PD_ERR_INVALID, "bad argument");
    std::vector<uint32_t> k;
    morton_keys(X, Tet, nT, k);
    std::memcpy(keys, k.data(), k.size() * 4);
    return PD_OK;
    PD_CATCH_INT
}
int pd_cholesky_factor(int n, const int* rowptr, const int* col, const float* val, int* nnz_l, int** lptr, int** lcol, float** lval)
{
    PD_TRY
    if (n <= 0 || !rowptr || !col || !val || !nnz_l || !lptr || !lcol || !lval) return fail(PD_ERR_INVALID, "bad argument");
    CsrMatrix A; A.n = n;
    A.rowPtr.assign(rowptr, rowptr + n + 1); A.col.assign(col, col + rowptr[n]); A.val.assign(val, val + rowptr[n]);
    CholFactor F;
    cholesky_factor(A, F);
    *nnz_l = (int)F.lCol.size();
    *lptr = (int*)std::malloc(F.lPtr.size() * sizeof(int)); *lcol = (int*)std::malloc(F.lCol.size() * sizeof(int)); *lval = (float*)std::malloc(F.lVal.size() * sizeof(float));
    std::memcpy(*lptr, F.lPtr.data(), F.lPtr.size() * sizeof(int)); std::memcpy(*lcol, F.lCol.data(), F.lCol.size() * sizeof(int));
    std::memcpy(*lval, F.lVal.data(), F.lVal.size() * sizeof(float));
    return PD_OK;
    PD_CATCH_INT
}
int pd_nested_dissection(int n, const int* rowptr, const int* col, const float* xyz, int* perm, int* nnz_l_natural, int* nnz_l_ordered)
{
    PD_TRY
    if (n <= 0 || !rowptr || !col || !xyz || !perm) return fail(PD_ERR_INVALID, "bad argument");
    CsrMatrix A; A.n = n;
    A.rowPtr.assign(rowptr, rowptr + n + 1); A.col.assign(col, col + rowptr[n]); A.val.assign((size_t)rowptr[n], 0.f);
    std::vector<int> p;
    nested_dissection_order(A, xyz, p);
    std::memcpy(perm, p.data(), (size_t)n * sizeof(int));
    if (nnz_l_natural || nnz_l_ordered) {
        // symbolic only: a diagonally dominant stand-in with the same pattern
        for (int i = 0; i < n; ++i)
            for (int e = A.rowPtr[i]; e < A.rowPtr[i + 1]; ++e) A.val[(size_t)e] = (A.col[(size_t)e] == i) ? (float)(A.rowPtr[i + 1] - A.rowPtr[i] + 1) : -1.f;
        CholFactor F;
        if (nnz_l_natural) { cholesky_factor(A, F); *nnz_l_natural = (int)F.lCol.size(); }
        if (nnz_l_ordered) { CsrMatrix B; permute_symmetric(A, p, B); cholesky_factor(B, F); *nnz_l_ordered = (int)F.lCol.size(); }
    }
    return PD_OK;
    PD_CATCH_INT
}
int pd_partition_vertices(int nV, int world, int* vbeg)
{
    if (nV < 0 || world <= 0 || !vbeg) return fail(PD_ERR_INVALID, "bad argument");
    std::vector<int> v;
    partition_vertices(nV, world, v);
    std::memcpy(vbeg, v.data(), v.size() * sizeof(int));
    return PD_OK;
}

// ------------------------------------------------------------------ multi-GPU plan (host only)
pd_rank_plan* pd_rank_plan_build(const pd_layout* g, int world, int rank)
{
    PD_TRY
    if (!g) { g_err = "layout is NULL"; return nullptr; }
    pd_rank_plan* p = new pd_rank_plan;
    try { build_rank_plan(g->L, world, rank, p->P, dist_trim_from_env()); } catch (...) { delete p; throw; }
    return p;
    PD_CATCH_PTR
}
void pd_rank_plan_free(pd_rank_plan* p) { delete p; }
int pd_rank_plan_counts(const pd_rank_plan* p, int c[7])
{
    if (!p || !c) return fail(PD_ERR_INVALID, "NULL argument");
    c[0] = p->P.nOwn; c[1] = p->P.nGhost; c[2] = (int)p->P.tiles.size(); c[3] = (int)p->P.neighbours.size();
    c[4] = (int)p->P.pushSrc.size(); c[5] = p->P.vbeg[(size_t)p->P.rank]; c[6] = p->P.nInteriorTiles;
    return PD_OK;
}
int pd_rank_plan_get(const pd_rank_plan* p, uint32_t* tiles, uint32_t* ghosts, int* nbr, int* nLocOf, uint32_t* ps, uint32_t* pdst, int* pr)
{
    if (!p) return fail(PD_ERR_INVALID, "plan is NULL");
    const RankPlan& P = p->P;
    if (tiles) std::memcpy(tiles, P.tiles.data(), P.tiles.size() * 4);
    if (ghosts) std::memcpy(ghosts, P.ghosts.data(), P.ghosts.size() * 4);
    if (nbr) std::memcpy(nbr, P.neighbours.data(), P.neighbours.size() * sizeof(int));
    if (nLocOf) std::memcpy(nLocOf, P.nLocOf.data(), P.nLocOf.size() * sizeof(int));
    if (ps) std::memcpy(ps, P.pushSrc.data(), P.pushSrc.size() * 4);
    if (pdst) std::memcpy(pdst, P.pushDst.data(), P.pushDst.size() * 4);
    if (pr) std::memcpy(pr, P.pushRank.data(), P.pushRank.size() * sizeof(int));
    return PD_OK;
}
pd_layout* pd_rank_layout(const pd_layout* g, const pd_rank_plan* p)
{
    PD_TRY
    if (!g || !p) { g_err = "NULL argument"; return nullptr; }
    pd_layout* l = new pd_layout;
    try { extract_rank_layout(g->L, p->P, l->L); } catch (...) { delete l; throw; }
    return l;
    PD_CATCH_PTR
}

// ------------------------------------------------------------------ engine
pd_engine* pd_create(const pd_scene* s, const pd_engine_options* o)
{
    PD_TRY
    if (!s) { g_err = "scene is NULL"; return nullptr; }
    EngineOptions eo;
    if (o) { eo.device = o->device; eo.rotMode = o->rot_mode; eo.reorder = o->reorder; eo.useGraph = o->use_graph; eo.ctasPerSm = o->ctas_per_sm;
             eo.rank = o->rank; eo.world = o->world < 1 ? 1 : o->world; eo.bodyKernel = o->body_kernel; }
    pd_engine* e = new pd_engine{nullptr};
    try { e->e = new Engine(s->s, eo); } catch (...) { delete e; throw; }
    return e;
    PD_CATCH_PTR
}

pd_engine* pd_create_from_json(const char* json_path, const char* context_name, const char* asset_root, const pd_engine_options* o)
{
    pd_scene* s = pd_scene_load_json(json_path, context_name, asset_root);
    if (!s) return nullptr;
    pd_engine* e = nullptr;
    if (s->s.precision != "float" && s->s.precision != "fp32") g_err = "context '" + s->s.name + "' is not a float (PD) context";
    else e = pd_create(s, o);
    pd_scene_free(s);
    return e;
}

void pd_destroy(pd_engine* e) { if (e) { delete e->e; delete e; } }

#define ENGINE_CALL(body)                                   \
    PD_TRY                                                  \
    if (!e || !e->e) return fail(PD_ERR_INVALID, "engine is NULL"); \
    body;                                                   \
    return PD_OK;                                           \
    PD_CATCH_INT

int pd_step(pd_engine* e, int n) { ENGINE_CALL(if (n < 0) return fail(PD_ERR_INVALID, "n_steps < 0"); e->e->step(n)) }
int pd_synchronize(pd_engine* e) { ENGINE_CALL(e->e->synchronize()) }
int pd_step_timed(pd_engine* e, int n, float* ms)
{
    ENGINE_CALL(if (n < 0 || !ms) return fail(PD_ERR_INVALID, "bad argument"); *ms = e->e->stepTimed(n))
}
int pd_set_params(pd_engine* e, const pd_params* p)
{
    ENGINE_CALL(if (!p) return fail(PD_ERR_INVALID, "params is NULL"); SolverParams sp = e->e->params(); from_c(p, sp); e->e->setParams(sp))
}
int pd_get_params(const pd_engine* e, pd_params* p)
{
    if (!e || !e->e || !p) return fail(PD_ERR_INVALID, "NULL argument");
    to_c(e->e->params(), p);
    return PD_OK;
}
int pd_set_global_solver(pd_engine* e, int solver)
{
    ENGINE_CALL(if (solver < 0 || solver > 2) return fail(PD_ERR_INVALID, "unknown solver"); SolverParams sp = e->e->params(); sp.globalSolver = solver; e->e->setParams(sp))
}
int pd_reset(pd_engine* e) { ENGINE_CALL(e->e->reset()) }
int pd_set_perf(pd_engine* e, int on) { ENGINE_CALL(e->e->setPerf(on != 0)) }
int pd_get_perf(const pd_engine* e, pd_perf* o)
{
    if (!e || !e->e || !o) return fail(PD_ERR_INVALID, "NULL argument");
    try { e->e->syncSolveStats(); } catch (const std::exception& ex) { return fail(PD_ERR_CUDA, ex.what()); }
    const PerfCounters& c = e->e->perf();
    o->local_step_ms = c.localStep; o->global_step_ms = c.globalStep; o->collision_fixed_ms = c.collisionFixed; o->collision_mesh_ms = c.collisionMesh;
    o->step_ms_total = c.stepMsTotal; o->steps = c.steps; o->pd_iterations = c.pdIterations; o->inner_iterations = c.innerIterations;
    o->kernel_launches = c.kernelLaunches;
    return PD_OK;
}
int pd_download(pd_engine* e, float* X, float* V, float* XT) { ENGINE_CALL(e->e->download(X, V, XT)) }
int pd_upload_state(pd_engine* e, const float* X, const float* V, const float* XT) { ENGINE_CALL(e->e->upload(X, V, XT); e->e->synchronize()) }
int pd_step_host(pd_engine* e, int n, const float* Xi, const float* Vi, const float* XTi, float* Xo, float* Vo, float* XTo)
{
    ENGINE_CALL(if (n < 0) return fail(PD_ERR_INVALID, "n_steps < 0"); e->e->stepHost(n, Xi, Vi, XTi, Xo, Vo, XTo))
}
int pd_step_host_owned(pd_engine* e, int n, const float* Xi, const float* Vi, const float* XTi, float* Xo, float* Vo, float* XTo)
{
    ENGINE_CALL(if (n < 0) return fail(PD_ERR_INVALID, "n_steps < 0"); e->e->stepHostOwned(n, Xi, Vi, XTi, Xo, Vo, XTo))
}
int pd_dist_owned_ids(const pd_engine* e, uint32_t* out)
{
    if (!e || !e->e || !out) return fail(PD_ERR_INVALID, "NULL argument");
    e->e->ownedIds(out);
    return PD_OK;
}
int pd_update_device(pd_engine* e, int n, float* dX, float* dV, float* dXT)
{
    ENGINE_CALL(if (n < 0) return fail(PD_ERR_INVALID, "n_steps < 0");
                e->e->importDevice(dX, dV, dXT); e->e->step(n); e->e->exportDevice(dX, dV, dXT); e->e->synchronize())
}
int pd_set_drag(pd_engine* e, const float* more, const float* off, const float* target)
{
    ENGINE_CALL(if (more && (!off || !target)) return fail(PD_ERR_INVALID, "offset_x and target are required with more_dbc"); e->e->setDrag(more, off, target))
}
int pd_set_drag_device(pd_engine* e, const float* dMore, const float* dOff, const float* target)
{
    ENGINE_CALL(if (dMore && (!dOff || !target)) return fail(PD_ERR_INVALID, "d_offset_x and target are required with d_more_dbc"); e->e->setDragDevice(dMore, dOff, target))
}
int pd_drag_select(pd_engine* e, int selectV, float controlMag, const float* target)
{
    ENGINE_CALL(if (!target) return fail(PD_ERR_INVALID, "target is NULL"); e->e->dragSelect(selectV, controlMag, target))
}
int pd_get_drag(pd_engine* e, float* more, float* off, float* dbcx, int* active)
{
    ENGINE_CALL(e->e->getDrag(more, off, dbcx); if (active) *active = e->e->dragActive() ? 1 : 0)
}
int pd_update_mu(pd_engine* e, const float* mu) { ENGINE_CALL(e->e->updateMu(mu)) }
int pd_update_mu_device(pd_engine* e, const float* dMu) { ENGINE_CALL(e->e->updateMuDevice(dMu)) }
int pd_get_setup(pd_engine* e, float* md, float* mdt2, float* DmInv, float* V0) { ENGINE_CALL(e->e->getSetup(md, mdt2, DmInv, V0)) }
int pd_get_system_matrix(pd_engine* e, int* nnz, int* rowptr, int* col, float* val)
{
    ENGINE_CALL(const CsrMatrix& A = e->e->systemMatrix();
                if (nnz) *nnz = (int)A.col.size();
                if (rowptr) std::memcpy(rowptr, A.rowPtr.data(), A.rowPtr.size() * sizeof(int));
                if (col) std::memcpy(col, A.col.data(), A.col.size() * sizeof(int));
                if (val) std::memcpy(val, A.val.data(), A.val.size() * sizeof(float)))
}
int pd_get_collision(pd_engine* e, float* tI, float* normals, long long* num_pairs) { ENGINE_CALL(e->e->getCollision(tI, normals, num_pairs)) }
int pd_get_solver_sizes(pd_engine* e, long long* nnzA, long long* nnzL)
{
    if (!e || !e->e) return fail(PD_ERR_INVALID, "engine is NULL");
    size_t a = 0, l = 0;
    e->e->solverSizes(a, l);
    if (nnzA) *nnzA = (long long)a;
    if (nnzL) *nnzL = (long long)l;
    return PD_OK;
}
int pd_get_solve_stats(pd_engine* e, float* err, int* pd_iters)
{
    ENGINE_CALL(e->e->syncSolveStats(); if (err) *err = e->e->lastError(); if (pd_iters) *pd_iters = e->e->lastPdIterations())
}
int pd_time_kernels(pd_engine* e, int reps, float* lms, float* vms)
{
    ENGINE_CALL(if (reps <= 0) return fail(PD_ERR_INVALID, "reps <= 0");
                if (lms) *lms = e->e->timeLocalKernelMs(reps); if (vms) *vms = e->e->timeVertexKernelMs(reps))
}
int pd_profile_local(pd_engine* e, unsigned long long* out) { ENGINE_CALL(if (!out) return fail(PD_ERR_INVALID, "out is NULL"); e->e->profileLocal(out)) }
int pd_engine_info(const pd_engine* e, int* nv, int* nt, int* ntiles, uint32_t* nslots, size_t* streamBytes, size_t* devBytes, int* lgrid)
{
    if (!e || !e->e) return fail(PD_ERR_INVALID, "engine is NULL");
    if (nv) *nv = e->e->numVerts();
    if (nt) *nt = e->e->numTets();
    if (ntiles) *ntiles = e->e->layout().nTiles;
    if (nslots) *nslots = e->e->layout().nSlots;
    if (streamBytes) *streamBytes = e->e->tileStreamBytes();
    if (devBytes) *devBytes = e->e->deviceBytes();
    if (lgrid) *lgrid = e->e->localGrid();
    return PD_OK;
}

int pd_engine_rot_mode(const pd_engine* e) { return (e && e->e) ? e->e->rotMode() : PD_ERR_INVALID; }

int pd_dist_window_handle(pd_engine* e, void* out64) { ENGINE_CALL(if (!out64) return fail(PD_ERR_INVALID, "out is NULL"); e->e->windowHandle(out64)) }
int pd_dist_connect(pd_engine* e, const void* handles) { ENGINE_CALL(if (!handles) return fail(PD_ERR_INVALID, "handles is NULL"); e->e->connectIpc(handles)) }
static int collect(pd_engine* const* engines, int n, std::vector<Engine*>& v)
{
    if (!engines || n <= 0) return fail(PD_ERR_INVALID, "no engines");
    for (int i = 0; i < n; ++i) { if (!engines[i] || !engines[i]->e) return fail(PD_ERR_INVALID, "engine is NULL"); v.push_back(engines[i]->e); }
    return PD_OK;
}
int pd_dist_connect_local(pd_engine* const* engines, int n)
{
    PD_TRY
    std::vector<Engine*> v;
    if (int rc = collect(engines, n, v)) return rc;
    Engine::connectLocal(v.data(), n);
    return PD_OK;
    PD_CATCH_INT
}
int pd_dist_step_lockstep(pd_engine* const* engines, int n, int nSteps)
{
    PD_TRY
    std::vector<Engine*> v;
    if (int rc = collect(engines, n, v)) return rc;
    if (nSteps < 0) return fail(PD_ERR_INVALID, "n_steps < 0");
    Engine::stepLockstep(v.data(), n, nSteps);
    return PD_OK;
    PD_CATCH_INT
}
int pd_dist_status(pd_engine* e, unsigned int* out) { ENGINE_CALL(if (!out) return fail(PD_ERR_INVALID, "out is NULL"); *out = e->e->distStatus()) }
int pd_dist_info(const pd_engine* e, int info[6])
{
    if (!e || !e->e || !info) return fail(PD_ERR_INVALID, "NULL argument");
    const RankPlan& P = e->e->plan();
    info[0] = e->e->numOwned(); info[1] = e->e->numVerts() - e->e->numOwned(); info[2] = (int)P.neighbours.size();
    info[3] = (int)P.pushSrc.size(); info[4] = e->e->numTets(); info[5] = e->e->layout().nTiles;
    return PD_OK;
}

// ------------------------------------------------------------------ IPC (double) linear back-ends
struct pd_linsolver { LinearSolver* s; };
pd_linsolver* pd_linsolver_create(int kind, int n, int max_iter, double tolerance, int device)
{
    PD_TRY
    pd_linsolver* h = new pd_linsolver{nullptr};
    try { h->s = new LinearSolver(kind, n, max_iter, tolerance, device); } catch (...) { delete h; throw; }
    return h;
    PD_CATCH_PTR
}
void pd_linsolver_destroy(pd_linsolver* h) { if (h) { delete h->s; delete h; } }
int pd_linsolver_solve_device(pd_linsolver* h, int n, const double* d_b, double* d_x, const double* d_A, int nz, const int* d_row, const int* d_col, const double* d_guess)
{
    PD_TRY
    if (!h || !h->s) return fail(PD_ERR_INVALID, "solver is NULL");
    h->s->solveDevice(n, d_b, d_x, d_A, nz, d_row, d_col, d_guess);
    return PD_OK;
    PD_CATCH_INT
}
int pd_linsolver_solve_host(pd_linsolver* h, int n, const double* b, double* x, const double* A, int nz, const int* row, const int* col, const double* guess)
{
    PD_TRY
    if (!h || !h->s) return fail(PD_ERR_INVALID, "solver is NULL");
    if (!b || !x || !A || !row || !col || nz <= 0) return fail(PD_ERR_INVALID, "NULL argument or empty matrix");
    h->s->solveHost(n, b, x, A, nz, row, col, guess);
    return PD_OK;
    PD_CATCH_INT
}
int pd_linsolver_stats(const pd_linsolver* h, int* iterations, double* residual, int* nnz)
{
    if (!h || !h->s) return fail(PD_ERR_INVALID, "solver is NULL");
    h->s->stats(iterations, residual, nnz);
    return PD_OK;
}

int pd_ccd_batch(int device, int n, const int* type, const uint32_t* verts, int num_verts, const float* X, const float* XTilde, float* toi, float* normals)
{
    PD_TRY
    if (n <= 0 || num_verts <= 0 || !type || !verts || !X || !XTilde || !toi || !normals) return fail(PD_ERR_INVALID, "bad argument");
    for (int i = 0; i < 4 * n; ++i) if (verts[i] >= (uint32_t)num_verts) return fail(PD_ERR_INVALID, "query names a vertex outside the arrays");
    ccd_batch(device, n, type, verts, num_verts, X, XTilde, toi, normals);
    return PD_OK;
    PD_CATCH_INT
}
int pd_rotation_batch(int device, int rot_mode, int n, const float* F, float* R, int* used_fast)
{
    PD_TRY
    if (n <= 0 || !F || !R) return fail(PD_ERR_INVALID, "bad argument");
    rotation_batch(device, rot_mode, n, F, R, used_fast);
    return PD_OK;
    PD_CATCH_INT
}

void* pd_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void pd_free_pinned(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
