// Scene assembly: context.json + TetGen assets -> merged host arrays.
// Follows src/context.cpp:222-387, src/simulation/simulationContext.cu:34-123 and
// src/simulation/dataLoader.cu:38-66,131-173,203-237,291-378 of the reference (behaviour, not code).
#include "scene.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <sstream>
#include <stdexcept>

#include "json_min.hpp"

namespace pdb200 {

// ---------------------------------------------------------------- TetGen readers
std::vector<float> load_node_file(const std::string& path, bool centralize)
{
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Unable to open file: " + path);
    std::string line;
    std::getline(f, line);
    int n = 0;
    { std::istringstream is(line); is >> n; }
    if (n <= 0) throw std::runtime_error("bad .node header: " + path);
    std::vector<float> X((size_t)n * 3, 0.f);
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int i = 0; i < n && std::getline(f, line); ++i) {
        std::istringstream is(line);
        int idx; float x = 0, y = 0, z = 0;          // parsed as float, extra columns ignored
        is >> idx >> x >> y >> z;
        X[3 * i] = x; X[3 * i + 1] = y; X[3 * i + 2] = z;
        cx += x; cy += y; cz += z;
    }
    if (centralize) {                                   // centroid to origin, then swap y <-> z
        const float fn = static_cast<float>(n);
        cx /= fn; cy /= fn; cz /= fn;
        for (int i = 0; i < n; ++i) {
            X[3 * i] -= cx; X[3 * i + 1] -= cy; X[3 * i + 2] -= cz;
            std::swap(X[3 * i + 1], X[3 * i + 2]);
        }
    }
    return X;
}

std::vector<uint32_t> load_ele_file(const std::string& path, int startIndex)
{
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("Unable to open file: " + path);
    std::string line;
    std::getline(f, line);
    int n = 0;
    { std::istringstream is(line); is >> n; }
    if (n <= 0) throw std::runtime_error("bad .ele header: " + path);
    std::vector<uint32_t> T((size_t)n * 4, 0u);
    for (int t = 0; t < n && std::getline(f, line); ++t) {
        std::istringstream is(line);
        int a = 0, b = 0, c = 0, d = 0, e = 0;         // five ints per row; the tet is the last four
        is >> a >> b >> c >> d >> e;
        T[4 * t + 0] = (uint32_t)(b - startIndex);
        T[4 * t + 1] = (uint32_t)(c - startIndex);
        T[4 * t + 2] = (uint32_t)(d - startIndex);
        T[4 * t + 3] = (uint32_t)(e - startIndex);
    }
    return T;
}

std::vector<uint32_t> load_face_file(const std::string& path, int startIndex)
{
    std::vector<uint32_t> F;
    std::ifstream f(path);
    if (!f.is_open()) return F;                        // the reference falls back to the tets' boundary faces (dataLoader.cu:73,92)
    std::string line;
    std::getline(f, line);
    int n = 0;
    { std::istringstream is(line); is >> n; }
    if (n <= 0) return F;
    F.assign((size_t)n * 3, 0u);
    for (int t = 0; t < n && std::getline(f, line); ++t) {
        std::istringstream is(line);
        int a = 0, b = 0, c = 0, d = 0, e = 0;         // index, three vertices, boundary marker
        is >> a >> b >> c >> d >> e;
        F[3 * t + 0] = (uint32_t)(b - startIndex); F[3 * t + 1] = (uint32_t)(c - startIndex); F[3 * t + 2] = (uint32_t)(d - startIndex);
    }
    return F;
}

// ---------------------------------------------------------------- glm-equivalent transforms
namespace {
void m4_identity(float M[16]) { std::memset(M, 0, 64); M[0] = M[5] = M[10] = M[15] = 1.f; }
void m4_translate(float M[16], const float v[3])
{
    for (int r = 0; r < 4; ++r) M[12 + r] = M[r] * v[0] + M[4 + r] * v[1] + M[8 + r] * v[2] + M[12 + r];
}
void m4_scale(float M[16], const float v[3])
{
    for (int r = 0; r < 4; ++r) { M[r] *= v[0]; M[4 + r] *= v[1]; M[8 + r] *= v[2]; }
}
void m4_rotate_axis(float M[16], float angle, int ax)
{   // glm::rotate about a unit coordinate axis (gtc/matrix_transform.inl:52-86)
    float a[3] = {0, 0, 0}; a[ax] = 1.f;
    const float c = std::cos(angle), s = std::sin(angle);
    const float t[3] = {(1.f - c) * a[0], (1.f - c) * a[1], (1.f - c) * a[2]};
    float R[3][3];
    R[0][0] = c + t[0] * a[0];            R[0][1] = 0 + t[0] * a[1] + s * a[2]; R[0][2] = 0 + t[0] * a[2] - s * a[1];
    R[1][0] = 0 + t[1] * a[0] - s * a[2]; R[1][1] = c + t[1] * a[1];            R[1][2] = 0 + t[1] * a[2] + s * a[0];
    R[2][0] = 0 + t[2] * a[0] + s * a[1]; R[2][1] = 0 + t[2] * a[1] - s * a[0]; R[2][2] = c + t[2] * a[2];
    float out[12];
    for (int k = 0; k < 3; ++k)
        for (int r = 0; r < 4; ++r) out[k * 4 + r] = M[r] * R[k][0] + M[4 + r] * R[k][1] + M[8 + r] * R[k][2];
    std::memcpy(M, out, sizeof(out));
}
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
}  // namespace

void model_matrix(const float pos[3], const float rot[3], const float scale[3], bool softBodyOrder, float M[16])
{
    m4_identity(M);
    m4_translate(M, pos);
    if (softBodyOrder) m4_scale(M, scale);
    m4_rotate_axis(M, radians(rot[0]), 0);
    m4_rotate_axis(M, radians(rot[1]), 1);
    m4_rotate_axis(M, radians(rot[2]), 2);
    if (!softBodyOrder) m4_scale(M, scale);
}

void transform_vertices(float* X, int nV, const float M[16])
{   // TransformVertices, utilities.cu:56-65: vec3(M * vec4(x, 1)), pairwise sum like glm's mat4*vec4
    for (int i = 0; i < nV; ++i) {
        const float x = X[3 * i], y = X[3 * i + 1], z = X[3 * i + 2];
        for (int r = 0; r < 3; ++r) {
            const float a = M[r] * x, b = M[4 + r] * y, c = M[8 + r] * z, d = M[12 + r] * 1.f;
            X[3 * i + r] = (a + b) + (c + d);
        }
    }
}

void plane_up(const float M[16], float up[3])
{   // normalize(vec3(transpose(inverse(M)) * (0,1,0,0))): row 1 of the inverse of the linear part
    double a[3][3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) a[r][c] = M[c * 4 + r];
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    float v[3];
    v[0] = (float)(-(a[1][0] * a[2][2] - a[1][2] * a[2][0]) / det);
    v[1] = (float)((a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det);
    v[2] = (float)(-(a[0][0] * a[1][2] - a[0][2] * a[1][0]) / det);
    const float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    up[0] = v[0] / len; up[1] = v[1] / len; up[2] = v[2] / len;
}

void cylinder_axis(const float M[16], float axis[3])
{
    const float len = std::sqrt(M[4] * M[4] + M[5] * M[5] + M[6] * M[6] + M[7] * M[7]);
    axis[0] = M[4] / len; axis[1] = M[5] / len; axis[2] = M[6] / len;
}

// ---------------------------------------------------------------- merge
void scene_add_body(Scene& s, const std::string& name, const std::vector<float>& X, const std::vector<uint32_t>& Tet,
                    float mass, float mu, float lambda, const std::vector<uint32_t>& dbc, const std::vector<uint32_t>* faces)
{
    const int nv = (int)(X.size() / 3), nt = (int)(Tet.size() / 4);
    const uint32_t vOff = (uint32_t)s.numVerts;
    const bool hasTri = faces && !faces->empty();
    s.bodyHasTri.push_back(hasTri ? 1 : 0);
    if (hasTri) {
        for (uint32_t v : *faces) s.Tri.push_back(v + vOff);
        s.triFather.insert(s.triFather.end(), faces->size() / 3, (uint32_t)s.bodyVertStart.size());
    }
    s.bodyVertStart.push_back(s.numVerts);
    s.bodyTetStart.push_back(s.numTets);
    s.bodyNames.push_back(name);
    s.X.insert(s.X.end(), X.begin(), X.end());
    for (uint32_t v : Tet) s.Tet.push_back(v + vOff);
    s.mass.insert(s.mass.end(), nv, mass);
    s.DBC.insert(s.DBC.end(), nv, 0.f);
    for (uint32_t d : dbc) if ((int)d < nv) s.DBC[vOff + d] = 1.f;
    s.mu.insert(s.mu.end(), nt, mu);
    s.lambda.insert(s.lambda.end(), nt, lambda);
    s.numVerts += nv;
    s.numTets += nt;
}

// ---------------------------------------------------------------- context.json
namespace {
std::string read_file(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("Failed to open JSON file: " + path);
    std::ostringstream ss; ss << f.rdbuf();
    return ss.str();
}
bool file_exists(const std::string& p) { std::ifstream f(p); return f.is_open(); }
std::string dir_of(const std::string& p)
{
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}
// The reference opens "../assets/x" from its build directory, one level below context.json.
std::string resolve_asset(const std::string& rel, const std::string& jsonDir, const std::string& assetRoot)
{
    if (rel.empty()) return rel;
    std::vector<std::string> cands;
    std::string stripped = rel;
    while (stripped.rfind("../", 0) == 0) stripped = stripped.substr(3);
    if (!assetRoot.empty()) {
        cands.push_back(assetRoot + "/" + stripped);
        std::string noAssets = stripped.rfind("assets/", 0) == 0 ? stripped.substr(7) : stripped;
        cands.push_back(assetRoot + "/" + noAssets);
    }
    cands.push_back(rel);
    cands.push_back(jsonDir + "/" + rel);
    cands.push_back(jsonDir + "/" + stripped);
    cands.push_back(jsonDir + "/build/" + rel);
    for (auto& c : cands) if (file_exists(c)) return c;
    return cands.front();
}
void vec3_from(const Json& j, float out[3])
{
    out[0] = (float)j[0].number(); out[1] = (float)j[1].number(); out[2] = (float)j[2].number();
}
// getJsonVec3, simulationContext.cu:26-33: override, else definition, else 0
void vec3_pick(const Json& j, const Json& def, const char* key, float out[3])
{
    if (j.contains(key)) vec3_from(j.at(key), out);
    else if (def.contains(key)) vec3_from(def.at(key), out);
    else out[0] = out[1] = out[2] = 0.f;
}
// ReadFixedBodies, context.cpp:222-317: override, else definition, else default
void vec3_pick_default(const Json& j, const Json& def, const char* key, float dflt, float out[3])
{
    if (j.contains(key)) vec3_from(j.at(key), out);
    else if (def.contains(key)) vec3_from(def.at(key), out);
    else out[0] = out[1] = out[2] = dflt;
}
}  // namespace

std::vector<std::string> list_contexts(const std::string& jsonPath)
{
    const std::string text = read_file(jsonPath);
    Json root = JsonParser(text).parse();
    std::vector<std::string> out;
    if (root.contains("contexts"))
        for (auto& c : root.at("contexts").arr) out.push_back(c.at("name").string());
    return out;
}

Scene load_context_json(const std::string& jsonPath, const std::string& contextName, const std::string& assetRoot)
{
    const std::string text = read_file(jsonPath);
    Json root = JsonParser(text).parse();
    const std::string jsonDir = dir_of(jsonPath);

    Scene sc;
    int numIterations = 10, tpb = 128;                       // context.cpp:331,345
    if (root.contains("num of iterations")) numIterations = (int)root.at("num of iterations").number();
    if (root.contains("threads per block")) tpb = (int)root.at("threads per block").number();

    std::map<std::string, const Json*> softDefs, fixedDefs;
    if (root.contains("softBodies")) for (auto& j : root.at("softBodies").arr) softDefs[j.at("name").string()] = &j;
    if (root.contains("fixedBodies")) for (auto& j : root.at("fixedBodies").arr) fixedDefs[j.at("name").string()] = &j;

    const Json* ctx = nullptr;
    if (root.contains("contexts"))
        for (auto& c : root.at("contexts").arr) {
            const std::string nm = c.at("name").string();
            if (contextName.empty()) { if (c.value("load", true)) { ctx = &c; break; } }
            else if (nm == contextName) { ctx = &c; break; }
        }
    if (!ctx) throw std::runtime_error("context not found: '" + contextName + "'");

    sc.name = ctx->at("name").string();
    sc.precision = ctx->value("precision", "double");        // simulationContext.cpp:56-59
    SolverParams& p = sc.params;
    p.numIterations = numIterations;
    p.threadsPerBlock = tpb;
    // Impl::Init, simulationContext.cu:48-58 (only present keys override the def.h defaults)
    if (ctx->contains("dt")) p.dt = (float)ctx->at("dt").number();
    if (ctx->contains("tolerance")) p.tol = (float)ctx->at("tolerance").number();
    if (ctx->contains("maxIterations")) p.maxIterations = (int)ctx->at("maxIterations").number();
    if (ctx->contains("dhat")) p.dhat = (float)ctx->at("dhat").number();
    if (ctx->contains("gravity")) p.gravity = (float)ctx->at("gravity").number();
    if (ctx->contains("damp")) p.damp = (float)ctx->at("damp").number();
    if (ctx->contains("muN")) p.muN = (float)ctx->at("muN").number();
    if (ctx->contains("muT")) p.muT = (float)ctx->at("muT").number();
    // optional keys the reference ignores
    if (ctx->contains("rho")) p.rho = (float)ctx->at("rho").number();
    if (ctx->contains("globalSolver")) p.globalSolver = (int)ctx->at("globalSolver").number();
    if (ctx->contains("handleCollision")) p.handleCollision = ctx->at("handleCollision").boolean() ? 1 : 0;
    if (p.tol < 1e-6f) p.tol = 1e-6f;                         // simulationContext.cpp:28-31

    if (ctx->contains("softBodies")) {
        for (auto& sb : ctx->at("softBodies").arr) {
            const std::string nm = sb.at("name").string();
            auto it = softDefs.find(nm);
            if (it == softDefs.end()) throw std::runtime_error("soft body definition not found: " + nm);
            const Json& def = *it->second;
            const std::string nodeFile = def.value("nodeFile", ""), eleFile = def.value("eleFile", "");
            if (!def.value("mshFile", "").empty() && nodeFile.empty())
                throw std::runtime_error("Gmsh .msh input is out of scope (no shipped asset uses it): " + nm);
            if (nodeFile.empty()) throw std::runtime_error("Msh or node file must be provided!!!");
            float pos[3], scale[3], rot[3];
            vec3_pick(sb, def, "pos", pos);
            vec3_pick(sb, def, "scale", scale);
            vec3_pick(sb, def, "rot", rot);
            const float mass = (float)sb.value("mass", def.value("mass", 1.0));
            const float mu = (float)sb.value("mu", def.value("mu", 1000.0));
            const float lambda = (float)sb.value("lambda", def.value("lambda", 1000.0));
            std::vector<uint32_t> dbc;
            const Json* dsrc = sb.contains("DBC") ? sb.find("DBC") : def.find("DBC");
            if (dsrc && !dsrc->is_null()) for (auto& d : dsrc->arr) dbc.push_back((uint32_t)d.number());
            const bool centralize = def.value("centralize", false);
            const int startIndex = (int)def.value("start index", 0.0);

            std::vector<float> X = load_node_file(resolve_asset(nodeFile, jsonDir, assetRoot), centralize);
            float M[16];
            model_matrix(pos, rot, scale, /*softBodyOrder=*/true, M);
            transform_vertices(X.data(), (int)(X.size() / 3), M);
            std::vector<uint32_t> T = load_ele_file(resolve_asset(eleFile, jsonDir, assetRoot), startIndex);
            std::string base = nodeFile.substr(nodeFile.find_last_of('/') + 1);
            const std::string faceFile = def.value("faceFile", "");
            std::vector<uint32_t> Fc;
            if (!faceFile.empty()) {
                std::string fp;
                try { fp = resolve_asset(faceFile, jsonDir, assetRoot); } catch (...) { fp.clear(); }
                if (!fp.empty()) Fc = load_face_file(fp, startIndex);
            }
            scene_add_body(sc, base, X, T, mass, mu, lambda, dbc, &Fc);
        }
    }
    if (ctx->contains("fixedBodies")) {
        for (auto& fb : ctx->at("fixedBodies").arr) {
            const std::string nm = fb.at("name").string();
            auto it = fixedDefs.find(nm);
            if (it == fixedDefs.end()) throw std::runtime_error("fixed body definition not found: " + nm);
            const Json& def = *it->second;
            float pos[3], scale[3], rot[3];
            vec3_pick_default(fb, def, "pos", 0.f, pos);
            vec3_pick_default(fb, def, "scale", 1.f, scale);
            vec3_pick_default(fb, def, "rot", 0.f, rot);
            const std::string type = def.at("type").string();
            FixedBody b;
            b.name = nm;
            if (type == "sphere") {
                float radius = 1.f;
                bool has = false;
                if (fb.contains("radius")) { radius = (float)fb.at("radius").number(); has = true; }
                else if (def.contains("radius")) { radius = (float)def.at("radius").number(); has = true; }
                const float s3[3] = {has ? radius : 1.f, has ? radius : 1.f, has ? radius : 1.f};
                b.type = FB_SPHERE; b.radius = radius;
                model_matrix(pos, rot, s3, false, b.model);
            } else if (type == "cylinder") {
                const float s3[3] = {scale[0], scale[1], scale[0]};   // context.cpp:291
                b.type = FB_CYLINDER; b.radius = scale[0];
                model_matrix(pos, rot, s3, false, b.model);
            } else if (type == "plane") {
                b.type = FB_PLANE;
                model_matrix(pos, rot, scale, false, b.model);
            } else {
                throw std::runtime_error("unknown fixed body type: " + type);
            }
            sc.fixed.push_back(b);
        }
    }
    return sc;
}

// ---------------------------------------------------------------- synthetic Kuhn grid
Scene make_kuhn_grid(int nx, int ny, int nz, float h, float jitter, uint32_t seed, const float origin[3],
                     float mass, float mu)
{
    Scene sc;
    sc.name = "kuhn_grid_" + std::to_string(nx) + "x" + std::to_string(ny) + "x" + std::to_string(nz);
    const int vx = nx + 1, vy = ny + 1, vz = nz + 1;
    const size_t nV = (size_t)vx * vy * vz, nT = (size_t)6 * nx * ny * nz;
    std::vector<float> X(nV * 3);
    std::mt19937 rng(seed);
    auto uni = [&]() { return (float)(rng() >> 8) * (1.0f / 16777216.0f); };   // [0,1), 24 bits
    for (int k = 0; k < vz; ++k)
        for (int j = 0; j < vy; ++j)
            for (int i = 0; i < vx; ++i) {
                const size_t v = ((size_t)k * vy + j) * vx + i;
                const float jx = jitter * (2.0f * uni() - 1.0f);
                const float jy = jitter * (2.0f * uni() - 1.0f);
                const float jz = jitter * (2.0f * uni() - 1.0f);
                X[3 * v + 0] = origin[0] + (float)i * h + jx;
                X[3 * v + 1] = origin[1] + (float)j * h + jy;
                X[3 * v + 2] = origin[2] + (float)k * h + jz;
            }
    // six tets around the main diagonal of each cell; odd permutations get two vertices swapped
    // so that every tet is positively oriented
    static const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {2, 1, 0}, {1, 0, 2}};
    std::vector<uint32_t> T(nT * 4);
    size_t t = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i)
                for (int pidx = 0; pidx < 6; ++pidx) {
                    int c[3] = {i, j, k};
                    uint32_t vid[4];
                    auto id = [&](const int q[3]) { return (uint32_t)(((size_t)q[2] * vy + q[1]) * vx + q[0]); };
                    vid[0] = id(c);
                    for (int s = 0; s < 3; ++s) { c[perm[pidx][s]] += 1; vid[s + 1] = id(c); }
                    if (pidx >= 3) std::swap(vid[1], vid[2]);
                    for (int s = 0; s < 4; ++s) T[4 * t + s] = vid[s];
                    ++t;
                }
    scene_add_body(sc, sc.name, X, T, mass, mu, 5000.f, {});
    return sc;
}

void write_tetgen(const Scene& s, const std::string& nodePath, const std::string& elePath)
{
    FILE* f = std::fopen(nodePath.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + nodePath);
    std::fprintf(f, "%d  3  0  0\n", s.numVerts);
    for (int i = 0; i < s.numVerts; ++i)
        std::fprintf(f, "%d  %.9g  %.9g  %.9g\n", i + 1, s.X[3 * i], s.X[3 * i + 1], s.X[3 * i + 2]);
    std::fclose(f);
    f = std::fopen(elePath.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + elePath);
    std::fprintf(f, "%d  4  0\n", s.numTets);
    for (int t = 0; t < s.numTets; ++t)
        std::fprintf(f, "%d  %u  %u  %u  %u\n", t + 1, s.Tet[4 * t] + 1, s.Tet[4 * t + 1] + 1, s.Tet[4 * t + 2] + 1, s.Tet[4 * t + 3] + 1);
    std::fclose(f);
}

}  // namespace pdb200
