// Host-side scene description: the merged soft bodies, fixed bodies and solver parameters
// that the reference assembles in SimulationCUDAContext::Impl<float>::Init
// (src/simulation/simulationContext.cu:34-123) and DataLoader (src/simulation/dataLoader.cu).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pdb200 {

enum FixedBodyType { FB_PLANE = 0, FB_SPHERE = 1, FB_CYLINDER = 2 };

struct FixedBody {
    int type = FB_PLANE;
    float model[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // glm column-major
    float radius = 1.f;   // Sphere::m_radius / Cylinder::m_radius (= scale.x)
    std::string name;
};

// def.h:82-100 SolverParams<float> (+ the global-solver choice the reference keeps in PdSolver)
struct SolverParams {
    float dt = 0.001f, damp = 0.999f, muN = 0.5f, muT = 0.5f, gravity = 9.8f;
    float dhat = 1e-2f, tol = 1e-2f, rho = 0.9992f;
    int numIterations = 1, maxIterations = 100;
    int handleCollision = 0;
    int globalSolver = 0;          // PdSolver::SolverType: 0 Jacobi, 1 Cholesky, 2 PCG-Jacobi
    int pcgMaxIter = 2000;         // pcgJacobi.h defaults
    float pcgTol = 1e-5f;
    int threadsPerBlock = 128;     // context.json "threads per block" (accepted, unused)
};

struct Scene {
    std::string name, precision = "float";
    int numVerts = 0, numTets = 0;
    std::vector<float> X;          // 3*numVerts, transformed rest positions (AoS xyz)
    std::vector<uint32_t> Tet;     // 4*numTets, merged numbering
    std::vector<float> mass, DBC;  // numVerts
    std::vector<float> mu, lambda; // numTets
    std::vector<int> bodyVertStart, bodyTetStart;   // per soft body (startIndices in the reference)
    std::vector<std::string> bodyNames;
    // surface triangles for the mesh-mesh collision pass (SolverData::Tri / dev_TriFathers, dataLoader.cu:68-127,343-369):
    // explicit ones (a body's .face file, or the caller's arrays for the whole scene); bodies without explicit triangles get
    // the boundary faces of their tets when a collision mesh is built (collision.cpp:scene_surface) -- never before, a 16 M-tet
    // grid should not pay for it
    std::vector<uint32_t> Tri, triFather;      // 3 per triangle (merged numbering); body of each triangle
    std::vector<uint8_t> bodyHasTri;           // per body: its triangles are in Tri
    bool triWholeScene = false;                // Tri / triFather ARE the scene's surface as given by the caller
    std::vector<FixedBody> fixed;
    SolverParams params;
};

// TetGen readers, dataLoader.cu:131-173 and :38-66
std::vector<float> load_node_file(const std::string& path, bool centralize);
std::vector<uint32_t> load_ele_file(const std::string& path, int startIndex);
std::vector<uint32_t> load_face_file(const std::string& path, int startIndex);      // dataLoader.cu:69-90 (empty when the file cannot be opened)

// utilities.cpp:141-150 (fixed bodies: T*Rx*Ry*Rz*S) / dataLoader.cu:214-220 (soft: T*S*Rx*Ry*Rz)
void model_matrix(const float pos[3], const float rot[3], const float scale[3], bool softBodyOrder, float M[16]);
void transform_vertices(float* X, int nV, const float M[16]);
void plane_up(const float M[16], float up[3]);        // rigid/plane.cpp:9
void cylinder_axis(const float M[16], float axis[3]); // fixedBodyData.cu:116

// Append one soft body (already transformed) to a scene: DataLoader::AllocData merge rules
void scene_add_body(Scene& s, const std::string& name, const std::vector<float>& X, const std::vector<uint32_t>& Tet,
                    float mass, float mu, float lambda, const std::vector<uint32_t>& dbc, const std::vector<uint32_t>* faces = nullptr);

// context.json -> Scene.  contextName empty = first context with "load" != false.
// assetRoot empty = resolve asset paths like the reference (relative to the json's build dir).
Scene load_context_json(const std::string& jsonPath, const std::string& contextName, const std::string& assetRoot);
std::vector<std::string> list_contexts(const std::string& jsonPath);

// Synthetic Kuhn 6-tet grid: nx*ny*nz cells of size h,
// vertices jittered by U(-jitter, jitter) from mt19937(seed), origin at `origin`.
Scene make_kuhn_grid(int nx, int ny, int nz, float h, float jitter, uint32_t seed, const float origin[3],
                     float mass, float mu);
// writes TetGen .node/.ele (start index 1) so other tools see identical input
void write_tetgen(const Scene& s, const std::string& nodePath, const std::string& elePath);

}  // namespace pdb200
