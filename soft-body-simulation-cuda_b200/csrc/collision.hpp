// Mesh-mesh collision of the PD path, host side.  Reference behaviour restated:
//   * surface triangles and their "father" bodies: DataLoader::loadEleFaceFile / AllocData (dataLoader.cu:68-127, 343-369) --
//     a body's .face file when the scene names one, else the faces that belong to exactly one of its tets, listed in the
//     order of their sorted vertex triples with the winding of the tet they come from;
//   * the queries CollisionDetection::BroadPhaseCCD forms for a pair of triangles whose swept boxes overlap (broadphase.cu:
//     267-305, 453-478): 3 vertex-face and 9 edge-edge tests with their vertex ids sorted the reference's way.
// What is NOT restated is the reference's LBVH rebuild per step (Morton codes, thrust sort, Karras splits, bvh.cu /
// broadphase.cu:28-211): the SET of overlapping leaf pairs does not depend on the tree, so this engine keeps ONE tree per
// body, built here once from the rest shape (median splits of the triangle centroids), and only refits its boxes every step.
#pragma once
#include <cstdint>
#include <vector>

#include "scene.hpp"

namespace pdb200 {

// explicit .face triangles where the scene has them, derived boundary faces elsewhere; ids in the scene's (original) numbering
void scene_surface(const Scene& s, std::vector<uint32_t>& tri, std::vector<uint32_t>& father);
// boundary faces of a tet range (dataLoader.cu:92-127): unique faces in ascending order of their sorted vertex triples
void boundary_faces(const uint32_t* Tet, int t0, int t1, std::vector<uint32_t>& tri);

struct CollisionMesh {
    int nTris = 0, nEdges = 0, nBodies = 0;
    std::vector<uint32_t> tri;        // 3 * nTris, original vertex ids, LEAF order (leaf k of the tree = triangle k here)
    std::vector<uint32_t> father;     // nTris: soft body of each triangle (SolverData::dev_TriFathers)
    std::vector<uint32_t> edge;       // 2 * nEdges: sorted endpoints, edges in ascending (v0, v1) order
    std::vector<uint32_t> triEdge;    // 3 * nTris: edge ids of the local edges (0,1), (0,2), (1,2) (edgeIndicesTable, broadphase.cu:246)
    // one binary tree per body over its triangles; node ids: internal nodes 0 .. nInternal-1, leaf k = nInternal + k
    int nInternal = 0;
    std::vector<int> left, right;     // nInternal: children (node ids)
    std::vector<int> parent;          // nInternal + nTris: parent node id, -1 for a root
    std::vector<int> bodyRoot;        // nBodies: root node of each body's tree (a leaf id when the body has one triangle), -1 = no triangles
};
// X: rest positions (3 * numVerts, original numbering) for the median splits
void build_collision_mesh(const Scene& s, CollisionMesh& M);

}  // namespace pdb200
