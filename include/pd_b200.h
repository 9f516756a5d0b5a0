/*
 * pd_b200.h -- C ABI of the B200-native projective-dynamics (PD) step engine.
 *
 * Drop-in boundary for the float PD path of GrahamZen/Soft-Body-Simulation-CUDA.  The
 * reference has no C ABI; its boundary is the C++ plugin interface
 *     template<typename Scalar> class Solver            src/simulation/solver/solver.h:11-28
 *     class PdSolver : public FEMSolver<float>          src/simulation/solver/projective/pdSolver.h:12-44
 * driven by SimulationCUDAContext (src/simulation/simulationContext.cpp:79-114).  Every entry
 * point below names the reference interface it replaces.  INTEGRATION.md shows the adapter
 * (`class B200PdSolver : public Solver<float>`, include/b200_pd_solver.h) a maintainer adds.
 *
 * Conventions: plain pointers and sizes only; all functions returning int return 0 on success
 * and a negative pd_status otherwise; the message is available from pd_last_error() (thread
 * local).  One engine = one simulation context on one GPU, driven by one host thread at a
 * time, on its own non-blocking CUDA stream.  There is no CPU fallback: creating an engine
 * without a sm_100 device fails loudly.
 *
 * Host arrays use the reference's layouts: positions/velocities are AoS float[3*numVerts]
 * (glm::vec3), tets are uint32[4*numTets] (indexType, def.h:4), original vertex numbering.
 */
#ifndef PD_B200_H
#define PD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pd_engine pd_engine;   /* one SimulationCUDAContext::Impl<float> + PdSolver        */
typedef struct pd_scene pd_scene;     /* host-side merged scene (DataLoader output), no GPU needed */
typedef struct pd_layout pd_layout;   /* host-side device-layout builder output, no GPU needed     */
typedef struct pd_rank_plan pd_rank_plan; /* host-side multi-GPU plan of one rank, no GPU needed    */

enum pd_status {
    PD_OK = 0,
    PD_ERR_INVALID = -1,      /* bad argument                                   */
    PD_ERR_IO = -2,           /* file / parse error                             */
    PD_ERR_CUDA = -3,         /* CUDA runtime error or no usable device         */
    PD_ERR_UNSUPPORTED = -4   /* feature outside the PD hot path                */
};

/* PdSolver::SolverType, pdSolver.h:15-18 (+ the PCG back-end of linear/pcgJacobi.cu) */
enum pd_global_solver { PD_JACOBI = 0, PD_CHOLESKY = 1, PD_PCG_JACOBI = 2 };

/* FixedBody subclasses, src/collision/rigid/{plane,sphere,cylinder}.h */
enum pd_fixed_body_type { PD_PLANE = 0, PD_SPHERE = 1, PD_CYLINDER = 2 };

typedef struct {
    int type;          /* pd_fixed_body_type                                                  */
    float model[16];   /* FixedBody::m_model, glm column-major (utilities.cpp:141-150)        */
    float radius;      /* Sphere::m_radius / Cylinder::m_radius ; unused for planes           */
} pd_fixed_body;

/* SolverParams<float>, src/def.h:82-100, plus the solver choice of PdSolver::SetGlobalSolver */
typedef struct {
    float dt, gravity, muN, muT, rho, tol, damp;
    int num_iterations;       /* context.json "num of iterations"                             */
    int global_solver;        /* pd_global_solver                                             */
    int pcg_max_iter;         /* PCGJacobiSolver max_iter (pcgJacobi.h)                       */
    float pcg_tol;            /* PCGJacobiSolver tolerance on ||r||_2                         */
    int handle_collision;     /* SolverParams::handleCollision: 1 = the mesh-mesh collision pass of PdSolver::Update
                               * (DetectCollision + CCDKernel, pdSolver.cu:218-225) between the velocity update and the
                               * fixed bodies; single-GPU engines */
    int threads_per_block;    /* accepted for interface parity, ignored                       */
} pd_params;

/* SolverData<float> as DataLoader::AllocData leaves it (dataLoader.cu:291-378), host side */
typedef struct {
    int num_verts, num_tets;
    const float* X;           /* 3*num_verts rest positions (already transformed)             */
    const uint32_t* Tet;      /* 4*num_tets                                                   */
    const float* mass;        /* num_verts  (SoftBodyAttribute::mass per vertex)              */
    const float* mu;          /* num_tets                                                     */
    const float* DBC;         /* num_verts, 1 = pinned (may be NULL = none)                   */
    int num_fixed;
    const pd_fixed_body* fixed;
    /* surface triangles for the mesh-mesh collision pass (SolverData::numTris / Tri / dev_TriFathers, dataLoader.cu:343-369):
     * num_tris = 0 or Tri = NULL: the boundary faces of the tets, ONE father body; TriFathers NULL: father 0 everywhere */
    int num_tris;
    const uint32_t* Tri;          /* 3*num_tris */
    const uint32_t* TriFathers;   /* num_tris: soft body of each triangle (pairs of the same father are skipped) */
} pd_scene_desc;

#define PD_ROT_AUTO (-1)
typedef struct {
    int device;               /* CUDA device ordinal                                          */
    int rot_mode;             /* PD_ROT_AUTO (default): the bit-faithful mode (1, input tet order) on a mesh that fits one
                               * tile (<= 256 tets: one warp-sized launch of the reference is deterministic and the mode
                               * costs nothing there), else 0;  0 Newton polar + SVD fallback;  1 always the reference's
                               * Jacobi SVD, sums in the reference's order (bit-exact parity mode, ~4x slower local step) */
    int reorder;              /* 1 (default) Morton tet order + first-touch vertex renumbering */
    int use_graph;            /* 1 (default) one CUDA graph per step                          */
    int ctas_per_sm;          /* 0 = occupancy query                                          */
    int rank, world;          /* multi-GPU: this engine is rank `rank` of `world` (default 0 of 1); one process per GPU */
    int body_kernel;          /* -1 (default) auto: a scene whose connected components all fit one CTA's shared memory (the
                               * batches of small bodies of BASELINE config 5) steps with one CTA per body and ONE launch per
                               * step; 0: always the tile kernels */
} pd_engine_options;

/* PdSolver::GetPerformanceData (solver.h:20, pdSolver.cu:26): the four named counters, ms */
typedef struct {
    float local_step_ms, global_step_ms, collision_fixed_ms, collision_mesh_ms;
    double step_ms_total;
    long long steps, pd_iterations, inner_iterations, kernel_launches;
} pd_perf;

const char* pd_last_error(void);
const char* pd_version(void);
void pd_default_params(pd_params* p);                 /* def.h:82-100 defaults                 */
void pd_default_options(pd_engine_options* o);

/* ---- scene (host only) : Context::LoadSimContext + Impl::Init + DataLoader -------------- */
/* context.cpp:319-387, simulationContext.cu:34-123.  context_name NULL/"" = first loadable.
 * asset_root NULL = resolve "../assets/..." like the reference does from its build dir.    */
pd_scene* pd_scene_load_json(const char* json_path, const char* context_name, const char* asset_root);
pd_scene* pd_scene_from_desc(const pd_scene_desc* desc, const pd_params* params);
/* the surface the collision pass would use: *num_tris is always set; tri (3 per triangle) / father are filled when not NULL */
int pd_scene_get_surface(const pd_scene*, int* num_tris, uint32_t* tri, uint32_t* father);
/* synthetic Kuhn 6-tet grid (bench configs 3/4) */
pd_scene* pd_scene_kuhn_grid(int nx, int ny, int nz, float h, float jitter, uint32_t seed,
                             const float origin[3], float mass, float mu);
/* batch of independent contexts (same solver parameters and fixed bodies) -> one scene; vertex/tet numbering is the
 * concatenation in argument order.  Context::mpSimContexts holds several contexts but steps one (context.cpp:546-555);
 * stepping many at once is the batch mode of BASELINE config 5. */
pd_scene* pd_scene_merge(const pd_scene* const* scenes, int n);
void pd_scene_free(pd_scene*);
int pd_scene_counts(const pd_scene*, int* num_verts, int* num_tets, int* num_fixed, int* num_bodies);
int pd_scene_get(const pd_scene*, float* X, uint32_t* Tet, float* mass, float* mu, float* DBC,
                 pd_fixed_body* fixed, int* body_vert_start);
int pd_scene_get_params(const pd_scene*, pd_params* out);
int pd_scene_set_params(pd_scene*, const pd_params* in);
int pd_scene_add_fixed(pd_scene*, const pd_fixed_body* fb);
int pd_scene_write_tetgen(const pd_scene*, const char* node_path, const char* ele_path);
/* dataLoader.cu:131-173 / :38-66 ; caller frees with pd_free */
int pd_load_node(const char* path, int centralize, float** X, int* num_verts);
int pd_load_ele(const char* path, int start_index, uint32_t** Tet, int* num_tets);
void pd_free(void*);
/* utilities.cpp:141-150 / dataLoader.cu:214-220, rigid/plane.cpp:9 */
void pd_model_matrix(const float pos[3], const float rot_deg[3], const float scale[3], int soft_body_order, float M[16]);
void pd_transform_vertices(float* X, int num_verts, const float M[16]);
void pd_plane_up(const float M[16], float up[3]);

/* ---- device layout (host only) : bit-exact partition / reorder / incidence checks -------- */
pd_layout* pd_layout_build(const pd_scene*, int reorder);
void pd_layout_free(pd_layout*);
int pd_layout_counts(const pd_layout*, int* num_tiles, uint32_t* num_slots, size_t* record_bytes, int* max_local);
int pd_layout_get(const pd_layout*, uint32_t* tet_order, uint32_t* vert_order, uint32_t* tet_new,
                  uint32_t* tile_tet_start, uint64_t* tile_rec_off, uint8_t* records,
                  uint32_t* vslot_ptr, uint32_t* vslot, uint32_t* vlist);
/* staging slot -> vertex map of every tile (num_tiles * 256 entries, 0xffffffff = empty): what the local kernel's
   position gather reads; the corner words of the tet records hold staging slots */
int pd_layout_get_vstage(const pd_layout*, uint32_t* vstage);
/* the device tile table the local kernel reads: num_tiles * 12 words (csrc/layout.hpp, TILE_META_WORDS) */
int pd_layout_tile_table(const pd_layout*, uint32_t* table);
/* matrix_diag of SolverPrepare (computeSiTSi, pdUtil.cu:16-24) as the engine computes it on the host, per LAYOUT vertex
 * (renumbered ids of a global layout, local ids of a rank's layout): sums in ascending global tet order on every world size */
int pd_layout_matrix_diag(const pd_layout*, float* matrix_diag /* layout's vertex count */);
int pd_morton_keys(const float* X, const uint32_t* Tet, int num_tets, uint32_t* keys);
int pd_partition_vertices(int num_verts, int world, int* vbeg /* world+1 */);
/* host-side prefactorisation of the small-mesh path: sparse Cholesky A = L L^T of a symmetric CSR matrix in the given
 * order (replaces cusolverSpXcsrcholAnalysis/Factor, cholesky.cu:152-157, and Eigen::SimplicialCholesky, pdSolver.cu:103).
 * L is returned by rows (ascending columns, diagonal last); free the three arrays with pd_free. */
int pd_cholesky_factor(int n, const int* rowptr, const int* col, const float* val, int* nnz_l, int** lptr, int** lcol, float** lval);
/* the fill-reducing order the engine factors in (the reference: AMD inside cusolverSpXcsrcholAnalysis / SimplicialCholesky,
 * cholesky.cu:72-131): geometric nested dissection of the matrix graph on the rows' rest positions xyz (n x 3).
 * perm[new] = old.  nnz_l_natural / nnz_l_ordered (may be NULL): non-zeros of L without and with the order (symbolic). */
int pd_nested_dissection(int n, const int* rowptr, const int* col, const float* xyz, int* perm, int* nnz_l_natural, int* nnz_l_ordered);

/* ---- multi-GPU plan (host only) : vertex partition, tile selection, ghosts, push lists --- new work,
 * the reference is single-GPU.  Checked bit for bit across ranks by the gloo tests. */
pd_rank_plan* pd_rank_plan_build(const pd_layout* global_layout, int world, int rank);
void pd_rank_plan_free(pd_rank_plan*);
/* counts[7] = {num_owned, num_ghosts, num_tiles, num_neighbours, num_push, first_owned_vertex (renumbered id),
 *              num_interior_tiles (they come first in `tiles`; only the tiles after them read ghost positions)} */
int pd_rank_plan_counts(const pd_rank_plan*, int counts[7]);
int pd_rank_plan_get(const pd_rank_plan*, uint32_t* tiles, uint32_t* ghosts, int* neighbours, int* n_loc_of /* world */,
                     uint32_t* push_src, uint32_t* push_dst, int* push_rank);
/* the rank's own layout (selected tiles, local vertex ids); vert_order maps local -> ORIGINAL vertex ids */
pd_layout* pd_rank_layout(const pd_layout* global_layout, const pd_rank_plan*);

/* ---- engine : PdSolver behind SimulationCUDAContext -------------------------------------- */
/* PdSolver::PdSolver + FEMSolver ctor (pdSolver.cu:22-27, femSolver.cu:6-17) */
pd_engine* pd_create(const pd_scene*, const pd_engine_options* opt /* NULL = defaults */);
pd_engine* pd_create_from_json(const char* json_path, const char* context_name, const char* asset_root,
                               const pd_engine_options* opt);
void pd_destroy(pd_engine*);
/* SimulationCUDAContext::Update -> PdSolver::Update (simulationContext.cpp:79-86, pdSolver.cu:210-232) */
int pd_step(pd_engine*, int n_steps);
int pd_synchronize(pd_engine*);
/* pd_step bracketed by CUDA events on the engine's stream; *device_ms = elapsed device time */
int pd_step_timed(pd_engine*, int n_steps, float* device_ms);
/* CopyUIToParams before every Update (simulationContext.cpp:17-35,85) */
int pd_set_params(pd_engine*, const pd_params*);
int pd_get_params(const pd_engine*, pd_params*);
/* SimulationCUDAContext::SetGlobalSolver (simulationContext.cpp:104-114) */
int pd_set_global_solver(pd_engine*, int solver);
/* SimulationCUDAContext::Reset + Solver::Reset (simulationContext.cu:233-243, solver.h:39-43) */
int pd_reset(pd_engine*);
/* Solver::SetPerf / GetPerformanceData (solver.h:19-20) */
int pd_set_perf(pd_engine*, int on);
int pd_get_perf(const pd_engine*, pd_perf* out);
/* state in the reference's host layout; any pointer may be NULL */
int pd_download(pd_engine*, float* X, float* V, float* XTilde);
int pd_upload_state(pd_engine*, const float* X, const float* V, const float* XTilde);
/* end to end on HOST buffers: upload state, n steps, download state (all inside the call) */
int pd_step_host(pd_engine*, int n_steps, const float* X_in, const float* V_in, const float* XTilde_in,
                 float* X_out, float* V_out, float* XTilde_out);
/* the same for ONE RANK'S SHARD of a multi-GPU engine: every array holds 3 * num_owned floats in the rank's local
 * owned-vertex order; pd_dist_owned_ids gives the original vertex id of each entry (num_owned of them).  Each
 * process keeps and moves only its part of SolverData<float>::X/V/XTilde (def.h:21-27). */
int pd_step_host_owned(pd_engine*, int n_steps, const float* X_in, const float* V_in, const float* XTilde_in,
                       float* X_out, float* V_out, float* XTilde_out);
int pd_dist_owned_ids(const pd_engine*, uint32_t* original_ids /* num_owned */);
/* adopt the reference's DEVICE arrays (SolverData<float>::X/V/XTilde, glm::vec3*, original
 * numbering): import -> n steps -> export, all on the engine's stream, then synchronise.
 * This is what B200PdSolver::Update calls (include/b200_pd_solver.h). */
int pd_update_device(pd_engine*, int n_steps, float* dX, float* dV, float* dXTilde);
/* Mouse-drag soft constraints: SolverData<float>::moreDBC[num_verts] / OffsetX[3*num_verts] (def.h:31-32) and
 * MouseSelection::target (def.h:14-18), written by Control_Kernel between two Updates (simulationContext.cu:202-231)
 * and consumed by PdSolver at pdUtil.cu:56-69,80-87,159-164,187-188,201-206: a vertex with moreDBC > 0 is held at
 * target + OffsetX for the whole step with zero velocity, and its massDt_2s becomes (m + moreDBC)/dt^2.
 * Original vertex numbering.  more_dbc NULL (or no positive entry) ends the drag (ResetMoreDBC(true)); pd_reset clears
 * it like SimulationCUDAContext::Reset (simulationContext.cu:240).  Single-GPU engines only (PD_ERR_UNSUPPORTED else). */
int pd_set_drag(pd_engine*, const float* more_dbc, const float* offset_x, const float target[3]);
/* the same from the reference's DEVICE arrays (what B200PdSolver::Update passes while mouseSelection.dragging) */
int pd_set_drag_device(pd_engine*, const float* d_more_dbc, const float* d_offset_x, const float target[3]);
/* Control_Kernel itself (simulationContext.cu:202-218: RADIUS_SQUARED 0.002, control_mag 10 at :229) on the engine's
 * current X: select_v = -1 clears; for headless callers that have no SimulationCUDAContext */
int pd_drag_select(pd_engine*, int select_v, float control_mag, const float target[3]);
/* parity checks: the engine's moreDBC / OffsetX / DBCX (computeSn overwrites DBCX of dragged vertices, pdUtil.cu:86) */
int pd_get_drag(pd_engine*, float* more_dbc, float* offset_x, float* dbcx, int* active);
/* Live stiffness edit: SolverData<float>::mu[num_tets] changed (SimulationCUDAContext::UpdateSoftBodyAttr -> FillData,
 * simulationContext.cu:165-176; original tet order).  As in the reference the new mu acts from the next Update on
 * (computeLocal reads it every iteration, pdUtil.cu:97,124) while matrix_diag and the assembled system matrix stay
 * SolverPrepare products: they follow only after pd_reset (pdSolver.cu:212-216). */
int pd_update_mu(pd_engine*, const float* mu);
int pd_update_mu_device(pd_engine*, const float* d_mu);
/* setup products for parity checks (original numbering): matrix_diag, massDt_2s, DmInv(9/tet,row-major), V0 */
int pd_get_setup(pd_engine*, float* matrix_diag, float* mass_dt2, float* DmInv, float* V0);
/* the scalar system matrix A^ = M/h^2 + sum_t w_t S^T (DmInv^T G)^T (DmInv^T G) S that SolverPrepare assembles as COO
 * (pdSolver.cu:62-77, pdUtil.cu:9-54), deduplicated CSR with ascending columns, in the engine's RENUMBERED vertex ids
 * (pd_layout_get: vert_order maps renumbered -> original).  Call with NULL arrays to get nnz first. */
int pd_get_system_matrix(pd_engine*, int* nnz, int* rowptr, int* col, float* val);
/* direct / CG modes: computeError of the last PD iteration (pdSolver.cu:243-253) and the PD iterations the last step ran
 * before sqrt(err) < tol (pdSolver.cu:164) */
int pd_get_solve_stats(pd_engine*, float* err, int* pd_iterations_last_step);
/* Mesh-mesh collision (pd_params.handle_collision = 1; PdSolver::Update, pdSolver.cu:218-225: DetectCollision + CCDKernel between
 * SolverStep and the fixed-body response).  The pass runs inside pd_step / pd_update_device on the scene's surface triangles
 * (pd_scene_desc.Tri or the tets' boundary faces; pairs of the same father body are skipped, as PdSolver asks for).  This
 * reads back what the LAST pass left in SolverData::dev_tIs / dev_Normals: tI[v] = 1 (free) or 0.5 (in a detected contact: the
 * vertex kept its position and got V = -(n . dx) n), the contact normals (3 per vertex), and the number of triangle pairs
 * whose swept boxes overlapped.  Any pointer may be NULL.  Single-GPU engines only. */
int pd_get_collision(pd_engine*, float* tI, float* normals, long long* num_pairs);
/* sizes of the direct / CG modes' setup products (0 until the mode has been prepared): non-zeros of the scalar system matrix A^
 * and of its Cholesky factor L in the nested-dissection order */
int pd_get_solver_sizes(pd_engine*, long long* nnz_A, long long* nnz_L);
/* measurement helpers used by bench.py: average device time (ms) of one launch of the local /
 * vertex kernel over `reps` back-to-back launches, CUDA events on the engine's stream */
int pd_time_kernels(pd_engine*, int reps, float* local_ms, float* vertex_ms);
/* measurement helper: clock64 totals per phase of the local kernel, 8 x u64 per CTA (local_grid CTAs):
 * [phase B, record fetch + gather wait, barrier, gather issue + producer + table loads, wait part C, phase C,
 *  unused, tiles processed] */
int pd_profile_local(pd_engine*, unsigned long long* out);
int pd_engine_info(const pd_engine*, int* num_verts, int* num_tets, int* num_tiles, uint32_t* num_slots,
                   size_t* tile_stream_bytes, size_t* device_bytes, int* local_grid);
/* the rotation mode the engine runs (PD_ROT_AUTO resolved): 0 or 1 */
int pd_engine_rot_mode(const pd_engine*);
/* ---- multi-GPU engine (options.world > 1): every rank creates its engine from the SAME scene, exchanges the
 * 64-byte window handles (e.g. torch.distributed.all_gather) and connects; pd_step then runs the ranks in
 * lock step through the phase tags of the halo positions in peer memory (no host synchronisation, no NCCL on the data path).
 * pd_download returns the rank's OWN vertices and zeros elsewhere (sum the ranks' arrays to combine). */
int pd_dist_window_handle(pd_engine*, void* out64);                       /* cudaIpcMemHandle_t */
int pd_dist_connect(pd_engine*, const void* handles /* world x 64 bytes, rank order */);
int pd_dist_connect_local(pd_engine* const* engines, int n);              /* all ranks in one process (tests) */
int pd_dist_step_lockstep(pd_engine* const* engines, int n, int n_steps); /* one process drives all ranks phase by phase */
int pd_dist_status(pd_engine*, unsigned int* halo_wait_timed_out);
/* info[6] = {num_owned, num_ghosts, num_neighbours, num_push, tets evaluated here, tiles evaluated here} */
int pd_dist_info(const pd_engine*, int info[6]);

/* test hook: the corotational projection (pdUtil.cu:112-122) of n row-major 3x3 matrices on
 * `device`; rot_mode as in pd_engine_options; used_fast (may be NULL) reports the path taken */
int pd_rotation_batch(int device, int rot_mode, int n, const float* F, float* R, int* used_fast);
/* ---- linear back-ends of the reference's IPC (double) solver, on the engine's fused CSR / CG kernels ---------------------------
 * LinearSolver<double>::Solve(N, d_b, d_x, A, nz, rowIdx, colIdx, d_guess) (src/simulation/solver/linear/linear.h:55-71) as
 * IPCSolver::SearchDirection calls it on the 3 nV x 3 nV Hessian in COO form with duplicates (IPC/ipc.cu:233-241):
 *   PD_LS_PCG_JACOBI  PCGJacobiSolver<double> (linear/pcgJacobi.cu:88-172; defaults max_iter 2000, ||r|| < 1e-5)
 *   PD_LS_CG_IC0      CGSolver<double>        (linear/cg.cu:81-247: IC(0)-preconditioned CG; defaults max_iter 100, ||r|| < 1e-6)
 * COO -> CSR (duplicates summed), the preconditioner and the whole CG loop run on the device in one cooperative kernel; no
 * cuSPARSE / cuBLAS / thrust.  max_iter / tolerance <= 0 select the reference's defaults.  The PD path does not use these. */
typedef struct pd_linsolver pd_linsolver;
#define PD_LS_CG_IC0 1
#define PD_LS_PCG_JACOBI 2
pd_linsolver* pd_linsolver_create(int kind, int n, int max_iter, double tolerance, int device);
void pd_linsolver_destroy(pd_linsolver*);
/* device pointers, as the reference passes them; the matrix arrays are not modified (the reference sorts them in place);
 * d_guess may be NULL (x0 = 0).  Synchronises before returning: d_x is complete. */
int pd_linsolver_solve_device(pd_linsolver*, int n, const double* d_b, double* d_x, const double* d_A, int nz, const int* d_row_idx, const int* d_col_idx, const double* d_guess);
int pd_linsolver_solve_host(pd_linsolver*, int n, const double* b, double* x, const double* A, int nz, const int* row_idx, const int* col_idx, const double* guess);
/* of the last solve: CG iterations, final ||r||_2, non-zeros of the assembled CSR */
int pd_linsolver_stats(const pd_linsolver*, int* iterations, double* residual, int* nnz);

/* test hook: the collision pass's continuous-collision test (ccdCollisionTest<float>, intersections.cu:312-355) on n queries of
 * host arrays: type[i] = 1 vertex-face / 2 edge-edge, verts = 4 vertex ids per query, X / XTilde = 3 per vertex;
 * toi[i] in [0, 1] (1 = no hit), normals = 3 per query */
int pd_ccd_batch(int device, int n, const int* type, const uint32_t* verts, int num_verts, const float* X, const float* XTilde, float* toi, float* normals);
/* pinned host memory for the e2e path */
void* pd_alloc_pinned(size_t bytes);
void pd_free_pinned(void*);

#ifdef __cplusplus
}
#endif
#endif /* PD_B200_H */
