/*
 * pd_oracle.h -- CPU restatement of the reference's float projective-dynamics step.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product (soft-body-simulation-cuda_b200/)
 * may include, link or call this. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Parity pin: the reference's own tests hold NO golden vectors for this path
 * (SURVEY.md section 4), so the pin is the reference's CUDA kernels compiled verbatim
 * (oracle/_ref, built by oracle/Makefile from /root/reference) and run on a B200;
 * outputs of those runs are committed under tests/golden/ (see tests/golden/README.md).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  All arithmetic is IEEE float, one rounding per operation
 * (compile with -ffp-contract=off), sequential tet order for the scatter.
 */
#ifndef PD_ORACLE_H
#define PD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fixed bodies in the form the collision kernels consume them
 * (src/simulation/fixedBodyData.cu:67-134) */
typedef struct {
    int   n_planes;      /* each: point p0[3], unit normal up[3]           */
    const float *plane_p0, *plane_up;
    int   n_spheres;     /* each: centre[3], radius                        */
    const float *sphere_c, *sphere_r;
    int   n_cyls;        /* each: centre[3], unit axis[3], radius          */
    const float *cyl_c, *cyl_axis, *cyl_r;
} o_fixed_bodies;

typedef struct {
    float dt, gravity, muN, muT, rho, tol;
    int   num_iterations;
    int   global_solver;     /* 0 Jacobi (Chebyshev), 1 direct (Cholesky), 2 PCG-Jacobi */
    int   pcg_max_iter;      /* pcgJacobi.cu: max_iter                       */
    float pcg_tol;           /* pcgJacobi.cu: tolerance on ||r||_2           */
    int   threads;           /* OpenMP threads for the local step (1 = sequential) */
} o_params;

typedef struct o_scene o_scene;

/* external/svd3_cuda/svd3_cuda.h:34-1041 ; A,U,V row-major a[r*3+c]; S = (s11,s22,s33) */
void o_svd3(const float A[9], float U[9], float S[3], float V[9]);

/* corotational projection used by PdUtil::computeLocal (pdUtil.cu:112-122): R = U V^T */
void o_rotation(const float F[9], float R[9]);

/* solverUtil.cuh:98-116 : DmInv (row-major, 9 per tet) and V0 */
void o_rest_shape(const float *X, const uint32_t *Tet, int nT, float *DmInv, float *V0);

/* utilities.cpp:141-150 (fixed bodies: T*Rx*Ry*Rz*S) and dataLoader.cu:214-220
 * (soft bodies: T*S*Rx*Ry*Rz); column-major 4x4 like glm; rot in degrees          */
void o_model_matrix(const float pos[3], const float rot[3], const float scale[3],
                    int soft_body_order, float M[16]);
/* utilities.cu:56-65 */
void o_transform_vertices(float *X, int nV, const float M[16]);
/* rigid/plane.cpp:9 + rigid/rigid.cpp:5 */
void o_plane_up(const float M[16], float up[3]);
/* fixedBodyData.cu:116 */
void o_cylinder_axis(const float M[16], float axis[3]);

/* dataLoader.cu:131-173 / :38-66 ; return count or <0 on error. X/Tet malloc'ed by callee */
int o_load_node(const char *path, int centralize, float **X);
int o_load_ele(const char *path, int start_index, uint32_t **Tet);
void o_free(void *p);

/* scene = merged bodies (dataLoader.cu:291-378) + solver state (pdSolver.cu) */
o_scene *o_scene_create(int nV, int nT, const float *X, const uint32_t *Tet,
                        const float *mass, const float *mu, const float *DBC,
                        const o_fixed_bodies *fb);
void o_scene_destroy(o_scene *);
/* PdSolver::Update (pdSolver.cu:210-232), n times. returns 0 on success */
int  o_scene_step(o_scene *, const o_params *, int n_steps);
/* SimulationCUDAContext::Reset (simulationContext.cu:233-243) */
void o_scene_reset(o_scene *);
void o_scene_get(const o_scene *, float *X, float *V, float *XTilde);
void o_scene_set(o_scene *, const float *X, const float *V, const float *XTilde);
/* mouse-drag soft constraints: SolverData::moreDBC[nV] / OffsetX[3nV] (def.h:31-32) and MouseSelection::target
 * (def.h:14-18), consumed at pdUtil.cu:56-69,80-87,159-164,187-188,201-206.  moreDBC NULL clears the drag. */
void o_scene_set_drag(o_scene *, const float *moreDBC, const float *OffsetX, const float target[3]);
/* Control_Kernel (simulationContext.cu:202-218) on the current X, then target */
void o_scene_drag_select(o_scene *, int select_v, float control_mag, const float target[3]);
void o_scene_get_drag(const o_scene *, float *moreDBC, float *OffsetX, float *DBCX);
/* live stiffness edit (UpdateSoftBodyAttr, simulationContext.cu:165-176): mu[nT]; setup products follow only after reset */
void o_scene_set_mu(o_scene *, const float *mu);
/* setup products, for unit checks: matrix_diag[nV], massDt_2s[nV], DmInv[9nT], V0[nT] */
void o_scene_get_setup(o_scene *, const o_params *, float *matrix_diag, float *massDt_2s,
                       float *DmInv, float *V0);
/* scalar system matrix A^ = diag(c) + sum_t P^T K_t P (pdUtil.cu:9-54), CSR, sorted cols.
 * Call with rowptr==NULL to get nnz. */
int  o_scene_system_matrix(o_scene *, const o_params *, int *rowptr, int *col, float *val);
/* statistics of the last step: total PCG iterations, PD iterations executed */
void o_scene_stats(const o_scene *, int *pd_iters, int *inner_iters);

/* mesh-mesh collision of PdSolver::Update (pdSolver.cu:218-225: DetectCollision with ignoreSelfCollision = true, then
 * CCDKernel): enable / disable (SolverParams::handleCollision) and, when Tri is not NULL, the surface triangles
 * (SolverData::Tri, 3 per triangle) with their father bodies (dev_TriFathers; NULL = all 0).  All triangle pairs are tried
 * instead of the reference's LBVH (same set of overlapping pairs). */
void o_scene_set_collision(o_scene *, int enable, int nTris, const uint32_t *Tri, const uint32_t *TriFathers);
/* SolverData::dev_tIs / dev_Normals after the last pass, and the number of (ordered) overlapping triangle pairs */
void o_scene_get_collision(const o_scene *, float *tI, float *normals, long long *pairs);
/* ccdCollisionTest<float> (intersections.cu:312-355) on one query: ee = edge-edge, q = its four vertex ids */
float o_ccd_test(int ee, const uint32_t q[4], const float *X, const float *XTilde, float normal[3]);

/* ---- double-precision twin of the Jacobi step (separates algorithmic from rounding
 * differences) ---- */
int  o_scene_step_f64(o_scene *, const o_params *, int n_steps);

#ifdef __cplusplus
}
#endif
#endif
