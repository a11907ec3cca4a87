/* xpic_b200.h -- C ABI of the B200-native ECSIM / ECSIMCorr / EC-CAPFIM step.
 *
 * xpic (vakurshakov/xpic) has no FFI boundary of its own: a scheme is a C++ subclass compiled
 * into libxpic.so.  This header is the seam a maintainer binds instead (INTEGRATION.md shows
 * the shim): every entry point names the reference member it replaces (file:line relative to
 * the reference tree).  Plain pointers and sizes only; all functions return 0 on success and a
 * non-zero code otherwise (the PetscErrorCode convention, src/interfaces/simulation.h:59,71-72),
 * with a message available from xb_last_error().  One context per rank / GPU; a context is not
 * thread-safe; no exception crosses the ABI.  All arithmetic is fp64, indices int32/int64.
 *
 * There is no CPU fallback: xb_create fails when no CUDA device is usable.
 */
#ifndef XPIC_B200_H
#define XPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xb_ctx xb_ctx;

/* Geometry of the periodic box and of this rank's z-slab.
 * Replaces World::initialize (src/utils/world.cpp:11-48) + the globals of src/constants.h:10-28.
 * The slab decomposition mirrors DMDA with -da_processors_z nranks (utils/configuration.cpp:111-130):
 * rank r owns planes [z0, z0+nzl) of every field, natural [z][y][x][c] order inside. */
typedef struct xb_grid {
  int32_t n[3];      /* global cells Nx, Ny, Nz (geom_nx.., src/constants.h:21-24) */
  double d[3];       /* dx, dy, dz */
  double dt;         /* time step */
  int32_t curl_sign; /* +1 = sources as read; -1 reproduces the golden files of tests/ecsim, tests/ecsimcorr */
  int32_t device;    /* CUDA device ordinal */
  int32_t rank;      /* z-slab index */
  int32_t nranks;    /* number of z-slabs (1 = single GPU) */
  int32_t track_ids; /* 1: carry a 64-bit particle id (parity / diagnostics), 0: do not */
  /* "da_boundary_x/y/z" (src/utils/configuration.cpp:88-108): XB_BOUNDARY_PERIODIC or XB_BOUNDARY_OPEN
   * (DM_BOUNDARY_NONE / DM_BOUNDARY_GHOSTED: nodes outside the box do not exist -- their matrix entries and
   * deposits are dropped, gathers read zero there, particles that leave are removed,
   * src/interfaces/particles.cpp:100-103).  This build: x and y periodic only; open z with scheme XB_ECSIM. */
  int32_t boundary[3];
} xb_grid;

enum { XB_BOUNDARY_PERIODIC = 0, XB_BOUNDARY_OPEN = 1 };

enum { XB_ECSIM = 0, XB_ECSIMCORR = 1, XB_ECCAPFIM = 2 }; /* "Simulation" key, src/interfaces/simulation.cpp:169-178 */

/* Named vectors: Simulation::E,B,B0 (src/interfaces/simulation.h:33-45, get_named_vector :135-143),
 * ecsim Ep,currI (src/impls/ecsim/simulation.h:27-28), ecsimcorr Ec,currJe
 * (src/impls/ecsimcorr/simulation.h:16-17), per-sort Particles::currI / currJe
 * (src/impls/ecsim/particles.h:22, src/impls/ecsimcorr/particles.h:22). */
enum { XB_E = 0, XB_B = 1, XB_B0 = 2, XB_EP = 3, XB_EC = 4, XB_CURRI = 5, XB_CURRJE = 6, XB_CURRI_SORT = 7, XB_CURRJE_SORT = 8,
       /* eccapfim: Simulation::J, Particles::J and E_hk of the last residual evaluation
        * (src/interfaces/simulation.h:36, src/interfaces/particles.h:41, src/impls/eccapfim/simulation.h:74) */
       XB_J = 9, XB_J_SORT = 10, XB_EHK = 11 };

/* Stages of timestep_implementation (src/impls/ecsim/simulation.cpp:145-155,
 * src/impls/ecsimcorr/simulation.cpp:21-32); same names as the PETSc log stages :495-508. */
enum {
  XB_STAGE_CLEAR_SOURCES = 0,
  XB_STAGE_FIRST_PUSH = 1,   /* push + re-binning + moments (current, mass matrices) */
  XB_STAGE_ADVANCE_FIELDS = 2, /* "Advance field" / "Predict field" */
  XB_STAGE_SECOND_PUSH = 3,
  XB_STAGE_CORRECT_FIELDS = 4, /* ecsimcorr only */
  XB_STAGE_FINAL_UPDATE = 5,
  XB_STAGE_COUNT = 6
};

enum { XB_SOLVER_PREDICT = 0, XB_SOLVER_CORRECT = 1 }; /* ecsim `ksp` / ecsimcorr `predict`,`correct` */

/* xb_scalar selectors: ecsimcorr::Particles members read by ecsimcorr::Energy
 * (src/impls/ecsimcorr/simulation.cpp:169-197) and Energy::calculate_kinetic (diagnostics/energy.cpp:61-107) */
enum { XB_KINETIC = 0, XB_PRED_W = 1, XB_CORR_W = 2, XB_PRED_DK = 3, XB_CORR_DK = 4, XB_LAMBDA_DK = 5, XB_ENERGY_MEMBER = 6, XB_J_DIFF_NORM = 7 };

/* Operator selectors for xb_spmv*: bit 0 = particle mass matrix L (matL), bit 1 = constant
 * M = 2I + dt^2/2 curl curl (matM, src/impls/ecsim/simulation.cpp:544-552); 3 = A = L + M. */
enum { XB_OP_L = 1, XB_OP_M = 2, XB_OP_A = 3 };

const char* xb_last_error(void);
int xb_version(void);

/* Number of fixed-offset coefficients per cell of the stencil-layout operator (369): the non-zeros of matL per
 * cell that Simulation::fill_matrix_indices generates (src/impls/ecsim/simulation.cpp:370-469), without indices. */
int xb_operator_ncoef(void);
/* Describe coefficient k: row component c1, column component c2 and column offset (dx,dy,dz). */
int xb_operator_coef_info(int k, int* c1, int* c2, int* dx, int* dy, int* dz);

/* ncclUniqueId for a multi-rank run (128 bytes), produced on rank 0 and passed to every
 * rank's xb_create (replaces MPI_Init / PETSC_COMM_WORLD, src/main.cpp:12). */
int xb_comm_unique_id(void* out128);

/* World::initialize + Simulation::initialize_implementation (src/impls/ecsim/simulation.cpp:122-143):
 * device vectors, Yee curls, solver defaults (rtol = atol = 1e-7, maxit 100, src/impls/ecsim/simulation.h:15-18). */
int xb_create(const xb_grid* grid, const void* comm_unique_id, xb_ctx** out);
/* Simulation::finalize (src/impls/ecsim/simulation.cpp:569-590) */
int xb_destroy(xb_ctx* ctx);

/* interfaces::Simulation::init_particles, one call per "Particles" JSON entry
 * (src/interfaces/simulation.tpp:7-79).  capacity = max particles this rank may ever hold. */
int xb_species_add(xb_ctx* ctx, double q, double m, double n, int32_t Np, int64_t capacity, int32_t* sid);

/* interfaces::Particles::add_particle (src/interfaces/particles.cpp:47-57): host AoS Points
 * {r[3], p[3]} (48 B each); points outside this rank's slab are skipped; *added = number kept.
 * ids may be NULL (then ids continue from the species' running counter). */
int xb_particles_append(xb_ctx* ctx, int32_t sid, const double* aos6, const uint64_t* ids, int64_t count, int64_t* added);
/* SetParticles{CoordinateInBox, MaxwellianMomentum} (src/commands/set_particles.cpp:19-43,
 * src/utils/particles_load.cpp:11-18,52-76) for the large synthetic configurations: `total` particles
 * of the GLOBAL box drawn from a counter-based generator (not the serial mt19937 stream; parity-size
 * runs upload the reference's own stream with xb_particles_append); this rank keeps those of its slab. */
int xb_particles_maxwellian(xb_ctx* ctx, int32_t sid, int64_t total, const double T[3], uint64_t seed, int32_t tov, int64_t* added);
int xb_particles_count(xb_ctx* ctx, int32_t sid, int64_t* count);
/* Host mirror of interfaces::Particles::storage (src/interfaces/particles.h:32), cell-major order. */
int xb_particles_download(xb_ctx* ctx, int32_t sid, double* aos6, uint64_t* ids, int64_t capacity, int64_t* count);

/* Owned z-slab of a named vector, natural [z][y][x][c] order, 3*Nx*Ny*nzl doubles
 * (VecGetArray / VecView of a DMDA global vector, e.g. diagnostics/field_view.cpp:98-118). */
int xb_field_upload(xb_ctx* ctx, int32_t which, int32_t sid, const double* host);
int xb_field_download(xb_ctx* ctx, int32_t which, int32_t sid, double* host);

/* KSPSetTolerances / -ksp_rtol etc. (src/impls/ecsim/simulation.cpp:558-567, ecsimcorr :114-136).
 * precond: 0 none, k>0 Chebyshev polynomial of degree k in the constant operator M. */
int xb_solver_set(xb_ctx* ctx, int32_t which, double rtol, double atol, int32_t maxit, int32_t restart, int32_t precond);
/* KSPGetIterationNumber / KSPGetResidualNorm / KSPGetConvergedReason (reason > 0 converged). */
int xb_solver_info(xb_ctx* ctx, int32_t which, int32_t* iterations, double* rnorm, int32_t* reason);

/* Simulation::timestep_implementation (ecsim/simulation.cpp:145, ecsimcorr/simulation.cpp:21).
 * Returns non-zero when a solve does not converge (KSPSetErrorIfNotConverged, :562). */
int xb_step(xb_ctx* ctx, int32_t scheme);
/* One stage of the step, for per-stage timing / parity: the stages timestep_implementation logs
 * (src/impls/ecsim/simulation.cpp:145-155, PetscLogStage names :495-508). */
int xb_stage(xb_ctx* ctx, int32_t scheme, int32_t stage);
/* The same step behind the reference-facing boundary with HOST buffers: uploads E, B, B0 (what
 * StepPresets commands may have edited), steps, downloads E, B and the kinetic energy of every
 * sort (what Energy::diagnose reads each step, src/interfaces/simulation.cpp:91-92).
 * kinetic may be NULL.  Particles stay resident (xb_particles_download refreshes the mirror). */
int xb_step_host(xb_ctx* ctx, int32_t scheme, double* E, double* B, const double* B0, double* kinetic);

/* k consecutive steps bracketed by CUDA events on the launching stream; *ms = device-timeline
 * milliseconds for all k steps (Simulation::calculate's loop, src/interfaces/simulation.cpp:75-96). */
int xb_run_steps(xb_ctx* ctx, int32_t scheme, int32_t k, double* ms);
int xb_run_steps_host(xb_ctx* ctx, int32_t scheme, int32_t k, double* E, double* B, const double* B0, double* kinetic, double* ms);

int xb_scalar(xb_ctx* ctx, int32_t sid, int32_t which, double* out);
/* Energy::calculate_kinetic's sums over all ranks (src/diagnostics/energy.cpp:61-107):
 * out = { sum vx, sum vy, sum vz, sum v^2, number of particles }. */
int xb_particle_moments(xb_ctx* ctx, int32_t sid, double out[5]);

/* Seconds / launches accumulated per stage since the last reset (SyncClock, utils/sync_clock.cpp:75-93). */
int xb_timing(xb_ctx* ctx, int32_t stage, double* seconds, int64_t* calls);
int xb_timing_reset(xb_ctx* ctx);
/* Number of this library's kernel launches since creation (measurement hook; PETSc's -log_view counts events instead). */
int xb_launch_count(xb_ctx* ctx, int64_t* launches);
/* Per-launch CUDA-event timing of the operator kernel (the MatMult event of PETSc's -log_view):
 * enable != 0 starts collecting; the query returns the launches seen and their summed ms. */
int xb_spmv_profile(xb_ctx* ctx, int32_t enable);
int xb_spmv_profile_read(xb_ctx* ctx, int64_t* launches, double* total_ms);
/* The same for the other kernel families INSIDE the step (PETSc log events at kernel-family granularity):
 * re-binning (Particles::update_cells, src/interfaces/particles.cpp:79-116, with the move of first_push fused),
 * moments (fill_ecsim_current, src/impls/ecsim/simulation.cpp:336-368), second push (ecsim/particles.cpp:175-192),
 * operator SpMV (MatMult), Chebyshev preconditioner steps (PCApply). */
enum { XB_FAMILY_SORT = 0, XB_FAMILY_MOMENTS = 1, XB_FAMILY_PUSH2 = 2, XB_FAMILY_SPMV = 3, XB_FAMILY_PRECOND = 4,
       /* parts of XB_FAMILY_MOMENTS: cell blocks of the owned planes, of the two ghost planes (multi-rank: includes
        * the wait for the neighbours' boundary particles), gather of the rows */
       XB_FAMILY_MOMENTS_CELLS = 5, XB_FAMILY_MOMENTS_GHOST = 6, XB_FAMILY_MOMENTS_ROWS = 7,
       /* parts of XB_FAMILY_SORT: move + key pass, migration (count table, payloads, keys of the arrivals; multi-rank), scan + scatter */
       XB_FAMILY_SORT_KEYS = 8, XB_FAMILY_SORT_MIGRATE = 9, XB_FAMILY_SORT_SCATTER = 10, XB_FAMILY_COUNT = 11 };
int xb_family_profile(xb_ctx* ctx, int32_t enable);
int xb_family_profile_read(xb_ctx* ctx, int32_t family, int64_t* launches, double* total_ms);
/* Energy::calculate_energy (src/diagnostics/energy.cpp:43-59): 0.5 * |v|^2 of a named vector, summed over all ranks. */
int xb_field_energy(xb_ctx* ctx, int32_t which, int32_t sid, double* out);
/* The sums Energy::calculate_energy forms with VecNorm and VecStrideSumAll (energy.cpp:43-59), over all ranks:
 * out = { sum of component x, y, z, sum of squares }. */
int xb_field_sums(xb_ctx* ctx, int32_t which, int32_t sid, double out[4]);

/* --- hooks for the operator sweep (BASELINE config 4) and per-kernel parity tests ----------- */
/* y = Op x with host vectors (owned slab, natural order).  MatMult, e.g. ecsimcorr/simulation.cpp:78 */
int xb_spmv(xb_ctx* ctx, int32_t op, const double* x, double* y);
/* Times `reps` device-resident SpMVs on pseudo-random x; returns average ms per SpMV (the MatMult event of
 * PETSc's -log_view, BASELINE config 4). */
int xb_spmv_bench(xb_ctx* ctx, int32_t op, int32_t reps, double* ms_per_spmv);
/* Stencil-layout coefficients of L: coef[k*ncells + cell], k < xb_operator_ncoef(), owned cells -- the values
 * MatSetValuesCOO sums into matL (src/impls/ecsim/simulation.cpp:359,366). */
int xb_operator_download(xb_ctx* ctx, double* coef);
int xb_operator_upload(xb_ctx* ctx, const double* coef);
/* Kernel variant switches for cross-checks: what = 0 selects the first pass of the moment
 * deposition (value 0: fused warp-specialised fp64 tensor-core kernel whose accumulator tiles are summed by the row
 * gather, default; the others stage the reference's 9 x 12 x 12 cell blocks: 4 the same kernel with the fold in shared
 * memory, 3: fused kernel without role split; 2: round-1 pipeline with field records in HBM; 1: scalar FMA);
 * what = 1 the Esirkepov deposit
 * of ecsimcorr (0: atomic-free DMMA cell blocks, default; 1: per-particle global fp64 reductions);
 * what = 2: canonical particle order inside every bin after a sort when ids are not tracked
 * (1: runs are bit-reproducible, costs one more pass over the particles; 0, default: keep the order
 * in which the scatter's integer atomics resolved -- results then differ run to run at round-off level;
 * with track_ids = 1 the order is always canonical, by id);
 * what = 3 the particle pass of eccapfim's residual evaluation (0: CTA-wide task machine, default;
 * 1: one thread per particle);
 * what = 4: eccapfim's per-particle Picard iteration of residual evaluation k + 1 starts from the velocity
 * evaluation k of the same step converged to (1, default) or from the start-of-step velocity every time as
 * the reference does (0, src/impls/eccapfim/particles.cpp:77-78); the converged particle state is the same
 * to the Picard tolerance, the number of field gathers per evaluation drops from ~3.7 to ~2;
 * what = 5: nanoseconds a warp of the warp-specialised moment kernel sleeps between two polls of its mbarrier
 * (measurement knob; 0 = spin). */
int xb_set_option(xb_ctx* ctx, int32_t what, int32_t value);
/* --- eccapfim (BASELINE config 5): xb_step / xb_stage with scheme XB_ECCAPFIM run
 * eccapfim::Simulation::timestep_implementation (src/impls/eccapfim/simulation.cpp:36-44); stages
 * XB_STAGE_CLEAR_SOURCES = init_iteration (:46-70), XB_STAGE_ADVANCE_FIELDS = calc_iteration (SNESSolve,
 * :72-104; non-zero return when it does not converge, :102), XB_STAGE_FINAL_UPDATE = after_iteration (:106-129).
 * SNESSetTolerances (simulation.cpp:384, defaults simulation.h:14-19) + the per-particle Crank-Nicolson
 * tolerance and iteration cap (particles.cpp:99-101: 0.5 * atol, 30).  depth = Anderson history,
 * cheb_degree = degree of the Chebyshev residual preconditioner (0 = plain fixed-point residual). */
int xb_nonlinear_set(xb_ctx* ctx, double atol, double rtol, double stol, int32_t maxit, int32_t depth, int32_t cheb_degree, double particle_tol,
                     int32_t particle_maxit);
/* SNESGetIterationNumber / SNESGetNumberFunctionEvals / SNESGetConvergedReason and the averages the
 * ConvergenceHistory diagnostic prints (src/impls/eccapfim/convergence_history.cpp:11-44). */
int xb_nonlinear_info(xb_ctx* ctx, int32_t* iterations, int32_t* fevals, int32_t* reason, double* fnorm, double* avg_cn, double* avg_cells);
/* SNESGetConvergenceHistory: |F| after every iteration of the last solve (length entries, <= capacity copied). */
int xb_nonlinear_history(xb_ctx* ctx, double* out, int32_t capacity, int32_t* length);
/* CUDA-event timing of the particle pass of every residual evaluation (the "form_current" event,
 * simulation.cpp:180-192): returns the totals so far, then enable >= 0 resets and switches collecting. */
int xb_nonlinear_profile(xb_ctx* ctx, int32_t enable, int64_t* evaluations, double* total_ms);
/* F(x) of form_iteration (simulation.cpp:132-155) for host x at the present E^n, B^n and particles;
 * XB_J / XB_J_SORT then hold the currents of that evaluation. */
int xb_eccapfim_function(xb_ctx* ctx, const double* x, double* f);

/* --- per-step diagnostics on the device (the reference walks Particles::storage on the host) -------
 * ParticlesChargeDensity::collect (src/diagnostics/charge_conservation.cpp:67-97): the charge density of
 * sort sid at the nodes, 2nd-order form factor; rho (may be NULL) receives the owned slab, Nx*Ny*nzl doubles.
 * The collection is remembered as the sort's latest density for xb_charge_conservation. */
int xb_charge_density(xb_ctx* ctx, int32_t sid, double* rho);
/* ChargeConservation::add_columns (charge_conservation.cpp:125-171): for every sort the 1- and 2-norm of
 * (rho_now - rho_previous) / dt + div J_sort (rho_now is collected here, rho_previous by the previous
 * call or xb_charge_density = ChargeConservation::initialize), then the same for the sum of the sorts with
 * the total current.  which_current: 0 = currJe (ecsimcorr, ecsimcorr/simulation.cpp:103-110),
 * 1 = J (eccapfim).  norms receives 2 * (number of sorts + 1) doubles, summed over all ranks. */
int xb_charge_conservation(xb_ctx* ctx, int32_t which_current, double* norms);

/* DistributionMoment::collect (src/diagnostics/distribution_moment.cpp:157-210) for sort sid: cell-centred
 * moments with the 1st-order form factor on the owned slab, natural [z][y][x][component] order,
 * Nx*Ny*nzl*components doubles.  moment: get_density (1 component), get_current (3), get_momentum_flux (6: xx xy xz yy yz zz),
 * get_momentum_flux_cyl (6: rr ra rz aa az zz about the box axis), the two diagonal forms (3) -- :212-313.
 * The region form takes `start` / `size` in cells ("region" of the diagnostic, field_view_builder.cpp:52-97): only
 * particles whose cell lies inside contribute (:180-181) and what they deposit outside is dropped unless the region
 * spans the whole axis; the output is still the whole owned slab (zeros outside the region). */
enum { XB_MOMENT_DENSITY = 0, XB_MOMENT_CURRENT = 1, XB_MOMENT_MOMENTUM_FLUX = 2, XB_MOMENT_MOMENTUM_FLUX_CYL = 3, XB_MOMENT_MOMENTUM_FLUX_DIAG = 4,
       XB_MOMENT_MOMENTUM_FLUX_DIAG_CYL = 5 };
int xb_distribution_moment(xb_ctx* ctx, int32_t sid, int32_t moment, double* out);
int xb_distribution_moment_region(xb_ctx* ctx, int32_t sid, int32_t moment, const int32_t start[3], const int32_t size[3], double* out);

/* VelocityDistribution::collect (src/diagnostics/velocity_distribution.cpp:116-166) for sort sid: the number density
 * n / Np of the particles whose cell centre lies inside the geometry (same six numbers as the commands below), binned
 * over two projections of the velocity -- get_vx_vy, get_vz_vxy, get_vr_vphi (:169-194) -- with ROUND_STEP(v, dv).
 * As in set_regions (:53-63) both axes start at ROUND_STEP(vmin[0], dv[0]) and hold ROUND_STEP(vmax[0] - vmin[0], dv[0])
 * bins; xb_velocity_distribution_size returns those two numbers.  out receives size * size doubles,
 * [second projection][first projection], summed over all ranks (the reference's VecScatter ADD, :163-164). */
enum { XB_PROJECTOR_VX_VY = 0, XB_PROJECTOR_VZ_VXY = 1, XB_PROJECTOR_VR_VPHI = 2 };
int xb_velocity_distribution_size(const double dv[2], const double vmin[2], const double vmax[2], int32_t* start, int32_t* size);
int xb_velocity_distribution(xb_ctx* ctx, int32_t sid, int32_t projector, int32_t geometry, const double p[6], const double dv[2],
                             const double vmin[2], const double vmax[2], double* out);

/* MomentumConservation::calculate (src/diagnostics/momentum_conservation.cpp:71-126) for sort sid with the
 * present E: out = { Px, Py, Pz, QEx, QEy, QEz }, P = m / Np * sum v, QE = q / Np * sum E(x_p) with the global
 * 2nd-order form factor, summed over all ranks.  The caller keeps P of the previous call for the
 * (P1 - P0) / dt - QE residual of add_columns (:29-68). */
int xb_momentum(xb_ctx* ctx, int32_t sid, double out[6]);

/* --- StepPresets commands on the device (src/commands) ---------------------------------------------------------
 * Geometry of a command: XB_GEOMETRY_BOX with p = { min[3], max[3] } (BoxGeometry, WithinBox, src/utils/geometries.cpp:3-9)
 * or XB_GEOMETRY_CYLINDER with p = { center[3], radius, height, 0 } (WithinCylinder, :12-19). */
enum { XB_GEOMETRY_BOX = 0, XB_GEOMETRY_CYLINDER = 1 };
/* FieldsDamping::execute (src/commands/fields_damping.cpp:16-112): E and B - B0 at the nodes whose cell centre lies
 * outside the geometry are multiplied by DampForBox / DampForCylinder's factor (:66-112); *damped_energy receives
 * the energy taken out, 0.5 f^2 (1 - damping^2) summed over all ranks (may be NULL). */
int xb_fields_damping(xb_ctx* ctx, int32_t geometry, const double p[6], double coefficient, double* damped_energy);
/* RemoveParticles::execute (src/commands/remove_particles.cpp:11-45): every particle of sort sid whose CELL corner
 * (x dx, y dy, z dz) lies outside the geometry is removed; out = { removed particles, removed kinetic energy
 * 0.5 m v^2 n / Np } over all ranks (may be NULL).  Collective in a multi-rank run. */
int xb_particles_remove(xb_ctx* ctx, int32_t sid, int32_t geometry, const double p[6], double out[2]);

/* Moments only at the present particle positions: fill_ecsim_current (ecsim/simulation.cpp:336-368). */
int xb_deposit(xb_ctx* ctx);
/* Solve (L? + M) x = b for host vectors with the given solver slot (KSPSolve). */
int xb_solve(xb_ctx* ctx, int32_t which, int32_t op, const double* b, double* x);
/* out = curl(f): positive != 0 -> Rotor::create_positive, else create_negative (utils/operators.cpp:175-213). */
int xb_curl(xb_ctx* ctx, int32_t positive, const double* f, double* out);
/* Time one kernel family in isolation on the resident state (the per-stage SyncClock of the reference,
 * src/utils/sync_clock.cpp:75-93, at kernel-family granularity): what = 0 first_push+sort, 1 deposit,
 * 2 second_push, 3 solve(predict).  Returns average ms. */
int xb_kernel_bench(xb_ctx* ctx, int32_t what, int32_t reps, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* XPIC_B200_H */
