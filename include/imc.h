/* imc.h — C ABI of the B200 Implicit Monte Carlo transport-step engine.
 *
 * Drop-in boundary for the transport step of simonbutson/MixedPrecisionIMC.jl.  The reference
 * has no FFI; its boundary is the set of Julia call sites in the time-step loop
 * (src/MixedPrecisionIMC.jl:138-146 and :167-171):
 *
 *     Update.update(inputs, mesh, simvars)                       -> imc_update
 *     Sourcing.sourcing(mesh, simvars, particles)                -> imc_source
 *     Transport.MC / MC_RW / MC2D(mesh, simvars, [rw], particles)-> imc_transport
 *     Clean.clean(particles)                                     -> imc_clean
 *     Tally.tally(inputs, mesh, simvars, particles)              -> imc_tally
 *     EnergyCheck.energychecker(inputs, mesh, simvars, particles)-> imc_energycheck
 *     Transport.randomwalk_table(aVals, prVals, ptVals, simvars) -> imc_rw_table
 *
 * A Julia shim with those exact module/function names `ccall`s these entry points
 * (INTEGRATION.md); MixedPrecisionIMC.main, the deck parser and the mesh generator stay Julia.
 *
 * Conventions
 *  - Every function returns 0 on success or a negative imc_status; the message is available
 *    from imc_last_error().  Nothing throws across the ABI.
 *  - All host arrays cross the boundary as Float64 (double), whatever the deck PRECISION is:
 *    values produced by the host in Float16/Float32 are exactly representable and are rounded
 *    back (exactly) inside.  2-D fields are column-major [xindex, yindex] (x fastest), as Julia
 *    stores them; multi-scale fields are [cell, scale] (scale plane slowest).
 *  - Host pointers are borrowed for the duration of the call only.  The engine owns the
 *    particle population and all device memory.
 *  - One handle = one GPU (or one oracle instance).  Calls on a handle must come from one
 *    thread, in time-step order; several handles may coexist.
 *
 * The same ABI is exported by two libraries: libimc_b200.so (CUDA, sm_100a — the product) and
 * oracle/_build/libimc_oracle.so (CPU restatement of the reference — test infrastructure only).
 */
#ifndef IMC_B200_H
#define IMC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMC_ABI_VERSION 1
#define IMC_MAX_SCALES 16

typedef struct imc_engine* imc_handle;

typedef enum {
  IMC_OK = 0,
  IMC_ERR_ARG = -1,        /* bad argument / inconsistent configuration */
  IMC_ERR_STATE = -2,      /* call out of order (e.g. transport before set_mesh) */
  IMC_ERR_CUDA = -3,       /* CUDA runtime error */
  IMC_ERR_NOMEM = -4,      /* host or device allocation failed */
  IMC_ERR_TAPE = -5,       /* replay tape exhausted */
  IMC_ERR_NUMERIC = -6,    /* NaN/Inf met where the reference would abort (InexactError etc.) */
  IMC_ERR_UNSUPPORTED = -7
} imc_status;

/* PRECISION (imc_input.jl:113-129) */
typedef enum { IMC_F16 = 0, IMC_F32 = 1, IMC_F64 = 2 } imc_precision;
/* LEFTBC/RIGHTBC/TOPBC/BOTTOMBC (imc_transport.jl:133-161, :625-696) */
typedef enum { IMC_REFLECT = 0, IMC_VACUUM = 1 } imc_bc;
/* random numbers: counter-based Philox, or replay of pre-drawn numbers */
typedef enum { IMC_RNG_PHILOX = 0, IMC_RNG_TAPE = 1 } imc_rng_mode;
/* energy-deposition / census accumulation
 *   ATOMIC : floating-point atomics (shared-memory privatised, one flush per block), Float64 totals
 *   FIXED  : 64-bit fixed-point atomics — order-free, bit-identical for any GPU count / schedule
 *   EXACT  : reference order — per-deposit records sorted by (cell, particle, segment) and summed
 *            sequentially (PAIRWISE = FALSE, imc_transport.jl:120) or with Julia's pairwise
 *            sum (PAIRWISE = TRUE, imc_transport.jl:198-205); O(segments) memory
 *   AUTO   : PAIRWISE = TRUE or PRECISION = FLOAT16 (summation order is part of the reference's result)
 *            -> EXACT while the records fit exact_record_budget, else FIXED (PAIRWISE) / ATOMIC;
 *            otherwise ATOMIC */
typedef enum { IMC_TALLY_AUTO = 0, IMC_TALLY_ATOMIC = 1, IMC_TALLY_FIXED = 2, IMC_TALLY_EXACT = 3 } imc_tally_mode;
/* tracking kernel variant */
typedef enum { IMC_TRACK_AUTO = 0, IMC_TRACK_HISTORY = 1, IMC_TRACK_REFILL = 2, IMC_TRACK_EVENT = 3 } imc_track_mode;

/* index of each boundary in imc_config.bc, in the reference's BC tuple order
 * (MixedPrecisionIMC.jl:123, :162) */
enum { IMC_BC_LEFT = 0, IMC_BC_RIGHT = 1, IMC_BC_TOP = 2, IMC_BC_BOTTOM = 3 };

typedef struct {
  int32_t struct_size;       /* sizeof(imc_config), for ABI checking */
  int32_t precision;         /* imc_precision */
  int32_t geometry;          /* 1 or 2  (GEOMETRY = 1D / 2D) */
  int32_t nx, ny;            /* cells (ny = 1 in 1-D) */
  int32_t bc[4];             /* imc_bc: left, right, top, bottom */
  int32_t linearized;        /* LINEARIZED = TRUE  (imc_update.jl:23, imc_tally.jl:71) */
  int32_t pairwise;          /* PAIRWISE = TRUE */
  int32_t randomwalk;        /* RANDOMWALK = TRUE (1-D only; selects MC_RW) */
  int32_t marshak_quirk;     /* uppercase(NAME) == "MARSHAK WAVE" (imc_update.jl:32) */
  int32_t n_scales;          /* length(ENERGYSCALES), 1..IMC_MAX_SCALES */
  double energyscales[IMC_MAX_SCALES]; /* sorted descending (imc_mesh.jl:130) */
  double distancescale, phys_c, phys_a, alpha; /* already parsed in deck precision */
  int64_t seed;              /* SEED */
  int64_t n_max;             /* NMAX (after the host's parse through PRECISION) */
  int32_t device;            /* CUDA device ordinal (ignored by the oracle) */
  int32_t rank, world;       /* particle sharding: this engine emits new-particle ordinals j with j % world == rank */
  int32_t rng_mode;          /* imc_rng_mode */
  int32_t tally_mode;        /* imc_tally_mode */
  int32_t track_mode;        /* imc_track_mode */
  int64_t exact_record_budget; /* max deposit records for IMC_TALLY_EXACT under AUTO (0 = default 2^28) */
} imc_config;

typedef struct {
  double totalenergy;        /* mesh.totalenergy (imc_sourcing.jl:121) */
  double emitted_sum;        /* sum(emittedenergy ./ escale)  (print at imc_sourcing.jl:75) */
  int64_t n_source;          /* n_source after the NMAX cap (imc_sourcing.jl:132-136) */
  int64_t n_new_global;      /* particles created this step over all ranks */
  int64_t n_new_local;       /* particles created by this engine */
  int64_t n_particles;       /* length(particles) on this engine after sourcing (:369) */
} imc_source_stats;

typedef struct {
  double lostenergy;         /* mesh.lostenergy after this call (accumulates until energycheck) */
  uint64_t segments;         /* loop iterations this call == increment of simvars.iterations (imc_transport.jl:73) */
  uint64_t segments_total;   /* cumulative, what the reference prints (Q13) */
  int64_t histories;         /* particles tracked */
  int64_t n_census, n_absorbed, n_escaped; /* outcomes */
  int64_t n_rw;              /* random-walk steps taken (MC_RW) */
  int64_t n_errors;          /* NaN/Inf distances or energies met (reference: print + sleep) */
  int32_t variant;           /* imc_track_mode actually used */
  int32_t tally_mode;        /* imc_tally_mode actually used */
  float kernel_ms;           /* device time of the tracking kernel(s) (CUDA events; 0 for the oracle) */
} imc_transport_stats;

typedef struct {
  double totalenergydep;     /* mesh.totalenergydep (imc_tally.jl:44-56) */
  double energy_increase;    /* sum(nrg_inc)  (print at :67) */
  double max_temp;           /* maximum(mesh.temp) (:78) */
  double total_energy_density; /* sum(matenergydens + radenergydens) (:136) */
} imc_tally_stats;

typedef struct {
  double radenergy;          /* sum(radenergydens .* dx [.* dy'])  (imc_energycheck.jl:24-29) */
  double radenergy_change;   /* radenergy - radenergyold */
  double lostenergy;         /* before the reset at :37 */
  double energy_error;       /* (totalenergy - totalenergydep - change - lost) / totalenergy (:34) */
} imc_energy_stats;

/* fields readable with imc_get_field (all returned as double) */
typedef enum {
  IMC_FIELD_TEMP = 0, IMC_FIELD_FLECK, IMC_FIELD_BETA, IMC_FIELD_BEE, IMC_FIELD_SIGMA_A, IMC_FIELD_SIGMA_S,
  IMC_FIELD_ENERGYDEP,      /* [Nc x Ns] */
  IMC_FIELD_EMITTEDENERGY,  /* [Nc x Ns] */
  IMC_FIELD_MATENERGYDENS, IMC_FIELD_RADENERGYDENS, IMC_FIELD_NRG_INC, IMC_FIELD_COUNT_
} imc_field;

int imc_abi_version(void);
/* "cuda-sm_100a" for the product library, "oracle-cpu" for the oracle */
const char* imc_backend(void);

int imc_create(const imc_config* cfg, imc_handle* out);
void imc_destroy(imc_handle h);
const char* imc_last_error(imc_handle h); /* h may be NULL: error of the last failed imc_create */

/* Mesh and material state produced by Mesh.mesh_generation (imc_mesh.jl:42-173).
 *   dx[nx], dy[ny] (dy may be NULL in 1-D)
 *   sigma_a_const/pow, sigma_s_const/pow [Nc]: columns 2 and 3 of mesh.sigma_a / sigma_s (already / distancescale)
 *   sigma_static[Nc]: column 1 of mesh.sigma (random-walk trigger, imc_transport.jl:289); may be NULL if !randomwalk
 *   bee, radsource, temp [Nc]
 *   tsurf_bottom[nx], tsurf_top[nx], tsurf_left[ny], tsurf_right[ny]; in 1-D only tsurf_left[1], tsurf_right[1] are read */
int imc_set_mesh(imc_handle h, const double* dx, const double* dy,
                 const double* sigma_a_const, const double* sigma_a_pow,
                 const double* sigma_s_const, const double* sigma_s_pow,
                 const double* sigma_static, const double* bee, const double* radsource, const double* temp,
                 const double* tsurf_bottom, const double* tsurf_top,
                 const double* tsurf_left, const double* tsurf_right);

/* Transport.randomwalk_table (imc_transport.jl:786-797): fills the engine's (aVals, prVals, ptVals)
 * with n entries of LinRange(a_lo, a_hi, n) (reference: 0, 10, 1000; MixedPrecisionIMC.jl:129).
 * Optional outputs (may be NULL) receive the tables. */
int imc_rw_table(imc_handle h, double a_lo, double a_hi, int32_t n, double* a_vals, double* pr_vals, double* pt_vals);

/* Update.update (imc_update.jl:12-70): beta, bee (linearized), sigma_a, sigma_s, fleck. */
int imc_update(imc_handle h, double dt);

/* Sourcing.sourcing (imc_sourcing.jl:12-370).  n_census_global < 0: use this engine's own count
 * (single GPU); otherwise the all-rank census count for the NMAX cap (:133-136). */
int imc_source(imc_handle h, double dt, int64_t n_input, double cellmin, int64_t step,
               int64_t n_census_global, imc_source_stats* out);

/* Transport.MC / MC_RW / MC2D (imc_transport.jl:13-210, :212-479, :483-732). */
int imc_transport(imc_handle h, double dt, int64_t step, imc_transport_stats* out);

/* Clean.clean (imc_clean.jl:6-19): stable removal of dead particles.  *n_alive = length(particles) afterwards.
 * The CUDA engine may leave a few dead entries in its list on populations above 2^22 (they are skipped by every
 * kernel and removed once they exceed 1/32 of the list, or before imc_get_particles); counts, order, ids and all
 * results are those of the compacted list.  Never with replay tapes, EXACT tallies or outcome records. */
int imc_clean(imc_handle h, int64_t* n_alive);

/* Tally.tally (imc_tally.jl:11-149) = imc_tally_local (census radiation tally into the reduce
 * buffer) + [multi-GPU: host all-reduces the reduce buffer] + imc_tally_finish (per-cell update). */
int imc_tally(imc_handle h, double t, double dt, imc_tally_stats* out);
int imc_tally_local(imc_handle h);
int imc_tally_finish(imc_handle h, double t, double dt, imc_tally_stats* out);

/* EnergyCheck.energychecker (imc_energycheck.jl:19-37): conservation residual; resets lostenergy,
 * updates radenergyold. */
int imc_energycheck(imc_handle h, imc_energy_stats* out);

/* Fused fast path: update -> source -> transport -> clean -> tally -> energycheck with no host
 * round trip in between.  Any of the stats pointers may be NULL. */
int imc_step(imc_handle h, double t, double dt, int64_t n_input, double cellmin, int64_t step,
             imc_source_stats* src, imc_transport_stats* trk, imc_tally_stats* tal, imc_energy_stats* chk);

/* The buffer a multi-GPU host must sum over ranks between imc_tally_local and imc_tally_finish:
 * [energydep Nc*Ns | radenergydens Nc | lostenergy | counters...], n slots of 8 bytes.  *kind says how to sum them:
 *   0  Float64 throughout (ATOMIC and EXACT tallies);
 *   1  int64 throughout (the engine accumulates in fixed point: FIXED tallies; sums are then independent of the GPU count).
 * The kind follows from the deck, the tally mode and replicated quantities alone, so it is the same on every rank.
 * *ptr is a device pointer for the CUDA library, a host pointer for the oracle (always kind 0).
 * The call waits for the engine's stream, i.e. for everything issued before it (the tracking kernel, imc_tally_local's
 * census tally): a host that runs its collective on another stream calls it immediately before reducing each part —
 * [energydep] may be reduced as soon as imc_transport has returned, the rest after imc_tally_local. */
int imc_reduce_buffer(imc_handle h, void** ptr, int64_t* n, int32_t* kind);

int imc_get_field(imc_handle h, int32_t field, double* dst, int64_t n);
/* overwrite mesh.temp (and optionally matenergydens when non-NULL): lets a host restart from saved fields */
int imc_set_state(imc_handle h, const double* temp, const double* matenergydens, const double* radenergydens);

/* The same two transfers in the field's own element type, with no Float64 staging on either side: the Julia
 * host's arrays are Array{T} (mesh.matenergydens, mesh.radenergydens, mesh.fleck ...; imc_mesh.jl:117-160), so
 * the shim passes pointer(mesh.x) straight through.  imc_field_elsize returns the element size in bytes:
 * sizeof(T) (2 / 4 / 8) for every field except IMC_FIELD_TEMP, which is 8 once the reference's mesh.temp has
 * turned Float64 (after the first LINEARIZED tally, imc_tally.jl:72) and sizeof(T) before.
 * `bytes` must equal n_elements * elsize.  The caller's buffers may be pageable or pinned host memory, or device memory of
 * this process (the copies use unified addressing): a multi-GPU host uploads the replicated state once, broadcasts it over
 * NVLink and hands every engine a device pointer (bench.py does for its end-to-end loop). */
int32_t imc_field_elsize(imc_handle h, int32_t field);
int imc_get_field_native(imc_handle h, int32_t field, void* dst, int64_t bytes);
int imc_set_state_native(imc_handle h, const void* temp, const void* matenergydens, const void* radenergydens);

/* Per-step history kept by the engine — the reference's mesh.temp_saved / matenergy_saved / radenergy_saved /
 * energyincrease_saved lists (imc_tally.jl:58, :138-142; written out by imc_output.jl:50-57).  After
 * imc_history_enable(h, capacity) every imc_tally / imc_tally_finish appends one snapshot of IMC_FIELD_TEMP (always
 * Float64: the exact image of the reference's value whatever its type at that step), IMC_FIELD_MATENERGYDENS,
 * IMC_FIELD_RADENERGYDENS and IMC_FIELD_NRG_INC (element type T) to device memory, with no host round trip; the host
 * fetches `count` consecutive snapshots when it wants them (end of run, or every k steps) with imc_history_get and may
 * empty the buffer with imc_history_clear.  When `capacity` snapshots are stored further steps are not recorded and are
 * counted in *dropped.  capacity == 0 frees the buffers and stops recording. */
int imc_history_enable(imc_handle h, int64_t capacity);
int imc_history_count(imc_handle h, int64_t* stored, int64_t* dropped);
int imc_history_get(imc_handle h, int32_t field, int64_t first, int64_t count, void* dst, int64_t bytes);
int imc_history_clear(imc_handle h);

/* The CUDA stream (cudaStream_t) every kernel of this engine is launched on, so that a host can bracket
 * calls with its own events; NULL for the oracle. */
void* imc_stream(imc_handle h);

/* Particle population in the reference's array-of-slots layout:
 *   1-D: 9 slots  [origin, time, cellindex, position, mu, freq, energy, startenergy, energyscale]
 *   2-D: 10 slots [time, xindex, yindex, xpos, ypos, mu, frq, energy, startenergy, energyscale]
 * indices are 1-based as in Julia; a dead particle has slot 8 == -1.0.  ids (may be NULL) are the
 * engine's 64-bit particle ids (Philox counter).  imc_num_particles is length(particles); when imc_clean has left dead
 * entries in the engine's list (see there), imc_get_particles removes every flagged entry first. */
int64_t imc_num_particles(imc_handle h);
/* diagnostic: CUDA kernels launched by this engine since creation (0 for the oracle) */
int64_t imc_kernel_launches(imc_handle h);
int imc_get_particles(imc_handle h, double* slots, uint64_t* ids, int64_t capacity);
int imc_set_particles(imc_handle h, const double* slots, const uint64_t* ids, int64_t n);

/* Replay mode (rng_mode = IMC_RNG_TAPE): pre-drawn numbers, draw-major: tape[k * n_slots + slot].
 *   transport tape: slot = particle position in the list when imc_transport is called;
 *                   uniforms are rand(T) values, exponentials are Float64 randexp() values.
 *   source tape:    slot = ordinal of the new particle in emission order; uniforms only. */
int imc_set_transport_tape(imc_handle h, const double* uniforms, int32_t n_uni,
                           const double* exponentials, int32_t n_exp, int64_t n_slots);
int imc_set_source_tape(imc_handle h, const double* uniforms, int32_t n_uni, int64_t n_slots);

/* Sourcing.sample_planck (imc_sourcing.jl:372-399): n frequencies h nu / k T drawn from the Planck spectrum with the
 * Fleck-Cummings series method, in the deck precision.  The reference defines the function but every call site is
 * commented out (grey transport, particle slot `frq` = 1.0: imc_sourcing.jl:171, :209, :342 ...), so nothing in the
 * step uses it; it is exported for hosts that switch the frequency sampling on.  Sample i draws from its own Philox
 * stream (seed, id = i, step) or, in replay mode, from slot i of the source tape.  The reference's loop does not
 * terminate when the first draw exceeds the largest value 90 nsum / pi^4 can reach in the deck precision (Float32:
 * 0.9999989); such a sample is returned as NaN after 100000 terms. */
int imc_sample_planck(imc_handle h, int64_t n, int64_t step, double* out);

/* Restart point inside the library (the reference has none: its state lives in Julia objects a host can copy with
 * `deepcopy(mesh)`, `deepcopy(particles)`; here the engine owns that state, MixedPrecisionIMC.jl:112-133).
 * op = IMC_CKPT_SAVE copies everything a later call can change — the particle population, the per-cell fields
 * (temp, fleck, sigma_a/s, beta, bee, the energy densities, energydep / emittedenergy), lostenergy, the scalars of
 * imc_*_stats — into a second set of device buffers; IMC_CKPT_RESTORE makes that copy current again (any number of
 * times); IMC_CKPT_DROP frees it.  The host restores its own step counter / t / dt.  bench.py uses it to time the
 * resident and the host-buffer path over the SAME time steps. */
enum { IMC_CKPT_SAVE = 0, IMC_CKPT_RESTORE = 1, IMC_CKPT_DROP = 2 };
int imc_checkpoint(imc_handle h, int32_t op);

/* per-particle outcome of the last imc_transport call, for replay checks:
 * event[i] = 0 census, 1 absorbed (energy cut-off), 2 escaped (VACUUM), 3 random-walk kill;
 * nseg[i] = segments tracked.  Indexed like the particle list before imc_clean. */
int imc_get_outcomes(imc_handle h, int32_t* event, int32_t* nseg, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* IMC_B200_H */
