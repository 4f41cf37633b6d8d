/* cslam.h — C ABI of the B200-native SLAM filter hot path (libcslam.so).
 *
 * This is the drop-in boundary for conan-slam's filter layer.  The reference has no
 * FFI: its boundary is the abstract C++ class `Slam` (slam/include/slam.h:56-977) that
 * test/main.cpp drives through std::shared_ptr<Slam> (test/main.cpp:89,204).  Each entry
 * point below names the `Slam` virtual it replaces (file:line under /root/reference).
 * The C++ adaptor (conan_slam_b200/host/slam_gpu.hpp) and the Python mirror
 * (conan_slam_b200/ekf.py, pf.py) re-expose these under the reference's own method names.
 *
 * Conventions (identical to the reference unless stated):
 *   - state X = [x, y, phi, l1x, l1y, ...]  (EKF.cpp:356-357), n = 3 + 2*N, FP64 on the device
 *   - landmark ids `idf` are 1-BASED map slots (EKF.cpp:356-357; table values EKF.cpp:212-226)
 *   - Z is 2 x m column-major = interleaved (range_i, bearing_i) pairs, radians
 *   - R, Q are 2 x 2 (4 doubles; symmetric, so row/column-major coincide)
 *   - resampled `keep` indices are 0-based int32
 *   - all state lives on the GPU behind the handle; host buffers are caller-owned
 *   - one handle = one CUDA stream; calls on a handle must be externally serialised
 *   - functions return CSLAM_OK or an error code and never throw; cslam_last_error()
 *     gives the message.  Numerically skipped updates (reference: slam.h:252-255 zero-gain
 *     path) are counted on the device and reported by cslam_*_sync().
 *   - there is NO CPU fallback: without a usable CUDA device every compute call fails
 *     with CSLAM_ERR_CUDA.
 */
#ifndef CSLAM_H
#define CSLAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSLAM_VERSION 100

#define CSLAM_OK              0
#define CSLAM_ERR_BAD_ARG     1
#define CSLAM_ERR_CAPACITY    2
#define CSLAM_ERR_CUDA        3
#define CSLAM_ERR_NCCL        4
#define CSLAM_ERR_UNSUPPORTED 5

/* Quirk switches.  0 reproduces the reference literally; each
 * bit selects the intended behaviour for ONE quirk.  Same values as oracle/slam_oracle.hpp. */
#define CSLAM_FLAG_REF_LITERAL   0u
#define CSLAM_FLAG_Q1_METRIC_S   (1u << 0) /* gain metric S=LL^T instead of L^T L      slam.h:250-260 */
#define CSLAM_FLAG_Q2_FULL_WIDTH (1u << 1) /* predict all cross-cov columns           EKF.cpp:442-443 */
#define CSLAM_FLAG_Q5_RETURN_ZN  (1u << 2) /* dataAssociate returns new features      EKF.cpp:308-325 */
#define CSLAM_FLAG_Q9_METRIC_S   (1u << 3) /* gaussEvaluate exponent uses S^-1        PF.cpp:287-300  */
#define CSLAM_FLAG_Q10_SEARCH    (1u << 4) /* resample Keep[c]=min{i:sel[c]<cumW[i]}  PF.cpp:566-574  */
#define CSLAM_FLAG_INTENDED      0x1Fu

/* Largest number of observations one update/gate/augment call accepts (they travel as
 * kernel parameters, no H2D copy); a joint (batch) update accepts at most 32 (rank 64). */
#define CSLAM_MAX_OBS       64
#define CSLAM_MAX_BATCH_OBS 32

typedef struct cslam_ekf cslam_ekf_t;
typedef struct cslam_pf cslam_pf_t;
typedef struct cslam_world cslam_world_t;

const char* cslam_last_error(void);
int cslam_version(void);
int cslam_device_count(int* count);
/* Diagnostics (bench.py): the FP64 tensor-core (DMMA.8x8x4) peak of `device` in TFLOP/s, measured by a short
 * register-resident microbenchmark (the denominator of the joint update's roofline; MEASURED_PEAKS.json carries
 * HBM and bf16 figures only). */
int cslam_dmma_peak(int device, double* tflops);
/* Diagnostics (bench.py): number of CUDA kernels this library has launched since it was loaded. */
unsigned long long cslam_kernel_launches(void);

/* ------------------------------------------------------------------ EKF-SLAM ---- */

/* Replaces `new EKF(LM, WP)` (EKF.cpp:3-7) + the driver-owned X(3)=0, P(3x3)=0
 * (test/main.cpp:107-108).  Pre-allocates X and the joint covariance for
 * `capacity_landmarks` landmarks (n_cap = 3 + 2*capacity; P is n_cap x ld FP64,
 * row-major, upper triangle authoritative) so augmentation never reallocates. */
int cslam_ekf_create(cslam_ekf_t** out, int capacity_landmarks, int device, unsigned flags);
/* Multi-GPU: one process per GPU, the joint covariance ROW-SHARDED over `world` ranks in
 * block-cyclic 128-row tiles (very large maps: 60,000 landmarks = 115 GB).  X and rows 0..2 of P
 * are replicated; per update only the observed landmark's two columns of P are exchanged (one
 * NCCL all-reduce of 2*n doubles), then every rank forms the gain and streams its own rows.
 * SPMD contract: every rank makes the same calls in the same order (accessors included — 
 * cslam_ekf_get_cov_block is a collective).  nccl_unique_id: 128 bytes from cslam_nccl_unique_id()
 * on rank 0, distributed by the host (torch.distributed / MPI / a file).  NCCL is dlopen'ed. */
int cslam_nccl_unique_id(void* out128);
int cslam_ekf_create_sharded(cslam_ekf_t** out, int capacity_landmarks, int device, unsigned flags, int rank,
                             int world, const void* nccl_unique_id);
/* Sharded handles exchange the observed columns of P once per scan.  After export on every rank / exchange on the
 * host / import (2 x 64-byte CUDA-IPC handles per rank, as for the particle filter), that exchange is fused into the
 * snapshot kernel: every rank writes the entries it stores straight into every peer's buffer over NVLink as flagged
 * 16-byte cells (the reader polls the cells it needs) — no collective, fence or flag on the critical path.  Without it the columns travel by one NCCL all-reduce. */
int cslam_ekf_ipc_export(cslam_ekf_t* h, void* out128);
int cslam_ekf_ipc_import(cslam_ekf_t* h, const void* all_ranks_128_each, int world);
int cslam_ekf_destroy(cslam_ekf_t* h);
/* Run this handle's kernels on a caller-provided cudaStream_t (e.g. a torch stream).  On a deferred-pass handle
 * (below) give that stream a priority above the default (cudaStreamCreateWithPriority): the small kernels of a scan
 * are then dispatched beside the running covariance pass instead of behind its grid.  The handle's own stream has it. */
int cslam_ekf_set_stream(cslam_ekf_t* h, void* cuda_stream);
/* Wait for the stream; *skipped_updates (nullable) = updates skipped as non-SPD since create/reset. */
int cslam_ekf_sync(cslam_ekf_t* h, int* skipped_updates);
/* Large (capacity >= 1023 landmarks) and sharded handles DEFER the covariance passes: every Kalman update of the
 * reference is P <- P - W1 W1^T (slam.h:260; the heading update slam.h:718 has the same shape), so the
 * updates of consecutive calls — the heading updates of the control steps and the sequential landmark
 * updates of a scan — accumulate as panel rows and ONE pass over the covariance applies up to 64 of them
 * (32 sequential landmark updates; 32 rows from four GPUs on, 16 without room for a second covariance array;
 * environment CSLAM_LAZY_BANK overrides)
 * (X, rows 0..2 of P and the landmarks' 2x2 diagonal blocks are always current; gains read the few other
 * entries of P they need through the pending terms).  Results are those of the eager sequence up to rounding.
 * cslam_ekf_flush applies everything that is pending (asynchronously, in stream order); accessors, augment,
 * the joint update, save and sync do so themselves.  pass_count: passes launched so far / rows pending. */
int cslam_ekf_flush(cslam_ekf_t* h);
int cslam_ekf_pass_count(cslam_ekf_t* h, unsigned long long* passes, int* pending_rows);
int cslam_ekf_n(const cslam_ekf_t* h);             /* state dimension n = X.rows()            */
int cslam_ekf_num_landmarks(const cslam_ekf_t* h); /* (n-3)/2                                  */
int cslam_ekf_capacity(const cslam_ekf_t* h);

/* Slam::predict(X,P,v,swa,Q,wb,dt)                      slam.h:841-847 -> EKF.cpp:406-455 */
int cslam_ekf_predict(cslam_ekf_t* h, double v, double swa, const double Q[4], double wb, double dt);
/* Slam::observeHeading(X,P,phi,useHeading)              slam.h:788 -> EKF.cpp:328-352 -> slam.h:700-725 */
int cslam_ekf_observe_heading(cslam_ekf_t* h, double phi, int use_heading);
/* k consecutive control steps — per step Slam::predict then Slam::observeHeading, i.e. the body of
 * test/main.cpp:140-168 for k iterations (the controls v[i], swa[i] and the measured heading phi[i] do not
 * depend on the filter, so a driver can hand over all steps up to the next observation at once).  On
 * small maps (n <= 1024: the reference's own 30-landmark world) all k steps run in ONE single-CTA
 * launch — the first "next" row of SURVEY.md §8f; larger / sharded maps run the per-step kernels.
 * Results are bit-identical to k calls of cslam_ekf_predict + cslam_ekf_observe_heading.
 * pose_trace (nullable, host, 3*k doubles): X[0..2] after each step (the reference prints X every
 * iteration, test/main.cpp:134-137); passing it makes the call synchronous. */
int cslam_ekf_control_steps(cslam_ekf_t* h, int k, const double* v, const double* swa, const double* phi,
                            int use_heading, const double Q[4], double wb, double dt, double* pose_trace);
/* Slam::dataAssociate(X,P,Z,R,gate1,gate2)              slam.h:482-487 -> EKF.cpp:235-326 (+131-144)
 * Per observation i: jbest[i] = 1-based nearest in-gate landmark (0 = none, lowest j wins ties),
 * is_new[i] = (jbest==0 && min_j nis > gate2).  nbest/outer are optional (nullable):
 * nbest[i] = nd of jbest (inf if none); outer[i] = min_j nis_j — equal to the reference's
 * `outer` whenever the reference consults it, i.e. when jbest==0. */
int cslam_ekf_gate(cslam_ekf_t* h, const double* Z, int m, const double R[4], double gate1, double gate2,
                   int32_t* jbest, uint8_t* is_new, double* nbest, double* outer);
/* Slam::update(X,P,Z,R,idf,batch)                       slam.h:938-943 -> EKF.cpp:481-496
 *   batch=0: singleUpdate EKF.cpp:457-479 (re-linearised per observation)
 *   batch=1: batchUpdate  EKF.cpp:93-129  (one joint rank-2m update for m <= CSLAM_MAX_BATCH_OBS = 32; more
 *            observations are applied as successive joint updates of 32, each chunk linearised at the state
 *            the previous one produced — an explicit, documented deviation: the reference has no rank limit)
 * both through Slam::choleskyUpdate slam.h:235-266.  Asynchronous. m == 0 is a no-op. */
int cslam_ekf_update(cslam_ekf_t* h, const double* Z, const int32_t* idf, int m, const double R[4], int batch);
/* The observation step of test/main.cpp:186-189 with known associations — update(X, P, ZF, RE, IDF, batch)
 * (EKF.cpp:481-496) immediately followed by augment(X, P, ZN, RE) (EKF.cpp:9-26) — as ONE call.  On small maps
 * (n + 2*mn <= 1024, single GPU, batch = 1: the reference's own 30-landmark world and its default switch
 * slam.h:101) the joint update and all augmentations run in ONE single-CTA launch instead of 5 + mn
 * (SURVEY.md §8f row 1), bit-identical to the two separate calls; otherwise it is exactly those two calls. */
int cslam_ekf_observe_step(cslam_ekf_t* h, const double* ZF, const int32_t* idf, int mf, const double* ZN, int mn,
                           const double R[4], int batch);
/* One observation cycle with NO host round trip between association and update — the call pair
 * test/main.cpp:193-195 (dataAssociate EKF.cpp:235-326, then update with batch=false -> singleUpdate
 * EKF.cpp:457-479) as one asynchronous submission: the gate kernel leaves the association indices in
 * device memory, every gain / covariance kernel reads its own index there and exits when no landmark
 * passed the gate.  jbest / is_new (nullable) are read back behind the gate kernel, overlapping the
 * updates; pass NULL for a fully asynchronous scan.  Results are identical to cslam_ekf_gate followed by
 * cslam_ekf_update(batch = 0) on the associated observations.  Works on sharded handles too (every rank
 * computes the same indices from its replicated gate inputs; SPMD call contract). */
int cslam_ekf_scan(cslam_ekf_t* h, const double* Z, int m, const double R[4], double gate1, double gate2,
                   int32_t* jbest, uint8_t* is_new);
/* Running total of observations that cslam_ekf_scan associated (and therefore applied as updates) since
 * the handle was created — the one read an asynchronous driver needs after a series of scans issued with
 * jbest == NULL.  Synchronises the stream. */
int cslam_ekf_scan_associations(cslam_ekf_t* h, unsigned long long* total);
/* Slam::augment(X,P,Z,R)                                slam.h:190-191 -> EKF.cpp:9-26 -> :28-91 */
int cslam_ekf_augment(cslam_ekf_t* h, const double* Z, int m, const double R[4]);

/* State / covariance accessors (the driver owns X and P in the reference; here they are
 * read back on request).  get_cov_block returns a dense row-major nr x nc block with the
 * lower triangle mirrored from the authoritative upper one. */
int cslam_ekf_get_state(cslam_ekf_t* h, double* X, int max_n);
int cslam_ekf_get_cov_block(cslam_ekf_t* h, int r0, int c0, int nr, int nc, double* out);
/* Principal sub-matrix of P for an arbitrary set of k <= 8192 state indices (0-based positions in X): the joint
 * marginal of the pose and a few landmarks without reading P back; out is dense row-major k x k.  Collective
 * on sharded handles. */
int cslam_ekf_get_cov_gather(cslam_ekf_t* h, const int32_t* idx, int k, double* out);
/* Landmark marginals without reading P back (visualisation of the uncertainty ellipses, README.md:15-22;
 * "next" row of SURVEY.md §8f): for the 1-based landmarks first .. first+count-1 the 2x2 diagonal block
 * packed as out[3*k + {0,1,2}] = (P_ff, P_f,f+1, P_f+1,f+1).  Collective on sharded handles. */
int cslam_ekf_get_landmark_covs(cslam_ekf_t* h, int first_landmark, int count, double* out);
/* Checkpoint / restore of X and the upper triangle of P (the reference keeps X, P in the driver and has
 * no persistence): header + X[n] + row i of P from the diagonal on, 4*n*(n+1) bytes of
 * covariance.  load() requires n <= the handle's capacity, the same quirk flags and the same sharding; it
 * validates the file (length, header) BEFORE it touches device state.  Sharded handles: every rank calls with
 * the same path and writes / reads its own rows in <path>.r<rank>of<world>. */
int cslam_ekf_save(cslam_ekf_t* h, const char* path);
int cslam_ekf_load(cslam_ekf_t* h, const char* path);
/* Load a state (tests / benchmarks / checkpoint restore): X has n entries, P is a dense
 * row-major n x n matrix (only j >= i is read) or NULL for zeros. */
int cslam_ekf_reset(cslam_ekf_t* h, const double* X, int n, const double* P);
/* Diagnostics (bench.py roofline): CUDA-event timing of every covariance-update kernel launch
 * (slam.h:260 / slam.h:718) on the handle's stream between begin and end.  *ms = summed kernel
 * time, *launches = how many, *bytes = summed ALGORITHMIC bytes (8*n*(n+1) per launch: one read
 * and one write of the upper triangle). */
int cslam_ekf_profile_begin(cslam_ekf_t* h, int max_launches);
int cslam_ekf_profile_end(cslam_ekf_t* h, double* ms, int* launches, double* bytes);
/* Device pointers for zero-copy consumers (bench, visualisers): X (n doubles), P (row-major, ld). */
int cslam_ekf_device_ptrs(cslam_ekf_t* h, void** dX, void** dP, size_t* ld);

/* ----------------------------------------------- observation front-end (simulated world) ---- */

/* Slam::getObservations(XTrue, LM, tags, maxRange) (slam.h:575-582 -> getVisibleLandmarks :608-683 ->
 * computeRangeBearing :339-368) with the world's landmarks resident on the device — the O(N) step right
 * before the filter hot path in test/main.cpp:177 ("next" row 2 of SURVEY.md §8f).  landmarks_2xN is
 * the reference's LM (2 x N column-major: x_i, y_i interleaved).  observe() returns the visible
 * landmarks in landmark order (as the reference's loop does): Z[2k], Z[2k+1] = range, unwrapped bearing;
 * tags[k] = 1-based landmark number; *m_out = how many are visible (only the first max_out are written).
 * Sensor noise (slam.h:168-178) stays with the caller: its draws are inputs. */
int cslam_world_create(cslam_world_t** out, const double* landmarks_2xN, int num_landmarks, int device);
int cslam_world_destroy(cslam_world_t* w);
int cslam_world_observe(cslam_world_t* w, const double x_true[3], double max_range, int max_out, double* Z,
                        int32_t* tags, int* m_out);

/* getObservations + Slam::dataAssociateTable(X, Z, idz, table) (slam.h:454-457 -> EKF.cpp:146-233), i.e.
 * test/main.cpp:177-186, without leaving the device in between ("next" row 2 of SURVEY.md §8f, second half): the
 * association table (the reference's mTABLE, slam.h:105) lives on the device; the visible landmarks are split into
 * known ones (ZF, idf = 1-based map slots) and new ones (ZN, which receive the slots num_map_landmarks + 1, ... in
 * list order, EKF.cpp:212-226) and ONE read-back returns both lists.  *mf / *mn are the true counts (at most
 * max_out entries are written to each list).  Follow with cslam_ekf_update + cslam_ekf_augment or
 * cslam_ekf_observe_step. */
int cslam_world_observe_associate(cslam_world_t* w, const double x_true[3], double max_range, int num_map_landmarks,
                                  int max_out, double* ZF, int32_t* idf, int* mf, double* ZN, int* mn);
int cslam_world_get_table(cslam_world_t* w, int32_t* table);
int cslam_world_reset_table(cslam_world_t* w);

/* -------------------------------------------------------- particle filter (FastSLAM) ---- */

/* Replaces `new PF(LM, WP)` + Slam::initializeParticles(n) (slam.h:688 -> PF.cpp:319-341):
 * w = 1/P, X = 0, P = 0, no features.  SoA over particles on the device. */
int cslam_pf_create(cslam_pf_t** out, int num_particles, int capacity_landmarks, int device, unsigned flags);
/* Multi-GPU: particles block-partitioned over `world` ranks (one process per GPU); num_particles_local
 * must be a multiple of 32 and equal on every rank.  Per-particle kernels are local.  Resampling:
 * the upper levels of the canonical scan are all-gathered (a few KB), the cumulative weights are
 * all-gathered (8 B per particle), every rank resolves its own output slots, and survivors that live
 * on other ranks are read straight from the owner's HBM over NVLink inside the gather kernel
 * (peer buffers mapped with CUDA IPC: export on every rank, exchange on the host, import).
 * Requires CSLAM_FLAG_Q10_SEARCH.  keep[] then holds GLOBAL particle indices.  SPMD call contract. */
int cslam_pf_create_sharded(cslam_pf_t** out, int num_particles_local, int capacity_landmarks, int device,
                            unsigned flags, int rank, int world, const void* nccl_unique_id);
int cslam_pf_ipc_export(cslam_pf_t* h, void* out640);
int cslam_pf_ipc_import(cslam_pf_t* h, const void* all_ranks_640_each, int world);
int cslam_pf_destroy(cslam_pf_t* h);
int cslam_pf_set_stream(cslam_pf_t* h, void* cuda_stream);
int cslam_pf_sync(cslam_pf_t* h, int* skipped_updates);
int cslam_pf_num_particles(const cslam_pf_t* h);
int cslam_pf_num_features(const cslam_pf_t* h);

/* Slam::predict(Particle_t&,v,swa,Q,wb,dt) for every particle     slam.h:858-863 -> PF.cpp:419-471 */
int cslam_pf_predict(cslam_pf_t* h, double v, double swa, const double Q[4], double wb, double dt);
/* Slam::observeHeading(Particle_t&,phi,use) for every particle     slam.h:796 -> PF.cpp:382-417 */
int cslam_pf_observe_heading(cslam_pf_t* h, double phi, int use_heading);
/* k consecutive control steps of every particle — per step Slam::predict then Slam::observeHeading,
 * test/main.cpp:279-286 for k iterations of the driver loop — in ONE launch: the pose block of a particle
 * is loaded once, stepped k times in registers and stored once.  Bit-identical to k calls of
 * cslam_pf_predict + cslam_pf_observe_heading. */
int cslam_pf_control_steps(cslam_pf_t* h, int k, const double* v, const double* swa, const double* phi,
                           int use_heading, const double Q[4], double wb, double dt);
/* Slam::sampleProposal(Particle_t&,Z,idf,R) for every particle     slam.h:881-884 -> PF.cpp:502-544.
 * xi: 3 standard-normal draws per particle, [p][3] (the values slam.h:753-764 would take
 * from Boost — an input).  xi_on_device != 0: xi is a device pointer. */
int cslam_pf_sample_proposal(cslam_pf_t* h, const double* Z, const int32_t* idf, int m, const double R[4],
                             const double* xi, int xi_on_device);
/* Slam::featureUpdate(Particle_t&,Z,idf,R) for every particle      slam.h:549-552 -> PF.cpp:222-277 */
int cslam_pf_feature_update(cslam_pf_t* h, const double* Z, const int32_t* idf, int m, const double R[4]);
/* Slam::resampleParticles(particles,numEffective,on)               slam.h:871-872 -> PF.cpp:473-500
 *   -> stratifiedResample PF.cpp:546-577 (+ stratifiedRandom :579-596).
 * u: one deviate per slot (an input).  keep (nullable, host, P int32) receives the
 * selected 0-based source index per slot; *neff the effective particle count; *resampled
 * whether the gather-copy ran (neff < num_effective && resample_on).  With num_effective = +infinity
 * ("resample every step") and keep == neff == NULL the call is fully asynchronous (no read-back). */
int cslam_pf_resample(cslam_pf_t* h, const double* u, int u_on_device, double num_effective, int resample_on,
                      int32_t* keep, double* neff, int* resampled);
/* Slam::addOneNewFeature(Particle_t&,Z,R) for every particle       slam.h:134 -> PF.cpp:9-60 */
int cslam_pf_add_features(cslam_pf_t* h, const double* Z, int m, const double R[4]);
/* particles[i].X = multivariateNormalGaussianDistribution(X,P,1); P = 0   test/main.cpp:319-325 */
int cslam_pf_sample_pose(cslam_pf_t* h, const double* xi, int xi_on_device);

/* Diagnostics (bench.py roofline): CUDA-event timing of the gather-copy of every resample between
 * begin and end; *bytes = summed algorithmic bytes (2 x 8 x (12 + 5*Nf) per local particle). */
int cslam_pf_profile_begin(cslam_pf_t* h, int max_resamples);
int cslam_pf_profile_end(cslam_pf_t* h, double* ms, int* resamples, double* bytes);

/* Accessors: weights (P), poses ([p][3]), pose covariances ([p][9] row-major),
 * features of one particle (XF [f][2], PF [f][4] row-major 2x2). */
int cslam_pf_get_weights(cslam_pf_t* h, double* w);
int cslam_pf_get_poses(cslam_pf_t* h, double* X);
int cslam_pf_get_pose_covs(cslam_pf_t* h, double* Pv);
int cslam_pf_get_features(cslam_pf_t* h, int particle, double* XF, double* PF);
/* Slam::extractFeaturesFromParticles (slam.h:517-539): the feature estimates of every particle in particle
 * order, XF[p][f][2], and optionally their packed 2x2 covariances PFp[p][f][3] = (xx, xy, yy).  Either may be NULL. */
int cslam_pf_get_features_all(cslam_pf_t* h, double* XF, double* PFp);
int cslam_pf_set_weights(cslam_pf_t* h, const double* w);
/* w[p] *= factor[p] (np doubles, host or device): an external likelihood term on the importance weights. */
int cslam_pf_scale_weights(cslam_pf_t* h, const double* factor, int on_device);
int cslam_pf_set_poses(cslam_pf_t* h, const double* X, const double* Pv);
/* Checkpoint / restore of the whole particle set: header + every struct-of-arrays row of
 * the current buffer.  load() needs a handle with the same particle count and enough landmark capacity.
 * Single-GPU handles. */
int cslam_pf_save(cslam_pf_t* h, const char* path);
int cslam_pf_load(cslam_pf_t* h, const char* path);
/* Slam::extractStatesFromParticles (slam.h:493-511): pose of the MINIMUM-weight particle (Q13). */
int cslam_pf_extract_state(cslam_pf_t* h, double X[3], int* index);

#ifdef __cplusplus
}
#endif
#endif /* CSLAM_H */
