/* dlpd.h -- C ABI of the B200 docking-search library (libdlpd.so).
 *
 * This is the drop-in boundary for the hot path of lupoglaz/DeepLocalProteinDocking: the rotation loop of
 * `Docker.dockSE3` (reference: src/Docker/Docker.py:211-236) and the three TorchProteinLibrary operators it
 * calls.  The reference has no FFI table of its own -- its boundary is the Python module surface of
 * TorchProteinLibrary (`_Volume` pybind module) -- so each entry point below names the reference call it
 * replaces.  INTEGRATION.md shows the Python-side binding (ctypes) a maintainer would add.
 *
 * Conventions
 *   - plain C types only; every `const float*` / `float*` volume pointer is a DEVICE pointer owned by the caller
 *     (a torch tensor's data_ptr()), fp32, contiguous, layout [..][x][y][z] with z fastest (NCDHW);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous on it unless
 *     stated otherwise;
 *   - return value 0 = success, otherwise a DLPD_E* code; dlpd_last_error() returns the message of the last
 *     failure on the calling thread;
 *   - a plan is not thread-safe; distinct plans are independent (every kernel-selection option, workspace and profile
 *     counter lives in the plan; the only process-wide state is the table of DEFAULT options new plans start from and
 *     the mutex-guarded workspace of the plan-less standalone operators);
 *   - there is no CPU fallback: every entry point fails with DLPD_ECUDA if no device is usable.
 */
#ifndef DLPD_H
#define DLPD_H

#ifdef __cplusplus
extern "C" {
#endif

#define DLPD_OK 0
#define DLPD_EINVAL 1      /* bad argument / unsupported size                       */
#define DLPD_ECUDA 2       /* CUDA runtime error (message in dlpd_last_error)       */
#define DLPD_ESTATE 3      /* call order violated (e.g. search before set_receptor) */

#define DLPD_CENTER_BOX 0  /* rotation centre L/2     (voxel i at i * resolution; SURVEY 8c choice 2)               */
#define DLPD_CENTER_GRID 1 /* rotation centre (L-1)/2 (literal affine_grid / grid_sample, align_corners = True:
                              what TorchProteinLibrary's VolumeRotation most plausibly does -- the Python default) */

typedef struct dlpd_plan dlpd_plan;

int dlpd_version(void);
const char* dlpd_last_error(void);

/* Non-zero if (L, C0, C1, H) is a compiled search configuration. */
int dlpd_search_supported(int L, int C0, int C1, int H);
/* Non-zero if the standalone operators accept box size L. */
int dlpd_op_supported(int L);

/* ---- search engine: replaces the body of the loop Docker.py:211-236 -------------------------------------
 * L            box size (Docker.__init__ box_size, Docker.py:18; 80 in local_test.py:55)
 * C0, C1       channels of the two resolution levels (16 @ L, 32 @ L/2: ProteinRepresentationModels.py:35-36)
 * H            hidden width of SimpleFilter (DockingModels.py:29: (C0+C1)/2)
 * max_rot      rotations processed per internal batch (workspace is sized for it)
 */
int dlpd_plan_create(dlpd_plan** plan, int L, int C0, int C1, int H, int max_rot, int center_mode, int device);
int dlpd_plan_destroy(dlpd_plan* plan);
/* bytes of device memory the plan owns */
long long dlpd_plan_bytes(const dlpd_plan* plan);

/* Receptor side, once per pair (Docker.py:204-206): feature volumes vol0 [C0][L^3], vol1 [C1][(L/2)^3] and the
 * clash volume `receptor.sum(dim=1)` [L^3].  Computes and keeps the receptor spectra (the reference re-transforms
 * the receptor inside every VolumeConvolution call, DockingModels.py:71). */
int dlpd_set_receptor(dlpd_plan* plan, const float* vol0, const float* vol1, const float* clash, void* stream);

/* Ligand side, once per pair (Docker.py:208-209): un-rotated feature volumes; `clash` [L^3] may be NULL when
 * per-rotation clash volumes are supplied to dlpd_search (re-projected rotated atoms, Docker.py:221-224). */
int dlpd_set_ligand(dlpd_plan* plan, const float* vol0, const float* vol1, const float* clash, void* stream);

/* SimpleFilter weights (DockingModels.py:28-32), HOST pointers: W1 [H][C0+C1], b1 [H], W2 [H], b2 [1]. */
int dlpd_set_filter(dlpd_plan* plan, const float* W1, const float* b1, const float* W2, const float* b2);

/* One batch of rotations (any n >= 1; processed in chunks of max_rot).
 *   R              device [n][9] row-major rotation matrices (r_batch_gpu, Docker.py:215)
 *   lig_clash_rot  device [n][L^3] already-rotated ligand clash volumes (Docker.py:221-224) or NULL to rotate the
 *                  clash volume given to dlpd_set_ligand with the same operator as the feature volumes
 *   tau            clash threshold (docking_model.threshold_clash, Docker.py:226)
 *   out_best       device [n][4]: score, x, y, z of argmin_t V[t] (first extraction of update_top, Docker.py:91-97;
 *                  ties -> first voxel in row-major order)
 *   out_scores     optional device [n][(2L)^3]: the masked score volume V (Docker.py:232); NULL in production */
int dlpd_search(dlpd_plan* plan, const float* R, const float* lig_clash_rot, int n, float tau, float* out_best,
                float* out_scores, void* stream);

/* dlpd_search that ALSO lists every translation whose masked score is <= score_max: the candidates of the global
 * top-K list kept by Docker.update_top (Docker.py:86-105) once the caller knows its current K-th best score, so the
 * K-iteration extraction loop never needs the score volume.
 *   cand        device [cand_cap][4] int32: (score as float bits, rotation index in [0,n), linear index (x*2L+y)*2L+z, 0),
 *               in no particular order
 *   cand_count  device [1]: number of qualifying translations; if it exceeds cand_cap the list is truncated and the
 *               caller must fall back to dlpd_search with out_scores for this batch */
int dlpd_search_candidates(dlpd_plan* plan, const float* R, const float* lig_clash_rot, int n, float tau, float score_max,
                           float* out_best, int* cand, int cand_cap, unsigned int* cand_count, void* stream);

/* The loop body of Docker.dockE3 (Docker.py:159-180): the ligand is re-projected and pushed through the representation
 * network for EVERY rotation (plain PyTorch, outside this library), so all its volumes arrive already rotated:
 *   lig0_rot [n][C0][L^3], lig1_rot [n][C1][(L/2)^3], lig_clash_rot [n][L^3]   (device)
 * Everything downstream (correlation, filter, mask, argmin) is the same fused path; dlpd_set_ligand is not needed. */
int dlpd_search_prerotated(dlpd_plan* plan, const float* lig0_rot, const float* lig1_rot, const float* lig_clash_rot, int n,
                           float tau, float* out_best, float* out_scores, void* stream);

/* Same with HOST buffers (R_host [n][9] in, best_host [n][4] out): stages through pinned memory, copies inside,
 * and synchronises the stream before returning.  The ligand clash volume given to dlpd_set_ligand is rotated. */
int dlpd_search_host(dlpd_plan* plan, const float* R_host, int n, float tau, float* best_host, void* stream);

/* ---- the reference's own per-rotation clash volumes, inside the library (Docker.py:214-224) -------------------------
 * dlpd_set_ligand_atoms keeps the ligand's bbox-centred atom coordinates (device [nat][3] float64, Angstrom; the typed
 * layout does not matter once summed over types) and the voxel size.  The *_atoms searches then take the float64
 * rotation matrices of the table (r_batch_cpu, Docker.py:214) and, chunk by chunk: cast them to fp32 for the volume
 * rotation (Docker.py:215), rotate + translate the atoms in float64, cast, project and sum over types (Docker.py:221-224,
 * = dlpd_project_clash) and run the fused search on the result.  Nothing per rotation crosses the host boundary.
 *   dlpd_search_atoms            R64 device [n][9] float64; otherwise as dlpd_search
 *   dlpd_search_candidates_atoms as dlpd_search_candidates
 *   dlpd_search_host_atoms       R_host [n][9] float64 and best_host [n][4] are HOST buffers: H2D, search, D2H and a stream
 *                                synchronisation inside -- the end-to-end call bench.py times (`e2e`). */
int dlpd_set_ligand_atoms(dlpd_plan* plan, const double* coords, int nat, float resolution, void* stream);
int dlpd_search_atoms(dlpd_plan* plan, const double* R64, int n, float tau, float* out_best, float* out_scores, void* stream);
int dlpd_search_candidates_atoms(dlpd_plan* plan, const double* R64, int n, float tau, float score_max, float* out_best,
                                 int* cand, int cand_cap, unsigned int* cand_count, void* stream);
int dlpd_search_host_atoms(dlpd_plan* plan, const double* R_host, int n, float tau, float* best_host, void* stream);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------------------
 * Per-kernel device time with CUDA events recorded on the launching stream around every launch of dlpd_search. */
#define DLPD_PROFILE_FWDZ0 0     /* gather + z transform, level 0 (+ clash channel) */
#define DLPD_PROFILE_FWDZ1 1     /* same, level 1 */
#define DLPD_PROFILE_PLANE0 2    /* y/x transforms, receptor multiply, inverse x/y, level 0 */
#define DLPD_PROFILE_PLANE1 3    /* same, level 1 */
#define DLPD_PROFILE_EPILOGUE 4  /* inverse z (level 0) + filter + mask + argmin */
#define DLPD_PROFILE_HIDDEN1 5   /* inverse z (level 1) + level-1 part of the filter's hidden layer */
#define DLPD_PROFILE_KINDS 6
int dlpd_profile_enable(dlpd_plan* plan, int on);
/* Synchronises `stream`, returns accumulated milliseconds and launch counts per kind, and resets them. */
int dlpd_profile_read(dlpd_plan* plan, double* ms, int* launches, void* stream);
/* Number of kernels this library has launched in the calling process so far. */
long long dlpd_launch_count(void);

/* The in-loop ligand clash re-projection of dockSE3 (Docker.py:221-224) for n rotations in one launch:
 * CoordsRotate + CoordsTranslate in float64, cast to float32, TypedCoords2Volume, sum over atom types.
 *   coords device [nat][3] float64 bbox-centred ligand atoms (Angstrom); R device [n][9] float64;
 *   out device [n][L^3] fp32 (overwritten) -- feed it to dlpd_search as lig_clash_rot. */
int dlpd_project_clash(const double* coords, const double* R, int nat, int n, int L, float resolution, float* out,
                       void* stream);

/* TorchProteinLibrary.Volume.TypedCoords2Volume(box_size, resolution).forward(coords, num_atoms_of_type, offsets)
 * (Docker.py:31,204,208; per rotation in dockE3, Docker.py:164): coords device [B][3 nmax] fp32 in the typed layout (atoms
 * of type t of structure b at slots offsets[b][t] .. + num_atoms_of_type[b][t] - 1), both int tables device [B][T];
 * out device [B][T][L^3] fp32 (overwritten): every atom adds exp(-r^2 / 2) to the 5^3 voxels around floor(x / res). */
int dlpd_project_typed(const float* coords, const int* num_atoms_of_type, const int* offsets, int B, int T, int nmax, int L,
                       float resolution, float* out, void* stream);

/* ---- kernel-selection options ----------------------------------------------------------------------------
 * Every plan owns its options: dlpd_plan_set_option changes ONE plan; dlpd_set_option changes the process-wide DEFAULTS that
 * plans created afterwards start from (and that the plan-less standalone operators use).
 * DLPD_OPT_PLANE_MODE     which kernel runs the per-(channel, kz) plane stage: AUTO = one CTA per plane when the plane
 *                         fits an SM's shared memory (N <= 160), else a thread-block cluster (8 CTAs at N = 256, 4 elsewhere) with distributed shared memory;
 *                         SINGLE forces the former (fails if the plane does not fit), CLUSTER prefers the latter wherever it is
 *                         compiled (box sizes 8, 16, 20, 40, 64, 80, 128).  The receptor spectra of
 *                         a plan are laid out for the mode in force at dlpd_set_receptor time: a search after a switch
 *                         that selects the other kernel fails with DLPD_ESTATE until dlpd_set_receptor is called again.
 * DLPD_OPT_EPILOGUE_TILE  0 = automatic, 1 = 2x1, 2 = 2x2 (x,y) columns per epilogue tile (classic epilogue)
 * DLPD_OPT_EPILOGUE_KIND  AUTO / CLASSIC; PIPELINED / DUO / FUSED are measured-and-rejected variants that exist only in builds
 *                         with -DDLPD_EXPERIMENTAL=1 (dlpd_build_flags) -- elsewhere selecting them fails with DLPD_EINVAL
 * DLPD_OPT_FILTER_KIND    how the epilogue evaluates the level-0 half of SimpleFilter's first layer (DockingModels.py:29,
 *                         C0 -> H):  TCGEN05 = tcgen05.mma kind::tf32 with the operand and accumulators in TMEM, every
 *                         factor split into TF32 hi + lo (three products per term, fp32 accumulation: agrees with the FP32
 *                         chain to ~3e-7 of the score range; needs C0 % 8 == 0 and H % 8 == 0);  FP32 = packed FFMA2 chain;
 *                         TENSOR = mma.sync fp16 split (experimental builds only);  AUTO = TCGEN05 where available, else FP32 */
#define DLPD_OPT_PLANE_MODE 1
#define DLPD_OPT_EPILOGUE_TILE 2
#define DLPD_OPT_EPILOGUE_KIND 3
#define DLPD_OPT_FILTER_KIND 4
#define DLPD_FILTER_AUTO 0
#define DLPD_FILTER_FP32 1
#define DLPD_FILTER_TENSOR 2
#define DLPD_FILTER_TCGEN05 3
#define DLPD_EPILOGUE_AUTO 0
#define DLPD_EPILOGUE_CLASSIC 1     /* one consumer team: transform then filter, next tile prefetched            */
#define DLPD_EPILOGUE_PIPELINED 2   /* load / transform / filter of consecutive half-tiles overlap (N <= 160)    */
#define DLPD_EPILOGUE_DUO 3         /* two independent half-tile CTAs per SM (N <= 160)                          */
#define DLPD_EPILOGUE_FUSED 4       /* half-tile ring; every thread interleaves a butterfly task of half-tile i
                                       with the filter of its voxel of half-tile i - 1 (N <= 160)                */
#define DLPD_EPILOGUE_TC_DUO 5      /* experimental builds; tcgen05 filter only: two persistent CTAs per SM on 2x1 half-tiles
                                       (transform passes of one overlap the TMEM staging / MMA waits of the other) */
#define DLPD_EPILOGUE_PIPE_TC 6     /* experimental builds: load / transform / tcgen05-scoring pipeline on half-tiles   */
#define DLPD_PLANE_AUTO 0
#define DLPD_PLANE_SINGLE 1
#define DLPD_PLANE_CLUSTER 2
int dlpd_set_option(int key, int value);
int dlpd_plan_set_option(dlpd_plan* plan, int key, int value);
/* what this build of the library contains */
#define DLPD_BUILD_EXPERIMENTAL 1   /* the rejected epilogue variants and the mma.sync filter are compiled in */
#define DLPD_BUILD_EMULATION 2      /* host emulation of the kernel bodies (tests/emu only; never shipped)      */
int dlpd_build_flags(void);
/* Frees the device workspace the standalone operators (dlpd_volume_correlate) keep per device. */
int dlpd_release_workspace(void);

/* ---- standalone operators (module-level drop-ins) --------------------------------------------------------
 * TorchProteinLibrary.Volume.VolumeRotation.forward(volume, R)   (Docker.py:40,218)
 *   out[b,c,p] = trilinear(vol[b,c], ctr + R_b^T (p - ctr)), zero outside;  vol/out [B][C][L^3], R [B][9] */
int dlpd_volume_rotate(const float* vol, const float* R, float* out, int B, int C, int L, int center_mode,
                       void* stream);
/* TorchProteinLibrary.Volume.VolumeConvolution.forward(v1, v2)   (DockingModels.py:71, Docker.py:225)
 *   out[b,c,d mod 2L] = sum_r v1[b,c,r+d] * v2[b,c,r];  v1,v2 [B][C][L^3] -> out [B][C][(2L)^3] */
int dlpd_volume_correlate(const float* v1, const float* v2, float* out, int B, int C, int L, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DLPD_H */
