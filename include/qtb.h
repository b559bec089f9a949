/* qtb.h — C ABI of libqtb.so, the B200-native (sm_100a) engine for QuantiT's block-sparse tensor hot path.
 *
 * The reference (AlexandreFoley/QuantiT) has no FFI: its boundary is the public C++ API in namespace quantit.
 * Every entry point below names the reference interface it replaces (path relative to the reference root, file:line).
 * A thin C++ adaptor converts `const quantit::btensor&` to the plain arrays taken here, using only
 * public accessors (btensor.h:215,216,250,276,289,292,795), and rebuilds a btensor from the arrays returned here
 * through the public raw constructor (btensor.h:168) with every block a zero-copy view of the device arena.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; all integers int64_t unless noted; all data fp64.
 *  - every function returns a qtb_status; no exception crosses the ABI; qtb_last_error() gives the message
 *    (thread-local). The adaptor maps statuses back to the exception types the reference throws (see enum).
 *  - a block tensor = (rank, sections per dim, section sizes, section charges, selection rule, sorted block table,
 *    ONE device arena). Charges are nc-component integer tuples; `mods[c]` = 0 for a Z component, N for C<N>
 *    (reference include/Conserved/quantity.h:62-239). Block order is ascending lexicographic on the block index
 *    (reference include/blockTensor/flat_map.h:31,151).
 *  - one host thread per context (the reference is single-threaded); work is stream-ordered on the
 *    context's CUDA stream. There is NO CPU fallback: every compute entry point fails with QTB_ERR_NO_DEVICE
 *    when no CUDA device is usable.
 */
#ifndef QTB_H
#define QTB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qtb_status
{
	QTB_OK = 0,
	QTB_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument  (btensor.cpp:390-395,501-517; bad_selection_rule, …) */
	QTB_ERR_OUT_OF_RANGE = 2,     /* std::out_of_range      (flat_map.h:232, btensor::block_at)                   */
	QTB_ERR_DOMAIN = 3,           /* std::domain_error      (btensor.cpp:399-404)                                 */
	QTB_ERR_LOGIC = 4,            /* std::logic_error       (NaN in eig2x2Mat, dmrg.cpp:566-569)                  */
	QTB_ERR_RUNTIME = 5,          /* std::runtime_error     (dmrg.cpp:261-263)                                    */
	QTB_ERR_CHECK = 6,            /* c10::Error from TORCH_CHECK (btensor.cpp:802,819,844-846)                    */
	QTB_ERR_CUDA = 7,             /* a CUDA runtime call failed                                                   */
	QTB_ERR_NO_DEVICE = 8         /* no usable CUDA device: the engine refuses to run (no CPU fallback)           */
} qtb_status;

typedef struct qtb_ctx qtb_ctx;       /* device + stream + memory pool + plan cache                    */
typedef struct qtb_tensor qtb_tensor; /* host block table + (shared, ref-counted) device arena         */

const char *qtb_last_error(void);
const char *qtb_version(void);

/* ---- context ------------------------------------------------------------------------------------------------ */
/* stream == NULL: the context creates its own non-blocking stream. Otherwise `stream` is a cudaStream_t the caller
 * owns (e.g. torch's current stream). */
qtb_status qtb_ctx_create(int device, void *stream, qtb_ctx **out);
void qtb_ctx_destroy(qtb_ctx *ctx);
qtb_status qtb_ctx_sync(qtb_ctx *ctx);
/* returns the engine's cached (free) device blocks to the driver; live tensors are untouched. The reference has no
 * counterpart (libtorch's CUDA caching allocator: c10::cuda::CUDACachingAllocator::emptyCache()). */
qtb_status qtb_ctx_trim(qtb_ctx *ctx);
/* Where the block-pair matching of a contraction runs (the two-pointer merge over "columns" of btensor::tensordot,
 * sources/btensor.cpp:2057-2108): mode 0 = host sort, 1 = device sort / segmented-match kernels (qtb_match.cu),
 * -1 (default) = by size (device from 8192 blocks on). The plan — output block set, pair order — is identical either
 * way. qtb_ctx_device_matches: how many plans of this context were matched on the device. */
qtb_status qtb_ctx_set_device_planner(qtb_ctx *ctx, int mode);
qtb_status qtb_ctx_device_matches(qtb_ctx *ctx, int64_t *count);
void *qtb_ctx_stream(qtb_ctx *ctx);
/* counters since context creation: [0] kernel launches of this library, [1] grouped-GEMM launches,
 * [2] plans built, [3] plan-cache hits, [4] bytes host->device, [5] bytes device->host,
 * [6] algorithmic GEMM flops issued (sum 2mnk), [7] device bytes currently allocated */
qtb_status qtb_ctx_counters(qtb_ctx *ctx, int64_t out[8]);

/* ---- charge-sector sharding across the GPUs of one node (no reference counterpart: the reference
 * is single-process, its only parallelism is the BLAS threading inside libtorch) -----------------------------------
 * One process (one qtb_ctx) per GPU, all ranks hold the SAME tensors (MPS, environments, MPO) and make the SAME calls.
 * With world > 1 the engine computes, on each rank, only the output blocks that rank owns — charge sectors of one
 * designated leg of the result, assigned by a longest-processing-time-first balance of the planner's flop counts —
 * for H_eff.psi (qtb_heff_apply, qtb_two_sites_update), the environment updates, the two-site theta inside qtb_dmrg
 * and the charge groups of the block SVD; the pieces are then summed into every rank's (zero-filled) full arena by ONE
 * collective per result. Adding zeros is exact, so an N-GPU run is bit-identical to the 1-GPU run.
 * The collective is supplied by the host as a callback: `allreduce(user, device_ptr, n_doubles, cuda_stream)` must
 * enqueue an in-place fp64 sum-allreduce over all ranks ordered after prior work of `cuda_stream` (ncclAllReduce on
 * that stream, or torch.distributed as quantit_b200/sharding.py does) and return 0 on success. */
typedef int (*qtb_allreduce_fn)(void *user, double *device_ptr, int64_t n, void *cuda_stream);
qtb_status qtb_ctx_set_sharding(qtb_ctx *ctx, int rank, int world, qtb_allreduce_fn allreduce, void *user);
/* The engine's own NCCL communicator (preferred over the callback: no host code on the data path). libnccl is bound
 * at run time with dlopen (`libnccl_path` may be NULL: "libnccl.so.2" as already mapped by the process, e.g. by torch).
 * Rank 0 obtains a 128-byte unique id, the host broadcasts it to every rank by any means, every rank then calls
 * qtb_ctx_init_nccl (collective: ncclCommInitRank). Also sets rank / world like qtb_ctx_set_sharding. */
qtb_status qtb_nccl_unique_id(const char *libnccl_path, char out[128]);
qtb_status qtb_ctx_init_nccl(qtb_ctx *ctx, int rank, int world, const char unique_id[128], const char *libnccl_path);
/* longest-processing-time-first assignment of n weighted sections to `world` ranks (pure host; deterministic: ties go
 * to the lower section index / lower rank). owner_out[n]. */
qtb_status qtb_lpt_assign(int64_t n, const double *weights, int world, int32_t *owner_out);

/* ---- tensors: storage (replaces class btensor's block list, btensor.h:105-110,821-837) ---------------------- */
/* Create from host data. `block_index` is [nblocks*rank] in any order (sorted internally); `host_data` holds the
 * blocks back to back in the GIVEN order, each C-contiguous with the dims implied by its sections; NULL = zeros.
 * Mirrors the raw constructor btensor.h:168 + check_tensor (btensor.cpp:405-467): blocks violating the selection
 * rule or repeated -> QTB_ERR_INVALID_ARGUMENT. */
qtb_status qtb_tensor_create(qtb_ctx *ctx, int64_t rank, int64_t nc, const int64_t *mods, const int64_t *nsec,
                             const int64_t *sec_sizes, const int64_t *cvals, const int64_t *sel, int64_t nblocks,
                             const int64_t *block_index, const double *host_data, qtb_tensor **out);
/* Zero-copy adoption of blocks that already live on the device (e.g. the data_ptr() of CUDA torch tensors held by a
 * quantit::btensor): `block_ptr[b]` is a device pointer, `block_strides` [nblocks*rank] in elements. The memory is
 * NOT owned; the caller keeps it alive while the handle (or any view derived from it) exists. */
qtb_status qtb_tensor_adopt(qtb_ctx *ctx, int64_t rank, int64_t nc, const int64_t *mods, const int64_t *nsec,
                            const int64_t *sec_sizes, const int64_t *cvals, const int64_t *sel, int64_t nblocks,
                            const int64_t *block_index, void *const *block_ptr, const int64_t *block_strides,
                            qtb_tensor **out);
void qtb_tensor_free(qtb_tensor *t);

/* structure queries (btensor::dim :276, section_numbers :289, section_sizes :215, get_cvals :292, selection_rule) */
int64_t qtb_tensor_rank(const qtb_tensor *t);
int64_t qtb_tensor_nc(const qtb_tensor *t);
int64_t qtb_tensor_nblocks(const qtb_tensor *t);
int64_t qtb_tensor_total_sections(const qtb_tensor *t);
int64_t qtb_tensor_numel(const qtb_tensor *t); /* stored elements */
qtb_status qtb_tensor_structure(const qtb_tensor *t, int64_t *nsec, int64_t *sec_sizes, int64_t *cvals, int64_t *sel,
                                int64_t *mods);
/* block table in block order: index [nb*rank], dims [nb*rank], strides [nb*rank] (elements), device pointers [nb].
 * Any output pointer may be NULL. */
qtb_status qtb_tensor_blocks(const qtb_tensor *t, int64_t *index, int64_t *dims, int64_t *strides, void **ptrs);
/* blocks back to back in block order, each C-contiguous (strided views are gathered on the device first) */
qtb_status qtb_tensor_download(qtb_ctx *ctx, const qtb_tensor *t, double *host_out);

/* ---- structural views (no data movement; the copy the reference pays later in permute_bl is fused into the GEMM
 *      operand load) ------------------------------------------------------------------------------------------ */
/* btensor::permute, btensor.cpp:1754-1802 */
qtb_status qtb_permute(qtb_ctx *ctx, const qtb_tensor *a, const int64_t *perm, qtb_tensor **out);
/* btensor::conj for real dtypes = inverse_cvals, btensor.cpp:2156-2172 */
qtb_status qtb_conj(qtb_ctx *ctx, const qtb_tensor *a, qtb_tensor **out);

/* ---- contraction (replaces btensor::tensordot, btensor.h:623, btensor.cpp:1971-2119) ------------------------ */
qtb_status qtb_tensordot(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, int64_t k, const int64_t *dims_a,
                         const int64_t *dims_b, qtb_tensor **out);
/* plan only: number of output blocks, matched pairs and algorithmic flops (sum over pairs of 2*m*n*k) */
qtb_status qtb_tensordot_plan_info(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, int64_t k,
                                   const int64_t *dims_a, const int64_t *dims_b, int64_t *n_out_blocks,
                                   int64_t *n_pairs, int64_t *flops);
/* re-run a contraction into an existing output of identical layout (benchmark loops; no allocation) */
qtb_status qtb_tensordot_into(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, int64_t k,
                              const int64_t *dims_a, const int64_t *dims_b, qtb_tensor *out);
/* One call from host buffers to host buffers (upload A,B; contract; download C). `c_index`/`c_data` may be NULL to
 * query sizes first (n_out_blocks, c_numel). This is the end-to-end entry the reference-side adaptor uses for CPU
 * btensors. */
qtb_status qtb_tensordot_host(qtb_ctx *ctx, int64_t nc, const int64_t *mods,
                              int64_t rank_a, const int64_t *nsec_a, const int64_t *sec_sizes_a, const int64_t *cvals_a,
                              const int64_t *sel_a, int64_t nblocks_a, const int64_t *index_a, const double *data_a,
                              int64_t rank_b, const int64_t *nsec_b, const int64_t *sec_sizes_b, const int64_t *cvals_b,
                              const int64_t *sel_b, int64_t nblocks_b, const int64_t *index_b, const double *data_b,
                              int64_t k, const int64_t *dims_a, const int64_t *dims_b,
                              int64_t *n_out_blocks, int64_t *c_numel, int64_t *c_index, double *c_data);

/* ---- elementwise / Lanczos vector ops (replace btensor mul/add_/div_/sum/sqrt, btensor.cpp:895-976,1204-1302,
 *      1499-1621,2666-2752 as used by dmrg.cpp:585-638) ------------------------------------------------------- */
/* out = alpha*a + beta*b with the union of the two block lists (btensor::add, btensor.cpp:2666-2752) */
qtb_status qtb_axpby(qtb_ctx *ctx, double alpha, const qtb_tensor *a, double beta, const qtb_tensor *b,
                     qtb_tensor **out);
/* btensor::add(other, alpha) = a + alpha*b exactly as the reference computes it (btensor.cpp:2666-2752): its
 * flat_map::merge (flat_map.h:350-425) also multiplies by alpha the blocks of `a` that sort before every block of `b`;
 * that observed behaviour is reproduced (it is visible in `psi_ip -= state*a0`, dmrg.cpp:595). qtb_axpby is the plain
 * linear combination. */
qtb_status qtb_add(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, double alpha, qtb_tensor **out);
/* <a,b> = tensordot over every index (dmrg.cpp:593): result written to *host_out (one device->host sync) */
qtb_status qtb_dot(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, double *host_out);
/* a *= s in place */
qtb_status qtb_scale_(qtb_ctx *ctx, qtb_tensor *a, double s);
/* out(…,k) = a(…,k) * d(k), d rank-1 over a's last dim; blocks with no partner in d are dropped
 * (btensor::mul_ broadcast, btensor.cpp:1204-1302 as used at dmrg.cpp:192,198) */
qtb_status qtb_mul_lastdim(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *d, qtb_tensor **out);

/* ---- block SVD + truncation (replaces quantit::svd(btensor, split, tol, min, max, pow),
 *      blockTensor/LinearAlgebra.h:115, btensor_linalg.cpp:390-534,657-755,805-809; compute_last_index
 *      LinearAlgebra.cpp:57-75) -------------------------------------------------------------------------------- */
/* truncate = 0: plain svd(A, split). max_size < 0 means "no maximum" (SIZE_MAX in the reference). */
qtb_status qtb_svd(qtb_ctx *ctx, const qtb_tensor *a, int64_t split, int truncate, double tol, int64_t min_size,
                   int64_t max_size, double pow, qtb_tensor **u, qtb_tensor **d, qtb_tensor **v);

/* truncate(U, d, V, max, min, tol, pow) / truncate(e, S, max, min, tol, pow) as free-standing operations (reference
 * blockTensor/LinearAlgebra.h:244-247, btensor_linalg.cpp:657-755,768-803): `d` is rank 1 (one section per sector,
 * descending inside a sector), `u` / `v` (either may be NULL) carry the sector as their last index. */
qtb_status qtb_truncate(qtb_ctx *ctx, const qtb_tensor *u, const qtb_tensor *d, const qtb_tensor *v, int64_t max_size,
                        int64_t min_size, double tol, double pow, qtb_tensor **u_out, qtb_tensor **d_out, qtb_tensor **v_out);
/* eigh(A, split) / eigh(A, split, tol, min, max, pow) of a block-symmetric matrix (reference
 * blockTensor/LinearAlgebra.h:159-192, btensor_linalg.cpp:294-389,816-829): A = U diag(e) U^T per charge group, e
 * ascending inside a group, U orthonormal, output structure like d / U of the SVD. The reference's own implementation
 * crashes (SIGSEGV) in the oracle build and its truncation moves the same tuple element twice (:789-790); the contract
 * here is the mathematical one, with the truncation threshold taken over |e|. */
qtb_status qtb_eigh(qtb_ctx *ctx, const qtb_tensor *a, int64_t split, int truncate, double tol, int64_t min_size,
                    int64_t max_size, double pow, qtb_tensor **e, qtb_tensor **u);

/* ---- structural reshapes (btensor::reshape(index_groups), btensor.cpp:2986-3024; btensor::reshape_as<mode>(other),
 *      :3026-3083). Metadata only for packed blocks; strided views are gathered first. ------------------------------ */
/* index_groups: the n_groups boundaries between the groups of consecutive dims that are merged (rank n_groups + 1 out) */
qtb_status qtb_reshape(qtb_ctx *ctx, const qtb_tensor *a, int64_t n_groups, const int64_t *index_groups, qtb_tensor **out);
/* overwrite_cvals = 0: reshape_mode::dims_only (keeps a's selection rule, checks the block fluxes);
 *                  1: reshape_mode::overwrite_c_vals (takes `like`'s selection rule, checks every block against it) */
qtb_status qtb_reshape_as(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *like, int overwrite_cvals, qtb_tensor **out);

/* ---- generalised contraction D = alpha*C + beta*A.B (btensor::tensorgdot, declared at btensor.h:624-627 and never
 *      defined in the reference; semantics of the dense tensorgdot, include/tensorgdot.h:22-86). When C has exactly the
 *      block table of A.B the combination is the epilogue of the grouped GEMM. ------------------------------------- */
qtb_status qtb_tensorgdot(qtb_ctx *ctx, const qtb_tensor *c, const qtb_tensor *a, const qtb_tensor *b, int64_t k,
                          const int64_t *dims_a, const int64_t *dims_b, double beta, double alpha, qtb_tensor **out);

/* ---- packed block tensor <-> file (structure, block table, arena image in one file; the
 *      reference has no btensor serialisation) ---------------------------------------------------------------------- */
qtb_status qtb_tensor_save(qtb_ctx *ctx, const qtb_tensor *t, const char *path);
qtb_status qtb_tensor_load(qtb_ctx *ctx, const char *path, qtb_tensor **out);

/* ---- two-site DMRG pieces (dmrg.cpp) ----------------------------------------------------------------------- */
/* details::hamil2site_times_state, dmrg.cpp:520-531 */
qtb_status qtb_heff_apply(qtb_ctx *ctx, const qtb_tensor *psi, const qtb_tensor *h2, const qtb_tensor *lenv,
                          const qtb_tensor *renv, qtb_tensor **out);
/* compute_left_env / compute_right_env, dmrg.cpp:424-493 */
qtb_status qtb_env_left(qtb_ctx *ctx, const qtb_tensor *h, const qtb_tensor *mps, const qtb_tensor *lenv,
                        qtb_tensor **out);
qtb_status qtb_env_right(qtb_ctx *ctx, const qtb_tensor *h, const qtb_tensor *mps, const qtb_tensor *renv,
                         qtb_tensor **out);
/* two_sites_update = one_step_lanczos + eig2x2Mat + recombination, dmrg.cpp:543-651. */
qtb_status qtb_two_sites_update(qtb_ctx *ctx, const qtb_tensor *psi, const qtb_tensor *h2, const qtb_tensor *lenv,
                                const qtb_tensor *renv, double *energy, qtb_tensor **psi_out);

/* ---- two-site DMRG driver (replaces quantit::dmrg(bMPO&, bMPS&, const dmrg_options&, dmrg_logger&), dmrg.h:39,
 *      dmrg.cpp:92-100 -> details::dmrg_impl :219-273, generate_env :370-409, compute_2sitesHamil :503-515, sweep :127-142,
 *      dmrg_2sites_update::operator() :163-206). Field names follow include/dmrg_options.h:15-34. ------------------- */
typedef struct qtb_dmrg_options
{
	double cutoff;                /* 1e-6   */
	double convergence_criterion; /* 1e-5   */
	int64_t maximum_bond;         /* < 0: unlimited (SIZE_MAX in the reference) */
	int64_t minimum_bond;         /* 4      */
	int64_t maximum_iterations;   /* 1000   */
} qtb_dmrg_options;
/* `mps` is in/out: the handles of the sites that were updated are freed and replaced. `oc` is the orthogonality
 * centre (in/out). `sweep_energy` / `sweep_seconds` / `sweep_mid_bond` (each [maximum_iterations], may be NULL) receive
 * what the reference's dmrg_log_sweeptime logger records (dmrg_logger.h:133-185). The state, its environments and
 * the Krylov vectors never leave the GPU; per two-site update the host reads back one 64-byte scalar record and the
 * singular values. */
qtb_status qtb_dmrg(qtb_ctx *ctx, int64_t length, qtb_tensor *const *mpo, qtb_tensor **mps, int64_t *oc,
                    const qtb_dmrg_options *options, double *energy, int64_t *n_sweeps, double *sweep_energy,
                    double *sweep_seconds, int64_t *sweep_mid_bond);

/* The same with a per-sweep callback in the role of the reference's dmrg_logger (include/dmrg_logger.h:22-60:
 * it_log_all(iteration, E, state) after every sweep): `log(user, iteration, energy, seconds, bond_dims, n)` receives
 * the n = length + 1 bond dimensions of the chain (left edge ... right edge). The adaptor forwards it to the caller's
 * dmrg_logger object while the run is in progress. */
typedef void (*qtb_dmrg_log_fn)(void *user, int64_t iteration, double energy, double seconds, const int64_t *bond_dims,
                                int64_t n);
qtb_status qtb_dmrg_logged(qtb_ctx *ctx, int64_t length, qtb_tensor *const *mpo, qtb_tensor **mps, int64_t *oc,
                           const qtb_dmrg_options *options, double *energy, int64_t *n_sweeps, qtb_dmrg_log_fn log,
                           void *user);

/* <a|obs|b> (obs != NULL: `obs` holds `length` rank-4 MPO tensors) or <a|b> (obs == NULL) of two bMPS of `length` sites.
 * Replaces quantit::contract(const bMPS&, const bMPS&, const bMPO&) and contract(const bMPS&, const bMPS&)
 * (reference include/MPT.h:724-727, sources/MPT.cpp:211-233, 275-292: identity edges on the outer bonds, all-ones on
 * the outer MPO bonds; b enters conjugated). The reference returns a rank-0 btensor; here the scalar itself. */
qtb_status qtb_contract(qtb_ctx *ctx, int64_t length, qtb_tensor *const *a, qtb_tensor *const *b, qtb_tensor *const *obs,
                        double *result);
/* Moves the orthogonality centre of a bMPS from *oc to `target` with untruncated block SVDs; the handles of the sites
 * that change are replaced in place and *oc is updated. Replaces bMPS::move_oc(int) (reference include/MPT.h:611,
 * sources/MPT.cpp:75-111). QTB_ERR_INVALID_ARGUMENT when target is outside the chain (std::invalid_argument there). */
qtb_status qtb_move_oc(qtb_ctx *ctx, int64_t length, qtb_tensor **mps, int64_t *oc, int64_t target);
/* Compresses the bonds of a bMPO by a left-to-right sweep of truncated block SVDs (tolerance `cutoff`, min size 1, no
 * maximum, pow 2); the handles are replaced in place. Replaces bMPO::coalesce(btensor::Scalar cutoff) (reference
 * include/MPT.h:699, sources/MPT.cpp:154-168). */
qtb_status qtb_coalesce(qtb_ctx *ctx, int64_t length, qtb_tensor **mpo, double cutoff);

#ifdef __cplusplus
}
#endif
#endif /* QTB_H */
