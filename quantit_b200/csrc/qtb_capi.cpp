// qtb_capi.cpp — the extern "C" boundary declared in include/qtb.h. No exception crosses it.
#include <cstring>
#include <string>

#include "qtb_core.h"
#include "qtb_ops.h"

using namespace qtb;

static thread_local std::string g_last_error;

struct qtb_ctx
{
	Ctx c;
};
struct qtb_tensor
{
	std::unique_ptr<Tensor> t;
};

template <class F>
static qtb_status guarded(F &&f)
{
	try
	{
		f();
		return QTB_OK;
	}
	catch (const Error &e)
	{
		g_last_error = e.what();
		return e.code;
	}
	catch (const std::bad_alloc &)
	{
		g_last_error = "out of host memory";
		return QTB_ERR_RUNTIME;
	}
	catch (const std::exception &e)
	{
		g_last_error = e.what();
		return QTB_ERR_RUNTIME;
	}
}

// entry points that take a context: the context's device becomes current first (a process may hold contexts on several
// devices; kernels, allocations and function attributes are per device)
template <class F>
static qtb_status guarded(qtb_ctx *ctx, F &&f)
{
	return guarded(
	    [&]()
	    {
		    QTB_REQUIRE(ctx != nullptr, QTB_ERR_INVALID_ARGUMENT, "null context");
		    QTB_CUDA(cudaSetDevice(ctx->c.device));
		    f();
	    });
}

static qtb_tensor *wrap(std::unique_ptr<Tensor> t)
{
	auto *h = new qtb_tensor;
	h->t = std::move(t);
	return h;
}

static Structure make_structure(int64_t rank, int64_t nc, const int64_t *mods, const int64_t *nsec,
                                const int64_t *sec_sizes, const int64_t *cvals, const int64_t *sel)
{
	QTB_REQUIRE(rank >= 0 && rank <= 8, QTB_ERR_INVALID_ARGUMENT, "rank must be in [0,8]");
	QTB_REQUIRE(nc >= 1 && nc <= 8, QTB_ERR_INVALID_ARGUMENT, "number of charge components must be in [1,8]");
	Structure st;
	st.rank = rank;
	st.ct.nc = nc;
	st.ct.mods.assign(nc, 0);
	if (mods)
		st.ct.mods.assign(mods, mods + nc);
	st.nsec.assign(nsec, nsec + rank);
	int64_t tot = 0;
	for (int64_t d = 0; d < rank; ++d)
	{
		QTB_REQUIRE(nsec[d] >= 0, QTB_ERR_INVALID_ARGUMENT, "negative section count");
		tot += nsec[d];
	}
	st.sec_sizes.assign(sec_sizes, sec_sizes + tot);
	st.cvals.assign(cvals, cvals + tot * nc);
	for (size_t i = 0; i < st.cvals.size(); ++i)
		st.cvals[i] = st.ct.norm(st.cvals[i], (int64_t)(i % nc));
	st.sel.assign(sel, sel + nc);
	for (int64_t c = 0; c < nc; ++c)
		st.sel[c] = st.ct.norm(st.sel[c], c);
	st.finalize();
	return st;
}

extern "C"
{

const char *qtb_last_error(void) { return g_last_error.c_str(); }
const char *qtb_version(void) { return "qtb 0.1 (sm_100a, fp64 DMMA grouped GEMM)"; }

qtb_status qtb_ctx_create(int device, void *stream, qtb_ctx **out)
{
	return guarded(
	    [&]()
	    {
		    QTB_REQUIRE(out != nullptr, QTB_ERR_INVALID_ARGUMENT, "null output pointer");
		    int ndev = 0;
		    cudaError_t e = cudaGetDeviceCount(&ndev);
		    if (e != cudaSuccess || ndev == 0)
			    throw Error(QTB_ERR_NO_DEVICE,
			                std::string("no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
			                    "): the qtb engine has no CPU fallback");
		    QTB_REQUIRE(device >= 0 && device < ndev, QTB_ERR_INVALID_ARGUMENT, "device ordinal out of range");
		    QTB_CUDA(cudaSetDevice(device));
		    auto h = std::make_unique<qtb_ctx>();
		    h->c.device = device;
		    if (stream)
			    h->c.stream = (cudaStream_t)stream;
		    else
		    {
			    QTB_CUDA(cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking));
			    h->c.own_stream = true;
		    }
		    cudaDeviceProp prop;
		    QTB_CUDA(cudaGetDeviceProperties(&prop, device));
		    h->c.sm_count = prop.multiProcessorCount;
		    // keep freed arenas cached in the stream-ordered pool: allocation stays off the critical path
		    QTB_CUDA(cudaDeviceGetDefaultMemPool(&h->c.pool, device));
		    uint64_t thresh = UINT64_MAX;
		    QTB_CUDA(cudaMemPoolSetAttribute(h->c.pool, cudaMemPoolAttrReleaseThreshold, &thresh));
		    *out = h.release();
	    });
}
void qtb_ctx_destroy(qtb_ctx *ctx)
{
	if (!ctx)
		return;
	cudaSetDevice(ctx->c.device);
	cudaStreamSynchronize(ctx->c.stream);
	ctx->c.plan_cache.clear();
	delete ctx;
}
qtb_status qtb_ctx_sync(qtb_ctx *ctx)
{
	return guarded(ctx, [&]() { QTB_CUDA(cudaStreamSynchronize(ctx->c.stream)); });
}
qtb_status qtb_ctx_set_device_planner(qtb_ctx *ctx, int mode)
{
	return guarded(ctx,
	    [&]()
	    {
		    QTB_REQUIRE(mode >= -1 && mode <= 1, QTB_ERR_INVALID_ARGUMENT, "device planner mode must be -1, 0 or 1");
		    ctx->c.planner_mode = mode;
	    });
}
qtb_status qtb_ctx_device_matches(qtb_ctx *ctx, int64_t *count)
{
	return guarded(ctx, [&]() { *count = ctx->c.device_matches; });
}
qtb_status qtb_ctx_trim(qtb_ctx *ctx)
{
	return guarded(ctx, [&]() { ctx->c.trim_cache(); });
}
void *qtb_ctx_stream(qtb_ctx *ctx) { return ctx ? (void *)ctx->c.stream : nullptr; }
qtb_status qtb_ctx_counters(qtb_ctx *ctx, int64_t out[8])
{
	return guarded(ctx, 
	    [&]()
	    {
		    for (int i = 0; i < 8; ++i)
			    out[i] = ctx->c.counters[i];
	    });
}

qtb_status qtb_tensor_create(qtb_ctx *ctx, int64_t rank, int64_t nc, const int64_t *mods, const int64_t *nsec,
                             const int64_t *sec_sizes, const int64_t *cvals, const int64_t *sel, int64_t nblocks,
                             const int64_t *block_index, const double *host_data, qtb_tensor **out)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx && out, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    Structure st = make_structure(rank, nc, mods, nsec, sec_sizes, cvals, sel);
		    *out = wrap(make_tensor(ctx->c, st, nblocks, block_index, host_data));
	    });
}

qtb_status qtb_tensor_adopt(qtb_ctx *ctx, int64_t rank, int64_t nc, const int64_t *mods, const int64_t *nsec,
                            const int64_t *sec_sizes, const int64_t *cvals, const int64_t *sel, int64_t nblocks,
                            const int64_t *block_index, void *const *block_ptr, const int64_t *block_strides,
                            qtb_tensor **out)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx && out, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    auto t = std::make_unique<Tensor>();
		    t->st = make_structure(rank, nc, mods, nsec, sec_sizes, cvals, sel);
		    t->nblocks = nblocks;
		    t->index.assign(block_index, block_index + nblocks * rank);
		    for (int64_t b = 0; b < nblocks; ++b)
			    QTB_REQUIRE(t->st.allowed(&t->index[b * rank]), QTB_ERR_INVALID_ARGUMENT,
			                "Invalid argument to construct a block tensor: a block violates the selection rule");
		    std::vector<i64> perm(nblocks, 0);
		    if (rank > 0)
			    perm = sort_blocks(rank, t->index);
		    t->dims_from_structure();
		    uintptr_t base = UINTPTR_MAX;
		    for (int64_t b = 0; b < nblocks; ++b)
		    {
			    QTB_REQUIRE(((uintptr_t)block_ptr[b] & 7) == 0, QTB_ERR_INVALID_ARGUMENT, "block pointer not 8-byte aligned");
			    base = std::min(base, (uintptr_t)block_ptr[b]);
		    }
		    if (nblocks == 0)
			    base = 0;
		    t->arena = std::make_shared<Arena>((double *)base);
		    t->arena->ctx = &ctx->c;
		    ctx->c.arenas.insert(t->arena.get()); // orphaned (not freed) if the context goes first
		    t->strides.resize(nblocks * rank);
		    t->offs.resize(nblocks);
		    for (int64_t nb = 0; nb < nblocks; ++nb)
		    {
			    const int64_t ob = perm[nb];
			    for (int64_t d = 0; d < rank; ++d)
				    t->strides[nb * rank + d] = block_strides[ob * rank + d];
			    t->offs[nb] = (int64_t)(((uintptr_t)block_ptr[ob] - base) / sizeof(double));
		    }
		    t->compute_hash();
		    *out = wrap(std::move(t));
	    });
}

void qtb_tensor_free(qtb_tensor *t) { delete t; }

int64_t qtb_tensor_rank(const qtb_tensor *t) { return t->t->st.rank; }
int64_t qtb_tensor_nc(const qtb_tensor *t) { return t->t->st.ct.nc; }
int64_t qtb_tensor_nblocks(const qtb_tensor *t) { return t->t->nblocks; }
int64_t qtb_tensor_total_sections(const qtb_tensor *t) { return t->t->st.total_sections(); }
int64_t qtb_tensor_numel(const qtb_tensor *t) { return t->t->numel(); }

qtb_status qtb_tensor_structure(const qtb_tensor *t, int64_t *nsec, int64_t *sec_sizes, int64_t *cvals, int64_t *sel,
                                int64_t *mods)
{
	return guarded(
	    [&]()
	    {
		    const Structure &st = t->t->st;
		    if (nsec)
			    std::copy(st.nsec.begin(), st.nsec.end(), nsec);
		    if (sec_sizes)
			    std::copy(st.sec_sizes.begin(), st.sec_sizes.end(), sec_sizes);
		    if (cvals)
			    std::copy(st.cvals.begin(), st.cvals.end(), cvals);
		    if (sel)
			    std::copy(st.sel.begin(), st.sel.end(), sel);
		    if (mods)
			    std::copy(st.ct.mods.begin(), st.ct.mods.end(), mods);
	    });
}
qtb_status qtb_tensor_blocks(const qtb_tensor *t, int64_t *index, int64_t *dims, int64_t *strides, void **ptrs)
{
	return guarded(
	    [&]()
	    {
		    const Tensor &x = *t->t;
		    if (index)
			    std::copy(x.index.begin(), x.index.end(), index);
		    if (dims)
			    std::copy(x.dims.begin(), x.dims.end(), dims);
		    if (strides)
			    std::copy(x.strides.begin(), x.strides.end(), strides);
		    if (ptrs)
			    for (int64_t b = 0; b < x.nblocks; ++b)
				    ptrs[b] = (void *)(x.arena->ptr + x.offs[b]);
	    });
}
qtb_status qtb_tensor_download(qtb_ctx *ctx, const qtb_tensor *t, double *host_out)
{
	return guarded(ctx, [&]() { download(ctx->c, *t->t, host_out); });
}

qtb_status qtb_permute(qtb_ctx *ctx, const qtb_tensor *a, const int64_t *perm, qtb_tensor **out)
{
	(void)ctx;
	return guarded(ctx, 
	    [&]()
	    {
		    std::vector<i64> p(perm, perm + a->t->st.rank);
		    *out = wrap(permute(*a->t, p));
	    });
}
qtb_status qtb_conj(qtb_ctx *ctx, const qtb_tensor *a, qtb_tensor **out)
{
	(void)ctx;
	return guarded(ctx, [&]() { *out = wrap(conj(*a->t)); });
}

qtb_status qtb_tensordot(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, int64_t k, const int64_t *dims_a,
                         const int64_t *dims_b, qtb_tensor **out)
{
	return guarded(ctx, 
	    [&]()
	    {
		    std::vector<i64> da(dims_a, dims_a + k), db(dims_b, dims_b + k);
		    *out = wrap(tensordot(ctx->c, *a->t, *b->t, da, db));
	    });
}
qtb_status qtb_tensordot_plan_info(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, int64_t k,
                                   const int64_t *dims_a, const int64_t *dims_b, int64_t *n_out_blocks,
                                   int64_t *n_pairs, int64_t *flops)
{
	return guarded(ctx, 
	    [&]()
	    {
		    std::vector<i64> da(dims_a, dims_a + k), db(dims_b, dims_b + k);
		    auto p = get_plan(ctx->c, *a->t, *b->t, da, db);
		    if (n_out_blocks)
			    *n_out_blocks = (int64_t)p->outs.size();
		    if (n_pairs)
			    *n_pairs = (int64_t)p->pairs.size();
		    if (flops)
			    *flops = p->flops;
	    });
}
qtb_status qtb_tensordot_into(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, int64_t k,
                              const int64_t *dims_a, const int64_t *dims_b, qtb_tensor *out)
{
	return guarded(ctx, 
	    [&]()
	    {
		    std::vector<i64> da(dims_a, dims_a + k), db(dims_b, dims_b + k);
		    tensordot_into(ctx->c, *a->t, *b->t, da, db, *out->t);
	    });
}

qtb_status qtb_tensordot_host(qtb_ctx *ctx, int64_t nc, const int64_t *mods, int64_t rank_a, const int64_t *nsec_a,
                              const int64_t *sec_sizes_a, const int64_t *cvals_a, const int64_t *sel_a,
                              int64_t nblocks_a, const int64_t *index_a, const double *data_a, int64_t rank_b,
                              const int64_t *nsec_b, const int64_t *sec_sizes_b, const int64_t *cvals_b,
                              const int64_t *sel_b, int64_t nblocks_b, const int64_t *index_b, const double *data_b,
                              int64_t k, const int64_t *dims_a, const int64_t *dims_b, int64_t *n_out_blocks,
                              int64_t *c_numel, int64_t *c_index, double *c_data)
{
	return guarded(ctx, 
	    [&]()
	    {
		    std::vector<i64> da(dims_a, dims_a + k), db(dims_b, dims_b + k);
		    Structure sa = make_structure(rank_a, nc, mods, nsec_a, sec_sizes_a, cvals_a, sel_a);
		    Structure sb = make_structure(rank_b, nc, mods, nsec_b, sec_sizes_b, cvals_b, sel_b);
		    if (c_data == nullptr)
		    { // size query: structure only, nothing is uploaded
			    Tensor ta, tb;
			    ta.st = sa;
			    tb.st = sb;
			    ta.nblocks = nblocks_a;
			    tb.nblocks = nblocks_b;
			    ta.index.assign(index_a, index_a + nblocks_a * rank_a);
			    tb.index.assign(index_b, index_b + nblocks_b * rank_b);
			    if (rank_a)
				    sort_blocks(rank_a, ta.index);
			    if (rank_b)
				    sort_blocks(rank_b, tb.index);
			    ta.dims_from_structure();
			    tb.dims_from_structure();
			    ta.layout_packed();
			    tb.layout_packed();
			    ta.compute_hash();
			    tb.compute_hash();
			    auto p = get_plan(ctx->c, ta, tb, da, db);
			    if (n_out_blocks)
				    *n_out_blocks = p->out_proto.nblocks;
			    if (c_numel)
				    *c_numel = p->out_proto.numel();
			    if (c_index)
				    std::copy(p->out_proto.index.begin(), p->out_proto.index.end(), c_index);
			    return;
		    }
		    auto ta = make_tensor(ctx->c, sa, nblocks_a, index_a, data_a);
		    auto tb = make_tensor(ctx->c, sb, nblocks_b, index_b, data_b);
		    auto tc = tensordot(ctx->c, *ta, *tb, da, db);
		    if (n_out_blocks)
			    *n_out_blocks = tc->nblocks;
		    if (c_numel)
			    *c_numel = tc->numel();
		    if (c_index)
			    std::copy(tc->index.begin(), tc->index.end(), c_index);
		    download(ctx->c, *tc, c_data);
	    });
}

qtb_status qtb_axpby(qtb_ctx *ctx, double alpha, const qtb_tensor *a, double beta, const qtb_tensor *b,
                     qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(axpby_dev(ctx->c, nullptr, alpha, *a->t, nullptr, beta, *b->t, false)); });
}
qtb_status qtb_add(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, double alpha, qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(axpby_dev(ctx->c, nullptr, 1.0, *a->t, nullptr, alpha, *b->t, false, alpha)); });
}
qtb_status qtb_dot(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *b, double *host_out)
{
	return guarded(ctx, 
	    [&]()
	    {
		    double *d = (double *)ctx_alloc(ctx->c, sizeof(double));
		    dot_dev(ctx->c, *a->t, *b->t, d, false);
		    QTB_CUDA(cudaMemcpyAsync(host_out, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->c.stream));
		    QTB_CUDA(cudaStreamSynchronize(ctx->c.stream));
		    ctx->c.counters[5] += sizeof(double);
		    ctx_free(ctx->c, d);
	    });
}
qtb_status qtb_scale_(qtb_ctx *ctx, qtb_tensor *a, double s)
{
	return guarded(ctx, 
	    [&]()
	    {
		    Tensor &t = *a->t;
		    QTB_REQUIRE(t.arena && t.arena->owned && t.packed_canonical(), QTB_ERR_INVALID_ARGUMENT,
		                "scale_: in-place scaling needs a packed tensor that owns its arena");
		    launch_scale(ctx->c, t.arena->ptr, t.arena->numel, nullptr, s, false);
	    });
}
qtb_status qtb_mul_lastdim(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *d, qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(mul_lastdim(ctx->c, *a->t, *d->t)); });
}

qtb_status qtb_svd(qtb_ctx *ctx, const qtb_tensor *a, int64_t split, int truncate, double tol, int64_t min_size,
                   int64_t max_size, double pow, qtb_tensor **u, qtb_tensor **d, qtb_tensor **v)
{
	return guarded(ctx, 
	    [&]()
	    {
		    std::unique_ptr<Tensor> tu, td, tv;
		    block_svd(ctx->c, *a->t, split, truncate != 0, tol, min_size, max_size, pow, tu, td, tv);
		    *u = wrap(std::move(tu));
		    *d = wrap(std::move(td));
		    *v = wrap(std::move(tv));
	    });
}

qtb_status qtb_eigh(qtb_ctx *ctx, const qtb_tensor *a, int64_t split, int truncate, double tol, int64_t min_size,
                    int64_t max_size, double pow, qtb_tensor **e, qtb_tensor **u)
{
	return guarded(ctx,
	    [&]()
	    {
		    std::unique_ptr<Tensor> te, tu;
		    block_eigh(ctx->c, *a->t, split, truncate != 0, tol, min_size, max_size, pow, te, tu);
		    *e = wrap(std::move(te));
		    *u = wrap(std::move(tu));
	    });
}
qtb_status qtb_truncate(qtb_ctx *ctx, const qtb_tensor *u, const qtb_tensor *d, const qtb_tensor *v, int64_t max_size,
                        int64_t min_size, double tol, double pow, qtb_tensor **u_out, qtb_tensor **d_out, qtb_tensor **v_out)
{
	return guarded(ctx,
	    [&]()
	    {
		    QTB_REQUIRE(d != nullptr && d_out != nullptr, QTB_ERR_INVALID_ARGUMENT, "truncate: d is required");
		    std::vector<const Tensor *> units;
		    if (u)
			    units.push_back(u->t.get());
		    if (v)
			    units.push_back(v->t.get());
		    std::unique_ptr<Tensor> td;
		    std::vector<std::unique_ptr<Tensor>> outs;
		    block_truncate(ctx->c, *d->t, units, tol, min_size, max_size, pow, td, outs);
		    *d_out = wrap(std::move(td));
		    size_t k = 0;
		    if (u)
			    *u_out = wrap(std::move(outs[k++]));
		    if (v)
			    *v_out = wrap(std::move(outs[k++]));
	    });
}
qtb_status qtb_reshape(qtb_ctx *ctx, const qtb_tensor *a, int64_t n_groups, const int64_t *index_groups, qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(reshape(ctx->c, *a->t, std::vector<i64>(index_groups, index_groups + n_groups))); });
}
qtb_status qtb_reshape_as(qtb_ctx *ctx, const qtb_tensor *a, const qtb_tensor *like, int overwrite_cvals, qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(reshape_as(ctx->c, *a->t, *like->t, overwrite_cvals != 0)); });
}
qtb_status qtb_tensorgdot(qtb_ctx *ctx, const qtb_tensor *c, const qtb_tensor *a, const qtb_tensor *b, int64_t k,
                          const int64_t *dims_a, const int64_t *dims_b, double beta, double alpha, qtb_tensor **out)
{
	return guarded(ctx,
	    [&]()
	    {
		    *out = wrap(tensorgdot(ctx->c, *c->t, *a->t, *b->t, std::vector<i64>(dims_a, dims_a + k),
		                           std::vector<i64>(dims_b, dims_b + k), beta, alpha));
	    });
}
qtb_status qtb_tensor_save(qtb_ctx *ctx, const qtb_tensor *t, const char *path)
{
	return guarded(ctx, [&]() { save_tensor(ctx->c, *t->t, path); });
}
qtb_status qtb_tensor_load(qtb_ctx *ctx, const char *path, qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(load_tensor(ctx->c, path)); });
}

qtb_status qtb_ctx_set_sharding(qtb_ctx *ctx, int rank, int world, qtb_allreduce_fn allreduce, void *user)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx != nullptr, QTB_ERR_INVALID_ARGUMENT, "null context");
		    QTB_REQUIRE(world >= 1 && rank >= 0 && rank < world, QTB_ERR_INVALID_ARGUMENT, "rank outside [0, world)");
		    QTB_REQUIRE(world == 1 || allreduce != nullptr || ctx->c.nccl_comm != nullptr, QTB_ERR_INVALID_ARGUMENT,
		                "sharding over more than one rank needs an allreduce callback or qtb_ctx_init_nccl");
		    ctx->c.rank = rank;
		    ctx->c.world = world;
		    ctx->c.allreduce_fn = allreduce;
		    ctx->c.allreduce_user = user;
	    });
}
qtb_status qtb_nccl_unique_id(const char *libnccl_path, char out[128])
{
	return guarded([&]() { nccl_unique_id(libnccl_path, out); });
}
qtb_status qtb_ctx_init_nccl(qtb_ctx *ctx, int rank, int world, const char unique_id[128], const char *libnccl_path)
{
	return guarded(ctx, [&]() { ctx_init_nccl(ctx->c, rank, world, unique_id, libnccl_path); });
}
qtb_status qtb_lpt_assign(int64_t n, const double *weights, int world, int32_t *owner_out)
{
	return guarded(
	    [&]()
	    {
		    QTB_REQUIRE(n >= 0 && world >= 1 && (n == 0 || (weights && owner_out)), QTB_ERR_INVALID_ARGUMENT, "bad argument");
		    const auto o = lpt_assign(std::vector<double>(weights, weights + n), world);
		    std::copy(o.begin(), o.end(), owner_out);
	    });
}

qtb_status qtb_heff_apply(qtb_ctx *ctx, const qtb_tensor *psi, const qtb_tensor *h2, const qtb_tensor *lenv,
                          const qtb_tensor *renv, qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(heff_apply(ctx->c, *psi->t, *h2->t, *lenv->t, *renv->t)); });
}
qtb_status qtb_env_left(qtb_ctx *ctx, const qtb_tensor *h, const qtb_tensor *mps, const qtb_tensor *lenv,
                        qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(env_left(ctx->c, *h->t, *mps->t, *lenv->t)); });
}
qtb_status qtb_env_right(qtb_ctx *ctx, const qtb_tensor *h, const qtb_tensor *mps, const qtb_tensor *renv,
                         qtb_tensor **out)
{
	return guarded(ctx, [&]() { *out = wrap(env_right(ctx->c, *h->t, *mps->t, *renv->t)); });
}
qtb_status qtb_two_sites_update(qtb_ctx *ctx, const qtb_tensor *psi, const qtb_tensor *h2, const qtb_tensor *lenv,
                                const qtb_tensor *renv, double *energy, qtb_tensor **psi_out)
{
	return guarded(ctx, [&]()
	               { *psi_out = wrap(two_sites_update(ctx->c, *psi->t, *h2->t, *lenv->t, *renv->t, energy)); });
}

qtb_status qtb_dmrg(qtb_ctx *ctx, int64_t length, qtb_tensor *const *mpo, qtb_tensor **mps, int64_t *oc,
                    const qtb_dmrg_options *options, double *energy, int64_t *n_sweeps, double *sweep_energy,
                    double *sweep_seconds, int64_t *sweep_mid_bond)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx && mpo && mps && oc && options && energy && n_sweeps, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    std::vector<const Tensor *> H(length);
		    std::vector<std::unique_ptr<Tensor>> psi(length);
		    for (int64_t i = 0; i < length; ++i)
		    {
			    H[i] = mpo[i]->t.get();
			    psi[i] = std::move(mps[i]->t); // ownership moves into the sweep and back (also on error)
		    }
		    try
		    {
			    dmrg(ctx->c, length, H.data(), psi, *oc, *options, *energy, *n_sweeps, sweep_energy, sweep_seconds,
			         sweep_mid_bond);
		    }
		    catch (...)
		    {
			    for (int64_t i = 0; i < length; ++i)
				    mps[i]->t = std::move(psi[i]);
			    throw;
		    }
		    for (int64_t i = 0; i < length; ++i)
			    mps[i]->t = std::move(psi[i]);
	    });
}

qtb_status qtb_dmrg_logged(qtb_ctx *ctx, int64_t length, qtb_tensor *const *mpo, qtb_tensor **mps, int64_t *oc,
                           const qtb_dmrg_options *options, double *energy, int64_t *n_sweeps, qtb_dmrg_log_fn log,
                           void *user)
{
	return guarded(ctx,
	    [&]()
	    {
		    QTB_REQUIRE(ctx && mpo && mps && oc && options && energy && n_sweeps, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    std::vector<const Tensor *> H(length);
		    std::vector<std::unique_ptr<Tensor>> psi(length);
		    for (int64_t i = 0; i < length; ++i)
		    {
			    H[i] = mpo[i]->t.get();
			    psi[i] = std::move(mps[i]->t);
		    }
		    try
		    {
			    dmrg(ctx->c, length, H.data(), psi, *oc, *options, *energy, *n_sweeps, nullptr, nullptr, nullptr, log, user);
		    }
		    catch (...)
		    {
			    for (int64_t i = 0; i < length; ++i)
				    mps[i]->t = std::move(psi[i]);
			    throw;
		    }
		    for (int64_t i = 0; i < length; ++i)
			    mps[i]->t = std::move(psi[i]);
	    });
}

qtb_status qtb_contract(qtb_ctx *ctx, int64_t length, qtb_tensor *const *a, qtb_tensor *const *b, qtb_tensor *const *obs,
                        double *result)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx && a && b && result && length >= 1, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    std::vector<const Tensor *> A(length), B(length), H(length);
		    for (int64_t i = 0; i < length; ++i)
		    {
			    QTB_REQUIRE(a[i] && b[i] && (!obs || obs[i]), QTB_ERR_INVALID_ARGUMENT, "null tensor handle");
			    A[i] = a[i]->t.get();
			    B[i] = b[i]->t.get();
			    H[i] = obs ? obs[i]->t.get() : nullptr;
		    }
		    *result = contract(ctx->c, length, A.data(), B.data(), obs ? H.data() : nullptr);
	    });
}

qtb_status qtb_move_oc(qtb_ctx *ctx, int64_t length, qtb_tensor **mps, int64_t *oc, int64_t target)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx && mps && oc && length >= 1, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    std::vector<std::unique_ptr<Tensor>> psi(length);
		    for (int64_t i = 0; i < length; ++i)
		    {
			    QTB_REQUIRE(mps[i] != nullptr, QTB_ERR_INVALID_ARGUMENT, "null tensor handle");
			    psi[i] = std::move(mps[i]->t); // ownership moves into the gauge walk and back (also on error)
		    }
		    try
		    {
			    move_oc(ctx->c, psi, *oc, target);
		    }
		    catch (...)
		    {
			    for (int64_t i = 0; i < length; ++i)
				    mps[i]->t = std::move(psi[i]);
			    throw;
		    }
		    for (int64_t i = 0; i < length; ++i)
			    mps[i]->t = std::move(psi[i]);
	    });
}

qtb_status qtb_coalesce(qtb_ctx *ctx, int64_t length, qtb_tensor **mpo, double cutoff)
{
	return guarded(ctx, 
	    [&]()
	    {
		    QTB_REQUIRE(ctx && mpo && length >= 1, QTB_ERR_INVALID_ARGUMENT, "null argument");
		    std::vector<std::unique_ptr<Tensor>> H(length);
		    for (int64_t i = 0; i < length; ++i)
		    {
			    QTB_REQUIRE(mpo[i] != nullptr, QTB_ERR_INVALID_ARGUMENT, "null tensor handle");
			    H[i] = std::move(mpo[i]->t);
		    }
		    try
		    {
			    coalesce(ctx->c, H, cutoff);
		    }
		    catch (...)
		    {
			    for (int64_t i = 0; i < length; ++i)
				    mpo[i]->t = std::move(H[i]);
			    throw;
		    }
		    for (int64_t i = 0; i < length; ++i)
			    mpo[i]->t = std::move(H[i]);
	    });
}

} // extern "C"
