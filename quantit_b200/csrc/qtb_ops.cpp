// qtb_ops.cpp — block-tensor vector ops and the two-site DMRG pieces, composed from the planner + kernels.
//
// Reference functions restated on the engine's data model (paths relative to the reference root):
//   btensor::add / add_            sources/btensor.cpp:2666-2752
//   btensor::mul_ (broadcast)      sources/btensor.cpp:1204-1302
//   hamil2site_times_state         sources/dmrg.cpp:520-531
//   compute_left_env/right_env     sources/dmrg.cpp:424-493
//   one_step_lanczos, eig2x2Mat, two_sites_update   sources/dmrg.cpp:543-651
#include <cstdlib>

#include "qtb_ops.h"

#include <algorithm>
#include <cstring>

namespace qtb
{

static void require_same_structure(const Tensor &a, const Tensor &b, const char *what)
{ // reference btensor::add_tensor_check, btensor.cpp:2592-2630
	QTB_REQUIRE(a.st.rank == b.st.rank && a.st.ct == b.st.ct && a.st.nsec == b.st.nsec &&
	                a.st.sec_sizes == b.st.sec_sizes && a.st.cvals == b.st.cvals && a.st.sel == b.st.sel,
	            QTB_ERR_INVALID_ARGUMENT, std::string(what) + ": the two tensors do not have the same block structure");
}

static const Tensor &packed(Ctx &ctx, const Tensor &t, std::unique_ptr<Tensor> &hold)
{
	if (t.packed_canonical())
		return t;
	hold = contiguous(ctx, t);
	return *hold;
}

std::unique_ptr<Tensor> axpby_dev(Ctx &ctx, const double *ca_ptr, double ca_mul, const Tensor &a_in,
                                  const double *cb_ptr, double cb_mul, const Tensor &b_in, bool divide_a,
                                  double merge_quirk)
{
	require_same_structure(a_in, b_in, "add");
	std::unique_ptr<Tensor> ha, hb;
	const Tensor &a = packed(ctx, a_in, ha);
	const Tensor &b = packed(ctx, b_in, hb);
	const i64 r = a.st.rank;
	auto out = std::make_unique<Tensor>();
	out->st = a.st;
	// sorted merge of the two block lists
	std::vector<std::pair<i64, i64>> src; // (block in a or -1, block in b or -1)
	i64 i = 0, j = 0;
	while (i < a.nblocks || j < b.nblocks)
	{
		int cmp;
		if (i == a.nblocks)
			cmp = 1;
		else if (j == b.nblocks)
			cmp = -1;
		else if (std::lexicographical_compare(a.idx(i), a.idx(i) + r, b.idx(j), b.idx(j) + r))
			cmp = -1;
		else if (std::lexicographical_compare(b.idx(j), b.idx(j) + r, a.idx(i), a.idx(i) + r))
			cmp = 1;
		else
			cmp = 0;
		const i64 *ix = cmp <= 0 ? a.idx(i) : b.idx(j);
		out->index.insert(out->index.end(), ix, ix + r);
		src.push_back({cmp <= 0 ? i : -1, cmp >= 0 ? j : -1});
		if (cmp <= 0)
			++i;
		if (cmp >= 0)
			++j;
	}
	out->nblocks = (i64)src.size();
	out->dims_from_structure();
	const i64 total = out->layout_packed();
	out->arena = std::make_shared<Arena>(&ctx, total);
	std::vector<VecSeg> segs;
	for (i64 ob = 0; ob < out->nblocks; ++ob)
	{
		VecSeg s{};
		s.n = out->block_numel(ob);
		s.o_off = out->offs[ob];
		s.a_scale = 1.0;
		if (merge_quirk != 1.0 && src[ob].first >= 0 &&
		    (b.nblocks == 0 || std::lexicographical_compare(out->idx(ob), out->idx(ob) + r, b.idx(0), b.idx(0) + r)))
			s.a_scale = merge_quirk; // reference flat_map::insert: elements of `this` below other's first key
		s.a_off = src[ob].first >= 0 ? a.offs[src[ob].first] : -1;
		s.b_off = src[ob].second >= 0 ? b.offs[src[ob].second] : -1;
		if (src[ob].first >= 0)
			QTB_REQUIRE(a.block_numel(src[ob].first) == s.n, QTB_ERR_INVALID_ARGUMENT, "add: block shape mismatch");
		if (src[ob].second >= 0)
			QTB_REQUIRE(b.block_numel(src[ob].second) == s.n, QTB_ERR_INVALID_ARGUMENT, "add: block shape mismatch");
		if (s.n)
			segs.push_back(s);
	}
	launch_axpby(ctx, segs, a.arena->ptr, b.arena->ptr, out->arena->ptr, ca_ptr, ca_mul, cb_ptr, cb_mul, divide_a);
	out->compute_hash();
	return out;
}

void dot_dev(Ctx &ctx, const Tensor &a_in, const Tensor &b_in, double *d_result, bool take_sqrt)
{ // tensordot over every index pair i<->i (dmrg.cpp:593): sum over the blocks present in BOTH operands.
  // b is used as given (the reference passes b.conj(): inverted charges, same values for real dtypes), so the
  // compatibility requirement is "charges pairwise inverse"; callers pass conj() where the reference does.
	QTB_REQUIRE(a_in.st.rank == b_in.st.rank && a_in.st.nsec == b_in.st.nsec, QTB_ERR_CHECK,
	            "contracted dimensions need to match");
	std::unique_ptr<Tensor> ha, hb;
	const Tensor &a = packed(ctx, a_in, ha);
	const Tensor &b = packed(ctx, b_in, hb);
	const i64 r = a.st.rank;
	std::vector<VecSeg> segs;
	i64 i = 0, j = 0;
	while (i < a.nblocks && j < b.nblocks)
	{
		if (std::lexicographical_compare(a.idx(i), a.idx(i) + r, b.idx(j), b.idx(j) + r))
			++i;
		else if (std::lexicographical_compare(b.idx(j), b.idx(j) + r, a.idx(i), a.idx(i) + r))
			++j;
		else
		{
			VecSeg s{};
			s.n = a.block_numel(i);
			QTB_REQUIRE(s.n == b.block_numel(j), QTB_ERR_CHECK, "dot: block shape mismatch");
			s.a_off = a.offs[i];
			s.b_off = b.offs[j];
			s.o_off = 0;
			s.a_scale = 1.0;
			if (s.n)
				segs.push_back(s);
			++i;
			++j;
		}
	}
	launch_dot(ctx, segs, a.arena->ptr, b.arena->ptr, d_result, take_sqrt);
}

std::unique_ptr<Tensor> mul_lastdim(Ctx &ctx, const Tensor &a, const Tensor &d_in)
{ // reference broadcast mul_ with a rank-1 tensor over the last dim (dmrg.cpp:192,198): the outer loop runs over the
  // blocks of d, the inner one over the blocks of a; a block of a without a partner in d is dropped; section charges
  // and selection rules multiply (mul_helpers::shape_compute, btensor.cpp:980-1084).
	QTB_REQUIRE(d_in.st.rank == 1 && a.st.rank >= 1, QTB_ERR_INVALID_ARGUMENT, "mul_lastdim: d must be rank 1");
	const i64 r = a.st.rank;
	QTB_REQUIRE(a.st.ct == d_in.st.ct && a.st.nsec[r - 1] == d_in.st.nsec[0], QTB_ERR_INVALID_ARGUMENT,
	            "mul_lastdim: incompatible sections");
	for (i64 s = 0; s < d_in.st.nsec[0]; ++s)
		QTB_REQUIRE(a.st.size_of(r - 1, s) == d_in.st.size_of(0, s), QTB_ERR_INVALID_ARGUMENT,
		            "mul_lastdim: section sizes differ");
	std::unique_ptr<Tensor> hd;
	const Tensor &d = packed(ctx, d_in, hd);
	auto out = std::make_unique<Tensor>();
	out->st = a.st;
	const i64 nc = a.st.ct.nc;
	for (i64 s = 0; s < d.st.nsec[0]; ++s)
		for (i64 c = 0; c < nc; ++c)
		{
			i64 &cv = out->st.cvals[(a.st.sec_off[r - 1] + s) * nc + c];
			cv = a.st.ct.norm(cv + d.st.charge_of(0, s)[c], c);
		}
	for (i64 c = 0; c < nc; ++c)
		out->st.sel[c] = a.st.ct.norm(a.st.sel[c] + d.st.sel[c], c);
	std::vector<i64> keep;
	for (i64 b = 0; b < a.nblocks; ++b)
		if (d.find_block(&a.idx(b)[r - 1]) >= 0)
			keep.push_back(b);
	out->nblocks = (i64)keep.size();
	for (i64 b : keep)
	{
		out->index.insert(out->index.end(), a.idx(b), a.idx(b) + r);
		out->dims.insert(out->dims.end(), a.dm(b), a.dm(b) + r); // actual block dims (may be trimmed below the section size)
	}
	const i64 total = out->layout_packed();
	out->arena = std::make_shared<Arena>(&ctx, total);
	std::unique_ptr<Tensor> ha;
	const Tensor &ap = packed(ctx, a, ha);
	std::vector<MulSeg> segs;
	for (i64 ob = 0; ob < out->nblocks; ++ob)
	{
		const i64 b = keep[ob];
		const i64 db = d.find_block(&a.idx(b)[r - 1]);
		MulSeg s{};
		s.n = ap.dm(b)[r - 1];
		QTB_REQUIRE(s.n == d.block_numel(db), QTB_ERR_INVALID_ARGUMENT, "mul_lastdim: block size mismatch");
		s.rows = s.n ? ap.block_numel(b) / s.n : 0;
		s.a_off = ap.offs[b];
		s.a_row_stride = s.n;
		s.a_col_stride = 1;
		s.d_off = d.offs[db];
		s.o_off = out->offs[ob];
		if (s.rows > 0 && s.n > 0)
			segs.push_back(s);
	}
	launch_mul_lastdim(ctx, segs, ap.arena->ptr, d.arena->ptr, out->arena->ptr);
	out->compute_hash();
	return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// A chain of three contractions that carries one free leg from the first operand to the result (the bra bond a' through
// H_eff.psi, the new ket bond through an environment update). Single GPU: three grouped-GEMM launches. Sharded
// (qtb_ctx_set_sharding, world > 1): the sections of that leg are balanced over the ranks with the planner's flop
// counts of all three steps, every rank runs the three steps for its sections only — no exchange between the steps,
// the leg's sector is preserved — and ONE allreduce of the (otherwise zero) result arena makes it whole on every rank.
namespace
{
struct ChainStep
{
	std::vector<i64> da, db;
	i64 owner_dim; // position of the carried leg in this step's output
};

// Second-level split of the carried leg ("split the largest sectors' GEMMs along M"): a view of `t`
// whose sections along `dim` larger than 2 * piece are cut into pieces of `piece` rows (same charge). Metadata only:
// every block becomes several blocks that address row ranges of the same memory. parent / start give, for every refined
// section, the section it came from and its first row inside it.
struct Refined
{
	std::unique_ptr<Tensor> t;
	std::vector<i64> parent, start;
};
Refined refine_leg(const Tensor &t, i64 dim, i64 piece)
{
	Refined out;
	const i64 r = t.st.rank, nc = t.st.ct.nc;
	std::vector<i64> first_sub(t.st.nsec[dim] + 1, 0);
	std::vector<i64> sizes;
	for (i64 s = 0; s < t.st.nsec[dim]; ++s)
	{
		const i64 n = t.st.size_of(dim, s);
		first_sub[s] = (i64)sizes.size();
		if (n >= 2 * piece)
			for (i64 r0 = 0; r0 < n; r0 += piece)
			{
				out.parent.push_back(s);
				out.start.push_back(r0);
				sizes.push_back(std::min(piece, n - r0));
			}
		else
		{
			out.parent.push_back(s);
			out.start.push_back(0);
			sizes.push_back(n);
		}
	}
	first_sub[t.st.nsec[dim]] = (i64)sizes.size();
	auto v = std::make_unique<Tensor>();
	v->st.rank = r;
	v->st.ct = t.st.ct;
	v->st.sel = t.st.sel;
	for (i64 d = 0; d < r; ++d)
	{
		if (d != dim)
		{
			v->st.nsec.push_back(t.st.nsec[d]);
			for (i64 s = 0; s < t.st.nsec[d]; ++s)
			{
				v->st.sec_sizes.push_back(t.st.size_of(d, s));
				const i64 *c = t.st.charge_of(d, s);
				v->st.cvals.insert(v->st.cvals.end(), c, c + nc);
			}
		}
		else
		{
			v->st.nsec.push_back((i64)sizes.size());
			for (size_t k = 0; k < sizes.size(); ++k)
			{
				v->st.sec_sizes.push_back(sizes[k]);
				const i64 *c = t.st.charge_of(d, out.parent[k]);
				v->st.cvals.insert(v->st.cvals.end(), c, c + nc);
			}
		}
	}
	v->st.finalize();
	// blocks: lexicographic order is preserved when the pieces of a block are emitted in place only if `dim` is the last
	// differing position; in general re-sort
	std::vector<i64> index, dims, strides, offs;
	for (i64 b = 0; b < t.nblocks; ++b)
	{
		const i64 s = t.idx(b)[dim];
		for (i64 k = first_sub[s]; k < first_sub[s + 1]; ++k)
		{
			for (i64 d = 0; d < r; ++d)
			{
				index.push_back(d == dim ? k : t.idx(b)[d]);
				dims.push_back(d == dim ? sizes[k] : t.dm(b)[d]);
				strides.push_back(t.sd(b)[d]);
			}
			offs.push_back(t.offs[b] + out.start[k] * t.sd(b)[dim]);
		}
	}
	const i64 nb = (i64)offs.size();
	std::vector<i64> sorted_index = index;
	const auto order = r > 0 ? sort_blocks(r, sorted_index) : std::vector<i64>(nb, 0);
	v->nblocks = nb;
	v->index = sorted_index;
	v->dims.resize(nb * r);
	v->strides.resize(nb * r);
	v->offs.resize(nb);
	for (i64 k = 0; k < nb; ++k)
	{
		const i64 o = order[k];
		std::copy(dims.begin() + o * r, dims.begin() + (o + 1) * r, v->dims.begin() + k * r);
		std::copy(strides.begin() + o * r, strides.begin() + (o + 1) * r, v->strides.begin() + k * r);
		v->offs[k] = offs[o];
	}
	v->arena = t.arena;
	v->compute_hash();
	out.t = std::move(v);
	return out;
}

constexpr i64 kShardPiece = 128; // rows of a piece of the carried leg: the row extent of the large GEMM tile

std::unique_ptr<Tensor> run_chain(Ctx &ctx, const Tensor &first, const Tensor *const rhs[3], const ChainStep steps[3])
{
	if (ctx.world <= 1)
	{
		auto t1 = tensordot(ctx, first, *rhs[0], steps[0].da, steps[0].db);
		auto t2 = tensordot(ctx, *t1, *rhs[1], steps[1].da, steps[1].db);
		return tensordot(ctx, *t2, *rhs[2], steps[2].da, steps[2].db);
	}
	// where does the carried leg enter? output dims of a contraction = free dims of A (ascending), then free dims of B
	const i64 ra = first.st.rank;
	std::vector<char> ca(ra, 0), cb(rhs[0]->st.rank, 0);
	for (auto d : steps[0].da)
		ca[d < 0 ? d + ra : d] = 1;
	for (auto d : steps[0].db)
		cb[d < 0 ? d + rhs[0]->st.rank : d] = 1;
	std::vector<std::pair<int, i64>> outdims; // (operand, dim)
	for (i64 d = 0; d < ra; ++d)
		if (!ca[d])
			outdims.push_back({0, d});
	for (i64 d = 0; d < rhs[0]->st.rank; ++d)
		if (!cb[d])
			outdims.push_back({1, d});
	const auto src = outdims[steps[0].owner_dim];
	static const bool msplit = !(std::getenv("QTB_SHARD_MSPLIT") && std::atoi(std::getenv("QTB_SHARD_MSPLIT")) == 0);
	const bool refine = msplit && steps[2].owner_dim == 0;
	Refined ref;
	const Tensor *A0 = &first, *B0 = rhs[0];
	if (refine)
	{
		ref = refine_leg(src.first == 0 ? first : *rhs[0], src.second, kShardPiece);
		(src.first == 0 ? A0 : B0) = ref.t.get();
	}
	std::shared_ptr<Plan> plans[3];
	plans[0] = get_plan(ctx, *A0, *B0, steps[0].da, steps[0].db);
	plans[1] = get_plan(ctx, plans[0]->out_proto, *rhs[1], steps[1].da, steps[1].db);
	plans[2] = get_plan(ctx, plans[1]->out_proto, *rhs[2], steps[2].da, steps[2].db);
	std::vector<double> w;
	for (int i = 0; i < 3; ++i)
		add_section_weights(*plans[i], steps[i].owner_dim, w);
	const auto owner = lpt_assign(w, ctx.world);
	// the carried leg keeps its sector through the chain: a rank only ever reads the intermediate blocks it wrote itself
	auto t1 = tensordot_owned(ctx, plans[0], *A0, *B0, steps[0].owner_dim, owner, false);
	auto t2 = tensordot_owned(ctx, plans[1], *t1, *rhs[1], steps[1].owner_dim, owner, false);
	t1.reset();
	if (!refine)
	{
		auto t3 = tensordot_owned(ctx, plans[2], *t2, *rhs[2], steps[2].owner_dim, owner);
		t2.reset();
		ctx.allreduce(t3->arena->ptr, t3->arena->numel);
		return t3;
	}
	// The result keeps the UNREFINED structure (bit-exact with the reference): every refined output block is a contiguous
	// row slab of the block it refines (the carried leg is the slowest dim of the result), so the last step writes its
	// tiles straight into the unrefined packed layout.
	std::shared_ptr<Plan> up[3];
	up[0] = get_plan(ctx, first, *rhs[0], steps[0].da, steps[0].db);
	up[1] = get_plan(ctx, up[0]->out_proto, *rhs[1], steps[1].da, steps[1].db);
	up[2] = get_plan(ctx, up[1]->out_proto, *rhs[2], steps[2].da, steps[2].db);
	const Tensor &fin = up[2]->out_proto, &rfin = plans[2]->out_proto;
	std::vector<i64> c_off(rfin.nblocks), fin_off(rfin.nblocks);
	std::vector<i64> uidx(fin.st.rank);
	for (i64 b = 0; b < rfin.nblocks; ++b)
	{
		const i64 k = rfin.idx(b)[0];
		std::copy(rfin.idx(b), rfin.idx(b) + rfin.st.rank, uidx.begin());
		uidx[0] = ref.parent[k];
		const i64 ub = fin.find_block(uidx.data());
		QTB_REQUIRE(ub >= 0, QTB_ERR_RUNTIME, "sharded chain: a refined output block has no parent block");
		const i64 row = fin.block_numel(ub) / std::max<i64>(fin.dm(ub)[0], 1);
		fin_off[b] = fin.offs[ub] + ref.start[k] * row;
	}
	auto out = std::make_unique<Tensor>(fin);
	out->arena = std::make_shared<Arena>(&ctx, up[2]->out_numel);
	bool any_empty = false;
	for (auto &o : plans[2]->outs)
		any_empty |= (o.pair_begin == o.pair_end);
	const bool gather = ctx.nccl_comm != nullptr && !any_empty &&
	                    !(std::getenv("QTB_SHARD_ALLREDUCE") && std::atoi(std::getenv("QTB_SHARD_ALLREDUCE")) == 1);
	if (gather)
	{ // Every rank writes its row slabs back to back into its chunk of a staging buffer (the last GEMM's output offsets are
	  // remapped), ONE in-place ncclAllGather makes the staging buffer whole on every rank — exactly the owned bytes cross
	  // NVLink — and a copy kernel puts the slabs at their place in the result's packed layout.
		std::vector<i64> fill(ctx.world, 0), local(rfin.nblocks);
		for (i64 b = 0; b < rfin.nblocks; ++b)
		{
			const int o = owner[rfin.idx(b)[0]];
			local[b] = fill[o];
			fill[o] += rfin.block_numel(b);
		}
		const i64 chunk = (*std::max_element(fill.begin(), fill.end()) + 1) & ~i64(1);
		double *stage = (double *)ctx_alloc(ctx, (size_t)std::max<i64>(chunk * ctx.world, 1) * sizeof(double));
		std::vector<GatherDesc> gd;
		for (i64 b = 0; b < rfin.nblocks; ++b)
		{
			const int o = owner[rfin.idx(b)[0]];
			c_off[b] = (i64)o * chunk + local[b];
			GatherDesc g{};
			g.src_off = c_off[b];
			g.dst_off = fin_off[b];
			g.numel = rfin.block_numel(b);
			g.rank = 1;
			g.dims[0] = g.numel;
			g.strides[0] = 1;
			if (g.numel > 0)
				gd.push_back(g);
		}
		tensordot_owned_into(ctx, plans[2], *t2, *rhs[2], steps[2].owner_dim, owner, c_off, stage);
		t2.reset();
		if (chunk > 0)
			nccl_allgather(ctx, stage, chunk);
		launch_gather(ctx, gd, stage, out->arena->ptr);
		ctx_free(ctx, stage);
		ctx.counters[0] += 1;
		return out;
	}
	if (up[2]->out_numel)
		QTB_CUDA(cudaMemsetAsync(out->arena->ptr, 0, up[2]->out_numel * sizeof(double), ctx.stream));
	tensordot_owned_into(ctx, plans[2], *t2, *rhs[2], steps[2].owner_dim, owner, fin_off, out->arena->ptr);
	t2.reset();
	ctx.allreduce(out->arena->ptr, out->arena->numel);
	return out;
}
} // namespace

std::unique_ptr<Tensor> heff_apply(Ctx &ctx, const Tensor &psi, const Tensor &h2, const Tensor &lenv,
                                   const Tensor &renv)
{ // reference hamil2site_times_state_impl, dmrg.cpp:520-531. Carried leg: the bra bond a' of the left environment
  // (t1 = [w,a',s1,s2,b], t2 = [a',b,s1',s2',w'], out = [a',s1',s2',b']).
	const Tensor *rhs[3] = {&psi, &h2, &renv};
	const ChainStep steps[3] = {{{0}, {0}, 1}, {{0, 2, 3}, {0, 4, 5}, 0}, {{1, 4}, {0, 1}, 0}};
	return run_chain(ctx, lenv, rhs, steps);
}
std::unique_ptr<Tensor> env_left(Ctx &ctx, const Tensor &h, const Tensor &mps, const Tensor &lenv)
{ // reference compute_left_env_impl, dmrg.cpp:424-459. Carried leg: the new ket bond b of the site tensor
  // (t1 = [w,a',s,b], t2 = [a',b,s',w'], out = [b,w',b']).
	auto mc = conj(mps);
	const Tensor *rhs[3] = {&mps, &h, mc.get()};
	const ChainStep steps[3] = {{{0}, {0}, 3}, {{0, 2}, {0, 3}, 1}, {{0, 2}, {0, 1}, 0}};
	return run_chain(ctx, lenv, rhs, steps);
}
std::unique_ptr<Tensor> env_right(Ctx &ctx, const Tensor &h, const Tensor &mps, const Tensor &renv)
{ // reference compute_right_env_impl, dmrg.cpp:468-493. Carried leg: the new ket bond a of the site tensor
  // (t1 = [w,b',a,s], t2 = [b',a,w',s'], out = [a,w',a']).
	auto mc = conj(mps);
	const Tensor *rhs[3] = {&mps, &h, mc.get()};
	const ChainStep steps[3] = {{{0}, {2}, 2}, {{0, 3}, {2, 3}, 1}, {{3, 0}, {1, 2}, 0}};
	return run_chain(ctx, renv, rhs, steps);
}

std::unique_ptr<Tensor> tensordot_sharded(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &da,
                                          const std::vector<i64> &db, i64 owner_dim)
{ // one contraction, output blocks sharded by the sections of output dim `owner_dim`, then made whole by an allreduce
	if (ctx.world <= 1)
		return tensordot(ctx, a, b, da, db);
	auto plan = get_plan(ctx, a, b, da, db);
	std::vector<double> w;
	add_section_weights(*plan, owner_dim, w);
	const auto owner = lpt_assign(w, ctx.world);
	auto out = tensordot_owned(ctx, plan, a, b, owner_dim, owner);
	ctx.allreduce(out->arena->ptr, out->arena->numel);
	return out;
}

std::unique_ptr<Tensor> two_sites_update(Ctx &ctx, const Tensor &psi, const Tensor &h2, const Tensor &lenv,
                                         const Tensor &renv, double *energy)
{ // reference one_step_lanczos_impl + eig2x2Mat_impl + two_sites_update_impl, dmrg.cpp:543-651.
  // scal: [0]=a0 [1]=b [2]=a1 [3]=E0 [4]=o [5]=n [6]=nan flag [7]=guarded b
	double *scal = (double *)ctx_alloc(ctx, 8 * sizeof(double));
	auto phi = heff_apply(ctx, psi, h2, lenv, renv);
	auto psic = conj(psi);
	dot_dev(ctx, *phi, *psic, scal + 0, false);                                // a0 = <phi, psi>
	// psi_ip -= state * a0  ==  psi_ip.add_(state*a0, -1) (btensor.h:473): with the reference's merge behaviour
	auto phi2 = axpby_dev(ctx, nullptr, 1.0, *phi, scal + 0, -1.0, psi, false, -1.0);
	phi.reset();
	auto phi2c = conj(*phi2);
	dot_dev(ctx, *phi2, *phi2c, scal + 1, true); // b = sqrt(<phi,phi>)
	launch_guard_norm(ctx, scal + 1, scal + 7);  // divide only when b >= 1e-15 (dmrg.cpp:597-603)
	launch_scale(ctx, phi2->arena->ptr, phi2->arena->numel, scal + 7, 1.0, true);
	auto hphi = heff_apply(ctx, *phi2, h2, lenv, renv);
	dot_dev(ctx, *phi2c, *hphi, scal + 2, false); // a1 = <phi, H phi>   (phi2c aliases phi2's storage)
	launch_eig2x2(ctx, scal);
	auto out = axpby_dev(ctx, scal + 4, 1.0, psi, scal + 5, 1.0, *phi2, false); // o*psi + n*phi
	double h[8];
	QTB_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
	QTB_CUDA(cudaStreamSynchronize(ctx.stream));
	ctx.counters[5] += sizeof(h);
	ctx_free(ctx, scal);
	QTB_REQUIRE(h[6] == 0.0, QTB_ERR_LOGIC, "nan found in output tensor");
	if (energy)
		*energy = h[3];
	return out;
}

} // namespace qtb
