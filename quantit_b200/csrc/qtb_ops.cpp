// qtb_ops.cpp — block-tensor vector ops and the two-site DMRG pieces, composed from the planner + kernels.
//
// Reference functions restated on the engine's data model (paths relative to the reference root):
//   btensor::add / add_            sources/btensor.cpp:2666-2752
//   btensor::mul_ (broadcast)      sources/btensor.cpp:1204-1302
//   hamil2site_times_state         sources/dmrg.cpp:520-531
//   compute_left_env/right_env     sources/dmrg.cpp:424-493
//   one_step_lanczos, eig2x2Mat, two_sites_update   sources/dmrg.cpp:543-651
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
std::unique_ptr<Tensor> run_chain(Ctx &ctx, const Tensor &first, const Tensor *const rhs[3], const ChainStep steps[3])
{
	if (ctx.world <= 1)
	{
		auto t1 = tensordot(ctx, first, *rhs[0], steps[0].da, steps[0].db);
		auto t2 = tensordot(ctx, *t1, *rhs[1], steps[1].da, steps[1].db);
		return tensordot(ctx, *t2, *rhs[2], steps[2].da, steps[2].db);
	}
	std::shared_ptr<Plan> plans[3];
	plans[0] = get_plan(ctx, first, *rhs[0], steps[0].da, steps[0].db);
	plans[1] = get_plan(ctx, plans[0]->out_proto, *rhs[1], steps[1].da, steps[1].db);
	plans[2] = get_plan(ctx, plans[1]->out_proto, *rhs[2], steps[2].da, steps[2].db);
	std::vector<double> w;
	for (int i = 0; i < 3; ++i)
		add_section_weights(*plans[i], steps[i].owner_dim, w);
	const auto owner = lpt_assign(w, ctx.world);
	auto t1 = tensordot_owned(ctx, plans[0], first, *rhs[0], steps[0].owner_dim, owner);
	auto t2 = tensordot_owned(ctx, plans[1], *t1, *rhs[1], steps[1].owner_dim, owner);
	t1.reset();
	auto t3 = tensordot_owned(ctx, plans[2], *t2, *rhs[2], steps[2].owner_dim, owner);
	t2.reset();
	ctx.allreduce(t3->arena->ptr, t3->arena->numel);
	return t3;
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
