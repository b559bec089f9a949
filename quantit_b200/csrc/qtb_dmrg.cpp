// qtb_dmrg.cpp — the two-site DMRG sweep driver on the engine's primitives.
//
// Reference (paths relative to the reference root): sources/dmrg.cpp
//   dmrg(bMPO&, bMPS&, options, logger)   :92-100      details::dmrg_impl            :219-273
//   generate_env_impl / trivial edges     :370-409     compute_2sitesHamil_impl      :503-515
//   sweep                                 :127-142     dmrg_2sites_update::operator() :163-206
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <map>

#include "qtb_ops.h"

namespace qtb
{

static std::unique_ptr<Tensor> trivial_edge(Ctx &ctx, const Tensor &state, i64 sdim, const Tensor &ham, i64 hdim)
{ // reference generate_env_impl, dmrg.cpp:381-391: ones over (inverse ket leg, inverse MPO leg, ket leg), neutral rule
	const i64 nc = state.st.ct.nc;
	Structure st;
	st.rank = 3;
	st.ct = state.st.ct;
	st.sel.assign(nc, 0);
	auto push = [&](const Tensor &t, i64 d, bool inv)
	{
		st.nsec.push_back(t.st.nsec[d]);
		for (i64 s = 0; s < t.st.nsec[d]; ++s)
		{
			st.sec_sizes.push_back(t.st.size_of(d, s));
			for (i64 c = 0; c < nc; ++c)
				st.cvals.push_back(inv ? st.ct.norm(-t.st.charge_of(d, s)[c], c) : t.st.charge_of(d, s)[c]);
		}
	};
	push(state, sdim, true);
	push(ham, hdim, true);
	push(state, sdim, false);
	st.finalize();
	std::vector<i64> index;
	std::vector<double> data;
	i64 idx[3];
	for (idx[0] = 0; idx[0] < st.nsec[0]; ++idx[0])
		for (idx[1] = 0; idx[1] < st.nsec[1]; ++idx[1])
			for (idx[2] = 0; idx[2] < st.nsec[2]; ++idx[2])
				if (st.allowed(idx))
				{
					index.insert(index.end(), idx, idx + 3);
					const i64 n = st.size_of(0, idx[0]) * st.size_of(1, idx[1]) * st.size_of(2, idx[2]);
					data.insert(data.end(), (size_t)n, 1.0);
				}
	return make_tensor(ctx, st, (i64)index.size() / 3, index.data(), data.data());
}

// ---------------------------------------------------------------------------------------------------------------------
// bMPS / bMPO chain contractions and gauge moves (reference sources/MPT.cpp)
// ---------------------------------------------------------------------------------------------------------------------
namespace
{
// eye_like(shape_from(A[dim_a]^-1, B[dim_b])) (x) ones_like(edge_shape_prep(H[dim_h])) permuted to [a^-1, w^-1, b]
// (reference contract(bMPS, bMPS, bMPO), MPT.cpp:218-225; eye_like btensor.cpp:2444-2458: identity blocks on the (i, i)
// section pairs the selection rule allows). Without an MPO the edge is the rank-2 eye itself (MPT.cpp:285-290).
std::unique_ptr<Tensor> eye_edge(Ctx &ctx, const Tensor &a, i64 dim_a, const Tensor &b, i64 dim_b, const Tensor *h, i64 dim_h)
{
	const i64 nc = a.st.ct.nc;
	Structure st;
	st.rank = h ? 3 : 2;
	st.ct = a.st.ct;
	st.sel.assign(nc, 0);
	auto push = [&](const Tensor &t, i64 d, bool inv)
	{
		st.nsec.push_back(t.st.nsec[d]);
		for (i64 s = 0; s < t.st.nsec[d]; ++s)
		{
			st.sec_sizes.push_back(t.st.size_of(d, s));
			for (i64 c = 0; c < nc; ++c)
				st.cvals.push_back(inv ? st.ct.norm(-t.st.charge_of(d, s)[c], c) : t.st.charge_of(d, s)[c]);
		}
	};
	push(a, dim_a, true);
	if (h)
		push(*h, dim_h, true);
	push(b, dim_b, false);
	st.finalize();
	std::vector<i64> index;
	std::vector<double> data;
	const i64 n = std::min(a.st.nsec[dim_a], b.st.nsec[dim_b]);
	const i64 nw = h ? h->st.nsec[dim_h] : 1;
	for (i64 i = 0; i < n; ++i)
		for (i64 j = 0; j < nw; ++j)
		{
			i64 idx[3];
			idx[0] = i;
			if (h)
			{
				idx[1] = j;
				idx[2] = i;
			}
			else
				idx[1] = i;
			if (!st.allowed(idx))
				continue;
			const i64 ra = a.st.size_of(dim_a, i), rb = b.st.size_of(dim_b, i), w = h ? h->st.size_of(dim_h, j) : 1;
			index.insert(index.end(), idx, idx + st.rank);
			for (i64 x = 0; x < ra; ++x)
				for (i64 y = 0; y < w; ++y)
					for (i64 z = 0; z < rb; ++z)
						data.push_back(x == z ? 1.0 : 0.0);
		}
	return make_tensor(ctx, st, (i64)index.size() / st.rank, index.data(), data.data());
}

double scalar_of(Ctx &ctx, const Tensor &t)
{ // a rank-0 result holds at most the single block {}
	QTB_REQUIRE(t.st.rank == 0, QTB_ERR_RUNTIME, "contract: the chain did not close to a scalar");
	if (t.nblocks == 0 || !t.arena || t.arena->numel == 0)
		return 0.0;
	double v = 0.0;
	download(ctx, t, &v);
	return v;
}
} // namespace

double contract(Ctx &ctx, i64 L, const Tensor *const *a, const Tensor *const *b, const Tensor *const *obs)
{ // reference contract(bMPS, bMPS, bMPO) MPT.cpp:211-233 and contract(bMPS, bMPS) MPT.cpp:275-292
	QTB_REQUIRE(L >= 1, QTB_ERR_INVALID_ARGUMENT, "contract: empty chain");
	for (i64 i = 0; i < L; ++i)
	{
		QTB_REQUIRE(a[i]->st.rank == 3 && b[i]->st.rank == 3, QTB_ERR_INVALID_ARGUMENT,
		            "Input bMPT is an invalid bMPS: one or more Tensors has rank differing from 3");
		QTB_REQUIRE(!obs || obs[i]->st.rank == 4, QTB_ERR_INVALID_ARGUMENT, "contract: MPO tensors must have rank 4");
	}
	auto left = eye_edge(ctx, *a[0], 0, *b[0], 0, obs ? obs[0] : nullptr, 0);
	auto right = eye_edge(ctx, *a[L - 1], 2, *b[L - 1], 2, obs ? obs[L - 1] : nullptr, 2);
	for (i64 i = 0; i < L; ++i)
	{
		auto bc = conj(*b[i]);
		if (obs)
		{
			auto t1 = tensordot(ctx, *left, *a[i], {0}, {0});     // [w, b', s, a]
			auto t2 = tensordot(ctx, *t1, *obs[i], {0, 2}, {0, 3}); // [b', a, s', w']
			left = tensordot(ctx, *t2, *bc, {0, 2}, {0, 1});      // [a, w', b'']
		}
		else
		{
			auto t1 = tensordot(ctx, *left, *a[i], {0}, {0}); // [b', s, a]
			left = tensordot(ctx, *t1, *bc, {0, 1}, {0, 1});  // [a, b'']
		}
	}
	std::unique_ptr<Tensor> res =
	    obs ? tensordot(ctx, *left, *right, {0, 1, 2}, {0, 1, 2}) : tensordot(ctx, *left, *right, {0, 1}, {0, 1});
	return scalar_of(ctx, *res);
}

void move_oc(Ctx &ctx, std::vector<std::unique_ptr<Tensor>> &mps, i64 &oc, i64 target)
{ // reference bMPS::move_oc, MPT.cpp:75-111
	const i64 L = (i64)mps.size();
	QTB_REQUIRE(target >= 0 && target < L, QTB_ERR_INVALID_ARGUMENT,
	            " Proposed orthogonality center falls outside the MPS");
	QTB_REQUIRE(oc >= 0 && oc < L, QTB_ERR_INVALID_ARGUMENT,
	            "orthogonality center position greater than the number of defined tensors.");
	while (target < oc)
	{ // svd(curr, 1): curr = (v*) permuted {2,0,1}; previous site absorbs u.d
		std::unique_ptr<Tensor> u, d, v;
		block_svd(ctx, *mps[oc], 1, false, 0.0, 0, -1, 2.0, u, d, v);
		auto vc = conj(*v);
		mps[oc] = permute(*vc, {2, 0, 1});
		auto ud = mul_lastdim(ctx, *u, *d);
		mps[oc - 1] = tensordot(ctx, *mps[oc - 1], *ud, {2}, {0});
		--oc;
	}
	while (target > oc)
	{ // svd(curr, 2): curr = u; next site absorbs (v.d)*
		std::unique_ptr<Tensor> u, d, v;
		block_svd(ctx, *mps[oc], 2, false, 0.0, 0, -1, 2.0, u, d, v);
		auto vd = mul_lastdim(ctx, *v, *d);
		auto dv = conj(*vd);
		mps[oc] = std::move(u);
		mps[oc + 1] = tensordot(ctx, *dv, *mps[oc + 1], {0}, {0});
		++oc;
	}
}

void coalesce(Ctx &ctx, std::vector<std::unique_ptr<Tensor>> &mpo, double cutoff)
{ // reference bMPO::coalesce, MPT.cpp:154-168: left-to-right sweep of truncated block SVDs over the MPO bonds
  // (svd(tens, 3, cutoff): min_size 1, max_size unlimited, pow 2 — btensor_linalg.cpp:811-816)
	const i64 L = (i64)mpo.size();
	for (i64 i = 0; i < L; ++i)
		QTB_REQUIRE(mpo[i]->st.rank == 4, QTB_ERR_INVALID_ARGUMENT, "coalesce: MPO tensors must have rank 4");
	for (i64 i = 0; i + 1 < L; ++i)
	{
		auto tens = permute(*mpo[i], {0, 1, 3, 2});
		std::unique_ptr<Tensor> u, d, v;
		block_svd(ctx, *tens, 3, true, cutoff, 1, -1, 2.0, u, d, v);
		auto ud = mul_lastdim(ctx, *u, *d); // U.mul_(d)
		auto vc = conj(*v);
		mpo[i + 1] = tensordot(ctx, *vc, *mpo[i + 1], {0}, {0});
		mpo[i] = permute(*ud, {0, 1, 3, 2});
	}
}

void dmrg(Ctx &ctx, i64 L, const Tensor *const *mpo, std::vector<std::unique_ptr<Tensor>> &mps, i64 &oc,
          const qtb_dmrg_options &opt, double &energy, i64 &n_sweeps, double *sweep_energy, double *sweep_seconds,
          i64 *sweep_mid_bond, qtb_dmrg_log_fn log_fn, void *log_user)
{
	QTB_REQUIRE(L >= 2, QTB_ERR_INVALID_ARGUMENT, "dmrg: at least two sites are required");
	for (i64 i = 0; i < L; ++i)
	{
		QTB_REQUIRE(mpo[i]->st.rank == 4, QTB_ERR_INVALID_ARGUMENT, "dmrg: MPO tensors must have rank 4");
		QTB_REQUIRE(mps[i]->st.rank == 3, QTB_ERR_INVALID_ARGUMENT,
		            "Input bMPT is an invalid bMPS: one or more Tensors has rank differing from 3");
	}
	QTB_REQUIRE(oc >= 0 && oc < L, QTB_ERR_INVALID_ARGUMENT,
	            "orthogonality center position greater than the number of defined tensors.");
	// environments Env[-1 .. L]
	std::map<i64, std::unique_ptr<Tensor>> env;
	env[-1] = trivial_edge(ctx, *mps[0], 0, *mpo[0], 0);
	env[L] = trivial_edge(ctx, *mps[L - 1], 2, *mpo[L - 1], 2);
	for (i64 i = 0; i < oc; ++i)
		env[i] = env_left(ctx, *mpo[i], *mps[i], *env[i - 1]);
	for (i64 i = L - 1; i > oc; --i)
		env[i] = env_right(ctx, *mpo[i], *mps[i], *env[i + 1]);
	// two-site MPO, dmrg.cpp:503-515
	std::vector<std::unique_ptr<Tensor>> h2(L - 1);
	for (i64 i = 0; i + 1 < L; ++i)
	{
		auto t = tensordot(ctx, *mpo[i], *mpo[i + 1], {2}, {0});
		h2[i] = permute(*t, {0, 1, 3, 4, 2, 5});
	}
	double E0 = 100000.0;
	const i64 nh = L - 1;
	const i64 n_step = nh - 1 + (nh == 1 ? 1 : 0);
	int step = (oc == 0) ? 1 : -1;
	if (nh == 1)
		step = 0;
	// a centre on the last site: the reference steps it back by one without regauging (dmrg.cpp:229-233) — the two-site
	// tensor of sites (L-2, L-1) is the centre either way, the left environments up to L-3 and the right edge are in place
	if (oc == L - 1)
		--oc;
	const i64 init_pos = oc;
	double *d_scal = (double *)ctx_alloc(ctx, 2 * sizeof(double));
	n_sweeps = 0;
	// QTB_PROFILE=1: per-phase wall time with a stream sync after every phase (diagnostics only, perturbs the timing)
	const bool prof = std::getenv("QTB_PROFILE") != nullptr;
	ctx.prof_level = prof ? std::atoi(std::getenv("QTB_PROFILE")) : 0;
	double tph[5] = {0, 0, 0, 0, 0};
	auto tick = [&](int ph, std::chrono::steady_clock::time_point &t)
	{
		if (!prof)
			return;
		cudaStreamSynchronize(ctx.stream);
		auto now = std::chrono::steady_clock::now();
		tph[ph] += std::chrono::duration<double, std::milli>(now - t).count();
		t = now;
	};
	for (i64 it = 0; it < opt.maximum_iterations; ++it)
	{
		auto t0 = std::chrono::steady_clock::now();
		double E = E0;
		for (i64 s = 0; s < 2 * n_step; ++s)
		{
			// ---- dmrg_2sites_update::operator(), dmrg.cpp:163-206 ----
			auto tp = std::chrono::steady_clock::now();
			auto theta = tensordot_sharded(ctx, *mps[oc], *mps[oc + 1], {2}, {0}, 0);
			tick(0, tp);
			auto theta2 = two_sites_update(ctx, *theta, *h2[oc], *env[oc - 1], *env[oc + 2], &E);
			theta.reset();
			tick(1, tp);
			std::unique_ptr<Tensor> u, d, v;
			block_svd(ctx, *theta2, 2, true, opt.cutoff, opt.minimum_bond, opt.maximum_bond, 2.0, u, d, v);
			theta2.reset();
			tick(2, tp);
			// d /= sqrt(sum(d^2))
			{
				auto dc = conj(*d);
				dot_dev(ctx, *d, *dc, d_scal, true);
				if (d->arena->numel > 0)
					launch_scale(ctx, d->arena->ptr, d->arena->numel, d_scal, 1.0, true);
			}
			if (step == 1)
			{
				auto vd = mul_lastdim(ctx, *v, *d);
				auto vdc = conj(*vd);
				mps[oc] = std::move(u);
				mps[oc + 1] = permute(*vdc, {2, 0, 1});
				env[oc] = env_left(ctx, *mpo[oc], *mps[oc], *env[oc - 1]);
			}
			else
			{
				auto vc = conj(*v);
				mps[oc] = mul_lastdim(ctx, *u, *d);
				mps[oc + 1] = permute(*vc, {2, 0, 1});
				env[oc + 1] = env_right(ctx, *mpo[oc + 1], *mps[oc + 1], *env[oc + 2]);
			}
			tick(3, tp);
			oc += step;
			if (oc == 0 || oc == L - 2)
				step = -step;
		}
		if (prof)
		{
			std::fprintf(stderr, "[qtb profile] sweep %ld: theta %.1f ms, lanczos %.1f ms, svd %.1f ms, absorb+env %.1f ms; plans built %ld, cache hits %ld, launches %ld\n",
			             (long)it, tph[0], tph[1], tph[2], tph[3], (long)ctx.counters[2], (long)ctx.counters[3], (long)ctx.counters[0]);
			tph[0] = tph[1] = tph[2] = tph[3] = 0;
			if (ctx.prof_level >= 2)
				ctx.prof_dump("contractions of this sweep (synchronised per call)");
		}
		QTB_CUDA(cudaStreamSynchronize(ctx.stream));
		const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		if (sweep_energy)
			sweep_energy[it] = E;
		if (sweep_seconds)
			sweep_seconds[it] = secs;
		if (sweep_mid_bond)
			sweep_mid_bond[it] = mps[L / 2]->st.dim_size(0);
		n_sweeps = it + 1;
		if (log_fn) // dmrg_logger::it_log_all (dmrg.cpp:243): energy, bond dimensions, wall time of the sweep
		{
			std::vector<i64> bonds(L + 1);
			for (i64 i = 0; i < L; ++i)
				bonds[i] = mps[i]->st.dim_size(0);
			bonds[L] = mps[L - 1]->st.dim_size(2);
			log_fn(log_user, it, E, secs, bonds.data(), L + 1);
		}
		const double Eold = E0;
		E0 = E;
		if (!(std::fabs((E0 - Eold) / E0) > opt.convergence_criterion)) // stops on NaN too (dmrg.cpp:247-254)
			break;
	}
	ctx_free(ctx, d_scal);
	QTB_REQUIRE(oc == init_pos, QTB_ERR_RUNTIME, "the orthogonality center finished somewhere surprising!");
	energy = E0;
}

} // namespace qtb
