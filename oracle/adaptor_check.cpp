// TEST INFRASTRUCTURE ONLY — the reference-side binding of INTEGRATION.md, compiled for real.
//
// adaptor_check links the UNMODIFIED reference library (oracle/_ref/libquantit_ref.so) AND loads the engine
// (quantit_b200/libqtb.so) through its C ABI (include/qtb.h, dlopen: the reference side needs no CUDA toolchain). The
// namespace qtb_bind below is the thin C++ adaptor a QuantiT maintainer would add: quantit::btensor -> plain arrays ->
// qtb_tensor (to_engine), qtb_tensor -> quantit::btensor (from_engine), status codes -> the reference's exception types.
// Each command runs ONE reference entry point and the engine's replacement on the same btensor inputs IN THE SAME
// PROCESS and compares structure (bit-exact) and values:
//   adaptor_check tdot A.qtbt B.qtbt dimsA dimsB        btensor::tensordot            (btensor.h:623)
//   adaptor_check svdt A.qtbt split tol min max pow     svd(A, split, tol, min, max, pow) (LinearAlgebra.h:115), through U.d.V^T
//   adaptor_check heis L maxbond cutoff conv maxit      dmrg(bMPO&, bMPS&, options)   (dmrg.h:39): energies of both runs
// Exit code 0 = parity within the printed tolerance.
#include <dlfcn.h>

#include "ref_io.h"

#include "../include/qtb.h"

namespace qtb_bind
{
struct Api
{
	void *h = nullptr;
#define QTB_FN(name) decltype(&::name) name = nullptr;
	QTB_FN(qtb_last_error)
	QTB_FN(qtb_ctx_create)
	QTB_FN(qtb_ctx_destroy)
	QTB_FN(qtb_tensor_create)
	QTB_FN(qtb_tensor_free)
	QTB_FN(qtb_tensor_rank)
	QTB_FN(qtb_tensor_nc)
	QTB_FN(qtb_tensor_nblocks)
	QTB_FN(qtb_tensor_total_sections)
	QTB_FN(qtb_tensor_numel)
	QTB_FN(qtb_tensor_structure)
	QTB_FN(qtb_tensor_blocks)
	QTB_FN(qtb_tensor_download)
	QTB_FN(qtb_tensordot)
	QTB_FN(qtb_svd)
	QTB_FN(qtb_dmrg)
	QTB_FN(qtb_contract)
#undef QTB_FN
	explicit Api(const std::string &path)
	{
		h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
		if (!h)
			throw std::runtime_error(std::string("cannot load the engine: ") + dlerror());
#define QTB_LOAD(name)                                                                                                 \
	name = reinterpret_cast<decltype(name)>(dlsym(h, #name));                                                          \
	if (!name)                                                                                                         \
		throw std::runtime_error("libqtb.so does not export " #name);
		QTB_LOAD(qtb_last_error)
		QTB_LOAD(qtb_ctx_create)
		QTB_LOAD(qtb_ctx_destroy)
		QTB_LOAD(qtb_tensor_create)
		QTB_LOAD(qtb_tensor_free)
		QTB_LOAD(qtb_tensor_rank)
		QTB_LOAD(qtb_tensor_nc)
		QTB_LOAD(qtb_tensor_nblocks)
		QTB_LOAD(qtb_tensor_total_sections)
		QTB_LOAD(qtb_tensor_numel)
		QTB_LOAD(qtb_tensor_structure)
		QTB_LOAD(qtb_tensor_blocks)
		QTB_LOAD(qtb_tensor_download)
		QTB_LOAD(qtb_tensordot)
		QTB_LOAD(qtb_svd)
		QTB_LOAD(qtb_dmrg)
		QTB_LOAD(qtb_contract)
#undef QTB_LOAD
	}
};

// status code -> the exception type the reference throws for the same condition
inline void check(const Api &api, qtb_status st)
{
	if (st == QTB_OK)
		return;
	const std::string msg = api.qtb_last_error();
	switch (st)
	{
	case QTB_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
	case QTB_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);
	case QTB_ERR_LOGIC: throw std::logic_error(msg);
	case QTB_ERR_CHECK: TORCH_CHECK(false, msg); // c10::Error, what the reference raises from its own TORCH_CHECKs
	default: throw std::runtime_error(msg);
	}
}

// quantit::btensor -> engine tensor: structure through the public accessors only (dim, section_numbers, section_sizes,
// section_cqtts, selection_rule, begin/end), block data uploaded in ONE transfer
inline qtb_tensor *to_engine(const Api &api, qtb_ctx *ctx, const btensor &t)
{
	const i64 rank = t.dim();
	auto sel = charge_ints(t.selection_rule->get());
	const i64 nc = (i64)sel.size();
	std::vector<i64> mods(nc, 0); // Z charges (the harness' make_charge covers Z and ZxZ)
	std::vector<i64> nsec(t.section_numbers().begin(), t.section_numbers().end()), sizes, cvals;
	for (i64 d = 0; d < rank; ++d)
	{
		auto [b, e] = t.section_sizes(d);
		sizes.insert(sizes.end(), b, e);
		auto [cb, ce] = t.section_cqtts(d);
		for (auto it = cb; it != ce; ++it)
		{
			auto c = charge_ints(*it);
			cvals.insert(cvals.end(), c.begin(), c.end());
		}
	}
	std::vector<i64> index;
	std::vector<double> data;
	i64 nblocks = 0;
	for (auto &blk : t)
	{
		index.insert(index.end(), std::get<0>(blk).begin(), std::get<0>(blk).end());
		auto c = std::get<1>(blk).to(torch::kFloat64).contiguous();
		data.insert(data.end(), c.data_ptr<double>(), c.data_ptr<double>() + c.numel());
		++nblocks;
	}
	qtb_tensor *out = nullptr;
	check(api, api.qtb_tensor_create(ctx, rank, nc, mods.data(), nsec.data(), sizes.data(), cvals.data(), sel.data(), nblocks,
	                                 index.data(), data.data(), &out));
	return out;
}

// engine tensor -> quantit::btensor through the structure-only constructor + block(idx) = tensor (btensor.h:135, :243)
inline btensor from_engine(const Api &api, qtb_ctx *ctx, const qtb_tensor *t)
{
	const i64 rank = api.qtb_tensor_rank(t), nc = api.qtb_tensor_nc(t), nb = api.qtb_tensor_nblocks(t);
	const i64 tot = api.qtb_tensor_total_sections(t);
	std::vector<i64> nsec(rank), sizes(tot), cvals(tot * nc), sel(nc), mods(nc);
	check(api, api.qtb_tensor_structure(t, nsec.data(), sizes.data(), cvals.data(), sel.data(), mods.data()));
	btensor::vec_list_t spec(rank);
	i64 k = 0;
	for (i64 d = 0; d < rank; ++d)
		for (i64 s = 0; s < nsec[d]; ++s, ++k)
			spec[d].emplace_back(static_cast<size_t>(sizes[k]), make_charge(&cvals[k * nc], nc));
	btensor out(spec, make_charge(sel.data(), nc), torch::TensorOptions().dtype(torch::kFloat64));
	std::vector<i64> index(nb * rank), dims(nb * rank);
	check(api, api.qtb_tensor_blocks(t, index.data(), dims.data(), nullptr, nullptr));
	std::vector<double> data((size_t)std::max<i64>(api.qtb_tensor_numel(t), 1));
	check(api, api.qtb_tensor_download(ctx, t, data.data()));
	size_t pos = 0;
	for (i64 b = 0; b < nb; ++b)
	{
		std::vector<i64> bi(index.begin() + b * rank, index.begin() + (b + 1) * rank);
		std::vector<i64> bd(dims.begin() + b * rank, dims.begin() + (b + 1) * rank);
		auto blk = torch::empty(bd, torch::kFloat64);
		std::memcpy(blk.data_ptr<double>(), data.data() + pos, sizeof(double) * blk.numel());
		pos += blk.numel();
		out.block(bi) = blk;
	}
	return out;
}

// the replaced definitions
inline btensor tensordot(const Api &api, qtb_ctx *ctx, const btensor &a, const btensor &b, const std::vector<i64> &d1,
                         const std::vector<i64> &d2)
{
	qtb_tensor *A = to_engine(api, ctx, a), *B = to_engine(api, ctx, b), *C = nullptr;
	check(api, api.qtb_tensordot(ctx, A, B, (i64)d1.size(), d1.data(), d2.data(), &C));
	btensor out = from_engine(api, ctx, C);
	api.qtb_tensor_free(A);
	api.qtb_tensor_free(B);
	api.qtb_tensor_free(C);
	return out;
}
inline std::tuple<btensor, btensor, btensor> svd(const Api &api, qtb_ctx *ctx, const btensor &a, size_t split, double tol,
                                                 size_t mn, size_t mx, double pw)
{
	qtb_tensor *A = to_engine(api, ctx, a), *U = nullptr, *D = nullptr, *V = nullptr;
	check(api, api.qtb_svd(ctx, A, (i64)split, 1, tol, (i64)mn, mx == std::numeric_limits<size_t>::max() ? -1 : (i64)mx, pw, &U,
	                       &D, &V));
	auto out = std::make_tuple(from_engine(api, ctx, U), from_engine(api, ctx, D), from_engine(api, ctx, V));
	for (auto *t : {A, U, D, V})
		api.qtb_tensor_free(t);
	return out;
}
inline double dmrg(const Api &api, qtb_ctx *ctx, const bMPO &H, bMPS &psi, const dmrg_options &o)
{
	const i64 L = (i64)H.size();
	std::vector<qtb_tensor *> h(L), p(L);
	for (i64 i = 0; i < L; ++i)
	{
		h[i] = to_engine(api, ctx, H[i]);
		p[i] = to_engine(api, ctx, psi[i]);
	}
	qtb_dmrg_options opt{o.cutoff, o.convergence_criterion, (i64)std::min<size_t>(o.maximum_bond, (size_t)1 << 40),
	                     (i64)o.minimum_bond, (i64)o.maximum_iterations};
	double E = 0;
	i64 oc = (i64)(size_t)psi.orthogonality_center, nsw = 0;
	check(api, api.qtb_dmrg(ctx, L, h.data(), p.data(), &oc, &opt, &E, &nsw, nullptr, nullptr, nullptr));
	for (i64 i = 0; i < L; ++i)
		psi[i] = from_engine(api, ctx, p[i]);
	for (i64 i = 0; i < L; ++i)
	{
		api.qtb_tensor_free(h[i]);
		api.qtb_tensor_free(p[i]);
	}
	std::printf("ENGINE_SWEEPS %ld\n", (long)nsw);
	return E;
}
} // namespace qtb_bind

static bool same_structure_impl(const btensor &a, const btensor &b);
static bool same_structure(const btensor &a, const btensor &b)
{
	const bool ok = same_structure_impl(a, b);
	if (!ok)
	{
		auto brief = [](const btensor &t)
		{
			std::ostringstream o;
			o << "rank " << t.dim() << " sel " << fmt::format("{}", t.selection_rule->get()) << " nsec [";
			for (auto n : t.section_numbers())
				o << n << ",";
			o << "] sizes [";
			for (i64 d = 0; d < (i64)t.dim(); ++d)
			{
				auto [b, e] = t.section_sizes(d);
				for (auto it = b; it != e; ++it)
					o << *it << ",";
				o << "|";
			}
			o << "] charges [";
			for (i64 d = 0; d < (i64)t.dim(); ++d)
			{
				auto [b, e] = t.section_cqtts(d);
				for (auto it = b; it != e; ++it)
					o << fmt::format("{}", *it) << ",";
				o << "|";
			}
			o << "] blocks";
			for (auto &blk : t)
			{
				o << " (";
				for (auto i : std::get<0>(blk))
					o << i << ",";
				o << ":";
				for (auto s : std::get<1>(blk).sizes())
					o << s << "x";
				o << ")";
			}
			return o.str();
		};
		std::cerr << "STRUCTURE MISMATCH\n reference: " << brief(a) << "\n engine:    " << brief(b) << "\n";
	}
	return ok;
}
static bool same_structure_impl(const btensor &a, const btensor &b)
{
	if (a.dim() != b.dim() || a.section_numbers() != b.section_numbers())
		return false;
	for (i64 d = 0; d < (i64)a.dim(); ++d)
	{
		auto [ab, ae] = a.section_sizes(d);
		auto [bb, be] = b.section_sizes(d);
		if (!std::equal(ab, ae, bb))
			return false;
		auto [ac, ace] = a.section_cqtts(d);
		auto [bc, bce] = b.section_cqtts(d);
		for (auto x = ac, y = bc; x != ace; ++x, ++y)
			if (charge_ints(*x) != charge_ints(*y))
				return false;
	}
	if (charge_ints(a.selection_rule->get()) != charge_ints(b.selection_rule->get()))
		return false;
	auto ia = a.begin();
	auto ib = b.begin();
	for (; ia != a.end() && ib != b.end(); ++ia, ++ib)
		if (std::get<0>(*ia) != std::get<0>(*ib) || std::get<1>(*ia).sizes() != std::get<1>(*ib).sizes())
			return false;
	return ia == a.end() && ib == b.end();
}
static double max_rel_err(const btensor &a, const btensor &b)
{
	double num = 0, den = 0;
	auto ia = a.begin();
	auto ib = b.begin();
	for (; ia != a.end(); ++ia, ++ib)
	{
		if (std::get<1>(*ia).numel() == 0)
			continue;
		num = std::max(num, (std::get<1>(*ia) - std::get<1>(*ib)).abs().max().item().toDouble());
		den = std::max(den, std::get<1>(*ia).abs().max().item().toDouble());
	}
	return den > 0 ? num / den : num;
}

int main(int argc, char **argv)
{
	std::vector<std::string> a(argv + 1, argv + argc);
	std::string lib = "quantit_b200/libqtb.so";
	for (size_t i = 0; i + 1 < a.size(); ++i)
		if (a[i] == "--lib")
		{
			lib = a[i + 1];
			a.erase(a.begin() + i, a.begin() + i + 2);
			break;
		}
	if (a.empty())
	{
		std::puts("usage: adaptor_check [--lib path/to/libqtb.so] <tdot|svdt|heis> ...");
		return 2;
	}
	// the reference is an fp64 code run with torch's default dtype set to double (its own tests and examples do this;
	// compact_dense_single allocates with the DEFAULT dtype whatever the input's, btensor_linalg.cpp:134)
	torch::set_default_dtype(torch::scalarTypeToTypeMeta(torch::kFloat64));
	try
	{
		if (a[0] == "selfcheck")
		{ // CPU-only check of the comparison itself: reference svd -> QTBT dump -> load -> same_structure
			auto A = load(a[1]);
			auto [U, d, V] = quantit::svd(A, std::stoul(a[2]), std::stod(a[3]), std::stoull(a[4]), std::stoull(a[5]), std::stod(a[6]));
			int ok = 1;
			for (auto *t : {&U, &d, &V})
			{
				dump(*t, "/tmp/_selfcheck.qtbt");
				ok &= (int)same_structure(*t, load("/tmp/_selfcheck.qtbt"));
			}
			std::printf("SELFCHECK %d\n", ok);
			return ok ? 0 : 1;
		}
		qtb_bind::Api api(lib);
		qtb_ctx *ctx = nullptr;
		qtb_bind::check(api, api.qtb_ctx_create(0, nullptr, &ctx));
		int rc = 0;
		if (a[0] == "tdot")
		{
			auto A = load(a[1]), B = load(a[2]);
			auto dA = csv(a[3]), dB = csv(a[4]);
			auto ref = A.tensordot(B, dA, dB);
			auto got = qtb_bind::tensordot(api, ctx, A, B, dA, dB);
			const bool st = same_structure(ref, got);
			const double err = st ? max_rel_err(ref, got) : 1.0;
			std::printf("TDOT structure_identical %d max_rel_err %.3e\n", (int)st, err);
			rc = (st && err <= 1e-12) ? 0 : 1;
		}
		else if (a[0] == "svdt")
		{
			auto A = load(a[1]);
			size_t split = std::stoul(a[2]);
			double tol = std::stod(a[3]), pw = std::stod(a[6]);
			size_t mn = std::stoull(a[4]), mx = std::stoull(a[5]);
			auto [U, d, V] = quantit::svd(A, split, tol, mn, mx, pw);
			auto [U2, d2, V2] = qtb_bind::svd(api, ctx, A, split, tol, mn, mx, pw);
			const bool st = same_structure(U, U2) && same_structure(d, d2) && same_structure(V, V2);
			double err = 1.0;
			if (st)
			{ // singular values directly; the factors through the gauge-invariant product U.d.V^T, computed by the reference
				err = max_rel_err(d, d2);
				const auto r = (i64)U.dim() - 1;
				auto rec = [&](const btensor &u, const btensor &s, const btensor &v)
				{ return u.mul(s).tensordot(v.conj(), {r}, {(i64)v.dim() - 1}); };
				err = std::max(err, max_rel_err(rec(U, d, V), rec(U2, d2, V2)));
			}
			std::printf("SVD structure_identical %d max_rel_err %.3e\n", (int)st, err);
			rc = (st && err <= 1e-11) ? 0 : 1;
		}
		else if (a[0] == "heis")
		{
			size_t L = std::stoul(a[1]), maxbond = std::stoull(a[2]);
			double cutoff = std::stod(a[3]), conv = std::stod(a[4]);
			size_t maxit = std::stoul(a[5]);
			MPO heis = Heisenberg(torch::tensor(-1.0), L);
			auto phys = btensor({{{1, Z(1)}, {1, Z(-1)}}}, any_quantity(Z(0)));
			auto lb = btensor({{{1, Z(0)}, {1, Z(-2)}, {1, Z(2)}, {1, Z(0)}, {1, Z(0)}}}, any_quantity(Z(0)));
			bMPO H = to_bMPO(std::move(heis), shape_from(lb, phys, lb.conj(), phys.conj()));
			H.coalesce();
			torch::manual_seed(1234);
			bMPS psi = random_bMPS(4, H, any_quantity(Z(L % 2)), {}, 0);
			bMPS psi2 = psi;
			dmrg_options opt(cutoff, conv, maxbond, 4, maxit);
			const double Eref = quantit::dmrg(H, psi, opt).item().toDouble();
			const double Eeng = qtb_bind::dmrg(api, ctx, H, psi2, opt);
			// the engine's final state, read back into a reference bMPS, evaluated by the REFERENCE's contract()
			const double Echk = contract(psi2, psi2, H).item().toDouble() / contract(psi2, psi2).item().toDouble();
			std::printf("DMRG E_reference %.14f E_engine %.14f <psi_engine|H|psi_engine> by the reference %.14f\n", Eref, Eeng, Echk);
			rc = (std::fabs(Eeng - Eref) <= 1e-10 * std::fabs(Eref) && std::fabs(Echk - Eeng) <= 1e-10 * std::fabs(Eref)) ? 0 : 1;
		}
		else
		{
			std::fprintf(stderr, "unknown command %s\n", a[0].c_str());
			rc = 2;
		}
		api.qtb_ctx_destroy(ctx);
		std::puts(rc == 0 ? "ADAPTOR_OK" : "ADAPTOR_MISMATCH");
		return rc;
	}
	catch (const std::exception &e)
	{
		std::fprintf(stderr, "ADAPTOR_EXCEPTION %s\n", e.what());
		return 3;
	}
}
