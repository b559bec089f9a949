// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// ref_harness: our own command-line driver around the UNMODIFIED reference library (AlexandreFoley/QuantiT,
// compiled from /root/reference/sources by oracle/Makefile into oracle/_ref/libquantit_ref.so). It reads block
// tensors in the QTBT dump format (see oracle/qtbt_format.md / oracle/qtb_oracle.py), runs ONE reference entry point
// on them and writes the result back as QTBT. It is used (a) to pin the numpy restatement oracle/qtb_oracle.py and
// to generate tests/golden/*, (b) as the "reference" CPU baseline timed by bench.py.
//
// Reference entry points exercised (all through the public C++ API):
//   btensor::tensordot                include/blockTensor/btensor.h:623   sources/btensor.cpp:1971
//   btensor::permute / conj           sources/btensor.cpp:1754 / :2156
//   svd(btensor, split[, tol,min,max,pow])   include/blockTensor/LinearAlgebra.h:87,115  sources/btensor_linalg.cpp:503,805
//   details::hamil2site_times_state   include/dmrg.h:61   sources/dmrg.cpp:520
//   compute_left_env/right_env        sources/dmrg.cpp:424-493 (defined, not declared in a header: declared below)
//   two_sites_update                  sources/dmrg.cpp:623-651 (same)
//   dmrg(bMPO&, bMPS&, options)       include/dmrg.h:39    sources/dmrg.cpp:92
#include "ref_io.h"

struct sweep_printer : public dmrg_logger
{
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void log_step(size_t) override {}
	void log_energy(const torch::Tensor &) override {}
	void log_energy(const btensor &) override {}
	void log_bond_dims(const MPS &) override {}
	void log_bond_dims(const bMPS &) override {}
	void it_log_all(size_t it, const btensor &E, const bMPS &state) override
	{
		auto t1 = std::chrono::steady_clock::now();
		double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		t0 = t1;
		i64 mid = state[state.size() / 2].sizes()[0];
		std::printf("SWEEP %zu E %.14f mid_bond %ld ms %.3f\n", it, E.item().toDouble(), (long)mid, ms);
		std::fflush(stdout);
	}
	void end_log_all(size_t, const btensor &, const bMPS &) override {}
};

static bMPO heisenberg_u1(size_t L)
{ // the reference's Heisenberg(J,L,shape) assigns wrong bond charges (sources/models.cpp:96);
  // go through its own to_bMPO (sources/models.cpp:70) with the conserving assignment {0, p1-p0, p0-p1, 0, 0}.
	MPO heis = Heisenberg(torch::tensor(-1.0), L);
	auto phys = btensor({{{1, Z(1)}, {1, Z(-1)}}}, any_quantity(Z(0)));
	auto lb = btensor({{{1, Z(0)}, {1, Z(-2)}, {1, Z(2)}, {1, Z(0)}, {1, Z(0)}}}, any_quantity(Z(0)));
	bMPO H = to_bMPO(std::move(heis), shape_from(lb, phys, lb.conj(), phys.conj()));
	H.coalesce();
	return H;
}
static bMPO hubbard_u1u1(size_t L, double U, double mu)
{ // python_binding/exemples/dmrg.py:41-50 ; sources/models.cpp:164-172
	auto phys = btensor({{{1, ZZ(0, 0)}, {1, ZZ(1, 1)}, {1, ZZ(1, -1)}, {1, ZZ(2, 0)}}}, any_quantity(ZZ(0, 0)));
	return Hubbard(torch::tensor(U), torch::tensor(mu), L, phys);
}

template <class F>
static void timed(int reps, F &&f)
{
	for (int r = 0; r < reps; ++r)
	{
		auto t0 = std::chrono::steady_clock::now();
		f();
		auto t1 = std::chrono::steady_clock::now();
		std::printf("TIME_MS %.6f\n", std::chrono::duration<double, std::milli>(t1 - t0).count());
	}
	std::fflush(stdout);
}

int main(int argc, char **argv)
{
	torch::set_default_dtype(torch::scalarTypeToTypeMeta(torch::kFloat64));
	torch::InferenceMode guard;
	std::vector<std::string> a(argv + 1, argv + argc);
	int threads = 1;
	int reps = 0;
	// trailing "--threads N" "--reps R"
	for (size_t i = 0; i + 1 < a.size();)
	{
		if (a[i] == "--threads")
		{
			threads = std::stoi(a[i + 1]);
			a.erase(a.begin() + i, a.begin() + i + 2);
		}
		else if (a[i] == "--reps")
		{
			reps = std::stoi(a[i + 1]);
			a.erase(a.begin() + i, a.begin() + i + 2);
		}
		else
			++i;
	}
	torch::set_num_threads(threads);
	at::init_num_threads();
	if (a.empty())
	{
		std::puts("usage: ref_harness <tdot|permute|conj|svd|svdt|heff|lenv|renv|update|mul|heis|hub|moveoc|coalesce|eigh> ...");
		return 2;
	}
	try
	{
		const auto &cmd = a[0];
		if (cmd == "tdot")
		{ // tdot A B dimsA dimsB OUT
			auto A = load(a[1]), B = load(a[2]);
			auto dA = csv(a[3]), dB = csv(a[4]);
			auto C = A.tensordot(B, dA, dB);
			dump(C, a[5]);
			timed(reps, [&]() { auto X = A.tensordot(B, dA, dB); });
		}
		else if (cmd == "permute")
		{
			auto A = load(a[1]);
			dump(A.permute(csv(a[2])), a[3]);
		}
		else if (cmd == "conj")
		{
			auto A = load(a[1]);
			dump(A.conj(), a[2]);
		}
		else if (cmd == "reshape")
		{ // reshape A "i,j" OUT : index groups are split at the listed positions (btensor.cpp:2986)
			auto A = load(a[1]);
			dump(A.reshape(csv(a[2])), a[3]);
		}
		else if (cmd == "svd")
		{ // svd A split OUT_U OUT_d OUT_V
			auto A = load(a[1]);
			size_t split = std::stoul(a[2]);
			auto [U, d, V] = quantit::svd(A, split);
			dump(U, a[3]);
			dump(d, a[4]);
			dump(V, a[5]);
			timed(reps, [&]() { auto X = quantit::svd(A, split); });
		}
		else if (cmd == "svdt")
		{ // svdt A split tol min max pow OUT_U OUT_d OUT_V
			auto A = load(a[1]);
			size_t split = std::stoul(a[2]);
			double tol = std::stod(a[3]);
			size_t mn = std::stoul(a[4]);
			size_t mx = std::stoull(a[5]);
			double pw = std::stod(a[6]);
			auto [U, d, V] = quantit::svd(A, split, tol, mn, mx, pw);
			dump(U, a[7]);
			dump(d, a[8]);
			dump(V, a[9]);
			timed(reps, [&]() { auto X = quantit::svd(A, split, tol, mn, mx, pw); });
		}
		else if (cmd == "heff")
		{ // heff PSI H2 L R OUT
			auto psi = load(a[1]), H2 = load(a[2]), L = load(a[3]), R = load(a[4]);
			dump(details::hamil2site_times_state(psi, H2, L, R), a[5]);
			timed(reps, [&]() { auto X = details::hamil2site_times_state(psi, H2, L, R); });
		}
		else if (cmd == "lenv")
		{ // lenv H Y L OUT
			auto H = load(a[1]), Y = load(a[2]), L = load(a[3]);
			dump(compute_left_env(H, Y, L), a[4]);
			timed(reps, [&]() { auto X = compute_left_env(H, Y, L); });
		}
		else if (cmd == "renv")
		{
			auto H = load(a[1]), Y = load(a[2]), R = load(a[3]);
			dump(compute_right_env(H, Y, R), a[4]);
			timed(reps, [&]() { auto X = compute_right_env(H, Y, R); });
		}
		else if (cmd == "update")
		{ // update PSI H2 L R OUT_E OUT_PSI     (one_step_lanczos + eig2x2Mat + recombination, dmrg.cpp:623-651)
			auto psi = load(a[1]), H2 = load(a[2]), L = load(a[3]), R = load(a[4]);
			auto [E, p] = two_sites_update(psi, H2, L, R);
			dump(E, a[5]);
			dump(p, a[6]);
			timed(reps, [&]() { auto X = two_sites_update(psi, H2, L, R); });
		}
		else if (cmd == "mul")
		{ // broadcasting elementwise product (btensor.cpp:1204-1302)
			auto A = load(a[1]), B = load(a[2]);
			dump(A.mul(B), a[3]);
		}
		else if (cmd == "add")
		{ // add A B alpha OUT  (btensor.cpp:2666-2752)
			auto A = load(a[1]), B = load(a[2]);
			dump(A.add(B, std::stod(a[3])), a[4]);
		}
		else if (cmd == "heis" or cmd == "hub")
		{ // heis L maxbond cutoff conv maxit seed [dumpdir]  — dumps MPO + initial MPS before, final MPS after
			size_t L = std::stoul(a[1]);
			size_t maxbond = std::stoull(a[2]);
			double cutoff = std::stod(a[3]), conv = std::stod(a[4]);
			size_t maxit = std::stoul(a[5]);
			size_t seed = std::stoul(a[6]);
			std::string dir = a.size() > 7 ? a[7] : "";
			bMPO H = cmd == "heis" ? heisenberg_u1(L) : hubbard_u1u1(L, 4.0, 2.0);
			torch::manual_seed(1234 + seed);
			any_quantity target = cmd == "heis" ? any_quantity(Z(L % 2)) : any_quantity(ZZ(Z(L), Z(0)));
			bMPS psi = random_bMPS(4, H, target, {}, seed);
			if (!dir.empty())
				for (size_t i = 0; i < L; ++i)
				{
					dump(H[i], dir + "/H_" + std::to_string(i) + ".qtbt");
					dump(psi[i], dir + "/psi0_" + std::to_string(i) + ".qtbt");
				}
			std::printf("OC %zu\n", (size_t)psi.orthogonality_center);
			sweep_printer logger;
			auto t0 = std::chrono::steady_clock::now();
			auto E = dmrg(H, psi, dmrg_options(cutoff, conv, maxbond, 4, maxit), logger);
			auto t1 = std::chrono::steady_clock::now();
			std::printf("E0 %.14f\nTOTAL_MS %.3f\n", E.item().toDouble(),
			            std::chrono::duration<double, std::milli>(t1 - t0).count());
			auto Ec = contract(psi, psi, H);
			auto Nc = contract(psi, psi);
			std::printf("CONTRACT_E %.14f NORM %.14f\n", Ec.item().toDouble(), Nc.item().toDouble());
			if (!dir.empty())
				for (size_t i = 0; i < L; ++i)
					dump(psi[i], dir + "/psiF_" + std::to_string(i) + ".qtbt");
		}
		else if (cmd == "eigh")
		{ // eigh A split OUT_d OUT_U
			auto A = load(a[1]);
			size_t split = std::stoul(a[2]);
			auto [d, U] = quantit::eigh(A, split);
			dump(d, a[3]);
			dump(U, a[4]);
			timed(reps, [&]() { auto X = quantit::eigh(A, split); });
		}
		else if (cmd == "coalesce")
		{ // coalesce HDIR L cutoff OUTDIR — bMPO::coalesce on HDIR/H_i.qtbt, dumps OUTDIR/Hc_i.qtbt
			size_t L = std::stoul(a[2]);
			double cutoff = std::stod(a[3]);
			bMPO H(L);
			for (size_t i = 0; i < L; ++i)
				H[i] = load(a[1] + "/H_" + std::to_string(i) + ".qtbt");
			H.coalesce(cutoff);
			for (size_t i = 0; i < L; ++i)
				dump(H[i], a[4] + "/Hc_" + std::to_string(i) + ".qtbt");
		}
		else if (cmd == "moveoc")
		{ // moveoc DIR PREFIX L oc target OUTDIR [HDIR] — bMPS::move_oc on DIR/PREFIX_i.qtbt, dumps OUTDIR/psiM_i.qtbt and
		  // prints contract(psi, psi) (and contract(psi, psi, H) when HDIR holds H_i.qtbt) before / after
			std::string dir = a[1], prefix = a[2];
			size_t L = std::stoul(a[3]), oc = std::stoul(a[4]), target = std::stoul(a[5]);
			std::string out = a[6];
			std::vector<btensor> sites;
			for (size_t i = 0; i < L; ++i)
				sites.push_back(load(dir + "/" + prefix + "_" + std::to_string(i) + ".qtbt"));
			bMPS psi(sites, oc);
			std::printf("NORM_BEFORE %.14f\n", contract(psi, psi).item().toDouble());
			psi.move_oc((int)target);
			std::printf("OC %zu\nNORM_AFTER %.14f\n", (size_t)psi.orthogonality_center, contract(psi, psi).item().toDouble());
			if (a.size() > 7)
			{
				bMPO H(L);
				for (size_t i = 0; i < L; ++i)
					H[i] = load(a[7] + "/H_" + std::to_string(i) + ".qtbt");
				std::printf("CONTRACT_E %.14f\n", contract(psi, psi, H).item().toDouble());
			}
			for (size_t i = 0; i < L; ++i)
				dump(psi[i], out + "/psiM_" + std::to_string(i) + ".qtbt");
		}
		else
		{
			std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
			return 2;
		}
	}
	catch (const std::exception &e)
	{
		std::fprintf(stderr, "REF_EXCEPTION %s\n", e.what());
		return 3;
	}
	return 0;
}
