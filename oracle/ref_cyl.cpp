// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// ref_cyl: the reference's own 4x8 Heisenberg cylinder input (BASELINE.json configs[3] is its width-6 sibling; the 32 MPO
// tensors in COO text form + 32 charge files under /root/reference/tests/2dHeisenberg are the only fixtures the
// reference ships, tests/2Dheisenberg.cpp:318-330). The MPO is built by the reference's OWN helper functions: the
// unmodified test translation unit is compiled in place (its `main` renamed), so string2structure / make_tensor /
// guess_btensor and the embedded fixture strings are exactly the reference's. Nothing is copied into the repo.
//   ref_cyl DUMPDIR maxbond cutoff conv maxit seed [--threads N]
// dumps H_i.qtbt (after bMPO::coalesce(), as the reference test does) and psi0_i.qtbt, then runs the reference's
// dmrg() and prints one SWEEP line per sweep (same format as ref_harness heis).
#define main reference_2dheisenberg_main
#include "2Dheisenberg.cpp" // -I$(REF)/tests
#undef main
#include "ref_io.h"

struct cyl_printer : public dmrg_logger
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

int main(int argc, char **argv)
{
	torch::set_default_dtype(torch::scalarTypeToTypeMeta(torch::kFloat64));
	torch::InferenceMode guard;
	std::vector<std::string> a(argv + 1, argv + argc);
	int threads = 1;
	for (size_t i = 0; i + 1 < a.size();)
		if (a[i] == "--threads")
		{
			threads = std::stoi(a[i + 1]);
			a.erase(a.begin() + i, a.begin() + i + 2);
		}
		else
			++i;
	torch::set_num_threads(threads);
	at::init_num_threads();
	if (a.size() < 6)
	{
		std::puts("usage: ref_cyl DUMPDIR maxbond cutoff conv maxit seed [--threads N]");
		return 2;
	}
	try
	{
		const std::string dir = a[0];
		const size_t maxbond = std::stoull(a[1]);
		const double cutoff = std::stod(a[2]), conv = std::stod(a[3]);
		const size_t maxit = std::stoul(a[4]), seed = std::stoul(a[5]);
		// reference tests/2Dheisenberg.cpp:292-318, verbatim in structure
		quantit::MPO heis(32);
		int i = 0;
		for (auto &tens : heis)
		{
			tens = make_tensor(string2structure(mpo_strings[i]));
			++i;
		}
		quantit::bMPO bheis(32);
		using cval = quantit::conserved::Z;
		i = 0;
		auto phys = quantit::btensor({{{1, cval(-1)}, {1, cval(1)}}}, any_quantity(cval(0)));
		auto physdag = phys.conj();
		auto leftbond = quantit::btensor({{{1, cval(0)}}}, any_quantity(cval(0)));
		for (auto &tens : bheis)
		{
			auto before_missing = shape_from(leftbond, phys);
			tens = guess_btensor(heis[i], before_missing, physdag, 1e-4);
			auto rightbond = tens.shape_from({0, 0, -1, 0}).set_selection_rule_(any_quantity(cval(0)));
			leftbond = rightbond.conj();
			++i;
		}
		bheis.coalesce();
		torch::manual_seed(1234 + seed);
		quantit::bMPS state = quantit::random_bMPS(4, bheis, any_quantity(cval(0)), {}, seed);
		for (size_t s = 0; s < 32; ++s)
		{
			dump(bheis[s], dir + "/H_" + std::to_string(s) + ".qtbt");
			dump(state[s], dir + "/psi0_" + std::to_string(s) + ".qtbt");
		}
		std::printf("OC %zu\n", (size_t)state.orthogonality_center);
		cyl_printer logger;
		auto E = quantit::dmrg(bheis, state, dmrg_options(cutoff, conv, maxbond, 4, maxit), logger);
		std::printf("E0 %.14f\n", E.item().toDouble());
		auto Ec = contract(state, state, bheis);
		auto Nc = contract(state, state);
		std::printf("CONTRACT_E %.14f NORM %.14f\n", Ec.item().toDouble(), Nc.item().toDouble());
	}
	catch (const std::exception &e)
	{
		std::fprintf(stderr, "REF_EXCEPTION %s\n", e.what());
		return 3;
	}
	return 0;
}
