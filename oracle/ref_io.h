// TEST INFRASTRUCTURE ONLY — QTBT I/O and charge conversion helpers around the reference's public btensor API, shared by
// oracle/ref_harness.cpp and oracle/adaptor_check.cpp. Never part of the product path.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "MPT.h"
#include "blockTensor/LinearAlgebra.h"
#include "blockTensor/btensor.h"
#include "dmrg.h"
#include "dmrg_logger.h"
#include "dmrg_options.h"
#include "models.h"

namespace quantit
{
// defined in the reference sources, absent from its headers
bMPO to_bMPO(MPO &&, btensor &&);
btensor compute_left_env(const btensor &Hamil, const btensor &MPS, const btensor &left_env);
btensor compute_right_env(const btensor &Hamil, const btensor &MPS, const btensor &right_env);
std::tuple<btensor, btensor> two_sites_update(const btensor &state, const btensor &hamil, const btensor &Lenv,
                                              const btensor &Renv);
benv_holder generate_env(const bMPO &hamiltonian, const bMPS &state);
} // namespace quantit

using namespace quantit;
using Z = conserved::Z;
using ZZ = quantity<conserved::Z, conserved::Z>;
using i64 = int64_t;

static std::vector<i64> parse_ints(const std::string &s)
{ // every (possibly signed) integer found in s
	std::vector<i64> out;
	size_t i = 0;
	while (i < s.size())
	{
		if (isdigit(s[i]) or (s[i] == '-' and i + 1 < s.size() and isdigit(s[i + 1])))
		{
			size_t j = i + 1;
			while (j < s.size() and isdigit(s[j]))
				++j;
			out.push_back(std::stoll(s.substr(i, j - i)));
			i = j;
		}
		else
			++i;
	}
	return out;
}
static std::vector<i64> charge_ints(any_quantity_cref q) { return parse_ints(fmt::format("{}", q)); }
static any_quantity make_charge(const i64 *v, i64 nc)
{
	if (nc == 1)
		return any_quantity(Z(static_cast<int16_t>(v[0])));
	if (nc == 2)
		return any_quantity(ZZ(Z(static_cast<int16_t>(v[0])), Z(static_cast<int16_t>(v[1]))));
	throw std::invalid_argument("harness supports Z and ZxZ charges only");
}

static void write_i64(std::ofstream &f, const i64 *p, size_t n) { f.write(reinterpret_cast<const char *>(p), 8 * n); }
static void read_i64(std::ifstream &f, i64 *p, size_t n) { f.read(reinterpret_cast<char *>(p), 8 * n); }

static void dump(const btensor &t, const std::string &path)
{
	std::ofstream f(path, std::ios::binary);
	f.write("QTBT0001", 8);
	i64 rank = t.dim();
	auto sel = charge_ints(t.selection_rule->get());
	i64 nc = sel.size();
	i64 nblocks = std::distance(t.begin(), t.end());
	i64 hdr[3] = {rank, nc, nblocks};
	write_i64(f, hdr, 3);
	std::vector<i64> nsec(t.section_numbers().begin(), t.section_numbers().end());
	write_i64(f, nsec.data(), nsec.size());
	for (i64 d = 0; d < rank; ++d)
	{
		auto [b, e] = t.section_sizes(d);
		std::vector<i64> s(b, e);
		write_i64(f, s.data(), s.size());
	}
	for (i64 d = 0; d < rank; ++d)
	{
		auto [b, e] = t.section_cqtts(d);
		for (auto it = b; it != e; ++it)
		{
			auto c = charge_ints(*it);
			write_i64(f, c.data(), c.size());
		}
	}
	write_i64(f, sel.data(), sel.size());
	for (auto &blk : t)
		write_i64(f, std::get<0>(blk).data(), rank);
	// per block: the dims actually held by the block tensor (the reference lets them drift from the section sizes
	// only through bugs; dumping them lets the checker see that).
	for (auto &blk : t)
	{
		auto sz = std::get<1>(blk).sizes();
		std::vector<i64> s(sz.begin(), sz.end());
		write_i64(f, s.data(), rank);
	}
	for (auto &blk : t)
	{
		auto c = std::get<1>(blk).to(torch::kFloat64).contiguous();
		f.write(reinterpret_cast<const char *>(c.data_ptr<double>()), 8 * c.numel());
	}
}

static btensor load(const std::string &path)
{
	std::ifstream f(path, std::ios::binary);
	if (!f)
		throw std::runtime_error("cannot open " + path);
	char magic[8];
	f.read(magic, 8);
	if (std::strncmp(magic, "QTBT0001", 8) != 0)
		throw std::runtime_error("bad magic in " + path);
	i64 hdr[3];
	read_i64(f, hdr, 3);
	i64 rank = hdr[0], nc = hdr[1], nblocks = hdr[2];
	std::vector<i64> nsec(rank);
	read_i64(f, nsec.data(), rank);
	i64 tot = 0;
	for (auto n : nsec)
		tot += n;
	std::vector<i64> sizes(tot), cv(tot * nc), sel(nc);
	read_i64(f, sizes.data(), tot);
	read_i64(f, cv.data(), tot * nc);
	read_i64(f, sel.data(), nc);
	btensor::vec_list_t spec(rank);
	i64 k = 0;
	for (i64 d = 0; d < rank; ++d)
		for (i64 s = 0; s < nsec[d]; ++s, ++k)
			spec[d].emplace_back(static_cast<size_t>(sizes[k]), make_charge(&cv[k * nc], nc));
	btensor out(spec, make_charge(sel.data(), nc), torch::TensorOptions().dtype(torch::kFloat64));
	std::vector<i64> idx(nblocks * rank), dims(nblocks * rank);
	read_i64(f, idx.data(), idx.size());
	read_i64(f, dims.data(), dims.size());
	for (i64 b = 0; b < nblocks; ++b)
	{
		std::vector<i64> bi(idx.begin() + b * rank, idx.begin() + (b + 1) * rank);
		std::vector<i64> bd(dims.begin() + b * rank, dims.begin() + (b + 1) * rank);
		auto t = torch::empty(bd, torch::kFloat64);
		f.read(reinterpret_cast<char *>(t.data_ptr<double>()), 8 * t.numel());
		out.block(bi) = t;
	}
	return out;
}

static std::vector<i64> csv(const std::string &s)
{
	if (s == "-" or s.empty())
		return {};
	return parse_ints(s);
}

