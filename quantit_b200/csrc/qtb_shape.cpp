// qtb_shape.cpp — structural reshapes, the generalised contraction and the packed-tensor file format.
//
//   reshape / reshape_as : reference btensor::reshape(index_groups) (sources/btensor.cpp:2986-3024) and
//                          btensor::reshape_as<mode>(other) (:3026-3083). Metadata only for packed blocks (the C-order
//                          flattening of a C-contiguous block is the same memory), a gather first for strided views.
//   tensorgdot           : D = alpha C + beta A.B, declared at reference include/blockTensor/btensor.h:624-627 (never
//                          defined there; semantics of the dense include/tensorgdot.h:22-86). When C has exactly the
//                          block table of A.B the linear combination is the epilogue of the grouped GEMM (one pass);
//                          otherwise contraction + union-merge axpby.
//   save / load          : one file per packed block tensor (structure, block table, arena image): SURVEY.md section
//                          8(f)4. The reference has no serialisation of btensors at all.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>

#include "qtb_ops.h"
#include "qtb_vec.h"

namespace qtb
{

namespace
{
const Tensor &packed_or_gather(Ctx &ctx, const Tensor &t, std::unique_ptr<Tensor> &hold)
{
	if (t.packed_canonical())
		return t;
	hold = contiguous(ctx, t);
	return *hold;
}
} // namespace

std::unique_ptr<Tensor> reshape(Ctx &ctx, const Tensor &a_in, const std::vector<i64> &index_groups)
{
	const i64 r = a_in.st.rank, nc = a_in.st.ct.nc;
	const i64 out_rank = (i64)index_groups.size() + 1;
	std::vector<i64> grp(out_rank + 1);
	grp[0] = 0;
	grp[out_rank] = r;
	for (i64 i = 0; i < (i64)index_groups.size(); ++i)
		grp[i + 1] = index_groups[i];
	for (i64 i = 0; i < out_rank; ++i)
		QTB_REQUIRE(grp[i] >= 0 && grp[i] <= grp[i + 1] && grp[i + 1] <= r, QTB_ERR_INVALID_ARGUMENT,
		            "reshape: index groups must be a non-decreasing list of positions inside [0, rank]");
	std::unique_ptr<Tensor> hold;
	const Tensor &a = packed_or_gather(ctx, a_in, hold);
	auto out = std::make_unique<Tensor>();
	Structure &st = out->st;
	st.rank = out_rank;
	st.ct = a.st.ct;
	st.sel = a.st.sel;
	// every combination of the grouped sections becomes a section (row-major, last index fastest, btensor.cpp:2832-2857):
	// size = product of the sizes, charge = product of the charges
	for (i64 gdim = 0; gdim < out_rank; ++gdim)
	{
		i64 count = 1;
		for (i64 d = grp[gdim]; d < grp[gdim + 1]; ++d)
			count *= a.st.nsec[d];
		st.nsec.push_back(count);
		std::vector<i64> ix(grp[gdim + 1] - grp[gdim], 0);
		for (i64 f = 0; f < count; ++f)
		{
			i64 size = 1;
			std::vector<i64> q(nc, 0);
			for (i64 d = grp[gdim]; d < grp[gdim + 1]; ++d)
			{
				const i64 s = ix[d - grp[gdim]];
				size *= a.st.size_of(d, s);
				for (i64 c = 0; c < nc; ++c)
					q[c] = a.st.ct.norm(q[c] + a.st.charge_of(d, s)[c], c);
			}
			st.sec_sizes.push_back(size);
			st.cvals.insert(st.cvals.end(), q.begin(), q.end());
			for (i64 k = (i64)ix.size() - 1; k >= 0; --k)
			{
				if (++ix[k] < a.st.nsec[grp[gdim] + k])
					break;
				ix[k] = 0;
			}
		}
	}
	st.finalize();
	out->nblocks = a.nblocks;
	out->index.resize(a.nblocks * out_rank);
	out->dims.resize(a.nblocks * out_rank);
	out->strides.resize(a.nblocks * out_rank);
	out->offs = a.offs;
	for (i64 b = 0; b < a.nblocks; ++b)
	{
		for (i64 gdim = out_rank - 1; gdim >= 0; --gdim)
		{
			i64 f = 0, size = 1;
			for (i64 d = grp[gdim]; d < grp[gdim + 1]; ++d)
			{
				f = f * a.st.nsec[d] + a.idx(b)[d];
				size *= a.dm(b)[d];
			}
			out->index[b * out_rank + gdim] = f;
			out->dims[b * out_rank + gdim] = size;
		}
		i64 s = 1;
		for (i64 gdim = out_rank - 1; gdim >= 0; --gdim)
		{ // the block is C-contiguous: merged dims keep the C-order strides
			out->strides[b * out_rank + gdim] = s;
			s *= out->dims[b * out_rank + gdim];
		}
	}
	// the row-major flattening is monotone: the block order is unchanged
	out->arena = a.arena;
	out->compute_hash();
	return out;
}

std::unique_ptr<Tensor> reshape_as(Ctx &ctx, const Tensor &a_in, const Tensor &like, bool overwrite_cvals)
{
	const i64 r = a_in.st.rank, ro = like.st.rank, nc = a_in.st.ct.nc;
	QTB_REQUIRE(a_in.st.ct == like.st.ct, QTB_ERR_INVALID_ARGUMENT, "reshape_as: different types of conserved quantities");
	auto nsections = [](const Structure &s)
	{
		i64 n = 1;
		for (auto v : s.nsec)
			n *= v;
		return n;
	};
	QTB_REQUIRE(nsections(a_in.st) == nsections(like.st), QTB_ERR_INVALID_ARGUMENT, "incompatible sections layouts");
	std::unique_ptr<Tensor> hold;
	const Tensor &a = packed_or_gather(ctx, a_in, hold);
	auto out = std::make_unique<Tensor>();
	out->st = like.st;
	if (!overwrite_cvals)
		out->st.sel = a.st.sel;
	out->nblocks = a.nblocks;
	out->index.resize(a.nblocks * ro);
	for (i64 b = 0; b < a.nblocks; ++b)
	{
		i64 f = 0;
		for (i64 d = 0; d < r; ++d)
			f = f * a.st.nsec[d] + a.idx(b)[d];
		for (i64 d = ro - 1; d >= 0; --d)
		{
			out->index[b * ro + d] = f % like.st.nsec[d];
			f /= like.st.nsec[d];
		}
		const i64 *ix = &out->index[b * ro];
		i64 numel = 1;
		for (i64 d = 0; d < ro; ++d)
			numel *= like.st.size_of(d, ix[d]);
		QTB_REQUIRE(numel == a.block_numel(b), QTB_ERR_INVALID_ARGUMENT, "incompatible block dimensions");
		if (overwrite_cvals)
			QTB_REQUIRE(out->st.allowed(ix), QTB_ERR_INVALID_ARGUMENT,
			            "a block of the original tensor is not allowed by the new selection rule");
		else
		{ // dims_only: the flux of the block must be the same in both descriptions
			for (i64 c = 0; c < nc; ++c)
			{
				i64 qa = 0, qo = 0;
				for (i64 d = 0; d < r; ++d)
					qa += a.st.charge_of(d, a.idx(b)[d])[c];
				for (i64 d = 0; d < ro; ++d)
					qo += like.st.charge_of(d, ix[d])[c];
				QTB_REQUIRE(a.st.ct.norm(qa, c) == a.st.ct.norm(qo, c), QTB_ERR_INVALID_ARGUMENT,
				            "incompatible conserved quantities");
			}
		}
	}
	out->dims_from_structure();
	out->strides.resize(a.nblocks * ro);
	for (i64 b = 0; b < a.nblocks; ++b)
	{
		i64 s = 1;
		for (i64 d = ro - 1; d >= 0; --d)
		{
			out->strides[b * ro + d] = s;
			s *= out->dims[b * ro + d];
		}
	}
	out->offs = a.offs;
	out->arena = a.arena;
	out->compute_hash();
	return out;
}

// ---------------------------------------------------------------------------------------------------------------------
std::unique_ptr<Tensor> tensorgdot(Ctx &ctx, const Tensor &c_in, const Tensor &a, const Tensor &b, const std::vector<i64> &da,
                                   const std::vector<i64> &db, double beta, double alpha)
{
	auto plan = get_plan(ctx, a, b, da, db);
	const Tensor &proto = plan->out_proto;
	QTB_REQUIRE(c_in.st.rank == proto.st.rank && c_in.st.ct == proto.st.ct && c_in.st.nsec == proto.st.nsec &&
	                c_in.st.sec_sizes == proto.st.sec_sizes && c_in.st.cvals == proto.st.cvals && c_in.st.sel == proto.st.sel,
	            QTB_ERR_INVALID_ARGUMENT, "tensorgdot: the added tensor does not have the structure of the contraction's result");
	bool any_empty = false;
	for (auto &o : plan->outs)
		any_empty |= (o.pair_begin == o.pair_end);
	if (c_in.layout_hash == proto.layout_hash && c_in.layout_hash2 == proto.layout_hash2 && !any_empty && plan->tile_cfg != 2)
	{ // same block table, same packed layout: the linear combination is the GEMM's epilogue
		auto out = std::make_unique<Tensor>(proto);
		out->arena = std::make_shared<Arena>(&ctx, plan->out_numel);
		launch_grouped_gemm(ctx, *plan, a.arena ? a.arena->ptr : nullptr, b.arena ? b.arena->ptr : nullptr, out->arena->ptr,
		                    nullptr, c_in.arena->ptr, alpha, beta);
		ctx.counters[0] += 0;
		return out;
	}
	auto t = tensordot(ctx, a, b, da, db);
	return axpby_dev(ctx, nullptr, alpha, c_in, nullptr, beta, *t, false);
}

// ---------------------------------------------------------------------------------------------------------------------
// file format "QTBPACK1": little endian, int64 header fields, then the tables, then the packed arena image
//   magic[8] | rank nc nblocks total_sections numel | mods[nc] | nsec[rank] | sec_sizes[ts] | cvals[ts*nc] | sel[nc] |
//   index[nblocks*rank] | data[numel] (blocks back to back in block order, each C-contiguous)
void save_tensor(Ctx &ctx, const Tensor &t, const char *path)
{
	std::vector<double> data((size_t)t.numel());
	download(ctx, t, data.data());
	std::FILE *f = std::fopen(path, "wb");
	QTB_REQUIRE(f != nullptr, QTB_ERR_RUNTIME, std::string("save: cannot open ") + path);
	const char magic[8] = {'Q', 'T', 'B', 'P', 'A', 'C', 'K', '1'};
	const i64 hdr[5] = {t.st.rank, t.st.ct.nc, t.nblocks, t.st.total_sections(), (i64)data.size()};
	bool ok = std::fwrite(magic, 1, 8, f) == 8 && std::fwrite(hdr, sizeof(i64), 5, f) == 5;
	auto put = [&](const std::vector<i64> &v) { ok = ok && std::fwrite(v.data(), sizeof(i64), v.size(), f) == v.size(); };
	put(t.st.ct.mods);
	put(t.st.nsec);
	put(t.st.sec_sizes);
	put(t.st.cvals);
	put(t.st.sel);
	put(t.index);
	ok = ok && std::fwrite(data.data(), sizeof(double), data.size(), f) == data.size();
	ok = (std::fclose(f) == 0) && ok;
	QTB_REQUIRE(ok, QTB_ERR_RUNTIME, std::string("save: short write to ") + path);
}

std::unique_ptr<Tensor> load_tensor(Ctx &ctx, const char *path)
{
	std::FILE *f = std::fopen(path, "rb");
	QTB_REQUIRE(f != nullptr, QTB_ERR_RUNTIME, std::string("load: cannot open ") + path);
	struct Closer
	{
		std::FILE *f;
		~Closer() { std::fclose(f); }
	} closer{f};
	char magic[8];
	i64 hdr[5];
	QTB_REQUIRE(std::fread(magic, 1, 8, f) == 8 && std::memcmp(magic, "QTBPACK1", 8) == 0 && std::fread(hdr, sizeof(i64), 5, f) == 5,
	            QTB_ERR_INVALID_ARGUMENT, std::string("load: not a packed block tensor file: ") + path);
	const i64 rank = hdr[0], nc = hdr[1], nb = hdr[2], ts = hdr[3], numel = hdr[4];
	QTB_REQUIRE(rank >= 0 && rank <= 8 && nc >= 1 && nc <= 8 && nb >= 0 && ts >= 0 && numel >= 0, QTB_ERR_INVALID_ARGUMENT,
	            "load: corrupt header");
	auto get = [&](std::vector<i64> &v, i64 n)
	{
		v.resize((size_t)n);
		QTB_REQUIRE(std::fread(v.data(), sizeof(i64), (size_t)n, f) == (size_t)n, QTB_ERR_INVALID_ARGUMENT, "load: truncated file");
	};
	Structure st;
	st.rank = rank;
	st.ct.nc = nc;
	get(st.ct.mods, nc);
	get(st.nsec, rank);
	get(st.sec_sizes, ts);
	get(st.cvals, ts * nc);
	get(st.sel, nc);
	std::vector<i64> index;
	get(index, nb * rank);
	std::vector<double> data((size_t)numel);
	QTB_REQUIRE(std::fread(data.data(), sizeof(double), (size_t)numel, f) == (size_t)numel, QTB_ERR_INVALID_ARGUMENT,
	            "load: truncated file");
	auto t = make_tensor(ctx, st, nb, index.data(), data.data());
	QTB_REQUIRE(t->numel() == numel, QTB_ERR_INVALID_ARGUMENT, "load: the data size does not match the block table");
	QTB_CUDA(cudaStreamSynchronize(ctx.stream)); // `data` is pageable: the upload must have consumed it before it goes away
	return t;
}

} // namespace qtb
