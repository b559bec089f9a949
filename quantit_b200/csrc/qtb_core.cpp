// qtb_core.cpp — context, arenas, block tables, structural views, the contraction planner and tensordot.
#include "qtb_core.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>

#include "qtb_vec.h"

namespace qtb
{

// =====================================================================================================================
// context / memory
// =====================================================================================================================
Ctx::~Ctx()
{
	ctx_destroy_nccl(*this);
	plan_cache.clear();
	plan_slab.reset();
	trim_cache();
	for (Arena *a : arenas) // tensors that outlive the context: their arenas are orphaned, the blocks released below
		a->ctx = nullptr;
	arenas.clear();
	for (auto &kv : live_blocks)
		cudaFree(kv.first);
	live_blocks.clear();
	if (ring_base)
		cudaFreeHost(ring_base);
	for (auto ev : ring_event)
		if (ev)
			cudaEventDestroy(ev);
	if (pinned)
		cudaFreeHost(pinned);
	if (pinned_gauge_)
		cudaFreeHost(pinned_gauge_);
	for (auto ev : aux_events)
		cudaEventDestroy(ev);
	for (auto st : aux_streams)
		cudaStreamDestroy(st);
	if (own_stream && stream)
		cudaStreamDestroy(stream);
}

void Ctx::ensure_aux_streams(int n)
{
	while ((int)aux_streams.size() < n)
	{ // aux stream 0 gets the greatest priority (the heaviest SVD lane runs there: its kernels are the critical path and
	  // must not queue behind the other lanes' GEMMs), the others the least
		int least = 0, greatest = 0;
		QTB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
		cudaStream_t st = nullptr;
		QTB_CUDA(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, aux_streams.empty() ? greatest : least));
		aux_streams.push_back(st);
	}
	while (aux_events.size() < aux_streams.size() + 1)
	{
		cudaEvent_t ev = nullptr;
		QTB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		aux_events.push_back(ev);
	}
}

unsigned long long *Ctx::pinned_gauge()
{
	if (!pinned_gauge_)
		QTB_CUDA(cudaMallocHost((void **)&pinned_gauge_, 64 * sizeof(unsigned long long)));
	return pinned_gauge_;
}

void *Ctx::pinned_buf(size_t bytes)
{
	if (bytes > pinned_bytes)
	{
		if (pinned)
		{
			cudaStreamSynchronize(stream);
			cudaFreeHost(pinned);
		}
		pinned_bytes = std::max<size_t>(bytes * 2, size_t(1) << 20);
		QTB_CUDA(cudaMallocHost(&pinned, pinned_bytes));
	}
	return pinned;
}

// Device memory: a size-class caching allocator on top of cudaMalloc. Everything the engine launches is ordered on the
// ONE stream of the context, so a freed block can be handed out again immediately (the next user is ordered after the
// last one) and neither path ever touches the driver once the working set is warm. The stream-ordered pool
// (cudaMallocAsync) was measured to spend tens of milliseconds per call growing and re-mapping the pool when the arena
// sizes drift from site to site of a D=4096 sweep (sizes of 0.1..3 GB, never twice the same).
// Size classes: powers of two below 1 MiB, eighths of an octave above (<= 12.5 % slack); a request may also take a block
// of up to two classes above its own. On cudaMalloc failure the cache is released to the driver and the call retried.
static size_t size_class(size_t bytes)
{
	if (bytes <= 512)
		return 512;
	int lg = 63 - __builtin_clzll((unsigned long long)(bytes - 1)); // floor(log2(bytes-1))
	if (bytes <= (size_t(1) << 20))
		return size_t(1) << (lg + 1);
	const size_t step = size_t(1) << (lg - 3);
	return (bytes + step - 1) / step * step;
}

void Ctx::trim_cache()
{
	cudaStreamSynchronize(stream);
	for (auto &kv : free_bins)
		for (void *p : kv.second)
			cudaFree(p);
	free_bins.clear();
	cached_bytes = 0;
}

void *ctx_alloc(Ctx &ctx, size_t bytes)
{
	const size_t cls = size_class(bytes);
	auto it = ctx.free_bins.lower_bound(cls);
	for (int tries = 0; it != ctx.free_bins.end() && tries < 3 && it->first <= cls + cls / 4; ++it, ++tries)
		if (!it->second.empty())
		{
			void *p = it->second.back();
			it->second.pop_back();
			ctx.cached_bytes -= it->first;
			ctx.live_blocks[p] = it->first;
			return p;
		}
	void *p = nullptr;
	cudaError_t e = cudaMalloc(&p, cls);
	if (e == cudaErrorMemoryAllocation)
	{
		cudaGetLastError();
		ctx.trim_cache();
		e = cudaMalloc(&p, cls);
	}
	QTB_CUDA(e);
	ctx.live_blocks[p] = cls;
	return p;
}
void ctx_free(Ctx &ctx, void *p)
{
	if (!p)
		return;
	auto it = ctx.live_blocks.find(p);
	if (it == ctx.live_blocks.end())
		return;
	ctx.free_bins[it->second].push_back(p);
	ctx.cached_bytes += it->second;
	ctx.live_blocks.erase(it);
}

// Small structure tables go through a pinned ring (owned by the context) so that the copy is truly asynchronous; the
// ring is only recycled after a stream synchronisation.
// host -> device copy of a small table through the pinned ring (truly asynchronous). The ring is cut into kRingParts
// parts, each closed by an event when the write position leaves it; a part is reused only after ITS event has completed
// — the copies that read it are done — instead of draining the whole stream.
void ctx_stage_copy(Ctx &ctx, void *d, const void *host, size_t bytes)
{
	if (bytes == 0)
		return;
	const size_t need = (bytes + 255) & ~size_t(255);
	if (ctx.ring_base == nullptr || need * Ctx::kRingParts > ctx.ring_size)
	{
		if (ctx.ring_base)
		{
			cudaStreamSynchronize(ctx.stream);
			cudaFreeHost(ctx.ring_base);
			ctx.ring_base = nullptr;
		}
		ctx.ring_size = std::max<size_t>(need * 2 * Ctx::kRingParts, size_t(64) << 20);
		QTB_CUDA(cudaMallocHost((void **)&ctx.ring_base, ctx.ring_size));
		ctx.ring_pos = 0;
		ctx.ring_part = 0;
		for (int k = 0; k < Ctx::kRingParts; ++k)
		{
			if (!ctx.ring_event[k])
				QTB_CUDA(cudaEventCreateWithFlags(&ctx.ring_event[k], cudaEventDisableTiming));
			ctx.ring_busy[k] = false;
		}
	}
	const size_t part_size = (ctx.ring_size / Ctx::kRingParts) & ~size_t(255);
	if (ctx.ring_pos + need > (size_t)(ctx.ring_part + 1) * part_size)
	{ // leave this part: close it with an event, move to the start of the next one
		QTB_CUDA(cudaEventRecord(ctx.ring_event[ctx.ring_part], ctx.stream));
		ctx.ring_busy[ctx.ring_part] = true;
		ctx.ring_part = (ctx.ring_part + 1) % Ctx::kRingParts;
		ctx.ring_pos = (size_t)ctx.ring_part * part_size;
		if (ctx.ring_busy[ctx.ring_part])
		{
			QTB_CUDA(cudaEventSynchronize(ctx.ring_event[ctx.ring_part]));
			ctx.ring_busy[ctx.ring_part] = false;
		}
	}
	std::memcpy(ctx.ring_base + ctx.ring_pos, host, bytes);
	QTB_CUDA(cudaMemcpyAsync(d, ctx.ring_base + ctx.ring_pos, bytes, cudaMemcpyHostToDevice, ctx.stream));
	ctx.ring_pos += need;
	ctx.counters[4] += (i64)bytes;
}

void *ctx_upload(Ctx &ctx, const void *host, size_t bytes)
{
	void *d = ctx_alloc(ctx, bytes);
	ctx_stage_copy(ctx, d, host, bytes);
	return d;
}

Arena::Arena(Ctx *c, i64 n) : numel(n), owned(true), ctx(c)
{
	size_t bytes = std::max<i64>(n, 1) * sizeof(double);
	ptr = (double *)ctx_alloc(*c, bytes);
	c->counters[7] += (i64)bytes;
	c->arenas.insert(this);
}
Arena::~Arena()
{
	if (ctx)
		ctx->arenas.erase(this);
	if (owned && ptr && ctx)
	{
		ctx_free(*ctx, ptr);
		ctx->counters[7] -= (i64)(std::max<i64>(numel, 1) * sizeof(double));
	}
}

static inline uint64_t mix64(uint64_t h, uint64_t v)
{
	h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
	h *= 0xff51afd7ed558ccdull;
	h ^= h >> 33;
	return h;
}

Plan::~Plan() {} // the tables live in PlanSlabs (released with the last plan that uses them)

// device copy of a plan table: carved from the context's current slab, uploaded through the pinned ring
void ctx_stage_copy(Ctx &ctx, void *d, const void *host, size_t bytes);
static void *plan_upload(Ctx &ctx, Plan &plan, const void *host, size_t bytes)
{
	constexpr size_t kSlab = size_t(32) << 20;
	const size_t need = (bytes + 255) & ~size_t(255);
	if (!ctx.plan_slab || ctx.plan_slab->used + need > ctx.plan_slab->size)
	{
		auto slab = std::make_shared<PlanSlab>();
		slab->size = std::max(kSlab, need);
		QTB_CUDA(cudaMalloc((void **)&slab->base, slab->size));
		ctx.plan_slab = slab;
	}
	void *d = ctx.plan_slab->base + ctx.plan_slab->used;
	ctx.plan_slab->used += need;
	if (plan.slabs.empty() || plan.slabs.back() != ctx.plan_slab)
		plan.slabs.push_back(ctx.plan_slab);
	ctx_stage_copy(ctx, d, host, bytes);
	return d;
}

// =====================================================================================================================
// charge-sector sharding
// =====================================================================================================================
void Ctx::allreduce(double *ptr, i64 n)
{
	if (world <= 1 || n <= 0)
		return;
	if (nccl_allreduce(*this, ptr, n))
	{
		counters[0] += 1;
		return;
	}
	QTB_REQUIRE(allreduce_fn != nullptr, QTB_ERR_RUNTIME, "sharding is enabled but no allreduce callback is registered");
	const int rc = allreduce_fn(allreduce_user, ptr, n, (void *)stream);
	QTB_REQUIRE(rc == 0, QTB_ERR_RUNTIME, "the allreduce callback failed with code " + std::to_string(rc));
	counters[0] += 1;
}

int gemm_grid_limit(const Ctx &ctx, int tile_cfg)
{ // 64x64 tiles: 2 CTAs of 256 threads per SM; 128x128: 1 CTA of 384 threads; skinny: 8 light CTAs per SM
	return ctx.sm_count * (tile_cfg == 0 ? 2 : (tile_cfg == 1 ? 1 : 8));
}

std::vector<int32_t> schedule_tiles(std::vector<GemmTile> &tiles, std::vector<double> &cost, int ncta)
{
	const size_t n = tiles.size();
	std::vector<size_t> order(n);
	std::iota(order.begin(), order.end(), size_t(0));
	std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return cost[x] > cost[y]; });
	std::vector<std::vector<size_t>> per(ncta);
	static const bool snake = std::getenv("QTB_SCHED") && std::string(std::getenv("QTB_SCHED")) == "snake";
	if (snake)
	{ // experiment switch: cost-sorted list dealt to the CTAs forwards, backwards, forwards, ...
		for (size_t k = 0; k < n; ++k)
		{
			const size_t round = k / ncta, pos = k % ncta;
			per[(round & 1) ? (ncta - 1 - pos) : pos].push_back(order[k]);
		}
	}
	// min-heap of (load, cta)
	using Slot = std::pair<double, int>;
	std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> heap;
	for (int c = 0; c < ncta; ++c)
		heap.push({0.0, c});
	for (size_t t : order)
	{
		if (snake)
			break;
		Slot s = heap.top();
		heap.pop();
		per[s.second].push_back(t);
		s.first += cost[t];
		heap.push(s);
	}
	std::vector<GemmTile> nt;
	std::vector<double> nc;
	nt.reserve(n);
	nc.reserve(n);
	std::vector<int32_t> begin(ncta + 1, 0);
	for (int c = 0; c < ncta; ++c)
	{
		begin[c] = (int32_t)nt.size();
		for (size_t t : per[c])
		{
			nt.push_back(tiles[t]);
			nc.push_back(cost[t]);
		}
	}
	begin[ncta] = (int32_t)nt.size();
	tiles.swap(nt);
	cost.swap(nc);
	return begin;
}

std::vector<int32_t> skinny_item_prefix(const std::vector<GemmTile> &tiles)
{ // streaming (MPO) kernel: prefix sums of the number of kSkinnyRows-row work items of every block record
	std::vector<int32_t> p(tiles.size() + 1, 0);
	for (size_t t = 0; t < tiles.size(); ++t)
	{
		const i64 items = ((i64)tiles[t].M + kSkinnyRows - 1) / kSkinnyRows;
		QTB_REQUIRE((i64)p[t] + items < (i64(1) << 31), QTB_ERR_INVALID_ARGUMENT, "too many work items in one contraction");
		p[t + 1] = p[t] + (int32_t)items;
	}
	return p;
}

std::vector<int32_t> lpt_assign(const std::vector<double> &weights, int world)
{ // longest processing time first: heaviest section to the least loaded rank; ties -> lower section / lower rank
	const size_t n = weights.size();
	std::vector<size_t> order(n);
	std::iota(order.begin(), order.end(), size_t(0));
	std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return weights[x] > weights[y]; });
	std::vector<double> load(std::max(world, 1), 0.0);
	std::vector<int32_t> owner(n, 0);
	for (size_t s : order)
	{
		int best = 0;
		for (int r = 1; r < (int)load.size(); ++r)
			if (load[r] < load[best])
				best = r;
		owner[s] = best;
		load[best] += weights[s];
	}
	return owner;
}

void add_section_weights(const Plan &plan, i64 owner_dim, std::vector<double> &weights)
{
	const Tensor &o = plan.out_proto;
	QTB_REQUIRE(owner_dim >= 0 && owner_dim < o.st.rank, QTB_ERR_INVALID_ARGUMENT, "owner dim out of range");
	if ((i64)weights.size() < o.st.nsec[owner_dim])
		weights.resize(o.st.nsec[owner_dim], 0.0);
	for (i64 b = 0; b < o.nblocks; ++b)
		weights[o.idx(b)[owner_dim]] += (double)plan.out_flops[b] + 1.0; // +1: empty blocks still cost a tile visit
}

static const Plan::Owned &owned_tiles(Ctx &ctx, const std::shared_ptr<Plan> &plan, i64 owner_dim,
                                      const std::vector<int32_t> &owner, const std::vector<i64> *c_off)
{
	uint64_t h = mix64(0x51a7d, (uint64_t)owner_dim * 131 + (uint64_t)ctx.rank * 7 + (uint64_t)ctx.world);
	for (auto r : owner)
		h = mix64(h, (uint64_t)r + 3);
	if (c_off)
		for (auto v : *c_off)
			h = mix64(h, (uint64_t)v + 11);
	auto it = plan->owned.find(h);
	if (it == plan->owned.end())
	{
		const Tensor &o = plan->out_proto;
		std::vector<GemmTile> mine;
		std::vector<double> cost;
		Plan::Owned ow;
		for (size_t t = 0; t < plan->tiles.size(); ++t)
			if (owner[o.idx(plan->tiles[t].out_blk)[owner_dim]] == ctx.rank)
			{
				mine.push_back(plan->tiles[t]);
				if (c_off) // the block is written somewhere else than the plan's own packed layout
					mine.back().c_off = (*c_off)[plan->tiles[t].out_blk];
				cost.push_back(plan->tile_cost[t]);
			}
		for (i64 ob = 0; ob < o.nblocks; ++ob)
			if (owner[o.idx(ob)[owner_dim]] == ctx.rank)
				ow.flops += plan->out_flops[ob];
		ow.ntiles = (int)mine.size();
		if (!mine.empty())
		{
			ow.ncta = std::max(1, std::min<int>(ow.ntiles, gemm_grid_limit(ctx, plan->tile_cfg)));
			const auto cb = plan->tile_cfg == 2 ? skinny_item_prefix(mine) : schedule_tiles(mine, cost, ow.ncta);
			const size_t tb = (mine.size() * sizeof(GemmTile) + 15) & ~size_t(15);
			std::vector<char> blob(tb + cb.size() * sizeof(int32_t));
			std::memcpy(blob.data(), mine.data(), mine.size() * sizeof(GemmTile));
			std::memcpy(blob.data() + tb, cb.data(), cb.size() * sizeof(int32_t));
			ow.d_tiles = (GemmTile *)plan_upload(ctx, *plan, blob.data(), blob.size());
			ow.d_cta_begin = (int32_t *)((char *)ow.d_tiles + tb);
		}
		it = plan->owned.emplace(h, ow).first;
	}
	return it->second;
}

std::unique_ptr<Tensor> tensordot_owned(Ctx &ctx, const std::shared_ptr<Plan> &plan, const Tensor &a, const Tensor &b,
                                        i64 owner_dim, const std::vector<int32_t> &owner, bool zero_rest)
{
	const Plan::Owned &ow = owned_tiles(ctx, plan, owner_dim, owner, nullptr);
	auto out = std::make_unique<Tensor>(plan->out_proto);
	out->arena = std::make_shared<Arena>(&ctx, plan->out_numel);
	if (plan->out_numel && zero_rest)
		QTB_CUDA(cudaMemsetAsync(out->arena->ptr, 0, plan->out_numel * sizeof(double), ctx.stream));
	else if (plan->out_numel)
	{ // only the owned blocks that no pair writes
		const Tensor &o = plan->out_proto;
		for (i64 ob = 0; ob < o.nblocks; ++ob)
			if (plan->outs[ob].pair_begin == plan->outs[ob].pair_end && owner[o.idx(ob)[owner_dim]] == ctx.rank && o.block_numel(ob) > 0)
				QTB_CUDA(cudaMemsetAsync(out->arena->ptr + o.offs[ob], 0, o.block_numel(ob) * sizeof(double), ctx.stream));
	}
	launch_grouped_gemm(ctx, *plan, a.arena ? a.arena->ptr : nullptr, b.arena ? b.arena->ptr : nullptr, out->arena->ptr, &ow);
	return out;
}

void tensordot_owned_into(Ctx &ctx, const std::shared_ptr<Plan> &plan, const Tensor &a, const Tensor &b, i64 owner_dim,
                          const std::vector<int32_t> &owner, const std::vector<i64> &c_off, double *c_arena)
{
	const Plan::Owned &ow = owned_tiles(ctx, plan, owner_dim, owner, &c_off);
	launch_grouped_gemm(ctx, *plan, a.arena ? a.arena->ptr : nullptr, b.arena ? b.arena->ptr : nullptr, c_arena, &ow);
}

// =====================================================================================================================
// structure / block tables
// =====================================================================================================================
void Structure::finalize()
{
	sec_off.assign(rank + 1, 0);
	for (i64 d = 0; d < rank; ++d)
		sec_off[d + 1] = sec_off[d] + nsec[d];
}
bool Structure::allowed(const i64 *index) const
{
	for (i64 c = 0; c < ct.nc; ++c)
	{
		i64 q = 0;
		for (i64 d = 0; d < rank; ++d)
			q += charge_of(d, index[d])[c];
		if (ct.norm(q, c) != ct.norm(sel[c], c))
			return false;
	}
	return true;
}
i64 Structure::dim_size(i64 dim) const
{
	i64 s = 0;
	for (i64 k = 0; k < nsec[dim]; ++k)
		s += size_of(dim, k);
	return s;
}

std::vector<i64> sort_blocks(i64 rank, std::vector<i64> &index)
{
	const i64 nb = rank ? (i64)index.size() / rank : (index.empty() ? 0 : 0);
	std::vector<i64> perm(nb);
	std::iota(perm.begin(), perm.end(), 0);
	std::stable_sort(perm.begin(), perm.end(),
	                 [&](i64 x, i64 y)
	                 {
		                 return std::lexicographical_compare(index.begin() + x * rank, index.begin() + (x + 1) * rank,
		                                                     index.begin() + y * rank, index.begin() + (y + 1) * rank);
	                 });
	std::vector<i64> sorted(index.size());
	for (i64 i = 0; i < nb; ++i)
		std::copy(index.begin() + perm[i] * rank, index.begin() + (perm[i] + 1) * rank, sorted.begin() + i * rank);
	index.swap(sorted);
	return perm;
}

bool Tensor::block_contiguous(i64 b) const
{
	i64 expect = 1;
	for (i64 d = st.rank - 1; d >= 0; --d)
	{
		if (dm(b)[d] != 1 && sd(b)[d] != expect)
			return false;
		expect *= dm(b)[d];
	}
	return true;
}
bool Tensor::packed_canonical() const
{
	for (i64 b = 0; b < nblocks; ++b)
		if (!block_contiguous(b))
			return false;
	return true;
}
static inline uint64_t mix64b(uint64_t h, uint64_t v)
{ // an independent mixer (splitmix64 finaliser over a different combination) for the second half of the cache key
	h = (h ^ v) * 0xbf58476d1ce4e5b9ull;
	h ^= h >> 29;
	h *= 0x94d049bb133111ebull;
	h ^= h >> 32;
	return h + 0x632be59bd9b4e019ull;
}
void Tensor::compute_hash()
{
	{
		uint64_t g = 0x9ae16a3b2f90404full;
		auto eat = [&](const std::vector<i64> &v)
		{
			g = mix64b(g, (uint64_t)v.size());
			for (auto x : v)
				g = mix64b(g, (uint64_t)x);
		};
		g = mix64b(g, (uint64_t)st.rank);
		g = mix64b(g, (uint64_t)nblocks);
		g = mix64b(g, (uint64_t)st.ct.nc);
		eat(st.ct.mods);
		eat(st.nsec);
		eat(st.sec_sizes);
		eat(st.cvals);
		eat(st.sel);
		eat(index);
		eat(dims);
		eat(strides);
		eat(offs);
		layout_hash2 = g;
	}
	uint64_t h = 0x1234567ull;
	h = mix64(h, (uint64_t)st.rank);
	h = mix64(h, (uint64_t)nblocks);
	h = mix64(h, (uint64_t)st.ct.nc);
	for (auto v : st.ct.mods)
		h = mix64(h, (uint64_t)v + 5);
	for (auto v : st.nsec)
		h = mix64(h, (uint64_t)v);
	for (auto v : st.sec_sizes)
		h = mix64(h, (uint64_t)v);
	for (auto v : st.cvals)
		h = mix64(h, (uint64_t)v);
	for (auto v : st.sel)
		h = mix64(h, (uint64_t)v);
	for (auto v : index)
		h = mix64(h, (uint64_t)v);
	for (auto v : dims)
		h = mix64(h, (uint64_t)v);
	for (auto v : strides)
		h = mix64(h, (uint64_t)v);
	for (auto v : offs)
		h = mix64(h, (uint64_t)v);
	layout_hash = h;
}
void Tensor::dims_from_structure()
{
	dims.resize(nblocks * st.rank);
	for (i64 b = 0; b < nblocks; ++b)
		for (i64 d = 0; d < st.rank; ++d)
			dims[b * st.rank + d] = st.size_of(d, index[b * st.rank + d]);
}
i64 Tensor::layout_packed()
{
	strides.resize(nblocks * st.rank);
	offs.resize(nblocks);
	i64 pos = 0;
	for (i64 b = 0; b < nblocks; ++b)
	{
		i64 s = 1;
		for (i64 d = st.rank - 1; d >= 0; --d)
		{
			strides[b * st.rank + d] = s;
			s *= dims[b * st.rank + d];
		}
		offs[b] = pos;
		pos += (s + kBlockAlign - 1) / kBlockAlign * kBlockAlign;
	}
	return pos;
}
i64 Tensor::find_block(const i64 *index_) const
{
	i64 lo = 0, hi = nblocks;
	const i64 r = st.rank;
	while (lo < hi)
	{
		i64 mid = (lo + hi) / 2;
		if (std::lexicographical_compare(idx(mid), idx(mid) + r, index_, index_ + r))
			lo = mid + 1;
		else
			hi = mid;
	}
	if (lo < nblocks && std::equal(idx(lo), idx(lo) + r, index_))
		return lo;
	return -1;
}

// =====================================================================================================================
// creation / download
// =====================================================================================================================
static void check_structure(const Structure &st)
{ // the structural half of reference check_tensor, btensor.cpp:405-467
	QTB_REQUIRE(st.rank >= 0 && (i64)st.nsec.size() == st.rank, QTB_ERR_INVALID_ARGUMENT,
	            "Invalid argument to construct a block tensor: rank and section table disagree");
	QTB_REQUIRE((i64)st.sel.size() == st.ct.nc, QTB_ERR_INVALID_ARGUMENT, "selection rule has the wrong arity");
	i64 tot = 0;
	for (auto n : st.nsec)
	{
		QTB_REQUIRE(n >= 0, QTB_ERR_INVALID_ARGUMENT, "negative section count");
		tot += n;
	}
	QTB_REQUIRE((i64)st.sec_sizes.size() == tot && (i64)st.cvals.size() == tot * st.ct.nc, QTB_ERR_INVALID_ARGUMENT,
	            "Invalid argument to construct a block tensor: section sizes / conserved values length mismatch");
	for (auto s : st.sec_sizes)
		QTB_REQUIRE(s >= 0, QTB_ERR_INVALID_ARGUMENT, "negative section size");
}

std::unique_ptr<Tensor> make_tensor(Ctx &ctx, const Structure &st_in, i64 nblocks, const i64 *block_index,
                                    const double *host_data)
{
	auto t = std::make_unique<Tensor>();
	t->st = st_in;
	t->st.finalize();
	check_structure(t->st);
	const i64 r = t->st.rank;
	t->nblocks = nblocks;
	t->index.assign(block_index, block_index + nblocks * r);
	for (i64 b = 0; b < nblocks; ++b)
	{
		for (i64 d = 0; d < r; ++d)
			QTB_REQUIRE(t->index[b * r + d] >= 0 && t->index[b * r + d] < t->st.nsec[d], QTB_ERR_INVALID_ARGUMENT,
			            "block index out of the section range");
		QTB_REQUIRE(t->st.allowed(&t->index[b * r]), QTB_ERR_INVALID_ARGUMENT,
		            "Invalid argument to construct a block tensor: a block violates the selection rule");
	}
	std::vector<i64> perm;
	if (r > 0)
		perm = sort_blocks(r, t->index);
	else
	{
		QTB_REQUIRE(nblocks <= 1, QTB_ERR_INVALID_ARGUMENT, "a rank-0 tensor holds at most one block");
		perm.assign(nblocks, 0);
	}
	for (i64 b = 1; b < nblocks; ++b)
		QTB_REQUIRE(!std::equal(t->idx(b - 1), t->idx(b - 1) + r, t->idx(b)), QTB_ERR_INVALID_ARGUMENT,
		            "Invalid argument to construct a block tensor: repeated block index");
	t->dims_from_structure();
	const i64 total = t->layout_packed();
	t->arena = std::make_shared<Arena>(&ctx, total);
	if (total > 0)
	{
		if (host_data == nullptr)
			QTB_CUDA(cudaMemsetAsync(t->arena->ptr, 0, total * sizeof(double), ctx.stream));
		else
		{
			bool sorted_already = true;
			for (i64 nb = 0; nb < nblocks; ++nb)
				sorted_already &= (perm[nb] == nb);
			if (sorted_already)
			{ // the arena image IS the caller's buffer: one DMA straight from it (direct when the buffer is pinned)
				QTB_CUDA(cudaMemcpyAsync(t->arena->ptr, host_data, total * sizeof(double), cudaMemcpyHostToDevice,
				                         ctx.stream));
			}
			else
			{
				// source offsets in the GIVEN order
				std::vector<i64> src_off(nblocks);
				std::vector<i64> numel_given(nblocks);
				for (i64 nb = 0; nb < nblocks; ++nb)
					numel_given[perm[nb]] = t->block_numel(nb);
				i64 pos = 0;
				for (i64 g = 0; g < nblocks; ++g)
				{
					src_off[g] = pos;
					pos += numel_given[g];
				}
				QTB_CUDA(cudaStreamSynchronize(ctx.stream)); // staging buffer may still be in flight
				double *stage = (double *)ctx.pinned_buf(total * sizeof(double));
				for (i64 nb = 0; nb < nblocks; ++nb)
					std::memcpy(stage + t->offs[nb], host_data + src_off[perm[nb]], t->block_numel(nb) * sizeof(double));
				QTB_CUDA(cudaMemcpyAsync(t->arena->ptr, stage, total * sizeof(double), cudaMemcpyHostToDevice,
				                         ctx.stream));
			}
			ctx.counters[4] += total * (i64)sizeof(double);
		}
	}
	t->compute_hash();
	return t;
}

std::unique_ptr<Tensor> contiguous(Ctx &ctx, const Tensor &t)
{
	auto out = std::make_unique<Tensor>();
	out->st = t.st;
	out->nblocks = t.nblocks;
	out->index = t.index;
	out->dims = t.dims;
	const i64 total = out->layout_packed();
	out->arena = std::make_shared<Arena>(&ctx, total);
	if (total > 0)
		QTB_CUDA(cudaMemsetAsync(out->arena->ptr, 0, total * sizeof(double), ctx.stream));
	std::vector<GatherDesc> descs;
	for (i64 b = 0; b < t.nblocks; ++b)
	{
		GatherDesc d{};
		d.src_off = t.offs[b];
		d.dst_off = out->offs[b];
		d.numel = t.block_numel(b);
		d.rank = (int)t.st.rank;
		QTB_REQUIRE(d.rank <= 8, QTB_ERR_INVALID_ARGUMENT, "rank > 8 is not supported");
		for (int k = 0; k < d.rank; ++k)
		{
			d.dims[k] = t.dm(b)[k];
			d.strides[k] = t.sd(b)[k];
		}
		if (d.numel > 0)
			descs.push_back(d);
	}
	launch_gather(ctx, descs, t.arena->ptr, out->arena->ptr);
	out->compute_hash();
	return out;
}

void download(Ctx &ctx, const Tensor &t, double *host_out)
{
	std::unique_ptr<Tensor> tmp;
	const Tensor *src = &t;
	if (!t.packed_canonical())
	{
		tmp = contiguous(ctx, t);
		src = tmp.get();
	}
	i64 total = 0;
	bool back_to_back = true;
	for (i64 b = 0; b < src->nblocks; ++b)
	{
		back_to_back &= (src->offs[b] == src->offs[0] + total);
		total += src->block_numel(b);
	}
	if (total == 0)
		return;
	if (back_to_back)
	{ // one DMA straight into the caller's buffer
		QTB_CUDA(cudaMemcpyAsync(host_out, src->arena->ptr + src->offs[0], total * sizeof(double),
		                         cudaMemcpyDeviceToHost, ctx.stream));
		QTB_CUDA(cudaStreamSynchronize(ctx.stream));
	}
	else
	{ // blocks of a view may sit anywhere in a larger arena
		i64 pos = 0;
		for (i64 b = 0; b < src->nblocks; ++b)
		{
			const i64 n = src->block_numel(b);
			if (n)
				QTB_CUDA(cudaMemcpyAsync(host_out + pos, src->arena->ptr + src->offs[b], n * sizeof(double),
				                         cudaMemcpyDeviceToHost, ctx.stream));
			pos += n;
		}
		QTB_CUDA(cudaStreamSynchronize(ctx.stream));
	}
	ctx.counters[5] += total * (i64)sizeof(double);
}

// =====================================================================================================================
// structural views
// =====================================================================================================================
std::unique_ptr<Tensor> permute(const Tensor &a, const std::vector<i64> &perm_in)
{ // reference btensor::permute, btensor.cpp:1754-1802: metadata only, blocks keep aliasing the same storage
	const i64 r = a.st.rank;
	QTB_REQUIRE((i64)perm_in.size() == r, QTB_ERR_INVALID_ARGUMENT, "permutation length differs from the tensor rank");
	std::vector<i64> perm(r);
	std::vector<char> seen(r, 0);
	for (i64 i = 0; i < r; ++i)
	{
		perm[i] = perm_in[i] < 0 ? perm_in[i] + r : perm_in[i];
		QTB_REQUIRE(perm[i] >= 0 && perm[i] < r && !seen[perm[i]], QTB_ERR_INVALID_ARGUMENT, "invalid permutation");
		seen[perm[i]] = 1;
	}
	auto out = std::make_unique<Tensor>();
	out->st.rank = r;
	out->st.ct = a.st.ct;
	out->st.sel = a.st.sel;
	out->st.nsec.resize(r);
	for (i64 i = 0; i < r; ++i)
	{
		const i64 p = perm[i];
		out->st.nsec[i] = a.st.nsec[p];
		for (i64 s = 0; s < a.st.nsec[p]; ++s)
		{
			out->st.sec_sizes.push_back(a.st.size_of(p, s));
			const i64 *c = a.st.charge_of(p, s);
			out->st.cvals.insert(out->st.cvals.end(), c, c + a.st.ct.nc);
		}
	}
	out->st.finalize();
	out->nblocks = a.nblocks;
	out->index.resize(a.index.size());
	for (i64 b = 0; b < a.nblocks; ++b)
		for (i64 i = 0; i < r; ++i)
			out->index[b * r + i] = a.index[b * r + perm[i]];
	std::vector<i64> order(a.nblocks);
	if (r > 0)
		order = sort_blocks(r, out->index);
	else
		std::iota(order.begin(), order.end(), 0);
	out->dims.resize(a.dims.size());
	out->strides.resize(a.strides.size());
	out->offs.resize(a.nblocks);
	for (i64 nb = 0; nb < a.nblocks; ++nb)
	{
		const i64 ob = order[nb];
		for (i64 i = 0; i < r; ++i)
		{
			out->dims[nb * r + i] = a.dims[ob * r + perm[i]];
			out->strides[nb * r + i] = a.strides[ob * r + perm[i]];
		}
		out->offs[nb] = a.offs[ob];
	}
	out->arena = a.arena;
	out->compute_hash();
	return out;
}

std::unique_ptr<Tensor> conj(const Tensor &a)
{ // reference btensor::conj for a real dtype: conj_only() is the identity, inverse_cvals_() flips every section
  // charge and the selection rule (btensor.cpp:2156-2172). Storage is shared.
	auto out = std::make_unique<Tensor>(a);
	const i64 nc = a.st.ct.nc;
	for (size_t i = 0; i < out->st.cvals.size(); ++i)
		out->st.cvals[i] = a.st.ct.norm(-a.st.cvals[i], (i64)(i % nc));
	for (i64 c = 0; c < nc; ++c)
		out->st.sel[c] = a.st.ct.norm(-a.st.sel[c], c);
	out->compute_hash();
	return out;
}

// =====================================================================================================================
// contraction planner
// =====================================================================================================================
namespace
{
// int32 offsets of the C-order flattening of `dimlist` of block b (relative to the block's first element)
std::vector<int32_t> flat_offsets(const Tensor &t, i64 b, const std::vector<i64> &dimlist)
{
	i64 n = 1;
	for (auto d : dimlist)
		n *= t.dm(b)[d];
	std::vector<int32_t> out((size_t)n);
	if (n == 0)
		return out;
	const size_t nd = dimlist.size();
	std::vector<i64> counter(nd, 0);
	i64 off = 0;
	for (i64 e = 0; e < n; ++e)
	{
		QTB_REQUIRE(off >= 0 && off < (i64(1) << 31), QTB_ERR_INVALID_ARGUMENT, "block extent exceeds 2^31 elements");
		out[(size_t)e] = (int32_t)off;
		for (i64 k = (i64)nd - 1; k >= 0; --k)
		{
			const i64 d = dimlist[k];
			if (++counter[k] < t.dm(b)[d])
			{
				off += t.sd(b)[d];
				break;
			}
			off -= (counter[k] - 1) * t.sd(b)[d];
			counter[k] = 0;
		}
	}
	return out;
}
} // namespace

static std::shared_ptr<Plan> build_plan(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a_in,
                                        const std::vector<i64> &dims_b_in)
{
	// QTB_PROFILE >= 3: where the host time of a plan goes (accumulated per context, printed by prof_dump)
	auto tp0 = std::chrono::steady_clock::now();
	auto lap = [&](int slot)
	{
		if (ctx.prof_level < 3)
			return;
		auto now = std::chrono::steady_clock::now();
		ctx.plan_phase_ms[slot] += std::chrono::duration<double, std::milli>(now - tp0).count();
		tp0 = now;
	};
	// ---- validation: reference compute_tdot_shape + check_product_compat<true>, btensor.cpp:841-883,783-825 ----
	QTB_REQUIRE(dims_a_in.size() == dims_b_in.size(), QTB_ERR_CHECK,
	            "both dimension lists should have the same length.");
	QTB_REQUIRE(a.st.ct == b.st.ct, QTB_ERR_CHECK, "the two tensors have different type of conserved quantities");
	const i64 k = (i64)dims_a_in.size();
	const i64 ra = a.st.rank, rb = b.st.rank;
	std::vector<i64> dims_a(k), dims_b(k);
	std::vector<char> ca(ra, 0), cb(rb, 0);
	for (i64 i = 0; i < k; ++i)
	{
		dims_a[i] = dims_a_in[i] < 0 ? dims_a_in[i] + ra : dims_a_in[i];
		dims_b[i] = dims_b_in[i] < 0 ? dims_b_in[i] + rb : dims_b_in[i];
		QTB_REQUIRE(dims_a[i] >= 0 && dims_a[i] < ra && dims_b[i] >= 0 && dims_b[i] < rb, QTB_ERR_CHECK,
		            "contracted dimension out of range");
		QTB_REQUIRE(!ca[dims_a[i]] && !cb[dims_b[i]], QTB_ERR_CHECK, "dim appears multiple times in the list of dims");
		ca[dims_a[i]] = cb[dims_b[i]] = 1;
	}
	const i64 nc = a.st.ct.nc;
	for (i64 i = 0; i < k; ++i)
	{
		const i64 s1 = a.st.nsec[dims_a[i]], s2 = b.st.nsec[dims_b[i]];
		QTB_REQUIRE(s1 == s2, QTB_ERR_CHECK,
		            "contracted dimensions need to match, but first has " + std::to_string(s1) +
		                " sections along dim " + std::to_string(dims_a[i]) + "  and second has " + std::to_string(s2) +
		                " sections along dim " + std::to_string(dims_b[i]));
		for (i64 s = 0; s < s1; ++s)
			for (i64 c = 0; c < nc; ++c)
				QTB_REQUIRE(a.st.ct.norm(a.st.charge_of(dims_a[i], s)[c] + b.st.charge_of(dims_b[i], s)[c], c) == 0,
				            QTB_ERR_CHECK,
				            "contracted conserved numbers need to sum to zero, but there is a violation when "
				            "contracting dim " +
				                std::to_string(dims_a[i]) + " of the left tensor with dim " +
				                std::to_string(dims_b[i]) + " of the right tensor");
	}
	std::vector<i64> free_a, free_b;
	for (i64 i = 0; i < ra; ++i)
		if (!ca[i])
			free_a.push_back(i);
	for (i64 i = 0; i < rb; ++i)
		if (!cb[i])
			free_b.push_back(i);
	const i64 nfa = (i64)free_a.size(), nfb = (i64)free_b.size();
	QTB_REQUIRE(nfa + nfb <= 8, QTB_ERR_INVALID_ARGUMENT, "output rank > 8 is not supported");

	auto plan = std::make_shared<Plan>();
	plan->ctx = &ctx;
	Tensor &out = plan->out_proto;
	// ---- output structure: reference compute_tdot_cval_sectSize, btensor.cpp:1908-1935 ----
	out.st.rank = nfa + nfb;
	out.st.ct = a.st.ct;
	out.st.sel.resize(nc);
	for (i64 c = 0; c < nc; ++c)
		out.st.sel[c] = a.st.ct.norm(a.st.sel[c] + b.st.sel[c], c);
	auto push_dim = [&](const Tensor &t, i64 d)
	{
		out.st.nsec.push_back(t.st.nsec[d]);
		for (i64 s = 0; s < t.st.nsec[d]; ++s)
		{
			out.st.sec_sizes.push_back(t.st.size_of(d, s));
			const i64 *c = t.st.charge_of(d, s);
			out.st.cvals.insert(out.st.cvals.end(), c, c + nc);
		}
	};
	for (auto d : free_a)
		push_dim(a, d);
	for (auto d : free_b)
		push_dim(b, d);
	out.st.finalize();

	lap(0);
	// ---- block-pair matching: reference two-pointer merge over "columns", btensor.cpp:2057-2108 ----
	// Equivalent formulation: every (A block, B block) pair with equal contracted block indices, grouped by the output
	// index (freeA, freeB) in ascending order, pairs inside a group in ascending contracted index.
	// Keys are mixed-radix encodings of the block indices (radix = section count of the dim): their numeric order is the
	// lexicographic order of the index vectors, and one sort of 128-bit integers replaces the map / vector compares.
	using u128 = unsigned __int128;
	for (i64 d = 0; d < ra; ++d)
		QTB_REQUIRE(a.st.nsec[d] < (i64(1) << 15), QTB_ERR_INVALID_ARGUMENT, "more than 32767 sections along one dim");
	for (i64 d = 0; d < rb; ++d)
		QTB_REQUIRE(b.st.nsec[d] < (i64(1) << 15), QTB_ERR_INVALID_ARGUMENT, "more than 32767 sections along one dim");
	auto encode = [](const Tensor &t, i64 blk, const std::vector<i64> &dl, u128 key)
	{
		for (auto d : dl)
			key = key * (u128)std::max<i64>(t.st.nsec[d], 1) + (u128)t.idx(blk)[d];
		return key;
	};
	struct Cand
	{
		u128 okey, ckey;
		i64 a, b;
	};
	std::vector<Cand> cands;
	// device matching (qtb_match.cu) when asked for, or by size; needs every key below 2^63
	bool on_device = false;
	{
		const int mode = ctx.planner_mode;
		const bool want = mode == 1 || (mode < 0 && a.nblocks + b.nblocks >= 8192);
		u128 rc = 1, rfa = 1, rfb = 1;
		for (i64 c = 0; c < k; ++c)
			rc *= (u128)std::max<i64>(b.st.nsec[dims_b[c]], 1);
		for (auto d : free_a)
			rfa *= (u128)std::max<i64>(a.st.nsec[d], 1);
		for (auto d : free_b)
			rfb *= (u128)std::max<i64>(b.st.nsec[d], 1);
		const u128 lim = (u128)1 << 62;
		if (want && rc < lim && rfa * rfb < lim && rfa < lim && rfb < lim && a.nblocks < (i64(1) << 30) && b.nblocks < (i64(1) << 30))
		{
			std::vector<unsigned long long> a_ck(a.nblocks), a_fk(a.nblocks), b_ck(b.nblocks), b_fk(b.nblocks);
			for (i64 i = 0; i < a.nblocks; ++i)
			{
				u128 ck = 0;
				for (i64 c = 0; c < k; ++c)
					ck = ck * (u128)std::max<i64>(b.st.nsec[dims_b[c]], 1) + (u128)a.idx(i)[dims_a[c]];
				a_ck[i] = (unsigned long long)ck;
				a_fk[i] = (unsigned long long)encode(a, i, free_a, 0);
			}
			for (i64 j = 0; j < b.nblocks; ++j)
			{
				b_ck[j] = (unsigned long long)encode(b, j, dims_b, 0);
				b_fk[j] = (unsigned long long)encode(b, j, free_b, 0);
			}
			std::vector<MatchRec> recs;
			if (device_match(ctx, a_ck, a_fk, b_ck, b_fk, (unsigned long long)rfb, recs))
			{
				on_device = true;
				cands.reserve(recs.size());
				for (auto &r : recs)
					cands.push_back({(u128)r.okey, (u128)r.ckey, (i64)r.a, (i64)r.b});
			}
		}
	}
	std::vector<std::pair<u128, i64>> b_sorted(on_device ? 0 : b.nblocks);
	for (i64 j = 0; j < b.nblocks && !on_device; ++j)
		b_sorted[j] = {encode(b, j, dims_b, 0), j};
	std::sort(b_sorted.begin(), b_sorted.end());
	for (i64 i = 0; i < a.nblocks && !on_device; ++i)
	{
		// contracted key in B's radices (section counts of contracted dims agree, checked above)
		u128 ck = 0;
		for (i64 c = 0; c < k; ++c)
			ck = ck * (u128)std::max<i64>(b.st.nsec[dims_b[c]], 1) + (u128)a.idx(i)[dims_a[c]];
		auto lo = std::lower_bound(b_sorted.begin(), b_sorted.end(), std::make_pair(ck, (i64)-1));
		if (lo == b_sorted.end() || lo->first != ck)
			continue;
		const u128 ka = encode(a, i, free_a, 0);
		for (auto it = lo; it != b_sorted.end() && it->first == ck; ++it)
			cands.push_back({encode(b, it->second, free_b, ka), ck, i, it->second});
	}
	if (!on_device)
		std::sort(cands.begin(), cands.end(),
		          [](const Cand &x, const Cand &y) { return x.okey != y.okey ? x.okey < y.okey : x.ckey < y.ckey; });

	lap(1);
	// ---- operand offset tables (the fused permute_bl, btensor.cpp:1843-1894) ----
	struct OpTab
	{
		int32_t r = -1, c = -1;
		int contig = 0;
		int32_t rs = -1, cs = -1; // affine strides of the two tables, -1 when a table is not affine
	};
	// merged (size, stride) description of the C-order flattening of `dimlist` of block blk: size-1 dims dropped, adjacent
	// dims merged when the outer stride equals inner stride x inner size. One entry (or none) = affine: the offset of
	// flat position m is m * stride and no table is needed (every block of the DMRG path with size-1 physical / MPO
	// sections). Returns -1 when the list does not collapse to one stride (the int32 tables are built then).
	auto affine_stride = [](const Tensor &t, i64 blk, const std::vector<i64> &dimlist) -> int32_t
	{
		i64 size = 1, stride = 0;
		bool have = false;
		for (auto d : dimlist)
		{
			const i64 n = t.dm(blk)[d], s = t.sd(blk)[d];
			if (n == 1)
				continue;
			if (n == 0)
				return 0;
			if (!have)
			{
				size = n;
				stride = s;
				have = true;
			}
			else if (stride == s * n)
			{ // outer dim walks exactly one inner extent per step: merge
				size *= n;
				stride = s;
			}
			else
				return -1;
		}
		if (stride < 0 || stride * std::max<i64>(size - 1, 0) >= (i64(1) << 31))
			return -1;
		return (int32_t)stride;
	};
	auto flat_extent = [](const Tensor &t, i64 blk, const std::vector<i64> &dimlist)
	{
		i64 n = 1;
		for (auto d : dimlist)
			n *= t.dm(blk)[d];
		return n;
	};
	std::vector<OpTab> atab(a.nblocks), btab(b.nblocks);
	auto push_pool = [&](const std::vector<int32_t> &v)
	{
		const int32_t pos = (int32_t)plan->offpool.size();
		QTB_REQUIRE(plan->offpool.size() + v.size() < (size_t(1) << 31), QTB_ERR_INVALID_ARGUMENT,
		            "operand offset tables exceed 2^31 entries");
		plan->offpool.insert(plan->offpool.end(), v.begin(), v.end());
		return pos;
	};
	// one operand: `outer` = the free dims (rows of A / columns of B), `inner` = the contracted dims
	auto make_tab = [&](const Tensor &t, i64 blk, const std::vector<i64> &outer, const std::vector<i64> &inner, OpTab &tb,
	                    bool a_side)
	{
		const int32_t os = affine_stride(t, blk, outer), ks = affine_stride(t, blk, inner);
		const i64 no = flat_extent(t, blk, outer), nk = flat_extent(t, blk, inner);
		// which direction is unit stride in memory (decides the shared-memory layout the producer fills)
		const bool k_unit = nk > 1 && ks == 1, o_unit = no > 1 && os == 1;
		if (a_side)
			tb.contig = (nk > 1) ? k_unit : !o_unit; // 1: k is the unit-stride direction of A
		else
			tb.contig = (no > 1) ? o_unit : !k_unit; // 1: n is the unit-stride direction of B
		if (os >= 0 && ks >= 0)
		{
			tb.rs = a_side ? os : ks;
			tb.cs = a_side ? ks : os;
			tb.r = tb.c = 0; // affine: the kernel never touches the pool
			return;
		}
		auto oo = flat_offsets(t, blk, outer);
		auto ko = flat_offsets(t, blk, inner);
		tb.rs = tb.cs = -1;
		tb.r = push_pool(a_side ? oo : ko);
		tb.c = push_pool(a_side ? ko : oo);
	};
	auto a_tab = [&](i64 i) -> OpTab &
	{
		OpTab &t = atab[i];
		if (t.r < 0)
			make_tab(a, i, free_a, dims_a, t, true);
		return t;
	};
	auto b_tab = [&](i64 j) -> OpTab &
	{
		OpTab &t = btab[j];
		if (t.r < 0)
			make_tab(b, j, free_b, dims_b, t, false);
		return t;
	};

	// ---- group into output blocks ----
	size_t pos = 0;
	std::vector<i64> Ms, Ns;
	while (pos < cands.size())
	{
		size_t end = pos + 1;
		while (end < cands.size() && cands[end].okey == cands[pos].okey)
			++end;
		for (auto d : free_a)
			out.index.push_back(a.idx(cands[pos].a)[d]);
		for (auto d : free_b)
			out.index.push_back(b.idx(cands[pos].b)[d]);
		GemmOut go{};
		i64 M = 1, N = 1;
		for (auto d : free_a)
			M *= a.dm(cands[pos].a)[d];
		for (auto d : free_b)
			N *= b.dm(cands[pos].b)[d];
		go.M = (int32_t)M;
		go.N = (int32_t)N;
		go.pair_begin = (int32_t)plan->pairs.size();
		for (size_t p = pos; p < end; ++p)
		{
			const i64 i = cands[p].a, j = cands[p].b;
			i64 Ka = 1, Kb = 1;
			for (i64 c = 0; c < k; ++c)
			{
				Ka *= a.dm(i)[dims_a[c]];
				Kb *= b.dm(j)[dims_b[c]];
			}
			// the reference surfaces a size mismatch as a torch::mm shape error (c10::Error)
			QTB_REQUIRE(Ka == Kb, QTB_ERR_CHECK,
			            "mat1 and mat2 shapes cannot be multiplied (contracted section sizes differ)");
			if (Ka == 0 || M == 0 || N == 0)
				continue;
			GemmPair gp{};
			gp.a_off = a.offs[i];
			gp.b_off = b.offs[j];
			gp.K = (int32_t)Ka;
			OpTab &ta = a_tab(i);
			OpTab &tb = b_tab(j);
			gp.a_roff = ta.r;
			gp.a_koff = ta.c;
			gp.a_kcontig = ta.contig;
			gp.b_koff = tb.r;
			gp.b_coff = tb.c;
			gp.b_ncontig = tb.contig;
			const bool a_aff = ta.rs >= 0 && ta.cs >= 0, b_aff = tb.rs >= 0 && tb.cs >= 0;
			gp.a_rs = a_aff ? ta.rs : -1;
			gp.a_ks = a_aff ? ta.cs : -1;
			gp.b_ks = b_aff ? tb.rs : -1;
			gp.b_cs = b_aff ? tb.cs : -1;
			// bulk-copy staging: the unit-stride runs of an affine operand go through cp.async.bulk (16-byte aligned source:
			// a run starting on an odd element is fetched from one element earlier, the consumers undo the shift)
			static const bool bulk_on = !(std::getenv("QTB_BULK") && std::atoi(std::getenv("QTB_BULK")) == 0);
			gp.shf = 0;
			if (bulk_on && a_aff && ((gp.a_kcontig && gp.a_ks == 1) || (!gp.a_kcontig && gp.a_rs == 1)))
				gp.shf |= 1 | (int32_t)((gp.a_off & 1) << 2) | (((gp.a_kcontig ? gp.a_rs : gp.a_ks) & 1) << 3);
			if (bulk_on && b_aff && ((gp.b_ncontig && gp.b_cs == 1) || (!gp.b_ncontig && gp.b_ks == 1)))
				gp.shf |= 2 | (int32_t)((gp.b_off & 1) << 4) | (((gp.b_ncontig ? gp.b_ks : gp.b_cs) & 1) << 5);
			plan->pairs.push_back(gp);
			plan->flops += 2 * M * N * Ka;
		}
		go.pair_end = (int32_t)plan->pairs.size();
		{
			i64 ksum = 0;
			for (int p = go.pair_begin; p < go.pair_end; ++p)
				ksum += plan->pairs[p].K;
			plan->out_flops.push_back(2 * M * N * ksum);
		}
		plan->outs.push_back(go);
		Ms.push_back(M);
		Ns.push_back(N);
		pos = end;
	}
	out.nblocks = (i64)plan->outs.size();
	out.dims_from_structure();
	plan->out_numel = out.layout_packed();
	for (i64 ob = 0; ob < out.nblocks; ++ob)
	{
		QTB_REQUIRE(out.block_numel(ob) == Ms[ob] * Ns[ob], QTB_ERR_CHECK,
		            "shape mismatch between the operand blocks and the output sections");
		plan->outs[ob].c_off = out.offs[ob];
	}
	out.compute_hash();

	lap(2);
	// ---- tiling ----
	auto count_tiles = [&](int bm, int bn, double &padded)
	{
		i64 n = 0;
		padded = 0;
		for (auto &o : plan->outs)
		{
			if (o.pair_end == o.pair_begin)
				continue;
			i64 ksum = 0;
			for (int p = o.pair_begin; p < o.pair_end; ++p)
				ksum += plan->pairs[p].K;
			const i64 tm = (o.M + bm - 1) / bm, tn = (o.N + bn - 1) / bn;
			n += tm * tn;
			padded += 2.0 * tm * bm * tn * bn * ksum;
		}
		return n;
	};
	double pad64 = 0, pad128 = 0;
	const i64 n64 = count_tiles(64, 64, pad64);
	const i64 n128 = count_tiles(128, 128, pad128);
	(void)n64;
	// 128x128 tiles when they fill the machine and do not waste much on block edges
	plan->tile_cfg = (n128 >= ctx.sm_count && pad128 <= 1.25 * pad64) ? 1 : 0;
	// Sharded contexts: a rank runs 1 / world of the tiles. When that share is only a few tiles per SM, 128 x 128 tiles
	// quantise badly (H_eff.psi at D=4096 on 8 ranks: 93 tiles for 148 SMs in the first contraction, 561 us against
	// 330 ideal; profiles/r2/s62.txt) while the 64 x 64 configuration runs the same large-K products at the same rate
	// (T2: 30.3 TFLOP/s with either, s63.txt) with four times the tiles. The arithmetic per output element is the same
	// sequence of DMMAs in both configurations, so the result stays bit-identical to the single-GPU run.
	static const bool rule_single = std::getenv("QTB_TILE_RULE") && std::atoi(std::getenv("QTB_TILE_RULE")) == 1; // experiment
	if (plan->tile_cfg == 1 && (ctx.world > 1 || rule_single) && n128 / ctx.world < 4 * (i64)ctx.sm_count)
		plan->tile_cfg = 0;
	if (const char *force = std::getenv("QTB_TILE")) // experiment switch: 64 / 128
		plan->tile_cfg = std::atoi(force) == 128 ? 1 : 0;
	// skinny: every product is [M x K].[K x N] with K, N <= 16 (contraction with the MPO): CUDA-core row kernel
	{
		int max_n = 0, max_k = 0;
		i64 max_m = 0;
		for (auto &o : plan->outs)
			if (o.pair_end > o.pair_begin)
			{
				max_n = std::max(max_n, (int)o.N);
				max_m = std::max<i64>(max_m, o.M);
			}
		for (auto &pr : plan->pairs)
			max_k = std::max(max_k, (int)pr.K);
		plan->max_n = max_n;
		if (max_n <= 16 && max_k <= 16 && max_m >= 64)
			plan->tile_cfg = 2;
	}
	const int bm = plan->tile_cfg == 2 ? kSkinnyRows : (plan->tile_cfg ? 128 : 64);
	const int bn = plan->tile_cfg == 2 ? (1 << 30) : bm;
	// cost model of a tile (cycles of its busiest consumer warp): per K chunk a fixed part (barrier wait, fragment address
	// set-up) + 16 cycles per DMMA.8x8x4 of a whole warp tile; plus the epilogue.
	const int BKc = 16;
	const int wm = plan->tile_cfg == 1 ? 64 : 32, wn = 32; // warp tile of the two tensor-core configurations
	for (size_t ob = 0; ob < plan->outs.size(); ++ob)
	{
		auto &o = plan->outs[ob];
		if (o.pair_end == o.pair_begin)
			continue;
		i64 nchunks = 0, ksum = 0;
		for (int p = o.pair_begin; p < o.pair_end; ++p)
		{
			nchunks += (plan->pairs[p].K + BKc - 1) / BKc;
			ksum += plan->pairs[p].K;
		}
		for (int m0 = 0; m0 < o.M; m0 += (plan->tile_cfg == 2 ? std::max(o.M, 1) : bm))
			for (int n0 = 0; n0 < o.N; n0 += bn)
			{
				double c;
				if (plan->tile_cfg == 2) // ONE record per output block: the kernel cuts it into kSkinnyRows-row work items itself
					c = 200.0 + (double)o.M / 256.0 * (double)ksum * (4.0 + o.N);
				else
				{
					if (plan->tile_cfg == 0)
					{ // 64 x 64: the four warps share the valid 8x8 atoms of the tile (gemm_warp_grid)
						int gm, gn, am, an;
						gemm_warp_grid(std::min(8, (int)(o.M - m0 + 7) / 8), std::min(8, (int)(o.N - n0 + 7) / 8), gm, gn, am, an);
						static const double cT0 = std::getenv("QTB_COST_T0") ? std::atof(std::getenv("QTB_COST_T0")) : 800.0;
						static const double cC0 = std::getenv("QTB_COST_C0") ? std::atof(std::getenv("QTB_COST_C0")) : 150.0;
						static const double cA = std::getenv("QTB_COST_A") ? std::atof(std::getenv("QTB_COST_A")) : 16.0;
						c = cT0 + (double)nchunks * (cC0 + cA * (BKc / 4) * am * an);
					}
					else // 128 x 128: the busiest consumer warp always computes its whole warp tile (unpredicated path)
						c = 800.0 + (double)nchunks * (150.0 + 16.0 * (BKc / 4) * (wm / 8) * (wn / 8));
				}
				{
					const GemmPair &p0 = plan->pairs[o.pair_begin];
					plan->tiles.push_back(GemmTile{o.c_off, o.M, o.N, m0, n0, o.pair_begin, o.pair_end, (int32_t)ob, p0.K,
					                               p0.a_kcontig | (p0.b_ncontig << 1), p0.shf});
				}
				plan->tile_cost.push_back(c);
			}
	}
	// Bulk (UBLKCP) staging pays with 128 x 128 tiles (runs of 1 KB); with 64 x 64 tiles the runs are 128-512 bytes and the
	// per-request cost of the TMA unit makes it slower than LDGSTS (configs[1]: 81 us against 75 us, profiles/r2/s17.txt)
	// With 64 x 64 tiles the same shifted layout is filled by 16-byte LDGSTS instead (QTB_VEC16=0: plain 8-byte copies).
	static const bool vec16_on = !(std::getenv("QTB_VEC16") && std::atoi(std::getenv("QTB_VEC16")) == 0);
	if (plan->tile_cfg == 2 || (plan->tile_cfg == 0 && !vec16_on))
	{
		for (auto &pr : plan->pairs)
			pr.shf = 0;
		for (auto &t : plan->tiles)
			t.shf0 = 0;
	}
	plan->ncta = std::max(1, std::min<int>((int)plan->tiles.size(), gemm_grid_limit(ctx, plan->tile_cfg)));
	if (plan->tile_cfg == 2)
		plan->cta_begin = skinny_item_prefix(plan->tiles); // work-item prefix sums: item w belongs to the block t with prefix[t] <= w
	else
		plan->cta_begin = schedule_tiles(plan->tiles, plan->tile_cost, plan->ncta);

	lap(3);
	// ---- upload ----
	auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
	const size_t s_outs = align(plan->outs.size() * sizeof(GemmOut));
	const size_t s_pairs = align(plan->pairs.size() * sizeof(GemmPair));
	const size_t s_tiles = align(plan->tiles.size() * sizeof(GemmTile)) + align(plan->cta_begin.size() * sizeof(int32_t));
	const size_t s_pool = align(plan->offpool.size() * sizeof(int32_t));
	const size_t total = s_outs + s_pairs + s_tiles + s_pool + 256;
	std::vector<char> blob(total, 0);
	std::memcpy(blob.data(), plan->outs.data(), plan->outs.size() * sizeof(GemmOut));
	std::memcpy(blob.data() + s_outs, plan->pairs.data(), plan->pairs.size() * sizeof(GemmPair));
	std::memcpy(blob.data() + s_outs + s_pairs, plan->tiles.data(), plan->tiles.size() * sizeof(GemmTile));
	std::memcpy(blob.data() + s_outs + s_pairs + align(plan->tiles.size() * sizeof(GemmTile)), plan->cta_begin.data(),
	            plan->cta_begin.size() * sizeof(int32_t));
	std::memcpy(blob.data() + s_outs + s_pairs + s_tiles, plan->offpool.data(), plan->offpool.size() * sizeof(int32_t));
	plan->d_blob = plan_upload(ctx, *plan, blob.data(), total);
	char *base = (char *)plan->d_blob;
	plan->d_outs = (GemmOut *)base;
	plan->d_pairs = (GemmPair *)(base + s_outs);
	plan->d_tiles = (GemmTile *)(base + s_outs + s_pairs);
	plan->d_cta_begin = (int32_t *)(base + s_outs + s_pairs + align(plan->tiles.size() * sizeof(GemmTile)));
	plan->d_offpool = (int32_t *)(base + s_outs + s_pairs + s_tiles);
	plan->d_counter = (int *)(base + s_outs + s_pairs + s_tiles + s_pool);
	ctx.counters[2] += 1;
	lap(4);
	return plan;
}

std::shared_ptr<Plan> get_plan(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a,
                               const std::vector<i64> &dims_b)
{
	uint64_t h = mix64(a.layout_hash, b.layout_hash * 3 + 1);
	h = mix64(h, dims_a.size());
	for (auto d : dims_a)
		h = mix64(h, (uint64_t)d + 17);
	for (auto d : dims_b)
		h = mix64(h, (uint64_t)d + 91);
	// The cache is keyed by a 128-bit digest: `h` indexes the map, and a hit is only accepted when an independently mixed
	// second digest of the same layouts (structure, charge type, block tables, offsets, dim lists) agrees as well.
	h = mix64(h, (uint64_t)ctx.world + 977); // the tile configuration depends on the number of ranks
	uint64_t h2 = mix64b(a.layout_hash2, b.layout_hash2 * 0x9e3779b97f4a7c15ull + 7);
	h2 = mix64b(h2, dims_a.size());
	for (auto d : dims_a)
		h2 = mix64b(h2, (uint64_t)d + 1);
	for (auto d : dims_b)
		h2 = mix64b(h2, (uint64_t)d + 1001);
	auto it = ctx.plan_cache.find(h);
	if (it != ctx.plan_cache.end() && it->second->key2 == h2)
	{
		ctx.counters[3] += 1;
		return it->second;
	}
	if (ctx.plan_cache.size() > 8192)
		ctx.plan_cache.clear();
	auto p = build_plan(ctx, a, b, dims_a, dims_b);
	p->key2 = h2;
	// adopted operands (qtb_tensor_adopt): the block offsets are the pointer deltas of separately allocated caller
	// blocks, different at every call — such a plan would never be hit again, so it is not kept
	const bool adopted = (a.arena && !a.arena->owned) || (b.arena && !b.arena->owned);
	if (!adopted)
		ctx.plan_cache[h] = p;
	return p;
}

void Ctx::prof_dump(const char *title)
{
	std::fprintf(stderr, "[qtb profile] %s\n", title);
	if (prof_level >= 3)
	{
		std::fprintf(stderr, "[qtb profile]   plan phases (ms): validate+structure %.1f, matching %.1f, pairs+tables %.1f, tiling+schedule %.1f, upload %.1f\n",
		             plan_phase_ms[0], plan_phase_ms[1], plan_phase_ms[2], plan_phase_ms[3], plan_phase_ms[4]);
		for (double &v : plan_phase_ms)
			v = 0;
	}
	for (auto &kv : prof)
		std::fprintf(stderr, "[qtb profile]   %-28s calls %6ld built %5ld plan %9.2f ms gemm %9.2f ms  %8.3f GFLOP  %7.2f TFLOP/s tiles %ld\n",
		             kv.first.c_str(), (long)kv.second.calls, (long)kv.second.built, kv.second.plan_ms, kv.second.gemm_ms,
		             kv.second.flops * 1e-9, kv.second.gemm_ms > 0 ? kv.second.flops / kv.second.gemm_ms * 1e-9 : 0.0,
		             (long)kv.second.tiles);
	prof.clear();
}

std::unique_ptr<Tensor> tensordot(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a,
                                  const std::vector<i64> &dims_b)
{
	std::chrono::steady_clock::time_point t0;
	const i64 built0 = ctx.counters[2];
	if (ctx.prof_level >= 2)
	{
		cudaStreamSynchronize(ctx.stream);
		t0 = std::chrono::steady_clock::now();
	}
	auto plan = get_plan(ctx, a, b, dims_a, dims_b);
	auto out = std::make_unique<Tensor>(plan->out_proto);
	out->arena = std::make_shared<Arena>(&ctx, plan->out_numel);
	std::chrono::steady_clock::time_point t1a;
	if (ctx.prof_level >= 2)
	{
		cudaStreamSynchronize(ctx.stream); // plan upload done: what follows is the kernel alone
		t1a = std::chrono::steady_clock::now();
	}
	bool any_empty = false;
	for (auto &o : plan->outs)
		any_empty |= (o.pair_begin == o.pair_end);
	if (any_empty && plan->out_numel)
		QTB_CUDA(cudaMemsetAsync(out->arena->ptr, 0, plan->out_numel * sizeof(double), ctx.stream));
	launch_grouped_gemm(ctx, *plan, a.arena ? a.arena->ptr : nullptr, b.arena ? b.arena->ptr : nullptr,
	                    out->arena->ptr);
	if (ctx.prof_level >= 2)
	{
		cudaStreamSynchronize(ctx.stream);
		auto t2 = std::chrono::steady_clock::now();
		std::string key = "r" + std::to_string(a.st.rank) + "xr" + std::to_string(b.st.rank) + " k" + std::to_string(dims_a.size()) +
		                  " a" + (dims_a.empty() ? std::string("-") : std::to_string(dims_a[0])) + " cfg" + std::to_string(plan->tile_cfg);
		auto &r = ctx.prof[key];
		r.calls += 1;
		r.built += ctx.counters[2] - built0;
		r.flops += plan->flops;
		r.tiles += (i64)plan->tiles.size();
		r.plan_ms += std::chrono::duration<double, std::milli>(t1a - t0).count(); // plan build + upload + arena allocation
		r.gemm_ms += std::chrono::duration<double, std::milli>(t2 - t1a).count();
	}
	return out;
}

void tensordot_into(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a,
                    const std::vector<i64> &dims_b, Tensor &out)
{
	auto plan = get_plan(ctx, a, b, dims_a, dims_b);
	QTB_REQUIRE(out.layout_hash == plan->out_proto.layout_hash && out.arena && out.arena->numel >= plan->out_numel,
	            QTB_ERR_INVALID_ARGUMENT, "tensordot_into: the output tensor does not have the planned layout");
	launch_grouped_gemm(ctx, *plan, a.arena->ptr, b.arena->ptr, out.arena->ptr);
}

} // namespace qtb
