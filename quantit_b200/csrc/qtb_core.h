// qtb_core.h — host-side data model of the engine: packed block tensors living in one device arena.
//
// Replaces the storage of quantit::btensor (reference include/blockTensor/btensor.h:105-110,821-837: a sorted vector
// of (block_index, torch::Tensor) pairs, one heap tensor per block) by
//   * ONE device allocation ("arena") per tensor, shared (ref-counted) between a tensor and its permute/conj views;
//   * a host block table {index[rank], dims[rank], strides[rank], offset} sorted lexicographically by index
//     (reference flat_map.h:31,151), so permute/conj are metadata-only exactly like the torch views the reference
//     creates (btensor.cpp:1781-1782, 2156-2172) and the copy is fused into the next contraction's operand load.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/qtb.h"

namespace qtb
{
using i64 = int64_t;

// ---- errors: C++ exceptions inside the library, translated to qtb_status at the ABI ------------------------------
struct Error : std::runtime_error
{
	qtb_status code;
	Error(qtb_status c, const std::string &m) : std::runtime_error(m), code(c) {}
};
#define QTB_CUDA(call)                                                                                                 \
	do                                                                                                                 \
	{                                                                                                                  \
		cudaError_t e__ = (call);                                                                                      \
		if (e__ != cudaSuccess)                                                                                        \
			throw ::qtb::Error(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? QTB_ERR_NO_DEVICE      \
			                                                                                   : QTB_ERR_CUDA,         \
			                   std::string(#call) + ": " + cudaGetErrorString(e__));                                   \
	} while (0)
#define QTB_REQUIRE(cond, code, msg)                                                                                   \
	do                                                                                                                 \
	{                                                                                                                  \
		if (!(cond))                                                                                                   \
			throw ::qtb::Error(code, msg);                                                                             \
	} while (0)

struct Ctx;

// ---- device arena ------------------------------------------------------------------------------------------------
struct Arena
{
	double *ptr = nullptr;
	i64 numel = 0;
	bool owned = true;
	Ctx *ctx = nullptr;
	Arena(Ctx *c, i64 n); // stream-ordered allocation from the context's pool
	Arena(double *p) : ptr(p), owned(false) {}
	~Arena();
	Arena(const Arena &) = delete;
	Arena &operator=(const Arena &) = delete;
};

// ---- charges -----------------------------------------------------------------------------------------------------
// nc-component integer tuples stored flat; mods[c]==0 -> Z (int16 in the reference), N -> C<N>.
struct ChargeType
{
	i64 nc = 0;
	std::vector<i64> mods;
	bool operator==(const ChargeType &o) const { return nc == o.nc && mods == o.mods; }
	i64 norm(i64 v, i64 c) const
	{
		i64 m = mods[c];
		if (m == 0)
			return v;
		v %= m;
		return v < 0 ? v + m : v;
	}
};

// ---- structure (what the reference calls the "shape" of a btensor) ------------------------------------------------
struct Structure
{
	i64 rank = 0;
	ChargeType ct;
	std::vector<i64> nsec;      // [rank]
	std::vector<i64> sec_off;   // [rank+1] prefix sums of nsec
	std::vector<i64> sec_sizes; // [sum nsec]
	std::vector<i64> cvals;     // [sum nsec * nc]
	std::vector<i64> sel;       // [nc]

	void finalize(); // builds sec_off
	i64 total_sections() const { return sec_off.empty() ? 0 : sec_off.back(); }
	i64 size_of(i64 dim, i64 sec) const { return sec_sizes[sec_off[dim] + sec]; }
	const i64 *charge_of(i64 dim, i64 sec) const { return &cvals[(sec_off[dim] + sec) * ct.nc]; }
	bool allowed(const i64 *index) const; // block_conservation_rule_test, btensor.cpp:333-345
	i64 dim_size(i64 dim) const;
};

struct Block
{
	i64 off = 0;    // element offset of the block's first element from the arena base
	i64 tab = 0;    // position of index/dims/strides in the flat tables (= block number * rank)
};

// ---- tensor ------------------------------------------------------------------------------------------------------
struct Tensor
{
	Structure st;
	i64 nblocks = 0;
	std::vector<i64> index;   // [nblocks*rank] sorted lexicographically
	std::vector<i64> dims;    // [nblocks*rank]
	std::vector<i64> strides; // [nblocks*rank] in elements
	std::vector<i64> offs;    // [nblocks] element offsets from arena->ptr
	std::shared_ptr<Arena> arena;
	uint64_t layout_hash = 0; // structure-only identity (index/dims/strides/offs): the plan-cache key component
	uint64_t layout_hash2 = 0; // an independent second hash of the same data: cache hits compare both (128-bit key)

	i64 rank() const { return st.rank; }
	const i64 *idx(i64 b) const { return &index[b * st.rank]; }
	const i64 *dm(i64 b) const { return &dims[b * st.rank]; }
	const i64 *sd(i64 b) const { return &strides[b * st.rank]; }
	i64 block_numel(i64 b) const
	{
		i64 n = 1;
		for (i64 d = 0; d < st.rank; ++d)
			n *= dims[b * st.rank + d];
		return n;
	}
	i64 numel() const
	{
		i64 n = 0;
		for (i64 b = 0; b < nblocks; ++b)
			n += block_numel(b);
		return n;
	}
	bool block_contiguous(i64 b) const;
	bool packed_canonical() const; // every block C-contiguous (the layout produced by allocate_packed)
	void compute_hash();
	// lay the blocks out back to back, C-contiguous; fills strides/offs and
	// returns the arena size in elements. dims must be set.
	i64 layout_packed();
	void dims_from_structure(); // dims[b] = section sizes of index[b]
	i64 find_block(const i64 *index_) const; // -1 if absent
};

constexpr i64 kBlockAlign = 1; // packed blocks sit back to back: the arena image equals the caller's flat buffer, so
                                // host<->device transfers are ONE DMA each (cp.async needs only 8-byte alignment)

// sorts blocks lexicographically; returns the permutation applied (new position -> old position)
std::vector<i64> sort_blocks(i64 rank, std::vector<i64> &index);

// ---- contraction plan ----------------------------------------------------------------------------------------------
// 64 x 64 configuration: how the four consumer warps share the 8 x 8 grid of 8x8 MMA atoms of a tile that holds only
// mv x nv valid atoms (block edges). The warps form a gm x gn grid, (2,2), (1,4) or (4,1), each owning am x an atoms
// (am, an <= 4); the shape with the fewest atoms on the busiest warp wins. A full tile is (2,2) with 4 x 4 atoms each.
// Used by the kernel (qtb_gemm.cu) and by the planner's cost model, which must agree.
#if defined(__CUDACC__)
#define QTB_HD __host__ __device__
#else
#define QTB_HD
#endif
QTB_HD inline void gemm_warp_grid(int mv, int nv, int &gm, int &gn, int &am, int &an)
{
	gm = 2;
	gn = 2;
	am = (mv + 1) >> 1;
	an = (nv + 1) >> 1;
	int best = am * an;
	if (mv <= 4)
	{ // one row of warps
		const int a = mv, b = (nv + 3) >> 2;
		if (a * b < best)
		{
			best = a * b;
			gm = 1;
			gn = 4;
			am = a;
			an = b;
		}
	}
	if (nv <= 4)
	{ // one column of warps
		const int a = (mv + 3) >> 2, b = nv;
		if (a * b < best)
		{
			gm = 4;
			gn = 1;
			am = a;
			an = b;
		}
	}
}

// one candidate pair of the device matching (qtb_match.cu), sorted by (okey, ckey); head = 1: first pair of an output block
struct MatchRec
{
	unsigned long long okey, ckey;
	int32_t a, b;
	int32_t head, pad;
};
static_assert(sizeof(MatchRec) == 32, "MatchRec is copied back as packed 32-byte records");

struct GemmTile
{ // one work item of the grouped GEMM: a tile of one output block, self-contained (the kernels fetch ONE descriptor per
  // item, no dependent second fetch of the block's record)
	i64 c_off;                    // element offset of the output block in the C arena
	int32_t M, N;                 // matrix dims of the output block
	int32_t m0, n0;               // tile origin inside the block matrix
	int32_t pair_begin, pair_end; // matched pairs of the block (ascending contracted index)
	int32_t out_blk;              // which output block (sharding filters on it)
	int32_t K0, flags0;           // K and (a_kcontig | b_ncontig << 1) of the first pair: a single-pair block (every block of
	                              // configs[1]) needs no pair-descriptor fetch on the consumer side at all
	int32_t shf0;                 // GemmPair::shf of the first pair
};
struct GemmOut
{
	i64 c_off;       // element offset of the output block in the C arena
	int32_t M, N;    // matrix dims of the output block
	int32_t pair_begin, pair_end;
};
struct GemmPair
{
	i64 a_off, b_off;        // element offsets of the operand blocks in their arenas
	int32_t K;               // contracted extent
	int32_t a_roff, a_koff;  // positions in the int32 offset pool: row offsets [M], k offsets [K] of the A block
	int32_t b_koff, b_coff;  // k offsets [K], column offsets [N] of the B block
	int32_t a_kcontig;       // 1: k is the unit-stride direction of A (else rows are)
	int32_t b_ncontig;       // 1: n is the unit-stride direction of B (else k is)
	int32_t a_rs, a_ks;      // affine fast path: offset(m,k) = m*a_rs + k*a_ks   (a_rs < 0: use the tables)
	int32_t b_ks, b_cs;      // affine fast path: offset(k,n) = k*b_ks + n*b_cs   (b_cs < 0: use the tables)
	int32_t shf;             // bulk-copy staging (cp.async.bulk, qtb_gemm.cu): bit 0 / 1: operand A / B is staged by 16-byte
	                         // aligned bulk copies of its unit-stride runs; bit 2 / 4: parity of a_off / b_off; bit 3 / 5:
	                         // parity of the operand's non-unit stride. A run that starts on an odd element is copied from
	                         // one element earlier and read back with a one-element shift derived from these parities.
};
static_assert(sizeof(GemmPair) == 64, "GemmPair is uploaded as a packed 64-byte record");

constexpr int kSkinnyRows = 1024; // rows of an output block per work item of the skinny (MPO) contraction kernel

// Plan tables live in large device slabs carved linearly (a cudaMalloc per plan — ~1000 plans per saturated sweep, no
// two of the same size — cost 0.2 ms each and was the largest part of the planner's host time); a slab goes back to the
// driver when the last plan that points into it is destroyed.
struct PlanSlab
{
	char *base = nullptr;
	size_t size = 0, used = 0;
	~PlanSlab()
	{
		if (base)
			cudaFree(base);
	}
};

struct Plan
{
	// output layout
	Tensor out_proto; // structure + block table + packed layout, no arena
	i64 out_numel = 0;
	// work
	std::vector<GemmOut> outs;
	std::vector<GemmPair> pairs;
	std::vector<GemmTile> tiles;     // grouped by CTA: CTA c runs tiles[cta_begin[c] .. cta_begin[c+1]), heaviest first
	std::vector<int32_t> cta_begin;  // [ncta + 1]
	std::vector<double> tile_cost;   // [tiles] modelled cycles (same order as tiles)
	int ncta = 0;
	std::vector<int32_t> offpool;
	i64 flops = 0;
	uint64_t key2 = 0; // second half of the cache key (see get_plan)
	int tile_cfg = 0; // 0: 64x64 tiles, 1: 128x128 tiles, 2: skinny (one thread per output row, N and K <= 16)
	int max_n = 0;    // largest N over the output blocks
	// device copies
	void *d_blob = nullptr;
	GemmOut *d_outs = nullptr;
	GemmPair *d_pairs = nullptr;
	GemmTile *d_tiles = nullptr;
	int32_t *d_cta_begin = nullptr;
	int32_t *d_offpool = nullptr;
	int *d_counter = nullptr;
	Ctx *ctx = nullptr;
	// charge-sector sharding: the tiles of the output blocks one rank owns, keyed by a hash of (owner dim, owner map, rank)
	struct Owned
	{
		GemmTile *d_tiles = nullptr; // the tiles of this rank, grouped by CTA (then int32 cta_begin[ncta + 1])
		int32_t *d_cta_begin = nullptr;
		int ntiles = 0, ncta = 0;
		i64 flops = 0;
	};
	std::unordered_map<uint64_t, Owned> owned;
	std::vector<std::shared_ptr<PlanSlab>> slabs; // keep the slabs of d_blob and of the owned tile lists alive
	std::vector<i64> out_flops; // [out blocks] 2 M N sum K
	~Plan();
};

// ---- context -------------------------------------------------------------------------------------------------------
struct Ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaMemPool_t pool = nullptr;
	int sm_count = 148;
	i64 counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	int planner_mode = -1;  // block-pair matching: 0 host, 1 device (qtb_match.cu), -1 by size (qtb_ctx_set_device_planner)
	i64 device_matches = 0; // contractions whose pairs were matched by the device kernels
	std::unordered_map<uint64_t, std::shared_ptr<Plan>> plan_cache;
	std::shared_ptr<PlanSlab> plan_slab; // the slab new plan tables are carved from
	// live arenas: a context that is destroyed before its tensors orphans them (Arena::ctx = nullptr, the device block is
	// released with the context) so that a late qtb_tensor_free is harmless
	std::unordered_set<Arena *> arenas;
	// pinned ring for small asynchronous uploads (plan tables, descriptors), recycled after a stream synchronisation
	static constexpr int kRingParts = 4;
	char *ring_base = nullptr;
	size_t ring_size = 0, ring_pos = 0;
	int ring_part = 0;
	cudaEvent_t ring_event[kRingParts] = {nullptr, nullptr, nullptr, nullptr};
	bool ring_busy[kRingParts] = {false, false, false, false};
	// kernels whose function attributes (dynamic shared memory opt-in, carve-out) were set for this context's device
	uint32_t kernel_attr_mask = 0;
	bool attr_once(int bit)
	{ // true the first time `bit` is asked for
		if (kernel_attr_mask & (1u << bit))
			return false;
		kernel_attr_mask |= 1u << bit;
		return true;
	}
	// caching device allocator (qtb_core.cpp: ctx_alloc / ctx_free)
	std::map<size_t, std::vector<void *>> free_bins; // size class -> free blocks
	std::unordered_map<void *, size_t> live_blocks;  // block -> size class
	size_t cached_bytes = 0;
	void trim_cache(); // returns every cached block to the driver (synchronises the stream)
	// auxiliary streams / events (block SVD lanes) and a small pinned read-back area; created on demand
	std::vector<cudaStream_t> aux_streams;
	std::vector<cudaEvent_t> aux_events; // [aux_streams.size() + 1]
	unsigned long long *pinned_gauge_ = nullptr;
	void ensure_aux_streams(int n);
	unsigned long long *pinned_gauge(); // 64 pinned words
	// pinned staging for plan uploads / small downloads
	void *pinned = nullptr;
	size_t pinned_bytes = 0;
	void *pinned_buf(size_t bytes);
	// QTB_PROFILE=2 diagnostics: per contraction signature {calls, plan-build ms, kernel ms (synchronised), flops, tiles}
	int prof_level = 0;
	struct ProfRec
	{
		i64 calls = 0, flops = 0, tiles = 0, built = 0;
		double plan_ms = 0, gemm_ms = 0;
	};
	std::map<std::string, ProfRec> prof;
	double plan_phase_ms[5] = {0, 0, 0, 0, 0};
	void prof_dump(const char *title);
	// charge-sector sharding (qtb_ctx_set_sharding)
	int rank = 0, world = 1;
	qtb_allreduce_fn allreduce_fn = nullptr;
	void *allreduce_user = nullptr;
	void allreduce(double *ptr, i64 n); // in-place sum over ranks, stream ordered; no-op when world == 1
	void *nccl_comm = nullptr;          // the engine's own communicator (qtb_ctx_init_nccl): preferred over the callback
	~Ctx();
};

// ---- NCCL bound at run time (qtb_nccl.cpp) ----------------------------------------------------------------------------
struct OwnedRange
{
	i64 off, n; // element range of an arena
	int owner;  // rank that computed it
};
void nccl_unique_id(const char *libpath, char out[128]);
void ctx_init_nccl(Ctx &ctx, int rank, int world, const char id_bytes[128], const char *libpath);
void ctx_destroy_nccl(Ctx &ctx);
bool nccl_allreduce(Ctx &ctx, double *ptr, i64 n);
bool nccl_exchange_ranges(Ctx &ctx, double *base, const std::vector<OwnedRange> &ranges);
bool nccl_allgather(Ctx &ctx, double *base, i64 chunk);

// ---- ops (host orchestration; kernels in the .cu files) --------------------------------------------------------------
std::shared_ptr<Plan> get_plan(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a,
                               const std::vector<i64> &dims_b);
std::unique_ptr<Tensor> tensordot(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a,
                                  const std::vector<i64> &dims_b);
void tensordot_into(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &dims_a,
                    const std::vector<i64> &dims_b, Tensor &out);
std::unique_ptr<Tensor> permute(const Tensor &a, const std::vector<i64> &perm);
std::unique_ptr<Tensor> conj(const Tensor &a);
std::unique_ptr<Tensor> make_tensor(Ctx &ctx, const Structure &st, i64 nblocks, const i64 *block_index,
                                    const double *host_data);
void download(Ctx &ctx, const Tensor &t, double *host_out);
std::unique_ptr<Tensor> contiguous(Ctx &ctx, const Tensor &t); // packed copy (gathers strided views)

// kernels (qtb_gemm.cu / qtb_vec.cu)
// cin != nullptr: c = alpha * cin + beta * (a . b), cin laid out like the output (the tensorgdot epilogue)
// device block-pair matching (qtb_match.cu); false = input beyond what the kernels hold, use the host path
bool device_match(Ctx &ctx, const std::vector<unsigned long long> &a_ck, const std::vector<unsigned long long> &a_fk,
                  const std::vector<unsigned long long> &b_ck, const std::vector<unsigned long long> &b_fk,
                  unsigned long long rb, std::vector<MatchRec> &out);
void launch_grouped_gemm(Ctx &ctx, const Plan &plan, const double *a, const double *b, double *c,
                         const Plan::Owned *owned = nullptr, const double *cin = nullptr, double alpha = 0.0,
                         double beta = 1.0);
// static tile schedule: longest-processing-time-first assignment of the cost-modelled tiles to `ncta` CTAs; reorders
// `tiles`/`cost` so that the tiles of one CTA are contiguous (heaviest first) and returns cta_begin[ncta + 1]
std::vector<int32_t> schedule_tiles(std::vector<GemmTile> &tiles, std::vector<double> &cost, int ncta);
int gemm_grid_limit(const Ctx &ctx, int tile_cfg); // CTAs the grouped GEMM keeps resident (SM count x CTAs per SM)
// sharding helpers (qtb_core.cpp)
std::vector<int32_t> lpt_assign(const std::vector<double> &weights, int world);
std::vector<int32_t> skinny_item_prefix(const std::vector<GemmTile> &tiles);
// one step of a sharded chain of contractions: computes only the output blocks whose section along `owner_dim` belongs
// to this rank (the rest of the freshly allocated arena is zero); `owner` maps sections of that dim to ranks.
// zero_rest: clear the whole output arena first (needed when the result is summed over the ranks; an intermediate of a
// chain that is only read through the blocks this rank owns does not need it — owned blocks without any matched pair
// are cleared individually)
std::unique_ptr<Tensor> tensordot_owned(Ctx &ctx, const std::shared_ptr<Plan> &plan, const Tensor &a, const Tensor &b,
                                        i64 owner_dim, const std::vector<int32_t> &owner, bool zero_rest = true);
// same, writing every owned output block at c_off[block] of an existing arena (the caller's layout, not the plan's)
void tensordot_owned_into(Ctx &ctx, const std::shared_ptr<Plan> &plan, const Tensor &a, const Tensor &b, i64 owner_dim,
                          const std::vector<int32_t> &owner, const std::vector<i64> &c_off, double *c_arena);
// a chain of contractions sharing one owner leg: accumulates the planner's flops per section of that leg
void add_section_weights(const Plan &plan, i64 owner_dim, std::vector<double> &weights);
void launch_gather_blocks(Ctx &ctx, const Tensor &src, double *dst_packed, const std::vector<i64> &dst_offs);

} // namespace qtb
