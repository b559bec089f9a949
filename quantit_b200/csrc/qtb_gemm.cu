// qtb_gemm.cu — the grouped fp64 block GEMM of the engine (sm_100a).
//
// Replaces, in ONE launch per contraction, what the reference does with one torch op per block / block pair
// (reference sources/btensor.cpp): permute_bl's per-block permute+reshape copies (:1843-1894, call sites K1),
// torch::mm for the first matched pair of an output block (:2095, K2) and addmm_ for the remaining pairs (:2102, K3).
//
//  * work unit = one BMxBN tile of one output block; the CTA walks the block's matched pair list (ascending
//    contracted block index, like the reference's two-pointer merge) and keeps the accumulators in registers across
//    pairs: no HBM round trip between pairs (the reference's addmm_ reads+writes C once per pair).
//  * operands are read straight from the (possibly permuted / strided) source blocks: element (m,k) of an operand
//    matrix lives at base + roff(m) + koff(k). roff/koff are affine (m*rs, k*ks) for almost every block of the DMRG
//    path; otherwise two int32 tables built by the planner. Either way the permute+reshape copy the reference
//    materialises is fused into the operand load.
//  * fp64 tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — the only fp64 MMA shape sm_100a has; tcgen05 has no
//    fp64 kind).
//  * warp-specialised: one producer warpgroup issues the cp.async operand loads into a STAGES-deep shared-memory ring
//    (completion tracked by mbarriers, cp.async.mbarrier.arrive), the consumer warps only do LDS + DMMA, so the tensor
//    pipe is not starved by address arithmetic. The producer runs ahead across tile boundaries (it prefetches the next
//    tile's operands while the consumers finish the current one), which is what keeps the small-block regime
//    (hundreds of 50x50x50 GEMMs per contraction) from being pipeline-fill bound.
//  * tiles are assigned to the CTAs by the planner (longest-processing-time-first over a cycle model, each CTA's list
//    heaviest first); work items are self-contained records fetched ahead of use. At ragged block edges a warp tile that
//    intersects the block is computed whole (clamped rows, zero-filled K tail), a warp tile outside it is skipped.
#include <cuda_runtime.h>
#include <cstdlib>

#include <algorithm>
#include <cstdint>

#include "qtb_core.h"

namespace qtb
{

// ---- PTX helpers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async8(unsigned smem, const void *gmem, bool valid)
{
	int sz = valid ? 8 : 0; // src-size 0 -> the 8 destination bytes are zero-filled
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned bar)
{ // the arrival fires when every cp.async issued so far by this thread has landed
	asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{ // adds `bytes` to the transaction count of the current phase (no arrival)
	asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (SASS UBLKCP): 16-byte aligned source / destination, size a
// multiple of 16; completion is signalled on the mbarrier as `bytes` transaction bytes.
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void *gmem, unsigned bytes, unsigned bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
	             "l"(gmem), "r"(bytes), "r"(bar)
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
	asm volatile("{\n"
	             ".reg .pred P1;\n"
	             "LAB_WAIT:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
	             "@P1 bra DONE;\n"
	             "bra LAB_WAIT;\n"
	             "DONE:\n"
	             "}\n" ::"r"(bar),
	             "r"(parity)
	             : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_, int MINB_, bool REALLOC_>
struct GemmCfg
{
	static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
	static constexpr bool REALLOC = REALLOC_;
	static constexpr bool kVec16 = (BM_ == 64 && BN_ == 64); // shifted-layout operands: 16-byte LDGSTS (else the bulk path)
	static constexpr bool kEdge = (BM_ == 64 && BN_ == 64 && WM_ == 32 && WN_ == 32); // per-tile warp grid + atom-count variants
	static constexpr int kWarpsM = BM / WM;
	static constexpr int kWarpsN = BN / WN;
	static constexpr int kConsWarps = kWarpsM * kWarpsN;
	static constexpr int kProdThreads = 128; // one warpgroup
	static constexpr int kThreads = kProdThreads + kConsWarps * 32;
	static constexpr int kPad = 4;
	// an operand tile is stored either "k-major" [rows][BK+4] or "row-major" [BK][rows+4]; reserve the larger
	static constexpr int kASize = (BM * (BK + kPad) > BK * (BM + kPad)) ? BM * (BK + kPad) : BK * (BM + kPad);
	static constexpr int kBSize = (BN * (BK + kPad) > BK * (BN + kPad)) ? BN * (BK + kPad) : BK * (BN + kPad);
	static constexpr int kStage = kASize + kBSize;
	// + descriptor ring: kTileRing work-item records and two pair records, staged by the producer warpgroup
	static constexpr int kTileRing = 16;
	static constexpr size_t kDescOff = (size_t(STAGES) * kStage * sizeof(double) + 2 * STAGES * sizeof(uint64_t) + 15) & ~size_t(15);
	static constexpr size_t kSmemBytes = kDescOff + kTileRing * 48 + 2 * 64;
	static constexpr int kAPerThread = BM * BK / kProdThreads;
	static constexpr int kBPerThread = BN * BK / kProdThreads;
	static_assert(BM * BK % kProdThreads == 0 && BN * BK % kProdThreads == 0, "tile/threads mismatch");
	static_assert(BM % 16 == 0 && BN % 16 == 0 && BK % 4 == 0, "tile shape");
	static_assert(kConsWarps % 4 == 0, "consumer warps must form whole warpgroups");
};


// Producer side: one operand tile [R rows x BK] of a block matrix into shared memory.
//   KC (k-contiguous source): consecutive threads walk k  -> smem layout [R][BK+4]   (k-major)
//   else                    : consecutive threads walk r  -> smem layout [BK][R+4]   (row-major)
// element (r,k) lives at base + roff(r) + koff(k); AFF: r*rs + k*ks, else the planner's int32 tables.
// FULL: the tile lies entirely inside the block (no bounds checks, no zero fill) — the hot path: one address
// computation + one LDGSTS per element.
__device__ __forceinline__ void cp_async8_full(unsigned smem, const void *gmem)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem), "l"(gmem));
}

// offpar / strpar: the one-element shift of the bulk-staged layout (GemmPair::shf), 0 / 0 for operands that never take
// the bulk path: run `i` of the tile (a row for KC, a k index otherwise) is stored shifted by (offpar + i * strpar) & 1.
template <class Cfg, int R, bool AFF, bool KC, bool FULL>
__device__ __forceinline__ void load_tile(unsigned smem_dst, const double *__restrict__ base,
                                          const int32_t *__restrict__ rtab, const int32_t *__restrict__ ktab, int rs,
                                          int ks, int r0, int Rmax, int k0, int K, int pt, int offpar, int strpar)
{
	constexpr int BK = Cfg::BK, PAD = Cfg::kPad, NT = Cfg::kProdThreads;
	constexpr int PER = R * BK / NT;
	if constexpr (KC)
	{
		const int k = pt % BK, rbase = pt / BK;
		const int kg = k0 + k;
		const bool kok = FULL || (kg < K);
		const int kc = FULL ? kg : (kok ? kg : 0);
		const double *src_k = base + (AFF ? (int64_t)kc * ks : (int64_t)ktab[kc]);
		static_assert((NT / BK) % 2 == 0, "row parity must be thread-constant");
		const unsigned dst = smem_dst + (rbase * (BK + PAD) + k + ((offpar + (rbase & 1) * strpar) & 1)) * 8;
#pragma unroll
		for (int i = 0; i < PER; ++i)
		{
			int r = r0 + rbase + i * (NT / BK);
			if constexpr (!FULL)
				r = r < Rmax ? r : Rmax - 1; // rows past the block edge re-read the last row: results are never stored
			const double *src = src_k + (AFF ? (int64_t)r * rs : (int64_t)rtab[r]);
			if constexpr (FULL)
				cp_async8_full(dst + i * ((NT / BK) * (BK + PAD) * 8), src);
			else
				cp_async8(dst + i * ((NT / BK) * (BK + PAD) * 8), src, kok); // k tail is zero-filled
		}
	}
	else
	{
		const int r = pt % R, kbase = pt / R;
		int rg = r0 + r;
		if constexpr (!FULL)
			rg = rg < Rmax ? rg : Rmax - 1;
		const double *src_r = base + (AFF ? (int64_t)rg * rs : (int64_t)rtab[rg]);
		const unsigned dst = smem_dst + (kbase * (R + PAD) + r) * 8;
#pragma unroll
		for (int i = 0; i < PER; ++i)
		{
			const int kg = k0 + kbase + i * (NT / R);
			const bool kok = FULL || (kg < K);
			const int kc = FULL ? kg : (kok ? kg : 0);
			const double *src = src_r + (AFF ? (int64_t)kc * ks : (int64_t)ktab[kc]);
			const unsigned sh8 = ((offpar + ((kbase + i * (NT / R)) & 1) * strpar) & 1) * 8;
			if constexpr (FULL)
				cp_async8_full(dst + i * ((NT / R) * (R + PAD) * 8) + sh8, src);
			else
				cp_async8(dst + i * ((NT / R) * (R + PAD) * 8) + sh8, src, kok);
		}
	}
}

// Affine operands (element (r,k) at base + r*rs + k*ks, one of the strides 1 — every block on the tensordot / DMRG hot
// path): the producer's fast path. The first version recomputed every element address from (r, k) in every K chunk
// (IMAD.WIDE + LEA pairs, row clamps, k clamps: ~300 instructions per thread and chunk on configs[1], where most tiles
// touch a block edge) and the producer warpgroup, not the tensor pipe, set the chunk rate (ncu source view,
// profiles/r2: producers 68 % busy issuing, consumers 22 % waiting on the full barrier; shrinking the DMMA work by a
// quarter did not move the kernel time). Here a thread's PER elements of a chunk are an arithmetic progression
// ptr + i*step (i-th row of its k column, or i-th k of its row), set up once per (tile, pair); a chunk costs one
// 64-bit add + one LDGSTS per element, elements past the block edge or the K tail are zero-filled (src-size 0: the
// address is never dereferenced) and are exactly the progression's tail i >= n_ok.
struct AffineRun
{
	const double *ptr; // element 0 of the current chunk
	long long step;    // elements between consecutive i
	long long kadv;    // elements per K chunk
	int nfix;          // kc: valid rows of this thread (0..PER); else: PER if the thread's row is inside the block, 0 if not
	int koff;          // kc: the thread's k inside a chunk; else: its first k
	unsigned dst;      // byte offset of element 0 inside the operand's stage slot
};

template <class Cfg, int R>
__device__ __forceinline__ AffineRun affine_setup(int kc, const double *__restrict__ base, int rs, int ks, int r0,
                                                  int Rmax, int pt)
{
	constexpr int BK = Cfg::BK, PAD = Cfg::kPad, NT = Cfg::kProdThreads, PER = R * BK / NT;
	AffineRun a;
	a.kadv = (long long)BK * ks;
	if (kc)
	{
		const int k = pt % BK, rbase = pt / BK;
		a.ptr = base + (long long)(r0 + rbase) * rs + (long long)k * ks;
		a.step = (long long)(NT / BK) * rs;
		const int left = (Rmax - r0 - rbase + (NT / BK) - 1) / (NT / BK);
		a.nfix = left < 0 ? 0 : (left > PER ? PER : left);
		a.koff = k;
		a.dst = (unsigned)(rbase * (BK + PAD) + k) * 8u;
	}
	else
	{
		const int r = pt % R, kbase = pt / R;
		a.ptr = base + (long long)(r0 + r) * rs + (long long)kbase * ks;
		a.step = (long long)(NT / R) * ks;
		a.nfix = (r0 + r < Rmax) ? PER : 0;
		a.koff = kbase;
		a.dst = (unsigned)(kbase * (R + PAD) + r) * 8u;
	}
	return a;
}

template <class Cfg, int R>
__device__ __forceinline__ void affine_issue(int kc, AffineRun &a, unsigned slot, int k0, int K)
{
	constexpr int BK = Cfg::BK, PAD = Cfg::kPad, NT = Cfg::kProdThreads, PER = R * BK / NT;
	// every address in its own register pair BEFORE the first copy issues: an LDGSTS holds its address registers until
	// the load/store unit takes it, and a single running pointer made every copy wait for the previous one to leave the
	// queue (ncu source view: long-scoreboard samples on each pointer increment)
	const double *src[PER];
#pragma unroll
	for (int i = 0; i < PER; ++i)
	{
		src[i] = a.ptr + i * a.step;
		asm volatile("" : "+l"(src[i])); // opaque: keeps ptxas from rematerialising the address into a shared register pair
	}
	a.ptr += a.kadv;
	const unsigned dst = slot + a.dst;
	int n_ok;
	if (kc)
		n_ok = (k0 + a.koff < K) ? a.nfix : 0;
	else
	{
		int left = (K - k0 - a.koff + (NT / R) - 1) / (NT / R);
		left = left < 0 ? 0 : left;
		n_ok = left < a.nfix ? left : a.nfix;
	}
	if (kc)
	{
		constexpr unsigned DS = (NT / BK) * (BK + PAD) * 8;
		if (n_ok == PER)
		{
#pragma unroll
			for (int i = 0; i < PER; ++i)
				cp_async8_full(dst + i * DS, src[i]);
		}
		else
		{
#pragma unroll
			for (int i = 0; i < PER; ++i)
				cp_async8(dst + i * DS, src[i], i < n_ok);
		}
	}
	else
	{
		constexpr unsigned DS = (NT / R) * (R + PAD) * 8;
		if (n_ok == PER)
		{
#pragma unroll
			for (int i = 0; i < PER; ++i)
				cp_async8_full(dst + i * DS, src[i]);
		}
		else
		{
#pragma unroll
			for (int i = 0; i < PER; ++i)
				cp_async8(dst + i * DS, src[i], i < n_ok);
		}
	}
}

// 16-byte variant of the affine fast path for operands whose unit-stride runs are staged in the shifted layout
// (GemmPair::shf, see the bulk path): the 8-byte LDGSTS is what bounds the producer on configs[1] (64 warp-level copies
// of 256 bytes per chunk; the load/store unit takes ~2000 cycles for them, more than the consumers need for the chunk's
// DMMAs), a 16-byte copy moves twice as much per instruction. A run (a row of BK elements when k is the unit-stride
// direction, else the tile's rows at one k) that starts on an odd element is fetched from one element earlier and
// lands one element late in its slot, exactly as in the bulk path, so source and destination are both 16-byte
// aligned; the consumers undo the shift. Pieces past the block edge / the K tail are zero-filled through src-size.
__device__ __forceinline__ void cp_async16(unsigned smem, const void *gmem, int bytes)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async16_full(unsigned smem, const void *gmem)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem), "l"(gmem));
}
struct VecRun
{
	const double *ptr; // piece j of run 0 of this thread, current chunk
	long long step;    // elements between this thread's consecutive runs
	long long kadv;    // elements per K chunk
	int nfix;          // kc: valid runs (rows) of this thread; else unused
	int run0;          // first run of this thread inside a chunk
	int j;             // piece index inside a run
	int sh;            // one-element shift of this thread's runs
	int lv;            // !kc: valid elements of a run (rows of the tile inside the block)
	unsigned dst;
};

template <class Cfg, int R>
__device__ __forceinline__ VecRun vec_setup(int kc, const double *__restrict__ base, int rs, int ks, int r0, int Rmax,
                                            int pt, int offpar, int strpar)
{
	constexpr int BK = Cfg::BK, PAD = Cfg::kPad, NT = Cfg::kProdThreads;
	VecRun a;
	if (kc)
	{
		constexpr int P = BK / 2, RP = NT / P, PASS = R / RP;
		a.j = pt % P;
		a.run0 = pt / P;
		a.sh = (offpar + (a.run0 & 1) * strpar) & 1;
		a.ptr = base + (long long)(r0 + a.run0) * rs + 2 * a.j - a.sh;
		a.step = (long long)RP * rs;
		a.kadv = BK;
		const int left = (Rmax - r0 - a.run0 + RP - 1) / RP;
		a.nfix = left < 0 ? 0 : (left > PASS ? PASS : left);
		a.lv = 0;
		a.dst = (unsigned)(a.run0 * (BK + PAD) + 2 * a.j) * 8u;
	}
	else
	{
		constexpr int P = R / 2, RP = NT / P;
		a.j = pt % P;
		a.run0 = pt / P;
		a.sh = (offpar + (a.run0 & 1) * strpar) & 1;
		a.ptr = base + (long long)a.run0 * ks + r0 + 2 * a.j - a.sh;
		a.step = (long long)RP * ks;
		a.kadv = (long long)BK * ks;
		a.nfix = 0;
		a.lv = Rmax - r0 < R ? Rmax - r0 : R;
		a.dst = (unsigned)(a.run0 * (R + PAD) + 2 * a.j) * 8u;
	}
	return a;
}

template <class Cfg, int R>
__device__ __forceinline__ void vec_issue(int kc, VecRun &a, unsigned slot, int k0, int K)
{
	constexpr int BK = Cfg::BK, PAD = Cfg::kPad, NT = Cfg::kProdThreads;
	constexpr int PASS = R * BK / 2 / NT; // 16-byte pieces per thread and chunk
	const double *src[PASS];
#pragma unroll
	for (int i = 0; i < PASS; ++i)
	{
		src[i] = a.ptr + i * a.step;
		asm volatile("" : "+l"(src[i]));
	}
	a.ptr += a.kadv;
	const unsigned dst = slot + a.dst;
	int n_ok, lv, P;
	unsigned DS;
	if (kc)
	{
		P = BK / 2;
		DS = (NT / (BK / 2)) * (BK + PAD) * 8;
		n_ok = a.nfix;
		lv = K - k0 < BK ? K - k0 : BK;
	}
	else
	{
		P = R / 2;
		DS = (NT / (R / 2)) * (R + PAD) * 8;
		int left = (K - k0 - a.run0 + (NT / (R / 2)) - 1) / (NT / (R / 2));
		left = left < 0 ? 0 : left;
		n_ok = left < PASS ? left : PASS;
		lv = a.lv;
	}
	int e = lv + a.sh - 2 * a.j; // elements of this piece inside the run: >= 2 whole piece, 1 half, <= 0 none
	const int sz = e >= 2 ? 16 : (e == 1 ? 8 : 0);
	if (n_ok == PASS && sz == 16)
	{
#pragma unroll
		for (int i = 0; i < PASS; ++i)
			cp_async16_full(dst + i * DS, src[i]);
	}
	else
	{
#pragma unroll
		for (int i = 0; i < PASS; ++i)
			cp_async16(dst + i * DS, src[i], i < n_ok ? sz : 0);
	}
	if (a.sh && a.j == P - 1)
	{ // the shifted run spills into one more piece (its last element)
		e -= 2;
		const int sz2 = e >= 2 ? 16 : (e == 1 ? 8 : 0);
#pragma unroll
		for (int i = 0; i < PASS; ++i)
			cp_async16(dst + i * DS + 16, src[i] + 2, i < n_ok ? sz2 : 0);
	}
}

template <class Cfg, int R>
__device__ __noinline__ void load_operand(int kcontig, bool affine, unsigned smem_dst, const double *__restrict__ base,
                                             const int32_t *__restrict__ rtab, const int32_t *__restrict__ ktab, int rs,
                                             int ks, int r0, int Rmax, int k0, int K, int pt, int offpar, int strpar)
{
	const bool full = (r0 + R <= Rmax) && (k0 + Cfg::BK <= K);
	if (affine)
	{
		if (kcontig)
		{
			if (full)
				load_tile<Cfg, R, true, true, true>(smem_dst, base, rtab, ktab, rs, ks, r0, Rmax, k0, K, pt, offpar, strpar);
			else
				load_tile<Cfg, R, true, true, false>(smem_dst, base, rtab, ktab, rs, ks, r0, Rmax, k0, K, pt, offpar, strpar);
		}
		else
		{
			if (full)
				load_tile<Cfg, R, true, false, true>(smem_dst, base, rtab, ktab, rs, ks, r0, Rmax, k0, K, pt, offpar, strpar);
			else
				load_tile<Cfg, R, true, false, false>(smem_dst, base, rtab, ktab, rs, ks, r0, Rmax, k0, K, pt, offpar, strpar);
		}
	}
	else
	{ // table-driven gather (blocks whose merged free / contracted dims are not a single stride): always bounds-safe
		if (kcontig)
			load_tile<Cfg, R, false, true, false>(smem_dst, base, rtab, ktab, rs, ks, r0, Rmax, k0, K, pt, offpar, strpar);
		else
			load_tile<Cfg, R, false, false, false>(smem_dst, base, rtab, ktab, rs, ks, r0, Rmax, k0, K, pt, offpar, strpar);
	}
}

// Consumer side: the DMMAs of one staged K chunk for one warp tile, with the operand layouts as template parameters.
// With run-time strides (first version) every fragment load cost an IMAD + IADD and the address registers left room
// for ONE live B fragment: ptxas emitted LDS -> 4 DMMA -> LDS -> 4 DMMA ..., each group waiting out the shared-memory
// latency (ncu source view of configs[1], profiles/r2: the short-scoreboard samples sit on the DMMA after every LDS,
// tensor pipe 44 % active). Here every offset is an immediate and the fragments of k-step kk + 1 are requested before
// the 16 (32) DMMAs of k-step kk issue.
// MI x NI: the 8x8 atoms this warp computes (the whole warp tile, or fewer on block edges of the 64 x 64 configuration).
template <class Cfg, bool AKC, bool BNC, int MI, int NI>
__device__ __forceinline__ void mma_chunk(double (&acc)[Cfg::WM / 8][Cfg::WN / 8][2], const double *__restrict__ As,
                                          const double *__restrict__ Bs, int wm0, int wn0, int g, int q, int shA, int shB)
{
	constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, PAD = Cfg::kPad;
	constexpr int sa_m = AKC ? (BK + PAD) : 1, sa_k = AKC ? 1 : (BM + PAD);
	constexpr int sb_k = BNC ? (BN + PAD) : 1, sb_n = BNC ? 1 : (BK + PAD);
	const double *Ap = As + (wm0 + g) * sa_m + q * sa_k + shA;
	const double *Bp = Bs + q * sb_k + (wn0 + g) * sb_n + shB;
	double af[2][MI], bf[2][NI];
#pragma unroll
	for (int i = 0; i < MI; ++i)
		af[0][i] = Ap[i * 8 * sa_m];
#pragma unroll
	for (int j = 0; j < NI; ++j)
		bf[0][j] = Bp[j * 8 * sb_n];
#pragma unroll
	for (int ks = 0; ks < BK / 4; ++ks)
	{
		const int cur = ks & 1, nxt = cur ^ 1;
		if (ks + 1 < BK / 4)
		{
#pragma unroll
			for (int i = 0; i < MI; ++i)
				af[nxt][i] = Ap[i * 8 * sa_m + (ks + 1) * 4 * sa_k];
#pragma unroll
			for (int j = 0; j < NI; ++j)
				bf[nxt][j] = Bp[(ks + 1) * 4 * sb_k + j * 8 * sb_n];
		}
#pragma unroll
		for (int i = 0; i < MI; ++i)
#pragma unroll
			for (int j = 0; j < NI; ++j)
				dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
	}
}

// All K chunks of one tile for one consumer warp that owns MIV x NJV atoms at (wm0, wn0); MIV == 0: the warp has no
// valid atom in this tile and only keeps the ring moving.
template <class Cfg, int MIV, int NJV>
__device__ __forceinline__ void consume_tile(double (&acc)[Cfg::WM / 8][Cfg::WN / 8][2], const GemmTile &ob,
                                             const GemmPair *__restrict__ pairs, int bulk_mask, const double *smem,
                                             unsigned full0, unsigned empty0, int &stage, unsigned &phase, int wm0,
                                             int wn0, int g, int q, int lane)
{
	constexpr int BK = Cfg::BK, STAGES = Cfg::STAGES;
	int K = ob.K0, lay = ob.flags0, shf = ob.shf0 & bulk_mask;
	for (int p = ob.pair_begin; p < ob.pair_end; ++p)
	{
		int K_next = 0, lay_next = 0, shf_next = 0;
		if (p + 1 < ob.pair_end)
		{
			K_next = pairs[p + 1].K;
			lay_next = pairs[p + 1].a_kcontig | (pairs[p + 1].b_ncontig << 1);
			shf_next = pairs[p + 1].shf & bulk_mask;
		}
		const int a_kc = lay & 1, b_nc = lay >> 1;
		// one-element shifts of the bulk-staged runs (see the producer): the run index is the fragment row g (A rows /
		// B columns when k is the unit-stride direction) or the fragment k index q; every other term of the run's
		// source address (tile origins, chunk origins, warp and MMA offsets) is even
		const int shA = (shf & 1) ? (((shf >> 2) & 1) + ((a_kc ? g : q) & 1) * ((shf >> 3) & 1)) & 1 : 0;
		const int shB = (shf & 2) ? (((shf >> 4) & 1) + ((b_nc ? q : g) & 1) * ((shf >> 5) & 1)) & 1 : 0;
		const int nchunk = (K + BK - 1) / BK;
		for (int ch = 0; ch < nchunk; ++ch)
		{
			mbar_wait(full0 + 8 * stage, phase);
			if (MIV > 0 && (bulk_mask & 0x100)) // bit 8 cleared: diagnostic run without the DMMAs (QTB_GEMM_DEBUG=1)
			{
				const double *As = smem + stage * Cfg::kStage;
				const double *Bs = As + Cfg::kASize;
				// rows / columns past the block edge inside an atom were clamped by the producer (finite data, never
				// stored) and the K tail is zero-filled: no predicate inside the loop
				if constexpr (MIV > 0)
				{
					if (lay == 0)
						mma_chunk<Cfg, false, false, MIV, NJV>(acc, As, Bs, wm0, wn0, g, q, shA, shB);
					else if (lay == 1)
						mma_chunk<Cfg, true, false, MIV, NJV>(acc, As, Bs, wm0, wn0, g, q, shA, shB);
					else if (lay == 2)
						mma_chunk<Cfg, false, true, MIV, NJV>(acc, As, Bs, wm0, wn0, g, q, shA, shB);
					else
						mma_chunk<Cfg, true, true, MIV, NJV>(acc, As, Bs, wm0, wn0, g, q, shA, shB);
				}
			}
			__syncwarp();
			if (lane == 0)
				mbar_arrive(empty0 + 8 * stage);
			if (++stage == STAGES)
			{
				stage = 0;
				phase ^= 1;
			}
		}
		K = K_next;
		lay = lay_next;
		shf = shf_next;
	}
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::MINB)
    grouped_gemm_kernel(const GemmTile *__restrict__ tiles, const int32_t *__restrict__ cta_begin,
                        const GemmPair *__restrict__ pairs, const int32_t *__restrict__ offpool,
                        const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C, int bulk_mask,
                        const double *__restrict__ Cin, double alpha, double beta)
{
	constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES;
	constexpr int PAD = Cfg::kPad;
	extern __shared__ __align__(16) double smem[];
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * Cfg::kStage);
	const unsigned full0 = smem_u32(bars);           // full[s]  : producer -> consumers
	const unsigned empty0 = smem_u32(bars + STAGES); // empty[s] : consumers -> producer
	const int tid = threadIdx.x;
	const int warp = tid >> 5;
	const int lane = tid & 31;

	// static schedule from the planner: this CTA runs tiles [t_begin, t_end), heaviest first (LPT over modelled cycles)
	const int t_begin = cta_begin[blockIdx.x], t_end = cta_begin[blockIdx.x + 1];
	// Descriptor ring in shared memory. First version: every thread kept the current / next / next-but-one work item and
	// two pair records in registers (80 registers of look-ahead); under the 128-register cap ptxas spilled them right
	// after the loads, which turned every prefetch into a blocking load (ncu source view of the kernel skeleton,
	// profiles/r2: 40 % of the producer samples on the STL after a descriptor LDG; the skeleton alone — no copies, no
	// DMMAs — took 27 us of the 56). Now ONE 4-byte word per producer thread is in flight: threads 0-15 hold the next
	// pair record, threads 32-43 the work item two tiles ahead; they are published to shared memory at the next pair /
	// tile boundary (named barrier of the producer warpgroup) and both sides read the records from there.
	static_assert(sizeof(GemmTile) == 48 && sizeof(GemmPair) == 64, "descriptor ring layout");
	int32_t *s_tile_w = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(smem) + Cfg::kDescOff);
	int32_t *s_pair_w = s_tile_w + Cfg::kTileRing * 12;
	const GemmTile *s_tiles = reinterpret_cast<const GemmTile *>(s_tile_w);
	const GemmPair *s_pairs = reinterpret_cast<const GemmPair *>(s_pair_w);
	if (tid == 0)
	{
		for (int s = 0; s < STAGES; ++s)
		{
			mbar_init(full0 + 8 * s, Cfg::kProdThreads);
			mbar_init(empty0 + 8 * s, Cfg::kConsWarps);
		}
	}
	if (tid < 24 && t_begin + tid / 12 < t_end)
		s_tile_w[((t_begin + tid / 12) % Cfg::kTileRing) * 12 + tid % 12] =
		    reinterpret_cast<const int32_t *>(tiles + t_begin + tid / 12)[tid % 12];
	__syncthreads();

	if (warp < 4)
	{
		// ================================================ PRODUCER ================================================
		if constexpr (Cfg::REALLOC)
			asm volatile("setmaxnreg.dec.sync.aligned.u32 104;\n");
		const int pt = tid; // 0..127
		int stage = 0;
		unsigned phase = 0;
		int32_t reg_pair = 0, reg_tile = 0; // the word of the next pair record / of work item t + 2 this thread carries
		if (t_begin < t_end)
		{
			if (pt < 16)
				reg_pair = reinterpret_cast<const int32_t *>(pairs + s_tiles[t_begin % Cfg::kTileRing].pair_begin)[pt];
			if (pt >= 32 && pt < 44 && t_begin + 2 < t_end)
				reg_tile = reinterpret_cast<const int32_t *>(tiles + t_begin + 2)[pt - 32];
		}
		int visit = 0;
		for (int t = t_begin; t < t_end; ++t)
		{
			GemmTile ob;
			int p = 0;
			for (bool first = true;; first = false)
			{
				// ---- pair boundary: publish the record carried in registers, read it back, start the next fetch ----
				if (pt < 16)
					s_pair_w[(visit & 1) * 16 + pt] = reg_pair;
				if (first && t > t_begin && t + 1 < t_end && pt >= 32 && pt < 44)
					s_tile_w[((t + 1) % Cfg::kTileRing) * 12 + pt - 32] = reg_tile;
				asm volatile("bar.sync 1, 128;\n" ::: "memory");
				if (first)
				{
					ob = s_tiles[t % Cfg::kTileRing];
					p = ob.pair_begin;
					if (pt >= 32 && pt < 44 && t + 2 < t_end)
						reg_tile = reinterpret_cast<const int32_t *>(tiles + t + 2)[pt - 32];
				}
				const GemmPair pr = s_pairs[visit & 1];
				++visit;
				{
					const int np = p + 1 < ob.pair_end ? p + 1 : (t + 1 < t_end ? s_tiles[(t + 1) % Cfg::kTileRing].pair_begin : -1);
					if (pt < 16 && np >= 0)
						reg_pair = reinterpret_cast<const int32_t *>(pairs + np)[pt];
				}
				const int M = ob.M, N = ob.N, m0 = ob.m0, n0 = ob.n0;
				const double *Ab = A + pr.a_off;
				const double *Bb = B + pr.b_off;
				const int32_t *aro = offpool + pr.a_roff;
				const int32_t *ako = offpool + pr.a_koff;
				const int32_t *bko = offpool + pr.b_koff;
				const int32_t *bco = offpool + pr.b_coff;
				const bool a_aff = pr.a_rs >= 0, b_aff = pr.b_cs >= 0;
				const int shf = pr.shf & bulk_mask;
				const int nchunk = (pr.K + BK - 1) / BK;
				// affine operands outside the bulk-staged layout: the pointer-progression fast path
				const bool a_fast = a_aff && !(shf & 1), b_fast = b_aff && !(shf & 2);
				// 64 x 64 configuration: shifted-layout operands take 16-byte LDGSTS instead of the bulk path
				const bool a_vec = Cfg::kVec16 && (shf & 1), b_vec = Cfg::kVec16 && (shf & 2);
				AffineRun fa, fb;
				VecRun va, vb;
				if (a_fast)
					fa = affine_setup<Cfg, BM>(pr.a_kcontig, Ab, pr.a_rs, pr.a_ks, m0, M, pt);
				if (b_fast)
					fb = affine_setup<Cfg, BN>(!pr.b_ncontig, Bb, pr.b_cs, pr.b_ks, n0, N, pt);
				if (a_vec)
					va = vec_setup<Cfg, BM>(pr.a_kcontig, Ab, pr.a_rs, pr.a_ks, m0, M, pt, (shf >> 2) & 1, (shf >> 3) & 1);
				if (b_vec)
					vb = vec_setup<Cfg, BN>(!pr.b_ncontig, Bb, pr.b_cs, pr.b_ks, n0, N, pt, (shf >> 4) & 1, (shf >> 5) & 1);
				for (int ch = 0; ch < nchunk; ++ch)
				{
					mbar_wait(empty0 + 8 * stage, phase ^ 1);
					const unsigned As = smem_u32(smem + stage * Cfg::kStage);
					const unsigned Bs = As + Cfg::kASize * 8;
					const unsigned fullb = full0 + 8 * stage;
					const int k0 = ch * BK;
					// Bulk staging (cp.async.bulk -> UBLKCP, the TMA unit): a K-full chunk of an affine operand is a set of
					// unit-stride runs (rows of BK elements, or BK runs along the rows); thread `pt` copies run `pt` with ONE
					// instruction. The source must be 16-byte aligned: a run that starts on an odd element is fetched from one
					// element earlier and lands shifted by one in its slot (every slot has 4 elements of padding); the consumers
					// undo the shift from the parities in GemmPair::shf. Rows / columns past the block edge are not copied
					// (whatever the slot holds only reaches accumulators that are never stored); the K tail chunk takes the
					// LDGSTS path below, which zero-fills and writes with the same shifts.
					const bool kfull = k0 + BK <= pr.K;
					const bool a_bulk = !Cfg::kVec16 && (shf & 1) && kfull, b_bulk = !Cfg::kVec16 && (shf & 2) && kfull;
					unsigned tx = 0, a_bytes = 0, b_bytes = 0, a_dst = 0, b_dst = 0;
					const double *a_src = nullptr, *b_src = nullptr;
					if (a_bulk)
					{
						int len;
						bool valid;
						if (pr.a_kcontig)
						{ // run = row m0 + pt, BK elements
							valid = pt < BM && m0 + pt < M;
							a_src = Ab + (int64_t)(m0 + pt) * pr.a_rs + k0;
							len = BK;
							a_dst = As + pt * (BK + PAD) * 8;
						}
						else
						{ // run = k index k0 + pt, the tile's rows
							valid = pt < BK;
							a_src = Ab + (int64_t)(k0 + pt) * pr.a_ks + m0;
							len = min(BM, M - m0);
							a_dst = As + pt * (BM + PAD) * 8;
						}
						if (valid)
						{
							const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(a_src) >> 3) & 1u;
							a_src -= sh;
							a_bytes = ((sh + len + 1) >> 1) << 4;
						}
					}
					if (b_bulk)
					{
						int len;
						bool valid;
						if (pr.b_ncontig)
						{ // run = k index k0 + pt, the tile's columns
							valid = pt < BK;
							b_src = Bb + (int64_t)(k0 + pt) * pr.b_ks + n0;
							len = min(BN, N - n0);
							b_dst = Bs + pt * (BN + PAD) * 8;
						}
						else
						{ // run = column n0 + pt, BK elements
							valid = pt < BN && n0 + pt < N;
							b_src = Bb + (int64_t)(n0 + pt) * pr.b_cs + k0;
							len = BK;
							b_dst = Bs + pt * (BK + PAD) * 8;
						}
						if (valid)
						{
							const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(b_src) >> 3) & 1u;
							b_src -= sh;
							b_bytes = ((sh + len + 1) >> 1) << 4;
						}
					}
					tx = a_bytes + b_bytes;
					if (tx)
					{
						mbar_expect_tx(fullb, tx);
						if (a_bytes)
							bulk_g2s(a_dst, a_src, a_bytes, fullb);
						if (b_bytes)
							bulk_g2s(b_dst, b_src, b_bytes, fullb);
					}
					if (!(bulk_mask & 0x200))
						; // bit 9 cleared: diagnostic run without the operand copies (QTB_GEMM_DEBUG=2)
					else if (a_fast)
						affine_issue<Cfg, BM>(pr.a_kcontig, fa, As, k0, pr.K);
					else if (a_vec)
						vec_issue<Cfg, BM>(pr.a_kcontig, va, As, k0, pr.K);
					else if (!a_bulk)
						load_operand<Cfg, BM>(pr.a_kcontig, a_aff, As, Ab, aro, ako, pr.a_rs, pr.a_ks, m0, M, k0, pr.K, pt,
						                      (shf & 1) ? (shf >> 2) & 1 : 0, (shf & 1) ? (shf >> 3) & 1 : 0);
					if (!(bulk_mask & 0x200))
						;
					else if (b_fast)
						affine_issue<Cfg, BN>(!pr.b_ncontig, fb, Bs, k0, pr.K);
					else if (b_vec)
						vec_issue<Cfg, BN>(!pr.b_ncontig, vb, Bs, k0, pr.K);
					else if (!b_bulk)
						load_operand<Cfg, BN>(!pr.b_ncontig, b_aff, Bs, Bb, bco, bko, pr.b_cs, pr.b_ks, n0, N, k0, pr.K, pt,
						                      (shf & 2) ? (shf >> 4) & 1 : 0, (shf & 2) ? (shf >> 5) & 1 : 0);
					mbar_arrive_cp_async(full0 + 8 * stage);
					if (++stage == STAGES)
					{
						stage = 0;
						phase ^= 1;
					}
				}
				if (++p >= ob.pair_end)
					break;
			}
		}
		// drain: the async arrivals must have fired before the CTA (and its shared memory) goes away
		asm volatile("cp.async.wait_all;\n" ::: "memory");
	}
	else
	{
		// ================================================ CONSUMERS ===============================================
		if constexpr (Cfg::REALLOC)
			asm volatile("setmaxnreg.inc.sync.aligned.u32 200;\n");
		const int cw = warp - 4;
		const int g = lane >> 2; // fragment row (A) / column (B) inside an 8x8x4 MMA
		const int q = lane & 3;  // fragment k index
		constexpr int MI = WM / 8;
		constexpr int NI = WN / 8;
		int stage = 0;
		unsigned phase = 0;
		// The consumers never wait on a descriptor at a tile boundary: the next work item is fetched while the current
		// one computes and carries the first pair's K / layout flags; further pairs are fetched one pair ahead. (ncu,
		// configs[1], first version: 4.7 long-scoreboard stall cycles per issued instruction from the dependent
		// tiles[] -> outs[] -> pairs[] fetches.)
		for (int t = t_begin; t < t_end; ++t)
		{
			// the producer published this record before it issued the chunks of tile t - 1 (the full-barrier wait orders
			// the read); the ring is deeper than the producer can run ahead (STAGES chunks)
			const GemmTile ob = s_tiles[t % Cfg::kTileRing];
			const int M = ob.M, N = ob.N, m0 = ob.m0, n0 = ob.n0;
			// atoms (8 rows x 8 columns) of this warp that intersect the block. 128 x 128: the static 2 x 4 warp grid, a warp
			// that intersects the block computes its whole tile. 64 x 64: the four warps share the valid atoms of the tile
			// (gemm_warp_grid) and run a loop compiled for exactly their atom count: on configs[1] (blocks of 1..136 rows)
			// the static grid left 22 % of the consumer samples on warps with no valid atom and 1.49x padded DMMA work.
			int wm0 = (cw / Cfg::kWarpsN) * WM, wn0 = (cw % Cfg::kWarpsN) * WN;
			int mi_valid, nj_valid;
			if constexpr (Cfg::kEdge)
			{
				const int mv = min(BM / 8, (M - m0 + 7) >> 3), nv = min(BN / 8, (N - n0 + 7) >> 3);
				int gm, gn, am, an;
				gemm_warp_grid(mv, nv, gm, gn, am, an);
				const int wi = cw / gn, wj = cw - wi * gn;
				wm0 = wi * am * 8;
				wn0 = wj * an * 8;
				mi_valid = max(0, min(am, mv - wi * am));
				nj_valid = max(0, min(an, nv - wj * an));
			}
			else
			{
				mi_valid = (M - m0 - wm0 + 7) / 8;
				mi_valid = mi_valid < 0 ? 0 : (mi_valid > MI ? MI : mi_valid);
				nj_valid = (N - n0 - wn0 + 7) / 8;
				nj_valid = nj_valid < 0 ? 0 : (nj_valid > NI ? NI : nj_valid);
			}
			const bool any = (mi_valid > 0) && (nj_valid > 0);

			double acc[MI][NI][2];
#pragma unroll
			for (int i = 0; i < MI; ++i)
#pragma unroll
				for (int j = 0; j < NI; ++j)
					acc[i][j][0] = acc[i][j][1] = 0.0;

#define QTB_CONSUME(MIV, NJV) \
	consume_tile<Cfg, MIV, NJV>(acc, ob, pairs, bulk_mask, smem, full0, empty0, stage, phase, wm0, wn0, g, q, lane)
			if (!any)
				QTB_CONSUME(0, 0);
			else if constexpr (Cfg::kEdge)
			{
				static_assert(!Cfg::kEdge || (MI == 4 && NI == 4), "edge variants are written for 4 x 4 atoms per warp");
				switch (mi_valid * 4 + nj_valid - 5)
				{
				case 0: QTB_CONSUME(1, 1); break;
				case 1: QTB_CONSUME(1, 2); break;
				case 2: QTB_CONSUME(1, 3); break;
				case 3: QTB_CONSUME(1, 4); break;
				case 4: QTB_CONSUME(2, 1); break;
				case 5: QTB_CONSUME(2, 2); break;
				case 6: QTB_CONSUME(2, 3); break;
				case 7: QTB_CONSUME(2, 4); break;
				case 8: QTB_CONSUME(3, 1); break;
				case 9: QTB_CONSUME(3, 2); break;
				case 10: QTB_CONSUME(3, 3); break;
				case 11: QTB_CONSUME(3, 4); break;
				case 12: QTB_CONSUME(4, 1); break;
				case 13: QTB_CONSUME(4, 2); break;
				case 14: QTB_CONSUME(4, 3); break;
				default: QTB_CONSUME(4, 4); break;
				}
			}
			else
				QTB_CONSUME(MI, NI);
#undef QTB_CONSUME

			// epilogue: the output block is a fresh packed row-major [M,N] matrix (bit 10 cleared: diagnostic run without it)
			if (any && (bulk_mask & 0x400))
			{
				double *Cb = C + ob.c_off;
#pragma unroll
				for (int i = 0; i < MI; ++i)
				{
					const int m = m0 + wm0 + i * 8 + g;
					if (m < M && i < mi_valid)
					{
#pragma unroll
						for (int j = 0; j < NI; ++j)
						{
							if (j >= nj_valid)
								continue; // atoms beyond this warp's share belong to another warp (64 x 64 configuration)
							const int n = n0 + wn0 + j * 8 + 2 * q;
							double *dst = Cb + (size_t)m * N + n;
							if (Cin != nullptr)
							{ // tensorgdot epilogue: D = alpha C + beta A.B (C has the output's packed layout)
								const double *src = Cin + ob.c_off + (size_t)m * N + n;
								if (n < N)
									acc[i][j][0] = alpha * src[0] + beta * acc[i][j][0];
								if (n + 1 < N)
									acc[i][j][1] = alpha * src[1] + beta * acc[i][j][1];
							}
							if (n + 1 < N)
							{
								if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)
									*reinterpret_cast<double2 *>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
								else
								{
									dst[0] = acc[i][j][0];
									dst[1] = acc[i][j][1];
								}
							}
							else if (n < N)
								dst[0] = acc[i][j][0];
						}
					}
				}
			}
		}
	}
}

// ---- skinny contractions ------------------------------------------------------------------------------------------------
// Step 2 of H_eff.psi and of the environment updates contracts the big intermediate with the MPO over (MPO bond,
// physical legs): every matched pair is a [M x K].[K x N] product with M = D_a * D_b (10^4..10^6) and K, N <= the MPO
// bond section size (1..3 for the models of the reference). That is a scaled sum of blocks, HBM bound: tiling it
// 64x64 for the tensor cores wastes 60/64 of every tile and floods the planner with millions of tiles. Here one thread
// owns one output row: it streams A(m, 0..K) of every pair (coalesced over m when the row stride is 1, which is the
// case on the DMRG path), reads the tiny B through the read-only cache (the same address for the whole warp) and keeps
// the N accumulators in registers. Work item = (output block, chunk of kSkinnyRows rows); a thread owns kSkinnyR rows
// 256 apart so that it has kSkinnyR independent loads in flight per pair; the pair descriptors of the block are staged
// in shared memory once per work item.
constexpr int kSkinnyR = kSkinnyRows / 256; // rows per thread: independent loads in flight
constexpr int kSkinnyN = 16;                // max N (and max K per pair) the planner routes here
constexpr int kSkinnyBatch = 32;            // pair descriptors staged in shared memory at a time

template <int NMAX>
__global__ void __launch_bounds__(256, (NMAX <= 4 ? 4 : 1)) skinny_gemm_kernel(const GemmTile *__restrict__ tiles, int ntiles,
                                                           const int32_t *__restrict__ item_prefix,
                                                           const GemmPair *__restrict__ pairs,
                                                           const int32_t *__restrict__ offpool,
                                                           const double *__restrict__ A, const double *__restrict__ B,
                                                           double *__restrict__ C)
{
	constexpr int R = kSkinnyR;
	__shared__ GemmPair sp[kSkinnyBatch];
	static_assert(sizeof(GemmPair) % 4 == 0, "descriptor copied as words");
	const int nitems = item_prefix[ntiles];
	for (int w = blockIdx.x; w < nitems; w += gridDim.x)
	{
		// work item w = chunk (w - prefix[t]) of block record t: binary search of the prefix sums (block-uniform)
		int lo = 0, hi = ntiles;
		while (hi - lo > 1)
		{
			const int mid = (lo + hi) >> 1;
			if (item_prefix[mid] <= w)
				lo = mid;
			else
				hi = mid;
		}
		GemmTile ob = tiles[lo];
		ob.m0 = (w - item_prefix[lo]) * kSkinnyRows;
		const GemmTile &tile = ob;
		const int M = ob.M, N = ob.N;
		int m[R];
		bool ok[R];
		double acc[R][NMAX];
#pragma unroll
		for (int i = 0; i < R; ++i)
		{
			m[i] = tile.m0 + i * 256 + threadIdx.x; // consecutive threads, consecutive rows: coalesced when the row stride is 1
			ok[i] = m[i] < M;
#pragma unroll
			for (int n = 0; n < NMAX; ++n)
				acc[i][n] = 0.0;
		}
		for (int pb = ob.pair_begin; pb < ob.pair_end; pb += kSkinnyBatch)
		{
			const int np = min(kSkinnyBatch, ob.pair_end - pb);
			__syncthreads();
			for (int e = threadIdx.x; e < np * (int)(sizeof(GemmPair) / 4); e += 256)
				reinterpret_cast<int32_t *>(sp)[e] = reinterpret_cast<const int32_t *>(pairs + pb)[e];
			__syncthreads();
			for (int p = 0; p < np; ++p)
			{
				const GemmPair &pr = sp[p];
				const bool a_aff = pr.a_rs >= 0, b_aff = pr.b_cs >= 0;
				const double *Ab = A + pr.a_off;
				const double *Bb = B + pr.b_off;
				int64_t ro[R];
#pragma unroll
				for (int i = 0; i < R; ++i)
					ro[i] = ok[i] ? (a_aff ? (int64_t)m[i] * pr.a_rs : (int64_t)offpool[pr.a_roff + m[i]]) : 0;
				for (int k = 0; k < pr.K; ++k)
				{
					const int64_t ko = a_aff ? (int64_t)k * pr.a_ks : (int64_t)offpool[pr.a_koff + k];
					double a[R];
#pragma unroll
					for (int i = 0; i < R; ++i)
						a[i] = ok[i] ? Ab[ro[i] + ko] : 0.0;
					const double *Bk = Bb + (b_aff ? (int64_t)k * pr.b_ks : (int64_t)offpool[pr.b_koff + k]);
#pragma unroll
					for (int n = 0; n < NMAX; ++n)
						if (n < N)
						{
							const double bv = __ldg(Bk + (b_aff ? (int64_t)n * pr.b_cs : (int64_t)offpool[pr.b_coff + n]));
#pragma unroll
							for (int i = 0; i < R; ++i)
								acc[i][n] += a[i] * bv;
						}
				}
			}
		}
#pragma unroll
		for (int i = 0; i < R; ++i)
			if (ok[i])
			{
				double *dst = C + ob.c_off + (size_t)m[i] * N;
#pragma unroll
				for (int n = 0; n < NMAX; ++n)
					if (n < N)
						dst[n] = acc[i][n];
			}
	}
}

//                     BM   BN  BK  WM  WN  ST MINB realloc
using Cfg64 = GemmCfg<64, 64, 16, 32, 32, 5, 2, false>;   // 4 consumer warps + producer warpgroup = 256 threads
using Cfg128 = GemmCfg<128, 128, 16, 64, 32, 4, 1, true>; // 8 consumer warps + producer warpgroup = 384 threads

template <class Cfg>
static void launch_cfg(Ctx &ctx, int which, const Plan &plan, const double *a, const double *b, double *c,
                       const GemmTile *d_tiles, const int32_t *d_cta_begin, int ncta, const double *cin, double alpha,
                       double beta)
{
	auto kern = grouped_gemm_kernel<Cfg>;
	if (ctx.attr_once(which)) // per context (= per device): the opt-in to > 48 KB of dynamic shared memory
		QTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
	// the bulk-staged layout's shift parities are planned on element offsets: they hold when the arena bases are 16-byte
	// aligned (always for the engine's own arenas; adopted blocks may not be)
	static const int dbg = std::getenv("QTB_GEMM_DEBUG") ? std::atoi(std::getenv("QTB_GEMM_DEBUG")) : 0; // diagnostics: wrong results
	const int bulk_mask = ~0 ^ ((reinterpret_cast<uintptr_t>(a) & 15) ? 1 : 0) ^ ((reinterpret_cast<uintptr_t>(b) & 15) ? 2 : 0) ^
	                      ((dbg & 7) << 8);
	kern<<<ncta, Cfg::kThreads, Cfg::kSmemBytes, ctx.stream>>>(d_tiles, d_cta_begin, plan.d_pairs,
	                                                            plan.d_offpool, a, b, c, bulk_mask, cin, alpha, beta);
	QTB_CUDA(cudaGetLastError());
}

void launch_grouped_gemm(Ctx &ctx, const Plan &plan, const double *a, const double *b, double *c,
                         const Plan::Owned *owned, const double *cin, double alpha, double beta)
{
	const GemmTile *d_tiles = owned ? owned->d_tiles : plan.d_tiles;
	const int32_t *d_cta_begin = owned ? owned->d_cta_begin : plan.d_cta_begin;
	const int ntiles = owned ? owned->ntiles : (int)plan.tiles.size();
	const int ncta = owned ? owned->ncta : plan.ncta;
	if (ntiles == 0)
		return;
	QTB_REQUIRE(cin == nullptr || plan.tile_cfg != 2, QTB_ERR_INVALID_ARGUMENT,
	            "the fused linear-combination epilogue is not available on the streaming (MPO) kernel");
	if (plan.tile_cfg == 2)
	{
		const int grid = ctx.sm_count * 8; // grid-stride over the work items (their count lives in the prefix array)
		if (plan.max_n <= 4)
			skinny_gemm_kernel<4><<<grid, 256, 0, ctx.stream>>>(d_tiles, ntiles, d_cta_begin, plan.d_pairs, plan.d_offpool, a, b, c);
		else
			skinny_gemm_kernel<kSkinnyN><<<grid, 256, 0, ctx.stream>>>(d_tiles, ntiles, d_cta_begin, plan.d_pairs,
			                                                            plan.d_offpool, a, b, c);
		QTB_CUDA(cudaGetLastError());
	}
	else if (plan.tile_cfg == 0)
		launch_cfg<Cfg64>(ctx, 0, plan, a, b, c, d_tiles, d_cta_begin, ncta, cin, alpha, beta);
	else
		launch_cfg<Cfg128>(ctx, 1, plan, a, b, c, d_tiles, d_cta_begin, ncta, cin, alpha, beta);
	ctx.counters[0] += 1;
	ctx.counters[1] += 1;
	ctx.counters[6] += owned ? owned->flops : plan.flops;
}

} // namespace qtb
