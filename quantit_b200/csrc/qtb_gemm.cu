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
//    matrix lives at base + roff[m] + koff[k]; the two int32 offset tables are built by the planner from the block's
//    dims/strides, so the permute+reshape the reference materialises is fused into the cp.async operand load.
//  * fp64 tensor cores: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 — the only fp64 MMA shape sm_100a has; tcgen05 has no
//    fp64 kind). Operand tiles are staged through a STAGES-deep cp.async pipeline in shared memory, laid out per
//    operand orientation so that both the async stores and the fragment loads are bank-conflict free.
//  * persistent CTAs pull tiles (sorted by decreasing cost by the planner) from an atomic counter.
#include <cuda_runtime.h>

#include <cstdint>

#include "qtb_core.h"

namespace qtb
{

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool valid)
{
	unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
	int sz = valid ? 8 : 0; // src-size 0 -> the 8 destination bytes are zero-filled
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
	asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES>
struct GemmCfg
{
	static constexpr int kWarpsM = BM / WM;
	static constexpr int kWarpsN = BN / WN;
	static constexpr int kThreads = kWarpsM * kWarpsN * 32;
	static constexpr int kPad = 4;
	// an operand tile is stored either "k-major" [rows][BK+4] or "row-major" [BK][rows+4]; reserve the larger
	static constexpr int kASize = (BM * (BK + kPad) > BK * (BM + kPad)) ? BM * (BK + kPad) : BK * (BM + kPad);
	static constexpr int kBSize = (BN * (BK + kPad) > BK * (BN + kPad)) ? BN * (BK + kPad) : BK * (BN + kPad);
	static constexpr int kStage = kASize + kBSize;
	static constexpr size_t kSmemBytes = size_t(STAGES) * kStage * sizeof(double);
	static constexpr int kAPerThread = BM * BK / kThreads;
	static constexpr int kBPerThread = BN * BK / kThreads;
	static_assert(BM * BK % kThreads == 0 && BN * BK % kThreads == 0, "tile/threads mismatch");
	static_assert(BM % 16 == 0 && BN % 16 == 0 && BK % 4 == 0, "tile shape");
};

// iterator over the (pair, k-chunk) steps of one output block
struct StepIter
{
	int pair;  // current pair index (absolute)
	int chunk; // k-chunk inside the pair
	int nchunk;
};

template <class Cfg, int BM, int BN, int BK, int WM, int WN, int STAGES>
__global__ void __launch_bounds__(Cfg::kThreads)
    grouped_gemm_kernel(const GemmTile *__restrict__ tiles, int ntiles, const GemmOut *__restrict__ outs,
                        const GemmPair *__restrict__ pairs, const int32_t *__restrict__ offpool,
                        const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ C,
                        int *__restrict__ counter)
{
	extern __shared__ __align__(16) double smem[];
	__shared__ int s_tile;
	const int tid = threadIdx.x;
	const int lane = tid & 31;
	const int warp = tid >> 5;
	const int wm0 = (warp / Cfg::kWarpsN) * WM;
	const int wn0 = (warp % Cfg::kWarpsN) * WN;
	const int g = lane >> 2; // fragment row (A) / column (B) inside an 8x8x4 MMA
	const int q = lane & 3;  // fragment k index
	constexpr int MI = WM / 8;
	constexpr int NI = WN / 8;

	for (;;)
	{
		if (tid == 0)
			s_tile = atomicAdd(counter, 1);
		__syncthreads();
		const int t = s_tile;
		__syncthreads();
		if (t >= ntiles)
			break;
		const GemmTile tile = tiles[t];
		const GemmOut ob = outs[tile.out_blk];
		const int M = ob.M, N = ob.N;
		const int m0 = tile.m0, n0 = tile.n0;

		double acc[MI][NI][2];
#pragma unroll
		for (int i = 0; i < MI; ++i)
#pragma unroll
			for (int j = 0; j < NI; ++j)
				acc[i][j][0] = acc[i][j][1] = 0.0;

		// total number of (pair, chunk) steps of this block
		int total_steps = 0;
		for (int p = ob.pair_begin; p < ob.pair_end; ++p)
			total_steps += (pairs[p].K + BK - 1) / BK;

		// producer iterator
		int p_pair = ob.pair_begin, p_chunk = 0;
		int p_nchunk = (pairs[p_pair].K + BK - 1) / BK;
		// consumer iterator
		int c_pair = ob.pair_begin, c_chunk = 0;
		int c_nchunk = p_nchunk;

		auto issue_load = [&](int stage)
		{
			const GemmPair pr = pairs[p_pair];
			double *As = smem + stage * Cfg::kStage;
			double *Bs = As + Cfg::kASize;
			const int k0 = p_chunk * BK;
			const double *Ab = A + pr.a_off;
			const double *Bb = B + pr.b_off;
			const int32_t *aro = offpool + pr.a_roff;
			const int32_t *ako = offpool + pr.a_koff;
			const int32_t *bko = offpool + pr.b_koff;
			const int32_t *bco = offpool + pr.b_coff;
			if (pr.a_kcontig)
			{ // consecutive threads walk k: smem layout [BM][BK+4]
#pragma unroll
				for (int i = 0; i < Cfg::kAPerThread; ++i)
				{
					const int e = tid + i * Cfg::kThreads;
					const int m = e / BK, k = e % BK;
					const bool ok = (m0 + m < M) && (k0 + k < pr.K);
					const double *src = ok ? Ab + aro[m0 + m] + ako[k0 + k] : Ab;
					cp_async8(As + m * (BK + Cfg::kPad) + k, src, ok);
				}
			}
			else
			{ // consecutive threads walk m: smem layout [BK][BM+4]
#pragma unroll
				for (int i = 0; i < Cfg::kAPerThread; ++i)
				{
					const int e = tid + i * Cfg::kThreads;
					const int k = e / BM, m = e % BM;
					const bool ok = (m0 + m < M) && (k0 + k < pr.K);
					const double *src = ok ? Ab + aro[m0 + m] + ako[k0 + k] : Ab;
					cp_async8(As + k * (BM + Cfg::kPad) + m, src, ok);
				}
			}
			if (pr.b_ncontig)
			{ // consecutive threads walk n: smem layout [BK][BN+4]
#pragma unroll
				for (int i = 0; i < Cfg::kBPerThread; ++i)
				{
					const int e = tid + i * Cfg::kThreads;
					const int k = e / BN, n = e % BN;
					const bool ok = (n0 + n < N) && (k0 + k < pr.K);
					const double *src = ok ? Bb + bko[k0 + k] + bco[n0 + n] : Bb;
					cp_async8(Bs + k * (BN + Cfg::kPad) + n, src, ok);
				}
			}
			else
			{ // consecutive threads walk k: smem layout [BN][BK+4]
#pragma unroll
				for (int i = 0; i < Cfg::kBPerThread; ++i)
				{
					const int e = tid + i * Cfg::kThreads;
					const int n = e / BK, k = e % BK;
					const bool ok = (n0 + n < N) && (k0 + k < pr.K);
					const double *src = ok ? Bb + bko[k0 + k] + bco[n0 + n] : Bb;
					cp_async8(Bs + n * (BK + Cfg::kPad) + k, src, ok);
				}
			}
			if (++p_chunk == p_nchunk)
			{
				p_chunk = 0;
				++p_pair;
				if (p_pair < ob.pair_end)
					p_nchunk = (pairs[p_pair].K + BK - 1) / BK;
			}
		};

		// prologue: fill STAGES-1 stages
		int issued = 0;
#pragma unroll
		for (int s = 0; s < STAGES - 1; ++s)
		{
			if (issued < total_steps)
			{
				issue_load(s);
				++issued;
			}
			cp_async_commit();
		}

		for (int step = 0; step < total_steps; ++step)
		{
			cp_async_wait<STAGES - 2>();
			__syncthreads();
			// refill the stage that was consumed in the previous iteration
			if (issued < total_steps)
			{
				issue_load((step + STAGES - 1) % STAGES);
				++issued;
			}
			cp_async_commit();

			const int stage = step % STAGES;
			const double *As = smem + stage * Cfg::kStage;
			const double *Bs = As + Cfg::kASize;
			const GemmPair pr = pairs[c_pair];
			const int sa_m = pr.a_kcontig ? (BK + Cfg::kPad) : 1;
			const int sa_k = pr.a_kcontig ? 1 : (BM + Cfg::kPad);
			const int sb_k = pr.b_ncontig ? (BN + Cfg::kPad) : 1;
			const int sb_n = pr.b_ncontig ? 1 : (BK + Cfg::kPad);
#pragma unroll
			for (int kk = 0; kk < BK; kk += 4)
			{
				double af[MI], bf[NI];
#pragma unroll
				for (int i = 0; i < MI; ++i)
					af[i] = As[(wm0 + i * 8 + g) * sa_m + (kk + q) * sa_k];
#pragma unroll
				for (int j = 0; j < NI; ++j)
					bf[j] = Bs[(kk + q) * sb_k + (wn0 + j * 8 + g) * sb_n];
#pragma unroll
				for (int i = 0; i < MI; ++i)
#pragma unroll
					for (int j = 0; j < NI; ++j)
						dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
			}
			if (++c_chunk == c_nchunk)
			{
				c_chunk = 0;
				++c_pair;
				if (c_pair < ob.pair_end)
					c_nchunk = (pairs[c_pair].K + BK - 1) / BK;
			}
		}
		cp_async_wait<0>();

		// epilogue: the output block is a fresh packed row-major [M,N] matrix
		double *Cb = C + ob.c_off;
#pragma unroll
		for (int i = 0; i < MI; ++i)
		{
			const int m = m0 + wm0 + i * 8 + g;
			if (m < M)
			{
#pragma unroll
				for (int j = 0; j < NI; ++j)
				{
					const int n = n0 + wn0 + j * 8 + 2 * q;
					double *dst = Cb + (size_t)m * N + n;
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
		__syncthreads(); // all warps done with the stages before the next tile's prologue overwrites them
	}
}

using Cfg64 = GemmCfg<64, 64, 16, 32, 32, 3>;
using Cfg128 = GemmCfg<128, 128, 16, 64, 32, 3>;

static int g_blocks_per_sm[2] = {0, 0};

template <class Cfg, int BM, int BN, int BK, int WM, int WN, int STAGES>
static void launch_cfg(Ctx &ctx, int which, const Plan &plan, const double *a, const double *b, double *c)
{
	auto kern = grouped_gemm_kernel<Cfg, BM, BN, BK, WM, WN, STAGES>;
	if (g_blocks_per_sm[which] == 0)
	{
		QTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes));
		int nb = 0;
		QTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, Cfg::kThreads, Cfg::kSmemBytes));
		g_blocks_per_sm[which] = nb > 0 ? nb : 1;
	}
	const int ntiles = (int)plan.tiles.size();
	int grid = ctx.sm_count * g_blocks_per_sm[which];
	if (grid > ntiles)
		grid = ntiles;
	QTB_CUDA(cudaMemsetAsync(plan.d_counter, 0, sizeof(int), ctx.stream));
	kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, ctx.stream>>>(plan.d_tiles, ntiles, plan.d_outs, plan.d_pairs,
	                                                            plan.d_offpool, a, b, c, plan.d_counter);
	QTB_CUDA(cudaGetLastError());
}

void launch_grouped_gemm(Ctx &ctx, const Plan &plan, const double *a, const double *b, double *c)
{
	if (plan.tiles.empty())
		return;
	if (plan.tile_cfg == 0)
		launch_cfg<Cfg64, 64, 64, 16, 32, 32, 3>(ctx, 0, plan, a, b, c);
	else
		launch_cfg<Cfg128, 128, 128, 16, 64, 32, 3>(ctx, 1, plan, a, b, c);
	ctx.counters[0] += 1;
	ctx.counters[1] += 1;
	ctx.counters[6] += plan.flops;
}

} // namespace qtb
