// Block-pair matching of a contraction on the device (north star: "block-pair selection-rule matching becomes an
// on-GPU sort / segmented-match kernel; its output block structure must be bit-exact with the reference").
// Reference: the two-pointer merge over "columns" of btensor::tensordot, sources/btensor.cpp:2057-2108, i.e. every
// (A block, B block) pair with equal contracted block indices, grouped by the output index (free A, free B) in
// ascending order, pairs inside a group in ascending contracted index. The host planner (qtb_core.cpp, build_plan)
// states that as two sorts of mixed-radix keys; this file is the same computation as kernels:
//   1. sort the B blocks by contracted key (one CTA, bitonic network in shared memory, ties by block number)
//   2. one thread per A block: equal range of its contracted key in the sorted B list -> candidate count
//   3. exclusive scan of the counts (one CTA)
//   4. one thread per A block writes its candidate records (output key, contracted key, a, b)
//   5. bitonic sort of the records by (output key, contracted key) in global memory (keys are unique)
//   6. segment heads: record p starts a new output block when its output key differs from record p - 1
// Two host round trips: the candidate total (to size the record buffer) and the sorted records themselves, which the
// planner needs on the host anyway to size the output arena and to build the tile schedule.
// Selected by qtb_ctx_set_device_planner (always / never / by size). The host matching of ~1000 blocks takes tens of
// microseconds, less than one kernel launch + synchronise, so "by size" switches over only for thousands of blocks
// (U(1)xU(1) tensors at bond dimension 8192 and beyond); the plan is identical either way (tests/test_gpu_planner.py).
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "qtb_core.h"
#include "qtb_vec.h"

namespace qtb
{
namespace
{
constexpr int kMatchSortMax = 4096; // B blocks the one-CTA sort holds in 48 KB of shared memory

__global__ void __launch_bounds__(1024) match_sort_b_kernel(const unsigned long long *__restrict__ b_ck, int nb,
                                                            unsigned long long *__restrict__ s_key, int *__restrict__ s_idx)
{
	extern __shared__ unsigned long long mk[];
	int n2 = 1;
	while (n2 < nb)
		n2 <<= 1;
	int *mi = reinterpret_cast<int *>(mk + n2);
	for (int i = threadIdx.x; i < n2; i += blockDim.x)
	{
		mk[i] = i < nb ? b_ck[i] : ~0ull;
		mi[i] = i < nb ? i : 0x7fffffff;
	}
	__syncthreads();
	for (int k = 2; k <= n2; k <<= 1)
		for (int j = k >> 1; j > 0; j >>= 1)
		{
			for (int i = threadIdx.x; i < n2; i += blockDim.x)
			{
				const int x = i ^ j;
				if (x > i)
				{
					const unsigned long long ka = mk[i], kb = mk[x];
					const int ia = mi[i], ib = mi[x];
					const bool a_first = ka < kb || (ka == kb && ia < ib);
					if (a_first != ((i & k) == 0))
					{
						mk[i] = kb;
						mk[x] = ka;
						mi[i] = ib;
						mi[x] = ia;
					}
				}
			}
			__syncthreads();
		}
	for (int i = threadIdx.x; i < nb; i += blockDim.x)
	{
		s_key[i] = mk[i];
		s_idx[i] = mi[i];
	}
}

__global__ void match_count_kernel(const unsigned long long *__restrict__ a_ck, int na,
                                   const unsigned long long *__restrict__ s_key, int nb, int *__restrict__ lo_out,
                                   int *__restrict__ cnt)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= na)
		return;
	const unsigned long long key = a_ck[i];
	int lo = 0, hi = nb;
	while (lo < hi)
	{
		const int mid = (lo + hi) >> 1;
		if (s_key[mid] < key)
			lo = mid + 1;
		else
			hi = mid;
	}
	int e = lo, hi2 = nb;
	while (e < hi2)
	{
		const int mid = (e + hi2) >> 1;
		if (s_key[mid] <= key)
			e = mid + 1;
		else
			hi2 = mid;
	}
	lo_out[i] = lo;
	cnt[i] = e - lo;
}

// exclusive scan of cnt[0..n) into off[0..n], off[n] = total; one CTA, chunks of blockDim.x with a running carry
__global__ void __launch_bounds__(1024) match_scan_kernel(const int *__restrict__ cnt, int n, int *__restrict__ off)
{
	__shared__ int sh[1024];
	__shared__ int carry;
	if (threadIdx.x == 0)
		carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += blockDim.x)
	{
		const int i = base + threadIdx.x;
		const int v = i < n ? cnt[i] : 0;
		sh[threadIdx.x] = v;
		__syncthreads();
		for (int o = 1; o < (int)blockDim.x; o <<= 1)
		{
			const int t = threadIdx.x >= (unsigned)o ? sh[threadIdx.x - o] : 0;
			__syncthreads();
			sh[threadIdx.x] += t;
			__syncthreads();
		}
		if (i < n)
			off[i] = carry + sh[threadIdx.x] - v;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1)
			carry += sh[threadIdx.x];
		__syncthreads();
	}
	if (threadIdx.x == 0)
		off[n] = carry;
}

__global__ void match_emit_kernel(const unsigned long long *__restrict__ a_ck, const unsigned long long *__restrict__ a_fk,
                                  int na, const unsigned long long *__restrict__ b_fk, const int *__restrict__ s_idx,
                                  const int *__restrict__ lo, const int *__restrict__ cnt, const int *__restrict__ off,
                                  unsigned long long rb, MatchRec *__restrict__ recs, int total, int n2)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < na)
	{
		const int c = cnt[i], l = lo[i], o = off[i];
		for (int t = 0; t < c; ++t)
		{
			const int j = s_idx[l + t];
			MatchRec r;
			r.okey = a_fk[i] * rb + b_fk[j];
			r.ckey = a_ck[i];
			r.a = i;
			r.b = j;
			r.head = 0;
			r.pad = 0;
			recs[o + t] = r;
		}
	}
	// padding records of the power-of-two sort buffer sort last
	for (int p = total + blockIdx.x * blockDim.x + threadIdx.x; p < n2; p += gridDim.x * blockDim.x)
	{
		MatchRec r;
		r.okey = ~0ull;
		r.ckey = ~0ull;
		r.a = r.b = -1;
		r.head = 0;
		r.pad = 0;
		recs[p] = r;
	}
}

__global__ void match_bitonic_kernel(MatchRec *__restrict__ recs, int n2, int j, int k)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n2)
		return;
	const int x = i ^ j;
	if (x <= i)
		return;
	const MatchRec ra = recs[i], rx = recs[x];
	const bool a_first = ra.okey < rx.okey || (ra.okey == rx.okey && ra.ckey < rx.ckey);
	const bool equal = ra.okey == rx.okey && ra.ckey == rx.ckey; // only padding records compare equal
	if (!equal && a_first != ((i & k) == 0))
	{
		recs[i] = rx;
		recs[x] = ra;
	}
}

__global__ void match_heads_kernel(MatchRec *__restrict__ recs, int total)
{
	const int p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p < total)
		recs[p].head = (p == 0 || recs[p].okey != recs[p - 1].okey) ? 1 : 0;
}
} // namespace

bool device_match(Ctx &ctx, const std::vector<unsigned long long> &a_ck, const std::vector<unsigned long long> &a_fk,
                  const std::vector<unsigned long long> &b_ck, const std::vector<unsigned long long> &b_fk,
                  unsigned long long rb, std::vector<MatchRec> &out)
{
	const int na = (int)a_ck.size(), nb = (int)b_ck.size();
	out.clear();
	if (na == 0 || nb == 0)
		return true;
	if (nb > kMatchSortMax)
		return false; // the caller falls back to the host matching
	auto up = [&](const std::vector<unsigned long long> &v)
	{ return (unsigned long long *)ctx_upload(ctx, v.data(), v.size() * sizeof(unsigned long long)); };
	unsigned long long *d_ack = up(a_ck), *d_afk = up(a_fk), *d_bck = up(b_ck), *d_bfk = up(b_fk);
	unsigned long long *d_skey = (unsigned long long *)ctx_alloc(ctx, (size_t)nb * 8);
	int *d_sidx = (int *)ctx_alloc(ctx, (size_t)nb * 4);
	int *d_lo = (int *)ctx_alloc(ctx, (size_t)na * 4), *d_cnt = (int *)ctx_alloc(ctx, (size_t)na * 4);
	int *d_off = (int *)ctx_alloc(ctx, (size_t)(na + 1) * 4);
	int nb2 = 1;
	while (nb2 < nb)
		nb2 <<= 1;
	match_sort_b_kernel<<<1, 1024, (size_t)nb2 * 12, ctx.stream>>>(d_bck, nb, d_skey, d_sidx);
	match_count_kernel<<<(na + 255) / 256, 256, 0, ctx.stream>>>(d_ack, na, d_skey, nb, d_lo, d_cnt);
	match_scan_kernel<<<1, 1024, 0, ctx.stream>>>(d_cnt, na, d_off);
	QTB_CUDA(cudaGetLastError());
	int total = 0;
	QTB_CUDA(cudaMemcpyAsync(&total, d_off + na, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
	QTB_CUDA(cudaStreamSynchronize(ctx.stream));
	ctx.counters[0] += 3;
	ctx.counters[5] += sizeof(int);
	if (total > 0)
	{
		int n2 = 1;
		while (n2 < total)
			n2 <<= 1;
		MatchRec *d_recs = (MatchRec *)ctx_alloc(ctx, (size_t)n2 * sizeof(MatchRec));
		match_emit_kernel<<<(std::max(na, 256) + 255) / 256, 256, 0, ctx.stream>>>(d_ack, d_afk, na, d_bfk, d_sidx, d_lo, d_cnt, d_off,
		                                                                          rb, d_recs, total, n2);
		int launches = 1;
		for (int k = 2; k <= n2; k <<= 1)
			for (int j = k >> 1; j > 0; j >>= 1, ++launches)
				match_bitonic_kernel<<<(n2 + 255) / 256, 256, 0, ctx.stream>>>(d_recs, n2, j, k);
		match_heads_kernel<<<(total + 255) / 256, 256, 0, ctx.stream>>>(d_recs, total);
		QTB_CUDA(cudaGetLastError());
		out.resize(total);
		QTB_CUDA(cudaMemcpyAsync(out.data(), d_recs, (size_t)total * sizeof(MatchRec), cudaMemcpyDeviceToHost, ctx.stream));
		QTB_CUDA(cudaStreamSynchronize(ctx.stream));
		ctx.counters[0] += launches + 1;
		ctx.counters[5] += (i64)total * (i64)sizeof(MatchRec);
		ctx_free(ctx, d_recs);
	}
	for (void *p : {(void *)d_ack, (void *)d_afk, (void *)d_bck, (void *)d_bfk, (void *)d_skey, (void *)d_sidx, (void *)d_lo,
	                (void *)d_cnt, (void *)d_off})
		ctx_free(ctx, p);
	ctx.device_matches += 1;
	return true;
}

} // namespace qtb
