// qtb_vec.cu — HBM-bound arena kernels: strided block gather, axpby over block segments, dot, scale, broadcast mul.
//
// Replace the per-block torch ops of the reference's Lanczos step and DMRG bookkeeping
// (reference sources/dmrg.cpp:585-638,188-199 -> sources/btensor.cpp mul/add_/div_/sum/sqrt, call sites K12-K14 of
// SURVEY.md §2.3): one launch per whole block tensor instead of one torch op per block, coefficients read from
// device memory so the Lanczos recurrence never synchronises with the host.
#include <cuda_runtime.h>

#include <cstdint>

#include "qtb_core.h"
#include "qtb_vec.h"

namespace qtb
{

// ---------------------------------------------------------------------------------------------------------------------
// strided gather: dst (packed, C-contiguous per block) <- src blocks with arbitrary strides
// ---------------------------------------------------------------------------------------------------------------------
__global__ void gather_kernel(const GatherDesc *__restrict__ descs, const GatherWork *__restrict__ work, int nwork,
                              const double *__restrict__ src, double *__restrict__ dst)
{
	for (int w = blockIdx.x; w < nwork; w += gridDim.x)
	{
		const GatherWork wk = work[w];
		const GatherDesc d = descs[wk.desc];
		const i64 end = min(wk.begin + (i64)kGatherChunk, d.numel);
		for (i64 e = wk.begin + threadIdx.x; e < end; e += blockDim.x)
		{
			i64 rem = e, so = 0;
#pragma unroll 1
			for (int k = d.rank - 1; k >= 0; --k)
			{
				const i64 c = rem % d.dims[k];
				rem /= d.dims[k];
				so += c * d.strides[k];
			}
			dst[d.dst_off + e] = src[d.src_off + so];
		}
	}
}

void launch_gather(Ctx &ctx, const std::vector<GatherDesc> &descs, const double *src, double *dst)
{
	if (descs.empty())
		return;
	std::vector<GatherWork> work;
	for (size_t i = 0; i < descs.size(); ++i)
		for (i64 b = 0; b < descs[i].numel; b += kGatherChunk)
			work.push_back({(int)i, 0, b});
	if (work.empty())
		return;
	auto d_desc = ctx_upload(ctx, descs.data(), descs.size() * sizeof(GatherDesc));
	auto d_work = ctx_upload(ctx, work.data(), work.size() * sizeof(GatherWork));
	int grid = (int)std::min<size_t>(work.size(), (size_t)ctx.sm_count * 8);
	gather_kernel<<<grid, 256, 0, ctx.stream>>>((const GatherDesc *)d_desc, (const GatherWork *)d_work,
	                                           (int)work.size(), src, dst);
	QTB_CUDA(cudaGetLastError());
	ctx_free(ctx, d_desc);
	ctx_free(ctx, d_work);
	ctx.counters[0] += 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// segment axpby: out[seg] = ca * a[seg] + cb * b[seg]; a or b may be absent (offset < 0)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void axpby_kernel(const VecSeg *__restrict__ segs, const VecWork *__restrict__ work, int nwork,
                             const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out,
                             const double *__restrict__ ca_ptr, double ca_mul, const double *__restrict__ cb_ptr,
                             double cb_mul, int divide_a)
{
	const double ca = (ca_ptr ? *ca_ptr : 1.0) * ca_mul;
	const double cb = (cb_ptr ? *cb_ptr : 1.0) * cb_mul;
	for (int w = blockIdx.x; w < nwork; w += gridDim.x)
	{
		const VecWork wk = work[w];
		const VecSeg s = segs[wk.seg];
		const i64 end = min(wk.begin + (i64)kVecChunk, s.n);
		for (i64 e = wk.begin + threadIdx.x; e < end; e += blockDim.x)
		{
			double v = 0.0;
			if (s.a_off >= 0)
				v = (divide_a ? a[s.a_off + e] / ca : ca * a[s.a_off + e]) * s.a_scale;
			if (s.b_off >= 0)
				v += cb * b[s.b_off + e];
			out[s.o_off + e] = v;
		}
	}
}

void launch_axpby(Ctx &ctx, const std::vector<VecSeg> &segs, const double *a, const double *b, double *out,
                  const double *ca_ptr, double ca_mul, const double *cb_ptr, double cb_mul, bool divide_a)
{
	std::vector<VecWork> work;
	for (size_t i = 0; i < segs.size(); ++i)
		for (i64 s = 0; s < segs[i].n; s += kVecChunk)
			work.push_back({(int)i, 0, s});
	if (work.empty())
		return;
	auto d_segs = ctx_upload(ctx, segs.data(), segs.size() * sizeof(VecSeg));
	auto d_work = ctx_upload(ctx, work.data(), work.size() * sizeof(VecWork));
	int grid = (int)std::min<size_t>(work.size(), (size_t)ctx.sm_count * 8);
	axpby_kernel<<<grid, 256, 0, ctx.stream>>>((const VecSeg *)d_segs, (const VecWork *)d_work, (int)work.size(), a, b,
	                                          out, ca_ptr, ca_mul, cb_ptr, cb_mul, divide_a ? 1 : 0);
	QTB_CUDA(cudaGetLastError());
	ctx_free(ctx, d_segs);
	ctx_free(ctx, d_work);
	ctx.counters[0] += 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// segment dot: result[0] = sum over segments of a[seg] . b[seg]   (deterministic two-stage reduction)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_reduce_sum(double v, double *sh)
{
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_down_sync(0xffffffffu, v, o);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0)
		sh[warp] = v;
	__syncthreads();
	if (warp == 0)
	{
		v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
		for (int o = 16; o > 0; o >>= 1)
			v += __shfl_down_sync(0xffffffffu, v, o);
	}
	__syncthreads();
	return v; // valid in thread 0
}

__global__ void dot_partial_kernel(const VecSeg *__restrict__ segs, const VecWork *__restrict__ work, int nwork,
                                   const double *__restrict__ a, const double *__restrict__ b,
                                   double *__restrict__ partial)
{
	__shared__ double sh[32];
	double acc = 0.0;
	for (int w = blockIdx.x; w < nwork; w += gridDim.x)
	{
		const VecWork wk = work[w];
		const VecSeg s = segs[wk.seg];
		const i64 end = min(wk.begin + (i64)kVecChunk, s.n);
		for (i64 e = wk.begin + threadIdx.x; e < end; e += blockDim.x)
			acc += a[s.a_off + e] * b[s.b_off + e];
	}
	acc = block_reduce_sum(acc, sh);
	if (threadIdx.x == 0)
		partial[blockIdx.x] = acc;
}
__global__ void dot_final_kernel(const double *__restrict__ partial, int n, double *__restrict__ result, int op)
{
	__shared__ double sh[32];
	double acc = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		acc += partial[i];
	acc = block_reduce_sum(acc, sh);
	if (threadIdx.x == 0)
		result[0] = op == 1 ? sqrt(acc) : acc;
}

void launch_dot(Ctx &ctx, const std::vector<VecSeg> &segs, const double *a, const double *b, double *d_result,
                bool take_sqrt)
{
	std::vector<VecWork> work;
	for (size_t i = 0; i < segs.size(); ++i)
		for (i64 s = 0; s < segs[i].n; s += kVecChunk)
			work.push_back({(int)i, 0, s});
	if (work.empty())
	{
		QTB_CUDA(cudaMemsetAsync(d_result, 0, sizeof(double), ctx.stream));
		return;
	}
	auto d_segs = ctx_upload(ctx, segs.data(), segs.size() * sizeof(VecSeg));
	auto d_work = ctx_upload(ctx, work.data(), work.size() * sizeof(VecWork));
	int grid = (int)std::min<size_t>(work.size(), (size_t)ctx.sm_count * 4);
	double *partial = (double *)ctx_alloc(ctx, grid * sizeof(double));
	dot_partial_kernel<<<grid, 256, 0, ctx.stream>>>((const VecSeg *)d_segs, (const VecWork *)d_work, (int)work.size(),
	                                                a, b, partial);
	dot_final_kernel<<<1, 256, 0, ctx.stream>>>(partial, grid, d_result, take_sqrt ? 1 : 0);
	QTB_CUDA(cudaGetLastError());
	ctx_free(ctx, partial);
	ctx_free(ctx, d_segs);
	ctx_free(ctx, d_work);
	ctx.counters[0] += 2;
}

// ---------------------------------------------------------------------------------------------------------------------
// in-place scale of a contiguous range
// ---------------------------------------------------------------------------------------------------------------------
__global__ void scale_kernel(double *__restrict__ x, i64 n, const double *__restrict__ c_ptr, double c_mul, int divide)
{
	const double c = (c_ptr ? *c_ptr : 1.0) * c_mul;
	for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
		x[i] = divide ? x[i] / c : x[i] * c;
}
void launch_scale(Ctx &ctx, double *x, i64 n, const double *c_ptr, double c_mul, bool divide)
{
	if (n <= 0)
		return;
	int grid = (int)std::min<i64>((n + 255) / 256, (i64)ctx.sm_count * 8);
	scale_kernel<<<grid, 256, 0, ctx.stream>>>(x, n, c_ptr, c_mul, divide ? 1 : 0);
	QTB_CUDA(cudaGetLastError());
	ctx.counters[0] += 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// out(r, k) = a(r, k) * d(k)  for packed blocks [rows, n] (last dim contiguous) and a contiguous d block
// ---------------------------------------------------------------------------------------------------------------------
__global__ void mul_lastdim_kernel(const MulSeg *__restrict__ segs, const VecWork *__restrict__ work, int nwork,
                                   const double *__restrict__ a, const double *__restrict__ d,
                                   double *__restrict__ out)
{
	for (int w = blockIdx.x; w < nwork; w += gridDim.x)
	{
		const VecWork wk = work[w];
		const MulSeg s = segs[wk.seg];
		const i64 total = s.rows * s.n;
		const i64 end = min(wk.begin + (i64)kVecChunk, total);
		for (i64 e = wk.begin + threadIdx.x; e < end; e += blockDim.x)
		{
			const i64 r = e / s.n, k = e % s.n;
			out[s.o_off + e] = a[s.a_off + r * s.a_row_stride + k * s.a_col_stride] * d[s.d_off + k];
		}
	}
}
void launch_mul_lastdim(Ctx &ctx, const std::vector<MulSeg> &segs, const double *a, const double *d, double *out)
{
	std::vector<VecWork> work;
	for (size_t i = 0; i < segs.size(); ++i)
		for (i64 s = 0; s < segs[i].rows * segs[i].n; s += kVecChunk)
			work.push_back({(int)i, 0, s});
	if (work.empty())
		return;
	auto d_segs = ctx_upload(ctx, segs.data(), segs.size() * sizeof(MulSeg));
	auto d_work = ctx_upload(ctx, work.data(), work.size() * sizeof(VecWork));
	int grid = (int)std::min<size_t>(work.size(), (size_t)ctx.sm_count * 8);
	mul_lastdim_kernel<<<grid, 256, 0, ctx.stream>>>((const MulSeg *)d_segs, (const VecWork *)d_work, (int)work.size(),
	                                                a, d, out);
	QTB_CUDA(cudaGetLastError());
	ctx_free(ctx, d_segs);
	ctx_free(ctx, d_work);
	ctx.counters[0] += 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// the 2x2 Krylov eigenproblem of the one-step Lanczos, on the device (reference eig2x2Mat_impl, dmrg.cpp:543-572)
// scal layout: [0]=a0 [1]=b [2]=a1 -> writes [3]=E0 [4]=o [5]=n [6]=flag (1.0 when a NaN was produced)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void eig2x2_kernel(double *scal)
{
	const double a0 = scal[0], b = scal[1], a1 = scal[2];
	const double crit = sqrt((a0 - a1) * (a0 - a1) + 4.0 * (b * b));
	const double E0 = (a0 + a1 - crit) / 2.0;
	const double delt = E0 - a1;
	double o = sqrt(delt / (-crit));
	const bool zero_o = ((o + E0) == E0) || isnan(o);
	double n = (b * o) / delt;
	if (zero_o)
	{
		n = 1.0;
		o = 0.0;
	}
	scal[3] = E0;
	scal[4] = o;
	scal[5] = n;
	scal[6] = (isnan(o) || isnan(n)) ? 1.0 : 0.0;
}
void launch_eig2x2(Ctx &ctx, double *scal)
{
	eig2x2_kernel<<<1, 1, 0, ctx.stream>>>(scal);
	QTB_CUDA(cudaGetLastError());
	ctx.counters[0] += 1;
}

// b >= 1e-15 ? b : 1   (the reference only divides by b when it is non-singular, dmrg.cpp:597-603)
__global__ void guard_norm_kernel(const double *b, double *out) { out[0] = (fabs(b[0]) >= 1e-15) ? b[0] : 1.0; }
void launch_guard_norm(Ctx &ctx, const double *b, double *out)
{
	guard_norm_kernel<<<1, 1, 0, ctx.stream>>>(b, out);
	QTB_CUDA(cudaGetLastError());
	ctx.counters[0] += 1;
}

} // namespace qtb
