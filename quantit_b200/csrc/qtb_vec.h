// qtb_vec.h — descriptors and launchers of the HBM-bound arena kernels (qtb_vec.cu)
#pragma once
#include <vector>

#include "qtb_core.h"

namespace qtb
{
constexpr int kGatherChunk = 4096; // elements per work item
constexpr int kVecChunk = 8192;

struct GatherDesc
{
	i64 src_off, dst_off, numel;
	int rank;
	int pad_;
	i64 dims[8];
	i64 strides[8];
};
struct GatherWork
{
	int desc;
	int pad_;
	i64 begin;
};
struct VecSeg
{
	i64 a_off, b_off, o_off, n; // a_off / b_off < 0: operand absent on this segment
	double a_scale;             // extra factor on the a term of this segment (1 except for the reference's merge quirk)
};
struct VecWork
{
	int seg;
	int pad_;
	i64 begin;
};
struct MulSeg
{
	i64 a_off, d_off, o_off, rows, n, a_row_stride, a_col_stride;
};

// stream-ordered scratch memory + small uploads through a pinned ring
void *ctx_alloc(Ctx &ctx, size_t bytes);
void ctx_free(Ctx &ctx, void *p);
void *ctx_upload(Ctx &ctx, const void *host, size_t bytes);

void launch_gather(Ctx &ctx, const std::vector<GatherDesc> &descs, const double *src, double *dst);
void launch_axpby(Ctx &ctx, const std::vector<VecSeg> &segs, const double *a, const double *b, double *out,
                  const double *ca_ptr, double ca_mul, const double *cb_ptr, double cb_mul, bool divide_a);
void launch_dot(Ctx &ctx, const std::vector<VecSeg> &segs, const double *a, const double *b, double *d_result,
                bool take_sqrt);
void launch_scale(Ctx &ctx, double *x, i64 n, const double *c_ptr, double c_mul, bool divide);
void launch_mul_lastdim(Ctx &ctx, const std::vector<MulSeg> &segs, const double *a, const double *d, double *out);
void launch_eig2x2(Ctx &ctx, double *scal);
void launch_guard_norm(Ctx &ctx, const double *b, double *out);
} // namespace qtb
