// qtb_svd.cu — block SVD + truncation (placeholder until the Jacobi kernels land in this round)
#include "qtb_ops.h"
namespace qtb
{
void block_svd(Ctx &, const Tensor &, i64, bool, double, i64, i64, double, std::unique_ptr<Tensor> &,
               std::unique_ptr<Tensor> &, std::unique_ptr<Tensor> &)
{
	throw Error(QTB_ERR_RUNTIME, "qtb_svd: not implemented in this build");
}
} // namespace qtb
