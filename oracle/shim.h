// TEST INFRASTRUCTURE ONLY. Force-included (-include) when compiling the UNMODIFIED reference sources
// under /root/reference against torch 2.11: restores two C++-frontend names that newer libtorch dropped.
// Used at: reference sources/btensor_linalg.cpp:354, sources/LinearAlgebra.cpp:144, sources/btensor.cpp:2519.
#pragma once
#include <torch/torch.h>
namespace torch
{
namespace linalg
{
inline std::tuple<at::Tensor, at::Tensor> eigh(const at::Tensor &a, c10::string_view uplo)
{
	return at::linalg_eigh(a, uplo);
}
inline at::Tensor vector_norm(const at::Tensor &a, const at::Scalar &ord, at::OptionalIntArrayRef dim, bool keepdim,
                              std::optional<at::ScalarType> dt)
{
	return at::linalg_vector_norm(a, ord, dim, keepdim, dt);
}
} // namespace linalg
} // namespace torch
