// qtb_ops.h — block-tensor vector ops and two-site DMRG pieces (qtb_ops.cpp), block SVD (qtb_svd.cu)
#pragma once
#include <memory>

#include "qtb_core.h"
#include "qtb_vec.h"

namespace qtb
{
// out = (ca)*a + (cb)*b over the union of the block lists; coefficients = (*ptr or 1) * mul, read on the device.
// divide_a: a's term is a / ca instead of ca * a.
// merge_quirk != 1: reproduces the reference's flat_map::merge behaviour (flat_map.h:350-425, last for_each): the blocks
// of `a` that sort before every block of `b` are additionally multiplied by the add_ coefficient (= merge_quirk).
std::unique_ptr<Tensor> axpby_dev(Ctx &ctx, const double *ca_ptr, double ca_mul, const Tensor &a,
                                  const double *cb_ptr, double cb_mul, const Tensor &b, bool divide_a,
                                  double merge_quirk = 1.0);
void dot_dev(Ctx &ctx, const Tensor &a, const Tensor &b, double *d_result, bool take_sqrt);
std::unique_ptr<Tensor> mul_lastdim(Ctx &ctx, const Tensor &a, const Tensor &d);
std::unique_ptr<Tensor> heff_apply(Ctx &ctx, const Tensor &psi, const Tensor &h2, const Tensor &lenv,
                                   const Tensor &renv);
std::unique_ptr<Tensor> env_left(Ctx &ctx, const Tensor &h, const Tensor &mps, const Tensor &lenv);
std::unique_ptr<Tensor> env_right(Ctx &ctx, const Tensor &h, const Tensor &mps, const Tensor &renv);
std::unique_ptr<Tensor> tensordot_sharded(Ctx &ctx, const Tensor &a, const Tensor &b, const std::vector<i64> &da,
                                          const std::vector<i64> &db, i64 owner_dim);
std::unique_ptr<Tensor> two_sites_update(Ctx &ctx, const Tensor &psi, const Tensor &h2, const Tensor &lenv,
                                         const Tensor &renv, double *energy);
// qtb_svd.cu
void block_svd(Ctx &ctx, const Tensor &a, i64 split, bool truncate, double tol, i64 min_size, i64 max_size,
               double pow, std::unique_ptr<Tensor> &u, std::unique_ptr<Tensor> &d, std::unique_ptr<Tensor> &v);
void block_eigh(Ctx &ctx, const Tensor &a, i64 split, bool truncate, double tol, i64 min_size, i64 max_size, double pow,
                std::unique_ptr<Tensor> &e, std::unique_ptr<Tensor> &u);
void block_truncate(Ctx &ctx, const Tensor &d, const std::vector<const Tensor *> &units, double tol, i64 min_size, i64 max_size,
                    double pow, std::unique_ptr<Tensor> &d_out, std::vector<std::unique_ptr<Tensor>> &units_out);
// qtb_shape.cpp
std::unique_ptr<Tensor> reshape(Ctx &ctx, const Tensor &a, const std::vector<i64> &index_groups);
std::unique_ptr<Tensor> reshape_as(Ctx &ctx, const Tensor &a, const Tensor &like, bool overwrite_cvals);
std::unique_ptr<Tensor> tensorgdot(Ctx &ctx, const Tensor &c, const Tensor &a, const Tensor &b, const std::vector<i64> &da,
                                   const std::vector<i64> &db, double beta, double alpha);
void save_tensor(Ctx &ctx, const Tensor &t, const char *path);
std::unique_ptr<Tensor> load_tensor(Ctx &ctx, const char *path);
// qtb_dmrg.cpp
// <a|obs|b> (obs != nullptr) or <a|b>: reference contract(bMPS, bMPS[, bMPO]), sources/MPT.cpp:211-233, 275-292
double contract(Ctx &ctx, i64 L, const Tensor *const *a, const Tensor *const *b, const Tensor *const *obs);
// reference bMPS::move_oc, sources/MPT.cpp:75-111 (untruncated block SVDs; site tensors are replaced in place)
void move_oc(Ctx &ctx, std::vector<std::unique_ptr<Tensor>> &mps, i64 &oc, i64 target);
// reference bMPO::coalesce(cutoff), sources/MPT.cpp:154-168 (MPO tensors replaced in place)
void coalesce(Ctx &ctx, std::vector<std::unique_ptr<Tensor>> &mpo, double cutoff);
void dmrg(Ctx &ctx, i64 L, const Tensor *const *mpo, std::vector<std::unique_ptr<Tensor>> &mps, i64 &oc,
          const qtb_dmrg_options &opt, double &energy, i64 &n_sweeps, double *sweep_energy, double *sweep_seconds,
          i64 *sweep_mid_bond, qtb_dmrg_log_fn log_fn = nullptr, void *log_user = nullptr);
} // namespace qtb
