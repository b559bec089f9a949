// qtb_svd_qr.cuh — QR preconditioning of the batched one-sided block Jacobi SVD (included by qtb_svd.cu only).
//
// One-sided Jacobi applied to A itself needs 13-18 outer sweeps on the graded matrices a DMRG theta produces once a
// charge group has ~64 column blocks (measured: profiles/r2/s1_svd_inner.txt; reproduced by the numpy model
// profiles/r2/jacobi_model.py). Following Drmac & Veselic ("New fast and accurate Jacobi SVD algorithm", SIAM J. Matrix
// Anal. Appl. 29, 2008) the iteration is run on X = R^T of a Householder factorisation A = Q R instead: R R^T is far
// closer to diagonal than A^T A and the same block Jacobi converges in 6-7 sweeps (the last two almost empty).
//   A = Q R,   R^T J = W S  (Jacobi, J accumulated under R^T)   =>   A = (Q J) S W^T :  U = Q J,  V = W.
// No accuracy is traded: Q is a product of Householder reflectors (orthogonal to rounding whatever the conditioning of
// A), and the left vectors U = Q J come out orthonormal even for zero singular values.
//
// Kernels (all groups of a call batched in every launch; fp64 tensor cores for the O(m n^2) parts):
//   qr_panel_kernel : Householder factorisation of one 32-column panel per group, the panel resident in the shared
//                     memory of a thread-block CLUSTER of 8 CTAs (row slabs), per column ONE all-to-all exchange of
//                     32 partial dot products through distributed shared memory + one cluster barrier; the same dot
//                     products give the compact-WY factor T (larft) for free.
//   qr_w_kernel     : W = V^T C      (DMMA, partial sums per 512-row chunk, summed in a fixed order: reproducible)
//   qr_tw_kernel    : TW = op(T) W   (tiny)
//   qr_apply_kernel : C -= V TW      (DMMA)
// used twice: for the trailing matrix during the factorisation (op(T) = T^T) and, after the Jacobi iteration, to apply
// Q to [J; 0] (op(T) = T, panels in reverse order).
#pragma once
#include <cooperative_groups.h>

namespace qtb
{
namespace
{
namespace cg = cooperative_groups;

constexpr int kQrB = 32;          // panel width
constexpr int kQrCluster = 8;     // CTAs per panel (portable cluster size)
constexpr int kQrSlabMax = 832;   // rows of a panel one CTA holds: 32 x 833 doubles = 213 KB
constexpr int kQrThreads = 256;
constexpr int kQrWRows = 512;     // rows per partial sum of W = V^T C
constexpr int kQrARows = 64;      // rows per CTA of the apply kernel

struct QrGroup
{
	i64 a_off;  // F (m x n, column-major, ld = m): on exit R in the upper triangle, the Householder vectors below it
	i64 u_off;  // U workspace (m x n, column-major, ld = m)
	i64 t_off;  // T factors: npanels x 32 x 32, row-major
	i64 w_off;  // partial W: nchunks x 32 x n
	i64 tw_off; // op(T) W: 32 x n
	i64 x_off;  // Jacobi workspace X = [R^T ; I] (2n x n, column-major, ld = 2n)
	int m, n;
};

__device__ __forceinline__ void qr_dmma(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
	             : "+d"(c0), "+d"(c1)
	             : "d"(a), "d"(b));
}

// shared memory of the panel kernel (doubles): P[32][SP] | exch[2][9][32] | T[32][33] | part[32]
__host__ __device__ inline size_t qr_panel_smem(int SP) { return (size_t)(32 * SP + 2 * 9 * 32 + 32 * 33 + 32) * sizeof(double); }
__host__ __device__ inline int qr_slab_rows(int mrows, int cluster)
{ // rows per CTA of the cluster; at least the panel width so that every pivot row lives in CTA 0
	int s = ((mrows + cluster - 1) / cluster + 7) & ~7;
	return s < kQrB ? kQrB : s;
}

// cluster size (launch attribute): kQrCluster for tall panels, 1 when the whole panel fits one CTA's shared memory (the
// per-column cluster barrier is then a CTA barrier: small groups factorise ~2x faster)
__global__ void __launch_bounds__(kQrThreads)
    qr_panel_kernel(const QrGroup *__restrict__ groups, double *__restrict__ ws, int k, int SPmax)
{
	extern __shared__ double qsm[];
	cg::cluster_group cluster = cg::this_cluster();
	const int rank = (int)cluster.block_rank();
	const int CS = (int)cluster.num_blocks();
	const QrGroup G = groups[blockIdx.x / CS];
	const int r0 = k * kQrB;
	const int w = min(kQrB, G.n - r0);
	const int mrows = G.m - r0;
	const int S = qr_slab_rows(mrows, CS);
	const int SP = SPmax; // uniform carve-up of the dynamic shared memory
	double *P = qsm;
	double *exch = P + 32 * SP;
	double *T = exch + 2 * 9 * 32;
	double *part = T + 32 * 33;
	const int lr0 = rank * S;
	const int nloc = max(0, min(S, mrows - lr0));
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	double *Ag = ws + G.a_off;
	for (int c = warp; c < 32; c += 8)
		for (int r = lane; r < nloc; r += 32)
			P[c * SP + r] = c < w ? Ag[(i64)(r0 + c) * G.m + r0 + lr0 + r] : 0.0;
	for (int e = tid; e < 32 * 33; e += kQrThreads)
		T[e] = 0.0;
	__syncthreads();

	for (int j = 0; j < w; ++j)
	{
		const int par = j & 1;
		const double *x = P + j * SP;
		const int rbeg = rank == 0 ? j + 1 : 0; // rows strictly below the pivot
		// ---- partial dot products x^T p_c over the rows below the pivot, all 32 columns: c > j gives the reflector's
		// action on the rest of the panel, c < j gives V^T v_j (the T factor), c = j the norm ----
		{
			double acc[4] = {0.0, 0.0, 0.0, 0.0};
			for (int r = rbeg + lane; r < nloc; r += 32)
			{
				const double xv = x[r];
#pragma unroll
				for (int t = 0; t < 4; ++t)
					acc[t] += xv * P[(warp + 8 * t) * SP + r];
			}
#pragma unroll
			for (int t = 0; t < 4; ++t)
			{
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
				if (lane == 0)
					part[warp + 8 * t] = acc[t];
			}
		}
		__syncthreads();
		{ // all-to-all: thread (dest, c) stores this CTA's partial (and, from CTA 0, the pivot row) into CTA `dest`
			const int dest = tid >> 5, c = tid & 31;
			if (dest < CS)
			{
				double *remote = cluster.map_shared_rank(exch, dest);
				remote[(par * 9 + rank) * 32 + c] = part[c];
				if (rank == 0)
					remote[(par * 9 + 8) * 32 + c] = P[c * SP + j];
			}
		}
		cluster.sync();
		const double *E = exch + par * 9 * 32;
		double dj = 0.0;
		for (int rk = 0; rk < CS; ++rk)
			dj += E[rk * 32 + j];
		const double alpha = E[8 * 32 + j];
		double tau = 0.0, inv = 0.0, beta = alpha;
		if (dj > 0.0)
		{
			const double nrm = sqrt(alpha * alpha + dj);
			beta = alpha >= 0.0 ? -nrm : nrm;
			inv = 1.0 / (alpha - beta);
			tau = (beta - alpha) / beta;
		}
		if (tau != 0.0)
		{
#pragma unroll
			for (int t = 0; t < 4; ++t)
			{
				const int c = warp + 8 * t;
				if (c > j && c < w)
				{
					double dc = 0.0;
					for (int rk = 0; rk < CS; ++rk)
						dc += E[rk * 32 + c];
					const double wc = E[8 * 32 + c] + dc * inv; // v^T p_c
					const double f = tau * wc * inv;
					double *pc = P + c * SP;
					for (int r = rbeg + lane; r < nloc; r += 32)
						pc[r] -= f * x[r];
					if (rank == 0 && lane == 0)
						pc[j] -= tau * wc;
				}
			}
		}
		if (rank == 0 && warp == 0)
		{ // T(0:j, j) = -tau T(0:j, 0:j) (V^T v_j),  T(j, j) = tau      (LAPACK dlarft, forward / columnwise)
			double s = 0.0;
			if (lane < j)
			{
				double dc = 0.0;
				for (int rk = 0; rk < CS; ++rk)
					dc += E[rk * 32 + lane];
				s = E[8 * 32 + lane] + dc * inv; // v_lane^T v_j
			}
			double tij = 0.0;
			for (int kk = 0; kk < j; ++kk)
			{
				const double sk = __shfl_sync(0xffffffffu, s, kk);
				if (lane <= kk)
					tij += T[lane * 33 + kk] * sk;
			}
			if (lane < j)
				T[lane * 33 + j] = -tau * tij;
			if (lane == j)
				T[j * 33 + j] = tau;
		}
		__syncthreads(); // every reader of x is done
		if (warp == (j & 7))
		{ // column j becomes (beta, v(2:))
			double *xj = P + j * SP;
			if (tau != 0.0)
				for (int r = rbeg + lane; r < nloc; r += 32)
					xj[r] *= inv;
			if (rank == 0 && lane == 0)
				xj[j] = beta;
		}
		__syncthreads();
	}
	for (int c = warp; c < w; c += 8)
		for (int r = lane; r < nloc; r += 32)
			Ag[(i64)(r0 + c) * G.m + r0 + lr0 + r] = P[c * SP + r];
	if (rank == 0)
	{
		double *Tg = ws + G.t_off + (i64)k * 1024;
		for (int e = tid; e < 1024; e += kQrThreads)
			Tg[e] = T[(e >> 5) * 33 + (e & 31)];
	}
}

// Householder vector entry of panel k: unit lower trapezoidal, the stored upper part belongs to R
__device__ __forceinline__ double qr_vmask(int lrow, int c, double val) { return lrow < c ? 0.0 : (lrow == c ? 1.0 : val); }

constexpr int kQrLd = 64 + 4;  // == 4 mod 16: conflict-free DMMA fragment loads
constexpr int kQrSub = 32;     // rows of V / C staged at a time by the W kernel
constexpr int kQrLdW = kQrSub + 4;

// W_part[chunk][c][col] = sum over the chunk's rows of V(r, c) C(r, col).  grid = (column tiles of 64, row chunks of 512,
// groups); 8 warps as 4 (c blocks of 8) x 2 (32 columns).  mode 0: C = trailing matrix of F (columns >= 32(k+1));
// mode 1: C = U workspace, all columns.
__global__ void __launch_bounds__(256) qr_w_kernel(const QrGroup *__restrict__ groups, double *__restrict__ ws, int k, int mode)
{
	__shared__ double sV[32 * kQrLdW];
	__shared__ double sC[64 * kQrLdW];
	const QrGroup G = groups[blockIdx.z];
	const int r0 = k * kQrB, mrows = G.m - r0, w = min(kQrB, G.n - r0);
	const int cbeg = mode == 0 ? r0 + kQrB : 0, cend = G.n;
	const int col0 = cbeg + 64 * blockIdx.x;
	const int rb = kQrWRows * blockIdx.y;
	if (col0 >= cend || rb >= mrows)
		return;
	const double *Vg = ws + G.a_off;
	const double *Cg = ws + (mode == 0 ? G.a_off : G.u_off);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
	const int ar = warp >> 1, ac = (warp & 1) * 32;
	double acc[4][2];
#pragma unroll
	for (int j = 0; j < 4; ++j)
		acc[j][0] = acc[j][1] = 0.0;
	for (int sub = 0; sub < kQrWRows; sub += kQrSub)
	{
		const int nr = min(kQrSub, mrows - rb - sub);
		if (nr <= 0)
			break;
		for (int e = threadIdx.x; e < 32 * kQrSub; e += 256)
		{
			const int c = e / kQrSub, r = e % kQrSub;
			double v = 0.0;
			if (c < w && r < nr)
			{
				const int lrow = rb + sub + r;
				v = qr_vmask(lrow, c, Vg[(i64)(r0 + c) * G.m + r0 + lrow]);
			}
			sV[c * kQrLdW + r] = v;
		}
		for (int e = threadIdx.x; e < 64 * kQrSub; e += 256)
		{
			const int c = e / kQrSub, r = e % kQrSub;
			double v = 0.0;
			if (col0 + c < cend && r < nr)
				v = Cg[(i64)(col0 + c) * G.m + r0 + rb + sub + r];
			sC[c * kQrLdW + r] = v;
		}
		__syncthreads();
#pragma unroll 4
		for (int kk = 0; kk < kQrSub; kk += 4)
		{
			const double af = sV[(ar * 8 + g) * kQrLdW + kk + q];
			double bf[4];
#pragma unroll
			for (int j = 0; j < 4; ++j)
				bf[j] = sC[(ac + j * 8 + g) * kQrLdW + kk + q];
#pragma unroll
			for (int j = 0; j < 4; ++j)
				qr_dmma(acc[j][0], acc[j][1], af, bf[j]);
		}
		__syncthreads();
	}
	double *Wp = ws + G.w_off + ((i64)blockIdx.y * 32 + ar * 8 + g) * G.n;
#pragma unroll
	for (int j = 0; j < 4; ++j)
	{
		const int col = col0 + ac + j * 8 + 2 * q;
		if (col < cend)
			Wp[col] = acc[j][0];
		if (col + 1 < cend)
			Wp[col + 1] = acc[j][1];
	}
}

// TW = op(T) (sum over chunks of W_part).  grid = (column tiles of 64, groups).  transT: op(T) = T^T (applying Q^T)
__global__ void __launch_bounds__(256) qr_tw_kernel(const QrGroup *__restrict__ groups, double *__restrict__ ws, int k, int mode)
{
	__shared__ double sW[32 * 65];
	__shared__ double sT[32 * 33];
	const QrGroup G = groups[blockIdx.y];
	const int r0 = k * kQrB, mrows = G.m - r0;
	const int cbeg = mode == 0 ? r0 + kQrB : 0, cend = G.n;
	const int col0 = cbeg + 64 * blockIdx.x;
	if (col0 >= cend)
		return;
	const int nch = (mrows + kQrWRows - 1) / kQrWRows;
	const double *Tg = ws + G.t_off + (i64)k * 1024;
	for (int e = threadIdx.x; e < 1024; e += 256)
		sT[(e >> 5) * 33 + (e & 31)] = Tg[e];
	for (int e = threadIdx.x; e < 32 * 64; e += 256)
	{
		const int c = e >> 6, col = col0 + (e & 63);
		double v = 0.0;
		if (col < cend)
			for (int ch = 0; ch < nch; ++ch)
				v += ws[G.w_off + ((i64)ch * 32 + c) * G.n + col];
		sW[c * 65 + (e & 63)] = v;
	}
	__syncthreads();
	double *TW = ws + G.tw_off;
	for (int e = threadIdx.x; e < 32 * 64; e += 256)
	{
		const int i = e >> 6, cc = e & 63;
		if (col0 + cc >= cend)
			continue;
		double v = 0.0;
		if (mode == 0)
		{ // T^T: sum_c T(c, i) W(c)   (T upper triangular: c <= i)
			for (int c = 0; c <= i; ++c)
				v += sT[c * 33 + i] * sW[c * 65 + cc];
		}
		else
		{
			for (int c = i; c < 32; ++c)
				v += sT[i * 33 + c] * sW[c * 65 + cc];
		}
		TW[(i64)i * G.n + col0 + cc] = v;
	}
}

// C(rows, cols) -= V(rows, 0:32) TW(0:32, cols).  grid = (column tiles of 64, row chunks of 64, groups), 8 warps x 8 rows
constexpr int kQrLdA = kQrARows + 4;
__global__ void __launch_bounds__(256) qr_apply_kernel(const QrGroup *__restrict__ groups, double *__restrict__ ws, int k, int mode)
{
	__shared__ double sV[32 * kQrLdA];
	__shared__ double sTW[32 * kQrLd];
	const QrGroup G = groups[blockIdx.z];
	const int r0 = k * kQrB, mrows = G.m - r0, w = min(kQrB, G.n - r0);
	const int cbeg = mode == 0 ? r0 + kQrB : 0, cend = G.n;
	const int col0 = cbeg + 64 * blockIdx.x;
	const int rb = kQrARows * blockIdx.y;
	if (col0 >= cend || rb >= mrows)
		return;
	const int nr = min(kQrARows, mrows - rb);
	const double *Vg = ws + G.a_off;
	double *Cg = ws + (mode == 0 ? G.a_off : G.u_off);
	const double *TW = ws + G.tw_off;
	for (int e = threadIdx.x; e < 32 * kQrARows; e += 256)
	{
		const int c = e / kQrARows, r = e % kQrARows;
		double v = 0.0;
		if (c < w && r < nr)
			v = qr_vmask(rb + r, c, Vg[(i64)(r0 + c) * G.m + r0 + rb + r]);
		sV[c * kQrLdA + r] = v;
	}
	for (int e = threadIdx.x; e < 32 * 64; e += 256)
	{
		const int i = e >> 6, cc = e & 63;
		sTW[i * kQrLd + cc] = col0 + cc < cend ? TW[(i64)i * G.n + col0 + cc] : 0.0;
	}
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
	const int rw0 = warp * 8;
	double acc[1][8][2];
#pragma unroll
	for (int j = 0; j < 8; ++j)
		acc[0][j][0] = acc[0][j][1] = 0.0;
#pragma unroll 2
	for (int kk = 0; kk < 32; kk += 4)
	{
		double bf[8];
		const double af = sV[(kk + q) * kQrLdA + rw0 + g];
#pragma unroll
		for (int j = 0; j < 8; ++j)
			bf[j] = sTW[(kk + q) * kQrLd + j * 8 + g];
#pragma unroll
		for (int j = 0; j < 8; ++j)
			qr_dmma(acc[0][j][0], acc[0][j][1], af, bf[j]);
	}
#pragma unroll
	for (int i = 0; i < 1; ++i)
	{
		const int r = rw0 + i * 8 + g;
		if (r < nr)
		{
#pragma unroll
			for (int j = 0; j < 8; ++j)
#pragma unroll
				for (int h = 0; h < 2; ++h)
				{
					const int col = col0 + j * 8 + 2 * q + h;
					if (col < cend)
						Cg[(i64)col * G.m + r0 + rb + r] -= acc[i][j][h];
				}
		}
	}
}

// X(0:n, 0:n) = R^T (lower triangular), 32 x 32 tiles through shared memory.  grid = (tiles, tiles, groups)
__global__ void __launch_bounds__(256) qr_rt_kernel(const QrGroup *__restrict__ groups, double *__restrict__ ws)
{
	__shared__ double tile[32][33];
	const QrGroup G = groups[blockIdx.z];
	const int tr = blockIdx.x, tc = blockIdx.y; // R tile (rows tr, columns tc), needed when tc >= tr
	if (tr * 32 >= G.n || tc * 32 >= G.n || tc < tr)
		return;
	const double *Ag = ws + G.a_off;
	double *Xg = ws + G.x_off;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	for (int cc = ty; cc < 32; cc += 8)
	{
		const int r = tr * 32 + tx, c = tc * 32 + cc;
		tile[cc][tx] = (r < G.n && c < G.n && r <= c) ? Ag[(i64)c * G.m + r] : 0.0;
	}
	__syncthreads();
	for (int rr = ty; rr < 32; rr += 8)
	{ // X(c, r) = R(r, c): X column r = tr * 32 + rr, X row c = tc * 32 + tx
		const int r = tr * 32 + rr, c = tc * 32 + tx;
		if (r < G.n && c < G.n)
			Xg[(i64)r * (2 * G.n) + c] = tile[tx][rr];
	}
}

// U(0:n, :) = J (the rotations accumulated under R^T), the rows below zero (written here: the U workspace doubles as the
// scratch of the column pre-sort).  grid = (column chunks, groups)
__global__ void __launch_bounds__(256) qr_j_kernel(const QrGroup *__restrict__ groups, double *__restrict__ ws)
{
	const QrGroup G = groups[blockIdx.y];
	const double *Xg = ws + G.x_off;
	double *Ug = ws + G.u_off;
	for (int c = blockIdx.x; c < G.n; c += gridDim.x)
		for (int r = threadIdx.x; r < G.m; r += 256)
			Ug[(i64)c * G.m + r] = r < G.n ? Xg[(i64)c * (2 * G.n) + G.n + r] : 0.0;
}

} // namespace
} // namespace qtb
