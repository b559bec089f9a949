// consumer inner-loop microbenchmark: LDS fragments + DMMA.8x8x4 on a static smem tile (no producer, no barriers)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int BM, int BN, int WM, int WN, int MODE>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32) k(double *out, int steps)
{
	constexpr int BK = 16, PAD = 4;
	extern __shared__ double smem[];
	double *As = smem, *Bs = smem + BM * (BK + PAD);
	for (int i = threadIdx.x; i < BM * (BK + PAD) + BK * (BN + PAD); i += blockDim.x)
		smem[i] = i * 1e-6;
	__syncthreads();
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
	constexpr int WN_ = BN / WN;
	const int wm0 = (warp / WN_) * WM, wn0 = (warp % WN_) * WN;
	constexpr int MI = WM / 8, NI = WN / 8;
	double acc[MI][NI][2];
#pragma unroll
	for (int i = 0; i < MI; ++i)
#pragma unroll
		for (int j = 0; j < NI; ++j)
			acc[i][j][0] = acc[i][j][1] = 0;
	const int sa_m = MODE ? (BK + PAD) : (BK + PAD) + (steps >> 30), sa_k = 1 + (MODE ? 0 : (steps >> 30));
	const int sb_k = (BN + PAD) + (MODE ? 0 : (steps >> 30)), sb_n = 1 + (MODE ? 0 : (steps >> 30));
	const double *Ap = As + (wm0 + g) * sa_m + q * sa_k;
	const double *Bp = Bs + q * sb_k + (wn0 + g) * sb_n;
	for (int s = 0; s < steps; ++s)
	{
#pragma unroll
		for (int kk = 0; kk < BK; kk += 4)
		{
			double af[MI], bf[NI];
#pragma unroll
			for (int i = 0; i < MI; ++i)
				af[i] = Ap[i * 8 * sa_m + kk * sa_k];
#pragma unroll
			for (int j = 0; j < NI; ++j)
				bf[j] = Bp[kk * sb_k + j * 8 * sb_n];
#pragma unroll
			for (int i = 0; i < MI; ++i)
#pragma unroll
				for (int j = 0; j < NI; ++j)
					dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
		}
	}
	double sum = 0;
#pragma unroll
	for (int i = 0; i < MI; ++i)
#pragma unroll
		for (int j = 0; j < NI; ++j)
			sum += acc[i][j][0] + acc[i][j][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}
template <int BM, int BN, int WM, int WN, int MODE>
void run(const char *name, int ctas_per_sm, double *d)
{
	constexpr int NT = (BM / WM) * (BN / WN) * 32;
	size_t sm = (BM * 20 + 16 * (BN + 4)) * 8;
	auto kern = k<BM, BN, WM, WN, MODE>;
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
	int steps = 4000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	kern<<<148 * ctas_per_sm, NT, sm>>>(d, 10);
	cudaEventRecord(e0);
	kern<<<148 * ctas_per_sm, NT, sm>>>(d, steps);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double fl = 148.0 * ctas_per_sm * steps * 2.0 * BM * BN * 16;
	printf("%-40s ctas/SM %d threads %4d : %7.2f TFLOP/s (%s)\n", name, ctas_per_sm, NT, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
	double *d;
	cudaMalloc(&d, 148 * 4 * 1024 * 8);
	run<128, 128, 64, 32, 1>("128x128 warp 64x32 const strides", 1, d);
	run<128, 128, 64, 32, 0>("128x128 warp 64x32 runtime strides", 1, d);
	run<128, 128, 32, 32, 1>("128x128 warp 32x32 (16 warps)", 1, d);
	run<128, 128, 32, 64, 1>("128x128 warp 32x64", 1, d);
	run<64, 64, 32, 32, 1>("64x64 warp 32x32", 1, d);
	run<64, 64, 32, 32, 1>("64x64 warp 32x32", 2, d);
	run<64, 64, 32, 32, 1>("64x64 warp 32x32", 3, d);
	run<128, 64, 32, 32, 1>("128x64 warp 32x32 (8 warps)", 1, d);
	run<128, 64, 32, 32, 1>("128x64 warp 32x32 (8 warps)", 2, d);
	run<64, 64, 16, 32, 1>("64x64 warp 16x32 (8 warps)", 2, d);
	return 0;
}
