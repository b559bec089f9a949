// DMMA.8x8x4 issue-rate microbenchmark: TFLOP/s vs warps per SM and independent accumulator chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters)
{
	double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
	double c[ILP][2];
#pragma unroll
	for (int i = 0; i < ILP; ++i)
		c[i][0] = c[i][1] = 0;
	for (int it = 0; it < iters; ++it)
	{
#pragma unroll
		for (int i = 0; i < ILP; ++i)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
			             : "+d"(c[i][0]), "+d"(c[i][1])
			             : "d"(a), "d"(b));
	}
	double s = 0;
#pragma unroll
	for (int i = 0; i < ILP; ++i)
		s += c[i][0] + c[i][1];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps, double *d)
{
	int iters = 20000;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	k<ILP><<<148, warps * 32>>>(d, 100);
	cudaEventRecord(e0);
	k<ILP><<<148, warps * 32>>>(d, iters);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double fl = 148.0 * warps * iters * ILP * 512.0;
	printf("warps/SM %2d ILP %2d : %7.2f TFLOP/s\n", warps, ILP, fl / ms / 1e9);
}
int main()
{
	double *d;
	cudaMalloc(&d, 148 * 1024 * 8);
	for (int w : {4, 8, 12, 16, 24, 32})
	{
		run<1>(w, d);
		run<2>(w, d);
		run<4>(w, d);
		run<8>(w, d);
		run<16>(w, d);
		run<32>(w, d);
	}
	return 0;
}
