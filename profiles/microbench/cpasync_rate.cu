// cp.async (LDGSTS) throughput per SM as a function of the copy size: 8-byte .ca copies (what the grouped GEMM's producer
// issues: operand rows start at arbitrary 8-byte offsets) against 16-byte .cg copies, same bytes, same access pattern
// (each warp instruction covers two contiguous 128-byte / 256-byte row segments), data resident in L2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cpasync_rate cpasync_rate.cu ; run: ./cpasync_rate
#include <cuda_runtime.h>
#include <cstdio>

template <int BYTES>
__global__ void __launch_bounds__(128) copy_kernel(const double *__restrict__ src, int rows_total, int iters)
{
	extern __shared__ __align__(16) double smem[];
	// tile: 64 rows x 16 doubles (one GEMM operand chunk), row stride 20 doubles in smem; 128 threads
	const int tid = threadIdx.x;
	const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
	long row0 = (long)blockIdx.x * 64;
	for (int it = 0; it < iters; ++it)
	{
		const int stage = it & 3;
		const unsigned sdst = sbase + stage * (64 * 20 * 8);
		if (BYTES == 8)
		{
			const int k = tid & 15, rb = tid >> 4;
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				const int r = rb + i * 8;
				const double *g = src + ((row0 + r) % rows_total) * 16 + k;
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sdst + (r * 20 + k) * 8), "l"(g));
			}
		}
		else
		{
			const int k2 = tid & 7, rb = tid >> 3; // 8 x 16-byte copies per row
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				const int r = rb + i * 16;
				const double *g = src + ((row0 + r) % rows_total) * 16 + 2 * k2;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sdst + (r * 20 + 2 * k2) * 8), "l"(g));
			}
		}
		asm volatile("cp.async.commit_group;\n");
		asm volatile("cp.async.wait_group 3;\n");
		row0 += (long)gridDim.x * 64;
	}
	asm volatile("cp.async.wait_all;\n");
	if (smem[tid] == 123.456)
		printf("x");
}

int main()
{
	const int rows_total = 1 << 20; // 128 MB > L2? 1M rows x 128 B = 128 MB; use 512K rows = 64 MB to stay in L2
	double *src;
	cudaMalloc(&src, (size_t)rows_total * 16 * 8);
	cudaMemset(src, 0, (size_t)rows_total * 16 * 8);
	const int iters = 2000, grid = 148 * 2;
	for (int rep = 0; rep < 2; ++rep)
		for (int bytes : {8, 16})
		{
			cudaEvent_t e0, e1;
			cudaEventCreate(&e0);
			cudaEventCreate(&e1);
			cudaEventRecord(e0);
			if (bytes == 8)
				copy_kernel<8><<<grid, 128, 4 * 64 * 20 * 8>>>(src, rows_total / 2, iters);
			else
				copy_kernel<16><<<grid, 128, 4 * 64 * 20 * 8>>>(src, rows_total / 2, iters);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms;
			cudaEventElapsedTime(&ms, e0, e1);
			const double gb = (double)grid * iters * 64 * 16 * 8 / 1e9;
			printf("cp.async %2d-byte: %.3f ms, %.1f GB/s total, %.2f bytes/clk/SM (1.965 GHz, 148 SMs)\n", bytes, ms, gb / (ms * 1e-3),
			       gb * 1e9 / (ms * 1e-3) / 148 / 1.965e9);
		}
	return 0;
}
