// mia_peaks.cu -- micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not carry:
// the pair kernels are bound by FP64 ALU issue, so bench.py measures the dependent-free DFMA rate of this GPU live and
// reports roofline.frac against it.  Separate tiny library (libmia_peaks.so): not part of the drop-in C ABI.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
// 8 independent DFMA chains per thread; 2 flop per DFMA.
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
	double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3,
		   x6 = x0 + 6e-3, x7 = x0 + 7e-3;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			x0 = fma(x0, a, b);
			x1 = fma(x1, a, b);
			x2 = fma(x2, a, b);
			x3 = fma(x3, a, b);
			x4 = fma(x4, a, b);
			x5 = fma(x5, a, b);
			x6 = fma(x6, a, b);
			x7 = fma(x7, a, b);
		}
	}
	out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

extern "C" {
// Runs `reps` launches of the DFMA kernel (after one warm-up launch) and returns the best TFLOP/s; <0 on CUDA error.
double mia_peak_fp64_tflops(int reps, int iters, double *scratch /* device, >= blocks*256 doubles */, int blocks) {
	cudaEvent_t e0, e1;
	if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
	double best = 0.0;
	for (int r = 0; r <= reps; r++) {
		cudaEventRecord(e0);
		k_dfma<<<blocks, 256>>>(scratch, iters, 0.999999, 1e-7);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		double flop = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
		double tf = flop / (ms * 1e-3) / 1e12;
		if (r > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return best;
}
}
