// reference pattern from the CUDA programming guide (libcu++ barrier + experimental TMA wrappers)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int BW = 32, BH = 32;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, float* out) {
	__shared__ alignas(128) float smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
	__shared__ barrier bar;
	if (threadIdx.x == 0) {
		init(&bar, blockDim.x);
		cde::fence_proxy_async_shared_cta();
	}
	__syncthreads();
	barrier::arrival_token token;
	if (threadIdx.x == 0) {
		cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
		token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
	} else {
		token = bar.arrive();
	}
	bar.wait(std::move(token));
	for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem_buffer[i / BW][i % BW];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
	const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
	int X = argc > 1 ? atoi(argv[1]) : 8, Y = argc > 2 ? atoi(argv[2]) : 4; int prom = argc > 3 ? atoi(argv[3]) : 0;
	const int w = 256, h = 256;
	std::vector<float> himg((size_t)w * h);
	for (int i = 0; i < w * h; i++) himg[i] = (float)(i % 1000);
	float *dimg, *dout;
	cudaMalloc(&dimg, himg.size() * 4);
	cudaMemcpy(dimg, himg.data(), himg.size() * 4, cudaMemcpyHostToDevice);
	cudaMalloc(&dout, BW * BH * 4);
	void* p = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	CUtensorMap map;
	cuuint64_t dims[2] = { (cuuint64_t)w, (cuuint64_t)h };
	cuuint64_t strides[1] = { (cuuint64_t)w * 4 };
	cuuint32_t box[2] = { BW, BH };
	cuuint32_t estr[2] = { 1, 1 };
	CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dimg, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		CU_TENSOR_MAP_SWIZZLE_NONE, prom ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode: %d\n", (int)r);
	kernel<<<1, 128>>>(map, X, Y, dout);
	cudaError_t e = cudaDeviceSynchronize();
	std::vector<float> hout(BW * BH);
	cudaMemcpy(hout.data(), dout, BW * BH * 4, cudaMemcpyDeviceToHost);
	printf("x=%d y=%d prom=%d: sync=%s out[0]=%g out[5*32+3]=%g (expect %g)\n", X, Y, prom, cudaGetErrorString(e), hout[0], hout[5*32+3], himg[(Y+5) * w + X + 3]);
	return 0;
}
