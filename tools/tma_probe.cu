// tma_probe.cu -- standalone probe of 2D TMA tile loads (descriptor placement / smem placement variants).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu ; run on a B200.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WAIT_DONE;\n\tbra.uni WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(
			smem_u32(bar)),
		"r"(parity)
		: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
		"l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
		: "memory");
}

__global__ void probe_bulk1d(const float* src, float* out, int n) {
	__shared__ __align__(128) float buf[1024];
	__shared__ __align__(8) uint64_t bar;
	const int lane = threadIdx.x & 31;
	if (lane == 0) mbar_init(&bar, 1);
	__syncwarp();
	if (lane == 0) {
		mbar_expect_tx(&bar, n * 4);
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)), "l"(src),
			"r"(n * 4), "r"(smem_u32(&bar))
			: "memory");
	}
	mbar_wait(&bar, 0);
	for (int i = lane; i < n; i += 32) out[i] = buf[i];
}

// variant 0: descriptor as __grid_constant__ param; variant 1: descriptor in global memory (gmap)
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, int variant, int bw, int bh, int x, int y, float* out, int off_floats) {
	extern __shared__ __align__(128) float smem[];
	float* slab = smem + off_floats;
	uint64_t* bar = (uint64_t*)slab;
	float* T = slab + 32;
	const int lane = threadIdx.x & 31;
	if (lane == 0) mbar_init(bar, 1);
	__syncwarp();
	if (lane == 0) {
		mbar_expect_tx(bar, bw * bh * 4);
		tma_load_2d(T, variant == 0 ? &pmap : gmap, x, y, bar);
	}
	mbar_wait(bar, 0);
	for (int i = lane; i < bw * bh; i += 32) out[i] = T[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
	const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
	int mode = argc > 1 ? atoi(argv[1]) : 0;
	const int w = 256, h = 200, bw = 32, bh = 29;
	std::vector<float> himg((size_t)w * h);
	for (int i = 0; i < w * h; i++) himg[i] = (float)(i % 1000);
	float *dimg, *dout;
	cudaMalloc(&dimg, himg.size() * 4);
	cudaMemcpy(dimg, himg.data(), himg.size() * 4, cudaMemcpyHostToDevice);
	cudaMalloc(&dout, bw * bh * 4);
	void* p = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
	printf("entry point: err=%d q=%d p=%p\n", (int)e, (int)q, p);
	CUtensorMap map;
	memset(&map, 0, sizeof(map));
	cuuint64_t dims[2] = { (cuuint64_t)w, (cuuint64_t)h };
	cuuint64_t strides[1] = { (cuuint64_t)w * 4 };
	cuuint32_t box[2] = { (cuuint32_t)bw, (cuuint32_t)bh };
	cuuint32_t estr[2] = { 1, 1 };
	CUresult r = ((EncodeTiledFn)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dimg, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
		CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("encode: %d\n", (int)r);
	CUtensorMap* gmap;
	cudaMalloc(&gmap, sizeof(map));
	cudaMemcpy(gmap, &map, sizeof(map), cudaMemcpyHostToDevice);
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
	std::vector<float> hout(bw * bh);
	if (mode == 2) {
		probe_bulk1d<<<1, 32>>>(dimg, dout, 512);
		e = cudaDeviceSynchronize();
		cudaMemcpy(hout.data(), dout, 512 * 4, cudaMemcpyDeviceToHost);
		printf("bulk1d: sync=%s out[7]=%g (expect 7) out[511]=%g\n", cudaGetErrorString(e), hout[7], hout[511]);
		return e != cudaSuccess;
	}
	for (int variant = mode; variant <= mode; variant++)
		for (int off = 0; off <= 4928; off += 4928) {
			cudaMemset(dout, 0, bw * bh * 4);
			probe<<<1, 32, 64 * 1024>>>(map, gmap, variant, bw, bh, 10, -2, dout, off);
			e = cudaDeviceSynchronize();
			cudaMemcpy(hout.data(), dout, bw * bh * 4, cudaMemcpyDeviceToHost);
			// expected: rows y=-2,-1 are zero, row 2 (y=0) starts at img[0][10]
			printf("variant %d off %d: sync=%s  out[0]=%g out[2*bw]=%g (expect %g) out[3*bw+5]=%g (expect %g)\n", variant, off,
				cudaGetErrorString(e), hout[0], hout[2 * bw], himg[10], hout[3 * bw + 5], himg[1 * w + 15]);
			if (e != cudaSuccess) return 1;
		}
	return 0;
}
