#include <cstdio>
#include "/root/repo/opencorr_b200/csrc/ocb_tmem.cuh"
__global__ void k(float* out) {
	__shared__ uint32_t base_s;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (warp == 0) ocb::tmem_alloc<128>(&base_s);
	ocb::tmem_fence_before_sync();
	__syncthreads();
	ocb::tmem_fence_after_sync();
	const uint32_t tb = ocb::tmem_warp_base(base_s, warp);
	for (int c = 0; c < 128; c += 4) ocb::tmem_st4(tb + c, threadIdx.x * 1000.f + c, threadIdx.x * 1000.f + c + 1, threadIdx.x * 1000.f + c + 2, threadIdx.x * 1000.f + c + 3);
	ocb::tmem_wait_st();
	float s = 0;
	for (int c = 0; c < 128; c += 2) { float a, b; ocb::tmem_ld2(tb + c, a, b); ocb::tmem_wait_ld(); s += a - b; }
	float v = ocb::tmem_ld1(tb + 77); ocb::tmem_wait_ld();
	out[threadIdx.x] = v + s;
	__syncthreads();
	if (warp == 0) ocb::tmem_dealloc<128>(base_s);
}
int main() { float* d; cudaMalloc(&d, 512); k<<<1, 128>>>(d); float h[128]; cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost); printf("%s %f %f %f\n", cudaGetErrorString(cudaGetLastError()), h[0], h[1], h[127]); return 0; }
