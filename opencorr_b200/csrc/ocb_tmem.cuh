// ocb_tmem.cuh -- Tensor Memory (TMEM, 256 KB per SM on sm_100a: 128 lanes x 512 columns x 32 bit) used as a per-lane
// scratchpad.  TMEM is reachable only through tcgen05.ld / tcgen05.st (SASS LDTM / STTM); with the 32x32b shape every thread
// of a warp reads or writes N consecutive columns of ITS OWN lane -- lane 32 * (warp % 4) + laneid, a warp can only touch its
// quarter of the 128 lanes.  That is exactly the access pattern of the IC-GN kernels' per-sample constants (a lane walks down
// its subset column), so they can live here instead of in shared memory, which is what limits the resident warps per SM.
// All instructions are .sync.aligned: every thread of the warp must execute them, convergently.
#pragma once
#include <stdint.h>

namespace ocb {

// One warp of the CTA allocates `COLS` columns (power of two >= 32) for the whole CTA and publishes the base address in smem.
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((uint64_t)__cvta_generic_to_shared(smem_result)), "n"(COLS) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// address = base + (first lane of the warp's quarter << 16) + column
__device__ __forceinline__ uint32_t tmem_warp_base(uint32_t base, int warp_in_cta) { return base + ((uint32_t)((warp_in_cta & 3) * 32) << 16); }

__device__ __forceinline__ void tmem_st1(uint32_t addr, float a) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr), "r"(__float_as_uint(a)) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t addr, float a, float b) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t addr, float a, float b, float c, float d) {
	asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
		"r"(__float_as_uint(c)), "r"(__float_as_uint(d))
		: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tmem_ld1(uint32_t addr) {
	uint32_t a;
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(addr) : "memory");
	return __uint_as_float(a);
}
__device__ __forceinline__ void tmem_ld2(uint32_t addr, float& a, float& b) {
	uint32_t x, y;
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr) : "memory");
	a = __uint_as_float(x);
	b = __uint_as_float(y);
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float& a, float& b, float& c, float& d) {
	uint32_t x, y, z, w;
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
	a = __uint_as_float(x);
	b = __uint_as_float(y);
	c = __uint_as_float(z);
	d = __uint_as_float(w);
}
// the registers written by tcgen05.ld may only be read after this
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

} // namespace ocb
