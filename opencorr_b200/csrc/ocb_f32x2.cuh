// ocb_f32x2.cuh -- packed fp32 pairs on sm_100a: fma/mul/add.rn.f32x2 (SASS FFMA2 / FMUL2 / FADD2).
// One instruction issues two IEEE round-to-nearest fp32 operations (bit-identical to the scalar ones), so
// arithmetic-heavy inner loops spend about half the issue slots on math; a scalar operand broadcast to both
// halves is an operand modifier (R.F32), not an extra move.  Blackwell only: these do not exist on sm_90.
#pragma once
#include <cuda_runtime.h>

namespace ocb {

typedef unsigned long long ocb_u64;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
	float2 d;
	asm("fma.rn.f32x2 %0, %1, %2, %3;"
		: "=l"(reinterpret_cast<ocb_u64&>(d))
		: "l"(reinterpret_cast<ocb_u64&>(a)), "l"(reinterpret_cast<ocb_u64&>(b)), "l"(reinterpret_cast<ocb_u64&>(c)));
	return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
	float2 d;
	asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<ocb_u64&>(d)) : "l"(reinterpret_cast<ocb_u64&>(a)), "l"(reinterpret_cast<ocb_u64&>(b)));
	return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
	float2 d;
	asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<ocb_u64&>(d)) : "l"(reinterpret_cast<ocb_u64&>(a)), "l"(reinterpret_cast<ocb_u64&>(b)));
	return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
	float2 d;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<ocb_u64&>(d)) : "l"(reinterpret_cast<ocb_u64&>(a)), "l"(reinterpret_cast<ocb_u64&>(b)));
	return d;
}
__device__ __forceinline__ float2 bcast2(float a) { return make_float2(a, a); }

// Bicubic weights of the reference's BC matrix (ocb_common.cuh bicubic_weights) as two pairs {w0, w1}, {w2, w3};
// the same Horner steps, two weights per instruction.
__device__ __forceinline__ void bicubic_weights2(float t, float2& w01, float2& w23) {
	const float s = 1.0f / 336.0f;
	const float2 tt = bcast2(t);
	w01 = ffma2(ffma2(ffma2(make_float2(-144.0f * s, 384.0f * s), tt, make_float2(342.0f * s, -702.0f * s)), tt, make_float2(-198.0f * s, -18.0f * s)), tt,
		make_float2(0.f, 1.0f));
	w23 = fmul2(ffma2(ffma2(make_float2(-384.0f * s, 144.0f * s), tt, make_float2(450.0f * s, -90.0f * s)), tt, make_float2(270.0f * s, -54.0f * s)), tt);
}

} // namespace ocb
