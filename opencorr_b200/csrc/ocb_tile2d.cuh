// ocb_tile2d.cuh -- per-warp 2D tile staging and bicubic B-spline sampling shared by icgn2d.cu and nr2d.cu.
#pragma once
#include "ocb_common.cuh"

namespace ocb {

// Stage a (rows x cols) window of a row-major image into smem (row pitch `cols`), origin (ox, oy),
// subtracting `shift`; pixels outside the image read as -shift.  Lanes run along x (coalesced).
__device__ __forceinline__ void stage_tile(float* dst, const float* __restrict__ img, int w, int h, int ox, int oy, int cols, int rows,
	float shift, int lane) {
	for (int col = lane; col < cols; col += 32) {
		const int gx = ox + col;
		const bool colok = gx >= 0 && gx < w;
#pragma unroll 8
		for (int row = 0; row < rows; row++) {
			const int gy = oy + row;
			float v = 0.f;
			if (colok && gy >= 0 && gy < h) v = __ldg(img + (size_t)gy * w + gx);
			dst[row * cols + col] = v - shift;
		}
	}
}

// Bicubic B-spline sample of the target at (X, Y), src/oc_cubic_bspline.cpp:134-181.
// fast: the 4x4 support lies inside the staged tile.  Otherwise read the image (caller guarantees
// 1 <= X < w-2, 1 <= Y < h-2).
__device__ __forceinline__ float bicubic_sample(const float* tile, int TW, int tx0, int ty0, const float* __restrict__ tar, int w,
	float X, float Y, bool fast) {
	const float xf = floorf(X), yf = floorf(Y);
	float wx[4], wy[4];
	bicubic_weights(X - xf, wx);
	bicubic_weights(Y - yf, wy);
	const int ix = (int)xf - 1, iy = (int)yf - 1;
	float t = 0.f;
	if (fast) {
		const float* q = tile + (iy - ty0) * TW + (ix - tx0);
#pragma unroll
		for (int nn = 0; nn < 4; nn++) {
			float row = fmaf(q[nn * TW + 3], wx[3], fmaf(q[nn * TW + 2], wx[2], fmaf(q[nn * TW + 1], wx[1], q[nn * TW] * wx[0])));
			t = fmaf(row, wy[nn], t);
		}
	} else {
		const float* q = tar + (size_t)iy * w + ix;
#pragma unroll
		for (int nn = 0; nn < 4; nn++) {
			const float* qq = q + (size_t)nn * w;
			float row = fmaf(__ldg(qq + 3), wx[3], fmaf(__ldg(qq + 2), wx[2], fmaf(__ldg(qq + 1), wx[1], __ldg(qq) * wx[0])));
			t = fmaf(row, wy[nn], t);
		}
	}
	return t;
}

} // namespace ocb
