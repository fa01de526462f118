// icgn2d.cu -- inverse-compositional Gauss-Newton subset registration, 2D, first-order (6
// parameters) and second-order (12 parameters) shape functions, for sm_100a.
//
// Replaces ICGN2D1::compute(POI2D*) (reference src/oc_icgn.cpp:144-341) and
// ICGN2D2::compute(POI2D*) (src/oc_icgn.cpp:685-898), including what ICGN2D*::prepare() feeds
// them (Gradient2D4, src/oc_gradient.cpp:37-79; BicubicBspline, src/oc_cubic_bspline.cpp:84-181).
//
// Mapping: ONE WARP PER POI, no block-level synchronisation.
//   setup   : the (2r+1)^2 reference subset is read once; zero-mean values f, gradients gx, gy
//             (4th-order differences recomputed from the image, nothing precomputed in HBM) go to
//             the warp's shared-memory slab; the Hessian, sum(sd) and sum(sd*f) are accumulated in
//             registers and reduced with warp shuffles; the Hessian is Cholesky-factorised in
//             registers (same arithmetic replicated in every lane).
//   iterate : a target tile (subset + bicubic support + slack) is staged in shared memory once per
//             POI; each iteration evaluates the bicubic interpolant from the 4x4 pixel block
//             with explicit fp32 weights (the reference's 64 B/pixel LUT is never materialised),
//             and accumulates ONE pass of sums: with d = (t - mean_ref) - f,
//                 sum d, sum d^2, sum f*d, sum sd_k*d
//             from which mean/norm of the warped target, ZNSSD and the Gauss-Newton right-hand
//             side follow algebraically (see DESIGN.md "single-pass IC-GN sums").
//   update  : solve with the Cholesky factors, compose W <- W * W(dp)^-1 in registers.
// Samples whose 4x4 support leaves the staged tile (large deformation gradients) fall back to
// global-memory reads of the target image, so results never depend on the tile size.
#include "ocb_kernels.h"

namespace ocb {

constexpr int ICGN2D_TILE_MARGIN = 2; // slack (pixels) around subset+support in the target tile

__host__ __device__ inline int icgn2d_tile_w(int rx) { return 2 * rx + 4 + 2 * ICGN2D_TILE_MARGIN; }
__host__ __device__ inline int icgn2d_warp_floats(int rx, int ry) {
	int n = (2 * rx + 1) * (2 * ry + 1);
	int t = icgn2d_tile_w(rx) * icgn2d_tile_w(ry);
	return ((3 * n + t) + 3) & ~3;
}

// W(p) of the second-order shape function (reference src/oc_deformation.cpp:301-350)
__device__ __forceinline__ void warp2d2_matrix(const float* p, float* W) {
	const float u = p[0], ux = p[1], uy = p[2], uxx = p[3], uxy = p[4], uyy = p[5];
	const float v = p[6], vx = p[7], vy = p[8], vxx = p[9], vxy = p[10], vyy = p[11];
	W[0] = 1.f + 2.f * ux + ux * ux + u * uxx;
	W[1] = 2.f * u * uxy + 2.f * (1.f + ux) * uy;
	W[2] = uy * uy + u * uyy;
	W[3] = 2.f * u * (1.f + ux);
	W[4] = 2.f * u * uy;
	W[5] = u * u;
	W[6] = 0.5f * (v * uxx + 2.f * (1.f + ux) * vx + u * vxx);
	W[7] = 1.f + uy * vx + ux * vy + v * uxy + u * vxy + vy + ux;
	W[8] = 0.5f * (v * uyy + 2.f * uy * (1.f + vy) + u * vyy);
	W[9] = v + v * ux + u * vx;
	W[10] = u + v * uy + u * vy;
	W[11] = u * v;
	W[12] = vx * vx + v * vxx;
	W[13] = 2.f * v * vxy + 2.f * vx * (1.f + vy);
	W[14] = 1.f + 2.f * vy + vy * vy + v * vyy;
	W[15] = 2.f * v * vx;
	W[16] = 2.f * v * (1.f + vy);
	W[17] = v * v;
	W[18] = 0.5f * uxx; W[19] = uxy; W[20] = 0.5f * uyy; W[21] = 1.f + ux; W[22] = uy; W[23] = u;
	W[24] = 0.5f * vxx; W[25] = vxy; W[26] = 0.5f * vyy; W[27] = vx; W[28] = 1.f + vy; W[29] = v;
	// row 5 = [0 0 0 0 0 1] is implicit
}

// rows <- rows * M^-1 for the 2x6 block `rows` (rows 3,4 of the running warp) and the 6x6 warp
// increment M whose last row is [0 0 0 0 0 1] (given as its first 5 rows, 30 floats).
// Gaussian elimination without pivoting: M = W(dp) is a perturbation of the identity.
__device__ __forceinline__ void right_divide_2x6(float* rows, float* M) {
	// Solve X M = R  <=>  for each row x of X: x M = r.  Eliminate column by column:
	// x_j = (r_j - sum_{i<j} x_i M[i][j]) / M[j][j] requires M upper triangular, so first reduce M
	// to upper-triangular form U = L^-1 M with row operations, accumulating L: X M = R  <=>
	// (X L) U = R.  Let Y = X L: solve Y U = R by forward substitution over columns, then
	// X = Y L^-1, applied by undoing the row operations in reverse order.
	float Lm[5][5]; // multipliers l[i][k], i > k (rows 0..4; row 5 of M is e5 and needs no elimination)
#pragma unroll
	for (int k = 0; k < 5; k++) {
		float inv = 1.0f / M[k * 6 + k];
#pragma unroll
		for (int i = k + 1; i < 5; i++) {
			float l = M[i * 6 + k] * inv;
			Lm[i][k] = l;
#pragma unroll
			for (int j = k + 1; j < 6; j++) M[i * 6 + j] -= l * M[k * 6 + j];
		}
	}
#pragma unroll
	for (int r = 0; r < 2; r++) {
		float y[6];
		// Y U = R, U upper triangular 6x6 (row 5 = e5)
#pragma unroll
		for (int j = 0; j < 6; j++) {
			float v = rows[r * 6 + j];
#pragma unroll
			for (int i = 0; i < j; i++) {
				if (i < 5) v -= y[i] * M[i * 6 + j];
			}
			y[j] = (j < 5) ? v / M[j * 6 + j] : v;
		}
		// X = Y L^-1 : x_k = y_k - sum_{i>k} x_i l[i][k], from the last column backwards (i < 5 only)
#pragma unroll
		for (int k = 4; k >= 0; k--) {
			float v = y[k];
#pragma unroll
			for (int i = k + 1; i < 5; i++) v -= y[i] * Lm[i][k];
			y[k] = v;
		}
#pragma unroll
		for (int j = 0; j < 6; j++) rows[r * 6 + j] = y[j];
	}
}

template <int NP>
__global__ void __launch_bounds__(128) icgn2d_kernel(Image2D img, float* __restrict__ pois, int n_poi, int rx, int ry,
	float conv_criterion, float stop_condition, int warps_per_block) {
	extern __shared__ __align__(16) float smem[];
	constexpr int NH = NP * (NP + 1) / 2;
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int sw = 2 * rx + 1, sh = 2 * ry + 1, N = sw * sh;
	const int TW = icgn2d_tile_w(rx), TH = icgn2d_tile_w(ry);
	float* sF = smem + (size_t)warp * icgn2d_warp_floats(rx, ry);
	float* sGx = sF + N;
	float* sGy = sGx + N;
	float* tile = sGy + N;
	const float* __restrict__ ref = img.ref;
	const float* __restrict__ tar = img.tar;
	const int w = img.w, h = img.h;
	const float inv_n = 1.0f / (float)N;

	for (int poi = blockIdx.x * warps_per_block + warp; poi < n_poi; poi += gridDim.x * warps_per_block) {
		float* P = pois + (size_t)poi * P2_N;
		const float rec = lane < P2_N ? P[lane] : 0.f;
		const float px = __shfl_sync(0xffffffffu, rec, P2_X);
		const float py = __shfl_sync(0xffffffffu, rec, P2_Y);
		const float u_in = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_U);
		const float v_in = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_V);
		const float zncc_in = __shfl_sync(0xffffffffu, rec, P2_ZNCC);
		// guard, reference src/oc_icgn.cpp:160-167 / :701-708 (NaN coordinates are rejected too)
		if (py - ry < 0 || px - rx < 0 || py + ry > h - 1 || px + rx > w - 1 || fabsf(u_in) >= w || fabsf(v_in) >= h
			|| zncc_in < 0 || is_nan_f(u_in) || is_nan_f(v_in) || is_nan_f(px) || is_nan_f(py)) {
			if (lane == 0) P[P2_ZNCC] = zncc_in >= 0 ? -3.f : zncc_in;
			continue;
		}
		__syncwarp();

		// ---------------- setup: reference subset, gradients, Hessian ----------------
		const int x0 = (int)px - rx, y0 = (int)py - ry;
		float s1 = 0.f;
		for (int i = lane; i < N; i += 32) {
			int r = i / sw, c = i - r * sw;
			float R = __ldg(ref + (size_t)(y0 + r) * w + (x0 + c));
			sF[i] = R;
			s1 += R;
		}
		const float ref_mean = warp_sum(s1) * inv_n; // Subset2D::zeroMeanNorm, src/oc_subset.cpp:46-53
		float H[NH], S[NP], SF[NP];
#pragma unroll
		for (int k = 0; k < NH; k++) H[k] = 0.f;
#pragma unroll
		for (int k = 0; k < NP; k++) { S[k] = 0.f; SF[k] = 0.f; }
		float f2 = 0.f;
		for (int i = lane; i < N; i += 32) {
			int r = i / sw, c = i - r * sw;
			int xg = x0 + c, yg = y0 + r;
			float f = sF[i] - ref_mean;
			sF[i] = f;
			f2 = fmaf(f, f, f2);
			const float* q = ref + (size_t)yg * w + xg;
			float gx = 0.f, gy = 0.f; // 2-pixel borders of the gradient maps are zero (src/oc_gradient.cpp:42,46)
			if (xg >= 2 && xg < w - 2) gx = grad4(__ldg(q - 2), __ldg(q - 1), __ldg(q + 1), __ldg(q + 2));
			if (yg >= 2 && yg < h - 2) gy = grad4(__ldg(q - 2 * (size_t)w), __ldg(q - (size_t)w), __ldg(q + (size_t)w), __ldg(q + 2 * (size_t)w));
			sGx[i] = gx;
			sGy[i] = gy;
			float xl = (float)(c - rx), yl = (float)(r - ry);
			float sd[NP];
			if constexpr (NP == 6) { // src/oc_icgn.cpp:191-196
				sd[0] = gx; sd[1] = gx * xl; sd[2] = gx * yl;
				sd[3] = gy; sd[4] = gy * xl; sd[5] = gy * yl;
			} else { // src/oc_icgn.cpp:725-745
				float xx = xl * xl * 0.5f, xy = xl * yl, yy = yl * yl * 0.5f;
				sd[0] = gx; sd[1] = gx * xl; sd[2] = gx * yl; sd[3] = gx * xx; sd[4] = gx * xy; sd[5] = gx * yy;
				sd[6] = gy; sd[7] = gy * xl; sd[8] = gy * yl; sd[9] = gy * xx; sd[10] = gy * xy; sd[11] = gy * yy;
			}
#pragma unroll
			for (int a = 0; a < NP; a++) {
				S[a] += sd[a];
				SF[a] = fmaf(sd[a], f, SF[a]);
#pragma unroll
				for (int b = 0; b <= a; b++) H[a * (a + 1) / 2 + b] = fmaf(sd[a], sd[b], H[a * (a + 1) / 2 + b]);
			}
		}
		f2 = warp_sum(f2);
#pragma unroll
		for (int k = 0; k < NH; k++) H[k] = warp_sum(H[k]);
#pragma unroll
		for (int k = 0; k < NP; k++) { S[k] = warp_sum(S[k]); SF[k] = warp_sum(SF[k]); }
		cholesky_packed<NP>(H);
		const float ref_norm = sqrtf(f2);

		// ---------------- target tile ----------------
		const int tx0 = (int)floorf(px + u_in) - rx - 1 - ICGN2D_TILE_MARGIN;
		const int ty0 = (int)floorf(py + v_in) - ry - 1 - ICGN2D_TILE_MARGIN;
		for (int i = lane; i < TW * TH; i += 32) {
			int ty = i / TW, tx = i - ty * TW;
			int gxp = tx0 + tx, gyp = ty0 + ty;
			float val = 0.f;
			if (gxp >= 0 && gxp < w && gyp >= 0 && gyp < h) val = __ldg(tar + (size_t)gyp * w + gxp);
			tile[i] = val;
		}
		__syncwarp();

		// ---------------- IC-GN iterations ----------------
		// running warp: NP==6 -> A = {W00,W01,W02,W10,W11,W12}; NP==12 -> rows 3,4 of the 6x6 warp
		float A[12];
		{
			const float ux = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_UX), uy = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_UY);
			const float vx = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_VX), vy = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_VY);
			if constexpr (NP == 6) {
				A[0] = 1.f + ux; A[1] = uy; A[2] = u_in; A[3] = vx; A[4] = 1.f + vy; A[5] = v_in;
			} else { // second-order terms of the incoming guess are dropped (src/oc_icgn.cpp:765-770)
				A[0] = 0.f; A[1] = 0.f; A[2] = 0.f; A[3] = 1.f + ux; A[4] = uy; A[5] = u_in;
				A[6] = 0.f; A[7] = 0.f; A[8] = 0.f; A[9] = vx; A[10] = 1.f + vy; A[11] = v_in;
			}
		}
		const float xmax = (float)(w - 2), ymax = (float)(h - 2);
		int iteration = 0;
		float dp_norm = 0.f, zncc = 0.f;
		bool left_image = false;
		float dp[NP];
		do {
			iteration++;
			float d1 = 0.f, d2 = 0.f, fd = 0.f;
			float SD[NP];
#pragma unroll
			for (int k = 0; k < NP; k++) SD[k] = 0.f;
			bool invalid = false;
			int r = 0, c = lane;
			while (c >= sw) { c -= sw; r++; }
			for (int i = lane; i < N; i += 32) {
				const float xl = (float)(c - rx), yl = (float)(r - ry);
				float wxp, wyp;
				if constexpr (NP == 6) { // Deformation2D1::warp, src/oc_deformation.cpp:94-105
					wxp = fmaf(A[0], xl, fmaf(A[1], yl, A[2]));
					wyp = fmaf(A[3], xl, fmaf(A[4], yl, A[5]));
				} else { // Deformation2D2::warp rows 3,4, src/oc_deformation.cpp:268-282
					const float m0 = xl * xl, m1 = xl * yl, m2 = yl * yl;
					wxp = fmaf(A[0], m0, fmaf(A[1], m1, fmaf(A[2], m2, fmaf(A[3], xl, fmaf(A[4], yl, A[5])))));
					wyp = fmaf(A[6], m0, fmaf(A[7], m1, fmaf(A[8], m2, fmaf(A[9], xl, fmaf(A[10], yl, A[11])))));
				}
				const float X = px + wxp, Y = py + wyp;
				// BicubicBspline::compute validity, src/oc_cubic_bspline.cpp:137-142 (NaN fails the test too)
				const bool ok = (X >= 1.f) && (Y >= 1.f) && (X < xmax) && (Y < ymax);
				if (!ok) {
					invalid = true;
				} else {
					const float xf = floorf(X), yf = floorf(Y);
					float wx[4], wy[4];
					bicubic_weights(X - xf, wx);
					bicubic_weights(Y - yf, wy);
					const int ix = (int)xf - 1, iy = (int)yf - 1;
					const int lx = ix - tx0, ly = iy - ty0;
					float t = 0.f;
					if (lx >= 0 && ly >= 0 && lx + 3 < TW && ly + 3 < TH) {
						const float* q = tile + ly * TW + lx;
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
					const float f = sF[i];
					const float d = (t - ref_mean) - f;
					d1 += d;
					d2 = fmaf(d, d, d2);
					fd = fmaf(f, d, fd);
					const float gxd = sGx[i] * d, gyd = sGy[i] * d;
					if constexpr (NP == 6) {
						SD[0] += gxd; SD[1] = fmaf(gxd, xl, SD[1]); SD[2] = fmaf(gxd, yl, SD[2]);
						SD[3] += gyd; SD[4] = fmaf(gyd, xl, SD[4]); SD[5] = fmaf(gyd, yl, SD[5]);
					} else {
						const float xx = xl * xl * 0.5f, xy = xl * yl, yy = yl * yl * 0.5f;
						SD[0] += gxd; SD[1] = fmaf(gxd, xl, SD[1]); SD[2] = fmaf(gxd, yl, SD[2]);
						SD[3] = fmaf(gxd, xx, SD[3]); SD[4] = fmaf(gxd, xy, SD[4]); SD[5] = fmaf(gxd, yy, SD[5]);
						SD[6] += gyd; SD[7] = fmaf(gyd, xl, SD[7]); SD[8] = fmaf(gyd, yl, SD[8]);
						SD[9] = fmaf(gyd, xx, SD[9]); SD[10] = fmaf(gyd, xy, SD[10]); SD[11] = fmaf(gyd, yy, SD[11]);
					}
				}
				c += 32;
				while (c >= sw) { c -= sw; r++; }
			}
			if (__any_sync(0xffffffffu, invalid)) { // src/oc_icgn.cpp:251-255
				left_image = true;
				break;
			}
			d1 = warp_sum(d1);
			d2 = warp_sum(d2);
			fd = warp_sum(fd);
#pragma unroll
			for (int k = 0; k < NP; k++) SD[k] = warp_sum(SD[k]);
			// warped-target statistics: g = t - mean(t) = f + (d - dbar)
			const float dbar = d1 * inv_n;
			const float g2 = f2 + 2.f * fd + (d2 - d1 * dbar);
			const float tar_norm = sqrtf(g2);
			const float factor = ref_norm / tar_norm; // src/oc_icgn.cpp:260
			zncc = (f2 + fd) / (ref_norm * tar_norm); // == 0.5*(2 - znssd), src/oc_icgn.cpp:263,320
			float b[NP];
#pragma unroll
			for (int k = 0; k < NP; k++) b[k] = factor * (SF[k] + SD[k] - dbar * S[k]) - SF[k];
			cholesky_solve<NP>(H, b, dp);
			if constexpr (NP == 6) {
				// W <- W * W(dp)^-1, 3x3 affine (src/oc_icgn.cpp:290)
				const float a = dp[1], bb = dp[2], cc = dp[0], d = dp[4], e = dp[5], ff = dp[3];
				const float det = (1.f + a) * (1.f + e) - bb * d;
				const float id = 1.0f / det;
				const float i00 = (1.f + e) * id, i01 = -bb * id, i02 = (bb * ff - cc * (1.f + e)) * id;
				const float i10 = -d * id, i11 = (1.f + a) * id, i12 = (cc * d - (1.f + a) * ff) * id;
				const float n00 = A[0] * i00 + A[1] * i10, n01 = A[0] * i01 + A[1] * i11, n02 = A[0] * i02 + A[1] * i12 + A[2];
				const float n10 = A[3] * i00 + A[4] * i10, n11 = A[3] * i01 + A[4] * i11, n12 = A[3] * i02 + A[4] * i12 + A[5];
				A[0] = n00; A[1] = n01; A[2] = n02; A[3] = n10; A[4] = n11; A[5] = n12;
				const float rx2 = (float)(rx * rx), ry2 = (float)(ry * ry);
				dp_norm = dp[0] * dp[0] + dp[1] * dp[1] * rx2 + dp[2] * dp[2] * ry2
					+ dp[3] * dp[3] + dp[4] * dp[4] * rx2 + dp[5] * dp[5] * ry2; // src/oc_icgn.cpp:296-306
			} else {
				float M[30];
				warp2d2_matrix(dp, M);
				right_divide_2x6(A, M); // rows 3,4 of W * W(dp)^-1 (src/oc_icgn.cpp:831)
				const int rx2 = rx * rx, ry2 = ry * ry;
				const float rxy2 = (float)(rx2 * ry2);
				const float rx4 = (float)(int)((float)(rx2 * rx2) * 0.25f); // float->int truncation, src/oc_icgn.cpp:840-841
				const float ry4 = (float)(int)((float)(ry2 * ry2) * 0.25f);
				dp_norm = dp[0] * dp[0] + dp[1] * dp[1] * (float)rx2 + dp[2] * dp[2] * (float)ry2
					+ dp[3] * dp[3] * rx4 + dp[5] * dp[5] * ry4 + dp[4] * dp[4] * rxy2
					+ dp[6] * dp[6] + dp[7] * dp[7] * (float)rx2 + dp[8] * dp[8] * (float)ry2
					+ dp[9] * dp[9] * rx4 + dp[11] * dp[11] * ry4 + dp[10] * dp[10] * rxy2;
			}
			dp_norm = sqrtf(dp_norm);
		} while ((float)iteration < stop_condition && dp_norm >= conv_criterion);

		if (left_image) {
			if (lane == 0) P[P2_ZNCC] = -3.f;
			__syncwarp();
			continue;
		}
		// ---------------- results, src/oc_icgn.cpp:310-340 / :859-897 ----------------
		if (lane == 0) {
			float u, v;
			if constexpr (NP == 6) {
				u = A[2]; v = A[5];
				P[P2_DEF + D2_U] = u; P[P2_DEF + D2_UX] = A[0] - 1.f; P[P2_DEF + D2_UY] = A[1];
				P[P2_DEF + D2_V] = v; P[P2_DEF + D2_VX] = A[3]; P[P2_DEF + D2_VY] = A[4] - 1.f;
			} else { // Deformation2D2::setDeformation(), src/oc_deformation.cpp:284-299
				u = A[5]; v = A[11];
				P[P2_DEF + D2_U] = u; P[P2_DEF + D2_UX] = A[3] - 1.f; P[P2_DEF + D2_UY] = A[4];
				P[P2_DEF + D2_UXX] = A[0] * 2.f; P[P2_DEF + D2_UXY] = A[1]; P[P2_DEF + D2_UYY] = A[2] * 2.f;
				P[P2_DEF + D2_V] = v; P[P2_DEF + D2_VX] = A[9]; P[P2_DEF + D2_VY] = A[10] - 1.f;
				P[P2_DEF + D2_VXX] = A[6] * 2.f; P[P2_DEF + D2_VXY] = A[7]; P[P2_DEF + D2_VYY] = A[8] * 2.f;
			}
			P[P2_U0] = u_in;
			P[P2_V0] = v_in;
			float zout = zncc;
			P[P2_ITER] = (float)iteration;
			P[P2_CONV] = dp_norm;
			P[P2_RX] = (float)rx;
			P[P2_RY] = (float)ry;
			if (dp_norm >= conv_criterion && (float)iteration >= stop_condition) zout = -4.f;
			if (is_nan_f(zout) || is_nan_f(u) || is_nan_f(v)) {
				P[P2_DEF + D2_U] = u_in;
				P[P2_DEF + D2_V] = v_in;
				zout = -5.f;
			}
			P[P2_ZNCC] = zout;
		}
		__syncwarp();
	}
}

// host-side launch ---------------------------------------------------------------------------
// Returns 0, or -1 when one warp's slab does not fit in shared memory.
int icgn2d_launch(int np, const Image2D& img, float* d_pois, size_t n, int rx, int ry, float conv, float stop,
	int sm_count, size_t smem_optin, cudaStream_t stream, cudaError_t* err) {
	const size_t per_warp = (size_t)icgn2d_warp_floats(rx, ry) * sizeof(float);
	int best_wpb = 0, best_warps = 0;
	for (int wpb = 4; wpb >= 1; wpb >>= 1) {
		size_t need = per_warp * wpb;
		if (need > smem_optin) continue;
		int blocks = (int)((228 * 1024) / (need + 1024));
		if (blocks > 32) blocks = 32;
		int warps = blocks * wpb;
		if (warps > best_warps) { best_warps = warps; best_wpb = wpb; }
	}
	if (best_wpb == 0) return -1;
	const size_t smem = per_warp * best_wpb;
	auto kern = (np == 6) ? icgn2d_kernel<6> : icgn2d_kernel<12>;
	*err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (*err != cudaSuccess) return -2;
	long long blocks_needed = ((long long)n + best_wpb - 1) / best_wpb;
	long long resident = (long long)sm_count * (best_warps / best_wpb);
	int grid = (int)(blocks_needed < resident * 4 ? blocks_needed : resident * 4);
	if (grid < 1) grid = 1;
	kern<<<grid, best_wpb * 32, smem, stream>>>(img, d_pois, (int)n, rx, ry, conv, stop, best_wpb);
	*err = cudaGetLastError();
	return *err == cudaSuccess ? 0 : -2;
}

} // namespace ocb
