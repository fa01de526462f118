// fftcc2d_w32.cu -- FFT-CC for the 32x32 window (subset radius 16, the headline configuration):
// ONE WARP PER POI, the 32-point transforms live entirely in registers.
//
// Same algorithm as fftcc2d_kernel (reference src/oc_fftcc.cpp:177-275), specialised:
//   lane = window column: each lane gathers its column of both windows (coalesced rows), the packed
//     z = ref + i*tar column is transformed along y with a fully unrolled radix-2 DIF FFT
//     (compile-time twiddles, natural order in -> bit-reversed order out, tracked statically);
//   transpose through a padded 32x33 shared tile; lane = ky: DIF FFT along x;
//   cross spectrum C = conj(A) B from Z(k) and Z(-k): the partner bin sits in lane (32-ky)%32 at a
//     statically known register, fetched with warp shuffles;
//   inverse: DIT FFT along kx (bit-reversed in -> natural out), transpose back, DIT along ky;
//   first-maximum argmax over registers + warp shuffle reduction.
// No CTA barrier, no integer division, ~2.6k warp instructions per POI (the generic kernel: ~23k).
#include "ocb_kernels.h"

namespace ocb {

__host__ __device__ constexpr int brev5(int i) {
	return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// (cos, sin) of 2*pi*k/32
__device__ __forceinline__ float tw32_cos(int k) {
	switch (k) {
	case 0: return 1.0f;
	case 1: return 0.98078528040323044913f;
	case 2: return 0.92387953251128675613f;
	case 3: return 0.83146961230254523708f;
	case 4: return 0.70710678118654752440f;
	case 5: return 0.55557023301960222474f;
	case 6: return 0.38268343236508977173f;
	case 7: return 0.19509032201612826785f;
	case 8: return 0.0f;
	case 9: return -0.19509032201612826785f;
	case 10: return -0.38268343236508977173f;
	case 11: return -0.55557023301960222474f;
	case 12: return -0.70710678118654752440f;
	case 13: return -0.83146961230254523708f;
	case 14: return -0.92387953251128675613f;
	default: return -0.98078528040323044913f;
	}
}
__device__ __forceinline__ float tw32_sin(int k) { return k < 8 ? tw32_cos(8 - k) : tw32_cos(k - 8); }

// (yr, yi) = (xr + i xi) * W, W = exp(-+ 2 pi i k / 32)  (minus: forward, plus: inverse)
template <bool INV>
__device__ __forceinline__ void mul_tw32(float xr, float xi, int k, float& yr, float& yi) {
	if (k == 0) {
		yr = xr;
		yi = xi;
	} else if (k == 8) { // -i (forward) / +i (inverse)
		yr = INV ? -xi : xi;
		yi = INV ? xr : -xr;
	} else {
		const float c = tw32_cos(k), s = INV ? -tw32_sin(k) : tw32_sin(k); // W = c - i s
		yr = fmaf(xr, c, xi * s);
		yi = fmaf(xi, c, -xr * s);
	}
}

// radix-2 decimation in frequency: natural order in, bit-reversed order out
template <bool INV>
__device__ __forceinline__ void fft32_dif(float* re, float* im) {
#pragma unroll
	for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
		for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
			for (int k = 0; k < half; k++) {
				const int i = base + k, j = i + half;
				const float ar = re[i], ai = im[i], br = re[j], bi = im[j];
				re[i] = ar + br;
				im[i] = ai + bi;
				mul_tw32<INV>(ar - br, ai - bi, k * (16 / half), re[j], im[j]);
			}
		}
	}
}

// radix-2 decimation in time: bit-reversed order in, natural order out
template <bool INV>
__device__ __forceinline__ void fft32_dit(float* re, float* im) {
#pragma unroll
	for (int half = 1; half <= 16; half <<= 1) {
#pragma unroll
		for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
			for (int k = 0; k < half; k++) {
				const int i = base + k, j = i + half;
				float tr, ti;
				mul_tw32<INV>(re[j], im[j], k * (16 / half), tr, ti);
				const float ar = re[i], ai = im[i];
				re[i] = ar + tr;
				im[i] = ai + ti;
				re[j] = ar - tr;
				im[j] = ai - ti;
			}
		}
	}
}

__device__ __forceinline__ void cross32(float zr, float zi, float nr, float ni, float& cr, float& ci) {
	// A = (z + conj n)/2, B = (z - conj n)/(2i), C = conj(A) B   (src/oc_fftcc.cpp:239-240)
	const float Ar = 0.5f * (zr + nr), Ai = 0.5f * (zi - ni);
	const float dr = 0.5f * (zr - nr), di = 0.5f * (zi + ni);
	const float Br = di, Bi = -dr;
	cr = Ar * Br + Ai * Bi;
	ci = Ar * Bi - Ai * Br;
}

constexpr int FFTW32_WARPS = 4;
constexpr int FFTW32_PITCH = 33;

__global__ void __launch_bounds__(FFTW32_WARPS * 32) fftcc2d_w32_kernel(Image2D img, float* __restrict__ pois, int n_poi) {
	__shared__ float s_re[FFTW32_WARPS][32 * FFTW32_PITCH];
	__shared__ float s_im[FFTW32_WARPS][32 * FFTW32_PITCH];
	constexpr int R = 16, NW = 32, M = NW * NW;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float* sre = s_re[warp];
	float* sim = s_im[warp];
	const int w = img.w, h = img.h;
	const int plane = (32 - lane) & 31;

	for (int poi = blockIdx.x * FFTW32_WARPS + warp; poi < n_poi; poi += gridDim.x * FFTW32_WARPS) {
		float* P = pois + (size_t)poi * P2_N;
		const float rec = lane < P2_N ? P[lane] : 0.f;
		const float px = __shfl_sync(0xffffffffu, rec, P2_X), py = __shfl_sync(0xffffffffu, rec, P2_Y);
		const float u0 = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_U), v0 = __shfl_sync(0xffffffffu, rec, P2_DEF + D2_V);
		// border guard: the POI is left untouched (src/oc_fftcc.cpp:190-196)
		if ((int)px < R || (int)px >= w - R || (int)py < R || (int)py >= h - R || (int)(px + u0) < R || (int)(px + u0) >= w - R
			|| (int)(py + v0) < R || (int)(py + v0) >= h - R || is_nan_f(px) || is_nan_f(py) || is_nan_f(u0) || is_nan_f(v0))
			continue;
		// gather: lane = column; float coordinate arithmetic then (int) truncation (src/oc_fftcc.cpp:204-219)
		float re[32], im[32];
		{
			const float rpx = px + lane - R;
			const int ax = (int)rpx, bx = (int)(rpx + u0);
			float sa = 0.f, sb = 0.f;
#pragma unroll
			for (int r = 0; r < 32; r++) {
				const float rpy = py + r - R;
				re[r] = __ldg(img.ref + (size_t)(int)rpy * w + ax);
				im[r] = __ldg(img.tar + (size_t)(int)(rpy + v0) * w + bx);
				sa += re[r];
				sb += im[r];
			}
			sa = warp_sum(sa) / (float)M;
			sb = warp_sum(sb) / (float)M;
#pragma unroll
			for (int r = 0; r < 32; r++) {
				re[r] -= sa;
				im[r] -= sb;
			}
		}
		float na = 0.f, nb = 0.f;
#pragma unroll
		for (int r = 0; r < 32; r++) {
			na = fmaf(re[r], re[r], na);
			nb = fmaf(im[r], im[r], nb);
		}
		na = warp_sum(na);
		nb = warp_sum(nb);

		fft32_dif<false>(re, im); // along y: register i holds ky = brev5(i)
		__syncwarp();
#pragma unroll
		for (int i = 0; i < 32; i++) {
			sre[brev5(i) * FFTW32_PITCH + lane] = re[i];
			sim[brev5(i) * FFTW32_PITCH + lane] = im[i];
		}
		__syncwarp();
#pragma unroll
		for (int i = 0; i < 32; i++) { // lane = ky, register i = x
			re[i] = sre[lane * FFTW32_PITCH + i];
			im[i] = sim[lane * FFTW32_PITCH + i];
		}
		fft32_dif<false>(re, im); // along x: register i holds kx = brev5(i)

		// cross spectrum; partner of (ky, kx) is (-ky, -kx): lane `plane`, register brev5((32 - kx) % 32)
#pragma unroll
		for (int i = 0; i < 32; i++) {
			const int ip = brev5((32 - brev5(i)) & 31);
			if (ip < i) continue;
			const float pr_i = __shfl_sync(0xffffffffu, re[i], plane), pi_i = __shfl_sync(0xffffffffu, im[i], plane);
			if (ip == i) {
				float cr, ci;
				cross32(re[i], im[i], pr_i, pi_i, cr, ci);
				re[i] = cr;
				im[i] = ci;
			} else {
				const float pr_p = __shfl_sync(0xffffffffu, re[ip], plane), pi_p = __shfl_sync(0xffffffffu, im[ip], plane);
				float c0r, c0i, c1r, c1i;
				cross32(re[i], im[i], pr_p, pi_p, c0r, c0i);
				cross32(re[ip], im[ip], pr_i, pi_i, c1r, c1i);
				re[i] = c0r; im[i] = c0i;
				re[ip] = c1r; im[ip] = c1i;
			}
		}

		fft32_dit<true>(re, im); // inverse along kx: register i = x
		__syncwarp();
#pragma unroll
		for (int i = 0; i < 32; i++) {
			sre[lane * FFTW32_PITCH + i] = re[i];
			sim[lane * FFTW32_PITCH + i] = im[i];
		}
		__syncwarp();
#pragma unroll
		for (int i = 0; i < 32; i++) { // lane = x, register i holds ky = brev5(i)
			re[i] = sre[brev5(i) * FFTW32_PITCH + lane];
			im[i] = sim[brev5(i) * FFTW32_PITCH + lane];
		}
		fft32_dit<true>(re, im); // inverse along ky: register i = y, lane = x

		// first maximum in linear order y*32 + x (src/oc_fftcc.cpp:246-255)
		float bv = -2.f;
		int bi = 0;
#pragma unroll
		for (int y = 0; y < 32; y++) {
			if (re[y] > bv) { bv = re[y]; bi = y * 32 + lane; }
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
			const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
			if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
		}
		if (lane == 0) {
			int du = bi & 31, dv = bi >> 5;
			if (du > R) du -= NW;
			if (dv > R) dv -= NW;
			P[P2_DEF + D2_U] = (float)du + u0;
			P[P2_DEF + D2_V] = (float)dv + v0;
			P[P2_U0] = u0;
			P[P2_V0] = v0;
			P[P2_ZNCC] = bv / (sqrtf(na * nb) * (float)M); // src/oc_fftcc.cpp:274
		}
		__syncwarp();
	}
}

int fftcc2d_w32_launch(const Image2D& img, float* d_pois, size_t n, int sm_count, cudaStream_t stream, cudaError_t* err) {
	long long blocks_needed = ((long long)n + FFTW32_WARPS - 1) / FFTW32_WARPS;
	long long grid = (long long)sm_count * 16;
	if (grid > blocks_needed) grid = blocks_needed;
	if (grid < 1) grid = 1;
	fftcc2d_w32_kernel<<<(int)grid, FFTW32_WARPS * 32, 0, stream>>>(img, d_pois, (int)n);
	*err = cudaGetLastError();
	return *err == cudaSuccess ? 0 : -2;
}

} // namespace ocb
