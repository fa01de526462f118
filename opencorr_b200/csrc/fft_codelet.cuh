// fft_codelet.cuh -- N-point complex FFTs held entirely in the registers of ONE thread, for any
// N = 2^a 3^b 5^c (N <= 64): fully unrolled in-place mixed-radix decimation-in-frequency with
// compile-time twiddles.  Input in natural order; the value left at position `pos` is frequency
// fft_freq_of<N>(pos) (the mixed-radix digit reversal), which costs nothing because every index is a
// compile-time constant after unrolling -- callers scatter through that map when they store.
// Used by fftcc2d_reg.cu (thread-per-row 2D transforms); fft32.cuh remains the warp-per-POI special
// case for N = 32.
#pragma once
#include "ocb_common.cuh"

namespace ocb {

// ---- compile-time sin/cos of 2*pi*j/n (double precision Taylor after exact octant reduction) ----
constexpr double cx_pi = 3.14159265358979323846264338327950288;
constexpr double cx_sin_small(double x) { // |x| <= pi/4
	double x2 = x * x, term = x, sum = x;
	for (int i = 1; i < 14; i++) {
		term *= -x2 / (double)((2 * i) * (2 * i + 1));
		sum += term;
	}
	return sum;
}
constexpr double cx_cos_small(double x) {
	double x2 = x * x, term = 1.0, sum = 1.0;
	for (int i = 1; i < 14; i++) {
		term *= -x2 / (double)((2 * i - 1) * (2 * i));
		sum += term;
	}
	return sum;
}
// cos / sin of 2*pi*j/n, j any integer, n > 0: the angle is reduced with integer arithmetic on 8j/n
constexpr double cx_cos2pi(long long j, long long n) {
	j %= n;
	if (j < 0) j += n;
	if (2 * j > n) j = n - j;                                  // cos is even about pi
	if (4 * j > n) return -cx_cos2pi(n - 2 * j, 2 * n);        // cos(x) = -cos(pi - x); (n/2 - j)/n = (n - 2j)/(2n)
	if (8 * j > n) return cx_sin_small(2.0 * cx_pi * (double)(n - 4 * j) / (double)(4 * n)); // cos(x) = sin(pi/2 - x)
	return cx_cos_small(2.0 * cx_pi * (double)j / (double)n);
}
constexpr double cx_sin2pi(long long j, long long n) { return cx_cos2pi(4 * j - n, 4 * n); } // sin(x) = cos(x - pi/2)

// radix schedule: 4 while divisible, then 2, 3, 5
__host__ __device__ constexpr int fft_radix_of(int n) { return n % 4 == 0 ? 4 : (n % 2 == 0 ? 2 : (n % 3 == 0 ? 3 : (n % 5 == 0 ? 5 : n))); }
__host__ __device__ constexpr bool fft_codelet_ok(int n) {
	if (n < 2 || n > 64) return false;
	while (n % 2 == 0) n /= 2;
	while (n % 3 == 0) n /= 3;
	while (n % 5 == 0) n /= 5;
	return n == 1;
}
// frequency index held at position `pos` after the in-place DIF transform of length n
__host__ __device__ constexpr int fft_freq_of_n(int pos, int n) {
	if (n == 1) return 0;
	const int r = fft_radix_of(n), m = n / r;
	return pos / m + r * fft_freq_of_n(pos % m, m);
}
template <int N>
__host__ __device__ constexpr int fft_freq_of(int pos) { return fft_freq_of_n(pos, N); }
// f(pos, freq) for pos = 0 .. N-1 with BOTH arguments true compile-time constants (std::integral_constant-like tags):
// called as  fft_for_each_pos<N>([&](auto pos, auto freq) { ... re[pos.value] ... freq.value ... });
// (a plain loop calling the constexpr map leaves a run-time recursive function with integer divisions in the kernel)
template <int V>
struct FftConst {
	static constexpr int value = V;
};
template <int N, int POS = 0, class F>
__host__ __device__ __forceinline__ void fft_for_each_pos(F&& f) {
	if constexpr (POS < N) {
		f(FftConst<POS>{}, FftConst<fft_freq_of_n(POS, N)>{});
		fft_for_each_pos<N, POS + 1>(f);
	}
}

// twiddle W_N^k = exp(-2 pi i k / N) as compile-time constants; the if-chain folds once k is a constant
template <int N, int K>
struct FftTw {
	static constexpr float c = (float)cx_cos2pi(K, N);
	static constexpr float s = (float)cx_sin2pi(K, N);
};
template <int N, int K = 0>
__host__ __device__ __forceinline__ float fft_tw_cos(int k) {
	if constexpr (K >= N) return 1.f;
	else return k == K ? FftTw<N, K>::c : fft_tw_cos<N, K + 1>(k);
}
template <int N, int K = 0>
__host__ __device__ __forceinline__ float fft_tw_sin(int k) {
	if constexpr (K >= N) return 0.f;
	else return k == K ? FftTw<N, K>::s : fft_tw_sin<N, K + 1>(k);
}

// (yr, yi) = (xr + i xi) * W_N^k (forward) or its conjugate (inverse)
template <int N, bool INV>
__host__ __device__ __forceinline__ void fft_mul_tw(float xr, float xi, int k, float& yr, float& yi) {
	if (k == 0) {
		yr = xr;
		yi = xi;
	} else if (4 * k == N) { // -i (forward) / +i (inverse)
		yr = INV ? -xi : xi;
		yi = INV ? xr : -xr;
	} else if (2 * k == N) {
		yr = -xr;
		yi = -xi;
	} else if (4 * k == 3 * N) { // +i (forward) / -i (inverse)
		yr = INV ? xi : -xi;
		yi = INV ? -xr : xr;
	} else {
		const float c = fft_tw_cos<N>(k), s = INV ? -fft_tw_sin<N>(k) : fft_tw_sin<N>(k); // W = c - i s
		yr = fmaf(xr, c, xi * s);
		yi = fmaf(xi, c, -xr * s);
	}
}

// one DIF stage on blocks of length L (L divides N): radix R = fft_radix_of(L), then recurse on L / R
template <int N, bool INV, int L>
__host__ __device__ __forceinline__ void fft_dif_stage(float* re, float* im) {
	if constexpr (L > 1) {
		constexpr int R = fft_radix_of(L), M = L / R, TS = N / L; // W_L^k = W_N^(k * TS)
		constexpr float sgn = INV ? -1.f : 1.f;
#pragma unroll
		for (int b = 0; b < N; b += L) {
#pragma unroll
			for (int k = 0; k < M; k++) {
				const int i0 = b + k;
				if constexpr (R == 4) {
					const float a0r = re[i0], a0i = im[i0], a1r = re[i0 + M], a1i = im[i0 + M];
					const float a2r = re[i0 + 2 * M], a2i = im[i0 + 2 * M], a3r = re[i0 + 3 * M], a3i = im[i0 + 3 * M];
					const float t0r = a0r + a2r, t0i = a0i + a2i, t1r = a0r - a2r, t1i = a0i - a2i;
					const float t2r = a1r + a3r, t2i = a1i + a3i, t3r = a1r - a3r, t3i = a1i - a3i;
					const float ur = sgn * t3i, ui = -sgn * t3r; // t3 * (-i) forward, (+i) inverse
					re[i0] = t0r + t2r;
					im[i0] = t0i + t2i;
					fft_mul_tw<N, INV>(t1r + ur, t1i + ui, (k * TS) % N, re[i0 + M], im[i0 + M]);
					fft_mul_tw<N, INV>(t0r - t2r, t0i - t2i, (2 * k * TS) % N, re[i0 + 2 * M], im[i0 + 2 * M]);
					fft_mul_tw<N, INV>(t1r - ur, t1i - ui, (3 * k * TS) % N, re[i0 + 3 * M], im[i0 + 3 * M]);
				} else if constexpr (R == 2) {
					const float a0r = re[i0], a0i = im[i0], a1r = re[i0 + M], a1i = im[i0 + M];
					re[i0] = a0r + a1r;
					im[i0] = a0i + a1i;
					fft_mul_tw<N, INV>(a0r - a1r, a0i - a1i, (k * TS) % N, re[i0 + M], im[i0 + M]);
				} else if constexpr (R == 3) {
					const float a0r = re[i0], a0i = im[i0], a1r = re[i0 + M], a1i = im[i0 + M], a2r = re[i0 + 2 * M], a2i = im[i0 + 2 * M];
					constexpr float c = -0.5f, sn = -0.86602540378443864676f * sgn; // W_3 = c + i sn
					const float t1r = a1r + a2r, t1i = a1i + a2i, t2r = a1r - a2r, t2i = a1i - a2i;
					const float mr = fmaf(c, t1r, a0r), mi = fmaf(c, t1i, a0i);
					const float rr = -sn * t2i, ri = sn * t2r; // i sn t2
					re[i0] = a0r + t1r;
					im[i0] = a0i + t1i;
					fft_mul_tw<N, INV>(mr + rr, mi + ri, (k * TS) % N, re[i0 + M], im[i0 + M]);
					fft_mul_tw<N, INV>(mr - rr, mi - ri, (2 * k * TS) % N, re[i0 + 2 * M], im[i0 + 2 * M]);
				} else {
					static_assert(R == 5, "fft codelet: only radices 2, 3, 4, 5");
					const float a0r = re[i0], a0i = im[i0], a1r = re[i0 + M], a1i = im[i0 + M], a2r = re[i0 + 2 * M], a2i = im[i0 + 2 * M];
					const float a3r = re[i0 + 3 * M], a3i = im[i0 + 3 * M], a4r = re[i0 + 4 * M], a4i = im[i0 + 4 * M];
					constexpr float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
					constexpr float s1 = -0.95105651629515357212f * sgn, s2 = -0.58778525229247312917f * sgn;
					const float p1r = a1r + a4r, p1i = a1i + a4i, m1r = a1r - a4r, m1i = a1i - a4i;
					const float p2r = a2r + a3r, p2i = a2i + a3i, m2r = a2r - a3r, m2i = a2i - a3i;
					const float e1r = fmaf(c2, p2r, fmaf(c1, p1r, a0r)), e1i = fmaf(c2, p2i, fmaf(c1, p1i, a0i));
					const float e2r = fmaf(c1, p2r, fmaf(c2, p1r, a0r)), e2i = fmaf(c1, p2i, fmaf(c2, p1i, a0i));
					const float o1r = -(s1 * m1i + s2 * m2i), o1i = s1 * m1r + s2 * m2r;
					const float o2r = -(s2 * m1i - s1 * m2i), o2i = s2 * m1r - s1 * m2r;
					re[i0] = a0r + p1r + p2r;
					im[i0] = a0i + p1i + p2i;
					fft_mul_tw<N, INV>(e1r + o1r, e1i + o1i, (k * TS) % N, re[i0 + M], im[i0 + M]);
					fft_mul_tw<N, INV>(e2r + o2r, e2i + o2i, (2 * k * TS) % N, re[i0 + 2 * M], im[i0 + 2 * M]);
					fft_mul_tw<N, INV>(e2r - o2r, e2i - o2i, (3 * k * TS) % N, re[i0 + 3 * M], im[i0 + 3 * M]);
					fft_mul_tw<N, INV>(e1r - o1r, e1i - o1i, (4 * k * TS) % N, re[i0 + 4 * M], im[i0 + 4 * M]);
				}
			}
		}
		fft_dif_stage<N, INV, M>(re, im);
	}
}

// in-place transform: natural order in, frequency fft_freq_of<N>(pos) at position pos out
template <int N, bool INV>
__host__ __device__ __forceinline__ void fft_reg(float* re, float* im) {
	static_assert(fft_codelet_ok(N), "fft_reg: N must be 2^a 3^b 5^c, 2 <= N <= 64");
	fft_dif_stage<N, INV, N>(re, im);
}

} // namespace ocb
