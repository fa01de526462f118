// Host-side check of the register FFT codelets (opencorr_b200/csrc/fft_codelet.cuh): the same templates compiled for
// the CPU, every supported length against a double-precision DFT, forward and round trip.  Built and run by
// tests/test_fft_codelet_host.py (needs nvcc, no GPU).  Exit code 0 = all lengths within tolerance.
#include <cstdio>
#include <cmath>
#include <complex>
#include <vector>
#include "fft_codelet.cuh"
using namespace ocb;
template <int N> double check() {
	float re[N], im[N];
	std::vector<std::complex<double>> x(N);
	for (int i = 0; i < N; i++) { re[i] = (float)std::sin(1.3 * i + 0.2) * 10.f + i; im[i] = (float)std::cos(0.7 * i * i) * 5.f; x[i] = { re[i], im[i] }; }
	float fr[N], fi[N];
	for (int i = 0; i < N; i++) { fr[i] = re[i]; fi[i] = im[i]; }
	fft_reg<N, false>(fr, fi);
	double err = 0, mag = 0;
	for (int pos = 0; pos < N; pos++) {
		int k = fft_freq_of<N>(pos);
		std::complex<double> s = 0;
		for (int n = 0; n < N; n++) s += x[n] * std::polar(1.0, -2 * M_PI * k * n / N);
		err = std::max(err, std::abs(s - std::complex<double>(fr[pos], fi[pos])));
		mag = std::max(mag, std::abs(s));
	}
	// inverse of the forward result (scatter to natural order first) must give N * x
	float gr[N], gi[N];
	for (int pos = 0; pos < N; pos++) { gr[fft_freq_of<N>(pos)] = fr[pos]; gi[fft_freq_of<N>(pos)] = fi[pos]; }
	fft_reg<N, true>(gr, gi);
	double err2 = 0;
	for (int pos = 0; pos < N; pos++) { int n = fft_freq_of<N>(pos); err2 = std::max(err2, std::abs(std::complex<double>(gr[pos], gi[pos]) / (double)N - x[n])); }
	printf("N=%2d fwd rel err %.2e  roundtrip err %.2e\n", N, err / mag, err2);
	return std::max(err / mag, err2 * 1e-2);
}
int main() {
	double worst = 0;
	worst = std::max(worst, check<8>()); worst = std::max(worst, check<10>()); worst = std::max(worst, check<12>());
	worst = std::max(worst, check<16>()); worst = std::max(worst, check<18>()); worst = std::max(worst, check<20>());
	worst = std::max(worst, check<24>()); worst = std::max(worst, check<30>()); worst = std::max(worst, check<32>());
	worst = std::max(worst, check<36>()); worst = std::max(worst, check<40>()); worst = std::max(worst, check<48>());
	worst = std::max(worst, check<50>()); worst = std::max(worst, check<54>()); worst = std::max(worst, check<60>());
	worst = std::max(worst, check<64>());
	printf("worst %.2e\n", worst);
	return worst < 5e-7 ? 0 : 1;
}
