// shim_bench.cpp -- end-to-end timing of the FFT-CC -> IC-GN path THROUGH THE C++ SHIM (include/opencorr/opencorr.h), i.e. the
// calls an OpenCorr program makes, with the caller's own pageable memory (Image2D pixels, std::vector<POI2D>): the number
// bench.py reports as "e2e_shim".  One step = what examples/test_2d_dic_fftcc_icgn1.cpp does per image pair:
//     fftcc.setImages; fftcc.compute(queue); icgn.setImages; icgn.prepare(); icgn.compute(queue)
// usage: shim_bench ref.pgm tar.pgm x0 y0 nx ny sx sy radius order conv stop steps warmup
// Multi-GPU: OPENCORR_B200_DEVICES=all (the shim's switch) shards compute(queue) over the visible devices.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "opencorr.h"

using namespace opencorr;

int main(int argc, char** argv)
{
	if (argc < 15) {
		std::fprintf(stderr, "usage: %s ref.pgm tar.pgm x0 y0 nx ny sx sy radius order conv stop steps warmup\n", argv[0]);
		return 2;
	}
	try {
		const auto t_start = std::chrono::steady_clock::now();
		Image2D ref_img(argv[1]), tar_img(argv[2]);
		const int x0 = std::atoi(argv[3]), y0 = std::atoi(argv[4]), nx = std::atoi(argv[5]), ny = std::atoi(argv[6]);
		const int sx = std::atoi(argv[7]), sy = std::atoi(argv[8]), r = std::atoi(argv[9]), order = std::atoi(argv[10]);
		const float conv = (float)std::atof(argv[11]), stop = (float)std::atof(argv[12]);
		const int steps = std::atoi(argv[13]), warmup = std::atoi(argv[14]);
		std::vector<POI2D> base;
		base.reserve((size_t)nx * ny);
		for (int i = 0; i < ny; i++)
			for (int j = 0; j < nx; j++) base.push_back(POI2D(Point2D((float)(x0 + j * sx), (float)(y0 + i * sy))));
		FFTCC2D fftcc(r, r, 1);
		ICGN2D1 icgn1(r, r, conv, stop, 1);
		ICGN2D2 icgn2(r, r, conv, stop, 1);
		std::vector<double> ms;
		std::vector<POI2D> queue;
		double first_step_ms = 0;
		for (int s = 0; s < warmup + steps; s++) {
			queue = base; // the step's input, prepared outside the clock
			const auto t0 = std::chrono::steady_clock::now();
			fftcc.setImages(ref_img, tar_img);
			fftcc.compute(queue);
			if (order == 1) {
				icgn1.setImages(ref_img, tar_img);
				icgn1.prepare();
				icgn1.compute(queue);
			} else {
				icgn2.setImages(ref_img, tar_img);
				icgn2.prepare();
				icgn2.compute(queue);
			}
			const auto t1 = std::chrono::steady_clock::now();
			const double dt = std::chrono::duration<double, std::milli>(t1 - t0).count();
			if (s == 0) first_step_ms = std::chrono::duration<double, std::milli>(t1 - t_start).count();
			if (s >= warmup) ms.push_back(dt);
		}
		size_t good = 0;
		double su = 0;
		for (const POI2D& p : queue)
			if (p.result.zncc >= 0) { good++; su += p.deformation.u; }
		std::printf("{\"n_poi\": %zu, \"converged\": %zu, \"mean_u\": %.6f, \"first_step_incl_start_up_ms\": %.3f, \"ms\": [", queue.size(), good,
			good ? su / good : 0.0, first_step_ms);
		for (size_t i = 0; i < ms.size(); i++) std::printf("%s%.4f", i ? ", " : "", ms[i]);
		std::printf("]}\n");
	} catch (const std::string& e) {
		std::fprintf(stderr, "shim_bench: %s\n", e.c_str());
		return 1;
	}
	return 0;
}
