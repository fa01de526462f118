// Host-only check of the C++ shim's file formats (include/opencorr/opencorr.h): BMP / single- and multi-page TIFF /
// .bin loaders and the result-table writers and loaders (POI2D, POI2DS, POI3D).  No GPU call is made: only classes
// that do not derive from DIC/DVC are touched.  Built and run by tests/test_shim_io_host.py.
//   usage: shim_io_test <dir>   (dir holds img.tif, stack.tif, vol.bin written by the Python side with known contents)
#include <cmath>
#include <cstdio>
#include <iomanip>
#include <sstream>

#include "opencorr.h"

using namespace opencorr;

static int fails = 0;
#define CHECK(c)                                                     \
	do {                                                             \
		if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } \
	} while (0)

int main(int argc, char** argv)
{
	if (argc < 2) return 2;
	const std::string dir = argv[1];
	try {
		// images: pixel (r, c) = (7 r + 3 c) % 251, volume voxel (z, y, x) = (5 z + 7 y + 3 x) % 251 (written by the test driver)
		Image2D img(dir + "/img.tif");
		CHECK(img.width == 16 && img.height == 12);
		for (int r = 0; r < img.height; r++)
			for (int c = 0; c < img.width; c++) CHECK(img.eg_mat(r, c) == (float)((7 * r + 3 * c) % 251));
		{ // every load / construction gets its own generation, also when a new image lands on the address of a destroyed one
			// (the engine's "is this pair already on the device?" test relies on it), and pixel buffers copy by value
			unsigned long long seen[4];
			const void* where[4];
			for (int i = 0; i < 4; i++) {
				Image2D again(dir + "/img.tif");
				seen[i] = again.generation;
				where[i] = &again;
				CHECK(again.eg_mat(3, 5) == (float)((7 * 3 + 3 * 5) % 251));
			}
			CHECK(where[0] == where[1]); // same stack slot ...
			for (int i = 1; i < 4; i++) CHECK(seen[i] != seen[i - 1] && seen[i] > img.generation); // ... never the same generation
			Image2D blank(16, 12);
			CHECK(blank.generation > seen[3] && blank.eg_mat(0, 0) == 0.f);
			Image2D copy = img;
			CHECK(copy.generation != img.generation); // a copy may be edited on its own: the engine must not take it for the original
			copy.eg_mat(2, 2) = -7.f;
			CHECK(img.eg_mat(2, 2) == (float)((7 * 2 + 3 * 2) % 251) && copy.eg_mat(2, 2) == -7.f);
			const unsigned long long g0 = blank.generation;
			blank.load(dir + "/img.tif");
			CHECK(blank.generation > g0 && blank.width == 16);
		}
		Image3D stack(dir + "/stack.tif");
		CHECK(stack.dim_x == 8 && stack.dim_y == 6 && stack.dim_z == 4);
		Image3D vol(dir + "/vol.bin");
		CHECK(vol.dim_x == 8 && vol.dim_y == 6 && vol.dim_z == 4);
		for (int z = 0; z < 4; z++)
			for (int y = 0; y < 6; y++)
				for (int x = 0; x < 8; x++) {
					CHECK(stack.vol_mat[z][y][x] == (float)((5 * z + 7 * y + 3 * x) % 251));
					CHECK(vol.vol_mat[z][y][x] == (float)((5 * z + 7 * y + 3 * x) % 251) + 0.5f);
				}

		// table round trips; the writers must produce what `ofstream << fixed << setprecision(8)` produces (src/oc_io.cpp:320-322)
		std::vector<POI2D> q2;
		std::vector<POI2DS> qs;
		std::vector<POI3D> q3;
		for (int i = 0; i < 50; i++) {
			POI2D a((float)(30 + 2 * i), 41.f);
			POI2DS s((float)(30 + 2 * i), 77.f);
			POI3D b((float)i, 2.f * i, 100.f - i);
			for (int k = 0; k < 12; k++) { a.deformation.p[k] = std::sin(0.37f * (i + k)) * (k == 0 ? 40.f : 1e-3f); b.deformation.p[k] = std::cos(0.11f * (i * k + 1)); }
			for (int k = 0; k < 6; k++) a.result.r[k] = (k == 3) ? (float)(i % 7) : std::cos(0.2f * i + k);
			for (int k = 0; k < 3; k++) { a.strain.e[k] = 1e-4f * (i - 20 + k); s.deformation.p[k] = 0.01f * (i + k) - 0.3f; }
			for (int k = 0; k < 9; k++) s.result.r[k] = 0.9f + 0.001f * (i + k);
			s.ref_coor = Point3D(-37.6f + i, -26.2f, 393.4f);
			s.tar_coor = Point3D(-41.6f + i, -27.0f, 394.3f);
			for (int k = 0; k < 6; k++) { s.strain.e[k] = -1e-3f * (k + 1) + 1e-5f * i; b.strain.e[k] = 2e-3f * k - 1e-5f * i; }
			for (int k = 0; k < 7; k++) b.result.r[k] = std::sin(0.05f * (i + 3 * k));
			a.subset_radius.x = a.subset_radius.y = 16.f;
			s.subset_radius.x = s.subset_radius.y = 9.f;
			b.subset_radius.x = b.subset_radius.y = b.subset_radius.z = 30.f;
			q2.push_back(a); qs.push_back(s); q3.push_back(b);
		}
		IO2D io;
		io.setDelimiter(",");
		io.setPath(dir + "/t2d.csv");
		io.saveTable2D(q2);
		std::vector<POI2D> r2 = io.loadTable2D();
		CHECK(r2.size() == q2.size());
		// values pass through 8 decimals: compare against the same rounding
		auto near8 = [](float a, float b) { return std::fabs((double)a - (double)b) <= 0.6e-8 + 1e-7 * std::fabs((double)b); };
		for (size_t i = 0; i < r2.size() && i < q2.size(); i++) {
			CHECK(r2[i].x == q2[i].x && r2[i].y == q2[i].y);
			CHECK(near8(r2[i].deformation.u, q2[i].deformation.u) && near8(r2[i].deformation.v, q2[i].deformation.v));
			for (int k = 0; k < 6; k++) CHECK(near8(r2[i].result.r[k], q2[i].result.r[k]));
			for (int k = 0; k < 3; k++) CHECK(near8(r2[i].strain.e[k], q2[i].strain.e[k]));
			CHECK(r2[i].subset_radius.x == 16.f && r2[i].subset_radius.y == 16.f);
		}
		{ // byte-level check of the first data row against iostream formatting
			std::ifstream f(dir + "/t2d.csv");
			std::string header, row;
			std::getline(f, header);
			std::getline(f, row);
			CHECK(header == "x,y,u,v,u0,v0,ZNCC,iteration,convergence,feature,exx,eyy,exy,subset_rx,subset_ry,");
			std::ostringstream o;
			o.setf(std::ios::fixed);
			o << std::setprecision(8);
			const POI2D& p = q2[0];
			o << p.x << "," << p.y << "," << p.deformation.u << "," << p.deformation.v << ",";
			for (int k = 0; k < 6; k++) o << p.result.r[k] << ",";
			for (int k = 0; k < 3; k++) o << p.strain.e[k] << ",";
			o << p.subset_radius.x << "," << p.subset_radius.y << ",";
			CHECK(row == o.str());
		}
		io.setPath(dir + "/t2ds.csv");
		io.saveTable2DS(qs);
		std::vector<POI2DS> rs = io.loadTable2DS();
		CHECK(rs.size() == qs.size());
		for (size_t i = 0; i < rs.size() && i < qs.size(); i++) {
			for (int k = 0; k < 3; k++) CHECK(near8(rs[i].deformation.p[k], qs[i].deformation.p[k]));
			for (int k = 0; k < 9; k++) CHECK(near8(rs[i].result.r[k], qs[i].result.r[k]));
			CHECK(near8(rs[i].ref_coor.x, qs[i].ref_coor.x) && near8(rs[i].tar_coor.z, qs[i].tar_coor.z));
			for (int k = 0; k < 6; k++) CHECK(near8(rs[i].strain.e[k], qs[i].strain.e[k]));
			CHECK(rs[i].subset_radius.x == 9.f);
		}
		IO3D io3;
		io3.setDelimiter(",");
		io3.setPath(dir + "/t3d.csv");
		io3.saveTable3D(q3);
		std::vector<POI3D> r3 = io3.loadTable3D();
		CHECK(r3.size() == q3.size());
		for (size_t i = 0; i < r3.size() && i < q3.size(); i++) {
			CHECK(r3[i].x == q3[i].x && r3[i].z == q3[i].z);
			for (int k = 0; k < 12; k++) CHECK(near8(r3[i].deformation.p[k], q3[i].deformation.p[k]));
			for (int k = 0; k < 7; k++) CHECK(near8(r3[i].result.r[k], q3[i].result.r[k]));
			for (int k = 0; k < 6; k++) CHECK(near8(r3[i].strain.e[k], q3[i].strain.e[k]));
			CHECK(r3[i].subset_radius.z == 30.f);
		}
		// a shipped-style table without the subset_rx/ry columns loads too (zeros)
		{
			std::ofstream f(dir + "/old.csv");
			f << "x,y,u,v,u0,v0,ZNCC,iteration,convergence,feature,exx,eyy,exy\n30,30,-0.4151558,-4.13913202,0,-4,0.99734747,5,0.00023785,0,-0.00046931,0.00296533,-0.00030154\n";
		}
		io.setPath(dir + "/old.csv");
		std::vector<POI2D> ro = io.loadTable2D();
		CHECK(ro.size() == 1 && ro[0].result.zncc == 0.99734747f && ro[0].strain.exy == -0.00030154f && ro[0].subset_radius.x == 0.f);
		// error behaviour: the reference throws std::string on unreadable files (src/oc_image.cpp:43)
		bool threw = false;
		try { Image2D missing(dir + "/nope.bmp"); } catch (const std::string&) { threw = true; }
		CHECK(threw);
	} catch (const std::string& e) {
		std::printf("exception: %s\n", e.c_str());
		return 1;
	}
	std::printf(fails ? "FAILED (%d)\n" : "ok\n", fails);
	return fails ? 1 : 0;
}
