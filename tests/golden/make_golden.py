"""Regenerates the committed fixtures under tests/golden/ from the reference checkout.

Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_golden.py
Only DATA is copied (images, sub-sampled result tables the reference ships as its regression
fixtures, examples/2d_dic and examples/dvc); no reference source code.

Outputs
  oht_cfrp_0.bmp, oht_cfrp_4.bmp    the 2D example pair (280x900, 8-bit), verbatim
  oht_cfrp_4_fftcc_icgn1_r16.npz    every 23rd row of the shipped result table + deformation table
  oht_cfrp_4_fftcc_iclm1_r16.npz    the same rows of the shipped ICLM2D1 table
  oht_cfrp_4_fftcc_nr1_r16.npz      every 23rd row of the shipped NR2D1 table (x,y,u,v,u0,v0,ZNCC,iteration,convergence)
                                    + a 96-row band (y in [370,560), all 100 columns, includes the specimen's hole)
                                    of x,y,u,v,ZNCC,exx,eyy,exy for the Strain test; rows with `band_check` have
                                    their whole 20-px neighbourhood inside the band
  torus_strain_crop.npz             a box of the shipped DVC table examples/dvc/Torus_def_sift_icgn1_r16.csv
                                    (x,y,z,u,v,w,ZNCC + 6 strains) for the 3D Strain test, same idea
  al_foam4_crop.npz                 z-slices [18,118) of the DVC example pair as uint8 (values are
                                    integral, 52..202) + the shipped CPU and GPU result rows of the
                                    196 POIs with z in {60,65,70,75}.  The 15-tap prefilter and the
                                    subsets of those POIs never reach within 7 voxels of the cut, so
                                    results on the crop equal results on the full volume.
"""
import os
import shutil

import numpy as np

REF = "/root/reference/examples"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ("oht_cfrp_0.bmp", "oht_cfrp_4.bmp"):
        shutil.copyfile(os.path.join(REF, "2d_dic", name), os.path.join(OUT, name))
    tab = np.genfromtxt(os.path.join(REF, "2d_dic", "oht_cfrp_4_fftcc_icgn1_r16.csv"), delimiter=",", skip_header=1)
    dtab = np.genfromtxt(os.path.join(REF, "2d_dic", "oht_cfrp_4_fftcc_icgn1_r16_deformation.csv"), delimiter=",", skip_header=1)
    sel = np.arange(0, tab.shape[0], 23)
    np.savez_compressed(os.path.join(OUT, "oht_cfrp_4_fftcc_icgn1_r16.npz"),
                        columns=np.array("x,y,u,v,u0,v0,ZNCC,iteration,convergence".split(",")),
                        table=tab[sel, :9], deformation_columns=np.array("x,y,u,ux,uy,v,vx,vy".split(",")),
                        deformation=dtab[sel, :8], rows=sel)

    # ICLM2D1 table shipped by the reference (examples/2d_dic/oht_cfrp_4_fftcc_iclm1_r16.csv, same POIs)
    itab = np.genfromtxt(os.path.join(REF, "2d_dic", "oht_cfrp_4_fftcc_iclm1_r16.csv"), delimiter=",", skip_header=1)
    np.savez_compressed(os.path.join(OUT, "oht_cfrp_4_fftcc_iclm1_r16.npz"),
                        columns=np.array("x,y,u,v,u0,v0,ZNCC,iteration,convergence".split(",")), table=itab[sel, :9], rows=sel)

    # NR2D1 + Strain table shipped by the reference (examples/2d_dic/oht_cfrp_4_fftcc_nr1_r16.csv)
    ntab = np.genfromtxt(os.path.join(REF, "2d_dic", "oht_cfrp_4_fftcc_nr1_r16.csv"), delimiter=",", skip_header=1)
    band = (ntab[:, 1] >= 370) & (ntab[:, 1] < 560)
    bt = ntab[band][:, [0, 1, 2, 3, 6, 10, 11, 12]].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "oht_cfrp_4_fftcc_nr1_r16.npz"),
                        columns=np.array("x,y,u,v,u0,v0,ZNCC,iteration,convergence".split(",")), table=ntab[sel, :9], rows=sel,
                        band_columns=np.array("x,y,u,v,ZNCC,exx,eyy,exy".split(",")), band=bt,
                        band_check=(bt[:, 1] >= 390) & (bt[:, 1] < 540))

    # DVC table with strains (examples/dvc/Torus_def_sift_icgn1_r16.csv; strain radius 30, min 5 neighbours,
    # examples/test_dvc_strain.cpp:48-54)
    tt = np.genfromtxt(os.path.join(REF, "dvc", "Torus_def_sift_icgn1_r16.csv"), delimiter=",", skip_header=1)
    box = (tt[:, 0] >= 470) & (tt[:, 0] <= 630) & (tt[:, 2] >= 351) & (tt[:, 2] <= 501)
    tb = tt[box][:, [0, 1, 2, 3, 4, 5, 9, 22, 23, 24, 25, 26, 27]].astype(np.float32)
    inner = (tb[:, 0] >= 500) & (tb[:, 0] <= 600) & (tb[:, 2] >= 381) & (tb[:, 2] <= 471)
    np.savez_compressed(os.path.join(OUT, "torus_strain_crop.npz"),
                        columns=np.array("x,y,z,u,v,w,ZNCC,exx,eyy,ezz,exy,eyz,ezx".split(",")), table=tb, check=inner)

    def load(p):
        d = np.fromfile(p, dtype=np.int32, count=3)
        v = np.fromfile(p, dtype=np.float32, offset=12)
        return v.reshape(d[2], d[1], d[0])

    z0, z1 = 18, 118
    ref = load(os.path.join(REF, "dvc", "al_foam4_0.bin"))[z0:z1]
    tar = load(os.path.join(REF, "dvc", "al_foam4_1.bin"))[z0:z1]
    assert np.all(ref == np.round(ref)) and ref.min() >= 0 and ref.max() <= 255
    assert np.all(tar == np.round(tar)) and tar.min() >= 0 and tar.max() <= 255
    cpu = np.genfromtxt(os.path.join(REF, "dvc", "al_foam4_1_fftcc_icgn1_r30.csv"), delimiter=",", skip_header=1)
    gpu = np.genfromtxt(os.path.join(REF, "dvc", "al_foam4_1_fftcc_icgn1(gpu)_r30.csv"), delimiter=",", skip_header=1)
    keep = cpu[:, 2] <= 75
    assert np.array_equal(cpu[keep, :3], gpu[keep, :3])
    np.savez_compressed(os.path.join(OUT, "al_foam4_crop.npz"), ref=ref.astype(np.uint8), tar=tar.astype(np.uint8),
                        z_offset=np.int32(z0),
                        cpu_columns=np.array("x,y,z,u,v,w,u0,v0,w0,ZNCC,iteration,convergence,ux,uy,uz,vx,vy,vz,wx,wy,wz".split(",")),
                        cpu_table=cpu[keep],
                        gpu_columns=np.array("x,y,z,u,v,w,u0,v0,w0,ZNCC,iteration,convergence,feature,ux,uy,uz,vx,vy,vz,wx,wy,wz".split(",")),
                        gpu_table=gpu[keep])
    for f in sorted(os.listdir(OUT)):
        print("%10d  %s" % (os.path.getsize(os.path.join(OUT, f)), f))


if __name__ == "__main__":
    main()
