// keff2D -- drop-in replacement of the reference's Keff2D / Keff2DGPU programs: reads input.txt
// (same keys) from the working directory, the same JPG structures, writes the same CSV rows and
// T-/Q-maps; the numerical path is libkeffb200 (C ABI).  Modes as in Keff2D/keff2D.cpp:
//   single image (:69-192), mesh-convergence sweep (:193-332), batch %05d.jpg (:333-485).
// Usage: keff2D [input.txt] [--check-inputs]   (--check-inputs: parse + decode only, no GPU)
#include <chrono>

#include "host_common.hpp"

using namespace kbhost;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

int main(int argc, char** argv) {
  std::string in = "input.txt";
  bool check_only = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--check-inputs")) check_only = true;
    else in = argv[i];
  }
  Options o;
  std::string err;
  if (!read_input_file(in, o, err)) { fprintf(stderr, "keff2D: %s\n", err.c_str()); return 1; }
  if (o.MeshIncreaseX < 1 || o.MeshIncreaseY < 1) { fprintf(stderr, "keff2D: MeshAmp must be an integer >= 1\n"); return 1; }
  keffb200_params p = params_from(o);
  auto t0 = clk::now();

  if (o.BatchFlag == 1 && o.MeshFlag == 0) {
    // ---- batch: %05d.jpg, img = 0 .. NumImages-1 (keff2D.cpp:341-360) ----
    std::vector<uint8_t> all;
    std::vector<double> poro(o.NumImg);
    int nx = 0, ny = 0;
    for (int img = 0; img < o.NumImg; img++) {
      char name[64];
      snprintf(name, sizeof name, "%05d.jpg", img);
      GrayImage g;
      if (!load_jpeg_gray(name, g)) { fprintf(stderr, "keff2D: %s: %s\n", name, g.err.c_str()); return 1; }
      if (g.channels != 1) fprintf(stderr, "keff2D: warning: %s has %d channels, using luminance\n", name, g.channels);
      std::vector<uint8_t> s;
      structure_from_image(g, o.MeshIncreaseX, o.MeshIncreaseY, s, poro[img]);
      if (img == 0) { nx = g.w * o.MeshIncreaseX; ny = g.h * o.MeshIncreaseY; }
      if (g.w * o.MeshIncreaseX != nx || g.h * o.MeshIncreaseY != ny) {
        fprintf(stderr, "keff2D: %s: all batch images must have the same size\n", name);
        return 1;
      }
      all.insert(all.end(), s.begin(), s.end());
    }
    if (check_only) { printf("batch ok: %d images %dx%d\n", o.NumImg, nx, ny); return 0; }
    if (keffb200_init(0)) return die("init");
    std::vector<keffb200_result> res(o.NumImg);
    std::vector<double> T;
    if (o.printTmap) T.resize(all.size());
    if (keffb200_solve2d_batch(all.data(), o.NumImg, nx, ny, &p, o.printTmap ? T.data() : nullptr, res.data()))
      return die("solve2d_batch");
    const double per = since(t0) / o.NumImg;
    for (int img = 0; img < o.NumImg; img++) {
      char name[64];
      snprintf(name, sizeof name, "%05d.csv", img);  // the reference names batch rows %05d.csv (keff2D.h:271)
      append_row(o, false, res[img].keff, res[img].q_left, res[img].q_right, res[img].iters, name, nx * ny, poro[img], per, 1);
      if (o.printTmap) {
        snprintf(name, sizeof name, "TMAP_%05d.csv", img);
        write_tmap(name, T.data() + (size_t)img * nx * ny, nx, ny, 1, false);
      }
    }
    if (o.verbose) printf("Done, saving output\n");
    keffb200_shutdown();
    return 0;
  }

  GrayImage g;
  if (!load_jpeg_gray(o.inputFilename, g)) { fprintf(stderr, "keff2D: %s: %s\n", o.inputFilename.c_str(), g.err.c_str()); return 1; }
  if (g.channels != 1) {
    // keff2D.cpp:60-63
    printf("Error: please enter a grascale image with 1 channel.\n Current number of channels = %d\n", g.channels);
    return 1;
  }
  if (check_only) {
    std::vector<uint8_t> s;
    double por;
    structure_from_image(g, o.MeshIncreaseX, o.MeshIncreaseY, s, por);
    printf("input ok: %dx%d porosity %.10f cells %zu ks %g kf %g TL %g TR %g tol %g maxit %ld\n", g.w, g.h, por, s.size(),
           o.TCsolid, o.TCfluid, o.TempLeft, o.TempRight, o.ConvergeCriteria, o.MAX_ITER);
    return 0;
  }
  if (keffb200_init(0)) return die("init");

  const int mlo = o.MeshFlag == 1 ? 1 : 0, mhi = o.MeshFlag == 1 ? o.MaxMesh : 0;
  std::vector<double> Tprev;
  int pnx = 0, pny = 0;
  for (int m = mlo; m <= mhi; m++) {
    t0 = o.MeshFlag == 1 ? clk::now() : t0;
    const int ax = o.MeshFlag == 1 ? m : o.MeshIncreaseX, ay = o.MeshFlag == 1 ? m : o.MeshIncreaseY;
    std::vector<uint8_t> solid;
    double porosity;
    structure_from_image(g, ax, ay, solid, porosity);
    const int nx = g.w * ax, ny = g.h * ay;
    std::vector<double> T((size_t)nx * ny), qL(ny), qR(ny), Tinit;
    if (!Tprev.empty()) {  // warm start: replicate the previous mesh's field (SolExpandSimple, keff2D.h:614-642)
      Tinit.resize(T.size());
      for (int r = 0; r < ny; r++)
        for (int c = 0; c < nx; c++) Tinit[(size_t)r * nx + c] = Tprev[(size_t)(r * pny / ny) * pnx + c * pnx / nx];
    }
    keffb200_result res;
    if (keffb200_solve2d(solid.data(), nx, ny, &p, Tinit.empty() ? nullptr : Tinit.data(), T.data(), qL.data(), qR.data(), &res))
      return die("solve2d");
    printf("Keff = %f, Total Iterations = %ld\n", res.keff, res.iters);
    if (o.printTmap == 1 && o.MeshFlag == 0) write_tmap(o.TMapName, T.data(), nx, ny, 1, false);
    if (o.printQmap == 1 && o.MeshFlag == 0) write_qmap2d(o.QMapName, T.data(), solid.data(), nx, ny, o.TCsolid, o.TCfluid, qL.data(), qR.data());
    const double secs = since(t0);
    printf("Residual = %f\n", res.energy_residual);
    append_row(o, false, res.keff, res.q_left, res.q_right, res.iters, o.inputFilename.c_str(), nx * ny, porosity, secs, o.numCores);
    printf("Time spent = %lf\n", secs);
    if (res.status != 0) fprintf(stderr, "keff2D: warning: stopped at MaxIter before convergence\n");
    Tprev.swap(T);
    pnx = nx;
    pny = ny;
  }
  keffb200_shutdown();
  return 0;
}
