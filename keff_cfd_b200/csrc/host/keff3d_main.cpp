// keff3D -- drop-in replacement of the reference's Keff3D program: input.txt (same keys incl. Width /
// Height / Depth / MeshAmpZ), "x,y,z" CSV structures, the 3D CSV row and the x,y,z,T map; the
// numerical path is libkeffb200.  Single volume (Keff3D/keff3D.cpp:59-186) and batch %05d.csv
// (:332-496).  Usage: keff3D [input.txt] [--check-inputs]
#include <chrono>

#include "host_common.hpp"

using namespace kbhost;
using clk = std::chrono::steady_clock;

int main(int argc, char** argv) {
  std::string in = "input.txt";
  bool check_only = false;
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--check-inputs")) check_only = true;
    else in = argv[i];
  }
  Options o;
  std::string err;
  if (!read_input_file(in, o, err)) { fprintf(stderr, "keff3D: %s\n", err.c_str()); return 1; }
  if (o.Width < 2 || o.Height < 1 || o.Depth < 1) { fprintf(stderr, "keff3D: Width/Height/Depth missing in %s\n", in.c_str()); return 1; }
  const keffb200_params p = params_from(o);
  const int nfiles = o.BatchFlag == 1 ? o.NumImg : 1;
  const size_t n = (size_t)o.Width * o.Height * o.Depth;
  if (!check_only && keffb200_init(0)) return die("init");
  for (int f = 0; f < nfiles; f++) {
    auto t0 = clk::now();
    char name[128];
    if (o.BatchFlag == 1) snprintf(name, sizeof name, "%05d.csv", f);
    else snprintf(name, sizeof name, "%s", o.inputFilename.c_str());
    std::vector<uint8_t> solid;
    if (!structure_from_csv(name, o.Width, o.Height, o.Depth, solid, err)) { fprintf(stderr, "keff3D: %s\n", err.c_str()); return 1; }
    size_t ns = 0;
    for (uint8_t v : solid) ns += v;
    // The reference's calcPorosity3D (keff3D.h:364-394) counts voxels equal to TCfluid, i.e. the SOLID
    // fraction when kf = 1; the void fraction is what the column name promises.
    const double porosity = 1.0 - (double)ns / (double)n;
    if (check_only) {
      printf("input ok: %s %dx%dx%d solid voxels %zu porosity %.10f\n", name, o.Width, o.Height, o.Depth, ns, porosity);
      continue;
    }
    std::vector<double> T;
    if (o.printTmap == 1) T.resize(n);
    keffb200_result r;
    if (keffb200_solve3d(solid.data(), o.Width, o.Height, o.Depth, &p, nullptr, o.printTmap == 1 ? T.data() : nullptr, nullptr,
                         nullptr, &r))
      return die("solve3d");
    if (o.verbose == 1) printf("Effective Thermal Conductivity = %g\n", r.keff);
    if (o.printTmap == 1) {
      char tn[160];
      if (o.BatchFlag == 1) snprintf(tn, sizeof tn, "TMAP_%05d.csv", f);
      else snprintf(tn, sizeof tn, "%s", o.TMapName.c_str());
      write_tmap(tn, T.data(), o.Width, o.Height, o.Depth, true);
    }
    if (o.printQmap == 1) printf("Heat Flux Map: Feature not currently available.\n");  // keff3D.cpp:179-181
    const double secs = std::chrono::duration<double>(clk::now() - t0).count();
    append_row(o, true, r.keff, r.q_left, r.q_right, r.iters, name, (int)n, porosity, secs, o.BatchFlag == 1 ? 1 : o.numCores);
    if (r.status != 0) fprintf(stderr, "keff3D: warning: stopped at MaxIter before convergence\n");
  }
  if (!check_only) keffb200_shutdown();
  return 0;
}
