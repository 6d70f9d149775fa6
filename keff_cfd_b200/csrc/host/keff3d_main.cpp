// keff3D -- drop-in replacement of the reference's Keff3D program: input.txt (same keys incl. Width /
// Height / Depth / MeshAmpZ), "x,y,z" CSV structures, the 3D CSV row and the x,y,z,T map; the
// numerical path is libkeffb200 on every visible GPU (large volumes are cut into z-slabs inside the
// library).  Modes as in Keff3D/keff3D.cpp: single volume (:59-186), mesh-convergence sweep
// m = 1..MaxMesh (:187-331), batch %05d.csv (:332-496).  Usage: keff3D [input.txt] [--check-inputs]
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
  if (!read_input_file(in, o, err)) { fprintf(stderr, "keff3D: %s\n", err.c_str()); return 1; }
  if (o.Width < 2 || o.Height < 1 || o.Depth < 1) { fprintf(stderr, "keff3D: Width/Height/Depth missing in %s\n", in.c_str()); return 1; }
  const keffb200_params p = params_from(o);
  const bool mesh_mode = o.MeshFlag == 1 && o.BatchFlag == 0;
  const int nfiles = o.BatchFlag == 1 ? o.NumImg : 1;
  const size_t n0 = (size_t)o.Width * o.Height * o.Depth;
  if (!check_only && keffb200_init_multi(0)) return die("init");
  for (int f = 0; f < nfiles; f++) {
    auto t0 = clk::now();
    char name[128];
    if (o.BatchFlag == 1) snprintf(name, sizeof name, "%05d.csv", f);
    else snprintf(name, sizeof name, "%s", o.inputFilename.c_str());
    std::vector<uint8_t> native;
    if (!structure_from_csv(name, o.Width, o.Height, o.Depth, native, err)) { fprintf(stderr, "keff3D: %s\n", err.c_str()); return 1; }
    size_t ns = 0;
    for (uint8_t v : native) ns += v;
    // The reference's calcPorosity3D (keff3D.h:364-394) counts voxels equal to TCfluid, i.e. the SOLID
    // fraction when kf = 1; the void fraction is what the column name promises.
    const double porosity = 1.0 - (double)ns / (double)n0;
    if (check_only) {
      printf("input ok: %s %dx%dx%d solid voxels %zu porosity %.10f\n", name, o.Width, o.Height, o.Depth, ns, porosity);
      continue;
    }
    // single volume / batch: the native mesh (the reference fixes the amplification to 1 there, keff3D.cpp:62-64);
    // mesh-convergence mode: every direction amplified by m = 1 .. MaxMesh (keff3D.cpp:189-196)
    const int mlo = 1, mhi = mesh_mode ? std::max(1, o.MaxMesh) : 1;
    std::vector<double> Tprev;
    int pnx = 0, pny = 0, pnz = 0;
    for (int m = mlo; m <= mhi; m++) {
      if (mesh_mode) t0 = clk::now();
      const int nx = o.Width * m, ny = o.Height * m, nz = o.Depth * m;
      const size_t n = (size_t)nx * ny * nz;
      std::vector<uint8_t> amplified;
      if (m > 1) amplify3d(native, o.Width, o.Height, o.Depth, m, m, m, amplified);
      const std::vector<uint8_t>& solid = m > 1 ? amplified : native;
      const bool want_T = o.printTmap == 1 || o.printQmap == 1 || (mesh_mode && m < mhi);
      std::vector<double> T, Tinit;
      if (want_T) T.resize(n);
      if (!Tprev.empty()) {
        // warm start: the previous mesh's field interpolated trilinearly (the reference replicates the native
        // field with SolExpandSimple and then overwrites it with the constant start, keff3D.cpp:262-266)
        Tinit.resize(n);
        if (keffb200_interpolate(Tprev.data(), pnx, pny, pnz, Tinit.data(), nx, ny, nz, o.TempLeft, o.TempRight)) return die("interpolate");
      }
      keffb200_result r;
      if (keffb200_solve3d(solid.data(), nx, ny, nz, &p, Tinit.empty() ? nullptr : Tinit.data(), want_T ? T.data() : nullptr, nullptr,
                           nullptr, &r))
        return die("solve3d");
      if (o.verbose == 1) {
        printf("Effective Thermal Conductivity = %g\n", r.keff);
        if (mesh_mode) printf("Time (s) = %g, Number of Elements = %zu\n", since(t0), n);
      }
      if (o.printTmap == 1 && !mesh_mode) {
        char tn[160];
        if (o.BatchFlag == 1) snprintf(tn, sizeof tn, "TMAP_%05d.csv", f);
        else snprintf(tn, sizeof tn, "%s", o.TMapName.c_str());
        write_tmap(tn, T.data(), nx, ny, nz, true);
      }
      if (o.printQmap == 1 && !mesh_mode) {
        // the map the reference announces as "Feature not currently available" (keff3D.cpp:179-181)
        std::vector<double> Q((size_t)nz * ny * (nx + 1));
        if (keffb200_flux_field(solid.data(), nx, ny, nz, &p, T.data(), Q.data())) return die("flux_field");
        char qn[160];
        if (o.BatchFlag == 1) snprintf(qn, sizeof qn, "QMAP_%05d.csv", f);
        else snprintf(qn, sizeof qn, "%s", o.QMapName.c_str());
        write_qmap(qn, Q.data(), nx, ny, nz, true);
      }
      append_row(o, true, r.keff, r.q_left, r.q_right, r.iters, name, (long)n, porosity, since(t0), o.BatchFlag == 1 ? 1 : o.numCores);
      if (r.status != 0) fprintf(stderr, "keff3D: warning: stopped at MaxIter before convergence\n");
      if (mesh_mode) {
        printf("--------------------------------------\n\n");
        Tprev.swap(T);
        pnx = nx;
        pny = ny;
        pnz = nz;
      }
    }
  }
  if (!check_only) keffb200_shutdown();
  return 0;
}
