// keff2D -- drop-in replacement of the reference's Keff2D / Keff2DGPU programs: reads input.txt
// (same keys) from the working directory, the same JPG structures, writes the same CSV rows and
// T-/Q-maps; the numerical path is libkeffb200 (C ABI).  Modes as in Keff2D/keff2D.cpp:
//   single image (:69-192), mesh-convergence sweep (:193-332), batch %05d.jpg (:333-485).
// Usage: keff2D [input.txt] [--check-inputs]   (--check-inputs: parse + decode only, no GPU)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>

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
    // The reference decodes and solves image by image inside an OpenMP loop (NumCores threads).  Here
    // NumCores threads only decode (JPEG -> structure bytes, straight into one staging array); the
    // images are handed to the GPU in chunks, so chunk k is being solved while chunk k+1 is decoded.
    if (o.NumImg < 1) { fprintf(stderr, "keff2D: NumImages must be >= 1 in batch mode\n"); return 1; }
    GrayImage first;
    if (!load_jpeg_gray("00000.jpg", first)) { fprintf(stderr, "keff2D: 00000.jpg: %s\n", first.err.c_str()); return 1; }
    const int nx = first.w * o.MeshIncreaseX, ny = first.h * o.MeshIncreaseY;
    const size_t cells = (size_t)nx * ny;
    // every visible GPU takes a share of each chunk (keffb200_init_multi); the staging array is page-locked so the
    // copies to the GPUs run at PCIe rate underneath the solves
    int n_gpus = 1;
    if (!check_only) {
      if (keffb200_init_multi(0)) return die("init");
      n_gpus = std::max(1, keffb200_device_count());
    }
    uint8_t* all = check_only ? (uint8_t*)malloc(cells * o.NumImg) : (uint8_t*)keffb200_host_alloc(cells * o.NumImg);
    const bool pinned = !check_only && all != nullptr;
    if (!all) all = (uint8_t*)malloc(cells * o.NumImg);
    if (!all) { fprintf(stderr, "keff2D: cannot allocate the image staging array\n"); return 1; }
    std::vector<double> poro(o.NumImg);
    std::vector<keffb200_result> res(o.NumImg);
    std::vector<double> T;
    if (o.printTmap && !check_only) T.resize(cells * o.NumImg);
    const int chunk = std::max(1, std::min(o.NumImg, 2048 * n_gpus));
    const int nchunks = (o.NumImg + chunk - 1) / chunk;
    std::vector<std::atomic<int>> left(nchunks);  // images of the chunk still being decoded
    for (int k = 0; k < nchunks; k++) left[k] = std::min(chunk, o.NumImg - k * chunk);
    std::atomic<int> next{0};
    std::atomic<bool> failed{false};
    std::mutex mu;
    std::condition_variable cv;
    std::string first_error;
    auto decode = [&]() {
      for (int img = next++; img < o.NumImg && !failed; img = next++) {
        char name[64];
        snprintf(name, sizeof name, "%05d.jpg", img);
        GrayImage g;
        std::string e;
        if (!load_jpeg_gray(name, g)) e = std::string(name) + ": " + g.err;
        else if (g.w * o.MeshIncreaseX != nx || g.h * o.MeshIncreaseY != ny) e = std::string(name) + ": all batch images must have the same size";
        if (!e.empty()) {
          std::lock_guard<std::mutex> lk(mu);
          if (first_error.empty()) first_error = e;
          failed = true;
          cv.notify_all();
          return;
        }
        if (g.channels != 1) fprintf(stderr, "keff2D: warning: %s has %d channels, using luminance\n", name, g.channels);
        std::vector<uint8_t> sl;
        structure_from_image(g, o.MeshIncreaseX, o.MeshIncreaseY, sl, poro[img]);
        memcpy(all + cells * img, sl.data(), cells);
        if (--left[img / chunk] == 0) {
          std::lock_guard<std::mutex> lk(mu);
          cv.notify_all();
        }
      }
    };
    const int nthreads = std::max(1, std::min(o.numCores, o.NumImg));
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back(decode);
    int rc = 0;
    for (int k = 0; k < nchunks && rc == 0; k++) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return failed || left[k] == 0; });
      }
      if (failed) break;
      if (check_only) continue;
      const int i0 = k * chunk, cnt = std::min(chunk, o.NumImg - i0);
      if (keffb200_solve2d_batch(all + cells * i0, cnt, nx, ny, &p, o.printTmap ? T.data() + cells * i0 : nullptr,
                                 res.data() + i0))
        rc = die("solve2d_batch");
    }
    if (rc != 0) failed = true;
    for (auto& th : pool) th.join();
    if (!first_error.empty()) { fprintf(stderr, "keff2D: %s\n", first_error.c_str()); return 1; }
    if (rc != 0) return rc;
    if (check_only) { printf("batch ok: %d images %dx%d\n", o.NumImg, nx, ny); return 0; }
    const double per = since(t0) / o.NumImg;
    for (int img = 0; img < o.NumImg; img++) {
      char name[64];
      snprintf(name, sizeof name, "%05d.csv", img);  // the reference names batch rows %05d.csv (keff2D.h:271)
      append_row(o, false, res[img].keff, res[img].q_left, res[img].q_right, res[img].iters, name, (long)nx * ny, poro[img], per, 1);
      if (o.printTmap) {
        snprintf(name, sizeof name, "TMAP_%05d.csv", img);
        write_tmap(name, T.data() + cells * img, nx, ny, 1, false);
      }
    }
    if (o.verbose)
      printf("Done, saving output (%d images, %d decode threads, %d GPU%s, %.3f s)\n", o.NumImg, nthreads, n_gpus, n_gpus > 1 ? "s" : "",
             since(t0));
    if (pinned) keffb200_host_free(all);
    else free(all);
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
    if (!Tprev.empty()) {
      // warm start from the previous mesh.  The reference replicates the native-mesh field (SolExpandSimple,
      // keff2D.h:614-642) -- and then overwrites it (keff2D.cpp:268); here the previous field is interpolated
      // bilinearly, with the two walls as nodes, so each refinement starts within discretisation error of its answer.
      Tinit.resize(T.size());
      if (keffb200_interpolate(Tprev.data(), pnx, pny, 1, Tinit.data(), nx, ny, 1, o.TempLeft, o.TempRight)) return die("interpolate");
    }
    keffb200_result res;
    if (keffb200_solve2d(solid.data(), nx, ny, &p, Tinit.empty() ? nullptr : Tinit.data(), T.data(), qL.data(), qR.data(), &res))
      return die("solve2d");
    printf("Keff = %f, Total Iterations = %ld\n", res.keff, res.iters);
    if (o.printTmap == 1 && o.MeshFlag == 0) write_tmap(o.TMapName, T.data(), nx, ny, 1, false);
    if (o.printQmap == 1 && o.MeshFlag == 0) {
      std::vector<double> Q((size_t)ny * (nx + 1));
      if (keffb200_flux_field(solid.data(), nx, ny, 1, &p, T.data(), Q.data())) return die("flux_field");
      write_qmap(o.QMapName, Q.data(), nx, ny, 1, false);
    }
    const double secs = since(t0);
    printf("Residual = %f\n", res.energy_residual);
    append_row(o, false, res.keff, res.q_left, res.q_right, res.iters, o.inputFilename.c_str(), (long)nx * ny, porosity, secs, o.numCores);
    printf("Time spent = %lf\n", secs);
    if (res.status != 0) fprintf(stderr, "keff2D: warning: stopped at MaxIter before convergence\n");
    Tprev.swap(T);
    pnx = nx;
    pny = ny;
  }
  keffb200_shutdown();
  return 0;
}
