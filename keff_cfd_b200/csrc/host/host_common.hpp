// host_common.hpp -- the drop-in shell around the C ABI: the reference's input.txt keys, structure
// readers and output writers.  Formats follow the reference byte for byte where a
// downstream script could notice (CSV rows, T-map / Q-map columns); the code is this repo's own.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/keffb200.h"
#include "image_gray.hpp"

namespace kbhost {

// The reference `options` struct (Keff2D/keff2D.h:15-36, Keff3D/keff3D.h:13-38).  Unlike the
// reference, every field has a defined default (the shipped input.txt values), so a missing key is
// not undefined behaviour.
struct Options {
  double TCsolid = 10, TCfluid = 1;
  int Width = 0, Height = 0, Depth = 0;
  int MeshIncreaseX = 1, MeshIncreaseY = 1, MeshIncreaseZ = 1;
  double TempLeft = 0, TempRight = 1;
  long MAX_ITER = 500000;
  double ConvergeCriteria = 1e-5;
  std::string inputFilename, outputFilename = "output.csv", TMapName = "TMAP.csv", QMapName = "QMAP.csv";
  int printTmap = 0, printQmap = 0, verbose = 0, numCores = 1, MeshFlag = 0, MaxMesh = 1, BatchFlag = 0, NumImg = 0;
};

// readInputFile (Keff2D/keff2D.h:114-218, Keff3D/keff3D.h:117-233): "key: value" lines, keys
// case-sensitive with the trailing colon, any order, anything else ignored.
inline bool read_input_file(const std::string& path, Options& o, std::string& err) {
  std::ifstream in(path);
  if (!in) { err = "cannot open " + path; return false; }
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string key, val;
    if (!(ls >> key >> val)) continue;
    const double d = atof(val.c_str());
    if (key == "ks:") o.TCsolid = d;
    else if (key == "kf:") o.TCfluid = d;
    else if (key == "Width:") o.Width = (int)d;
    else if (key == "Height:") o.Height = (int)d;
    else if (key == "Depth:") o.Depth = (int)d;
    else if (key == "MeshAmpX:") o.MeshIncreaseX = (int)d;
    else if (key == "MeshAmpY:") o.MeshIncreaseY = (int)d;
    else if (key == "MeshAmpZ:") o.MeshIncreaseZ = (int)d;
    else if (key == "InputName:") o.inputFilename = val;
    else if (key == "TR:") o.TempRight = d;
    else if (key == "TL:") o.TempLeft = d;
    else if (key == "OutputName:") o.outputFilename = val;
    else if (key == "printTMap:") o.printTmap = (int)d;
    else if (key == "TMapName:") o.TMapName = val;
    else if (key == "printQMap:") o.printQmap = (int)d;
    else if (key == "QMapName:") o.QMapName = val;
    else if (key == "Convergence:") o.ConvergeCriteria = d;
    else if (key == "MaxIter:") o.MAX_ITER = (long)d;
    else if (key == "Verbose:") o.verbose = (int)d;
    else if (key == "NumCores:") o.numCores = d < 1 ? 1 : (int)d;  // "Entered number of cores not valid. Default = 1."
    else if (key == "MeshFlag:") o.MeshFlag = (int)d;
    else if (key == "MaxMesh:") o.MaxMesh = (int)d;
    else if (key == "RunBatch:") o.BatchFlag = (int)d;
    else if (key == "NumImages:") o.NumImg = (int)d;
  }
  return true;
}

inline keffb200_params params_from(const Options& o) {
  keffb200_params p;
  keffb200_default_params(&p);
  p.ks = o.TCsolid;
  p.kf = o.TCfluid;
  p.tl = o.TempLeft;
  p.tr = o.TempRight;
  p.convergence = o.ConvergeCriteria;
  p.max_iter = o.MAX_ITER;
  return p;
}

// 2D structure: pixel < 150 is fluid, else solid, replicated by the mesh amplification
// (Keff2D/keff2D.cpp:108-121); porosity = fraction of pixels < 150 (calcPorosity, keff2D.h:306-331).
inline void structure_from_image(const GrayImage& g, int ampX, int ampY, std::vector<uint8_t>& solid, double& porosity) {
  const int nx = g.w * ampX, ny = g.h * ampY;
  solid.resize((size_t)nx * ny);
  size_t fluid = 0;
  for (size_t i = 0; i < g.pix.size(); i++) fluid += g.pix[i] < 150;
  porosity = (double)fluid / (double)g.pix.size();
  for (int r = 0; r < ny; r++)
    for (int c = 0; c < nx; c++) solid[(size_t)r * nx + c] = g.pix[(size_t)(r / ampY) * g.w + c / ampX] >= 150;
}

// 3D structure: text lines "x,y,z" of solid voxels (readImage3D, Keff3D/keff3D.h:237-280).
inline bool structure_from_csv(const std::string& path, int W, int H, int D, std::vector<uint8_t>& solid, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  std::vector<char> buf;
  {
    char tmp[1 << 16];
    size_t r;
    while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
  }
  fclose(f);
  buf.push_back('\0');
  solid.assign((size_t)W * H * D, 0);
  // the reference reads with fscanf(" %d,%d,%d") until it fails: same grammar (white space, optional sign, digits,
  // commas), same stop at the first thing that is not a triple -- parsed in place (a 512^3 structure has 10^8 lines)
  const char* p = buf.data();
  auto skip_ws = [&]() { while (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t' || *p == '\f' || *p == '\v') p++; };
  auto integer = [&](int& v) -> bool {
    skip_ws();
    bool neg = false;
    if (*p == '-' || *p == '+') neg = *p++ == '-';
    if (*p < '0' || *p > '9') return false;
    long t = 0;
    while (*p >= '0' && *p <= '9') {
      if (t < (1L << 40)) t = t * 10 + (*p - '0');
      p++;
    }
    v = (int)(neg ? -t : t);
    return true;
  };
  size_t bad = 0;
  int x, y, z;
  while (true) {
    if (!integer(x) || *p != ',') break;
    p++;
    if (!integer(y) || *p != ',') break;
    p++;
    if (!integer(z)) break;
    if (x < 0 || y < 0 || z < 0 || x >= W || y >= H || z >= D) { bad++; continue; }
    solid[((size_t)z * H + y) * W + x] = 1;
  }
  if (bad) { err = "coordinates outside Width/Height/Depth in " + path; return false; }
  return true;
}

// createOutput rows.  2D: Keff2D/keff2D.h:245-246; 3D: Keff3D/keff3D.h:308-309 (adds MeshIncreaseZ).
// nElements is an int in the reference (it overflows beyond 2^31 cells); here it is printed from a long, which gives
// the same text wherever the reference's is right.
inline bool append_row(const Options& o, bool three_d, double keff, double Q1, double Q2, long iters, const char* input,
                       long nElements, double porosity, double seconds, int nCores) {
  FILE* f = fopen(o.outputFilename.c_str(), "a+");
  if (!f) return false;
  if (three_d)
    fprintf(f, "%.10lf,%f,%f,%ld,%.10lf,%s,%ld,%d,%d,%d,%f,%f,%f,%f,%f,%f,%d\n", keff, Q1, Q2, iters, o.ConvergeCriteria,
            input, nElements, o.MeshIncreaseX, o.MeshIncreaseY, o.MeshIncreaseZ, porosity, o.TCsolid, o.TCfluid,
            o.TempLeft, o.TempRight, seconds, nCores);
  else
    fprintf(f, "%.10lf,%f,%f,%ld,%.10lf,%s,%ld,%d,%d,%f,%f,%f,%f,%f,%f,%d\n", keff, Q1, Q2, iters, o.ConvergeCriteria, input,
            nElements, o.MeshIncreaseX, o.MeshIncreaseY, porosity, o.TCsolid, o.TCfluid, o.TempLeft, o.TempRight, seconds,
            nCores);
  fclose(f);
  return true;
}

// printTMAP: 2D "x,y,T" (keff2D.h:349-381), 3D "x,y,z,T" (keff3D.h:397-437)
inline bool write_tmap(const std::string& name, const double* T, int nx, int ny, int nz, bool three_d) {
  FILE* f = fopen(name.c_str(), "w+");
  if (!f) return false;
  fprintf(f, three_d ? "x,y,z,T\n" : "x,y,T\n");
  for (int k = 0; k < nz; k++)
    for (int i = 0; i < ny; i++)
      for (int j = 0; j < nx; j++) {
        const double v = T[((size_t)k * ny + i) * nx + j];
        if (three_d) fprintf(f, "%d,%d,%d,%f\n", j, i, k, v);
        else fprintf(f, "%d,%d,%f\n", j, i, v);
      }
  fclose(f);
  return true;
}

// printQMAP (keff2D.h:427-479): x-direction face fluxes, nx+1 values per row: the wall fluxes at j = 0 and j = nx,
// harmonic-mean face fluxes in between -- evaluated on the GPU (keffb200_flux_field); 2D "x,y,Q", and the 3D map the
// reference announces but never wrote (keff3D.cpp:179-181) as "x,y,z,Q".
inline bool write_qmap(const std::string& name, const double* Qx, int nx, int ny, int nz, bool three_d) {
  FILE* f = fopen(name.c_str(), "w+");
  if (!f) return false;
  fprintf(f, three_d ? "x,y,z,Q\n" : "x,y,Q\n");
  for (int k = 0; k < nz; k++)
    for (int i = 0; i < ny; i++)
      for (int j = 0; j <= nx; j++) {
        const double q = Qx[((size_t)k * ny + i) * (nx + 1) + j];
        if (three_d) fprintf(f, "%d,%d,%d,%f\n", j, i, k, q);
        else fprintf(f, "%d,%d,%f\n", j, i, q);
      }
  fclose(f);
  return true;
}

// 3D structure replicated by an integer factor per direction (the mesh-convergence mode, Keff3D/keff3D.cpp:226-243)
inline void amplify3d(const std::vector<uint8_t>& s, int W, int H, int D, int ax, int ay, int az, std::vector<uint8_t>& out) {
  const int nx = W * ax, ny = H * ay, nz = D * az;
  out.resize((size_t)nx * ny * nz);
  for (int k = 0; k < nz; k++)
    for (int j = 0; j < ny; j++) {
      const uint8_t* src = s.data() + ((size_t)(k / az) * H + j / ay) * W;
      uint8_t* dst = out.data() + ((size_t)k * ny + j) * nx;
      for (int i = 0; i < nx; i++) dst[i] = src[i / ax];
    }
}

inline int die(const char* what) {
  fprintf(stderr, "keffb200: %s: %s\n", what, keffb200_last_error());
  return 1;
}

}  // namespace kbhost
