// jpeg2raw in.jpg out.raw -- test helper: decode with jpeg_gray.hpp, write "w h channels\n" + raw bytes.
#include "jpeg_gray.hpp"
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  kbhost::GrayImage g;
  if (!kbhost::load_jpeg_gray(argv[1], g)) { fprintf(stderr, "%s\n", g.err.c_str()); return 1; }
  FILE* f = fopen(argv[2], "wb");
  fprintf(f, "%d %d %d\n", g.w, g.h, g.channels);
  fwrite(g.pix.data(), 1, g.pix.size(), f);
  fclose(f);
  return 0;
}
