// jpeg2raw in.(jpg|png|bmp|pgm) out.raw -- test helper: decode with image_gray.hpp, write "w h channels\n" + raw bytes.
#include "image_gray.hpp"
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  kbhost::GrayImage g;
  if (!kbhost::load_image_gray(argv[1], g)) { fprintf(stderr, "%s\n", g.err.c_str()); return 1; }
  FILE* f = fopen(argv[2], "wb");
  fprintf(f, "%d %d %d\n", g.w, g.h, g.channels);
  fwrite(g.pix.data(), 1, g.pix.size(), f);
  fclose(f);
  return 0;
}
