// stbgray in out.raw -- TEST INFRASTRUCTURE: the reference's own image decoder (the stb_image.h it vendors,
// Keff2D/keff2D.h:300: stbi_load(name, &w, &h, &channels, 1)), compiled from /root/reference where it lies by
// oracle/Makefile into oracle/_ref/stbgray.  Writes "w h channels\n" + the raw 8-bit channel, the format of the
// host programs' jpeg2raw, so tests/test_host_cpp.py can require the two decoders to agree bit for bit.
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <cstdio>
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  int w = 0, h = 0, ch = 0;
  unsigned char* p = stbi_load(argv[1], &w, &h, &ch, 1);
  if (!p) { fprintf(stderr, "%s\n", stbi_failure_reason()); return 1; }
  FILE* f = fopen(argv[2], "wb");
  fprintf(f, "%d %d %d\n", w, h, ch);
  fwrite(p, 1, (size_t)w * h, f);
  fclose(f);
  return 0;
}
