// one_round_native.cpp — driver of the free oneRound() (reference src/my_utilities.cpp:263-315: fresh solver, kernel
// threshold 100, outliers KEPT with the sqrt(thr/chi) weight, <= 50 rounds, stop at 5 % relative improvement; fewer
// than 10 correspondences -> the input pose comes back untouched).  No executable of the reference calls it; this
// binary exists so that the third driver of the solver is exercised through the host mirror.
//   one_round_native <in.bin> <out.bin>
// in.bin : int32 n_world, n_image, n_pairs; float pose[12]; float world[3*n_world]; float image[2*n_image];
//          int32 pairs[2*n_pairs]      out.bin: float pose[12]
#include <cstdio>
#include <iostream>
#include <vector>

#include "cam.h"
#include "my_utilities.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: one_round_native <in.bin> <out.bin>" << std::endl;
    return 2;
  }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[3];
  vo::Iso3f pose;
  if (std::fread(hdr, 4, 3, f) != 3 || std::fread(pose.m, 4, 12, f) != 12) return 2;
  pr::Vector3fVector world(hdr[0]);
  pr::Vector2fVector image(hdr[1]);
  pr::IntPairVector pairs(hdr[2]);
  if (hdr[0] && std::fread(world.data(), 12, hdr[0], f) != (size_t)hdr[0]) return 2;
  if (hdr[1] && std::fread(image.data(), 8, hdr[1], f) != (size_t)hdr[1]) return 2;
  if (hdr[2] && std::fread(pairs.data(), 8, hdr[2], f) != (size_t)hdr[2]) return 2;
  std::fclose(f);
  Cam cam;
  pr::Camera picp_cam(cam.getHeight(), cam.getWidth(), cam.getEigenCamera(), vo::Iso3f::Identity());
  const vo::Iso3f out = oneRound(pose, picp_cam, world, image, pairs);
  f = std::fopen(argv[2], "wb");
  if (!f || std::fwrite(out.m, 4, 12, f) != 12) return 2;
  std::fclose(f);
  return 0;
}
