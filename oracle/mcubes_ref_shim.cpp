// TEST INFRASTRUCTURE ONLY.  Thin C entry point around the REFERENCE's own marching cubes
// (/root/reference/convocc/src/utils/libmcubes/marchingcubes.{h,cpp}, compiled from where they lie by oracle/Makefile into
// oracle/_ref/libmcubes_ref.so): the call libmcubes.marching_cubes(volume, isovalue) makes (pywrapper.cpp:90-107 --
// lower = 0, upper = shape - 1, the volume read through an (int, int, int) functor, i.e. with the template's x + 0.5
// coordinates truncated back to grid indices) without the numpy / Cython wrapping.  No reference source is copied here.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "marchingcubes.h"

namespace {
struct DenseVolume {
  const double* p;
  long ny, nz;
  double operator()(int x, int y, int z) const { return p[((long)x * ny + y) * nz + z]; }
};
}  // namespace

extern "C" {
// vol: (nx, ny, nz) C-order doubles.  Outputs are malloc'ed (free with mcref_free): vertices 3 doubles each, triangles 3 indices each.
int mcref_marching_cubes(const double* vol, int nx, int ny, int nz, double isovalue, double** vertices, unsigned long long* n_vertices,
                         unsigned long long** triangles, unsigned long long* n_triangles) {
  double lower[3] = {0, 0, 0};
  double upper[3] = {(double)(nx - 1), (double)(ny - 1), (double)(nz - 1)};
  std::vector<double> v;
  std::vector<size_t> t;
  mc::marching_cubes<double>(lower, upper, nx, ny, nz, DenseVolume{vol, ny, nz}, isovalue, v, t);
  *n_vertices = v.size() / 3;
  *n_triangles = t.size() / 3;
  *vertices = (double*)malloc(sizeof(double) * (v.size() + 1));
  *triangles = (unsigned long long*)malloc(sizeof(unsigned long long) * (t.size() + 1));
  if (!*vertices || !*triangles) return 1;
  memcpy(*vertices, v.data(), sizeof(double) * v.size());
  for (size_t i = 0; i < t.size(); ++i) (*triangles)[i] = t[i];
  return 0;
}
void mcref_free(void* p) { free(p); }
}
