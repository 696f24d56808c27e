// Occupancy post-step on the GPU: marching cubes on the decoded logit grid, in place of the reference's
// eval_points -> .cpu() -> libmcubes round trip (convocc/src/conv_onet/generation.py:123-186,
// convocc/src/utils/libmcubes/marchingcubes.h:23-193).
//
// The reference walks the cells sequentially (x outermost, z innermost), appends a vertex the first time an edge is met
// (each cell owns the three edges at its far corner -- table edges 6, 5, 10 -- plus, on the low faces of the volume, the
// edges nobody else owns) and looks shared vertices up in a two-slab index cache.  The SAME mesh -- same vertices, same
// duplicates on the volume faces, same ORDER of vertices and triangles, bit-identical float64 coordinates -- comes out of three
// data-parallel passes:
//   1. classify : one thread per cell -> case index, the set of edges the cell creates, vertex / triangle-corner counts
//   2. scan     : exclusive prefix sum of the counts in cell order (= the reference's append positions)
//   3. emit     : one thread per cell -> its vertices at its scanned offset (in the reference's per-cell creation order),
//                 its triangles with shared vertices resolved through the owner cell's offset + rank
// Vertex arithmetic is float64 with explicit round-to-nearest operations in the reference's order (no FMA contraction).
// HBM-bound integer / index work: 8 value loads per cell from a 128^3 grid (L2-resident), 12 B of bookkeeping per cell.
#include "common.cuh"
#include "mc_tables.cuh"

namespace ddmi {
namespace mcubes {

__constant__ unsigned long long c_tri[256] = DDMI_MC_TRIANGLES_INIT;

struct Dims {
  int nx, ny, nz;      // the caller's grid
  int pad;             // layers of pad_value around it (the reference pads by 1 to close the mesh)
  int cx, cy, cz;      // cells of the padded volume
};

__device__ __forceinline__ double value_at(const float* __restrict__ grid, const Dims& d, double pad_value, int i, int j, int k) {
  i -= d.pad; j -= d.pad; k -= d.pad;
  if ((unsigned)i >= (unsigned)d.nx || (unsigned)j >= (unsigned)d.ny || (unsigned)k >= (unsigned)d.nz) return pad_value;
  return (double)__ldg(grid + ((size_t)i * d.ny + j) * d.nz + k);
}
// corner m of cell (i, j, k): marchingcubes.h:59-63
__device__ __forceinline__ void load_cell(const float* __restrict__ grid, const Dims& d, double pad_value, int i, int j, int k,
                                          double (&v)[8]) {
  v[0] = value_at(grid, d, pad_value, i, j, k);
  v[1] = value_at(grid, d, pad_value, i + 1, j, k);
  v[2] = value_at(grid, d, pad_value, i + 1, j + 1, k);
  v[3] = value_at(grid, d, pad_value, i, j + 1, k);
  v[4] = value_at(grid, d, pad_value, i, j, k + 1);
  v[5] = value_at(grid, d, pad_value, i + 1, j, k + 1);
  v[6] = value_at(grid, d, pad_value, i + 1, j + 1, k + 1);
  v[7] = value_at(grid, d, pad_value, i, j + 1, k + 1);
}
__device__ __forceinline__ uint32_t case_index(const double (&v)[8], double iso) {
  uint32_t ci = 0;
#pragma unroll
  for (int m = 0; m < 8; ++m) ci |= (v[m] <= iso ? 1u : 0u) << m;   // marchingcubes.h:65-68
  return ci;
}
// edges a case uses (= the reference's edge_table entry) and its number of triangle corners
__device__ __forceinline__ void case_edges(uint32_t ci, uint32_t& edges, uint32_t& corners) {
  unsigned long long w = c_tri[ci];
  edges = 0;
  corners = 0;
#pragma unroll
  for (int m = 0; m < 15; ++m) {
    const uint32_t e = (uint32_t)(w >> (4 * m)) & 0xF;
    if (e != 0xF) {
      edges |= 1u << e;
      ++corners;
    }
  }
}
// edges whose vertex THIS cell appends (marchingcubes.h:73-182): 6, 5, 10 always; the others only on the low faces
__device__ __forceinline__ uint32_t created_edges(uint32_t edges, int i, int j, int k) {
  uint32_t own = 0x040 | 0x020 | 0x400;
  if (j == 0 || k == 0) own |= 0x001;
  if (k == 0) own |= 0x002 | 0x004;
  if (i == 0 || k == 0) own |= 0x008;
  if (j == 0) own |= 0x010 | 0x200;
  if (i == 0) own |= 0x080 | 0x800;
  if (i == 0 || j == 0) own |= 0x100;
  return edges & own;
}

// ---- pass 1
__global__ void __launch_bounds__(256)
classify_kernel(const float* __restrict__ grid, Dims d, double pad_value, double iso, long long ncells,
                uint32_t* __restrict__ info, unsigned long long* __restrict__ counts) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int k = (int)(c % d.cz), j = (int)((c / d.cz) % d.cy), i = (int)(c / ((long long)d.cz * d.cy));
  double v[8];
  load_cell(grid, d, pad_value, i, j, k, v);
  const uint32_t ci = case_index(v, iso);
  uint32_t edges, corners;
  case_edges(ci, edges, corners);
  const uint32_t cr = created_edges(edges, i, j, k);
  info[c] = cr | (ci << 12);
  counts[c] = (unsigned long long)__popc(cr) | ((unsigned long long)corners << 32);
}

// ---- pass 2: exclusive scan of packed (vertex count | corner count << 32) words, 2048 per block
constexpr int SCAN_T = 256, SCAN_PER = 8, SCAN_BLOCK = SCAN_T * SCAN_PER;
__device__ __forceinline__ unsigned long long block_exclusive(unsigned long long x, unsigned long long* sh, unsigned long long& total) {
  // exclusive scan of one value per thread across the block (both 32-bit halves stay below 2^32 by the host-side check)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  unsigned long long woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_T / 32; ++w) {
    if (w < warp) woff += sh[w];
    tot += sh[w];
  }
  __syncthreads();
  total = tot;
  return woff + inc - x;
}
__global__ void __launch_bounds__(SCAN_T)
scan_local_kernel(unsigned long long* __restrict__ data, long long n, unsigned long long* __restrict__ block_sums) {
  __shared__ unsigned long long sh[SCAN_T / 32];
  const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_PER;
  unsigned long long v[SCAN_PER], s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_PER; ++e) {
    v[e] = base + e < n ? data[base + e] : 0ull;
    s += v[e];
  }
  unsigned long long total;
  unsigned long long off = block_exclusive(s, sh, total);
#pragma unroll
  for (int e = 0; e < SCAN_PER; ++e) {
    if (base + e < n) data[base + e] = off;
    off += v[e];
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_T)
scan_sums_kernel(unsigned long long* __restrict__ block_sums, int nblocks, unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long sh[SCAN_T / 32];
  unsigned long long carry = 0;
  for (int b0 = 0; b0 < nblocks; b0 += SCAN_T) {
    const int b = b0 + threadIdx.x;
    const unsigned long long x = b < nblocks ? block_sums[b] : 0ull;
    unsigned long long total;
    const unsigned long long ex = block_exclusive(x, sh, total);
    if (b < nblocks) block_sums[b] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) {
    totals[0] = carry & 0xFFFFFFFFull;   // vertices
    totals[1] = carry >> 32;             // triangle corners
  }
}
__global__ void __launch_bounds__(SCAN_T)
scan_add_kernel(unsigned long long* __restrict__ data, long long n, const unsigned long long* __restrict__ block_sums) {
  const unsigned long long add = block_sums[blockIdx.x];
  const long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_PER;
#pragma unroll
  for (int e = 0; e < SCAN_PER; ++e)
    if (base + e < n) data[base + e] += add;
}

// ---- pass 3
struct Affine {     // Generator3D.extract_mesh's vertex post-processing (generation.py:170-186), applied per component in its order:
  int on;           // v = mul * ((((v - sub0) - sub1) / div[axis]) - sub2)
  double sub0, sub1, div[3], sub2, mul;
};
// mc_isovalue_interpolation (marchingcubes.cpp:290-297), round-to-nearest operations in the reference's order
__device__ __forceinline__ double interp(double iso, double f1, double f2, double x1, double x2) {
  if (f2 == f1) return __ddiv_rn(__dadd_rn(x2, x1), 2.0);
  return __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(iso, f1)), __dsub_rn(f2, f1)), x1);
}
__device__ __forceinline__ void put_vertex(double* __restrict__ out, unsigned long long idx, double x, double y, double z, const Affine& a) {
  double p[3] = {x, y, z};
  if (a.on) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      p[c] = __dmul_rn(a.mul, __dsub_rn(__ddiv_rn(__dsub_rn(__dsub_rn(p[c], a.sub0), a.sub1), a.div[c]), a.sub2));
  }
  out[3 * idx] = p[0];
  out[3 * idx + 1] = p[1];
  out[3 * idx + 2] = p[2];
}
// rank of the owner cell's slot vertex (0: edge 6, 1: edge 5, 2: edge 10) in its creation order 6, 5, 10, ...
__device__ __forceinline__ uint32_t slot_rank(uint32_t created, int slot) {
  const uint32_t has6 = (created >> 6) & 1, has5 = (created >> 5) & 1;
  return slot == 0 ? 0u : (slot == 1 ? has6 : has6 + has5);
}

__global__ void __launch_bounds__(256)
emit_kernel(const float* __restrict__ grid, Dims d, double pad_value, double iso, long long ncells,
            const uint32_t* __restrict__ info, const unsigned long long* __restrict__ offs, Affine aff,
            double* __restrict__ vertices, long long* __restrict__ triangles) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint32_t inf = info[c], ci = inf >> 12, cr = inf & 0xFFF;
  if (ci == 0 || ci == 255) return;
  const int k = (int)(c % d.cz), j = (int)((c / d.cz) % d.cy), i = (int)(c / ((long long)d.cz * d.cy));
  double v[8];
  load_cell(grid, d, pad_value, i, j, k, v);
  const unsigned long long off = offs[c];
  unsigned long long vnext = off & 0xFFFFFFFFull;
  const unsigned long long tbase = off >> 32;
  const double x = i + 0.5, y = j + 0.5, z = k + 0.5, x1 = i + 1.5, y1 = j + 1.5, z1 = k + 1.5;
  const unsigned long long w = c_tri[ci];
  uint32_t edges = 0;
#pragma unroll
  for (int m = 0; m < 15; ++m) {
    const uint32_t e = (uint32_t)(w >> (4 * m)) & 0xF;
    if (e != 0xF) edges |= 1u << e;
  }
  long long idx[12];
  const long long sz = d.cz, sy = (long long)d.cz * d.cy;
  // vertex of a shared edge: owner cell `oc`, slot s
  auto shared = [&](long long oc, int s) { return (long long)((offs[oc] & 0xFFFFFFFFull) + slot_rank(info[oc] & 0xFFF, s)); };
  // the reference's order of appends within a cell (marchingcubes.h:73-182)
  if (edges & 0x040) { idx[6] = (long long)vnext; put_vertex(vertices, vnext++, interp(iso, v[6], v[7], x1, x), y1, z1, aff); }
  if (edges & 0x020) { idx[5] = (long long)vnext; put_vertex(vertices, vnext++, x1, interp(iso, v[5], v[6], y, y1), z1, aff); }
  if (edges & 0x400) { idx[10] = (long long)vnext; put_vertex(vertices, vnext++, x1, y1, interp(iso, v[2], v[6], z, z1), aff); }
  if (edges & 0x001) {
    if (cr & 0x001) { idx[0] = (long long)vnext; put_vertex(vertices, vnext++, interp(iso, v[0], v[1], x, x1), y, z, aff); }
    else idx[0] = shared(c - sy * 0 - sz - 1, 0);                                   // cell (i, j - 1, k - 1), edge 6
  }
  if (edges & 0x002) {
    if (cr & 0x002) { idx[1] = (long long)vnext; put_vertex(vertices, vnext++, x1, interp(iso, v[1], v[2], y, y1), z, aff); }
    else idx[1] = shared(c - 1, 1);                                                 // (i, j, k - 1), edge 5
  }
  if (edges & 0x004) {
    if (cr & 0x004) { idx[2] = (long long)vnext; put_vertex(vertices, vnext++, interp(iso, v[2], v[3], x1, x), y1, z, aff); }
    else idx[2] = shared(c - 1, 0);                                                 // (i, j, k - 1), edge 6
  }
  if (edges & 0x008) {
    if (cr & 0x008) { idx[3] = (long long)vnext; put_vertex(vertices, vnext++, x, interp(iso, v[3], v[0], y1, y), z, aff); }
    else idx[3] = shared(c - sy - 1, 1);                                            // (i - 1, j, k - 1), edge 5
  }
  if (edges & 0x010) {
    if (cr & 0x010) { idx[4] = (long long)vnext; put_vertex(vertices, vnext++, interp(iso, v[4], v[5], x, x1), y, z1, aff); }
    else idx[4] = shared(c - sz, 0);                                                // (i, j - 1, k), edge 6
  }
  if (edges & 0x080) {
    if (cr & 0x080) { idx[7] = (long long)vnext; put_vertex(vertices, vnext++, x, interp(iso, v[7], v[4], y1, y), z1, aff); }
    else idx[7] = shared(c - sy, 1);                                                // (i - 1, j, k), edge 5
  }
  if (edges & 0x100) {
    if (cr & 0x100) { idx[8] = (long long)vnext; put_vertex(vertices, vnext++, x, y, interp(iso, v[0], v[4], z, z1), aff); }
    else idx[8] = shared(c - sy - sz, 2);                                           // (i - 1, j - 1, k), edge 10
  }
  if (edges & 0x200) {
    if (cr & 0x200) { idx[9] = (long long)vnext; put_vertex(vertices, vnext++, x1, y, interp(iso, v[1], v[5], z, z1), aff); }
    else idx[9] = shared(c - sz, 2);                                                // (i, j - 1, k), edge 10
  }
  if (edges & 0x800) {
    if (cr & 0x800) { idx[11] = (long long)vnext; put_vertex(vertices, vnext++, x, y1, interp(iso, v[3], v[7], z, z1), aff); }
    else idx[11] = shared(c - sy, 2);                                               // (i - 1, j, k), edge 10
  }
#pragma unroll
  for (int m = 0; m < 15; ++m) {
    const uint32_t e = (uint32_t)(w >> (4 * m)) & 0xF;
    if (e != 0xF) triangles[tbase + m] = idx[e];
  }
}

}  // namespace mcubes

static bool mc_dims(int nx, int ny, int nz, int pad, mcubes::Dims* d, long long* ncells) {
  if (nx < 1 || ny < 1 || nz < 1 || pad < 0 || pad > 1) return false;
  d->nx = nx; d->ny = ny; d->nz = nz; d->pad = pad;
  d->cx = nx + 2 * pad - 1; d->cy = ny + 2 * pad - 1; d->cz = nz + 2 * pad - 1;
  if (d->cx < 1 || d->cy < 1 || d->cz < 1) return false;
  *ncells = (long long)d->cx * d->cy * d->cz;
  return *ncells < (1ll << 28);     // at most 12 vertices / 15 corners per cell: both scanned counts stay below 2^32
}
// workspace: info (u32 per cell) | counts / offsets (u64 per cell) | block sums (u64 per 2048 cells)
static void mc_layout(long long ncells, size_t* off_counts, size_t* off_sums, size_t* total) {
  const size_t a = ((size_t)ncells * 4 + 255) & ~(size_t)255;
  const size_t b = (size_t)ncells * 8;
  const size_t nb = (size_t)((ncells + mcubes::SCAN_BLOCK - 1) / mcubes::SCAN_BLOCK);
  *off_counts = a;
  *off_sums = a + b;
  *total = a + b + nb * 8 + 256;
}

int mcubes_workspace_bytes(int nx, int ny, int nz, int pad, unsigned long long* bytes) {
  mcubes::Dims d;
  long long ncells;
  DDMI_REQUIRE(mc_dims(nx, ny, nz, pad, &d, &ncells), "marching cubes: grid %d x %d x %d (pad %d) is empty or has 2^28 cells or more", nx, ny, nz, pad);
  size_t oc, os, tot;
  mc_layout(ncells, &oc, &os, &tot);
  *bytes = tot;
  return DDMI_OK;
}

int launch_mcubes_count(const float* grid, int nx, int ny, int nz, int pad, double pad_value, double iso, void* workspace,
                        unsigned long long workspace_bytes, unsigned long long* totals_dev, cudaStream_t st) {
  using namespace mcubes;
  Dims d;
  long long ncells;
  DDMI_REQUIRE(mc_dims(nx, ny, nz, pad, &d, &ncells), "marching cubes: grid %d x %d x %d (pad %d) is empty or has 2^28 cells or more", nx, ny, nz, pad);
  size_t oc, os, tot;
  mc_layout(ncells, &oc, &os, &tot);
  DDMI_REQUIRE(workspace_bytes >= tot, "marching cubes workspace is %llu bytes, %zu needed", workspace_bytes, tot);
  uint32_t* info = (uint32_t*)workspace;
  unsigned long long* counts = (unsigned long long*)((char*)workspace + oc);
  unsigned long long* sums = (unsigned long long*)((char*)workspace + os);
  const int nblocks = (int)((ncells + SCAN_BLOCK - 1) / SCAN_BLOCK);
  classify_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(grid, d, pad_value, iso, ncells, info, counts);
  scan_local_kernel<<<nblocks, SCAN_T, 0, st>>>(counts, ncells, sums);
  scan_sums_kernel<<<1, SCAN_T, 0, st>>>(sums, nblocks, totals_dev);
  scan_add_kernel<<<nblocks, SCAN_T, 0, st>>>(counts, ncells, sums);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

int launch_mcubes_emit(const float* grid, int nx, int ny, int nz, int pad, double pad_value, double iso, const void* workspace,
                       const double* affine, double* vertices, long long* triangles, cudaStream_t st) {
  using namespace mcubes;
  Dims d;
  long long ncells;
  DDMI_REQUIRE(mc_dims(nx, ny, nz, pad, &d, &ncells), "marching cubes: bad grid %d x %d x %d (pad %d)", nx, ny, nz, pad);
  size_t oc, os, tot;
  mc_layout(ncells, &oc, &os, &tot);
  Affine a = {};
  if (affine) {
    a.on = 1;
    a.sub0 = affine[0]; a.sub1 = affine[1]; a.div[0] = affine[2]; a.div[1] = affine[3]; a.div[2] = affine[4];
    a.sub2 = affine[5]; a.mul = affine[6];
  }
  emit_kernel<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(grid, d, pad_value, iso, ncells, (const uint32_t*)workspace,
                                                               (const unsigned long long*)((const char*)workspace + oc), a,
                                                               vertices, triangles);
  DDMI_CUDA(cudaGetLastError());
  return DDMI_OK;
}

}  // namespace ddmi
