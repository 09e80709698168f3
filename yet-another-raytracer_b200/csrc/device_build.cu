// device_build.cu -- L4QBVH::new (reference qbvh.rs:251-361) on the GPU: the SAME tree, byte for byte, as
// host_qbvh.cpp builds (tests/test_gpu_build.py compares the two), without the host-side sorts.
//
// What makes a level-synchronous build possible: the SHAPE of the reference tree -- which index ranges become
// nodes, halves and leaves, the post-order node numbering, the leaf encodings -- is a function of the triangle
// count alone (ranges are always cut at len/2, qbvh.rs:281-302).  Only three things depend on the geometry:
//   * the permutation: every range that is split is first sorted by triangle-AABB centroid along the axis
//     with the largest centroid extent (split, qbvh.rs:636-693), ties by original index (host_qbvh.cpp);
//   * the three split axes stored per node;
//   * the boxes.
// So the host lays out the shape (O(nodes), no geometry), and per BINARY level the device does
//   bounds of every segment (atomics on order-preserving u64 images of the f64 centroids)
//   -> axis per segment -> per-triangle keys -> two stable radix sorts (cub::DeviceRadixSort):
//      by centroid key (ties keep ascending triangle index), then by segment start (puts every triangle
//      back into its own segment; triangles of finished ranges keep their position),
// then gathers the triangle records and fills the boxes bottom-up, one launch per node height.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "device_build.h"

namespace yart {
namespace {

#define BUILD_TRY(expr)                                                    \
  do {                                                                     \
    cudaError_t e__ = (expr);                                              \
    if (e__ != cudaSuccess) {                                              \
      err = std::string(#expr) + ": " + cudaGetErrorString(e__);           \
      return false;                                                        \
    }                                                                      \
  } while (0)

struct Seg {          // a range that split() sorts at one binary level
  uint32_t lo, hi;
  uint32_t node;      // the node that stores this split's axis ...
  uint32_t role;      // ... as top (0), left (1) or right (2) axis
};
struct TopoNode {     // one 4-wide node: child ranges and ids (shape only)
  uint32_t lo[4], hi[4], id[4];
};

// The shape of construct() (qbvh.rs:253-347): same recursion as host_qbvh.cpp without touching geometry.
struct Topology {
  std::vector<std::vector<Seg>> levels;          // [binary depth] -> segments in ascending lo
  std::vector<TopoNode> nodes;                   // post-order, root last
  std::vector<std::vector<uint32_t>> by_height;  // node ids of each height (1 = all children are leaves)
  uint32_t n_leaves = 0;

  uint32_t construct(uint32_t lo, uint32_t hi, uint32_t depth, uint32_t& height) {
    const uint32_t n = hi - lo;
    height = 0;
    if (n == 0) return 0xFFFFFFFFu;
    if (n <= 4) {
      n_leaves++;
      return lo | (1u << 31) | (n << 27); // qbvh.rs:270
    }
    if (levels.size() < (size_t)depth + 2) levels.resize((size_t)depth + 2);
    const uint32_t mid = lo + n / 2, lmid = lo + (mid - lo) / 2, rmid = mid + (hi - mid) / 2;
    const size_t s_top = levels[depth].size();
    levels[depth].push_back({lo, hi, 0, 0});
    const size_t s_half = levels[depth + 1].size();
    levels[depth + 1].push_back({lo, mid, 0, 1});
    levels[depth + 1].push_back({mid, hi, 0, 2});
    TopoNode nd;
    const uint32_t cl[4] = {lo, lmid, mid, rmid}, chi[4] = {lmid, mid, rmid, hi};
    uint32_t hmax = 0;
    for (int k = 0; k < 4; ++k) {
      uint32_t h;
      nd.lo[k] = cl[k];
      nd.hi[k] = chi[k];
      nd.id[k] = construct(cl[k], chi[k], depth + 2, h);
      hmax = h > hmax ? h : hmax;
    }
    const uint32_t id = (uint32_t)nodes.size();
    nodes.push_back(nd);
    levels[depth][s_top].node = id;
    levels[depth + 1][s_half].node = id;
    levels[depth + 1][s_half + 1].node = id;
    height = hmax + 1;
    if (by_height.size() < (size_t)height + 1) by_height.resize((size_t)height + 1);
    by_height[height].push_back(id);
    return id;
  }
};

// order-preserving u64 image of a finite f64 (and of +-inf); -0.0 is folded into +0.0 first because the
// host comparator treats them as equal
__device__ __forceinline__ unsigned long long key_of(double c) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(c + 0.0);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

// triangle AABB (triangle.rs:20-45) and its centroid (hittable.rs:12-22), exactly as host_qbvh.cpp::prepare
__global__ void k_prepare(const float* __restrict__ pos, uint32_t n, float* __restrict__ bmin, float* __restrict__ bmax,
                          double* __restrict__ cen, uint32_t* __restrict__ perm) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pos + (size_t)i * 9;
  for (int a = 0; a < 3; ++a) {
    const float lo = fminf(fminf(p[a], p[3 + a]), p[6 + a]);
    const float hi = fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]);
    bmin[(size_t)i * 3 + a] = lo;
    bmax[(size_t)i * 3 + a] = hi;
    cen[(size_t)i * 3 + a] = ((double)hi + (double)lo) / 2.0;
  }
  perm[i] = i;
}

// last segment whose lo <= pos, or -1
__device__ __forceinline__ int find_seg(const Seg* segs, uint32_t n_segs, uint32_t pos) {
  int a = 0, b = (int)n_segs - 1, r = -1;
  while (a <= b) {
    const int m = (a + b) >> 1;
    if (segs[m].lo <= pos) { r = m; a = m + 1; } else { b = m - 1; }
  }
  if (r >= 0 && pos >= segs[r].hi) r = -1;
  return r;
}

__global__ void k_init_bounds(unsigned long long* bounds, uint32_t n_segs) { // [seg][6]: min xyz, max xyz
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_segs * 6u) bounds[i] = ((i % 6u) < 3u) ? ~0ull : 0ull;
}

// centroid bounds of every segment of this level (split, qbvh.rs:640-665)
__global__ void k_seg_bounds(const Seg* __restrict__ segs, uint32_t n_segs, const uint32_t* __restrict__ perm,
                             const double* __restrict__ cen, uint32_t n, int* __restrict__ seg_of_pos,
                             unsigned long long* __restrict__ bounds) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  int s = -1;
  unsigned long long k[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  if (pos < n) {
    s = find_seg(segs, n_segs, pos);
    seg_of_pos[pos] = s;
    if (s >= 0) {
      const double* c = cen + (size_t)perm[pos] * 3;
      for (int a = 0; a < 3; ++a) k[a] = k[3 + a] = key_of(c[a]);
    }
  }
  // one atomic per warp when the whole warp sits in one segment (the big segments of the upper levels)
  const int s0 = __shfl_sync(0xffffffffu, s, 0);
  if (__all_sync(0xffffffffu, s == s0)) {
    if (s0 < 0) return;
    for (int off = 16; off > 0; off >>= 1)
      for (int a = 0; a < 3; ++a) {
        const unsigned long long lo = __shfl_xor_sync(0xffffffffu, k[a], off), hi = __shfl_xor_sync(0xffffffffu, k[3 + a], off);
        k[a] = lo < k[a] ? lo : k[a];
        k[3 + a] = hi > k[3 + a] ? hi : k[3 + a];
      }
    if ((threadIdx.x & 31) == 0)
      for (int a = 0; a < 3; ++a) {
        atomicMin(&bounds[(size_t)s0 * 6 + a], k[a]);
        atomicMax(&bounds[(size_t)s0 * 6 + 3 + a], k[3 + a]);
      }
  } else if (s >= 0) {
    for (int a = 0; a < 3; ++a) {
      atomicMin(&bounds[(size_t)s * 6 + a], k[a]);
      atomicMax(&bounds[(size_t)s * 6 + 3 + a], k[3 + a]);
    }
  }
}

// the axis rule of split (qbvh.rs:667-676) and its place in the node (top | left<<2 | right<<4)
__global__ void k_seg_axis(const Seg* __restrict__ segs, uint32_t n_segs, const unsigned long long* __restrict__ bounds,
                           uint32_t* __restrict__ seg_axis, uint32_t* __restrict__ node_axes) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  const unsigned long long* b = bounds + (size_t)s * 6;
  const double ex = value_of(b[3]) - value_of(b[0]), ey = value_of(b[4]) - value_of(b[1]), ez = value_of(b[5]) - value_of(b[2]);
  uint32_t axis = 0;
  if (ey > ex) axis = 1;
  if (ez > fmax(ey, ex)) axis = 2;
  seg_axis[s] = axis;
  atomicOr(&node_axes[segs[s].node], axis << (2u * segs[s].role));
}

// Per TRIANGLE (index = original id): the centroid key along its segment's axis and the start of its segment.
// A triangle outside every segment of this level is in a finished range: it keeps its position.
__global__ void k_tri_keys(const Seg* __restrict__ segs, const int* __restrict__ seg_of_pos, const uint32_t* __restrict__ seg_axis,
                           const uint32_t* __restrict__ perm, const double* __restrict__ cen, uint32_t n,
                           unsigned long long* __restrict__ cen_key, uint32_t* __restrict__ place_key, uint32_t* __restrict__ iota) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const uint32_t tri = perm[pos];
  const int s = seg_of_pos[pos];
  if (s >= 0) {
    cen_key[tri] = key_of(cen[(size_t)tri * 3 + seg_axis[s]]);
    place_key[tri] = segs[s].lo;
  } else {
    cen_key[tri] = 0ull;
    place_key[tri] = pos;
  }
  iota[pos] = pos;
}

__global__ void k_gather_u32(const uint32_t* __restrict__ table, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = table[idx[i]];
}

// FlatTri / FlatTriShade in tree order (host_qbvh.cpp, end of build_qbvh)
__global__ void k_emit_tris(const uint32_t* __restrict__ perm, uint32_t n, const float* __restrict__ pos, const double* __restrict__ nrm,
                            const float* __restrict__ uv, FlatTri* __restrict__ tris, FlatTriShade* __restrict__ shade) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t src = perm[i];
  const float* p = pos + (size_t)src * 9;
  FlatTri t;
  for (int j = 0; j < 3; ++j) { t.v0[j] = p[j]; t.v1[j] = p[3 + j]; t.v2[j] = p[6 + j]; }
  t.orig = src;
  t.pad1 = t.pad2 = 0;
  tris[i] = t;
  FlatTriShade s;
  for (int k = 0; k < 3; ++k) {
    for (int j = 0; j < 3; ++j) s.n[k][j] = nrm[(size_t)src * 9 + k * 3 + j];
    for (int j = 0; j < 2; ++j) s.uv[k][j] = uv[(size_t)src * 6 + k * 2 + j];
  }
  shade[i] = s;
}

// boxes of the nodes of one height: a leaf child's box is the union of its triangles' boxes, a node child's box
// the union of that node's four child boxes (finished by an earlier launch).  fminf / fmaxf only select, so
// the result does not depend on the order of the union (qbvh.rs:261-279, 303-345).
__global__ void k_node_boxes(const uint32_t* __restrict__ ids, uint32_t n_ids, const TopoNode* __restrict__ topo,
                             const uint32_t* __restrict__ perm, const float* __restrict__ bmin, const float* __restrict__ bmax,
                             const uint32_t* __restrict__ node_axes, FlatNode* __restrict__ nodes) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_ids * 4u) return;
  const uint32_t id = ids[w >> 2], k = w & 3u;
  const TopoNode& t = topo[id];
  const uint32_t cid = t.id[k];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {FLT_MAX, FLT_MAX, FLT_MAX}; // absent child (host_qbvh.cpp)
  if (cid != 0xFFFFFFFFu) {
    mx[0] = mx[1] = mx[2] = -FLT_MAX;
    if (cid >> 31) {
      for (uint32_t i = t.lo[k]; i < t.hi[k]; ++i) {
        const uint32_t tri = perm[i];
        for (int a = 0; a < 3; ++a) {
          mn[a] = fminf(mn[a], bmin[(size_t)tri * 3 + a]);
          mx[a] = fmaxf(mx[a], bmax[(size_t)tri * 3 + a]);
        }
      }
    } else {
      const FlatNode& c = nodes[cid];
      for (int j = 0; j < 4; ++j) {
        if (c.child[j] == 0xFFFFFFFFu) continue;
        mn[0] = fminf(mn[0], c.min_x[j]); mx[0] = fmaxf(mx[0], c.max_x[j]);
        mn[1] = fminf(mn[1], c.min_y[j]); mx[1] = fmaxf(mx[1], c.max_y[j]);
        mn[2] = fminf(mn[2], c.min_z[j]); mx[2] = fmaxf(mx[2], c.max_z[j]);
      }
    }
  }
  FlatNode& o = nodes[id];
  // __fadd_rn(x, 0): -0.0 -> +0.0, as host_qbvh.cpp stores it (a min/max fold's choice between equal zeros
  // depends on the fold order)
  o.min_x[k] = __fadd_rn(mn[0], 0.0f); o.max_x[k] = __fadd_rn(mx[0], 0.0f);
  o.min_y[k] = __fadd_rn(mn[1], 0.0f); o.max_y[k] = __fadd_rn(mx[1], 0.0f);
  o.min_z[k] = __fadd_rn(mn[2], 0.0f); o.max_z[k] = __fadd_rn(mx[2], 0.0f);
  o.child[k] = cid;
  if (k == 0) {
    o.axes = node_axes[id];
    o.pad[0] = o.pad[1] = o.pad[2] = 0;
  }
}

// One device allocation carved into the work arrays (a cudaMalloc per array costs more than the whole
// build of a small mesh); freed on every exit path.
struct Scratch {
  size_t total = 0;
  unsigned char* base = nullptr;
  std::vector<std::pair<void**, size_t>> wants;
  ~Scratch() { cudaFree(base); }
  template <class T>
  void want(T** out, size_t count) {
    wants.push_back({reinterpret_cast<void**>(out), total});
    total += (((count ? count : 1) * sizeof(T)) + 255) & ~(size_t)255;
  }
  cudaError_t commit() {
    cudaError_t e = cudaMalloc((void**)&base, total ? total : 256);
    if (e != cudaSuccess) return e;
    for (auto& w : wants) *w.first = base + w.second;
    return cudaSuccess;
  }
};

inline uint32_t blocks_for(size_t n) { return (uint32_t)((n + 255) / 256); }

} // namespace

bool build_qbvh_device(cudaStream_t stream, const yart_trimesh& mesh, DeviceQbvh& out, std::string& err) {
  memset(&out, 0, sizeof(out));
  if (!mesh.positions || !mesh.normals || !mesh.uvs) {
    err = "trimesh has null arrays";
    return false;
  }
  if (mesh.n_tris <= 4) { // same refusals as the host builder (SURVEY A-17, qbvh.rs:270)
    err = "L4QBVH needs more than 4 triangles (the reference panics on such meshes, qbvh.rs:383-384)";
    return false;
  }
  if (mesh.n_tris >= (1u << 27)) {
    err = "L4QBVH child ids hold 27 bits of triangle index (qbvh.rs:270)";
    return false;
  }
  const uint32_t n = mesh.n_tris;

  // ---- shape (host, no geometry) ----
  Topology topo;
  uint32_t height = 0;
  const uint32_t root = topo.construct(0, n, 0, height);
  const uint32_t n_nodes = (uint32_t)topo.nodes.size();
  size_t max_segs = 1, total_segs = 0;
  for (const auto& l : topo.levels) {
    max_segs = l.size() > max_segs ? l.size() : max_segs;
    total_segs += l.size();
  }

  // ---- device buffers ----
  Scratch scratch;
  float *d_pos, *d_uv, *d_bmin, *d_bmax;
  double *d_nrm, *d_cen;
  uint32_t *d_perm, *d_perm2, *d_place, *d_place_sorted, *d_place_sorted2, *d_iota, *d_vals, *d_seg_axis, *d_node_axes, *d_ids;
  unsigned long long *d_cen_key, *d_cen_key2, *d_bounds;
  int* d_seg_of_pos;
  Seg* d_segs;
  TopoNode* d_topo;
  size_t tmp_a = 0, tmp_b = 0; // (sizing calls: no work is done, the pointers are not touched)
  d_cen_key = d_cen_key2 = nullptr;
  d_iota = d_vals = d_place_sorted = d_place_sorted2 = d_perm2 = nullptr;
  BUILD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_a, d_cen_key, d_cen_key2, d_iota, d_vals, (int)n, 0, 64, stream));
  BUILD_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_b, d_place_sorted, d_place_sorted2, d_vals, d_perm2, (int)n, 0, 32, stream));
  const size_t tmp_bytes = tmp_a > tmp_b ? tmp_a : tmp_b;
  unsigned char* d_tmp;
  scratch.want(&d_pos, (size_t)n * 9);
  scratch.want(&d_nrm, (size_t)n * 9);
  scratch.want(&d_uv, (size_t)n * 6);
  scratch.want(&d_bmin, (size_t)n * 3);
  scratch.want(&d_bmax, (size_t)n * 3);
  scratch.want(&d_cen, (size_t)n * 3);
  scratch.want(&d_perm, n);
  scratch.want(&d_perm2, n);
  scratch.want(&d_place, n);
  scratch.want(&d_place_sorted, n);
  scratch.want(&d_place_sorted2, n);
  scratch.want(&d_iota, n);
  scratch.want(&d_vals, n);
  scratch.want(&d_cen_key, n);
  scratch.want(&d_cen_key2, n);
  scratch.want(&d_seg_of_pos, n);
  scratch.want(&d_segs, total_segs);
  scratch.want(&d_bounds, max_segs * 6);
  scratch.want(&d_seg_axis, max_segs);
  scratch.want(&d_node_axes, n_nodes);
  scratch.want(&d_topo, n_nodes);
  scratch.want(&d_ids, n_nodes);
  scratch.want(&d_tmp, tmp_bytes);
  BUILD_TRY(scratch.commit());

  // the results outlive this call
  FlatNode* d_nodes = nullptr;
  FlatTri* d_tris = nullptr;
  FlatTriShade* d_shade = nullptr;
  BUILD_TRY(cudaMalloc((void**)&d_nodes, (size_t)n_nodes * sizeof(FlatNode)));
  cudaError_t e2 = cudaMalloc((void**)&d_tris, (size_t)n * sizeof(FlatTri));
  cudaError_t e3 = e2 == cudaSuccess ? cudaMalloc((void**)&d_shade, (size_t)n * sizeof(FlatTriShade)) : e2;
  if (e3 != cudaSuccess) {
    cudaFree(d_nodes);
    cudaFree(d_tris);
    err = std::string("cudaMalloc (tree): ") + cudaGetErrorString(e3);
    return false;
  }
  auto fail = [&]() {
    cudaFree(d_nodes);
    cudaFree(d_tris);
    cudaFree(d_shade);
    return false;
  };
#define BUILD_TRY2(expr)                                                   \
  do {                                                                     \
    cudaError_t e__ = (expr);                                              \
    if (e__ != cudaSuccess) {                                              \
      err = std::string(#expr) + ": " + cudaGetErrorString(e__);           \
      return fail();                                                       \
    }                                                                      \
  } while (0)

  // ---- uploads ----
  std::vector<Seg> flat_segs;
  flat_segs.reserve(total_segs);
  std::vector<size_t> level_off;
  for (const auto& l : topo.levels) {
    level_off.push_back(flat_segs.size());
    flat_segs.insert(flat_segs.end(), l.begin(), l.end());
  }
  std::vector<uint32_t> flat_ids;
  std::vector<size_t> height_off;
  for (const auto& h : topo.by_height) {
    height_off.push_back(flat_ids.size());
    flat_ids.insert(flat_ids.end(), h.begin(), h.end());
  }
  BUILD_TRY2(cudaMemcpyAsync(d_pos, mesh.positions, (size_t)n * 9 * sizeof(float), cudaMemcpyHostToDevice, stream));
  BUILD_TRY2(cudaMemcpyAsync(d_nrm, mesh.normals, (size_t)n * 9 * sizeof(double), cudaMemcpyHostToDevice, stream));
  BUILD_TRY2(cudaMemcpyAsync(d_uv, mesh.uvs, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, stream));
  if (total_segs)
    BUILD_TRY2(cudaMemcpyAsync(d_segs, flat_segs.data(), total_segs * sizeof(Seg), cudaMemcpyHostToDevice, stream));
  BUILD_TRY2(cudaMemcpyAsync(d_topo, topo.nodes.data(), (size_t)n_nodes * sizeof(TopoNode), cudaMemcpyHostToDevice, stream));
  BUILD_TRY2(cudaMemcpyAsync(d_ids, flat_ids.data(), flat_ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  BUILD_TRY2(cudaMemsetAsync(d_node_axes, 0, (size_t)n_nodes * sizeof(uint32_t), stream));

  // ---- the permutation, one binary level at a time ----
  k_prepare<<<blocks_for(n), 256, 0, stream>>>(d_pos, n, d_bmin, d_bmax, d_cen, d_perm);
  for (size_t d = 0; d < topo.levels.size(); ++d) {
    const uint32_t n_segs = (uint32_t)topo.levels[d].size();
    if (n_segs == 0) continue;
    const Seg* segs = d_segs + level_off[d];
    k_init_bounds<<<blocks_for((size_t)n_segs * 6), 256, 0, stream>>>(d_bounds, n_segs);
    k_seg_bounds<<<blocks_for(n), 256, 0, stream>>>(segs, n_segs, d_perm, d_cen, n, d_seg_of_pos, d_bounds);
    k_seg_axis<<<blocks_for(n_segs), 256, 0, stream>>>(segs, n_segs, d_bounds, d_seg_axis, d_node_axes);
    k_tri_keys<<<blocks_for(n), 256, 0, stream>>>(segs, d_seg_of_pos, d_seg_axis, d_perm, d_cen, n, d_cen_key, d_place, d_iota);
    // (1) all triangles by centroid key; the input values are 0..n-1, so equal keys stay in ascending index
    size_t tb = tmp_bytes;
    BUILD_TRY2(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_cen_key, d_cen_key2, d_iota, d_vals, (int)n, 0, 64, stream));
    // (2) stable by segment start: every triangle returns to its own range, now in (key, index) order
    k_gather_u32<<<blocks_for(n), 256, 0, stream>>>(d_place, d_vals, n, d_place_sorted);
    tb = tmp_bytes;
    BUILD_TRY2(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_place_sorted, d_place_sorted2, d_vals, d_perm2, (int)n, 0, 32, stream));
    std::swap(d_perm, d_perm2);
  }

  // ---- records and boxes ----
  k_emit_tris<<<blocks_for(n), 256, 0, stream>>>(d_perm, n, d_pos, d_nrm, d_uv, d_tris, d_shade);
  for (size_t h = 1; h < topo.by_height.size(); ++h) {
    const uint32_t cnt = (uint32_t)topo.by_height[h].size();
    if (cnt == 0) continue;
    k_node_boxes<<<blocks_for((size_t)cnt * 4), 256, 0, stream>>>(d_ids + height_off[h], cnt, d_topo, d_perm, d_bmin, d_bmax,
                                                                  d_node_axes, d_nodes);
  }
  BUILD_TRY2(cudaGetLastError());
  FlatNode root_node;
  BUILD_TRY2(cudaMemcpyAsync(&root_node, d_nodes + root, sizeof(FlatNode), cudaMemcpyDeviceToHost, stream));
  BUILD_TRY2(cudaStreamSynchronize(stream));
#undef BUILD_TRY2

  out.nodes = d_nodes;
  out.tris = d_tris;
  out.shade = d_shade;
  out.n_nodes = n_nodes;
  out.n_tris = n;
  out.root = root;
  out.n_leaves = topo.n_leaves;
  out.height = height;
  out.max_stack = 3 * height + 1;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int k = 0; k < 4; ++k) {
    if (root_node.child[k] == 0xFFFFFFFFu) continue;
    mn[0] = fminf(mn[0], root_node.min_x[k]); mx[0] = fmaxf(mx[0], root_node.max_x[k]);
    mn[1] = fminf(mn[1], root_node.min_y[k]); mx[1] = fmaxf(mx[1], root_node.max_y[k]);
    mn[2] = fminf(mn[2], root_node.min_z[k]); mx[2] = fmaxf(mx[2], root_node.max_z[k]);
  }
  for (int a = 0; a < 3; ++a) {
    out.bbox_min[a] = (double)mn[a];
    out.bbox_max[a] = (double)mx[a];
  }
  return true;
}

void free_qbvh_device(DeviceQbvh& q) {
  cudaFree(q.nodes);
  cudaFree(q.tris);
  cudaFree(q.shade);
  memset(&q, 0, sizeof(q));
}

} // namespace yart
