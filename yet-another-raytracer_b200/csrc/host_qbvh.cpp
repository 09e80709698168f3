// host_qbvh.cpp -- L4QBVH::new (reference qbvh.rs:251-361) built on the host straight into the
// flat device layout of host_common.h.
//
// Same tree as the reference: recursive median split on the longest centroid axis
// (split, qbvh.rs:636-693), leaves of <= 4 triangles (qbvh.rs:261-279), two binary levels per
// 4-wide node, post-order node emission so the root is the LAST node (qbvh.rs:308-345).
// Differences that do not change any closest-hit result:
//   * Rust's sort_unstable_by leaves the order of equal centroids unspecified; ties are broken
//     by original triangle index here (and in the oracle), so the tree is reproducible;
//   * boxes are stored as f32 -- lossless, every coordinate is an f32 OBJ value widened to
//     f64 by the reference (triangle.rs:118-122); absent children get +FLT_MAX boxes where the
//     reference stores f64::MAX (qbvh.rs:570-572): both make t0 == t1 so `tfar > tnear` fails;
//   * leaves keep the three VERTICES (the device forms e1 = v1 - v0 in f64 exactly like
//     precompute_soa_triangle, qbvh.rs:614-616) instead of a HashMap of f64 SoA blocks.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>

#include "host_common.h"

namespace yart {
namespace {

struct Builder {
  const yart_trimesh& mesh;
  std::vector<uint32_t> order;   // permutation: tree position -> original triangle
  std::vector<double> cen;       // [n][3] centroid of the triangle AABB (hittable.rs:12-22)
  std::vector<float> bmin, bmax; // [n][3] triangle AABB (triangle.rs:20-45), exact in f32
  FlatQbvh& out;

  struct Box {
    float mn[3], mx[3];
    bool some;
  };

  explicit Builder(const yart_trimesh& m, FlatQbvh& o) : mesh(m), out(o) {}

  void prepare() {
    const uint32_t n = mesh.n_tris;
    order.resize(n);
    std::iota(order.begin(), order.end(), 0u);
    cen.resize((size_t)n * 3);
    bmin.resize((size_t)n * 3);
    bmax.resize((size_t)n * 3);
    for (uint32_t i = 0; i < n; ++i) {
      const float* p = mesh.positions + (size_t)i * 9;
      for (int a = 0; a < 3; ++a) {
        // f64::min / f64::max folds over the three vertices; on f32-exact inputs fminf/fmaxf
        // select the same element
        float lo = fminf(fminf(p[a], p[3 + a]), p[6 + a]);
        float hi = fmaxf(fmaxf(p[a], p[3 + a]), p[6 + a]);
        bmin[(size_t)i * 3 + a] = lo;
        bmax[(size_t)i * 3 + a] = hi;
        cen[(size_t)i * 3 + a] = ((double)hi + (double)lo) / 2.0;
      }
    }
  }

  // split (qbvh.rs:636-693): choose the axis, sort the range, halves at len/2
  uint32_t split(size_t lo, size_t hi) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = lo; i < hi; ++i) {
      const double* c = &cen[(size_t)order[i] * 3];
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::fmin(mn[a], c[a]);
        mx[a] = std::fmax(mx[a], c[a]);
      }
    }
    const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    uint32_t axis = 0;
    if (ey > ex) axis = 1;
    if (ez > std::fmax(ey, ex)) axis = 2;
    const double* c = cen.data();
    std::sort(order.begin() + lo, order.begin() + hi, [c, axis](uint32_t a, uint32_t b) {
      const double ca = c[(size_t)a * 3 + axis], cb = c[(size_t)b * 3 + axis];
      return ca < cb || (ca == cb && a < b);
    });
    return axis;
  }

  static Box merge(const Box& a, const Box& b) {
    if (!a.some) return b;
    if (!b.some) return a;
    Box r;
    r.some = true;
    for (int k = 0; k < 3; ++k) {
      r.mn[k] = fminf(a.mn[k], b.mn[k]);
      r.mx[k] = fmaxf(a.mx[k], b.mx[k]);
    }
    return r;
  }

  // construct (qbvh.rs:253-347); returns the subtree box, its child id and its node height
  Box construct(size_t lo, size_t hi, uint32_t& id, uint32_t& height) {
    const size_t n = hi - lo;
    Box none;
    none.some = false;
    height = 0;
    if (n == 0) {
      id = 0xFFFFFFFFu;
      return none;
    }
    if (n <= 4) {
      Box b;
      b.some = true;
      for (int a = 0; a < 3; ++a) {
        b.mn[a] = bmin[(size_t)order[lo] * 3 + a];
        b.mx[a] = bmax[(size_t)order[lo] * 3 + a];
      }
      for (size_t i = lo + 1; i < hi; ++i)
        for (int a = 0; a < 3; ++a) {
          b.mn[a] = fminf(b.mn[a], bmin[(size_t)order[i] * 3 + a]);
          b.mx[a] = fmaxf(b.mx[a], bmax[(size_t)order[i] * 3 + a]);
        }
      id = (uint32_t)lo | (1u << 31) | ((uint32_t)n << 27); // qbvh.rs:270
      out.n_leaves++;
      return b;
    }
    const uint32_t top = split(lo, hi);
    const size_t mid = lo + n / 2;
    const uint32_t la = split(lo, mid);
    const size_t lmid = lo + (mid - lo) / 2;
    uint32_t ids[4], hs[4];
    Box bx[4];
    bx[0] = construct(lo, lmid, ids[0], hs[0]);
    bx[1] = construct(lmid, mid, ids[1], hs[1]);
    const uint32_t ra = split(mid, hi);
    const size_t rmid = mid + (hi - mid) / 2;
    bx[2] = construct(mid, rmid, ids[2], hs[2]);
    bx[3] = construct(rmid, hi, ids[3], hs[3]);

    FlatNode nd;
    for (int k = 0; k < 4; ++k) {
      const bool s = bx[k].some;
      // (+ 0.0f turns a -0.0 plane into +0.0: which of two equal zeros a min/max fold keeps depends on the
      // fold order, and no slab test can tell them apart; this makes the stored tree canonical, so the GPU
      // builder -- another fold order -- produces the same bytes)
      nd.min_x[k] = s ? bx[k].mn[0] + 0.0f : FLT_MAX;
      nd.min_y[k] = s ? bx[k].mn[1] + 0.0f : FLT_MAX;
      nd.min_z[k] = s ? bx[k].mn[2] + 0.0f : FLT_MAX;
      nd.max_x[k] = s ? bx[k].mx[0] + 0.0f : FLT_MAX;
      nd.max_y[k] = s ? bx[k].mx[1] + 0.0f : FLT_MAX;
      nd.max_z[k] = s ? bx[k].mx[2] + 0.0f : FLT_MAX;
      nd.child[k] = ids[k];
    }
    nd.axes = top | (la << 2) | (ra << 4);
    nd.pad[0] = nd.pad[1] = nd.pad[2] = 0;
    out.nodes.push_back(nd);
    id = (uint32_t)(out.nodes.size() - 1);
    height = 1 + std::max(std::max(hs[0], hs[1]), std::max(hs[2], hs[3]));
    return merge(merge(bx[0], bx[1]), merge(bx[2], bx[3]));
  }
};

} // namespace

bool build_qbvh(const yart_trimesh& mesh, FlatQbvh& out, std::string& err) {
  out = FlatQbvh();
  if (!mesh.positions || !mesh.normals || !mesh.uvs) {
    err = "trimesh has null arrays";
    return false;
  }
  if (mesh.n_tris <= 4) {
    // the reference builds zero nodes and then underflows in hit (SURVEY A-17)
    err = "L4QBVH needs more than 4 triangles (the reference panics on such meshes, qbvh.rs:383-384)";
    return false;
  }
  if (mesh.n_tris >= (1u << 27)) {
    err = "L4QBVH child ids hold 27 bits of triangle index (qbvh.rs:270)";
    return false;
  }
  Builder b(mesh, out);
  b.prepare();
  uint32_t root_id, height;
  Builder::Box box = b.construct(0, mesh.n_tris, root_id, height);
  out.root = root_id;
  out.height = height;
  out.max_stack = 3 * height + 1;
  for (int a = 0; a < 3; ++a) {
    out.bbox_min[a] = (double)(box.mn[a] + 0.0f);
    out.bbox_max[a] = (double)(box.mx[a] + 0.0f);
  }
  out.tris.resize(mesh.n_tris);
  out.shade.resize(mesh.n_tris);
  for (uint32_t pos = 0; pos < mesh.n_tris; ++pos) {
    const uint32_t src = b.order[pos];
    const float* p = mesh.positions + (size_t)src * 9;
    FlatTri& t = out.tris[pos];
    for (int j = 0; j < 3; ++j) {
      t.v0[j] = p[j];
      t.v1[j] = p[3 + j];
      t.v2[j] = p[6 + j];
    }
    t.orig = src;
    t.pad1 = t.pad2 = 0;
    FlatTriShade& s = out.shade[pos];
    for (int k = 0; k < 3; ++k) {
      for (int j = 0; j < 3; ++j) s.n[k][j] = mesh.normals[(size_t)src * 9 + k * 3 + j];
      for (int j = 0; j < 2; ++j) s.uv[k][j] = mesh.uvs[(size_t)src * 6 + k * 2 + j];
    }
  }
  return true;
}

} // namespace yart
