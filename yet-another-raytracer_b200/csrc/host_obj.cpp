// host_obj.cpp -- Wavefront OBJ front end with the semantics the reference gets from
// tobj 4.0.2 + GPU_LOAD_OPTIONS (triangulate, single_index, ignore points/lines), followed by
// the per-triangle assembly of TriangleMesh::from_obj (reference triangle.rs:110-168):
//   * positions / normals / texcoords are parsed as f32 and widened (triangle.rs:118-134);
//   * polygons are fan-triangulated from their first vertex: (0,k,k+1);
//   * a corner without `vn` gets the f64 face normal (v1-v0)x(v2-v0) normalised
//     (triangle.rs:146-153), a corner without `vt` gets (0,0) (triangle.rs:154-158);
//   * triangles are emitted in file order -- that index is the prim_id of the whole library.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host_common.h"

namespace yart {

namespace {
thread_local std::string g_global_error;

struct Corner {
  long v, vt, vn; // 0-based, -1 = absent
};

inline const char* skip_ws(const char* p) {
  while (*p == ' ' || *p == '\t') ++p;
  return p;
}

// resolve an OBJ index (1-based, negative = relative to the end) to 0-based, -1 if invalid
inline long fix_index(long idx, size_t count) {
  if (idx > 0) return idx - 1;
  if (idx < 0) return (long)count + idx;
  return -1;
}

bool parse_corner(const char*& p, size_t nv, size_t nvt, size_t nvn, Corner& c) {
  c.v = c.vt = c.vn = -1;
  char* end = nullptr;
  long iv = strtol(p, &end, 10);
  if (end == p) return false;
  c.v = fix_index(iv, nv);
  p = end;
  if (*p == '/') {
    ++p;
    if (*p != '/') {
      long it = strtol(p, &end, 10);
      if (end != p) {
        c.vt = fix_index(it, nvt);
        p = end;
      }
    }
    if (*p == '/') {
      ++p;
      long in = strtol(p, &end, 10);
      if (end != p) {
        c.vn = fix_index(in, nvn);
        p = end;
      }
    }
  }
  return true;
}
} // namespace

void set_global_error(const std::string& msg) { g_global_error = msg; }
const std::string& global_error() { return g_global_error; }

bool load_obj(const std::string& path, TriSoup& out, std::string& err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) {
    err = "cannot open OBJ file '" + path + "': " + strerror(errno);
    return false;
  }
  std::vector<char> text;
  {
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    text.resize((size_t)(sz > 0 ? sz : 0) + 1);
    size_t got = sz > 0 ? fread(text.data(), 1, (size_t)sz, f) : 0;
    text[got] = '\0';
    fclose(f);
  }
  std::vector<float> pos, nrm, tex;
  std::vector<Corner> poly;
  out.positions.clear();
  out.normals.clear();
  out.uvs.clear();

  auto emit = [&](const Corner& a, const Corner& b, const Corner& c) -> bool {
    const Corner* cs[3] = {&a, &b, &c};
    double v[3][3];
    for (int k = 0; k < 3; ++k) {
      if (cs[k]->v < 0 || (size_t)cs[k]->v * 3 + 2 >= pos.size()) return false;
      for (int j = 0; j < 3; ++j) {
        float x = pos[(size_t)cs[k]->v * 3 + j];
        out.positions.push_back(x);
        v[k][j] = (double)x;
      }
    }
    // default_normal (triangle.rs:146-148): (v1 - v0) x (v2 - v0), unit_vector, all in f64
    const double e1[3] = {v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2]};
    const double e2[3] = {v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2]};
    const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2],
                 cz = e1[0] * e2[1] - e1[1] * e2[0];
    const double len = std::sqrt(cx * cx + cy * cy + cz * cz);
    const double dn[3] = {cx / len, cy / len, cz / len};
    for (int k = 0; k < 3; ++k) {
      const bool has_n = cs[k]->vn >= 0 && (size_t)cs[k]->vn * 3 + 2 < nrm.size();
      for (int j = 0; j < 3; ++j) out.normals.push_back(has_n ? (double)nrm[(size_t)cs[k]->vn * 3 + j] : dn[j]);
    }
    for (int k = 0; k < 3; ++k) {
      const bool has_t = cs[k]->vt >= 0 && (size_t)cs[k]->vt * 2 + 1 < tex.size();
      for (int j = 0; j < 2; ++j) out.uvs.push_back(has_t ? tex[(size_t)cs[k]->vt * 2 + j] : 0.0f);
    }
    return true;
  };

  char* line = text.data();
  size_t line_no = 0;
  while (*line) {
    char* eol = line;
    while (*eol && *eol != '\n') ++eol;
    char saved = *eol;
    *eol = '\0';
    if (eol > line && eol[-1] == '\r') eol[-1] = '\0';
    ++line_no;
    const char* p = skip_ws(line);
    if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
      char* e;
      p += 1;
      for (int j = 0; j < 3; ++j) {
        float x = strtof(p, &e);
        if (e == p) {
          err = path + ":" + std::to_string(line_no) + ": malformed vertex";
          return false;
        }
        pos.push_back(x);
        p = e;
      }
    } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
      char* e;
      p += 2;
      for (int j = 0; j < 3; ++j) {
        float x = strtof(p, &e);
        if (e == p) {
          err = path + ":" + std::to_string(line_no) + ": malformed normal";
          return false;
        }
        nrm.push_back(x);
        p = e;
      }
    } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
      char* e;
      p += 2;
      float u = strtof(p, &e);
      if (e == p) {
        err = path + ":" + std::to_string(line_no) + ": malformed texcoord";
        return false;
      }
      p = e;
      float v = strtof(p, &e);
      if (e == p) v = 0.0f;
      tex.push_back(u);
      tex.push_back(v);
    } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
      p += 1;
      poly.clear();
      for (;;) {
        p = skip_ws(p);
        if (!*p) break;
        Corner c;
        if (!parse_corner(p, pos.size() / 3, tex.size() / 2, nrm.size() / 3, c)) {
          err = path + ":" + std::to_string(line_no) + ": malformed face";
          return false;
        }
        poly.push_back(c);
      }
      // points and lines are ignored (GPU_LOAD_OPTIONS); polygons fan out from corner 0
      for (size_t k = 1; k + 1 < poly.size(); ++k) {
        if (!emit(poly[0], poly[k], poly[k + 1])) {
          err = path + ":" + std::to_string(line_no) + ": face index out of range";
          return false;
        }
      }
    }
    *eol = saved;
    line = *eol ? eol + 1 : eol;
  }
  if (out.positions.empty()) {
    err = "OBJ file '" + path + "' holds no faces";
    return false;
  }
  return true;
}

} // namespace yart
