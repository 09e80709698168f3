// host_api.cpp -- the context-free half of the C ABI (include/yart.h): OBJ loading, QBVH build,
// scene presets, option resolution.  No CUDA in this file.
#include <cmath>
#include <cstring>
#include <exception>
#include <new>

#include "host_common.h"

struct yart_objfile {
  yart::TriSoup soup;
};
struct yart_preset {
  yart::Preset p;
};

// extern "C" bodies must not let C++ exceptions cross the ABI (a std::bad_alloc while loading a large OBJ, ...)
#define YART_HOST_GUARD_BEGIN try {
#define YART_HOST_GUARD_END(fn_name)                                                          \
  }                                                                                           \
  catch (const std::bad_alloc&) {                                                             \
    yart::set_global_error(std::string(fn_name) + ": out of host memory");                    \
    return YART_ERR_NOMEM;                                                                    \
  }                                                                                           \
  catch (const std::exception& e) {                                                           \
    yart::set_global_error(std::string(fn_name) + ": " + e.what());                           \
    return YART_ERR_INVALID;                                                                  \
  }                                                                                           \
  catch (...) {                                                                               \
    yart::set_global_error(std::string(fn_name) + ": unknown C++ exception");                 \
    return YART_ERR_INVALID;                                                                  \
  }

extern "C" {

const char* yart_version(void) { return "yart-b200 0.2 (abi 2)"; }
const char* yart_last_error_global(void) { return yart::global_error().c_str(); }

int yart_obj_load(const char* path, yart_objfile** out) {
  if (!path || !out) {
    yart::set_global_error("yart_obj_load: null argument");
    return YART_ERR_INVALID;
  }
  yart_objfile* o = new (std::nothrow) yart_objfile();
  if (!o) return YART_ERR_NOMEM;
  struct Drop {
    yart_objfile* p;
    ~Drop() { delete p; }
  } drop{o};
  YART_HOST_GUARD_BEGIN
  std::string err;
  if (!yart::load_obj(path, o->soup, err)) {
    yart::set_global_error(err);
    return YART_ERR_IO;
  }
  drop.p = nullptr;
  *out = o;
  return YART_OK;
  YART_HOST_GUARD_END("yart_obj_load")
}
void yart_obj_free(yart_objfile* obj) { delete obj; }
int yart_obj_trimesh(const yart_objfile* obj, yart_trimesh* out) {
  if (!obj || !out) {
    yart::set_global_error("yart_obj_trimesh: null argument");
    return YART_ERR_INVALID;
  }
  *out = obj->soup.view();
  return YART_OK;
}

int yart_qbvh_build(const yart_trimesh* mesh, yart_qbvh** out) {
  if (!mesh || !out) {
    yart::set_global_error("yart_qbvh_build: null argument");
    return YART_ERR_INVALID;
  }
  yart_qbvh* q = new (std::nothrow) yart_qbvh();
  if (!q) return YART_ERR_NOMEM;
  struct Drop {
    yart_qbvh* p;
    ~Drop() { delete p; }
  } drop{q};
  YART_HOST_GUARD_BEGIN
  std::string err;
  if (!yart::build_qbvh(*mesh, q->q, err)) {
    yart::set_global_error(err);
    return YART_ERR_INVALID;
  }
  q->n_tris = mesh->n_tris;
  drop.p = nullptr;
  *out = q;
  return YART_OK;
  YART_HOST_GUARD_END("yart_qbvh_build")
}
void yart_qbvh_free(yart_qbvh* q) { delete q; }
int yart_qbvh_get_info(const yart_qbvh* q, yart_qbvh_info* out) {
  if (!q || !out) {
    yart::set_global_error("yart_qbvh_get_info: null argument");
    return YART_ERR_INVALID;
  }
  memset(out, 0, sizeof(*out));
  out->n_nodes = (uint32_t)q->q.nodes.size();
  out->n_leaves = q->q.n_leaves;
  out->n_tris = q->n_tris;
  out->height = q->q.height;
  out->root = q->q.root;
  out->max_stack = q->q.max_stack;
  for (int a = 0; a < 3; ++a) {
    out->bbox_min[a] = q->q.bbox_min[a];
    out->bbox_max[a] = q->q.bbox_max[a];
  }
  return YART_OK;
}
const void* yart_qbvh_nodes(const yart_qbvh* q) { return q ? q->q.nodes.data() : nullptr; }
const void* yart_qbvh_tris(const yart_qbvh* q) { return q ? q->q.tris.data() : nullptr; }
const void* yart_qbvh_shade(const yart_qbvh* q) { return q ? q->q.shade.data() : nullptr; }

int yart_preset_build(const char* name, const char* assets_dir, uint64_t seed, yart_preset** out) {
  if (!name || !assets_dir || !out) {
    yart::set_global_error("yart_preset_build: null argument");
    return YART_ERR_INVALID;
  }
  yart_preset* p = new (std::nothrow) yart_preset();
  if (!p) return YART_ERR_NOMEM;
  struct Drop {
    yart_preset* p;
    ~Drop() { delete p; }
  } drop{p};
  YART_HOST_GUARD_BEGIN
  std::string err;
  if (!yart::build_preset(name, assets_dir, seed, p->p, err)) {
    yart::set_global_error(err);
    return err.rfind("unknown scene", 0) == 0 ? YART_ERR_INVALID : YART_ERR_IO;
  }
  drop.p = nullptr;
  *out = p;
  return YART_OK;
  YART_HOST_GUARD_END("yart_preset_build")
}
void yart_preset_free(yart_preset* p) { delete p; }
const char* yart_preset_note(const yart_preset* p) { return p ? p->p.note.c_str() : ""; }
const yart_scene_desc* yart_preset_scene(const yart_preset* p) { return p ? &p->p.scene.desc : nullptr; }
int yart_preset_get_info(const yart_preset* p, yart_preset_info* out) {
  if (!p || !out) {
    yart::set_global_error("yart_preset_get_info: null argument");
    return YART_ERR_INVALID;
  }
  *out = p->p.info;
  return YART_OK;
}
int yart_preset_count(void) { return 13; }
const char* yart_preset_name(int i) { return (i >= 0 && i < 13) ? yart::kPresetNames[i] : nullptr; }

// resolve_dimensions (main.rs:166-186)
void yart_resolve_dimensions(uint32_t default_w, uint32_t default_h, uint32_t w_override, uint32_t h_override,
                             uint32_t* w, uint32_t* h) {
  const double aspect = (double)default_w / (double)default_h;
  if (w_override && h_override) {
    *w = w_override;
    *h = h_override;
  } else if (w_override) {
    *w = w_override;
    *h = (uint32_t)std::fmax(std::round((double)w_override / aspect), 1.0);
  } else if (h_override) {
    *w = (uint32_t)std::fmax(std::round((double)h_override * aspect), 1.0);
    *h = h_override;
  } else {
    *w = default_w;
    *h = default_h;
  }
}

int yart_preset_camera(const yart_preset* p, uint32_t width, uint32_t height, double vfov, double aperture,
                       yart_camera* out) {
  if (!p || !out || !width || !height) {
    yart::set_global_error("yart_preset_camera: bad argument");
    return YART_ERR_INVALID;
  }
  yart::camera_for(p->p.info, width, height, vfov, aperture, out);
  return YART_OK;
}

} // extern "C"
