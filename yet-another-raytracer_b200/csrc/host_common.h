// host_common.h -- shared declarations of the host-side front end (no CUDA here).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/yart.h"

namespace yart {

// error plumbing for context-free calls (yart_last_error_global)
void set_global_error(const std::string& msg);
const std::string& global_error();

// A triangle soup in tobj emission order (what TriangleMesh::from_obj builds, triangle.rs:110-168)
struct TriSoup {
  std::vector<float> positions; // [n][3][3]
  std::vector<double> normals;  // [n][3][3]
  std::vector<float> uvs;       // [n][3][2]
  uint32_t n_tris() const { return (uint32_t)(positions.size() / 9); }
  yart_trimesh view() const {
    yart_trimesh m;
    m.n_tris = n_tris();
    m._pad = 0;
    m.positions = positions.data();
    m.normals = normals.data();
    m.uvs = uvs.data();
    return m;
  }
};

// tobj::load_obj(path, GPU_LOAD_OPTIONS) + the per-triangle assembly of triangle.rs:115-168
bool load_obj(const std::string& path, TriSoup& out, std::string& err);

// ---- flattened L4QBVH: the device layout (DESIGN.md "Data layout in HBM") -------------------
// Node: 128 bytes = one L2 line = four 32-byte sectors, each fetched with one 256-bit load
// (LDG.E.256): {min_x,max_x} {min_y,max_y} {min_z,max_z} {children,axes}.  Lane k of every float[4]
// is child k (LL, LR, RL, RR).
struct alignas(128) FlatNode {
  float min_x[4], max_x[4];
  float min_y[4], max_y[4];
  float min_z[4], max_z[4];
  uint32_t child[4]; // bit31 leaf | count<<27 | first triangle (tree order);  inner: node index;
                     // 0xFFFFFFFF = absent (its box is +FLT_MAX, never hit)
  uint32_t axes;     // top | left<<2 | right<<4   (QBVHNode.top_axis/left_axis/right_axis)
  uint32_t pad[3];
};
static_assert(sizeof(FlatNode) == 128, "node must be one 128-byte line");

// Triangle geometry in tree order, 48 bytes: three float4 (vertex xyz + one payload word).
struct alignas(16) FlatTri {
  float v0[3];
  uint32_t orig; // ORIGINAL triangle index (tobj emission order) = the reported prim_id
  float v1[3];
  uint32_t pad1;
  float v2[3];
  uint32_t pad2;
};
static_assert(sizeof(FlatTri) == 48, "triangle record must be 48 bytes");

// Shading attributes in tree order, read once per shaded hit (96 bytes).
struct alignas(16) FlatTriShade {
  double n[3][3]; // vertex normals
  float uv[3][2];
};
static_assert(sizeof(FlatTriShade) == 96, "shade record must be 96 bytes");

struct FlatQbvh {
  std::vector<FlatNode> nodes;
  std::vector<FlatTri> tris;
  std::vector<FlatTriShade> shade;
  uint32_t root = 0;
  uint32_t n_leaves = 0;
  uint32_t height = 0;    // number of node levels above the leaves
  uint32_t max_stack = 0; // upper bound of the traversal stack depth (3*height + 1)
  double bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
};

// L4QBVH::new (qbvh.rs:251-361) straight into the flat layout.
bool build_qbvh(const yart_trimesh& mesh, FlatQbvh& out, std::string& err);

// ---- scene presets (main.rs:211-432 + scenes.rs) -------------------------------------------
struct OwnedScene {
  std::vector<yart_object> objects, lights;
  std::vector<TriSoup> soups;
  std::vector<yart_trimesh> meshes;
  std::vector<std::vector<yart_object>> group_members;
  std::vector<yart_group> groups;
  std::vector<yart_material> materials;
  std::vector<yart_texture> textures;
  std::vector<yart_perlin> perlins;
  std::vector<std::vector<uint8_t>> image_data;
  std::vector<yart_image> images;
  double background[3] = {0, 0, 0};
  yart_scene_desc desc;
  void finalize(); // (re)build `desc` and the borrowed views from the owned vectors
};

struct Preset {
  OwnedScene scene;
  yart_preset_info info;
  std::string note; // deviations a user must be told about (a stand-in mesh for a missing OBJ file); empty otherwise
};
bool build_preset(const std::string& name, const std::string& assets_dir, uint64_t seed, Preset& out,
                  std::string& err);
extern const char* const kPresetNames[13];

void camera_for(const yart_preset_info& info, uint32_t width, uint32_t height, double vfov, double aperture,
                yart_camera* out);

} // namespace yart

// the opaque handle of include/yart.h (host_api.cpp and yart_device.cu both create it)
struct yart_qbvh {
  yart::FlatQbvh q;
  uint32_t n_tris;
};
