// host_presets.cpp -- the reference's 13 scene presets as plain-C scene descriptions.
//
// Mirrors build_scene_preset (reference main.rs:211-432: defaults, background, camera, the
// separate sampling `lights` list) and the world builders of scenes.rs.  Every value below is
// the reference's; each builder cites the lines it restates.  Randomised scenes draw from a
// Philox stream keyed by the caller's seed where the reference uses the OS-seeded thread_rng.
//
// Documented deviations:
//   * next-week-final: `boxes2_size` is taken after the 1000 spheres are added (the reference
//     reads it before, scenes.rs:411-424, and then recurses forever), and the ceiling light is
//     wrapped in FlipFace like cornell's (scenes.rs:355-358 forgets it, leaving the scene unlit).
//   * bunny / teapot: bunny.obj and teapot.obj are not shipped; sycee.obj stands in.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/yart_rng.h"
#include "host_common.h"

namespace yart {

const char* const kPresetNames[13] = {
    "random-scene", "two-spheres", "two-perlin-spheres", "earth",        "simple-light",
    "cornell-box",  "cornell-box-smoke", "next-week-final", "teapot",    "bunny",
    "three-spheres", "sycee",      "david"};

namespace {

const double kPi = 3.14159265358979323846264338327950288;

// sequential draws for scene construction: Philox4x32-10, counter = (n, 0, 0, 0xSCENE)
struct SceneRng {
  uint32_t key[2];
  uint32_t n = 0;
  explicit SceneRng(uint64_t seed) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
  }
  double next() {
    uint32_t c0 = n++, c1 = 0, c2 = 0, c3 = 0x5CE4E000u, k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      uint64_t p0 = (uint64_t)YART_PHILOX_M0 * c0, p1 = (uint64_t)YART_PHILOX_M1 * c2;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += YART_PHILOX_W0;
      k1 += YART_PHILOX_W1;
    }
    uint64_t a = (uint64_t)c0 | ((uint64_t)c1 << 32);
    return (double)(a >> 11) * (1.0 / 9007199254740992.0);
  }
  double range(double lo, double hi) { return lo + (hi - lo) * next(); }
  uint32_t below(uint32_t n_) { // gen_range(0..n)
    uint32_t k = (uint32_t)(next() * (double)n_);
    return k >= n_ ? n_ - 1 : k;
  }
};

struct B { // scene builder helpers
  OwnedScene& s;
  explicit B(OwnedScene& sc) : s(sc) {}

  uint32_t tex_solid(double r, double g, double b) {
    yart_texture t;
    memset(&t, 0, sizeof(t));
    t.kind = YART_TEX_SOLID;
    t.rgb_a[0] = r; t.rgb_a[1] = g; t.rgb_a[2] = b;
    s.textures.push_back(t);
    return (uint32_t)s.textures.size() - 1;
  }
  uint32_t tex_checker(const double odd[3], const double even[3]) {
    yart_texture t;
    memset(&t, 0, sizeof(t));
    t.kind = YART_TEX_CHECKER;
    for (int i = 0; i < 3; ++i) { t.rgb_a[i] = odd[i]; t.rgb_b[i] = even[i]; }
    s.textures.push_back(t);
    return (uint32_t)s.textures.size() - 1;
  }
  uint32_t tex_noise(uint32_t type, double scale, SceneRng& rng) { // NoiseTexture::new (texture.rs:250-257)
    yart_perlin p;                                                  // Perlin::new (texture.rs:96-111)
    for (int i = 0; i < 256; ++i) p.ranfloat[i] = rng.range(0.0, 1.0);
    for (int i = 0; i < 256; ++i)
      for (int k = 0; k < 3; ++k) p.ranvec[i][k] = rng.range(-1.0, 1.0);
    int32_t* perms[3] = {p.perm_x, p.perm_y, p.perm_z};
    for (int a = 0; a < 3; ++a) {
      for (int i = 0; i < 256; ++i) perms[a][i] = i;
      for (int i = 255; i >= 1; --i) { // permute (texture.rs:185-192): target in 0..i, never i
        uint32_t target = rng.below((uint32_t)i);
        int32_t tmp = perms[a][i];
        perms[a][i] = perms[a][target];
        perms[a][target] = tmp;
      }
    }
    s.perlins.push_back(p);
    yart_texture t;
    memset(&t, 0, sizeof(t));
    t.kind = YART_TEX_NOISE;
    t.noise_type = type;
    t.scale = scale;
    t.perlin = (uint32_t)s.perlins.size() - 1;
    s.textures.push_back(t);
    return (uint32_t)s.textures.size() - 1;
  }
  uint32_t tex_image(uint32_t image) {
    yart_texture t;
    memset(&t, 0, sizeof(t));
    t.kind = YART_TEX_IMAGE;
    t.image = image;
    s.textures.push_back(t);
    return (uint32_t)s.textures.size() - 1;
  }
  uint32_t mat(uint32_t kind, uint32_t tex, double fuzz = 0.0) {
    yart_material m;
    memset(&m, 0, sizeof(m));
    m.kind = kind;
    m.texture = tex;
    m.fuzz = fuzz;
    s.materials.push_back(m);
    return (uint32_t)s.materials.size() - 1;
  }
  uint32_t lambertian(double r, double g, double b) { return mat(YART_MAT_LAMBERTIAN, tex_solid(r, g, b)); }
  uint32_t diffuse_light(double r, double g, double b) { return mat(YART_MAT_DIFFUSE_LIGHT, tex_solid(r, g, b)); }
  uint32_t no_material() { return mat(YART_MAT_NONE, 0); }
  uint32_t sf66() { // material.rs:178-185 (N-SF66 Sellmeier terms, c in nm^2)
    yart_material m;
    memset(&m, 0, sizeof(m));
    m.kind = YART_MAT_DIELECTRIC;
    m.sellmeier_b[0] = 2.0245976; m.sellmeier_b[1] = 0.470187196; m.sellmeier_b[2] = 2.59970433;
    m.sellmeier_c[0] = 0.0147053225 * 1e6; m.sellmeier_c[1] = 0.0692998276 * 1e6; m.sellmeier_c[2] = 161.817601 * 1e6;
    s.materials.push_back(m);
    return (uint32_t)s.materials.size() - 1;
  }

  static yart_object blank(uint32_t kind, uint32_t material) {
    yart_object o;
    memset(&o, 0, sizeof(o));
    o.kind = kind;
    o.material = material;
    o.cos_theta = 1.0;
    return o;
  }
  static yart_object sphere(double cx, double cy, double cz, double r, uint32_t material) {
    yart_object o = blank(YART_OBJ_SPHERE, material);
    o.p[0] = cx; o.p[1] = cy; o.p[2] = cz; o.p[3] = r;
    return o;
  }
  static yart_object rect(uint32_t kind, double a0, double a1, double b0, double b1, double k, uint32_t material) {
    yart_object o = blank(kind, material);
    o.p[0] = a0; o.p[1] = a1; o.p[2] = b0; o.p[3] = b1; o.p[4] = k;
    return o;
  }
  static yart_object box(double x0, double y0, double z0, double x1, double y1, double z1, uint32_t material) {
    yart_object o = blank(YART_OBJ_BOX, material);
    o.p[0] = x0; o.p[1] = y0; o.p[2] = z0; o.p[3] = x1; o.p[4] = y1; o.p[5] = z1;
    return o;
  }
  static yart_object triangle(const double v[9], const double uv[6], uint32_t material) {
    yart_object o = blank(YART_OBJ_TRIANGLE, material);
    for (int i = 0; i < 9; ++i) o.p[i] = v[i];
    for (int k = 0; k < 3; ++k) { o.p[9 + 3 * k] = 0.0; o.p[10 + 3 * k] = 1.0; o.p[11 + 3 * k] = 0.0; }
    for (int i = 0; i < 6; ++i) o.p[18 + i] = uv[i];
    return o;
  }
  static void rotate_y(yart_object& o, double degrees) { // RotateY::new (hittable.rs:166-169)
    double radians = degrees * kPi / 180.0;
    o.sin_theta = std::sin(radians);
    o.cos_theta = std::cos(radians);
    o.wrap |= YART_WRAP_ROTATE_Y;
  }
  static void translate(yart_object& o, double x, double y, double z) {
    o.offset[0] = x; o.offset[1] = y; o.offset[2] = z;
    o.wrap |= YART_WRAP_TRANSLATE;
  }
  void medium(yart_object& o, double density, double r, double g, double b) { // ConstantMedium::new (hittable.rs:264-271)
    o.neg_inv_density = -1.0 / density;
    o.material = mat(YART_MAT_ISOTROPIC, tex_solid(r, g, b));
    o.wrap |= YART_WRAP_MEDIUM;
  }
  void add(const yart_object& o) { s.objects.push_back(o); }
  void add_light(const yart_object& o) { s.lights.push_back(o); }

  // the two ground triangles shared by sycee/teapot/bunny/three-spheres (e.g. scenes.rs:450-479)
  void ground_quad(double hx, double hz, uint32_t material) {
    const double a[9] = {-hx, 0.0, -hz, hx, 0.0, -hz, hx, 0.0, hz};
    const double auv[6] = {0.0, 0.0, 1.0, 0.0, 1.0, 1.0};
    const double b[9] = {-hx, 0.0, -hz, -hx, 0.0, hz, hx, 0.0, hz};
    const double buv[6] = {0.0, 0.0, 0.0, 1.0, 1.0, 1.0};
    add(triangle(a, auv, material));
    add(triangle(b, buv, material));
  }
  bool load_mesh(const std::string& path, uint32_t& index, std::string& err) {
    s.soups.emplace_back();
    if (!load_obj(path, s.soups.back(), err)) return false;
    index = (uint32_t)s.soups.size() - 1;
    return true;
  }
  // bunny.obj / teapot.obj are not shipped with the reference (its .MISSING_LARGE_BLOBS lists them; the Rust binary
  // stops with "Failed to load OBJ file", triangle.rs:112-113).  Use the real file when the caller's assets directory
  // has it; otherwise sycee.obj stands in -- and the preset SAYS so (yart_preset_note), or, with YART_STRICT_ASSETS=1
  // in the environment, fails like the reference.
  bool load_mesh_or_stand_in(const std::string& assets, const char* wanted, uint32_t& index, std::string& note, std::string& err) {
    const std::string real = assets + "/" + wanted;
    if (FILE* f = fopen(real.c_str(), "rb")) {
      fclose(f);
      return load_mesh(real, index, err);
    }
    const char* strict = getenv("YART_STRICT_ASSETS");
    if (strict && *strict && *strict != '0') {
      err = "Failed to load OBJ file '" + real + "' (not shipped with the reference; YART_STRICT_ASSETS forbids the sycee.obj stand-in)";
      return false;
    }
    note = std::string("input/") + wanted + " is not shipped with the reference (.MISSING_LARGE_BLOBS): sycee.obj stands in for it";
    return load_mesh(assets + "/sycee.obj", index, err);
  }
  bool load_earth(const std::string& assets, uint32_t& image, std::string& err) { // ImageTexture::new("input/earthmap.jpg")
    const std::string path = assets + "/earthmap_1024x512.rgb8";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
      err = "cannot open '" + path + "' (decoded earthmap.jpg, 1024x512 RGB8)";
      return false;
    }
    std::vector<uint8_t> d((size_t)1024 * 512 * 3);
    size_t got = fread(d.data(), 1, d.size(), f);
    fclose(f);
    if (got != d.size()) {
      err = "short read on '" + path + "'";
      return false;
    }
    s.image_data.push_back(std::move(d));
    yart_image im;
    im.rgb8 = nullptr; // fixed up by finalize()
    im.width = 1024;
    im.height = 512;
    s.images.push_back(im);
    image = (uint32_t)s.images.size() - 1;
    return true;
  }
  static yart_object mesh(uint32_t index, uint32_t material) {
    yart_object o = blank(YART_OBJ_MESH, material);
    o.index = index;
    return o;
  }
  yart_object group(std::vector<yart_object>&& members) {
    s.group_members.push_back(std::move(members));
    yart_object o = blank(YART_OBJ_GROUP, 0);
    o.index = (uint32_t)s.group_members.size() - 1;
    return o;
  }
};

void set_defaults(yart_preset_info& i) { // main.rs:212-220
  memset(&i, 0, sizeof(i));
  i.width = 1200; i.height = 800; i.samples_per_pixel = 100; i.max_depth = 50; i.workers = 30;
  i.vfov = 20.0; i.aperture = 0.0;
}
void set_view(yart_preset_info& i, double fx, double fy, double fz, double ax, double ay, double az, const char* file) {
  i.lookfrom[0] = fx; i.lookfrom[1] = fy; i.lookfrom[2] = fz;
  i.lookat[0] = ax; i.lookat[1] = ay; i.lookat[2] = az;
  snprintf(i.output_filename, sizeof(i.output_filename), "%s", file);
}
void sky(OwnedScene& s) { s.background[0] = 0.7; s.background[1] = 0.8; s.background[2] = 1.0; }

// cornell walls + light shared by cornell_box / cornell_box_smoke (scenes.rs:171-209, 245-283)
void cornell_shell(B& b, double lx0, double lx1, double lz0, double lz1, double emit, uint32_t& white) {
  uint32_t red = b.lambertian(0.65, 0.05, 0.05);
  white = b.lambertian(0.73, 0.73, 0.73);
  uint32_t green = b.lambertian(0.12, 0.45, 0.15);
  uint32_t light = b.diffuse_light(emit, emit, emit);
  b.add(B::rect(YART_OBJ_YZ_RECT, 0.0, 555.0, 0.0, 555.0, 555.0, green));
  b.add(B::rect(YART_OBJ_YZ_RECT, 0.0, 555.0, 0.0, 555.0, 0.0, red));
  yart_object l = B::rect(YART_OBJ_XZ_RECT, lx0, lx1, lz0, lz1, 554.0, light);
  l.wrap |= YART_WRAP_FLIP_FACE;
  b.add(l);
  b.add(B::rect(YART_OBJ_XZ_RECT, 0.0, 555.0, 0.0, 555.0, 0.0, white));
  b.add(B::rect(YART_OBJ_XZ_RECT, 0.0, 555.0, 0.0, 555.0, 555.0, white));
  b.add(B::rect(YART_OBJ_XY_RECT, 0.0, 555.0, 0.0, 555.0, 555.0, white));
}

} // namespace

void OwnedScene::finalize() {
  meshes.clear();
  for (const TriSoup& t : soups) meshes.push_back(t.view());
  groups.clear();
  for (const auto& g : group_members) {
    yart_group gg;
    gg.members = g.data();
    gg.n_members = (uint32_t)g.size();
    gg._pad = 0;
    groups.push_back(gg);
  }
  for (size_t i = 0; i < images.size(); ++i) images[i].rgb8 = image_data[i].data();
  memset(&desc, 0, sizeof(desc));
  desc.objects = objects.data();     desc.n_objects = (uint32_t)objects.size();
  desc.lights = lights.data();       desc.n_lights = (uint32_t)lights.size();
  desc.meshes = meshes.data();       desc.n_meshes = (uint32_t)meshes.size();
  desc.groups = groups.data();       desc.n_groups = (uint32_t)groups.size();
  desc.materials = materials.data(); desc.n_materials = (uint32_t)materials.size();
  desc.textures = textures.data();   desc.n_textures = (uint32_t)textures.size();
  desc.perlins = perlins.data();     desc.n_perlins = (uint32_t)perlins.size();
  desc.images = images.data();       desc.n_images = (uint32_t)images.size();
  for (int k = 0; k < 3; ++k) desc.background_rgb[k] = background[k];
}

bool build_preset(const std::string& name, const std::string& assets, uint64_t seed, Preset& out, std::string& err) {
  out = Preset();
  OwnedScene& s = out.scene;
  yart_preset_info& info = out.info;
  set_defaults(info);
  B b(s);
  SceneRng rng(seed);

  if (name == "random-scene") { // scenes.rs:21-95, main.rs:230-238
    b.add(B::sphere(0.0, -1000.0, 0.0, 1000.0, b.lambertian(0.5, 0.5, 0.5)));
    for (int a = -11; a < 11; ++a)
      for (int c = -11; c < 11; ++c) {
        double choose_mat = rng.next();
        double cx = (double)a + 0.9 * rng.next(), cy = 0.2, cz = (double)c + 0.9 * rng.next();
        double dx = cx - 4.0, dy = cy - 0.2, dz = cz - 0.0;
        if (std::sqrt(dx * dx + dy * dy + dz * dz) > 0.9) {
          if (choose_mat < 0.8) {
            double r = rng.range(-1.0, 1.0), g = rng.range(-1.0, 1.0), bl = rng.range(-1.0, 1.0);
            b.add(B::sphere(cx, cy, cz, 0.2, b.lambertian(r, g, bl)));
          } else if (choose_mat < 0.95) {
            double r = rng.range(0.5, 1.0), g = rng.range(0.5, 1.0), bl = rng.range(0.5, 1.0);
            double fuzz = rng.range(0.0, 0.5);
            b.add(B::sphere(cx, cy, cz, 0.2, b.mat(YART_MAT_METAL, b.tex_solid(r, g, bl), fuzz)));
          } else {
            b.add(B::sphere(cx, cy, cz, 0.2, b.sf66()));
          }
        }
      }
    b.add(B::sphere(0.0, 1.0, 0.0, 1.0, b.sf66()));
    b.add(B::sphere(-4.0, 1.0, 0.0, 1.0, b.lambertian(0.4, 0.2, 0.1)));
    b.add(B::sphere(4.0, 1.0, 0.0, 1.0, b.mat(YART_MAT_METAL, b.tex_solid(0.7, 0.6, 0.5), 0.0)));
    sky(s);
    set_view(info, 13.0, 2.0, 3.0, 0.0, 0.0, 0.0, "random_scene.png");
    info.aperture = 0.1;
    info.samples_per_pixel = 1000;
  } else if (name == "two-spheres") { // scenes.rs:97-118, main.rs:239-245
    const double odd[3] = {0.2, 0.3, 0.1}, even[3] = {0.9, 0.9, 0.9};
    b.add(B::sphere(0.0, -10.0, 0.0, 10.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_checker(odd, even))));
    b.add(B::sphere(0.0, 10.0, 0.0, 10.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_checker(odd, even))));
    sky(s);
    set_view(info, 13.0, 2.0, 3.0, 0.0, 0.0, 0.0, "two_spheres.png");
  } else if (name == "two-perlin-spheres") { // scenes.rs:120-138, main.rs:246-252
    b.add(B::sphere(0.0, -1000.0, 0.0, 1000.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_noise(YART_NOISE_MARBLE, 4.0, rng))));
    b.add(B::sphere(0.0, 2.0, 0.0, 2.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_noise(YART_NOISE_MARBLE, 4.0, rng))));
    sky(s);
    set_view(info, 13.0, 2.0, 30.0, 0.0, 0.0, 0.0, "two_perlin_spheres.png");
  } else if (name == "earth") { // scenes.rs:140-149, main.rs:253-259
    uint32_t img;
    if (!b.load_earth(assets, img, err)) return false;
    b.add(B::sphere(0.0, 0.0, 0.0, 2.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_image(img))));
    sky(s);
    set_view(info, 13.0, 2.0, 3.0, 0.0, 0.0, 0.0, "earth.png");
  } else if (name == "simple-light") { // scenes.rs:151-169, main.rs:260-267
    b.add(B::sphere(0.0, -1000.0, 0.0, 1000.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_noise(YART_NOISE_MARBLE, 4.0, rng))));
    b.add(B::sphere(0.0, 2.0, 0.0, 2.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_noise(YART_NOISE_MARBLE, 4.0, rng))));
    b.add(B::rect(YART_OBJ_XY_RECT, 3.0, 5.0, 1.0, 3.0, -2.0, b.diffuse_light(4.0, 4.0, 4.0)));
    set_view(info, 26.0, 3.0, 6.0, 0.0, 2.0, 0.0, "simple_light.png");
    info.samples_per_pixel = 400;
  } else if (name == "cornell-box") { // scenes.rs:171-243, main.rs:268-286
    uint32_t white;
    cornell_shell(b, 213.0, 343.0, 227.0, 332.0, 15.0, white);
    yart_object box1 = B::box(0.0, 0.0, 0.0, 165.0, 330.0, 165.0, white);
    B::rotate_y(box1, 15.0);
    B::translate(box1, 265.0, 0.0, 295.0);
    b.add(box1);
    b.add(B::sphere(190.0, 90.0, 190.0, 90.0, b.sf66()));
    uint32_t none = b.no_material();
    b.add_light(B::rect(YART_OBJ_XZ_RECT, 213.0, 343.0, 227.0, 332.0, 554.0, none));
    b.add_light(B::sphere(190.0, 90.0, 190.0, 90.0, none));
    set_view(info, 278.0, 278.0, -800.0, 278.0, 278.0, 0.0, "cornell_box.png");
    info.width = 600; info.height = 600; info.samples_per_pixel = 100; info.vfov = 40.0;
  } else if (name == "cornell-box-smoke") { // scenes.rs:245-318, main.rs:287-297
    uint32_t white;
    cornell_shell(b, 113.0, 443.0, 127.0, 432.0, 7.0, white);
    yart_object box1 = B::box(0.0, 0.0, 0.0, 165.0, 330.0, 165.0, white);
    B::rotate_y(box1, 15.0);
    B::translate(box1, 265.0, 0.0, 295.0);
    b.medium(box1, 0.01, 0.0, 0.0, 0.0);
    yart_object box2 = B::box(0.0, 0.0, 0.0, 165.0, 165.0, 165.0, white);
    B::rotate_y(box2, -18.0);
    B::translate(box2, 130.0, 0.0, 65.0);
    b.medium(box2, 0.01, 1.0, 1.0, 1.0);
    b.add(box1);
    b.add(box2);
    set_view(info, 278.0, 278.0, -800.0, 278.0, 278.0, 0.0, "cornell_box_smoke.png");
    info.width = 600; info.height = 600; info.samples_per_pixel = 200; info.vfov = 40.0;
  } else if (name == "next-week-final") { // scenes.rs:320-431, main.rs:298-311
    uint32_t ground = b.lambertian(0.48, 0.83, 0.53);
    std::vector<yart_object> boxes1;
    for (int i = 0; i < 20; ++i)
      for (int j = 0; j < 20; ++j) {
        double w = 100.0;
        double x0 = -1000.0 + (double)i * w, z0 = -1000.0 + (double)j * w, y0 = 0.0;
        double x1 = x0 + w, y1 = rng.range(1.0, 101.0), z1 = z0 + w;
        boxes1.push_back(B::box(x0, y0, z0, x1, y1, z1, ground));
      }
    b.add(b.group(std::move(boxes1)));
    yart_object light = B::rect(YART_OBJ_XZ_RECT, 123.0, 423.0, 147.0, 412.0, 554.0, b.diffuse_light(7.0, 7.0, 7.0));
    light.wrap |= YART_WRAP_FLIP_FACE; // deviation, see the file header
    b.add(light);
    yart_object ms = B::blank(YART_OBJ_MOVING_SPHERE, b.lambertian(0.7, 0.3, 0.1));
    ms.p[0] = 400.0; ms.p[1] = 400.0; ms.p[2] = 200.0;
    ms.p[3] = 400.0 + 30.0; ms.p[4] = 400.0 + 0.0; ms.p[5] = 200.0 + 0.0;
    ms.p[6] = 0.0; ms.p[7] = 1.0; ms.p[8] = 50.0;
    b.add(ms);
    b.add(B::sphere(260.0, 150.0, 45.0, 50.0, b.sf66()));
    b.add(B::sphere(0.0, 150.0, 145.0, 50.0, b.mat(YART_MAT_METAL, b.tex_solid(0.8, 0.8, 0.9), 1.0)));
    b.add(B::sphere(360.0, 150.0, 145.0, 70.0, b.sf66()));
    yart_object fog1 = B::sphere(360.0, 150.0, 145.0, 70.0, 0);
    b.medium(fog1, 0.2, 0.2, 0.4, 0.9);
    b.add(fog1);
    yart_object fog2 = B::sphere(0.0, 0.0, 0.0, 5000.0, 0);
    b.medium(fog2, 0.0001, 1.0, 1.0, 1.0);
    b.add(fog2);
    uint32_t img;
    if (!b.load_earth(assets, img, err)) return false;
    b.add(B::sphere(400.0, 200.0, 400.0, 100.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_image(img))));
    b.add(B::sphere(220.0, 280.0, 300.0, 80.0, b.mat(YART_MAT_LAMBERTIAN, b.tex_noise(YART_NOISE_MARBLE, 0.1, rng))));
    uint32_t white = b.lambertian(0.73, 0.73, 0.73);
    std::vector<yart_object> boxes2;
    for (int i = 0; i < 1000; ++i) {
      double x = rng.range(0.0, 165.0), y = rng.range(0.0, 165.0), z = rng.range(0.0, 165.0);
      boxes2.push_back(B::sphere(x, y, z, 10.0, white));
    }
    yart_object g2 = b.group(std::move(boxes2));
    B::rotate_y(g2, 15.0);
    B::translate(g2, -100.0, 270.0, 395.0);
    b.add(g2);
    b.add_light(B::rect(YART_OBJ_XZ_RECT, 123.0, 423.0, 147.0, 412.0, 554.0, b.no_material()));
    set_view(info, 478.0, 278.0, -600.0, 278.0, 278.0, 0.0, "the_next_week_final_scene.png");
    info.width = 100; info.height = 100; info.samples_per_pixel = 1000; info.vfov = 40.0;
  } else if (name == "teapot") { // scenes.rs:482-533, main.rs:312-332 (teapot.obj missing -> sycee.obj)
    uint32_t light = b.diffuse_light(5.0, 5.0, 5.0), ground = b.lambertian(0.5, 0.5, 0.5), glass = b.sf66();
    b.add(B::sphere(30.0, 40.0, -30.0, 20.0, light));
    b.add(B::sphere(-20.0, 10.0, 50.0, 10.0, light));
    uint32_t m;
    if (!b.load_mesh_or_stand_in(assets, "teapot.obj", m, out.note, err)) return false;
    b.add(B::mesh(m, glass));
    b.ground_quad(80.0, 120.0, ground);
    uint32_t none = b.no_material();
    b.add_light(B::sphere(30.0, 40.0, -30.0, 20.0, none));
    b.add_light(B::sphere(-20.0, 10.0, 50.0, 10.0, none));
    set_view(info, 5.0, 50.0, 60.0, 0.0, 5.0, 0.0, "teapot.png");
    info.width = 1000; info.height = 1000; info.samples_per_pixel = 8000; info.vfov = 30.0; info.aperture = 0.001;
  } else if (name == "bunny") { // scenes.rs:535-579, main.rs:333-349 (bunny.obj missing -> sycee.obj)
    uint32_t glass = b.sf66(), ground = b.lambertian(0.5, 0.5, 0.5), light = b.diffuse_light(5.0, 5.0, 5.0);
    uint32_t m;
    if (!b.load_mesh_or_stand_in(assets, "bunny.obj", m, out.note, err)) return false;
    b.add(B::mesh(m, glass));
    b.ground_quad(20.0, 30.0, ground);
    b.add(B::sphere(0.0, 6.0, -2.0, 2.0, light)); // emitter at z=-2, sampling light at z=+2 (A-11)
    b.add_light(B::sphere(0.0, 6.0, 2.0, 2.0, b.no_material()));
    set_view(info, 0.0, 2.0, 10.0, 0.0, 1.0, 0.0, "bunny.png");
    info.width = 1000; info.height = 1000; info.samples_per_pixel = 50; info.vfov = 30.0; info.aperture = 0.1;
  } else if (name == "three-spheres") { // scenes.rs:626-699, main.rs:350-366
    uint32_t glass = b.sf66(), ground = b.lambertian(0.5, 0.5, 0.5), light = b.diffuse_light(5.0, 5.0, 5.0);
    b.ground_quad(20.0, 30.0, ground);
    b.add(B::sphere(0.0, 1.0, 0.0, 1.0, glass));
    b.add(B::sphere(0.0, 1.3, 0.0, -0.7, glass));
    b.add(B::sphere(0.0, 0.65, 0.0, -0.35, glass));
    b.add(B::sphere(0.0, 0.325, 0.0, -0.125, glass));
    b.add(B::sphere(0.0, 6.0, 2.0, 2.0, light));
    b.add_light(B::sphere(0.0, 6.0, 2.0, 2.0, b.no_material()));
    set_view(info, 1.0, 5.0, -8.0, 0.0, 1.0, 0.0, "three_spheres.png");
    info.width = 1000; info.height = 1000; info.samples_per_pixel = 5000; info.vfov = 30.0; info.aperture = 0.1;
  } else if (name == "sycee") { // scenes.rs:433-480, main.rs:367-383
    uint32_t light = b.diffuse_light(5.0, 5.0, 5.0), ground = b.lambertian(0.5, 0.5, 0.5), glass = b.sf66();
    b.add(B::sphere(0.0, 6.0, 2.0, 2.0, light));
    uint32_t m;
    if (!b.load_mesh(assets + "/sycee.obj", m, err)) return false;
    b.add(B::mesh(m, glass));
    b.ground_quad(20.0, 30.0, ground);
    b.add_light(B::sphere(0.0, 6.0, 2.0, 2.0, b.no_material()));
    set_view(info, 1.0, 5.0, -8.0, 0.0, 1.0, 0.0, "sycee.png");
    info.width = 1000; info.height = 1000; info.samples_per_pixel = 5000; info.vfov = 30.0; info.aperture = 0.1;
  } else if (name == "david") { // scenes.rs:581-624, main.rs:384-421
    uint32_t light = b.diffuse_light(5.0, 5.0, 5.0), white = b.lambertian(1.0, 1.0, 1.0), glass = b.sf66();
    uint32_t m; // the reference loads david.obj twice; one copy serves both instances here
    if (!b.load_mesh(assets + "/david.obj", m, err)) return false;
    b.add(B::mesh(m, white));
    yart_object inst = B::mesh(m, glass);
    B::rotate_y(inst, 300.0);
    B::translate(inst, 50.0, 0.0, 50.0);
    b.add(inst);
    const double lp[5][3] = {{1200.0, 1300.0, 800.0}, {-1200.0, 1300.0, 800.0}, {1200.0, 1300.0, -800.0},
                             {1200.0, -1300.0, -800.0}, {1200.0, 1300.0, -800.0}};
    uint32_t none = b.no_material();
    for (int i = 0; i < 5; ++i) b.add(B::sphere(lp[i][0], lp[i][1], lp[i][2], 700.0, light));
    for (int i = 0; i < 5; ++i) b.add_light(B::sphere(lp[i][0], lp[i][1], lp[i][2], 700.0, none));
    set_view(info, 50.0, 120.0, 300.0, 0.0, 120.0, 0.0, "david.png");
    info.width = 600; info.height = 600; info.samples_per_pixel = 10000; info.vfov = 20.0; info.aperture = 0.001;
  } else {
    err = "unknown scene '" + name + "' (see main.rs:61-76 for the list)";
    return false;
  }
  s.finalize();
  return true;
}

// render()'s camera (main.rs:605-625): vup (0,1,0), focus distance 10, shutter [0,1)
void camera_for(const yart_preset_info& info, uint32_t width, uint32_t height, double vfov, double aperture,
                yart_camera* out) {
  for (int k = 0; k < 3; ++k) {
    out->lookfrom[k] = info.lookfrom[k];
    out->lookat[k] = info.lookat[k];
  }
  out->vup[0] = 0.0; out->vup[1] = 1.0; out->vup[2] = 0.0;
  out->vfov_degrees = vfov < 0.0 ? info.vfov : vfov;
  out->aspect_ratio = (double)width / (double)height;
  out->aperture = aperture < 0.0 ? info.aperture : aperture;
  out->focus_dist = 10.0;
  out->time0 = 0.0;
  out->time1 = 1.0;
}

} // namespace yart
