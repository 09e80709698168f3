/* yart_main.c -- the reference's `main` (raytracer/src/main.rs:777-782: parse -> resolve options -> render)
 * written against the C ABI alone: plain C99, no Python, no torch, nothing but include/yart.h and
 * libyart_b200.so.  It is the proof that the boundary is usable from any host language the way a Rust
 * `-sys` crate would use it, and tests/test_c_host.py checks that its image equals the one
 * the Python binding produces.
 *
 *   yart_main --info                          host-only: presets, OBJ load, QBVH build (runs without a GPU)
 *   yart_main SCENE W H SPP OUT.ppm [SEED]    render on GPU 0, write a binary PPM (RGB of the RGBA result)
 *
 * build: gcc -std=c99 -O2 -Iinclude examples/yart_main.c -o yart_main -L<pkg dir> -lyart_b200 -Wl,-rpath,<pkg dir>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "yart.h"

static int fail(const char* what, const yart_ctx* ctx) {
  fprintf(stderr, "yart_main: %s: %s\n", what, ctx ? yart_last_error(ctx) : yart_last_error_global());
  return 1;
}

static int info(const char* assets) {
  int i, n = yart_preset_count();
  char path[1024];
  yart_objfile* obj = NULL;
  yart_trimesh mesh;
  yart_qbvh* q = NULL;
  yart_qbvh_info qi;
  uint32_t w, h;
  printf("%s, %d presets:", yart_version(), n);
  for (i = 0; i < n; ++i) printf(" %s", yart_preset_name(i));
  printf("\n");
  yart_resolve_dimensions(600, 600, 1920, 0, &w, &h); /* --width 1920 keeps the preset's aspect (main.rs:166-186) */
  printf("resolve_dimensions(600x600, --width 1920) = %ux%u\n", w, h);
  snprintf(path, sizeof path, "%s/cube.obj", assets);
  if (yart_obj_load(path, &obj) != YART_OK) return fail("yart_obj_load", NULL);
  if (yart_obj_trimesh(obj, &mesh) != YART_OK) return fail("yart_obj_trimesh", NULL);
  if (yart_qbvh_build(&mesh, &q) != YART_OK) return fail("yart_qbvh_build", NULL);
  if (yart_qbvh_get_info(q, &qi) != YART_OK) return fail("yart_qbvh_get_info", NULL);
  printf("cube.obj: %u triangles -> %u nodes, %u leaves, root %u\n", mesh.n_tris, qi.n_nodes, qi.n_leaves, qi.root);
  yart_qbvh_free(q);
  yart_obj_free(obj);
  printf("devices: %d\n", yart_device_count());
  return 0;
}

int main(int argc, char** argv) {
  const char* assets = getenv("YART_ASSETS") ? getenv("YART_ASSETS") : "assets/_unpacked";
  if (argc >= 2 && strcmp(argv[1], "--info") == 0) return info(assets);
  if (argc < 6) {
    fprintf(stderr, "usage: %s --info | SCENE WIDTH HEIGHT SPP OUT.ppm [SEED]\n", argv[0]);
    return 2;
  }
  {
    const char* scene = argv[1];
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]), spp = (uint32_t)atoi(argv[4]);
    const uint64_t seed = argc > 6 ? (uint64_t)strtoull(argv[6], NULL, 10) : 1;
    yart_preset* preset = NULL;
    yart_preset_info pi;
    yart_ctx* ctx = NULL;
    yart_camera cam;
    yart_render_opts o;
    yart_stats st;
    double* film;
    unsigned char* rgba;
    FILE* f;
    size_t i;

    if (yart_preset_build(scene, assets, seed, &preset) != YART_OK) return fail("yart_preset_build", NULL);
    if (yart_preset_get_info(preset, &pi) != YART_OK) return fail("yart_preset_get_info", NULL);
    if (yart_ctx_create(0, &ctx) != YART_OK) return fail("yart_ctx_create (no CPU fallback)", NULL);
    if (yart_ctx_set_scene(ctx, yart_preset_scene(preset)) != YART_OK) return fail("yart_ctx_set_scene", ctx);
    if (yart_preset_camera(preset, w, h, -1.0, -1.0, &cam) != YART_OK) return fail("yart_preset_camera", NULL);

    film = (double*)calloc((size_t)w * h * 3, sizeof(double));
    rgba = (unsigned char*)malloc((size_t)w * h * 4);
    if (!film || !rgba) return 3;
    memset(&o, 0, sizeof o);
    o.width = w;
    o.height = h;
    o.sample_begin = 0;
    o.sample_end = spp;
    o.max_depth = pi.max_depth; /* RenderDefaults (main.rs:109-120) */
    o.order = YART_ORDER_NEAR;
    o.seed = seed;
    if (yart_render(ctx, &cam, &o, film, &st) != YART_OK) return fail("yart_render", ctx);
    if (yart_film_finalize(ctx, film, w, h, spp, 0, rgba) != YART_OK) return fail("yart_film_finalize", ctx);
    printf("%s %ux%u %u spp: %llu paths, %llu rays, %.2f ms on the device, %.1f Mrays/s, %llu launches\n", scene, w, h, spp,
           (unsigned long long)st.paths, (unsigned long long)st.rays, st.gpu_ms, (double)st.rays / st.gpu_ms / 1e3,
           (unsigned long long)st.kernel_launches);

    f = fopen(argv[5], "wb");
    if (!f) return 4;
    fprintf(f, "P6\n%u %u\n255\n", w, h);
    for (i = 0; i < (size_t)w * h; ++i) fwrite(rgba + i * 4, 1, 3, f);
    fclose(f);
    free(film);
    free(rgba);
    yart_ctx_destroy(ctx);
    yart_preset_free(preset);
  }
  return 0;
}
