// nt_host.hpp — C++ host side above the C ABI (include/nettracer_b200.h): the part a NetTracer host
// keeps — scene set-up, camera resolution, image output — written in C++ because the reference is
// compiled (Java) code and no JDK exists in this image.  Mirrors nettracer_b200/scene.py one to one;
// the reference's own classes cannot be mirrored by name (/root/reference/README:1-3 holds no source).
// Header-only; link with -lnettracer_b200.  All arithmetic of Camera::resolve follows SPEC-PROVISIONAL §2.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/nettracer_b200.h"

namespace nthost {

struct Material {
    double r = 1, g = 1, b = 1, ka = 0.1, kd = 0.8, ks = 0, shininess = 1, kr = 0, kt = 0, ior = 1;
};

struct Vec3 { double x, y, z; };
inline Vec3 operator-(Vec3 a, Vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline Vec3 operator*(Vec3 a, double s) { return { a.x * s, a.y * s, a.z * s }; }
inline double dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline Vec3 normalize(Vec3 v) { return v * (1.0 / std::sqrt(dot(v, v))); }

struct Camera {
    Vec3 eye{ 0, 0, 5 }, at{ 0, 0, 0 }, up{ 0, 1, 0 };
    double vfov_deg = 45;
    // SPEC-PROVISIONAL §2: the only place a transcendental (tan) is used.
    nt_camera resolve(uint32_t width, uint32_t height) const {
        const Vec3 w = normalize(eye - at), u = normalize(cross(up, w)), v = cross(w, u);
        const double hh = std::tan(vfov_deg * (M_PI / 180.0) / 2.0), hw = hh * width / height;
        const Vec3 p00 = { -w.x - hw * u.x + hh * v.x, -w.y - hw * u.y + hh * v.y, -w.z - hw * u.z + hh * v.z };
        const Vec3 dx = u * (2.0 * hw / width), dy = v * (-(2.0 * hh / height));
        nt_camera c;
        const Vec3 a[4] = { eye, p00, dx, dy };
        double *dst[4] = { c.eye, c.p00, c.dx, c.dy };
        for (int i = 0; i < 4; ++i) { dst[i][0] = a[i].x; dst[i][1] = a[i].y; dst[i][2] = a[i].z; }
        return c;
    }
};

class Scene {
  public:
    double ambient[3] = { 1, 1, 1 }, background[3] = { 0, 0, 0 };
    int add_material(const Material &m) {
        const double row[10] = { m.r, m.g, m.b, m.ka, m.kd, m.ks, m.shininess, m.kr, m.kt, m.ior };
        materials_.insert(materials_.end(), row, row + 10);
        return (int)(materials_.size() / 10) - 1;
    }
    void add_sphere(Vec3 c, double r, int mat) { push(spheres_, { c.x, c.y, c.z, r }); sphere_mat_.push_back(mat); }
    void add_plane(Vec3 n, double d, int mat) {
        const double len = std::sqrt(dot(n, n));
        push(planes_, { n.x / len, n.y / len, n.z / len, d / len });
        plane_mat_.push_back(mat);
    }
    void add_triangle(Vec3 a, Vec3 b, Vec3 c, int mat) { push(triangles_, { a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z }); triangle_mat_.push_back(mat); }
    void add_light(Vec3 p, Vec3 rgb) { push(lights_, { p.x, p.y, p.z, rgb.x, rgb.y, rgb.z }); }

    nt_scene_desc desc() const {
        nt_scene_desc d{};
        d.struct_size = sizeof d;
        d.n_spheres = (uint32_t)sphere_mat_.size(); d.n_planes = (uint32_t)plane_mat_.size();
        d.n_triangles = (uint32_t)triangle_mat_.size(); d.n_materials = (uint32_t)(materials_.size() / 10);
        d.n_lights = (uint32_t)(lights_.size() / 6);
        d.spheres = spheres_.data(); d.sphere_mat = sphere_mat_.data();
        d.planes = planes_.data(); d.plane_mat = plane_mat_.data();
        d.triangles = triangles_.data(); d.triangle_mat = triangle_mat_.data();
        d.materials = materials_.data(); d.lights = lights_.data();
        for (int i = 0; i < 3; ++i) { d.ambient[i] = ambient[i]; d.background[i] = background[i]; }
        return d;
    }

  private:
    static void push(std::vector<double> &v, std::initializer_list<double> x) { v.insert(v.end(), x); }
    std::vector<double> spheres_, planes_, triangles_, materials_, lights_;
    std::vector<int32_t> sphere_mat_, plane_mat_, triangle_mat_;
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
    if (rc != NT_OK) throw Error(rc, "nettracer_b200 error " + std::to_string(rc) + ": " + nt_last_error());
}

// One scene resident on one GPU.  No CPU fallback: the constructor throws when no sm_100 device exists.
class Renderer {
  public:
    Renderer(const Scene &scene, int device = 0) {
        const nt_scene_desc d = scene.desc();
        check(nt_scene_create(&d, device, &h_));
    }
    ~Renderer() { nt_scene_destroy(h_); }
    Renderer(const Renderer &) = delete;
    Renderer &operator=(const Renderer &) = delete;

    std::vector<uint8_t> render(const Camera &cam, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth,
                                nt_precision precision = NT_F64_STRICT, nt_render_stats *stats = nullptr) {
        nt_render_params p{};
        p.struct_size = sizeof p;
        p.width = w; p.height = h; p.spp = spp; p.max_depth = max_depth; p.precision = precision;
        p.camera = cam.resolve(w, h);
        p.shard_index = 0; p.shard_count = 1; p.band_rows = 16; p.layout = NT_LAYOUT_FULL;
        std::vector<uint8_t> rgba((size_t)w * h * 4);
        check(nt_render(h_, &p, rgba.data(), (size_t)w * 4, stats));
        return rgba;
    }

  private:
    nt_scene *h_ = nullptr;
};

// The same scene on several GPUs of this process (nt_multi_*): interleaved row bands, every GPU stores its bands
// straight into the host frame.  Pass a page-locked buffer (PinnedFrame) to avoid the staging copy.
class MultiRenderer {
  public:
    MultiRenderer(const Scene &scene, const std::vector<int> &devices) {
        const nt_scene_desc d = scene.desc();
        check(nt_multi_create(&d, devices.data(), (int)devices.size(), &h_));
    }
    ~MultiRenderer() { nt_multi_destroy(h_); }
    MultiRenderer(const MultiRenderer &) = delete;
    MultiRenderer &operator=(const MultiRenderer &) = delete;
    int devices() const { return nt_multi_device_count(h_); }

    void render_into(uint8_t *rgba, size_t stride, const Camera &cam, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth,
                     nt_precision precision = NT_F64_STRICT, nt_render_stats *stats = nullptr, uint32_t band_rows = 8) {
        nt_render_params p{};
        p.struct_size = sizeof p;
        p.width = w; p.height = h; p.spp = spp; p.max_depth = max_depth; p.precision = precision;
        p.camera = cam.resolve(w, h);
        p.shard_index = 0; p.shard_count = 1; p.band_rows = band_rows; p.layout = NT_LAYOUT_FULL;
        check(nt_multi_render(h_, &p, rgba, stride, stats));
    }
    std::vector<uint8_t> render(const Camera &cam, uint32_t w, uint32_t h, uint32_t spp, uint32_t max_depth,
                                nt_precision precision = NT_F64_STRICT, nt_render_stats *stats = nullptr) {
        std::vector<uint8_t> rgba((size_t)w * h * 4);
        render_into(rgba.data(), (size_t)w * 4, cam, w, h, spp, max_depth, precision, stats);
        return rgba;
    }

  private:
    nt_multi *h_ = nullptr;
};

inline void write_ppm(const std::string &path, const std::vector<uint8_t> &rgba, uint32_t w, uint32_t h) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fprintf(f, "P6\n%u %u\n255\n", w, h);
    for (size_t i = 0; i < (size_t)w * h; ++i) std::fwrite(&rgba[4 * i], 1, 3, f);
    std::fclose(f);
}

} // namespace nthost
