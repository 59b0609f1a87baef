// render_demo.cpp — minimal C++ host: builds a small scene, renders it through the C ABI, writes a PPM.
//   ./render_demo out.ppm [width height spp depth f64|f32]
// Exit code 3 + message when no sm_100 GPU is usable (there is no CPU fallback).
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "nt_host.hpp"

int main(int argc, char **argv) {
    using namespace nthost;
    const std::string out = argc > 1 ? argv[1] : "demo.ppm";
    const uint32_t w = argc > 2 ? std::atoi(argv[2]) : 320, h = argc > 3 ? std::atoi(argv[3]) : 180;
    const uint32_t spp = argc > 4 ? std::atoi(argv[4]) : 4, depth = argc > 5 ? std::atoi(argv[5]) : 4;
    const nt_precision prec = argc > 6 && !std::strcmp(argv[6], "f32") ? NT_F32_FAST : NT_F64_STRICT;

    Scene s;
    s.background[0] = 0.05; s.background[1] = 0.07; s.background[2] = 0.12;
    Material matte; matte.r = 0.7; matte.g = 0.7; matte.b = 0.72; matte.kd = 0.8; matte.kr = 0.15;
    Material glass; glass.kd = 0.05; glass.ks = 0.5; glass.shininess = 120; glass.kr = 0.1; glass.kt = 0.85; glass.ior = 1.5;
    Material red; red.r = 0.85; red.g = 0.2; red.b = 0.15; red.ks = 0.4; red.shininess = 40;
    Material mirror; mirror.kd = 0.15; mirror.ks = 0.6; mirror.shininess = 100; mirror.kr = 0.75;
    const int m0 = s.add_material(matte), m1 = s.add_material(glass), m2 = s.add_material(red), m3 = s.add_material(mirror);
    s.add_plane({ 0, 1, 0 }, 0.0, m0);
    s.add_plane({ 0.2, 0.1, 1 }, -9.0, m0); // tilted back wall: exercises the general plane path
    s.add_sphere({ -1.6, 1.0, 0.0 }, 1.0, m1);
    s.add_sphere({ 1.2, 0.8, -0.8 }, 0.8, m2);
    s.add_sphere({ 0.2, 0.5, 1.6 }, 0.5, m3);
    s.add_triangle({ -3.5, 0.0, -2.5 }, { -1.5, 0.0, -3.5 }, { -2.5, 2.4, -3.0 }, m2);
    s.add_light({ -4, 7, 5 }, { 0.7, 0.68, 0.65 });
    s.add_light({ 5, 6, 2 }, { 0.35, 0.38, 0.45 });
    Camera cam; cam.eye = { 0.3, 2.2, 7.5 }; cam.at = { 0, 0.8, 0 }; cam.vfov_deg = 42;

    try {
        Renderer r(s, 0);
        nt_render_stats st{};
        const auto img = r.render(cam, w, h, spp, depth, prec, &st);
        write_ppm(out, img, w, h);
        std::cout << "rendered " << w << "x" << h << " spp " << spp << " depth " << depth << ": "
                  << (st.rays_primary + st.rays_secondary + st.rays_shadow) << " rays, kernel " << st.kernel_ms << " ms -> " << out << "\n";
    } catch (const Error &e) {
        std::cerr << e.what() << "\n";
        return e.code == NT_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
