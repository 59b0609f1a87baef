// render_demo.cpp — minimal C++ host: builds a small scene, renders it through the C ABI, writes a PPM.
//   ./render_demo out.ppm [width height spp depth f64|f32 [n_gpus [scene]]]
// n_gpus > 1 (or 0 = all visible) renders through nt_multi_render: one process, one host thread and stream per GPU,
// interleaved row bands.  scene: "demo" (default) or "cornell" = the Cornell-style box of BASELINE.json configs[1..2]
// with fixed sphere positions (tests build the same scene through the Python host and compare with the oracle).
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
    int n_gpus = argc > 7 ? std::atoi(argv[7]) : 1;
    const bool cornell = argc > 8 && !std::strcmp(argv[8], "cornell");

    Scene s;
    Camera cam;
    if (cornell) {
        s.background[0] = 0.02; s.background[1] = 0.02; s.background[2] = 0.03;
        Material white; white.r = white.g = white.b = 0.75; white.ka = 0.08; white.kd = 0.85;
        Material red = white; red.g = red.b = 0.15;
        Material green = white; green.r = green.b = 0.15;
        Material floor; floor.r = 0.6; floor.g = 0.6; floor.b = 0.65; floor.ka = 0.08; floor.kd = 0.7; floor.ks = 0.2; floor.shininess = 40; floor.kr = 0.2;
        Material mirror; mirror.r = 0.9; mirror.g = 0.9; mirror.b = 0.95; mirror.ka = 0.02; mirror.kd = 0.15; mirror.ks = 0.6; mirror.shininess = 120; mirror.kr = 0.75;
        Material glass; glass.r = 0.95; glass.g = 0.98; glass.b = 1.0; glass.ka = 0; glass.kd = 0.05; glass.ks = 0.5; glass.shininess = 200; glass.kr = 0.1; glass.kt = 0.85; glass.ior = 1.5;
        Material blue; blue.r = 0.2; blue.g = 0.35; blue.b = 0.85; blue.ka = 0.1; blue.kd = 0.7; blue.ks = 0.4; blue.shininess = 30;
        const int mw = s.add_material(white), mr = s.add_material(red), mg = s.add_material(green), mf = s.add_material(floor),
                  mm = s.add_material(mirror), mgl = s.add_material(glass), mb = s.add_material(blue);
        s.add_plane({ 0, 1, 0 }, 0.0, mf); s.add_plane({ 0, -1, 0 }, -10.0, mw);
        s.add_plane({ 1, 0, 0 }, -6.0, mr); s.add_plane({ -1, 0, 0 }, -6.0, mg);
        s.add_plane({ 0, 0, 1 }, -8.0, mw); s.add_plane({ 0, 0, -1 }, -16.0, mw);
        const int mats[8] = { mm, mgl, mb, mm, mgl, mb, mm, mgl };
        for (int k = 0; k < 8; ++k) {
            const double r = 0.95 + 0.05 * k, cx = -4.2 + (k % 4) * 2.8, cz = -3.5 + (k / 4) * 4.5;
            s.add_sphere({ cx, r + (k % 3 == 1 ? 0.5 * k : 0.0), cz }, r, mats[k]);
        }
        s.add_light({ -3.0, 9.2, 4.0 }, { 0.65, 0.62, 0.6 });
        s.add_light({ 3.5, 8.8, -2.0 }, { 0.45, 0.47, 0.5 });
        cam.eye = { 0.0, 5.0, 15.0 }; cam.at = { 0.0, 3.2, 0.0 }; cam.vfov_deg = 42;
    } else {
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
    cam.eye = { 0.3, 2.2, 7.5 }; cam.at = { 0, 0.8, 0 }; cam.vfov_deg = 42;
    }

    try {
        nt_render_stats st{};
        std::vector<uint8_t> img;
        if (n_gpus == 1) {
            Renderer r(s, 0);
            img = r.render(cam, w, h, spp, depth, prec, &st);
        } else {
            int visible = 0;
            check(nt_device_count(&visible));
            if (n_gpus == 0) n_gpus = visible;
            std::vector<int> devices;
            for (int i = 0; i < n_gpus; ++i) devices.push_back(i);
            MultiRenderer r(s, devices);
            img = r.render(cam, w, h, spp, depth, prec, &st); // first frame: includes lazy allocations
            img = r.render(cam, w, h, spp, depth, prec, &st);
            std::cout << "multi-GPU: " << r.devices() << " devices, " << st.total_ms << " ms per frame end to end\n";
        }
        write_ppm(out, img, w, h);
        std::cout << "rendered " << w << "x" << h << " spp " << spp << " depth " << depth << ": "
                  << (st.rays_primary + st.rays_secondary + st.rays_shadow) << " rays, kernel " << st.kernel_ms << " ms -> " << out << "\n";
    } catch (const Error &e) {
        std::cerr << e.what() << "\n";
        return e.code == NT_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
