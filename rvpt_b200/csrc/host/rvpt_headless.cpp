/*
 * rvpt_headless — the reference's main() (src/rvpt/main.cpp:88-159) without a
 * window: build the demo scene, run the update()/draw() loop for N frames, dump
 * the result image as a binary PPM.
 *
 *   rvpt_headless <model.obj> [--width W] [--height H] [--frames N] [--bounces B] [--aa A]
 *                 [--translate x y z] [--rotate x y z] [--out image.ppm] [--rgba8-accum]
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rvpt_host.h"

using namespace rvpt_b200;

int main(int argc, char** argv)
{
    if (argc < 2 || !std::strcmp(argv[1], "--help"))
    {
        std::fprintf(stderr,
                     "usage: %s <model.obj> [--width W] [--height H] [--frames N] [--bounces B] [--aa A]\n"
                     "       [--translate x y z] [--rotate x y z] [--out image.ppm] [--rgba8-accum]\n",
                     argv[0]);
        return argc < 2 ? 2 : 0;
    }
    std::string model = argv[1], out = "rvpt.ppm";
    uint32_t width = 1024, height = 512; /* main.cpp:96-97 */
    int frames = 64, bounces = 8, aa = 1;
    vec3 translate, rotate;
    uint32_t flags = 0;
    for (int i = 2; i < argc; ++i)
    {
        auto need = [&](int n) {
            if (i + n >= argc)
            {
                std::fprintf(stderr, "missing value after %s\n", argv[i]);
                std::exit(2);
            }
        };
        if (!std::strcmp(argv[i], "--width")) need(1), width = (uint32_t)std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--height")) need(1), height = (uint32_t)std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--frames")) need(1), frames = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--bounces")) need(1), bounces = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--aa")) need(1), aa = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--out")) need(1), out = argv[++i];
        else if (!std::strcmp(argv[i], "--rgba8-accum")) flags |= RVPT_B200_FLAG_ACCUM_RGBA8;
        else if (!std::strcmp(argv[i], "--translate"))
        {
            need(3);
            translate = vec3((float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3]));
            i += 3;
        }
        else if (!std::strcmp(argv[i], "--rotate"))
        {
            need(3);
            rotate = vec3((float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3]));
            i += 3;
        }
        else
        {
            std::fprintf(stderr, "unknown option %s\n", argv[i]);
            return 2;
        }
    }

    RVPT rvpt(width, height, 0, flags);
    std::string err;
    if (!load_model(rvpt, model, 1, &err)) /* main.cpp:102 */
    {
        std::fprintf(stderr, "[ERROR: MODEL-LOADING] %s\n", err.c_str());
        return 1;
    }
    /* Setup Demo Scene, main.cpp:105-107 */
    rvpt.add_material(Material(vec4(1, 1, 1, 0), vec4(0.1, 0.4, 0.6, 0), Material::Type::LAMBERT));
    rvpt.add_material(Material(vec4(1.0, 1.0, 1.0, 0), vec4(0, 0, 0, 0), Material::Type::LAMBERT));
    if (!rvpt.initialize())
    {
        std::fprintf(stderr, "failed to initialize RVPT: %s\n", rvpt.last_error().c_str());
        return 1;
    }
    rvpt.render_settings.max_bounces = bounces;
    rvpt.render_settings.aa = aa;
    rvpt.scene_camera.translate(translate);
    rvpt.scene_camera.rotate(rotate);

    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < frames; ++f) /* main.cpp:139-155 */
    {
        rvpt.update();
        if (!rvpt.draw())
        {
            std::fprintf(stderr, "draw failed: %s\n", rvpt.last_error().c_str());
            return 1;
        }
    }
    std::vector<uint8_t> rgba;
    if (!rvpt.read_output(rgba))
    {
        std::fprintf(stderr, "read-back failed: %s\n", rvpt.last_error().c_str());
        return 1;
    }
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%d frames of %ux%u (last current_frame = %u) in %.3f ms: %.1f Msamples/s\n", frames, width,
                height, rvpt.render_settings.current_frame, s * 1e3,
                (double)frames * width * height * aa / s / 1e6);

    FILE* fp = std::fopen(out.c_str(), "wb");
    if (!fp)
    {
        std::fprintf(stderr, "cannot write %s\n", out.c_str());
        return 1;
    }
    std::fprintf(fp, "P6\n%u %u\n255\n", width, height);
    for (size_t p = 0; p < (size_t)width * height; ++p) std::fwrite(&rgba[4 * p], 1, 3, fp);
    std::fclose(fp);
    rvpt.shutdown();
    return 0;
}
