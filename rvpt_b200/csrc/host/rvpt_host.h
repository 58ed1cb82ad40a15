/*
 * rvpt_host.h — C++ host mirror of the reference's scene / camera / renderer
 * facade for the compute path, glm- and Vulkan-free, on top of the C ABI.
 *
 * A user of RVPT finds the same surface: `Triangle(v0, v1, v2, material_id)`
 * (src/rvpt/geometry.h:76-111), `Material(albedo, emission, type)`
 * (src/rvpt/material.h:9-26), `Camera::translate/rotate/get_data`
 * (src/rvpt/camera.{h,cpp}), and `RVPT::add_triangle/add_material/initialize/
 * update/draw/shutdown` with public `scene_camera` and `render_settings`
 * (src/rvpt/rvpt.h:27-90). `draw()` calls rvpt_b200_render_frame instead of
 * recording and submitting a Vulkan compute dispatch (rvpt.cpp:346-354).
 */
#pragma once

#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/rvpt_abi.h"

namespace rvpt_b200
{

struct vec3
{
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct vec4
{
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(double x_, double y_, double z_, double w_)
        : x((float)x_), y((float)y_), z((float)z_), w((float)w_) {}
};

/* geometry.h:76-111 — 64 bytes, face normal packed into the three .w */
struct Triangle
{
    Triangle() = default;
    Triangle(const vec3& vertex0, const vec3& vertex1, const vec3& vertex2, int material_id);
    vec4 vertex0, vertex1, vertex2, material_id;
};

/* material.h:9-26 — 48 bytes */
struct Material
{
    enum class Type { LAMBERT, MIRROR, DIELECTRIC };
    Material(const vec4& albedo, const vec4& emission, Type type);
    vec4 albedo, emission, data;
};
static_assert(sizeof(Triangle) == sizeof(rvpt_triangle), "Triangle layout");
static_assert(sizeof(Material) == sizeof(rvpt_material), "Material layout");

/* camera.h:13-57 */
class Camera
{
public:
    explicit Camera(float aspect) : aspect(aspect) {}
    void translate(const vec3& in_translation); /* in camera space, camera.cpp:29-33 */
    void rotate(const vec3& in_rotation);       /* degrees, camera.cpp:35-39 */
    void set_fov(float in_fov) { fov = in_fov; }
    void set_scale(float in_scale) { scale = in_scale; }
    void set_camera_mode(int in_mode) { mode = in_mode; }
    void clamp_vertical_view_angle(bool clamp) { vertical_view_angle_clamp = clamp; }
    float get_fov() const noexcept { return fov; }
    float get_scale() const noexcept { return scale; }
    int get_camera_mode() const noexcept { return mode; }
    std::array<float, 20> get_data() const; /* camera.cpp:55-66 */

private:
    int mode = 0;
    float fov = 90.f, scale = 4.f, aspect;
    bool vertical_view_angle_clamp = false;
    vec3 translation{}, rotation{};
};

/* RenderSettings, rvpt.h:77-89 */
struct RenderSettings
{
    int max_bounces = 8;
    int aa = 1;
    uint32_t current_frame = 1;
    int camera_mode = 0;
    int top_left_render_mode = 9;
    int top_right_render_mode = 9;
    int bottom_left_render_mode = 9;
    int bottom_right_render_mode = 9;
    float split_ratio[2] = {0.5f, 0.5f};
};
static_assert(sizeof(RenderSettings) == sizeof(rvpt_render_settings), "RenderSettings layout");

class RVPT
{
public:
    RVPT(uint32_t width, uint32_t height, int device = 0, uint32_t flags = 0);
    ~RVPT();
    RVPT(const RVPT&) = delete;
    RVPT& operator=(const RVPT&) = delete;

    bool initialize(); /* builds the BVH, permutes, uploads (rvpt.cpp:57-94) */
    bool update();     /* frame-counter rule (rvpt.cpp:96-111) */
    bool draw();       /* one frame through the C ABI (rvpt.cpp:346-354) */
    void shutdown();

    void add_material(Material material) { materials.push_back(material); }
    void add_triangle(Triangle triangle) { triangles.push_back(triangle); }
    bool read_output(std::vector<uint8_t>& rgba8);   /* W*H*4 */
    bool read_radiance(std::vector<float>& rgba_f32); /* W*H*4 */
    const std::string& last_error() const { return error; }
    uint32_t width() const { return w; }
    uint32_t height() const { return h; }

    Camera scene_camera;
    RenderSettings render_settings;

private:
    struct PreviousFrameState
    {
        RenderSettings settings;
        std::array<float, 20> camera_data{};
        bool valid = false;
    } previous_frame_state;

    uint32_t w, h;
    int device;
    uint32_t flags;
    rvpt_b200_ctx* ctx = nullptr;
    std::vector<Triangle> triangles, sorted_triangles;
    std::vector<Material> materials;
    std::vector<rvpt_bvh_node> bvh_nodes;
    std::string error;
    bool fail();
};

/* load_model(), main.cpp:12-62: positions + faces of an OBJ file, polygons fan
 * triangulated, every triangle gets `material_id`. Returns false on I/O error. */
bool load_model(RVPT& rvpt, const std::string& inputfile, int material_id, std::string* err);

} /* namespace rvpt_b200 */
