/* rvpt_host.cpp — see rvpt_host.h. */
#include "rvpt_host.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rvpt_b200
{

Triangle::Triangle(const vec3& v0, const vec3& v1, const vec3& v2, int mat)
    : vertex0(v0.x, v0.y, v0.z, 0), vertex1(v1.x, v1.y, v1.z, 0), vertex2(v2.x, v2.y, v2.z, 0),
      material_id(mat, 0, 0, 0)
{
    /* normal = normalize(cross(v1 - v0, v2 - v0)), geometry.h:87-90 */
    const float ax = v1.x - v0.x, ay = v1.y - v0.y, az = v1.z - v0.z;
    const float bx = v2.x - v0.x, by = v2.y - v0.y, bz = v2.z - v0.z;
    float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const float inv = 1.0f / std::sqrt(nx * nx + ny * ny + nz * nz);
    vertex0.w = nx * inv;
    vertex1.w = ny * inv;
    vertex2.w = nz * inv;
}

Material::Material(const vec4& albedo_, const vec4& emission_, Type type)
    : albedo(albedo_), emission(emission_), data()
{
    data.x = (float)type;
}

void Camera::translate(const vec3& t)
{
    /* translation += vec3(camera_matrix * vec4(in_translation, 0)) */
    const float tr[3] = {translation.x, translation.y, translation.z};
    const float rot[3] = {rotation.x, rotation.y, rotation.z};
    float m[20];
    rvpt_b200_camera_data(tr, rot, aspect, fov, scale, m);
    translation.x += m[0] * t.x + m[4] * t.y + m[8] * t.z;
    translation.y += m[1] * t.x + m[5] * t.y + m[9] * t.z;
    translation.z += m[2] * t.x + m[6] * t.y + m[10] * t.z;
}

void Camera::rotate(const vec3& r)
{
    rotation.x += r.x;
    rotation.y += r.y;
    rotation.z += r.z;
    if (vertical_view_angle_clamp) rotation.y = std::fmin(90.f, std::fmax(-90.f, rotation.y));
}

std::array<float, 20> Camera::get_data() const
{
    std::array<float, 20> out{};
    const float tr[3] = {translation.x, translation.y, translation.z};
    const float rot[3] = {rotation.x, rotation.y, rotation.z};
    rvpt_b200_camera_data(tr, rot, aspect, fov, scale, out.data());
    return out;
}

RVPT::RVPT(uint32_t width, uint32_t height, int device_, uint32_t flags_)
    : scene_camera((float)width / (float)height), w(width), h(height), device(device_), flags(flags_)
{
}

RVPT::~RVPT() { shutdown(); }

bool RVPT::fail()
{
    error = rvpt_b200_last_error(ctx);
    return false;
}

bool RVPT::initialize()
{
    if (triangles.empty() || materials.empty())
    {
        error = "add_triangle/add_material must be called before initialize() (main.cpp:102-109)";
        return false;
    }
    if (rvpt_b200_create(&ctx, device, w, h, flags)) return fail();
    /* top_level_bvh = build_bvh(triangles); sorted = permute_primitives (rvpt.cpp:84-86) */
    bvh_nodes.resize(2 * triangles.size());
    std::vector<uint32_t> perm(triangles.size());
    size_t n_nodes = 0;
    if (rvpt_b200_build_bvh(reinterpret_cast<const rvpt_triangle*>(triangles.data()), triangles.size(),
                            bvh_nodes.data(), &n_nodes, perm.data()))
    {
        error = "BVH build failed";
        return false;
    }
    bvh_nodes.resize(n_nodes);
    sorted_triangles.resize(triangles.size());
    for (size_t i = 0; i < perm.size(); ++i) sorted_triangles[i] = triangles[perm[i]];
    if (rvpt_b200_upload_scene(ctx, bvh_nodes.data(), bvh_nodes.size(),
                               reinterpret_cast<const rvpt_triangle*>(sorted_triangles.data()),
                               sorted_triangles.size(),
                               reinterpret_cast<const rvpt_material*>(materials.data()),
                               materials.size()))
        return fail();
    return true;
}

bool RVPT::update()
{
    const auto camera_data = scene_camera.get_data();
    render_settings.camera_mode = scene_camera.get_camera_mode();
    /* PreviousFrameState::operator== (rvpt.cpp:21-29): split ratio, the four render
     * modes, the camera mode and the camera block — NOT max_bounces / aa */
    const RenderSettings& a = previous_frame_state.settings;
    const RenderSettings& b = render_settings;
    const bool same = previous_frame_state.valid && a.split_ratio[0] == b.split_ratio[0] &&
                      a.split_ratio[1] == b.split_ratio[1] &&
                      a.top_left_render_mode == b.top_left_render_mode &&
                      a.top_right_render_mode == b.top_right_render_mode &&
                      a.bottom_left_render_mode == b.bottom_left_render_mode &&
                      a.bottom_right_render_mode == b.bottom_right_render_mode &&
                      a.camera_mode == b.camera_mode &&
                      previous_frame_state.camera_data == camera_data;
    if (!same)
    {
        render_settings.current_frame = 0;
        previous_frame_state.settings = render_settings;
        previous_frame_state.camera_data = camera_data;
        previous_frame_state.valid = true;
    }
    else
        render_settings.current_frame++;
    return true;
}

bool RVPT::draw()
{
    const auto camera_data = scene_camera.get_data();
    if (rvpt_b200_render_frame(ctx, reinterpret_cast<const rvpt_render_settings*>(&render_settings),
                               camera_data.data()))
        return fail();
    return true;
}

void RVPT::shutdown()
{
    if (ctx) rvpt_b200_destroy(ctx);
    ctx = nullptr;
}

bool RVPT::read_output(std::vector<uint8_t>& rgba8)
{
    rgba8.resize((size_t)w * h * 4);
    if (rvpt_b200_read_output_rgba8(ctx, rgba8.data())) return fail();
    return true;
}

bool RVPT::read_radiance(std::vector<float>& rgba)
{
    rgba.resize((size_t)w * h * 4);
    if (rvpt_b200_read_accum_f32(ctx, rgba.data())) return fail();
    return true;
}

bool load_model(RVPT& rvpt, const std::string& inputfile, int material_id, std::string* err)
{
    std::ifstream in(inputfile);
    if (!in)
    {
        if (err) *err = "cannot open " + inputfile;
        return false;
    }
    std::vector<vec3> verts;
    std::string line;
    while (std::getline(in, line))
    {
        std::istringstream ls(line);
        std::string tag;
        ls >> tag;
        if (tag == "v")
        {
            double x, y, z;
            ls >> x >> y >> z;
            verts.emplace_back((float)x, (float)y, (float)z);
        }
        else if (tag == "f")
        {
            std::vector<long> idx;
            std::string tok;
            while (ls >> tok)
            {
                const long i = std::strtol(tok.c_str(), nullptr, 10); /* v, v/t, v//n, v/t/n */
                idx.push_back(i > 0 ? i - 1 : (long)verts.size() + i);
            }
            for (size_t k = 1; k + 1 < idx.size(); ++k)
            {
                if (idx[0] < 0 || idx[k] < 0 || idx[k + 1] < 0 || (size_t)idx[0] >= verts.size() ||
                    (size_t)idx[k] >= verts.size() || (size_t)idx[k + 1] >= verts.size())
                {
                    if (err) *err = "face index out of range in " + inputfile;
                    return false;
                }
                rvpt.add_triangle(Triangle(verts[idx[0]], verts[idx[k]], verts[idx[k + 1]], material_id));
            }
        }
    }
    return true;
}

} /* namespace rvpt_b200 */
