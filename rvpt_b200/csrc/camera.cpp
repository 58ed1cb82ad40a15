/*
 * camera.cpp — the 80-byte camera block of Camera::get_data()
 * (src/rvpt/camera.cpp:17-25, 55-66) without glm: mat = I, then translate,
 * rotate(rotation.x about UP), rotate(rotation.y about RIGHT),
 * rotate(rotation.z about FORWARD), each post-multiplied like
 * glm::translate / glm::rotate do; params = aspect, radians(fov), scale, 0.
 *
 * glm is not vendored in the reference tree; its float rounding only shapes
 * the *inputs* of the hot path, so the block produced here is
 * what both the engine and the oracle are fed.
 */
#include <cmath>
#include <cstring>

#include "../../include/rvpt_abi.h"

namespace
{

struct Mat4
{
    float c[4][4]; /* c[column][row] */
};

Mat4 identity()
{
    Mat4 m;
    std::memset(&m, 0, sizeof(m));
    for (int i = 0; i < 4; ++i) m.c[i][i] = 1.0f;
    return m;
}

/* m * T(v): only the last column changes */
Mat4 translate(const Mat4& m, const float v[3])
{
    Mat4 r = m;
    for (int row = 0; row < 4; ++row)
        r.c[3][row] = m.c[0][row] * v[0] + m.c[1][row] * v[1] + m.c[2][row] * v[2] + m.c[3][row];
    return r;
}

/* m * R(angle, axis), axis a unit vector (Rodrigues) */
Mat4 rotate(const Mat4& m, float angle, const float axis[3])
{
    const float c = std::cos(angle), s = std::sin(angle);
    const float t[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
    float rot[3][3];
    rot[0][0] = c + t[0] * axis[0];
    rot[0][1] = t[0] * axis[1] + s * axis[2];
    rot[0][2] = t[0] * axis[2] - s * axis[1];
    rot[1][0] = t[1] * axis[0] - s * axis[2];
    rot[1][1] = c + t[1] * axis[1];
    rot[1][2] = t[1] * axis[2] + s * axis[0];
    rot[2][0] = t[2] * axis[0] + s * axis[1];
    rot[2][1] = t[2] * axis[1] - s * axis[0];
    rot[2][2] = c + t[2] * axis[2];
    Mat4 r = m;
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 4; ++row)
            r.c[col][row] =
                m.c[0][row] * rot[col][0] + m.c[1][row] * rot[col][1] + m.c[2][row] * rot[col][2];
    return r;
}

float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

} /* namespace */

extern "C" void rvpt_b200_camera_data(const float translation[3], const float rotation_deg[3],
                                      float aspect, float fov_deg, float scale, float out[20])
{
    static const float UP[3] = {0, 1, 0}, RIGHT[3] = {1, 0, 0}, FORWARD[3] = {0, 0, 1};
    Mat4 m = identity();
    m = translate(m, translation);
    m = rotate(m, radians(rotation_deg[0]), UP);
    m = rotate(m, radians(rotation_deg[1]), RIGHT);
    m = rotate(m, radians(rotation_deg[2]), FORWARD);
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) out[4 * col + row] = m.c[col][row];
    out[16] = aspect;
    out[17] = radians(fov_deg);
    out[18] = scale;
    out[19] = 0.0f;
}
